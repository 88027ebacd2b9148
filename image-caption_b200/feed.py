"""Data feed for the caption path.

The reference's `TrainDataset.__getitem__` (core/dataset.py:12-18) returns `features[image_idx]` -- a [R, 2048] fp32
array, 300 KB -- for every CAPTION (five per image), the DataLoader collates them on the host and the wrapper ships
[B, R, 2048] fp32 over PCIe each step (core/models.py:115-120): 75 MB / step at batch 256.  At the rates of the CUDA
path (~55 k samples/s, 16 GB/s of features) that host gather + copy is the bottleneck, not the model.

* `RegionCache`   -- the whole split's region features live in HBM once, already packed the way the encoder's
                     embedding GEMM reads them ([features | positions | 0-pad], compute dtype; COCO train2014 with
                     37 regions is 18 GB in bf16).  A step sends B image numbers + the captions (22 KB).
* `PrefetchLoader`-- for data that does not fit: pinned double-buffered host->device copies on a copy stream, one
                     batch ahead of the compute stream.

Both feed the SAME kernels: a cached batch is bit-identical to the same images passed as fp32 tensors."""
from typing import Iterable, Iterator, Optional, Sequence

import numpy as np
import torch

from ._native import call
from .engine import BF16, F32, RegionBatch


def _cpu_slice(x, lo: int, hi: int) -> torch.Tensor:
    if torch.is_tensor(x):
        return x[lo:hi]
    return torch.from_numpy(np.ascontiguousarray(x[lo:hi], dtype=np.float32))      # ndarray / memmap / h5 dataset


class RegionCache:
    """features [n, R, Df] / positions [n, R, Dp] (numpy arrays, memmaps or CPU tensors, fp32) -> device-resident packed
    rows [n, R, Kc] in the model's compute dtype + validity bytes [n, R] (a region is padding iff its position row is
    all zero, model.py:202-209).  Built chunk by chunk through a pinned staging buffer with the library's own packing
    kernels, so the cached rows are exactly what `encode()` would have packed from the fp32 inputs."""

    def __init__(self, model, features, positions, chunk_images: int = 1024):
        eng = model._engine()
        n, R, Df = features.shape
        Dp = positions.shape[2]
        cfg = eng.cfg
        assert positions.shape[:2] == (n, R) and Df == cfg.encode_dim_features and Dp == cfg.encode_dim_positions
        Kc = eng._cat_width()
        self.num_images, self.regions, self.dim_features, self.dim_positions = n, R, Df, Dp
        self.engine_key = (Kc, eng.act, Df, Dp)
        dev = eng.dev
        self.xcat = torch.zeros(n, R, Kc, dtype=eng.tdt, device=dev)      # pad columns stay zero
        self.valid = torch.empty(n, R, dtype=torch.uint8, device=dev)
        self.err = torch.zeros(1, dtype=torch.int32, device=dev)
        chunk = max(1, min(chunk_images, n))
        hf = torch.empty(chunk, R, Df, dtype=torch.float32).pin_memory()
        hp = torch.empty(chunk, R, Dp, dtype=torch.float32).pin_memory()
        df, dp = torch.empty_like(hf, device=dev), torch.empty_like(hp, device=dev)
        st = torch.cuda.current_stream(dev)
        esz = self.xcat.element_size()
        for lo in range(0, n, chunk):
            m = min(chunk, n - lo)
            st.synchronize()                                              # staging buffers are free again
            hf[:m].copy_(_cpu_slice(features, lo, lo + m))
            hp[:m].copy_(_cpu_slice(positions, lo, lo + m))
            df[:m].copy_(hf[:m], non_blocking=True)
            dp[:m].copy_(hp[:m], non_blocking=True)
            rows = m * R
            dst = self.xcat.data_ptr() + lo * R * Kc * esz
            call("icap_copy2d", df.data_ptr(), F32, Df, dst, eng.act, Kc, rows, Df, 0, st.cuda_stream)
            call("icap_copy2d", dp.data_ptr(), F32, Dp, dst + Df * esz, eng.act, Kc, rows, Dp, 0, st.cuda_stream)
            call("icap_region_valid", dp.data_ptr(), rows, Dp, self.valid.data_ptr() + lo * R, None, st.cuda_stream)
        st.synchronize()

    @property
    def nbytes(self) -> int:
        return self.xcat.numel() * self.xcat.element_size() + self.valid.numel()

    def batch(self, image_idx) -> RegionBatch:
        """image numbers (CPU or device tensor / sequence, int32 or int64) -> the handle the model API takes as
        `object_features` (with `position_features=None`)."""
        idx = torch.as_tensor(image_idx)
        if idx.dtype not in (torch.int32, torch.int64):
            idx = idx.long()
        if not idx.is_cuda:
            lo, hi = int(idx.min()), int(idx.max())
            if lo < 0 or hi >= self.num_images:
                raise IndexError(f"image index {lo if lo < 0 else hi} outside the cache of {self.num_images} images")
        return RegionBatch(self, idx.contiguous())

    def check(self) -> None:
        """Raises if any gather since the last check() saw an out-of-range device-side index (one host sync)."""
        if int(self.err.item()):
            self.err.zero_()
            raise IndexError("a RegionBatch carried an image index outside the cache")


class PrefetchLoader:
    """Wraps an iterable of tuples of CPU tensors (a DataLoader): each batch is staged into one of `depth` pinned
    buffer sets and copied to the device on a copy stream while the previous batch computes.  Yields tuples of device
    tensors that stay valid until `depth - 1` more batches have been drawn."""

    def __init__(self, loader: Iterable, device, depth: int = 2):
        assert depth >= 2
        self.loader, self.dev, self.depth = loader, torch.device(device), depth
        self.copy_stream = torch.cuda.Stream(device=self.dev)
        self._host: list = [None] * depth
        self._devb: list = [None] * depth
        self._free: list = [None] * depth          # event: the compute stream is done with slot i

    def __len__(self):
        return len(self.loader)

    def _stage(self, slot: int, batch: Sequence[torch.Tensor]):
        batch = [torch.as_tensor(t) for t in batch]
        shapes = [(tuple(t.shape), t.dtype) for t in batch]
        if self._devb[slot] is None or [(tuple(t.shape), t.dtype) for t in self._devb[slot]] != shapes:
            self._devb[slot] = [torch.empty(s, dtype=d, device=self.dev) for s, d in shapes]
            self._host[slot] = None
        if self._free[slot] is not None:
            self._free[slot].synchronize()         # the previous user of this slot has consumed it
        if all(t.is_pinned() for t in batch):      # already page-locked (a pin_memory=True DataLoader): no staging copy
            self._host[slot] = batch               # keeps the source alive until the copy below has run
        else:
            if self._host[slot] is None or any(not h.is_pinned() or h.shape != t.shape or h.dtype != t.dtype
                                               for h, t in zip(self._host[slot], batch)):
                self._host[slot] = [torch.empty(sh, dtype=d).pin_memory() for sh, d in shapes]
            for h, t in zip(self._host[slot], batch):
                h.copy_(t)
        ev = torch.cuda.Event()
        with torch.cuda.stream(self.copy_stream):
            for d, h in zip(self._devb[slot], self._host[slot]):
                d.copy_(h, non_blocking=True)
            ev.record(self.copy_stream)
        return ev

    def __iter__(self) -> Iterator[tuple]:
        it = iter(self.loader)
        pending: list = []                          # (slot, ready event)
        slot = 0
        for _ in range(self.depth - 1):
            b = next(it, None)
            if b is None:
                break
            pending.append((slot, self._stage(slot, b)))
            slot = (slot + 1) % self.depth
        while pending:
            cur, ready = pending.pop(0)
            b = next(it, None)
            if b is not None:
                pending.append((slot, self._stage(slot, b)))
                slot = (slot + 1) % self.depth
            main = torch.cuda.current_stream(self.dev)
            main.wait_event(ready)
            yield tuple(self._devb[cur])
            done = torch.cuda.Event()
            done.record(torch.cuda.current_stream(self.dev))
            self._free[cur] = done
