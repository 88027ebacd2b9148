"""Dataset shims.  The reference's TrainDataset / TestDataset (core/dataset.py:8-52) read COCO artefacts
(hickle feature arrays + pickled caption vectors) that cannot exist offline; `SyntheticCaptionDataset`
yields tensors of the same shapes / dtypes / conventions (features.py:101-118, preprocess.py:121-134,303-345)."""
import numpy as np
import torch
from torch.utils.data import Dataset

from core.utils import load_coco_data


class TrainDataset(Dataset):
    """core/dataset.py:8-30: one item per CAPTION -> (features[image], positions[image], caption, image_idx)."""

    def __init__(self, data_path, split):
        self.data = load_coco_data(data_path=data_path, split=split)

    def __getitem__(self, index):
        image_idx = self.data['image_idxs'][index]
        return np.asarray(self.data['features'][image_idx]), np.asarray(self.data['positions'][image_idx]), \
            self.data['captions'][index], image_idx

    def __len__(self):
        return len(self.data['captions'])

    @property
    def len_image(self):
        return len(self.data['positions'])

    @property
    def data_dict(self):
        return self.data


class TestDataset(TrainDataset):
    """core/dataset.py:33-52: (features[image], positions[image], image_idx)."""

    def __init__(self, data_path, split='test'):
        super().__init__(data_path, split)

    def __getitem__(self, index):
        image_idx = self.data['image_idxs'][index]
        return np.asarray(self.data['features'][image_idx]), np.asarray(self.data['positions'][image_idx]), image_idx


class IndexedCaptions(Dataset):
    """The same items WITHOUT the region arrays: (image_idx, caption) / (image_idx,).  Used with a device-resident
    `RegionCache` of `dataset.data['features'|'positions']`: the host never gathers or ships the
    300 KB of fp32 features per caption that dataset.py:12-18 returns."""

    def __init__(self, dataset, with_captions=True, unique_images=False):
        d = dataset.data
        self.image_idxs = np.asarray(d['image_idxs'], dtype=np.int64)
        self.captions = np.asarray(d['captions'], dtype=np.int32) if with_captions else None
        if unique_images:                      # evaluation: decode every image once, not once per caption
            self.image_idxs = np.unique(self.image_idxs)
            assert not with_captions
        self.len_image = len(d['positions'])

    def __len__(self):
        return len(self.image_idxs)

    def __getitem__(self, i):
        if self.captions is None:
            return (int(self.image_idxs[i]),)
        return int(self.image_idxs[i]), self.captions[i]


class SyntheticCaptionDataset(Dataset):
    def __init__(self, num_images, num_objects, dim_features, dim_positions, caption_len, num_vocab,
                 captions_per_image=5, seed=1234, with_captions=True):
        g = torch.Generator().manual_seed(seed)
        R = num_objects + 1                                        # region 0 = whole image (features.py:101)
        n_valid = torch.randint((R + 1) // 2, R + 1, (num_images,), generator=g)
        valid = (torch.arange(R)[None, :] < n_valid[:, None]).unsqueeze(-1)
        self.features = torch.randn(num_images, R, dim_features, generator=g).abs_() * valid
        pos = torch.zeros(num_images, R, dim_positions)
        xy = torch.rand(num_images, R, 4, generator=g)
        pos[:, :, 0] = torch.minimum(xy[:, :, 0], xy[:, :, 1])
        pos[:, :, 2] = torch.maximum(xy[:, :, 0], xy[:, :, 1])
        pos[:, :, 1] = torch.minimum(xy[:, :, 2], xy[:, :, 3])
        pos[:, :, 3] = torch.maximum(xy[:, :, 2], xy[:, :, 3])
        cls = torch.randint(4, dim_positions, (num_images, R), generator=g)
        pos.scatter_(2, cls.unsqueeze(-1), (torch.rand(num_images, R, generator=g) * 0.99 + 0.01).unsqueeze(-1))
        pos[:, 0] = 0
        pos[:, 0, 2:4] = 1
        self.positions = pos * valid
        self.with_captions = with_captions
        self.len_image = num_images
        n = num_images * captions_per_image if with_captions else num_images
        self.image_idx = torch.arange(n) // (captions_per_image if with_captions else 1)
        if with_captions:
            cap = torch.zeros(n, caption_len, dtype=torch.int32)
            cap[:, 0] = 1
            lens = torch.randint(min(5, caption_len - 2), caption_len - 1, (n,), generator=g)
            words = torch.randint(4, num_vocab, (n, caption_len), generator=g, dtype=torch.int32)
            for i in range(n):
                cap[i, 1:1 + int(lens[i])] = words[i, :int(lens[i])]
                cap[i, 1 + int(lens[i])] = 2
            self.captions = cap

    @property
    def data(self):
        """The reference's data dict (utils.py:32-64) over the synthetic arrays."""
        d = {'features': self.features, 'positions': self.positions, 'image_idxs': self.image_idx.numpy()}
        if self.with_captions:
            d['captions'] = self.captions.numpy()
        return d

    def __len__(self):
        return len(self.image_idx)

    def __getitem__(self, i):
        j = int(self.image_idx[i])
        if self.with_captions:
            return self.features[j], self.positions[j], self.captions[i], j
        return self.features[j], self.positions[j], j
