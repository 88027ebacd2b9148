"""Dataset shims.  The reference's TrainDataset / TestDataset (core/dataset.py:8-52) read COCO artefacts
(hickle feature arrays + pickled caption vectors) that cannot exist offline; `SyntheticCaptionDataset`
yields tensors of the same shapes / dtypes / conventions (features.py:101-118, preprocess.py:121-134,303-345)."""
import torch
from torch.utils.data import Dataset


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

    def __len__(self):
        return len(self.image_idx)

    def __getitem__(self, i):
        j = int(self.image_idx[i])
        if self.with_captions:
            return self.features[j], self.positions[j], self.captions[i], j
        return self.features[j], self.positions[j], j
