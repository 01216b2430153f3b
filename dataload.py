"""Dataset side of the drop-in (`data.target: dataload.CATER` / `dataload.MovingMnistLMDB` in config/*.yaml).

The reference's readers (/root/reference/dataload.py:75-490) need lmdb / decord / nltk and the CATER-GEN /
Moving-MNIST files; they are data preparation, outside the sampling hot path ("next" row N4).
What the hot path needs from this module is the batch-dict contract (dataload.py:260,370):

    'images' f32 [L, C, H, W] per item ([-1,1] CATER, [-0.5,0.5] MNIST), 'text' i64 [T] = [CLS]=1 .. [SEP]=2,
    'speed' f32 scalar in [0,1), 'video_id' str (CATER only; the entry deletes it before the model)

`SyntheticCaptionVideos` produces exactly that from a seed, so `main_mage.py --split test --synthetic N` runs
the whole entry without the datasets.  The real-dataset class names resolve (so a saved config.yaml still
instantiates) and fail loudly when asked for data.
"""
from __future__ import annotations

import torch
from torch.utils.data import Dataset

from mage_b200 import synthetic as syn


class SyntheticCaptionVideos(Dataset):
    def __init__(self, model_params: dict, n_items: int, seed: int = 1234, text_len: int = 20, with_video_id: bool = True):
        self.n = int(n_items)
        self.batch = syn.make_batch(model_params, self.n, seed=seed, text_len=text_len) if self.n else {}
        self.with_video_id = with_video_id

    def __len__(self):
        return self.n

    def __getitem__(self, i):
        item = {k: v[i] for k, v in self.batch.items()}
        if self.with_video_id:
            item["video_id"] = f"synthetic_{i:06d}"
        return item


def collate_fn(items):
    """Pads 'text' to the longest caption with 0 like the reference's collate (dataload.py:262-271, 372-380)."""
    out = {}
    T = max(int(it["text"].shape[0]) for it in items)
    text = torch.zeros(len(items), T, dtype=torch.long)
    for i, it in enumerate(items):
        text[i, : it["text"].shape[0]] = it["text"]
    out["text"] = text
    out["images"] = torch.stack([it["images"] for it in items])
    if "speed" in items[0]:
        out["speed"] = torch.stack([torch.as_tensor(it["speed"], dtype=torch.float32) for it in items])
    if "video_id" in items[0]:
        out["video_id"] = [it["video_id"] for it in items]
    return out


class _RealDatasetOutOfScope(Dataset):
    def __init__(self, *args, **kwargs):
        raise NotImplementedError(
            f"{type(self).__name__}: the LMDB/decord readers of the reference (dataload.py:75-490) are data preparation, "
            "outside the sampling hot path (SURVEY.md §8f N4); run main_mage.py with --synthetic N, or pass batches "
            "with the documented dict contract to MAGE.autoregressive_generate")


class CATER(_RealDatasetOutOfScope):
    pass


class MovingMnistLMDB(_RealDatasetOutOfScope):
    pass
