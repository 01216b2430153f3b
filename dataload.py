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


# Word-level vocabularies of the reference's datasets (token ids are part of the checkpoint contract: they index
# text_encoder.token_embedding).  Moving MNIST: dataload.py:199-203; CATER-GEN-v1 / -v2: dataload.py:299-312.
_SPECIAL = ["[PAD]", "[CLS]", "[SEP]"]
VOCABS = {
    "mnist": _SPECIAL + list("0123456789") + ["the", "digit", "and", "is", "are", "bouncing", "moving", "here", "there", "around",
                                              "jumping", "up", "down", "left", "right", "then", "."],
    "caterv1": _SPECIAL + ["the", "cone", "snitch", "is", "sliding", "picked", "placed", "containing", "rotating", "and", "to", "up", "(",
                           ")", "1", "2", "3", "-1", "-2", "-3", ",", ".", "first", "second", "third", "fourth", "quadrant"],
    "caterv2": _SPECIAL + ["the", "cone", "snitch", "is", "sliding", "picked", "placed", "containing", "and", "to", "up", "sphere",
                           "cylinder", "cube", "small", "medium", "large", "metal", "rubber", "gold", "gray", "red", "blue", "green",
                           "brown", "purple", "cyan", "yellow", "(", ")", "1", "2", "3", "-1", "-2", "-3", ",", ".", "rotating", "while",
                           "contained", "still", "first", "second", "third", "fourth", "quadrant"],
}
VOCABS = {k: {w: i for i, w in enumerate(v)} for k, v in VOCABS.items()}


def encode_caption(text: str, dataset: str) -> torch.Tensor:
    """caption -> int64 [T] = [CLS] words [SEP] (dataload.py:215-224 / :325-334).  Moving MNIST captions are split on white
    space; CATER captions go through nltk.word_tokenize in the reference, which for this grammar is: words, signed integers and
    the punctuation ( ) , . as separate tokens.  Unknown words raise KeyError like the reference's dict lookup."""
    import re
    vocab = VOCABS[dataset]
    words = text.split() if dataset == "mnist" else re.findall(r"-?\d+|[A-Za-z]+|[(),.]", text)
    return torch.tensor([vocab["[CLS]"]] + [vocab[w] for w in words] + [vocab["[SEP]"]], dtype=torch.long)


def decode_caption(tokens, dataset: str) -> str:
    """inverse table lookup (dataload.py:226-237)."""
    rev = {i: w for w, i in VOCABS[dataset].items()}
    return " ".join(rev[int(t)] for t in tokens)


def load_first_frame(path: str, channels: int, size: int) -> torch.Tensor:
    """image file -> [1, C, size, size] float in the dataset's range: CATER Resize(128) + Normalize(0.5, 0.5) -> [-1, 1]
    (dataload.py:282-286), Moving MNIST x/255 - 0.5 (dataload.py:254)."""
    import numpy as np
    from PIL import Image
    im = Image.open(path).convert("L" if channels == 1 else "RGB").resize((size, size), Image.BILINEAR)
    x = torch.from_numpy(np.asarray(im, dtype=np.float32) / 255.0)
    x = x.unsqueeze(0) if channels == 1 else x.permute(2, 0, 1)
    return (x - 0.5 if channels == 1 else (x - 0.5) / 0.5).unsqueeze(0).contiguous()


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
