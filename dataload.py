"""Dataset side of the drop-in (`data.target: dataload.CATER` / `dataload.MovingMnistLMDB` in config/*.yaml).

The reference's readers (/root/reference/dataload.py:75-490) need lmdb / decord / nltk and the CATER-GEN /
Moving-MNIST files (SURVEY.md §8f "next" row N4).  What the hot path needs from this module is the batch-dict contract
(dataload.py:260,370):

    'images' f32 [L, C, H, W] per item ([-1,1] CATER, [-0.5,0.5] MNIST), 'text' i64 [T] = [CLS]=1 .. [SEP]=2,
    'speed' f32 scalar in [0,1), 'video_id' str (CATER only; the entry deletes it before the model)

`SyntheticCaptionVideos` produces exactly that from a seed, so `main_mage.py --split test --synthetic N` runs
the whole entry without the datasets.  `CATER` / `MovingMnistLMDB` are the real readers (same kwargs, items, collate and
speed-based frame sub-sampling as the reference); `decord` / `lmdb` are used when installed, dependency-free on-disk forms
of the same content (`.npy` clips or frame directories; a pickled list) otherwise.
"""
from __future__ import annotations

import os

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
    def __init__(self, model_params: dict, n_items: int, seed: int = 1234, text_len: int = 20, with_video_id: bool = True,
                 frames: int = 1):
        self.n = int(n_items)
        self.batch = syn.make_batch(model_params, self.n, seed=seed, text_len=text_len, frames=frames) if self.n else {}
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


# ----------------------------------------------------------------------------------------------------------------------
# Real-dataset readers: same constructor kwargs, item dict and collate as the reference's classes
# (/root/reference/dataload.py:183-381), including the speed-based temporal sub-sampling.  The reference decodes videos with
# `decord` and reads Moving MNIST from an `lmdb` file; both are used here when importable, and each reader also accepts a
# dependency-free on-disk form of the same content so the front end runs (and is tested) without them.
# ----------------------------------------------------------------------------------------------------------------------
def speed_subsample_indices(frame_num: int, speed: float, sample_speed, min_interval: float):
    """dataload.py:244-248 / :346-349: a clip played at `speed` in [0,1) keeps every `interval`-th frame, interval =
    max(min_interval, speed * (hi - lo) + lo) with [lo, hi] = sample_speed; indices = floor(linspace(0, n-1, round(n / interval)))."""
    import numpy as np
    interval = max(min_interval, speed * (sample_speed[-1] - sample_speed[0]) + sample_speed[0])
    return np.floor(np.linspace(0, frame_num - 1, round(frame_num / interval), endpoint=True)).astype(np.int32)


def _pad_to_length(images: torch.Tensor, frames_length: int) -> torch.Tensor:
    """dataload.py:256-257 / :353-354: a clip shorter than frames_length repeats its last frame."""
    if images.shape[0] < frames_length:
        images = torch.cat([images, images[-1].unsqueeze(0).repeat(frames_length - images.shape[0], 1, 1, 1)], dim=0)
    return images


def _no_bert(path):
    if path is not None:
        raise NotImplementedError("the BERT tokenizer path (dataload.py:15-73, pytorch_transformers) is not part of any shipped config; "
                                  "the word-level vocabularies are")


class _VideoFrames:
    """Random access to the RGB frames [T, H, W, 3] uint8 of one video: `decord.VideoReader` for real video files (what the
    reference uses, dataload.py:343), or -- dependency-free -- a `.npy` / `.npz` array or a directory of frame images."""

    def __init__(self, path: str):
        import numpy as np
        self.np = np
        self.arr = self.vr = self.files = None
        if os.path.isdir(path):
            self.files = sorted(os.path.join(path, f) for f in os.listdir(path) if f.lower().endswith((".png", ".jpg", ".jpeg", ".bmp")))
        elif path.endswith(".npy"):
            self.arr = np.load(path, mmap_mode="r")
        elif path.endswith(".npz"):
            z = np.load(path)
            self.arr = z[z.files[0]]
        else:
            try:
                from decord import VideoReader
            except ImportError as e:
                raise ImportError(f"decoding {path!r} needs `decord` (the reference's reader); without it store the clip as a .npy "
                                  "[T,H,W,3] uint8 array or a directory of frame images") from e
            self.vr = VideoReader(path)

    def __len__(self):
        return len(self.files) if self.files is not None else len(self.arr) if self.arr is not None else len(self.vr)

    def get_batch(self, idx):
        np = self.np
        if self.vr is not None:
            return self.vr.get_batch(list(idx)).asnumpy()
        if self.arr is not None:
            return np.stack([np.asarray(self.arr[int(i)]) for i in idx])
        from PIL import Image
        return np.stack([np.asarray(Image.open(self.files[int(i)]).convert("RGB")) for i in idx])


def _cater_transform(frames, size: int = 128) -> torch.Tensor:
    """The reference's default CATER transform (dataload.py:282-286): Resize(128) (utils/videotransforms.py:270-287: shorter side
    to 128, aspect kept, NEAREST on PIL images), ClipToTensor (uint8 -> [0,1]), Normalize(0.5, 0.5) -> [-1, 1].  Returns [T,C,H,W]."""
    import numpy as np
    from PIL import Image
    out = []
    for f in frames:
        im = Image.fromarray(f)
        w, h = im.size
        if not ((w <= h and w == size) or (h <= w and h == size)):
            if w < h:
                nw, nh = size, int(size * h / w)
            else:
                nw, nh = int(size * w / h), size
            im = im.resize((nw, nh), Image.NEAREST)
        out.append(np.asarray(im, dtype=np.float32) / 255.0)
    x = torch.from_numpy(np.stack(out)).permute(0, 3, 1, 2)
    return ((x - 0.5) / 0.5).contiguous()


class CATER(Dataset):
    """dataload.py:273-381.  `<data_root>/<split>_{explicit|ambiguous}.json` = {"0": {"video": rel_path, "caption": str}, ...}."""

    def __init__(self, dataset: str, data_root: str, split: str, frames_length: int, sample_speed: list, image_transform=None,
                 tokenizer_path=None, randomness=False):
        import json
        _no_bert(tokenizer_path)
        mode = "ambiguous" if randomness else "explicit"
        with open(os.path.join(data_root, f"{split}_{mode}.json"), "r") as fp:
            self.anno = json.load(fp)
        self.dataset, self.data_root, self.transform = dataset, data_root, image_transform or _cater_transform
        self.frames_length, self.sample_speed, self.randomness = frames_length, sample_speed, randomness
        self.vocab = VOCABS[dataset]
        self.tokenizer, self.padding_idx = None, self.vocab["[PAD]"]

    def __len__(self):
        return len(self.anno)

    def encode(self, x):
        return encode_caption(x, self.dataset).numpy()

    def decode(self, tokens):
        return " " + decode_caption(tokens, self.dataset)

    def __getitem__(self, idx):
        import random
        entry = self.anno[str(idx)]
        video_path = os.path.join(self.data_root, entry["video"])
        vid = _VideoFrames(video_path)
        speed = random.random()
        choice_idx = speed_subsample_indices(len(vid), speed, self.sample_speed, 3.0)
        images = self.transform(vid.get_batch(choice_idx)[: self.frames_length])
        return {"video_id": os.path.basename(video_path), "images": _pad_to_length(images, self.frames_length),
                "text": torch.tensor(self.encode(entry["caption"]), dtype=torch.long), "speed": torch.tensor(speed, dtype=torch.float)}

    def collate_fn(self, data):
        return collate_fn(data)


class MovingMnistLMDB(Dataset):
    """dataload.py:183-271.  `<data_root><split>.lmdb` holds pickled (video uint8 [T,1,H,W], caption) tuples under the keys
    b"0", b"1", ... (LmdbReader, :75-181); without `lmdb`, `<data_root><split>.pkl` -- a pickled list of the same tuples -- is read."""

    def __init__(self, data_root: str, split: str, frames_length: int, sample_speed: list, image_transform=None, bert_path=None,
                 eos_token=0):
        import pickle
        _no_bert(bert_path)
        self.txn = self.items = None
        lmdb_path, pkl_path = data_root + split + ".lmdb", data_root + split + ".pkl"
        if os.path.exists(lmdb_path):
            try:
                import lmdb
            except ImportError as e:
                raise ImportError(f"{lmdb_path} needs the `lmdb` package (or convert it to {pkl_path}: a pickled list of "
                                  "(video uint8 [T,1,H,W], caption) tuples)") from e
            env = lmdb.open(lmdb_path, subdir=False, readonly=True, lock=False, readahead=False)
            self.txn = env.begin()
            self.n = env.stat()["entries"]
        else:
            with open(pkl_path, "rb") as fp:
                self.items = pickle.load(fp)
            self.n = len(self.items)
        self.transform, self.frames_length, self.sample_speed = image_transform, frames_length, sample_speed
        self.vocab = VOCABS["mnist"]
        self.tokenizer, self.padding_idx = None, self.vocab["[PAD]"]

    def __len__(self):
        return self.n

    def encode(self, x):
        return encode_caption(x, "mnist").numpy()

    def decode(self, tokens):
        return " " + decode_caption(tokens, "mnist")

    def __getitem__(self, idx):
        import pickle
        import random
        images_raw, caption = pickle.loads(self.txn.get(f"{idx}".encode("ascii"))) if self.txn is not None else self.items[idx]
        speed = random.random()
        choice_idx = speed_subsample_indices(images_raw.shape[0], speed, self.sample_speed, 1.0)
        images_raw = images_raw[choice_idx][: self.frames_length]
        if self.transform is not None:
            image = self.transform(images_raw.transpose(0, 2, 3, 1)).permute(1, 0, 2, 3)
        else:
            image = torch.tensor(images_raw / 255. - 0.5, dtype=torch.float)
        return {"images": _pad_to_length(image, self.frames_length), "text": torch.tensor(self.encode(caption), dtype=torch.long),
                "speed": torch.tensor(speed, dtype=torch.float)}

    def collate_fn(self, data):
        return collate_fn(data)
