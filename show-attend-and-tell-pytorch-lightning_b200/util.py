"""Data formats on the caller side of the hot path: the preprocessed caption JSON, the dataset
that yields the `(img, encoded_captions, lengths)` batches SAT.train_batch / val_batch consume, and the length-bucketed
sampler.  Counterpart of the reference's util.py:16-87 (same class names, constructor arguments and sampling behaviour) --
host-side Python, nothing here touches the GPU.

JSON layout (preprocess.ipynb cell 17, read at util.py:21-29 / train.py:238-242):
  vocab_stoi {word: id}, vocab_size, embed_dim, pretrained_embedding, min_count, max_cap_length,
  train / val / test: {samples, img_paths [n], encoded_captions [n][ncap][max_cap_length+2], lengths [n][ncap]}
with captions encoded as <START> w1 .. wk <END> <PAD>... and lengths = number of targets (k + 1).
"""
import json
from collections import defaultdict

import numpy as np
import torch
from torch.utils.data import Dataset
from torch.utils.data.sampler import Sampler


def json_loader(path):
    with open(path) as f:
        return json.load(f)


def pil_loader(path):
    from PIL import Image
    with open(path, "rb") as f:
        return Image.open(f).convert("RGB")


class CocoCaptionDataset(Dataset):
    """One item = (image tensor, LongTensor [ncap, max_cap_length+2] of word ids, LongTensor [ncap] of target counts)."""

    def __init__(self, jsonpath, split="train", transforms=None):
        data = jsonpath if isinstance(jsonpath, dict) else json_loader(jsonpath)
        self.json = data
        self.split = split
        self.transforms = transforms
        self.vocab_stoi = data["vocab_stoi"]
        self.vocab_itos = {i: w for w, i in self.vocab_stoi.items()}
        part = data[split]
        self.img_paths, self.encoded_captions, self.lengths = part["img_paths"], part["encoded_captions"], part["lengths"]
        if not (len(self.img_paths) == len(self.encoded_captions) == len(self.lengths)):
            raise AssertionError("img_paths / encoded_captions / lengths of split %r differ in length" % split)

    def stoi(self, s):
        return int(self.vocab_stoi.get(s, self.vocab_stoi["<UNK>"]))

    def itos(self, i):
        return str(self.vocab_itos.get(int(i), "<UNK>"))

    def __len__(self):
        return len(self.img_paths)

    def _to_tensor(self, img):
        if self.transforms is not None:
            return self.transforms(img)
        arr = np.asarray(img, dtype=np.uint8)                      # default transform: ToTensor (HWC uint8 -> CHW float in [0,1])
        return torch.from_numpy(arr.copy()).permute(2, 0, 1).float().div_(255.0)

    def __getitem__(self, idx):
        img = self._to_tensor(pil_loader(self.img_paths[idx]))
        return img, torch.as_tensor(self.encoded_captions[idx], dtype=torch.long), torch.as_tensor(self.lengths[idx], dtype=torch.long)


class BucketSampler(Sampler):
    """Samples grouped by their number of targets (sum of the caption lengths of a sample), longest group first, shuffled
    inside each group with numpy's global RNG on every pass -- batches then hold captions of similar length, and the largest
    batch comes first (util.py:48-87)."""

    def __init__(self, lengths, batch_size, indices=None):
        self.lengths = lengths
        self.batch_size = batch_size
        self.indices = list(indices) if indices else list(range(len(lengths)))
        groups = defaultdict(list)
        for i, per_caption in zip(self.indices, self.lengths):
            groups[int(sum(per_caption))].append(i)
        self.grouped_indices = [groups[k] for k in sorted(groups, reverse=True)]

    def __iter__(self):
        order = []
        for g in self.grouped_indices:
            np.random.shuffle(g)
            order.extend(g)
        return iter(order)

    def __len__(self):
        return len(self.lengths)
