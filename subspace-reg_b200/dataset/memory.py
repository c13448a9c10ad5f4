"""Replay memory of the incremental sessions (API of reference dataset/memory.py: `additems`, `data`, `labels`, len()).

Exemplars are appended once per session and read as two dense tensors every session, so the chunks are kept as
they arrive and concatenated lazily (one torch.cat per session instead of one per `additems`)."""
import torch
from torch.utils.data import Dataset


class Memory(Dataset):
    def __init__(self):
        self._chunks = []          # [(images, labels)] in arrival order
        self._dense = None         # cached (data, labels) of everything appended so far

    def additems(self, data, label):
        if data.shape[0] != label.shape[0]:
            raise ValueError("Memory.additems: %d images but %d labels" % (data.shape[0], label.shape[0]))
        self._chunks.append((data, label))
        self._dense = None

    def _materialise(self):
        if self._dense is None and self._chunks:
            if len(self._chunks) > 1:
                merged = (torch.cat([c[0] for c in self._chunks], dim=0), torch.cat([c[1] for c in self._chunks], dim=0))
                self._chunks = [merged]
            self._dense = self._chunks[0]
        return self._dense

    @property
    def data(self):
        dense = self._materialise()
        return None if dense is None else dense[0]

    @property
    def labels(self):
        dense = self._materialise()
        return None if dense is None else dense[1]

    def __getitem__(self, item):
        dense = self._materialise()
        return dense[0][item], dense[1][item]

    def __len__(self):
        return sum(int(c[1].shape[0]) for c in self._chunks)
