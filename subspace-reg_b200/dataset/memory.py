"""Drop-in for reference dataset/memory.py:4-28: replay memory as two growing tensors (data, labels)."""
import torch
from torch.utils.data import Dataset


class Memory(Dataset):
    def __init__(self):
        self.data = None
        self.labels = None

    def additems(self, data, label):
        if self.data is None:
            self.data, self.labels = data, label
        else:
            self.data = torch.cat((self.data, data), dim=0)
            self.labels = torch.cat((self.labels, label), dim=0)

    def __getitem__(self, item):
        return self.data[item], self.labels[item]

    def __len__(self):
        return 0 if self.labels is None else len(self.labels)
