"""Drop-in for reference models/util.py: create_model (:6-35) and get_embeds (:50-67)."""
import pickle

import numpy as np
import torch


def create_model(name, n_cls, opt, vocab=None, dataset='miniImageNet'):
    """Same factory contract as the reference: resnet12/resnet18 with avg_pool=True, drop_rate=0.1, dropblock_size=5."""
    from . import model_dict
    if dataset not in ('miniImageNet', 'tieredImageNet'):
        raise NotImplementedError('dataset not supported: {}'.format(dataset))
    if name not in model_dict:
        raise NotImplementedError('model {} not supported in dataset {}:'.format(name, dataset))
    return model_dict[name](avg_pool=True, drop_rate=0.1, dropblock_size=5, num_classes=n_cls, vocab=vocab, opt=opt)


def get_embeds(embed_pth, vocab, dim=500):
    """Label embeddings: mean of the per-word vectors of each label; an out-of-vocabulary word resets the running sum
    to a float64 zero vector, exactly like the reference (the stacked result is then float64 until ``.float()``)."""
    with open(embed_pth, "rb") as f:
        table = pickle.load(f)
    rows = []
    for token in vocab:
        words = token.split(' ')
        acc = 0
        for w in words:
            acc = acc + table[w] if w in table else np.zeros(dim)
        rows.append(torch.from_numpy(np.asarray(acc / len(words))))
    dtype = torch.float64 if any(r.dtype == torch.float64 for r in rows) else rows[0].dtype
    return torch.stack([r.to(dtype) for r in rows], 0)
