"""Oracle (test infrastructure): restatement of the weight regularisers of the incremental-session path.

Follows reference models/resnet_language.py:
  LangPuller.__init__/update_novel_embeds/create_pulling_mapping/forward  :21-87   (semantic / mapping pullers)
  LangPuller.loss1                                                       :89-90
  LangPuller.get_projected_weight                                        :92-97   (distance2subspace)
  ResNet.regloss / reglossnovel                                          :229-240
and models/util.py:50-67 (get_embeds).
"""
import os
import pickle

import numpy as np
import torch
import torch.nn.functional as F


def get_embeds(embed_pth, vocab, dim=500):
    """models/util.py:50-67: per-label mean of word vectors; an out-of-vocabulary word RESETS the running sum to a
    float64 zero vector (so 'komondor' becomes all-zero and the stack is promoted to float64)."""
    with open(embed_pth, "rb") as f:
        table = pickle.load(f)
    rows = []
    for token in vocab:
        words = token.split(' ')
        acc = 0
        for w in words:
            if w in table:
                acc = acc + table[w]
            else:
                acc = np.zeros(dim)
        acc = acc / len(words)
        rows.append(torch.from_numpy(np.asarray(acc)))
    dt = torch.float64 if any(r.dtype == torch.float64 for r in rows) else rows[0].dtype
    return torch.stack([r.to(dt) for r in rows], 0)


class Puller(object):
    """State of LangPuller: label embeddings of the base vocabulary and of the current session's novel vocabulary."""

    def __init__(self, opt, vocab_base, vocab_novel):
        self.opt = opt
        self.path = os.path.join(opt.word_embed_path, "{0}_dim{1}.pickle".format(opt.dataset, opt.word_embed_size))
        self.base_embeds = self._load(vocab_base)
        self.novel_embeds = self._load(vocab_novel)
        self.mapping = None

    def _load(self, vocab):
        e = get_embeds(self.path, vocab).float()
        return e[:, :300] if self.opt.glove else e          # first 300 dims are GloVe (:51-54)

    def update_novel(self, vocab_novel):                     # :56-65
        self.novel_embeds = self._load(vocab_novel)

    def set_mapping(self, state_dict, base_weight_size=640):  # create_pulling_mapping :67-72
        # LinearMap(indim, outdim) is CONSTRUCTED (default nn.Linear init, which draws from the CPU generator) before
        # load_state_dict overwrites it - every session; the draws matter for the dropout masks that follow.
        torch.nn.Linear(self.novel_embeds.size(1), base_weight_size)
        self.mapping = (state_dict['map.weight'].float(), state_dict['map.bias'].float())

    def pullers(self, base_weight):                          # forward :74-87
        if self.mapping is None:
            scores = self.novel_embeds @ self.base_embeds.t()
            scores = torch.softmax(scores / self.opt.temperature, dim=1)
            return scores @ base_weight
        return F.linear(self.novel_embeds, self.mapping[0], self.mapping[1]).detach()


def projected_weight(base_weight, weights):
    """get_projected_weight :92-97 (differentiable w.r.t. `weights`)."""
    Q, _ = torch.linalg.qr(base_weight.t(), mode='reduced')   # == torch.qr(., some=True): Q is [640, 60]
    mut = weights @ Q
    mutnorm = mut / torch.norm(Q.t(), dim=1).unsqueeze(0)
    return mutnorm @ Q.t()


def pull_loss(pull, inspired, weights):                      # loss1 :89-90
    return pull * torch.norm(inspired - weights) ** 2


def drift_loss(lmbd, current_rows, anchor):                  # regloss / reglossnovel :229-240 (un-squared norm)
    return lmbd * torch.norm(current_rows - anchor)
