"""Drop-in for the helpers of reference eval/util.py that sit on the incremental-session path:
AverageMeter :9-24, accuracy :26-40, freeze_backbone_weights :62-69, get_optim :92-102, get_vocabs :112-129,
drop_a_dim :131-138, log_episode :148-183."""
import numpy as np
import torch

from srb200 import ops


class AverageMeter(object):
    """Computes and stores the average and current value"""

    def __init__(self):
        self.reset()

    def reset(self):
        self.val = 0
        self.avg = 0
        self.sum = 0
        self.count = 0

    def update(self, val, n=1):
        self.val = val
        self.sum += val * n
        self.count += n
        self.avg = self.sum / self.count


def percent(hits, batch_size):
    """hits * (100 / batch) with the reference's fp32 rounding (correct_k.mul_(100.0 / batch_size))."""
    return torch.tensor([float(hits)], dtype=torch.float32).mul_(100.0 / batch_size)


def accuracy(output, target, topk=(1,)):
    """Top-k accuracy (percent) of CUDA logits: rank of the true class from the scoring kernel."""
    for k in topk:
        if k not in (1, 5):
            raise NotImplementedError("srb200 accuracy supports k in {1, 5} (the reference only uses these)")
    r = ops.score_logits(output.detach().contiguous(), target)
    c = r["counts"].cpu().tolist()
    return [percent(c[0] if k == 1 else c[1], target.size(0)) for k in topk]


def freeze_backbone_weights(backbone, opt, epoch, exclude=['classifier.transform']):
    if opt.freeze_backbone_at == epoch:
        print("Freezing the backbone.")
        for name, param in backbone.named_parameters():
            param.requires_grad = False
            if any(map(lambda s: name.startswith(s), exclude)):
                print("Not frozen: ", name)
                param.requires_grad = True


def get_optim(net, opt):
    """Same optimiser objects as the reference (used only when a caller drives the epoch loop itself; the fused
    driver in eval/language_eval.py keeps optimiser state inside sr_head_run)."""
    if opt.adam:
        return torch.optim.Adam(net.parameters(), lr=opt.learning_rate, weight_decay=0.0005)
    return torch.optim.SGD(net.parameters(), lr=opt.learning_rate, momentum=opt.momentum, weight_decay=opt.weight_decay)


def get_vocabs(base_loader=None, novel_loader=None, query_ys=None):
    vocab_all = []
    vocab_base = None
    if base_loader is not None:
        vocab_base = [name for name in base_loader.dataset.label2human if name != '']
        vocab_all += vocab_base
    vocab_novel, orig2id = None, None
    if novel_loader is not None:
        novel_ids = np.sort(np.unique(query_ys))
        label2human_novel = novel_loader.dataset.label2human
        vocab_novel = [label2human_novel[i] for i in novel_ids]
        orig2id = dict(zip(novel_ids, len(vocab_base) + np.arange(len(novel_ids))))
        vocab_all += vocab_novel
    return vocab_base, vocab_all, vocab_novel, orig2id


def drop_a_dim(data):
    support_xs, support_ys, query_xs, query_ys = data
    batch_size, _, height, width, channel = support_xs.size()
    support_xs = support_xs.view(-1, height, width, channel)
    query_xs = query_xs.view(-1, height, width, channel)
    support_ys = support_ys.view(-1).detach().numpy()
    query_ys = query_ys.view(-1).detach().numpy()
    return (support_xs, support_ys, query_xs, query_ys)


def log_episode(novel_labels, vocab_novel, epoch, novel_acc, base_acc, running_base, running_novel):
    avg_score = (novel_acc + base_acc) / 2
    running_avg = (running_base + running_novel) / 2
    print('\n{:25} {:}\n{:25} {:}\n{:25} {:}\n{:25} {:.4f}\n{:25} {:.4f}\n{:25} {:.4f}\n{:25} {:.4f}\n{:25} {:.4f}\n'
          '{:25} {:.4f}\n'.format("Classes:", novel_labels, "Labels:", vocab_novel, "Fine-tuning epochs:", epoch - 1,
                                  "Novel acc:", novel_acc, "Base acc:", base_acc, "Average:", avg_score,
                                  "Runnning Base Avg:", running_base, "Running Novel Avg:", running_novel,
                                  "Running Average:", running_avg), flush=True)
