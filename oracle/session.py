"""Oracle (test infrastructure): restatement of the incremental-session loop on a plain state dict.

Follows reference eval/language_eval.py:
  validate                              :18-43
  eval_base                             :46-69
  few_shot_finetune_incremental_test    :71-454
and eval/util.py: accuracy :26-40, get_optim :92-102, get_vocabs :112-129, drop_a_dim :131-138.

Two schedules produce the same numbers:
  'literal'  every epoch re-runs the backbone on support, memory and all query sets, exactly like the reference
             (this is the schedule timed as the CPU baseline);
  'cached'   the eval-mode features are computed once per session after the train-mode epoch 1 and re-used; the
             per-block forward counters are advanced as if the literal schedule had run.  Used to make oracle runs
             that finish in seconds-to-minutes for the GPU parity tests.
"""
import time

import numpy as np
import torch
import torch.nn.functional as F

from . import backbone as bb
from . import regularizer as rg


def accuracy(output, target, topk=(1,)):
    """eval/util.py:26-40 -> list of 1-element tensors (percent)."""
    maxk = max(topk)
    _, pred = output.topk(maxk, 1, True, True)
    correct = pred.t().eq(target.view(1, -1).expand_as(pred.t()))
    return [correct[:k].reshape(-1).float().sum(0, keepdim=True).mul_(100.0 / target.size(0)) for k in topk]


class Timers(object):
    def __init__(self):
        self.train_s = 0.0       # epoch-loop body up to and including the convergence bookkeeping
        self.score_s = 0.0       # validate + eval_base
        self.loop_s = 0.0        # the whole `while stop_condition` body (language_eval.py:242-350): train step + validate
        self.steps = 0
        self.images_scored = 0
        self.images_backbone = 0


def run_sessions(sd, world, n_sessions=None, schedule='literal', ckpt=None, model='resnet18', keep_w_trajectory=False,
                 verbose=False, probe_rows=0):
    """-> dict record.  `sd` (state dict of fp32 tensors) is modified in place, like the reference's `net`."""
    opt = world.opt
    assert not opt.linear_bias, "backbones are trained with --no_linear_bias; the bias branches are dead (resnet_language.py:239)"
    plan = bb.block_plan(model, opt.no_dropblock)
    counters = bb.new_counters(plan)
    T = Timers()
    rec = dict(sessions=[], weighted=[], novel=[], base=[], timers=T)

    torch.manual_seed(opt.set_seed)                                    # :101-102
    np.random.seed(opt.set_seed)
    base_weight = sd['classifier.weight'].detach().clone()             # :106-107
    n_base_cls = base_weight.shape[0]

    def fwd(x, train):
        T.images_backbone += x.shape[0]
        return bb.features(sd, plan, x, train, counters)

    def logits_of(feat):
        return F.linear(feat, sd['classifier.weight'])

    bs = world.base_support_loader.batches[0] if world.base_support_loader is not None else None
    if bs is not None:                                                 # :112-116 (drop_a_dim)
        base_support_xs = bs[0].view(-1, *bs[0].shape[2:])
        base_support_ys = bs[1].view(-1).numpy()
    base_x, base_y = world.base_val_loader.batches[0][:2]              # :121
    base_x, base_y = base_x.squeeze(0), base_y.squeeze(0)
    mem_x, mem_y = None, None                                          # dataset/memory.py
    mem_f = None

    def eval_base_imgs():                                              # :46-69
        t0 = time.perf_counter()
        with torch.no_grad():
            out = logits_of(fwd(base_x, False))
            acc1 = accuracy(out, base_y, (1, 5))[0]
        T.score_s += time.perf_counter() - t0
        T.images_scored += base_x.shape[0]
        return float(np.mean([acc1[0].item()])), torch.argmax(out, 1), out

    acc_base0, _, _ = eval_base_imgs()                                    # :128
    rec['weighted'].append(acc_base0)
    rec['base0'] = acc_base0
    iter_num = n_sessions if n_sessions is not None else (8 if opt.continual else opt.neval_episodes)   # :132-136

    vocab_base_names = [n for n in world.base_val_loader.dataset.label2human if n != '']
    l2h_novel = world.meta_valloader.dataset.label2human
    query_x_list, query_y_list = [], []
    reserve = None
    puller = None
    acc_novel_sum, acc_base_sum = 0.0, 0.0

    for idx in range(iter_num):
        ep = world.meta_valloader.batches[idx % len(world.meta_valloader.batches)]
        support_xs = ep[0].view(-1, *ep[0].shape[2:])                 # drop_a_dim :131-138
        support_ys = ep[1].view(-1).numpy()
        query_xs = ep[2].view(-1, *ep[2].shape[2:])
        query_ys = ep[3].view(-1).numpy()
        if bs is not None:
            support_xs = torch.cat([support_xs, base_support_xs], 0)  # :149-150
        if idx > 0:
            prev_vocab_base, prev_vocab_novel = vocab_base, vocab_novel
        novel_ids = np.sort(np.unique(query_ys))                      # get_vocabs :112-129
        vocab_novel = [l2h_novel[i] for i in novel_ids]
        vocab_base = list(vocab_base_names)
        orig2id = dict(zip(novel_ids, len(vocab_base) + np.arange(len(novel_ids))))
        if idx == 0:
            orig_base_num = len(vocab_base)
        else:
            vocab_base = prev_vocab_base + prev_vocab_novel            # :166-167
        W = sd['classifier.weight']
        if idx == 1:                                                   # :172-185
            reserve = W.detach().clone()[-opt.n_ways:, :]
        elif idx > 1:
            reserve = torch.cat((reserve, W.detach().clone()[-opt.n_ways:, :]), 0)
        for k in list(orig2id.keys()):                                 # :193-196
            orig2id[k] = orig2id[k] + idx * opt.n_ways
        query_ys_id = torch.LongTensor([orig2id[y] for y in query_ys])
        support_ys_id = torch.LongTensor([orig2id[y] for y in support_ys])
        query_x_list.append(query_xs)                                  # :199-204
        query_y_list.append(query_ys_id)
        if bs is not None:
            support_ys_id = torch.cat([support_ys_id, torch.from_numpy(base_support_ys)])   # :207-209

        # augment_base_classifier_ (resnet_language.py:202-226): default nn.Linear init on the CPU generator
        novel_init = torch.nn.Linear(W.shape[1], len(novel_ids), bias=False).weight.detach()
        W = torch.cat([W.detach(), novel_init], 0).requires_grad_(True)
        sd['classifier.weight'] = W

        use_pull = opt.label_pull is not None and getattr(opt, 'pulling', None) == "regularize"
        pullers = None
        if use_pull:                                                   # :218-228
            if idx == 0:
                puller = rg.Puller(opt, vocab_base, vocab_novel)
            else:
                puller.update_novel(vocab_novel)
            if opt.attraction_override == "mapping_linear_label2image":
                puller.set_mapping(ckpt[opt.attraction_override])
            pullers = puller.pullers(base_weight[:orig_base_num, :])

        if opt.adam:                                                   # get_optim, eval/util.py:92-102
            optimizer = torch.optim.Adam([W], lr=opt.learning_rate, weight_decay=0.0005)
        else:
            optimizer = torch.optim.SGD([W], lr=opt.learning_rate, momentum=opt.momentum, weight_decay=opt.weight_decay)

        train_loss, epoch, stable_epochs = 15, 1, 0                    # :234-239
        stable = opt.target_train_loss == 0
        n_vb = len(vocab_base)
        terms, w_traj = [], []
        sup_f = None
        go = True
        while go:
            t0 = t_loop0 = time.perf_counter()
            train_mode = epoch == 1        # net.train() at :211, validate() leaves the net in eval mode from epoch 2 on
            if schedule == 'literal' or train_mode:
                f_s = fwd(support_xs, train_mode)
                f_m = fwd(mem_x, train_mode) if (opt.memory_replay and mem_y is not None) else None
            else:
                if sup_f is None:          # first eval-mode epoch: build the session's feature cache
                    with torch.no_grad():
                        sup_f = fwd(support_xs, False)
                        mem_f = fwd(mem_x, False) if (opt.memory_replay and mem_y is not None) else None
                else:                      # keep BasicBlock.num_batches_tracked in step with the literal schedule
                    for k in counters:
                        counters[k] += 1 + (1 if mem_f is not None else 0)
                f_s, f_m = sup_f, mem_f
            output = F.linear(f_s, W)
            ce_s = F.cross_entropy(output, support_ys_id)              # :252-253
            loss = ce_s
            ce_m = None
            if f_m is not None:                                        # :256-258
                ce_m = F.cross_entropy(F.linear(f_m, W), mem_y)
                loss = loss + ce_m
            reg_b = reg_n = pull = None
            if opt.lmbd_reg_transform_w is not None:                   # :261-265
                reg_b = rg.drift_loss(opt.lmbd_reg_transform_w, W[:base_weight.size(0), :], base_weight)
                loss = loss + reg_b
            if opt.lmbd_reg_novel is not None and idx > 0:             # :268-274
                reg_n = rg.drift_loss(opt.lmbd_reg_novel, W[n_base_cls:n_base_cls + reserve.size(0), :], reserve)
                loss = loss + reg_n
            if use_pull:                                               # :277-290
                if opt.attraction_override == "distance2subspace":
                    pullers = rg.projected_weight(base_weight, W[n_vb:, :])
                pull = rg.pull_loss(opt.label_pull, pullers, W[n_vb:, :])
                loss = loss + pull
            optimizer.zero_grad()
            loss.backward()
            optimizer.step()
            with torch.no_grad():                                      # :298-318
                lv = loss.item()
                if stable:
                    if abs(lv - train_loss) < opt.convergence_epsilon:
                        stable_epochs += 1
                    else:
                        stable_epochs = 0
                    if stable_epochs == opt.stable_epochs:
                        go = False
                acc1, acc5 = accuracy(output, support_ys_id, (1, 5))
                train_loss = lv
                if (epoch >= opt.max_novel_epochs) or (train_loss <= opt.target_train_loss and
                                                       epoch >= opt.min_novel_epochs + 1):
                    go = False
            terms.append([lv, ce_s.item(), 0.0 if ce_m is None else ce_m.item(), 0.0 if reg_b is None else reg_b.item(),
                          0.0 if reg_n is None else reg_n.item(), 0.0 if pull is None else pull.item(),
                          acc1[0].item(), acc5[0].item()])
            if keep_w_trajectory:
                w_traj.append(W.detach().clone())
            T.train_s += time.perf_counter() - t0
            T.steps += 1
            # validate (:321-326, :18-43): every epoch in the reference; only the last epoch's result is consumed
            t0 = time.perf_counter()
            if schedule == 'literal' or not go:
                with torch.no_grad():
                    q_feats = [fwd(q, False) for q in query_x_list]
                    test_acc, test_preds, q_logits = [], [], []
                    for qf, qy in zip(q_feats, query_y_list):
                        out = F.linear(qf, W)
                        a1 = accuracy(out, qy, (1, 5))[0]
                        test_acc.append(a1[0])
                        test_preds.append(torch.argmax(out, 1))
                        q_logits.append(out)
                T.images_scored += sum(q.shape[0] for q in query_x_list)
            else:
                for k in counters:
                    counters[k] += len(query_x_list)
            T.score_s += time.perf_counter() - t0
            T.loop_s += time.perf_counter() - t_loop0
            epoch += 1

        if opt.memory_replay:                                          # :353-359
            inds = np.random.choice(opt.n_shots, opt.memory_replay)
            inds = np.tile(5 * np.arange(5) + inds, (5, 1)) + (np.tile(np.arange(0, 125, 25), (5, 1))).T
            inds = inds.flatten()
            if mem_y is None:
                mem_x, mem_y = support_xs[inds, :], support_ys_id[inds]
            else:
                mem_x = torch.cat((mem_x, support_xs[inds, :]), 0)
                mem_y = torch.cat((mem_y, support_ys_id[inds]), 0)
        sd['classifier.weight'] = W.detach()
        acc_base_, base_pred, base_logits = eval_base_imgs()                        # :362-367
        test_acc_r = [round(i.item(), 2) for i in test_acc]            # :370-376
        test_acc_m = float(np.array(test_acc_r).mean())
        acc_base_sum += acc_base_
        acc_novel_sum += test_acc_m
        w1 = 60 if opt.dataset == "miniImageNet" else 200              # :383-393
        w2 = len(vocab_base) + len(vocab_novel) - 60
        weighted = (w1 * acc_base_ + w2 * test_acc_m) / (w1 + w2)
        rec['weighted'].append(round(weighted, 2))
        rec['novel'].append(round(test_acc_m, 2))
        rec['base'].append(round(acc_base_, 2))
        srec = dict(epochs=epoch - 1, terms=np.asarray(terms, dtype=np.float64), W=W.detach().clone(),
                    novel_session_acc=test_acc_r, query_pred=[p.clone() for p in test_preds],
                    query_logits=[q.clone() for q in q_logits], base_pred=base_pred.clone(), base_logits=base_logits.clone(),
                    acc_base=acc_base_,
                    memory_inds=inds.copy() if opt.memory_replay else None, vocab_novel=list(vocab_novel))
        if keep_w_trajectory:
            srec['W_traj'] = torch.stack(w_traj)
        if probe_rows:
            with torch.no_grad():
                saved = dict(counters)
                srec['probe_feat'] = bb.features(sd, plan, support_xs[:probe_rows], False, counters)
                counters.update(saved)
        srec['bn'] = {k: v.clone() for k, v in sd.items() if 'running_' in k or 'num_batches_tracked' in k}
        rec['sessions'].append(srec)
        if verbose:
            print("oracle session %d: epochs %d loss %.6f novel %s base %.2f" % (idx + 1, epoch - 1, terms[-1][0],
                                                                               test_acc_r, acc_base_), flush=True)
    rec['acc_novel_avg'] = acc_novel_sum / iter_num                    # AverageMeter.avg, :454
    rec['acc_base_avg'] = acc_base_sum / iter_num
    rec['counters'] = dict(counters)
    return rec
