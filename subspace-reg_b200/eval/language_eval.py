"""Drop-in for reference eval/language_eval.py: validate (:18-43), eval_base (:46-69) and
few_shot_finetune_incremental_test (:71-454) with the reference's signatures, prints and return values.

B200-first restructuring of the session loop (same numbers, different schedule):
  * epoch 1 of a session is the only train-mode backbone pass (batch-stat BN, running-stat update, dropout /
    DropBlock with the CPU generator's masks) -> its features feed ONE head step;
  * from epoch 2 on the backbone is in eval mode with fixed weights and statistics, so its outputs never change:
    support, memory, every past session's queries and the base batch go through the tcgen05 backbone ONCE per
    session into a feature cache;
  * all remaining epochs run inside one persistent device kernel (sr_head_run) that applies the reference's
    stopping rule on the device; the host reads back one status word and the loss trace;
  * validate / eval_base score the cached features (sr_eval_logits) after the last epoch - the reference computes
    them every epoch but only consumes the last (:370-376);
  * BasicBlock.num_batches_tracked is advanced as if every skipped forward had happened (DropBlock's gamma).
"""
from __future__ import print_function

import itertools
import sys
import time

import numpy as np
import torch

from dataset.memory import Memory
from models.resnet_language import LangPuller
from srb200 import _lib as L
from srb200 import ops
from srb200 import rng
from .util import AverageMeter, drop_a_dim, freeze_backbone_weights, get_vocabs, log_episode, percent


def _score(net, feat, labels, confusion=None):
    r = ops.eval_logits(feat, net.classifier.weight.detach().contiguous(), labels, confusion)
    counts = r["counts"].cpu().tolist()
    n = feat.shape[0]
    return percent(counts[0], n), percent(counts[1], n), r["loss_sum"].item() / n, r["pred"].cpu().numpy().astype(np.int64), r


def validate(query_xs, query_ys_id, net, criterion, opt, epoch):
    """Per past session: top-1 / top-5 (percent, 1-element tensors), mean CE and argmax predictions.
    Side effect kept from the reference: the network is left in eval mode."""
    net.eval()
    with torch.no_grad():
        if isinstance(query_xs, list):
            acc1, acc5, losses, preds = [], [], [], []
            for i, item in enumerate(query_xs):
                r = validate(item, query_ys_id[i], net, criterion, opt, epoch)
                acc1.append(r[0])
                acc5.append(r[1])
                losses.append(r[2])
                preds.append(r[3])
            return acc1, acc5, losses, preds
        feat = net.features(query_xs.cuda())
        a1, a5, loss, pred, _ = _score(net, feat, query_ys_id.cuda())
        return a1[0], a5[0], loss, pred


def eval_base(net, base_batch, criterion, vocab_all=None, df=None, return_preds=False):
    if df is not None:
        raise NotImplementedError("the visualisation dataframe (vis=True) relies on DataFrame.append, removed in pandas 2")
    net.eval()
    with torch.no_grad():
        input, target, *_ = base_batch
        feat = net.features(input.squeeze(0).cuda())
        a1, _, _, pred, _ = _score(net, feat, target.squeeze(0).cuda())
    acc = np.mean([a1[0].item()])
    if return_preds:
        return acc, pred
    return acc


def few_shot_finetune_incremental_test(net, ckpt, criterion, meta_valloader, base_val_loader, opt, vis=False,
                                       base_support_loader=None):
    if vis or opt.track_weights or opt.track_label_inspired_weights or opt.save_preds_0:
        raise NotImplementedError("vis / track_weights / track_label_inspired_weights / save_preds_0 are the reference's "
                                  "pandas dumps (broken under pandas >= 2); they are outside the B200 hot path")
    if criterion is not None and not isinstance(criterion, torch.nn.CrossEntropyLoss):
        raise NotImplementedError("the fused head implements nn.CrossEntropyLoss (mean reduction) only")
    if net.classifier.bias is not None:
        raise NotImplementedError("classifier bias: the reference's bias branches are dead code (--no_linear_bias)")
    if opt.freeze_backbone_at != 1:
        raise NotImplementedError("freeze_backbone_at != 1 trains the backbone, which is outside the incremental-session path")
    record = dict(sessions=[], timers=dict(train_s=0.0, score_s=0.0, backbone_imgs=0, steps=0, images_scored=0),
                  phases=dict(setup=0.0, train_pass=0.0, head1=0.0, cache=0.0, head=0.0, score=0.0))
    ph = record['phases']
    # Phase times are DEVICE times between CUDA events on the current stream, resolved once at the end of the run: a
    # session is queued without any host synchronisation (the host only waits once per session, for its results).
    phase_events = []

    def _mark():
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        return ev
    few_shot_finetune_incremental_test.last_record = record
    net._last_record = record            # (the function attribute is shared between threads; this one is not)
    tm = record['timers']

    acc_novel, acc_base = [AverageMeter() for _ in range(2)]
    weighted_avg_l, acc_novel_list, acc_base_list = [[] for _ in range(3)]

    if getattr(opt, 'conv_precision', None):     # 'bf16' | 'bf16x3' (not a reference flag: the B200 precision tier)
        net.set_conv_precision(opt.conv_precision)

    rng.manual_seed(opt.set_seed)          # torch.manual_seed + np.random.seed (of this thread's run, srb200/rng.py)

    # (the reference deep-copies the whole network - language_eval.py:106-107 - only to read this clone off the copy)
    base_weight, base_bias = net._get_base_weights()
    base_weight = base_weight.cuda().contiguous()
    dev = base_weight.device
    n_base_cls = net.num_classes
    W_cols = base_weight.shape[1]

    base_valloader_it = itertools.cycle(iter(base_val_loader))
    meta_valloader_it = itertools.cycle(iter(meta_valloader))
    if base_support_loader is not None:
        base_support_it = itertools.cycle(iter(base_support_loader))
        base_support_xs, base_support_ys, *_ = drop_a_dim(next(base_support_it))
        base_support_xs = base_support_xs.cuda(non_blocking=True)   # joined to every session's support set ON the device

    novel_query_collection = None
    novel_query_collection_id = None
    base_batch = next(base_valloader_it)
    base_x = base_batch[0].squeeze(0).cuda(non_blocking=True)
    base_y = base_batch[1].squeeze(0).cuda(non_blocking=True)

    if opt.memory_replay:
        memory = Memory()

    # Initial validation on base samples.
    ev0 = _mark()
    acc_base_ = eval_base(net, (base_x, base_y), criterion)
    phase_events.append(('score', ev0, _mark()))
    tm['images_scored'] += base_x.shape[0]
    base_y_host = base_batch[1].squeeze(0).cpu().numpy()
    confusion_run = torch.zeros((100, 100), dtype=torch.int64, device=dev)   # (gold, predicted) over every scoring of the run
    weighted_avg_l.append(acc_base_)
    record['base0'] = acc_base_

    iter_num = opt.neval_episodes
    if opt.continual:
        iter_num = getattr(opt, 'n_sessions_override', 8)   # 8 sessions for miniImageNet (reference literal)
        basec_map = ckpt['training_classes']

    # A session's accuracy bookkeeping is finished one session late (finish_session below), while the device is busy with
    # the next session; whatever the next session prints before that point is held back so that the log keeps the
    # reference's order.
    pending_finish = None
    held_prints = []

    def say(*args):
        line = " ".join(str(a) for a in args) + "\n"
        if pending_finish is not None:
            held_prints.append(line)
        else:
            sys.stdout.write(line)

    for idx in range(iter_num):
        say("\n**** Iteration {}/{} ****\n".format(idx + 1, opt.neval_episodes))
        support_xs, support_ys, query_xs, query_ys = drop_a_dim(next(meta_valloader_it))
        if base_support_loader is not None:
            # (concatenating on the host would turn pinned loader tensors into a pageable one: a host memcpy plus a
            # synchronous staged H2D copy per session, 14 % of an end-to-end sweep)
            support_xs = torch.cat([support_xs.cuda(non_blocking=True), base_support_xs], 0)

        if idx > 0:
            prev_vocab_base = vocab_base
            prev_vocab_novel = vocab_novel
        vocab_base, vocab_all, vocab_novel, orig2id = get_vocabs(base_val_loader, meta_valloader, query_ys)
        say("Vocab base: ", vocab_base)
        say("Vocab novel: ", vocab_novel)
        if idx == 0:
            orig_base_num = len(vocab_base)
        if idx > 0:
            vocab_base = prev_vocab_base + prev_vocab_novel

        # Previous sessions' novel weights as they were when their session ended (reglossnovel anchors).
        if idx == 1:
            novel_weight_to_reserve = net.classifier.weight.clone().detach()[-opt.n_ways:, :].requires_grad_(False)
            say(f"Novel weight to reserve is of shape {novel_weight_to_reserve.shape} at session {idx+1}.")
        if idx > 1:
            new_novel_set = net.classifier.weight.clone().detach()[-opt.n_ways:, :].requires_grad_(False)
            novel_weight_to_reserve = torch.cat((novel_weight_to_reserve, new_novel_set), 0)
            say(f"Novel weight to reserve is of shape {novel_weight_to_reserve.shape} at session {idx+1}.")

        novel_labels = np.sort(np.unique(query_ys))
        say("Novel labels: ", novel_labels)
        for k, v in orig2id.items():
            orig2id[k] = v + idx * opt.n_ways
        query_ys_id = torch.LongTensor([orig2id[y] for y in query_ys])
        support_ys_id = torch.LongTensor([orig2id[y] for y in support_ys])

        query_xs_d = query_xs.cuda(non_blocking=True)
        if novel_query_collection_id is None:
            novel_query_collection = [query_xs_d]
            novel_query_collection_id = [query_ys_id.cuda(non_blocking=True)]
            query_ids_host = [query_ys_id.numpy()]
        else:
            novel_query_collection.append(query_xs_d)
            novel_query_collection_id.append(query_ys_id.cuda(non_blocking=True))
            query_ids_host.append(query_ys_id.numpy())

        if base_support_loader is not None:
            support_ys_id = torch.cat([support_ys_id, torch.from_numpy(base_support_ys)])

        if (idx == 0 and not net.engine().mask_prefetch_alive()
                and getattr(opt, 'attraction_override', None) != "mapping_linear_label2image"):
            # Start drawing the dropout masks of the WHOLE run on the host thread right away: from here on the CPU
            # generator is only consumed by each session's nn.Linear(640, n_ways) init and by the masks themselves (a
            # wrong guess is detected when a mask is taken and costs nothing but the prefetch).  Session 1's own masks
            # then overlap its set-up and the first convolutions instead of being drawn in line.
            skip = opt.n_ways * W_cols
            fwds = [(skip, support_xs.shape[0])]
            for j in range(1, iter_num):
                fwds.append((skip, support_xs.shape[0]))
                if opt.memory_replay == 1:
                    fwds.append((0, 25 * j))
            net.engine().start_mask_prefetch(fwds)

        # DropBlock masks of this session's train-mode forward(s): their gamma is known now (it depends on the epochs the
        # previous sessions ran), so a second host thread draws them while this thread sets the session up and launches
        # the first blocks (each mask is checked against the live generator when it is consumed).
        nbt0 = next(iter(net.block_counters().values()))
        ahead = [(support_xs.shape[0], nbt0 + 1)]
        if opt.memory_replay and len(memory) > 0:
            ahead.append((len(memory), nbt0 + 2))
        net.engine().start_dropblock_ahead(ahead)

        net.train()
        net.augment_base_classifier_(len(novel_labels))

        use_pull = opt.label_pull is not None and getattr(opt, 'pulling', None) == "regularize"
        pull_mode, pull_t, q_rows = L.SR_PULL_NONE, None, 0
        if use_pull:
            if idx == 0:
                lang_puller = LangPuller(opt, vocab_base, vocab_novel)
            else:
                lang_puller.update_novel_embeds(vocab_novel)
            if opt.attraction_override == "mapping_linear_label2image":
                lang_puller.create_pulling_mapping(ckpt[opt.attraction_override])
            pullers = lang_puller(base_weight[:orig_base_num, :])
            if opt.attraction_override == "distance2subspace":
                qt, q_rows, _ = lang_puller.factor(base_weight)   # constant for the whole run (torch.qr every epoch in the reference)
                pull_mode, pull_t = L.SR_PULL_PROJECT, qt
            else:
                pull_mode, pull_t = L.SR_PULL_FIXED, pullers.detach().contiguous()

        opt.stable = True if opt.target_train_loss == 0 else False
        freeze_backbone_weights(net, opt, 1, exclude=["classifier"])
        t_train0 = time.perf_counter()
        support_xs_d = support_xs if support_xs.is_cuda else support_xs.cuda(non_blocking=True)
        support_ys_d = support_ys_id.cuda(non_blocking=True)
        n_sup = support_xs_d.shape[0]
        has_mem = bool(opt.memory_replay and len(memory) > 0)
        n_mem = len(memory) if has_mem else 0
        counters0 = next(iter(net.block_counters().values()))

        # ---- epoch 1: the session's only train-mode forward(s); BN running statistics move here ----
        tp0 = _mark()
        ph['setup'] += time.perf_counter() - t_train0
        f_train = net.features(support_xs_d)
        if has_mem:
            f_train = torch.cat([f_train, net.features(memory.data)], 0)
        tm['backbone_imgs'] += n_sup + n_mem
        if idx + 1 < iter_num and not net.engine().mask_prefetch_alive():
            # the dropout masks of EVERY remaining session are drawn on a host thread, working ahead while the GPU runs
            # the cache builds and head loops (each mask is verified against the live generator when it is consumed)
            skip = opt.n_ways * W_cols                       # a session's nn.Linear(640, n_ways) init
            if use_pull and opt.attraction_override == "mapping_linear_label2image":
                skip += lang_puller.novel_embeds.size(1) * W_cols + W_cols   # LinearMap re-created every session
            fwds = []
            for j in range(idx + 1, iter_num):
                fwds.append((skip, n_sup))
                if opt.memory_replay == 1:
                    fwds.append((0, n_mem + 25 * (j - idx)))
            net.engine().start_mask_prefetch(fwds)
        tp1 = _mark()
        W = net.classifier.weight.data
        reserve = novel_weight_to_reserve.contiguous() if (opt.lmbd_reg_novel is not None and idx > 0) else None
        head = ops.HeadSession(
            f_train, n_sup, 0, support_ys_d, W, n_base_cls, len(novel_labels), n_memory=n_mem, memory_row0=n_sup,
            labels_memory=memory.labels if has_mem else None,
            base_weight=base_weight if opt.lmbd_reg_transform_w is not None else None, reserve_weight=reserve,
            pull_mode=pull_mode, pull=pull_t, q_rows=q_rows,
            lmbd_base=opt.lmbd_reg_transform_w or 0.0, lmbd_novel=opt.lmbd_reg_novel or 0.0,
            gamma=opt.label_pull if use_pull else 0.0, adam=bool(opt.adam), lr=opt.learning_rate, momentum=opt.momentum,
            weight_decay=0.0005 if opt.adam else opt.weight_decay,   # get_optim (eval/util.py:92-102) hard-codes Adam's
            stable=opt.stable, convergence_epsilon=opt.convergence_epsilon,
            stable_epochs=opt.stable_epochs, target_train_loss=opt.target_train_loss,
            min_novel_epochs=opt.min_novel_epochs, max_novel_epochs=opt.max_novel_epochs)
        head.run(1, defer=True)
        net.eval()                                  # validate()'s side effect after epoch 1
        tp2 = _mark()

        # ---- eval-mode feature cache, part A: the rows the head trains on (support | memory) ----
        with torch.no_grad():
            cache = net.engine().eval_features(torch.cat([support_xs_d] + ([memory.data] if has_mem else []), 0))
        tm['backbone_imgs'] += cache.shape[0]
        tp3 = _mark()

        # ---- epochs 2.. on the device until the stopping rule fires (chained to epoch 1 on the device) ----
        ops.align_runs()     # (several runs sharing this GPU start their head loops together; no-op for a run alone)
        tp3h = _mark()
        head.run(max(opt.max_novel_epochs - 1, 1), feat=cache, support_row0=0, memory_row0=n_sup, defer=True)
        head.post()          # status / trace copies are queued HERE: the host reads them while part B below still runs
        tp4 = _mark()

        # ---- part B, queued behind the head loop: features of the queries of every session so far | base batch, and the
        # scoring of the last epoch - validate (:321-326) + eval_base (:362-367) - in ONE launch over those rows.  Nothing
        # of the next session's set-up depends on it, so it is not waited for here: the host reads the epoch count as soon
        # as the head loop ends and prepares session idx+2 while the device works through part B (finish_session below).
        n_query = sum(q.shape[0] for q in novel_query_collection)
        with torch.no_grad():
            cache_q = net.engine().eval_features(torch.cat(novel_query_collection + [base_x], 0))
        tm['backbone_imgs'] += cache_q.shape[0]
        tp4b = _mark()
        labels_all = torch.cat(novel_query_collection_id + [base_y])
        confusion = torch.zeros((100, 100), dtype=torch.int64, device=dev)   # (gold id, predicted id) of this session
        scored = ops.eval_logits(cache_q, net.classifier.weight.detach(), labels_all, confusion)
        confusion_run += confusion
        tp5 = _mark()
        phase_events += [('train_pass', tp0, tp1), ('head1', tp1, tp2), ('cache', tp2, tp3), ('head', tp3h, tp4),
                         ('cache', tp4, tp4b), ('score', tp4b, tp5)]

        # the previous session's scores have long been computed: finish its bookkeeping (and its prints, in the reference's
        # order) before this session's own results are read
        if pending_finish is not None:
            pending_finish()
            pending_finish = None
            sys.stdout.write("".join(held_prints))
            held_prints.clear()

        # ---- the session's one wait: for the head loop (epoch counts, loss trace) ----
        head.collect()
        rescored = False
        while not head.stopped:                     # (only if the chained launch ran out of epochs before the rule fired)
            head.run(max(opt.max_novel_epochs - head.epochs, 1))
            rescored = True
        if rescored:
            confusion_run -= confusion
            confusion.zero_()
            scored = ops.eval_logits(cache_q, net.classifier.weight.detach(), labels_all, confusion)
            confusion_run += confusion
        pred_host = torch.empty(scored["pred"].shape, dtype=torch.int32, pin_memory=True)
        pred_host.copy_(scored["pred"], non_blocking=True)
        scored_ev = torch.cuda.Event(blocking=True)
        scored_ev.record()
        trace = torch.cat(head.traces, 0).numpy()
        epoch = head.epochs + 1
        tm['train_s'] += time.perf_counter() - t_train0
        tm['steps'] += head.epochs
        # the reference's every-10th-epoch log lines, formatted in one go (fp32 percentages like `percent`)
        es = np.arange(10, head.epochs + 1, 10)
        if es.size:
            to_pct = np.float32(100.0 / n_sup)
            acc1 = trace[es - 1, 6].astype(np.float32) * to_pct
            acc5 = trace[es - 1, 7].astype(np.float32) * to_pct
            lines = []
            for k, e in enumerate(es):
                if use_pull:
                    lines.append("PULL:  %s\n" % float(trace[e - 1, 5]))
                lines.append('Novel Epoch {:4d}\tTrain Loss {:10.4f}\tAcc@1 {:10.3f}\tAcc@5 {:10.3f}\n'.format(
                    int(e), float(trace[e - 1, 0]), acc1[k], acc5[k]))
            sys.stdout.write("".join(lines))

        # ---- the reference's forward-call count, for DropBlock's schedule ----
        n_calls = head.epochs * (1 + (1 if has_mem else 0) + len(novel_query_collection)) + 1
        done = next(iter(net.block_counters().values())) - counters0
        net.advance_block_counters(n_calls - done)

        inds = None
        if opt.memory_replay:
            inds = rng.np_random().choice(opt.n_shots, opt.memory_replay)
            margin = 5 * np.arange(5)
            offset = np.arange(0, 125, 25)
            inds = np.tile(margin + inds, (5, 1)) + (np.tile(offset, (5, 1))).T
            inds = inds.flatten()
            inds_d = torch.from_numpy(inds).to(dev)
            memory.additems(support_xs_d[inds_d, :], support_ys_d[inds_d])

        def finish_session(scored=scored, scored_ev=scored_ev, pred_host=pred_host, n_queries=[q.shape[0] for q in
                           novel_query_collection], query_ids=list(query_ids_host), confusion=confusion, trace=trace,
                           epoch=epoch, epochs=head.epochs, novel_labels=novel_labels, vocab_base=list(vocab_base),
                           vocab_novel=list(vocab_novel), inds=inds, n_query=n_query, cache=cache, f_train=f_train,
                           W_snap=None if getattr(opt, 'light_record', False) else net.classifier.weight.detach().clone(),
                           bn_snap=None if getattr(opt, 'light_record', False) else {
                               k: v.detach().clone() for k, v in net.state_dict().items()
                               if 'running_' in k or 'num_batches_tracked' in k}):
            """Accuracy bookkeeping of a session once its scores are on the host (language_eval.py:321-404)."""
            scored_ev.synchronize()
            pred_all = pred_host.numpy().astype(np.int64)
            test_acc, query_ys_pred, query_logits = [], [], []
            r0 = 0
            for nq, qy_h in zip(n_queries, query_ids):
                pred = pred_all[r0:r0 + nq]
                test_acc.append(percent(int((pred == qy_h).sum()), nq)[0])
                query_ys_pred.append(pred)
                query_logits.append(scored["logits"][r0:r0 + nq])
                r0 += nq
            base_pred = pred_all[r0:r0 + base_x.shape[0]]
            acc_base_ = np.mean([percent(int((base_pred == base_y_host).sum()), base_x.shape[0])[0].item()])
            record['confusion_last'] = confusion
            record['confusion'] = confusion_run
            tm['images_scored'] += n_query + base_x.shape[0]

            test_acc = [round(i.item(), 2) for i in test_acc]
            print("Novel session accuracies: ", test_acc)
            novel_session_acc = list(test_acc)
            test_acc = np.array(test_acc).mean()

            acc_base.update(acc_base_)
            acc_novel.update(test_acc)
            w1 = 60 if opt.dataset == "miniImageNet" else 200
            w2 = len(vocab_base) + len(vocab_novel) - 60
            weighted_avg = (w1 * acc_base_ + w2 * test_acc) / (w1 + w2)
            weighted_avg_l.append(round(weighted_avg, 2))
            acc_novel_list.append(round(test_acc, 2))
            acc_base_list.append(round(acc_base_, 2))
            print(f"***Running weighted avg: {weighted_avg}")
            log_episode(novel_labels, vocab_novel, epoch, test_acc, acc_base_, acc_base.avg, acc_novel.avg)

            sess = dict(
                epochs=epochs, terms=trace.astype(np.float64),
                novel_session_acc=novel_session_acc, query_pred=[torch.from_numpy(p) for p in query_ys_pred],
                query_logits=query_logits, base_pred=torch.from_numpy(base_pred), acc_base=float(acc_base_),
                memory_inds=inds.copy() if inds is not None else None, vocab_novel=list(vocab_novel))
            if W_snap is not None:   # snapshots for the parity tests (~100 small device copies per session)
                sess.update(W=W_snap, probe_feat=cache[:8].clone(), train_feat=f_train[:8].clone(), bn=bn_snap)
            record['sessions'].append(sess)

        pending_finish = finish_session

    if pending_finish is not None:
        pending_finish()
        sys.stdout.write("".join(held_prints))
    torch.cuda.current_stream().synchronize()
    for name, e0, e1 in phase_events:
        ph[name] += e0.elapsed_time(e1) * 1e-3
    tm['score_s'] = ph['score']
    record.update(weighted=weighted_avg_l, novel=acc_novel_list, base=acc_base_list, acc_novel_avg=acc_novel.avg,
                  acc_base_avg=acc_base.avg, counters=net.block_counters())
    print("Overall continual accuracies: ", weighted_avg_l)
    print("Novel only incremental: ", acc_novel_list)
    print("Base only incremental: ", acc_base_list)
    return acc_novel.avg, acc_base.avg
