"""Host-side logic that needs no GPU: drop-in module surface, RNG parity of model construction, synthetic world,
label-embedding loader, session bookkeeping helpers."""
import os

import numpy as np
import pytest
import torch


def test_model_surface_and_state_dict_keys():
    """State-dict key set of the reference ResNet-18: 22 convs + 44 BN affine + classifier.weight, 66 BN
    buffers; same RNG consumption as the oracle's restatement of ResNet.__init__."""
    import models
    from models.util import create_model
    from oracle import init as oinit
    from srb200 import synthetic
    assert models.model_pool == ['resnet12', 'resnet18'] and set(models.model_dict) == set(models.model_pool)
    opt = synthetic.default_opt(3)
    net = synthetic.init_model(create_model, opt, 3)
    sd = net.state_dict()
    ref = oinit.init_state_dict(3)
    assert list(sd.keys()) == list(ref.keys())
    params = dict(net.named_parameters())
    assert len(params) == 22 + 44 + 1 and 'classifier.weight' in params and 'classifier.bias' not in sd
    assert sum(1 for k in sd if 'running_' in k or 'num_batches' in k) == 66
    for k in sd:
        assert torch.equal(sd[k], ref[k]), k
    assert net.num_classes == 60 and net.classifier.weight.shape == (60, 640)


def test_augment_and_base_weights_cpu():
    from models.util import create_model
    from srb200 import synthetic
    opt = synthetic.default_opt(1)
    net = synthetic.init_model(create_model, opt, 1)
    w0, b0 = net._get_base_weights()
    assert b0 is None and not w0.requires_grad and w0.data_ptr() != net.classifier.weight.data_ptr()
    torch.manual_seed(7)
    net.augment_base_classifier_(5)
    torch.manual_seed(7)
    expect = torch.nn.Linear(640, 5, bias=False).weight.detach()
    W = net.classifier.weight
    assert W.shape == (65, 640) and W.requires_grad and isinstance(W, torch.nn.Parameter)
    assert torch.equal(W[:60].detach(), w0) and torch.equal(W[60:].detach(), expect)
    assert net.num_classes == 60          # stays 60, as in the reference
    assert [n for n, _ in net.named_parameters() if n.startswith('classifier')] == ['classifier.weight']
    import copy
    net2 = copy.deepcopy(net)
    assert torch.equal(net2.classifier.weight, net.classifier.weight)
    counters = net.block_counters()
    assert list(counters) == ['layer1.0', 'layer2.0', 'layer3.0', 'layer3.1', 'layer4.0', 'layer4.1']
    net.advance_block_counters(7)
    assert set(net.block_counters().values()) == {7}


def test_block_plan_matches_reference_quirks():
    from models.util import create_model
    from oracle import backbone as obb
    from srb200 import synthetic
    net = synthetic.init_model(create_model, synthetic.default_opt(1), 1)
    plan = obb.block_plan('resnet18', True)
    mine = net._blocks()
    for a, b in zip(plan, mine):
        for k in ('prefix', 'cin', 'cout', 'pool', 'downsample', 'drop_block', 'block_size'):
            assert a[k] == b[k], (a['prefix'], k)
    # layerX.0 of a two-block stage is plain dropout (use_se lands in the drop_block slot), the last block is DropBlock
    assert [b['drop_block'] for b in plan] == [False, False, False, True, False, True]
    opt5 = synthetic.default_opt(1, no_dropblock=False)
    net5 = synthetic.init_model(create_model, opt5, 1)
    assert [b['block_size'] for b in net5._blocks()] == [1, 1, 1, 5, 1, 5]


def test_synthetic_world_contract():
    from srb200 import synthetic
    w = synthetic.make_world(4, n_sessions=3, n_base_batch=16)
    base, sessions = synthetic.class_split(4)
    assert len(base) == 60 and len(set(base)) == 60 and all(len(s) == 5 for s in sessions)
    assert not (set(base) & set(np.concatenate(sessions)))
    sx, sy, qx, qy = w.meta_valloader.batches[1]
    assert sx.shape == (1, 125, 3, 84, 84) and qx.shape == (1, 125, 3, 84, 84)
    assert np.array_equal(sy.view(-1).numpy(), np.tile(np.repeat(sessions[1], 5), 5))
    assert np.array_equal(qy.view(-1).numpy(), np.repeat(sessions[1], 25))
    l2h = w.base_val_loader.dataset.label2human
    assert len(l2h) == 100 and sum(1 for n in l2h if n) == 60
    w2 = synthetic.make_world(4, n_sessions=3, n_base_batch=16)
    assert torch.equal(w2.meta_valloader.batches[2][0], w.meta_valloader.batches[2][0])      # deterministic
    from eval.util import drop_a_dim, get_vocabs
    a, b, c, d = drop_a_dim(w.meta_valloader.batches[0])
    assert a.shape == (125, 3, 84, 84) and b.dtype == np.int64
    vb, va, vn, o2i = get_vocabs(w.base_val_loader, w.meta_valloader, d)
    assert len(vb) == 60 and len(vn) == 5 and sorted(o2i.values()) == [60, 61, 62, 63, 64]


def test_get_embeds_matches_oracle(word_embed_dir):
    from models.util import get_embeds
    from oracle import regularizer as rg
    from srb200 import synthetic
    path = os.path.join(word_embed_dir, "miniImageNet_dim500.pickle")
    a = get_embeds(path, synthetic.LABELS)
    b = rg.get_embeds(path, synthetic.LABELS)
    assert a.dtype == b.dtype == torch.float64 and torch.equal(a, b)
    assert float(a[synthetic.LABELS.index('komondor')].abs().sum()) == 0.0


def test_percent_rounding_like_reference():
    from eval.util import percent
    from oracle.session import accuracy
    out = torch.zeros(125, 7)
    out[:, 3] = 1.0
    tgt = torch.full((125,), 3, dtype=torch.int64)
    tgt[:38] = 2
    ref = accuracy(out, tgt, (1,))[0]
    assert torch.equal(percent(87, 125), ref)


def test_memory_dropin():
    from dataset.memory import Memory
    m = Memory()
    assert len(m) == 0
    m.additems(torch.zeros(25, 3, 4, 4), torch.arange(25))
    m.additems(torch.ones(25, 3, 4, 4), torch.arange(25))
    assert len(m) == 50 and m.data.shape == (50, 3, 4, 4) and m[30][1].item() == 5


def test_unsupported_options_fail_loudly():
    from eval.language_eval import few_shot_finetune_incremental_test
    from models.util import create_model
    from srb200 import synthetic
    w = synthetic.make_world(1, n_sessions=1, n_base_batch=4)
    net = synthetic.init_model(create_model, w.opt, 1)
    w.opt.track_weights = True
    with pytest.raises(NotImplementedError):
        few_shot_finetune_incremental_test(net, {}, None, w.meta_valloader, w.base_val_loader, w.opt)
    w.opt.track_weights = False
    w.opt.freeze_backbone_at = 3
    with pytest.raises(NotImplementedError):
        few_shot_finetune_incremental_test(net, {}, None, w.meta_valloader, w.base_val_loader, w.opt)


def test_episode_front_end_matches_the_reference_sampler(tmp_path):
    """dataset/mini_imagenet.py (image store split + episode sampler, no PIL) against tests/golden/episodes.pt, recorded by
    oracle/make_episode_golden.py from the UNMODIFIED reference classes on the same synthetic store: same base test set,
    same base exemplars, same eight disjoint novel sessions (which images, in which order, with which labels), and
    bit-identical normalised pixels; the raw uint8 path returns the same images."""
    import argparse
    import numpy as np
    import torch
    from dataset.mini_imagenet import ImageNet, MetaImageNet
    from dataset.transform_cfg import transforms_test_options
    from srb200 import synthetic
    plain = transforms_test_options['A'][1]      # ToTensor + Normalize (what the golden run passed on both branches)
    gold = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "episodes.pt"), weights_only=False)
    root = synthetic.write_image_store(str(tmp_path / "store"))
    for seed, g in gold.items():
        a = argparse.Namespace(data_root=root, data_aug=False, set_seed=seed, continual=True, n_ways=5, n_shots=5, n_queries=15,
                               n_test_runs=8, eval_mode="few-shot-incremental-fine-tune", n_aug_support_samples=5,
                               n_base_aug_support_samples=1, n_base_support_samples=1)
        base = ImageNet(args=a, split='train', phase='test')
        assert len(base) == g['base_len'] and np.array_equal(np.asarray(base.labels), g['base_labels'])
        assert list(base.label2human) == g['label2human']
        probe = list(range(0, len(base), 97))
        raw_base = ImageNet(args=a, split='train', phase='test', raw=True)
        ids = synthetic.image_ids(torch.stack([raw_base[i][0] for i in probe]).numpy())
        assert np.array_equal(ids, g['base_probe_ids'])
        assert torch.equal(torch.stack([base[i][0] for i in probe[:3]]), g['base_probe_x'])
        assert all(base[i][1] == base.labels[i] - min(base.labels) and base[i][2] == i for i in probe[:5])

        for raw in (False, True):
            sup = MetaImageNet(args=a, split='train', phase='train', train_transform=plain, test_transform=plain,
                               fix_seed=True, use_episodes=False, raw=raw)
            assert len(sup) == g['exemplar_len']
            for item, want in zip((0, 3), g['exemplars']):
                sx, sy, qx, qy = sup[item]
                assert np.array_equal(np.asarray(sy), want['ys']) and np.array_equal(np.asarray(qy), want['ys'])
                if raw:
                    assert sx.dtype == torch.uint8 and np.array_equal(synthetic.image_ids(sx.numpy()), want['ids'])
                else:
                    assert sx.dtype == torch.float32 and tuple(sx.shape[1:]) == (3, 4, 4)
            val = MetaImageNet(args=a, split='val', train_transform=plain, test_transform=plain, fix_seed=True,
                               use_episodes=False, disjoint_classes=True, raw=raw)
            assert len(val) == g['val_len'] and list(val.label2human) == g['val_label2human']
            for item, want in enumerate(g['sessions']):
                sx, sy, qx, qy = val[item]
                assert np.array_equal(np.asarray(sy), want['sup_ys']) and np.array_equal(np.asarray(qy), want['qry_ys'])
                if raw:
                    assert np.array_equal(synthetic.image_ids(sx.numpy()), want['sup_ids'])
                    assert np.array_equal(synthetic.image_ids(qx.numpy()), want['qry_ids'])
                elif item == 0:
                    assert torch.equal(sx[:10], want['sup_x'])


def test_crop_flip_parameters_reproduce_torchvision_on_the_host():
    """dataset.transform_cfg.draw_crop_flip draws what torchvision's RandomCrop(84, padding=8) + RandomHorizontalFlip draw
    (same order, same generator): applying its parameters with plain NumPy indexing gives the pixels of the reference's
    support transform, and both leave torch's CPU generator in the same state.  (The GPU kernel applies the same
    parameters: tests/test_gpu_kernels.py::test_support_augmentation_on_device_matches_torchvision.)"""
    from dataset import transform_cfg
    from dataset.mini_imagenet import _normalise
    support_tf = transform_cfg.transforms_test_options['A'][0]
    if support_tf is None:
        pytest.skip("torchvision / PIL not installed")
    rng = np.random.RandomState(1)
    x8 = rng.randint(0, 256, size=(12, 84, 84, 3)).astype(np.uint8)
    torch.manual_seed(99)
    want = torch.stack([support_tf(img) for img in x8])
    state = torch.get_rng_state()
    torch.manual_seed(99)
    ij, flip = transform_cfg.draw_crop_flip(12)
    assert torch.equal(torch.get_rng_state(), state)
    padded = np.zeros((12, 100, 100, 3), dtype=np.uint8)
    padded[:, 8:92, 8:92] = x8
    out = np.stack([padded[k, ij[k, 0]:ij[k, 0] + 84, ij[k, 1]:ij[k, 1] + 84] for k in range(12)])
    out = np.stack([o[:, ::-1] if flip[k] else o for k, o in enumerate(out)])
    assert torch.equal(_normalise(out), want)


def test_rng_thread_share_between_local_ranks(monkeypatch):
    """dist.init() divides the host cores between the local ranks for the mask-generator walkers."""
    from srb200 import dist as sdist
    monkeypatch.delenv("SRB_RNG_THREADS", raising=False)
    monkeypatch.setenv("LOCAL_WORLD_SIZE", "8")
    monkeypatch.setenv("WORLD_SIZE", "1")
    monkeypatch.setattr(os, "sched_getaffinity", lambda pid: set(range(32)), raising=False)
    sdist.init()
    assert os.environ["SRB_RNG_THREADS"] == "2"          # 32 cores / 8 ranks = 4 per rank: two walkers, never fewer
    monkeypatch.delenv("SRB_RNG_THREADS", raising=False)
    monkeypatch.setenv("LOCAL_WORLD_SIZE", "1")
    sdist.init()
    assert os.environ["SRB_RNG_THREADS"] == "4"          # a rank with the whole host still uses at most four
    monkeypatch.setenv("SRB_RNG_THREADS", "3")
    sdist.init()
    assert os.environ["SRB_RNG_THREADS"] == "3"          # an explicit setting wins


def test_support_augmentation_is_never_dropped(tmp_path):
    """The reference's support transform (RandomCrop(84, padding 8) + RandomHorizontalFlip, transform_cfg.py:32-40) on the
    uint8 front end: MetaImageNet(raw=True) with transforms_test_options['A'] returns the SAME pixels torchvision produces
    (before ToTensor / Normalize) and leaves torch's generator in the same state; a transform it cannot express on uint8
    (ColorJitter, the reference's default when train_transform is None) raises instead of being skipped."""
    import argparse
    import numpy as np
    import pytest
    import torch
    from dataset.mini_imagenet import MetaImageNet
    from dataset.transform_cfg import transforms_test_options, mean, std
    from srb200 import synthetic
    root = synthetic.write_image_store(str(tmp_path / "store"), side=84, per_class=30, n_classes=100, light=True)
    a = argparse.Namespace(data_root=root, data_aug=False, set_seed=3, continual=True, n_ways=5, n_shots=5, n_queries=15,
                           n_test_runs=8, eval_mode="few-shot-incremental-fine-tune", n_aug_support_samples=5,
                           n_base_aug_support_samples=1, n_base_support_samples=1)
    sup_t, qry_t = transforms_test_options['A']
    ref = MetaImageNet(args=a, split='val', train_transform=sup_t, test_transform=qry_t, fix_seed=True, disjoint_classes=True)
    raw = MetaImageNet(args=a, split='val', train_transform=sup_t, test_transform=qry_t, fix_seed=True, disjoint_classes=True,
                       raw=True)
    torch.manual_seed(7)
    sx, sy, qx, qy = ref[0]
    state_ref = torch.get_rng_state()
    torch.manual_seed(7)
    rx, ry, rqx, rqy = raw[0]
    assert torch.equal(torch.get_rng_state(), state_ref)
    assert rx.dtype == torch.uint8 and tuple(rx.shape) == (125, 84, 84, 3)
    m = torch.tensor(mean).view(1, 3, 1, 1)
    sd = torch.tensor(std).view(1, 3, 1, 1)
    want = rx.permute(0, 3, 1, 2).float().div(255).sub(m).div(sd)
    assert torch.equal(want, sx)                       # same crop, same flip, same pixels
    assert not torch.equal(rx[:25], rx[25:50])         # the tiled copies really are augmented differently
    assert torch.equal(rqx.permute(0, 3, 1, 2).float().div(255).sub(m).div(sd), qx)
    with pytest.raises(NotImplementedError):
        MetaImageNet(args=a, split='val', fix_seed=True, disjoint_classes=True, raw=True)[0]


def test_reference_format_checkpoint_round_trip(tmp_path):
    """a checkpoint in the layout train_supervised.py:194-202 writes (argparse.Namespace opt, numpy-keyed
    training_classes) is readable by the callers' PLAIN torch.load(path) once the shadow `models` package is imported
    (torch >= 2.6 defaults to weights_only=True and would reject it), and its state dict loads into the shadow model."""
    import argparse
    import numpy as np
    import torch
    import models                                      # registers the safe globals
    from models.util import create_model
    from srb200 import synthetic
    from srb200.checkpoint import load_backbone, load_reference_checkpoint, save_reference_checkpoint
    opt = synthetic.default_opt(1)
    net = synthetic.init_model(create_model, opt, 1)
    path = str(tmp_path / "resnet18_last.pth")
    mapping = {'map.weight': torch.randn(640, 300), 'map.bias': torch.randn(640)}
    save_reference_checkpoint(path, net, {np.int64(7 + i): np.int64(i) for i in range(60)}, synthetic.LABELS,
                              opt=argparse.Namespace(model='resnet18', continual=True), mapping=mapping)
    ckpt = torch.load(path)                            # exactly what eval_incremental.py:86 does
    assert set(ckpt) == {'opt', 'model', 'training_classes', 'label2human', 'mapping_linear_label2image'}
    assert all(isinstance(k, np.int64) for k in ckpt['training_classes']) and ckpt['training_classes'][np.int64(7)] == 0
    assert ckpt['opt'].model == 'resnet18' and ckpt['label2human'] == synthetic.LABELS
    assert 'classifier.bias' not in ckpt['model'] and len(ckpt['model']) == 67 + 66     # SURVEY 8b key set
    other = synthetic.init_model(create_model, opt, 2)
    load_backbone(other, load_reference_checkpoint(path))
    for (k, a), (_, b) in zip(net.state_dict().items(), other.state_dict().items()):
        assert torch.equal(a, b), k
    opt_b = synthetic.default_opt(1, linear_bias=True)
    import pytest
    with pytest.raises(ValueError):
        load_backbone(synthetic.init_model(create_model, opt_b, 1), ckpt)


def test_lockstep_rendezvous_protocol():
    """srb200.concurrent.Lockstep (the rendezvous that makes the head loops of the runs sharing a GPU start together): every
    align() returns only after all runs of the group have reached it, each run waits for exactly the events the OTHER runs
    recorded in that round, a run that leaves (finished or failed) releases the rest, and a group of one never waits."""
    import threading
    from srb200.concurrent import Lockstep
    n, rounds = 3, 5
    log = [[] for _ in range(n)]          # per run: the events it waited for, round by round
    counter = [0] * n
    tls = threading.local()

    def record():
        counter[tls.i] += 1
        return (tls.i, counter[tls.i])

    def wait(ev):
        log[tls.i].append(ev)

    group = Lockstep(n, record=record, wait=wait)
    arrived = []

    def run(i):
        tls.i = i
        for r in range(rounds):
            arrived.append((r, i))
            group.align(slot=i)
            # nobody is past round r before everybody has arrived at round r
            assert {(r, j) for j in range(n)} <= set(arrived)
        if i == 0:
            group.leave()

    ts = [threading.Thread(target=run, args=(i,)) for i in range(n)]
    for t in ts:
        t.start()
    for t in ts:
        t.join(timeout=30)
    assert not any(t.is_alive() for t in ts)
    for i in range(n):
        assert log[i] == [(j, r + 1) for r in range(rounds) for j in range(n) if j != i]
    # after a run has left, the others pass straight through
    assert group.broken
    group.align(slot=1)
    # a failing run must not keep the others waiting: leave() breaks a barrier that is being waited on
    g2 = Lockstep(2, record=lambda: None, wait=lambda e: None)
    done = []
    t = threading.Thread(target=lambda: (g2.align(slot=0), done.append(1)))
    t.start()
    g2.leave()
    t.join(timeout=30)
    assert done == [1] and g2.broken
    # a group of one never waits
    g1 = Lockstep(1, record=lambda: (_ for _ in ()).throw(AssertionError("no event for a single run")), wait=None)
    g1.align(slot=0)
