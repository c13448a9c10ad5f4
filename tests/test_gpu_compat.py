"""The per-op (autograd) surface of the drop-in modules: an epoch loop written exactly like the reference's
(eval/language_eval.py:242-295: net(x) -> criterion -> regloss -> reglossnovel -> get_projected_weight -> loss1 ->
backward -> optimizer.step) must land on the same weights as the fused persistent head kernel."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _setup(word_embed_dir, **over):
    from models.util import create_model
    from srb200 import synthetic
    world = synthetic.make_world(2, n_sessions=1, n_base_batch=8, word_embed_path=word_embed_dir, **over)
    net = synthetic.init_model(create_model, world.opt, 2).cuda()
    return world, net


@pytest.mark.parametrize("mode", ["distance2subspace", "semantic", "mapping"])
def test_reference_style_loop_matches_fused_head(mode, word_embed_dir):
    from eval.util import drop_a_dim, freeze_backbone_weights, get_optim, get_vocabs
    from models.resnet_language import LangPuller
    from srb200 import ops, _lib as L
    over = {}
    if mode == "semantic":
        over = dict(attraction_override=None, glove=True, label_pull=0.2, temperature=3.0)
    elif mode == "mapping":
        over = dict(attraction_override="mapping_linear_label2image", glove=True, label_pull=0.1)
    world, net = _setup(word_embed_dir, **over)
    opt = world.opt
    ckpt = {}
    if mode == "mapping":
        g = torch.Generator().manual_seed(5)
        ckpt["mapping_linear_label2image"] = {"map.weight": torch.randn(640, 300, generator=g) * 0.05,
                                              "map.bias": torch.randn(640, generator=g) * 0.01}
    criterion = torch.nn.CrossEntropyLoss()
    base_weight, _ = net._get_base_weights()
    support_xs, support_ys, query_xs, query_ys = drop_a_dim(world.meta_valloader.batches[0])
    vocab_base, vocab_all, vocab_novel, orig2id = get_vocabs(world.base_val_loader, world.meta_valloader, query_ys)
    support_ys_id = torch.LongTensor([orig2id[y] for y in support_ys]).cuda()
    x = support_xs[:64].cuda()
    y = support_ys_id[:64]
    net.eval()
    torch.manual_seed(11)
    net.augment_base_classifier_(5)
    W_start = net.classifier.weight.detach().clone()
    lang_puller = LangPuller(opt, vocab_base, vocab_novel)
    if mode == "mapping":
        lang_puller.create_pulling_mapping(ckpt["mapping_linear_label2image"])
    pullers = lang_puller(base_weight[:60, :])
    optimizer = get_optim(net, opt)
    freeze_backbone_weights(net, opt, 1, exclude=["classifier"])
    feat = net.features(x).detach()
    losses = []
    epochs = 12
    for epoch in range(1, epochs + 1):                       # reference loop body, eval-mode features
        output = net(x)
        loss = criterion(output, y)
        loss = loss + net.regloss(opt.lmbd_reg_transform_w, base_weight, None)
        if opt.attraction_override == "distance2subspace":
            pullers = lang_puller.get_projected_weight(base_weight, net.classifier.weight[60:, :])
        loss = loss + lang_puller.loss1(opt.label_pull, pullers, net.classifier.weight[60:, :])
        optimizer.zero_grad()
        loss.backward()
        optimizer.step()
        losses.append(loss.item())
    W_loop = net.classifier.weight.detach().clone()

    Wf = W_start.clone()
    if mode == "distance2subspace":
        qt, q, _ = lang_puller.factor(base_weight)
        pm, pt = L.SR_PULL_PROJECT, qt
    else:
        q = 0
        pm, pt = L.SR_PULL_FIXED, pullers.detach().contiguous()
    hs = ops.HeadSession(feat, 64, 0, y, Wf, 60, 5, base_weight=base_weight.contiguous(), pull_mode=pm, pull=pt, q_rows=q,
                         lmbd_base=opt.lmbd_reg_transform_w, gamma=opt.label_pull, lr=opt.learning_rate,
                         momentum=opt.momentum, weight_decay=opt.weight_decay, stable=False, target_train_loss=-1.0,
                         min_novel_epochs=0, max_novel_epochs=10 ** 6)
    tr = hs.run(epochs)
    np.testing.assert_allclose(tr[:, 0].numpy(), np.asarray(losses), rtol=1e-5)
    err = ((Wf - W_loop).abs().max() / W_loop.abs().max()).item()
    assert err < 1e-5, err


def test_semantic_pullers_and_linear_map_vs_torch(word_embed_dir):
    """LangPuller.forward (softmax mix / LinearMap) against plain fp64 torch on the same embeddings."""
    from models.resnet_language import LangPuller
    from srb200 import synthetic
    opt = synthetic.default_opt(1, word_embed_path=word_embed_dir, attraction_override=None, glove=True, temperature=3.0)
    base, sessions = synthetic.class_split(1)
    vb = [synthetic.LABELS[c] for c in base]
    vn = [synthetic.LABELS[c] for c in sessions[0]]
    lp = LangPuller(opt, vb, vn)
    W0 = torch.randn(60, 640, device="cuda") * 0.04
    got = lp(W0)
    En, Eb = lp.novel_embeds.double(), lp.base_embeds.double()
    want = torch.softmax(En @ Eb.t() / 3.0, 1) @ W0.double()
    assert ((got.double() - want).abs().max() / want.abs().max()).item() < 1e-5
    sd = {"map.weight": torch.randn(640, 300) * 0.05, "map.bias": torch.randn(640) * 0.01}
    lp.create_pulling_mapping(sd)
    got = lp(W0)
    want = En @ sd["map.weight"].double().cuda().t() + sd["map.bias"].double().cuda()
    assert ((got.double() - want).abs().max() / want.abs().max()).item() < 1e-5


def test_validate_and_eval_base_api(word_embed_dir):
    """validate / eval_base keep the reference's return structure and leave the net in eval mode."""
    from eval.language_eval import eval_base, validate
    world, net = _setup(word_embed_dir)
    q = world.meta_valloader.batches[0][2].view(-1, 3, 84, 84)[:20]
    yq = torch.randint(0, 60, (20,))
    net.train()
    a1, a5, loss, pred = validate(q, yq, net, torch.nn.CrossEntropyLoss(), world.opt, 1)
    assert not net.training and pred.shape == (20,) and 0.0 <= float(a1) <= float(a5) <= 100.0 and loss > 0
    l1, l5, ll, lp = validate([q, q], [yq, yq], net, None, world.opt, 1)
    assert len(l1) == 2 and float(l1[0]) == float(a1) and (lp[1] == pred).all()
    acc = eval_base(net, world.base_val_loader.batches[0], None)
    logits = net(world.base_val_loader.batches[0][0].cuda())
    want = (logits.argmax(1).cpu() == world.base_val_loader.batches[0][1]).float().mean().item() * 100
    assert abs(acc - want) < 1e-4


def test_fit_linear_map_matches_learn_mapping_recipe(word_embed_dir):
    """learn_mapping.py:41-67 on the B200 kernels vs the same loop in fp64 torch: 200 full-batch SGD steps (lr 1.0,
    wd 5e-4, MSE) from the same init."""
    from srb200 import mapping
    g = torch.Generator().manual_seed(0)
    E = torch.randn(60, 300, generator=g) * 0.3
    T = torch.randn(60, 640, generator=g) * 0.05
    init = {'map.weight': torch.randn(640, 300, generator=g) * 0.02, 'map.bias': torch.zeros(640)}
    import time
    res = {}
    for fused in (True, False):
        mapping.fit_linear_map(E.cuda(), T.cuda(), epochs=5, lr=1.0, weight_decay=5e-4, init=init, fused=fused)      # warm-up
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        res[fused] = mapping.fit_linear_map(E.cuda(), T.cuda(), epochs=200, lr=1.0, weight_decay=5e-4, init=init, fused=fused)
        torch.cuda.synchronize()
        print("fit_linear_map (%s): %.1f us per full-batch step (60 x 300 -> 640)" %
              ("one launch: sr_fit_linear_map" if fused else "5 launches per step, host-bound", (time.perf_counter() - t0) * 1e6 / 200))
    Ec, Tc = E.cuda(), T.cuda()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    mapping.fit_linear_map(Ec, Tc, epochs=1000, lr=1.0, weight_decay=5e-4, init=init)
    e1.record()
    torch.cuda.synchronize()
    print("fit_linear_map, the reference's 1000 steps in one launch: %.2f ms of device time (%.2f us per step)" %
          (e0.elapsed_time(e1), e0.elapsed_time(e1)))
    sd, losses = res[True]
    # the one-launch fit against the per-op route (same arithmetic, different summation order)
    for k in sd:
        rel = ((sd[k] - res[False][0][k]).norm() / res[False][0][k].norm()).item()
        assert rel < 1e-5, (k, rel)
    np.testing.assert_allclose(losses, res[False][1], rtol=1e-5)
    W = init['map.weight'].double().clone().requires_grad_(True)
    b = init['map.bias'].double().clone().requires_grad_(True)
    opt = torch.optim.SGD([W, b], lr=1.0, weight_decay=5e-4)
    ref = []
    for _ in range(200):
        loss = torch.nn.functional.mse_loss(E.double() @ W.t() + b, T.double())
        opt.zero_grad()
        loss.backward()
        opt.step()
        ref.append(loss.item())
    np.testing.assert_allclose(losses, ref, rtol=1e-4)
    assert ((sd['map.weight'].cpu().double() - W.detach()).norm() / W.detach().norm()).item() < 1e-5
    assert ((sd['map.bias'].cpu().double() - b.detach()).norm() / b.detach().norm()).item() < 1e-4
    assert ((sd['map.weight'].cpu().double() - W.detach()).abs().max() / W.detach().abs().max()).item() < 1e-4
    assert ((sd['map.bias'].cpu().double() - b.detach()).abs().max() / (b.detach().abs().max() + 1e-12)).item() < 1e-3


def test_description_embedding_pullers(tmp_path, word_embed_dir):
    """Config 3 (ii): 768-d label-description embeddings through the softmax-mix puller and through a LinearMap fitted
    with the learn_mapping recipe; oracle = the restated LangPuller with the same embeddings assigned directly."""
    import pickle
    from models.resnet_language import LangPuller
    from oracle import regularizer as rg
    from srb200 import mapping, synthetic
    g = torch.Generator().manual_seed(4)
    table = {name: torch.randn(768, generator=g) * 0.2 for name in synthetic.LABELS}
    path = tmp_path / "miniImageNet_bert-base-cased_layer6.pickle"
    with open(path, "wb") as f:
        pickle.dump(table, f)
    base, sessions = synthetic.class_split(1)
    vb = [synthetic.LABELS[c] for c in base]
    vn = [synthetic.LABELS[c] for c in sessions[0]]
    opt = synthetic.default_opt(1, word_embed_path=word_embed_dir, attraction_override=None, temperature=3.0,
                                description_embed_path=str(path))
    lp = LangPuller(opt, vb, vn)
    assert lp.base_embeds.shape == (60, 768) and lp.novel_embeds.shape == (5, 768)
    W0 = torch.randn(60, 640, generator=g) * 0.04
    o = rg.Puller.__new__(rg.Puller)
    o.opt, o.mapping = opt, None
    o.base_embeds = torch.stack([table[n] for n in vb])
    o.novel_embeds = torch.stack([table[n] for n in vn])
    want = o.pullers(W0)
    got = lp(W0.cuda())
    assert ((got.cpu() - want).abs().max() / want.abs().max()).item() < 1e-5
    sd, _ = mapping.fit_linear_map(lp.base_embeds, W0.cuda(), epochs=50, seed=9)
    lp.create_pulling_mapping({k: v.cpu() for k, v in sd.items()})
    o.set_mapping({k: v.cpu() for k, v in sd.items()})
    assert ((lp(W0.cuda()).cpu() - o.pullers(W0)).abs().max() / o.pullers(W0).abs().max()).item() < 1e-5
