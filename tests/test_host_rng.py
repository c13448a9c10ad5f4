"""The host replay of PyTorch's CPU Bernoulli stream (sr_host_bernoulli) and the mask prefetcher built on it."""
import ctypes as C

import pytest
import torch

from srb200 import host_rng, _lib as L


@pytest.fixture(autouse=True)
def _need_replay():
    if not host_rng.replay_available():
        pytest.skip("this torch build does not use the serial mt19937 bernoulli path; torch itself draws the masks")


@pytest.mark.parametrize("kind,p,shape", [(0, 0.9, (7, 64, 21, 21)), (0, 0.5, (1, 3, 1001)), (1, 0.00123, (5, 320, 10, 10)),
                                          (1, 0.5, (100003,)), (0, 1.0, (1000,)), (1, 0.0, (1000,))])
def test_replay_equals_torch(kind, p, shape):
    torch.manual_seed(77)
    torch.rand(5)
    s0 = torch.get_rng_state()
    want = host_rng._torch_draw(shape, p, kind)
    s_want = torch.get_rng_state()
    torch.set_rng_state(s0)
    got, ones = host_rng._replay_draw(shape, p, kind)
    assert torch.equal(want, got) and ones == int(want.sum())
    assert torch.equal(torch.get_rng_state(), s_want)
    # and the stream continues identically afterwards (the next session's nn.Linear init)
    a = torch.nn.Linear(640, 5, bias=False).weight.clone()
    torch.set_rng_state(s_want)
    b = torch.nn.Linear(640, 5, bias=False).weight.clone()
    assert torch.equal(a, b)


def test_dropout_mask_equals_functional_dropout():
    import torch.nn.functional as F
    shape = (9, 64, 42, 42)
    torch.manual_seed(5)
    m = F.dropout(torch.ones(shape), 0.1, True)
    s1 = torch.get_rng_state()
    torch.manual_seed(5)
    keep, _ = host_rng.bernoulli_u8(shape, 1 - 0.1, 0)
    assert torch.equal((m != 0).to(torch.uint8), keep) and torch.equal(s1, torch.get_rng_state())
    assert float(m.max()) == float(torch.ones(1).div_(1 - 0.1))


def test_skip_matches_linear_init_and_dropblock_draws():
    torch.manual_seed(11)
    s0 = torch.get_rng_state()
    torch.nn.Linear(640, 5, bias=False)
    torch.distributions.Bernoulli(0.01).sample((3, 320, 10, 10))
    s1 = torch.get_rng_state()
    blob = s0.clone()
    n = 5 * 640 + 3 * 320 * 100
    assert L.load().sr_host_bernoulli(C.c_void_p(blob.data_ptr()), blob.numel(), 2, 0.0, n, None) == 0
    assert torch.equal(blob, s1)


def test_mask_prefetch_is_exact_and_self_checking():
    shapes = [(4, 64, 42, 42), (4, 160, 21, 21)]
    torch.manual_seed(21)
    start = torch.get_rng_state()
    # what the main thread will do later: Linear init, mask 0, a DropBlock draw, mask 1
    torch.nn.Linear(640, 5, bias=False)
    m0 = torch.empty(shapes[0], dtype=torch.uint8).bernoulli_(0.9)
    torch.distributions.Bernoulli(0.02).sample((4, 320, 10, 10))
    m1 = torch.empty(shapes[1], dtype=torch.uint8).bernoulli_(0.9)
    end = torch.get_rng_state()

    torch.set_rng_state(start)
    bufs = [torch.empty(s, dtype=torch.uint8) for s in shapes]
    pf = host_rng.MaskPrefetch([('skip', 3200), ('draw', 'a', bufs[0], 0.9), ('skip', 4 * 320 * 100), ('draw', 'b', bufs[1], 0.9)])
    torch.nn.Linear(640, 5, bias=False)
    got0 = pf.take('a', shapes[0])
    assert got0 is not None and torch.equal(got0[0], m0)
    torch.distributions.Bernoulli(0.02).sample((4, 320, 10, 10))
    got1 = pf.take('b', shapes[1])
    assert got1 is not None and torch.equal(got1[0], m1) and torch.equal(torch.get_rng_state(), end)

    # a consumer the plan did not foresee: the prefetch must refuse and leave the generator alone
    torch.set_rng_state(start)
    pf = host_rng.MaskPrefetch([('skip', 3200), ('draw', 'a', bufs[0], 0.9)])
    torch.nn.Linear(640, 5, bias=False)
    torch.rand(1)
    live = torch.get_rng_state()
    assert pf.take('a', shapes[0]) is None and torch.equal(torch.get_rng_state(), live)


def test_dropblock_keep_matches_the_torch_formulation():
    """sr_host_dropblock == 1 - (union of the seeds shifted over a bs x bs window), as _compute_block_mask builds it
    (reference models/resnet_language.py:327-352), and its return value is the count of kept positions."""
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(3)
    for bs, hs in ((5, 6), (3, 3), (1, 5), (2, 4)):
        seeds = (torch.rand(4, 7, hs, hs, generator=g) < 0.08).to(torch.uint8)
        left, right = int((bs - 1) / 2), int(bs / 2)
        padded = F.pad(seeds, (left, right, left, right))
        for i in range(bs):
            for j in range(bs):
                padded[:, :, i:i + hs, j:j + hs] = torch.maximum(padded[:, :, i:i + hs, j:j + hs], seeds)
        want = 1 - padded
        out = torch.empty(4, 7, hs + bs - 1, hs + bs - 1, dtype=torch.uint8)
        kept = host_rng.dropblock_keep(seeds.contiguous(), bs, out)
        assert torch.equal(out, want)
        assert kept == int(want.sum())
    # no seeds at all: everything is kept
    out = torch.zeros(2, 3, 7, 7, dtype=torch.uint8)
    assert host_rng.dropblock_keep(torch.zeros(2, 3, 5, 5, dtype=torch.uint8), 3, out) == out.numel() and bool(out.all())


def test_held_regions_can_be_drawn_off_the_live_generator():
    """A 'hold' step keeps the generator states around a region whose probability is not known yet (DropBlock seeds);
    drawing that region later from the held state gives exactly what torch draws live at that point, and the masks after
    it are still accepted."""
    torch.manual_seed(77)
    torch.rand(3)
    shape_a, shape_h, shape_b = (2, 8, 21, 21), (2, 320, 6, 6), (2, 16, 10, 10)
    bufs = [torch.empty(shape_a, dtype=torch.uint8), torch.empty(shape_b, dtype=torch.uint8)]
    n_h = 2 * 320 * 6 * 6
    pf = host_rng.MaskPrefetch([('draw', 'a', bufs[0], 0.9), ('hold', 'h', n_h), ('draw', 'b', bufs[1], 0.9)])
    if not pf.ok:
        pytest.skip("replay not available on this torch build")
    assert pf.take('a', shape_a) is not None
    before, after = pf.hold_state('h')
    assert torch.equal(torch.get_rng_state(), before)
    seeds = torch.empty(shape_h, dtype=torch.uint8)
    state = before.clone()
    ones = host_rng.replay_into(state, 0.0123, 1, seeds)
    assert torch.equal(state, after) and torch.equal(torch.get_rng_state(), before)      # the live generator did not move
    live = torch.bernoulli(torch.tensor(0.0123).expand(shape_h)).to(torch.uint8)            # what the reference draws here
    assert torch.equal(live, seeds) and ones == int(live.sum())
    assert torch.equal(torch.get_rng_state(), after)
    assert pf.take('b', shape_b) is not None


JW = 1 << 18   # SR_MT_JUMP_WORDS


@pytest.mark.parametrize("pre,n", [(0, 5), (7, 617), (7, 618), (100, 624 * 3 + 1), (623, JW - 1), (1, JW), (300, JW + 77),
                                   (17, 3 * JW + 12345), (5, 8 * JW - 3)])
def test_jump_ahead_equals_drawing(pre, n):
    """sr_host_mt_advance (polynomial jump-ahead, csrc/mt_jump.cpp) leaves the generator byte-identical to drawing n words
    one by one: against torch itself for short distances, against the replay's skip for long ones."""
    lib = L.load()
    table = torch.zeros((8, 624), dtype=torch.int32)
    assert lib.sr_mt_jump_table(C.c_void_p(table.data_ptr()), 8) == 0
    torch.manual_seed(11 + pre)
    if pre:
        torch.empty(pre, dtype=torch.int32).random_()
    s0 = torch.get_rng_state()
    a = s0.clone()
    assert lib.sr_host_mt_advance(C.c_void_p(a.data_ptr()), a.numel(), n, C.c_void_p(table.data_ptr()), 8) == 0
    b = s0.clone()
    assert lib.sr_host_bernoulli(C.c_void_p(b.data_ptr()), b.numel(), 2, 0.0, n, None) == 0
    assert torch.equal(a, b)
    if n < 5000:
        torch.empty(n, dtype=torch.int32).random_()
        assert torch.equal(torch.get_rng_state(), a)
    # a distance the table does not cover is refused, not approximated
    c = s0.clone()
    assert lib.sr_host_mt_advance(C.c_void_p(c.data_ptr()), c.numel(), 10 * JW, C.c_void_p(table.data_ptr()), 8) != 0
