"""Keep-masks drawn on the GPU (csrc/mask.cu: jump-ahead walkers over torch's CPU mt19937 stream) against torch's own CPU
draws: bit-identical masks, DropBlock keep-masks / scale, and generator state (integer work: the bar is equality)."""
import pytest
import torch
import torch.nn.functional as F

from srb200 import device_rng

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _need_device_path():
    if not device_rng.available(torch.device("cuda", 0)):
        pytest.skip("device mask path disabled (SRB_MASKS=host or this torch build draws bernoulli_ differently)")


@pytest.mark.parametrize("batch,pre", [(3, 0), (20, 5), (37, 623)])
def test_forward_masks_equal_torch(batch, pre):
    """The six mask draws of one ResNet-18 train-mode forward (dropout after layer1 / layer2, DropBlock seeds in layers 3 and
    4) in one sr_device_bernoulli call: every byte equals what torch's CPU generator draws, the generator ends in the same
    state, and the next nn.Linear init (the next session's classifier rows) is the same."""
    dev = torch.device("cuda", 0)
    shapes = [(0, 0.9, (batch, 64, 42, 42)), (0, 0.9, (batch, 160, 21, 21)), (1, 0.0173, (batch, 320, 6, 6)),
              (1, 0.0173, (batch, 320, 6, 6)), (1, 0.004, (batch, 640, 1, 1)), (1, 0.5, (batch, 640, 1, 1))]
    torch.manual_seed(100 + batch)
    if pre:
        torch.empty(pre, dtype=torch.int32).random_()
    s0 = torch.get_rng_state()
    want = []
    for kind, p, shape in shapes:
        if kind == 0:
            want.append(torch.empty(shape, dtype=torch.uint8).bernoulli_(p))
        else:
            want.append(torch.bernoulli(torch.tensor(p).expand(shape)).to(torch.uint8))
    s_want = torch.get_rng_state()
    lin_want = torch.nn.Linear(640, 5, bias=False).weight.clone()
    torch.set_rng_state(s0)
    outs = [torch.empty(shape, dtype=torch.uint8, device=dev) for _, _, shape in shapes]
    device_rng.draw([(k, p, o) for (k, p, _), o in zip(shapes, outs)])
    assert torch.equal(torch.get_rng_state(), s_want)
    for w, o in zip(want, outs):
        assert torch.equal(w, o.cpu())
    assert torch.equal(torch.nn.Linear(640, 5, bias=False).weight, lin_want)


def test_long_stream_crosses_many_jump_boundaries():
    """9 M words = 35 walkers; checked against the host replay of the same stream (itself pinned to torch)."""
    from srb200 import host_rng
    if not host_rng.replay_available():
        pytest.skip("host replay unavailable")
    dev = torch.device("cuda", 0)
    torch.manual_seed(9)
    torch.rand(77)
    s0 = torch.get_rng_state()
    shape = (4500001,)
    want, _ = host_rng._replay_draw(shape, 0.9, 0)
    s_want = torch.get_rng_state()
    torch.set_rng_state(s0)
    out = torch.empty(shape, dtype=torch.uint8, device=dev)
    device_rng.draw([(0, 0.9, out)])
    assert torch.equal(torch.get_rng_state(), s_want)
    assert torch.equal(want, out.cpu())


def test_dropblock_on_device_matches_the_torch_formulation():
    """sr_dropblock_keep == 1 - (union of the seeds shifted over a bs x bs window) (resnet_language.py:327-352) and the scale
    is the reference's fp32 numel / kept."""
    g = torch.Generator().manual_seed(3)
    for bs, hs in ((5, 6), (3, 3), (1, 5), (2, 4)):
        seeds = (torch.rand(4, 7, hs, hs, generator=g) < 0.08).to(torch.uint8)
        left, right = int((bs - 1) / 2), int(bs / 2)
        padded = F.pad(seeds, (left, right, left, right))
        for i in range(bs):
            for j in range(bs):
                padded[:, :, i:i + hs, j:j + hs] = torch.maximum(padded[:, :, i:i + hs, j:j + hs], seeds)
        want = 1 - padded
        keep = torch.empty(4, 7, hs + bs - 1, hs + bs - 1, dtype=torch.uint8, device="cuda")
        scale = torch.zeros(4, dtype=torch.float32, device="cuda")
        device_rng.dropblock_keep(seeds.cuda().contiguous(), bs, keep, scale)
        assert torch.equal(keep.cpu(), want)
        ref_scale = torch.tensor(float(want.numel()), dtype=torch.float32) / torch.tensor(float(want.sum()), dtype=torch.float32)
        assert float(scale[0]) == float(ref_scale)
