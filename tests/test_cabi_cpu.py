"""CPU-side checks of the C-ABI boundary: the in-tree library loads, exports every symbol include/srb200.h declares,
the ctypes mirrors have the C structs' sizes, and the product path refuses CPU tensors (no fallback)."""
import ctypes
import os
import re
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "srb200.h")


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    g.build()
    from srb200 import _lib
    return _lib.load()


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sr_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(lib):
    from srb200 import _lib
    names = declared_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), "libsrb200.so does not export %s" % n
    assert sorted(_lib.EXPORTS) == names, "srb200._lib.EXPORTS is out of sync with include/srb200.h"


def test_version_and_error_string(lib):
    assert lib.sr_version() >= 100
    assert isinstance(lib.sr_last_error(), bytes)


def test_struct_layouts_match_c(tmp_path):
    from srb200 import _lib
    prog = tmp_path / "sizes.c"
    prog.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "srb200.h"\nint main(void){'
                    'printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu ", sizeof(sr_conv_panel), sizeof(sr_conv_args), sizeof(sr_bn_apply_args),'
                    'sizeof(sr_head_args), sizeof(sr_eval_args), offsetof(sr_head_args, convergence_epsilon),'
                    'offsetof(sr_head_args, workspace_bytes), sizeof(sr_train_block_args), offsetof(sr_train_block_args, stats));'
                    'printf("%zu %zu %zu\\n", sizeof(sr_eval_block), sizeof(sr_backbone_eval_args), offsetof(sr_backbone_eval_args, features));'
                    ' return 0;}\n')
    exe = tmp_path / "sizes"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(prog), "-o", str(exe)], check=True)
    out = [int(x) for x in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()]
    assert out[0] == ctypes.sizeof(_lib.ConvPanel)
    assert out[1] == ctypes.sizeof(_lib.ConvArgs)
    assert out[2] == ctypes.sizeof(_lib.BnApplyArgs)
    assert out[3] == ctypes.sizeof(_lib.HeadArgs)
    assert out[4] == ctypes.sizeof(_lib.EvalArgs)
    assert out[5] == _lib.HeadArgs.convergence_epsilon.offset
    assert out[6] == _lib.HeadArgs.workspace_bytes.offset
    assert out[7] == ctypes.sizeof(_lib.TrainBlockArgs) and out[8] == _lib.TrainBlockArgs.stats.offset
    assert out[9] == ctypes.sizeof(_lib.EvalBlock) and out[10] == ctypes.sizeof(_lib.BackboneEvalArgs)
    assert out[11] == _lib.BackboneEvalArgs.features.offset


def test_argument_validation_without_gpu(lib):
    """Bad arguments are rejected on the host before anything touches a device."""
    from srb200 import _lib
    a = _lib.ConvArgs()
    a.n_panels = 3
    assert lib.sr_conv(ctypes.byref(a), None) == -1
    assert b"n_panels" in lib.sr_last_error()
    h = _lib.HeadArgs()
    assert lib.sr_head_run(ctypes.byref(h), None) == -1
    assert lib.sr_pack_input(None, None, None, 1, 3, 84, 84, 16, None) == -1
    # the entry points added in round 2: mask generator, DropBlock, the one-launch mapping fit
    import torch
    state = torch.get_rng_state().clone()
    regs = (_lib.MaskRegion * 2)()
    assert lib.sr_device_bernoulli(None, 0, regs, 2, None, None, 0, None, 0, None) == -1
    regs[0].kind, regs[0].n, regs[0].out = 1, 3, 4096          # an odd word count before the last region is refused
    regs[1].kind, regs[1].n, regs[1].out = 0, 4, 8192
    assert lib.sr_device_bernoulli(ctypes.c_void_p(state.data_ptr()), state.numel(), regs, 2, None, None, 0, None, 0, None) == -1
    assert b"odd word count" in lib.sr_last_error()
    regs[0].n = 1 << 20                                         # more words than the (empty) jump table covers
    assert lib.sr_device_bernoulli(ctypes.c_void_p(state.data_ptr()), state.numel(), regs, 2, None, None, 0, None, 0, None) == -1
    assert b"jump polynomials" in lib.sr_last_error()
    assert torch.equal(state, torch.get_rng_state())           # a refused call leaves the generator state alone
    assert lib.sr_dropblock_keep(None, 1, 6, 6, 5, None, None, None) == -1
    assert lib.sr_fit_linear_map(None, None, None, None, 60, 300, 640, 10, 1.0, 5e-4, None, None, 0, None) == -1
    assert lib.sr_fit_linear_map_workspace_bytes(60, 300, 640, 1000) > 0
    assert lib.sr_fit_linear_map_workspace_bytes(4096, 4096, 640, 10) == 0     # does not fit in shared memory: per-op route
    assert lib.sr_host_mt_advance(None, 0, 5, None, 0) == -1


def test_no_cpu_fallback():
    """CPU tensors are refused loudly everywhere on the product path."""
    from models.util import create_model
    from srb200 import ops, synthetic
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.pack_input(torch.zeros(1, 3, 84, 84))
    opt = synthetic.default_opt(1)
    net = synthetic.init_model(create_model, opt, 1).eval()
    with pytest.raises(RuntimeError, match="CUDA"):
        net(torch.zeros(2, 3, 84, 84))
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.subspace_factor(torch.zeros(60, 640))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "subspace-reg_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), os.path.join(dp, f)


def _conv_plan(lib, B, H, cin, cout, epi=0, taps=9, second=None):
    from srb200 import _lib as L
    a = L.ConvArgs()
    a.batch, a.height, a.width, a.cout = B, H, H, cout
    a.n_panels = 1 if second is None else 2
    a.panel[0].cin_pad, a.panel[0].taps = (cin + 15) // 16 * 16, taps
    if second is not None:
        a.panel[1].cin_pad, a.panel[1].taps = (second + 15) // 16 * 16, 1
    a.epilogue = epi
    out = (ctypes.c_int32 * 16)()
    assert lib.sr_conv_plan(ctypes.byref(a), out) == 0, lib.sr_last_error()
    keys = "TW TH TN stacked tiles_w tiles_h n_cta n_splits tmem_stages reuse row_bytes ring stage_bytes dyn_smem ncb last_k".split()
    return dict(zip(keys, list(out)))


def test_conv_plan_for_every_backbone_layer(lib):
    """The host-side launch plan (no GPU needed) for every conv shape of the RFS ResNet-18 at the bench batch: the large
    maps must get the row-stacked one-image tile that tap reuse needs (a utilisation-first picker once chose 4x4x8 boxes
    there and silently disabled reuse), rows are 128 bytes from 64 channels up, the ring has >= 3 stages, >= 90 % of the
    UMMA rows are real pixels, pooled tiles keep their 2x2 windows inside a CTA, and everything fits in 227 KB."""
    ACT, POOL, AVG, RAW = 0, 1, 2, 3
    layers = [(84, 3, 64, ACT, None), (84, 64, 64, ACT, None), (84, 64, 64, POOL, 3),
              (42, 64, 160, ACT, None), (42, 160, 160, ACT, None), (42, 160, 160, POOL, 64),
              (21, 160, 320, ACT, None), (21, 320, 320, ACT, None), (21, 320, 320, POOL, 160),
              (10, 320, 320, ACT, None), (10, 320, 640, ACT, None), (10, 640, 640, ACT, None), (10, 640, 640, POOL, 320),
              (5, 640, 640, ACT, None), (5, 640, 640, AVG, None)]
    for B in (1024, 185, 2):
        for H, cin, cout, epi, second in layers:
            for mode in (epi, RAW):
                p = _conv_plan(lib, B, H, cin, cout, mode, 9, second if mode == epi else None)
                tag = (B, H, cin, cout, mode)
                assert p['n_cta'] * p['n_splits'] == cout and p['n_cta'] % 32 == 0 and p['n_cta'] <= 256, tag
                assert p['tmem_stages'] * 2 * p['n_cta'] <= 512, tag
                assert p['TW'] * p['TH'] * p['TN'] >= 115 and p['TW'] * p['TH'] * p['TN'] <= 128, tag
                assert p['ring'] >= 3 and p['dyn_smem'] <= 227 * 1024 - 1024, tag      # 1 KB of static shared memory
                assert p['row_bytes'] == (128 if cin >= 64 else 32), tag
                if mode == POOL:
                    assert p['stacked'] or p['TH'] % 2 == 0, tag
                if H >= 42:
                    assert (p['TW'], p['TH'], p['TN'], p['stacked']) == (42, 3, 1, 1), tag
                if H == 84:
                    assert p['reuse'] == 1, tag
                if cin == 160:
                    assert (p['ncb'], p['last_k']) == (3, 2), tag      # 64 + 64 + 32 channels: the zero K-steps are skipped


def test_integration_doc_lists_every_entry_point():
    """INTEGRATION.md names every function include/srb200.h declares (with the reference call site it replaces)."""
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    missing = [f for f in declared_functions() if f not in doc and f.replace("sr_linear_", "sr_linear_") not in doc]
    assert not missing, missing


def test_head_plan_for_every_session_shape(lib):
    """sr_head_plan (host only): every session of BASELINE config 2 runs on the paper-size persistent kernel; without a CTA
    budget it takes the fastest shape (4 rows / 8 columns per CTA, 82-92 CTAs: two such cooperative launches do NOT fit on 148
    SMs), with the budget of three runs per GPU (148 // 3 = 49) every session's launch is <= 49 CTAs so that three are
    resident together; config 5 goes to the tensor-core head, a mid-size problem to the SIMT tiles."""
    from srb200 import _lib

    def plan(ns, nm, ncls, dim, budget, n_new=5, q=60, n_base=60):
        a = _lib.HeadArgs()
        a.n_support, a.n_memory, a.n_classes, a.dim = ns, nm, ncls, dim
        a.n_base, a.n_new, a.n_prev_novel = n_base, n_new, max(ncls - n_base - n_new, 0)
        a.pull_mode, a.q_rows, a.optimizer, a.cta_budget = _lib.SR_PULL_PROJECT, q, _lib.SR_OPT_SGD, budget
        out = (ctypes.c_int32 * 4)()
        assert lib.sr_head_plan(ctypes.byref(a), out) == 0
        return list(out)

    for s in range(1, 9):
        ns, nm, ncls = 185, 25 * (s - 1), 60 + 5 * s
        kind, rows, cols, ctas = plan(ns, nm, ncls, 640, 0)
        assert (kind, rows, cols) == (1, 4, 8) and ctas == max((ns + nm + 3) // 4, 80) + 2 and 2 * ctas > 148
        for budget in (74, 49):
            kind, rows, cols, ctas = plan(ns, nm, ncls, 640, budget)
            assert kind == 1 and ctas <= budget and (148 // budget) * ctas <= 148, (s, budget, ctas)
            assert ctas == max((ns + nm + rows - 1) // rows, 640 // cols) + 2
    assert plan(10000, 0, 1100, 512, 0, n_new=100, q=512, n_base=1000)[0] == 3      # BASELINE config 5
    assert plan(9000, 1000, 300, 512, 0, n_new=50, q=256)[0] == 0                   # neither small nor large
    assert lib.sr_head_plan(None, None) == -1
