"""The reference's OWN entry scripts, unmodified, on top of the B200 shadow packages:

  * eval_incremental.main()  - argparse, the three DataLoaders over dataset.mini_imagenet, torch.load of a reference-format
                               checkpoint file, create_model / load_state_dict, few_shot_finetune_incremental_test;
  * learn_mapping.main()     - fits LinearMap on the checkpoint's classifier rows and stores it back into the checkpoint;
  * eval/language_eval.py    - the reference's own session loop (per-op mode: every library call of the loop lands in one
                               srb200 kernel through the autograd wrappers) against the fused driver on the same inputs.

The scripts are staged under baseline/_ref/ by oracle/stage_reference.py in the build container (git-ignored, they travel
to the GPU box with the snapshot); the tests skip when they are absent.  Each run is a fresh interpreter whose sys.path
decides which packages shadow which: [shadow, _ref] = the reference callers over the B200 modules and the fused driver;
[_ref, shadow] = additionally the reference's own eval/ package (only eval/ is staged, so models/ and dataset/ still
resolve to the B200 packages).
"""
import os
import re
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "subspace-reg_b200")
REF = os.path.join(ROOT, "baseline", "_ref")

DRIVER = r'''
import sys, runpy
sys.path[:0] = PATHS
sys.argv = ARGV
import torch
if SEED is not None:
    torch.manual_seed(SEED)
if WHAT == "eval":
    import eval_incremental
    eval_incremental.main()
else:
    import learn_mapping
    learn_mapping.main(*MAPARGS)
import srb200._lib as L, eval.language_eval as le
print("LOADED_LIB", L.LIB_PATH, "EVAL_MODULE", le.__file__)
'''


def _run(paths, argv, cwd, what="eval", mapargs=(), env=None, seed=None, timeout=900):
    code = (DRIVER.replace("PATHS", repr(list(paths))).replace("ARGV", repr(list(argv))).replace("WHAT", repr(what))
            .replace("MAPARGS", repr(tuple(mapargs))).replace("SEED", repr(seed)))
    e = dict(os.environ)
    e.pop("PYTHONPATH", None)
    e.update(env or {})
    r = subprocess.run([sys.executable, "-c", code], cwd=cwd, env=e, capture_output=True, text=True, timeout=timeout)
    assert r.returncode == 0, "driver failed:\n%s\n%s" % (r.stdout[-3000:], r.stderr[-3000:])
    return r.stdout


def _lists(out):
    got = {}
    for key in ("Overall continual accuracies", "Novel only incremental", "Base only incremental"):
        m = re.findall(re.escape(key) + r":\s*\[([^\]]*)\]", out)
        assert m, "missing %r in the output" % key
        got[key] = [float(v) for v in m[-1].replace("np.float64(", "").replace(")", "").split(",")]
    return got


@pytest.fixture(scope="module")
def staged(tmp_path_factory, word_embed_dir):
    if not os.path.isfile(os.path.join(REF, "eval_incremental.py")):
        pytest.skip("baseline/_ref is not staged (run `python -m oracle.stage_reference` in the build container)")
    sys.path.insert(0, PKG)
    from srb200 import synthetic
    from srb200.checkpoint import save_reference_checkpoint
    from models.util import create_model
    work = tmp_path_factory.mktemp("integration")
    data_root = work / "data"
    # a miniImageNet-format store (100 classes x 560 images of 84x84: the --continual split takes 500 / 50 / 50+ per base class)
    synthetic.write_image_store(str(data_root / "miniImageNet"), n_classes=100, per_class=560, side=84, light=True,
                                class_names=synthetic.LABELS)
    # word_embeds/ next to the working directory (learn_mapping.py reads the relative path)
    os.symlink(word_embed_dir, str(work / "word_embeds"))
    # a reference-format checkpoint file of a random-init ResNet-18 (train_supervised.py:194-202)
    import argparse
    from dataset.mini_imagenet import ImageNet
    seed = 3
    a = argparse.Namespace(data_root=str(data_root / "miniImageNet"), set_seed=seed, continual=True, data_aug=True)
    base = ImageNet(args=a, split='train', phase='test', raw=True)
    opt = synthetic.default_opt(seed)
    net = synthetic.init_model(create_model, opt, seed)
    training_classes = {i: i for i in range(60)}
    ckpt_path = str(work / "resnet18_last.pth")
    save_reference_checkpoint(ckpt_path, net, training_classes, base.label2human, opt=argparse.Namespace(**vars(opt)))
    return dict(work=str(work), data_root=str(data_root), ckpt=ckpt_path, seed=seed)


def _argv(s, ckpt, extra=()):
    return ["eval_incremental.py", "--model", "resnet18", "--model_path", ckpt, "--data_root", s['data_root'],
            "--no_dropblock", "--n_shots", "5", "--n_queries", "25", "--classifier", "linear",
            "--eval_mode", "few-shot-incremental-fine-tune", "--min_novel_epochs", "1", "--max_novel_epochs", "3",
            "--learning_rate", "0.002", "--momentum", "0.9", "--weight_decay", "5e-4", "--freeze_backbone_at", "1",
            "--test_base_batch_size", "400", "--continual", "--num_workers", "0", "--lmbd_reg_transform_w", "0.2",
            "--lmbd_reg_novel", "0.1", "--target_train_loss", "0.0", "--stable_epochs", "10",
            "--convergence_epsilon", "1e-4", "--n_base_support_samples", "1", "--memory_replay", "1",
            "--set_seed", str(s['seed']), "--word_embed_path", os.path.join(s['work'], "word_embeds")] + list(extra)


SUBSPACE = ["--label_pull", "1.0", "--attraction_override", "distance2subspace"]


def test_unmodified_eval_incremental_over_the_shadow_packages(staged):
    """eval_incremental.main() end to end from a checkpoint FILE and an image STORE: host fp32 front end (torchvision crop /
    flip on the support copies) and the uint8 front end (SRB_RAW_U8=1: crop / flip on uint8, ToTensor + Normalize inside
    sr_pack_input_u8) must print the same accuracy lists - the two front ends produce bit-identical pixels."""
    s = staged
    out = _run([PKG, REF], _argv(s, s['ckpt'], SUBSPACE), s['work'])
    assert "LOADED_LIB" in out and "libsrb200.so" in out
    assert os.path.join("subspace-reg_b200", "eval", "language_eval.py") in out      # the fused driver ran
    a = _lists(out)
    assert len(a["Overall continual accuracies"]) == 9 and len(a["Novel only incremental"]) == 8
    assert "val_acc_novel" in out and "val_acc_base" in out
    out_u8 = _run([PKG, REF], _argv(s, s['ckpt'], SUBSPACE), s['work'], env={"SRB_RAW_U8": "1"})
    assert _lists(out_u8) == a
    print("eval_incremental.main(): weighted accuracies", a["Overall continual accuracies"])


def test_reference_language_eval_per_op_matches_the_fused_driver(staged):
    """The reference's own eval/language_eval.py (literal schedule, torch.optim.SGD, loss.backward()) over the shadow models:
    every op of its loop is one srb200 kernel behind an autograd wrapper.  Same store, checkpoint and seed as the fused
    driver -> same accuracy lists (3 epochs per session; a prediction may flip where two logits tie to fp32 rounding)."""
    s = staged
    fused = _lists(_run([PKG, REF], _argv(s, s['ckpt'], SUBSPACE), s['work']))
    out = _run([REF, PKG], _argv(s, s['ckpt'], SUBSPACE), s['work'])
    assert os.path.join("baseline", "_ref", "eval", "language_eval.py") in out       # the reference's loop ran
    per_op = _lists(out)
    for key in fused:
        assert len(fused[key]) == len(per_op[key])
        worst = max(abs(x - y) for x, y in zip(fused[key], per_op[key]))
        print("%s: fused %s\n    per-op %s (max diff %.2f)" % (key, fused[key], per_op[key], worst))
        assert worst <= 0.81, key         # at most one query image of 125 (0.8 points) / a few base images of 400


def test_unmodified_learn_mapping_then_mapping_mode(staged):
    """learn_mapping.main() (1000 SGD steps on LinearMap through the per-op kernels) writes the mapping into the checkpoint;
    the native fit (srb200.mapping.fit_linear_map) gives the same map; eval_incremental then runs in the mapping mode of
    scripts/continual/slurm_linear_mapping.sh from that file."""
    s = staged
    sys.path.insert(0, PKG)
    from models.util import get_embeds
    from srb200 import mapping
    from srb200.checkpoint import load_reference_checkpoint
    save_path = os.path.join(s['work'], "resnet18_last_with_mapping.pth")
    out = _run([PKG, REF], ["learn_mapping.py"], s['work'], what="map", mapargs=(s['work'], s['ckpt'], save_path), seed=11)
    assert "Epoch [1000/1000]" in out
    ckpt = load_reference_checkpoint(save_path)
    got = ckpt['mapping_linear_label2image']
    assert tuple(got['map.weight'].shape) == (640, 300) and tuple(got['map.bias'].shape) == (640,)
    labels = [n for n in ckpt['label2human'] if n != '']
    emb = get_embeds(os.path.join(s['work'], "word_embeds", "miniImageNet_dim500.pickle"), labels).float()[:, :300].cuda().contiguous()
    want, losses = mapping.fit_linear_map(emb, ckpt['model']['classifier.weight'].cuda().contiguous(), seed=11)
    for k in got:
        rel = ((got[k].cuda() - want[k]).norm() / want[k].norm()).item()
        print("learn_mapping.main() vs native fit, %s: rel %.2e" % (k, rel))
        assert rel < 1e-4, (k, rel)
    out = _run([PKG, REF], _argv(s, save_path, ["--label_pull", "0.1", "--glove", "--attraction_override",
                                               "mapping_linear_label2image"]), s['work'])
    assert len(_lists(out)["Overall continual accuracies"]) == 9
