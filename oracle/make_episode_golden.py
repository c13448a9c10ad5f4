"""Golden episodes of the reference's data front-end (test infrastructure; run in the build container only).

Imports the UNMODIFIED reference dataset classes (/root/reference/dataset/mini_imagenet.py: ImageNet :13-178,
MetaImageNet :182-430) on the synthetic image store of srb200.synthetic.write_image_store and records which images /
labels every episode of the incremental evaluation uses (eval_incremental.py:53-76: base test set, base exemplar
episodes, eight disjoint novel sessions), with the deterministic test transform on both branches.  Output:
tests/golden/episodes.pt (a few hundred KB).  tests/test_host_logic.py replays it against dataset/mini_imagenet.py.
"""
import argparse
import os
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "subspace-reg_b200"))
from srb200 import synthetic  # noqa: E402


def reference_classes():
    sys.path.insert(0, "/root/reference")
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_mini_imagenet", "/root/reference/dataset/mini_imagenet.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def make_args(data_root, seed):
    a = argparse.Namespace()
    a.data_root, a.data_aug, a.set_seed, a.continual = data_root, False, seed, True
    a.n_ways, a.n_shots, a.n_queries, a.n_test_runs = 5, 5, 15, 8
    a.eval_mode = "few-shot-incremental-fine-tune"
    a.n_aug_support_samples, a.n_base_aug_support_samples, a.n_base_support_samples = 5, 1, 1
    return a


def ids_of(x_norm, mean, std):
    """(class, index) from normalised NCHW float tensors produced by ToTensor + Normalize."""
    px = x_norm[:, :, 0, 0] * torch.tensor(std) + torch.tensor(mean)
    px = torch.round(px * 255).to(torch.int64).numpy()
    return np.stack([px[:, 0], px[:, 1] * 256 + px[:, 2]], 1)


def main():
    ref = reference_classes()
    from PIL import Image
    import torchvision.transforms as T
    out = {}
    with tempfile.TemporaryDirectory() as d:
        synthetic.write_image_store(d)
        for seed in (5, 11):
            args = make_args(d, seed)
            base = ref.ImageNet(args=args, split='train', phase='test')
            norm = T.Compose([lambda x: Image.fromarray(x), T.ToTensor(), base.normalize])
            base = ref.ImageNet(args=args, split='train', phase='test', transform=norm)
            xs = torch.stack([base[i][0] for i in range(0, len(base), 97)])
            rec = dict(base_len=len(base), base_labels=np.asarray(base.labels), label2human=list(base.label2human),
                       base_probe_ids=ids_of(xs, base.mean, base.std), base_probe_x=xs[:3].clone())
            sup = ref.MetaImageNet(args=args, split='train', phase='train', train_transform=norm, test_transform=norm,
                                   fix_seed=True, use_episodes=False)
            rec['exemplar_len'] = len(sup)
            rec['exemplars'] = []
            for item in (0, 3):
                sx, sy, _, _ = sup[item]
                rec['exemplars'].append(dict(ids=ids_of(sx, base.mean, base.std), ys=np.asarray(sy)))
            val = ref.MetaImageNet(args=args, split='val', train_transform=norm, test_transform=norm, fix_seed=True,
                                   use_episodes=False, disjoint_classes=True)
            rec['val_len'] = len(val)
            rec['val_label2human'] = list(val.label2human)
            rec['sessions'] = []
            for item in range(8):
                sx, sy, qx, qy = val[item]
                e = dict(sup_ids=ids_of(sx, base.mean, base.std), sup_ys=np.asarray(sy), qry_ids=ids_of(qx, base.mean, base.std),
                         qry_ys=np.asarray(qy))
                if item == 0:
                    e['sup_x'] = sx[:10].clone()
                rec['sessions'].append(e)
            out[seed] = rec
    path = os.path.join(ROOT, "tests", "golden", "episodes.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
