"""miniImageNet image store + episode sampler of the incremental evaluation, without PIL / torchvision on the hot path.

Drop-in for the two classes `eval_incremental.py:53-76` builds (reference `dataset/mini_imagenet.py`: `ImageNet` :13-178,
`MetaImageNet` :182-430): same constructor arguments, same attributes (`imgs`, `labels`, `cat2label`, `label2human`,
`classes`, ...), same items.  What is reproduced exactly is everything that decides WHICH images an episode contains -
the NumPy generator is seeded and consumed call for call like the reference does (class shuffle and base / val split of
the `--continual` store :60-117, class order :270-276, per-item seeding, disjoint class groups, support / query draws,
support tiling :279-350) - and the deterministic tail of its transforms (ToTensor + Normalize).  Pinned on
`tests/golden/episodes.pt`, written by `oracle/make_episode_golden.py` from the unmodified reference.

Two ways to get the pixels:
  * default: normalised fp32 NCHW tensors, like the reference's loaders (computed batch-wise with torch ops that round
    exactly like ToTensor + Normalize);
  * `raw=True` (or SRB_RAW_U8=1 in the environment, for callers that cannot pass the flag): the uint8 NHWC images
    themselves - feed them to `ResNet.features` / `BackboneEngine`, whose first kernel (`sr_pack_input_u8`) normalises
    and packs them on the GPU; the support branch's random crop / flip is applied to the uint8 pixels with torchvision's
    own draws.
Callables passed as `transform` / `train_transform` / `test_transform` are applied per image exactly like the reference
does (that is where its PIL augmentations would go; they are not part of this package).

Not on the incremental-session path and therefore not here: contrastive sampling (`is_sample`) and the XtarNet episode
files (`use_episodes`).
"""
import os
import pickle

import numpy as np
import torch

from . import transform_cfg
from .transform_cfg import mean as _MEAN, std as _STD


def _normalise(batch_u8_nhwc):
    """uint8 [n,H,W,3] -> fp32 [n,3,H,W]: x / 255, then (x - mean) / std, rounding like ToTensor + Normalize."""
    t = torch.from_numpy(np.ascontiguousarray(batch_u8_nhwc)).permute(0, 3, 1, 2).to(torch.float32).div(255)
    m = torch.as_tensor(_MEAN, dtype=torch.float32).view(1, -1, 1, 1)
    s = torch.as_tensor(_STD, dtype=torch.float32).view(1, -1, 1, 1)
    return t.sub_(m).div_(s)


def _apply(transform, batch_u8_nhwc, raw):
    """transform: None / kind 'plain' = ToTensor + Normalize; a callable = applied per image like the reference does.
    raw: return uint8 NHWC (normalisation happens in sr_pack_input_u8); a 'crop_flip' transform is then applied HERE on the
    uint8 pixels with the draws torchvision would make (same generator consumption, same pixels), anything else that is
    not plain (ColorJitter, user callables) cannot be expressed on uint8 and raises instead of being dropped."""
    kind = getattr(transform, 'srb_kind', None) if transform is not None else 'plain'
    if raw:
        if kind == 'plain':
            return torch.from_numpy(np.ascontiguousarray(batch_u8_nhwc))
        if kind == 'crop_flip':
            ij, flip = transform_cfg.draw_crop_flip(len(batch_u8_nhwc), size=batch_u8_nhwc.shape[1])
            return torch.from_numpy(transform_cfg.apply_crop_flip_u8(batch_u8_nhwc, ij.numpy(), flip.numpy()))
        raise NotImplementedError("srb200: raw=True supports the plain and the crop+flip transforms (transforms_test_options"
                                  "['A']); %r would be silently dropped" % (kind or transform,))
    if kind == 'plain':
        return _normalise(batch_u8_nhwc)
    return torch.stack([transform(img) for img in batch_u8_nhwc])


class ImageNet(object):
    """Flat view of one split: `self[i] -> (image, label - min(labels), i)`."""

    def __init__(self, args, split='train', phase=None, is_sample=False, k=4096, transform=None, raw=None):
        if is_sample:
            raise NotImplementedError("contrastive sampling is outside the incremental-session path")
        if raw is None:      # callers that cannot pass the flag (the unmodified eval_incremental.py) select it by environment
            raw = os.environ.get("SRB_RAW_U8", "0") == "1"
        self.split, self.phase, self.raw = split, phase, raw
        self.data_aug = getattr(args, 'data_aug', False)
        self.mean, self.std = list(_MEAN), list(_STD)
        self.transform = transform
        np.random.seed(args.set_seed)                          # the store split below draws from this stream

        if args.continual:
            fname = "all.pickle"
        elif split == "train":
            fname = 'miniImageNet_category_split_train_phase_{}.pickle'.format(phase)
        else:
            fname = 'miniImageNet_category_split_{}.pickle'.format(split)
        with open(os.path.join(args.data_root, fname), 'rb') as f:
            blob = pickle.load(f, encoding='latin1')
        imgs, labels, cat2label = blob['data'], list(blob['labels']), dict(blob['catname2label'])

        if args.continual:
            # 100 classes -> 60 base (relabelled 0..59 in sorted order) + 40 novel; base images -> 500/50/50 per class
            order = np.arange(100)
            np.random.shuffle(order)
            base_classes = np.sort(order[:60])
            novel_classes = set(int(c) for c in order[60:])
            if split == "train":
                relabel = {int(c): i for i, c in enumerate(base_classes)}
                members = [i for i, lab in enumerate(labels) if int(lab) in relabel]
                np.random.shuffle(members)
                n_base = len(base_classes)
                cuts = {"train": (0, 500 * n_base), "val": (500 * n_base, 550 * n_base), "test": (550 * n_base, None)}
                if phase not in cuts:
                    raise ValueError("Phase {} is unrecognized for split train.".format(phase))
                lo, hi = cuts[phase]
                keep = np.asarray(members[lo:hi], dtype=np.int64)
                labels = [relabel[int(labels[i])] for i in keep]
                imgs = imgs[keep, :]
                cat2label = {name: relabel[int(v)] for name, v in cat2label.items() if int(v) in relabel}
            elif split == "val":
                keep = np.asarray([i for i, lab in enumerate(labels) if int(lab) in novel_classes], dtype=np.int64)
                labels = [labels[i] for i in keep]
                imgs = imgs[keep, :]
                cat2label = {name: v for name, v in cat2label.items() if int(v) in novel_classes}
            else:
                raise ValueError("No such split as {}.".format(split))

        self.imgs, self.labels, self.cat2label = imgs, labels, cat2label
        self.global_labels = self.labels
        self.label2human = [""] * 100
        with open(os.path.join(args.data_root, 'class_labels.txt'), 'r') as f:
            for line in f:
                cat, human = line.strip().lower().split(' ')
                if cat in cat2label:
                    self.label2human[cat2label[cat]] = " ".join(human.split('_'))
        self.k, self.is_sample = k, False
        self._label_min = min(self.labels) if len(self.labels) else 0

    def __getitem__(self, item):
        img = np.asarray(self.imgs[item]).astype('uint8')
        return _apply(self.transform, img[None], self.raw)[0], self.labels[item] - self._label_min, item

    def __len__(self):
        return len(self.labels)


class MetaImageNet(ImageNet):
    """Episodes: `self[i] -> (support_xs, support_ys, query_xs, query_ys)`."""

    def __init__(self, args, split, phase=None, train_transform=None, test_transform=None, fix_seed=True,
                 use_episodes=False, disjoint_classes=False, raw=None):
        if use_episodes:
            raise NotImplementedError("XtarNet episode files are outside the incremental-session path")
        super(MetaImageNet, self).__init__(args, split, phase, raw=raw)
        if split != "train":
            assert phase is None
        self.fix_seed, self.use_episodes, self.disjoint_classes = fix_seed, False, disjoint_classes
        self.n_ways, self.n_shots, self.n_queries = args.n_ways, args.n_shots, args.n_queries
        self.n_test_runs, self.eval_mode = args.n_test_runs, args.eval_mode
        self.n_aug_support_samples = args.n_aug_support_samples
        self.n_base_aug_support_samples = args.n_base_aug_support_samples
        self.n_base_support_samples = args.n_base_support_samples
        # like the reference (:242-262): no train_transform = RandomCrop + ColorJitter + RandomHorizontalFlip, no
        # test_transform = ToTensor + Normalize
        self.train_transform = transform_cfg.transforms_options['A'][0] if train_transform is None else train_transform
        self.test_transform = test_transform
        # images of a class, in store order; classes in order of first appearance, then shuffled with the run's seed
        lab = np.asarray(self.labels)
        self._members = {}
        self.classes = []
        for i, c in enumerate(lab):
            if c not in self._members:
                self._members[c] = []
                self.classes.append(c)
            self._members[c].append(i)
        self._members = {c: np.asarray(v, dtype=np.int64) for c, v in self._members.items()}
        if self.fix_seed:
            np.random.seed(args.set_seed)
            np.random.shuffle(self.classes)

    @property
    def data(self):
        """class -> list of its images (the reference's attribute; built on demand here)."""
        return {c: [self.imgs[i] for i in idx] for c, idx in self._members.items()}

    def _class_images(self, cls):
        return np.asarray(self.imgs[self._members[cls]]).astype('uint8')

    def __getitem__(self, item):
        base_exemplars = self.split == "train" and self.phase == "train" and self.n_base_support_samples > 0
        if self.fix_seed:
            np.random.seed(item)
        if base_exemplars:
            # n_base_support_samples images of EVERY base class: what the memory of old classes is filled with
            picked = np.random.choice(self.classes, len(self.classes), False)
            xs, ys = [], []
            for cls in np.sort(picked):
                imgs = self._class_images(cls)
                take = np.random.choice(range(imgs.shape[0]), self.n_base_support_samples, False)
                xs.append(imgs[take])
                ys.append([cls] * self.n_base_support_samples)
            xs, ys = np.array(xs), np.array(ys)
            xs = xs.reshape((-1,) + xs.shape[2:])
            if self.n_base_aug_support_samples > 1:
                xs = np.tile(xs, (self.n_base_aug_support_samples, 1, 1, 1))
                ys = np.tile(ys.reshape((-1,)), (self.n_base_aug_support_samples))
            sx = _apply(self.train_transform, xs, self.raw)
            sx = sx if self.raw else sx.float()
            return sx, ys, sx, ys                                   # (the query is a dummy, like the reference's)

        if self.disjoint_classes:                                    # sessions walk the shuffled class list n_ways at a time
            picked, self.classes = self.classes[:self.n_ways], self.classes[self.n_ways:]
        else:
            picked = np.random.choice(self.classes, self.n_ways, False)
        sup, sup_y, qry, qry_y = [], [], [], []
        keep_class_ids = self.eval_mode in ["few-shot-incremental-fine-tune"]
        for pos, cls in enumerate(np.sort(picked)):
            imgs = self._class_images(cls)
            s_ids = np.random.choice(range(imgs.shape[0]), self.n_shots, False)
            rest = np.setxor1d(np.arange(imgs.shape[0]), s_ids)
            q_ids = np.random.choice(rest, self.n_queries, False)
            lbl = cls if keep_class_ids else pos
            sup.append(imgs[s_ids])
            sup_y.append([lbl] * self.n_shots)
            qry.append(imgs[q_ids])
            qry_y.append([lbl] * q_ids.shape[0])
        sup, sup_y, qry, qry_y = np.array(sup), np.array(sup_y), np.array(qry), np.array(qry_y)
        qry = qry.reshape((-1,) + qry.shape[2:])
        qry_y = qry_y.reshape((-1,))
        sup = sup.reshape((-1,) + sup.shape[2:])
        if self.n_aug_support_samples > 1:                           # the 5x support tiling (augmented copies upstream)
            sup = np.tile(sup, (self.n_aug_support_samples, 1, 1, 1))
            sup_y = np.tile(sup_y.reshape((-1,)), (self.n_aug_support_samples))
        sx, qx = _apply(self.train_transform, sup, self.raw), _apply(self.test_transform, qry, self.raw)
        if not self.raw:
            sx, qx = sx.float(), qx.float()
        return sx, sup_y, qx, qry_y

    def __len__(self):
        if self.split == "train" and self.phase == "train" and self.disjoint_classes:
            return 8
        return self.n_test_runs
