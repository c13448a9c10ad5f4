"""Synthetic miniImageNet-shaped world: class split, images, loaders, checkpoint dict, opt.

The objects built here are consumed unchanged by the reference's ``few_shot_finetune_incremental_test``
(eval/language_eval.py:71), by the oracle restatement and by the B200 path, so all three see identical inputs.
Everything is generated on the CPU from seeded generators (no dataset or checkpoint files are needed).
"""
import os
import pickle
from types import SimpleNamespace

import numpy as np
import torch
import torch.nn.functional as F

# The 100 miniImageNet class names (keys of the reference's description_embeds pickles, in file order).
LABELS = [
    'house finch', 'robin', 'triceratops', 'green mamba', 'harvestman', 'toucan', 'jellyfish', 'dugong', 'walker hound',
    'saluki', 'gordon setter', 'komondor', 'boxer', 'tibetan mastiff', 'french bulldog', 'newfoundland',
    'miniature poodle', 'arctic fox', 'ladybug', 'three-toed sloth', 'rock beauty', 'aircraft carrier', 'ashcan',
    'barrel', 'beer bottle', 'carousel', 'chime', 'clog', 'cocktail shaker', 'dishrag', 'dome', 'file', 'fire screen',
    'frying pan', 'hair slide', 'holster', 'lipstick', 'oboe', 'organ', 'parallel bars', 'pencil box', 'photocopier',
    'prayer rug', 'reel', 'slot', 'snorkel', 'solar dish', 'spider web', 'stage', 'tank', 'tile roof', 'tobacco shop',
    'unicycle', 'upright', 'wok', 'worm fence', 'yawl', 'street sign', 'consomme', 'hotdog', 'orange', 'cliff', 'bolete',
    'ear', 'dalmatian', 'nematode', 'ant', 'black-footed ferret', 'king crab', 'lion', 'vase', 'golden retriever',
    'mixing bowl', 'malamute', 'african hunting dog', 'cuirass', 'bookshop', 'crate', 'hourglass', 'electric guitar',
    'trifle', 'school bus', 'theater curtain', 'scoreboard', 'horizontal bar', 'combination lock', 'catamaran', 'poncho',
    'miniskirt', 'ibizan hound', 'white wolf', 'rhinoceros beetle', 'garbage truck', 'carton', 'ipod', 'meerkat',
    'missile', 'cannon', 'goose', 'coral reef',
]

IMG = 84


def class_split(seed, n_base=60, n_ways=5, n_sessions=8):
    """Seed-dependent base/novel split and disjoint sessions (mirrors dataset/mini_imagenet.py:70-78, 314-323)."""
    rng = np.random.RandomState(seed)
    perm = rng.permutation(100)
    base = np.sort(perm[:n_base])
    novel = perm[n_base:]
    sessions = [np.sort(novel[n_ways * s:n_ways * (s + 1)]) for s in range(n_sessions)]
    return base, sessions


class _Dataset(object):
    def __init__(self, label2human):
        self.label2human = label2human


class ListLoader(object):
    """Minimal DataLoader stand-in: an iterable of pre-built batches with a ``.dataset.label2human`` list."""

    def __init__(self, batches, label2human):
        self.batches = batches
        self.dataset = _Dataset(label2human)

    def __iter__(self):
        return iter(self.batches)

    def __len__(self):
        return len(self.batches)


def _images(gen, means, classes):
    """x = mu_class + eps, eps ~ N(0, 1): fp32 [len(classes), 3, 84, 84]."""
    classes = np.asarray(classes)
    eps = torch.randn((len(classes), 3, IMG, IMG), generator=gen)
    return means[torch.from_numpy(classes).long()] + eps


def default_opt(seed=1, **over):
    """Canonical hyper-parameters of scripts/continual/slurm_subspace_reg.sh:33-54."""
    opt = SimpleNamespace(
        model='resnet18', dataset='miniImageNet', set_seed=seed, n_ways=5, n_shots=5, n_queries=25,
        n_aug_support_samples=5, n_base_support_samples=1, memory_replay=1, neval_episodes=8, continual=True,
        classifier='linear', eval_mode='few-shot-incremental-fine-tune', min_novel_epochs=20, max_novel_epochs=1000,
        learning_rate=0.002, momentum=0.9, weight_decay=5e-4, adam=False, freeze_backbone_at=1,
        lmbd_reg_transform_w=0.2, lmbd_reg_novel=0.1, label_pull=1.0, pulling='regularize',
        attraction_override='distance2subspace', target_train_loss=0.0, stable_epochs=10, convergence_epsilon=1e-4,
        temperature=1.0, word_embed_size=500, word_embed_path='word_embeds', use_synonyms=False, glove=False,
        no_dropblock=True, linear_bias=False, track_weights=False, track_label_inspired_weights=False,
        save_preds_0=False, verbose=False, attention=None, push_away=None, use_episodes=False, test_base_batch_size=2000,
        num_workers=0, split='val', vis=False)
    for k, v in over.items():
        setattr(opt, k, v)
    return opt


def make_world(seed=1, n_sessions=8, n_base_batch=1000, opt=None, **opt_over):
    """Build (opt, ckpt_meta, loaders) for one seed.

    Returns a SimpleNamespace with: opt, base_classes, sessions, label2human_base/novel, base_val_loader,
    base_support_loader, meta_valloader, training_classes.  The model/ckpt['model'] is created separately by
    ``init_model`` so the caller can choose the reference, the oracle or the B200 ``models`` package."""
    if opt is None:
        opt = default_opt(seed, **opt_over)
    base, sessions = class_split(seed, 60, opt.n_ways, 8)
    gen = torch.Generator().manual_seed(seed)
    mu = torch.randn((100, 3, 8, 8), generator=gen) * 0.5
    means = F.interpolate(mu, size=(IMG, IMG), mode='nearest')
    base_map = {int(c): i for i, c in enumerate(base)}

    # fixed base exemplars: one per base class, labels 0..59 (language_eval.py:112-116)
    bs_x = _images(gen, means, base)
    bs_y = torch.arange(len(base), dtype=torch.int64)
    dummy = torch.zeros((1, 1, 3, IMG, IMG))   # drop_a_dim views the (unused) query slot too
    base_support_loader = ListLoader([(bs_x.unsqueeze(0), bs_y.unsqueeze(0), dummy, torch.zeros((1, 1), dtype=torch.int64))],
                                     _l2h(base))
    # fixed base evaluation batch (language_eval.py:121)
    by = torch.randint(0, len(base), (n_base_batch,), generator=gen)
    bx = _images(gen, means, base[by.numpy()])
    base_val_loader = ListLoader([(bx, by)], _l2h(base))

    episodes = []
    n_aug, shots, ways, nq = opt.n_aug_support_samples, opt.n_shots, opt.n_ways, opt.n_queries
    for s in range(n_sessions):
        cls = sessions[s]
        sy = np.tile(np.repeat(cls, shots), n_aug)          # aug*25 + way*5 + shot
        qy = np.repeat(cls, nq)
        sx = _images(gen, means, sy)
        qx = _images(gen, means, qy)
        episodes.append((sx.unsqueeze(0), torch.from_numpy(sy).long().unsqueeze(0), qx.unsqueeze(0),
                         torch.from_numpy(qy).long().unsqueeze(0)))
    novel_all = np.concatenate(sessions)
    meta_valloader = ListLoader(episodes, _l2h(novel_all))
    return SimpleNamespace(opt=opt, seed=seed, base_classes=base, sessions=sessions[:n_sessions],
                           base_val_loader=base_val_loader, base_support_loader=base_support_loader,
                           meta_valloader=meta_valloader, training_classes=base_map,
                           label2human=[LABELS[c] if c in base_map else '' for c in range(100)])


def _l2h(classes):
    s = set(int(c) for c in classes)
    return [LABELS[c] if c in s else '' for c in range(100)]


def init_model(create_model, opt, seed):
    """torch.manual_seed(seed); create_model('resnet18', 60, opt) - identical RNG consumption for every models package."""
    torch.manual_seed(seed)
    return create_model(opt.model, 60, opt)


def make_ckpt(model, world):
    return {'model': {k: v.clone() for k, v in model.state_dict().items()},
            'training_classes': dict(world.training_classes), 'label2human': list(world.label2human)}


def write_word_embeds(npz_path, out_dir, dataset='miniImageNet', dim=500):
    """Re-create ``{dataset}_dim{dim}.pickle`` (dict word -> float32[dim], the format models/util.get_embeds reads)
    from the committed fixture tests/golden/word_embeds_dim500.npz."""
    z = np.load(npz_path, allow_pickle=False)
    words = [str(w) for w in z['words']]
    vecs = z['vectors'].astype(np.float32)
    os.makedirs(out_dir, exist_ok=True)
    path = os.path.join(out_dir, "{0}_dim{1}.pickle".format(dataset, dim))
    with open(path, 'wb') as f:
        pickle.dump({w: vecs[i] for i, w in enumerate(words)}, f)
    return path


def write_image_store(out_dir, n_classes=100, per_class=600, side=4, seed=0, light=False, class_names=None):
    """A miniImageNet-format store for the data front-end tests: ``all.pickle`` ({'data': uint8 [N,side,side,3], 'labels',
    'catname2label'}, the layout dataset/mini_imagenet.py reads with --continual) + ``class_labels.txt``.  Images are
    tiny; pixel (0,0) encodes (class, index // 256, index % 256) so a sampled image can be identified, the rest is noise.
    Class names are LABELS-like single tokens so that get_vocabs works (or `class_names`, e.g. LABELS, for runs that look
    label embeddings up).  light=True: full-size stores (side 84) are filled by tiling a pool of 509 random images with a
    per-class brightness offset - seconds instead of a minute for the 1.2 GB a 100 x 560 store takes."""
    rng = np.random.RandomState(seed)
    n = n_classes * per_class
    if light:
        pool = rng.randint(0, 200, size=(509, side, side, 3)).astype(np.uint8)
        data = pool[np.arange(n) % 509]
    else:
        data = rng.randint(0, 256, size=(n, side, side, 3)).astype(np.uint8)
    labels = []
    order = rng.permutation(n)                     # classes interleaved like a real store
    cls_of = np.repeat(np.arange(n_classes), per_class)[order]
    counters = np.zeros(n_classes, dtype=np.int64)
    if light:
        data += (cls_of % 56).astype(np.uint8)[:, None, None, None]      # a little class structure (values stay < 256)
    for i in range(n):
        c = int(cls_of[i])
        k = int(counters[c])
        counters[c] += 1
        data[i, 0, 0] = (c, k // 256, k % 256)
        labels.append(c)
    os.makedirs(out_dir, exist_ok=True)
    with open(os.path.join(out_dir, "all.pickle"), 'wb') as f:
        pickle.dump({'data': data, 'labels': labels, 'catname2label': {"n%08d" % c: c for c in range(n_classes)}}, f,
                    protocol=4)
    with open(os.path.join(out_dir, "class_labels.txt"), 'w') as f:
        for c in range(n_classes):
            name = "class_%d" % c if class_names is None else "_".join(class_names[c].split(' '))
            f.write("n%08d %s\n" % (c, name))
    return out_dir


def image_ids(imgs_u8_nhwc):
    """(class, index) of images written by write_image_store, from pixel (0,0)."""
    a = np.asarray(imgs_u8_nhwc).astype(np.int64)
    return np.stack([a[:, 0, 0, 0], a[:, 0, 0, 1] * 256 + a[:, 0, 0, 2]], 1)
