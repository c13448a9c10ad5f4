"""Reference-format checkpoints: what `train_supervised.py:194-202` writes and `eval_incremental.py:86-110` /
`learn_mapping.py:27-39` read.

    {'opt': argparse.Namespace,                      # training options (ignored by the evaluation)
     'model': state_dict,                            # reference key set: layer{1..4}.{0,1}.{conv,bn}{1,2,3}.*, ...downsample.{0,1}.*,
                                                     #   classifier.weight[, classifier.bias]; NCHW fp32 conv weights
     'training_classes': {np.int64: np.int64},       # original class id -> 0..59 (dataset/mini_imagenet.py:76), NUMPY keys
     'label2human': [100 x str],                     # '' for classes that are not base classes
     'mapping_linear_label2image': {'map.weight': [640, e], 'map.bias': [640]}}   # optional, added by learn_mapping.py:37-39,67

Two things stand between such a file and the unmodified callers on a current PyTorch:
  * `torch.load(path)` defaults to weights_only=True since 2.6 and rejects argparse.Namespace and the numpy scalars of
    `training_classes`.  `allow_reference_checkpoints()` allow-lists exactly those types; the shadow `models` package calls it
    on import, so `eval_incremental.py:86` and `learn_mapping.py:28` work as written.
  * the conv weights are repacked (OIHW fp32 -> [cout][tap][cin_pad] bf16 operand planes, eval-mode BN folded in) lazily by
    BackboneEngine the first time the model runs after `load_state_dict`; nothing is needed here.
"""
import argparse

import numpy as np
import torch

_ALLOWED = [False]


def allow_reference_checkpoints():
    """Allow-list the non-tensor types of a reference checkpoint for torch.load(weights_only=True).  Idempotent."""
    if _ALLOWED[0]:
        return
    safe = [argparse.Namespace, np.dtype, np.int64, np.int32, np.float64, np.float32, np.bool_, np.ndarray]
    try:
        from numpy._core import multiarray as _ma
    except ImportError:                                   # numpy < 2
        from numpy.core import multiarray as _ma
    safe += [_ma.scalar, _ma._reconstruct]
    for name in ("Int64DType", "Int32DType", "Float64DType", "Float32DType", "BoolDType"):
        t = getattr(getattr(np, "dtypes", None), name, None)
        if t is not None:
            safe.append(t)
    torch.serialization.add_safe_globals(safe)
    _ALLOWED[0] = True


def save_reference_checkpoint(path, model, training_classes, label2human, opt=None, mapping=None):
    """Write `model` (a models.resnet_language.ResNet, any device) in the reference's checkpoint layout."""
    state = {'opt': opt if opt is not None else argparse.Namespace(),
             'model': {k: v.detach().cpu().clone() for k, v in model.state_dict().items()},
             'training_classes': {np.int64(k): np.int64(v) for k, v in dict(training_classes).items()},
             'label2human': list(label2human)}
    if mapping is not None:
        state['mapping_linear_label2image'] = {k: v.detach().cpu().clone() for k, v in mapping.items()}
    torch.save(state, path)
    return state


def load_reference_checkpoint(path, map_location=None):
    allow_reference_checkpoints()
    ckpt = torch.load(path, map_location=map_location)
    missing = [k for k in ('model', 'training_classes', 'label2human') if k not in ckpt]
    if missing:
        raise KeyError("srb200: %s is not a reference continual checkpoint (missing %s)" % (path, missing))
    return ckpt


def load_backbone(model, ckpt):
    """model.load_state_dict(ckpt['model']) with the reference's bias convention (eval_incremental.py:99-110): a checkpoint
    without classifier.bias needs a model built with opt.linear_bias = False."""
    has_bias = 'classifier.bias' in ckpt['model']
    if has_bias != (model.classifier.bias is not None):
        raise ValueError("srb200: checkpoint %s classifier.bias but the model was built with linear_bias=%s" %
                         ("has" if has_bias else "has no", model.classifier.bias is not None))
    model.load_state_dict(ckpt['model'])
    if getattr(model, '_engine', None) is not None:
        model._engine.invalidate()
    return model
