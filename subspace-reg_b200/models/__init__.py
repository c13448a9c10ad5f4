"""Drop-in for the reference's ``models`` package: ``model_pool`` (names accepted by ``--model``) and ``model_dict``
(name -> constructor), both derived from the constructors this package actually implements."""
from . import resnet_language as _rl

model_dict = {fn.__name__: fn for fn in (_rl.resnet12, _rl.resnet18)}
model_pool = sorted(model_dict)
resnet12, resnet18 = _rl.resnet12, _rl.resnet18

# reference checkpoints hold argparse.Namespace / numpy scalars: make the callers' plain torch.load(path) accept them
from srb200.checkpoint import allow_reference_checkpoints as _allow
_allow()
