"""Native fit of the label-embedding -> classifier-weight map (reference learn_mapping.py:41-67).

LinearMap(e, 640) is fitted to the base classifier rows by full-batch gradient descent on nn.MSELoss:
1000 steps, lr 1.0, weight_decay 5e-4, no momentum.  The whole fit is ONE launch (sr_fit_linear_map: the output dimensions
are independent, every CTA fits its rows of the map out of shared memory); when the inputs do not fit in shared memory,
or with fused=False, every step is five srb200 launches (sr_linear_fwd, sr_mse_grad, sr_linear_bwd, sr_sgd_update x2).
The result is the state dict the reference stores under ckpt['mapping_linear_label2image'] and
LangPuller.create_pulling_mapping consumes.
"""
import ctypes as C

import torch

from . import _lib as L
from . import ops


def fit_linear_map(label_embeds, targets, epochs=1000, lr=1.0, weight_decay=5e-4, seed=None, init=None, fused=True):
    """label_embeds [n, e], targets [n, d] (CUDA fp32) -> ({'map.weight': [d, e], 'map.bias': [d]}, loss trace list)."""
    n, e = label_embeds.shape
    d = targets.shape[1]
    dev = label_embeds.device
    if init is None:
        if seed is not None:
            torch.manual_seed(seed)
        lin = torch.nn.Linear(e, d)                     # default init, drawn on the CPU generator like the reference
        W, b = lin.weight.detach().to(dev).contiguous(), lin.bias.detach().to(dev).contiguous()
    else:
        W, b = init['map.weight'].to(dev).clone().contiguous(), init['map.bias'].to(dev).clone().contiguous()
    X = label_embeds.contiguous()
    T = targets.contiguous()
    lib = L.load()
    losses = torch.zeros(epochs, dtype=torch.float32, device=dev)
    st = ops._stream()
    ws_bytes = int(lib.sr_fit_linear_map_workspace_bytes(n, e, d, epochs)) if fused else 0
    if ws_bytes > 0:
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        L.check(lib.sr_fit_linear_map(ops._ptr(X, torch.float32, "label_embeds"), ops._ptr(T, torch.float32, "targets"),
                                      ops._ptr(W), ops._ptr(b), n, e, d, epochs, lr, weight_decay, ops._ptr(losses),
                                      ops._ptr(ws), ws_bytes, st), "sr_fit_linear_map")
        ops.LAUNCHES.add(2)
        return {'map.weight': W, 'map.bias': b}, losses.cpu().tolist()
    dy = torch.empty((n, d), dtype=torch.float32, device=dev)
    for ep in range(epochs):
        y = ops.linear_fwd(X, W, b)
        L.check(lib.sr_mse_grad(ops._ptr(y), ops._ptr(T), n * d, ops._ptr(dy), C.c_void_p(losses.data_ptr() + 4 * ep), st),
                "sr_mse_grad")
        dW, db = ops.linear_bwd(dy, X, True)
        L.check(lib.sr_sgd_update(ops._ptr(W), ops._ptr(dW), W.numel(), lr, weight_decay, st), "sr_sgd_update")
        L.check(lib.sr_sgd_update(ops._ptr(b), ops._ptr(db), b.numel(), lr, weight_decay, st), "sr_sgd_update")
        ops.LAUNCHES.add(3)
    return {'map.weight': W, 'map.bias': b}, losses.cpu().tolist()
