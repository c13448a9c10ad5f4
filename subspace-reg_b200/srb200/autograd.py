"""torch.autograd.Function wrappers: one srb200 kernel launch per reference library call, so the reference's own
epoch loop (loss.backward() reaching only classifier.weight) runs unmodified on the B200 modules.

forward/backward arithmetic is in libsrb200.so (sr_linear_fwd/bwd, sr_sqdist, sr_diff_scale, sr_project_rows).
"""
import torch

from . import ops


def _c(t):
    return t if t.is_contiguous() else t.contiguous()


class _Linear(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b):
        x, w = _c(x.detach()), _c(w.detach())
        ctx.save_for_backward(x)
        ctx.has_bias = b is not None
        ctx.need_w = True
        return ops.linear_fwd(x, w, None if b is None else _c(b.detach()))

    @staticmethod
    def backward(ctx, gy):
        (x,) = ctx.saved_tensors
        if ctx.needs_input_grad[0]:
            raise NotImplementedError("srb200: gradients w.r.t. the features (backbone training) are out of scope")
        dw, db = ops.linear_bwd(_c(gy), x, ctx.has_bias)
        return None, dw, db


def linear(x, w, b=None):
    """F.linear(x, w, b) with gradients for w and b only (the backbone is frozen on this path)."""
    return _Linear.apply(x, w, b)


class _ScaledSqDist(torch.autograd.Function):
    """s * ||a - b||_F^2  (LangPuller.loss1)."""

    @staticmethod
    def forward(ctx, a, b, s):
        a, b = _c(a.detach()), _c(b.detach())
        ctx.save_for_backward(a, b)
        ctx.s = s
        return (ops.sqdist(a, b) * s).reshape(())

    @staticmethod
    def backward(ctx, g):
        a, b = ctx.saved_tensors
        g = _c(g.reshape(1))
        ga = ops.diff_scale(a, b, 2.0 * ctx.s, g) if ctx.needs_input_grad[0] else None
        gb = ops.diff_scale(b, a, 2.0 * ctx.s, g) if ctx.needs_input_grad[1] else None
        return ga, gb, None


class _ScaledDist(torch.autograd.Function):
    """s * ||a - b||_F  (ResNet.regloss / reglossnovel; zero sub-gradient at a == b like torch.norm)."""

    @staticmethod
    def forward(ctx, a, b, s):
        a, b = _c(a.detach()), _c(b.detach())
        sq = ops.sqdist(a, b)
        ctx.save_for_backward(a, b, sq)
        ctx.s = s
        return (sq.sqrt() * s).reshape(())

    @staticmethod
    def backward(ctx, g):
        a, b, sq = ctx.saved_tensors
        g = _c(g.reshape(1))
        ga = ops.diff_scale(a, b, ctx.s, g, sq) if ctx.needs_input_grad[0] else None
        gb = ops.diff_scale(b, a, ctx.s, g, sq) if ctx.needs_input_grad[1] else None
        return ga, gb, None


def scaled_sqdist(s, a, b):
    return _ScaledSqDist.apply(a, b, s)


def scaled_dist(s, a, b):
    return _ScaledDist.apply(a, b, s)


class _Project(torch.autograd.Function):
    """rows -> their projection on span(base); the projector is symmetric, so backward is the same kernel."""

    @staticmethod
    def forward(ctx, w, qt, q_rows, identity):
        ctx.qt, ctx.q_rows, ctx.identity = qt, q_rows, identity
        w = _c(w.detach())
        return w.clone() if identity else ops.project_rows(w, qt, q_rows)

    @staticmethod
    def backward(ctx, g):
        g = _c(g)
        return (g.clone() if ctx.identity else ops.project_rows(g, ctx.qt, ctx.q_rows)), None, None, None


def project_rows(w, qt, q_rows, identity):
    return _Project.apply(w, qt, q_rows, identity)
