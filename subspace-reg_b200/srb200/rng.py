"""The CPU random streams of one run, addressable per thread.

The reference consumes two process-global generators (torch's default CPU generator: classifier-row init, dropout and
DropBlock masks; numpy's legacy global RandomState: the replay-memory shot choice) after seeding both with --set_seed
(eval/language_eval.py:101-102).  A single run on the main thread uses exactly those globals, so callers that seed or
read them (eval_incremental.py:36-37) see the reference's behaviour.  To run several seeds CONCURRENTLY in one process
(srb200.concurrent: one host thread + CUDA stream per seed) every worker thread binds private generators with `scope()`;
a private torch.Generator / numpy RandomState seeded with s produces the same stream as the seeded global one, so the
numbers of a run do not depend on how many runs share the process.
"""
import contextlib
import math
import threading

import numpy as np
import torch

_tls = threading.local()


def generator():
    """The torch CPU generator of this thread's run (torch.default_generator unless a scope is bound)."""
    return getattr(_tls, "gen", None) or torch.default_generator


def private():
    return getattr(_tls, "gen", None) is not None


def np_random():
    """numpy's global generator interface (seed / choice / shuffle ...) for this thread's run."""
    return getattr(_tls, "np", None) or np.random


def get_state():
    return generator().get_state()


def set_state(state):
    generator().set_state(state)


def manual_seed(seed):
    """torch.manual_seed + np.random.seed for this thread's run."""
    if private():
        _tls.gen.manual_seed(int(seed))
        _tls.np.seed(int(seed))
    else:
        torch.manual_seed(seed)
        np.random.seed(seed)


@contextlib.contextmanager
def scope():
    """Bind private generators to the calling thread for the duration of one run."""
    prev = (getattr(_tls, "gen", None), getattr(_tls, "np", None))
    _tls.gen, _tls.np = torch.Generator(), np.random.RandomState()
    try:
        yield
    finally:
        _tls.gen, _tls.np = prev


def linear_init(out_features, in_features, bias):
    """The tensors nn.Linear(in_features, out_features, bias) would be initialised with (kaiming_uniform(a=sqrt(5)) weight,
    U(+-1/sqrt(fan_in)) bias), drawn from this thread's generator with the same consumption: -> (weight, bias | None)."""
    if not private():
        lin = torch.nn.Linear(in_features, out_features, bias=bias)
        return lin.weight.detach(), (lin.bias.detach() if bias else None)
    g = _tls.gen
    w = torch.empty(out_features, in_features)
    torch.nn.init.kaiming_uniform_(w, a=math.sqrt(5), generator=g)
    b = None
    if bias:
        bound = 1 / math.sqrt(in_features) if in_features > 0 else 0
        b = torch.empty(out_features)
        torch.nn.init.uniform_(b, -bound, bound, generator=g)
    return w, b
