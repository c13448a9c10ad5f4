import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "subspace-reg_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "slow: long CPU test")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def word_embed_dir(tmp_path_factory):
    """word_embeds/miniImageNet_dim500.pickle rebuilt from the committed fixture."""
    from srb200 import synthetic
    d = tmp_path_factory.mktemp("word_embeds")
    synthetic.write_word_embeds(os.path.join(GOLDEN, "word_embeds_dim500.npz"), str(d))
    return str(d)
