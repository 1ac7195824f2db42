import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def _have_gpu():
    try:
        import slam_kinectfusion_b200 as kfb
        return kfb.load_library().kfb_device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    # `-m gpu` on a box without a GPU must fail loudly, not skip: the product has no CPU path.
    pass


@pytest.fixture(scope="session")
def kfo():
    from oracle import kfo as m
    m.build()
    return m


@pytest.fixture(scope="session")
def kfb():
    import slam_kinectfusion_b200 as m
    return m


def small_intr(kfo_or_kfb, w=320, h=240):
    s = w / 640.0
    return dict(width=w, height=h, fx=525.0 * s, fy=525.0 * s, cx=(319.5 + 0.5) * s - 0.5, cy=(239.5 + 0.5) * s - 0.5)


def make_pair(kfo, kfb, dims=64, w=320, h=240, **over):
    """Matching oracle / product descriptors for a small configuration."""
    ki = small_intr(None, w, h)
    Ko = kfo.Intr(**ki)
    Kb = kfb.Intrinsics(**ki)
    Po = kfo.default_params(dims)
    Pb = kfb.default_params(dims)
    for k, v in over.items():
        setattr(Po, k, v)
        setattr(Pb, k, v)
    return Ko, Kb, Po, Pb


def lsb_stats(a, b):
    """(exact fraction, max abs diff, count of |diff|>1) between two int16 arrays."""
    d = np.abs(a.astype(np.int32) - b.astype(np.int32))
    return float((d == 0).mean()), int(d.max()), int((d > 1).sum())
