"""pytest configuration: the ``gpu`` marker, import paths and golden-fixture loading."""
import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """GPU tests are skipped (not failed) when no device is visible, e.g. in the build container."""
    try:
        import torch

        has_gpu = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device visible")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    with np.load(os.path.join(GOLDEN, name + ".npz")) as z:
        return {k: z[k] for k in z.files}


def golden_names(prefix):
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN, prefix + "*.npz")))


def golden_kwargs(g):
    """Hyper-parameters stored as ``kw_<name>`` entries by make_golden_multiscale.py / make_golden_images.py."""
    out = {}
    for k, v in g.items():
        if k.startswith("kw_"):
            v = v.item()
            out[k[3:]] = None if v == "None" else v
    return out


@pytest.fixture(params=["small-kernels", "tiled-kernels"])
def sinkhorn_path(request):
    """Small Sinkhorn problems are served by the one-launch-per-iteration kernels (csrc/b200ot_small.cu); the golden
    cases are all small, so every test that takes this fixture runs twice: as routed by default, and with that path
    disabled so that the tiled TMA kernels (the ones that run at N = 1e6) are checked on the same fixtures."""
    from geomloss_b200 import sinkhorn_small

    keep = sinkhorn_small.SMALL_MAX
    if request.param == "tiled-kernels":
        sinkhorn_small.SMALL_MAX = 0
    yield request.param
    sinkhorn_small.SMALL_MAX = keep
