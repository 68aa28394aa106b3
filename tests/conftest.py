import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture
def golden():
    import numpy as np

    def load(name):
        return dict(np.load(os.path.join(GOLDEN, name + ".npz")))

    return load


@pytest.fixture(params=["one_video_per_warp", "two_videos_per_warp"])
def pair_mode(request):
    """Run a test on both fast paths of the C <= 16 chain-constrained shapes (hsmm_set_pair_min_videos)."""
    import action_segmentation_b200 as pkg
    prev = pkg._lib.set_pair_min_videos(0 if request.param == "two_videos_per_warp" else -1)
    yield request.param
    pkg._lib.set_pair_min_videos(prev)
