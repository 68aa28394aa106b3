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


@pytest.fixture(params=["one_video_per_warp", "two_videos_per_warp", "mixed_in_one_launch"])
def pair_mode(request):
    """Run a test on the fast paths of the C <= 16 chain-constrained shapes (hsmm_set_pair_min_videos): one video per warp,
    two videos per warp, and -- for grouped launches -- both families inside one launch (hsmm_set_mixed_min_videos)."""
    import action_segmentation_b200 as pkg
    mixed = request.param == "mixed_in_one_launch"
    prev_m = pkg._lib.set_mixed_min_videos(0 if mixed else -1)   # read before the pair switch: "never pair" hides it
    prev = pkg._lib.set_pair_min_videos({"one_video_per_warp": -1, "two_videos_per_warp": 0}.get(request.param, 1 << 30))
    yield request.param
    pkg._lib.set_pair_min_videos(prev)
    pkg._lib.set_mixed_min_videos(prev_m)
