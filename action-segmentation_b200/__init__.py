"""hsmm_b200: B200-native (sm_100a) HSMM hot path behind the reference's `--classifier semimarkov` API.

Package directory is `action-segmentation_b200/`; import it as `action_segmentation_b200` (the
repo-root shim of that name loads this directory)."""
from . import _lib, evaluation, hsmm, semimarkov_utils  # noqa: F401
from .args import HsmmArgs  # noqa: F401
from ._lib import HsmmError  # noqa: F401
from .semimarkov import SemiMarkovModel  # noqa: F401
from .semimarkov_modules import HsmmScores, SemiMarkovModule  # noqa: F401

__all__ = ["SemiMarkovModule", "SemiMarkovModel", "HsmmScores", "HsmmError", "HsmmArgs", "hsmm", "semimarkov_utils", "evaluation"]
