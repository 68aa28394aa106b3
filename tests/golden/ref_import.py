"""Import the UNMODIFIED reference HSMM modules from /root/reference (this container only).

The reference needs two third-party packages that are not installable here:
  * torch_struct  -> oracle/torch_struct_shim.py (restated algorithm, see its header)
  * editdistance  -> a tiny pure-Python Levenshtein (only imported, never on the HSMM path)
Nothing in the -m gpu tests, smoke() or bench.py may call this: /root/reference does not exist
on the GPU box.  It is used by make_golden.py and by CPU tests that skip when the tree is absent.
"""
import importlib
import os
import sys
import types

REF_ROOT = "/root/reference"
REF_SRC = os.path.join(REF_ROOT, "src")
REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def reference_available():
    return os.path.isdir(os.path.join(REF_SRC, "models", "semimarkov"))


def _levenshtein(a, b):
    prev = list(range(len(b) + 1))
    for i, x in enumerate(a, 1):
        cur = [i]
        for j, y in enumerate(b, 1):
            cur.append(min(prev[j] + 1, cur[j - 1] + 1, prev[j - 1] + (x != y)))
        prev = cur
    return prev[-1]


def load_reference():
    """Returns (semimarkov_modules, semimarkov_utils) of the reference, imported as-is."""
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    if REPO not in sys.path:
        sys.path.insert(0, REPO)
    shim = importlib.import_module("oracle.torch_struct_shim")
    sys.modules.setdefault("torch_struct", shim)
    if "editdistance" not in sys.modules:
        ed = types.ModuleType("editdistance")
        ed.eval = _levenshtein
        sys.modules["editdistance"] = ed
    if REF_SRC not in sys.path:
        sys.path.insert(0, REF_SRC)
    mods = importlib.import_module("models.semimarkov.semimarkov_modules")
    utils = importlib.import_module("models.semimarkov.semimarkov_utils")
    return mods, utils


class RefArgs:
    """Minimal argparse namespace the reference module reads (semimarkov_modules.py:54-65,
    semimarkov.py:16-31)."""

    def __init__(self, **kw):
        self.sm_max_span_length = 20
        self.sm_supervised_state_smoothing = 1e-2
        self.sm_supervised_length_smoothing = 1e-1
        self.sm_supervised_method = "closed-form"
        self.sm_feature_projection = False
        self.sm_init_non_projection_parameters_from = None
        self.sm_train_discriminatively = False
        self.__dict__.update(kw)
