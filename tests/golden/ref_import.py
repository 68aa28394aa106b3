"""Import the UNMODIFIED reference modules: from /root/reference (this container) or from the git-ignored
copy oracle/_ref/src that oracle/make_ref.py makes and gpurun ships to the GPU box.

The reference needs two third-party packages that are not installable here:
  * torch_struct  -> oracle/torch_struct_shim.py (restated algorithm, see its header)
  * editdistance  -> a tiny pure-Python Levenshtein
and `ReduceLROnPlateau(verbose=...)` (models/model.py:30-36) no longer exists in torch 2.11: the keyword is dropped
by a wrapper installed on torch.optim.lr_scheduler for the duration of the import context.
Test / baseline infrastructure only: the product package never imports this.
"""
import importlib
import os
import sys
import types

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
CANDIDATES = ["/root/reference/src", os.path.join(REPO, "oracle", "_ref", "src")]


def ref_src():
    for c in CANDIDATES:
        if os.path.isdir(os.path.join(c, "models", "semimarkov")):
            return c
    return None


REF_SRC = ref_src()


def reference_available():
    return ref_src() is not None


def _levenshtein(a, b):
    prev = list(range(len(b) + 1))
    for i, x in enumerate(a, 1):
        cur = [i]
        for j, y in enumerate(b, 1):
            cur.append(min(prev[j] + 1, cur[j - 1] + 1, prev[j - 1] + (x != y)))
        prev = cur
    return prev[-1]


def install_shims():
    if REPO not in sys.path:
        sys.path.insert(0, REPO)
    shim = importlib.import_module("oracle.torch_struct_shim")
    sys.modules.setdefault("torch_struct", shim)
    if "editdistance" not in sys.modules:
        ed = types.ModuleType("editdistance")
        ed.eval = _levenshtein
        sys.modules["editdistance"] = ed
    import torch
    sched = torch.optim.lr_scheduler
    if not getattr(sched.ReduceLROnPlateau, "_hsmm_compat", False):
        base = sched.ReduceLROnPlateau

        class ReduceLROnPlateau(base):  # accepts and ignores the removed `verbose` keyword
            _hsmm_compat = True

            def __init__(self, *a, verbose=None, **kw):
                super().__init__(*a, **kw)

        sched.ReduceLROnPlateau = ReduceLROnPlateau
    src = ref_src()
    if src is None:
        raise RuntimeError("reference sources not found (looked in %s); run `python oracle/make_ref.py` where "
                           "/root/reference exists" % ", ".join(CANDIDATES))
    if src not in sys.path:
        sys.path.insert(0, src)
    return src


def load_reference():
    """Returns (semimarkov_modules, semimarkov_utils) of the reference, imported as-is."""
    install_shims()
    mods = importlib.import_module("models.semimarkov.semimarkov_modules")
    utils = importlib.import_module("models.semimarkov.semimarkov_utils")
    return mods, utils


def load_reference_module(name):
    """Any other module of the reference's src/ tree, e.g. 'models.semimarkov.semimarkov', 'evaluation.accuracy', 'main'."""
    install_shims()
    return importlib.import_module(name)


if REPO not in sys.path:
    sys.path.insert(0, REPO)
from action_segmentation_b200.args import HsmmArgs as RefArgs  # noqa: E402,F401  (kept under its old name for make_golden.py)
