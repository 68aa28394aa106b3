"""TEST / BASELINE INFRASTRUCTURE ONLY -- never imported by the product path.

Copies the reference's OWN Python sources (/root/reference/src, pure Python, no build step) into the
git-ignored directory oracle/_ref/src so that they travel to the GPU box with the gpurun snapshot
(/root/reference does not exist there).  Nothing is copied into the repository's history.

    python oracle/make_ref.py          # also run by __graft_entry__.build() when /root/reference exists

Users: tests/golden/ref_import.py (falls back to oracle/_ref when /root/reference is absent), i.e. the
parity tests that run the UNMODIFIED reference `SemiMarkovModule` / `SemiMarkovModel` / `Accuracy` beside
the CUDA path, and `bench.py --impl reference` (cpu_baseline.kind = "reference").
The two third-party packages the reference needs and this image lacks are shimmed at import time:
torch_struct -> oracle/torch_struct_shim.py, editdistance -> a pure-Python Levenshtein.
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/src"
DST = os.path.join(HERE, "_ref", "src")


def make_ref(verbose=True):
    if not os.path.isdir(SRC):
        if verbose:
            print("oracle/make_ref.py: %s not present; keeping %s as it is" % (SRC, DST))
        return os.path.isdir(DST)
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    shutil.copytree(SRC, DST, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    n = sum(len(fs) for _, _, fs in os.walk(DST))
    if verbose:
        print("oracle/make_ref.py: copied %d files of the reference's src/ into %s" % (n, DST))
    return True


if __name__ == "__main__":
    sys.exit(0 if make_ref() else 1)
