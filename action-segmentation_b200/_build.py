"""Build libhsmm_b200.so in-tree with nvcc for sm_100a (no torch involved)."""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libhsmm_b200.so")
SOURCES = ["hsmm_api.cu", "hsmm_dp_host.cu", "hsmm_dp_vit.cu", "hsmm_dp_fwd.cu", "hsmm_dp_fwd_xp.cu", "hsmm_dp_bwd.cu",
           "hsmm_dp_bwd_xp.cu", "hsmm_dp_lin_fwd.cu", "hsmm_dp_lin_bwd.cu", "hsmm_dp_vit2.cu", "hsmm_dp_gen.cu", "hsmm_dp_pair.cu", "hsmm_dp_group.cu", "hsmm_aux.cu", "hsmm_emission_tc.cu", "hsmm_wsums_tc.cu"]
NVCC_FLAGS = ["-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "-Xptxas", "-v"] + os.environ.get("HSMM_EXTRA_NVCC_FLAGS", "").split()


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: cannot build libhsmm_b200.so")


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "hsmm_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """Compile every .cu for sm_100a and link the shared library next to this file."""
    if not force and not needs_build():
        return LIB
    nvcc = _nvcc()
    objdir = os.path.join(HERE, "csrc", "_obj")
    os.makedirs(objdir, exist_ok=True)
    procs = []
    for src in SOURCES:
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        log = open(os.path.join(objdir, src.replace(".cu", ".ptxas.log")), "w")
        cmd = [nvcc] + NVCC_FLAGS + ["-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, obj, log, subprocess.Popen(cmd, stdout=log, stderr=subprocess.STDOUT)))
    objs = []
    for src, obj, log, p in procs:
        rc = p.wait()
        log.close()
        if rc != 0:
            raise RuntimeError("nvcc failed on %s:\n%s" % (src, open(log.name).read()[-4000:]))
        objs.append(obj)
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static", "-o", LIB] + objs
    subprocess.check_call(cmd)
    if verbose:
        print("built", LIB)
    return LIB


if __name__ == "__main__":
    build(force=True, verbose=True)
