"""ctypes binding of libhsmm_b200.so (include/hsmm_b200.h).  PyTorch supplies device memory and the
stream; every compute step is one of the library's CUDA kernels.  There is no CPU fallback: a
missing library or a missing GPU is an error."""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libhsmm_b200.so")

EXPORTS = [
    "hsmm_version", "hsmm_last_error", "hsmm_emission", "hsmm_emission_workspace_bytes", "hsmm_viterbi_workspace_bytes", "hsmm_logz_saved_bytes",
    "hsmm_viterbi", "hsmm_logz_forward", "hsmm_logz_backward", "hsmm_weighted_feature_sums", "hsmm_gold_score",
    "hsmm_feature_moments", "hsmm_onehot_weights", "hsmm_dp_variant", "hsmm_launch_count", "hsmm_set_linear_window", "hsmm_set_generic_dp", "hsmm_upload_ragged", "hsmm_upload_ragged_mapped", "hsmm_dp_grouped", "hsmm_set_pair_min_videos", "hsmm_set_mixed_min_videos",
]

_lib = None


class DpTask(ctypes.Structure):
    """`hsmm_dp_task` of include/hsmm_b200.h (one batch of a grouped DP launch)."""
    _p, _i = ctypes.c_void_p, ctypes.c_int
    _fields_ = [("em", _p), ("ldc", _i), ("init", _p), ("trans", _p), ("trans_list", _p), ("lenp", _p), ("end", _p),
                ("offset", _p), ("lengths", _p), ("order", _p), ("class_ids", _p),
                ("B", _i), ("Tmax", _i), ("C", _i), ("K", _i), ("flags", _i),
                ("out_spans", _p), ("out_labels", _p), ("out_score", _p), ("workspace", _p),
                ("out_logz", _p), ("saved", _p),
                ("grad_logz", _p), ("d_init", _p), ("d_trans", _p), ("d_len", _p), ("d_em", _p), ("trans_list2", _p)]


class HsmmError(RuntimeError):
    pass


def load():
    """dlopen the in-tree library and declare the prototypes.  Raises if it was never built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise HsmmError(
            "libhsmm_b200.so is missing (%s). Build it with `python -c 'import __graft_entry__ as g; g.build()'`; "
            "there is no CPU fallback for the HSMM path." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    p, i, f, sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_size_t
    lib.hsmm_version.restype = i
    lib.hsmm_last_error.restype = ctypes.c_char_p
    lib.hsmm_launch_count.restype = ctypes.c_uint64
    lib.hsmm_dp_variant.restype = ctypes.c_char_p
    lib.hsmm_dp_variant.argtypes = [i, i, i, i]
    lib.hsmm_set_linear_window.restype = i
    lib.hsmm_set_linear_window.argtypes = [i]
    lib.hsmm_set_generic_dp.restype = i
    lib.hsmm_set_generic_dp.argtypes = [i]
    lib.hsmm_viterbi_workspace_bytes.restype = sz
    lib.hsmm_viterbi_workspace_bytes.argtypes = [i, i, i, i]
    lib.hsmm_logz_saved_bytes.restype = sz
    lib.hsmm_logz_saved_bytes.argtypes = [i, i, i, i, i]
    lib.hsmm_emission.argtypes = [p, p, p, p, p, p, p, i, i, i, i, i, p, p, p, p, p]
    lib.hsmm_emission_workspace_bytes.restype = sz
    lib.hsmm_emission_workspace_bytes.argtypes = [i, i]
    lib.hsmm_viterbi.argtypes = [p, i, p, p, p, p, p, p, p, p, p, i, i, i, i, p, p, p, p, p]
    lib.hsmm_logz_forward.argtypes = [p, i, p, p, p, p, p, p, p, p, i, i, i, i, i, p, p, p]
    lib.hsmm_logz_backward.argtypes = [p, i, p, p, p, p, p, p, p, p, i, i, i, i, i, p, p, p, p, p, p]
    lib.hsmm_weighted_feature_sums.argtypes = [p, p, i, p, i, i, i, i, p, p, p]
    lib.hsmm_gold_score.argtypes = [p, i, p, p, p, p, p, p, p, p, i, i, i, i, p, p, p, p, p, p]
    lib.hsmm_feature_moments.argtypes = [p, p, i, i, i, p, p, p]
    lib.hsmm_onehot_weights.argtypes = [p, p, i, i, i, i, p, p]
    lib.hsmm_upload_ragged.argtypes = [p, p, p, i, i, i, p]
    lib.hsmm_upload_ragged.restype = i
    lib.hsmm_upload_ragged_mapped.argtypes = [p, p, p, i, i, i, p]
    lib.hsmm_upload_ragged_mapped.restype = i
    lib.hsmm_dp_grouped.argtypes = [i, i, ctypes.POINTER(DpTask), p]
    lib.hsmm_dp_grouped.restype = i
    lib.hsmm_set_pair_min_videos.argtypes = [i]
    lib.hsmm_set_pair_min_videos.restype = i
    lib.hsmm_set_mixed_min_videos.argtypes = [i]
    lib.hsmm_set_mixed_min_videos.restype = i
    for name in ("hsmm_emission", "hsmm_viterbi", "hsmm_logz_forward", "hsmm_logz_backward", "hsmm_weighted_feature_sums",
                 "hsmm_gold_score", "hsmm_feature_moments", "hsmm_onehot_weights"):
        getattr(lib, name).restype = i
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        raise HsmmError("%s failed (%d): %s" % (what, rc, load().hsmm_last_error().decode()))


def launch_count():
    return int(load().hsmm_launch_count())


def dp_variant(C, K, mode, sparse=False, f64_state=False):
    return load().hsmm_dp_variant(int(C), int(K), int(mode), int(bool(sparse)) | (2 if f64_state else 0)).decode()


def set_generic_dp(force):
    """Send every DP call to the general kernels (hsmm_set_generic_dp); returns the previous setting."""
    return bool(load().hsmm_set_generic_dp(int(bool(force))))


def set_pair_min_videos(n):
    """Minimum number of videos in a call for the two-videos-per-warp kernels (hsmm_set_pair_min_videos)."""
    return int(load().hsmm_set_pair_min_videos(int(n)))


def set_mixed_min_videos(n):
    """Minimum number of videos of a grouped launch for the mixed-family kernels (hsmm_set_mixed_min_videos)."""
    return int(load().hsmm_set_mixed_min_videos(int(n)))


def set_linear_window(enabled):
    """Enable/disable the linear-window DP kernels (hsmm_set_linear_window); returns the previous setting."""
    return bool(load().hsmm_set_linear_window(int(bool(enabled))))
