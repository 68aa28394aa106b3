"""Tensor-level entry points over the C-ABI (include/hsmm_b200.h).

Each function takes CUDA tensors, allocates outputs/workspaces with torch and enqueues the library's
kernels on torch's current stream.  The autograd Functions return the parameter-side gradients the
reference obtains by back-propagating through pytorch-struct and log_hsmm
(/root/reference/src/models/semimarkov/semimarkov.py:284-286)."""
import ctypes
import math

import torch

from . import _lib

LOG_2PI = math.log(2.0 * math.pi)


def _p(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _lib.HsmmError("hsmm_b200 runs on CUDA tensors only (no CPU fallback); got a %s tensor" % t.device)


def _f32(t):
    return None if t is None else t.detach().to(torch.float32).contiguous()


def ldc_of(C):
    return (C + 3) // 4 * 4


SPARSE_WIDTH = 4  # HSMM_SPARSE_WIDTH
FLAG_F64_STATE = 1  # HSMM_FLAG_F64_STATE


def sparse_transition_lists(allowed, device):
    """(pred, succ) int32 (C, 4) device tensors from a boolean [to, from] matrix of NOT-masked
    transitions, or None when some class has more than 4 unmasked neighbours (-> dense kernels)."""
    allowed = allowed.cpu()
    C = allowed.shape[0]
    if int(allowed.sum(dim=1).max()) > SPARSE_WIDTH or int(allowed.sum(dim=0).max()) > SPARSE_WIDTH:
        return None
    pred = torch.full((C, SPARSE_WIDTH), -1, dtype=torch.int32)
    succ = torch.full((C, SPARSE_WIDTH), -1, dtype=torch.int32)
    for c in range(C):
        p = torch.nonzero(allowed[c, :]).flatten()
        pred[c, :len(p)] = p.to(torch.int32)
        q = torch.nonzero(allowed[:, c]).flatten()
        succ[c, :len(q)] = q.to(torch.int32)
    return pred.to(device).contiguous(), succ.to(device).contiguous()


def prepare_lengths(lengths, device):
    """int32 device lengths + processing order (longest first) for load balance.  Host lengths go through
    pinned memory so that the copy never synchronises with earlier work on the stream."""
    if lengths.is_cuda:
        lengths_dev = lengths.to(device=device, dtype=torch.int32).contiguous()
    else:
        lengths_dev = lengths.to(torch.int32).contiguous().pin_memory().to(device, non_blocking=True)
    # longest first: with one video per warp the CTAs of short videos retire early and free their registers for the
    # kernels queued behind (dealing every CTA one video of each length quantile was measured: the CTAs then all end
    # together and the step is 8 % slower, r02q)
    order = torch.argsort(lengths_dev, descending=True, stable=True).to(torch.int32).contiguous()
    return lengths_dev, order


def emission_params(means, cov_diag):
    """(w, bias, inv_var, row_const) of the tied-diagonal Gaussian, cf.
    semimarkov_modules.py:324-362: log N(x; mu_c, diag(var)) = x.w_c + bias_c + rowc(x)."""
    inv_var = 1.0 / cov_diag
    w = means * inv_var
    bias = -0.5 * (means * w).sum(dim=1)
    D = means.shape[1]
    row_const = (-0.5 * torch.log(cov_diag.double()).sum() - 0.5 * D * LOG_2PI).to(torch.float32).reshape(1)  # stays on device
    return w, bias, inv_var, row_const


def emission_scores(features, means, cov_diag, penalty, lengths_i32, tensor_cores=True, params=None):
    """hsmm_emission: returns (em (B,T,ldc), rowterm (B,T), offset (B) float64).  `tensor_cores=False` withholds
    the workspace, which selects the SIMT kernel (used by the tests to compare the two)."""
    _need_cuda(features, means, cov_diag, penalty, lengths_i32)
    lib = _lib.load()
    B, T, D = features.shape
    C = means.shape[0]
    ldc = ldc_of(C)
    X = _f32(features)
    w, bias, inv_var, row_const = emission_params(_f32(means), _f32(cov_diag)) if params is None else params
    w, bias, inv_var = w.contiguous(), bias.contiguous(), inv_var.contiguous()
    pen = _f32(penalty)
    if pen is not None and tuple(pen.shape) != (B, T, C):
        raise _lib.HsmmError("constraints must be (B, T, C) = %s, got %s" % ((B, T, C), tuple(pen.shape)))
    em = torch.empty(B, T, ldc, device=X.device, dtype=torch.float32)
    rowterm = torch.empty(B, T, device=X.device, dtype=torch.float32)
    offset = torch.empty(B, device=X.device, dtype=torch.float64)
    ws_bytes = lib.hsmm_emission_workspace_bytes(D, C) if tensor_cores else 0
    ws = torch.empty(ws_bytes, device=X.device, dtype=torch.uint8) if ws_bytes else None
    _lib.check(lib.hsmm_emission(_p(X), _p(w), _p(bias), _p(inv_var), _p(row_const), _p(pen), _p(lengths_i32), B, T, D, C, ldc,
                                 _p(em), _p(rowterm), _p(offset), _p(ws), _stream()), "hsmm_emission")
    return em, rowterm, offset


def viterbi_decode(em, C, init, trans, lenp, end, offset, lengths_i32, order=None, class_ids=None, want_labels=True,
                   want_score=True, trans_pred=None):
    """hsmm_viterbi: returns (spans (B,T+1) int64, labels (B,T) int64 or None, score (B) float64 or None)."""
    _need_cuda(em, init, trans, lenp, end, offset, lengths_i32, order, class_ids)
    lib = _lib.load()
    B, T, ldc = em.shape
    K = lenp.shape[0]
    init, trans, lenp, end = _f32(init), _f32(trans), _f32(lenp), _f32(end)
    ws = torch.empty(lib.hsmm_viterbi_workspace_bytes(B, T, C, K), device=em.device, dtype=torch.uint8)
    spans = torch.empty(B, T + 1, device=em.device, dtype=torch.int64)
    labels = torch.empty(B, T, device=em.device, dtype=torch.int64) if want_labels else None
    score = torch.empty(B, device=em.device, dtype=torch.float64) if want_score else None
    _lib.check(lib.hsmm_viterbi(_p(em), ldc, _p(init), _p(trans), _p(trans_pred), _p(lenp), _p(end), _p(offset),
                                _p(lengths_i32), _p(order), _p(class_ids), B, T, C, K, _p(spans), _p(labels), _p(score),
                                _p(ws), _stream()), "hsmm_viterbi")
    return spans, labels, score


def logz_forward(em, C, init, trans, lenp, end, offset, lengths_i32, order=None, trans_pred=None, f64_state=False,
                 saved_bytes=None):
    """hsmm_logz_forward: returns (logz (B) float64, saved workspace).  `f64_state` (HSMM_FLAG_F64_STATE): keep the
    per-class DP state in double -- for score tensors that carry the -1e4 narration penalty."""
    _need_cuda(em, init, trans, lenp, end, offset, lengths_i32, order)
    lib = _lib.load()
    B, T, ldc = em.shape
    K = lenp.shape[0]
    flags = FLAG_F64_STATE if f64_state else 0
    nbytes = lib.hsmm_logz_saved_bytes(B, T, C, K, flags) if saved_bytes is None else saved_bytes
    saved = torch.empty(nbytes, device=em.device, dtype=torch.uint8)
    logz = torch.empty(B, device=em.device, dtype=torch.float64)
    _lib.check(lib.hsmm_logz_forward(_p(em), ldc, _p(init), _p(trans), _p(trans_pred), _p(lenp), _p(end), _p(offset),
                                     _p(lengths_i32), _p(order), B, T, C, K, flags, _p(logz), _p(saved), _stream()),
               "hsmm_logz_forward")
    return logz, saved


def logz_backward(em, C, init, trans, lenp, end, lengths_i32, order, grad_logz, saved, out=None, trans_succ=None,
                  f64_state=False, d_em=None):
    """hsmm_logz_backward: returns (d_init (C), d_trans (C,C), d_len (K,C), d_em (B,T,ldc)).
    `out` may carry pre-allocated (zeroed) d_init/d_trans/d_len views of a packed gradient buffer."""
    lib = _lib.load()
    B, T, ldc = em.shape
    K = lenp.shape[0]
    dev = em.device
    if out is None:
        d_init = torch.zeros(C, device=dev)
        d_trans = torch.zeros(C, C, device=dev)
        d_len = torch.zeros(K, C, device=dev)
    else:
        d_init, d_trans, d_len = out
    if d_em is None:
        d_em = torch.empty(B, T, ldc, device=dev, dtype=torch.float32)
    g = _f32(grad_logz)
    _lib.check(lib.hsmm_logz_backward(_p(em), ldc, _p(init), _p(trans), _p(trans_succ), _p(lenp), _p(end), _p(lengths_i32),
                                      _p(order), _p(g), B, T, C, K, FLAG_F64_STATE if f64_state else 0, _p(saved), _p(d_init),
                                      _p(d_trans), _p(d_len), _p(d_em), _stream()), "hsmm_logz_backward")
    return d_init, d_trans, d_len, d_em


def _ptr(t):
    return None if t is None else t.data_ptr()


def grouped_dp(mode, batches):
    """hsmm_dp_grouped: the Viterbi (mode 0) / forward (1) / backward (2) pass of several task-homogeneous batches in ONE
    launch per kernel family.  `batches`: list of dicts with the per-batch tensors of the single-batch entry points
    (em, C, init, trans, lenp, end, offset, lengths_i32, order, trans_list [, class_ids, want_labels, f64_state, grad,
    saved, out=(d_init, d_trans, d_len), d_em]).  Returns a list of per-batch results shaped like the single-batch
    functions': mode 0 (spans, labels, None), mode 1 (logz, saved), mode 2 (d_init, d_trans, d_len, d_em), mode 3
    (forward and backward in one launch; needs trans_list, trans_list2 = successors, grad)
    (logz, saved, d_init, d_trans, d_len, d_em)."""
    lib = _lib.load()
    n = len(batches)
    arr = (_lib.DpTask * n)()
    keep, results = [], []
    for i, bt in enumerate(batches):
        em = bt["em"]
        B, T, ldc = em.shape
        C = bt["C"]
        init, trans, lenp, end = _f32(bt["init"]), _f32(bt["trans"]), _f32(bt["lenp"]), _f32(bt.get("end"))
        K = lenp.shape[0]
        flags = FLAG_F64_STATE if bt.get("f64_state") else 0
        t = arr[i]
        t.em, t.ldc, t.init, t.trans, t.trans_list, t.lenp, t.end = _ptr(em), ldc, _ptr(init), _ptr(trans), _ptr(bt["trans_list"]), \
            _ptr(lenp), _ptr(end)
        t.offset, t.lengths, t.order, t.class_ids = _ptr(bt.get("offset")), _ptr(bt["lengths_i32"]), _ptr(bt.get("order")), \
            _ptr(bt.get("class_ids"))
        t.B, t.Tmax, t.C, t.K, t.flags = B, T, C, K, flags
        keep.append((init, trans, lenp, end))
        dev = em.device
        if mode == 0:
            ws = torch.empty(lib.hsmm_viterbi_workspace_bytes(B, T, C, K), device=dev, dtype=torch.uint8)
            spans = torch.empty(B, T + 1, device=dev, dtype=torch.int64)
            labels = torch.empty(B, T, device=dev, dtype=torch.int64) if bt.get("want_labels", True) else None
            t.out_spans, t.out_labels, t.out_score, t.workspace = _ptr(spans), _ptr(labels), None, _ptr(ws)
            keep.append(ws)
            results.append((spans, labels, None))
        elif mode == 1:
            saved = torch.empty(lib.hsmm_logz_saved_bytes(B, T, C, K, flags), device=dev, dtype=torch.uint8)
            logz = torch.empty(B, device=dev, dtype=torch.float64)
            t.out_logz, t.saved = _ptr(logz), _ptr(saved)
            results.append((logz, saved))
        elif mode == 3:
            saved = torch.empty(lib.hsmm_logz_saved_bytes(B, T, C, K, flags), device=dev, dtype=torch.uint8)
            logz = torch.empty(B, device=dev, dtype=torch.float64)
            if bt.get("out") is None:
                d_init, d_trans, d_len = torch.zeros(C, device=dev), torch.zeros(C, C, device=dev), torch.zeros(K, C, device=dev)
            else:
                d_init, d_trans, d_len = bt["out"]
            d_em = bt.get("d_em")
            if d_em is None:
                d_em = torch.empty(B, T, ldc, device=dev, dtype=torch.float32)
            g = _f32(bt["grad"])
            keep.append(g)
            t.out_logz, t.saved, t.trans_list2 = _ptr(logz), _ptr(saved), _ptr(bt["trans_list2"])
            t.grad_logz, t.d_init, t.d_trans, t.d_len, t.d_em = _ptr(g), _ptr(d_init), _ptr(d_trans), _ptr(d_len), _ptr(d_em)
            results.append((logz, saved, d_init, d_trans, d_len, d_em))
        else:
            if bt.get("out") is None:
                d_init, d_trans, d_len = torch.zeros(C, device=dev), torch.zeros(C, C, device=dev), torch.zeros(K, C, device=dev)
            else:
                d_init, d_trans, d_len = bt["out"]
            d_em = bt.get("d_em")
            if d_em is None:
                d_em = torch.empty(B, T, ldc, device=dev, dtype=torch.float32)
            g = _f32(bt["grad"])
            keep.append(g)
            t.saved, t.grad_logz, t.d_init, t.d_trans, t.d_len, t.d_em = _ptr(bt["saved"]), _ptr(g), _ptr(d_init), _ptr(d_trans), \
                _ptr(d_len), _ptr(d_em)
            results.append((d_init, d_trans, d_len, d_em))
    _lib.check(lib.hsmm_dp_grouped(mode, n, arr, _stream()), "hsmm_dp_grouped")
    return results


def weighted_feature_sums(features, weights, C, lengths_i32):
    """hsmm_weighted_feature_sums: returns (wx (C,D), wsum (C))."""
    _need_cuda(features, weights, lengths_i32)
    lib = _lib.load()
    B, T, D = features.shape
    X = _f32(features)
    wx = torch.zeros(C, D, device=X.device)
    wsum = torch.zeros(C, device=X.device)
    _lib.check(lib.hsmm_weighted_feature_sums(_p(X), _p(weights), weights.shape[2], _p(lengths_i32), B, T, D, C, _p(wx),
                                              _p(wsum), _stream()), "hsmm_weighted_feature_sums")
    return wx, wsum


def feature_moments(features, lengths_i32):
    """hsmm_feature_moments: returns (sum x (D), sum x^2 (D)) in float64."""
    _need_cuda(features, lengths_i32)
    lib = _lib.load()
    B, T, D = features.shape
    X = _f32(features)
    sx = torch.zeros(D, device=X.device, dtype=torch.float64)
    sx2 = torch.zeros(D, device=X.device, dtype=torch.float64)
    _lib.check(lib.hsmm_feature_moments(_p(X), _p(lengths_i32), B, T, D, _p(sx), _p(sx2), _stream()), "hsmm_feature_moments")
    return sx, sx2


def upload_ragged(host, dev, lengths_host_i32, lengths_dev_i32=None):
    """hsmm_upload_ragged: copy the live rows of a padded pinned host batch (B,T,width) into the device buffer `dev`
    (same shape; its padding rows are left as they are) on the current stream.  Returns the bytes enqueued.
    With `lengths_dev_i32` (the lengths on the device) and a pinned `host` the copy is ONE kernel reading the host rows
    over PCIe (hsmm_upload_ragged_mapped) instead of one copy-engine transfer per video."""
    lib = _lib.load()
    if host.is_cuda or not dev.is_cuda or host.dtype != torch.float32 or dev.dtype != torch.float32:
        raise _lib.HsmmError("upload_ragged: host must be a CPU float32 tensor and dev a CUDA float32 tensor")
    if tuple(host.shape) != tuple(dev.shape) or not host.is_contiguous() or not dev.is_contiguous():
        raise _lib.HsmmError("upload_ragged: host and dev must be contiguous and of the same (B, T, width) shape")
    B, T, W = host.shape
    lh = lengths_host_i32.to(torch.int32).contiguous()
    if lengths_dev_i32 is not None and host.is_pinned() and W % 4 == 0:
        _need_cuda(lengths_dev_i32)
        _lib.check(lib.hsmm_upload_ragged_mapped(ctypes.c_void_p(host.data_ptr()), _p(dev), _p(lengths_dev_i32), B, T, W, _stream()),
                   "hsmm_upload_ragged_mapped")
        return int(lh.clamp(0, T).sum()) * W * 4
    _lib.check(lib.hsmm_upload_ragged(ctypes.c_void_p(host.data_ptr()), _p(dev), ctypes.c_void_p(lh.data_ptr()), B, T, W,
                                      _stream()), "hsmm_upload_ragged")
    return int(lh.clamp(0, T).sum()) * W * 4


def onehot_weights(labels_i32, C, lengths_i32):
    _need_cuda(labels_i32, lengths_i32)
    lib = _lib.load()
    B, T = labels_i32.shape
    ldc = ldc_of(C)
    out = torch.empty(B, T, ldc, device=labels_i32.device, dtype=torch.float32)
    _lib.check(lib.hsmm_onehot_weights(_p(labels_i32), _p(lengths_i32), B, T, C, ldc, _p(out), _stream()), "hsmm_onehot_weights")
    return out


class HsmmLogZ(torch.autograd.Function):
    """logZ_b of every video from (features, class means, tied diagonal variance, init, trans, len).

    forward : hsmm_emission -> hsmm_logz_forward
    backward: hsmm_logz_backward -> hsmm_weighted_feature_sums; gradients w.r.t. the class means, init,
              trans and the length table (the tiny parameter transforms stay in torch autograd)."""

    @staticmethod
    def forward(ctx, features, means, cov_diag, penalty, init, trans, lenp, end, lengths_i32, order, sparse=None,
                em_pack=None):
        """`em_pack` = (em, rowterm, offset) already computed by `emission_scores` for the same inputs (the combined
        train + decode call scores the emissions once)."""
        C = means.shape[0]
        em, rowterm, offset = emission_scores(features, means, cov_diag, penalty, lengths_i32) if em_pack is None else em_pack
        init_f, trans_f, lenp_f, end_f = _f32(init), _f32(trans), _f32(lenp), _f32(end)
        pred, succ = (None, None) if sparse is None else sparse
        # narration constraints put -1e4 offsets into the scores: keep the per-class DP state in double
        xp = penalty is not None
        logz, saved = logz_forward(em, C, init_f, trans_f, lenp_f, end_f, offset, lengths_i32, order, trans_pred=pred,
                                   f64_state=xp)
        ctx.save_for_backward(features, means, cov_diag, em, init_f, trans_f, lenp_f, lengths_i32, saved)
        ctx.end, ctx.order, ctx.C, ctx.succ, ctx.xp = end_f, order, C, succ, xp
        ctx.mark_non_differentiable(rowterm)
        return logz.to(torch.float32), logz, rowterm

    @staticmethod
    def backward(ctx, g32, g64, _g_row):
        features, means, cov_diag, em, init, trans, lenp, lengths_i32, saved = ctx.saved_tensors
        C = ctx.C
        g = torch.zeros_like(g32) if g32 is None else g32
        if g64 is not None:
            g = g + g64.to(g.dtype)
        d_init, d_trans, d_len, d_em = logz_backward(em, C, init, trans, lenp, ctx.end, lengths_i32, ctx.order, g, saved,
                                                     trans_succ=ctx.succ, f64_state=ctx.xp)
        d_means = None
        if ctx.needs_input_grad[1]:
            wx, wsum = weighted_feature_sums(features, d_em, C, lengths_i32)
            d_means = (wx - wsum[:, None] * means) / cov_diag[None, :]
        d_pen = d_em[:, :, :C] if ctx.needs_input_grad[3] else None
        return None, d_means, None, d_pen, d_init, d_trans, d_len, None, None, None, None, None


class HsmmGoldScore(torch.autograd.Function):
    """Score of given segmentations (generative objective p(x, y), semimarkov_modules.py:647-655)."""

    @staticmethod
    def forward(ctx, features, means, cov_diag, penalty, init, trans, lenp, end, lengths_i32, spans_i32):
        lib = _lib.load()
        C = means.shape[0]
        em, rowterm, offset = emission_scores(features, means, cov_diag, penalty, lengths_i32)
        init_f, trans_f, lenp_f, end_f = _f32(init), _f32(trans), _f32(lenp), _f32(end)
        B, T, ldc = em.shape
        K = lenp_f.shape[0]
        score = torch.empty(B, device=em.device, dtype=torch.float64)
        _lib.check(lib.hsmm_gold_score(_p(em), ldc, _p(init_f), _p(trans_f), _p(lenp_f), _p(end_f), _p(offset), _p(lengths_i32),
                                       _p(spans_i32), None, B, T, C, K, _p(score), None, None, None, None, _stream()),
                   "hsmm_gold_score")
        ctx.save_for_backward(features, means, cov_diag, em, init_f, trans_f, lenp_f, lengths_i32, spans_i32, offset)
        ctx.end, ctx.C = end_f, C
        return score.to(torch.float32)

    @staticmethod
    def backward(ctx, g):
        lib = _lib.load()
        features, means, cov_diag, em, init, trans, lenp, lengths_i32, spans_i32, offset = ctx.saved_tensors
        C = ctx.C
        B, T, ldc = em.shape
        K = lenp.shape[0]
        dev = em.device
        d_init = torch.zeros(C, device=dev)
        d_trans = torch.zeros(C, C, device=dev)
        d_len = torch.zeros(K, C, device=dev)
        d_em = torch.empty(B, T, ldc, device=dev)
        score = torch.empty(B, device=dev, dtype=torch.float64)
        gg = _f32(g)
        _lib.check(lib.hsmm_gold_score(_p(em), ldc, _p(init), _p(trans), _p(lenp), _p(ctx.end), _p(offset), _p(lengths_i32),
                                       _p(spans_i32), _p(gg), B, T, C, K, _p(score), _p(d_init), _p(d_trans), _p(d_len),
                                       _p(d_em), _stream()), "hsmm_gold_score")
        d_means = None
        if ctx.needs_input_grad[1]:
            wx, wsum = weighted_feature_sums(features, d_em, C, lengths_i32)
            d_means = (wx - wsum[:, None] * means) / cov_diag[None, :]
        return None, d_means, None, None, d_init, d_trans, d_len, None, None, None
