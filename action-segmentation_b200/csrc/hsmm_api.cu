// C-ABI of libhsmm_b200.so (see include/hsmm_b200.h).  Argument checking, variant dispatch, error
// reporting.  No allocation, no synchronisation: every call enqueues kernels on the caller's stream.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/hsmm_b200.h"
#include "hsmm_common.cuh"

namespace hsmm {

static thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int check_launch(const char* what) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return HSMM_ERR_CUDA;
    }
    return HSMM_OK;
}

static int num_sms() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
            n = 148;
    }
    return n;
}

// implemented in the kernel translation units
bool dp_reg_supported(int C, int L, int mode, bool sparse, bool xp);
bool dp_lin_used(int C, int L, int mode, bool sparse, bool xp);
int dp_pair_set_min_videos(int n);
int dp_mixed_set_min_videos(int n);
int dp_lin_set_enabled(int on);
const char* dp_reg_name(int C, int L, int mode, bool sparse, bool xp);
int dp_reg_launch(DpParams p, int mode, cudaStream_t st);
size_t dp_gen_scratch_bytes(int C, int L);
bool dp_gen_shape_ok(int C);
int dp_gen_launch(DpParams p, int mode, void* scratch, cudaStream_t st);
int dp_group_launch(const DpParams* ps, int n, int mode, cudaStream_t st);

// HSMM_FORCE_GENERIC=1 / hsmm_set_generic_dp(1): every DP call takes the general kernels (tests, A/B runs)
static std::atomic<int> g_force_gen{-1};
static bool force_gen() {
    int v = g_force_gen.load(std::memory_order_relaxed);
    if (v < 0) {
        const char* e = getenv("HSMM_FORCE_GENERIC");
        v = (e && e[0] == '1') ? 1 : 0;
        g_force_gen.store(v, std::memory_order_relaxed);
    }
    return v == 1;
}
// does the call (mode, hint, precision) run on the general kernels?
static bool use_gen(int C, int L, int mode, bool sparse, bool xp) {
    return force_gen() || !dp_reg_supported(C, L, mode, sparse, xp);
}
// could ANY call on this shape need the general kernels' scratch area?  (the size queries do not know the hints)
static bool may_use_gen(int C, int L, bool logz, bool xp) {
    if (force_gen()) return true;
    for (int sparse = 0; sparse < 2; ++sparse) {
        if (logz) {
            if (!dp_reg_supported(C, L, 1, sparse, xp) || !dp_reg_supported(C, L, 2, sparse, xp)) return true;
        } else if (!dp_reg_supported(C, L, 0, sparse, false)) {
            return true;
        }
    }
    return false;
}
static size_t align256(size_t n) { return (n + 255) / 256 * 256; }
int launch_emission(const float*, const float*, const float*, const float*, const float*, const float*, const int32_t*, int, int, int,
                    int, int, float*, float*, double*, cudaStream_t);
int launch_upload_ragged(const float*, float*, const int32_t*, int, int, int, int, cudaStream_t);
size_t emission_tc_workspace_bytes(int D, int C);
int launch_emission_tc(const float*, const float*, const float*, const float*, const float*, const float*, const int32_t*, int, int,
                       int, int, int, float*, float*, double*, void*, int, cudaStream_t);
int launch_weighted_sums(const float*, const float*, int, const int32_t*, int, int, int, int, float*, float*, int, cudaStream_t);
int launch_moments(const float*, const int32_t*, int, int, int, double*, double*, int, cudaStream_t);
int launch_onehot(const int32_t*, const int32_t*, int, int, int, int, float*, int, cudaStream_t);
int launch_gold(const float*, int, const float*, const float*, const float*, const float*, const double*, const int32_t*,
                const int32_t*, const float*, int, int, int, int, double*, float*, float*, float*, float*, cudaStream_t);

static int check_dims(const char* fn, int B, int Tmax, int C, int K, int ldc) {
    if (B <= 0 || Tmax <= 0 || C <= 0) {
        set_error("%s: empty batch (B=%d Tmax=%d C=%d)", fn, B, Tmax, C);
        return HSMM_ERR_SHAPE;
    }
    if (K < 2) {
        set_error("%s: K=%d leaves no usable segment length (need K >= 2)", fn, K);
        return HSMM_ERR_SHAPE;
    }
    if (ldc < C) {
        set_error("%s: ldc=%d < C=%d", fn, ldc, C);
        return HSMM_ERR_SHAPE;
    }
    if (C > 65535 || K > 65535) {
        set_error("%s: C or K exceeds the 16-bit back-pointer fields", fn);
        return HSMM_ERR_SHAPE;
    }
    return HSMM_OK;
}

struct Saved {  // layout of the `saved` buffer of hsmm_logz_forward
    void* fbeta;
    void* fgamma;
    float* fdelta;
    double* logz2;
    float* fflag;
    float* bflag;
};
static size_t plane_elems(int B, int Tmax, int C) { return (size_t)B * (Tmax + 1) * (size_t)((C + 3) / 4 * 4); }
static size_t saved_bytes(int B, int Tmax, int C, bool xp) {
    const size_t st = xp ? sizeof(double) : sizeof(float);
    size_t n = 2 * plane_elems(B, Tmax, C) * st + (size_t)B * sizeof(double);  // fbeta, fgamma, logz2
    n += ((size_t)B * (Tmax + 1) + 2 * (size_t)B) * sizeof(float);              // fdelta, fflag, bflag
    return n + 16;
}
static Saved carve(void* saved, int B, int Tmax, int C, bool xp) {
    Saved s;
    const size_t st = xp ? sizeof(double) : sizeof(float);
    const size_t n = plane_elems(B, Tmax, C);  // even (ldc is a multiple of 4): the double planes stay 8-byte aligned
    char* base = reinterpret_cast<char*>(saved);
    s.logz2 = reinterpret_cast<double*>(base);
    base += (size_t)B * sizeof(double);
    s.fbeta = base;
    s.fgamma = base + n * st;
    s.fdelta = reinterpret_cast<float*>(base + 2 * n * st);
    s.fflag = s.fdelta + (size_t)B * (Tmax + 1);
    s.bflag = s.fflag + B;
    return s;
}

}  // namespace hsmm

using namespace hsmm;

extern "C" {

int hsmm_version(void) { return 100; }
const char* hsmm_last_error(void) { return g_err; }
uint64_t hsmm_launch_count(void) { return g_launches.load(); }

const char* hsmm_dp_variant(int C, int K, int mode, int flags) {
    const int L = K - 1;
    const bool sparse = (flags & HSMM_VARIANT_SPARSE) != 0, xp = (flags & HSMM_VARIANT_F64_STATE) != 0 && mode != 0;
    if (!use_gen(C, L, mode, sparse, xp)) {
        const char* nm = dp_reg_name(C, L, mode, sparse, xp);
        if (!dp_lin_used(C, L, mode, sparse, xp)) return nm;
        static thread_local char buf[160];
        snprintf(buf, sizeof(buf), "lin+%s", nm);
        return buf;
    }
    return dp_gen_shape_ok(C) ? "general (one CTA per video, f64 prefix-sum window in global memory)" : "unsupported";
}

int hsmm_set_generic_dp(int force) {
    const int prev = force_gen() ? 1 : 0;
    g_force_gen.store(force ? 1 : 0, std::memory_order_relaxed);
    return prev;
}

int hsmm_set_linear_window(int enabled) { return dp_lin_set_enabled(enabled); }
int hsmm_set_pair_min_videos(int n) { return dp_pair_set_min_videos(n); }
int hsmm_set_mixed_min_videos(int n) { return dp_mixed_set_min_videos(n); }

static size_t viterbi_base_bytes(int B, int Tmax, int C) {
    // [beta / back-pointer plane][predecessor plane][normaliser increments (B, Tmax+1)][flags (B)]
    return 2 * plane_elems(B, Tmax, C) * sizeof(uint32_t) + ((size_t)B * (Tmax + 1) + (size_t)B) * sizeof(float);
}

size_t hsmm_viterbi_workspace_bytes(int B, int Tmax, int C, int K) {
    size_t n = align256(viterbi_base_bytes(B, Tmax, C));
    if (K >= 2 && may_use_gen(C, K - 1, false, false)) n += dp_gen_scratch_bytes(C, K - 1);  // general kernels' scratch
    return n;
}

size_t hsmm_logz_saved_bytes(int B, int Tmax, int C, int K, int flags) {
    const bool xp = (flags & HSMM_FLAG_F64_STATE) != 0;
    size_t n = align256(saved_bytes(B, Tmax, C, xp));
    if (K >= 2 && may_use_gen(C, K - 1, true, xp)) n += dp_gen_scratch_bytes(C, K - 1);
    return n;
}

int hsmm_emission(const float* X, const float* w, const float* bias, const float* inv_var, const float* row_const,
                  const float* penalty, const int32_t* lengths, int B, int Tmax, int D, int C, int ldc, float* em,
                  float* rowterm, double* offset, void* workspace, void* stream) {
    if (!X || !w || !bias || !inv_var || !row_const || !lengths || !em || !rowterm || !offset) {
        set_error("hsmm_emission: null pointer");
        return HSMM_ERR_ARG;
    }
    if (B <= 0 || Tmax <= 0 || D <= 0 || C <= 0 || ldc < C) {
        set_error("hsmm_emission: bad shape B=%d Tmax=%d D=%d C=%d ldc=%d", B, Tmax, D, C, ldc);
        return HSMM_ERR_SHAPE;
    }
    if (B > 65535) {
        set_error("hsmm_emission: B=%d exceeds grid.y; split the batch", B);
        return HSMM_ERR_SHAPE;
    }
    // tensor-core path (TMA + tcgen05, 3xTF32) when the shape is eligible and a workspace was given
    const int rc = launch_emission_tc(X, w, bias, inv_var, row_const, penalty, lengths, B, Tmax, D, C, ldc, em, rowterm, offset,
                                      workspace, num_sms(), (cudaStream_t)stream);
    if (rc <= 0) return rc;
    return launch_emission(X, w, bias, inv_var, row_const, penalty, lengths, B, Tmax, D, C, ldc, em, rowterm, offset,
                           (cudaStream_t)stream);
}

size_t hsmm_emission_workspace_bytes(int D, int C) { return emission_tc_workspace_bytes(D, C); }

int hsmm_viterbi(const float* em, int ldc, const float* init, const float* trans, const int32_t* trans_pred, const float* lenp,
                 const float* end, const double* offset, const int32_t* lengths, const int32_t* order, const int32_t* class_ids, int B,
                 int Tmax, int C, int K, int64_t* out_spans, int64_t* out_labels, double* out_score, void* workspace,
                 void* stream) {
    if (!em || !init || !trans || !lenp || !lengths || !out_spans || !workspace) {
        set_error("hsmm_viterbi: null pointer");
        return HSMM_ERR_ARG;
    }
    int rc = check_dims("hsmm_viterbi", B, Tmax, C, K, ldc);
    if (rc) return rc;
    if (ldc != (C + 3) / 4 * 4) {
        set_error("hsmm_viterbi: ldc must be C rounded up to a multiple of 4 (got %d for C=%d)", ldc, C);
        return HSMM_ERR_SHAPE;
    }
    DpParams p;
    memset(&p, 0, sizeof(p));
    p.em = em; p.init = init; p.trans = trans; p.lenp = lenp; p.end = end; p.offset = offset;
    p.lengths = lengths; p.order = order; p.B = B; p.Tmax = Tmax; p.C = C; p.L = K - 1; p.ldc = ldc;
    p.bp = reinterpret_cast<uint32_t*>(workspace); p.class_ids = class_ids; p.spans = out_spans; p.labels = out_labels;
    p.vbeta = reinterpret_cast<float*>(workspace);
    p.vpred = reinterpret_cast<uint32_t*>(workspace) + plane_elems(B, Tmax, C);
    p.vdelta = reinterpret_cast<float*>(p.vpred + plane_elems(B, Tmax, C));
    p.vflag = p.vdelta + (size_t)B * (Tmax + 1);
    p.score = out_score; p.trans_pred = trans_pred;
    if (!use_gen(C, p.L, 0, trans_pred != nullptr, false)) return dp_reg_launch(p, 0, (cudaStream_t)stream);
    return dp_gen_launch(p, 0, reinterpret_cast<char*>(workspace) + align256(viterbi_base_bytes(B, Tmax, C)), (cudaStream_t)stream);
}

int hsmm_logz_forward(const float* em, int ldc, const float* init, const float* trans, const int32_t* trans_pred, const float* lenp,
                      const float* end, const double* offset, const int32_t* lengths, const int32_t* order, int B, int Tmax, int C, int K,
                      int flags, double* out_logz, void* saved, void* stream) {
    if (!em || !init || !trans || !lenp || !lengths || !out_logz || !saved) {
        set_error("hsmm_logz_forward: null pointer");
        return HSMM_ERR_ARG;
    }
    int rc = check_dims("hsmm_logz_forward", B, Tmax, C, K, ldc);
    if (rc) return rc;
    if (ldc != (C + 3) / 4 * 4) {
        set_error("hsmm_logz_forward: ldc must be C rounded up to a multiple of 4");
        return HSMM_ERR_SHAPE;
    }
    DpParams p;
    memset(&p, 0, sizeof(p));
    p.em = em; p.init = init; p.trans = trans; p.lenp = lenp; p.end = end; p.offset = offset;
    p.lengths = lengths; p.order = order; p.B = B; p.Tmax = Tmax; p.C = C; p.L = K - 1; p.ldc = ldc;
    p.xp = (flags & HSMM_FLAG_F64_STATE) ? 1 : 0;
    Saved s = carve(saved, B, Tmax, C, p.xp != 0);
    p.fbeta = s.fbeta; p.fgamma = s.fgamma; p.fdelta = s.fdelta; p.logz2 = s.logz2; p.fflag = s.fflag; p.bflag = s.bflag;
    p.logz = out_logz;
    p.trans_pred = trans_pred;
    // a shape whose BACKWARD pass needs the general kernels also runs its forward pass there: the general forward
    // keeps its state in double and only rounds when it stores the planes, so the posteriors of the (double) backward
    // pass do not inherit the float recursion's accumulated rounding over thousands of frames
    if (use_gen(C, p.L, 1, trans_pred != nullptr, p.xp != 0) || use_gen(C, p.L, 2, trans_pred != nullptr, p.xp != 0))
        return dp_gen_launch(p, 1, reinterpret_cast<char*>(saved) + align256(saved_bytes(B, Tmax, C, p.xp != 0)), (cudaStream_t)stream);
    return dp_reg_launch(p, 1, (cudaStream_t)stream);
}

int hsmm_logz_backward(const float* em, int ldc, const float* init, const float* trans, const int32_t* trans_succ, const float* lenp,
                       const float* end, const int32_t* lengths, const int32_t* order, const float* grad_logz, int B, int Tmax, int C, int K,
                       int flags, const void* saved, float* d_init, float* d_trans, float* d_len, float* d_em, void* stream) {
    if (!em || !init || !trans || !lenp || !lengths || !grad_logz || !saved || !d_init || !d_trans || !d_len || !d_em) {
        set_error("hsmm_logz_backward: null pointer");
        return HSMM_ERR_ARG;
    }
    int rc = check_dims("hsmm_logz_backward", B, Tmax, C, K, ldc);
    if (rc) return rc;
    if (ldc != (C + 3) / 4 * 4) {  // the saved planes are laid out with this stride
        set_error("hsmm_logz_backward: ldc must be C rounded up to a multiple of 4 (got %d for C=%d)", ldc, C);
        return HSMM_ERR_SHAPE;
    }
    DpParams p;
    memset(&p, 0, sizeof(p));
    p.em = em; p.init = init; p.trans = trans; p.lenp = lenp; p.end = end;
    p.lengths = lengths; p.order = order; p.B = B; p.Tmax = Tmax; p.C = C; p.L = K - 1; p.ldc = ldc;
    p.xp = (flags & HSMM_FLAG_F64_STATE) ? 1 : 0;
    Saved s = carve(const_cast<void*>(saved), B, Tmax, C, p.xp != 0);
    p.fbeta = s.fbeta; p.fgamma = s.fgamma; p.fdelta = s.fdelta; p.logz2 = s.logz2; p.fflag = s.fflag; p.bflag = s.bflag;
    p.trans_succ = trans_succ;
    p.grad = grad_logz; p.d_init = d_init; p.d_trans = d_trans; p.d_len = d_len; p.d_em = d_em;
    if (use_gen(C, p.L, 2, trans_succ != nullptr, p.xp != 0))
        return dp_gen_launch(p, 2, reinterpret_cast<char*>(const_cast<void*>(saved)) + align256(saved_bytes(B, Tmax, C, p.xp != 0)),
                             (cudaStream_t)stream);
    return dp_reg_launch(p, 2, (cudaStream_t)stream);
}

int hsmm_dp_grouped(int mode, int n, const hsmm_dp_task* tasks, void* stream) {
    if (!tasks || n <= 0 || mode < 0 || mode > 3) {
        set_error("hsmm_dp_grouped: bad arguments (mode=%d n=%d)", mode, n);
        return HSMM_ERR_ARG;
    }
    if (n > GROUP_MAX) {
        set_error("hsmm_dp_grouped: at most %d tasks per call (got %d)", GROUP_MAX, n);
        return HSMM_ERR_SHAPE;
    }
    static thread_local DpParams ps[GROUP_MAX];
    for (int i = 0; i < n; ++i) {
        const hsmm_dp_task& t = tasks[i];
        if (!t.em || !t.init || !t.trans || !t.trans_list || !t.lenp || !t.lengths) {
            set_error("hsmm_dp_grouped: task %d: null pointer", i);
            return HSMM_ERR_ARG;
        }
        int rc = check_dims("hsmm_dp_grouped", t.B, t.Tmax, t.C, t.K, t.ldc);
        if (rc) return rc;
        if (t.ldc != (t.C + 3) / 4 * 4 || t.flags != tasks[0].flags) {
            set_error("hsmm_dp_grouped: task %d: ldc must be C rounded up to 4 and flags must agree across the group", i);
            return HSMM_ERR_SHAPE;
        }
        DpParams& p = ps[i];
        memset(&p, 0, sizeof(p));
        p.em = t.em; p.init = t.init; p.trans = t.trans; p.lenp = t.lenp; p.end = t.end; p.offset = t.offset;
        p.lengths = t.lengths; p.order = t.order; p.B = t.B; p.Tmax = t.Tmax; p.C = t.C; p.L = t.K - 1; p.ldc = t.ldc;
        if (mode == 0) {
            if (!t.out_spans || !t.workspace) {
                set_error("hsmm_dp_grouped: task %d: Viterbi needs out_spans and workspace", i);
                return HSMM_ERR_ARG;
            }
            p.bp = reinterpret_cast<uint32_t*>(t.workspace); p.class_ids = t.class_ids; p.spans = t.out_spans; p.labels = t.out_labels;
            p.vbeta = reinterpret_cast<float*>(t.workspace);
            p.vpred = reinterpret_cast<uint32_t*>(t.workspace) + plane_elems(t.B, t.Tmax, t.C);
            p.vdelta = reinterpret_cast<float*>(p.vpred + plane_elems(t.B, t.Tmax, t.C));
            p.vflag = p.vdelta + (size_t)t.B * (t.Tmax + 1);
            p.score = t.out_score; p.trans_pred = t.trans_list;
        } else {
            if (!t.saved || ((mode == 1 || mode == 3) && !t.out_logz) ||
                (mode >= 2 && (!t.grad_logz || !t.d_init || !t.d_trans || !t.d_len || !t.d_em)) || (mode == 3 && !t.trans_list2)) {
                set_error("hsmm_dp_grouped: task %d: missing buffer for mode %d", i, mode);
                return HSMM_ERR_ARG;
            }
            p.xp = (t.flags & HSMM_FLAG_F64_STATE) ? 1 : 0;
            Saved sv = carve(t.saved, t.B, t.Tmax, t.C, p.xp != 0);
            p.fbeta = sv.fbeta; p.fgamma = sv.fgamma; p.fdelta = sv.fdelta; p.logz2 = sv.logz2; p.fflag = sv.fflag; p.bflag = sv.bflag;
            if (mode == 1 || mode == 3) {
                p.logz = t.out_logz; p.trans_pred = t.trans_list;
            }
            if (mode >= 2) {
                p.trans_succ = mode == 3 ? t.trans_list2 : t.trans_list; p.grad = t.grad_logz;
                p.d_init = t.d_init; p.d_trans = t.d_trans; p.d_len = t.d_len; p.d_em = t.d_em;
            }
        }
    }
    if (force_gen()) {
        set_error("hsmm_dp_grouped: not available while the general kernels are forced (hsmm_set_generic_dp)");
        return HSMM_ERR_SHAPE;
    }
    return dp_group_launch(ps, n, mode, (cudaStream_t)stream);
}

int hsmm_weighted_feature_sums(const float* X, const float* weights, int ldc, const int32_t* lengths, int B, int Tmax, int D,
                               int C, float* out_wx, float* out_wsum, void* stream) {
    if (!X || !weights || !lengths || !out_wx || !out_wsum) {
        set_error("hsmm_weighted_feature_sums: null pointer");
        return HSMM_ERR_ARG;
    }
    if (B <= 0 || Tmax <= 0 || D <= 0 || C <= 0 || ldc < C) {
        set_error("hsmm_weighted_feature_sums: bad shape");
        return HSMM_ERR_SHAPE;
    }
    return launch_weighted_sums(X, weights, ldc, lengths, B, Tmax, D, C, out_wx, out_wsum, num_sms(), (cudaStream_t)stream);
}

int hsmm_gold_score(const float* em, int ldc, const float* init, const float* trans, const float* lenp, const float* end,
                    const double* offset, const int32_t* lengths, const int32_t* spans, const float* grad_score, int B,
                    int Tmax, int C, int K, double* out_score, float* d_init, float* d_trans, float* d_len, float* d_em,
                    void* stream) {
    if (!em || !init || !trans || !lenp || !lengths || !spans || !out_score) {
        set_error("hsmm_gold_score: null pointer");
        return HSMM_ERR_ARG;
    }
    if (grad_score && (!d_init || !d_trans || !d_len || !d_em)) {
        set_error("hsmm_gold_score: gradient requested without output buffers");
        return HSMM_ERR_ARG;
    }
    int rc = check_dims("hsmm_gold_score", B, Tmax, C, K, ldc);
    if (rc) return rc;
    return launch_gold(em, ldc, init, trans, lenp, end, offset, lengths, spans, grad_score, B, Tmax, C, K - 1, out_score,
                       d_init, d_trans, d_len, d_em, (cudaStream_t)stream);
}

int hsmm_feature_moments(const float* X, const int32_t* lengths, int B, int Tmax, int D, double* out_sum_x,
                         double* out_sum_x2, void* stream) {
    if (!X || !lengths || !out_sum_x || !out_sum_x2) {
        set_error("hsmm_feature_moments: null pointer");
        return HSMM_ERR_ARG;
    }
    if (B <= 0 || Tmax <= 0 || D <= 0) {
        set_error("hsmm_feature_moments: bad shape");
        return HSMM_ERR_SHAPE;
    }
    return launch_moments(X, lengths, B, Tmax, D, out_sum_x, out_sum_x2, num_sms(), (cudaStream_t)stream);
}

int hsmm_upload_ragged(const float* host, float* dev, const int32_t* lengths_host, int B, int Tmax, int width, void* stream) {
    if (!host || !dev || !lengths_host) {
        set_error("hsmm_upload_ragged: null pointer");
        return HSMM_ERR_ARG;
    }
    if (B <= 0 || Tmax <= 0 || width <= 0) {
        set_error("hsmm_upload_ragged: bad shape B=%d Tmax=%d width=%d", B, Tmax, width);
        return HSMM_ERR_SHAPE;
    }
    const size_t row = (size_t)width * sizeof(float), vid = (size_t)Tmax * row;
    // runs of full-length videos are contiguous in both buffers: one copy per run, otherwise one per video
    int b = 0;
    while (b < B) {
        int len = lengths_host[b] < 0 ? 0 : (lengths_host[b] > Tmax ? Tmax : lengths_host[b]);
        int e = b + 1;
        size_t bytes = (size_t)len * row;
        if (len == Tmax) {
            while (e < B && lengths_host[e] >= Tmax) ++e;
            bytes = (size_t)(e - b) * vid;
        }
        if (bytes) {
            cudaError_t err = cudaMemcpyAsync(reinterpret_cast<char*>(dev) + (size_t)b * vid,
                                              reinterpret_cast<const char*>(host) + (size_t)b * vid, bytes,
                                              cudaMemcpyHostToDevice, (cudaStream_t)stream);
            if (err != cudaSuccess) {
                set_error("hsmm_upload_ragged: %s", cudaGetErrorString(err));
                return HSMM_ERR_CUDA;
            }
        }
        b = e;
    }
    return HSMM_OK;
}

int hsmm_upload_ragged_mapped(const float* host, float* dev, const int32_t* lengths, int B, int Tmax, int width, void* stream) {
    if (!host || !dev || !lengths) {
        set_error("hsmm_upload_ragged_mapped: null pointer");
        return HSMM_ERR_ARG;
    }
    if (B <= 0 || Tmax <= 0 || width <= 0 || width % 4 != 0) {
        set_error("hsmm_upload_ragged_mapped: bad shape B=%d Tmax=%d width=%d (width must be a multiple of 4)", B, Tmax, width);
        return HSMM_ERR_SHAPE;
    }
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, host) != cudaSuccess || at.type != cudaMemoryTypeHost || !at.devicePointer) {
        cudaGetLastError();
        set_error("hsmm_upload_ragged_mapped: host is not pinned memory the device can address (use hsmm_upload_ragged)");
        return HSMM_ERR_ARG;
    }
    if ((reinterpret_cast<uintptr_t>(at.devicePointer) & 15) || (reinterpret_cast<uintptr_t>(dev) & 15)) {
        set_error("hsmm_upload_ragged_mapped: buffers must be 16-byte aligned");
        return HSMM_ERR_ARG;
    }
    // 96 CTAs x 512 threads x 2 loads of 16 B = 1.5 MB in flight (r02q, configs[1] e2e: 8 CTAs 46.5, 24 CTAs 46.7, 48 CTAs
    // 47.1, 96 CTAs 47.8 GB/s; the per-video copy-engine path 44.8 GB/s); they use ~10 K registers each and leave the rest
    // of the SMs to the step that computes meanwhile
    static const int ctas = [] { const char* e = getenv("HSMM_UPLOAD_CTAS"); return e && atoi(e) > 0 ? atoi(e) : 96; }();
    return launch_upload_ragged(reinterpret_cast<const float*>(at.devicePointer), dev, lengths, B, Tmax, width, ctas,
                                (cudaStream_t)stream);
}

int hsmm_onehot_weights(const int32_t* labels, const int32_t* lengths, int B, int Tmax, int C, int ldc, float* weights,
                        void* stream) {
    if (!labels || !lengths || !weights) {
        set_error("hsmm_onehot_weights: null pointer");
        return HSMM_ERR_ARG;
    }
    if (B <= 0 || Tmax <= 0 || C <= 0 || ldc < C) {
        set_error("hsmm_onehot_weights: bad shape");
        return HSMM_ERR_SHAPE;
    }
    return launch_onehot(labels, lengths, B, Tmax, C, ldc, weights, num_sms(), (cudaStream_t)stream);
}

}  // extern "C"
