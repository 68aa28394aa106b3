/*
 * hsmm_b200 -- C-ABI of the B200-native HSMM hot path (libhsmm_b200.so).
 *
 * Drop-in boundary for the reference's `--classifier semimarkov` path
 * (dpfried/action-segmentation).  Every entry point names the reference interface it replaces
 * (paths relative to /root/reference/src).  The reference never calls native code of its own:
 * the arithmetic below is what models/semimarkov/semimarkov_modules.py builds out of torch ops
 * and hands to the un-vendored pytorch-struct (`SemiMarkovCRF`) / genbmm.  INTEGRATION.md shows
 * the ctypes stub a reference maintainer would add.
 *
 * Conventions
 *   - all pointers are DEVICE pointers owned by the caller (PyTorch), row-major, fp32 unless noted;
 *   - `stream` is a cudaStream_t passed as void*; every call is asynchronous on that stream and
 *     allocates nothing;
 *   - return value 0 = ok, <0 = argument / shape / capacity / CUDA error; the message is available
 *     from hsmm_last_error() (thread-local);
 *   - B videos, Tmax padded frames, D feature dims, C valid classes (EOS excluded), K rows of the
 *     length table (usable segment lengths 1..K-1, already clamped by the caller to the padded
 *     batch length as semimarkov_modules.py:450-452 does);
 *   - class scores use a leading dimension `ldc >= C` (multiple of 4 recommended);
 *   - transition scores are indexed [to, from] (semimarkov_modules.py:153-155, 320-322);
 *   - `end` is the EOS row of the augmented transition matrix (semimarkov_modules.py:462-471):
 *     0 where a class may end the video, -1e9 otherwise; NULL = every class may end;
 *   - `order` (optional) is the processing order of the videos, longest first, for load balance;
 *   - `trans_pred` / `trans_succ` (optional, (C, HSMM_SPARSE_WIDTH) int32, ascending, -1 padded): for every
 *     class the predecessors (successors) whose transition is NOT masked.  With
 *     --sm_constrain_transitions (semimarkov.py:41-54, data/crosstask.py:328-388) the matrix is a
 *     chain with self loops, so <= 2 entries per class survive the -1e9 mask
 *     (semimarkov_modules.py:298-322); the kernels then visit only the listed entries.  It is a
 *     hint, not a semantic change: a video whose result is degenerate (<= -1e8, i.e. no path avoids
 *     the masked transitions) is recomputed against the dense matrix inside the same kernel.
 *     NULL = dense.
 *
 * Score of a segmentation:
 *   init[c_0] + sum_i (len[l_i, c_i] + sum_{t in seg_i} em[t, c_i]) + sum_{i>=1} trans[c_i, c_{i-1}] + end[c_last]
 */
#ifndef HSMM_B200_H
#define HSMM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HSMM_OK 0
#define HSMM_ERR_ARG (-1)
#define HSMM_ERR_SHAPE (-2)
#define HSMM_ERR_CUDA (-3)
#define HSMM_SPARSE_WIDTH 4
/* flags of hsmm_logz_forward / hsmm_logz_backward / hsmm_logz_saved_bytes:
 * HSMM_FLAG_F64_STATE keeps the O(C) per-frame DP state (and the saved forward quantities) in double.
 * Set it when `em` carries large finite offsets -- the -1e4 narration penalty of
 * models/semimarkov/semimarkov.py:25,227-232 -- so that classes whose scores differ by multiples of
 * 1e4 keep their O(1) information; the O(C*K) span window stays in float either way. */
#define HSMM_FLAG_F64_STATE 1

/* Library version (major*100 + minor) and the last error message of the calling thread. */
int hsmm_version(void);
const char* hsmm_last_error(void);

/*
 * Emission scoring.  Replaces SemiMarkovModule._emission_log_probs_with_means / emission_log_probs
 * (models/semimarkov/semimarkov_modules.py:324-381): per-frame tied-diagonal Gaussian log density
 * of every valid class (+ additive narration constraints, semimarkov.py:227-232).
 *
 *   log N(x; mu_c, diag(var)) = x.w_c + bias_c + rowc(x),   w_c = mu_c / var,
 *   bias_c = -0.5 sum_d mu_cd^2 / var_d,   rowc(x) = -0.5 sum_d x_d^2 / var_d + row_const,
 *   row_const = -0.5 sum_d log var_d - 0.5 D log(2 pi).
 *
 * The class-independent part is kept apart so that the DP runs on well-conditioned numbers:
 *   em[b,t,c]  = x.w_c + bias_c + penalty[b,t,c] - shift[b,t],  shift = max_c(...)   (<= 0, best class 0)
 *   rowterm[b,t] = rowc(x_bt) + shift[b,t]        =>  elp[b,t,c] = em[b,t,c] + rowterm[b,t]
 *   offset[b]  = sum_{t < lengths[b]} rowterm[b,t]   (double; added to logZ / Viterbi scores)
 * Frames t >= lengths[b] get em = 0, rowterm = 0.
 *
 * X (B,Tmax,D); w (C,D); bias (C); inv_var (D); row_const: DEVICE pointer to one float (so that the caller never
 * synchronises to build it); penalty (B,Tmax,C) or NULL;
 * em (B,Tmax,ldc) out; rowterm (B,Tmax) out; offset (B) out (double, overwritten).
 * workspace: hsmm_emission_workspace_bytes(D, C) bytes (16-byte aligned) or NULL.  With a workspace and an
 * eligible shape (C <= 512, D % 4 == 0, X 16-byte aligned) the contraction runs on the tensor cores
 * (TMA-fed tcgen05.mma, operands split 3xTF32 so that the result is fp32-accurate; C <= 64 in ONE launch that
 * also shifts and stores, C > 64 in one launch per block of 64 classes followed by a row-shift kernel);
 * otherwise on a SIMT fp32 kernel.  hsmm_emission_workspace_bytes returns 0 for shapes without a tensor-core plan.
 */
size_t hsmm_emission_workspace_bytes(int D, int C);
int hsmm_emission(const float* X, const float* w, const float* bias, const float* inv_var, const float* row_const,
                  const float* penalty, const int32_t* lengths, int B, int Tmax, int D, int C, int ldc,
                  float* em, float* rowterm, double* offset, void* workspace, void* stream);

/*
 * Workspace sizes (bytes) for the DP entry points below, so that the caller allocates.
 *   hsmm_viterbi_workspace_bytes : back-pointer table
 *   hsmm_logz_saved_bytes        : forward quantities kept for hsmm_logz_backward (depends on flags)
 */
size_t hsmm_viterbi_workspace_bytes(int B, int Tmax, int C, int K);
size_t hsmm_logz_saved_bytes(int B, int Tmax, int C, int K, int flags);

/*
 * Max-plus Viterbi decode.  Replaces `SemiMarkovCRF(scores, lengths).argmax` +
 * `SemiMarkovCRF.struct.from_parts` + the valid->global id remap
 * (models/semimarkov/semimarkov_modules.py:660-696) and semimarkov_utils.spans_to_labels
 * (semimarkov_utils.py:51-63), without materialising the (B,T,K,C+1,C+1) potentials of
 * log_hsmm (semimarkov_modules.py:416-523).
 *
 * em (B,Tmax,ldc); init (C); trans (C,C) [to,from]; lenp (K,C) rows k=0..K-1; end (B,C) or NULL;
 * offset (B) double or NULL; lengths (B); order (B) or NULL; class_ids (C+1) int32 or NULL
 * (local -> global ids, last entry = EOS id; NULL = identity with EOS = C).
 * out_spans (B,Tmax+1) int64: class id at segment starts, -1 inside segments and past the EOS,
 *   EOS id at position lengths[b] (the reference's span encoding);
 * out_labels (B,Tmax) int64 or NULL: per-frame class ids (frames >= lengths[b] get the EOS id);
 * out_score (B) double or NULL: best path score (+ offset).
 * workspace: hsmm_viterbi_workspace_bytes(...) bytes.
 */
int hsmm_viterbi(const float* em, int ldc, const float* init, const float* trans, const int32_t* trans_pred,
                 const float* lenp, const float* end, const double* offset, const int32_t* lengths,
                 const int32_t* order, const int32_t* class_ids, int B, int Tmax, int C, int K,
                 int64_t* out_spans, int64_t* out_labels, double* out_score,
                 void* workspace, void* stream);

/*
 * Log-semiring forward pass.  Replaces `SemiMarkovCRF(scores, lengths).partition`
 * (models/semimarkov/semimarkov_modules.py:624,657): out_logz[b] = log sum over segmentations
 * (+ offset[b]).  `saved` (hsmm_logz_saved_bytes) receives what the backward pass needs.
 */
int hsmm_logz_forward(const float* em, int ldc, const float* init, const float* trans, const int32_t* trans_pred,
                      const float* lenp, const float* end, const double* offset, const int32_t* lengths,
                      const int32_t* order, int B, int Tmax, int C, int K, int flags, double* out_logz, void* saved,
                      void* stream);

/*
 * Backward pass / expected counts.  Replaces `loss.backward()` through pytorch-struct and log_hsmm
 * (models/semimarkov/semimarkov.py:284-286): with grad_logz[b] = d loss / d logZ_b it ACCUMULATES
 *   d_init (C)     += sum_b g_b * P(first segment has class c)
 *   d_trans (C,C)  += sum_b g_b * E[# transitions c1 -> c2]            ([to,from])
 *   d_len (K,C)    += sum_b g_b * E[# segments of class c and length k]
 * and WRITES d_em (B,Tmax,ldc) = g_b * P(frame t has class c) (0 for t >= lengths[b]).
 * Must follow hsmm_logz_forward with the same inputs, flags and `saved` buffer.
 */
int hsmm_logz_backward(const float* em, int ldc, const float* init, const float* trans, const int32_t* trans_succ,
                       const float* lenp, const float* end, const int32_t* lengths, const int32_t* order,
                       const float* grad_logz, int B, int Tmax, int C, int K, int flags, const void* saved,
                       float* d_init, float* d_trans, float* d_len, float* d_em, void* stream);

/*
 * Grouped DP launches: the Viterbi / forward / backward pass of SEVERAL task-homogeneous batches in one call.
 * The reference trains on mini-batches of one task at a time (data/corpus.py:613-644; models/semimarkov/semimarkov.py:
 * 222-266): every batch has its own class set, hence its own init / trans / len tables.  Calling the three entry points
 * above once per batch launches kernels that each keep only a few dozen CTAs busy for as long as that batch's longest
 * video lasts.  Here one kernel per family covers all the batches (parameter blocks indexed by block id), which is what
 * fills a B200 when a whole split is decoded or a large step is taken.
 *   mode 0: hsmm_viterbi, 1: hsmm_logz_forward, 2: hsmm_logz_backward -- same semantics, same buffers per task;
 *   mode 3: forward AND backward in one launch (a video's backward pass starts as soon as its own forward pass ends;
 *           grad_logz must be known up front, e.g. 1/B for the mean log-likelihood of semimarkov_modules.py:657).
 * Envelope: every task needs its sparse transition list (`trans_list` = predecessors for modes 0, 1 and 3, successors for
 * mode 2; `trans_list2` = successors for mode 3), K - 1 <= 20, C <= 32, ldc = C rounded up to 4, and the same `flags` for
 * all tasks; n <= 32 tasks per call.
 * Outside the envelope the call returns HSMM_ERR_SHAPE and the caller uses the per-batch entry points.
 */
typedef struct {
    const float* em; int ldc;
    const float* init; const float* trans; const int32_t* trans_list; const float* lenp; const float* end;
    const double* offset; const int32_t* lengths; const int32_t* order; const int32_t* class_ids;
    int B, Tmax, C, K, flags;
    int64_t* out_spans; int64_t* out_labels; double* out_score; void* workspace;   /* mode 0 */
    double* out_logz; void* saved;                                                   /* modes 1, 2 */
    const float* grad_logz; float* d_init; float* d_trans; float* d_len; float* d_em; /* modes 2, 3 */
    const int32_t* trans_list2;                                                      /* mode 3: successors */
} hsmm_dp_task;
int hsmm_dp_grouped(int mode, int n, const hsmm_dp_task* tasks, void* stream);

/*
 * Class-weighted feature sums.  The reduction behind d loss / d gaussian_means (autograd through
 * emission_log_probs, semimarkov_modules.py:324-381) and behind the supervised class means
 * r^T X of semimarkov_sufficient_stats (semimarkov_utils.py:74-126):
 *   out_wx (C,D)  += sum_{b,t<len_b} weights[b,t,c] * X[b,t,:]
 *   out_wsum (C)  += sum_{b,t<len_b} weights[b,t,c]
 * weights (B,Tmax,ldc).  Both outputs are accumulated into (caller zeroes them).  Frames t >= lengths[b] are never
 * read into the sums, whatever they hold (NaN included), in X or in weights.
 * Eligible shapes (D <= 896, D % 4 == 0, C <= 96, ldc % 4 == 0, X and weights 16-byte aligned) run on the tensor cores
 * (TMA-fed tcgen05.mma over MN-major tf32 operands, 3xTF32 so that the sums are fp32-accurate; one launch per block of
 * 224 features x 32 classes, so D <= 224 and C <= 32 is a single pass over X); the others on a SIMT fp32 kernel.
 */
int hsmm_weighted_feature_sums(const float* X, const float* weights, int ldc, const int32_t* lengths,
                               int B, int Tmax, int D, int C, float* out_wx, float* out_wsum, void* stream);

/*
 * Score of given (gold) segmentations.  Replaces `SemiMarkovCRF.struct.to_parts` + `struct().score`
 * (models/semimarkov/semimarkov_modules.py:626-655) by a direct gather-sum over the segments.
 * spans (B,Tmax) int32 in LOCAL class ids, -1 = continuation (semimarkov_utils.labels_to_spans).
 * out_score (B) double (+ offset).  With grad_score != NULL it also accumulates the (one-hot) counts
 * d_init/d_trans/d_len and writes d_em exactly like hsmm_logz_backward.
 * A segmentation the model cannot score -- a start label < -1 or >= C, frame 0 not a segment start, a segment longer
 * than K-1 -- gives out_score[b] = NaN (the reference raises inside struct.to_parts / score for those).
 */
int hsmm_gold_score(const float* em, int ldc, const float* init, const float* trans, const float* lenp,
                    const float* end, const double* offset, const int32_t* lengths, const int32_t* spans,
                    const float* grad_score, int B, int Tmax, int C, int K, double* out_score,
                    float* d_init, float* d_trans, float* d_len, float* d_em, void* stream);

/*
 * Supervised sufficient statistics over labelled frames.  Replaces the numpy/sklearn pass of
 * semimarkov_sufficient_stats + get_diagonal_covariances (semimarkov_utils.py:66-126):
 *   out_sum_x (D) += sum x,  out_sum_x2 (D) += sum x^2   (double; tied diagonal variance)
 * over the frames t < lengths[b].  Class sums use hsmm_weighted_feature_sums with one-hot weights.
 */
int hsmm_feature_moments(const float* X, const int32_t* lengths, int B, int Tmax, int D,
                         double* out_sum_x, double* out_sum_x2, void* stream);

/*
 * Host -> device copy of a zero-padded batch as `padding_colate` delivers it (models/model.py:42-63): only the live
 * rows t < lengths[b] of every video cross PCIe (CrossTask-shaped batches are ~1/3 padding).  `host` (B,Tmax,width)
 * should be pinned, `dev` (B,Tmax,width) is the device buffer (rows >= lengths[b] are left untouched: zero them once),
 * `lengths_host` is a HOST array.  Asynchronous on `stream`.
 */
int hsmm_upload_ragged(const float* host, float* dev, const int32_t* lengths_host, int B, int Tmax, int width, void* stream);
/*
 * The same copy done by a kernel: `host` must be pinned memory the device can address (cudaHostAlloc / torch
 * pin_memory), `lengths` a DEVICE array, width % 4 == 0, both buffers 16-byte aligned.  One launch instead of one
 * copy-engine transfer per video (each costs ~4 us of set-up: 45 instead of 55 GB/s at 1 MB per video).
 */
int hsmm_upload_ragged_mapped(const float* host, float* dev, const int32_t* lengths, int B, int Tmax, int width, void* stream);

/* One-hot weights from labels: weights[b,t,c] = (labels[b,t] == c) for t < lengths[b] else 0. */
int hsmm_onehot_weights(const int32_t* labels, const int32_t* lengths, int B, int Tmax, int C, int ldc,
                        float* weights, void* stream);

/* Introspection used by tests/bench: name of the DP kernel variant picked for a shape
 * ("reg<KR,S>/...", "lin+reg<...>", "general ...") and how many kernels the library has launched. */
#define HSMM_VARIANT_SPARSE 1    /* sparse transition lists given */
#define HSMM_VARIANT_F64_STATE 2 /* HSMM_FLAG_F64_STATE set on the DP call (NOT the same bit as HSMM_FLAG_F64_STATE) */
const char* hsmm_dp_variant(int C, int K, int mode /*0 viterbi, 1 forward, 2 backward*/,
                            int flags /* HSMM_VARIANT_* */);
uint64_t hsmm_launch_count(void);

/* Shapes beyond the register-resident kernels' envelope (hsmm_dp_variant names the kernel family of a shape) run on
 * general kernels -- one CTA per video, the span window in prefix-sum form in double precision, any K and up to 1024
 * classes -- so no shape the reference accepts is rejected; C > 1024 returns HSMM_ERR_SHAPE.  The workspace size
 * queries above already include their scratch area.  hsmm_set_generic_dp(1) (environment: HSMM_FORCE_GENERIC=1,
 * read before the first call) sends EVERY DP call to the general kernels (tests, A/B runs); set it before
 * querying workspace sizes.  Returns the previous setting. */
int hsmm_set_generic_dp(int force);

/* hsmm_logz_forward / hsmm_logz_backward run a linear-window (block floating point) kernel on the shapes that
 * fit one warp per video and recompute the videos it cannot certify with the log-domain kernel; results do not
 * depend on the switch.  hsmm_set_linear_window(0) keeps every video on the log-domain kernels (A/B runs);
 * returns the previous setting.  Environment: HSMM_DISABLE_LIN=1.  hsmm_dp_variant reports "lin+" in front
 * of the variant name when the linear-window kernel is used for that shape. */
int hsmm_set_linear_window(int enabled);

/* Chain-constrained shapes with C <= 16 classes and K - 1 <= 20 have a second fast path that carries TWO videos per warp
 * (1.6-1.7x fewer instructions per frame); it is used when a call (or a whole hsmm_dp_grouped group) has at least this
 * many videos -- below that a launch is latency-bound and the one-video-per-warp kernels finish sooner.  Default 4096
 * (environment: HSMM_PAIR_MIN_VIDEOS); 0 = always, negative = never.  Results do not depend on it.  Returns the
 * previous value. */
int hsmm_set_pair_min_videos(int n);

/* hsmm_dp_grouped, float state, forward+backward (mode 3): a group below the threshold above but with at least this many
 * videos runs its C <= 16 tasks two videos per warp and its other tasks one video per warp in ONE launch (a third fewer
 * warps for the same videos; two launches would run one after the other).  Default 1024 (environment:
 * HSMM_MIXED_MIN_VIDEOS); 0 = always, negative = never; never when hsmm_set_pair_min_videos is negative.  Results do not
 * depend on it.  Returns the previous value. */
int hsmm_set_mixed_min_videos(int n);

#ifdef __cplusplus
}
#endif
#endif /* HSMM_B200_H */
