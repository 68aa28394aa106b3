// Register-resident semi-Markov DP kernels (sm_100a).
//
// One group of W warps works on one video; a CTA carries VPB independent groups.  A lane owns one
// class c and one k-slice j of that class (S slices per class inside a warp, CPW = 32/S classes per
// warp).  The span window never touches memory: lane (c, j) keeps, in registers,
//
//     A[i] ~ beta[n-k][c] + em[n-k..n-1, c],   k = j*KR + i + 1,  i = 0..KR-1
//
// i.e. the score of "some prefix, then a class-c segment that started k frames ago", updated every
// frame by one shift-and-add (A[i] <- A[i-1] + em[n-1,c]); the element that crosses a slice boundary
// moves with one shuffle.  This is the direct window sum of the reference (no prefix-sum
// cancellation), one add per (frame, class, length).  The duration scores len[k,c] are static per
// register slot: held in registers (LREG) or, for long windows, in a per-thread shared-memory column.
//
//   phase 1: gamma[n][c]  = (+)_k  A_k + len[k,c]                       (registers + slice shuffles)
//   phase 2: beta[n][c2]  = (+)_c1 gamma[n][c1] + trans[c2,c1]
//
// Transition modes (TM): 0 dense, matrix row in registers (one warp per video); 1 dense, matrix in
// shared memory; 2 sparse: the caller lists, per class, the <= 4 predecessors (successors for the
// backward pass) that are not masked (-1e9) -- the ordering-constrained models are chains -- and the
// kernel visits only those; a video whose result comes out degenerate (<= -1e8: no unmasked path)
// is recomputed in the same kernel against the dense matrix, so results never depend on the hint.
//
// Numerics.  All quantities of frame n are kept relative to a running normaliser nu_n (nu_{n+1} =
// nu_n + max_c gamma~[n][c]), so values stay O(1) however long the video is; nu is accumulated in
// double.  The log-semiring sums are single-pass: every term of class c at frame n is bounded above by
// r = max(gamma_prev + e + maxstep_c, beta + e + len[1,c]) (maxstep_c = max_k len[k,c] - len[k-1,c]),
// so the window is stored relative to r and sum_k ex2(A_k + len_k) can neither overflow nor (except
// when the whole mass sat in the slot that just left the window -- detected, exact two-pass fallback)
// underflow.  Base-2 domain, ex2/lg2 on the MUFU pipe.
//
// The (B,T,K,C,C) potentials of the reference (semimarkov_modules.py:416-523) are never formed.
#pragma once
#include <stdlib.h>

#include <type_traits>

#include "hsmm_common.cuh"

namespace hsmm {

// Per-class DP state type.  XP ("extended precision") keeps the O(C) per-frame quantities (beta, gamma,
// eta, zeta, the window reference) in double while the O(C*L) span window stays in float relative to
// that reference: with -1e4 narration penalties (semimarkov.py:25,227-232) classes whose scores differ
// by multiples of 1e4 all matter, and a float cannot hold O(1) information next to 1e4-sized offsets.
template <bool XP>
using state_t = typename std::conditional<XP, double, float>::type;
__device__ __forceinline__ float smax(float a, float b) { return fmaxf(a, b); }
__device__ __forceinline__ double smax(double a, double b) { return fmax(a, b); }
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v = fmax(v, __shfl_xor_sync(FULL, v, off));
    return v;
}
template <int S>
__device__ __forceinline__ double slice_max(double v) {
#pragma unroll
    for (int off = 32 / S; off < 32; off <<= 1) v = fmax(v, __shfl_xor_sync(FULL, v, off));
    return v;
}

constexpr int F = 4;    // frames per register prefetch chunk
constexpr int SPW = 4;  // sparse transition list width (HSMM_SPARSE_WIDTH)
constexpr size_t kSmemCap = 220 * 1024;
constexpr float DEGENERATE = -1.0e8f;  // natural-log units
constexpr float TINY = 9.094947e-13f;  // 2^-40

template <int S>
struct Lay {
    static constexpr int CPW = 32 / S;
    static constexpr int CRR = (CPW + S - 1) / S;  // dense transitions per lane when held in registers
};

__host__ __device__ inline int ld_trans(int W, int S) {
    const int cpw = 32 / S;
    const int cpad = W * cpw;
    return ((cpad + 31) / 32) * 32 + (S > 1 ? cpw : 0);
}

// ---------------------------------------------------------------------------------------------
// forward: Viterbi (VIT) or log-partition (FWD)
// ---------------------------------------------------------------------------------------------
template <bool VIT, bool XP, int KR, int S, int TM, bool LREG, int MAXT>
__device__ __forceinline__ void dp_forward_kernel_body(const DpParams& p, const int bid) {
    static_assert(!(VIT && XP), "extended-precision state is a log-semiring option");
    using ST = state_t<XP>;
    constexpr int CPW = Lay<S>::CPW;
    constexpr int CRR = Lay<S>::CRR;
    extern __shared__ __align__(16) float smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int W = p.W, C = p.C, L = p.L, ldc = p.ldc, Tmax = p.Tmax;
    const int G = W * 32;
    const int slot = warp / W, wig = warp - slot * W;
    const int gtid = wig * 32 + lane;
    const int cpad = W * CPW;
    const int ldT = ld_trans(W, S);
    const float SC = VIT ? 1.0f : LOG2E;
    constexpr int FG = (LREG && MAXT <= 512) ? F : 1;  // frames per prefetch group (register budget of the big variants)

    // shared layout: [transT C*ldT (TM==1)] [len columns KR*G (!LREG)] [per group: gamma 2*cpad (ST), warp maxima 2*W]
    float* transT = smem;
    float* lens = smem + (TM == 1 ? C * ldT : 0);
    float* gbase = lens + (LREG ? 0 : KR * G);
    gbase += (XP ? ((gbase - smem) & 1) : 0);  // 8-byte alignment of the double gamma rows
    constexpr int STW = sizeof(ST) / sizeof(float);
    const int per_group = 2 * cpad * STW + 2 * W;
    ST* gam_s = reinterpret_cast<ST*>(gbase + slot * per_group);
    float* wmax_s = gbase + slot * per_group + 2 * cpad * STW;

    const int cl = lane % CPW, j = lane / CPW;
    const int c = wig * CPW + cl;
    const bool valid = c < C;
    const bool owner = valid && j == 0;

    if constexpr (TM == 1) {
        for (int i = threadIdx.x; i < C * C; i += blockDim.x) {
            const int c2 = i / C, c1 = i - c2 * C;
            transT[c1 * ldT + c2] = p.trans[i] * SC;
        }
    }
    if constexpr (!LREG) {
        if (slot == 0) {
#pragma unroll 1
            for (int i = 0; i < KR; ++i) {
                const int k = j * KR + i + 1;
                lens[i * G + gtid] = (valid && k <= L) ? p.lenp[(size_t)k * C + c] * SC : NEG;
            }
        }
    }
    if constexpr (TM == 1 || !LREG) __syncthreads();

    const int vidx = bid * p.VPB + slot;
    if (vidx >= p.B) return;
    const int b = p.order ? p.order[vidx] : vidx;
    const int T = p.lengths[b];
    const int bar_id = 1 + slot;
    if constexpr (!VIT) {
        if (p.only_flagged && p.fflag[b] != 2.0f) return;  // certified by the linear-window kernel (hsmm_dp_lin.cuh)
    } else {
        if (p.only_flagged && p.vflag[b] == 0.0f) return;  // decoded by the deferred-arg-max kernel (hsmm_dp_vit2.cuh)
    }

    // ---- per-lane constants ----------------------------------------------------------------
    float ln[LREG ? KR : 1];
    const float* lnp = lens + gtid;
#define LN(i) (LREG ? ln[LREG ? (i) : 0] : lnp[(i) * G])
    float maxstep = NEG;  // max_k len[k] - len[k-1] over the usable lengths of this class
    {
        float prev = NEG;
#pragma unroll
        for (int i = 0; i < KR; ++i) {
            const int k = j * KR + i + 1;
            const float v = (valid && k <= L) ? p.lenp[(size_t)k * C + c] * SC : NEG;
            if constexpr (LREG) ln[i] = v;
            if (i == 0 && j > 0) prev = (valid && k - 1 <= L) ? p.lenp[(size_t)(k - 1) * C + c] * SC : NEG;
            if (k >= 2 && k <= L && valid) maxstep = fmaxf(maxstep, v - prev);
            prev = v;
        }
        maxstep = slice_max<S>(maxstep);
        if (maxstep < -1.0e29f) maxstep = 0.0f;  // L == 1 or unused lane: no old slot survives
    }
    const float ln_first = valid ? p.lenp[(size_t)C + c] * SC : NEG;  // len[1, c]
    const float init_c = valid ? p.init[c] * SC : NEG;
    const float* endb = p.end ? p.end + (size_t)b * C : nullptr;

    float tr[TM == 0 ? CRR : 1];
    float trmax = 0.0f;  // dense FWD: row maximum (upper bound of the transition term)
    if constexpr (TM == 0) {
        // blocked assignment: slice j owns c1 in [j*CRR, (j+1)*CRR)
#pragma unroll
        for (int i = 0; i < CRR; ++i) {
            const int c1 = j * CRR + i;
            tr[i] = (valid && c1 < C) ? p.trans[(size_t)c * C + c1] * SC : NEG;
        }
    }
    if constexpr (!VIT && TM != 2) {
        float m = NEG;
        if (valid)
            for (int c1 = 0; c1 < C; ++c1) m = fmaxf(m, p.trans[(size_t)c * C + c1] * SC);
        trmax = valid ? m : 0.0f;
        if constexpr (TM == 0) {
#pragma unroll
            for (int i = 0; i < CRR; ++i) tr[i] -= trmax;  // masked entries stay ~NEG
        }
    }
    int pidx[TM == 2 ? SPW : 1];
    float pval[TM == 2 ? SPW : 1];
    if constexpr (TM == 2) {
#pragma unroll
        for (int q = 0; q < SPW; ++q) {
            const int c1 = valid ? p.trans_pred[c * SPW + q] : -1;
            pidx[q] = c1 >= 0 ? c1 : 0;
            pval[q] = c1 >= 0 ? p.trans[(size_t)c * C + c1] * SC : NEG;
        }
    }

    const float* em_b = p.em + (size_t)b * Tmax * ldc;
    const size_t row0 = (size_t)b * (Tmax + 1);
    ST* const fbeta = reinterpret_cast<ST*>(p.fbeta);
    ST* const fgamma = reinterpret_cast<ST*>(p.fgamma);
    if (!VIT && owner) fbeta[row0 * ldc + c] = init_c;

    bool dense_pass = (TM != 2);  // TM == 2: first pass sparse, second (rare) pass dense from global memory
    ST final_v = 0;               // VIT: best score; FWD: log2 Z; both relative to nu_T
    int final_c = 0;
    double nu = 0.0;

#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
        float A[KR];
#pragma unroll
        for (int i = 0; i < KR; ++i) A[i] = NEG;
        ST beta = init_c;   // beta^[n-1][c], relative to nu_n
        ST gprev = NEG;     // gamma~[n-1][c], relative to nu_{n-1}
        ST rref = 0;        // FWD: rho_{n-1}; the window is stored against r_{n-1} = e_{n-1} + rho_{n-1}
        ST eprev = 0;       // FWD: e_{n-1}
        float gmprev = 0.0f;  // gm_{n-1} (a float by construction, also in XP: it is only a normaliser)
        nu = 0.0;
        const bool use_smem = (TM != 2) || dense_pass || W > 1;

        // emission prefetch: the next group of FG frames is in flight while this group is processed (static register
        // names -- a rotating ring makes every frame wait for the load issued one frame earlier); running output
        // pointers: no per-frame 64-bit index math
        const float* ep = em_b + c;
        float enext[FG];
#pragma unroll
        for (int f = 0; f < FG; ++f) enext[f] = (valid && f < T) ? __ldg(ep + f * ldc) : 0.0f;
        ep += FG * ldc;
        ST* gout = fgamma + (row0 + 1) * ldc + c;      // gamma[n]
        ST* bout = fbeta + (row0 + 1) * ldc + c;       // beta[n]
        uint32_t* pout = p.bp + (row0 + 1) * ldc + c;  // back-pointers of frame n
        float* dout = p.fdelta + row0 + 1;

#pragma unroll 1
        for (int n0 = 1; n0 <= T; n0 += FG) {
            float ecur[FG];
#pragma unroll
            for (int f = 0; f < FG; ++f) ecur[f] = enext[f];
#pragma unroll
            for (int f = 0; f < FG; ++f) enext[f] = (valid && n0 - 1 + FG + f < T) ? __ldg(ep + f * ldc) : 0.0f;
            ep += FG * ldc;
#pragma unroll
            for (int f = 0; f < FG; ++f) {
                const int n = n0 + f;
                if (n > T) break;
                const float efr = ecur[f];
                const ST e = (ST)efr * (ST)SC;
                nu += (double)gmprev;
                ST gamma;
                int bk = 0;
                // ---- phase 1 -----------------------------------------------------------------
                if constexpr (VIT) {
                    const float eo = e - gmprev;
                    float carry = 0.0f;
                    if (S > 1) carry = __shfl_up_sync(FULL, A[KR - 1], CPW);
#pragma unroll
                    for (int i = KR - 1; i > 0; --i) A[i] = A[i - 1] + eo;
                    A[0] = (j == 0) ? beta + e : carry + eo;
                    // two independent arg-max chains (even / odd slots); ties go to the smaller k
                    float b0 = A[0] + LN(0), b1 = NEG;
                    int k0 = 0, k1 = 1;
#pragma unroll
                    for (int i = 1; i < KR; ++i) {
                        const float v = A[i] + LN(i);
                        if (i & 1) {
                            if (v > b1 || i == 1) {
                                b1 = v;
                                k1 = i;
                            }
                        } else {
                            if (v > b0) {
                                b0 = v;
                                k0 = i;
                            }
                        }
                    }
                    if (KR > 1 && (b1 > b0 || (b1 == b0 && k1 < k0))) {
                        b0 = b1;
                        k0 = k1;
                    }
                    bk = k0 + j * KR + 1;
                    slice_argmax<S>(b0, bk);
                    gamma = valid ? b0 : NEG;
                } else {
                    // every term <= r_n = e + rho; the window is stored relative to that reference.  The
                    // (possibly huge: -1e4 narration penalty, -1e9 masks) emission never meets an O(1)
                    // number before it has cancelled: the shift of the old slots is
                    // r_{n-1} - gm_{n-1} - rho_n = (e_{n-1} - rho_n) + (rho_{n-1} - gm_{n-1}).
                    const ST rho = smax(gprev - (ST)gmprev + (ST)maxstep, beta + (ST)ln_first);
                    const float eo = (float)((eprev - rho) + (rref - (ST)gmprev));
                    float carry = 0.0f;
                    if (S > 1) carry = __shfl_up_sync(FULL, A[KR - 1], CPW);
#pragma unroll
                    for (int i = KR - 1; i > 0; --i) A[i] = A[i - 1] + eo;
                    A[0] = (j == 0) ? (float)(beta - rho) : carry + eo;
                    rref = rho;
                    eprev = e;
                    float sp[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
                    for (int i = 0; i < KR; ++i) sp[i & 3] += ex2(A[i] + LN(i));
                    float s = slice_sum<S>((sp[0] + sp[1]) + (sp[2] + sp[3]));
                    float mfix = 0.0f;
                    const bool bad = valid && !(s > TINY);
                    if (__any_sync(FULL, bad)) {
                        // the mass left the window: exact two-pass on the same registers
                        float m = A[0] + LN(0);
#pragma unroll
                        for (int i = 1; i < KR; ++i) m = fmaxf(m, A[i] + LN(i));
                        m = slice_max<S>(m);
                        float s2p = 0.0f;
#pragma unroll
                        for (int i = 0; i < KR; ++i) s2p += ex2(A[i] + LN(i) - m);
                        s2p = slice_sum<S>(s2p);
                        if (bad) {
                            s = s2p;
                            mfix = m;
                        }
                    }
                    gamma = valid ? (e + rho) + (ST)(mfix + lg2(s)) : (ST)NEG;
                }
                gprev = gamma;
                // ---- group maximum of gamma: the normaliser increment ----------------------------
                float gm = warp_max_redux(owner ? (float)gamma : NEG);
                ST* gs = gam_s + (n & 1) * cpad;
                if (W > 1 && lane == 0) wmax_s[(n & 1) * W + wig] = gm;
                if (use_smem) {
                    if (j == 0) gs[wig * CPW + cl] = valid ? gamma : (ST)NEG;
                    group_sync(W, bar_id);
                }
                if (W > 1) {
                    gm = NEG;
                    for (int q = 0; q < W; ++q) gm = fmaxf(gm, wmax_s[(n & 1) * W + q]);
                }
                if (!VIT && owner) *gout = gamma;
                gout += ldc;
                if (n == T) {
                    if (VIT && owner) *pout = (uint32_t)bk << 16;
                    break;
                }
                if (!VIT && gtid == 0) *dout = gm;
                ++dout;
                gmprev = gm;
                // ---- phase 2: beta^[n][c2] = (+)_c1 gamma~[n][c1] + trans[c2,c1] - gm ---------------
                if constexpr (VIT) {
                    float best = NEG;
                    int bc = 0;
                    if (TM == 2 && !dense_pass) {
                        if constexpr (TM == 2) {
#pragma unroll
                            for (int q = 0; q < SPW; ++q) {
                                const float gv = (W == 1) ? __shfl_sync(FULL, (float)gamma, pidx[q]) : (float)gs[pidx[q]];
                                const float v = gv + pval[q];
                                if (v > best || q == 0) {
                                    best = v;
                                    bc = pidx[q];
                                }
                            }
                        }
                    } else if constexpr (TM == 0) {
#pragma unroll
                        for (int i = 0; i < CRR; ++i) {
                            const float v = (float)gs[j * CRR + i] + tr[i];  // padding: gs = NEG, tr = NEG
                            if (v > best || i == 0) {
                                best = v;
                                bc = j * CRR + i;
                            }
                        }
                        slice_argmax<S>(best, bc);
                    } else if constexpr (TM == 1) {
                        bc = j;
                        for (int c1 = j; c1 < C; c1 += S) {
                            const float v = (float)gs[c1] + transT[c1 * ldT + c];
                            if (v > best || c1 == j) {
                                best = v;
                                bc = c1;
                            }
                        }
                        slice_argmax<S>(best, bc);
                    } else {  // TM == 2, dense fallback straight from global memory (rare)
                        bc = j;
                        for (int c1 = j; c1 < C; c1 += S) {
                            const float v = (float)gs[c1] + (valid ? __ldg(p.trans + (size_t)c * C + c1) : NEG);
                            if (v > best || c1 == j) {
                                best = v;
                                bc = c1;
                            }
                        }
                        slice_argmax<S>(best, bc);
                    }
                    beta = valid ? best - gm : NEG;
                    if (owner) *pout = ((uint32_t)bk << 16) | (uint32_t)bc;
                    pout += ldc;
                } else {
                    if (TM == 2 && !dense_pass) {
                        if constexpr (TM == 2) {
                            ST v[SPW];
                            ST m = NEG;
#pragma unroll
                            for (int q = 0; q < SPW; ++q) {
                                const ST gv = (W == 1) ? __shfl_sync(FULL, gamma, pidx[q]) : gs[pidx[q]];
                                v[q] = gv + (ST)pval[q];
                                m = smax(m, v[q]);
                            }
                            float s = 0.0f;
#pragma unroll
                            for (int q = 0; q < SPW; ++q) s += ex2((float)(v[q] - m));
                            beta = valid ? (m - (ST)gm) + (ST)lg2(s) : (ST)NEG;
                        }
                    } else {
                        // single pass against the bound gm + trmax; exact two-pass when it underflows
                        float s = 0.0f;
                        if constexpr (TM == 0) {
                            float sp[2] = {0.0f, 0.0f};
#pragma unroll
                            for (int i = 0; i < CRR; ++i) sp[i & 1] += ex2((float)(gs[j * CRR + i] - (ST)gm) + tr[i]);
                            s = sp[0] + sp[1];
                        } else if constexpr (TM == 1) {
                            const float off = trmax;
                            for (int c1 = j; c1 < C; c1 += S) s += ex2((float)(gs[c1] - (ST)gm) + (transT[c1 * ldT + c] - off));
                        }
                        s = slice_sum<S>(s);
                        ST mfix = 0;
                        const bool bad = (TM == 2) || (valid && !(s > TINY));
                        if (__any_sync(FULL, bad)) {
                            ST m = NEG;
                            for (int c1 = j; c1 < C; c1 += S)
                                m = smax(m, gs[c1] + (ST)(valid ? __ldg(p.trans + (size_t)c * C + c1) * SC : NEG));
                            m = slice_max<S>(m);
                            float s2p = 0.0f;
                            for (int c1 = j; c1 < C; c1 += S)
                                s2p += ex2((float)(gs[c1] + (ST)(valid ? __ldg(p.trans + (size_t)c * C + c1) * SC : NEG) - m));
                            s2p = slice_sum<S>(s2p);
                            if (bad) {
                                s = s2p;
                                mfix = (m - (ST)gm) - (ST)trmax;
                            }
                        }
                        beta = valid ? ((ST)trmax + mfix) + (ST)lg2(s) : (ST)NEG;
                    }
                    if (owner) *bout = beta;
                    bout += ldc;
                }
            }
        }

        // ---- termination -------------------------------------------------------------------
        // gamma~[T] sits in gam_s (dense / multi-warp) or only in registers (sparse, one warp): put it
        // in shared memory in every case so that one code path finishes the video.
        ST* gT = gam_s + (T & 1) * cpad;
        if (!use_smem) {
            if (j == 0) gT[wig * CPW + cl] = valid ? gprev : (ST)NEG;
            group_sync(W, bar_id);
        }
        if constexpr (!VIT) {
            ST m = NEG;
            for (int cc = lane; cc < C; cc += 32) m = smax(m, gT[cc] + (ST)(endb ? endb[cc] * SC : 0.0f));
            m = warp_max(m);
            float s = 0.0f;
            for (int cc = lane; cc < C; cc += 32) s += ex2((float)(gT[cc] + (ST)(endb ? endb[cc] * SC : 0.0f) - m));
            s = warp_sum(s);
            final_v = m + (ST)lg2(s);
        } else {
            float best = NEG;
            int bc = 0x7fffffff;
            for (int cc = lane; cc < C; cc += 32) {
                const float v = (float)gT[cc] + (endb ? endb[cc] : 0.0f);
                if (v > best || bc == 0x7fffffff) {
                    best = v;
                    bc = cc;
                }
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                const float ov = __shfl_xor_sync(FULL, best, off);
                const int oc = __shfl_xor_sync(FULL, bc, off);
                if (ov > best || (ov == best && oc < bc)) {
                    best = ov;
                    bc = oc;
                }
            }
            final_v = best;
            final_c = bc;
        }
        if (dense_pass) break;
        // sparse pass: accept unless the result is degenerate (no path through the listed transitions)
        const double total = (nu + (double)final_v) * (VIT ? 1.0 : LN2);
        if (total > (double)DEGENERATE) break;
        dense_pass = true;
        group_sync(W, bar_id);
    }

    if constexpr (!VIT) {
        if (gtid == 0) {
            p.logz2[b] = (double)final_v;
            p.fflag[b] = ((TM == 2 && dense_pass) ? 1.0f : 0.0f) + (p.only_flagged ? 4.0f : 0.0f);  // bit 2: recomputed here
            p.logz[b] = (nu + (double)final_v) * LN2 + (p.offset ? p.offset[b] : 0.0);
        }
    } else {
        // Viterbi: prefill outputs, then walk the back-pointers (warp 0 of the group).
        const int eos = p.class_ids ? p.class_ids[C] : C;
        int64_t* sp = p.spans + (size_t)b * (Tmax + 1);
        for (int i = gtid; i <= Tmax; i += G) sp[i] = (i == T) ? (int64_t)eos : (int64_t)-1;
        int64_t* lab = p.labels ? p.labels + (size_t)b * Tmax : nullptr;
        if (lab)
            for (int i = T + gtid; i < Tmax; i += G) lab[i] = eos;
        group_sync(W, bar_id);
        if (wig != 0) return;
        if (lane == 0 && p.score) p.score[b] = nu + (double)final_v + (p.offset ? p.offset[b] : 0.0);
        int n = T, cc = final_c;
        while (n > 0) {
            const uint32_t v = __ldcg(p.bp + (row0 + n) * ldc + cc);
            int k = (int)(v >> 16);
            k = k < 1 ? 1 : (k > n ? n : k);
            const int start = n - k;
            const int64_t cid = p.class_ids ? p.class_ids[cc] : cc;
            if (lane == 0) sp[start] = cid;
            if (lab)
                for (int t = start + lane; t < n; t += 32) lab[t] = cid;
            if (start > 0) {
                const uint32_t u = __ldcg(p.bp + (row0 + start) * ldc + cc);
                cc = (int)(u & 0xffffu);
                if (cc >= C) cc = C - 1;
            }
            n = start;
        }
    }
#undef LN
}

template <bool VIT, bool XP, int KR, int S, int TM, bool LREG, int MAXT>
__global__ void __launch_bounds__(MAXT) dp_forward_kernel(const DpParams p) {
    dp_forward_kernel_body<VIT, XP, KR, S, TM, LREG, MAXT>(p, blockIdx.x);
}
template <bool VIT, bool XP, int KR, int S, int TM, bool LREG, int MAXT>
__global__ void __launch_bounds__(MAXT) dp_forward_kernel_grouped(const __grid_constant__ DpGroup g) {
    int local;
    const int t = group_find(g, blockIdx.x, local);
    dp_forward_kernel_body<VIT, XP, KR, S, TM, LREG, MAXT>(g.t[t], local);
}

// ---------------------------------------------------------------------------------------------
// backward: expected counts
// ---------------------------------------------------------------------------------------------
template <bool XP, int KR, int S, int TM, bool LREG, int MAXT>
__device__ __forceinline__ void dp_backward_kernel_body(const DpParams& p, const int bid) {
    using ST = state_t<XP>;
    constexpr int CPW = Lay<S>::CPW;
    constexpr int CRR = Lay<S>::CRR;
    extern __shared__ __align__(16) float smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int W = p.W, C = p.C, L = p.L, ldc = p.ldc, Tmax = p.Tmax;
    const int G = W * 32;
    const int slot = warp / W, wig = warp - slot * W;
    const int gtid = wig * 32 + lane;
    const int cpad = W * CPW;
    const int ldT = ld_trans(W, S);
    const float SC = LOG2E;

    // shared: [trans C*ldT (TM==1)] [len columns KR*G (!LREG)] [per group: zeta 2*cpad (ST), Etr C*ldT (TM != 0)]
    float* trans_s = smem;
    float* lens = smem + (TM == 1 ? C * ldT : 0);
    float* gbase = lens + (LREG ? 0 : KR * G);
    gbase += (XP ? ((gbase - smem) & 1) : 0);
    constexpr int STW = sizeof(ST) / sizeof(float);
    const int per_group = 2 * cpad * STW + (TM == 0 ? 0 : C * ldT + (XP ? ((C * ldT) & 1) : 0));
    ST* zet_s = reinterpret_cast<ST*>(gbase + slot * per_group);
    float* etr_s = gbase + slot * per_group + 2 * cpad * STW;  // TM == 1 always, TM == 2 only in the dense fallback

    const int cl = lane % CPW, j = lane / CPW;
    const int c = wig * CPW + cl;
    const bool valid = c < C;
    const bool owner = valid && j == 0;

    if constexpr (TM == 1) {
        for (int i = threadIdx.x; i < C * C; i += blockDim.x) {
            const int c2 = i / C, c1 = i - c2 * C;
            trans_s[c2 * ldT + c1] = p.trans[i] * SC;
        }
    }
    if constexpr (TM != 0) {
        for (int g = 0; g < p.VPB; ++g) {
            float* e = gbase + g * per_group + 2 * cpad * STW;
            for (int i = threadIdx.x; i < C * ldT; i += blockDim.x) e[i] = 0.0f;
        }
    }
    if constexpr (!LREG) {
        if (slot == 0) {
#pragma unroll 1
            for (int i = 0; i < KR; ++i) {
                const int k = j * KR + i + 1;
                lens[i * G + gtid] = (valid && k <= L) ? p.lenp[(size_t)k * C + c] * SC : NEG;
            }
        }
    }
    __syncthreads();

    const int vidx = bid * p.VPB + slot;
    if (vidx >= p.B) return;
    const int b = p.order ? p.order[vidx] : vidx;
    const int T = p.lengths[b];
    const int bar_id = 1 + slot;
    if (p.only_flagged && p.bflag[b] == 0.0f) return;  // done by the linear-window kernel (hsmm_dp_lin.cuh)
    const bool dense_pass = (TM != 2) || (((int)p.fflag[b]) & 1);
    const bool use_smem = (TM != 2) || dense_pass || W > 1;

    float Bq[KR], El[KR];
    float ln[LREG ? KR : 1];
    const float* lnp = lens + gtid;
#define LN(i) (LREG ? ln[LREG ? (i) : 0] : lnp[(i) * G])
    float maxstep = NEG;
    {
        float prev = NEG;
#pragma unroll
        for (int i = 0; i < KR; ++i) {
            const int k = j * KR + i + 1;
            const float v = (valid && k <= L) ? p.lenp[(size_t)k * C + c] * SC : NEG;
            if constexpr (LREG) ln[i] = v;
            if (i == 0 && j > 0) prev = (valid && k - 1 <= L) ? p.lenp[(size_t)(k - 1) * C + c] * SC : NEG;
            if (k >= 2 && k <= L && valid) maxstep = fmaxf(maxstep, v - prev);
            prev = v;
            Bq[i] = NEG;
            El[i] = 0.0f;
        }
        maxstep = slice_max<S>(maxstep);
        if (maxstep < -1.0e29f) maxstep = 0.0f;
    }
    const float ln_first = valid ? p.lenp[(size_t)C + c] * SC : NEG;

    float tr[TM == 0 ? CRR : 1], Et[TM == 0 ? CRR : 1];
    if constexpr (TM == 0) {
#pragma unroll
        for (int i = 0; i < CRR; ++i) {
            const int c2 = j * CRR + i;
            tr[i] = (valid && c2 < C) ? p.trans[(size_t)c2 * C + c] * SC : NEG;
            Et[i] = 0.0f;
        }
    }
    int sidx[TM == 2 ? SPW : 1];
    float sval[TM == 2 ? SPW : 1], Es[TM == 2 ? SPW : 1];
    if constexpr (TM == 2) {
#pragma unroll
        for (int q = 0; q < SPW; ++q) {
            const int c2 = valid ? p.trans_succ[c * SPW + q] : -1;
            sidx[q] = c2 >= 0 ? c2 : 0;
            sval[q] = c2 >= 0 ? p.trans[(size_t)c2 * C + c] * SC : NEG;
            Es[q] = 0.0f;
        }
    }
    const float endc = valid ? (p.end ? p.end[(size_t)b * C + c] : 0.0f) * SC : NEG;
    const ST lzrel = (ST)p.logz2[b];  // log2 Z relative to nu_T
    const float w = p.grad[b];
    const float init_c = valid ? p.init[c] * SC : NEG;

    const float* em_b = p.em + (size_t)b * Tmax * ldc;
    const size_t row0 = (size_t)b * (Tmax + 1);
    const ST* fb = reinterpret_cast<const ST*>(p.fbeta) + row0 * ldc;
    const ST* fg = reinterpret_cast<const ST*>(p.fgamma) + row0 * ldc;
    const float* fd = p.fdelta + row0;
    float* dem = p.d_em + (size_t)b * Tmax * ldc;

    // Frame n: zeta^[n] = zeta[n] - mu_{n+1}, eta~[n] = eta[n] - mu_n with mu_n = logZ - nu_n, so that
    // posteriors are exp(forward + backward) of O(1) numbers:
    //   S[n,c] = exp(beta^[n][c] + zeta^[n][c]),  F[n,c] = exp(gamma~[n][c] + eta~[n][c]).
    ST eta = valid ? (ST)endc - lzrel : (ST)NEG;  // eta~[T]
    ST zprev = NEG;                               // zeta^[n+1][c]
    ST rref = 0, eprev = 0;                       // rho_{n+1}, e_{n+1}
    float occ = 0.0f, comp = 0.0f;                // Kahan-compensated occupancy
    float Fprev = valid ? w * ex2((float)(__ldg(fg + (size_t)T * ldc + c) + (ST)endc - lzrel)) : 0.0f;
    float Sprev = 0.0f;
    float gm_next = 0.0f;  // gm_{n+1}

    for (int i = T * ldc + gtid; i < Tmax * ldc; i += G) dem[i] = 0.0f;  // frames beyond the video

    float enext[F], dnext[F];
    ST bnext[F], gnext[F];
#pragma unroll
    for (int f = 0; f < F; ++f) {
        const int n = T - 1 - f;
        const bool ok = valid && n >= 0;
        dnext[f] = (n >= 1) ? __ldg(fd + n) : 0.0f;  // gm_n
        enext[f] = ok ? __ldg(em_b + (size_t)n * ldc + c) : 0.0f;
        bnext[f] = (ok && n > 0) ? __ldg(fb + (size_t)n * ldc + c) : (ST)0;
        gnext[f] = (ok && n > 0) ? __ldg(fg + (size_t)n * ldc + c) : (ST)0;
    }
    for (int n0 = T - 1; n0 >= 0; n0 -= F) {
        float ecur[F], dcur[F];
        ST bcur[F], gcur[F];
#pragma unroll
        for (int f = 0; f < F; ++f) {
            ecur[f] = enext[f];
            bcur[f] = bnext[f];
            gcur[f] = gnext[f];
            dcur[f] = dnext[f];
        }
#pragma unroll
        for (int f = 0; f < F; ++f) {
            const int n = n0 - F - f;
            const bool ok = valid && n >= 0;
            dnext[f] = (n >= 1) ? __ldg(fd + n) : 0.0f;
            enext[f] = ok ? __ldg(em_b + (size_t)n * ldc + c) : 0.0f;
            bnext[f] = (ok && n > 0) ? __ldg(fb + (size_t)n * ldc + c) : (ST)0;
            gnext[f] = (ok && n > 0) ? __ldg(fg + (size_t)n * ldc + c) : (ST)0;
        }
#pragma unroll
        for (int f = 0; f < F; ++f) {
            const int n = n0 - f;
            if (n < 0) break;
            const ST e = (ST)ecur[f] * (ST)SC;
            // ---- phase 1: zeta^[n][c] (single pass against the bound r) and length counts ---------
            // window relative to r_n = e_n + rho_n; large emissions cancel before they meet O(1) numbers
            // (see dp_forward_kernel): shift of the old slots = (e_{n+1} - rho_n) + (rho_{n+1} - gm_{n+1}).
            const ST rho = smax(zprev - (ST)gm_next + (ST)maxstep, eta + (ST)ln_first);
            const float eo = (float)((eprev - rho) + (rref - (ST)gm_next));
            float carry = 0.0f;
            if (S > 1) carry = __shfl_up_sync(FULL, Bq[KR - 1], CPW);
#pragma unroll
            for (int i = KR - 1; i > 0; --i) Bq[i] = Bq[i - 1] + eo;
            Bq[0] = (j == 0) ? (float)(eta - rho) : carry + eo;
            rref = rho;
            eprev = e;
            const ST betan = (n == 0) ? (ST)init_c : bcur[f];
            // forward + backward exponent: three numbers of which two may be huge and cancel -> summed in
            // double.  Clamped: when it would overflow the class has (s <= TINY) and is redone below.
            const double fb2 = (double)betan + (double)e + (double)rho;
            const float coef0 = valid ? w * ex2(fminf((float)fb2, 100.0f)) : 0.0f;
            float sp[2] = {0.0f, 0.0f};
#pragma unroll
            for (int i = 0; i < KR; ++i) {
                const float pr = ex2(Bq[i] + LN(i));
                sp[i & 1] += pr;
                El[i] = fmaf(pr, coef0, El[i]);
            }
            float s = slice_sum<S>(sp[0] + sp[1]);
            float mfix = 0.0f;
            float coef = coef0;
            const bool bad = valid && !(s > TINY);
            if (__any_sync(FULL, bad)) {
                float m = Bq[0] + LN(0);
#pragma unroll
                for (int i = 1; i < KR; ++i) m = fmaxf(m, Bq[i] + LN(i));
                m = slice_max<S>(m);
                const float coefx = valid ? w * ex2((float)(fb2 + (double)m)) : 0.0f;
                float s2p = 0.0f;
#pragma unroll
                for (int i = 0; i < KR; ++i) {
                    const float v = Bq[i] + LN(i);
                    const float px = ex2(v - m);
                    s2p += px;
                    if (bad) El[i] += px * coefx - ex2(v) * coef0;  // replace the underflowed contribution
                }
                s2p = slice_sum<S>(s2p);
                if (bad) {
                    s = s2p;
                    mfix = m;
                    coef = coefx;
                }
            }
            const ST zeta = valid ? (e + rho) + (ST)(mfix + lg2(s)) : (ST)NEG;
            const float Sc = coef * s;
            zprev = zeta;
            // ---- occupancy of frame n ---------------------------------------------------------
            {
                const float y = (Fprev - Sprev) - comp;
                const float tsum = occ + y;
                comp = (tsum - occ) - y;
                occ = tsum;
            }
            if (j == 0 && c < ldc) dem[(size_t)n * ldc + c] = valid ? occ : 0.0f;
            Sprev = Sc;
            if (n == 0) {
                // P(first segment has class c) is a distribution over c: normalise it by its own sum rather than by
                // the forward pass's log Z (after T frames the float roundings of the two directions have drifted
                // apart, and frame 0 is where the whole difference shows; see dp_lin_backward_kernel)
                float tot;
                if (W == 1) {
                    tot = warp_sum(owner ? Sc : 0.0f);
                } else {
                    ST* ss0 = zet_s + (n & 1) * cpad;
                    group_sync(W, bar_id);
                    if (j == 0) ss0[wig * CPW + cl] = valid ? (ST)Sc : (ST)0;
                    group_sync(W, bar_id);
                    tot = 0.0f;
                    for (int cc = 0; cc < cpad; ++cc) tot += (float)ss0[cc];
                }
                if (owner && tot != 0.0f) atomicAdd(p.d_init + c, Sc * (w / tot));
                break;
            }
            const float gm_n = dcur[f];
            // ---- phase 2: eta~[n][c1] = (+)_c2 trans[c2,c1] + zeta^[n][c2] - gm_n; transition counts --
            ST* zs = zet_s + (n & 1) * cpad;
            if (use_smem) {
                if (j == 0) zs[wig * CPW + cl] = valid ? zeta : (ST)NEG;
                group_sync(W, bar_id);
            }
            const ST gam = gcur[f];
            if (TM == 2 && !dense_pass) {
                if constexpr (TM == 2) {
                    ST v[SPW];
                    ST m2 = NEG;
#pragma unroll
                    for (int q = 0; q < SPW; ++q) {
                        const ST zv = (W == 1) ? __shfl_sync(FULL, zeta, sidx[q]) : zs[sidx[q]];
                        v[q] = zv + (ST)sval[q];
                        m2 = smax(m2, v[q]);
                    }
                    const float coef2 = valid ? w * ex2((float)((gam + m2) - (ST)gm_n)) : 0.0f;
                    float s2 = 0.0f;
#pragma unroll
                    for (int q = 0; q < SPW; ++q) {
                        const float pq = ex2((float)(v[q] - m2));
                        s2 += pq;
                        Es[q] = fmaf(pq, coef2, Es[q]);
                    }
                    eta = valid ? (m2 - (ST)gm_n) + (ST)lg2(s2) : (ST)NEG;
                    Fprev = coef2 * s2;
                }
            } else {
                ST m2 = NEG;
                if constexpr (TM == 0) {
#pragma unroll
                    for (int i = 0; i < CRR; ++i) m2 = smax(m2, zs[j * CRR + i] + (ST)tr[i]);
                } else if constexpr (TM == 1) {
                    for (int c2 = j; c2 < C; c2 += S) m2 = smax(m2, zs[c2] + (ST)trans_s[c2 * ldT + c]);
                } else {
                    for (int c2 = j; c2 < C; c2 += S)
                        m2 = smax(m2, zs[c2] + (ST)(valid ? __ldg(p.trans + (size_t)c2 * C + c) * SC : NEG));
                }
                m2 = slice_max<S>(m2);
                const float coef2 = valid ? w * ex2((float)((gam + m2) - (ST)gm_n)) : 0.0f;
                float s2 = 0.0f;
                if constexpr (TM == 0) {
#pragma unroll
                    for (int i = 0; i < CRR; ++i) {
                        const float pq = ex2((float)(zs[j * CRR + i] + (ST)tr[i] - m2));
                        s2 += pq;
                        Et[i] = fmaf(pq, coef2, Et[i]);
                    }
                } else if constexpr (TM == 1) {
                    for (int c2 = j; c2 < C; c2 += S) {
                        const float pq = ex2((float)(zs[c2] + (ST)trans_s[c2 * ldT + c] - m2));
                        s2 += pq;
                        if (valid) etr_s[c2 * ldT + c] += pq * coef2;
                    }
                } else {
                    for (int c2 = j; c2 < C; c2 += S) {
                        const float pq =
                            ex2((float)(zs[c2] + (ST)(valid ? __ldg(p.trans + (size_t)c2 * C + c) * SC : NEG) - m2));
                        s2 += pq;
                        if (valid) etr_s[c2 * ldT + c] += pq * coef2;
                    }
                }
                s2 = slice_sum<S>(s2);
                eta = valid ? (m2 - (ST)gm_n) + (ST)lg2(s2) : (ST)NEG;
                Fprev = coef2 * s2;
            }
            gm_next = gm_n;
        }
    }
    // ---- flush the per-video counts ------------------------------------------------------------
    if (valid) {
#pragma unroll
        for (int i = 0; i < KR; ++i) {
            const int k = j * KR + i + 1;
            if (k <= L) atomicAdd(p.d_len + (size_t)k * C + c, El[i]);
        }
        if constexpr (TM == 0) {
#pragma unroll
            for (int i = 0; i < CRR; ++i) {
                const int c2 = j * CRR + i;
                if (c2 < C) atomicAdd(p.d_trans + (size_t)c2 * C + c, Et[i]);
            }
        } else {
            if (TM == 2 && !dense_pass) {
                if constexpr (TM == 2) {
                    if (j == 0) {
#pragma unroll
                        for (int q = 0; q < SPW; ++q)
                            if (p.trans_succ[c * SPW + q] >= 0) atomicAdd(p.d_trans + (size_t)sidx[q] * C + c, Es[q]);
                    }
                }
            } else {
                for (int c2 = j; c2 < C; c2 += S) atomicAdd(p.d_trans + (size_t)c2 * C + c, etr_s[c2 * ldT + c]);
            }
        }
    }
#undef LN
}

template <bool XP, int KR, int S, int TM, bool LREG, int MAXT>
__global__ void __launch_bounds__(MAXT) dp_backward_kernel(const DpParams p) {
    dp_backward_kernel_body<XP, KR, S, TM, LREG, MAXT>(p, blockIdx.x);
}
template <bool XP, int KR, int S, int TM, bool LREG, int MAXT>
__global__ void __launch_bounds__(MAXT) dp_backward_kernel_grouped(const __grid_constant__ DpGroup g) {
    int local;
    const int t = group_find(g, blockIdx.x, local);
    dp_backward_kernel_body<XP, KR, S, TM, LREG, MAXT>(g.t[t], local);
}

// ---------------------------------------------------------------------------------------------
// host-side variant table
// ---------------------------------------------------------------------------------------------
struct RegVariant {
    int KR, S;
    bool lreg;  // duration scores in registers (else per-thread shared-memory column)
};
// capacity L <= KR*S, classes per warp 32/S
static const RegVariant kVariants[] = {
    {10, 2, true}, {20, 1, true}, {13, 4, true}, {25, 2, true}, {25, 4, true}, {25, 8, true}, {32, 1, true},
    {50, 4, false}, {50, 8, false}, {63, 8, false},
    {50, 2, true},  // C <= 16 with windows of 51..100 frames in ONE warp (the S6 shape: C = 11, K = 100)
};
constexpr int kNumVariants = sizeof(kVariants) / sizeof(kVariants[0]);
constexpr int kMaxThreadsSmall1 = 128, kMaxThreadsSmall = 512;
// long windows keep KR partial scores per thread: fewer threads per CTA so that they stay in registers
constexpr int max_threads_big(int KR, int mode) { return mode == 2 ? (KR > 50 ? 256 : 384) : (KR > 50 ? 384 : 640); }

struct RegChoice {
    int v;    // variant index, -1 = none
    int W;    // warps per video
    int VPB;  // videos per block
    int tm;   // transition mode
    size_t smem;
};

static size_t smem_bytes(const RegVariant& rv, int C, int W, int vpb, int tm, int mode, bool xp) {
    const int cpw = 32 / rv.S, cpad = W * cpw, ldT = ld_trans(W, rv.S), G = W * 32;
    const int stw = xp ? 2 : 1;  // state words (double / float)
    size_t fl = 0;
    if (tm == 1) fl += (size_t)C * ldT;
    if (!rv.lreg) fl += (size_t)rv.KR * G;
    if (xp) fl += 1;  // alignment pad of the double rows
    if (mode == 2)
        fl += (size_t)vpb * (2 * cpad * stw + (tm != 0 ? C * ldT + 1 : 0));
    else
        fl += (size_t)vpb * (2 * cpad * stw + 2 * W);
    return fl * sizeof(float);
}

static RegChoice choose(int C, int L, int mode, bool sparse, bool xp = false) {
    RegChoice best{-1, 0, 0, 0, 0};
    double best_cost = 1e30;
    for (int v = 0; v < kNumVariants; ++v) {
        const RegVariant& rv = kVariants[v];
        if (rv.KR * rv.S < L) continue;
        const int cpw = 32 / rv.S;
        const int W = (C + cpw - 1) / cpw;
        if (v == 10 && W != 1) continue;  // 50 window registers per lane: the multi-warp instantiations (512 threads) would spill
        const bool small1 = rv.lreg && W == 1 && rv.S <= 4;  // S = 8: 4 classes x 8 slices, row split would overrun
        const int tm = sparse ? 2 : (small1 ? 0 : 1);
        const int maxt = rv.lreg ? (small1 ? kMaxThreadsSmall1 : kMaxThreadsSmall) : max_threads_big(rv.KR, mode);
        if (W * 32 > maxt) continue;
        int vpb = small1 ? 4 : maxt / (W * 32);
        if (vpb > 8) vpb = 8;
        if (!rv.lreg) vpb = 1;
        while (vpb > 1 && smem_bytes(rv, C, W, vpb, tm, mode, xp) > kSmemCap) --vpb;
        const size_t sm = smem_bytes(rv, C, W, vpb, tm, mode, xp);
        if (sm > kSmemCap) continue;
        // issue slots per frame ~ warps * (window + transitions per lane) (+ barrier cost when W > 1)
        const int crr = tm == 2 ? SPW : (tm == 0 ? (cpw + rv.S - 1) / rv.S : (C + rv.S - 1) / rv.S);
        const double cost = (double)W * (rv.KR * (rv.lreg ? 1.0 : 1.3) + crr) + (W > 1 ? 6.0 * W : 0.0);
        if (cost < best_cost) {
            best_cost = cost;
            best = RegChoice{v, W, vpb, tm, sm};
        }
    }
    return best;
}

// MODE: 0 Viterbi, 1 log-semiring forward, 2 backward.  XP: extended-precision per-class state.
static int dp_num_sms() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
            n = 148;
    }
    return n;
}

template <int MODE, bool XP, int KR, int S, int TM, bool LREG, int MAXT>
static int launch_one(const DpParams& p0, const RegChoice& ch, cudaStream_t st) {
    // videos per CTA: as many as the thread budget allows for big batches, fewer when that would leave SMs without a
    // CTA (512 videos at 8 per CTA are 64 CTAs on a 148-SM part: configs[0])
    DpParams p = p0;
    int vpb = ch.VPB;
    const int want = dp_num_sms();  // one CTA per SM is enough: concurrent calls on other streams fill the rest
    while (vpb > 1 && (p.B + vpb - 1) / vpb < want) --vpb;
    p.VPB = vpb;
    const RegVariant rv{KR, S, LREG};
    const size_t smem = (vpb == ch.VPB) ? ch.smem : smem_bytes(rv, p.C, ch.W, vpb, TM, MODE, XP);
    const int blocks = (p.B + vpb - 1) / vpb;
    const int threads = vpb * ch.W * 32;
    cudaError_t e = cudaSuccess;
    if constexpr (MODE == 0) {
        auto k = dp_forward_kernel<true, false, KR, S, TM, LREG, MAXT>;
        if (smem > 48 * 1024) e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) k<<<blocks, threads, smem, st>>>(p);
    } else if constexpr (MODE == 1) {
        auto k = dp_forward_kernel<false, XP, KR, S, TM, LREG, MAXT>;
        if (smem > 48 * 1024) e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) k<<<blocks, threads, smem, st>>>(p);
    } else {
        auto k = dp_backward_kernel<XP, KR, S, TM, LREG, MAXT>;
        if (smem > 48 * 1024) e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) k<<<blocks, threads, smem, st>>>(p);
    }
    if (e != cudaSuccess) {
        set_error("cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
        return -3;
    }
    return check_launch("dp_reg kernel");
}

constexpr int kMaxThreadsMid = 384;  // backward, W * VPB <= 12 warps: 170 registers instead of 128 (no spills in the frame loop)
template <int MODE, bool XP, int KR, int S>
static int launch_small(const DpParams& p, const RegChoice& ch, cudaStream_t st) {
    if constexpr (MODE == 2 && !XP) {
        if (ch.W > 1 && ch.W * ch.VPB * 32 <= kMaxThreadsMid && ch.W * ch.VPB * 32 > 256) {
            return ch.tm == 2 ? launch_one<MODE, XP, KR, S, 2, true, kMaxThreadsMid>(p, ch, st)
                              : launch_one<MODE, XP, KR, S, 1, true, kMaxThreadsMid>(p, ch, st);
        }
    }
    if (ch.tm == 2) {
        return ch.W == 1 ? launch_one<MODE, XP, KR, S, 2, true, kMaxThreadsSmall1>(p, ch, st)
                         : launch_one<MODE, XP, KR, S, 2, true, kMaxThreadsSmall>(p, ch, st);
    }
    return ch.tm == 0 ? launch_one<MODE, XP, KR, S, 0, true, kMaxThreadsSmall1>(p, ch, st)
                      : launch_one<MODE, XP, KR, S, 1, true, kMaxThreadsSmall>(p, ch, st);
}
template <int MODE, bool XP, int KR, int S>
static int launch_big(const DpParams& p, const RegChoice& ch, cudaStream_t st) {
    return ch.tm == 2 ? launch_one<MODE, XP, KR, S, 2, false, max_threads_big(KR, MODE)>(p, ch, st)
                      : launch_one<MODE, XP, KR, S, 1, false, max_threads_big(KR, MODE)>(p, ch, st);
}

template <int MODE, bool XP>
static int launch_mode(const DpParams& p, const RegChoice& ch, cudaStream_t st) {
    switch (ch.v) {
        case 0: return launch_small<MODE, XP, 10, 2>(p, ch, st);
        case 1: return launch_small<MODE, XP, 20, 1>(p, ch, st);
        case 2: return launch_small<MODE, XP, 13, 4>(p, ch, st);
        case 3: return launch_small<MODE, XP, 25, 2>(p, ch, st);
        case 4: return launch_small<MODE, XP, 25, 4>(p, ch, st);
        case 5: return launch_small<MODE, XP, 25, 8>(p, ch, st);
        case 6: return launch_small<MODE, XP, 32, 1>(p, ch, st);
        case 7: return launch_big<MODE, XP, 50, 4>(p, ch, st);
        case 8: return launch_big<MODE, XP, 50, 8>(p, ch, st);
        case 9: return launch_big<MODE, XP, 63, 8>(p, ch, st);
        case 10: return launch_small<MODE, XP, 50, 2>(p, ch, st);
    }
    set_error("no register-resident DP variant for this shape");
    return -2;
}

// one translation unit per (mode, precision) so that the variants compile in parallel
#define HSMM_DP_DEFINE_LAUNCHER(NAME, MODE, XP)                                   \
    int NAME(DpParams p, cudaStream_t st) {                                       \
        const bool sparse = (MODE == 2) ? (p.trans_succ != nullptr) : (p.trans_pred != nullptr); \
        RegChoice ch = choose(p.C, p.L, MODE, sparse, XP);                        \
        if (ch.v < 0) {                                                           \
            set_error("shape C=%d L=%d not supported by the register-resident DP", p.C, p.L); \
            return -2;                                                            \
        }                                                                         \
        p.W = ch.W;                                                               \
        p.VPB = ch.VPB;                                                           \
        return launch_mode<MODE, XP>(p, ch, st);                                  \
    }

}  // namespace hsmm
