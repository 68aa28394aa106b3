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
#include "hsmm_common.cuh"

namespace hsmm {

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
template <bool VIT, int KR, int S, int TM, bool LREG, int MAXT>
__global__ void __launch_bounds__(MAXT) dp_forward_kernel(const DpParams p) {
    constexpr int CPW = Lay<S>::CPW;
    constexpr int CRR = Lay<S>::CRR;
    extern __shared__ float smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int W = p.W, C = p.C, L = p.L, ldc = p.ldc, Tmax = p.Tmax;
    const int G = W * 32;
    const int slot = warp / W, wig = warp - slot * W;
    const int gtid = wig * 32 + lane;
    const int cpad = W * CPW;
    const int ldT = ld_trans(W, S);
    const float SC = VIT ? 1.0f : LOG2E;

    // shared layout: [transT C*ldT (TM==1)] [len columns KR*G (!LREG)] [per group: gamma 2*cpad, warp maxima 2*W]
    float* transT = smem;
    float* lens = smem + (TM == 1 ? C * ldT : 0);
    float* gbase = lens + (LREG ? 0 : KR * G);
    const int per_group = 2 * cpad + 2 * W;
    float* gam_s = gbase + slot * per_group;
    float* wmax_s = gam_s + 2 * cpad;

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

    const int vidx = blockIdx.x * p.VPB + slot;
    if (vidx >= p.B) return;
    const int b = p.order ? p.order[vidx] : vidx;
    const int T = p.lengths[b];
    const int bar_id = 1 + slot;

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
    if (!VIT && owner) p.fbeta[row0 * ldc + c] = init_c;

    bool dense_pass = (TM != 2);  // TM == 2: first pass sparse, second (rare) pass dense from global memory
    float final_v = 0.0f;         // VIT: best score; FWD: log2 Z; both relative to nu_T
    int final_c = 0;
    double nu = 0.0;

#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
        float A[KR];
#pragma unroll
        for (int i = 0; i < KR; ++i) A[i] = NEG;
        float beta = init_c;  // beta^[n-1][c], relative to nu_n
        float gprev = NEG;    // gamma~[n-1][c], relative to nu_{n-1}
        float rref = 0.0f;    // FWD: reference the window is stored against
        float gmprev = 0.0f;  // gm_{n-1}
        nu = 0.0;
        const bool use_smem = (TM != 2) || dense_pass || W > 1;

        float enext[F];
#pragma unroll
        for (int f = 0; f < F; ++f) enext[f] = (valid && f < T) ? __ldg(em_b + (size_t)f * ldc + c) : 0.0f;

        for (int n0 = 1; n0 <= T; n0 += F) {
            float ecur[F];
#pragma unroll
            for (int f = 0; f < F; ++f) ecur[f] = enext[f];
#pragma unroll
            for (int f = 0; f < F; ++f) {
                const int t = n0 - 1 + F + f;
                enext[f] = (valid && t < T) ? __ldg(em_b + (size_t)t * ldc + c) : 0.0f;
            }
#pragma unroll
            for (int f = 0; f < F; ++f) {
                const int n = n0 + f;
                if (n > T) break;
                const float e = ecur[f] * SC;
                nu += (double)gmprev;
                float gamma;
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
                    // every term <= rnew; the window is stored relative to that reference
                    const float rnew = fmaxf(gprev - gmprev + e + maxstep, beta + e + ln_first);
                    const float eo = e - gmprev + (rref - rnew);
                    float carry = 0.0f;
                    if (S > 1) carry = __shfl_up_sync(FULL, A[KR - 1], CPW);
#pragma unroll
                    for (int i = KR - 1; i > 0; --i) A[i] = A[i - 1] + eo;
                    A[0] = (j == 0) ? (beta + e) - rnew : carry + eo;
                    rref = rnew;
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
                    gamma = valid ? rnew + mfix + lg2(s) : NEG;
                }
                gprev = gamma;
                // ---- group maximum of gamma: the normaliser increment ----------------------------
                float gm = warp_max(owner ? gamma : NEG);
                float* gs = gam_s + (n & 1) * cpad;
                if (W > 1 && lane == 0) wmax_s[(n & 1) * W + wig] = gm;
                if (use_smem) {
                    if (j == 0) gs[wig * CPW + cl] = valid ? gamma : NEG;
                    group_sync(W, bar_id);
                }
                if (W > 1) {
                    gm = NEG;
                    for (int q = 0; q < W; ++q) gm = fmaxf(gm, wmax_s[(n & 1) * W + q]);
                }
                if (!VIT && owner) p.fgamma[(row0 + n) * ldc + c] = gamma;
                if (n == T) {
                    if (VIT && owner) p.bp[(row0 + n) * ldc + c] = (uint32_t)bk << 16;
                    break;
                }
                if (!VIT && gtid == 0) p.fdelta[row0 + n] = gm;
                gmprev = gm;
                // ---- phase 2: beta^[n][c2] = (+)_c1 gamma~[n][c1] + trans[c2,c1] - gm ---------------
                if constexpr (VIT) {
                    float best = NEG;
                    int bc = 0;
                    if (TM == 2 && !dense_pass) {
                        if constexpr (TM == 2) {
#pragma unroll
                            for (int q = 0; q < SPW; ++q) {
                                const float gv = (W == 1) ? __shfl_sync(FULL, gamma, pidx[q]) : gs[pidx[q]];
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
                            const float v = gs[j * CRR + i] + tr[i];  // padding: gs = NEG, tr = NEG
                            if (v > best || i == 0) {
                                best = v;
                                bc = j * CRR + i;
                            }
                        }
                        slice_argmax<S>(best, bc);
                    } else if constexpr (TM == 1) {
                        bc = j;
                        for (int c1 = j; c1 < C; c1 += S) {
                            const float v = gs[c1] + transT[c1 * ldT + c];
                            if (v > best || c1 == j) {
                                best = v;
                                bc = c1;
                            }
                        }
                        slice_argmax<S>(best, bc);
                    } else {  // TM == 2, dense fallback straight from global memory (rare)
                        bc = j;
                        for (int c1 = j; c1 < C; c1 += S) {
                            const float v = gs[c1] + (valid ? __ldg(p.trans + (size_t)c * C + c1) : NEG);
                            if (v > best || c1 == j) {
                                best = v;
                                bc = c1;
                            }
                        }
                        slice_argmax<S>(best, bc);
                    }
                    beta = valid ? best - gm : NEG;
                    if (owner) p.bp[(row0 + n) * ldc + c] = ((uint32_t)bk << 16) | (uint32_t)bc;
                } else {
                    if (TM == 2 && !dense_pass) {
                        if constexpr (TM == 2) {
                            float v[SPW];
                            float m = NEG;
#pragma unroll
                            for (int q = 0; q < SPW; ++q) {
                                const float gv = (W == 1) ? __shfl_sync(FULL, gamma, pidx[q]) : gs[pidx[q]];
                                v[q] = gv + pval[q];
                                m = fmaxf(m, v[q]);
                            }
                            float s = 0.0f;
#pragma unroll
                            for (int q = 0; q < SPW; ++q) s += ex2(v[q] - m);
                            beta = valid ? (m - gm) + lg2(s) : NEG;
                        }
                    } else {
                        // single pass against the bound gm + trmax; exact two-pass when it underflows
                        float s = 0.0f;
                        if constexpr (TM == 0) {
                            float sp[2] = {0.0f, 0.0f};
#pragma unroll
                            for (int i = 0; i < CRR; ++i) sp[i & 1] += ex2((gs[j * CRR + i] - gm) + tr[i]);
                            s = sp[0] + sp[1];
                        } else if constexpr (TM == 1) {
                            const float off = gm + trmax;
                            for (int c1 = j; c1 < C; c1 += S) s += ex2(gs[c1] + transT[c1 * ldT + c] - off);
                        }
                        s = slice_sum<S>(s);
                        float mfix = 0.0f;
                        const bool bad = (TM == 2) || (valid && !(s > TINY));
                        if (__any_sync(FULL, bad)) {
                            float m = NEG;
                            for (int c1 = j; c1 < C; c1 += S)
                                m = fmaxf(m, gs[c1] + (valid ? __ldg(p.trans + (size_t)c * C + c1) * SC : NEG));
                            m = slice_max<S>(m);
                            float s2p = 0.0f;
                            for (int c1 = j; c1 < C; c1 += S)
                                s2p += ex2(gs[c1] + (valid ? __ldg(p.trans + (size_t)c * C + c1) * SC : NEG) - m);
                            s2p = slice_sum<S>(s2p);
                            if (bad) {
                                s = s2p;
                                mfix = m - gm - trmax;
                            }
                        }
                        beta = valid ? trmax + mfix + lg2(s) : NEG;
                    }
                    if (owner) p.fbeta[(row0 + n) * ldc + c] = beta;
                }
            }
        }

        // ---- termination -------------------------------------------------------------------
        // gamma~[T] sits in gam_s (dense / multi-warp) or only in registers (sparse, one warp): put it
        // in shared memory in every case so that one code path finishes the video.
        float* gT = gam_s + (T & 1) * cpad;
        if (!use_smem) {
            if (j == 0) gT[wig * CPW + cl] = valid ? gprev : NEG;
            group_sync(W, bar_id);
        }
        if constexpr (!VIT) {
            float m = NEG;
            for (int cc = lane; cc < C; cc += 32) m = fmaxf(m, gT[cc] + (endb ? endb[cc] * SC : 0.0f));
            m = warp_max(m);
            float s = 0.0f;
            for (int cc = lane; cc < C; cc += 32) s += ex2(gT[cc] + (endb ? endb[cc] * SC : 0.0f) - m);
            s = warp_sum(s);
            final_v = m + lg2(s);
        } else {
            float best = NEG;
            int bc = 0x7fffffff;
            for (int cc = lane; cc < C; cc += 32) {
                const float v = gT[cc] + (endb ? endb[cc] : 0.0f);
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
            p.logz2[b] = final_v;
            p.fflag[b] = (TM == 2 && dense_pass) ? 1.0f : 0.0f;
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

// ---------------------------------------------------------------------------------------------
// backward: expected counts
// ---------------------------------------------------------------------------------------------
template <int KR, int S, int TM, bool LREG, int MAXT>
__global__ void __launch_bounds__(MAXT) dp_backward_kernel(const DpParams p) {
    constexpr int CPW = Lay<S>::CPW;
    constexpr int CRR = Lay<S>::CRR;
    extern __shared__ float smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int W = p.W, C = p.C, L = p.L, ldc = p.ldc, Tmax = p.Tmax;
    const int G = W * 32;
    const int slot = warp / W, wig = warp - slot * W;
    const int gtid = wig * 32 + lane;
    const int cpad = W * CPW;
    const int ldT = ld_trans(W, S);
    const float SC = LOG2E;

    // shared: [trans C*ldT (TM==1)] [len columns KR*G (!LREG)] [per group: zeta 2*cpad, Etr C*ldT (TM != 0)]
    float* trans_s = smem;
    float* lens = smem + (TM == 1 ? C * ldT : 0);
    float* gbase = lens + (LREG ? 0 : KR * G);
    const int per_group = 2 * cpad + (TM == 0 ? 0 : C * ldT);
    float* zet_s = gbase + slot * per_group;
    float* etr_s = zet_s + 2 * cpad;  // TM == 1 always, TM == 2 only in the dense fallback

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
            float* e = gbase + g * per_group + 2 * cpad;
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

    const int vidx = blockIdx.x * p.VPB + slot;
    if (vidx >= p.B) return;
    const int b = p.order ? p.order[vidx] : vidx;
    const int T = p.lengths[b];
    const int bar_id = 1 + slot;
    const bool dense_pass = (TM != 2) || (p.fflag[b] != 0.0f);
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
    const float lzrel = p.logz2[b];  // log2 Z relative to nu_T
    const float w = p.grad[b];
    const float init_c = valid ? p.init[c] * SC : NEG;

    const float* em_b = p.em + (size_t)b * Tmax * ldc;
    const size_t row0 = (size_t)b * (Tmax + 1);
    const float* fb = p.fbeta + row0 * ldc;
    const float* fg = p.fgamma + row0 * ldc;
    const float* fd = p.fdelta + row0;
    float* dem = p.d_em + (size_t)b * Tmax * ldc;

    // Frame n: zeta^[n] = zeta[n] - mu_{n+1}, eta~[n] = eta[n] - mu_n with mu_n = logZ - nu_n, so that
    // posteriors are exp(forward + backward) of O(1) numbers:
    //   S[n,c] = exp(beta^[n][c] + zeta^[n][c]),  F[n,c] = exp(gamma~[n][c] + eta~[n][c]).
    float eta = valid ? endc - lzrel : NEG;  // eta~[T]
    float zprev = NEG;                       // zeta^[n+1][c]
    float rref = 0.0f;
    float occ = 0.0f, comp = 0.0f;  // Kahan-compensated occupancy
    float Fprev = valid ? w * ex2(__ldg(fg + (size_t)T * ldc + c) + endc - lzrel) : 0.0f;
    float Sprev = 0.0f;
    float gm_next = 0.0f;  // gm_{n+1}

    for (int i = T * ldc + gtid; i < Tmax * ldc; i += G) dem[i] = 0.0f;  // frames beyond the video

    float enext[F], bnext[F], gnext[F], dnext[F];
#pragma unroll
    for (int f = 0; f < F; ++f) {
        const int n = T - 1 - f;
        const bool ok = valid && n >= 0;
        dnext[f] = (n >= 1) ? __ldg(fd + n) : 0.0f;  // gm_n
        enext[f] = ok ? __ldg(em_b + (size_t)n * ldc + c) : 0.0f;
        bnext[f] = (ok && n > 0) ? __ldg(fb + (size_t)n * ldc + c) : 0.0f;
        gnext[f] = (ok && n > 0) ? __ldg(fg + (size_t)n * ldc + c) : 0.0f;
    }
    for (int n0 = T - 1; n0 >= 0; n0 -= F) {
        float ecur[F], bcur[F], gcur[F], dcur[F];
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
            bnext[f] = (ok && n > 0) ? __ldg(fb + (size_t)n * ldc + c) : 0.0f;
            gnext[f] = (ok && n > 0) ? __ldg(fg + (size_t)n * ldc + c) : 0.0f;
        }
#pragma unroll
        for (int f = 0; f < F; ++f) {
            const int n = n0 - f;
            if (n < 0) break;
            const float e = ecur[f] * SC;
            // ---- phase 1: zeta^[n][c] (single pass against the bound r) and length counts ---------
            const float rnew = fmaxf(zprev - gm_next + e + maxstep, eta + e + ln_first);
            const float eo = e - gm_next + (rref - rnew);
            float carry = 0.0f;
            if (S > 1) carry = __shfl_up_sync(FULL, Bq[KR - 1], CPW);
#pragma unroll
            for (int i = KR - 1; i > 0; --i) Bq[i] = Bq[i - 1] + eo;
            Bq[0] = (j == 0) ? (eta + e) - rnew : carry + eo;
            rref = rnew;
            const float betan = (n == 0) ? init_c : bcur[f];
            // exponent clamped: when it would overflow the class has (s <= TINY) and is redone below
            const float coef0 = valid ? w * ex2(fminf(betan + rnew, 100.0f)) : 0.0f;
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
                const float coefx = valid ? w * ex2(betan + rnew + m) : 0.0f;
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
            const float zeta = valid ? rnew + mfix + lg2(s) : NEG;
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
                if (owner) atomicAdd(p.d_init + c, Sc);
                break;
            }
            const float gm_n = dcur[f];
            // ---- phase 2: eta~[n][c1] = (+)_c2 trans[c2,c1] + zeta^[n][c2] - gm_n; transition counts --
            float* zs = zet_s + (n & 1) * cpad;
            if (use_smem) {
                if (j == 0) zs[wig * CPW + cl] = valid ? zeta : NEG;
                group_sync(W, bar_id);
            }
            const float gam = gcur[f];
            if (TM == 2 && !dense_pass) {
                if constexpr (TM == 2) {
                    float v[SPW];
                    float m2 = NEG;
#pragma unroll
                    for (int q = 0; q < SPW; ++q) {
                        const float zv = (W == 1) ? __shfl_sync(FULL, zeta, sidx[q]) : zs[sidx[q]];
                        v[q] = zv + sval[q];
                        m2 = fmaxf(m2, v[q]);
                    }
                    const float coef2 = valid ? w * ex2(gam + m2 - gm_n) : 0.0f;
                    float s2 = 0.0f;
#pragma unroll
                    for (int q = 0; q < SPW; ++q) {
                        const float pq = ex2(v[q] - m2);
                        s2 += pq;
                        Es[q] = fmaf(pq, coef2, Es[q]);
                    }
                    eta = valid ? (m2 - gm_n) + lg2(s2) : NEG;
                    Fprev = coef2 * s2;
                }
            } else {
                float m2 = NEG;
                if constexpr (TM == 0) {
#pragma unroll
                    for (int i = 0; i < CRR; ++i) m2 = fmaxf(m2, zs[j * CRR + i] + tr[i]);
                } else if constexpr (TM == 1) {
                    for (int c2 = j; c2 < C; c2 += S) m2 = fmaxf(m2, zs[c2] + trans_s[c2 * ldT + c]);
                } else {
                    for (int c2 = j; c2 < C; c2 += S)
                        m2 = fmaxf(m2, zs[c2] + (valid ? __ldg(p.trans + (size_t)c2 * C + c) * SC : NEG));
                }
                m2 = slice_max<S>(m2);
                const float coef2 = valid ? w * ex2(gam + m2 - gm_n) : 0.0f;
                float s2 = 0.0f;
                if constexpr (TM == 0) {
#pragma unroll
                    for (int i = 0; i < CRR; ++i) {
                        const float pq = ex2(zs[j * CRR + i] + tr[i] - m2);
                        s2 += pq;
                        Et[i] = fmaf(pq, coef2, Et[i]);
                    }
                } else if constexpr (TM == 1) {
                    for (int c2 = j; c2 < C; c2 += S) {
                        const float pq = ex2(zs[c2] + trans_s[c2 * ldT + c] - m2);
                        s2 += pq;
                        if (valid) etr_s[c2 * ldT + c] += pq * coef2;
                    }
                } else {
                    for (int c2 = j; c2 < C; c2 += S) {
                        const float pq = ex2(zs[c2] + (valid ? __ldg(p.trans + (size_t)c2 * C + c) * SC : NEG) - m2);
                        s2 += pq;
                        if (valid) etr_s[c2 * ldT + c] += pq * coef2;
                    }
                }
                s2 = slice_sum<S>(s2);
                eta = valid ? (m2 - gm_n) + lg2(s2) : NEG;
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

static size_t smem_bytes(const RegVariant& rv, int C, int W, int vpb, int tm, int mode) {
    const int cpw = 32 / rv.S, cpad = W * cpw, ldT = ld_trans(W, rv.S), G = W * 32;
    size_t fl = 0;
    if (tm == 1) fl += (size_t)C * ldT;
    if (!rv.lreg) fl += (size_t)rv.KR * G;
    if (mode == 2)
        fl += (size_t)vpb * (2 * cpad + (tm != 0 ? C * ldT : 0));
    else
        fl += (size_t)vpb * (2 * cpad + 2 * W);
    return fl * sizeof(float);
}

static RegChoice choose(int C, int L, int mode, bool sparse) {
    RegChoice best{-1, 0, 0, 0, 0};
    double best_cost = 1e30;
    for (int v = 0; v < kNumVariants; ++v) {
        const RegVariant& rv = kVariants[v];
        if (rv.KR * rv.S < L) continue;
        const int cpw = 32 / rv.S;
        const int W = (C + cpw - 1) / cpw;
        const bool small1 = rv.lreg && W == 1 && rv.S <= 4;  // S = 8: 4 classes x 8 slices, row split would overrun
        const int tm = sparse ? 2 : (small1 ? 0 : 1);
        const int maxt = rv.lreg ? (small1 ? kMaxThreadsSmall1 : kMaxThreadsSmall) : max_threads_big(rv.KR, mode);
        if (W * 32 > maxt) continue;
        int vpb = small1 ? 4 : maxt / (W * 32);
        if (vpb > 8) vpb = 8;
        if (!rv.lreg) vpb = 1;
        while (vpb > 1 && smem_bytes(rv, C, W, vpb, tm, mode) > kSmemCap) --vpb;
        const size_t sm = smem_bytes(rv, C, W, vpb, tm, mode);
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

template <int MODE, int KR, int S, int TM, bool LREG, int MAXT>
static int launch_one(const DpParams& p, const RegChoice& ch, cudaStream_t st) {
    const int blocks = (p.B + ch.VPB - 1) / ch.VPB;
    const int threads = ch.VPB * ch.W * 32;
    cudaError_t e = cudaSuccess;
    if constexpr (MODE == 0) {
        auto k = dp_forward_kernel<true, KR, S, TM, LREG, MAXT>;
        if (ch.smem > 48 * 1024) e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ch.smem);
        if (e == cudaSuccess) k<<<blocks, threads, ch.smem, st>>>(p);
    } else if constexpr (MODE == 1) {
        auto k = dp_forward_kernel<false, KR, S, TM, LREG, MAXT>;
        if (ch.smem > 48 * 1024) e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ch.smem);
        if (e == cudaSuccess) k<<<blocks, threads, ch.smem, st>>>(p);
    } else {
        auto k = dp_backward_kernel<KR, S, TM, LREG, MAXT>;
        if (ch.smem > 48 * 1024) e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ch.smem);
        if (e == cudaSuccess) k<<<blocks, threads, ch.smem, st>>>(p);
    }
    if (e != cudaSuccess) {
        set_error("cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
        return -3;
    }
    return check_launch("dp_reg kernel");
}

template <int MODE, int KR, int S>
static int launch_small(const DpParams& p, const RegChoice& ch, cudaStream_t st) {
    if (ch.tm == 2) {
        return ch.W == 1 ? launch_one<MODE, KR, S, 2, true, kMaxThreadsSmall1>(p, ch, st)
                         : launch_one<MODE, KR, S, 2, true, kMaxThreadsSmall>(p, ch, st);
    }
    return ch.tm == 0 ? launch_one<MODE, KR, S, 0, true, kMaxThreadsSmall1>(p, ch, st)
                      : launch_one<MODE, KR, S, 1, true, kMaxThreadsSmall>(p, ch, st);
}
template <int MODE, int KR, int S>
static int launch_big(const DpParams& p, const RegChoice& ch, cudaStream_t st) {
    return ch.tm == 2 ? launch_one<MODE, KR, S, 2, false, max_threads_big(KR, MODE)>(p, ch, st)
                      : launch_one<MODE, KR, S, 1, false, max_threads_big(KR, MODE)>(p, ch, st);
}

template <int MODE>
static int launch_mode(const DpParams& p, const RegChoice& ch, cudaStream_t st) {
    switch (ch.v) {
        case 0: return launch_small<MODE, 10, 2>(p, ch, st);
        case 1: return launch_small<MODE, 20, 1>(p, ch, st);
        case 2: return launch_small<MODE, 13, 4>(p, ch, st);
        case 3: return launch_small<MODE, 25, 2>(p, ch, st);
        case 4: return launch_small<MODE, 25, 4>(p, ch, st);
        case 5: return launch_small<MODE, 25, 8>(p, ch, st);
        case 6: return launch_small<MODE, 32, 1>(p, ch, st);
        case 7: return launch_big<MODE, 50, 4>(p, ch, st);
        case 8: return launch_big<MODE, 50, 8>(p, ch, st);
        case 9: return launch_big<MODE, 63, 8>(p, ch, st);
    }
    set_error("no register-resident DP variant for this shape");
    return -2;
}

// exported to hsmm_api.cu
bool dp_reg_supported(int C, int L, int mode, bool sparse) { return choose(C, L, mode, sparse).v >= 0; }

const char* dp_reg_name(int C, int L, int mode, bool sparse) {
    static thread_local char buf[112];
    RegChoice ch = choose(C, L, mode, sparse);
    if (ch.v < 0) return "none";
    static const char* tmn[] = {"trans-reg", "trans-smem", "trans-sparse"};
    snprintf(buf, sizeof(buf), "reg<KR=%d,S=%d>/%s/%s/W=%d/VPB=%d/smem=%zu", kVariants[ch.v].KR, kVariants[ch.v].S,
             tmn[ch.tm], kVariants[ch.v].lreg ? "len-reg" : "len-smem", ch.W, ch.VPB, ch.smem);
    return buf;
}

int dp_reg_launch(DpParams p, int mode, cudaStream_t st) {
    const bool sparse = (mode == 2) ? (p.trans_succ != nullptr) : (p.trans_pred != nullptr);
    RegChoice ch = choose(p.C, p.L, mode, sparse);
    if (ch.v < 0) {
        set_error("shape C=%d L=%d not supported by the register-resident DP", p.C, p.L);
        return -2;
    }
    p.W = ch.W;
    p.VPB = ch.VPB;
    if (mode == 0) return launch_mode<0>(p, ch, st);
    if (mode == 1) return launch_mode<1>(p, ch, st);
    return launch_mode<2>(p, ch, st);
}

}  // namespace hsmm
