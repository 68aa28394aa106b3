// Register-resident semi-Markov DP kernels (sm_100a).
//
// One group of W warps works on one video; a CTA carries VPB independent groups.  A lane owns one
// class c and one k-slice j of that class (S slices per class inside a warp, CPW = 32/S classes per
// warp).  The span window never touches memory: lane (c, j) keeps, in registers,
//
//     A[i] = beta[n-k][c] + em[n-k..n-1, c],   k = j*KR + i + 1,  i = 0..KR-1
//
// i.e. the score of "some prefix, then a class-c segment that started k frames ago", updated every
// frame by one shift-and-add (A[i] <- A[i-1] + em[n-1,c]); the element that crosses a slice boundary
// moves with one shuffle.  This is the direct window sum of the reference (no prefix-sum
// cancellation), one add per (frame, class, length).  The duration scores len[k,c] are static per
// register slot: held in registers (LREG) or, for long windows, in a per-thread shared-memory column.
//
//   phase 1: gamma[n][c]  = (+)_k  A_k + len[k,c]                       (registers + slice shuffles)
//   phase 2: beta[n][c2]  = (+)_c1 gamma[n][c1] + trans[c2,c1]          (gamma through shared memory;
//                                                                        trans in registers (TREG, one
//                                                                        warp per video) or shared)
//
// The (B,T,K,C,C) potentials of the reference (semimarkov_modules.py:416-523) are never formed.
// Max-plus (Viterbi, with back-pointers and in-kernel back-trace) and log-semiring (base-2 domain,
// ex2/lg2 on the MUFU pipe) share the code path.  The backward kernel mirrors the recursion in
// reverse time and accumulates the expected counts in registers.
#include "hsmm_common.cuh"

namespace hsmm {

constexpr int F = 4;  // frames per register prefetch chunk
constexpr size_t kSmemCap = 220 * 1024;

template <int S>
struct Lay {
    static constexpr int CPW = 32 / S;
    static constexpr int CRR = (CPW + S - 1) / S;  // transitions per lane when held in registers
};

__host__ __device__ inline int ld_trans(int W, int S) {
    const int cpw = 32 / S;
    const int cpad = W * cpw;
    return ((cpad + 31) / 32) * 32 + (S > 1 ? cpw : 0);
}

// ---------------------------------------------------------------------------------------------
// forward: Viterbi (VIT) or log-partition (FWD)
// ---------------------------------------------------------------------------------------------
template <bool VIT, int KR, int S, bool TREG, bool LREG, int MAXT>
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

    // shared layout: [transT C*ldT (!TREG)] [len columns KR*G (!LREG)] [per group: gamma 2*cpad]
    float* transT = smem;
    float* lens = smem + (TREG ? 0 : C * ldT);
    float* gbase = lens + (LREG ? 0 : KR * G);
    float* gam_s = gbase + slot * 2 * cpad;

    const int cl = lane % CPW, j = lane / CPW;
    const int c = wig * CPW + cl;
    const bool valid = c < C;
    const bool owner = valid && j == 0;

    if constexpr (!TREG) {
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
    if constexpr (!TREG || !LREG) __syncthreads();

    const int vidx = blockIdx.x * p.VPB + slot;
    if (vidx >= p.B) return;
    const int b = p.order ? p.order[vidx] : vidx;
    const int T = p.lengths[b];
    const int bar_id = 1 + slot;

    float A[KR];
    float ln[LREG ? KR : 1];
    const float* lnp = lens + gtid;
#pragma unroll
    for (int i = 0; i < KR; ++i) {
        A[i] = NEG;
        if constexpr (LREG) {
            const int k = j * KR + i + 1;
            ln[i] = (valid && k <= L) ? p.lenp[(size_t)k * C + c] * SC : NEG;
        }
    }
#define LN(i) (LREG ? ln[LREG ? (i) : 0] : lnp[(i) * G])
    float beta = valid ? p.init[c] * SC : NEG;
    float tr[TREG ? CRR : 1];
    if constexpr (TREG) {
#pragma unroll
        for (int i = 0; i < CRR; ++i) {
            const int c1 = i * S + j;
            tr[i] = (valid && c1 < C) ? p.trans[(size_t)c * C + c1] * SC : NEG;
        }
    }

    const float* em_b = p.em + (size_t)b * Tmax * ldc;
    const size_t row0 = (size_t)b * (Tmax + 1);
    if (!VIT && owner) p.fbeta[row0 * ldc + c] = beta;

    // Running normaliser: every stored quantity of frame n is relative to nu_n = sum_{m<=n} delta_m,
    // delta_n = max_c gamma~[n-1][c] (lagged), so values stay O(1) however long the video is.
    float delta = 0.0f;
    double nu = 0.0;

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
            const float e = ecur[f] * SC - delta;
            nu += (double)delta;
            if (!VIT && gtid == 0) p.fdelta[row0 + n] = delta;
            // ---- shift-and-add the span window --------------------------------------------
            float carry = 0.0f;
            if (S > 1) carry = __shfl_up_sync(FULL, A[KR - 1], CPW);
#pragma unroll
            for (int i = KR - 1; i > 0; --i) A[i] = A[i - 1] + e;
            A[0] = (j == 0 ? beta : carry) + e;
            // ---- phase 1: reduce over k ---------------------------------------------------
            float gamma;
            int bk = 0;
            if constexpr (VIT) {
                float best = A[0] + LN(0);
#pragma unroll
                for (int i = 1; i < KR; ++i) {
                    const float v = A[i] + LN(i);
                    if (v > best) {
                        best = v;
                        bk = i;
                    }
                }
                bk += j * KR + 1;
                slice_argmax<S>(best, bk);
                gamma = best;
            } else {
                float m = A[0] + LN(0);
#pragma unroll
                for (int i = 1; i < KR; ++i) m = fmaxf(m, A[i] + LN(i));
                m = slice_max<S>(m);
                float s = 0.0f;
#pragma unroll
                for (int i = 0; i < KR; ++i) s += ex2(A[i] + LN(i) - m);
                s = slice_sum<S>(s);
                gamma = m + lg2(s);
            }
            float* gs = gam_s + (n & 1) * cpad;
            if (j == 0) gs[wig * CPW + cl] = valid ? gamma : NEG;
            if (!VIT && owner) p.fgamma[(row0 + n) * ldc + c] = gamma;
            group_sync(W, bar_id);
            if (n == T) {
                if (VIT && owner) p.bp[(row0 + n) * ldc + c] = (uint32_t)bk << 16;
                break;
            }
            // ---- phase 2: transitions -----------------------------------------------------
            float gm = NEG;
            if constexpr (VIT) {
                float best = NEG;
                int bc = j;
                if constexpr (TREG) {
#pragma unroll
                    for (int i = 0; i < CRR; ++i) {
                        const int c1 = i * S + j;
                        if (c1 < C) {
                            const float gv = gs[c1];
                            gm = fmaxf(gm, gv);
                            const float v = gv + tr[i];
                            if (v > best || i == 0) {
                                best = v;
                                bc = c1;
                            }
                        }
                    }
                } else {
                    for (int c1 = j; c1 < C; c1 += S) {
                        const float gv = gs[c1];
                        gm = fmaxf(gm, gv);
                        const float v = gv + transT[c1 * ldT + c];
                        if (v > best || c1 == j) {
                            best = v;
                            bc = c1;
                        }
                    }
                }
                slice_argmax<S>(best, bc);
                beta = best;
                if (owner) p.bp[(row0 + n) * ldc + c] = ((uint32_t)bk << 16) | (uint32_t)bc;
            } else {
                float m = NEG;
                if constexpr (TREG) {
#pragma unroll
                    for (int i = 0; i < CRR; ++i) {
                        const int c1 = i * S + j;
                        if (c1 < C) {
                            const float gv = gs[c1];
                            gm = fmaxf(gm, gv);
                            m = fmaxf(m, gv + tr[i]);
                        }
                    }
                } else {
                    for (int c1 = j; c1 < C; c1 += S) {
                        const float gv = gs[c1];
                        gm = fmaxf(gm, gv);
                        m = fmaxf(m, gv + transT[c1 * ldT + c]);
                    }
                }
                m = slice_max<S>(m);
                float s = 0.0f;
                if constexpr (TREG) {
#pragma unroll
                    for (int i = 0; i < CRR; ++i) {
                        const int c1 = i * S + j;
                        if (c1 < C) s += ex2(gs[c1] + tr[i] - m);
                    }
                } else {
                    for (int c1 = j; c1 < C; c1 += S) s += ex2(gs[c1] + transT[c1 * ldT + c] - m);
                }
                s = slice_sum<S>(s);
                beta = m + lg2(s);
                if (owner) p.fbeta[(row0 + n) * ldc + c] = beta;
            }
            if (!valid) beta = NEG;
            delta = slice_max<S>(gm);
        }
    }

    // ---- termination ---------------------------------------------------------------------
    const float* gT = gam_s + (T & 1) * cpad;
    const float* endb = p.end ? p.end + (size_t)b * C : nullptr;
    if constexpr (!VIT) {
        if (wig == 0) {
            float m = NEG;
            for (int cc = lane; cc < C; cc += 32) m = fmaxf(m, gT[cc] + (endb ? endb[cc] * SC : 0.0f));
            m = warp_max(m);
            float s = 0.0f;
            for (int cc = lane; cc < C; cc += 32) s += ex2(gT[cc] + (endb ? endb[cc] * SC : 0.0f) - m);
            s = warp_sum(s);
            if (lane == 0) {
                const float lzrel = m + lg2(s);  // relative to nu_T
                p.logz2[b] = lzrel;
                p.logz[b] = (nu + (double)lzrel) * LN2 + (p.offset ? p.offset[b] : 0.0);
            }
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
        if (lane == 0 && p.score) p.score[b] = nu + (double)best + (p.offset ? p.offset[b] : 0.0);
        int n = T, cc = bc;
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
template <int KR, int S, bool TREG, bool LREG, int MAXT>
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

    // shared: [trans C*ldT (!TREG)] [len columns KR*G (!LREG)] [per group: zeta 2*cpad, Etr C*ldT (!TREG)]
    float* trans_s = smem;
    float* lens = smem + (TREG ? 0 : C * ldT);
    float* gbase = lens + (LREG ? 0 : KR * G);
    const int per_group = 2 * cpad + (TREG ? 0 : C * ldT);
    float* zet_s = gbase + slot * per_group;
    float* etr_s = zet_s + 2 * cpad;

    const int cl = lane % CPW, j = lane / CPW;
    const int c = wig * CPW + cl;
    const bool valid = c < C;
    const bool owner = valid && j == 0;

    if constexpr (!TREG) {
        for (int i = threadIdx.x; i < C * C; i += blockDim.x) {
            const int c2 = i / C, c1 = i - c2 * C;
            trans_s[c2 * ldT + c1] = p.trans[i] * SC;
        }
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
    if constexpr (!TREG || !LREG) __syncthreads();

    const int vidx = blockIdx.x * p.VPB + slot;
    if (vidx >= p.B) return;
    const int b = p.order ? p.order[vidx] : vidx;
    const int T = p.lengths[b];
    const int bar_id = 1 + slot;

    float Bq[KR], El[KR];
    float ln[LREG ? KR : 1];
    const float* lnp = lens + gtid;
#pragma unroll
    for (int i = 0; i < KR; ++i) {
        Bq[i] = NEG;
        El[i] = 0.0f;
        if constexpr (LREG) {
            const int k = j * KR + i + 1;
            ln[i] = (valid && k <= L) ? p.lenp[(size_t)k * C + c] * SC : NEG;
        }
    }
#define LN(i) (LREG ? ln[LREG ? (i) : 0] : lnp[(i) * G])
    float tr[TREG ? CRR : 1], Et[TREG ? CRR : 1];
    if constexpr (TREG) {
#pragma unroll
        for (int i = 0; i < CRR; ++i) {
            const int c2 = i * S + j;
            tr[i] = (valid && c2 < C) ? p.trans[(size_t)c2 * C + c] * SC : NEG;
            Et[i] = 0.0f;
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
    float* dem = p.d_em + (size_t)b * Tmax * ldc;
    const float* fd = p.fdelta + row0;

    // backward quantities of frame n are relative to mu_n = logZ - nu_n (see the forward kernel), so
    // posteriors are exp(forward~ + backward~) of O(1) numbers.
    float eta = valid ? endc - lzrel : NEG;
    float occ = 0.0f, comp = 0.0f;  // Kahan-compensated occupancy
    float Fprev = valid ? w * ex2(__ldg(fg + (size_t)T * ldc + c) + endc - lzrel) : 0.0f;
    float Sprev = 0.0f;

    // frames beyond the video: zero gradient
    for (int i = T * ldc + gtid; i < Tmax * ldc; i += G) dem[i] = 0.0f;

    float enext[F], bnext[F], gnext[F], dnext[F];
#pragma unroll
    for (int f = 0; f < F; ++f) {
        const int n = T - 1 - f;
        const bool ok = valid && n >= 0;
        dnext[f] = (n >= 0) ? __ldg(fd + n + 1) : 0.0f;
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
            dnext[f] = (n >= 0) ? __ldg(fd + n + 1) : 0.0f;
            enext[f] = ok ? __ldg(em_b + (size_t)n * ldc + c) : 0.0f;
            bnext[f] = (ok && n > 0) ? __ldg(fb + (size_t)n * ldc + c) : 0.0f;
            gnext[f] = (ok && n > 0) ? __ldg(fg + (size_t)n * ldc + c) : 0.0f;
        }
#pragma unroll
        for (int f = 0; f < F; ++f) {
            const int n = n0 - f;
            if (n < 0) break;
            const float e = ecur[f] * SC - dcur[f];
            float carry = 0.0f;
            if (S > 1) carry = __shfl_up_sync(FULL, Bq[KR - 1], CPW);
#pragma unroll
            for (int i = KR - 1; i > 0; --i) Bq[i] = Bq[i - 1] + e;
            Bq[0] = (j == 0 ? eta : carry) + e;
            // ---- phase 1: zeta[n][c] and length counts -------------------------------------
            float m = Bq[0] + LN(0);
#pragma unroll
            for (int i = 1; i < KR; ++i) m = fmaxf(m, Bq[i] + LN(i));
            m = slice_max<S>(m);
            const float betan = (n == 0) ? init_c : bcur[f];
            const float coef = valid ? w * ex2(betan + m) : 0.0f;
            float s = 0.0f;
#pragma unroll
            for (int i = 0; i < KR; ++i) {
                const float pr = ex2(Bq[i] + LN(i) - m);
                s += pr;
                El[i] = fmaf(pr, coef, El[i]);
            }
            s = slice_sum<S>(s);
            const float zeta = m + lg2(s);
            const float Sc = coef * s;
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
            float* zs = zet_s + (n & 1) * cpad;
            if (j == 0) zs[wig * CPW + cl] = valid ? zeta : NEG;
            group_sync(W, bar_id);
            // ---- phase 2: eta[n][c1] and transition counts ------------------------------------
            float m2 = NEG;
            if constexpr (TREG) {
#pragma unroll
                for (int i = 0; i < CRR; ++i) {
                    const int c2 = i * S + j;
                    if (c2 < C) m2 = fmaxf(m2, zs[c2] + tr[i]);
                }
            } else {
                for (int c2 = j; c2 < C; c2 += S) m2 = fmaxf(m2, zs[c2] + trans_s[c2 * ldT + c]);
            }
            m2 = slice_max<S>(m2);
            const float coef2 = valid ? w * ex2(gcur[f] + m2) : 0.0f;
            float s2 = 0.0f;
            if constexpr (TREG) {
#pragma unroll
                for (int i = 0; i < CRR; ++i) {
                    const int c2 = i * S + j;
                    if (c2 < C) {
                        const float pr = ex2(zs[c2] + tr[i] - m2);
                        s2 += pr;
                        Et[i] = fmaf(pr, coef2, Et[i]);
                    }
                }
            } else {
                for (int c2 = j; c2 < C; c2 += S) {
                    const float pr = ex2(zs[c2] + trans_s[c2 * ldT + c] - m2);
                    s2 += pr;
                    if (valid) etr_s[c2 * ldT + c] += pr * coef2;
                }
            }
            s2 = slice_sum<S>(s2);
            eta = valid ? m2 + lg2(s2) : NEG;
            Fprev = coef2 * s2;
        }
    }
    // ---- flush the per-video counts ------------------------------------------------------------
    if (valid) {
#pragma unroll
        for (int i = 0; i < KR; ++i) {
            const int k = j * KR + i + 1;
            if (k <= L) atomicAdd(p.d_len + (size_t)k * C + c, El[i]);
        }
        if constexpr (TREG) {
#pragma unroll
            for (int i = 0; i < CRR; ++i) {
                const int c2 = i * S + j;
                if (c2 < C) atomicAdd(p.d_trans + (size_t)c2 * C + c, Et[i]);
            }
        } else {
            for (int c2 = j; c2 < C; c2 += S) atomicAdd(p.d_trans + (size_t)c2 * C + c, etr_s[c2 * ldT + c]);
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
constexpr int kMaxThreadsSmallTreg = 128, kMaxThreadsSmall = 512;
// long windows keep KR partial scores per thread: fewer threads per CTA so that they stay in registers
constexpr int max_threads_big(int KR, int mode) { return mode == 2 ? (KR > 50 ? 256 : 384) : (KR > 50 ? 384 : 576); }

struct RegChoice {
    int v;      // variant index, -1 = none
    int W;      // warps per video
    int VPB;    // videos per block
    bool treg;  // transitions in registers
    size_t smem;
};

static size_t smem_bytes(const RegVariant& rv, int C, int W, int vpb, bool treg, int mode) {
    const int cpw = 32 / rv.S, cpad = W * cpw, ldT = ld_trans(W, rv.S), G = W * 32;
    size_t fl = 0;
    if (!treg) fl += (size_t)C * ldT;
    if (!rv.lreg) fl += (size_t)rv.KR * G;
    fl += (size_t)vpb * (2 * cpad + ((mode == 2 && !treg) ? C * ldT : 0));
    return fl * sizeof(float);
}

static RegChoice choose(int C, int L, int mode) {
    RegChoice best{-1, 0, 0, false, 0};
    double best_cost = 1e30;
    for (int v = 0; v < kNumVariants; ++v) {
        const RegVariant& rv = kVariants[v];
        if (rv.KR * rv.S < L) continue;
        const int cpw = 32 / rv.S;
        const int W = (C + cpw - 1) / cpw;
        const bool treg = rv.lreg && (W == 1);
        const int maxt = rv.lreg ? (treg ? kMaxThreadsSmallTreg : kMaxThreadsSmall) : max_threads_big(rv.KR, mode);
        if (W * 32 > maxt) continue;
        int vpb = treg ? 4 : maxt / (W * 32);
        if (vpb > 8) vpb = 8;
        if (!rv.lreg) vpb = 1;
        while (vpb > 1 && smem_bytes(rv, C, W, vpb, treg, mode) > kSmemCap) --vpb;
        const size_t sm = smem_bytes(rv, C, W, vpb, treg, mode);
        if (sm > kSmemCap) continue;
        // issue slots per frame ~ warps * (window + transitions per lane) (+ barrier cost when W > 1)
        const int crr = treg ? (cpw + rv.S - 1) / rv.S : (C + rv.S - 1) / rv.S;
        const double cost = (double)W * (rv.KR * (rv.lreg ? 1.0 : 1.3) + crr) + (W > 1 ? 6.0 * W : 0.0);
        if (cost < best_cost) {
            best_cost = cost;
            best = RegChoice{v, W, vpb, treg, sm};
        }
    }
    return best;
}

template <int MODE, int KR, int S, bool TREG, bool LREG, int MAXT>
static int launch_one(const DpParams& p, const RegChoice& ch, cudaStream_t st) {
    const int blocks = (p.B + ch.VPB - 1) / ch.VPB;
    const int threads = ch.VPB * ch.W * 32;
    cudaError_t e = cudaSuccess;
    if constexpr (MODE == 0) {
        auto k = dp_forward_kernel<true, KR, S, TREG, LREG, MAXT>;
        if (ch.smem > 48 * 1024) e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ch.smem);
        if (e == cudaSuccess) k<<<blocks, threads, ch.smem, st>>>(p);
    } else if constexpr (MODE == 1) {
        auto k = dp_forward_kernel<false, KR, S, TREG, LREG, MAXT>;
        if (ch.smem > 48 * 1024) e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ch.smem);
        if (e == cudaSuccess) k<<<blocks, threads, ch.smem, st>>>(p);
    } else {
        auto k = dp_backward_kernel<KR, S, TREG, LREG, MAXT>;
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
    return ch.treg ? launch_one<MODE, KR, S, true, true, kMaxThreadsSmallTreg>(p, ch, st)
                   : launch_one<MODE, KR, S, false, true, kMaxThreadsSmall>(p, ch, st);
}
template <int MODE, int KR, int S>
static int launch_big(const DpParams& p, const RegChoice& ch, cudaStream_t st) {
    return launch_one<MODE, KR, S, false, false, max_threads_big(KR, MODE)>(p, ch, st);
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
bool dp_reg_supported(int C, int L, int mode) { return choose(C, L, mode).v >= 0; }

const char* dp_reg_name(int C, int L, int mode) {
    static thread_local char buf[96];
    RegChoice ch = choose(C, L, mode);
    if (ch.v < 0) return "none";
    snprintf(buf, sizeof(buf), "reg<KR=%d,S=%d>/%s/%s/W=%d/VPB=%d/smem=%zu", kVariants[ch.v].KR, kVariants[ch.v].S,
             ch.treg ? "trans-reg" : "trans-smem", kVariants[ch.v].lreg ? "len-reg" : "len-smem", ch.W, ch.VPB, ch.smem);
    return buf;
}

int dp_reg_launch(DpParams p, int mode, cudaStream_t st) {
    RegChoice ch = choose(p.C, p.L, mode);
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
