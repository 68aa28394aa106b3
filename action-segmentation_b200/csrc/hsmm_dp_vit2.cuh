// Max-plus Viterbi with deferred length arg-max, for the one-warp-per-video shapes with L <= 32 (sm_100a).
//
// The forward sweep of hsmm_dp_reg.cuh spends most of its issue slots on the per-(class, length) compare/select
// that tracks WHICH length attains gamma[n][c] = max_k A_k + len[k,c].  Only one (n, c) per decoded segment ever
// needs that answer, so this kernel keeps the sweep to add + add + 3-input max per element, stores per frame
//     beta^[n][c]  (score of "a class-c segment starts at n", relative to the running normaliser),
//     pred[n][c]   (arg-max predecessor class of that start: the cheap arg-max over <= 4 / C transitions),
//     gm_n         (normaliser increments),
// and recomputes the length arg-max during the back-trace: for a segment ending at n in class c, lane k-1 rebuilds
// A_k in exactly the order the sweep used -- (beta^[n-k] + e) + eo + eo ... , one shuffle + add per step -- so the
// candidates are bit-identical to the sweep's registers and the decoded path is the one dp_forward_kernel<VIT>
// returns (ties to the smaller length / class).  Videos with no path through a sparse transition hint are flagged
// and decoded by dp_forward_kernel<VIT> against the dense matrix (DpParams::only_flagged).
#pragma once
#include "hsmm_dp_reg.cuh"

namespace hsmm {

template <int KR, int S, int TM>
__device__ __forceinline__ void dp_vit2_kernel_body(const DpParams& p, const int bid) {
    constexpr int CPW = Lay<S>::CPW;
    constexpr int CRR = Lay<S>::CRR;
    extern __shared__ __align__(16) float smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int C = p.C, L = p.L, ldc = p.ldc, Tmax = p.Tmax;
    float* gam_s = smem + warp * (2 * CPW + 2);

    const int cl = lane % CPW, j = lane / CPW;
    const int c = cl;
    const bool valid = c < C;
    const bool owner = valid && j == 0;

    const int vidx = bid * (blockDim.x >> 5) + warp;
    if (vidx >= p.B) return;
    const int b = p.order ? p.order[vidx] : vidx;
    const int T = p.lengths[b];

    float ln[KR];
#pragma unroll
    for (int i = 0; i < KR; ++i) {
        const int k = j * KR + i + 1;
        ln[i] = (valid && k <= L) ? p.lenp[(size_t)k * C + c] : NEG;
    }
    const float init_c = valid ? p.init[c] : NEG;
    const float* endb = p.end ? p.end + (size_t)b * C : nullptr;

    float tr[TM == 0 ? CRR : 1];
    if constexpr (TM == 0) {
#pragma unroll
        for (int i = 0; i < CRR; ++i) {
            const int c1 = j * CRR + i;
            tr[i] = (valid && c1 < C) ? p.trans[(size_t)c * C + c1] : NEG;
        }
    }
    int pidx[TM == 2 ? SPW : 1];
    float pval[TM == 2 ? SPW : 1];
    if constexpr (TM == 2) {
#pragma unroll
        for (int q = 0; q < SPW; ++q) {
            const int c1 = valid ? p.trans_pred[c * SPW + q] : -1;
            pidx[q] = c1 >= 0 ? c1 : 0;
            pval[q] = c1 >= 0 ? p.trans[(size_t)c * C + c1] : NEG;
        }
    }

    const float* em_b = p.em + (size_t)b * Tmax * ldc;
    const size_t row0 = (size_t)b * (Tmax + 1);
    float* const vbeta = p.vbeta + row0 * ldc;
    uint32_t* const vpred = p.vpred + row0 * ldc;
    float* const vdelta = p.vdelta + row0;
    if (owner) vbeta[c] = init_c;

    float A[KR];
#pragma unroll
    for (int i = 0; i < KR; ++i) A[i] = NEG;
    float beta = init_c, gprev = NEG, gmprev = 0.0f, nu4 = 0.0f;
    double nu = 0.0;

    const float* ep = em_b + c;
    float enext[F];
#pragma unroll
    for (int f = 0; f < F; ++f) enext[f] = (valid && f < T) ? __ldg(ep + f * ldc) : 0.0f;
    ep += F * ldc;
    float* bout = vbeta + ldc + c;
    uint32_t* pout = vpred + ldc + c;
    float* dout = vdelta + 1;

#pragma unroll 1
    for (int n0 = 1; n0 <= T; n0 += F) {
        float ecur[F];
#pragma unroll
        for (int f = 0; f < F; ++f) ecur[f] = enext[f];
#pragma unroll
        for (int f = 0; f < F; ++f) enext[f] = (valid && n0 - 1 + F + f < T) ? __ldg(ep + f * ldc) : 0.0f;
        ep += F * ldc;
#pragma unroll
        for (int f = 0; f < F; ++f) {
            const int n = n0 + f;
            if (n > T) break;
            const float e = ecur[f];
            nu4 += gmprev;
            // ---- phase 1: gamma~[n][c] = max_k A_k + len[k,c] (value only) ---------------------------
            const float eo = e - gmprev;
            float carry = 0.0f;
            if (S > 1) carry = __shfl_up_sync(FULL, A[KR - 1], CPW);
#pragma unroll
            for (int i = KR - 1; i > 0; --i) A[i] = A[i - 1] + eo;
            A[0] = (j == 0) ? beta + e : carry + eo;
            float m0 = A[0] + ln[0], m1 = NEG;
#pragma unroll
            for (int i = 1; i + 1 < KR; i += 2) {
                const float v0 = A[i] + ln[i], v1 = A[i + 1] + ln[i + 1];
                if ((i >> 1) & 1)
                    m1 = fmaxf(m1, fmaxf(v0, v1));
                else
                    m0 = fmaxf(m0, fmaxf(v0, v1));
            }
            if ((KR & 1) == 0) m1 = fmaxf(m1, A[KR - 1] + ln[KR - 1]);
            float gamma = slice_max<S>(fmaxf(m0, m1));
            gamma = valid ? gamma : NEG;
            gprev = gamma;
            const float gm = warp_max_redux(owner ? gamma : NEG);
            if (n == T) break;
            if (lane == 0) *dout = gm;
            ++dout;
            gmprev = gm;
            // ---- phase 2: beta^[n][c2] = max_c1 gamma~[n][c1] + trans[c2,c1] - gm, with its arg-max -----
            float best = NEG;
            int bc = 0;
            if constexpr (TM == 2) {
#pragma unroll
                for (int q = 0; q < SPW; ++q) {
                    const float v = __shfl_sync(FULL, gamma, pidx[q]) + pval[q];
                    if (v > best || q == 0) {
                        best = v;
                        bc = pidx[q];
                    }
                }
            } else {
                float* gs = gam_s + (n & 1) * CPW;
                if (j == 0) gs[cl] = gamma;
                __syncwarp();
#pragma unroll
                for (int i = 0; i < CRR; ++i) {
                    const float v = gs[j * CRR + i] + tr[i];
                    if (v > best || i == 0) {
                        best = v;
                        bc = j * CRR + i;
                    }
                }
                slice_argmax<S>(best, bc);
            }
            beta = valid ? best - gm : NEG;
            if (owner) {
                *bout = beta;
                *pout = (uint32_t)bc;
            }
            bout += ldc;
            pout += ldc;
        }
        nu += (double)nu4;
        nu4 = 0.0f;
    }

    // ---- termination: best class at T (ties to the smaller class) -----------------------------------
    float* gT = gam_s + (T & 1) * CPW;
    __syncwarp();
    if (j == 0) gT[cl] = valid ? gprev : NEG;
    __syncwarp();
    float fv = NEG;
    for (int cc = lane; cc < C; cc += 32) fv = fmaxf(fv, gT[cc] + (endb ? endb[cc] : 0.0f));  // one class per lane (C <= 32)
    const float final_v = warp_max_redux(fv);
    const unsigned fm = __ballot_sync(FULL, lane < C && fv == final_v);
    int cc = fm ? (__ffs(fm) - 1) : 0;
    const double total = nu + (double)final_v;
    const bool degenerate = (TM == 2) && !(total > (double)DEGENERATE);
    if (lane == 0) p.vflag[b] = degenerate ? 1.0f : 0.0f;
    if (degenerate) return;  // dp_forward_kernel<VIT> decodes this video against the dense matrix

    // ---- outputs: prefill, then walk back ---------------------------------------------------------------
    const int eos = p.class_ids ? p.class_ids[C] : C;
    int64_t* sp = p.spans + (size_t)b * (Tmax + 1);
    for (int i = lane; i <= Tmax; i += 32) sp[i] = (i == T) ? (int64_t)eos : (int64_t)-1;
    int64_t* lab = p.labels ? p.labels + (size_t)b * Tmax : nullptr;
    if (lab)
        for (int i = T + lane; i < Tmax; i += 32) lab[i] = eos;
    if (lane == 0 && p.score) p.score[b] = total + (p.offset ? p.offset[b] : 0.0);
    __syncwarp();

    int n = T;
    while (n > 0) {
        const int kmax = L < n ? L : n;  // <= 32
        const int k = lane + 1;
        const bool act = k <= kmax;
        const int m_own = n - k + 1;  // frame whose emission opens the length-k candidate
        const float e_own = act ? __ldcg(em_b + (size_t)(m_own - 1) * ldc + cc) : 0.0f;
        const float g_own = (act && m_own >= 2) ? __ldcg(vdelta + m_own - 1) : 0.0f;  // gm_{m_own - 1}
        const float b_own = act ? __ldcg(vbeta + (size_t)(n - k) * ldc + cc) : NEG;
        const uint32_t p_own = act ? __ldcg(vpred + (size_t)(n - k) * ldc + cc) : 0u;
        const float l_own = act ? __ldg(p.lenp + (size_t)k * C + cc) : NEG;
        const float eo_own = e_own - g_own;
        float acc = b_own + e_own;  // A[0] of the sweep at frame m_own
#pragma unroll
        for (int s = 1; s < 32; ++s) {
            if (s >= kmax) break;
            const float t = __shfl_up_sync(FULL, eo_own, s);  // eo of frame m_own + s
            if (lane >= s) acc += t;
        }
        const float v = act ? acc + l_own : NEG;
        const float vm = warp_max_redux(v);
        const unsigned km = __ballot_sync(FULL, act && v == vm);
        const int kstar = km ? __ffs(km) : 1;  // smallest length attaining the maximum
        const int start = n - kstar;
        const int c1 = (int)__shfl_sync(FULL, p_own, kstar - 1);
        const int64_t cid = p.class_ids ? p.class_ids[cc] : cc;
        if (lane == 0) sp[start] = cid;
        if (lab)
            for (int t = start + lane; t < n; t += 32) lab[t] = cid;
        cc = (c1 < C) ? c1 : C - 1;
        n = start;
    }
}

template <int KR, int S, int TM>
__global__ void __launch_bounds__(128) dp_vit2_kernel(const DpParams p) {
    dp_vit2_kernel_body<KR, S, TM>(p, blockIdx.x);
}
template <int KR, int S, int TM>
__global__ void __launch_bounds__(128) dp_vit2_kernel_grouped(const __grid_constant__ DpGroup g) {
    int local;
    const int t = group_find(g, blockIdx.x, local);
    dp_vit2_kernel_body<KR, S, TM>(g.t[t], local);
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static inline bool vit2_eligible(const RegChoice& ch, int L) {
    return ch.v >= 0 && ch.W == 1 && kVariants[ch.v].lreg && L <= 32 && (ch.tm == 0 || ch.tm == 2) &&
           (ch.v == 0 || ch.v == 1 || ch.v == 2 || ch.v == 3 || ch.v == 6);
}

template <int KR, int S>
static int launch_vit2_tm(const DpParams& p, int tm, cudaStream_t st) {
    constexpr int VPB = 4;
    const int blocks = (p.B + VPB - 1) / VPB;
    const size_t smem = VPB * (2 * (32 / S) + 2) * sizeof(float);
    if (tm == 2)
        dp_vit2_kernel<KR, S, 2><<<blocks, VPB * 32, smem, st>>>(p);
    else
        dp_vit2_kernel<KR, S, 0><<<blocks, VPB * 32, smem, st>>>(p);
    return check_launch("dp_vit2 kernel");
}

static int launch_vit2(const DpParams& p, const RegChoice& ch, cudaStream_t st) {
    switch (ch.v) {
        case 0: return launch_vit2_tm<10, 2>(p, ch.tm, st);
        case 1: return launch_vit2_tm<20, 1>(p, ch.tm, st);
        case 2: return launch_vit2_tm<13, 4>(p, ch.tm, st);
        case 3: return launch_vit2_tm<25, 2>(p, ch.tm, st);
        case 6: return launch_vit2_tm<32, 1>(p, ch.tm, st);
    }
    set_error("no deferred-arg-max Viterbi variant for this shape");
    return -2;
}

}  // namespace hsmm
