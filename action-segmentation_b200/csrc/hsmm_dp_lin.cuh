// Block-floating-point ("linear window") log-semiring kernels for the one-warp-per-video shapes (sm_100a).
//
// Same decomposition, state layout and saved-tensor format as hsmm_dp_reg.cuh (lane = class c x k-slice j,
// span window of KR registers per lane, phase 1 over lengths, phase 2 over transitions), but the window holds
// LINEAR values relative to a per-class log2 reference:
//
//     P[i] = 2^( beta[n-k][c] + em[n-k..n-1,c] - r_n[c] ),   k = j*KR + i + 1,
//     r_n  = e_n + rho_n,   rho_n = max(gamma[n-1][c] - gm_{n-1}, beta[n-1][c] + len[1,c])
//
// so a frame costs one multiply (shift by f = 2^(r_{n-1} - gm_{n-1} - rho_n)) and one fused multiply-add
// (gamma sum against the static 2^len[k,c]) per (class, length) instead of add + add + ex2 + add, and only
// O(1) MUFU operations per class: f, the entering element, and lg2 of the sum.  The transition phase stays in
// the log domain.  This is what makes the path FP32-issue-bound instead of MUFU-bound.
//
// What a float window cannot hold is an element more than 2^126 below its class reference: it is flushed to
// zero.  The kernel carries a rigorous upper bound of the flushed mass relative to the reference
// (U <- U - lg2 s_n per frame, because ref_{n+1} >= s_n * ref_n * (common factor); floor 2^-121 for new losses;
// two alternating blocks of L frames because an element lives at most L-1 frames) and FLAGS the video when
//   * a class sum s_n <= 2^-30 of its reference, or
//   * the bound says flushed mass could exceed 2^-24 of a class sum, or
//   * the length table is unsuitable (a usable length below 2^-100, or 2^(maxstep - min len) > 2^110), or
//   * the result is degenerate (no unmasked path).
// Flagged videos are recomputed by the log-domain kernels of hsmm_dp_reg.cuh (launched right behind with
// DpParams::only_flagged), so results never depend on which path ran.  Classes whose reference is below -1e8
// (no unmasked path reaches them yet) carry no window at all.
#pragma once
#include "hsmm_dp_reg.cuh"

namespace hsmm {

constexpr float LIN_TINY = 9.3132257e-10f;  // 2^-30
constexpr float LIN_DEAD = -1.0e8f;         // log2 units: below this a class has no unmasked path
constexpr float LIN_FLOOR = -121.0f;        // lg2 of (L <= 32 newly flushed elements, each < 2^-126)
constexpr float LIN_RELEVANT = -24.0f;

struct LinTracker {
    float ua, ub;
    int cnt;
    __device__ __forceinline__ void init() {
        ua = NEG;
        ub = NEG;
        cnt = 0;
    }
    // g = -lg2(s_n); returns the bound of lg2(flushed mass / class sum) at this frame
    __device__ __forceinline__ float step(float g, float lnmax, int L) {
        ua += g;
        ub = fmaxf(ub + g, LIN_FLOOR);
        const float chk = fmaxf(ua, ub) + lnmax + g;
        if (++cnt >= L) {
            ua = ub;
            ub = NEG;
            cnt = 0;
        }
        return chk;
    }
};

// ---------------------------------------------------------------------------------------------
// forward (log-partition)
// ---------------------------------------------------------------------------------------------
// XP: per-class state (beta, gamma, the window reference) and the saved planes in double, window in float -- for score
// tensors that carry the -1e4 narration penalty (see state_t in hsmm_dp_reg.cuh): the penalty enters the reference
// and cancels in the window shift, in double, before it meets an O(1) number.
template <bool XP, int KR, int S, int TM>
__device__ __forceinline__ void dp_lin_forward_kernel_body(const DpParams& p, const int bid) {
    using ST = state_t<XP>;
    constexpr int CPW = Lay<S>::CPW;
    constexpr int CRR = Lay<S>::CRR;
    extern __shared__ __align__(16) float smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int C = p.C, L = p.L, ldc = p.ldc, Tmax = p.Tmax;
    const float SC = LOG2E;
    float* gam_s = smem + warp * (2 * CPW + 2);  // per video: two gamma rows (dense phase 2, termination)

    const int cl = lane % CPW, j = lane / CPW;
    const int c = cl;
    const bool valid = c < C;
    const bool owner = valid && j == 0;

    const int vidx = bid * (blockDim.x >> 5) + warp;
    if (vidx >= p.B) return;
    const int b = p.order ? p.order[vidx] : vidx;
    const int T = p.lengths[b];

    // ---- per-lane constants ----------------------------------------------------------------
    float pl[KR];
    float lnmax = NEG, lnmin = -NEG, maxstep = NEG;
    {
        float prev = NEG;
#pragma unroll
        for (int i = 0; i < KR; ++i) {
            const int k = j * KR + i + 1;
            const bool use = valid && k <= L;
            const float v = use ? p.lenp[(size_t)k * C + c] * SC : NEG;
            pl[i] = use ? ex2(v) : 0.0f;
            if (use) {
                lnmax = fmaxf(lnmax, v);
                lnmin = fminf(lnmin, v);
            }
            if (i == 0 && j > 0) prev = (valid && k - 1 <= L) ? p.lenp[(size_t)(k - 1) * C + c] * SC : NEG;
            if (k >= 2 && use) maxstep = fmaxf(maxstep, v - prev);
            prev = v;
        }
        lnmax = slice_max<S>(lnmax);
        lnmin = -slice_max<S>(-lnmin);
        maxstep = slice_max<S>(maxstep);
        if (maxstep < -1.0e29f) maxstep = 0.0f;
    }
    // the window must stay inside float range: P <= 2^(max(maxstep,0) - len), len >= 2^-100
    bool bad = valid && (lnmin < -100.0f || fmaxf(maxstep, 0.0f) - lnmin > 110.0f);
    if (__any_sync(FULL, bad)) {
        // a length table this path cannot hold (e.g. Poisson tails at K = 100) is known before the first frame: hand the
        // video to the log-domain kernel at once instead of after a full, discarded pass
        if (lane == 0) p.fflag[b] = 2.0f;
        return;
    }
    const float ln_first = valid ? p.lenp[(size_t)C + c] * SC : NEG;  // len[1, c]
    const float init_c = valid ? p.init[c] * SC : NEG;
    const float* endb = p.end ? p.end + (size_t)b * C : nullptr;

    float tr[TM == 0 ? CRR : 1];
    float trmax = 0.0f;
    if constexpr (TM == 0) {
        float m = NEG;
        if (valid)
            for (int c1 = 0; c1 < C; ++c1) m = fmaxf(m, p.trans[(size_t)c * C + c1] * SC);
        trmax = valid ? m : 0.0f;
#pragma unroll
        for (int i = 0; i < CRR; ++i) {
            const int c1 = j * CRR + i;
            tr[i] = (valid && c1 < C) ? p.trans[(size_t)c * C + c1] * SC - trmax : NEG;
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
    if (owner) fbeta[row0 * ldc + c] = (ST)init_c;

    float P[KR];
#pragma unroll
    for (int i = 0; i < KR; ++i) P[i] = 0.0f;
    ST beta = init_c;  // beta^[n-1][c], relative to nu_n
    ST gprev = NEG;    // gamma~[n-1][c], relative to nu_{n-1}
    ST rref = 0, eprev = 0;
    float gmprev = 0.0f;
    double nu = 0.0;
    float nu4 = 0.0f;  // normaliser increments of the current group of 4 frames
    LinTracker trk;
    trk.init();

    // emission prefetch: the next group of F frames is in flight while this group is processed (static register
    // names: a rotating ring would make every frame wait for the load issued one frame earlier)
    const float* ep = em_b + c;
    float enext[F];
#pragma unroll
    for (int f = 0; f < F; ++f) enext[f] = (valid && f < T) ? __ldg(ep + f * ldc) : 0.0f;
    ep += F * ldc;
    ST* gout = fgamma + (row0 + 1) * ldc + c;
    ST* bout = fbeta + (row0 + 1) * ldc + c;
    float* dout = p.fdelta + row0 + 1;

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
        const float efr = ecur[f];
        const ST e = (ST)efr * (ST)SC;
        nu4 += gmprev;
        // ---- phase 1: shift the window, add the entering element, sum against 2^len ----------
        const ST rho = smax(gprev - (ST)gmprev, beta + (ST)ln_first);
        const bool dead = rho < (ST)LIN_DEAD;
        const float eo = (float)((eprev - rho) + (rref - (ST)gmprev));
        const float fac = dead ? 0.0f : ex2(fminf(eo, 100.0f));
        float carry = 0.0f;
        if (S > 1) carry = __shfl_up_sync(FULL, P[KR - 1], CPW);
#pragma unroll
        for (int i = KR - 1; i > 0; --i) P[i] = P[i - 1] * fac;
        P[0] = (j == 0) ? (dead ? 0.0f : ex2(fminf((float)(beta - rho), 100.0f))) : carry * fac;
        rref = rho;
        eprev = e;
        // packed FMAs (FFMA2): two (class, length) elements per instruction, two independent accumulator pairs
        float sp[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
        for (int i = 0; i + 1 < KR; i += 2) {
            if ((i >> 1) & 1)
                ffma2(sp[2], sp[3], P[i], P[i + 1], pl[i], pl[i + 1]);
            else
                ffma2(sp[0], sp[1], P[i], P[i + 1], pl[i], pl[i + 1]);
        }
        if (KR & 1) sp[0] = fmaf(P[KR - 1], pl[KR - 1], sp[0]);
        float s = slice_sum<S>((sp[0] + sp[1]) + (sp[2] + sp[3]));
        const bool live = valid && !dead;
        bad |= live && !(s > LIN_TINY);
        s = dead ? 1.0f : fmaxf(s, 1.0e-37f);
        const float lg = lg2(s);
        bad |= live && trk.step(-lg, lnmax, L) > LIN_RELEVANT;
        const ST gamma = valid ? (e + rho) + (ST)lg : (ST)NEG;
        gprev = gamma;
        const float gm = warp_max_redux(owner ? (float)gamma : NEG);
        if (owner) *gout = gamma;
        gout += ldc;
        if (n == T) break;
        if (lane == 0) *dout = gm;
        ++dout;
        gmprev = gm;
        // ---- phase 2 (log domain): beta^[n][c2] = (+)_c1 gamma~[n][c1] + trans[c2,c1] - gm -------------
        if constexpr (TM == 2) {
            ST v[SPW];
            ST m = NEG;
#pragma unroll
            for (int q = 0; q < SPW; ++q) {
                v[q] = __shfl_sync(FULL, gamma, pidx[q]) + (ST)pval[q];
                m = smax(m, v[q]);
            }
            float s2 = 0.0f;
#pragma unroll
            for (int q = 0; q < SPW; ++q) s2 += ex2((float)(v[q] - m));
            beta = valid ? (m - (ST)gm) + (ST)lg2(s2) : (ST)NEG;
        } else {
            // dense transitions: float state only (the XP instantiations are launched for sparse lists only)
            float* gs = gam_s + (n & 1) * CPW;
            if (j == 0) gs[cl] = valid ? (float)gamma : NEG;
            __syncwarp();
            float sq[2] = {0.0f, 0.0f};
#pragma unroll
            for (int i = 0; i < CRR; ++i) sq[i & 1] += ex2((gs[j * CRR + i] - gm) + tr[i]);
            float s2 = slice_sum<S>(sq[0] + sq[1]);
            float mfix = 0.0f;
            const bool under = valid && !(s2 > TINY);
            if (__any_sync(FULL, under)) {  // exact two-pass (cf. dp_forward_kernel)
                float m = NEG;
                for (int c1 = j; c1 < C; c1 += S) m = fmaxf(m, gs[c1] + (valid ? __ldg(p.trans + (size_t)c * C + c1) * SC : NEG));
                m = slice_max<S>(m);
                float s2p = 0.0f;
                for (int c1 = j; c1 < C; c1 += S)
                    s2p += ex2(gs[c1] + (valid ? __ldg(p.trans + (size_t)c * C + c1) * SC : NEG) - m);
                s2p = slice_sum<S>(s2p);
                if (under) {
                    s2 = s2p;
                    mfix = (m - gm) - trmax;
                }
            }
            beta = valid ? (ST)((trmax + mfix) + lg2(s2)) : (ST)NEG;
        }
        if (owner) *bout = beta;
        bout += ldc;
        }
        nu += (double)nu4;  // one double add per group of F frames
        nu4 = 0.0f;
    }

    // ---- termination: gamma[T][c] sits in the owner lanes' registers (one warp per video, C <= 32) ---------------
    const ST vfin = owner ? gprev + (ST)(endb ? endb[c] * SC : 0.0f) : (ST)NEG;
    const ST m = warp_max(vfin);
    const float sfin = warp_sum(owner ? ex2((float)(vfin - m)) : 0.0f);
    const ST final_v = m + (ST)lg2(sfin);
    const double total = (nu + (double)final_v) * LN2;
    bad |= !(total > (double)DEGENERATE);  // degenerate or NaN: the log-domain kernel decides
    const bool flagged = __any_sync(FULL, bad);
    if (lane == 0) {
        p.logz2[b] = (double)final_v;
        p.fflag[b] = flagged ? 2.0f : 0.0f;
        p.logz[b] = total + (p.offset ? p.offset[b] : 0.0);
    }
}

template <bool XP, int KR, int S, int TM>
__global__ void __launch_bounds__(128, (!XP && KR <= 20 && TM == 2) ? 5 : 1) dp_lin_forward_kernel(const DpParams p) {
    dp_lin_forward_kernel_body<XP, KR, S, TM>(p, blockIdx.x);
}
template <bool XP, int KR, int S, int TM>
__global__ void __launch_bounds__(128, (!XP && KR <= 20 && TM == 2) ? 5 : 1) dp_lin_forward_kernel_grouped(const __grid_constant__ DpGroup g) {
    int local;
    const int t = group_find(g, blockIdx.x, local);
    dp_lin_forward_kernel_body<XP, KR, S, TM>(g.t[t], local);
}

// ---------------------------------------------------------------------------------------------
// backward (expected counts)
// ---------------------------------------------------------------------------------------------
template <bool XP, int KR, int S, int TM>
__device__ __forceinline__ void dp_lin_backward_kernel_body(const DpParams& p, const int bid) {
    using ST = state_t<XP>;
    constexpr int CPW = Lay<S>::CPW;
    constexpr int CRR = Lay<S>::CRR;
    constexpr int FB = 2;  // frames per prefetch group (four streams are prefetched: register budget)
    extern __shared__ __align__(16) float smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int C = p.C, L = p.L, ldc = p.ldc, Tmax = p.Tmax;
    const float SC = LOG2E;
    float* zet_s = smem + warp * (2 * CPW + 2);

    const int cl = lane % CPW, j = lane / CPW;
    const int c = cl;
    const bool valid = c < C;
    const bool owner = valid && j == 0;

    const int vidx = bid * (blockDim.x >> 5) + warp;
    if (vidx >= p.B) return;
    const int b = p.order ? p.order[vidx] : vidx;
    const int T = p.lengths[b];
    if (((int)p.fflag[b]) & 3) {  // the forward pass fell back to the dense matrix: so does the backward pass
        if (lane == 0) p.bflag[b] = 1.0f;
        return;
    }

    float Q[KR], El[KR], pl[KR];
    float lnmax = NEG, lnmin = -NEG, maxstep = NEG;
    {
        float prev = NEG;
#pragma unroll
        for (int i = 0; i < KR; ++i) {
            const int k = j * KR + i + 1;
            const bool use = valid && k <= L;
            const float v = use ? p.lenp[(size_t)k * C + c] * SC : NEG;
            pl[i] = use ? ex2(v) : 0.0f;
            if (use) {
                lnmax = fmaxf(lnmax, v);
                lnmin = fminf(lnmin, v);
            }
            if (i == 0 && j > 0) prev = (valid && k - 1 <= L) ? p.lenp[(size_t)(k - 1) * C + c] * SC : NEG;
            if (k >= 2 && use) maxstep = fmaxf(maxstep, v - prev);
            prev = v;
            Q[i] = 0.0f;
            El[i] = 0.0f;
        }
        lnmax = slice_max<S>(lnmax);
        lnmin = -slice_max<S>(-lnmin);
        maxstep = slice_max<S>(maxstep);
        if (maxstep < -1.0e29f) maxstep = 0.0f;
    }
    // reasons for handing the video to the log-domain kernel (bflag = 1 + bits): 2 class sum collapsed, 4 flushed-mass
    // bound, 8 length table, 16 NaN
    int why = (valid && (lnmin < -100.0f || fmaxf(maxstep, 0.0f) - lnmin > 110.0f)) ? 8 : 0;
    if (__any_sync(FULL, why != 0)) {  // known before the first frame: no discarded pass (cf. the forward kernel)
        if (lane == 0) p.bflag[b] = 9.0f;
        return;
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
    const ST lzrel = (ST)p.logz2[b];
    const float w = p.grad[b];
    const float init_c = valid ? p.init[c] * SC : NEG;

    const size_t row0 = (size_t)b * (Tmax + 1);
    const ST* fg0 = reinterpret_cast<const ST*>(p.fgamma) + row0 * ldc + c;
    float* dem = p.d_em + (size_t)b * Tmax * ldc;

    ST eta = valid ? (ST)endc - lzrel : (ST)NEG;  // eta~[T]
    ST zprev = NEG;                               // zeta^[n+1][c]
    ST rref = 0, eprev = 0;
    float occ = 0.0f, comp = 0.0f;
    float Fprev = valid ? w * ex2((float)(__ldcg(fg0 + (size_t)T * ldc) + (ST)endc - lzrel)) : 0.0f;
    float Sprev = 0.0f;
    float gm_next = 0.0f;
    float S0 = 0.0f;
    LinTracker trk;
    trk.init();

    for (int i = T * ldc + lane; i < Tmax * ldc; i += 32) dem[i] = 0.0f;  // frames beyond the video

    // running pointers (frame n0 of the current group); the next group of F frames is in flight meanwhile
    const float* pe = p.em + (size_t)b * Tmax * ldc + (size_t)(T - 1) * ldc + c;
    const ST* pb = reinterpret_cast<const ST*>(p.fbeta) + row0 * ldc + (size_t)(T - 1) * ldc + c;
    const ST* pg = fg0 + (size_t)(T - 1) * ldc;
    const float* pd = p.fdelta + row0 + (T - 1);
    float* pdem = dem + (size_t)(T - 1) * ldc + c;
    const bool wr_dem = (j == 0 && c < ldc);
    float enext[FB], dnext[FB];
    ST bnext[FB], gnext[FB];
#pragma unroll
    for (int f = 0; f < FB; ++f) {
        const int nn = T - 1 - f;
        const bool ok = valid && nn > 0;
        enext[f] = (valid && nn >= 0) ? __ldg(pe - f * ldc) : 0.0f;
        bnext[f] = ok ? __ldcg(pb - f * ldc) : (ST)0;
        gnext[f] = ok ? __ldcg(pg - f * ldc) : (ST)0;
        dnext[f] = (nn >= 1) ? __ldcg(pd - f) : 0.0f;
    }

#pragma unroll 1
    for (int n0 = T - 1; n0 >= 0; n0 -= FB) {
        float ecurv[FB], dcurv[FB];
        ST bcurv[FB], gcurv[FB];
#pragma unroll
        for (int f = 0; f < FB; ++f) {
            ecurv[f] = enext[f];
            bcurv[f] = bnext[f];
            gcurv[f] = gnext[f];
            dcurv[f] = dnext[f];
        }
        pe -= FB * ldc;
        pb -= FB * ldc;
        pg -= FB * ldc;
        pd -= FB;
#pragma unroll
        for (int f = 0; f < FB; ++f) {
            const int nn = n0 - FB - f;
            const bool ok = valid && nn > 0;
            enext[f] = (valid && nn >= 0) ? __ldg(pe - f * ldc) : 0.0f;
            bnext[f] = ok ? __ldcg(pb - f * ldc) : (ST)0;
            gnext[f] = ok ? __ldcg(pg - f * ldc) : (ST)0;
            dnext[f] = (nn >= 1) ? __ldcg(pd - f) : 0.0f;
        }
#pragma unroll
        for (int f = 0; f < FB; ++f) {
        const int n = n0 - f;
        if (n < 0) break;
        const float ecur = ecurv[f], gm_n = dcurv[f];
        const ST bcur = bcurv[f], gcur = gcurv[f];
        const ST e = (ST)ecur * (ST)SC;
        // ---- phase 1: zeta^[n][c] and the length counts ------------------------------------------
        const ST rho = smax(zprev - (ST)gm_next, eta + (ST)ln_first);
        const bool dead = rho < (ST)LIN_DEAD;
        const float eo = (float)((eprev - rho) + (rref - (ST)gm_next));
        const float fac = dead ? 0.0f : ex2(fminf(eo, 100.0f));
        float carry = 0.0f;
        if (S > 1) carry = __shfl_up_sync(FULL, Q[KR - 1], CPW);
#pragma unroll
        for (int i = KR - 1; i > 0; --i) Q[i] = Q[i - 1] * fac;
        Q[0] = (j == 0) ? (dead ? 0.0f : ex2(fminf((float)(eta - rho), 100.0f))) : carry * fac;
        rref = rho;
        eprev = e;
        const ST betan = (n == 0) ? (ST)init_c : bcur;
        // forward + backward exponent: masked scores (-1e9) cancel between the two directions -> double
        const double fb2 = (double)betan + (double)e + (double)rho;
        const float coef0 = valid ? w * ex2(fminf((float)fb2, 100.0f)) : 0.0f;
        float sp[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
        for (int i = 0; i + 1 < KR; i += 2) {  // packed FMAs (FFMA2)
            if ((i >> 1) & 1)
                ffma2(sp[2], sp[3], Q[i], Q[i + 1], pl[i], pl[i + 1]);
            else
                ffma2(sp[0], sp[1], Q[i], Q[i + 1], pl[i], pl[i + 1]);
            ffma2(El[i], El[i + 1], Q[i], Q[i + 1], coef0, coef0);
        }
        if (KR & 1) {
            sp[0] = fmaf(Q[KR - 1], pl[KR - 1], sp[0]);
            El[KR - 1] = fmaf(Q[KR - 1], coef0, El[KR - 1]);
        }
        float s = slice_sum<S>((sp[0] + sp[1]) + (sp[2] + sp[3]));
        const bool live = valid && !dead;
        const bool bad_tiny = live && !(s > LIN_TINY);
        const float Sc = dead ? 0.0f : coef0 * s;
        s = dead ? 1.0f : fmaxf(s, 1.0e-37f);
        const float lg = lg2(s);
        const bool bad_trk = live && trk.step(-lg, lnmax, L) > LIN_RELEVANT;
        why |= (bad_tiny ? 2 : 0) | (bad_trk ? 4 : 0);
        const ST zeta = valid ? (e + rho) + (ST)lg : (ST)NEG;
        zprev = zeta;
        // ---- occupancy of frame n -------------------------------------------------------------
        {
            const float y = (Fprev - Sprev) - comp;
            const float tsum = occ + y;
            comp = (tsum - occ) - y;
            occ = tsum;
        }
        if (wr_dem) *pdem = valid ? occ : 0.0f;
        pdem -= ldc;
        Sprev = Sc;
        if (n == 0) {
            S0 = Sc;
            break;
        }
        // ---- phase 2: eta~[n][c1] = (+)_c2 trans[c2,c1] + zeta^[n][c2] - gm_n; transition counts ------
        if constexpr (TM == 2) {
            ST v[SPW];
            ST m2 = NEG;
#pragma unroll
            for (int q = 0; q < SPW; ++q) {
                v[q] = __shfl_sync(FULL, zeta, sidx[q]) + (ST)sval[q];
                m2 = smax(m2, v[q]);
            }
            const float coef2 = valid ? w * ex2((float)((gcur + m2) - (ST)gm_n)) : 0.0f;
            float s2 = 0.0f;
#pragma unroll
            for (int q = 0; q < SPW; ++q) {
                const float pq = ex2((float)(v[q] - m2));
                s2 += pq;
                Es[q] = fmaf(pq, coef2, Es[q]);
            }
            eta = valid ? (m2 - (ST)gm_n) + (ST)lg2(s2) : (ST)NEG;
            Fprev = coef2 * s2;
        } else {
            // dense transitions: float state only (the XP instantiations are launched for sparse lists only)
            float* zs = zet_s + (n & 1) * CPW;
            if (j == 0) zs[cl] = valid ? (float)zeta : NEG;
            __syncwarp();
            float m2 = NEG;
#pragma unroll
            for (int i = 0; i < CRR; ++i) m2 = fmaxf(m2, zs[j * CRR + i] + tr[i]);
            m2 = slice_max<S>(m2);
            const float coef2 = valid ? w * ex2(((float)gcur + m2) - gm_n) : 0.0f;
            float s2 = 0.0f;
#pragma unroll
            for (int i = 0; i < CRR; ++i) {
                const float pq = ex2(zs[j * CRR + i] + tr[i] - m2);
                s2 += pq;
                Et[i] = fmaf(pq, coef2, Et[i]);
            }
            s2 = slice_sum<S>(s2);
            eta = valid ? (ST)((m2 - gm_n) + lg2(s2)) : (ST)NEG;
            Fprev = coef2 * s2;
        }
        gm_next = gm_n;
        }
    }

    // a NaN anywhere (it cannot happen on unflagged videos) must also send the video to the log-domain kernel
    why |= (valid && !(S0 == S0)) ? 16 : 0;
    why = __reduce_or_sync(FULL, why);
    const bool flagged = why != 0;
    if (lane == 0) p.bflag[b] = flagged ? (float)(1 + why) : 0.0f;
    if (flagged) return;
    // ---- flush the per-video counts ------------------------------------------------------------
    // P(first segment has class c) is a distribution over c: normalise it by its own sum instead of by the forward
    // pass's log Z -- after T frames the two directions' float roundings differ by ~1e-7 * sqrt(T)..T, and frame 0 is
    // where the whole difference would show
    {
        const float tot = warp_sum(owner ? S0 : 0.0f);
        if (owner && tot != 0.0f) atomicAdd(p.d_init + c, S0 * (w / tot));
    }
    if (valid) {
#pragma unroll
        for (int i = 0; i < KR; ++i) {
            const int k = j * KR + i + 1;
            if (k <= L) atomicAdd(p.d_len + (size_t)k * C + c, El[i] * pl[i]);
        }
        if constexpr (TM == 0) {
#pragma unroll
            for (int i = 0; i < CRR; ++i) {
                const int c2 = j * CRR + i;
                if (c2 < C) atomicAdd(p.d_trans + (size_t)c2 * C + c, Et[i]);
            }
        } else {
            if (j == 0) {
#pragma unroll
                for (int q = 0; q < SPW; ++q)
                    if (p.trans_succ[c * SPW + q] >= 0) atomicAdd(p.d_trans + (size_t)sidx[q] * C + c, Es[q]);
            }
        }
    }
}

template <bool XP, int KR, int S, int TM>
// f64 state: 173 registers unbounded = two CTAs per SM; bounded to three it compiles to 162 without spills
__global__ void __launch_bounds__(128, (TM == 2 && KR <= 20) ? (XP ? 3 : 4) : 1) dp_lin_backward_kernel(const DpParams p) {
    dp_lin_backward_kernel_body<XP, KR, S, TM>(p, blockIdx.x);
}
// the same body compiled for three CTAs per SM (<= 168 registers instead of 128: ~10 % fewer instructions per frame, no
// constant reloads in the loop) -- for launches of at most 3 x SMs CTAs, which are resident at once either way
template <bool XP, int KR, int S, int TM>
__global__ void __launch_bounds__(128, 3) dp_lin_backward_kernel_wide(const DpParams p) {
    dp_lin_backward_kernel_body<XP, KR, S, TM>(p, blockIdx.x);
}
template <bool XP, int KR, int S, int TM>
__global__ void __launch_bounds__(128, (TM == 2 && KR <= 20) ? (XP ? 3 : 4) : 1) dp_lin_backward_kernel_grouped(const __grid_constant__ DpGroup g) {
    int local;
    const int t = group_find(g, blockIdx.x, local);
    dp_lin_backward_kernel_body<XP, KR, S, TM>(g.t[t], local);
}

// ---------------------------------------------------------------------------------------------
// host side: which shapes take the linear-window path
// ---------------------------------------------------------------------------------------------
// variants of kVariants that have a linear-window instantiation (one warp per video, lengths in registers)
static inline bool lin_variant(int v) { return v == 0 || v == 1 || v == 2 || v == 3 || v == 6 || v == 10; }

static inline bool lin_eligible(const RegChoice& ch, bool xp) {
    // extended-precision state (narration penalties): sparse transition lists only, the two CrossTask variants
    if (xp) return ch.v >= 0 && ch.W == 1 && (ch.v == 0 || ch.v == 1) && ch.tm == 2;
    return ch.v >= 0 && ch.W == 1 && kVariants[ch.v].lreg && lin_variant(ch.v) && (ch.tm == 0 || ch.tm == 2);
}

template <int MODE, bool XP, int KR, int S, int TM>
static int launch_lin_one(const DpParams& p, cudaStream_t st) {
    constexpr int VPB = 4;
    const int blocks = (p.B + VPB - 1) / VPB;
    const size_t smem = VPB * (2 * (32 / S) + 2) * sizeof(float);
    if constexpr (MODE == 1) {
        dp_lin_forward_kernel<XP, KR, S, TM><<<blocks, VPB * 32, smem, st>>>(p);
    } else if constexpr (TM == 2 && KR <= 20 && !XP) {
        if (blocks <= 3 * dp_num_sms())
            dp_lin_backward_kernel_wide<XP, KR, S, TM><<<blocks, VPB * 32, smem, st>>>(p);
        else
            dp_lin_backward_kernel<XP, KR, S, TM><<<blocks, VPB * 32, smem, st>>>(p);
    } else {
        dp_lin_backward_kernel<XP, KR, S, TM><<<blocks, VPB * 32, smem, st>>>(p);
    }
    return check_launch("dp_lin kernel");
}

template <int MODE, int KR, int S>
static int launch_lin_tm(const DpParams& p, int tm, cudaStream_t st) {
    return tm == 2 ? launch_lin_one<MODE, false, KR, S, 2>(p, st) : launch_lin_one<MODE, false, KR, S, 0>(p, st);
}

template <int MODE>
static int launch_lin(const DpParams& p, const RegChoice& ch, cudaStream_t st) {
    if (p.xp) {  // lin_eligible: sparse lists, variants 0 and 1
        if (ch.v == 0) return launch_lin_one<MODE, true, 10, 2, 2>(p, st);
        if (ch.v == 1) return launch_lin_one<MODE, true, 20, 1, 2>(p, st);
        set_error("no extended-precision linear-window DP variant for this shape");
        return -2;
    }
    switch (ch.v) {
        case 0: return launch_lin_tm<MODE, 10, 2>(p, ch.tm, st);
        case 1: return launch_lin_tm<MODE, 20, 1>(p, ch.tm, st);
        case 2: return launch_lin_tm<MODE, 13, 4>(p, ch.tm, st);
        case 3: return launch_lin_tm<MODE, 25, 2>(p, ch.tm, st);
        case 6: return launch_lin_tm<MODE, 32, 1>(p, ch.tm, st);
        case 10: return launch_lin_tm<MODE, 50, 2>(p, ch.tm, st);
    }
    set_error("no linear-window DP variant for this shape");
    return -2;
}

}  // namespace hsmm
