// Two videos per warp: the chain-constrained CrossTask shapes with at most 16 classes (sm_100a).
//
// The one-warp-per-video kernels (hsmm_dp_lin.cuh, hsmm_dp_vit2.cuh) spend ~120 of their ~150-200 warp instructions per
// frame OUTSIDE the span window -- transition phase, reference / normaliser bookkeeping, flag tracking, loads, stores,
// loop control -- and that part costs the same whether 7 or 32 lanes carry a class.  More than half of the CrossTask tasks
// have C = 2s+1 <= 16 classes (data/crosstask.py:437-448 with the step counts of the 18 primary tasks): here a warp
// carries TWO videos, lanes 0-15 and 16-31, each lane one class and the whole window (KR = 20 >= L lengths, no k-slices),
// so the fixed part is paid once per PAIR of frames.  Videos are paired in processing order (longest first), i.e. with a
// neighbour of nearly the same length; the shorter one's lanes idle (predicated off) for the last few frames.
//
// Same numerics, saved-tensor format, flag protocol (fflag / bflag / vflag + the log-domain kernels launched behind with
// DpParams::only_flagged) and outputs as the kernels these replace; sparse transition lists only, float state only.
#pragma once
#include "hsmm_dp_lin.cuh"

namespace hsmm {

__device__ __forceinline__ float half_max(float v) {  // over the 16 lanes of a half warp
#pragma unroll
    for (int off = 8; off > 0; off >>= 1) v = fmaxf(v, __shfl_xor_sync(FULL, v, off));
    return v;
}
__device__ __forceinline__ float half_sum(float v) {
#pragma unroll
    for (int off = 8; off > 0; off >>= 1) v += __shfl_xor_sync(FULL, v, off);
    return v;
}
// per-half maximum with two whole-warp redux instructions (cheaper than four shuffle + max steps)
__device__ __forceinline__ float half_max_redux(float v, int sub) {
    const float a = warp_max_redux(sub == 0 ? v : NEG);
    const float b = warp_max_redux(sub == 1 ? v : NEG);
    return sub ? b : a;
}

struct PairLane {
    int sub, c, vidx, b, T, Tw;
    bool have, valid;
    unsigned hmask;
    __device__ __forceinline__ void init(const DpParams& p, const int bid) {
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        sub = lane >> 4;
        c = lane & 15;
        hmask = sub ? 0xffff0000u : 0x0000ffffu;
        vidx = (bid * (blockDim.x >> 5) + warp) * 2 + sub;
        have = vidx < p.B;
        b = have ? (p.order ? p.order[vidx] : vidx) : 0;
        T = have ? p.lengths[b] : 0;
        Tw = max(T, __shfl_xor_sync(FULL, T, 16));
        valid = have && c < p.C;
    }
};

// ---------------------------------------------------------------------------------------------
// forward (log-partition), linear window
// ---------------------------------------------------------------------------------------------
template <int KR>
__device__ __forceinline__ void dp_pair_forward_kernel_body(const DpParams& p, const int bid) {
    PairLane q;
    q.init(p, bid);
    if (!__any_sync(FULL, q.have)) return;
    const int C = p.C, L = p.L, ldc = p.ldc, Tmax = p.Tmax;
    const int sub = q.sub, c = q.c, b = q.b, T = q.T;
    const bool valid = q.valid;
    const float SC = LOG2E;

    float pl[KR];
    float lnmax = NEG, lnmin = -NEG, maxstep = NEG;
    {
        float prev = NEG;
#pragma unroll
        for (int i = 0; i < KR; ++i) {
            const int k = i + 1;
            const bool use = valid && k <= L;
            const float v = use ? p.lenp[(size_t)k * C + c] * SC : NEG;
            pl[i] = use ? ex2(v) : 0.0f;
            if (use) {
                lnmax = fmaxf(lnmax, v);
                lnmin = fminf(lnmin, v);
            }
            if (k >= 2 && use) maxstep = fmaxf(maxstep, v - prev);
            prev = v;
        }
        if (maxstep < -1.0e29f) maxstep = 0.0f;
    }
    bool bad = valid && (lnmin < -100.0f || fmaxf(maxstep, 0.0f) - lnmin > 110.0f);
    if (__any_sync(FULL, bad)) {  // unsuitable length table (one task: both videos): known before the first frame
        if (c == 0 && q.have) p.fflag[b] = 2.0f;
        return;
    }
    const float ln_first = valid ? p.lenp[(size_t)C + c] * SC : NEG;
    const float init_c = valid ? p.init[c] * SC : NEG;
    const float endc = valid ? (p.end ? p.end[(size_t)b * C + c] : 0.0f) * SC : NEG;

    int pidx[SPW];
    float pval[SPW];
#pragma unroll
    for (int s = 0; s < SPW; ++s) {
        const int c1 = valid ? p.trans_pred[c * SPW + s] : -1;
        pidx[s] = (c1 >= 0 ? c1 : 0) + 16 * sub;
        pval[s] = c1 >= 0 ? p.trans[(size_t)c * C + c1] * SC : NEG;
    }

    const float* em_b = p.em + (size_t)b * Tmax * ldc;
    const size_t row0 = (size_t)b * (Tmax + 1);
    float* const fbeta = reinterpret_cast<float*>(p.fbeta);
    float* const fgamma = reinterpret_cast<float*>(p.fgamma);
    if (valid) fbeta[row0 * ldc + c] = init_c;

    float P[KR];
#pragma unroll
    for (int i = 0; i < KR; ++i) P[i] = 0.0f;
    float beta = init_c, gprev = NEG, rref = 0.0f, eprev = 0.0f, gmprev = 0.0f;
    double nu = 0.0, nufin = 0.0;
    float nu4 = 0.0f, gfin = NEG;
    LinTracker trk;
    trk.init();

    const float* ep = em_b + c;
    float enext[F];
#pragma unroll
    for (int f = 0; f < F; ++f) enext[f] = (valid && f < T) ? __ldg(ep + f * ldc) : 0.0f;
    ep += F * ldc;
    float* gout = fgamma + (row0 + 1) * ldc + c;
    float* bout = fbeta + (row0 + 1) * ldc + c;
    float* dout = p.fdelta + row0 + 1;

#pragma unroll 1
    for (int n0 = 1; n0 <= q.Tw; n0 += F) {
        float ecur[F];
#pragma unroll
        for (int f = 0; f < F; ++f) ecur[f] = enext[f];
#pragma unroll
        for (int f = 0; f < F; ++f) enext[f] = (valid && n0 - 1 + F + f < T) ? __ldg(ep + f * ldc) : 0.0f;
        ep += F * ldc;
#pragma unroll
        for (int f = 0; f < F; ++f) {
            const int n = n0 + f;
            if (n > q.Tw) break;
            const bool in = n <= T;  // this half's video still runs (its lanes compute on, unobserved, afterwards)
            const float e = ecur[f] * SC;
            nu4 += gmprev;
            const float rho = fmaxf(gprev - gmprev, beta + ln_first);
            const bool dead = rho < LIN_DEAD;
            const float eo = (eprev - rho) + (rref - gmprev);
            const float fac = dead ? 0.0f : ex2(fminf(eo, 100.0f));
#pragma unroll
            for (int i = KR - 1; i > 0; --i) P[i] = P[i - 1] * fac;
            P[0] = dead ? 0.0f : ex2(fminf(beta - rho, 100.0f));
            rref = rho;
            eprev = e;
            float sp[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
            for (int i = 0; i + 1 < KR; i += 2) {
                if ((i >> 1) & 1)
                    ffma2(sp[2], sp[3], P[i], P[i + 1], pl[i], pl[i + 1]);
                else
                    ffma2(sp[0], sp[1], P[i], P[i + 1], pl[i], pl[i + 1]);
            }
            if (KR & 1) sp[0] = fmaf(P[KR - 1], pl[KR - 1], sp[0]);
            float s = (sp[0] + sp[1]) + (sp[2] + sp[3]);
            const bool live = valid && !dead && in;
            bad |= live && !(s > LIN_TINY);
            s = dead ? 1.0f : fmaxf(s, 1.0e-37f);
            const float lg = lg2(s);
            bad |= live && trk.step(-lg, lnmax, L) > LIN_RELEVANT;
            const float gamma = valid ? (e + rho) + lg : NEG;
            gprev = gamma;
            const float gm = half_max_redux(gamma, sub);
            if (valid && in) *gout = gamma;
            gout += ldc;
            if (n == T) {
                gfin = gamma;
                nufin = nu + (double)nu4;
            }
            if (c == 0 && q.have && n < T) *dout = gm;
            ++dout;
            gmprev = gm;
            float v[SPW];
            float m = NEG;
#pragma unroll
            for (int s2 = 0; s2 < SPW; ++s2) {
                v[s2] = __shfl_sync(FULL, gamma, pidx[s2]) + pval[s2];
                m = fmaxf(m, v[s2]);
            }
            float s2sum = 0.0f;
#pragma unroll
            for (int s2 = 0; s2 < SPW; ++s2) s2sum += ex2(v[s2] - m);
            beta = valid ? (m - gm) + lg2(s2sum) : NEG;
            if (valid && n < T) *bout = beta;
            bout += ldc;
        }
        nu += (double)nu4;
        nu4 = 0.0f;
    }

    const float vfin = valid ? gfin + endc : NEG;
    const float m = half_max(vfin);
    const float sfin = half_sum(valid ? ex2(vfin - m) : 0.0f);
    const float final_v = m + lg2(sfin);
    const double total = (nufin + (double)final_v) * LN2;
    bad |= q.have && !(total > (double)DEGENERATE);
    const bool flagged = (__ballot_sync(FULL, bad) & q.hmask) != 0u;
    if (c == 0 && q.have) {
        p.logz2[b] = (double)final_v;
        p.fflag[b] = flagged ? 2.0f : 0.0f;
        p.logz[b] = total + (p.offset ? p.offset[b] : 0.0);
    }
}

template <int KR>
__global__ void __launch_bounds__(128) dp_pair_forward_kernel(const DpParams p) {
    dp_pair_forward_kernel_body<KR>(p, blockIdx.x);
}
template <int KR>
__global__ void __launch_bounds__(128) dp_pair_forward_kernel_grouped(const __grid_constant__ DpGroup g) {
    int local;
    const int t = group_find(g, blockIdx.x, local);
    dp_pair_forward_kernel_body<KR>(g.t[t], local);
}

// ---------------------------------------------------------------------------------------------
// backward (expected counts), linear window
// ---------------------------------------------------------------------------------------------
template <int KR>
__device__ __forceinline__ void dp_pair_backward_kernel_body(const DpParams& p, const int bid) {
    constexpr int FB = 2;
    PairLane q;
    q.init(p, bid);
    if (!__any_sync(FULL, q.have)) return;
    const int C = p.C, L = p.L, ldc = p.ldc, Tmax = p.Tmax;
    const int sub = q.sub, c = q.c, b = q.b, T = q.T;
    const float SC = LOG2E;
    // a video whose forward pass fell back to the dense matrix is left to the log-domain kernel (as in dp_lin_backward)
    const bool dense_fwd = q.have && (((int)p.fflag[b]) & 3);
    if (dense_fwd && c == 0) p.bflag[b] = 1.0f;
    const bool have = q.have && !dense_fwd;
    const bool valid = have && c < C;

    float Q[KR], El[KR], pl[KR];
    float lnmax = NEG, lnmin = -NEG, maxstep = NEG;
    {
        float prev = NEG;
#pragma unroll
        for (int i = 0; i < KR; ++i) {
            const int k = i + 1;
            const bool use = valid && k <= L;
            const float v = use ? p.lenp[(size_t)k * C + c] * SC : NEG;
            pl[i] = use ? ex2(v) : 0.0f;
            if (use) {
                lnmax = fmaxf(lnmax, v);
                lnmin = fminf(lnmin, v);
            }
            if (k >= 2 && use) maxstep = fmaxf(maxstep, v - prev);
            prev = v;
            Q[i] = 0.0f;
            El[i] = 0.0f;
        }
        if (maxstep < -1.0e29f) maxstep = 0.0f;
    }
    int why = (valid && (lnmin < -100.0f || fmaxf(maxstep, 0.0f) - lnmin > 110.0f)) ? 8 : 0;
    {
        // an unsuitable length table is known before the first frame: flag the video now; leave when neither half has
        // a video left to process (cf. the forward kernel)
        const bool half_flag = (__ballot_sync(FULL, why != 0) & q.hmask) != 0u;
        if (half_flag && c == 0 && have) p.bflag[b] = 9.0f;
        if (__ballot_sync(FULL, valid && !half_flag) == 0u) return;
    }
    const float ln_first = valid ? p.lenp[(size_t)C + c] * SC : NEG;

    int sidx[SPW];
    float sval[SPW], Es[SPW];
#pragma unroll
    for (int s = 0; s < SPW; ++s) {
        const int c2 = valid ? p.trans_succ[c * SPW + s] : -1;
        sidx[s] = (c2 >= 0 ? c2 : 0) + 16 * sub;
        sval[s] = c2 >= 0 ? p.trans[(size_t)c2 * C + c] * SC : NEG;
        Es[s] = 0.0f;
    }
    const float endc = valid ? (p.end ? p.end[(size_t)b * C + c] : 0.0f) * SC : NEG;
    const float lzrel = have ? (float)p.logz2[b] : 0.0f;
    const float w = have ? p.grad[b] : 0.0f;
    const float init_c = valid ? p.init[c] * SC : NEG;

    const size_t row0 = (size_t)b * (Tmax + 1);
    const float* fg0 = reinterpret_cast<const float*>(p.fgamma) + row0 * ldc + c;
    float* dem = p.d_em + (size_t)b * Tmax * ldc;

    float eta = valid ? endc - lzrel : NEG;
    float zprev = NEG, rref = 0.0f, eprev = 0.0f;
    float occ = 0.0f, comp = 0.0f;
    float Fprev = valid ? w * ex2(__ldcg(fg0 + (size_t)T * ldc) + endc - lzrel) : 0.0f;
    float Sprev = 0.0f, gm_next = 0.0f, S0 = 0.0f;
    LinTracker trk;
    trk.init();

    if (have)
        for (int i = T * ldc + c; i < Tmax * ldc; i += 16) dem[i] = 0.0f;  // frames beyond the video

    // this half walks its OWN video from its own last frame: frame n = T - 1 - it
    const float* pe = p.em + (size_t)b * Tmax * ldc + (size_t)(T - 1) * ldc + c;
    const float* pb = reinterpret_cast<const float*>(p.fbeta) + row0 * ldc + (size_t)(T - 1) * ldc + c;
    const float* pg = fg0 + (size_t)(T - 1) * ldc;
    const float* pd = p.fdelta + row0 + (T - 1);
    float* pdem = dem + (size_t)(T - 1) * ldc + c;
    const bool wr_dem = have && c < ldc;
    float enext[FB], bnext[FB], gnext[FB], dnext[FB];
#pragma unroll
    for (int f = 0; f < FB; ++f) {
        const int nn = T - 1 - f;
        const bool ok = valid && nn > 0;
        enext[f] = (valid && nn >= 0) ? __ldg(pe - f * ldc) : 0.0f;
        bnext[f] = ok ? __ldcg(pb - f * ldc) : 0.0f;
        gnext[f] = ok ? __ldcg(pg - f * ldc) : 0.0f;
        dnext[f] = (have && nn >= 1) ? __ldcg(pd - f) : 0.0f;
    }

#pragma unroll 1
    for (int it0 = 0; it0 < q.Tw; it0 += FB) {
        float ecurv[FB], bcurv[FB], gcurv[FB], dcurv[FB];
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
            const int nn = T - 1 - it0 - FB - f;
            const bool ok = valid && nn > 0;
            enext[f] = (valid && nn >= 0) ? __ldg(pe - f * ldc) : 0.0f;
            bnext[f] = ok ? __ldcg(pb - f * ldc) : 0.0f;
            gnext[f] = ok ? __ldcg(pg - f * ldc) : 0.0f;
            dnext[f] = (have && nn >= 1) ? __ldcg(pd - f) : 0.0f;
        }
#pragma unroll
        for (int f = 0; f < FB; ++f) {
            if (it0 + f >= q.Tw) break;
            const int n = T - 1 - it0 - f;
            const bool in = n >= 0;  // this half's video still runs
            const float bcur = bcurv[f], gcur = gcurv[f], gm_n = dcurv[f];
            const float e = ecurv[f] * SC;
            const float rho = fmaxf(zprev - gm_next, eta + ln_first);
            const bool dead = rho < LIN_DEAD;
            const float eo = (eprev - rho) + (rref - gm_next);
            const float fac = dead ? 0.0f : ex2(fminf(eo, 100.0f));
#pragma unroll
            for (int i = KR - 1; i > 0; --i) Q[i] = Q[i - 1] * fac;
            Q[0] = dead ? 0.0f : ex2(fminf(eta - rho, 100.0f));
            rref = rho;
            eprev = e;
            const float betan = (n == 0) ? init_c : bcur;
            const double fb2 = (double)betan + (double)e + (double)rho;
            const float coef0 = (valid && in) ? w * ex2(fminf((float)fb2, 100.0f)) : 0.0f;
            float sp[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
            for (int i = 0; i + 1 < KR; i += 2) {
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
            float s = (sp[0] + sp[1]) + (sp[2] + sp[3]);
            const bool live = valid && !dead && in;
            const bool bad_tiny = live && !(s > LIN_TINY);
            const float Sc = dead ? 0.0f : coef0 * s;
            s = dead ? 1.0f : fmaxf(s, 1.0e-37f);
            const float lg = lg2(s);
            const bool bad_trk = live && trk.step(-lg, lnmax, L) > LIN_RELEVANT;
            why |= (bad_tiny ? 2 : 0) | (bad_trk ? 4 : 0);
            const float zeta = valid ? (e + rho) + lg : NEG;
            zprev = zeta;
            {
                const float y = (Fprev - Sprev) - comp;
                const float tsum = occ + y;
                comp = (tsum - occ) - y;
                occ = tsum;
            }
            if (wr_dem && in) *pdem = valid ? occ : 0.0f;
            pdem -= ldc;
            Sprev = Sc;
            if (n == 0) S0 = Sc;
            // phase 2 (for n <= 0 it runs on zeros with coef2 = 0: nothing is accumulated)
            float v[SPW];
            float m2 = NEG;
#pragma unroll
            for (int s2 = 0; s2 < SPW; ++s2) {
                v[s2] = __shfl_sync(FULL, zeta, sidx[s2]) + sval[s2];
                m2 = fmaxf(m2, v[s2]);
            }
            const float coef2 = (valid && n > 0) ? w * ex2((gcur + m2) - gm_n) : 0.0f;
            float s2sum = 0.0f;
#pragma unroll
            for (int s2 = 0; s2 < SPW; ++s2) {
                const float pq = ex2(v[s2] - m2);
                s2sum += pq;
                Es[s2] = fmaf(pq, coef2, Es[s2]);
            }
            eta = valid ? (m2 - gm_n) + lg2(s2sum) : NEG;
            Fprev = coef2 * s2sum;
            gm_next = gm_n;
        }
    }

    why |= (valid && !(S0 == S0)) ? 16 : 0;
    const unsigned wb = __ballot_sync(FULL, why != 0) & q.hmask;
    int whyh = why;
#pragma unroll
    for (int off = 8; off > 0; off >>= 1) whyh |= __shfl_xor_sync(FULL, whyh, off);
    const bool flagged = wb != 0u;
    if (c == 0 && have) p.bflag[b] = flagged ? (float)(1 + whyh) : 0.0f;
    const float tot = half_sum(valid ? S0 : 0.0f);
    if (flagged || !valid) return;
    if (tot != 0.0f) atomicAdd(p.d_init + c, S0 * (w / tot));
#pragma unroll
    for (int i = 0; i < KR; ++i) {
        const int k = i + 1;
        if (k <= L) atomicAdd(p.d_len + (size_t)k * C + c, El[i] * pl[i]);
    }
#pragma unroll
    for (int s = 0; s < SPW; ++s)
        if (p.trans_succ[c * SPW + s] >= 0) atomicAdd(p.d_trans + (size_t)(sidx[s] - 16 * sub) * C + c, Es[s]);
}

template <int KR>
__global__ void __launch_bounds__(128, 4) dp_pair_backward_kernel(const DpParams p) {
    dp_pair_backward_kernel_body<KR>(p, blockIdx.x);
}
template <int KR>
__global__ void __launch_bounds__(128, 4) dp_pair_backward_kernel_grouped(const __grid_constant__ DpGroup g) {
    int local;
    const int t = group_find(g, blockIdx.x, local);
    dp_pair_backward_kernel_body<KR>(g.t[t], local);
}

// ---------------------------------------------------------------------------------------------
// Viterbi with deferred length arg-max (cf. dp_vit2_kernel)
// ---------------------------------------------------------------------------------------------
template <int KR>
__device__ __forceinline__ void dp_pair_vit_kernel_body(const DpParams& p, const int bid) {
    PairLane q;
    q.init(p, bid);
    if (!__any_sync(FULL, q.have)) return;
    const int lane = threadIdx.x & 31;
    const int C = p.C, L = p.L, ldc = p.ldc, Tmax = p.Tmax;
    const int sub = q.sub, c = q.c, b = q.b, T = q.T;
    const bool valid = q.valid;

    float ln[KR];
#pragma unroll
    for (int i = 0; i < KR; ++i) {
        const int k = i + 1;
        ln[i] = (valid && k <= L) ? p.lenp[(size_t)k * C + c] : NEG;
    }
    const float init_c = valid ? p.init[c] : NEG;
    const float endc = valid ? (p.end ? p.end[(size_t)b * C + c] : 0.0f) : NEG;
    int pidx[SPW];
    float pval[SPW];
#pragma unroll
    for (int s = 0; s < SPW; ++s) {
        const int c1 = valid ? p.trans_pred[c * SPW + s] : -1;
        pidx[s] = (c1 >= 0 ? c1 : 0) + 16 * sub;
        pval[s] = c1 >= 0 ? p.trans[(size_t)c * C + c1] : NEG;
    }

    const float* em_b = p.em + (size_t)b * Tmax * ldc;
    const size_t row0 = (size_t)b * (Tmax + 1);
    float* const vbeta = p.vbeta + row0 * ldc;
    uint32_t* const vpred = p.vpred + row0 * ldc;
    float* const vdelta = p.vdelta + row0;
    if (valid) vbeta[c] = init_c;

    float A[KR];
#pragma unroll
    for (int i = 0; i < KR; ++i) A[i] = NEG;
    float beta = init_c, gmprev = 0.0f, nu4 = 0.0f, gfin = NEG;
    double nu = 0.0, nufin = 0.0;

    const float* ep = em_b + c;
    float enext[F];
#pragma unroll
    for (int f = 0; f < F; ++f) enext[f] = (valid && f < T) ? __ldg(ep + f * ldc) : 0.0f;
    ep += F * ldc;
    float* bout = vbeta + ldc + c;
    uint32_t* pout = vpred + ldc + c;
    float* dout = vdelta + 1;

#pragma unroll 1
    for (int n0 = 1; n0 <= q.Tw; n0 += F) {
        float ecur[F];
#pragma unroll
        for (int f = 0; f < F; ++f) ecur[f] = enext[f];
#pragma unroll
        for (int f = 0; f < F; ++f) enext[f] = (valid && n0 - 1 + F + f < T) ? __ldg(ep + f * ldc) : 0.0f;
        ep += F * ldc;
#pragma unroll
        for (int f = 0; f < F; ++f) {
            const int n = n0 + f;
            if (n > q.Tw) break;
            const float e = ecur[f];
            nu4 += gmprev;
            const float eo = e - gmprev;
#pragma unroll
            for (int i = KR - 1; i > 0; --i) A[i] = A[i - 1] + eo;
            A[0] = beta + e;
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
            const float gamma = valid ? fmaxf(m0, m1) : NEG;
            const float gm = half_max_redux(gamma, sub);
            if (n == T) {
                gfin = gamma;
                nufin = nu + (double)nu4;
            }
            if (c == 0 && q.have && n < T) *dout = gm;
            ++dout;
            gmprev = gm;
            float best = NEG;
            int bc = 0;
#pragma unroll
            for (int s = 0; s < SPW; ++s) {
                const float v = __shfl_sync(FULL, gamma, pidx[s]) + pval[s];
                if (v > best || s == 0) {
                    best = v;
                    bc = pidx[s] - 16 * sub;
                }
            }
            beta = valid ? best - gm : NEG;
            if (valid && n < T) {
                *bout = beta;
                *pout = (uint32_t)bc;
            }
            bout += ldc;
            pout += ldc;
        }
        nu += (double)nu4;
        nu4 = 0.0f;
    }

    // ---- termination per half: best class at T (ties to the smaller class) ------------------------------------------
    const float fv = valid ? gfin + endc : NEG;
    const float final_h = half_max(fv);
    const unsigned fmask = __ballot_sync(FULL, valid && fv == final_h) & q.hmask;
    const int cc_h = fmask ? ((__ffs(fmask) - 1) & 15) : 0;
    const double total_h = nufin + (double)final_h;
    const bool degen_h = q.have && !(total_h > (double)DEGENERATE);
    if (c == 0 && q.have) p.vflag[b] = degen_h ? 1.0f : 0.0f;
    __syncwarp();

    // ---- back-trace: the whole warp walks one video at a time (lane k-1 rebuilds the length-k candidate) ----------------
#pragma unroll 1
    for (int h = 0; h < 2; ++h) {
        const int src = 16 * h;
        const bool hv = __shfl_sync(FULL, (int)q.have, src) != 0;
        const bool dg = __shfl_sync(FULL, (int)degen_h, src) != 0;
        if (!hv || dg) continue;  // missing, or left to dp_forward_kernel<VIT> against the dense matrix
        const int bb = __shfl_sync(FULL, b, src);
        const int TT = __shfl_sync(FULL, T, src);
        int cc = __shfl_sync(FULL, cc_h, src);
        const double tot = __shfl_sync(FULL, total_h, src);
        const float* em_v = p.em + (size_t)bb * Tmax * ldc;
        const size_t r0 = (size_t)bb * (Tmax + 1);
        const float* vb = p.vbeta + r0 * ldc;
        const uint32_t* vp = p.vpred + r0 * ldc;
        const float* vd = p.vdelta + r0;
        const int eos = p.class_ids ? p.class_ids[C] : C;
        int64_t* sp = p.spans + (size_t)bb * (Tmax + 1);
        for (int i = lane; i <= Tmax; i += 32) sp[i] = (i == TT) ? (int64_t)eos : (int64_t)-1;
        int64_t* lab = p.labels ? p.labels + (size_t)bb * Tmax : nullptr;
        if (lab)
            for (int i = TT + lane; i < Tmax; i += 32) lab[i] = eos;
        if (lane == 0 && p.score) p.score[bb] = tot + (p.offset ? p.offset[bb] : 0.0);
        __syncwarp();
        int n = TT;
        while (n > 0) {
            const int kmax = L < n ? L : n;  // <= 32
            const int k = lane + 1;
            const bool act = k <= kmax;
            const int m_own = n - k + 1;
            const float e_own = act ? __ldcg(em_v + (size_t)(m_own - 1) * ldc + cc) : 0.0f;
            const float g_own = (act && m_own >= 2) ? __ldcg(vd + m_own - 1) : 0.0f;
            const float b_own = act ? __ldcg(vb + (size_t)(n - k) * ldc + cc) : NEG;
            const uint32_t p_own = act ? __ldcg(vp + (size_t)(n - k) * ldc + cc) : 0u;
            const float l_own = act ? __ldg(p.lenp + (size_t)k * C + cc) : NEG;
            const float eo_own = e_own - g_own;
            float acc = b_own + e_own;
#pragma unroll
            for (int s = 1; s < 32; ++s) {
                if (s >= kmax) break;
                const float t = __shfl_up_sync(FULL, eo_own, s);
                if (lane >= s) acc += t;
            }
            const float v = act ? acc + l_own : NEG;
            const float vm = warp_max_redux(v);
            const unsigned km = __ballot_sync(FULL, act && v == vm);
            const int kstar = km ? __ffs(km) : 1;
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
}

template <int KR>
__global__ void __launch_bounds__(128) dp_pair_vit_kernel(const DpParams p) {
    dp_pair_vit_kernel_body<KR>(p, blockIdx.x);
}
template <int KR>
__global__ void __launch_bounds__(128) dp_pair_vit_kernel_grouped(const __grid_constant__ DpGroup g) {
    int local;
    const int t = group_find(g, blockIdx.x, local);
    dp_pair_vit_kernel_body<KR>(g.t[t], local);
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
constexpr int PAIR_KR = 20;
static inline bool pair_shape_ok(int C, int L, bool sparse, bool xp) { return sparse && !xp && C <= 16 && L <= PAIR_KR; }

template <int MODE>
static int launch_pair(const DpParams& p, cudaStream_t st) {
    constexpr int WPB = 4;  // warps per CTA = 8 videos
    const int warps = (p.B + 1) / 2;
    const int blocks = (warps + WPB - 1) / WPB;
    if constexpr (MODE == 0)
        dp_pair_vit_kernel<PAIR_KR><<<blocks, WPB * 32, 0, st>>>(p);
    else if constexpr (MODE == 1)
        dp_pair_forward_kernel<PAIR_KR><<<blocks, WPB * 32, 0, st>>>(p);
    else
        dp_pair_backward_kernel<PAIR_KR><<<blocks, WPB * 32, 0, st>>>(p);
    return check_launch("dp_pair kernel");
}

}  // namespace hsmm
