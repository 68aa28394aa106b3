// General semi-Markov DP kernels: any number of classes up to 1024 and any span length (sm_100a).
//
// The register-resident kernels (hsmm_dp_reg.cuh) hold the C x L span window of a video in the registers of one
// warp group; beyond their envelope -- the backward pass at C = 48 / K = 500 or C = 133 / K >= 100, anything with
// more than ~224 classes such as `valid_classes=None` over CrossTask's 284 global classes
// (/root/reference/src/models/semimarkov/semimarkov_modules.py:604-606) -- these kernels take over, so that no shape
// the reference accepts is "unsupported".  They are the slow path by design: throughput comes from the kernels above.
//
// One CTA per video, threads = (k-slice j, class c) with the 32 lanes of a warp on 32 consecutive classes (every
// load below is coalesced over c).  The window is not kept as running sums but in prefix-sum form, in DOUBLE:
//
//     forward   gamma[n][c] = cs[n][c] + (+)_k ( bp[n-k][c] + len[k][c] ),     bp[m][c] = beta[m][c] - cs[m][c]
//     backward  zeta[n][c]  = ss[n][c] + (+)_k ( q[n+k][c]  + len[k][c] ),     q[m][c]  = eta[m][c]  - ss[m][c]
//
// (cs / ss = prefix / suffix sums of the class's emission scores), with bp / q in a ring of L+1 rows per CTA in global
// memory (L2-resident).  Double precision makes the -1e4 narration penalties and -1e9 masks harmless, so the same
// kernels serve HSMM_FLAG_F64_STATE.  The forward log-sums are two-pass (max, then sum of ex2); the backward pass is
// single-pass because every term is a POSTERIOR once the known log Z is subtracted -- sum_k 2^(q + len + beta + ss -
// logZ) = P(a class-c segment starts at n) <= 1 -- so nothing can overflow and what underflows is below 1e-38 of a count.
//
// Saved-tensor format and back-pointer format are those of hsmm_dp_reg.cuh (values relative to the running normaliser
// nu_n, increments in fdelta), so forward and backward may come from different kernel families.
#include "hsmm_common.cuh"

namespace hsmm {

constexpr double DNEG = -1.0e300;
constexpr double LOG2E_D = 1.4426950408889634;
constexpr int GEN_MAX_CTAS = 296;
constexpr int GEN_MAX_C = 1024;

struct GenPlan {
    int Cp, TPC, NT, R;
    size_t ring_elems, el_elems, etr_elems, per_cta_bytes, transT_bytes, total_bytes;
};

static GenPlan gen_plan(int C, int L) {
    GenPlan g;
    g.Cp = (C + 31) / 32 * 32;
    int tpc = 1024 / g.Cp;
    const int most = L > C ? L : C;
    if (tpc > most) tpc = most;
    if (tpc < 1) tpc = 1;
    g.TPC = tpc;
    g.NT = g.Cp * tpc;
    g.R = L + 1;
    g.ring_elems = (size_t)g.R * g.Cp;
    g.el_elems = (size_t)(L + 1) * g.Cp;
    g.etr_elems = (size_t)C * g.Cp;
    g.per_cta_bytes = g.ring_elems * sizeof(double) + (g.el_elems + g.etr_elems) * sizeof(float);
    g.per_cta_bytes = (g.per_cta_bytes + 255) / 256 * 256;
    g.transT_bytes = ((size_t)C * g.Cp * sizeof(float) + 255) / 256 * 256;
    g.total_bytes = g.transT_bytes + (size_t)GEN_MAX_CTAS * g.per_cta_bytes + 256;
    return g;
}

size_t dp_gen_scratch_bytes(int C, int L) { return gen_plan(C, L).total_bytes; }
bool dp_gen_shape_ok(int C) { return C >= 1 && C <= GEN_MAX_C; }

// transT[c1 * Cp + c2] = trans[c2, c1] * scale  (the forward pass reads the predecessors of c2 coalesced over c2)
__global__ void gen_transpose_kernel(const float* __restrict__ trans, int C, int Cp, float scale, float* __restrict__ out) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < C * Cp; i += gridDim.x * blockDim.x) {
        const int c1 = i / Cp, c2 = i - c1 * Cp;
        out[i] = c2 < C ? trans[(size_t)c2 * C + c1] * scale : NEG;
    }
}

// 2^x for a double exponent: the integer part goes into the float's exponent field exactly, only the fraction
// (|f| <= 0.5) meets the MUFU unit -- casting the whole exponent (|x| ~ 30) to float would cost 2e-6 of accuracy
__device__ __forceinline__ float ex2d(double x) {
    if (!(x > -140.0)) return 0.0f;
    x = fmin(x, 126.0);
    const double r = rint(x);
    const float f = ex2((float)(x - r));            // in [2^-0.5, 2^0.5]
    const int e = (int)r;
    if (e >= -125) return __int_as_float(__float_as_int(f) + (e << 23));
    return f * __int_as_float((e + 127 + 24) << 23) * 5.9604644775390625e-08f;  // denormal range: scale in two steps
}

template <typename ST>
struct GenSmem {
    double red_m[1024];
    float red_s[1024];
    int red_i[1024];
    double cls[GEN_MAX_C];   // gamma[n][.] (forward) / zeta[n][.] (backward)
    double aux[GEN_MAX_C];   // A_c / G_c of the backward pass
    double bc_d;
    float bc_f;
    int bc_i;
};

// ---------------------------------------------------------------------------------------------
// forward: MODE 0 = max-plus Viterbi (natural-log units), MODE 1 = log-partition (log2 units)
// ---------------------------------------------------------------------------------------------
template <int MODE, typename ST>
__global__ void __launch_bounds__(1024) dp_gen_forward_kernel(const DpParams p, const GenPlan g, char* scratch) {
    constexpr bool VIT = MODE == 0;
    __shared__ GenSmem<ST> sm;
    const int C = p.C, L = p.L, ldc = p.ldc, Tmax = p.Tmax, Cp = g.Cp, TPC = g.TPC, R = g.R;
    const int tid = threadIdx.x, c = tid % Cp, j = tid / Cp;
    const bool valid = c < C, owner = valid && j == 0;
    const double SC = VIT ? 1.0 : LOG2E_D;
    const float* transT = reinterpret_cast<const float*>(scratch);
    double* ring = reinterpret_cast<double*>(scratch + g.transT_bytes + (size_t)blockIdx.x * g.per_cta_bytes);
    ST* const fbeta = reinterpret_cast<ST*>(p.fbeta);
    ST* const fgamma = reinterpret_cast<ST*>(p.fgamma);

    for (int vid = blockIdx.x; vid < p.B; vid += gridDim.x) {
        const int b = p.order ? p.order[vid] : vid;
        const int T = p.lengths[b];
        const size_t row0 = (size_t)b * (Tmax + 1);
        const float* em_b = p.em + (size_t)b * Tmax * ldc;
        double cs = 0.0, nu = 0.0, gamma_abs = DNEG;
        int bk = 0;
        if (owner) {
            const double i2 = (double)p.init[c] * SC;
            ring[c] = i2;
            if (!VIT) fbeta[row0 * ldc + c] = (ST)i2;
        }
        if (j == 0 && !valid) sm.cls[c] = DNEG;
        __syncthreads();
        for (int n = 1; n <= T; ++n) {
            if (owner) cs += (double)__ldg(em_b + (size_t)(n - 1) * ldc + c) * SC;
            // ---- phase 1: over the span lengths -------------------------------------------------
            const int kmax = min(L, n);
            double m = DNEG;
            int mk = 0x7fffffff;
            float s = 0.0f;
            if (valid) {
                for (int k = j + 1; k <= kmax; k += TPC) {
                    const double v = ring[(size_t)((n - k) % R) * Cp + c] + (double)__ldg(p.lenp + (size_t)k * C + c) * SC;
                    if (v > m) {
                        m = v;
                        mk = k;
                    }
                }
                if (!VIT && m > DNEG) {
                    for (int k = j + 1; k <= kmax; k += TPC) {
                        const double v = ring[(size_t)((n - k) % R) * Cp + c] + (double)__ldg(p.lenp + (size_t)k * C + c) * SC;
                        s += ex2d(v - m);
                    }
                }
            }
            sm.red_m[tid] = m;
            sm.red_s[tid] = s;
            sm.red_i[tid] = mk;
            __syncthreads();
            if (owner) {
                double M = sm.red_m[c];
                int K0 = sm.red_i[c];
                for (int jj = 1; jj < TPC; ++jj) {
                    const double mm = sm.red_m[jj * Cp + c];
                    const int kk = sm.red_i[jj * Cp + c];
                    if (mm > M || (mm == M && kk < K0)) {
                        M = mm;
                        K0 = kk;
                    }
                }
                if (VIT) {
                    gamma_abs = cs + M;
                    bk = K0;
                } else {
                    float S = 0.0f;
                    for (int jj = 0; jj < TPC; ++jj) S += sm.red_s[jj * Cp + c] * ex2d(sm.red_m[jj * Cp + c] - M);
                    gamma_abs = cs + M + log2((double)S);
                }
                sm.cls[c] = gamma_abs;
            }
            __syncthreads();
            // ---- normaliser increment gm_n = max_c gamma[n][c] - nu_n (a float, as stored) ------------
            if (tid < 32) {
                double mx = DNEG;
                for (int cc = tid; cc < C; cc += 32) mx = fmax(mx, sm.cls[cc]);
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) mx = fmax(mx, __shfl_xor_sync(FULL, mx, off));
                if (tid == 0) sm.bc_f = (float)(mx - nu);
            }
            __syncthreads();
            const float gm = sm.bc_f;
            if (!VIT && owner) fgamma[(row0 + n) * ldc + c] = (ST)(gamma_abs - nu);
            if (n == T) {
                if (VIT && owner) p.bp[(row0 + n) * ldc + c] = (uint32_t)bk << 16;
                break;
            }
            if (!VIT && tid == 0) p.fdelta[row0 + n] = gm;
            const double nu_next = nu + (double)gm;
            // ---- phase 2: over the predecessor classes ------------------------------------------------
            m = DNEG;
            mk = 0x7fffffff;
            s = 0.0f;
            if (valid) {
                for (int c1 = j; c1 < C; c1 += TPC) {
                    const double v = sm.cls[c1] + (double)__ldg(transT + (size_t)c1 * Cp + c);
                    if (v > m) {
                        m = v;
                        mk = c1;
                    }
                }
                if (!VIT && m > DNEG) {
                    for (int c1 = j; c1 < C; c1 += TPC) s += ex2d(sm.cls[c1] + (double)__ldg(transT + (size_t)c1 * Cp + c) - m);
                }
            }
            sm.red_m[tid] = m;
            sm.red_s[tid] = s;
            sm.red_i[tid] = mk;
            __syncthreads();
            if (owner) {
                double M = sm.red_m[c];
                int C0 = sm.red_i[c];
                for (int jj = 1; jj < TPC; ++jj) {
                    const double mm = sm.red_m[jj * Cp + c];
                    const int kk = sm.red_i[jj * Cp + c];
                    if (mm > M || (mm == M && kk < C0)) {
                        M = mm;
                        C0 = kk;
                    }
                }
                double beta_abs;
                if (VIT) {
                    beta_abs = M;
                    p.bp[(row0 + n) * ldc + c] = ((uint32_t)bk << 16) | (uint32_t)C0;
                } else {
                    float S = 0.0f;
                    for (int jj = 0; jj < TPC; ++jj) S += sm.red_s[jj * Cp + c] * ex2d(sm.red_m[jj * Cp + c] - M);
                    beta_abs = M + log2((double)S);
                    fbeta[(row0 + n) * ldc + c] = (ST)(beta_abs - nu_next);
                }
                ring[(size_t)(n % R) * Cp + c] = beta_abs - cs;
            }
            nu = nu_next;
            __syncthreads();
        }
        // ---- termination: gamma[T][.] is in sm.cls ------------------------------------------------------
        __syncthreads();
        const float* endb = p.end ? p.end + (size_t)b * C : nullptr;
        if (tid < 32) {
            double mx = DNEG;
            int bc = 0x7fffffff;
            for (int cc = tid; cc < C; cc += 32) {
                const double v = sm.cls[cc] + (endb ? (double)endb[cc] * SC : 0.0);
                if (v > mx || bc == 0x7fffffff) {
                    mx = v;
                    bc = cc;
                }
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                const double ov = __shfl_xor_sync(FULL, mx, off);
                const int oc = __shfl_xor_sync(FULL, bc, off);
                if (ov > mx || (ov == mx && oc < bc)) {
                    mx = ov;
                    bc = oc;
                }
            }
            double fin = mx;
            if (!VIT) {
                float s2 = 0.0f;
                for (int cc = tid; cc < C; cc += 32) s2 += ex2d(sm.cls[cc] + (endb ? (double)endb[cc] * SC : 0.0) - mx);
                s2 = warp_sum(s2);
                fin = mx + log2((double)s2);
            }
            if (tid == 0) {
                sm.bc_d = fin;
                sm.bc_i = bc;
                if (!VIT) {
                    p.logz2[b] = fin - nu;
                    p.fflag[b] = 0.0f;
                    p.bflag[b] = 0.0f;
                    p.logz[b] = fin * LN2 + (p.offset ? p.offset[b] : 0.0);
                } else if (p.score) {
                    p.score[b] = fin + (p.offset ? p.offset[b] : 0.0);
                }
            }
        }
        if (VIT) {
            // prefill the outputs, then walk the back-pointers (same encoding as dp_forward_kernel<VIT>)
            const int eos = p.class_ids ? p.class_ids[C] : C;
            int64_t* sp = p.spans + (size_t)b * (Tmax + 1);
            for (int i = tid; i <= Tmax; i += blockDim.x) sp[i] = (i == T) ? (int64_t)eos : (int64_t)-1;
            int64_t* lab = p.labels ? p.labels + (size_t)b * Tmax : nullptr;
            if (lab)
                for (int i = T + tid; i < Tmax; i += blockDim.x) lab[i] = eos;
            __syncthreads();  // back-pointers, prefill and sm.bc_i are visible
            if (tid < 32) {
                int n = T, cc = sm.bc_i;
                while (n > 0) {
                    const uint32_t v = __ldcg(p.bp + (row0 + n) * ldc + cc);
                    int k = (int)(v >> 16);
                    k = k < 1 ? 1 : (k > n ? n : k);
                    const int start = n - k;
                    const int64_t cid = p.class_ids ? p.class_ids[cc] : cc;
                    if (tid == 0) sp[start] = cid;
                    if (lab)
                        for (int t = start + tid; t < n; t += 32) lab[t] = cid;
                    if (start > 0) {
                        const uint32_t u = __ldcg(p.bp + (row0 + start) * ldc + cc);
                        cc = (int)(u & 0xffffu);
                        if (cc >= C) cc = C - 1;
                    }
                    n = start;
                }
            }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// backward: expected counts from the saved forward quantities
// ---------------------------------------------------------------------------------------------
template <typename ST>
__global__ void __launch_bounds__(1024) dp_gen_backward_kernel(const DpParams p, const GenPlan g, char* scratch) {
    __shared__ GenSmem<ST> sm;
    const int C = p.C, L = p.L, ldc = p.ldc, Tmax = p.Tmax, Cp = g.Cp, TPC = g.TPC, R = g.R;
    const int tid = threadIdx.x, c = tid % Cp, j = tid / Cp;
    const bool valid = c < C, owner = valid && j == 0;
    char* mine = scratch + g.transT_bytes + (size_t)blockIdx.x * g.per_cta_bytes;
    double* ring = reinterpret_cast<double*>(mine);
    float* El = reinterpret_cast<float*>(mine + g.ring_elems * sizeof(double));
    float* Etr = El + g.el_elems;
    const ST* const fbeta = reinterpret_cast<const ST*>(p.fbeta);
    const ST* const fgamma = reinterpret_cast<const ST*>(p.fgamma);

    for (int vid = blockIdx.x; vid < p.B; vid += gridDim.x) {
        const int b = p.order ? p.order[vid] : vid;
        const int T = p.lengths[b];
        const size_t row0 = (size_t)b * (Tmax + 1);
        const float* em_b = p.em + (size_t)b * Tmax * ldc;
        float* dem = p.d_em + (size_t)b * Tmax * ldc;
        const float w = p.grad[b];
        // nu_T = sum of the stored normaliser increments (double)
        {
            double part = 0.0;
            for (int n = 1 + tid; n < T; n += blockDim.x) part += (double)p.fdelta[row0 + n];
            part = warp_sum(part);
            if ((tid & 31) == 0) sm.red_m[tid >> 5] = part;
            __syncthreads();
            if (tid == 0) {
                double tot = 0.0;
                for (int i = 0; i < (int)(blockDim.x >> 5); ++i) tot += sm.red_m[i];
                sm.bc_d = tot;
            }
            __syncthreads();
        }
        const double nuT = sm.bc_d;
        const double lz2 = nuT + p.logz2[b];
        for (size_t i = tid; i < g.el_elems + g.etr_elems; i += blockDim.x) El[i] = 0.0f;
        for (int i = T * ldc + tid; i < Tmax * ldc; i += blockDim.x) dem[i] = 0.0f;
        double ss = 0.0, occ = 0.0, Fnext = 0.0, Snext = 0.0;
        if (owner) {
            const double end2 = p.end ? (double)p.end[(size_t)b * C + c] * LOG2E_D : 0.0;
            const double gT = (double)fgamma[(row0 + T) * ldc + c] + nuT;
            Fnext = (double)w * (double)ex2d(gT + end2 - lz2);
            ring[(size_t)(T % R) * Cp + c] = end2;  // q[T] = eta[T] - ss[T]
        }
        if (j == 0 && !valid) sm.cls[c] = DNEG;
        double nu_np1 = nuT;  // nu_{n+1}
        __syncthreads();
        for (int n = T - 1; n >= 0; --n) {
            if (owner) {
                ss += (double)__ldg(em_b + (size_t)n * ldc + c) * LOG2E_D;
                const double beta_abs = (n == 0) ? (double)p.init[c] * LOG2E_D : (double)fbeta[(row0 + n) * ldc + c] + nu_np1;
                sm.aux[c] = beta_abs + ss - lz2;
            }
            __syncthreads();
            // ---- phase 1: zeta[n][c] and the length counts -------------------------------------------
            const int kmax = min(L, T - n);
            float s = 0.0f;
            if (valid) {
                const double Ac = sm.aux[c];
                for (int k = j + 1; k <= kmax; k += TPC) {
                    const double v = ring[(size_t)((n + k) % R) * Cp + c] + (double)__ldg(p.lenp + (size_t)k * C + c) * LOG2E_D;
                    const float pk = ex2d(v + Ac);
                    s += pk;
                    El[(size_t)k * Cp + c] += w * pk;
                }
            }
            sm.red_s[tid] = s;
            __syncthreads();
            if (owner) {
                float Ssum = 0.0f;
                for (int jj = 0; jj < TPC; ++jj) Ssum += sm.red_s[jj * Cp + c];
                const double Sc = (double)w * (double)Ssum;
                sm.cls[c] = Ssum > 0.0f ? (log2((double)Ssum) - sm.aux[c]) + ss : DNEG;  // zeta[n][c] = ss[n][c] + (+)_k (q + len)
                occ += Fnext - Snext;
                dem[(size_t)n * ldc + c] = (float)occ;
                Snext = Sc;
                if (n == 0) sm.red_m[c] = Sc;
            } else if (j == 0 && c < ldc) {
                dem[(size_t)n * ldc + c] = 0.0f;
            }
            if (n == 0) {
                // P(first segment has class c), normalised by its own sum (see dp_lin_backward_kernel)
                __syncthreads();
                if (owner) {
                    double tot = 0.0;
                    for (int cc = 0; cc < C; ++cc) tot += sm.red_m[cc];
                    if (tot != 0.0) atomicAdd(p.d_init + c, (float)(sm.red_m[c] * ((double)w / tot)));
                }
                break;
            }
            const double nu_n = nu_np1 - (double)p.fdelta[row0 + n];
            __syncthreads();  // zeta visible; aux free
            if (owner) sm.aux[c] = (double)fgamma[(row0 + n) * ldc + c] + nu_n - lz2;
            __syncthreads();
            // ---- phase 2: eta[n][c1] and the transition counts -----------------------------------------
            s = 0.0f;
            if (valid) {
                const double Gc = sm.aux[c];
                for (int c2 = j; c2 < C; c2 += TPC) {
                    const double v = (double)__ldg(p.trans + (size_t)c2 * C + c) * LOG2E_D + sm.cls[c2];
                    const float pq = ex2d(v + Gc);
                    s += pq;
                    Etr[(size_t)c2 * Cp + c] += w * pq;
                }
            }
            sm.red_s[tid] = s;
            __syncthreads();
            if (owner) {
                float Fsum = 0.0f;
                for (int jj = 0; jj < TPC; ++jj) Fsum += sm.red_s[jj * Cp + c];
                Fnext = (double)w * (double)Fsum;
                ring[(size_t)(n % R) * Cp + c] = Fsum > 0.0f ? (log2((double)Fsum) - sm.aux[c]) - ss : DNEG;  // q[n]
            }
            nu_np1 = nu_n;
            __syncthreads();
        }
        __syncthreads();
        // ---- flush the per-video counts -------------------------------------------------------------------
        if (valid) {
            for (int k = j + 1; k <= L; k += TPC) {
                const float v = El[(size_t)k * Cp + c];
                if (v != 0.0f) atomicAdd(p.d_len + (size_t)k * C + c, v);
            }
            for (int c2 = j; c2 < C; c2 += TPC) {
                const float v = Etr[(size_t)c2 * Cp + c];
                if (v != 0.0f) atomicAdd(p.d_trans + (size_t)c2 * C + c, v);
            }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
int dp_gen_launch(DpParams p, int mode, void* scratch, cudaStream_t st) {
    if (!dp_gen_shape_ok(p.C)) {
        set_error("general DP kernels support up to %d classes (got C=%d)", GEN_MAX_C, p.C);
        return -2;
    }
    if (!scratch) {
        set_error("general DP kernels need the scratch area of the workspace (C=%d L=%d)", p.C, p.L);
        return -1;
    }
    const GenPlan g = gen_plan(p.C, p.L);
    char* sc = reinterpret_cast<char*>(((uintptr_t)scratch + 255) / 256 * 256);
    const int grid = p.B < GEN_MAX_CTAS ? p.B : GEN_MAX_CTAS;
    if (mode != 2) {
        gen_transpose_kernel<<<(p.C * g.Cp + 255) / 256, 256, 0, st>>>(p.trans, p.C, g.Cp, mode == 0 ? 1.0f : LOG2E,
                                                                        reinterpret_cast<float*>(sc));
        int rc = check_launch("gen_transpose_kernel");
        if (rc) return rc;
    }
    if (mode == 0) {
        dp_gen_forward_kernel<0, float><<<grid, g.NT, 0, st>>>(p, g, sc);
    } else if (mode == 1) {
        if (p.xp)
            dp_gen_forward_kernel<1, double><<<grid, g.NT, 0, st>>>(p, g, sc);
        else
            dp_gen_forward_kernel<1, float><<<grid, g.NT, 0, st>>>(p, g, sc);
    } else {
        if (p.xp)
            dp_gen_backward_kernel<double><<<grid, g.NT, 0, st>>>(p, g, sc);
        else
            dp_gen_backward_kernel<float><<<grid, g.NT, 0, st>>>(p, g, sc);
    }
    return check_launch("dp_gen kernel");
}

}  // namespace hsmm
