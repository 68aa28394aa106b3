// Streaming kernels around the DP (sm_100a): emission scoring, class-weighted feature sums,
// feature moments, one-hot weights, gold-segmentation score.  All are HBM-bound passes over the
// (B, Tmax, D) feature tensor or the (B, Tmax, C) score tensor.
#include "hsmm_common.cuh"

namespace hsmm {

// ---------------------------------------------------------------------------------------------
// Emission scoring (SIMT fp32 path)
//   em'[t,c] = x_t . w_c + bias_c (+ penalty)      tile: TF frames x CB classes, D in chunks of DK
// ---------------------------------------------------------------------------------------------
constexpr int E_TF = 64;   // frames per CTA tile
constexpr int E_CB = 32;   // classes per pass
constexpr int E_DK = 64;   // feature chunk
constexpr int E_LD = E_DK + 4;
constexpr int E_THREADS = 256;

__global__ void __launch_bounds__(E_THREADS)
emission_kernel(const float* __restrict__ X, const float* __restrict__ w, const float* __restrict__ bias,
                const float* __restrict__ inv_var, const float* __restrict__ row_const_p, const float* __restrict__ penalty,
                const int32_t* __restrict__ lengths, int Tmax, int D, int C, int ldc, int cs,
                float* __restrict__ em, float* __restrict__ rowterm, double* __restrict__ offset) {
    extern __shared__ float sm[];
    float* Xs = sm;                       // [E_TF][E_LD]
    float* Ws = Xs + E_TF * E_LD;         // [E_CB][E_LD]
    float* Vs = Ws + E_CB * E_LD;         // [E_DK] inverse variances of the chunk
    float* Os = Vs + E_DK;                // [E_TF][cs] class scores of the tile
    float* Rs = Os + E_TF * cs;           // [E_TF] row terms
    __shared__ double red[E_THREADS / 32];

    const float row_const = __ldg(row_const_p);
    const int b = blockIdx.y;
    const int t0 = blockIdx.x * E_TF;
    const int T = lengths[b];
    const int tid = threadIdx.x;
    float* em_t = em + ((size_t)b * Tmax + t0) * ldc;
    const int nrows = min(E_TF, Tmax - t0);
    if (t0 >= T) {  // padding tile: zeros
        for (int i = tid; i < nrows * ldc; i += E_THREADS) em_t[i] = 0.0f;
        for (int i = tid; i < nrows; i += E_THREADS) rowterm[(size_t)b * Tmax + t0 + i] = 0.0f;
        return;
    }
    const int nval = min(E_TF, T - t0);  // frames of the tile that belong to the video
    const float* Xb = X + ((size_t)b * Tmax + t0) * D;
    const bool vec = (D % 4 == 0) && ((reinterpret_cast<uintptr_t>(X) & 15) == 0);

    const int tf = tid & 15;   // frames tf + 16 r
    const int tc = tid >> 4;   // classes 2 tc, 2 tc + 1 of the block
    float rowsq[4] = {0.f, 0.f, 0.f, 0.f};

    for (int cb = 0; cb < C; cb += E_CB) {
        float acc[4][2];
#pragma unroll
        for (int r = 0; r < 4; ++r) acc[r][0] = acc[r][1] = 0.0f;
        for (int d0 = 0; d0 < D; d0 += E_DK) {
            const int dk = min(E_DK, D - d0);
            __syncthreads();
            // stage X chunk (zero-filled beyond the video / beyond D)
            if (vec) {
                for (int i = tid; i < E_TF * (E_DK / 4); i += E_THREADS) {
                    const int r = i / (E_DK / 4), q = i - r * (E_DK / 4);
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (r < nval && q * 4 < dk) v = __ldg(reinterpret_cast<const float4*>(Xb + (size_t)r * D + d0) + q);
                    *reinterpret_cast<float4*>(Xs + r * E_LD + q * 4) = v;
                }
            } else {
                for (int i = tid; i < E_TF * E_DK; i += E_THREADS) {
                    const int r = i / E_DK, q = i - r * E_DK;
                    Xs[r * E_LD + q] = (r < nval && q < dk) ? __ldg(Xb + (size_t)r * D + d0 + q) : 0.0f;
                }
            }
            for (int i = tid; i < E_CB * E_DK; i += E_THREADS) {
                const int r = i / E_DK, q = i - r * E_DK;
                Ws[r * E_LD + q] = (cb + r < C && q < dk) ? __ldg(w + (size_t)(cb + r) * D + d0 + q) : 0.0f;
            }
            if (tid < E_DK) Vs[tid] = (tid < dk) ? __ldg(inv_var + d0 + tid) : 0.0f;
            __syncthreads();
#pragma unroll 4
            for (int q = 0; q < E_DK; q += 4) {
                float4 x[4];
#pragma unroll
                for (int r = 0; r < 4; ++r) x[r] = *reinterpret_cast<const float4*>(Xs + (tf + 16 * r) * E_LD + q);
                const float4 w0 = *reinterpret_cast<const float4*>(Ws + (2 * tc) * E_LD + q);
                const float4 w1 = *reinterpret_cast<const float4*>(Ws + (2 * tc + 1) * E_LD + q);
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    acc[r][0] = fmaf(x[r].x, w0.x, acc[r][0]);
                    acc[r][0] = fmaf(x[r].y, w0.y, acc[r][0]);
                    acc[r][0] = fmaf(x[r].z, w0.z, acc[r][0]);
                    acc[r][0] = fmaf(x[r].w, w0.w, acc[r][0]);
                    acc[r][1] = fmaf(x[r].x, w1.x, acc[r][1]);
                    acc[r][1] = fmaf(x[r].y, w1.y, acc[r][1]);
                    acc[r][1] = fmaf(x[r].z, w1.z, acc[r][1]);
                    acc[r][1] = fmaf(x[r].w, w1.w, acc[r][1]);
                }
                if (cb == 0 && tc == 0) {
                    const float4 iv = *reinterpret_cast<const float4*>(Vs + q);
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        rowsq[r] = fmaf(x[r].x * x[r].x, iv.x, rowsq[r]);
                        rowsq[r] = fmaf(x[r].y * x[r].y, iv.y, rowsq[r]);
                        rowsq[r] = fmaf(x[r].z * x[r].z, iv.z, rowsq[r]);
                        rowsq[r] = fmaf(x[r].w * x[r].w, iv.w, rowsq[r]);
                    }
                }
            }
        }
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int fr = tf + 16 * r;
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int c = cb + 2 * tc + u;
                if (c < C) Os[fr * cs + c] = acc[r][u] + __ldg(bias + c);
            }
        }
    }
    if (tc == 0) {
#pragma unroll
        for (int r = 0; r < 4; ++r) Rs[tf + 16 * r] = -0.5f * rowsq[r] + row_const;
    }
    __syncthreads();
    // ---- per-frame shift by the best class, 4 threads per frame ------------------------------
    {
        const int fr = tid >> 2, part = tid & 3;
        // shift = best penalised score; the penalty itself (-1e4 per offending frame, semimarkov.py:25,232)
        // is added AFTER the shift so that the large number meets an O(1) one exactly once, like the
        // reference's elp + constraints (semimarkov_modules.py:379-380)
        const float* pen_r = (penalty && fr < nval) ? penalty + ((size_t)b * Tmax + t0 + fr) * C : nullptr;
        float m = NEG;
        for (int c = part; c < C; c += 4) m = fmaxf(m, Os[fr * cs + c] + (pen_r ? __ldg(pen_r + c) : 0.0f));
        m = fmaxf(m, __shfl_xor_sync(FULL, m, 1));
        m = fmaxf(m, __shfl_xor_sync(FULL, m, 2));
        double contrib = 0.0;
        __syncthreads();
        if (part == 0) {
            float rt = 0.0f;
            if (fr < nval) {
                rt = Rs[fr] + m;
                contrib = (double)rt;
            } else {
                m = 0.0f;
            }
            Rs[fr] = m;  // reuse as the shift
            if (fr < nrows) rowterm[(size_t)b * Tmax + t0 + fr] = rt;
        }
        contrib = warp_sum(contrib);
        if ((tid & 31) == 0) red[tid >> 5] = contrib;
    }
    __syncthreads();
    if (tid == 0) {
        double s = 0.0;
        for (int i = 0; i < E_THREADS / 32; ++i) s += red[i];
        atomicAdd(offset + b, s);
    }
    for (int i = tid; i < nrows * ldc; i += E_THREADS) {
        const int fr = i / ldc, c = i - fr * ldc;
        float v = 0.0f;
        if (fr < nval && c < C) {
            v = Os[fr * cs + c] - Rs[fr];
            if (penalty) v += __ldg(penalty + ((size_t)b * Tmax + t0 + fr) * C + c);
        }
        em_t[i] = v;
    }
}

int launch_emission(const float* X, const float* w, const float* bias, const float* inv_var, const float* row_const,
                    const float* penalty, const int32_t* lengths, int B, int Tmax, int D, int C, int ldc,
                    float* em, float* rowterm, double* offset, cudaStream_t st) {
    cudaError_t e = cudaMemsetAsync(offset, 0, sizeof(double) * B, st);
    if (e != cudaSuccess) {
        set_error("memset offset: %s", cudaGetErrorString(e));
        return -3;
    }
    const int cs = C | 1;  // odd leading dimension: conflict-free column scans
    const size_t smem = (size_t)(E_TF * E_LD + E_CB * E_LD + E_DK + E_TF * cs + E_TF) * sizeof(float);
    if (smem > 200 * 1024) {
        set_error("emission: C=%d too large for the tile buffer", C);
        return -2;
    }
    if (smem > 48 * 1024) {
        e = cudaFuncSetAttribute(emission_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) {
            set_error("emission smem attr: %s", cudaGetErrorString(e));
            return -3;
        }
    }
    dim3 grid((Tmax + E_TF - 1) / E_TF, B);
    emission_kernel<<<grid, E_THREADS, smem, st>>>(X, w, bias, inv_var, row_const, penalty, lengths, Tmax, D, C, ldc, cs,
                                                    em, rowterm, offset);
    return check_launch("emission_kernel");
}

// ---------------------------------------------------------------------------------------------
// Class-weighted feature sums: out_wx[c][d] += sum_f wgt[f][c] x[f][d]   (a (C x F).(F x D) contraction)
//
// Persistent CTAs over tiles of W_TF consecutive frames of one video.  Thread = (4 feature dims, frame
// phase): it streams its float4 of x straight from global memory (a warp reads 512 contiguous bytes of
// a frame; W_UNROLL loads in flight per thread), takes the CP class weights of the frame as broadcast
// reads from shared memory and keeps a CP x 4 accumulator tile in registers across ALL its tiles; the
// four frame phases are reduced through shared memory at the end and flushed with one atomic per entry.
// ---------------------------------------------------------------------------------------------
constexpr int W_TF = 64;       // frames per tile
constexpr int W_DQ = 64;       // float4 columns per CTA (256 feature dims)
constexpr int W_THREADS = 256;
constexpr int W_UNROLL = 8;

// CP classes per CTA, split over CH thread groups (CPT = CP / CH accumulator rows per thread: 48 registers at
// CP = 24 instead of 96, so that two CTAs fit an SM); the remaining 4 / CH groups take different frames.
template <int CP, int CH>
__global__ void __launch_bounds__(W_THREADS, (CP / CH <= 16 ? 2 : 1))
weighted_sums_kernel(const float* __restrict__ X, const float* __restrict__ wgt, int ldc,
                     const int32_t* __restrict__ lengths, int B, int Tmax, int D, int C, int tiles_per_video,
                     float* __restrict__ out_wx, float* __restrict__ out_wsum) {
    constexpr int CPT = CP / CH;   // classes per thread
    constexpr int PH = 4 / CH;     // frame phases
    static_assert(CPT % 4 == 0 && PH >= 1, "class tile");
    __shared__ __align__(16) float Ws[W_TF][CP];
    __shared__ __align__(16) float Red[4][W_DQ * 4];
    const int tid = threadIdx.x;
    const int dq = tid & (W_DQ - 1), grp = tid / W_DQ;
    const int ph = grp % PH, ch = grp / PH;
    const int d0 = blockIdx.y * (W_DQ * 4) + dq * 4;  // first feature dim of this thread
    const int cb = blockIdx.z * CP;                   // first class of this CTA
    const int nc = min(CP, C - cb);
    const bool dok = d0 < D;                          // D % 4 == 0 on this path

    float acc[CPT][4];
#pragma unroll
    for (int c = 0; c < CPT; ++c) acc[c][0] = acc[c][1] = acc[c][2] = acc[c][3] = 0.0f;
    float wsum = 0.0f;  // thread c < nc (of the d-block 0 CTAs) accumulates the column sums

    // The class weights of the NEXT tile are fetched (float4 per thread, into registers) while the current tile streams
    // its features: a plain load -> store-to-shared phase per tile left every warp of the CTA waiting on HBM latency
    // in front of a barrier (40 % of the stall samples in ncu).
    constexpr int NQ = CP / 4;                                       // float4 per frame
    constexpr int NLD = (W_TF * NQ + W_THREADS - 1) / W_THREADS;     // float4 per thread and tile
    float4 wreg[NLD];
    // live tiles only, dealt round-robin to the persistent CTAs (TileCursor, hsmm_common.cuh)
    TileCursor cur;
    struct Tile { int b, t0, nf; };
    auto locate = [&](int g, Tile& tl) {
        int vb, j, vlen;
        if (!cur.locate(g, lengths, B, W_TF, vb, j, vlen)) return false;
        tl.b = vb;
        tl.t0 = j * W_TF;
        tl.nf = min(W_TF, vlen - tl.t0);
        return true;
    };
    auto issue = [&](const Tile& tl) {
#pragma unroll
        for (int l = 0; l < NLD; ++l) {
            const int i = tid + l * W_THREADS;
            const int f = i / NQ, c0 = (i - f * NQ) * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (i < W_TF * NQ && f < tl.nf && cb + c0 < ldc)
                v = __ldg(reinterpret_cast<const float4*>(wgt + ((size_t)tl.b * Tmax + tl.t0 + f) * ldc + cb + c0));
            if (c0 + 0 >= nc) v.x = 0.f;
            if (c0 + 1 >= nc) v.y = 0.f;
            if (c0 + 2 >= nc) v.z = 0.f;
            if (c0 + 3 >= nc) v.w = 0.f;
            wreg[l] = v;
        }
    };
    int g = blockIdx.x;
    Tile tl, nx;
    bool have = locate(g, tl);
    if (have) issue(tl);
    while (have) {
        const int b = tl.b, t0 = tl.t0, nf = tl.nf;
        __syncthreads();
#pragma unroll
        for (int l = 0; l < NLD; ++l) {
            const int i = tid + l * W_THREADS;
            if (i < W_TF * NQ) {
                const int f = i / NQ, c0 = (i - f * NQ) * 4;
                *reinterpret_cast<float4*>(&Ws[f][c0]) = wreg[l];
            }
        }
        __syncthreads();
        g += gridDim.x;
        const bool have_next = locate(g, nx);
        if (have_next) issue(nx);
        if (blockIdx.y == 0 && tid < nc) {
            for (int f = 0; f < nf; ++f) wsum += Ws[f][tid];
        }
        if (dok) {
            const float* xp = X + ((size_t)b * Tmax + t0) * D + d0;
#pragma unroll
            for (int f0 = 0; f0 < W_TF / PH; f0 += W_UNROLL) {
                float4 x[W_UNROLL];
#pragma unroll
                for (int u = 0; u < W_UNROLL; ++u) {
                    const int f = (f0 + u) * PH + ph;
                    x[u] = (f < nf) ? __ldg(reinterpret_cast<const float4*>(xp + (size_t)f * D)) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
#pragma unroll
                for (int u = 0; u < W_UNROLL; ++u) {
                    const int f = (f0 + u) * PH + ph;
#pragma unroll
                    for (int c4 = 0; c4 < CPT / 4; ++c4) {
                        // packed fp32 FMAs (FFMA2): one instruction updates the accumulators of two classes
                        const float4 w = *reinterpret_cast<const float4*>(&Ws[f][ch * CPT + c4 * 4]);
                        const float xs[4] = {x[u].x, x[u].y, x[u].z, x[u].w};
#pragma unroll
                        for (int d = 0; d < 4; ++d) {
                            ffma2(acc[c4 * 4 + 0][d], acc[c4 * 4 + 1][d], w.x, w.y, xs[d], xs[d]);
                            ffma2(acc[c4 * 4 + 2][d], acc[c4 * 4 + 3][d], w.z, w.w, xs[d], xs[d]);
                        }
                    }
                }
            }
        }
        tl = nx;
        have = have_next;
    }
    // reduce the frame phases through shared memory (one class row per group at a time), one atomic per entry
#pragma unroll
    for (int c = 0; c < CPT; ++c) {
        __syncthreads();
        *reinterpret_cast<float4*>(&Red[grp][dq * 4]) = make_float4(acc[c][0], acc[c][1], acc[c][2], acc[c][3]);
        __syncthreads();
        // thread (dd, h) sums the PH phase rows of class half h
        for (int i = tid; i < CH * W_DQ * 4; i += W_THREADS) {
            const int h = i / (W_DQ * 4), dd = i - h * (W_DQ * 4);
            const int cls = h * CPT + c;
            const int d = blockIdx.y * (W_DQ * 4) + dd;
            if (cls < nc && d < D) {
                float v = 0.0f;
#pragma unroll
                for (int q = 0; q < PH; ++q) v += Red[h * PH + q][dd];
                atomicAdd(out_wx + (size_t)(cb + cls) * D + d, v);
            }
        }
    }
    if (blockIdx.y == 0 && tid < nc) atomicAdd(out_wsum + cb + tid, wsum);
}

// generic fallback (D % 4 != 0 or unaligned X): thread <-> feature dim, scalar loads
__global__ void __launch_bounds__(256)
weighted_sums_generic_kernel(const float* __restrict__ X, const float* __restrict__ wgt, int ldc,
                             const int32_t* __restrict__ lengths, int B, int Tmax, int D, int C, int tiles_per_video,
                             float* __restrict__ out_wx, float* __restrict__ out_wsum) {
    constexpr int TF = 32, CB = 16;
    __shared__ float Ws[TF][CB + 1];
    const int tid = threadIdx.x;
    const int ntiles = B * tiles_per_video;
    for (int cb = 0; cb < C; cb += CB) {
        const int nc = min(CB, C - cb);
        for (int d0 = 0; d0 < D; d0 += 256) {
            const int d = d0 + tid;
            float acc[CB];
#pragma unroll
            for (int c = 0; c < CB; ++c) acc[c] = 0.0f;
            float wsum = 0.0f;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                const int b = tile / tiles_per_video;
                const int t0 = (tile - b * tiles_per_video) * TF;
                const int T = lengths[b];
                if (t0 >= T) continue;
                const int nf = min(TF, T - t0);
                __syncthreads();
                for (int i = tid; i < TF * CB; i += 256) {
                    const int f = i / CB, c = i - f * CB;
                    Ws[f][c] = (f < nf && c < nc) ? __ldg(wgt + ((size_t)b * Tmax + t0 + f) * ldc + cb + c) : 0.0f;
                }
                __syncthreads();
                if (d0 == 0 && tid < nc)
                    for (int f = 0; f < nf; ++f) wsum += Ws[f][tid];
                if (d < D) {
                    const float* xp = X + ((size_t)b * Tmax + t0) * D + d;
                    for (int f = 0; f < nf; ++f) {
                        const float x = __ldg(xp + (size_t)f * D);
#pragma unroll
                        for (int c = 0; c < CB; ++c) acc[c] = fmaf(Ws[f][c], x, acc[c]);
                    }
                }
            }
            if (d < D) {
#pragma unroll
                for (int c = 0; c < CB; ++c)
                    if (c < nc) atomicAdd(out_wx + (size_t)(cb + c) * D + d, acc[c]);
            }
            if (d0 == 0 && tid < nc) atomicAdd(out_wsum + cb + tid, wsum);
        }
    }
}

int launch_weighted_sums_tc(const float* X, const float* wgt, int ldc, const int32_t* lengths, int B, int Tmax, int D, int C,
                            float* out_wx, float* out_wsum, int num_sms, cudaStream_t st);  // hsmm_wsums_tc.cu

int launch_weighted_sums(const float* X, const float* wgt, int ldc, const int32_t* lengths, int B, int Tmax, int D, int C,
                         float* out_wx, float* out_wsum, int num_sms, cudaStream_t st) {
    {
        // tensor-core path for the shapes it takes (aligned, D <= 896 in feature blocks of 224, C <= 96 in class blocks of
        // 32: one launch per block pair); 1 = not eligible -> SIMT kernels below
        const int rc = launch_weighted_sums_tc(X, wgt, ldc, lengths, B, Tmax, D, C, out_wx, out_wsum, num_sms, st);
        if (rc != 1) return rc;
    }
    if (D % 4 != 0 || (reinterpret_cast<uintptr_t>(X) & 15) || ldc % 4 != 0 || (reinterpret_cast<uintptr_t>(wgt) & 15)) {
        const int tpv = (Tmax + 31) / 32;
        int grid = num_sms * 4;
        if (grid > B * tpv) grid = B * tpv;
        if (grid < 1) grid = 1;
        weighted_sums_generic_kernel<<<grid, 256, 0, st>>>(X, wgt, ldc, lengths, B, Tmax, D, C, tpv, out_wx, out_wsum);
        return check_launch("weighted_sums_generic_kernel");
    }
    const int tiles_per_video = (Tmax + W_TF - 1) / W_TF;
    const int dblocks = (D + W_DQ * 4 - 1) / (W_DQ * 4);
    const int cp = C <= 8 ? 8 : (C <= 16 ? 16 : (C <= 24 ? 24 : 32));
    const int cblocks = (C + cp - 1) / cp;
    int gx = (num_sms * 2) / (dblocks * cblocks);  // two resident CTAs per SM
    if (gx < 1) gx = 1;
    if (gx > B * tiles_per_video) gx = B * tiles_per_video;
    dim3 grid(gx, dblocks, cblocks);
    switch (cp) {
        case 8: weighted_sums_kernel<8, 1><<<grid, W_THREADS, 0, st>>>(X, wgt, ldc, lengths, B, Tmax, D, C, tiles_per_video, out_wx, out_wsum); break;
        case 16: weighted_sums_kernel<16, 1><<<grid, W_THREADS, 0, st>>>(X, wgt, ldc, lengths, B, Tmax, D, C, tiles_per_video, out_wx, out_wsum); break;
        case 24: weighted_sums_kernel<24, 2><<<grid, W_THREADS, 0, st>>>(X, wgt, ldc, lengths, B, Tmax, D, C, tiles_per_video, out_wx, out_wsum); break;
        default: weighted_sums_kernel<32, 2><<<grid, W_THREADS, 0, st>>>(X, wgt, ldc, lengths, B, Tmax, D, C, tiles_per_video, out_wx, out_wsum); break;
    }
    return check_launch("weighted_sums_kernel");
}

// ---------------------------------------------------------------------------------------------
// Feature moments: sum x, sum x^2 per dimension (double accumulation across tiles)
// ---------------------------------------------------------------------------------------------
constexpr int M_TF = 64;
__global__ void __launch_bounds__(256)
moments_kernel(const float* __restrict__ X, const int32_t* __restrict__ lengths, int B, int Tmax, int D,
               int tiles_per_video, double* __restrict__ sx, double* __restrict__ sx2) {
    const int ntiles = B * tiles_per_video;
    for (int d = threadIdx.x; d < D; d += blockDim.x) {
        double a = 0.0, a2 = 0.0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const int b = tile / tiles_per_video;
            const int t0 = (tile - b * tiles_per_video) * M_TF;
            const int T = lengths[b];
            if (t0 >= T) continue;
            const int nf = min(M_TF, T - t0);
            const float* xp = X + ((size_t)b * Tmax + t0) * D + d;
            float s = 0.0f, s2 = 0.0f, c1 = 0.0f, c2 = 0.0f;  // Kahan inside the tile
            for (int f = 0; f < nf; ++f) {
                const float x = __ldg(xp + (size_t)f * D);
                float y = x - c1, t = s + y;
                c1 = (t - s) - y;
                s = t;
                y = x * x - c2;
                t = s2 + y;
                c2 = (t - s2) - y;
                s2 = t;
            }
            a += (double)s;
            a2 += (double)s2;
        }
        atomicAdd(sx + d, a);
        atomicAdd(sx2 + d, a2);
    }
}

int launch_moments(const float* X, const int32_t* lengths, int B, int Tmax, int D, double* sx, double* sx2, int num_sms,
                   cudaStream_t st) {
    const int tiles_per_video = (Tmax + M_TF - 1) / M_TF;
    int grid = num_sms * 4;
    if (grid > B * tiles_per_video) grid = B * tiles_per_video;
    if (grid < 1) grid = 1;
    moments_kernel<<<grid, 256, 0, st>>>(X, lengths, B, Tmax, D, tiles_per_video, sx, sx2);
    return check_launch("moments_kernel");
}

// ---------------------------------------------------------------------------------------------
__global__ void onehot_kernel(const int32_t* __restrict__ labels, const int32_t* __restrict__ lengths, int B, int Tmax,
                              int C, int ldc, float* __restrict__ out) {
    const size_t total = (size_t)B * Tmax * ldc;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % ldc);
        const size_t bt = i / ldc;
        const int t = (int)(bt % Tmax), b = (int)(bt / Tmax);
        out[i] = (t < lengths[b] && c < C && labels[bt] == c) ? 1.0f : 0.0f;
    }
}

// ---------------------------------------------------------------------------------------------
// Ragged upload by the SMs: live rows of a padded batch in MAPPED pinned host memory -> device buffer
// ---------------------------------------------------------------------------------------------
// One copy-engine transfer per video costs ~3.8 us of set-up (a 1 MB cudaMemcpyAsync reaches 45 GB/s where a 1 GB one
// reaches 55 GB/s on this box); a few CTAs reading the host rows over PCIe with 16-byte loads pay it once.  Work item =
// (video, 16 KB piece) over the PADDED index space, pieces behind the end of a video are skipped.
constexpr int UP_PIECE = 1024;  // uint4 per piece (16 KB)
constexpr int UP_THREADS = 512;
__global__ void __launch_bounds__(UP_THREADS) upload_ragged_kernel(const uint4* __restrict__ host, uint4* __restrict__ dev,
                                                                  const int32_t* __restrict__ lengths, int B, int Tmax,
                                                                  long long row_q /* uint4 per row */) {
    const long long vid_q = (long long)Tmax * row_q;
    const long long ppv = (vid_q + UP_PIECE - 1) / UP_PIECE;
    const long long items = (long long)B * ppv;
    for (long long it = blockIdx.x; it < items; it += gridDim.x) {
        const int b = (int)(it / ppv);
        const long long q0 = (it - (long long)b * ppv) * UP_PIECE;
        const long long live = (long long)min(max(lengths[b], 0), Tmax) * row_q;
        if (q0 >= live) continue;
        const long long n = min((long long)UP_PIECE, live - q0);
        const uint4* src = host + (long long)b * vid_q + q0;
        uint4* dst = dev + (long long)b * vid_q + q0;
        // UP_PIECE / UP_THREADS = 2 independent loads per thread in flight
        uint4 v[UP_PIECE / UP_THREADS];
#pragma unroll
        for (int k = 0; k < UP_PIECE / UP_THREADS; ++k) {
            const int i = threadIdx.x + k * UP_THREADS;
            if (i < n) v[k] = __ldcs(src + i);
        }
#pragma unroll
        for (int k = 0; k < UP_PIECE / UP_THREADS; ++k) {
            const int i = threadIdx.x + k * UP_THREADS;
            if (i < n) dst[i] = v[k];
        }
    }
}

int launch_upload_ragged(const float* host_mapped, float* dev, const int32_t* lengths, int B, int Tmax, int width, int ctas,
                         cudaStream_t st) {
    upload_ragged_kernel<<<ctas, UP_THREADS, 0, st>>>(reinterpret_cast<const uint4*>(host_mapped), reinterpret_cast<uint4*>(dev),
                                                     lengths, B, Tmax, (long long)width / 4);
    return check_launch("upload_ragged_kernel");
}

int launch_onehot(const int32_t* labels, const int32_t* lengths, int B, int Tmax, int C, int ldc, float* out, int num_sms,
                  cudaStream_t st) {
    onehot_kernel<<<num_sms * 8, 256, 0, st>>>(labels, lengths, B, Tmax, C, ldc, out);
    return check_launch("onehot_kernel");
}

// ---------------------------------------------------------------------------------------------
// Gold-segmentation score: one warp per video, 32 frames per step.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
gold_score_kernel(const float* __restrict__ em, int ldc, const float* __restrict__ init, const float* __restrict__ trans,
                  const float* __restrict__ lenp, const float* __restrict__ end, const double* __restrict__ offset,
                  const int32_t* __restrict__ lengths, const int32_t* __restrict__ spans, const float* __restrict__ grad,
                  int B, int Tmax, int C, int L, double* __restrict__ out, float* d_init, float* d_trans, float* d_len,
                  float* d_em) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= B) return;
    const int b = warp;
    const int T = lengths[b];
    const int32_t* sp = spans + (size_t)b * Tmax;
    const float* em_b = em + (size_t)b * Tmax * ldc;
    const float g = grad ? grad[b] : 0.0f;
    float* dem = grad ? d_em + (size_t)b * Tmax * ldc : nullptr;
    if (dem)
        for (int i = lane; i < Tmax * ldc; i += 32) dem[i] = 0.0f;
    __syncwarp();
    double acc = 0.0;
    // A gold segmentation the model cannot score -- a start label outside [0, C) (the wrapper writes -2 for global ids
    // that are not valid classes of the batch), frame 0 not a segment start, a segment longer than K-1 -- makes the
    // reference raise inside struct.to_parts / score (semimarkov_modules.py:626-655); here the video's score is NaN.
    bool bad = false;
    int cur_c = -1, cur_s = 0;  // running segment (class, start), uniform across the warp
    for (int t0 = 0; t0 < T; t0 += 32) {
        const int t = t0 + lane;
        const int s = (t < T) ? sp[t] : -1;
        bad |= (s < -1) || (s >= C) || (t == 0 && T > 0 && s < 0);
        const unsigned starts = __ballot_sync(FULL, s >= 0);
        // class and start of the segment covering frame t
        const unsigned upto = starts & (0xffffffffu >> (31 - lane));
        int my_c = cur_c, my_s = cur_s;
        int src = upto ? 31 - __clz(upto) : 0;
        const int sc = __shfl_sync(FULL, s, src);
        if (upto) {
            my_c = sc;
            my_s = t0 + src;
        }
        // previous segment as seen from a start at lane: the one covering frame t-1
        const unsigned before = starts & ((1u << lane) - 1u);
        int pc = cur_c, ps = cur_s;
        int psrc = before ? 31 - __clz(before) : 0;
        const int psc = __shfl_sync(FULL, s, psrc);
        if (before) {
            pc = psc;
            ps = t0 + psrc;
        }
        if (t < T && my_c >= 0 && my_c < C) {
            acc += (double)em_b[(size_t)t * ldc + my_c];
            if (dem) dem[(size_t)t * ldc + my_c] = g;
            if (s >= 0) {
                if (t == 0) {
                    acc += (double)init[my_c];
                    if (grad) atomicAdd(d_init + my_c, g);
                } else if (pc >= 0 && pc < C) {
                    const int l = t - ps;
                    acc += (double)trans[(size_t)my_c * C + pc];
                    bad |= (l > L);
                    if (l >= 1 && l <= L) acc += (double)lenp[(size_t)l * C + pc];
                    if (grad) {
                        atomicAdd(d_trans + (size_t)my_c * C + pc, g);
                        if (l >= 1 && l <= L) atomicAdd(d_len + (size_t)l * C + pc, g);
                    }
                }
            }
        }
        // carry the last segment of this chunk
        if (starts) {
            const int last = 31 - __clz(starts);
            cur_c = __shfl_sync(FULL, s, last);
            cur_s = t0 + last;
        }
    }
    if (lane == 0 && cur_c >= 0 && cur_c < C) {
        const int l = T - cur_s;
        bad |= (l > L);
        if (l >= 1 && l <= L) {
            acc += (double)lenp[(size_t)l * C + cur_c];
            if (grad) atomicAdd(d_len + (size_t)l * C + cur_c, g);
        }
        if (end) acc += (double)end[(size_t)b * C + cur_c];
    }
    acc = warp_sum(acc);
    bad = __any_sync(FULL, bad);
    if (lane == 0) out[b] = bad ? __longlong_as_double(0x7ff8000000000000LL) : acc + (offset ? offset[b] : 0.0);
}

int launch_gold(const float* em, int ldc, const float* init, const float* trans, const float* lenp, const float* end,
                const double* offset, const int32_t* lengths, const int32_t* spans, const float* grad, int B, int Tmax,
                int C, int L, double* out, float* d_init, float* d_trans, float* d_len, float* d_em, cudaStream_t st) {
    const int blocks = (B * 32 + 127) / 128;
    gold_score_kernel<<<blocks, 128, 0, st>>>(em, ldc, init, trans, lenp, end, offset, lengths, spans, grad, B, Tmax, C, L,
                                              out, d_init, d_trans, d_len, d_em);
    return check_launch("gold_score_kernel");
}

}  // namespace hsmm
