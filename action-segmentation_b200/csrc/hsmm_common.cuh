// Shared device/host helpers for the HSMM kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

namespace hsmm {

constexpr float NEG = -1.0e30f;            // absorbing "minus infinity" that never produces NaN
constexpr float LOG2E = 1.4426950408889634f;
constexpr double LN2 = 0.6931471805599453;
constexpr unsigned FULL = 0xffffffffu;

// ---- error plumbing (host) ----------------------------------------------------------------
void set_error(const char* fmt, ...);
extern std::atomic<uint64_t> g_launches;
int check_launch(const char* what);

// ---- device helpers -----------------------------------------------------------------------
#ifdef HSMM_ACCURATE_MUFU  // experiment switch: libm-accurate exp2/log2 instead of the MUFU approximations
__device__ __forceinline__ float ex2(float x) { return exp2f(x); }
__device__ __forceinline__ float lg2(float x) { return log2f(x); }
#else
__device__ __forceinline__ float ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float lg2(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
#endif

// Barrier over the `nwarps` warps that cooperate on one video.  One warp: __syncwarp.
// Several warps: a named barrier (ids 1..15; id 0 is __syncthreads).
__device__ __forceinline__ void group_sync(int nwarps, int bar_id) {
    if (nwarps == 1) {
        __syncwarp();
    } else {
        asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(nwarps * 32) : "memory");
    }
}

// Combine a value across the S k-slices of a class inside one warp (lanes cl + j*CPW).
template <int S>
__device__ __forceinline__ float slice_max(float v) {
#pragma unroll
    for (int off = 32 / S; off < 32; off <<= 1) v = fmaxf(v, __shfl_xor_sync(FULL, v, off));
    return v;
}
template <int S>
__device__ __forceinline__ float slice_sum(float v) {
#pragma unroll
    for (int off = 32 / S; off < 32; off <<= 1) v += __shfl_xor_sync(FULL, v, off);
    return v;
}
// arg-max with ties going to the smaller index
template <int S>
__device__ __forceinline__ void slice_argmax(float& v, int& idx) {
#pragma unroll
    for (int off = 32 / S; off < 32; off <<= 1) {
        float ov = __shfl_xor_sync(FULL, v, off);
        int oi = __shfl_xor_sync(FULL, idx, off);
        if (ov > v || (ov == v && oi < idx)) {
            v = ov;
            idx = oi;
        }
    }
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v = fmaxf(v, __shfl_xor_sync(FULL, v, off));
    return v;
}
// packed fp32 FMA (sm_100a FFMA2): (d0, d1) += (a0, a1) * (b0, b1), each lane rounded like fmaf
__device__ __forceinline__ void ffma2(float& d0, float& d1, float a0, float a1, float b0, float b1) {
    unsigned long long a, b, c;
    asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(a0), "f"(a1));
    asm("mov.b64 %0, {%1, %2};" : "=l"(b) : "f"(b0), "f"(b1));
    asm("mov.b64 %0, {%1, %2};" : "=l"(c) : "f"(d0), "f"(d1));
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(c) : "l"(a), "l"(b));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(d0), "=f"(d1) : "l"(c));
}
// whole-warp maximum in one instruction (sm_100a: redux.sync.max.f32 -> CREDUX.MAX.F32)
__device__ __forceinline__ float warp_max_redux(float v) {
    float r;
    asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(v));
    return r;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(FULL, v, off);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(FULL, v, off);
    return v;
}


// ---- balanced enumeration of the frame tiles of a ragged batch ------------------------------------
// Video b contributes ceil(len_b / TF) tiles of TF frames; tile g of the batch-wide enumeration is tile j of video
// vb.  A persistent CTA takes g = blockIdx.x, blockIdx.x + gridDim.x, ...: every CTA gets the same number of LIVE
// tiles (+-1), whereas striding over the padded (B x ceil(Tmax/TF)) grid and skipping the dead tiles leaves the CTAs
// with binomially distributed work (measured: SMs idle for 25 % of the kernel on U[1000,3000]-frame videos).
// The cursor is warp-cooperative and monotone: all 32 lanes call locate() with the same, non-decreasing g; it walks
// the lengths 32 videos at a time (one coalesced load + a shuffle scan per window).
struct TileCursor {
    int b0 = 0;     // first video of the current window of 32
    int base = 0;   // tiles before video b0
    int incl = 0;   // lane i: inclusive tile count of videos b0 .. b0+i
    int len = 0;    // lane i: length of video b0+i (0 past the batch)
    bool loaded = false;

    __device__ __forceinline__ void load(const int32_t* __restrict__ lengths, int B, int TF) {
        const int lane = threadIdx.x & 31;
        const int b = b0 + lane;
        len = (b < B) ? max(lengths[b], 0) : 0;
        int v = (len + TF - 1) / TF;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const int o = __shfl_up_sync(FULL, v, off);
            if (lane >= off) v += o;
        }
        incl = v;
        loaded = true;
    }
    // false when g is past the last tile of the batch
    __device__ __forceinline__ bool locate(int g, const int32_t* __restrict__ lengths, int B, int TF, int& vb, int& j, int& vlen) {
        if (!loaded) load(lengths, B, TF);
        for (;;) {
            if (b0 >= B) return false;
            const int tot = __shfl_sync(FULL, incl, 31);
            if (g < base + tot) break;
            base += tot;
            b0 += 32;
            load(lengths, B, TF);
        }
        const unsigned m = __ballot_sync(FULL, g < base + incl);
        const int l = __ffs(m) - 1;
        const int incl_l = __shfl_sync(FULL, incl, l);
        vlen = __shfl_sync(FULL, len, l);
        vb = b0 + l;
        j = g - base - (incl_l - (vlen + TF - 1) / TF);
        return true;
    }
};

// ---- DP parameter block -------------------------------------------------------------------
struct DpParams {
    // inputs
    const float* em;       // (B, Tmax, ldc)
    const float* init;     // (C)
    const float* trans;    // (C, C) [to, from]
    const float* lenp;     // (K, C), rows 0..K-1
    const float* end;      // (B, C) or null
    const double* offset;  // (B) or null
    const int32_t* lengths;
    const int32_t* order;  // (B) or null
    const int32_t* trans_pred;  // (C, 4) unmasked predecessors of each class (-1 padded) or null = dense
    const int32_t* trans_succ;  // (C, 4) unmasked successors of each class (-1 padded) or null = dense
    int B, Tmax, C, L, ldc;
    int W;    // warps per video
    int VPB;  // videos per block
    // viterbi
    uint32_t* bp;             // (B, Tmax+1, ldc): (k << 16) | c1
    const int32_t* class_ids;  // (C+1) or null
    int64_t* spans;           // (B, Tmax+1)
    int64_t* labels;          // (B, Tmax) or null
    double* score;            // (B) or null
    // deferred-arg-max Viterbi (hsmm_dp_vit2.cuh); vbeta aliases bp
    float* vbeta;     // (B, Tmax+1, ldc) beta^[n][c]
    uint32_t* vpred;  // (B, Tmax+1, ldc) arg-max predecessor class of a segment start
    float* vdelta;    // (B, Tmax+1) normaliser increments
    float* vflag;     // (B) 1: no path through the sparse transition hint -> dp_forward_kernel<VIT> decodes the video
    // forward
    int xp;         // 1: per-class state (and fbeta/fgamma) in double -- see state_t in hsmm_dp_reg.cuh
    void* fbeta;    // (B, Tmax+1, ldc)  beta[n], n = 0..T-1   (log2 domain; float, double when xp)
    void* fgamma;   // (B, Tmax+1, ldc)  gamma[n], n = 1..T    (log2 domain; float, double when xp)
    float* fdelta;  // (B, Tmax+1)  per-frame normaliser increments delta_n, n = 1..T
    double* logz2;  // (B) log2 Z relative to the accumulated normaliser nu_T
    float* fflag;   // (B) bit 0: recomputed against the dense matrix (sparse hint degenerate); 2: the linear-window
                    //     forward kernel could not certify the video -> the log-domain kernel recomputes it and
                    //     sets bit 2 (value 4 or 5)
    float* bflag;   // (B) 1: the linear-window backward kernel left the video to the log-domain kernel
    int only_flagged;  // log-domain kernels: process only the videos flagged by the linear-window kernels
    double* logz;   // (B)
    // backward
    const float* grad;  // (B)
    float* d_init;
    float* d_trans;  // (C, C)
    float* d_len;    // (K, C)
    float* d_em;     // (B, Tmax, ldc)
};

// ---- grouped launches: one kernel over the batches of several tasks (class sets) ------------------
// The parameter blocks travel as ONE kernel argument (by value, a few KB: works under CUDA-graph capture and needs no
// device allocation); CTA `blockIdx.x` belongs to task t with first[t] <= blockIdx.x < first[t+1].
constexpr int GROUP_MAX = 32;
struct DpGroup {
    int n;
    int first[GROUP_MAX + 1];
    DpParams t[GROUP_MAX];
};
__device__ __forceinline__ int group_find(const DpGroup& g, int bid, int& local) {
    int t = 0;
    while (t + 1 < g.n && bid >= g.first[t + 1]) ++t;
    local = bid - g.first[t];
    return t;
}

}  // namespace hsmm
