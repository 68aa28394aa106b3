// Class-weighted feature sums on the 5th-generation tensor cores (sm_100a):
//
//   out_wx[c][d] += sum_f wgt[f][c] * x[f][d],   out_wsum[c] += sum_f wgt[f][c]      (frames f of all videos)
//
// the reduction behind d/d gaussian_means (loss.backward() through the emission scores, semimarkov.py:284-286) and the
// supervised class means (semimarkov_utils.py:74-126).  It is a (D x F).(F x C) contraction whose reduction dimension
// is the FRAME index, so both operands are "MN-major" for the tensor core: a frame is one 128-byte row of 32
// consecutive feature dims (A = X^T, M = feature dim) or of the <= 32 class weights (B = W^T, N = class).  That is
// the shared-memory image of the emission kernel's TMA boxes, in the one swizzle MN-major tf32 operands may use
// (128-byte rows, 32-byte pieces permuted by row & 3: CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B <-> UMMA SWIZZLE_128B_BASE32B).
// The SIMT kernel (hsmm_aux.cu) spends 194 warp instructions per frame on this; here the FMAs run on the tensor
// pipe and the SM only splits the operands (3xTF32: x = xb + xs, w = wb + ws; xb.wb + xs.wb + xb.ws in fp32, the
// dropped term is ~2^-22 relative).
//
// Persistent CTAs (one per SM); a tile is 32 consecutive frames of ONE video, live tiles only, dealt round-robin
// (TileCursor).  One ring stage = one tile:
//   warp 0      TMA producer: <= 7 boxes of 32 frames x 32 dims of X and one box of 32 frames x 32 weights
//   warp 1      MMA issuer (one lane): per 8 frames and per half of the feature dims (M = 128) three
//               tcgen05.mma.kind::tf32 products in two instructions (xb.[wb|ws] with N = 64, xs.wb with N = 32); the
//               accumulators (2 feature halves x 128 lanes x 64 columns) stay in TMEM for the whole kernel
//   warps 2..9  converters: remainders xs = x - (x & 0xffffe000) into the stage's second half (the landed chunk itself is
//               the big operand), weights split in place, rows behind the end of the video and class columns >= C
//               forced to zero; at the end warps 2..5 read the accumulators and flush them with atomics.
// Feature column 255 of every stage holds the constant 1 (chunk slot 7 is never loaded), so that accumulator row
// 255 is the column sum of the weights (out_wsum).
#include "hsmm_tc.cuh"

namespace hsmm {

namespace wtc {

using namespace tc;

constexpr int KC = 32;                       // floats per 128-byte row
constexpr int NCH = 8;                       // chunk slots per stage (256 feature columns)
constexpr int NSTAGE = 3;
constexpr int CONV_THREADS = 256;
constexpr int THREADS = 64 + CONV_THREADS;
constexpr int TMEM_COLS = 128;                // 2 feature halves x (NPAD columns x.wb + NPAD columns xb.ws)
constexpr int NPAD = 32;                     // classes per accumulator (UMMA N)
// TF = frames per tile (32: 72 KB stages, one CTA per SM)
template <int TF>
struct Geo {
    static constexpr int CHUNK_BYTES = TF * 128;        // one chunk of one tile (4 KB at TF = 32)
    static constexpr int XPART = NCH * CHUNK_BYTES;     // 32 KB
    static constexpr int WPART = TF * 128;              // 4 KB
    static constexpr int STAGE_BYTES = 2 * XPART + 2 * WPART;  // big + small of X and of the weights: 72 KB
    static constexpr size_t SMEM_BYTES = (size_t)NSTAGE * STAGE_BYTES + 128;
    static constexpr int XUNITS = (NCH - 1) * TF * 8;   // 16-byte units of the loadable chunk slots
    static constexpr int XIT = (XUNITS + CONV_THREADS - 1) / CONV_THREADS;
};

struct Params {
    const int32_t* lengths;
    float* out_wx;    // already offset to (first class, first feature) of this pass
    float* out_wsum;  // offset to the first class of this pass; null for the passes over later feature blocks
    int B, Tmax, D, C;  // D, C: feature dims / classes of THIS pass (<= 224, <= 32)
    int ldo;            // row stride of out_wx = feature dims of the whole problem
    int nchunk;  // ceil(D / 32) <= 7
};

// (THREADS, 2): <= 96 registers, 30 K per CTA: two DP CTAs stay resident beside this HBM-bound kernel
template <int TF>
__global__ void __launch_bounds__(THREADS, 2)
weighted_sums_tc_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w, const Params p) {
    constexpr int CHUNK_BYTES = Geo<TF>::CHUNK_BYTES, XPART = Geo<TF>::XPART, WPART = Geo<TF>::WPART;
    constexpr int STAGE_BYTES = Geo<TF>::STAGE_BYTES, XIT = Geo<TF>::XIT;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    if (smem_u32(smem_raw) & 1023u) __trap();
    uint8_t* st_s = smem_raw;
    uint64_t* bars = reinterpret_cast<uint64_t*>(st_s + (size_t)NSTAGE * STAGE_BYTES);
    uint64_t* full = bars;             // TMA -> converters
    uint64_t* conv = full + NSTAGE;    // converters -> MMA
    uint64_t* empty = conv + NSTAGE;   // MMA -> TMA
    uint64_t* done = empty + NSTAGE;   // last MMA -> epilogue
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < NSTAGE; ++s) {
            mbar_init(full + s, 1);
            mbar_init(conv + s, CONV_THREADS);
            mbar_init(empty + s, 1);
        }
        mbar_init(done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // the ring starts as zeros (chunk slots that are never loaded stay zero) except the ones column: feature 255 =
    // element 3 of the logical 16-byte unit 7 of every frame row of chunk slot 7 (32-byte piece 3 ^ (row & 3), upper half)
    for (int i = threadIdx.x; i < NSTAGE * STAGE_BYTES / 16; i += THREADS)
        reinterpret_cast<float4*>(st_s)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();
    if (threadIdx.x < NSTAGE * TF) {
        const int s = threadIdx.x / TF, f = threadIdx.x % TF;
        float* rowp = reinterpret_cast<float*>(st_s + (size_t)s * STAGE_BYTES + 7 * CHUNK_BYTES + f * 128 + ((((3 ^ (f & 3)) << 1) | 1) * 16));
        rowp[3] = 1.0f;
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    TileCursor cur;
    int vb, vj, vlen;
    if (warp == 0) {
        // ===================== TMA producer =====================
        int st = 0;
        uint32_t ph = 0;
        for (int g = blockIdx.x; cur.locate(g, p.lengths, p.B, TF, vb, vj, vlen); g += gridDim.x) {
            if (lane == 0) {
                const int row0 = vb * p.Tmax + vj * TF;
                uint8_t* sb = st_s + (size_t)st * STAGE_BYTES;
                mbar_wait(empty + st, ph ^ 1);
                mbar_arrive_expect_tx(full + st, (uint32_t)(p.nchunk * CHUNK_BYTES + WPART));
                for (int ch = 0; ch < p.nchunk; ++ch) tma_load_2d(sb + ch * CHUNK_BYTES, &tmap_x, full + st, ch * KC, row0);
                tma_load_2d(sb + 2 * XPART, &tmap_w, full + st, 0, row0);
                if (++st == NSTAGE) {
                    st = 0;
                    ph ^= 1;
                }
            }
            st = __shfl_sync(FULL, st, 0);
            ph = __shfl_sync(FULL, ph, 0);
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (whole warp converged, one elected lane issues) =====================
        // D (f32), A = B = tf32, both MN-major, N = 32, M = 128
        const uint32_t idesc_n = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(NPAD >> 3) << 17) |
                                 ((uint32_t)(128 >> 4) << 24);
        // the big and the small weights lie WPART bytes apart: as one B operand with N = 2 NPAD (LBO = WPART) xb meets
        // both in a single MMA -- two MMAs per (8 frames, feature half) instead of three
        const uint32_t idesc_2n = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(2 * NPAD >> 3) << 17) |
                                  ((uint32_t)(128 >> 4) << 24);
        const uint32_t leader = elect_one();
        const int total_tiles = count_tiles(p.lengths, p.B, TF);
        const int my_tiles = (total_tiles > (int)blockIdx.x) ? (total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
        const uint64_t xdesc0 = smem_desc_sw128b32_mn(0, CHUNK_BYTES, 512);   // address field added per MMA
        const uint64_t wdesc0 = smem_desc_sw128b32_mn(0, WPART, 512);
        const uint32_t st0 = smem_u32(st_s);
        int st = 0;
        uint32_t ph = 0;
        uint32_t accum = 0;
        for (int it = 0; it < my_tiles; ++it) {
            mbar_wait(conv + st, ph);
            tc_fence_after();
            const uint32_t xb0 = st0 + st * STAGE_BYTES;
            const uint32_t wb0 = xb0 + 2 * XPART;
#pragma unroll
            for (int ks = 0; ks < TF / 8; ++ks) {
                const uint64_t wb = desc_at(wdesc0, wb0 + ks * 1024);
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    const uint64_t xb = desc_at(xdesc0, xb0 + half * 4 * CHUNK_BYTES + ks * 1024);
                    const uint64_t xs = desc_at(xb, XPART);
                    const uint32_t d_tmem = tmem_base + half * 2 * NPAD;
                    tc_mma_tf32_lead(d_tmem, xb, wb, idesc_2n, accum, leader);   // xb.[wb | ws]
                    tc_mma_tf32_lead(d_tmem, xs, wb, idesc_n, 1, leader);        // xs.wb
                }
                accum = 1;
            }
            tc_commit_lead(empty + st, leader);
            if (++st == NSTAGE) {
                st = 0;
                ph ^= 1;
            }
        }
        tc_commit_lead(done, leader);
    } else {
        // ===================== converters =====================
        const int tc_id = threadIdx.x - 64;   // 0..255
        int st = 0;
        uint32_t ph = 0;
        bool any = false;
        for (int g = blockIdx.x; cur.locate(g, p.lengths, p.B, TF, vb, vj, vlen); g += gridDim.x) {
            any = true;
            const int nf = min(TF, vlen - vj * TF);   // live frames of the tile
            uint8_t* sb = st_s + (size_t)st * STAGE_BYTES;
            mbar_wait(full + st, ph);
            // X: 16-byte unit u of the stage <-> (chunk u / (8 TF), frame row (u >> 3) % TF); elementwise, in place
            const int xunits = p.nchunk * TF * 8;
            float4 x[XIT];
#pragma unroll
            for (int i = 0; i < XIT; ++i) {
                const int u = tc_id + i * CONV_THREADS;
                x[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (u < xunits) x[i] = *reinterpret_cast<const float4*>(sb + (size_t)u * 16);
            }
            // weights: unit <-> (frame row tc_id >> 3, physical 16-byte slot tc_id & 7)
            const int wf = tc_id >> 3;
            const int wpj = tc_id & 7;                       // physical 16-byte slot; its 32-byte piece is permuted by (row & 3)
            const int wc0 = (((((wpj >> 1) ^ (wf & 3)) << 1) | (wpj & 1))) * 4;   // first class of the unit
            const bool wmine = tc_id < TF * 8;
            float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
            if (wmine) w = *reinterpret_cast<const float4*>(sb + 2 * XPART + (size_t)tc_id * 16);
            const int xf = (tc_id >> 3) & (TF - 1);         // frame row of this thread's X units (the same for every i)
            const bool xlive = xf < nf;
#pragma unroll
            for (int i = 0; i < XIT; ++i) {
                const int u = tc_id + i * CONV_THREADS;
                if (u < xunits) {
                    float4 v = x[i];
                    if (!xlive) v = make_float4(0.f, 0.f, 0.f, 0.f);
                    float4 big, sml;
                    big.x = __uint_as_float(__float_as_uint(v.x) & TF32_MASK);
                    big.y = __uint_as_float(__float_as_uint(v.y) & TF32_MASK);
                    big.z = __uint_as_float(__float_as_uint(v.z) & TF32_MASK);
                    big.w = __uint_as_float(__float_as_uint(v.w) & TF32_MASK);
                    sml.x = v.x - big.x;
                    sml.y = v.y - big.y;
                    sml.z = v.z - big.z;
                    sml.w = v.w - big.w;
                    // the landed chunk itself is the big operand (kind::tf32 ignores the 13 low mantissa bits of a word);
                    // only rows behind the end of the video are overwritten (with zeros: they may hold anything)
                    if (!xlive) *reinterpret_cast<float4*>(sb + (size_t)u * 16) = big;
                    *reinterpret_cast<float4*>(sb + XPART + (size_t)u * 16) = sml;
                }
            }
            if (wmine) {
                if (wf >= nf) w = make_float4(0.f, 0.f, 0.f, 0.f);
                if (wc0 + 0 >= p.C) w.x = 0.f;
                if (wc0 + 1 >= p.C) w.y = 0.f;
                if (wc0 + 2 >= p.C) w.z = 0.f;
                if (wc0 + 3 >= p.C) w.w = 0.f;
                float4 big, sml;
                big.x = __uint_as_float(__float_as_uint(w.x) & TF32_MASK);
                big.y = __uint_as_float(__float_as_uint(w.y) & TF32_MASK);
                big.z = __uint_as_float(__float_as_uint(w.z) & TF32_MASK);
                big.w = __uint_as_float(__float_as_uint(w.w) & TF32_MASK);
                sml.x = w.x - big.x;
                sml.y = w.y - big.y;
                sml.z = w.z - big.z;
                sml.w = w.w - big.w;
                *reinterpret_cast<float4*>(sb + 2 * XPART + (size_t)tc_id * 16) = big;
                *reinterpret_cast<float4*>(sb + 2 * XPART + WPART + (size_t)tc_id * 16) = sml;
            }
            fence_proxy_async();  // generic-proxy writes -> visible to the tensor core (async proxy)
            mbar_arrive(conv + st);
            if (++st == NSTAGE) {
                st = 0;
                ph ^= 1;
            }
        }
        // ===================== final flush (warps 2..5: TMEM lane quarter = warp % 4) =====================
        if (warp < 6 && any) {
            mbar_wait(done, 0);
            tc_fence_after();
            const int q = warp & 3;
            const int r = q * 32 + lane;
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                float v[NPAD], lo[NPAD];
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + half * 2 * NPAD;
                tc_ld16(taddr, v);
                tc_ld16(taddr + 16, v + 16);
                tc_ld16(taddr + NPAD, lo);
                tc_ld16(taddr + NPAD + 16, lo + 16);
                tc_wait_ld();
#pragma unroll
                for (int c = 0; c < NPAD; ++c) v[c] += lo[c];
                const int d = half * 128 + r;
#pragma unroll
                for (int c = 0; c < NPAD; ++c) {
                    if (c < p.C) {
                        if (d < p.D) atomicAdd(p.out_wx + (size_t)c * p.ldo + d, v[c]);
                        if (d == 255 && p.out_wsum) atomicAdd(p.out_wsum + c, v[c]);
                    }
                }
            }
            tc_fence_before();
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

}  // namespace wtc

// returns 1 when the shape / alignment is not eligible (caller falls back to the SIMT kernel), 0 on launch, < 0 on error.
// One pass covers <= 224 feature dims x <= 32 classes (the TMEM accumulators); larger problems are tiled into passes over
// feature blocks (X is read once in total) and class blocks (X is read once per class block): the README's D = 300
// feature set (run_crosstask_i3d-resnet-audio.sh:13) takes 2 passes, Breakfast's C = 48 takes 2.  More than 3 class
// blocks would re-read X more often than the SIMT kernel is slower: those shapes return 1.
template <int TF>
static int launch_weighted_sums_tc_t(const float* X, const float* wgt, int ldc, const int32_t* lengths, int B, int Tmax, int D, int C,
                                     float* out_wx, float* out_wsum, int ctas, cudaStream_t st) {
    using namespace wtc;
    constexpr int DBLK = 7 * KC, CBLK = NPAD;
    const int ncb = (C + CBLK - 1) / CBLK, ndb = (D + DBLK - 1) / DBLK;
    const long long rows = (long long)B * Tmax;
    if (rows < 1 || rows >= (1ll << 31) - TF) return 1;
    const long long max_tiles = (long long)B * ((Tmax + TF - 1) / TF);
    int grid = ctas < max_tiles ? ctas : (int)max_tiles;
    if (grid < 1) grid = 1;
    cudaError_t e = cudaFuncSetAttribute(weighted_sums_tc_kernel<TF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Geo<TF>::SMEM_BYTES);
    if (e != cudaSuccess) {
        set_error("weighted_sums_tc smem attr: %s", cudaGetErrorString(e));
        return -3;
    }
    // every tensor map must encode before the first launch: a failure has to fall back to the SIMT kernel cleanly
    CUtensorMap mx[4], mw[3];
    for (int db = 0; db < ndb; ++db) {
        const int d0 = db * DBLK, dn = (D - d0 < DBLK) ? D - d0 : DBLK;
        if (!tc::make_map(&mx[db], X + d0, (uint64_t)rows, (uint64_t)dn, (uint64_t)D, KC, TF, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) return 1;
    }
    for (int cb = 0; cb < ncb; ++cb) {
        const int c0 = cb * CBLK, cn = (ldc - c0 < CBLK) ? ldc - c0 : CBLK;
        if (!tc::make_map(&mw[cb], wgt + c0, (uint64_t)rows, (uint64_t)cn, (uint64_t)ldc, KC, TF, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) return 1;
    }
    for (int cb = 0; cb < ncb; ++cb) {
        const int c0 = cb * CBLK, cn = (C - c0 < CBLK) ? C - c0 : CBLK;
        for (int db = 0; db < ndb; ++db) {
            const int d0 = db * DBLK, dn = (D - d0 < DBLK) ? D - d0 : DBLK;
            Params p;
            p.lengths = lengths; p.out_wx = out_wx + (size_t)c0 * D + d0; p.out_wsum = db == 0 ? out_wsum + c0 : nullptr;
            p.B = B; p.Tmax = Tmax; p.D = dn; p.C = cn; p.ldo = D;
            p.nchunk = (dn + KC - 1) / KC;
            weighted_sums_tc_kernel<TF><<<grid, THREADS, Geo<TF>::SMEM_BYTES, st>>>(mx[db], mw[cb], p);
            const int rc = check_launch("weighted_sums_tc_kernel");
            if (rc) return rc;
        }
    }
    return 0;
}

int launch_weighted_sums_tc(const float* X, const float* wgt, int ldc, const int32_t* lengths, int B, int Tmax, int D, int C,
                            float* out_wx, float* out_wsum, int num_sms, cudaStream_t st) {
    using namespace wtc;
    constexpr int DBLK = 7 * KC, CBLK = NPAD;
    if (D % 4 != 0 || D < 4 || C < 1 || ldc % 4 != 0 || ldc < C) return 1;
    if ((C + CBLK - 1) / CBLK > 3 || (D + DBLK - 1) / DBLK > 4) return 1;
    if ((reinterpret_cast<uintptr_t>(X) & 15) || (reinterpret_cast<uintptr_t>(wgt) & 15)) return 1;
    // tiles of 32 frames, one CTA per SM.  Tiles of 16 frames with two CTAs (two rings of three 36 KB stages) per SM were
    // measured (r02q): same results, same time (configs[1] step 4.84 vs 4.82 ms) -- unlike the emission kernel this one
    // is not bound by its ring
    return launch_weighted_sums_tc_t<32>(X, wgt, ldc, lengths, B, Tmax, D, C, out_wx, out_wsum, num_sms, st);
}

}  // namespace hsmm
