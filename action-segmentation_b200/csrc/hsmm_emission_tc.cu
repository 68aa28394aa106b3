// Emission scoring on the 5th-generation tensor cores (sm_100a): TMA -> shared memory -> tcgen05.mma
// (kind::tf32, accumulators in TMEM) -> tcgen05.ld epilogue.
//
//   em'[t,c] = x_t . w_c + bias_c (+ penalty) - shift_t        (semimarkov_modules.py:324-381)
//
// is a (frames x D) . (D x C) contraction with arithmetic intensity C/2 flop/B: HBM-bound, so the job of
// the kernel is to stream X exactly once at full bandwidth while the tensor pipe does the flops.
// Plain TF32 (10-bit mantissa) is NOT accurate enough for the 1e-4 tolerance on the marginals at
// D = 200, so operands are split into a tf32-exact "big" part and a "small" remainder (x = xb + xs, w = wb + ws)
// and the MMAs accumulate xb.wb + xs.wb + xb.ws in fp32 (3xTF32: the dropped xs.ws term is ~2^-22 relative).
//
// Persistent CTAs (one per SM), 320 threads:
//   warp 0      TMA producer: X chunks of 128 frames x 32 floats (one 128-byte swizzle row per frame) into a ring
//   warp 1      MMA issuer (whole warp converged, one elected lane issues) + TMEM allocation.  Per k-step of 8 floats:
//               xb.[wb; ws] with A = the landed chunk in shared memory and N = 2 NPAD, then xs.wb with A = the
//               remainders in TENSOR MEMORY and N = NPAD.  Two accumulators of 2 NPAD columns (double buffer).
//   warps 2..5  converters, thread <-> frame <-> TMEM lane: read the landed row once (conflict-free through the
//               swizzle), accumulate the row term -0.5 sum x^2/var (handed to the epilogue warps through shared
//               memory) and store the remainders xs = x - (x & 0xffffe000) with tcgen05.st into the stage's 32 TMEM
//               columns.  They write no shared memory: the chunk itself is the big operand (kind::tf32 ignores the 13
//               low mantissa bits of a word, which the parity tests pin).
//   warps 6..9  epilogue (TMEM -> registers: the two accumulator halves added, bias, penalty, per-frame shift, row
//               term, f64 per-video offset), concurrently with the conversion of the following tiles.
// A tile is 128 consecutive frames of ONE video (the last tile of a video runs into the padding / the next video's
// rows, which are scored and dropped); only live tiles exist and they are dealt round-robin to the CTAs (TileCursor).
#include "hsmm_tc.cuh"

namespace hsmm {

namespace etc {

using namespace tc;

constexpr int TILE_M = 128;       // frames per tile = UMMA M
constexpr int KC = 32;            // floats per chunk = one 128-byte swizzle row
constexpr int CHUNK_BYTES = TILE_M * KC * 4;  // 16 KB
constexpr int STAGE_BYTES = CHUNK_BYTES;      // the landed chunk is the big operand; the remainders go to TMEM
constexpr int XS_BASE = 256;                  // TMEM columns [256, 256 + 32 nstage): remainders xs of the ring stages (A operand)
constexpr int THREADS = 64 + 128 + 128;        // producer, MMA issuer, 4 converter warps, 4 epilogue warps
constexpr int MAX_STAGES = 6;

struct Params {
    const float* bias;
    const float* inv_var;
    const float* penalty;
    const int32_t* lengths;
    float* em;
    float* rowterm;
    double* offset;
    const float* row_const;  // device scalar
    int B, Tmax, D, C, ldc;
    int wcols;    // columns of em this launch writes (a multiple of 4; == ldc unless `raw`)
    int raw;      // class-block passes for C > 64: 0 = the whole class set in one launch; 1 / 2 = first / later block of
                  // <= 64 classes: em gets the UNSHIFTED scores of the block's columns, rowterm (first block only) the
                  // class-independent term, offset nothing -- emission_finish_kernel shifts the rows afterwards
    int npad;     // classes padded to a multiple of 16 (UMMA N)
    int nchunk;   // ceil(D / 32)
    int nstage;   // ring depth
};

template <int NB, bool HALF, int MINB>  // NPAD = 16 * NB; HALF: 256 TMEM columns (NB <= 2, at most 4 stages), so that two CTAs share an SM
// MINB = 2 caps the kernel at 96 registers -- 30 K registers per CTA, so that two CTAs of the DP kernels (12-16 K
// registers each) stay resident beside it and use the issue slots this HBM-bound kernel leaves idle
__global__ void __launch_bounds__(THREADS, MINB)
emission_tc_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w, const Params p) {
    constexpr int NPAD = 16 * NB;
    // One accumulator = 2 NPAD TMEM columns: [0, NPAD) collects xb.wb + xs.wb, [NPAD, 2 NPAD) collects xb.ws (the big and
    // small class weights sit behind each other in shared memory, so xb meets both in ONE MMA with N = 2 NPAD: two MMAs
    // per k-step instead of three); the epilogue adds the two halves.  Two accumulators (double buffer).
    constexpr int ACC_STRIDE = 2 * NPAD;
    constexpr int TMEM_COLS = HALF ? 256 : 512;   // accumulators in [0, 4 NPAD), xs slots behind XSB
    constexpr int XSB = HALF ? XS_BASE / 2 : XS_BASE;
    static_assert(!HALF || NB <= 2, "256 TMEM columns: accumulators must end at column 128");
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // carve: [W: nchunk x (big NPAD x 128 B, small NPAD x 128 B)] [stages] [bias NPAD] [inv_var nchunk*32] [barriers]
    // (offset arithmetic on the __shared__ symbol keeps the address space visible to the compiler: LDS/STS, not generic LD/ST)
    uint8_t* base = smem_raw;  // 1024-byte aligned by declaration (the 128-byte swizzle atoms need it)
    if (smem_u32(smem_raw) & 1023u) __trap();
    uint8_t* w_s = base;
    const int w_chunk_bytes = 2 * NPAD * 128;
    uint8_t* st_s = w_s + (size_t)p.nchunk * w_chunk_bytes;
    float* bias_s = reinterpret_cast<float*>(st_s + (size_t)p.nstage * STAGE_BYTES);
    float* iv_s = bias_s + NPAD;
    uint64_t* bars = reinterpret_cast<uint64_t*>(iv_s + p.nchunk * KC);
    uint64_t* full = bars;                     // TMA -> converters
    uint64_t* conv = full + MAX_STAGES;        // converters -> MMA
    uint64_t* empty = conv + MAX_STAGES;       // MMA -> TMA
    uint64_t* tfull = empty + MAX_STAGES;      // MMA -> epilogue   [2]
    uint64_t* tempty = tfull + 2;              // epilogue -> MMA   [2]
    uint64_t* wbar = tempty + 2;
    uint64_t* rfull = wbar + 1;                // converters -> epilogue: row terms of a tile   [2]
    uint64_t* rempty = rfull + 2;              // epilogue -> converters                         [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(rempty + 2);
    float* rowsq_s = reinterpret_cast<float*>(tmem_slot + 4);  // [2][128]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int klast = (p.D - (p.nchunk - 1) * KC + 7) / 8;  // k-steps (of 8) in the last chunk

    if (threadIdx.x == 0) {
        for (int s = 0; s < p.nstage; ++s) {
            mbar_init(full + s, 1);
            mbar_init(conv + s, 128);
            mbar_init(empty + s, 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(tfull + a, 1);
            mbar_init(tempty + a, 128);
            mbar_init(rfull + a, 128);
            mbar_init(rempty + a, 128);
        }
        mbar_init(wbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (warp >= 2) {
        for (int i = threadIdx.x - 64; i < NPAD; i += THREADS - 64) bias_s[i] = (i < p.C) ? p.bias[i] : 0.0f;
        for (int i = threadIdx.x - 64; i < p.nchunk * KC; i += THREADS - 64) iv_s[i] = (i < p.D) ? p.inv_var[i] : 0.0f;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    TileCursor cur;   // every warp walks the same enumeration of live tiles
    int vb, vj, vlen;
    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            mbar_arrive_expect_tx(wbar, (uint32_t)(p.nchunk * w_chunk_bytes));
            for (int ch = 0; ch < p.nchunk; ++ch) {
                tma_load_2d(w_s + (size_t)ch * w_chunk_bytes, &tmap_w, wbar, ch * KC, 0);
                tma_load_2d(w_s + (size_t)ch * w_chunk_bytes + NPAD * 128, &tmap_w, wbar, ch * KC, NPAD);
            }
        }
        int st = 0;
        uint32_t ph = 0;
        for (int g = blockIdx.x; cur.locate(g, p.lengths, p.B, TILE_M, vb, vj, vlen); g += gridDim.x) {
            if (lane == 0) {
                const int row0 = vb * p.Tmax + vj * TILE_M;
                for (int ch = 0; ch < p.nchunk; ++ch) {
                    mbar_wait(empty + st, ph ^ 1);
                    mbar_arrive_expect_tx(full + st, CHUNK_BYTES);
                    tma_load_2d(st_s + (size_t)st * STAGE_BYTES, &tmap_x, full + st, ch * KC, row0);
                    if (++st == p.nstage) {
                        st = 0;
                        ph ^= 1;
                    }
                }
            }
            st = __shfl_sync(FULL, st, 0);
            ph = __shfl_sync(FULL, ph, 0);
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (whole warp converged, one elected lane issues) =====================
        const uint32_t idesc_n = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NPAD >> 3) << 17) | ((uint32_t)(TILE_M >> 4) << 24);
        const uint32_t idesc_2n = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(2 * NPAD >> 3) << 17) | ((uint32_t)(TILE_M >> 4) << 24);
        const uint32_t leader = elect_one();
        const int total_tiles = count_tiles(p.lengths, p.B, TILE_M);
        const int my_tiles = (total_tiles > (int)blockIdx.x) ? (total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
        mbar_wait(wbar, 0);
        const uint64_t desc0 = smem_desc_sw128(0);   // descriptor of shared address 0: the address field is added per MMA
        const uint32_t st0 = smem_u32(st_s), w0 = smem_u32(w_s);
        int st = 0;
        uint32_t ph = 0;
        int acc = 0;
        uint32_t accph = 0;
        for (int it = 0; it < my_tiles; ++it) {
            mbar_wait(tempty + acc, accph ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * ACC_STRIDE;
            uint32_t accum = 0;
            for (int ch = 0; ch < p.nchunk; ++ch) {
                mbar_wait(conv + st, ph);
                tc_fence_after();
                const uint64_t xb = desc_at(desc0, st0 + st * STAGE_BYTES);
                const uint32_t xs = tmem_base + XSB + st * KC;   // 128 lanes x 32 columns of remainders
                const uint64_t wb = desc_at(desc0, w0 + ch * w_chunk_bytes);   // NPAD rows big, then NPAD rows small
                const int ksteps = (ch == p.nchunk - 1) ? klast : 4;
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    if (ks < ksteps) {
                        const uint32_t ko = ks * 32;  // 8 tf32 = 32 bytes inside the 128-byte swizzle row
                        tc_mma_tf32_lead(d_tmem, desc_at(xb, ko), desc_at(wb, ko), idesc_2n, accum, leader);   // xb.[wb; ws]
                        accum = 1;
                        tc_mma_tf32_ts_lead(d_tmem, xs + ks * 8, desc_at(wb, ko), idesc_n, 1, leader);        // xs.wb
                    }
                }
                tc_commit_lead(empty + st, leader);  // the stage may be refilled once these MMAs have read it
                if (++st == p.nstage) {
                    st = 0;
                    ph ^= 1;
                }
            }
            tc_commit_lead(tfull + acc, leader);
            if (++acc == 2) {
                acc = 0;
                accph ^= 1;
            }
        }
    } else if (warp < 6) {
        // ===================== converters (128 threads, thread <-> frame = TMEM lane) =====================
        // x = xb + xs with xb = the 19 upper bits of x: the landed chunk ITSELF is the big operand (kind::tf32 reads the
        // upper 19 bits of a word and ignores the 13 low mantissa bits -- pinned by the parity tests), and the remainder
        // xs goes straight from registers to tensor memory, the A operand of the second MMA.  The converters write no
        // shared memory at all: per 16 KB chunk the port carries TMA 16 + this read 16 + operand reads 16 + 12 KB.
        const int q = warp & 3;               // TMEM lane quarter this warp may access
        const int r = q * 32 + lane;          // row inside the tile
        int st = 0;
        uint32_t ph = 0;
        int acc = 0;
        uint32_t accph = 0;
        for (int g = blockIdx.x; cur.locate(g, p.lengths, p.B, TILE_M, vb, vj, vlen); g += gridDim.x) {
            float rowsq = 0.0f;
            for (int ch = 0; ch < p.nchunk; ++ch) {
                mbar_wait(full + st, ph);
                const uint8_t* xrow = st_s + (size_t)st * STAGE_BYTES + (size_t)r * 128;
                float4 x[8];
#pragma unroll
                for (int lj = 0; lj < 8; ++lj)   // logical 16-byte unit lj sits in physical slot lj ^ (r & 7): conflict-free
                    x[lj] = *reinterpret_cast<const float4*>(xrow + ((lj ^ (r & 7)) * 16));
                float q0 = 0.0f, q1 = 0.0f;
                float sm[32];
#pragma unroll
                for (int lj = 0; lj < 8; ++lj) {
                    const float4 iv = *reinterpret_cast<const float4*>(iv_s + ch * KC + lj * 4);
                    q0 = fmaf(x[lj].x * x[lj].x, iv.x, q0);
                    q1 = fmaf(x[lj].y * x[lj].y, iv.y, q1);
                    q0 = fmaf(x[lj].z * x[lj].z, iv.z, q0);
                    q1 = fmaf(x[lj].w * x[lj].w, iv.w, q1);
                    sm[4 * lj + 0] = x[lj].x - __uint_as_float(__float_as_uint(x[lj].x) & TF32_MASK);
                    sm[4 * lj + 1] = x[lj].y - __uint_as_float(__float_as_uint(x[lj].y) & TF32_MASK);
                    sm[4 * lj + 2] = x[lj].z - __uint_as_float(__float_as_uint(x[lj].z) & TF32_MASK);
                    sm[4 * lj + 3] = x[lj].w - __uint_as_float(__float_as_uint(x[lj].w) & TF32_MASK);
                }
                rowsq += q0 + q1;
                tc_st32(tmem_base + ((uint32_t)(q * 32) << 16) + XSB + st * KC, sm);
                tc_wait_st();
                tc_fence_before();
                mbar_arrive(conv + st);
                if (++st == p.nstage) {
                    st = 0;
                    ph ^= 1;
                }
            }
            // the tile's row terms go to the epilogue warps (slot = accumulator parity)
            mbar_wait(rempty + acc, accph ^ 1);
            rowsq_s[acc * TILE_M + r] = rowsq;
            mbar_arrive(rfull + acc);
            if (++acc == 2) {
                acc = 0;
                accph ^= 1;
            }
        }
    } else {
        // ===================== epilogue (128 threads, thread <-> frame) =====================
        const int q = warp & 3;              // TMEM lane quarter this warp may read
        const int r = q * 32 + lane;         // row inside the tile
        const long long total_rows = (long long)p.B * p.Tmax;
        const float row_const = __ldg(p.row_const);
        int acc = 0;
        uint32_t accph = 0;

        // Rows behind the last live tile of their video are zero-filled here (the interface promises em = 0,
        // rowterm = 0 for t >= length), in blocks of 128 flattened rows dealt round-robin; this runs while the
        // ring fills, before the first accumulator is ready.
        {
            const long long nblk = (total_rows + TILE_M - 1) / TILE_M;
            for (long long fb = blockIdx.x; fb < nblk; fb += gridDim.x) {
                const long long row = fb * TILE_M + r;
                if (row >= total_rows) continue;
                const int b = (int)(row / p.Tmax);
                const int t = (int)(row - (long long)b * p.Tmax);
                const int len = max(p.lengths[b], 0);
                if (t < (len + TILE_M - 1) / TILE_M * TILE_M) continue;
                float* em_r = p.em + (size_t)row * p.ldc;
                for (int c4 = 0; c4 * 4 < p.wcols; ++c4) *reinterpret_cast<float4*>(em_r + c4 * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
                if (p.raw < 2) p.rowterm[row] = 0.0f;
            }
        }

        for (int g = blockIdx.x; cur.locate(g, p.lengths, p.B, TILE_M, vb, vj, vlen); g += gridDim.x) {
            mbar_wait(rfull + acc, accph);
            const float rowsq = rowsq_s[acc * TILE_M + r];
            mbar_arrive(rempty + acc);

            const int t = vj * TILE_M + r;
            const bool in_range = t < p.Tmax;   // rows past Tmax belong to the next video: scored, not stored
            const bool live = t < vlen;
            const long long row = (long long)vb * p.Tmax + t;
            float v[NPAD];
            mbar_wait(tfull + acc, accph);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * ACC_STRIDE;
#pragma unroll
            for (int i = 0; i < NB; ++i) {
                float lo[16];
                tc_ld16(taddr + 16 * i, v + 16 * i);
                tc_ld16(taddr + NPAD + 16 * i, lo);
                tc_wait_ld();
#pragma unroll
                for (int c = 0; c < 16; ++c) v[16 * i + c] += lo[c];
            }
            tc_fence_before();
            mbar_arrive(tempty + acc);

            double contrib = 0.0;
            if (in_range) {
                float* em_r = p.em + (size_t)row * p.ldc;
                float rt = 0.0f;
                if (live && p.raw) {
#pragma unroll
                    for (int c = 0; c < NPAD; ++c) v[c] = (c < p.C) ? v[c] + bias_s[c] : 0.0f;
                    rt = -0.5f * rowsq + row_const;
                } else if (live) {
                    const float* pen_r = p.penalty ? p.penalty + (size_t)row * p.C : nullptr;
                    float m = NEG;
#pragma unroll
                    for (int c = 0; c < NPAD; ++c) {
                        v[c] += bias_s[c];
                        if (c < p.C) m = fmaxf(m, v[c] + (pen_r ? __ldg(pen_r + c) : 0.0f));
                    }
                    // shift by the best penalised score; the penalty (-1e4 per offending frame) is added AFTER
                    // the shift so that the large number meets an O(1) one exactly once, like the reference's
                    // elp + constraints (semimarkov_modules.py:379-380)
#pragma unroll
                    for (int c = 0; c < NPAD; ++c) {
                        float o = 0.0f;
                        if (c < p.C) {
                            o = v[c] - m;
                            if (pen_r) o += __ldg(pen_r + c);
                        }
                        v[c] = o;
                    }
                    rt = (-0.5f * rowsq + row_const) + m;
                    contrib = (double)rt;
                } else {
#pragma unroll
                    for (int c = 0; c < NPAD; ++c) v[c] = 0.0f;
                }
#pragma unroll
                for (int c4 = 0; c4 < NPAD / 4; ++c4)
                    if (c4 * 4 < p.wcols)
                        *reinterpret_cast<float4*>(em_r + c4 * 4) = make_float4(v[c4 * 4], v[c4 * 4 + 1], v[c4 * 4 + 2], v[c4 * 4 + 3]);
                if (p.raw < 2) p.rowterm[row] = rt;
            }
            // per-video offset: the tile lies inside one video, one f64 atomic per warp
            contrib = warp_sum(contrib);
            if (lane == 0 && contrib != 0.0) atomicAdd(p.offset + vb, contrib);

            if (++acc == 2) {
                acc = 0;
                accph ^= 1;
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// w (C, D) -> wsplit (2*npad, dpad): rows [0, npad) tf32-exact big parts, rows [npad, 2 npad) remainders; zero padded
__global__ void emission_split_w_kernel(const float* __restrict__ w, int C, int D, int npad, int dpad, float* __restrict__ out) {
    const int n = npad * dpad;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int c = i / dpad, d = i - c * dpad;
        float x = (c < C && d < D) ? w[(size_t)c * D + d] : 0.0f;
        const float big = __uint_as_float(__float_as_uint(x) & TF32_MASK);
        out[i] = big;
        out[n + i] = x - big;
    }
}

// Second half of the class-block passes (C > 64): every live row of em holds the unshifted scores of all C classes and
// rowterm the class-independent term.  One warp per row: m = max_c (score + penalty), em = score - m (+ penalty),
// rowterm += m, offset[b] += sum of rowterm -- what the single-launch epilogue does in registers.
constexpr int FIN_ROWS = 64;  // rows of one video per CTA
__global__ void __launch_bounds__(256) emission_finish_kernel(float* __restrict__ em, int ldc, int C, const float* __restrict__ penalty,
                                                              const int32_t* __restrict__ lengths, int B, int Tmax,
                                                              float* __restrict__ rowterm, double* __restrict__ offset) {
    const int tpv = (Tmax + FIN_ROWS - 1) / FIN_ROWS;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long ntiles = (long long)B * tpv;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int b = (int)(tile / tpv);
        const int t0 = (int)(tile - (long long)b * tpv) * FIN_ROWS;
        const int len = min(max(lengths[b], 0), Tmax);
        if (t0 >= len) continue;
        double contrib = 0.0;
        for (int t = t0 + warp; t < min(t0 + FIN_ROWS, len); t += 8) {
            const long long row = (long long)b * Tmax + t;
            float* em_r = em + (size_t)row * ldc;
            const float* pen_r = penalty ? penalty + (size_t)row * C : nullptr;
            float m = NEG;
            for (int c = lane; c < C; c += 32) m = fmaxf(m, em_r[c] + (pen_r ? __ldg(pen_r + c) : 0.0f));
            m = warp_max(m);
            for (int c = lane; c < C; c += 32) {
                float o = em_r[c] - m;
                if (pen_r) o += __ldg(pen_r + c);
                em_r[c] = o;
            }
            if (lane == 0) {
                const float rt = rowterm[row] + m;
                rowterm[row] = rt;
                contrib += (double)rt;
            }
        }
        if (lane == 0 && contrib != 0.0) atomicAdd(offset + b, contrib);
    }
}

struct Plan {
    int npad, nchunk, nstage;
    size_t smem;
    bool half;  // two CTAs per SM, 256 TMEM columns each
};
constexpr int CBLK = 64;      // classes per pass when C > 64
constexpr int MAX_CBLK = 8;   // C <= 512

static bool plan(int D, int C, Plan* pl) {
    if (C > 64 || D % 4 != 0 || D < 4) return false;
    pl->npad = (C + 15) / 16 * 16;
    pl->nchunk = (D + KC - 1) / KC;
    const size_t fixed = (size_t)pl->nchunk * 2 * pl->npad * 128 + (size_t)(pl->npad + pl->nchunk * KC) * 4 + (3 * MAX_STAGES + 9) * 8 + 16 +
                         2 * (size_t)TILE_M * 4;
    // 8 KB of the 227 KB stay free when the ring still gets three stages: the DP kernels' CTAs (~1-2 KB each) must fit
    // beside this one (measured: the ring depth beyond three stages does not matter, the TMA pattern alone reaches
    // 6 TB/s with three -- tools/tma_stream_bench.cu)
    size_t cap = 227 * 1024;
    if (fixed + 3 * (size_t)STAGE_BYTES <= cap - 8 * 1024) cap -= 8 * 1024;
    if (fixed + 2 * (size_t)STAGE_BYTES > cap) return false;
    int ns = (int)((cap - fixed) / STAGE_BYTES);
    if (ns > MAX_STAGES) ns = MAX_STAGES;
    pl->half = false;
    {
        // two co-resident CTAs (each its own producer / converters / issuer / epilogue, 256 TMEM columns, 3-4 stages):
        // one CTA's ring refill overlaps the other's, and the prologue of the next task's launch overlaps this one's tail
        static const int two = [] { const char* e = getenv("HSMM_EMISSION_TWO_CTAS"); return e ? atoi(e) : 1; }();  // A/B switch, r02q: 934 -> 795 us
        const size_t per_cta = (227 * 1024) / 2 - 1024;
        if (two && pl->npad <= 32 && fixed + 3 * (size_t)STAGE_BYTES <= per_cta) {
            ns = (int)((per_cta - fixed) / STAGE_BYTES);
            if (ns > 4) ns = 4;
            pl->half = true;
        }
    }
    pl->nstage = ns;
    pl->smem = fixed + (size_t)ns * STAGE_BYTES;
    if (pl->half) return true;
    // the kernel allocates all 512 TMEM columns: never two of its CTAs on one SM (the second would sit in tcgen05.alloc)
    if (pl->smem < 117 * 1024) pl->smem = 117 * 1024;
    return true;
}

}  // namespace etc

size_t emission_tc_workspace_bytes(int D, int C) {
    etc::Plan pl;
    size_t total = 0;
    if (C > etc::CBLK * etc::MAX_CBLK) return 0;
    for (int c0 = 0; c0 < C; c0 += etc::CBLK) {  // one split weight table per class block
        if (!etc::plan(D, C - c0 < etc::CBLK ? C - c0 : etc::CBLK, &pl)) return 0;
        total += (size_t)2 * pl.npad * pl.nchunk * etc::KC * sizeof(float);
    }
    return total;
}

// returns 1 when the shape / alignment is not eligible (caller falls back to the SIMT kernel), 0 on launch, < 0 on error
int launch_emission_tc(const float* X, const float* w, const float* bias, const float* inv_var, const float* row_const,
                       const float* penalty, const int32_t* lengths, int B, int Tmax, int D, int C, int ldc, float* em,
                       float* rowterm, double* offset, void* workspace, int num_sms, cudaStream_t st) {
    using namespace etc;
    if (!workspace || C < 1 || C > CBLK * MAX_CBLK) return 1;
    if ((reinterpret_cast<uintptr_t>(X) & 15) || (reinterpret_cast<uintptr_t>(workspace) & 15) ||
        (reinterpret_cast<uintptr_t>(em) & 15) || ldc % 4 != 0 || ldc < C)
        return 1;
    const long long rows = (long long)B * Tmax;
    if (rows >= (1ll << 31) - TILE_M) return 1;
    // C <= 64: one launch scores, shifts and stores.  C > 64: one launch per block of 64 classes (X is re-read per block:
    // 800 B a frame against the block's 64 x D multiply-adds the SIMT kernel would issue) + the finishing kernel
    const int nblk = (C + CBLK - 1) / CBLK;
    Plan pl[MAX_CBLK];
    CUtensorMap mx, mw[MAX_CBLK];
    float* wsp[MAX_CBLK];
    {
        float* wp = reinterpret_cast<float*>(workspace);
        for (int k = 0; k < nblk; ++k) {
            const int c0 = k * CBLK, cn = C - c0 < CBLK ? C - c0 : CBLK;
            if (!plan(D, cn, &pl[k])) return 1;
            const int wcols = (k + 1 < nblk) ? CBLK : ldc - c0;
            if (wcols > pl[k].npad) return 1;
            const int dpad = pl[k].nchunk * KC;
            wsp[k] = wp;
            if (!make_map(&mw[k], wp, (uint64_t)(2 * pl[k].npad), (uint64_t)dpad, (uint64_t)dpad, KC, (uint32_t)pl[k].npad)) return 1;
            wp += (size_t)2 * pl[k].npad * dpad;
        }
    }
    if (!make_map(&mx, X, (uint64_t)rows, (uint64_t)D, (uint64_t)D, KC, TILE_M)) return 1;

    cudaError_t e = cudaMemsetAsync(offset, 0, sizeof(double) * B, st);
    if (e != cudaSuccess) {
        set_error("memset offset: %s", cudaGetErrorString(e));
        return -3;
    }
    const long long max_tiles = (long long)B * ((Tmax + TILE_M - 1) / TILE_M);
    for (int k = 0; k < nblk; ++k) {
        const long long want = pl[k].half ? 2ll * num_sms : num_sms;
        int grid = want < max_tiles ? (int)want : (int)max_tiles;
        if (grid < 1) grid = 1;
        const int c0 = k * CBLK, cn = C - c0 < CBLK ? C - c0 : CBLK;
        const int dpad = pl[k].nchunk * KC;
        emission_split_w_kernel<<<(pl[k].npad * dpad + 255) / 256, 256, 0, st>>>(w + (size_t)c0 * D, cn, D, pl[k].npad, dpad, wsp[k]);
        int rc = check_launch("emission_split_w_kernel");
        if (rc) return rc;

        Params p;
        p.bias = bias + c0; p.inv_var = inv_var; p.penalty = penalty; p.lengths = lengths; p.em = em + c0; p.rowterm = rowterm;
        p.offset = offset; p.row_const = row_const; p.B = B; p.Tmax = Tmax; p.D = D; p.C = cn; p.ldc = ldc;
        p.raw = nblk == 1 ? 0 : (k == 0 ? 1 : 2);
        p.wcols = nblk == 1 ? ldc : ((k + 1 < nblk) ? CBLK : ldc - c0);
        p.npad = pl[k].npad; p.nchunk = pl[k].nchunk; p.nstage = pl[k].nstage;

#define HSMM_ETC_LAUNCH(NB, H, M)                                                                                           \
    {                                                                                                                   \
        e = cudaFuncSetAttribute(emission_tc_kernel<NB, H, M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl[k].smem); \
        if (e == cudaSuccess) emission_tc_kernel<NB, H, M><<<grid, THREADS, pl[k].smem, st>>>(mx, mw[k], p);                    \
    }
        // min-blocks 2 (<= 96 registers) up to 32 classes: two of these CTAs, or one beside two DP CTAs, share an SM.  Above,
        // shared memory allows one CTA per SM anyway and the epilogue's 4 NB + 16 accumulator registers per row spill under
        // that cap (r02s, graph-timed, 1.8 M frames: C = 48 2.33 -> 2.76 TB/s, C = 64 1.27 -> 2.22 TB/s without it)
        switch (pl[k].npad / 16 + (pl[k].half ? 10 : 0)) {
            case 1: HSMM_ETC_LAUNCH(1, false, 2) break;
            case 2: HSMM_ETC_LAUNCH(2, false, 2) break;
            case 3: HSMM_ETC_LAUNCH(3, false, 1) break;
            case 11: HSMM_ETC_LAUNCH(1, true, 2) break;
            case 12: HSMM_ETC_LAUNCH(2, true, 2) break;
            default: HSMM_ETC_LAUNCH(4, false, 1) break;
        }
#undef HSMM_ETC_LAUNCH
        if (e != cudaSuccess) {
            set_error("emission_tc smem attr: %s", cudaGetErrorString(e));
            return -3;
        }
        rc = check_launch("emission_tc_kernel");
        if (rc) return rc;
    }
    if (nblk > 1) {
        long long ft = (long long)B * ((Tmax + FIN_ROWS - 1) / FIN_ROWS);
        const int fgrid = (int)(ft < (long long)num_sms * 8 ? ft : (long long)num_sms * 8);
        emission_finish_kernel<<<fgrid < 1 ? 1 : fgrid, 256, 0, st>>>(em, ldc, C, penalty, lengths, B, Tmax, rowterm, offset);
        return check_launch("emission_finish_kernel");
    }
    return 0;
}

}  // namespace hsmm
