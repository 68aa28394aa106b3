// Microbenchmark: how fast can one persistent CTA per SM stream a (rows x 200) fp32 matrix from HBM into a shared-memory
// ring, by access pattern?  No compute: a consumer warp releases every stage as soon as it lands.
//   mode 0  2-D tensor map, box 32 floats x 128 rows (128-byte swizzle), chunk after chunk of the same 128 rows
//           -- the emission kernel's pattern (7 boxes of 16 KB per 128-row tile, rows 800 bytes apart)
//   mode 1  1-D cp.async.bulk of contiguous 16 KB pieces (whole rows, no tensor map)
//   mode 2  2-D tensor map, box 32 floats x 32 rows (the tensor-core weighted-sums pattern, 4 KB boxes)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tma_stream_bench.bin tools/tma_stream_bench.cu -lcuda
// Run:   tools/tma_stream_bench.bin [stages] [rows]
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n.reg .pred p;\nWAIT_LOOP:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra WAIT_DONE;\nbra WAIT_LOOP;\nWAIT_DONE:\n}\n" ::"r"(
            smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                     smem_u32(dst)),
                 "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void bulk_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

constexpr int D = 200;
constexpr int STAGE = 16384;
constexpr int MAXS = 12;

__global__ void __launch_bounds__(64, 1)
stream_kernel(const __grid_constant__ CUtensorMap m128, const __grid_constant__ CUtensorMap m32,
              const float* X, long long rows, int mode, int nstage) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)nstage * STAGE);
    uint64_t* empty = full + MAXS;
    if (threadIdx.x == 0) {
        for (int s = 0; s < nstage; ++s) {
            mbar_init(full + s, 1);
            mbar_init(empty + s, 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const long long ntiles = rows / 128;   // 128-row tiles (102400 contiguous bytes each)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // every tile is moved as a sequence of stage-sized pieces; count them per mode
    if (lane != 0) return;
    int st = 0;
    uint32_t ph = 0;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        for (int pc = 0; pc < 7; ++pc) {
            if (warp == 0) {
                mbar_wait(empty + st, ph ^ 1);
                uint8_t* dst = smem + (size_t)st * STAGE;
                if (mode == 0) {
                    mbar_arrive_expect_tx(full + st, STAGE);
                    tma_load_2d(dst, &m128, full + st, pc * 32, (int)(tile * 128));
                } else if (mode == 1) {
                    const uint32_t bytes = (pc < 6) ? STAGE : (128 * D * 4 - 6 * STAGE);   // 6 x 16 KB + 4096
                    mbar_arrive_expect_tx(full + st, bytes);
                    bulk_load_1d(dst, reinterpret_cast<const uint8_t*>(X) + (size_t)tile * 128 * D * 4 + (size_t)pc * STAGE, bytes, full + st);
                } else {
                    // 4 boxes of 32 rows x 32 floats = one 16 KB stage: the 4 row groups of chunk pc
                    mbar_arrive_expect_tx(full + st, STAGE);
                    for (int q = 0; q < 4; ++q) tma_load_2d(dst + q * 4096, &m32, full + st, pc * 32, (int)(tile * 128 + q * 32));
                }
            } else {
                mbar_wait(full + st, ph);
                mbar_arrive(empty + st);
            }
            if (++st == nstage) {
                st = 0;
                ph ^= 1;
            }
        }
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static bool make_map(EncodeTiledFn fn, CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint32_t box_cols, uint32_t box_rows) {
    cuuint64_t gdim[2] = {cols, rows};
    cuuint64_t gstr[1] = {cols * sizeof(float)};
    cuuint32_t box[2] = {box_cols, box_rows};
    cuuint32_t estr[2] = {1, 1};
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

int main(int argc, char** argv) {
    const int nstage_arg = argc > 1 ? atoi(argv[1]) : 0;
    const long long rows = argc > 2 ? atoll(argv[2]) : 128LL * 148 * 64;   // 1.2 M rows = 970 MB
    float* X;
    CK(cudaMalloc(&X, (size_t)rows * D * 4 + 65536));
    CK(cudaMemset(X, 0, (size_t)rows * D * 4 + 65536));
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres));
    EncodeTiledFn fn = reinterpret_cast<EncodeTiledFn>(ptr);
    CUtensorMap m128, m32;
    if (!make_map(fn, &m128, X, rows, D, 32, 128) || !make_map(fn, &m32, X, rows, D, 32, 32)) {
        printf("tensor map encode failed\n");
        return 1;
    }
    int dev = 0, sms = 0;
    CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a));
    CK(cudaEventCreate(&b));
    const int stage_list[] = {3, 5, 8, 12};
    for (int mode = 0; mode < 3; ++mode) {
        for (int si = 0; si < 4; ++si) {
            const int ns = nstage_arg ? nstage_arg : stage_list[si];
            if (nstage_arg && si) break;
            const size_t smem = (size_t)ns * STAGE + 2 * MAXS * 8 + 64;
            CK(cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            float best = 1e9f;
            for (int rep = 0; rep < 4; ++rep) {
                CK(cudaEventRecord(a));
                stream_kernel<<<sms, 64, smem>>>(m128, m32, X, rows, mode, ns);
                CK(cudaEventRecord(b));
                CK(cudaEventSynchronize(b));
                float ms;
                CK(cudaEventElapsedTime(&ms, a, b));
                if (rep && ms < best) best = ms;
            }
            CK(cudaGetLastError());
            const double bytes = (mode == 1) ? (double)(rows / 128) * 128 * D * 4 : (double)(rows / 128) * 7 * STAGE * (200.0 / 224.0);
            printf("mode %d stages %2d (%3d KB in flight/SM): %.3f ms  %.0f GB/s of X\n", mode, ns, ns * 16, best, bytes / best / 1e6);
        }
    }
    return 0;
}
