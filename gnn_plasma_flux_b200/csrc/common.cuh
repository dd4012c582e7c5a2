// Shared device helpers: mbarrier / bulk-copy (TMA) PTX wrappers, the packed
// weight layout, and the error plumbing of the C ABI.  sm_100a only.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/fluxgnn.h"

namespace fluxgnn {

constexpr int kF = FLUXGNN_INPUT_DIM;   // node feature width  (src/graph_constructor.py:32)
constexpr int kH = FLUXGNN_HIDDEN;      // hidden width        (src/config.py:21)
constexpr int kMaxL = FLUXGNN_MAX_LAYERS;
constexpr int kMaxHops = FLUXGNN_MAX_HOPS;

// ---------------------------------------------------------------------------
// Packed weight layout (floats).  Small vectors first, then the streamed
// matrices.  One "layer" of the stream is a [2H] x [H] Linear split in two
// K-major halves (src/flux_gnn.py:60,66 multiply [h, other] by W^T):
//   half 0 = the half applied to the neighbour mean / to h_col   (W[:, H:])
//   half 1 = the half applied to the node itself   / to h_row   (W[:, :H])
// each stored as Wt[k][q] = W[weight_column(q)][half_off + k], i.e. 128 rows of
// 128 floats with the columns permuted for the GEMM's register tile, cut into
// chunks of kChunkK rows -- the unit of one bulk copy.
// ---------------------------------------------------------------------------
#ifndef FLUXGNN_CHUNK_K
#define FLUXGNN_CHUNK_K 32
#endif
constexpr int kChunkK = FLUXGNN_CHUNK_K;          // k-rows per streamed chunk (multiple of 8)
constexpr int kChunkFloats = kChunkK * kH;        // 32 x 128 floats = 16 KiB
constexpr int kChunksPerHalf = kH / kChunkK;      // 4
static_assert(kChunkK % 8 == 0 && kH % kChunkK == 0, "chunk must tile K and keep the swizzle phase");
constexpr int kHalfFloats = kH * kH;              // 16384
constexpr int kLayerFloats = 2 * kHalfFloats;     // 32768

struct SmallParams {            // offsets (floats) inside the small block
    static constexpr int w_in = 0;                          // [F][H]  (feature-major)
    static constexpr int b_in = w_in + kF * kH;             // [H]
    static constexpr int b_upd = b_in + kH;                 // [kMaxL][H]
    static constexpr int b_e1 = b_upd + kMaxL * kH;         // [H]
    static constexpr int w_e2 = b_e1 + kH;                  // [H]
    static constexpr int b_e2 = w_e2 + kH;                  // [1] (+pad)
    static constexpr int count = b_e2 + kH;                 // padded to a multiple of 128 floats
};
static_assert(SmallParams::count % 128 == 0, "stream must start 512-byte aligned");

// Logical output feature stored at physical column q of a streamed weight row:
// thread tx of the GEMM reads physical columns 4tx..4tx+3 and 64+4tx..64+4tx+3 and
// owns the features tx + 16 j, j = 0..7 (see col_of() in hybrid_kernel_impl.cuh).
__host__ __device__ inline int weight_column(int q) {
    const int half = q >> 6, r = q & 63;
    return (r >> 2) + 16 * ((r & 3) + 4 * half);
}

__host__ __device__ inline size_t packed_floats(int L) {
    return (size_t)SmallParams::count + (size_t)(L + 1) * kLayerFloats;
}

// ---------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}

__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t"
        "}"
        ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}

// 1-D bulk copy global -> shared through the TMA unit (SASS: UBLKCP), completing
// `bytes` on the mbarrier.  dst/src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// Named barrier over a subset of the CTA's warps (id 1..15; 0 is __syncthreads).
__device__ __forceinline__ void named_sync(int id, int threads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

// ---------------------------------------------------------------------------
// error plumbing
// ---------------------------------------------------------------------------
int set_error(int code, const char* fmt, ...);
int cuda_fail(cudaError_t err, const char* what);
void count_launch(int n = 1);

#define FLUXGNN_CUDA_OK(expr)                                                   \
    do {                                                                         \
        cudaError_t err__ = (expr);                                              \
        if (err__ != cudaSuccess) return ::fluxgnn::cuda_fail(err__, #expr);    \
    } while (0)

}  // namespace fluxgnn
