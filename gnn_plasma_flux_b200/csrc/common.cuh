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

// ---- thread-block clusters: rank, remote shared-memory addresses, cluster-scope mbarrier signalling ----
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t cluster_map(uint32_t local_smem_addr, uint32_t rank) {      // shared::cluster address
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ float4 ld_cluster_f4(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared::cluster.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void cluster_sync_all() {         // every thread of every CTA of the cluster
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {      // release at cluster scope
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {   // acquire at cluster scope
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAITC_LOOP:\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra WAITC_DONE;\n\t"
        "bra WAITC_LOOP;\n\t"
        "WAITC_DONE:\n\t"
        "}"
        ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}

// Named barrier over a subset of the CTA's warps (id 1..15; 0 is __syncthreads).
__device__ __forceinline__ void named_sync(int id, int threads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

// ---------------------------------------------------------------------------
// tcgen05 (5th-generation tensor core) wrappers: TMEM allocation, UMMA issue,
// completion -> mbarrier, TMEM -> register loads.  cta_group::1 only.
// ---------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result, uint32_t ncols) {   // one full warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
                 "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {        // the allocating warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// generic-proxy shared-memory writes -> visible to the async proxy (UMMA operand reads)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// Shared-memory matrix descriptor, K-major operand, SWIZZLE_128B: rows of 128 bytes, 8-row
// groups 1024 bytes apart (SBO), start address in 16-byte units, descriptor version 1 (sm_100).
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
    return (uint64_t)((smem_addr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}

// Instruction descriptor for kind::tf32, fp32 accumulate, A and B K-major, shape M x N (x8).
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]; one thread issues for the CTA.
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// All previously issued UMMAs of this thread complete -> one arrival on the mbarrier
// (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}

// One lane of the converged warp (elect.sync).
__device__ __forceinline__ bool elect_one_lane() {
    uint32_t pred;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(pred));
    return pred != 0;
}

// TMEM -> registers: the warp's 32 lanes x `N` consecutive 32-bit columns, thread = lane.
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void tmem_ld4(uint32_t taddr, float (&v)[4]) {
    uint32_t r[4];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(taddr)
                 : "memory");
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = __uint_as_float(r[i]);
}

// fp32 -> tf32 (round to nearest, ties away): an fp32 bit pattern with the low 13 mantissa bits clear
__device__ __forceinline__ float to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

// ---------------------------------------------------------------------------
// Tensor-path weight stream (fluxgnn_pack_weights_tc): small block as above, then per
// layer 16 "units" of 16 KiB, each the exact shared-memory image of a K-major,
// 128-byte-swizzled UMMA A operand [128 output features] x [32 k]:
//   unit index = (kb * 2 + blk) * 2 + part,  kb = k / 32 (0..3),
//   blk  0 = W[:, H:] (neighbour / h_col half), 1 = W[:, :H] (self / h_row half),
//   part 0 = tf32(W), 1 = tf32(W - tf32(W))   (the 3xTF32 split; plain tf32 skips part 1)
//   element (n, kk) at byte n*128 + (((kk >> 2) ^ (n & 7)) << 4) + (kk & 3) * 4.
// ---------------------------------------------------------------------------
constexpr int kTcUnitFloats = kH * 32;                 // 4096 floats = 16 KiB
constexpr int kTcUnitsPerLayer = 16;
__host__ __device__ inline size_t packed_tc_floats(int L) {
    return (size_t)SmallParams::count + (size_t)(L + 1) * kTcUnitsPerLayer * kTcUnitFloats;
}

// ---------------------------------------------------------------------------
// 16-bit tensor-path weight stream (fluxgnn_pack_weights_tc16; hybrid_tc16_kernel.cu): small block as
// above (fp32), then per layer 8 units of 16 KiB, each the shared-memory image of a K-major,
// 128-byte-swizzled UMMA A operand [128 output features] x [64 k] of 16-bit numbers:
//   unit index = (kb * 2 + blk) * 2 + part,  kb = k / 64 (0..1), blk as above,
//   part 0 = r16(S W), 1 = r16(S W - part 0),  S = kTc16WeightScale, r16 = fp16 or bf16 rounding
//   element (n, kk) at byte n*128 + (((kk >> 3) ^ (n & 7)) << 4) + (kk & 7) * 2.
// ---------------------------------------------------------------------------
constexpr int kTc16TileRows = 256;
constexpr float kTc16WeightScale = 256.0f;
constexpr int kTc16UnitBytes = kH * 64 * 2;            // 16 KiB
constexpr int kTc16UnitsPerLayer = 8;
__host__ __device__ inline size_t packed_tc16_bytes(int L) {
    return (size_t)SmallParams::count * sizeof(float) + (size_t)(L + 1) * kTc16UnitsPerLayer * kTc16UnitBytes;
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
