// Micro-benchmark: cycles per tcgen05.mma (kind::tf32 / kind::f16) for the issue patterns the tile kernel
// can use.  One CTA per SM, one issuing thread, operands = zero-filled shared memory.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I gnn_plasma_flux_b200/csrc -o umma_probe scripts/probes/umma_probe.cu
#include <cstdio>
#include <cstdlib>
#include "common.cuh"
using namespace fluxgnn;

__device__ __forceinline__ void umma_f16(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
// kind::f16 instruction descriptor: fp16 A/B (format 0), fp32 accumulate
// descriptor with a chosen layout type: 2 = SWIZZLE_128B, 4 = SWIZZLE_64B, 6 = SWIZZLE_32B, 0 = none (LBO/SBO for 8x16B core matrices)
__device__ __forceinline__ uint64_t desc_layout(uint32_t addr, int layout) {
    if (layout == 2) return umma_desc_sw128(addr);
    const uint64_t sbo = layout == 4 ? 512 : layout == 6 ? 256 : 128;     // 8-row group pitch
    const uint64_t lbo = layout == 0 ? 4096 : 16;                          // none: next K core matrix far away
    return (uint64_t)((addr & 0x3FFFF) >> 4) | ((lbo >> 4) << 16) | ((sbo >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)layout << 61);
}
__host__ __device__ constexpr uint32_t idesc_f16(int M, int N) { return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24); }

struct Params { int kind16, N, nacc, per_commit, same_a, reps, M = 128, layout = 2, kstep = 32; };

__global__ void __launch_bounds__(128, 1) probe(Params p, long long* out) {
    extern __shared__ __align__(1024) unsigned char raw[];
    unsigned char* sm = (unsigned char*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar, bar2;
    __shared__ uint32_t tmem_base;
    for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) ((uint32_t*)sm)[i] = 0;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_init(&bar2, 1); mbar_fence_init(); }
    if (threadIdx.x < 32) tmem_alloc(&tmem_base, 512);
    tc_fence_before(); __syncthreads(); tc_fence_after();
    fence_proxy_async();
    if (threadIdx.x == 0) {
        const uint32_t tm = tmem_base;
        const uint32_t a0 = smem_u32(sm), b0 = smem_u32(sm + 64 * 1024);     // A: 4 units of 16 KiB; B: up to 256 rows x 128 B x 2
        const uint32_t idesc = p.kind16 ? idesc_f16(p.M, p.N) : umma_idesc_tf32(p.M, p.N);
        uint32_t phase = 0;
        long long t0 = clock64();
        int n = 0;
        for (int r = 0; r < p.reps; ++r) {
            for (int u = 0; u < 16; ++u) {                       // 16 "units" of 4 k-steps
                const uint32_t d = tm + (uint32_t)((u % p.nacc) * p.N);
                const uint32_t wa = a0 + (p.same_a ? 0 : (u & 3) * 16384);
                for (int ks = 0; ks < 4; ++ks) {
                    const uint64_t ad = desc_layout(wa + ks * p.kstep, p.layout), bd = desc_layout(b0 + ks * p.kstep, p.layout);
                    if (p.kind16) umma_f16(d, ad, bd, idesc, 1); else umma_tf32(d, ad, bd, idesc, 1);
                    ++n;
                }
                if (p.per_commit && (u % p.per_commit) == p.per_commit - 1) {
                    umma_commit(&bar2);                          // stage-release style commit nobody waits on
                }
            }
        }
        umma_commit(&bar);
        // wait for the final commit: count phases completed so far
        mbar_wait(&bar, phase);
        long long t1 = clock64();
        if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = n; }
    }
    tc_fence_before(); __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tmem_base, 512);
}

int main() {
    long long* out; cudaMalloc(&out, 16);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    struct { const char* name; Params p; } cases[] = {
        {"tf32 N=128, 1 accumulator, no intermediate commits", {0, 128, 1, 0, 0, 8}},
        {"tf32 N=128, 2 accumulators alternating per unit", {0, 128, 2, 0, 0, 8}},
        {"tf32 N=128, 2 accumulators, commit per unit", {0, 128, 2, 1, 0, 8}},
        {"tf32 N=128, 2 accumulators, commit per 4 units", {0, 128, 2, 4, 0, 8}},
        {"tf32 N=128, 4 accumulators", {0, 128, 4, 0, 0, 8}},
        {"tf32 N=256, 1 accumulator", {0, 256, 1, 0, 0, 8}},
        {"tf32 N=256, 2 accumulators", {0, 256, 2, 0, 0, 8}},
        {"tf32 N=64, 2 accumulators", {0, 64, 2, 0, 0, 8}},
        {"tf32 N=128, 2 acc, same A unit", {0, 128, 2, 0, 1, 8}},
        {"f16 N=128, 2 accumulators", {1, 128, 2, 0, 0, 8}},
        {"f16 N=128, 2 accumulators, commit per unit", {1, 128, 2, 1, 0, 8}},
        {"f16 N=256, 2 accumulators", {1, 256, 2, 0, 0, 8}},
        {"f16 N=64, 2 accumulators", {1, 64, 2, 0, 0, 8}},
        {"f16 N=32, 2 accumulators", {1, 32, 2, 0, 0, 8}},
        {"f16 N=160, 2 accumulators", {1, 160, 2, 0, 0, 8}},
        {"f16 N=192, 2 accumulators", {1, 192, 2, 0, 0, 8}},
        {"f16 N=224, 2 accumulators", {1, 224, 2, 0, 0, 8}},
        {"f16 M=64 N=128, 2 accumulators", {1, 128, 2, 0, 0, 8, 64}},
        {"f16 M=64 N=256, 2 accumulators", {1, 256, 2, 0, 0, 8, 64}},
        {"f16 N=128, SWIZZLE_64B", {1, 128, 2, 0, 0, 8, 128, 4}},
        {"f16 N=128, SWIZZLE_32B", {1, 128, 2, 0, 0, 8, 128, 6}},
        {"f16 N=128, no swizzle", {1, 128, 2, 0, 0, 8, 128, 0, 256}},
        {"f16 N=256, SWIZZLE_32B", {1, 256, 2, 0, 0, 8, 128, 6}},
        {"f16 N=256, no swizzle", {1, 256, 2, 0, 0, 8, 128, 0, 256}},
    };
    for (auto& c : cases) {
        for (int grid : {148}) {
            cudaMemset(out, 0, 16);
            probe<<<grid, 128, 200 * 1024>>>(c.p, out);
            cudaError_t e = cudaDeviceSynchronize();
            long long h[2]; cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
            printf("%-55s grid %3d: %7.1f clk per UMMA (%lld UMMAs)%s\n", c.name, grid, (double)h[0] / (double)h[1], h[1],
                   e == cudaSuccess ? "" : cudaGetErrorString(e));
        }
    }
    return 0;
}
