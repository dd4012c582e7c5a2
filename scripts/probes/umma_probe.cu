// Micro-benchmark: cycles per tcgen05.mma (kind::tf32 / kind::f16) by shape and by issue style.
// One CTA per SM, operands = zero-filled shared memory, everything compile-time so that the issue
// loop is as tight as the tile kernel's.
//   kUniform = true : the whole warp walks the loop, one elected lane issues (descriptors stay in uniform
//                     registers, UTCHMMA instructions go out back to back) -- measures the tensor pipe;
//   kUniform = false: everything inside `if (threadIdx.x == 0)` -- the compiler wraps every UTCHMMA in R2UR
//                     moves and an ELECT / BRA.U.ANY loop: the measurement is that lone thread's issue cost.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I gnn_plasma_flux_b200/csrc \
//        -o scripts/probes/umma_probe scripts/probes/umma_probe.cu
#include <cstdio>
#include <cstdlib>
#include "common.cuh"
using namespace fluxgnn;

__device__ __forceinline__ void umma_f16(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__host__ __device__ constexpr uint32_t idesc_f16(int M, int N) { return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24); }

constexpr int kReps = 8, kUnits = 16;      // 16 units of 4 k-steps per repetition (the tile kernel's weight units)

template <bool kF16, int M, int N, int kAcc, bool kCommitPerUnit, bool kUniform, bool kBmn = false>
__global__ void __launch_bounds__(128, 1) probe(long long* out) {
    extern __shared__ __align__(1024) unsigned char sm[];
    __shared__ uint64_t bar, bar2;
    __shared__ uint32_t tmem_base;
    for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) ((uint32_t*)sm)[i] = 0;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_init(&bar2, 1); mbar_fence_init(); }
    if (threadIdx.x < 32) tmem_alloc(&tmem_base, 512);
    tc_fence_before(); __syncthreads(); tc_fence_after();
    fence_proxy_async();
    const uint32_t idesc = (kF16 ? idesc_f16(M, N) : umma_idesc_tf32(M, N)) | (kBmn ? (1u << 16) : 0u);
    // kBmn: B operand MN-major (rows contiguous), SWIZZLE_128B atoms of [8 k][64 rows], LBO 16 KiB, SBO 1 KiB
    const uint32_t baddr = smem_u32(sm + 64 * 1024);
    const uint64_t a0 = umma_desc_sw128(smem_u32(sm));
    const uint64_t b0 = kBmn ? ((uint64_t)((baddr & 0x3FFFF) >> 4) | ((uint64_t)(16384 >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) |
                                ((uint64_t)1 << 46) | ((uint64_t)2 << 61))
                             : umma_desc_sw128(baddr);
    constexpr uint64_t kBStep = kBmn ? (2048 >> 4) : 2;
    auto body = [&](bool leader) {
        const uint32_t tm = tmem_base;
        for (int r = 0; r < kReps; ++r) {
#pragma unroll
            for (int u = 0; u < kUnits; ++u) {
                const uint32_t d = tm + (uint32_t)((u % kAcc) * N);
                const uint64_t ad = a0 + (uint64_t)((u & 3) * (16384 >> 4));
                if (leader) {
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        if (kF16) umma_f16(d, ad + 2 * ks, b0 + kBStep * ks, idesc, 1);
                        else umma_tf32(d, ad + 2 * ks, b0 + kBStep * ks, idesc, 1);
                    }
                    if (kCommitPerUnit) umma_commit(&bar2);          // stage-release style commit nobody waits on
                }
            }
        }
        if (leader) umma_commit(&bar);
    };
    long long t0 = 0;
    if (kUniform) {
        if (threadIdx.x < 32) {
            const bool leader = elect_one_lane();
            t0 = clock64();
            body(leader);
            mbar_wait(&bar, 0);
            if (blockIdx.x == 0 && leader) out[0] = clock64() - t0;
        }
    } else if (threadIdx.x == 0) {
        t0 = clock64();
        body(true);
        mbar_wait(&bar, 0);
        if (blockIdx.x == 0) out[0] = clock64() - t0;
    }
    tc_fence_before(); __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tmem_base, 512);
}

template <bool kF16, int M, int N, int kAcc, bool kCommitPerUnit, bool kUniform, bool kBmn = false>
static void run(const char* name, long long* out) {
    auto k = probe<kF16, M, N, kAcc, kCommitPerUnit, kUniform, kBmn>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaMemset(out, 0, 8);
    k<<<148, 128, 200 * 1024>>>(out);
    cudaError_t e = cudaDeviceSynchronize();
    long long h = 0;
    cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
    const int n = kReps * kUnits * 4;
    const double clk = (double)h / n;
    const double flop = 2.0 * M * N * (kF16 ? 16 : 8);
    printf("%-62s %7.1f clk per UMMA  %6.0f flop/clk/SM %s\n", name, clk, flop / clk, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
    long long* out; cudaMalloc(&out, 16);
    run<true, 128, 256, 2, false, true>("f16  M=128 N=256, uniform issue", out);
    run<true, 128, 128, 2, false, true>("f16  M=128 N=128, uniform issue", out);
    run<true, 128, 64, 2, false, true>("f16  M=128 N=64,  uniform issue", out);
    run<true, 128, 32, 2, false, true>("f16  M=128 N=32,  uniform issue", out);
    run<true, 64, 256, 2, false, true>("f16  M=64  N=256, uniform issue", out);
    run<true, 128, 128, 1, false, true>("f16  M=128 N=128, one accumulator chain", out);
    run<true, 128, 128, 2, true, true>("f16  M=128 N=128, commit after every 4 UMMAs", out);
    run<true, 128, 256, 2, true, true>("f16  M=128 N=256, commit after every 4 UMMAs", out);
    run<true, 128, 128, 2, false, true, true>("f16  M=128 N=128, B operand MN-major", out);
    run<true, 128, 256, 2, false, true, true>("f16  M=128 N=256, B operand MN-major", out);
    run<false, 128, 256, 2, false, true>("tf32 M=128 N=256, uniform issue", out);
    run<false, 128, 128, 2, false, true>("tf32 M=128 N=128, uniform issue", out);
    run<true, 128, 256, 2, false, false>("f16  M=128 N=256, issued inside `if (threadIdx.x == 0)`", out);
    run<true, 128, 128, 2, false, false>("f16  M=128 N=128, issued inside `if (threadIdx.x == 0)`", out);
    run<true, 128, 64, 2, false, false>("f16  M=128 N=64,  issued inside `if (threadIdx.x == 0)`", out);
    return 0;
}
