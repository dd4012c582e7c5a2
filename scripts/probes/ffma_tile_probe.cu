// Micro-benchmark: FP32-pipe throughput of a register-tiled GEMM inner loop fed from shared memory, for
// the tile shapes the FP32 tile kernel could use.  Per k-step a thread loads TM row values and TN column
// values with 128-bit shared-memory loads and issues TM*TN/2 packed FMAs (fma.rn.f32x2).  Operand bytes
// per FMA: 8x8 -> 1.0, 16x8 / 8x16 -> 0.75, 16x16 -> 0.5.  One CTA per SM; warps = 128*128 / (32*TM*TN).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o scripts/probes/ffma_tile_probe scripts/probes/ffma_tile_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int kK = 128, kRows = 128, kCols = 128;

template <int TM, int TN>
__global__ void __launch_bounds__(kRows* kCols / (TM * TN), 1) probe(float* out, int iters) {
    extern __shared__ __align__(16) float sm[];
    float* As = sm;                 // [k][128 rows]
    float* Bs = sm + kK * kRows;    // [k][128 cols]
    for (int i = threadIdx.x; i < kK * (kRows + kCols); i += blockDim.x) sm[i] = 1e-3f * (float)(i & 63);
    __syncthreads();
    constexpr int NX = kCols / TN;                      // threads along the columns
    const int tx = threadIdx.x % NX, ty = threadIdx.x / NX;
    float2 acc[TM][TN / 2];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN / 2; ++j) acc[i][j] = make_float2(0.f, 0.f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll 8
        for (int k = 0; k < kK; ++k) {
            float a[TM];
            float2 b[TN / 2];
#pragma unroll
            for (int i = 0; i < TM; i += 4) {
                const float4 v = *reinterpret_cast<const float4*>(As + k * kRows + ty * TM + i);
                a[i] = v.x; a[i + 1] = v.y; a[i + 2] = v.z; a[i + 3] = v.w;
            }
#pragma unroll
            for (int j = 0; j < TN; j += 4) {
                const float4 v = *reinterpret_cast<const float4*>(Bs + k * kCols + tx * TN + j);
                b[j / 2] = make_float2(v.x, v.y); b[j / 2 + 1] = make_float2(v.z, v.w);
            }
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN / 2; ++j) acc[i][j] = __ffma2_rn(make_float2(a[i], a[i]), b[j], acc[i][j]);
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN / 2; ++j) s += acc[i][j].x + acc[i][j].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// the same loop with scalar FFMA (one instruction per multiply-add)
template <int TM, int TN>
__global__ void __launch_bounds__(kRows* kCols / (TM * TN), 1) probe_scalar(float* out, int iters) {
    extern __shared__ __align__(16) float sm[];
    float* As = sm;
    float* Bs = sm + kK * kRows;
    for (int i = threadIdx.x; i < kK * (kRows + kCols); i += blockDim.x) sm[i] = 1e-3f * (float)(i & 63);
    __syncthreads();
    constexpr int NX = kCols / TN;
    const int tx = threadIdx.x % NX, ty = threadIdx.x / NX;
    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll 8
        for (int k = 0; k < kK; ++k) {
            float a[TM], b[TN];
#pragma unroll
            for (int i = 0; i < TM; i += 4) {
                const float4 v = *reinterpret_cast<const float4*>(As + k * kRows + ty * TM + i);
                a[i] = v.x; a[i + 1] = v.y; a[i + 2] = v.z; a[i + 3] = v.w;
            }
#pragma unroll
            for (int j = 0; j < TN; j += 4) {
                const float4 v = *reinterpret_cast<const float4*>(Bs + k * kCols + tx * TN + j);
                b[j] = v.x; b[j + 1] = v.y; b[j + 2] = v.z; b[j + 3] = v.w;
            }
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) s += acc[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int TM, int TN, bool kScalar = false>
static void run(const char* name, float* out) {
    const int threads = kRows * kCols / (TM * TN), iters = 200;
    const size_t smem = (size_t)kK * (kRows + kCols) * sizeof(float);
    auto kern = kScalar ? probe_scalar<TM, TN> : probe<TM, TN>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    kern<<<148, threads, smem>>>(out, 2);
    cudaEventRecord(e0);
    kern<<<148, threads, smem>>>(out, iters);
    cudaEventRecord(e1);
    cudaError_t e = cudaDeviceSynchronize();
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    const double flop = 2.0 * kRows * kCols * kK * iters * 148;
    cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, kern);
    printf("%-10s %4d threads, %3d regs: %6.1f TFLOP/s %s\n", name, threads, fa.numRegs, flop / (ms * 1e-3) / 1e12,
           e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
    float* out; cudaMalloc(&out, 148 * 1024 * sizeof(float));
    run<8, 8>("8x8", out);
    run<16, 8>("16x8", out);
    run<8, 16>("8x16", out);
    run<16, 16>("16x16", out);
    run<4, 8>("4x8", out);
    run<8, 8, true>("8x8 FFMA", out);
    run<16, 8, true>("16x8 FFMA", out);
    run<8, 16, true>("8x16 FFMA", out);
    return 0;
}
