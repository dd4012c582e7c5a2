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


// Warp-uniform column operand: a warp owns 16 columns and ALL 128 rows (lane = 4 consecutive rows), so the row operand
// is one conflict-free 128-bit load per lane (512 distinct bytes per warp) and the 16 column values are the same address
// for every lane (broadcast).  Same 64 FMAs per thread and k-step as the 8x8 tile.
template <bool kUniformViaShuffle>
__global__ void __launch_bounds__(256, 1) probe_uniform_b(float* out, int iters) {
    extern __shared__ __align__(16) float sm[];
    float* As = sm;
    float* Bs = sm + kK * kRows;
    for (int i = threadIdx.x; i < kK * (kRows + kCols); i += blockDim.x) sm[i] = 1e-3f * (float)(i & 63);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float2 acc[4][8];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = make_float2(0.f, 0.f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll 8
        for (int k = 0; k < kK; ++k) {
            const float4 a = *reinterpret_cast<const float4*>(As + k * kRows + lane * 4);
            float2 b[8];
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
                const float4 v = *reinterpret_cast<const float4*>(Bs + k * kCols + warp * 16 + j);
                b[j / 2] = make_float2(v.x, v.y); b[j / 2 + 1] = make_float2(v.z, v.w);
            }
            const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = __ffma2_rn(make_float2(av[i], av[i]), b[j], acc[i][j]);
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) s += acc[i][j].x + acc[i][j].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// 8 rows x 8 columns per thread, lane = (4 row groups) x (8 column groups) as in the tile kernel, but the column
// operand is read with 64-bit loads (a half-warp phase then covers 16 lanes = two row groups x 8 column groups).
__global__ void __launch_bounds__(256, 1) probe_8x8_lds64(float* out, int iters) {
    extern __shared__ __align__(16) float sm[];
    float* As = sm;
    float* Bs = sm + kK * kRows;
    for (int i = threadIdx.x; i < kK * (kRows + kCols); i += blockDim.x) sm[i] = 1e-3f * (float)(i & 63);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ty = (warp >> 1) * 4 + (lane >> 3), tx = (warp & 1) * 8 + (lane & 7);
    float2 acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = make_float2(0.f, 0.f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll 8
        for (int k = 0; k < kK; ++k) {
            float a[8];
            float2 b[4];
#pragma unroll
            for (int i = 0; i < 8; i += 4) {
                const float4 v = *reinterpret_cast<const float4*>(As + k * kRows + ty * 8 + i);
                a[i] = v.x; a[i + 1] = v.y; a[i + 2] = v.z; a[i + 3] = v.w;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = *reinterpret_cast<const float2*>(Bs + k * kCols + j * 32 + tx * 2);
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = __ffma2_rn(make_float2(a[i], a[i]), b[j], acc[i][j]);
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) s += acc[i][j].x + acc[i][j].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}


// 2 rows x 32 columns per thread: a warp owns 64 rows (lane = 2 consecutive rows, one 64-bit load) and 32 columns
// (eight warp-uniform 128-bit loads): 4 warps per 64-row group, so two independent row groups per tile remain possible.
__global__ void __launch_bounds__(256, 1) probe_2x32_uniform(float* out, int iters) {
    extern __shared__ __align__(16) float sm[];
    float* As = sm;
    float* Bs = sm + kK * kRows;
    for (int i = threadIdx.x; i < kK * (kRows + kCols); i += blockDim.x) sm[i] = 1e-3f * (float)(i & 63);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int row0 = (warp >> 2) * 64 + lane * 2, col0 = (warp & 3) * 32;
    float2 acc[2][16];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[i][j] = make_float2(0.f, 0.f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll 8
        for (int k = 0; k < kK; ++k) {
            const float2 a = *reinterpret_cast<const float2*>(As + k * kRows + row0);
            float2 b[16];
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
                const float4 v = *reinterpret_cast<const float4*>(Bs + k * kCols + col0 + j);
                b[j / 2] = make_float2(v.x, v.y); b[j / 2 + 1] = make_float2(v.z, v.w);
            }
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                acc[0][j] = __ffma2_rn(make_float2(a.x, a.x), b[j], acc[0][j]);
                acc[1][j] = __ffma2_rn(make_float2(a.y, a.y), b[j], acc[1][j]);
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 16; ++j) s += acc[i][j].x + acc[i][j].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// The tile kernel's own mapping: lane = (4 row groups of 8 rows) x (8 column groups), rows by two broadcast 128-bit loads,
// columns 4tx..4tx+3 and 64+4tx.. by two conflict-free 128-bit loads.
__global__ void __launch_bounds__(256, 1) probe_8x8_kernel_mapping(float* out, int iters) {
    extern __shared__ __align__(16) float sm[];
    float* As = sm;
    float* Bs = sm + kK * kRows;
    for (int i = threadIdx.x; i < kK * (kRows + kCols); i += blockDim.x) sm[i] = 1e-3f * (float)(i & 63);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ty = (warp >> 1) * 4 + (lane >> 3), tx = (warp & 1) * 8 + (lane & 7);
    float2 acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = make_float2(0.f, 0.f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll 8
        for (int k = 0; k < kK; ++k) {
            float a[8];
            float2 b[4];
#pragma unroll
            for (int i = 0; i < 8; i += 4) {
                const float4 v = *reinterpret_cast<const float4*>(As + k * kRows + ty * 8 + i);
                a[i] = v.x; a[i + 1] = v.y; a[i + 2] = v.z; a[i + 3] = v.w;
            }
            const float4 b0 = *reinterpret_cast<const float4*>(Bs + k * kCols + tx * 4);
            const float4 b1 = *reinterpret_cast<const float4*>(Bs + k * kCols + 64 + tx * 4);
            b[0] = make_float2(b0.x, b0.y); b[1] = make_float2(b0.z, b0.w);
            b[2] = make_float2(b1.x, b1.y); b[3] = make_float2(b1.z, b1.w);
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = __ffma2_rn(make_float2(a[i], a[i]), b[j], acc[i][j]);
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) s += acc[i][j].x + acc[i][j].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// Both halves of a layer in ONE pass: 128 rows x 256 columns (Z and Y), 256 threads with 8 rows x (8 + 8) columns each in the
// tile kernel's mapping.  Per k-step a thread loads 8 row values once for 128 multiply-adds (two broadcast 128-bit loads)
// and 16 column values (four conflict-free 128-bit loads): 0.75 operand bytes per multiply-add instead of 1.0.
__global__ void __launch_bounds__(256, 1) probe_8x16_two_halves(float* out, int iters) {
    extern __shared__ __align__(16) float sm[];
    float* As = sm;                       // [k][128 rows]
    float* Bs = sm + kK * kRows;          // [k][256 cols]: Z half then Y half
    for (int i = threadIdx.x; i < kK * (kRows + 2 * kCols); i += blockDim.x) sm[i] = 1e-3f * (float)(i & 63);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ty = (warp >> 1) * 4 + (lane >> 3), tx = (warp & 1) * 8 + (lane & 7);
    float2 acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = make_float2(0.f, 0.f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll 4
        for (int k = 0; k < kK; ++k) {
            float a[8];
            float2 b[8];
#pragma unroll
            for (int i = 0; i < 8; i += 4) {
                const float4 v = *reinterpret_cast<const float4*>(As + k * kRows + ty * 8 + i);
                a[i] = v.x; a[i + 1] = v.y; a[i + 2] = v.z; a[i + 3] = v.w;
            }
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const float4 b0 = *reinterpret_cast<const float4*>(Bs + k * 2 * kCols + h * kCols + tx * 4);
                const float4 b1 = *reinterpret_cast<const float4*>(Bs + k * 2 * kCols + h * kCols + 64 + tx * 4);
                b[4 * h] = make_float2(b0.x, b0.y); b[4 * h + 1] = make_float2(b0.z, b0.w);
                b[4 * h + 2] = make_float2(b1.x, b1.y); b[4 * h + 3] = make_float2(b1.z, b1.w);
            }
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = __ffma2_rn(make_float2(a[i], a[i]), b[j], acc[i][j]);
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) s += acc[i][j].x + acc[i][j].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename K>
static void run_kernel(const char* name, K kern, int threads, float* out) {
    const int iters = 200;
    const size_t smem = (size_t)kK * (kRows + kCols) * sizeof(float);
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
    printf("%-22s %4d threads, %3d regs: %6.1f TFLOP/s %s\n", name, threads, fa.numRegs, flop / (ms * 1e-3) / 1e12,
           e == cudaSuccess ? "" : cudaGetErrorString(e));
}

// 128 x 256 outputs per k (both halves of a layer)
template <typename K>
static void run_kernel2(const char* name, K kern, int threads, float* out) {
    const int iters = 200;
    const size_t smem = (size_t)kK * (kRows + 2 * kCols) * sizeof(float);
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    kern<<<148, threads, smem>>>(out, 2);
    cudaEventRecord(e0);
    kern<<<148, threads, smem>>>(out, iters);
    cudaEventRecord(e1);
    cudaError_t e = cudaDeviceSynchronize();
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    const double flop = 2.0 * kRows * 2 * kCols * kK * iters * 148;
    cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, kern);
    printf("%-22s %4d threads, %3d regs: %6.1f TFLOP/s %s\n", name, threads, fa.numRegs, flop / (ms * 1e-3) / 1e12,
           e == cudaSuccess ? "" : cudaGetErrorString(e));
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
    run_kernel("4x16 uniform cols", probe_uniform_b<false>, 256, out);
    run_kernel("8x8 cols by LDS.64", probe_8x8_lds64, 256, out);
    run_kernel("8x8 kernel mapping", probe_8x8_kernel_mapping, 256, out);
    run_kernel("2x32 uniform cols", probe_2x32_uniform, 256, out);
    run_kernel2("8x16 both halves", probe_8x16_two_halves, 256, out);
    run<8, 8, true>("8x8 FFMA", out);
    run<16, 8, true>("16x8 FFMA", out);
    run<8, 16, true>("8x16 FFMA", out);
    return 0;
}
