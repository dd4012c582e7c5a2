// Spectral field solve for power-of-two grids by FFT (sm_100a).
//
//   E = Re ifft( i * fft(n - 1) / k ),  k = 2*pi*fftfreq(nx, L/nx),  E_hat(0) = 0
//   (src/baseline_solver.py:26,59-68; the Nyquist bin vanishes under Re()).
//
// nx <= 2^14: one CTA per IC, the whole transform lives in shared memory:
//     load rho -> DIF forward FFT (bit-reversed spectrum) -> multiply -> DIT inverse -> store Re.
// nx  > 2^14: four-step factorisation nx = N1 * N2, N2 = 2^14, element n = n1*N2 + n2:
//     A  columns: for a tile of T consecutive n2, length-N1 FFT over n1, times W_nx^(n2*k1)  -> Y[k1][n2]
//     B  rows   : for each k1, length-N2 FFT over n2 -> k2, multiply by i/k(k1 + N1*k2),
//                 inverse FFT k2 -> n2, times conj twiddle                                      (in place)
//     C  columns: inverse length-N1 FFT over k1 -> n1, real part / nx -> E[n1*N2 + n2]
// No transposes: forward transforms are decimation-in-frequency (natural in, bit-reversed
// out), inverse ones decimation-in-time (bit-reversed in, natural out), and the spectral
// multiply is index-agnostic.  Shared-memory radix-2 stages with sincospi twiddles.
#include "common.cuh"
#include "field_kernels.cuh"

namespace fluxgnn {

namespace {

constexpr int kFftThreads = 512;

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

__device__ __forceinline__ int bitrev(int v, int bits) { return (int)(__brev((unsigned)v) >> (32 - bits)); }

// exp(sign * 2*pi*i * q / n), n a power of two <= 2^14: the argument of sincospif is exact.
__device__ __forceinline__ float2 twiddle(int q, int n, float sign) {
    float s, c;
    sincospif(sign * 2.0f * (float)q / (float)n, &s, &c);
    return make_float2(c, s);
}

// `cnt` interleaved transforms of length n = 2^bits: element idx of transform t at s[idx*cnt + t].
// Forward (sign -1): DIF, natural order in -> bit-reversed order out.
__device__ void fft_dif(float2* s, int bits, int cnt, float sign) {
    const int n = 1 << bits;
    const int work = (n >> 1) * cnt;
    for (int h = n >> 1; h >= 1; h >>= 1) {
        for (int w = threadIdx.x; w < work; w += blockDim.x) {
            const int t = w % cnt, b = w / cnt;
            const int j = b & (h - 1);
            const int i0 = ((b - j) << 1) + j, i1 = i0 + h;
            const float2 u = s[i0 * cnt + t], v = s[i1 * cnt + t];
            s[i0 * cnt + t] = make_float2(u.x + v.x, u.y + v.y);
            s[i1 * cnt + t] = cmul(make_float2(u.x - v.x, u.y - v.y), twiddle(j * (n / (2 * h)), n, sign));
        }
        __syncthreads();
    }
}

// Inverse of the above ordering: DIT, bit-reversed order in -> natural order out.
__device__ void fft_dit(float2* s, int bits, int cnt, float sign) {
    const int n = 1 << bits;
    const int work = (n >> 1) * cnt;
    for (int h = 1; h < n; h <<= 1) {
        for (int w = threadIdx.x; w < work; w += blockDim.x) {
            const int t = w % cnt, b = w / cnt;
            const int j = b & (h - 1);
            const int i0 = ((b - j) << 1) + j, i1 = i0 + h;
            const float2 u = s[i0 * cnt + t];
            const float2 v = cmul(s[i1 * cnt + t], twiddle(j * (n / (2 * h)), n, sign));
            s[i0 * cnt + t] = make_float2(u.x + v.x, u.y + v.y);
            s[i1 * cnt + t] = make_float2(u.x - v.x, u.y - v.y);
        }
        __syncthreads();
    }
}

// W_nx^(sign * r), 0 <= r < nx: r/nx is exact in float up to nx = 2^24, beyond that use double
__device__ __forceinline__ float2 big_twiddle(long long r, long long nx, float sign) {
    if (nx <= (1LL << 24)) {
        float s, c;
        sincospif(sign * 2.0f * (float)r / (float)nx, &s, &c);
        return make_float2(c, s);
    }
    double s, c;
    sincospi((double)sign * 2.0 * (double)r / (double)nx, &s, &c);
    return make_float2((float)c, (float)s);
}

// spectrum bin kbin (0..nx-1) of rho -> bin of E, including the 1/nx of the inverse transform
__device__ __forceinline__ float2 spectral_multiply(float2 v, long long kbin, long long nx, double length) {
    if (kbin == 0 || 2 * kbin == nx) return make_float2(0.f, 0.f);
    const long long m = (2 * kbin < nx) ? kbin : kbin - nx;
    const float f = (float)(length / (6.283185307179586476925 * (double)m * (double)nx));   // 1/(k * nx)
    return make_float2(-v.y * f, v.x * f);                                                  // i * v / k
}

}  // namespace

// ---------------------------------------------------------------------------
// whole transform in one CTA (nx = 2^bits <= 2^14)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kFftThreads) poisson_fft_small_kernel(const float* __restrict__ n, long long n_stride,
                                                                        float* __restrict__ E, long long e_stride,
                                                                        int bits, double length) {
    extern __shared__ float2 sfft[];
    const int nx = 1 << bits;
    const float* src = n + (size_t)blockIdx.x * n_stride;
    for (int i = threadIdx.x; i < nx; i += blockDim.x) sfft[i] = make_float2(__fsub_rn(src[i], 1.0f), 0.f);
    __syncthreads();
    fft_dif(sfft, bits, 1, -1.f);
    for (int i = threadIdx.x; i < nx; i += blockDim.x) sfft[i] = spectral_multiply(sfft[i], bitrev(i, bits), nx, length);
    __syncthreads();
    fft_dit(sfft, bits, 1, +1.f);
    float* dst = E + (size_t)blockIdx.x * e_stride;
    for (int i = threadIdx.x; i < nx; i += blockDim.x) dst[i] = sfft[i].x;
}

// ---------------------------------------------------------------------------
// four-step, pass A: forward column transforms.  grid = (N2 / T, B)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kFftThreads) poisson_fft_cols_fwd_kernel(const float* __restrict__ n, long long n_stride,
                                                                           float2* __restrict__ Y, int bits1, int bits2,
                                                                           int T) {
    extern __shared__ float2 sfft[];
    const int N1 = 1 << bits1;
    const long long N2 = 1LL << bits2, nx = (long long)N1 << bits2;
    const long long n2_0 = (long long)blockIdx.x * T;
    const float* src = n + (size_t)blockIdx.y * n_stride;
    for (int w = threadIdx.x; w < N1 * T; w += blockDim.x) {
        const int t = w % T, n1 = w / T;
        sfft[w] = make_float2(__fsub_rn(src[(size_t)n1 * N2 + n2_0 + t], 1.0f), 0.f);
    }
    __syncthreads();
    fft_dif(sfft, bits1, T, -1.f);
    float2* dst = Y + (size_t)blockIdx.y * nx;
    for (int w = threadIdx.x; w < N1 * T; w += blockDim.x) {
        const int t = w % T, pos = w / T;
        const int k1 = bitrev(pos, bits1);
        const long long n2 = n2_0 + t;
        dst[(size_t)k1 * N2 + n2] = cmul(sfft[w], big_twiddle((n2 * k1) % nx, nx, -1.f));      // W_nx^(n2*k1)
    }
}

// pass B: rows.  grid = (N1, B); one length-N2 row per CTA, in place.
__global__ void __launch_bounds__(kFftThreads) poisson_fft_rows_kernel(float2* __restrict__ Y, int bits1, int bits2,
                                                                       double length) {
    extern __shared__ float2 sfft[];
    const int N1 = 1 << bits1, N2 = 1 << bits2;
    const long long nx = (long long)N1 << bits2;
    const int k1 = blockIdx.x;
    float2* row = Y + (size_t)blockIdx.y * nx + (size_t)k1 * N2;
    for (int i = threadIdx.x; i < N2; i += blockDim.x) sfft[i] = row[i];
    __syncthreads();
    fft_dif(sfft, bits2, 1, -1.f);
    for (int i = threadIdx.x; i < N2; i += blockDim.x) {
        const long long kbin = (long long)k1 + (long long)N1 * bitrev(i, bits2);
        sfft[i] = spectral_multiply(sfft[i], kbin, nx, length);
    }
    __syncthreads();
    fft_dit(sfft, bits2, 1, +1.f);
    for (int i = threadIdx.x; i < N2; i += blockDim.x) {
        row[i] = cmul(sfft[i], big_twiddle(((long long)i * k1) % nx, nx, +1.f));              // conj twiddle
    }
}

// pass C: inverse column transforms, real part out.  grid = (N2 / T, B)
__global__ void __launch_bounds__(kFftThreads) poisson_fft_cols_inv_kernel(const float2* __restrict__ Y,
                                                                           float* __restrict__ E, long long e_stride,
                                                                           int bits1, int bits2, int T) {
    extern __shared__ float2 sfft[];
    const int N1 = 1 << bits1;
    const long long N2 = 1LL << bits2, nx = (long long)N1 << bits2;
    const long long n2_0 = (long long)blockIdx.x * T;
    const float2* src = Y + (size_t)blockIdx.y * nx;
    for (int w = threadIdx.x; w < N1 * T; w += blockDim.x) {
        const int t = w % T, k1 = w / T;
        sfft[w] = src[(size_t)k1 * N2 + n2_0 + t];
    }
    __syncthreads();
    fft_dif(sfft, bits1, T, +1.f);                 // natural k1 in -> bit-reversed n1 out
    float* dst = E + (size_t)blockIdx.y * e_stride;
    for (int w = threadIdx.x; w < N1 * T; w += blockDim.x) {
        const int t = w % T, pos = w / T;
        dst[(size_t)bitrev(pos, bits1) * N2 + n2_0 + t] = sfft[w].x;
    }
}

// ---------------------------------------------------------------------------
// host
// ---------------------------------------------------------------------------
static int ilog2_exact(long long v) {
    int b = 0;
    while ((1LL << b) < v) ++b;
    return ((1LL << b) == v) ? b : -1;
}

bool poisson_fft_supported(int nx) {
    const int bits = ilog2_exact(nx);
    return bits >= kFftMinBits && bits <= kFftMaxBits;
}

size_t poisson_fft_workspace_bytes(int B, int nx) {
    const int bits = ilog2_exact(nx);
    if (bits <= kFftRowBits) return 0;
    return (size_t)B * (size_t)nx * sizeof(float2);
}

int launch_poisson_fft(const float* n, long long n_stride, float* E, long long e_stride, int B, int nx,
                       double length, void* workspace, cudaStream_t stream) {
    const int bits = ilog2_exact(nx);
    if (bits < kFftMinBits || bits > kFftMaxBits)
        return set_error(FLUXGNN_EUNSUP, "FFT field solve needs nx = 2^%d..2^%d, got %d", kFftMinBits, kFftMaxBits, nx);
    if (bits <= kFftRowBits) {
        const size_t smem = (size_t)nx * sizeof(float2);
        FLUXGNN_CUDA_OK(cudaFuncSetAttribute(poisson_fft_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)smem));
        const int threads = nx / 2 < kFftThreads ? (nx / 2 < 32 ? 32 : nx / 2) : kFftThreads;
        poisson_fft_small_kernel<<<B, threads, smem, stream>>>(n, n_stride, E, e_stride, bits, length);
        FLUXGNN_CUDA_OK(cudaGetLastError());
        count_launch();
        return FLUXGNN_OK;
    }
    if (workspace == nullptr)
        return set_error(FLUXGNN_EINVAL, "field solve for nx=%d needs fluxgnn_poisson_workspace_bytes() of scratch", nx);
    if (B > 65535) return set_error(FLUXGNN_EUNSUP, "field solve for nx=%d handles at most 65535 ICs per call", nx);
    const int bits2 = kFftRowBits, bits1 = bits - bits2;
    const int N1 = 1 << bits1, N2 = 1 << bits2;
    const int T = (1 << kFftRowBits) / N1;                   // N1 * T = 2^14 complex = 128 KiB
    const size_t smem = (size_t)(1 << kFftRowBits) * sizeof(float2);
    float2* Y = (float2*)workspace;
    FLUXGNN_CUDA_OK(cudaFuncSetAttribute(poisson_fft_cols_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    FLUXGNN_CUDA_OK(cudaFuncSetAttribute(poisson_fft_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    FLUXGNN_CUDA_OK(cudaFuncSetAttribute(poisson_fft_cols_inv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 gcol((unsigned)(N2 / T), (unsigned)B), grow((unsigned)N1, (unsigned)B);
    poisson_fft_cols_fwd_kernel<<<gcol, kFftThreads, smem, stream>>>(n, n_stride, Y, bits1, bits2, T);
    FLUXGNN_CUDA_OK(cudaGetLastError());
    poisson_fft_rows_kernel<<<grow, kFftThreads, smem, stream>>>(Y, bits1, bits2, length);
    FLUXGNN_CUDA_OK(cudaGetLastError());
    poisson_fft_cols_inv_kernel<<<gcol, kFftThreads, smem, stream>>>(Y, E, e_stride, bits1, bits2, T);
    FLUXGNN_CUDA_OK(cudaGetLastError());
    count_launch(3);
    return FLUXGNN_OK;
}

}  // namespace fluxgnn
