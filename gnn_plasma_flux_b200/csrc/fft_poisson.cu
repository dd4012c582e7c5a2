// Spectral field solve for power-of-two grids by FFT (sm_100a).
//
//   E = Re ifft( i * fft(n - 1) / k ),  k = 2*pi*fftfreq(nx, L/nx),  E_hat(0) = 0
//   (src/baseline_solver.py:26,59-68; the Nyquist bin vanishes under Re()).
//
// Real-input formulation: rho (nx reals) is read as M = nx/2 complex numbers
// z_j = rho_2j + i rho_2j+1.  With Z = FFT_M(z) and k' = (M - k) mod M,
//   rho_hat[k]   = A + W_nx^k B,   A = (Z[k] + conj Z[k'])/2,  B = -i (Z[k] - conj Z[k'])/2,
//   rho_hat[k+M] = conj(rho_hat[k'])                      (rho is real),
// the multiplier i/k is applied to both, and the result is packed back the same way
//   Zt[k] = (E_hat[k] + E_hat[k+M])/2 + i W_nx^-k (E_hat[k] - E_hat[k+M])/2,
// so that IFFT_M(Zt) = E_2j + i E_2j+1: half the transform length, half the memory traffic.
//
// nx <= 2^15: one CTA per IC, everything in shared memory (DIF forward -> pair-wise spectral
//             step -> DIT inverse; the bit-reversed order in between never has to be undone).
// nx  > 2^15: four-step factorisation M = N1 * N2, N2 = 2^12, element j = j1*N2 + j2:
//     A  columns: tiles of T consecutive j2, length-N1 FFT over j1, times W_M^(j2*k1)   -> Y[k1][j2]
//     B  rows   : bins k and M-k live in rows k1 and N1-k1, so one CTA takes BOTH rows
//                 (interleaved in shared memory): forward FFTs, pair-wise spectral step,
//                 inverse FFTs, conj twiddle, in place.  Rows 0 and N1/2 pair with themselves.
//     C  columns: inverse length-N1 FFT over k1 -> (E_2j, E_2j+1)
// No transposes.  Each shared-memory round trip is a radix-16 transform in registers.
#include "common.cuh"
#include "field_kernels.cuh"

namespace fluxgnn {

namespace {

constexpr int kFftThreads = 512;       // whole-transform kernel (upper bound; it launches nx/16)
constexpr int kFftStepThreads = 256;   // four-step kernels: 3 CTAs per SM by registers and shared memory

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

// Shared-memory address of element idx of interleaved transform t.  A lone transform is
// padded by one element per 16 so that the strided register-radix accesses of the short
// block lengths spread over all banks; interleaved transforms are conflict-free as they are.
__device__ __forceinline__ int saddr(int idx, int t, int cnt) {
    return cnt == 1 ? idx + (idx >> 4) : idx * cnt + t;
}

// v * exp(sign * 2*pi*i * Q / 16) for a compile-time sixteenth-turn Q: trivial roots cost nothing,
// eighth turns two multiplies, the rest one complex multiply.
template <int Q>
__device__ __forceinline__ float2 mul_root16(float2 v, float sign) {
    constexpr int q = Q & 15;
    if (q == 0) return v;
    if (q == 8) return make_float2(-v.x, -v.y);
    if (q == 4) return make_float2(-sign * v.y, sign * v.x);                 // * (sign i)
    if (q == 12) return make_float2(sign * v.y, -sign * v.x);                // * (-sign i)
    constexpr float kC[16] = {1.f, 0.92387953251128674f, 0.70710678118654752f, 0.38268343236508977f,
                              0.f, -0.38268343236508977f, -0.70710678118654752f, -0.92387953251128674f,
                              -1.f, -0.92387953251128674f, -0.70710678118654752f, -0.38268343236508977f,
                              0.f, 0.38268343236508977f, 0.70710678118654752f, 0.92387953251128674f};
    constexpr float kS[16] = {0.f, 0.38268343236508977f, 0.70710678118654752f, 0.92387953251128674f,
                              1.f, 0.92387953251128674f, 0.70710678118654752f, 0.38268343236508977f,
                              0.f, -0.38268343236508977f, -0.70710678118654752f, -0.92387953251128674f,
                              -1.f, -0.92387953251128674f, -0.70710678118654752f, -0.38268343236508977f};
    const float c = kC[q], sn = sign * kS[q];
    return make_float2(v.x * c - v.y * sn, v.x * sn + v.y * c);
}

__host__ __device__ constexpr int brev_bits(int v, int bits) {
    int r = 0;
    for (int b = 0; b < bits; ++b) r |= ((v >> b) & 1) << (bits - 1 - b);
    return r;
}
__host__ __device__ constexpr int ilog2c(int v) { return v <= 1 ? 0 : 1 + ilog2c(v >> 1); }

// R-point DFT in registers, decimation in frequency: natural order in, bit-reversed order out,
// constant roots only (stage of register-span `half`: W_(2 half)^mm on the lower output).
template <int R, int HALF = R / 2>
struct DifStages {
    __device__ static __forceinline__ void run(float2 (&v)[R], float sign) {
        constexpr int kStep = 8 / HALF;                    // sixteenth-turns per unit of mm
        unroll<0>(v, sign, kStep);
        DifStages<R, HALF / 2>::run(v, sign);
    }
    template <int A>
    __device__ static __forceinline__ void unroll(float2 (&v)[R], float sign, int) {
        if constexpr (A < R) {
            if constexpr ((A & HALF) == 0) {
                constexpr int mm = A & (HALF - 1);
                const float2 u = v[A], x = v[A + HALF];
                v[A] = make_float2(u.x + x.x, u.y + x.y);
                v[A + HALF] = mul_root16<mm * (8 / HALF)>(make_float2(u.x - x.x, u.y - x.y), sign);
            }
            unroll<A + 1>(v, sign, 0);
        }
    }
};
template <int R>
struct DifStages<R, 0> {
    __device__ static __forceinline__ void run(float2 (&)[R], float) {}
};

// The matching decimation-in-time network: bit-reversed order in, natural order out.
template <int R, int HALF = 1>
struct DitStages {
    __device__ static __forceinline__ void run(float2 (&v)[R], float sign) {
        unroll<0>(v, sign);
        DitStages<R, HALF * 2>::run(v, sign);
    }
    template <int A>
    __device__ static __forceinline__ void unroll(float2 (&v)[R], float sign) {
        if constexpr (A < R) {
            if constexpr ((A & HALF) == 0) {
                constexpr int mm = A & (HALF - 1);
                const float2 u = v[A], x = mul_root16<mm * (8 / HALF)>(v[A + HALF], sign);
                v[A] = make_float2(u.x + x.x, u.y + x.y);
                v[A + HALF] = make_float2(u.x - x.x, u.y - x.y);
            }
            unroll<A + 1>(v, sign);
        }
    }
};
template <int R>
struct DitStages<R, R> {
    __device__ static __forceinline__ void run(float2 (&)[R], float) {}
};

// v * w^Q with w, w^2, w^4, w^8 given: the power is assembled from its binary digits, so only four
// powers live in registers and no product is more than four multiplications away from the sincos.
template <int Q>
__device__ __forceinline__ float2 mul_power(float2 v, float2 p1, float2 p2, float2 p4, float2 p8) {
    if constexpr (Q & 1) v = cmul(v, p1);
    if constexpr (Q & 2) v = cmul(v, p2);
    if constexpr (Q & 4) v = cmul(v, p4);
    if constexpr (Q & 8) v = cmul(v, p8);
    return v;
}

template <int R, int A = 1>
__device__ __forceinline__ void twiddle_outputs(float2 (&v)[R], float2 p1, float2 p2, float2 p4, float2 p8) {
    if constexpr (A < R) {
        v[A] = mul_power<brev_bits(A, ilog2c(R))>(v[A], p1, p2, p4, p8);
        twiddle_outputs<R, A + 1>(v, p1, p2, p4, p8);
    }
}

// One radix-R pass over blocks of length L (R | L): every work item owns the R elements
// blk*L + base + m*(L/R).
//   DIF (kDit = false): y = DFT_R(v), then y_q *= W_L^(base*q); output q sits in register bitrev(q),
//                       i.e. sub-block bitrev(q) of the block -- the layout the next passes expect.
//   DIT (kDit = true) : the transpose: register a *= W_L^(base*bitrev(a)), then the DIT network.
template <int R, bool kDit>
__device__ __forceinline__ void radix_pass(float2* s, int n, int L, int cnt, float sign) {
    const int sub = L / R;                         // stride between a work item's elements (power of two)
    const int items = (n / R) * cnt;
    const int cnt_bits = 31 - __clz(cnt), sub_bits = 31 - __clz(sub);
    for (int w = threadIdx.x; w < items; w += blockDim.x) {
        const int t = w & (cnt - 1), q = w >> cnt_bits;
        const int base = q & (sub - 1), blk = q >> sub_bits;
        const int e0 = blk * L + base;
        float2 v[R];
#pragma unroll
        for (int m = 0; m < R; ++m) v[m] = s[saddr(e0 + m * sub, t, cnt)];
        const float2 p1 = twiddle(base, L, sign);                     // W_L^base and its squarings
        const float2 p2 = cmul(p1, p1), p4 = cmul(p2, p2), p8 = cmul(p4, p4);
        if (!kDit) {
            DifStages<R>::run(v, sign);
            if (sub > 1) twiddle_outputs<R>(v, p1, p2, p4, p8);     // sub == 1: base == 0, all twiddles are 1
        } else {
            if (sub > 1) twiddle_outputs<R>(v, p1, p2, p4, p8);
            DitStages<R>::run(v, sign);
        }
#pragma unroll
        for (int m = 0; m < R; ++m) s[saddr(e0 + m * sub, t, cnt)] = v[m];
    }
    __syncthreads();
}

template <bool kDit>
__device__ __forceinline__ void radix_dispatch(float2* s, int n, int L, int R, int cnt, float sign) {
    switch (R) {
        case 16: radix_pass<16, kDit>(s, n, L, cnt, sign); break;
        case 8: radix_pass<8, kDit>(s, n, L, cnt, sign); break;
        case 4: radix_pass<4, kDit>(s, n, L, cnt, sign); break;
        default: radix_pass<2, kDit>(s, n, L, cnt, sign); break;
    }
}

// `cnt` interleaved transforms of length n = 2^bits (addressing: saddr()).
// Forward ordering: DIF, natural order in -> bit-reversed order out.  Radix-16 passes over
// block lengths n, n/16, ...; the last pass takes the remaining 2, 4 or 8.
__device__ void fft_dif(float2* s, int bits, int cnt, float sign) {
    const int n = 1 << bits;
    int L = n;
    while (L > 1) {
        const int R = L >= 16 ? 16 : L;
        radix_dispatch<false>(s, n, L, R, cnt, sign);
        L /= R;
    }
}

// Inverse ordering: DIT, bit-reversed order in -> natural order out: the same passes, backwards.
__device__ void fft_dit(float2* s, int bits, int cnt, float sign) {
    const int n = 1 << bits;
    int L = 1 << (bits & 3);                       // block length of the short pass (1 = none)
    if (L > 1) radix_dispatch<true>(s, n, L, L, cnt, sign);
    while (L < n) {
        L *= 16;
        radix_dispatch<true>(s, n, L, 16, cnt, sign);
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

// The pair-wise spectral step for bins k and kp = (M - k) mod M of the half-length transform
// (see the header): in  zk = Z[k], zp = Z[kp];  out  Zt[k], Zt[kp].  scale = L / (2 pi M) carries the
// 1/M of the inverse transform; wk = W_nx^k.
__device__ __forceinline__ void spectral_pair(float2 zk, float2 zp, long long k, long long M, float2 wk, float scale,
                                              float2& ok, float2& op) {
    if (k == 0) {                      // rho_hat[0] and the Nyquist bin: both multipliers are zero
        ok = make_float2(0.f, 0.f);
        op = ok;
        return;
    }
    const float2 a = make_float2(0.5f * (zk.x + zp.x), 0.5f * (zk.y - zp.y));       // A[k];  A[kp] = conj
    const float2 b = make_float2(0.5f * (zk.y + zp.y), -0.5f * (zk.x - zp.x));      // B[k];  B[kp] = conj
    const float2 wb = cmul(wk, b);
    const float2 rk = make_float2(a.x + wb.x, a.y + wb.y);                          // rho_hat[k]
    // W_nx^kp = -conj(W_nx^k):  rho_hat[kp] = conj(A) - conj(wk) conj(B) = conj(A - wk B)
    const float2 rp = make_float2(a.x - wb.x, -(a.y - wb.y));                       // rho_hat[kp]
    const long long kp = M - k;
    // bin numbers are < 2^24, exact in fp32; one rounding in the quotient
    const float fk = __fdiv_rn(scale, (float)k), fkM = -__fdiv_rn(scale, (float)kp);     // k - M = -kp
    const float fp = __fdiv_rn(scale, (float)kp), fpM = -__fdiv_rn(scale, (float)k);     // kp - M = -k
    // E_hat[q] = i f(q) rho_hat[q];  rho_hat[k+M] = conj(rho_hat[kp]),  rho_hat[kp+M] = conj(rho_hat[k])
    const float2 ek = make_float2(-fk * rk.y, fk * rk.x), ekM = make_float2(fkM * rp.y, fkM * rp.x);
    const float2 ep = make_float2(-fp * rp.y, fp * rp.x), epM = make_float2(fpM * rk.y, fpM * rk.x);
    // Zt[k] = (ek + ekM)/2 + i conj(wk) (ek - ekM)/2 ;   Zt[kp] likewise with W_nx^-kp = -wk
    const float2 sk = make_float2(0.5f * (ek.x + ekM.x), 0.5f * (ek.y + ekM.y));
    const float2 tk = cmul(make_float2(wk.x, -wk.y), make_float2(0.5f * (ek.x - ekM.x), 0.5f * (ek.y - ekM.y)));
    ok = make_float2(sk.x - tk.y, sk.y + tk.x);
    const float2 sp = make_float2(0.5f * (ep.x + epM.x), 0.5f * (ep.y + epM.y));
    const float2 tp = cmul(make_float2(-wk.x, -wk.y), make_float2(0.5f * (ep.x - epM.x), 0.5f * (ep.y - epM.y)));
    op = make_float2(sp.x - tp.y, sp.y + tp.x);
}

// Spectral step inside ONE row of length n2 = 2^bits2 that pairs with itself (row k1 = 0 or k1 = N1/2):
// position p holds k2 = bitrev(p); the partner bin sits at position pp.
__device__ __forceinline__ void spectral_self_row(float2* s, int bits2, int k1, int N1, long long M, long long nx,
                                                  float scale) {
    const int n2 = 1 << bits2;
    for (int p = threadIdx.x; p < n2; p += blockDim.x) {
        const int k2 = bitrev(p, bits2);
        int pp;
        if (k1 == 0) pp = bitrev((n2 - k2) & (n2 - 1), bits2);      // k' = N1 * ((N2 - k2) mod N2)
        else pp = n2 - 1 - p;                                       // k' = (N1 - k1) + N1 * (N2 - 1 - k2), same row
        if (pp < p) continue;                                       // each unordered pair once
        const long long k = (long long)k1 + (long long)N1 * k2;
        float2 ok, op;
        const float2 zk = s[saddr(p, 0, 1)], zp = s[saddr(pp, 0, 1)];
        spectral_pair(zk, zp, k, M, big_twiddle(k, nx, -1.f), scale, ok, op);
        s[saddr(p, 0, 1)] = ok;
        if (pp != p) s[saddr(pp, 0, 1)] = op;
    }
    __syncthreads();
}

}  // namespace

// ---------------------------------------------------------------------------
// whole transform in one CTA (M = nx/2 = 2^bits <= 2^14 complex points)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kFftThreads) poisson_fft_small_kernel(const float* __restrict__ n, long long n_stride,
                                                                        float* __restrict__ E, long long e_stride,
                                                                        int bits, double length) {
    extern __shared__ float2 sfft[];
    const int M = 1 << bits;
    const float2* src = reinterpret_cast<const float2*>(n + (size_t)blockIdx.x * n_stride);
    for (int j = threadIdx.x; j < M; j += blockDim.x) {
        const float2 v = src[j];
        sfft[saddr(j, 0, 1)] = make_float2(__fsub_rn(v.x, 1.0f), __fsub_rn(v.y, 1.0f));
    }
    __syncthreads();
    fft_dif(sfft, bits, 1, -1.f);
    spectral_self_row(sfft, bits, 0, 1, M, 2LL * M, (float)(length / (6.283185307179586476925 * (double)M)));
    fft_dit(sfft, bits, 1, +1.f);
    float2* dst = reinterpret_cast<float2*>(E + (size_t)blockIdx.x * e_stride);
    for (int j = threadIdx.x; j < M; j += blockDim.x) dst[j] = sfft[saddr(j, 0, 1)];
}

// ---------------------------------------------------------------------------
// four-step, pass A: forward column transforms.  grid = (N2 / T, B)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kFftStepThreads, 3) poisson_fft_cols_fwd_kernel(const float* __restrict__ n, long long n_stride,
                                                                                  float2* __restrict__ Y, int bits1, int bits2,
                                                                                  int T) {
    extern __shared__ float2 sfft[];
    const int N1 = 1 << bits1;
    const long long N2 = 1LL << bits2, M = (long long)N1 << bits2;
    const long long j2_0 = (long long)blockIdx.x * T;
    const int t_bits = 31 - __clz(T);
    const float2* src = reinterpret_cast<const float2*>(n + (size_t)blockIdx.y * n_stride);
    for (int w = threadIdx.x; w < N1 * T; w += blockDim.x) {
        const int t = w & (T - 1), j1 = w >> t_bits;
        const float2 v = src[(size_t)j1 * N2 + j2_0 + t];
        sfft[w] = make_float2(__fsub_rn(v.x, 1.0f), __fsub_rn(v.y, 1.0f));
    }
    __syncthreads();
    fft_dif(sfft, bits1, T, -1.f);
    float2* dst = Y + (size_t)blockIdx.y * M;
    for (int w = threadIdx.x; w < N1 * T; w += blockDim.x) {
        const int t = w & (T - 1), pos = w >> t_bits;
        const int k1 = bitrev(pos, bits1);
        const long long j2 = j2_0 + t;
        dst[(size_t)k1 * N2 + j2] = cmul(sfft[w], big_twiddle((j2 * k1) & (M - 1), M, -1.f));      // W_M^(j2*k1)
    }
}

// pass B: row pairs.  grid = (N1/2 + 1, B): pair 0 -> row 0 alone, pair N1/2 -> row N1/2 alone,
// pair q -> rows q and N1 - q, interleaved in shared memory (element idx of row t at s[2*idx + t]).
__global__ void __launch_bounds__(kFftStepThreads, 3) poisson_fft_rows_kernel(float2* __restrict__ Y, int bits1, int bits2,
                                                                              double length) {
    extern __shared__ float2 sfft[];
    const int N1 = 1 << bits1, N2 = 1 << bits2;
    const long long M = (long long)N1 << bits2, nx = 2 * M;
    const float scale = (float)(length / (6.283185307179586476925 * (double)M));
    const int q = blockIdx.x;
    float2* base = Y + (size_t)blockIdx.y * M;
    if (q == 0 || 2 * q == N1) {
        // ---- a row that pairs with itself ----------------------------------------------------
        float2* row = base + (size_t)q * N2;
        for (int i = threadIdx.x; i < N2; i += blockDim.x) sfft[saddr(i, 0, 1)] = row[i];
        __syncthreads();
        fft_dif(sfft, bits2, 1, -1.f);
        spectral_self_row(sfft, bits2, q, N1, M, nx, scale);
        fft_dit(sfft, bits2, 1, +1.f);
        for (int i = threadIdx.x; i < N2; i += blockDim.x)
            row[i] = cmul(sfft[saddr(i, 0, 1)], big_twiddle(((long long)i * q) & (M - 1), M, +1.f));   // conj twiddle
        return;
    }
    // ---- rows k1 = q and k1' = N1 - q: bin (k1, k2) pairs with (k1', N2 - 1 - k2), i.e. position N2-1-p ----
    const int k1a = q, k1b = N1 - q;
    float2* rowa = base + (size_t)k1a * N2;
    float2* rowb = base + (size_t)k1b * N2;
    for (int i = threadIdx.x; i < N2; i += blockDim.x) {
        sfft[2 * i] = rowa[i];
        sfft[2 * i + 1] = rowb[i];
    }
    __syncthreads();
    fft_dif(sfft, bits2, 2, -1.f);
    for (int p = threadIdx.x; p < N2; p += blockDim.x) {
        const int pp = N2 - 1 - p;
        const long long k = (long long)k1a + (long long)N1 * bitrev(p, bits2);
        float2 ok, op;
        spectral_pair(sfft[2 * p], sfft[2 * pp + 1], k, M, big_twiddle(k, nx, -1.f), scale, ok, op);
        sfft[2 * p] = ok;
        sfft[2 * pp + 1] = op;
    }
    __syncthreads();
    fft_dit(sfft, bits2, 2, +1.f);
    for (int i = threadIdx.x; i < N2; i += blockDim.x) {
        rowa[i] = cmul(sfft[2 * i], big_twiddle(((long long)i * k1a) & (M - 1), M, +1.f));
        rowb[i] = cmul(sfft[2 * i + 1], big_twiddle(((long long)i * k1b) & (M - 1), M, +1.f));
    }
}

// pass C: inverse column transforms, (E_2j, E_2j+1) out.  grid = (N2 / T, B)
__global__ void __launch_bounds__(kFftStepThreads, 3) poisson_fft_cols_inv_kernel(const float2* __restrict__ Y,
                                                                                  float* __restrict__ E, long long e_stride,
                                                                                  int bits1, int bits2, int T) {
    extern __shared__ float2 sfft[];
    const int N1 = 1 << bits1;
    const long long N2 = 1LL << bits2, M = (long long)N1 << bits2;
    const long long j2_0 = (long long)blockIdx.x * T;
    const int t_bits = 31 - __clz(T);
    const float2* src = Y + (size_t)blockIdx.y * M;
    for (int w = threadIdx.x; w < N1 * T; w += blockDim.x) {
        const int t = w & (T - 1), k1 = w >> t_bits;
        sfft[w] = src[(size_t)k1 * N2 + j2_0 + t];
    }
    __syncthreads();
    fft_dif(sfft, bits1, T, +1.f);                 // natural k1 in -> bit-reversed j1 out
    float2* dst = reinterpret_cast<float2*>(E + (size_t)blockIdx.y * e_stride);
    for (int w = threadIdx.x; w < N1 * T; w += blockDim.x) {
        const int t = w & (T - 1), pos = w >> t_bits;
        dst[(size_t)bitrev(pos, bits1) * N2 + j2_0 + t] = sfft[w];
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
    if (bits - 1 <= kFftRowBits) return 0;
    return (size_t)B * (size_t)(nx / 2) * sizeof(float2);
}

int launch_poisson_fft(const float* n, long long n_stride, float* E, long long e_stride, int B, int nx,
                       double length, void* workspace, cudaStream_t stream) {
    const int bits = ilog2_exact(nx);
    if (bits < kFftMinBits || bits > kFftMaxBits)
        return set_error(FLUXGNN_EUNSUP, "FFT field solve needs nx = 2^%d..2^%d, got %d", kFftMinBits, kFftMaxBits, nx);
    if ((n_stride & 1) || (e_stride & 1) || ((uintptr_t)n & 7) || ((uintptr_t)E & 7))
        return set_error(FLUXGNN_EINVAL, "FFT field solve needs 8-byte aligned density / field rows");
    const int mbits = bits - 1;                               // M = nx/2 complex points
    if (mbits <= kFftRowBits) {
        const int M = 1 << mbits;
        const size_t smem = (size_t)(M + M / 16) * sizeof(float2);
        FLUXGNN_CUDA_OK(cudaFuncSetAttribute(poisson_fft_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)smem));
        const int threads = M / 16 < kFftThreads ? (M / 16 < 64 ? 64 : M / 16) : kFftThreads;
        poisson_fft_small_kernel<<<B, threads, smem, stream>>>(n, n_stride, E, e_stride, mbits, length);
        FLUXGNN_CUDA_OK(cudaGetLastError());
        count_launch();
        return FLUXGNN_OK;
    }
    if (workspace == nullptr)
        return set_error(FLUXGNN_EINVAL, "field solve for nx=%d needs fluxgnn_poisson_workspace_bytes() of scratch", nx);
    if (B > 65535) return set_error(FLUXGNN_EUNSUP, "field solve for nx=%d handles at most 65535 ICs per call", nx);
    // rows as long as configured, but never leave the column transform longer than the column tile
    int bits2 = FLUXGNN_FFT_STEP_ROW_BITS;
    if (mbits - bits2 > FLUXGNN_FFT_STEP_COL_BITS - 1) bits2 = mbits - (FLUXGNN_FFT_STEP_COL_BITS - 1);
    const int bits1 = mbits - bits2;
    const int N1 = 1 << bits1, N2 = 1 << bits2;
    const int T = (1 << FLUXGNN_FFT_STEP_COL_BITS) / N1;     // N1 * T complex per column tile
    const size_t smem = (size_t)(1 << FLUXGNN_FFT_STEP_COL_BITS) * sizeof(float2);
    const size_t smem_row = (size_t)(2 * N2 + N2 / 8) * sizeof(float2);    // two interleaved rows, or one padded row
    float2* Y = (float2*)workspace;
    FLUXGNN_CUDA_OK(cudaFuncSetAttribute(poisson_fft_cols_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    FLUXGNN_CUDA_OK(cudaFuncSetAttribute(poisson_fft_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_row));
    FLUXGNN_CUDA_OK(cudaFuncSetAttribute(poisson_fft_cols_inv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 gcol((unsigned)(N2 / T), (unsigned)B), grow((unsigned)(N1 / 2 + 1), (unsigned)B);
    poisson_fft_cols_fwd_kernel<<<gcol, kFftStepThreads, smem, stream>>>(n, n_stride, Y, bits1, bits2, T);
    FLUXGNN_CUDA_OK(cudaGetLastError());
    poisson_fft_rows_kernel<<<grow, kFftStepThreads, smem_row, stream>>>(Y, bits1, bits2, length);
    FLUXGNN_CUDA_OK(cudaGetLastError());
    poisson_fft_cols_inv_kernel<<<gcol, kFftStepThreads, smem, stream>>>(Y, E, e_stride, bits1, bits2, T);
    FLUXGNN_CUDA_OK(cudaGetLastError());
    count_launch(3);
    return FLUXGNN_OK;
}

}  // namespace fluxgnn
