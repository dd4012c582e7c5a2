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
#include <cuda.h>

#include <cstring>

#include "common.cuh"
#include "field_kernels.cuh"

namespace fluxgnn {

namespace {

constexpr int kFftThreads = 512;       // whole-transform kernel (upper bound; it launches nx/16)
constexpr int kFftStepThreads = 256;   // four-step kernels
#ifndef FLUXGNN_FFT_ROW_CTAS
#define FLUXGNN_FFT_ROW_CTAS 2         // row kernel: CTAs per SM the register budget is set for
#endif
#define FLUXGNN_FFT_COL_CTAS 2         // column kernels: CTAs per SM the register budget is set for (3 measured slower:
                                       // the L1 data pipe, not occupancy, limits them -- 32-byte segments cost a wavefront each)

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
    // packed: (c, c) * v + (-sn, sn) * (v.y, v.x)
    return __ffma2_rn(make_float2(c, c), v, __fmul2_rn(make_float2(-sn, sn), make_float2(v.y, v.x)));
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
                v[A] = __fadd2_rn(u, x);
                v[A + HALF] = mul_root16<mm * (8 / HALF)>(__fadd2_rn(u, make_float2(-x.x, -x.y)), sign);
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
                v[A] = __fadd2_rn(u, x);
                v[A + HALF] = __fadd2_rn(u, make_float2(-x.x, -x.y));
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
// The diagonal spectral multiplier of the distributed solve (see launch_poisson_dist_local): the local bin kl of rank
// kr is the global bin k = kr + G kl of an nx-point COMPLEX transform; E_hat = i rho_hat / k_phys with the signed
// wavenumber (k > nx/2 -> k - nx), zero for k = 0 and for the Nyquist bin (src/baseline_solver.py:62-66: the real
// part drops it; here two real signals share one complex transform, so it must be removed explicitly).
__device__ __forceinline__ float2 spectral_diag(float2 z, long long k, long long nx, float scale) {
    if (k == 0 || 2 * k == nx) return make_float2(0.f, 0.f);
    const float kk = (float)(2 * k < nx ? k : k - nx);            // |kk| < 2^24: exact
    const float f = __fdividef(scale, kk);
    return make_float2(-f * z.y, f * z.x);
}

// kDist = false: the real-input half-length solve of one IC.  kDist = true: in-place complex transform of one of the
// P signals of a rank (no n0 subtraction, diagonal multiplier of global bins kr + G kl).
template <bool kDist>
__global__ void __launch_bounds__(kFftThreads) poisson_fft_small_kernel(const float* __restrict__ n, long long n_stride,
                                                                        float* __restrict__ E, long long e_stride,
                                                                        int bits, double length, int G, int kr) {
    extern __shared__ float2 sfft[];
    const int M = 1 << bits;
    const float2* src = reinterpret_cast<const float2*>(n + (size_t)blockIdx.x * n_stride);
    for (int j = threadIdx.x; j < M; j += blockDim.x) {
        const float2 v = src[j];
        sfft[saddr(j, 0, 1)] = kDist ? v : make_float2(__fsub_rn(v.x, 1.0f), __fsub_rn(v.y, 1.0f));
    }
    __syncthreads();
    fft_dif(sfft, bits, 1, -1.f);
    if constexpr (kDist) {
        const long long nx = (long long)M * G;
        const float scale = (float)(length / (6.283185307179586476925 * (double)M));     // carries the 1/M of the local inverse
        for (int p = threadIdx.x; p < M; p += blockDim.x)
            sfft[saddr(p, 0, 1)] = spectral_diag(sfft[saddr(p, 0, 1)], (long long)kr + (long long)G * bitrev(p, bits), nx, scale);
        __syncthreads();
    } else {
        spectral_self_row(sfft, bits, 0, 1, M, 2LL * M, (float)(length / (6.283185307179586476925 * (double)M)));
    }
    fft_dit(sfft, bits, 1, +1.f);
    float2* dst = reinterpret_cast<float2*>(E + (size_t)blockIdx.x * e_stride);
    for (int j = threadIdx.x; j < M; j += blockDim.x) dst[j] = sfft[saddr(j, 0, 1)];
}

// ---------------------------------------------------------------------------
// Four-step kernels.  Index split of the half-length transform: element j = j1*N2 + j2, bin
// k = k1 + N1*k2, N2 = 2^12.  The whole thing is ONE decimation-in-frequency network over M points
// whose first log2(N1) stages (kernel A) work on columns and whose last 12 stages (kernel B) work
// on rows; the "four-step twiddle" W_M^(j2*k1) is simply the inter-pass twiddle of that network
// (W_L^(base*q) with base = b1*N2 + j2), so it costs one sincospi per 16 elements instead of one
// per element.  Kernel C is the transposed (decimation-in-time) network of kernel A.
// Y keeps its rows in POSITION order (row pos holds k1 = bitrev(pos)): no scatter anywhere.
// The first pass of every kernel loads straight from global memory into registers and the last
// pass stores straight from registers; shared memory is touched only between passes.
// Arithmetic is packed (add/mul/fma.f32x2): a complex add is one instruction.
// ---------------------------------------------------------------------------
namespace {

constexpr int kRowBits = FLUXGNN_FFT_STEP_ROW_BITS;                // 12
constexpr int kRowLen = 1 << kRowBits;                             // 4096
constexpr int kRowPad = kRowLen + kRowLen / 16;                    // padded row in shared memory
static_assert(kRowBits == 12 && kFftStepThreads == 256, "row kernel is written for 4096 = 16*16*16 and 256 threads");

__device__ __forceinline__ float2 pmul(float2 v, float2 w) {       // v * w, two packed instructions
    return __ffma2_rn(make_float2(w.x, w.x), v, __fmul2_rn(make_float2(-w.y, w.y), make_float2(v.y, v.x)));
}
__device__ __forceinline__ float2 psqr(float2 w) {
    return make_float2(fmaf(w.x, w.x, -w.y * w.y), 2.f * w.x * w.y);
}
__host__ __device__ constexpr int top_bit(int q) { return q <= 1 ? q : 2 * top_bit(q >> 1); }

// pw[q] = w^q, q < R; every power is at most four multiplications away from the sincospi.
// (template recursion: the indices must be compile-time constants or the arrays leave the registers)
template <int R, int Q = 2>
__device__ __forceinline__ void twiddle_powers_from(float2 (&pw)[R]) {
    if constexpr (Q < R) {
        constexpr int hb = top_bit(Q);
        if constexpr (Q == hb) pw[Q] = psqr(pw[Q / 2]);
        else pw[Q] = pmul(pw[Q - hb], pw[hb]);
        twiddle_powers_from<R, Q + 1>(pw);
    }
}
template <int R>
__device__ __forceinline__ void twiddle_powers(float2 w, float2 (&pw)[R]) {
    pw[0] = make_float2(1.f, 0.f);
    if constexpr (R > 1) pw[1] = w;
    twiddle_powers_from<R>(pw);
}
// register a (which holds / will hold output q = bitrev(a)) *= w^bitrev(a)
template <int R, int A = 1>
__device__ __forceinline__ void apply_powers(float2 (&v)[R], const float2 (&pw)[R]) {
    if constexpr (A < R) {
        v[A] = pmul(v[A], pw[brev_bits(A, ilog2c(R))]);
        apply_powers<R, A + 1>(v, pw);
    }
}

// exp(sign * 2 pi i * num / 2^den_bits), num < 2^24: the argument is exact in fp32
__device__ __forceinline__ float2 unit_root(int num, int den_bits, float sign) {
    float s, c;
    sincospif(sign * (float)num * __int_as_float((127 + 1 - den_bits) << 23), &s, &c);
    return make_float2(c, s);
}

// The twiddles of one work item: all R powers in registers (14 + 15 multiplications per 16 points).
template <int R>
struct TwiddleSet {
    float2 pw[R];
    __device__ __forceinline__ void init(float2 w) { twiddle_powers<R>(w, pw); }
    __device__ __forceinline__ void apply(float2 (&v)[R]) const { apply_powers<R>(v, pw); }
};

// L2 residency hints (FLUXGNN_FFT_L2_HINTS): the spectrum Y (8 B per complex point, 64 MiB at 2^24
// cells) is written by kernel A, rewritten in place by B and consumed by C; marking it evict_last
// and the streamed density / consumed spectrum evict_first keeps it in the 126 MB L2 between kernels.
#ifndef FLUXGNN_FFT_L2_HINTS
#define FLUXGNN_FFT_L2_HINTS 1
#endif
enum class Hint { kNone, kFirst, kLast };
template <Hint H>
__device__ __forceinline__ uint64_t make_policy() {
    uint64_t p = 0;
#if FLUXGNN_FFT_L2_HINTS
    if constexpr (H == Hint::kFirst) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    if constexpr (H == Hint::kLast) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
#endif
    return p;
}
template <Hint H>
__device__ __forceinline__ float2 gload(const float2* p, uint64_t pol) {
#if FLUXGNN_FFT_L2_HINTS
    if constexpr (H != Hint::kNone) {
        float2 v;
        asm("ld.global.L2::cache_hint.v2.f32 {%0,%1}, [%2], %3;" : "=f"(v.x), "=f"(v.y) : "l"(p), "l"(pol));
        return v;
    }
#endif
    return *p;
}
template <Hint H>
__device__ __forceinline__ void gstore(float2* p, float2 v, uint64_t pol) {
#if FLUXGNN_FFT_L2_HINTS
    if constexpr (H != Hint::kNone) {
        asm volatile("st.global.L2::cache_hint.v2.f32 [%0], {%1,%2}, %3;" ::"l"(p), "f"(v.x), "f"(v.y), "l"(pol) : "memory");
        return;
    }
#endif
    *p = v;
}

__device__ __forceinline__ int cpad(int e) { return e + ((e >> 5) << 2); }     // column tile: 4 pads per 32
// Column-tile layouts (index of point e = row * T + column):
//   0  padded (cpad): the plain and cp.async-staged kernels
//   1  dense: what a TMA tensor copy writes / reads without swizzle
//   2  dense + SWIZZLE_128B: the 16-byte chunk index (byte address bits 4..6) XORed with address bits 7..9, the
//      pattern the TMA unit applies; it spreads the short-stride passes over all banks like the padding does
template <int LAYOUT>
__device__ __forceinline__ int cidx(int e) {
    if constexpr (LAYOUT == 0) return cpad(e);
    else if constexpr (LAYOUT == 1) return e;
    else return e ^ (((e >> 4) & 7) << 1);
}
__device__ __forceinline__ int rpad(int i) { return i + (i >> 4); }             // row: 1 pad per 16

// ---- kernel A / C: column passes -------------------------------------------------------------
// One pass over blocks of 2^LB rows (j1 units) of a tile of N1 x T elements held at s[cpad(j1*T + t)].
// kInv = false: DIF, forward, natural j1 in -> position order out;  kInv = true: the transposed DIT pass.
template <int BITS1, int TILE_BITS, int LB, bool kInv, int THREADS, bool kSub1 = true, bool kStaged = false, int LAYOUT = 0,
          bool kStagedOut = false>
__device__ __forceinline__ void column_pass(const float2* __restrict__ gin, float2* __restrict__ gout, float2* s,
                                            int j2_0) {
    constexpr int RB = LB >= 4 ? 4 : LB, R = 1 << RB, SUBB = LB - RB, SUB = 1 << SUBB;
    constexpr int TB = TILE_BITS - BITS1, T = 1 << TB;
    // forward: the pass over the longest blocks comes first and reads global memory; inverse: last and writes it
    // (kStaged: the tile was copied into `s` asynchronously beforehand, so the first pass reads shared memory too)
    constexpr bool kFirst = kInv ? (SUBB == 0) : (LB == BITS1);
    constexpr bool kGlobalIn = kFirst && !kStaged;
    constexpr bool kGlobalOut = !kStagedOut && (kInv ? (LB == BITS1) : (SUBB == 0));
    constexpr int ITEMS = (1 << TILE_BITS) >> RB;
    static_assert(ITEMS % THREADS == 0, "tile too small for the CTA");
    constexpr int ITERS = ITEMS / THREADS;
    constexpr bool kSharedTw = (SUB * T <= THREADS);   // (t, base) do not depend on the iteration
    // forward: density in (streamed), spectrum out (kept);  inverse: spectrum in (consumed), field out
    constexpr Hint kHin = Hint::kFirst, kHout = kInv ? Hint::kNone : Hint::kLast;
    const uint64_t pol_in = kGlobalIn ? make_policy<kHin>() : 0, pol_out = kGlobalOut ? make_policy<kHout>() : 0;
    const float sign = kInv ? 1.f : -1.f;
    TwiddleSet<R> pw;
    if constexpr (kSharedTw) {
        const int w = threadIdx.x, t = w & (T - 1), base = (w >> TB) & (SUB - 1);
        pw.init(unit_root((base << kRowBits) + j2_0 + t, LB + kRowBits, sign));
    }
#pragma unroll
    for (int it = 0; it < ITERS; ++it) {
        const int w = threadIdx.x + it * THREADS;
        const int t = w & (T - 1), q = w >> TB, base = q & (SUB - 1), blk = q >> SUBB;
        const int j1_0 = (blk << LB) + base;
        float2 v[R];
        if constexpr (kGlobalIn) {
#pragma unroll
            for (int m = 0; m < R; ++m)
                v[m] = gload<kHin>(gin + (((size_t)(j1_0 + (m << SUBB)) << kRowBits) + j2_0 + t), pol_in);
        } else {
#pragma unroll
            for (int m = 0; m < R; ++m) v[m] = s[cidx<LAYOUT>(((j1_0 + (m << SUBB)) << TB) + t)];
        }
        if constexpr (kFirst && !kInv && kSub1) {    // rho = n - n0 (src/baseline_solver.py:60)
#pragma unroll
            for (int m = 0; m < R; ++m) v[m] = make_float2(__fsub_rn(v[m].x, 1.0f), __fsub_rn(v[m].y, 1.0f));
        }
        if constexpr (!kSharedTw) pw.init(unit_root((base << kRowBits) + j2_0 + t, LB + kRowBits, sign));
        if constexpr (!kInv) {
            DifStages<R>::run(v, sign);
            pw.apply(v);
        } else {
            pw.apply(v);
            DitStages<R>::run(v, sign);
        }
        if constexpr (kGlobalOut) {
#pragma unroll
            for (int m = 0; m < R; ++m)
                gstore<kHout>(gout + (((size_t)(j1_0 + (m << SUBB)) << kRowBits) + j2_0 + t), v[m], pol_out);
        } else {
#pragma unroll
            for (int m = 0; m < R; ++m) s[cidx<LAYOUT>(((j1_0 + (m << SUBB)) << TB) + t)] = v[m];
        }
    }
    if constexpr (!kGlobalOut) __syncthreads();
}

template <int BITS1, int TILE_BITS, int LB, int THREADS, bool kSub1 = true, bool kStaged = false, int LAYOUT = 0,
          bool kStagedOut = false>
__device__ __forceinline__ void columns_forward(const float2* gin, float2* gout, float2* s, int j2_0) {
    column_pass<BITS1, TILE_BITS, LB, false, THREADS, kSub1, kStaged, LAYOUT, kStagedOut>(gin, gout, s, j2_0);
    constexpr int SUBB = LB - (LB >= 4 ? 4 : LB);
    if constexpr (SUBB > 0) columns_forward<BITS1, TILE_BITS, SUBB, THREADS, kSub1, kStaged, LAYOUT, kStagedOut>(gin, gout, s, j2_0);
}
template <int BITS1, int TILE_BITS, int LB, int THREADS, bool kStaged = false, int LAYOUT = 0, bool kStagedOut = false>
__device__ __forceinline__ void columns_inverse(const float2* gin, float2* gout, float2* s, int j2_0) {
    column_pass<BITS1, TILE_BITS, LB, true, THREADS, true, kStaged, LAYOUT, kStagedOut>(gin, gout, s, j2_0);
    if constexpr (LB < BITS1) columns_inverse<BITS1, TILE_BITS, LB + 4, THREADS, kStaged, LAYOUT, kStagedOut>(gin, gout, s, j2_0);
}
__host__ __device__ constexpr int first_inverse_lb(int bits1) { return bits1 < 4 ? bits1 : ((bits1 & 3) ? (bits1 & 3) : 4); }
__host__ __device__ constexpr int column_tile_bits(int bits1) {
    return bits1 + 2 > FLUXGNN_FFT_STEP_COL_BITS ? bits1 + 2 : FLUXGNN_FFT_STEP_COL_BITS;    // at least 4 columns per tile
}
// a 2^14 tile (one CTA per SM by shared memory) runs 512 threads, a 2^13 tile 256 at FLUXGNN_FFT_COL_CTAS per SM
__host__ __device__ constexpr int column_threads(int bits1) { return column_tile_bits(bits1) <= 13 ? kFftStepThreads : 512; }

}  // namespace

// pass A: forward column stages.  grid = (N2 / T, B).  kSub1: the input is the density (rho = n - 1 is formed
// on the fly); otherwise it is taken as it is (the complex signals of the distributed solve).
template <int BITS1, bool kSub1 = true>
__global__ void __launch_bounds__(column_threads(BITS1), column_tile_bits(BITS1) <= 13 ? FLUXGNN_FFT_COL_CTAS : 1)
poisson_fft_cols_fwd_kernel(const float* __restrict__ n, long long n_stride, float2* __restrict__ Y) {
    extern __shared__ float2 sfft[];
    constexpr int TILE_BITS = column_tile_bits(BITS1);
    const int j2_0 = blockIdx.x << (TILE_BITS - BITS1);
    const float2* src = reinterpret_cast<const float2*>(n + (size_t)blockIdx.y * n_stride);
    float2* dst = Y + ((size_t)blockIdx.y << (BITS1 + kRowBits));
    columns_forward<BITS1, TILE_BITS, BITS1, column_threads(BITS1), kSub1>(src, dst, sfft, j2_0);
}

// pass C: inverse column stages, (E_2j, E_2j+1) out.  grid = (N2 / T, B)
template <int BITS1>
__global__ void __launch_bounds__(column_threads(BITS1), column_tile_bits(BITS1) <= 13 ? FLUXGNN_FFT_COL_CTAS : 1)
poisson_fft_cols_inv_kernel(const float2* __restrict__ Y, float* __restrict__ E, long long e_stride) {
    extern __shared__ float2 sfft[];
    constexpr int TILE_BITS = column_tile_bits(BITS1);
    const int j2_0 = blockIdx.x << (TILE_BITS - BITS1);
    const float2* src = Y + ((size_t)blockIdx.y << (BITS1 + kRowBits));
    float2* dst = reinterpret_cast<float2*>(E + (size_t)blockIdx.y * e_stride);
    columns_inverse<BITS1, TILE_BITS, first_inverse_lb(BITS1), column_threads(BITS1)>(src, dst, sfft, j2_0);
}

// ---- column passes with asynchronously staged tiles (long grids) ------------------------------------------------
// The plain column kernels above load a tile (2^13 points as 32-byte pieces 32 KiB apart), transform it and store it,
// strictly in sequence per CTA: with two CTAs per SM only ~20 KiB per SM are in flight on average and the passes reach
// about half of the HBM bandwidth.  Here ONE persistent 512-thread CTA per SM owns two tile buffers: while it transforms
// the tile in one of them, cp.async (16-byte copies that bypass the registers and L1) fills the other with the next
// tile, so a whole tile per SM is always in flight.  Same passes on the same padded layout, hence identical results.
constexpr int kStagedThreads = 512;
constexpr int kStagedTileBits = 13;
constexpr int kStagedTile = (1 << kStagedTileBits) + ((1 << kStagedTileBits) >> 5) * 4;       // float2 elements per buffer

__device__ __forceinline__ void cp_async_16(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// copy tile `it` = (IC b, column tile) of a [B][N1][4096] float2 matrix into the padded buffer: row j1 = T points
template <int BITS1>
__device__ __forceinline__ void stage_tile(float2* buf, const float2* base, long long ic_stride, int it, int tiles) {
    constexpr int TB = kStagedTileBits - BITS1, T = 1 << TB, N1 = 1 << BITS1;
    constexpr int kChunksPerRow = T / 2;                              // 16-byte chunks (2 points) per row
    const int b = it / tiles, tile = it - b * tiles;
    const float2* src = base + (size_t)b * ic_stride + (tile << TB);
    for (int c = threadIdx.x; c < N1 * kChunksPerRow; c += kStagedThreads) {
        const int j1 = c / kChunksPerRow, p = (c - j1 * kChunksPerRow) * 2;
        cp_async_16(buf + cpad((j1 << TB) + p), src + ((size_t)j1 << kRowBits) + p);
    }
}

// kInv = false: pass A (density or complex signal -> Y);  kInv = true: pass C (Y -> field).  in / out strides in float2 per IC.
template <int BITS1, bool kInv, bool kSub1>
__global__ void __launch_bounds__(kStagedThreads, 1) poisson_fft_cols_staged_kernel(const float2* __restrict__ in,
                                                                                   long long in_stride,
                                                                                   float2* __restrict__ out,
                                                                                   long long out_stride, int tiles,
                                                                                   int total) {
    extern __shared__ __align__(16) float2 sfft[];
    constexpr int TB = kStagedTileBits - BITS1;
    int cur = 0;
    if ((int)blockIdx.x < total) stage_tile<BITS1>(sfft, in, in_stride, blockIdx.x, tiles);
    cp_async_commit();
    for (int it = blockIdx.x; it < total; it += gridDim.x, cur ^= 1) {
        const int nit = it + gridDim.x;
        float2* tile_buf = sfft + cur * kStagedTile;
        if (nit < total) stage_tile<BITS1>(sfft + (cur ^ 1) * kStagedTile, in, in_stride, nit, tiles);
        cp_async_commit();
        cp_async_wait<1>();                           // this tile's copies have landed (the next tile's may be in flight)
        __syncthreads();
        const int b = it / tiles, tile = it - b * tiles;
        float2* dst = out + (size_t)b * out_stride;
        if constexpr (kInv)
            columns_inverse<BITS1, kStagedTileBits, first_inverse_lb(BITS1), kStagedThreads, true>(nullptr, dst, tile_buf, tile << TB);
        else
            columns_forward<BITS1, kStagedTileBits, BITS1, kStagedThreads, kSub1, true>(nullptr, dst, tile_buf, tile << TB);
        __syncthreads();                              // the buffer may be refilled by the next iteration's copies
    }
    cp_async_wait<0>();
}

// ---- column passes through the TMA unit (long grids) ------------------------------------------------------------
// The column tile (N1 rows of T points, 32..128 bytes each, 32 KiB apart) is one 3-D tensor box per 256 rows for the TMA
// unit (cp.async.bulk.tensor, SASS UTMALDG / UTMASTG): loads and stores then cost no load/store-unit wavefronts at all
// -- which, at one wavefront per 32-byte piece, is what bounds the kernels above -- and no registers.  One persistent
// 512-thread CTA per SM rotates three dense 64 KiB tile buffers (next tile loading, this tile being transformed in
// place, previous tile being stored); the buffers use the unit's 128-byte swizzle, which keeps every pass free of bank
// conflicts the way the padding does in the other kernels.  Same passes, same arithmetic, identical results.
constexpr int kTmaThreads = 512;
constexpr int kTmaTileBits = 13;
constexpr int kTmaBoxRows = 256;

__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, int c0, int c1, int c2, const void* src) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%1, %2, %3}], [%4];"
                 ::"l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(src)) : "memory");
}
__device__ __forceinline__ void tma_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }

template <int BITS1, bool kInv, bool kSub1, int LAYOUT>
__global__ void __launch_bounds__(kTmaThreads, 1) poisson_fft_cols_tma_kernel(const __grid_constant__ CUtensorMap map_in,
                                                                             const __grid_constant__ CUtensorMap map_out,
                                                                             int tiles, int total) {
    extern __shared__ __align__(1024) unsigned char tma_smem[];
    constexpr int TB = kTmaTileBits - BITS1, T = 1 << TB, N1 = 1 << BITS1;
    constexpr int kTile = 1 << kTmaTileBits;                           // points per tile buffer (dense)
    constexpr uint32_t kTileBytes = kTile * 8u;
    unsigned char* base = tma_smem + ((1024u - (smem_u32(tma_smem) & 1023u)) & 1023u);      // the swizzle needs 1024-byte alignment
    float2* bufs = reinterpret_cast<float2*>(base);
    uint64_t* full = reinterpret_cast<uint64_t*>(base + 3 * kTileBytes);
    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int i = 0; i < 3; ++i) mbar_init(&full[i], 1);
        mbar_fence_init();
    }
    __syncthreads();
    auto issue_load = [&](int it, int slot) {                          // thread 0
        const int b = it / tiles, tile = it - b * tiles;
        mbar_arrive_expect_tx(&full[slot], kTileBytes);
        for (int r = 0; r < N1; r += kTmaBoxRows)
            tma_load_3d(bufs + (size_t)slot * kTile + r * T, &map_in, 2 * (tile << TB), r, b, &full[slot]);
    };
    const int stride = gridDim.x;
    if (tid == 0) {
        if ((int)blockIdx.x < total) issue_load(blockIdx.x, 0);
        if ((int)blockIdx.x + stride < total) issue_load(blockIdx.x + stride, 1);
    }
    int k = 0;
    for (int it = blockIdx.x; it < total; it += stride, ++k) {
        const int slot = k % 3;
        float2* buf = bufs + (size_t)slot * kTile;
        mbar_wait(&full[slot], (uint32_t)((k / 3) & 1));
        const int b = it / tiles, tile = it - b * tiles;
        if constexpr (kInv)
            columns_inverse<BITS1, kTmaTileBits, first_inverse_lb(BITS1), kTmaThreads, true, LAYOUT, true>(nullptr, nullptr, buf, tile << TB);
        else
            columns_forward<BITS1, kTmaTileBits, BITS1, kTmaThreads, kSub1, true, LAYOUT, true>(nullptr, nullptr, buf, tile << TB);
        fence_proxy_async();                          // the tile (generic-proxy stores) becomes visible to the TMA unit
        __syncthreads();
        if (tid == 0) {
            for (int r = 0; r < N1; r += kTmaBoxRows) tma_store_3d(&map_out, 2 * (tile << TB), r, b, buf + r * T);
            tma_commit();
            // the buffer of iteration k-1 is the target of the load for iteration k+2: its store must have read it
            if (it + 2 * stride < total) {
                tma_wait_read<1>();
                issue_load(it + 2 * stride, (k + 2) % 3);
            }
        }
    }
    if (tid == 0) tma_wait_read<0>();                 // shared memory must outlive the stores that read it
}

namespace {

// ---- kernel B: rows ----------------------------------------------------------------------------
// Row passes over 4096 = 16*16*16 points; thread `tid` owns the same work item in both rows of a pair,
// so every twiddle set is generated once and used twice.
template <bool kInv>
__device__ __forceinline__ void row_butterfly(float2 (&v)[16], const float2 (&pw)[16], bool twiddled) {
    const float sign = kInv ? 1.f : -1.f;
    if constexpr (!kInv) {
        DifStages<16>::run(v, sign);
        if (twiddled) apply_powers<16>(v, pw);
    } else {
        if (twiddled) apply_powers<16>(v, pw);
        DitStages<16>::run(v, sign);
    }
}

// alpha/beta form of the pair-wise spectral step (header comment):  with f(k) = scale/k,
// s = (f(k) + f(M-k))/2, d = (f(k) - f(M-k))/2 and w = W_nx^k = (wx, wy),
//   Zt[k]   = i (d + s wy) Z[k]    - s wx conj(Z[M-k])
//   Zt[M-k] = i (-d + s wy) Z[M-k] + s wx conj(Z[k])
// (algebraically identical to untangle -> multiply by i/k -> re-tangle).  k in 1..M-1.
// s = scale/2 * M / (k (M-k)),  d = scale/2 * (M - 2k) / (k (M-k)): one reciprocal per pair (the XU pipe was the
// limiter with two divisions and two int->float conversions per pair); kf = (float)k is exact (k < 2^24).
__device__ __forceinline__ void spectral_pair_ab(float2& zk, float2& zp, float kf, float Mf, float2 w, float half_scale) {
    const float kpf = Mf - kf;
    const float r = __fdividef(half_scale, kf * kpf);
    const float s = r * Mf, d = r * (kpf - kf);
    const float swy = s * w.y, be = s * w.x;
    const float al = swy + d, alp = swy - d;
    const float2 a = zk, b = zp;
    zk = make_float2(-al * a.y - be * b.x, al * a.x + be * b.y);
    zp = make_float2(-alp * b.y + be * a.x, alp * b.x - be * a.y);
}

// W_32^(-c) = exp(-2 pi i c / 32), c = 0..15
__device__ __forceinline__ float2 root32_conj(int c) {
    constexpr float kC[16] = {1.f, 0.98078528040323044f, 0.92387953251128674f, 0.83146961230254524f,
                              0.70710678118654752f, 0.55557023301960222f, 0.38268343236508977f, 0.19509032201612827f,
                              0.f, -0.19509032201612827f, -0.38268343236508977f, -0.55557023301960222f,
                              -0.70710678118654752f, -0.83146961230254524f, -0.92387953251128674f, -0.98078528040323044f};
    constexpr float kS[16] = {0.f, 0.19509032201612827f, 0.38268343236508977f, 0.55557023301960222f,
                              0.70710678118654752f, 0.83146961230254524f, 0.92387953251128674f, 0.98078528040323044f,
                              1.f, 0.98078528040323044f, 0.92387953251128674f, 0.83146961230254524f,
                              0.70710678118654752f, 0.55557023301960222f, 0.38268343236508977f, 0.19509032201612827f};
    return make_float2(kC[c], -kS[c]);
}

}  // namespace

// pass B: row pairs.  grid = (N1/2 + 1, B): pair q -> the rows holding k1 = q and k1 = N1 - q
// (rows 0 and N1/2 pair with themselves and take the shared-memory spectral step).
__global__ void __launch_bounds__(kFftStepThreads, FLUXGNN_FFT_ROW_CTAS) poisson_fft_rows_kernel(float2* __restrict__ Y, int bits1,
                                                                              float scale) {
    extern __shared__ float2 sfft[];
    const int N1 = 1 << bits1, M = N1 << kRowBits;
    const int tid = threadIdx.x;
    const int k1a = blockIdx.x, k1b = (N1 - k1a) & (N1 - 1);
    const bool self = (k1a == k1b);
    float2* base = Y + ((size_t)blockIdx.y << (bits1 + kRowBits));
    float2* rowa = base + ((size_t)bitrev(k1a, bits1) << kRowBits);
    float2* rowb = base + ((size_t)bitrev(k1b, bits1) << kRowBits);
    float2* sa = sfft;
    float2* sb = sfft + kRowPad;
    float2 va[16], vb[16], pw[16];
    const uint64_t pol = make_policy<Hint::kLast>();

    // ---- forward pass 1: blocks of 4096, stride 256, straight from global memory ----
#pragma unroll
    for (int m = 0; m < 16; ++m) va[m] = gload<Hint::kLast>(rowa + tid + 256 * m, pol);
    if (!self) {
#pragma unroll
        for (int m = 0; m < 16; ++m) vb[m] = gload<Hint::kLast>(rowb + tid + 256 * m, pol);
    }
    twiddle_powers<16>(unit_root(tid, kRowBits, -1.f), pw);
    row_butterfly<false>(va, pw, true);
#pragma unroll
    for (int m = 0; m < 16; ++m) sa[rpad(tid + 256 * m)] = va[m];
    if (!self) {
        row_butterfly<false>(vb, pw, true);
#pragma unroll
        for (int m = 0; m < 16; ++m) sb[rpad(tid + 256 * m)] = vb[m];
    }
    __syncthreads();
    // ---- forward pass 2: blocks of 256, stride 16 ----
    const int e2 = (tid >> 4) * 256 + (tid & 15);
    twiddle_powers<16>(unit_root(tid & 15, 8, -1.f), pw);
#pragma unroll
    for (int m = 0; m < 16; ++m) va[m] = sa[rpad(e2 + 16 * m)];
    row_butterfly<false>(va, pw, true);
#pragma unroll
    for (int m = 0; m < 16; ++m) sa[rpad(e2 + 16 * m)] = va[m];
    if (!self) {
#pragma unroll
        for (int m = 0; m < 16; ++m) vb[m] = sb[rpad(e2 + 16 * m)];
        row_butterfly<false>(vb, pw, true);
#pragma unroll
        for (int m = 0; m < 16; ++m) sb[rpad(e2 + 16 * m)] = vb[m];
    }
    __syncthreads();
    // ---- forward pass 3 (blocks of 16), spectral step, inverse pass 3 ----
    if (!self) {
        // position p = 16 tid + m of row a (bin k1a + N1*bitrev(p)) pairs with position 4095 - p of row b:
        // this thread takes block tid of row a and block 255 - tid of row b, and the pairing stays in registers.
        const int ea = tid * 16, eb = (255 - tid) * 16;
#pragma unroll
        for (int m = 0; m < 16; ++m) va[m] = sa[rpad(ea + m)];
#pragma unroll
        for (int m = 0; m < 16; ++m) vb[m] = sb[rpad(eb + m)];
        row_butterfly<false>(va, pw, false);
        row_butterfly<false>(vb, pw, false);
        // k = k1a + N1*bitrev12(16 tid + m) = k0 + bitrev4(m) * M/16;  W_nx^k = W_nx^k0 * W_32^bitrev4(m)
        const int k0 = k1a + (bitrev(tid, 8) << bits1);
        const float2 w0 = unit_root(k0, bits1 + kRowBits + 1, -1.f);
        const float k0f = (float)k0, Mf = (float)M, M16f = (float)(M >> 4);
#pragma unroll
        for (int m = 0; m < 16; ++m) {
            const int c = brev_bits(m, 4);
            spectral_pair_ab(va[m], vb[15 - m], k0f + (float)c * M16f, Mf, pmul(w0, root32_conj(c)), 0.5f * scale);
        }
        row_butterfly<true>(va, pw, false);
        row_butterfly<true>(vb, pw, false);
#pragma unroll
        for (int m = 0; m < 16; ++m) sa[rpad(ea + m)] = va[m];
#pragma unroll
        for (int m = 0; m < 16; ++m) sb[rpad(eb + m)] = vb[m];
    } else {
#pragma unroll
        for (int m = 0; m < 16; ++m) va[m] = sa[rpad(tid * 16 + m)];
        row_butterfly<false>(va, pw, false);
#pragma unroll
        for (int m = 0; m < 16; ++m) sa[rpad(tid * 16 + m)] = va[m];
        __syncthreads();
        spectral_self_row(sa, kRowBits, k1a, N1, M, 2LL * M, scale);      // ends with a barrier
#pragma unroll
        for (int m = 0; m < 16; ++m) va[m] = sa[rpad(tid * 16 + m)];
        row_butterfly<true>(va, pw, false);
#pragma unroll
        for (int m = 0; m < 16; ++m) sa[rpad(tid * 16 + m)] = va[m];
    }
    __syncthreads();
    // ---- inverse pass 2 ----
    twiddle_powers<16>(unit_root(tid & 15, 8, +1.f), pw);
#pragma unroll
    for (int m = 0; m < 16; ++m) va[m] = sa[rpad(e2 + 16 * m)];
    row_butterfly<true>(va, pw, true);
#pragma unroll
    for (int m = 0; m < 16; ++m) sa[rpad(e2 + 16 * m)] = va[m];
    if (!self) {
#pragma unroll
        for (int m = 0; m < 16; ++m) vb[m] = sb[rpad(e2 + 16 * m)];
        row_butterfly<true>(vb, pw, true);
#pragma unroll
        for (int m = 0; m < 16; ++m) sb[rpad(e2 + 16 * m)] = vb[m];
    }
    __syncthreads();
    // ---- inverse pass 1, straight to global memory ----
    twiddle_powers<16>(unit_root(tid, kRowBits, +1.f), pw);
#pragma unroll
    for (int m = 0; m < 16; ++m) va[m] = sa[rpad(tid + 256 * m)];
    row_butterfly<true>(va, pw, true);
#pragma unroll
    for (int m = 0; m < 16; ++m) gstore<Hint::kLast>(rowa + tid + 256 * m, va[m], pol);
    if (!self) {
#pragma unroll
        for (int m = 0; m < 16; ++m) vb[m] = sb[rpad(tid + 256 * m)];
        row_butterfly<true>(vb, pw, true);
#pragma unroll
        for (int m = 0; m < 16; ++m) gstore<Hint::kLast>(rowb + tid + 256 * m, vb[m], pol);
    }
}

// ---------------------------------------------------------------------------
// Fused classical step for long grids (BASELINE.json configs[4]; src/baseline_solver.py:80-101).
// Inside a multi-step rollout the field E_s is consumed only by the next finite-volume update, and that update
// is cell-local except for one neighbour on each side.  So pass C of step s (inverse column stages -> E_s), the
// finite-volume update s -> s+1 and pass A of step s+1 (forward column stages of rho' = n' - 1) run as ONE column
// kernel: the thread that ends the inverse network with E of 16 (row, column) positions in registers is the thread
// that starts the forward network with rho' of the same positions.  Per step and cell the kernel reads the
// spectrum (4 B), n and u (8 B) and writes n', u' (8 B) and the new spectrum (4 B): 24 B, plus 8 B for the row
// kernel = 32 B per cell-update instead of 44, in 2 launches instead of 4; E is never written between steps.
// Between steps n and u live in a TILE-MAJOR private layout P[b][tile][row j1][2T floats] -- the 2T cells a column
// tile owns in matrix row j1 are contiguous and a whole tile is one contiguous block, so a warp's accesses are
// 256 contiguous bytes instead of eight 32-byte pieces 32 KiB apart -- and the three neighbour values a tile needs
// per matrix row travel through side arrays H[b][3][tile][j1] (last n, last u, first u of the tile's row), which
// the owning tile writes and its neighbours read as contiguous runs.  Both are double-buffered.
// ---------------------------------------------------------------------------
struct FusedColsArgs {
    float2* Y;                       // [B][N1][4096] spectrum, transformed in place tile by tile
    const float* Pn_in;              // tile-major n, u of the current step
    const float* Pu_in;
    float* Pn_out;
    float* Pu_out;
    const float* H_in;               // [B][3][tiles][N1]
    float* H_out;
    float* nat_out;                  // nullable [B][3][nx]: n', u' also in the natural layout (last step of a rollout)
    float c, dt, nu, dx2;
};

template <int BITS1, int TILE_BITS, int LB, int THREADS>
__device__ __forceinline__ void columns_inverse_to_smem(const float2* gin, float2* s, int j2_0) {
    if constexpr (LB < BITS1) {
        column_pass<BITS1, TILE_BITS, LB, true, THREADS>(gin, nullptr, s, j2_0);
        columns_inverse_to_smem<BITS1, TILE_BITS, LB + 4, THREADS>(gin, s, j2_0);
    }
}

// One persistent CTA of 512 threads per SM walks the column tiles.  For every tile the n and u blocks (64 KiB each,
// contiguous in the tile-major layout) and the three neighbour rows (N1 floats each) are fetched by 1-D bulk copies
// (cp.async.bulk, the TMA unit) that run under the inverse column stages; the next tile's copies are issued as soon
// as the finite-volume update has read this tile's, and run under the forward column stages.  The spectrum tile (32-byte
// pieces 32 KiB apart) and the outputs go through ordinary loads / stores.
constexpr int kFusedThreads = 512;
constexpr int kFusedTileBits = 13;
constexpr int kFusedWork = (1 << kFusedTileBits) + ((1 << kFusedTileBits) >> 5) * 4;          // float2 elements, padded

template <int BITS1>
__global__ void __launch_bounds__(kFusedThreads, 1) baseline_fused_cols_kernel(FusedColsArgs a, int tiles, int total) {
    extern __shared__ __align__(128) float2 sfft[];
    constexpr int TILE_BITS = kFusedTileBits, THREADS = kFusedThreads;
    constexpr int TB = TILE_BITS - BITS1, T = 1 << TB, N1 = 1 << BITS1;
    constexpr int SUBB = BITS1 - 4;               // the middle pass is the radix-16 pass over whole columns
    constexpr int ITEMS = (1 << TILE_BITS) >> 4, ITERS = ITEMS / THREADS;
    static_assert(BITS1 >= 5 && BITS1 <= 11 && ITEMS % THREADS == 0, "fused column kernel: unsupported tile");
    constexpr uint32_t kTileBytes = (1u << TILE_BITS) * 8u, kHaloBytes = N1 * 4u;
    float2* sN = sfft + kFusedWork;                                   // [N1][T] float2 = the tile's n block
    float2* sU = sN + (1 << TILE_BITS);
    float* sH = reinterpret_cast<float*>(sU + (1 << TILE_BITS));      // [3][N1]: left n, left u, right u
    uint64_t* bar = reinterpret_cast<uint64_t*>(sH + 3 * N1);
    const size_t pcells = (size_t)N1 << (kRowBits + 1);               // cells per IC
    const size_t hplane = (size_t)tiles << BITS1;
    const int tid = threadIdx.x;
    const float rdx2 = fv_reciprocal(a.dx2);
    if (tid == 0) {
        mbar_init(bar, 1);
        mbar_fence_init();
    }
    __syncthreads();
    // thread 0: bulk copies of work item `it` (tile-major n, u blocks; neighbour rows where they are contiguous)
    auto issue = [&](int it) {
        const int b = it / tiles, tile = it - b * tiles;
        const uint32_t bytes = 2 * kTileBytes + (tile > 0 ? 2 * kHaloBytes : 0) + (tile < tiles - 1 ? kHaloBytes : 0);
        mbar_arrive_expect_tx(bar, bytes);
        bulk_g2s(sN, a.Pn_in + (size_t)b * pcells + ((size_t)tile << (TILE_BITS + 1)), kTileBytes, bar);
        bulk_g2s(sU, a.Pu_in + (size_t)b * pcells + ((size_t)tile << (TILE_BITS + 1)), kTileBytes, bar);
        const float* Hin = a.H_in + (size_t)b * 3 * hplane;
        if (tile > 0) {
            bulk_g2s(sH, Hin + ((size_t)(tile - 1) << BITS1), kHaloBytes, bar);
            bulk_g2s(sH + N1, Hin + hplane + ((size_t)(tile - 1) << BITS1), kHaloBytes, bar);
        }
        if (tile < tiles - 1) bulk_g2s(sH + 2 * N1, Hin + 2 * hplane + ((size_t)(tile + 1) << BITS1), kHaloBytes, bar);
    };
    if (tid == 0 && (int)blockIdx.x < total) issue(blockIdx.x);
    uint32_t parity = 0;
    for (int it = blockIdx.x; it < total; it += gridDim.x, parity ^= 1) {
        const int b = it / tiles, tile = it - b * tiles;
        const int j2_0 = tile << TB;
        float2* Y = a.Y + ((size_t)b << (BITS1 + kRowBits));
        const float* Hin = a.H_in + (size_t)b * 3 * hplane;
        // the grid's first / last tile: its outer neighbours are the far tile's rows shifted by one matrix row
        // (periodic wrap), which is not 16-byte aligned -> ordinary loads
        if (tile == 0) {
            const float* hn = Hin + ((size_t)(tiles - 1) << BITS1);
            for (int j = tid; j < N1; j += THREADS) {
                const int jl = (j == 0) ? N1 - 1 : j - 1;
                sH[j] = __ldg(hn + jl);
                sH[N1 + j] = __ldg(hn + hplane + jl);
            }
        }
        if (tile == tiles - 1) {
            const float* hu = Hin + 2 * hplane;
            for (int j = tid; j < N1; j += THREADS) sH[2 * N1 + j] = __ldg(hu + ((j == N1 - 1) ? 0 : j + 1));
        }
        // ---- pass C of the previous step, all but its last stage set ----
        columns_inverse_to_smem<BITS1, TILE_BITS, first_inverse_lb(BITS1), THREADS>(Y, sfft, j2_0);
        mbar_wait(bar, parity);
        float2* Qn = reinterpret_cast<float2*>(a.Pn_out + (size_t)b * pcells) + ((size_t)tile << TILE_BITS);
        float2* Qu = reinterpret_cast<float2*>(a.Pu_out + (size_t)b * pcells) + ((size_t)tile << TILE_BITS);
        float* Hout = a.H_out + (size_t)b * 3 * hplane + ((size_t)tile << BITS1);
#pragma unroll
        for (int rep = 0; rep < ITERS; ++rep) {
            const int w = tid + rep * THREADS;
            const int t = w & (T - 1), base = w >> TB;                    // rows base + m*SUB, column t
            float2 v[16];
#pragma unroll
            for (int m = 0; m < 16; ++m) v[m] = sfft[cpad(((base + (m << SUBB)) << TB) + t)];
            TwiddleSet<16> pw;
            pw.init(unit_root((base << kRowBits) + j2_0 + t, BITS1 + kRowBits, +1.f));
            pw.apply(v);
            DitStages<16>::run(v, +1.f);                                   // v[m] = (E_2j, E_2j+1) at (row base + m*SUB, column t)
            // ---- finite-volume update of the two cells of every position (src/baseline_solver.py:84-94) ----
#pragma unroll
            for (int m = 0; m < 16; ++m) {
                const int j1 = base + (m << SUBB);
                const int e = (j1 << TB) + t;
                const float2 n2 = sN[e], u2 = sU[e];
                float nl, ul, ur;
                if (t == 0) {
                    nl = sH[j1];
                    ul = sH[N1 + j1];
                } else {
                    nl = sN[e - 1].y;
                    ul = sU[e - 1].y;
                }
                ur = (t == T - 1) ? sH[2 * N1 + j1] : sU[e + 1].x;
                float2 nn, un;
                fv_pair(nl, n2, ul, u2, ur, v[m], a.c, a.dt, a.nu, a.dx2, rdx2, nn, un);
                Qn[e] = nn;
                Qu[e] = un;
                if (t == T - 1) {
                    Hout[j1] = nn.y;
                    Hout[hplane + j1] = un.y;
                }
                if (t == 0) Hout[2 * hplane + j1] = un.x;
                if (a.nat_out != nullptr) {
                    float2* so = reinterpret_cast<float2*>(a.nat_out + (size_t)b * 3 * pcells) + ((size_t)j1 << kRowBits) + j2_0 + t;
                    so[0] = nn;
                    so[pcells >> 1] = un;
                }
                v[m] = __fadd2_rn(nn, make_float2(-1.0f, -1.0f));                 // rho' = n' - n0
            }
            // ---- first stage set of pass A of the next step ----
            DifStages<16>::run(v, -1.f);
#pragma unroll
            for (int q = 0; q < 16; ++q) pw.pw[q].y = -pw.pw[q].y;         // W^-1 powers = conjugates
            pw.apply(v);
#pragma unroll
            for (int m = 0; m < 16; ++m) sfft[cpad(((base + (m << SUBB)) << TB) + t)] = v[m];
        }
        __syncthreads();                              // n, u, neighbour rows are consumed; the work tile is complete
        if (it + (int)gridDim.x < total) {
            const int nit = it + gridDim.x;
            if (tid == 0) issue(nit);
            // the next tile's spectrum pieces (one 32-byte sector per matrix row) on their way into the L2
            const int nb = nit / tiles, ntile = nit - nb * tiles;
            const float2* Yn = a.Y + ((size_t)nb << (BITS1 + kRowBits)) + (ntile << TB);
            for (int row = tid; row < N1; row += THREADS)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(Yn + ((size_t)row << kRowBits)));
        }
        columns_forward<BITS1, TILE_BITS, SUBB, THREADS>(nullptr, Y, sfft, j2_0);
        __syncthreads();                              // the work tile is free for the next tile's inverse stages
    }
}

// natural [B][3][nx] -> tile-major n, u and the halo side arrays (once per fused rollout).  TB = log2(columns per tile).
__global__ void __launch_bounds__(256) baseline_to_tiles_kernel(const float* __restrict__ state, float* __restrict__ Pn,
                                                                float* __restrict__ Pu, float* __restrict__ H, int bits1,
                                                                int tb, int B) {
    const size_t M = (size_t)1 << (bits1 + kRowBits);
    const int T = 1 << tb, tiles = kRowLen >> tb;
    const size_t hplane = (size_t)tiles << bits1;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < M * B; idx += (size_t)gridDim.x * blockDim.x) {
        const int b = (int)(idx >> (bits1 + kRowBits));
        const size_t j = idx & (M - 1);
        const int j1 = (int)(j >> kRowBits), j2 = (int)(j & (kRowLen - 1)), tile = j2 >> tb, t = j2 & (T - 1);
        const float2 n2 = reinterpret_cast<const float2*>(state + (size_t)b * 6 * M)[j];
        const float2 u2 = reinterpret_cast<const float2*>(state + (size_t)b * 6 * M + 2 * M)[j];
        const size_t e = (size_t)b * M + (((size_t)tile << bits1) + j1) * T + t;
        reinterpret_cast<float2*>(Pn)[e] = n2;
        reinterpret_cast<float2*>(Pu)[e] = u2;
        float* Hb = H + (size_t)b * 3 * hplane;
        const size_t hme = ((size_t)tile << bits1) + j1;
        if (t == T - 1) {
            Hb[hme] = n2.y;
            Hb[hplane + hme] = u2.y;
        }
        if (t == 0) Hb[2 * hplane + hme] = u2.x;
    }
}

// pass B of the distributed solve: two INDEPENDENT rows per CTA (positions 2 bx, 2 bx + 1; they share every
// twiddle set), forward row stages, diagonal multiplier, inverse row stages, in place.  grid = (N1/2, P).
__global__ void __launch_bounds__(kFftStepThreads, FLUXGNN_FFT_ROW_CTAS) poisson_fft_rows_diag_kernel(
    float2* __restrict__ Y, int bits1, float scale, int G, int kr) {
    extern __shared__ float2 sfft[];
    const int tid = threadIdx.x;
    const long long Ml = 1LL << (bits1 + kRowBits), nx = Ml * G;
    float2* base = Y + ((size_t)blockIdx.y << (bits1 + kRowBits));
    float2* rowa = base + ((size_t)(2 * blockIdx.x) << kRowBits);
    float2* rowb = rowa + kRowLen;
    const int k1a = bitrev(2 * blockIdx.x, bits1), k1b = bitrev(2 * blockIdx.x + 1, bits1);
    float2* sa = sfft;
    float2* sb = sfft + kRowPad;
    float2 va[16], vb[16], pw[16];
    const uint64_t pol = make_policy<Hint::kLast>();
#pragma unroll
    for (int m = 0; m < 16; ++m) va[m] = gload<Hint::kLast>(rowa + tid + 256 * m, pol);
#pragma unroll
    for (int m = 0; m < 16; ++m) vb[m] = gload<Hint::kLast>(rowb + tid + 256 * m, pol);
    twiddle_powers<16>(unit_root(tid, kRowBits, -1.f), pw);
    row_butterfly<false>(va, pw, true);
    row_butterfly<false>(vb, pw, true);
#pragma unroll
    for (int m = 0; m < 16; ++m) sa[rpad(tid + 256 * m)] = va[m];
#pragma unroll
    for (int m = 0; m < 16; ++m) sb[rpad(tid + 256 * m)] = vb[m];
    __syncthreads();
    const int e2 = (tid >> 4) * 256 + (tid & 15);
    twiddle_powers<16>(unit_root(tid & 15, 8, -1.f), pw);
#pragma unroll
    for (int m = 0; m < 16; ++m) va[m] = sa[rpad(e2 + 16 * m)];
#pragma unroll
    for (int m = 0; m < 16; ++m) vb[m] = sb[rpad(e2 + 16 * m)];
    row_butterfly<false>(va, pw, true);
    row_butterfly<false>(vb, pw, true);
#pragma unroll
    for (int m = 0; m < 16; ++m) sa[rpad(e2 + 16 * m)] = va[m];
#pragma unroll
    for (int m = 0; m < 16; ++m) sb[rpad(e2 + 16 * m)] = vb[m];
    __syncthreads();
    // forward pass 3 (blocks of 16), multiplier, inverse pass 3: block tid of both rows stays in registers
    const int e3 = tid * 16;
#pragma unroll
    for (int m = 0; m < 16; ++m) va[m] = sa[rpad(e3 + m)];
#pragma unroll
    for (int m = 0; m < 16; ++m) vb[m] = sb[rpad(e3 + m)];
    row_butterfly<false>(va, pw, false);
    row_butterfly<false>(vb, pw, false);
    // register m holds local bin kl = k1 + N1 (bitrev8(tid) + 256 bitrev4(m)) = kl0 + bitrev4(m) Ml/16
    const long long kl0a = (long long)k1a + ((long long)bitrev(tid, 8) << bits1);
    const long long kl0b = (long long)k1b + ((long long)bitrev(tid, 8) << bits1);
#pragma unroll
    for (int m = 0; m < 16; ++m) {
        const long long step = (long long)brev_bits(m, 4) * (Ml >> 4);
        va[m] = spectral_diag(va[m], (long long)kr + (long long)G * (kl0a + step), nx, scale);
        vb[m] = spectral_diag(vb[m], (long long)kr + (long long)G * (kl0b + step), nx, scale);
    }
    row_butterfly<true>(va, pw, false);
    row_butterfly<true>(vb, pw, false);
#pragma unroll
    for (int m = 0; m < 16; ++m) sa[rpad(e3 + m)] = va[m];
#pragma unroll
    for (int m = 0; m < 16; ++m) sb[rpad(e3 + m)] = vb[m];
    __syncthreads();
    twiddle_powers<16>(unit_root(tid & 15, 8, +1.f), pw);
#pragma unroll
    for (int m = 0; m < 16; ++m) va[m] = sa[rpad(e2 + 16 * m)];
#pragma unroll
    for (int m = 0; m < 16; ++m) vb[m] = sb[rpad(e2 + 16 * m)];
    row_butterfly<true>(va, pw, true);
    row_butterfly<true>(vb, pw, true);
#pragma unroll
    for (int m = 0; m < 16; ++m) sa[rpad(e2 + 16 * m)] = va[m];
#pragma unroll
    for (int m = 0; m < 16; ++m) sb[rpad(e2 + 16 * m)] = vb[m];
    __syncthreads();
    twiddle_powers<16>(unit_root(tid, kRowBits, +1.f), pw);
#pragma unroll
    for (int m = 0; m < 16; ++m) va[m] = sa[rpad(tid + 256 * m)];
#pragma unroll
    for (int m = 0; m < 16; ++m) vb[m] = sb[rpad(tid + 256 * m)];
    row_butterfly<true>(va, pw, true);
    row_butterfly<true>(vb, pw, true);
#pragma unroll
    for (int m = 0; m < 16; ++m) gstore<Hint::kLast>(rowa + tid + 256 * m, va[m], pol);
#pragma unroll
    for (int m = 0; m < 16; ++m) gstore<Hint::kLast>(rowb + tid + 256 * m, vb[m], pol);
}

// ---------------------------------------------------------------------------
// Distributed field solve (one grid of nx cells split over G ranks, S = nx/G cells each).
// Two density rows (ICs 2p, 2p+1) share one COMPLEX signal z = rho_a + i rho_b, so the multiplier i/k is diagonal
// (no k <-> nx-k pairing across ranks) and the solve returns E_a + i E_b.  The nx-point transform is decimated in
// frequency over the rank index: element (jr, jl) = jr S + jl, bin k = kr + G kl,
//   y[kr][jl]  = W_nx^(jl kr) sum_jr W_G^(jr kr) z[jr][jl]        poisson_rank_dft_kernel (between two all-to-alls)
//   Z[kr+G kl] = FFT_S(y[kr])[kl]                                  local: passes A, B, C above with the diagonal multiplier
// and the inverse runs the same steps backwards.  Every exchanged buffer is a flat array cut into G equal chunks.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) poisson_pack_pairs_kernel(const float* __restrict__ n, long long ic_stride,
                                                                 float2* __restrict__ z, int B, int S) {
    const long long total = (long long)((B + 1) / 2) * S;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int p = (int)(idx / S), j = (int)(idx - (long long)p * S);
        const float a = __fsub_rn(__ldg(n + (size_t)(2 * p) * ic_stride + j), 1.0f);
        const float b = (2 * p + 1 < B) ? __fsub_rn(__ldg(n + (size_t)(2 * p + 1) * ic_stride + j), 1.0f) : 0.f;
        z[idx] = make_float2(a, b);
    }
}

__global__ void __launch_bounds__(256) poisson_unpack_pairs_kernel(const float2* __restrict__ e, float* __restrict__ E,
                                                                   long long ic_stride, int B, int S) {
    const long long total = (long long)((B + 1) / 2) * S;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int p = (int)(idx / S), j = (int)(idx - (long long)p * S);
        const float2 v = e[idx];
        E[(size_t)(2 * p) * ic_stride + j] = v.x;
        if (2 * p + 1 < B) E[(size_t)(2 * p + 1) * ic_stride + j] = v.y;
    }
}

// in[q][e], out[r][e], e < chunk; flat index of element e = flat0 + e, jl = flat index mod S.
//   forward: out[kr][e] = W_nx^-(jl kr) sum_q W_G^-(q kr) in[q][e]
//   inverse: out[jr][e] = (1/G) sum_kr W_G^+(jr kr) W_nx^+(jl kr) in[kr][e]
template <int G>
__global__ void __launch_bounds__(256) poisson_rank_dft_kernel(const float2* __restrict__ in, float2* __restrict__ out,
                                                               long long chunk, long long flat0, int S, long long nx,
                                                               int inverse) {
    const float sign = inverse ? 1.f : -1.f;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < chunk; e += (long long)gridDim.x * blockDim.x) {
        const long long jl = (flat0 + e) & (long long)(S - 1);
        float2 v[G], pw[G];
#pragma unroll
        for (int q = 0; q < G; ++q) v[q] = in[(size_t)q * chunk + e];
        twiddle_powers<G>(big_twiddle(jl, nx, sign), pw);             // pw[r] = W_nx^(sign jl r)
        if (inverse) {
#pragma unroll
            for (int q = 1; q < G; ++q) v[q] = pmul(v[q], pw[q]);
        }
#pragma unroll
        for (int r = 0; r < G; ++r) {
            float2 acc = v[0];
#pragma unroll
            for (int q = 1; q < G; ++q) {
                // W_G^(sign q r): eighth turns at most for G <= 8, sixteenth turns for G = 16
                constexpr int kTurn = 16 / G;
                float2 t;
                switch (((q * r) & (G - 1)) * kTurn) {
                    case 0: t = v[q]; break;
                    case 1: t = mul_root16<1>(v[q], sign); break;
                    case 2: t = mul_root16<2>(v[q], sign); break;
                    case 3: t = mul_root16<3>(v[q], sign); break;
                    case 4: t = mul_root16<4>(v[q], sign); break;
                    case 5: t = mul_root16<5>(v[q], sign); break;
                    case 6: t = mul_root16<6>(v[q], sign); break;
                    case 7: t = mul_root16<7>(v[q], sign); break;
                    case 8: t = mul_root16<8>(v[q], sign); break;
                    case 9: t = mul_root16<9>(v[q], sign); break;
                    case 10: t = mul_root16<10>(v[q], sign); break;
                    case 11: t = mul_root16<11>(v[q], sign); break;
                    case 12: t = mul_root16<12>(v[q], sign); break;
                    case 13: t = mul_root16<13>(v[q], sign); break;
                    case 14: t = mul_root16<14>(v[q], sign); break;
                    default: t = mul_root16<15>(v[q], sign); break;
                }
                acc = __fadd2_rn(acc, t);
            }
            if (inverse) acc = make_float2(acc.x * (1.0f / G), acc.y * (1.0f / G));
            else if (r > 0) acc = pmul(acc, pw[r]);
            out[(size_t)r * chunk + e] = acc;
        }
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

// One column pass of the four-step solve: inverse = 0: density n -> Y (pass A); 1: Y -> field E (pass C).
// Tensor map of a [B] x [N1 rows] x [8192 floats] array (row pitch 32 KiB, IC stride given in floats), box = one
// 256-row piece of a column tile; the driver entry point is looked up once (no link-time dependency on libcuda).
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
        else
            (void)cudaGetLastError();
    }
    return fn;
}

static int column_tile_map(CUtensorMap* map, const void* base, long long ic_stride_floats, int B, int bits1, bool swizzle) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (fn == nullptr) return set_error(FLUXGNN_ECUDA, "cuTensorMapEncodeTiled is not available from this driver");
    const int T = 1 << (kTmaTileBits - bits1);
    const cuuint64_t dims[3] = {(cuuint64_t)2 * kRowLen, (cuuint64_t)1 << bits1, (cuuint64_t)B};
    const cuuint64_t strides[2] = {(cuuint64_t)2 * kRowLen * sizeof(float), (cuuint64_t)ic_stride_floats * sizeof(float)};
    const cuuint32_t box[3] = {(cuuint32_t)(2 * T), (cuuint32_t)kTmaBoxRows, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult rc = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(base), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                           CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) return set_error(FLUXGNN_ECUDA, "cuTensorMapEncodeTiled failed (%d)", (int)rc);
    return FLUXGNN_OK;
}

// 0 = plain / cp.async kernels, 1 = TMA without swizzle, 2 = TMA with the 128-byte swizzle (default for long grids)
static int tma_cols_mode(int bits1, long long tiles_total) {
    // Measured (profiles/r2_field_solve_variants.md): 5 % / 3 % faster than the plain kernels for 128- / 64-byte tile rows
    // (N1 = 512 / 1024), 2 % slower for 32-byte rows (N1 = 2048, where the dense tile has 4-way bank conflicts in the
    // radix-8 pass), so the default takes the TMA kernels for N1 <= 1024 only.  FLUXGNN_FFT_TMA = 0 / 1 / 2 forces
    // plain / dense / swizzled (test hook and experiment switch).
    const char* env = getenv("FLUXGNN_FFT_TMA");
    const bool forced = env != nullptr && env[0] >= '0' && env[0] <= '2';
    const int want = forced ? env[0] - '0' : (bits1 <= 10 ? 2 : 0);
    if (want == 0 || bits1 < 9 || bits1 > 11 || tiles_total < 2 * 148 || encode_tiled_fn() == nullptr) return 0;
    // SWIZZLE_128B needs 128-byte tile rows (16 points, N1 = 512); narrower boxes fault with it, so they stay dense
    return (want == 2 && bits1 != 9) ? 1 : want;
}

template <bool kInv, bool kSub1>
static int launch_cols_tma(const float* in, long long in_stride_floats, float* out, long long out_stride_floats, int B,
                           int bits1, int mode, cudaStream_t stream) {
    CUtensorMap map_in, map_out;
    int rc = column_tile_map(&map_in, in, in_stride_floats, B, bits1, mode == 2);
    if (rc != FLUXGNN_OK) return rc;
    rc = column_tile_map(&map_out, out, out_stride_floats, B, bits1, mode == 2);
    if (rc != FLUXGNN_OK) return rc;
    const int tiles = kRowLen >> (kTmaTileBits - bits1), total = tiles * B;
    const size_t smem = (size_t)3 * (8u << kTmaTileBits) + 64 + 1024;
    int dev = 0, sms = 0;
    FLUXGNN_CUDA_OK(cudaGetDevice(&dev));
    FLUXGNN_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int grid = total < sms ? total : sms;
#define FLUXGNN_TMA_ONE(BITS1, LAYOUT)                                                                                     \
    {                                                                                                                      \
        FLUXGNN_CUDA_OK(cudaFuncSetAttribute(poisson_fft_cols_tma_kernel<BITS1, kInv, kSub1, LAYOUT>,                       \
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                      \
        poisson_fft_cols_tma_kernel<BITS1, kInv, kSub1, LAYOUT><<<grid, kTmaThreads, smem, stream>>>(map_in, map_out, tiles, total); \
    }
#define FLUXGNN_TMA(BITS1)                                          \
    case BITS1:                                                     \
        if (mode == 2) FLUXGNN_TMA_ONE(BITS1, 2) else FLUXGNN_TMA_ONE(BITS1, 1) \
        break;
    switch (bits1) {
        FLUXGNN_TMA(9) FLUXGNN_TMA(10) FLUXGNN_TMA(11)
        default: return set_error(FLUXGNN_EUNSUP, "TMA column pass: unsupported column length 2^%d", bits1);
    }
#undef FLUXGNN_TMA
#undef FLUXGNN_TMA_ONE
    FLUXGNN_CUDA_OK(cudaGetLastError());
    count_launch();
    return FLUXGNN_OK;
}

// Staged (double-buffered, persistent) column pass: in / out are [B] matrices of N1 x 4096 points with the given IC
// strides (float2).  Returns FLUXGNN_EUNSUP when this shape keeps the plain kernels.
static bool staged_cols_ok(int bits1, long long tiles_total) {
    // opt-in (FLUXGNN_FFT_STAGING=1): measured SLOWER than the plain kernels (50 / 45 us against 40 / 36 us per pass at
    // 2^24 cells): the passes are bound by load/store-unit wavefronts (one per 32-byte piece), not by load latency
    const char* on = getenv("FLUXGNN_FFT_STAGING");
    if (on == nullptr || on[0] != '1') return false;
    return bits1 >= 5 && bits1 <= 11 && tiles_total >= 2 * 148;        // 2^13-point tiles; enough tiles to pipeline
}

template <bool kInv, bool kSub1>
static int launch_cols_staged(const float2* in, long long in_stride, float2* out, long long out_stride, int B, int bits1,
                              cudaStream_t stream) {
    const int tiles = kRowLen >> (kStagedTileBits - bits1), total = tiles * B;
    const size_t smem = (size_t)2 * kStagedTile * sizeof(float2);
    int dev = 0, sms = 0;
    FLUXGNN_CUDA_OK(cudaGetDevice(&dev));
    FLUXGNN_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int grid = total < sms ? total : sms;
#define FLUXGNN_STAGED(BITS1)                                                                                              \
    case BITS1:                                                                                                            \
        FLUXGNN_CUDA_OK(cudaFuncSetAttribute(poisson_fft_cols_staged_kernel<BITS1, kInv, kSub1>,                            \
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                      \
        poisson_fft_cols_staged_kernel<BITS1, kInv, kSub1><<<grid, kStagedThreads, smem, stream>>>(in, in_stride, out,      \
                                                                                                out_stride, tiles, total); \
        break;
    switch (bits1) {
        FLUXGNN_STAGED(5) FLUXGNN_STAGED(6) FLUXGNN_STAGED(7) FLUXGNN_STAGED(8) FLUXGNN_STAGED(9) FLUXGNN_STAGED(10)
        FLUXGNN_STAGED(11)
        default: return set_error(FLUXGNN_EUNSUP, "staged column pass: unsupported column length 2^%d", bits1);
    }
#undef FLUXGNN_STAGED
    FLUXGNN_CUDA_OK(cudaGetLastError());
    count_launch();
    return FLUXGNN_OK;
}

int launch_poisson_fft_cols(const float* n, long long n_stride, float2* Y, float* E, long long e_stride, int B, int nx,
                            int inverse, cudaStream_t stream) {
    const int bits1 = ilog2_exact(nx) - 1 - kRowBits;
    if (bits1 >= 9 && bits1 <= 11) {
        const int mode = tma_cols_mode(bits1, (long long)B * (kRowLen >> (kTmaTileBits - bits1)));
        if (mode != 0) {
            const long long ystride = (long long)2 << (bits1 + kRowBits);                    // floats per IC of Y
            if (!inverse) return launch_cols_tma<false, true>(n, n_stride, reinterpret_cast<float*>(Y), ystride, B, bits1, mode, stream);
            return launch_cols_tma<true, true>(reinterpret_cast<const float*>(Y), ystride, E, e_stride, B, bits1, mode, stream);
        }
    }
    if (bits1 <= 11 && staged_cols_ok(bits1, (long long)B * (kRowLen >> (kStagedTileBits - bits1)))) {
        const long long ystride = (long long)1 << (bits1 + kRowBits);
        if (!inverse)
            return launch_cols_staged<false, true>(reinterpret_cast<const float2*>(n), n_stride / 2, Y, ystride, B, bits1, stream);
        return launch_cols_staged<true, true>(Y, ystride, reinterpret_cast<float2*>(E), e_stride / 2, B, bits1, stream);
    }
    const int tile_bits = column_tile_bits(bits1);
    const int T = 1 << (tile_bits - bits1);
    const size_t tile = (size_t)1 << tile_bits;
    const size_t smem = (tile + (tile >> 5) * 4) * sizeof(float2);
    dim3 gcol((unsigned)(kRowLen / T), (unsigned)B);
#define FLUXGNN_FFT_COLS(BITS1, WHICH, ...)                                                                          \
    case BITS1:                                                                                                     \
        FLUXGNN_CUDA_OK(cudaFuncSetAttribute(WHICH<BITS1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        WHICH<BITS1><<<gcol, column_threads(BITS1), smem, stream>>>(__VA_ARGS__);                                          \
        break;
#define FLUXGNN_FFT_COLS_ALL(WHICH, ...)                                                       \
    switch (bits1) {                                                                          \
        FLUXGNN_FFT_COLS(3, WHICH, __VA_ARGS__) FLUXGNN_FFT_COLS(4, WHICH, __VA_ARGS__)       \
        FLUXGNN_FFT_COLS(5, WHICH, __VA_ARGS__) FLUXGNN_FFT_COLS(6, WHICH, __VA_ARGS__)       \
        FLUXGNN_FFT_COLS(7, WHICH, __VA_ARGS__) FLUXGNN_FFT_COLS(8, WHICH, __VA_ARGS__)       \
        FLUXGNN_FFT_COLS(9, WHICH, __VA_ARGS__) FLUXGNN_FFT_COLS(10, WHICH, __VA_ARGS__)      \
        FLUXGNN_FFT_COLS(11, WHICH, __VA_ARGS__) FLUXGNN_FFT_COLS(12, WHICH, __VA_ARGS__)     \
        default: return set_error(FLUXGNN_EUNSUP, "four-step FFT: unsupported column length 2^%d", bits1); \
    }
    if (!inverse) {
        FLUXGNN_FFT_COLS_ALL(poisson_fft_cols_fwd_kernel, n, n_stride, Y)
    } else {
        FLUXGNN_FFT_COLS_ALL(poisson_fft_cols_inv_kernel, Y, E, e_stride)
    }
#undef FLUXGNN_FFT_COLS_ALL
#undef FLUXGNN_FFT_COLS
    FLUXGNN_CUDA_OK(cudaGetLastError());
    count_launch();
    return FLUXGNN_OK;
}

// Pass B of the four-step solve (row pairs, spectral step), in place on Y.
int launch_poisson_fft_rows(float2* Y, int B, int nx, double length, cudaStream_t stream) {
    const int mbits = ilog2_exact(nx) - 1, bits1 = mbits - kRowBits;
    const size_t smem_row = (size_t)2 * kRowPad * sizeof(float2);
    const float scale = (float)(length / (6.283185307179586476925 * (double)(1LL << mbits)));
    dim3 grow((unsigned)((1 << bits1) / 2 + 1), (unsigned)B);
    FLUXGNN_CUDA_OK(cudaFuncSetAttribute(poisson_fft_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_row));
    poisson_fft_rows_kernel<<<grow, kFftStepThreads, smem_row, stream>>>(Y, bits1, scale);
    FLUXGNN_CUDA_OK(cudaGetLastError());
    count_launch();
    return FLUXGNN_OK;
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
        FLUXGNN_CUDA_OK(cudaFuncSetAttribute(poisson_fft_small_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)smem));
        const int threads = M / 16 < kFftThreads ? (M / 16 < 64 ? 64 : M / 16) : kFftThreads;
        poisson_fft_small_kernel<false><<<B, threads, smem, stream>>>(n, n_stride, E, e_stride, mbits, length, 1, 0);
        FLUXGNN_CUDA_OK(cudaGetLastError());
        count_launch();
        return FLUXGNN_OK;
    }
    if (workspace == nullptr)
        return set_error(FLUXGNN_EINVAL, "field solve for nx=%d needs fluxgnn_poisson_workspace_bytes() of scratch", nx);
    if (B > 65535) return set_error(FLUXGNN_EUNSUP, "field solve for nx=%d handles at most 65535 ICs per call", nx);
    float2* Y = (float2*)workspace;
    int rc = launch_poisson_fft_cols(n, n_stride, Y, nullptr, 0, B, nx, 0, stream);
    if (rc != FLUXGNN_OK) return rc;
    rc = launch_poisson_fft_rows(Y, B, nx, length, stream);
    if (rc != FLUXGNN_OK) return rc;
    return launch_poisson_fft_cols(nullptr, 0, Y, E, e_stride, B, nx, 1, stream);
}

// ---- fused classical step: host side --------------------------------------------------------------------
bool baseline_fused_supported(int nx, float dx2) {
    const int bits = ilog2_exact(nx);
    if (bits < 0) return false;
    // the fused kernel divides by dx^2 with the FMA sequence of fv_div() only (see fv_reciprocal in field_kernels.cuh)
    unsigned ubits;
    memcpy(&ubits, &dx2, sizeof(ubits));
    if (!(dx2 > 0.f && dx2 < 1.f) || (ubits & 0x7fffffu) == 0x7fffffu || (ubits >> 23) == 0u) return false;
    const int bits1 = bits - 1 - kRowBits;
    return bits1 >= 8 && bits1 <= 11;                 // column tiles of 2^13 points, 4..32 columns; nx = 2^21 .. 2^24
}

static void fused_geometry(int nx, int* bits1, int* tb, size_t* cells, size_t* hfloats) {
    const int bits = ilog2_exact(nx);
    *bits1 = bits - 1 - kRowBits;
    *tb = kFusedTileBits - *bits1;
    *cells = (size_t)nx;
    *hfloats = (size_t)3 * (kRowLen >> *tb) << *bits1;     // 3 planes x tiles x N1
}

size_t baseline_fused_workspace_floats(int B, int nx) {
    if (!baseline_fused_supported(nx, 0.5f)) return 0;
    int bits1, tb;
    size_t cells, hfloats;
    fused_geometry(nx, &bits1, &tb, &cells, &hfloats);
    return (size_t)B * 2 * (2 * cells + hfloats);
}

// fused_ws = [slot 0: Pn | Pu | H][slot 1: Pn | Pu | H]
static float* fused_slot(float* ws, int slot, int B, size_t cells, size_t hfloats) {
    return ws + (size_t)slot * B * (2 * cells + hfloats);
}

int launch_baseline_to_tiles(const float* state, float* fused_ws, int slot, int B, int nx, cudaStream_t stream) {
    int bits1, tb;
    size_t cells, hfloats;
    fused_geometry(nx, &bits1, &tb, &cells, &hfloats);
    float* base = fused_slot(fused_ws, slot, B, cells, hfloats);
    baseline_to_tiles_kernel<<<148 * 8, 256, 0, stream>>>(state, base, base + (size_t)B * cells, base + (size_t)2 * B * cells,
                                                          bits1, tb, B);
    FLUXGNN_CUDA_OK(cudaGetLastError());
    count_launch();
    return FLUXGNN_OK;
}

int launch_baseline_fused_cols(float2* Y, float* fused_ws, int slot_in, float* nat_out, int B, int nx,
                               float c, float dt, float nu, float dx2, cudaStream_t stream) {
    int bits1, tb;
    size_t cells, hfloats;
    fused_geometry(nx, &bits1, &tb, &cells, &hfloats);
    const float* in = fused_slot(fused_ws, slot_in, B, cells, hfloats);
    float* out = fused_slot(fused_ws, 1 - slot_in, B, cells, hfloats);
    FusedColsArgs a;
    a.Y = Y;
    a.Pn_in = in;
    a.Pu_in = in + (size_t)B * cells;
    a.H_in = in + (size_t)2 * B * cells;
    a.Pn_out = out;
    a.Pu_out = out + (size_t)B * cells;
    a.H_out = out + (size_t)2 * B * cells;
    a.nat_out = nat_out;
    a.c = c; a.dt = dt; a.nu = nu; a.dx2 = dx2;
    const int tiles = kRowLen >> tb, total = tiles * B;
    const size_t smem = (size_t)kFusedWork * sizeof(float2) + 2 * ((size_t)sizeof(float2) << kFusedTileBits) +
                        ((size_t)3 << bits1) * sizeof(float) + 16;
    int dev = 0, sms = 0;
    FLUXGNN_CUDA_OK(cudaGetDevice(&dev));
    FLUXGNN_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int grid = total < sms ? total : sms;
#define FLUXGNN_FUSED(BITS1)                                                                                                  \
    case BITS1:                                                                                                               \
        FLUXGNN_CUDA_OK(cudaFuncSetAttribute(baseline_fused_cols_kernel<BITS1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        baseline_fused_cols_kernel<BITS1><<<grid, kFusedThreads, smem, stream>>>(a, tiles, total);                            \
        break;
    switch (bits1) {
        FLUXGNN_FUSED(8) FLUXGNN_FUSED(9) FLUXGNN_FUSED(10) FLUXGNN_FUSED(11)
        default: return set_error(FLUXGNN_EUNSUP, "fused classical step: unsupported column length 2^%d", bits1);
    }
#undef FLUXGNN_FUSED
    FLUXGNN_CUDA_OK(cudaGetLastError());
    count_launch();
    return FLUXGNN_OK;
}

// ---- distributed solve: host side ------------------------------------------------------------------------
static int dist_shape_ok(int S, int G) {
    const int sbits = ilog2_exact(S), gbits = ilog2_exact(G);
    if (sbits < 8 || sbits > kFftMaxBits - 1 || gbits < 0 || G > 16 || sbits + gbits > 27)
        return set_error(FLUXGNN_EUNSUP, "distributed field solve: slab of %d cells over %d ranks (need powers of two, "
                                         "slab >= 256 cells, <= 16 ranks, grid <= 2^27 cells)", S, G);
    return FLUXGNN_OK;
}

int launch_poisson_dist_pack(const float* n, long long ic_stride, int B, int S, float2* z, int unpack, float* E,
                             cudaStream_t stream) {
    const long long total = (long long)((B + 1) / 2) * S;
    const unsigned blocks = (unsigned)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
    if (unpack) poisson_unpack_pairs_kernel<<<blocks, 256, 0, stream>>>(z, E, ic_stride, B, S);
    else poisson_pack_pairs_kernel<<<blocks, 256, 0, stream>>>(n, ic_stride, z, B, S);
    FLUXGNN_CUDA_OK(cudaGetLastError());
    count_launch();
    return FLUXGNN_OK;
}

int launch_poisson_rank_dft(const float2* in, float2* out, int G, long long chunk, long long flat0, int S, int inverse,
                            cudaStream_t stream) {
    int rc = dist_shape_ok(S, G);
    if (rc != FLUXGNN_OK) return rc;
    const long long nx = (long long)S * G;
    const unsigned blocks = (unsigned)((chunk + 255) / 256 < 148 * 16 ? (chunk + 255) / 256 : 148 * 16);
    switch (G) {
        case 1: poisson_rank_dft_kernel<1><<<blocks, 256, 0, stream>>>(in, out, chunk, flat0, S, nx, inverse); break;
        case 2: poisson_rank_dft_kernel<2><<<blocks, 256, 0, stream>>>(in, out, chunk, flat0, S, nx, inverse); break;
        case 4: poisson_rank_dft_kernel<4><<<blocks, 256, 0, stream>>>(in, out, chunk, flat0, S, nx, inverse); break;
        case 8: poisson_rank_dft_kernel<8><<<blocks, 256, 0, stream>>>(in, out, chunk, flat0, S, nx, inverse); break;
        default: poisson_rank_dft_kernel<16><<<blocks, 256, 0, stream>>>(in, out, chunk, flat0, S, nx, inverse); break;
    }
    FLUXGNN_CUDA_OK(cudaGetLastError());
    count_launch();
    return FLUXGNN_OK;
}

// In-place solve of the P complex signals y[P][S] of rank kr (local bins kl <-> global bins kr + G kl); `scratch`
// holds another P*S complex numbers (unused for S <= 2^14).
int launch_poisson_dist_local(float2* y, float2* scratch, int P, int S, int G, int kr, double length, cudaStream_t stream) {
    int rc = dist_shape_ok(S, G);
    if (rc != FLUXGNN_OK) return rc;
    if (P < 1 || P > 65535 || kr < 0 || kr >= G) return set_error(FLUXGNN_EINVAL, "distributed field solve: P=%d rank=%d of %d", P, kr, G);
    const int mbits = ilog2_exact(S);
    if (mbits <= kFftRowBits) {
        const int M = 1 << mbits;
        const size_t smem = (size_t)(M + M / 16) * sizeof(float2);
        FLUXGNN_CUDA_OK(cudaFuncSetAttribute(poisson_fft_small_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const int threads = M / 16 < kFftThreads ? (M / 16 < 64 ? 64 : M / 16) : kFftThreads;
        poisson_fft_small_kernel<true><<<P, threads, smem, stream>>>((const float*)y, 2LL * S, (float*)y, 2LL * S, mbits, length, G, kr);
        FLUXGNN_CUDA_OK(cudaGetLastError());
        count_launch();
        return FLUXGNN_OK;
    }
    if (scratch == nullptr) return set_error(FLUXGNN_EINVAL, "distributed field solve: scratch required for slabs above 2^14 cells");
    const int bits1 = mbits - kRowBits;
    const int tile_bits = column_tile_bits(bits1);
    const int T = 1 << (tile_bits - bits1);
    const size_t tile = (size_t)1 << tile_bits;
    const size_t smem = (tile + (tile >> 5) * 4) * sizeof(float2);
    const size_t smem_row = (size_t)2 * kRowPad * sizeof(float2);
    const float scale = (float)(length / (6.283185307179586476925 * (double)S));
    dim3 gcol((unsigned)(kRowLen / T), (unsigned)P), grow((unsigned)((1 << bits1) / 2), (unsigned)P);
    FLUXGNN_CUDA_OK(cudaFuncSetAttribute(poisson_fft_rows_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_row));
    if (bits1 >= 9 && bits1 <= 11) {
        const int mode = tma_cols_mode(bits1, (long long)P * (kRowLen / T));
        if (mode != 0) {
            rc = launch_cols_tma<false, false>((const float*)y, 2LL * S, (float*)scratch, 2LL * S, P, bits1, mode, stream);
            if (rc != FLUXGNN_OK) return rc;
            poisson_fft_rows_diag_kernel<<<grow, kFftStepThreads, smem_row, stream>>>(scratch, bits1, scale, G, kr);
            FLUXGNN_CUDA_OK(cudaGetLastError());
            count_launch();
            return launch_cols_tma<true, true>((const float*)scratch, 2LL * S, (float*)y, 2LL * S, P, bits1, mode, stream);
        }
    }
    if (staged_cols_ok(bits1, (long long)P * (kRowLen / T))) {
        rc = launch_cols_staged<false, false>(y, S, scratch, S, P, bits1, stream);
        if (rc != FLUXGNN_OK) return rc;
        poisson_fft_rows_diag_kernel<<<grow, kFftStepThreads, smem_row, stream>>>(scratch, bits1, scale, G, kr);
        FLUXGNN_CUDA_OK(cudaGetLastError());
        count_launch();
        return launch_cols_staged<true, true>(scratch, S, y, S, P, bits1, stream);
    }
#define FLUXGNN_DIST_COLS(BITS1)                                                                                         \
    case BITS1:                                                                                                          \
        FLUXGNN_CUDA_OK(cudaFuncSetAttribute(poisson_fft_cols_fwd_kernel<BITS1, false>,                                  \
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                   \
        FLUXGNN_CUDA_OK(cudaFuncSetAttribute(poisson_fft_cols_inv_kernel<BITS1>,                                         \
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                   \
        poisson_fft_cols_fwd_kernel<BITS1, false><<<gcol, column_threads(BITS1), smem, stream>>>((const float*)y, 2LL * S, scratch); \
        poisson_fft_rows_diag_kernel<<<grow, kFftStepThreads, smem_row, stream>>>(scratch, bits1, scale, G, kr);         \
        poisson_fft_cols_inv_kernel<BITS1><<<gcol, column_threads(BITS1), smem, stream>>>(scratch, (float*)y, 2LL * S);  \
        break;
    switch (bits1) {
        FLUXGNN_DIST_COLS(3) FLUXGNN_DIST_COLS(4) FLUXGNN_DIST_COLS(5) FLUXGNN_DIST_COLS(6) FLUXGNN_DIST_COLS(7)
        FLUXGNN_DIST_COLS(8) FLUXGNN_DIST_COLS(9) FLUXGNN_DIST_COLS(10) FLUXGNN_DIST_COLS(11) FLUXGNN_DIST_COLS(12)
        default: return set_error(FLUXGNN_EUNSUP, "distributed field solve: unsupported column length 2^%d", bits1);
    }
#undef FLUXGNN_DIST_COLS
    FLUXGNN_CUDA_OK(cudaGetLastError());
    count_launch(3);
    return FLUXGNN_OK;
}

}  // namespace fluxgnn
