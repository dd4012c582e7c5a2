// Kernels of field_kernels.cu (declarations for api.cu).
#pragma once

#include <cuda_runtime.h>

namespace fluxgnn {

// Every intermediate is rounded to fp32 exactly where numpy rounds it
// (python-float scalars are weak, so c, dt, nu and dx^2 act as float32).
struct FvOut { float n, u, fn; };
// Correctly rounded x / b without the IEEE division subroutine (whose special-case path nearly every cell takes here:
// the second difference of a smooth u is a few ulps or exactly zero, and 99 % of the divisions left the fast path).
// Markstein: with r = RN(1/b), q0 = RN(x r), e = x - q0 b (exact in an FMA), q = RN(q0 + e r) is RN(x/b) unless b's
// significand is all ones; fv_reciprocal() returns 0 for such a b, for b >= 1 (quotients could become subnormal) and
// for subnormal b, which selects __fdiv_rn.  Signed zeros and non-finite quotients are passed through from q0.
__device__ __forceinline__ float fv_reciprocal(float b) {
    const unsigned bits = __float_as_uint(b);
    const bool ok = b > 0.f && b < 1.f && (bits & 0x7fffffu) != 0x7fffffu && (bits >> 23) != 0u;
    return ok ? __frcp_rn(b) : 0.f;
}
__device__ __forceinline__ float fv_div(float x, float b, float r) {
    if (r == 0.f) return __fdiv_rn(x, b);
    const float q0 = __fmul_rn(x, r);
    const float e = __fmaf_rn(-q0, b, x);
    const float q = __fmaf_rn(e, r, q0);
    return (x == 0.f || !(fabsf(q0) < __int_as_float(0x7f800000))) ? q0 : q;
}

__device__ __forceinline__ FvOut fv_cell(float nm, float n0, float um, float u0, float up, float e0,
                                         float c, float dt, float nu, float dx2, float rdx2) {
    FvOut o;
    o.fn = __fmul_rn(n0, u0);                                                          // :70-71
    const float fnm = __fmul_rn(nm, um);
    o.n = __fsub_rn(n0, __fmul_rn(c, __fsub_rn(o.fn, fnm)));                           // :85-86
    const float fu = __fmul_rn(__fmul_rn(0.5f, u0), u0);                               // :73-74
    const float fum = __fmul_rn(__fmul_rn(0.5f, um), um);
    const float u_adv = __fsub_rn(u0, __fmul_rn(c, __fsub_rn(fu, fum)));               // :90-91
    const float lap = fv_div(__fadd_rn(__fsub_rn(up, __fmul_rn(2.0f, u0)), um), dx2, rdx2);   // :76-78
    o.u = __fadd_rn(u_adv, __fmul_rn(dt, __fadd_rn(e0, __fmul_rn(nu, lap))));          // :94
    return o;
}

// The same update for the two cells of a float2 (cells i, i+1), packed fp32x2 where that is safe: every operation is
// the per-component round-to-nearest operation of fv_cell, in the same order, so the results are bit-identical.
// Wherever a PRODUCT feeds a SUM the sum is written as two scalar operations: ptxas (12.9) contracts mul.rn.f32x2 +
// add.rn.f32x2 into one FFMA2 -- which it never does for scalar .rn operations -- and that would skip the rounding of
// the product (measured: n' and u' off by an ulp whenever the flux difference is not below rounding; -fmad=false does
// not prevent it).  The one packed product-into-sum left, up - 2u, is exact either way because 2u is exact.
//   nl, ul: n and u of cell i-1;  ur: u of cell i+2;  r = fv_reciprocal(dx2) must be non-zero.
__device__ __forceinline__ float2 fv_neg2(float2 a) { return make_float2(-a.x, -a.y); }
__device__ __forceinline__ void fv_pair(float nl, float2 n2, float ul, float2 u2, float ur, float2 e2, float c, float dt,
                                        float nu, float dx2, float r, float2& n_new, float2& u_new) {
    const float2 um = make_float2(ul, u2.x), up = make_float2(u2.y, ur);
    const float2 c2 = make_float2(c, c), half2 = make_float2(0.5f, 0.5f), two2 = make_float2(2.0f, 2.0f);
    const float2 fn = __fmul2_rn(n2, u2);                                             // :70-71
    const float fnl = __fmul_rn(nl, ul);
    const float2 tn = __fmul2_rn(c2, make_float2(__fsub_rn(fn.x, fnl), __fsub_rn(fn.y, fn.x)));
    n_new = make_float2(__fsub_rn(n2.x, tn.x), __fsub_rn(n2.y, tn.y));               // :85-86
    const float2 fu = __fmul2_rn(__fmul2_rn(half2, u2), u2);                          // :73-74
    const float ful = __fmul_rn(__fmul_rn(0.5f, ul), ul);
    const float2 tu = __fmul2_rn(c2, make_float2(__fsub_rn(fu.x, ful), __fsub_rn(fu.y, fu.x)));
    const float2 u_adv = make_float2(__fsub_rn(u2.x, tu.x), __fsub_rn(u2.y, tu.y));   // :90-91
    const float2 x = __fadd2_rn(__fadd2_rn(up, fv_neg2(__fmul2_rn(two2, u2))), um);   // :76-78
    const float2 r2 = make_float2(r, r), b2 = make_float2(dx2, dx2);
    const float2 q0 = __fmul2_rn(x, r2);
    const float2 e = __ffma2_rn(fv_neg2(q0), b2, x);
    float2 lap = __ffma2_rn(e, r2, q0);
    const float inf = __int_as_float(0x7f800000);
    lap.x = (x.x == 0.f || !(fabsf(q0.x) < inf)) ? q0.x : lap.x;
    lap.y = (x.y == 0.f || !(fabsf(q0.y) < inf)) ? q0.y : lap.y;
    const float2 v = __fmul2_rn(make_float2(nu, nu), lap);
    const float2 z = __fmul2_rn(make_float2(dt, dt), make_float2(__fadd_rn(e2.x, v.x), __fadd_rn(e2.y, v.y)));
    u_new = make_float2(__fadd_rn(u_adv.x, z.x), __fadd_rn(u_adv.y, z.y));           // :94
}

__global__ void poisson_table_kernel(int nx, double length, double* gtab);
__global__ void poisson_direct_kernel(const float* n, long long n_stride, float* E, long long e_stride,
                                      const double* gtab, int nx);
__global__ void baseline_fv_kernel(const float* in, float* out, float* flux_n, int B, int nx,
                                   float c, float dt, float nu, float dx2);
constexpr int kBaselineSmallMaxNx = 1024;     // persistent one-CTA-per-IC classical rollout (direct field solve, nx^2 work per step)
__global__ void baseline_small_rollout_kernel(const float* in, float* out, float* traj, float* flux_n, const double* gtab, int B,
                                              int nx, int steps, int record_every, float c, float dt, float nu, float dx2);
__global__ void baseline_fv_slab_kernel(const float* in, float* out, float* flux_n, int B, int owned, int halo,
                                        int out_ld, int out_off, int vec, float c, float dt, float nu, float dx2,
                                        float* left_out, float* right_out);
__global__ void pack_weights_kernel(const float* w_in, const float* b_in, const float* w_upd, const float* b_upd,
                                    const float* w_e1, const float* b_e1, const float* w_e2, const float* b_e2,
                                    int L, float* packed);

__global__ void pack_weights_tc_kernel(const float* w_in, const float* b_in, const float* w_upd, const float* b_upd,
                                       const float* w_e1, const float* b_e1, const float* w_e2, const float* b_e2,
                                       int L, float* packed);
__global__ void pack_weights_tc16_kernel(const float* w_in, const float* b_in, const float* w_upd, const float* b_upd,
                                         const float* w_e1, const float* b_e1, const float* w_e2, const float* b_e2,
                                         int L, int format, unsigned char* packed);
__global__ void rollout_metrics_kernel(const float* pred, const float* truth, int nx, float* out);
__global__ void ffma_probe_kernel(float* out, int iters, float a, float b);
__global__ void ffma2_probe_kernel(float* out, int iters, float a, float b);
constexpr int kFfmaProbeFlopsPerIter = 2 * 8 * 16;   // per thread per iteration

constexpr int kPoissonDirectMaxNx = 12288;   // rho staged in 48 KiB of shared memory

// fft_poisson.cu: power-of-two grids, 2^kFftMinBits <= nx <= 2^kFftMaxBits
constexpr int kFftMinBits = 8;
constexpr int kFftRowBits = 14;              // longest complex transform done inside one CTA (nx = 2^15 reals)
#ifndef FLUXGNN_FFT_STEP_ROW_BITS
#define FLUXGNN_FFT_STEP_ROW_BITS 12         // four-step: row length 2^12; a CTA holds a PAIR of rows (66 KiB -> 3 CTAs per SM)
#endif
#ifndef FLUXGNN_FFT_STEP_COL_BITS
#define FLUXGNN_FFT_STEP_COL_BITS 13         // four-step: N1 * T = 2^13 complex per column tile (64 KiB)
#endif
constexpr int kFftMaxBits = 25;
bool poisson_fft_supported(int nx);
size_t poisson_fft_workspace_bytes(int B, int nx);
int launch_poisson_fft(const float* n, long long n_stride, float* E, long long e_stride, int B, int nx,
                       double length, void* workspace, cudaStream_t stream);

// fused classical step (fft_poisson.cu): inverse column stages + finite-volume update + forward column stages in one
// kernel, n and u in a tile-major private layout.  Available for four-step grids whose column tiles are at most 32
// columns wide (nx >= 2^21).
bool baseline_fused_supported(int nx, float dx2);
size_t baseline_fused_workspace_floats(int B, int nx);        // 2 x (Pn, Pu) + 2 x H
int launch_baseline_to_tiles(const float* state, float* fused_ws, int slot, int B, int nx, cudaStream_t stream);
int launch_baseline_fused_cols(float2* Y, float* fused_ws, int slot_in, float* nat_out, int B, int nx,
                               float c, float dt, float nu, float dx2, cudaStream_t stream);
int launch_poisson_fft_rows(float2* Y, int B, int nx, double length, cudaStream_t stream);
int launch_poisson_fft_cols(const float* n, long long n_stride, float2* Y, float* E, long long e_stride, int B, int nx,
                            int inverse, cudaStream_t stream);

// distributed field solve (fft_poisson.cu): pack / unpack pairs of ICs, the cross-rank DFT stage, the local solve
int launch_poisson_dist_pack(const float* n, long long ic_stride, int B, int S, float2* z, int unpack, float* E,
                             cudaStream_t stream);
int launch_poisson_rank_dft(const float2* in, float2* out, int G, long long chunk, long long flat0, int S, int inverse,
                            cudaStream_t stream);
int launch_poisson_dist_local(float2* y, float2* scratch, int P, int S, int G, int kr, double length, cudaStream_t stream);

// scan_poisson.cu: classical rollout with the field reconstructed by a prefix sum and certified against the spectral
// operator (long grids, nx >= 4096, nx % 8 == 0); the state between steps is (n, u) only
bool baseline_scan_supported(int B, int nx);
size_t baseline_scan_workspace_bytes(int B, int nx, int sms);
cudaError_t launch_baseline_rollout_scan(const float* state_in, float* state_out, int B, int nx, double length, float c, float dt,
                                         float nu, float dx2, int steps, int record_every, float* traj, float* flux_n,
                                         double tol, void* workspace, int* flag, int sms, cudaStream_t stream, int* launches);

// scan_poisson.cu, slab form (domain decomposition): per-rank sums + message, reconstruction from the gathered messages
bool scan_slab_supported(int B, int S);
size_t scan_slab_workspace_bytes(int B, int S, int sms);
cudaError_t launch_scan_slab_sums(const float* n, long long n_ld, int B, int S, long long j_base, void* workspace, void* msg,
                                  int sms, cudaStream_t stream, void* const* bases = nullptr, long long offset = 0, int rank = 0,
                                  int world = 0);
cudaError_t launch_scan_slab_field(const float* n, long long n_ld, float* E, long long e_ld, int B, int S, int rank, int ranks,
                                   double length, const void* msg_all, void* workspace, double tol, int step, int* flag, int sms,
                                   cudaStream_t stream, float* E_left = nullptr, float* E_right = nullptr, int halo = 0);
cudaError_t launch_scan_slab_certify(int B, int S, int ranks, double length, const void* msg_all, double tol, int step, int* flag,
                                     cudaStream_t stream);

// field_kernels.cu: peer-memory (symmetric memory, NVLink P2P) halo stores and all-gather of a domain-decomposed grid
__global__ void peer_halo_push_kernel(const float* ext, float* ext_left, float* ext_right, int B, int owned, int halo,
                                      int ch0, int ch1);
__global__ void peer_allgather_kernel(const uint4* src, long long bytes, void* const* bases, long long offset, int rank);

}  // namespace fluxgnn
