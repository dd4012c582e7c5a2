// Field solve, classical finite-volume step and weight packing kernels (sm_100a).
//
//   poisson_table_kernel   g = Re ifft(i/k)                    src/baseline_solver.py:26,59-68
//   poisson_direct_kernel  E = g (*) (n - 1), any nx           src/baseline_solver.py:59-68
//   baseline_fv_kernel     upwind / forward-Euler + viscosity  src/baseline_solver.py:70-94
//   pack_weights_kernel    state_dict -> streaming layout      src/flux_gnn.py:17-38
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "common.cuh"
#include "field_kernels.cuh"

namespace fluxgnn {

// ---------------------------------------------------------------------------
// g_j = Re (1/N) sum_{m != 0} (i/k_m) e^{2 pi i m j/N},  k_m = 2 pi m / length
//     = -(length / (pi N)) * sum_{m=1..M} sin(2 pi m j / N) / m,   M = floor((N-1)/2)
// (the m = N/2 term of an even N is purely imaginary and vanishes under Re()).
// One block per j, exact integer reduction of m*j mod N keeps the phase accurate.
// ---------------------------------------------------------------------------
__global__ void poisson_table_kernel(int nx, double length, double* __restrict__ gtab) {
    const int j = blockIdx.x;
    const int M = (nx - 1) / 2;
    double s = 0.0;
    for (int m = 1 + threadIdx.x; m <= M; m += blockDim.x) {
        const long long r = ((long long)m * j) % nx;
        s += sinpi(2.0 * (double)r / (double)nx) / (double)m;
    }
    __shared__ double red[32];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        s = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.0;
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (threadIdx.x == 0) gtab[j] = -(length / (3.14159265358979323846 * (double)nx)) * s;
    }
}

// One block per (IC = blockIdx.x, slab of 256 outputs = blockIdx.y).  rho staged in shared memory as fp32
// (rho = n - 1 rounded to fp32 exactly as numpy does), fp64 accumulation.
__global__ void __launch_bounds__(256) poisson_direct_kernel(const float* __restrict__ n, long long n_stride,
                                                             float* __restrict__ E, long long e_stride,
                                                             const double* __restrict__ gtab, int nx) {
    extern __shared__ float rho[];
    const int ic = blockIdx.x;
    const float* src = n + (size_t)ic * n_stride;
    for (int i = threadIdx.x; i < nx; i += blockDim.x) rho[i] = __fsub_rn(src[i], 1.0f);
    __syncthreads();
    const int j = blockIdx.y * blockDim.x + threadIdx.x;
    if (j >= nx) return;
    double acc0 = 0.0, acc1 = 0.0;
    int d = j;                       // d = (j - i) mod nx
    int i = 0;
    for (; i + 1 < nx; i += 2) {
        acc0 = fma(__ldg(gtab + d), (double)rho[i], acc0);
        d = (d == 0) ? nx - 1 : d - 1;
        acc1 = fma(__ldg(gtab + d), (double)rho[i + 1], acc1);
        d = (d == 0) ? nx - 1 : d - 1;
    }
    if (i < nx) acc0 = fma(__ldg(gtab + d), (double)rho[i], acc0);
    E[(size_t)ic * e_stride + j] = (float)(acc0 + acc1);
}

// One thread per 4 consecutive cells when nx % 4 == 0 (128-bit loads/stores, the two halo
// cells come from the neighbouring quads through the read-only cache); scalar otherwise.
__global__ void __launch_bounds__(256) baseline_fv_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                          float* __restrict__ flux_n, int B, int nx,
                                                          float c, float dt, float nu, float dx2) {
    const float rdx2 = fv_reciprocal(dx2);
    if ((nx & 3) == 0) {
        const int quads = nx >> 2;
        const long long total = (long long)B * quads;
        for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
             idx += (long long)gridDim.x * blockDim.x) {
            const int ic = (int)(idx / quads);
            const int i = (int)(idx - (long long)ic * quads) << 2;
            const float* pn = in + (size_t)ic * 3 * nx;
            const float* pu = pn + nx;
            const float* pe = pu + nx;
            const float4 n4 = *reinterpret_cast<const float4*>(pn + i);
            const float4 u4 = *reinterpret_cast<const float4*>(pu + i);
            const float4 e4 = *reinterpret_cast<const float4*>(pe + i);
            const int im = (i == 0) ? nx - 1 : i - 1;
            const int ip = (i + 4 == nx) ? 0 : i + 4;
            const float nm = __ldg(pn + im), um = __ldg(pu + im), up = __ldg(pu + ip);
            const FvOut a = fv_cell(nm, n4.x, um, u4.x, u4.y, e4.x, c, dt, nu, dx2, rdx2);
            const FvOut b = fv_cell(n4.x, n4.y, u4.x, u4.y, u4.z, e4.y, c, dt, nu, dx2, rdx2);
            const FvOut d = fv_cell(n4.y, n4.z, u4.y, u4.z, u4.w, e4.z, c, dt, nu, dx2, rdx2);
            const FvOut e = fv_cell(n4.z, n4.w, u4.z, u4.w, up, e4.w, c, dt, nu, dx2, rdx2);
            float* po = out + (size_t)ic * 3 * nx;
            *reinterpret_cast<float4*>(po + i) = make_float4(a.n, b.n, d.n, e.n);
            *reinterpret_cast<float4*>(po + nx + i) = make_float4(a.u, b.u, d.u, e.u);
            if (flux_n != nullptr)
                *reinterpret_cast<float4*>(flux_n + (size_t)ic * nx + i) = make_float4(a.fn, b.fn, d.fn, e.fn);
        }
        return;
    }
    const long long total = (long long)B * nx;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int ic = (int)(idx / nx);
        const int i = (int)(idx - (long long)ic * nx);
        const int im = (i == 0) ? nx - 1 : i - 1;
        const int ip = (i == nx - 1) ? 0 : i + 1;
        const float* pn = in + (size_t)ic * 3 * nx;
        const float* pu = pn + nx;
        const float* pe = pu + nx;
        const FvOut o = fv_cell(pn[im], pn[i], pu[im], pu[i], pu[ip], pe[i], c, dt, nu, dx2, rdx2);
        float* po = out + (size_t)ic * 3 * nx;
        po[i] = o.n;
        po[nx + i] = o.u;
        if (flux_n != nullptr) flux_n[idx] = o.fn;
    }
}

// Short grids (the reference's default is 64 cells): the whole classical rollout of one IC in ONE persistent CTA, state
// and field-solve table in shared memory -- instead of two launches per step (the 50-step timing protocol of
// scripts/evaluation/benchmark_timing.py was 100 launches of ~7 us).  Same arithmetic as baseline_fv_kernel +
// poisson_direct_kernel (fv_cell; two alternating float64 accumulators over i, table index counting down), so the
// results are bit-identical to the launch-per-step path.  Dynamic shared memory: 6 nx floats + nx doubles.
__global__ void __launch_bounds__(256) baseline_small_rollout_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                                     float* __restrict__ traj, float* __restrict__ flux_n,
                                                                     const double* __restrict__ gtab, int B, int nx, int steps,
                                                                     int record_every, float c, float dt, float nu, float dx2) {
    extern __shared__ __align__(16) unsigned char small_raw[];
    double* g = reinterpret_cast<double*>(small_raw);
    float* sn = reinterpret_cast<float*>(g + nx);
    float* su = sn + nx;
    float* se = su + nx;
    float* rho = se + nx;
    float* nn = rho + nx;
    float* un = nn + nx;
    const int ic = blockIdx.x, tid = threadIdx.x, T = blockDim.x;
    const float rdx2 = fv_reciprocal(dx2);
    const float* src = in + (size_t)ic * 3 * nx;
    for (int i = tid; i < nx; i += T) {
        sn[i] = src[i]; su[i] = src[nx + i]; se[i] = src[2 * nx + i];
        g[i] = gtab[i];
    }
    __syncthreads();
    const size_t state_floats = (size_t)B * 3 * nx;
    for (int t = 0; t < steps; ++t) {
        for (int i = tid; i < nx; i += T) {
            const int im = (i == 0) ? nx - 1 : i - 1, ip = (i == nx - 1) ? 0 : i + 1;
            const FvOut o = fv_cell(sn[im], sn[i], su[im], su[i], su[ip], se[i], c, dt, nu, dx2, rdx2);
            nn[i] = o.n; un[i] = o.u;
            if (flux_n != nullptr) flux_n[((size_t)t * B + ic) * nx + i] = o.fn;
        }
        __syncthreads();
        for (int i = tid; i < nx; i += T) {
            sn[i] = nn[i]; su[i] = un[i];
            rho[i] = __fsub_rn(nn[i], 1.0f);
        }
        __syncthreads();
        for (int j = tid; j < nx; j += T) {
            double acc0 = 0.0, acc1 = 0.0;
            int d = j, i = 0;
            for (; i + 1 < nx; i += 2) {
                acc0 = fma(g[d], (double)rho[i], acc0);
                d = (d == 0) ? nx - 1 : d - 1;
                acc1 = fma(g[d], (double)rho[i + 1], acc1);
                d = (d == 0) ? nx - 1 : d - 1;
            }
            if (i < nx) acc0 = fma(g[d], (double)rho[i], acc0);
            se[j] = (float)(acc0 + acc1);
        }
        __syncthreads();
        if (traj != nullptr && (t + 1) % record_every == 0) {
            float* dst = traj + (size_t)((t + 1) / record_every - 1) * state_floats + (size_t)ic * 3 * nx;
            for (int i = tid; i < nx; i += T) { dst[i] = sn[i]; dst[nx + i] = su[i]; dst[2 * nx + i] = se[i]; }
        }
    }
    float* dst = out + (size_t)ic * 3 * nx;
    for (int i = tid; i < nx; i += T) { dst[i] = sn[i]; dst[nx + i] = su[i]; dst[2 * nx + i] = se[i]; }
}

// The same update for one slab of a domain-decomposed grid (baseline-only row): `in` is the extended
// state [B][3][owned + 2*halo] whose ghost cells replace the periodic wrap (halo >= 1; the stencil needs one cell);
// n', u' go to out[B][3][out_ld] at column out_off + cell (E' comes from the distributed field solve).
__global__ void __launch_bounds__(256) baseline_fv_slab_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                               float* __restrict__ flux_n, int B, int owned, int halo,
                                                               int out_ld, int out_off, int vec,
                                                               float c, float dt, float nu, float dx2,
                                                               float* __restrict__ left_out, float* __restrict__ right_out) {
    const float rdx2 = fv_reciprocal(dx2);
    const int ld = owned + 2 * halo;
    // slabs over peer memory (left_out / right_out = the ring neighbours' next extended states, nullable): n', u' of the
    // first / last `halo` owned cells also go into the neighbours' ghost zones -- the halo exchange is this store
    auto edge_store = [&](int ic, int cell, float nv, float uv) {
        if (left_out != nullptr && cell < halo) {
            float* pl = left_out + (size_t)ic * 3 * out_ld + out_off + owned + cell;
            pl[0] = nv;
            pl[out_ld] = uv;
        }
        if (right_out != nullptr && cell >= owned - halo) {
            float* pr = right_out + (size_t)ic * 3 * out_ld + out_off - owned + cell;
            pr[0] = nv;
            pr[out_ld] = uv;
        }
    };
    if (vec) {
        const int quads = owned >> 2;
        const long long total = (long long)B * quads;
        for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
             idx += (long long)gridDim.x * blockDim.x) {
            const int ic = (int)(idx / quads);
            const int i = (int)(idx - (long long)ic * quads) << 2;
            const float* pn = in + (size_t)ic * 3 * ld + halo + i;
            const float* pu = pn + ld;
            const float* pe = pu + ld;
            const float4 n4 = *reinterpret_cast<const float4*>(pn);
            const float4 u4 = *reinterpret_cast<const float4*>(pu);
            const float4 e4 = *reinterpret_cast<const float4*>(pe);
            const float nm = __ldg(pn - 1), um = __ldg(pu - 1), up = __ldg(pu + 4);
            const FvOut a = fv_cell(nm, n4.x, um, u4.x, u4.y, e4.x, c, dt, nu, dx2, rdx2);
            const FvOut b = fv_cell(n4.x, n4.y, u4.x, u4.y, u4.z, e4.y, c, dt, nu, dx2, rdx2);
            const FvOut d = fv_cell(n4.y, n4.z, u4.y, u4.z, u4.w, e4.z, c, dt, nu, dx2, rdx2);
            const FvOut e = fv_cell(n4.z, n4.w, u4.z, u4.w, up, e4.w, c, dt, nu, dx2, rdx2);
            float* po = out + (size_t)ic * 3 * out_ld + out_off + i;
            *reinterpret_cast<float4*>(po) = make_float4(a.n, b.n, d.n, e.n);
            *reinterpret_cast<float4*>(po + out_ld) = make_float4(a.u, b.u, d.u, e.u);
            if (flux_n != nullptr)
                *reinterpret_cast<float4*>(flux_n + (size_t)ic * owned + i) = make_float4(a.fn, b.fn, d.fn, e.fn);
            if (i < halo || i + 4 > owned - halo) {
                edge_store(ic, i, a.n, a.u); edge_store(ic, i + 1, b.n, b.u);
                edge_store(ic, i + 2, d.n, d.u); edge_store(ic, i + 3, e.n, e.u);
            }
        }
        return;
    }
    const long long total = (long long)B * owned;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int ic = (int)(idx / owned);
        const int i = (int)(idx - (long long)ic * owned);
        const float* pn = in + (size_t)ic * 3 * ld + halo + i;
        const float* pu = pn + ld;
        const FvOut o = fv_cell(pn[-1], pn[0], pu[-1], pu[0], pu[1], pu[ld], c, dt, nu, dx2, rdx2);
        float* po = out + (size_t)ic * 3 * out_ld + out_off + i;
        po[0] = o.n;
        po[out_ld] = o.u;
        edge_store(ic, i, o.n, o.u);
        if (flux_n != nullptr) flux_n[idx] = o.fn;
    }
}

// FP32 pipe probe: 16 independent FFMA chains per thread, register operands only.
// Used by bench.py to measure the FFMA roofline of the box it runs on.
__global__ void __launch_bounds__(256) ffma_probe_kernel(float* __restrict__ out, int iters, float a, float b) {
    float acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = (float)(threadIdx.x + i);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int rep = 0; rep < 8; ++rep)
#pragma unroll
            for (int i = 0; i < 16; ++i) acc[i] = fmaf(acc[i], a, b);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += acc[i];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// Same probe with packed fma.rn.f32x2 (SASS FFMA2): 8 independent float2 chains.
__global__ void __launch_bounds__(256) ffma2_probe_kernel(float* __restrict__ out, int iters, float a, float b) {
    float2 acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = make_float2((float)(threadIdx.x + i), (float)(threadIdx.x - i));
    const float2 a2 = make_float2(a, a), b2 = make_float2(b, b);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int rep = 0; rep < 8; ++rep)
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] = __ffma2_rn(acc[i], a2, b2);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += acc[i].x + acc[i].y;
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// packed small block + (L+1) x 2 K-major halves (see common.cuh)
__global__ void pack_weights_kernel(const float* __restrict__ w_in, const float* __restrict__ b_in,
                                    const float* __restrict__ w_upd, const float* __restrict__ b_upd,
                                    const float* __restrict__ w_e1, const float* __restrict__ b_e1,
                                    const float* __restrict__ w_e2, const float* __restrict__ b_e2,
                                    int L, float* __restrict__ packed) {
    const size_t total = packed_floats(L);
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        float v = 0.f;
        if (idx < (size_t)SmallParams::count) {
            const int o = (int)idx;
            if (o < SmallParams::b_in) {                       // w_in[f][n] <- W_in[n][f]
                const int f = o / kH, n = o % kH;
                v = w_in[n * kF + f];
            } else if (o < SmallParams::b_upd) {
                v = b_in[o - SmallParams::b_in];
            } else if (o < SmallParams::b_e1) {
                const int q = o - SmallParams::b_upd;
                v = (q < L * kH) ? b_upd[q] : 0.f;
            } else if (o < SmallParams::w_e2) {
                v = b_e1[o - SmallParams::b_e1];
            } else if (o < SmallParams::b_e2) {
                v = w_e2[o - SmallParams::w_e2];
            } else if (o == SmallParams::b_e2) {
                v = b_e2[0];
            }
        } else {
            const size_t q = idx - SmallParams::count;
            const int layer = (int)(q / kLayerFloats);
            const int r = (int)(q % kLayerFloats);
            const int half = r / kHalfFloats;                  // 0: W[:, H:], 1: W[:, :H]
            const int k = (r % kHalfFloats) / kH;
            const int n = weight_column(r % kH);
            const float* W = (layer < L) ? (w_upd + (size_t)layer * kH * 2 * kH) : w_e1;
            v = W[(size_t)n * 2 * kH + (half == 0 ? kH : 0) + k];
        }
        packed[idx] = v;
    }
}

// Rollout diagnostics of the reference's evaluation scripts, one block per stored state (t, b):
//   out[(t*B + b)*8 + 0..2] = mean((pred - true)^2) per channel (n, u, E)   scripts/evaluation/evaluate_all.py:130-132
//   out[.. + 3] = 0.5 * mean(u^2 + E^2)   (energy)                          evaluate_all.py:136-137
//   out[.. + 4] = mean(n)                 (charge)                          evaluate_all.py:142-143
//   out[.. + 5] = number of non-finite values in the state                  evaluate_long_rollout.py:56-60
// fp64 accumulation; `truth` may be null (MSE columns are 0 then).
__global__ void __launch_bounds__(256) rollout_metrics_kernel(const float* __restrict__ pred,
                                                              const float* __restrict__ truth, int nx,
                                                              float* __restrict__ out) {
    const size_t sid = blockIdx.x;                       // flattened (t, b)
    const float* p = pred + sid * 3 * nx;
    const float* q = truth ? truth + sid * 3 * nx : nullptr;
    double acc[6] = {0, 0, 0, 0, 0, 0};
    for (int i = threadIdx.x; i < nx; i += blockDim.x) {
        const float n = p[i], u = p[nx + i], e = p[2 * (size_t)nx + i];
        if (q != nullptr) {
            const double dn = (double)n - q[i], du = (double)u - q[nx + i], de = (double)e - q[2 * (size_t)nx + i];
            acc[0] += dn * dn; acc[1] += du * du; acc[2] += de * de;
        }
        acc[3] += (double)u * u + (double)e * e;
        acc[4] += n;
        acc[5] += (isfinite(n) ? 0.0 : 1.0) + (isfinite(u) ? 0.0 : 1.0) + (isfinite(e) ? 0.0 : 1.0);
    }
    __shared__ double red[6][8];
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        double v = acc[k];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0) red[k][threadIdx.x >> 5] = v;
    }
    __syncthreads();
    if (threadIdx.x < 6) {
        double v = 0.0;
        for (int w = 0; w < 8; ++w) v += red[threadIdx.x][w];
        const int k = threadIdx.x;
        float r;
        if (k < 3) r = (float)(v / nx);
        else if (k == 3) r = (float)(0.5 * v / nx);
        else if (k == 4) r = (float)(v / nx);
        else r = (float)v;
        out[sid * 8 + k] = r;
    }
    if (threadIdx.x >= 6 && threadIdx.x < 8) out[sid * 8 + threadIdx.x] = 0.f;
}

// Tensor-path stream: small block + pre-swizzled UMMA operand images (layout: common.cuh)
// 16-bit tensor-path packing (layout in common.cuh): one thread per 16-bit element of the stream,
// the small block is copied as fp32 by the first threads.  format 0 = fp16, 1 = bf16.
__global__ void pack_weights_tc16_kernel(const float* __restrict__ w_in, const float* __restrict__ b_in,
                                         const float* __restrict__ w_upd, const float* __restrict__ b_upd,
                                         const float* __restrict__ w_e1, const float* __restrict__ b_e1,
                                         const float* __restrict__ w_e2, const float* __restrict__ b_e2,
                                         int L, int format, unsigned char* __restrict__ packed) {
    float* small = reinterpret_cast<float*>(packed);
    unsigned short* stream = reinterpret_cast<unsigned short*>(packed + (size_t)SmallParams::count * sizeof(float));
    const size_t elems = (size_t)(L + 1) * kTc16UnitsPerLayer * (kTc16UnitBytes / 2);
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < elems + SmallParams::count;
         idx += (size_t)gridDim.x * blockDim.x) {
        if (idx < (size_t)SmallParams::count) {
            const int o = (int)idx;
            float v = 0.f;
            if (o < SmallParams::b_in) {
                const int f = o / kH, n = o % kH;
                v = w_in[n * kF + f];
            } else if (o < SmallParams::b_upd) {
                v = b_in[o - SmallParams::b_in];
            } else if (o < SmallParams::b_e1) {
                const int q = o - SmallParams::b_upd;
                v = (q < L * kH) ? b_upd[q] : 0.f;
            } else if (o < SmallParams::w_e2) {
                v = b_e1[o - SmallParams::b_e1];
            } else if (o < SmallParams::b_e2) {
                v = w_e2[o - SmallParams::w_e2];
            } else if (o == SmallParams::b_e2) {
                v = b_e2[0];
            }
            small[o] = v;
            continue;
        }
        const size_t q = idx - SmallParams::count;
        constexpr int kUnitElems = kTc16UnitBytes / 2;               // 8192
        const int layer = (int)(q / ((size_t)kTc16UnitsPerLayer * kUnitElems));
        const int u = (int)((q / kUnitElems) % kTc16UnitsPerLayer);
        const int within = (int)(q % kUnitElems);
        const int row = within >> 6;                                 // output feature n (128 bytes = 64 elements per row)
        const int chunk_phys = (within & 63) >> 3, e = within & 7;
        const int kk = ((chunk_phys ^ (row & 7)) << 3) | e;          // k inside the 64-wide K-block
        const int kb = u >> 2, blk = (u >> 1) & 1, part = u & 1;
        const float* W = (layer < L) ? (w_upd + (size_t)layer * kH * 2 * kH) : w_e1;
        const float w = kTc16WeightScale * W[(size_t)row * 2 * kH + (blk == 0 ? kH : 0) + kb * 64 + kk];
        unsigned short bits;
        if (format == 1) {
            const __nv_bfloat16 hi = __float2bfloat16_rn(w);
            bits = part == 0 ? __bfloat16_as_ushort(hi) : __bfloat16_as_ushort(__float2bfloat16_rn(w - __bfloat162float(hi)));
        } else {
            const __half hi = __float2half_rn(w);
            bits = part == 0 ? __half_as_ushort(hi) : __half_as_ushort(__float2half_rn(w - __half2float(hi)));
        }
        stream[q] = bits;
    }
}

__global__ void pack_weights_tc_kernel(const float* __restrict__ w_in, const float* __restrict__ b_in,
                                       const float* __restrict__ w_upd, const float* __restrict__ b_upd,
                                       const float* __restrict__ w_e1, const float* __restrict__ b_e1,
                                       const float* __restrict__ w_e2, const float* __restrict__ b_e2,
                                       int L, float* __restrict__ packed) {
    const size_t total = packed_tc_floats(L);
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        float v = 0.f;
        if (idx < (size_t)SmallParams::count) {
            const int o = (int)idx;
            if (o < SmallParams::b_in) {
                const int f = o / kH, n = o % kH;
                v = w_in[n * kF + f];
            } else if (o < SmallParams::b_upd) {
                v = b_in[o - SmallParams::b_in];
            } else if (o < SmallParams::b_e1) {
                const int q = o - SmallParams::b_upd;
                v = (q < L * kH) ? b_upd[q] : 0.f;
            } else if (o < SmallParams::w_e2) {
                v = b_e1[o - SmallParams::b_e1];
            } else if (o < SmallParams::b_e2) {
                v = w_e2[o - SmallParams::w_e2];
            } else if (o == SmallParams::b_e2) {
                v = b_e2[0];
            }
        } else {
            const size_t q = idx - SmallParams::count;
            const int layer = (int)(q / ((size_t)kTcUnitsPerLayer * kTcUnitFloats));
            const int u = (int)((q / kTcUnitFloats) % kTcUnitsPerLayer);
            const int within = (int)(q % kTcUnitFloats);
            const int row = within >> 5;                        // output feature n (128 bytes per row)
            const int chunk_phys = (within & 31) >> 2, e = within & 3;
            const int kk = ((chunk_phys ^ (row & 7)) << 2) | e; // k inside the 32-wide K-block
            const int kb = u >> 2, blk = (u >> 1) & 1, part = u & 1;
            const float* W = (layer < L) ? (w_upd + (size_t)layer * kH * 2 * kH) : w_e1;
            const float w = W[(size_t)row * 2 * kH + (blk == 0 ? kH : 0) + kb * 32 + kk];
            uint32_t hi_bits;
            asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi_bits) : "f"(w));
            const float hi = __uint_as_float(hi_bits);
            if (part == 0) {
                v = hi;
            } else {
                uint32_t lo_bits;
                asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lo_bits) : "f"(w - hi));
                v = __uint_as_float(lo_bits);
            }
        }
        packed[idx] = v;
    }
}

// ---- peer-memory exchange of a domain-decomposed grid (NVLink P2P stores, no NCCL on the step path) -------------
// The ranks' extended states live in symmetric memory (every rank has every other rank's buffer mapped).  After a step a
// rank stores its first / last `halo` cells straight into the ghost regions of its ring neighbours' extended states:
//   left neighbour's right ghosts  ext_left [b][ch][halo + owned + i] = ext[b][ch][halo + i]
//   right neighbour's left ghosts  ext_right[b][ch][i]               = ext[b][ch][owned + i]          i < halo
// for the channels ch0 <= ch < ch1.  One thread per (b, ch, side, i).
__global__ void __launch_bounds__(256) peer_halo_push_kernel(const float* __restrict__ ext, float* __restrict__ ext_left,
                                                             float* __restrict__ ext_right, int B, int owned, int halo,
                                                             int ch0, int ch1) {
    const int ld = owned + 2 * halo;
    const long long total = (long long)B * (ch1 - ch0) * 2 * halo;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(t % halo);
        const int side = (int)((t / halo) & 1);
        const long long bc = t / (2 * halo);
        const int ch = ch0 + (int)(bc % (ch1 - ch0));
        const long long b = bc / (ch1 - ch0);
        const size_t row = ((size_t)b * 3 + ch) * ld;
        if (side == 0) ext_left[row + halo + owned + i] = ext[row + halo + i];
        else ext_right[row + i] = ext[row + owned + i];
    }
}

// All-gather by peer stores: block p copies this rank's `bytes` (a multiple of 16) into slot `rank` of the gather
// buffer of rank p, which sits `offset` bytes into p's symmetric allocation (bases[p]).
__global__ void __launch_bounds__(256) peer_allgather_kernel(const uint4* __restrict__ src, long long bytes,
                                                             void* const* __restrict__ bases, long long offset, int rank) {
    uint4* dst = reinterpret_cast<uint4*>(static_cast<char*>(bases[blockIdx.x]) + offset + (long long)rank * bytes);
    for (long long i = threadIdx.x; i < bytes / 16; i += blockDim.x) dst[i] = src[i];
}

}  // namespace fluxgnn
