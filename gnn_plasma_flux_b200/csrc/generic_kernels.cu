// FluxGNN on the ring for architectures other than MODEL_CONFIG (src/flux_gnn.py:11-67 with any
// input_dim <= 16, hidden_dim in {16, 32, 64, 128}, 1..8 layers): the reference's own smoke test and
// notebooks build FluxGNN(4, 64, 3) (examples/smoke_test.py:50-56) and the class default is (2, 32, 2).
// Same algebra as the tuned kernel (hybrid_kernel_impl.cuh): per layer  h' = relu(Wa h + mean_nbr(Wb h) + b)
// with the neighbour mean taken over the +-radius window of the chain, edge readout
// w2 . relu(W1a h_row + W1b h_col + b1) + b2.  Plain FP32-pipe code, one CTA per 128-row window
// (activations in shared memory, weights K-major through the read-only cache): correct for every such
// model, not tuned -- the 128-wide MODEL_CONFIG network never takes this path.
#include <stdint.h>

#include "common.cuh"
#include "field_kernels.cuh"
#include "hybrid_kernel.cuh"
#include "tile_common.cuh"

namespace fluxgnn {

int launch_poisson_for_generic(const float* n, long long ns, float* E, long long es, const double* gtab, int B, int nx,
                               double length, void* fft_ws, cudaStream_t stream);      // api.cu

namespace {

constexpr int kGenThreads = 256;
constexpr int kGenRows = 128;
constexpr int kGenMaxF = 16;

// packed (floats), every matrix K-major ([k][n], n contiguous):
//   w_in [F][H], b_in [H], per layer { Wself [H][H], Wnbr [H][H], b [H] }, edge { W1a [H][H], W1b [H][H], b1 [H] }, w2 [H], b2 [+pad to 4]
__host__ __device__ inline size_t generic_floats(int F, int H, int L) {
    return (size_t)F * H + H + (size_t)(L + 1) * (2 * (size_t)H * H + H) + H + 4;
}

__global__ void generic_pack_kernel(const float* __restrict__ w_in, const float* __restrict__ b_in,
                                    const float* __restrict__ w_upd, const float* __restrict__ b_upd,
                                    const float* __restrict__ w_e1, const float* __restrict__ b_e1,
                                    const float* __restrict__ w_e2, const float* __restrict__ b_e2,
                                    int F, int H, int L, float* __restrict__ packed) {
    const size_t total = generic_floats(F, H, L);
    const size_t per = 2 * (size_t)H * H + H;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        size_t o = idx;
        float v = 0.f;
        if (o < (size_t)F * H) {
            v = w_in[(o % H) * F + o / H];                                       // [f][n] <- W_in[n][f]
        } else if ((o -= (size_t)F * H) < (size_t)H) {
            v = b_in[o];
        } else if ((o -= H) < per * (L + 1)) {
            const int l = (int)(o / per);
            const size_t q = o % per;
            const float* W = (l < L) ? w_upd + (size_t)l * H * 2 * H : w_e1;      // nn.Linear [H][2H]: [:, :H] self/row, [:, H:] nbr/col
            const float* bias = (l < L) ? b_upd + (size_t)l * H : b_e1;
            if (q < (size_t)H * H) v = W[(q % H) * 2 * H + q / H];
            else if (q < 2 * (size_t)H * H) v = W[((q - (size_t)H * H) % H) * 2 * H + H + (q - (size_t)H * H) / H];
            else v = bias[q - 2 * (size_t)H * H];
        } else if ((o -= per * (L + 1)) < (size_t)H) {
            v = w_e2[o];
        } else if (o == (size_t)H) {
            v = b_e2[0];
        }
        packed[idx] = v;
    }
}

// out[r][n] = sum_k in[r][k] * Wt[k][n] for 8 rows per sweep (activations read as 128-bit broadcasts).
template <int H>
__device__ __forceinline__ void gen_dense_rows(const float* __restrict__ in, const float* __restrict__ Wt, float* __restrict__ out,
                                               int n, int rgroup, int ngroups) {
    for (int r0 = rgroup * 8; r0 < kGenRows; r0 += ngroups * 8) {
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll 2
        for (int k = 0; k < H; k += 4) {
            const float w0 = __ldg(Wt + (size_t)(k + 0) * H + n), w1 = __ldg(Wt + (size_t)(k + 1) * H + n);
            const float w2 = __ldg(Wt + (size_t)(k + 2) * H + n), w3 = __ldg(Wt + (size_t)(k + 3) * H + n);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float4 a = *reinterpret_cast<const float4*>(in + (r0 + i) * H + k);
                acc[i] = fmaf(a.w, w3, fmaf(a.z, w2, fmaf(a.y, w1, fmaf(a.x, w0, acc[i]))));
            }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) out[(r0 + i) * H + n] = acc[i];
    }
}

struct GenericArgs {
    const float* packed;
    const float* feats;        // nullable [B][nx][F] (FluxGNN.forward's node_features)
    const float* state;        // nullable [B][3][nx] + x[nx] (F == 4: n, u, E, x)
    const float* x;
    float* flux_edges;         // nullable [B][2*hops*nx]
    float* face_flux;          // nullable [B][nx]
    int F, L, B, nx, radius, hops, halo, valid, tiles_per_ic;
};

// One CTA per window: tile row jl holds cell (t*valid - halo + jl) mod nx of its IC; rows [halo, halo+valid) are owned,
// the rest is recomputed halo (the receptive field of the network, L*radius + hops cells per side).  A window longer
// than the grid simply holds some cells twice (identical values), so every nx >= 1 takes the same path.
template <int H>
__global__ void __launch_bounds__(kGenThreads) generic_forward_kernel(GenericArgs a) {
    extern __shared__ __align__(16) float gsm[];
    float* h = gsm;                              // [128][H]
    float* P = h + kGenRows * H;                 // self / row half
    float* Q = P + kGenRows * H;                 // neighbour / col half
    const int tid = threadIdx.x, n = tid % H, rgroup = tid / H, ngroups = kGenThreads / H;
    const int ic = blockIdx.x / a.tiles_per_ic, t = blockIdx.x - ic * a.tiles_per_ic;
    const int nx = a.nx, F = a.F;
    const float* w_in = a.packed;
    const float* b_in = w_in + (size_t)F * H;
    const float* layers = b_in + H;
    const size_t per = 2 * (size_t)H * H + H;
    const float* w2 = layers + per * (a.L + 1);
    auto cell_of = [&](int jl) {
        long long g = (long long)t * a.valid - a.halo + jl;
        g %= nx;
        return (int)(g < 0 ? g + nx : g);
    };
    // input layer: h = relu(W_in x + b)                                   (src/flux_gnn.py:49)
    for (int r = rgroup; r < kGenRows; r += ngroups) {
        const int cell = cell_of(r);
        float v = __ldg(b_in + n);
        if (a.feats != nullptr) {
            const float* f = a.feats + ((size_t)ic * nx + cell) * F;
            for (int k = 0; k < F; ++k) v = fmaf(__ldg(w_in + (size_t)k * H + n), __ldg(f + k), v);
        } else {
            const float* s = a.state + (size_t)ic * 3 * nx + cell;
            v = fmaf(__ldg(w_in + n), __ldg(s), v);
            v = fmaf(__ldg(w_in + H + n), __ldg(s + nx), v);
            v = fmaf(__ldg(w_in + 2 * H + n), __ldg(s + 2 * (size_t)nx), v);
            v = fmaf(__ldg(w_in + 3 * H + n), __ldg(a.x + cell), v);
        }
        h[r * H + n] = fmaxf(v, 0.f);
    }
    __syncthreads();
    // message passing: every node has exactly 2*radius incoming edges on the ring          (src/flux_gnn.py:53-60)
    const float inv_deg = 1.0f / (float)(2 * a.radius);
    for (int l = 0; l < a.L; ++l) {
        const float* W = layers + per * l;
        gen_dense_rows<H>(h, W, P, n, rgroup, ngroups);
        gen_dense_rows<H>(h, W + (size_t)H * H, Q, n, rgroup, ngroups);
        __syncthreads();
        const float b = __ldg(W + 2 * (size_t)H * H + n);
        for (int r = rgroup; r < kGenRows; r += ngroups) {
            float s = 0.f;
            for (int k = 1; k <= a.radius; ++k) {
                const int lo = r - k, hi = r + k;             // outside the window: only rows whose value is never used
                s += (lo >= 0 ? Q[lo * H + n] : 0.f) + (hi < kGenRows ? Q[hi * H + n] : 0.f);
            }
            h[r * H + n] = fmaxf(P[r * H + n] + s * inv_deg + b, 0.f);
        }
        __syncthreads();
    }
    // edge readout: flux(row -> col) = w2 . relu(W1a h_row + W1b h_col + b1) + b2             (src/flux_gnn.py:63-66)
    const float* We = layers + per * a.L;
    gen_dense_rows<H>(h, We, P, n, rgroup, ngroups);
    gen_dense_rows<H>(h, We + (size_t)H * H, Q, n, rgroup, ngroups);
    __syncthreads();
    const float* b1 = We + 2 * (size_t)H * H;
    const float b2 = __ldg(w2 + H);
    const int warp = tid >> 5, lane = tid & 31, warps = kGenThreads >> 5;
    for (int jl = a.halo + warp; jl < a.halo + a.valid; jl += warps) {
        if ((long long)t * a.valid + (jl - a.halo) >= nx) break;             // last window of an IC: beyond the grid
        const int cell = cell_of(jl);
        for (int k = 1; k <= a.hops; ++k) {
            float sf = 0.f, sb = 0.f;
            for (int m = lane; m < H; m += 32) {
                const float w = __ldg(w2 + m), bb = __ldg(b1 + m);
                sf = fmaf(w, fmaxf(P[jl * H + m] + Q[(jl + k) * H + m] + bb, 0.f), sf);      // edge cell -> cell+k
                sb = fmaf(w, fmaxf(P[(jl + k) * H + m] + Q[jl * H + m] + bb, 0.f), sb);      // edge cell+k -> cell
            }
            for (int o = 16; o > 0; o >>= 1) {
                sf += __shfl_xor_sync(0xffffffffu, sf, o);
                sb += __shfl_xor_sync(0xffffffffu, sb, o);
            }
            if (lane == 0) {
                sf += b2;
                sb += b2;
                if (a.flux_edges != nullptr) {
                    float* fe = a.flux_edges + (size_t)ic * 2 * a.hops * nx + (size_t)2 * (k - 1) * nx + cell;
                    fe[0] = sf;
                    fe[nx] = sb;
                }
                if (k == 1 && a.face_flux != nullptr) a.face_flux[(size_t)ic * nx + cell] = 0.5f * (sf + sb);   // src/hybrid_solver.py:45-48
            }
        }
    }
}

// n' = n - c (F_i - F_{i-1}),  u' = u - c (u_i^2/2 - u_{i-1}^2/2) + dt E          (src/hybrid_solver.py:51-58)
__global__ void __launch_bounds__(256) generic_fv_kernel(const float* __restrict__ in, const float* __restrict__ face,
                                                         float* __restrict__ out, int B, int nx, float c, float dt) {
    HybridArgs a{};
    a.c = c;
    a.dt = dt;
    const long long total = (long long)B * nx;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int ic = (int)(idx / nx), i = (int)(idx - (long long)ic * nx), im = (i == 0) ? nx - 1 : i - 1;
        const float* s = in + (size_t)ic * 3 * nx;
        const float* f = face + (size_t)ic * nx;
        float n_new, u_new;
        tile_fv_update_values(a, s[i], s[nx + i], s[nx + im], s[2 * (size_t)nx + i], f[i], f[im], n_new, u_new);
        float* o = out + (size_t)ic * 3 * nx;
        o[i] = n_new;
        o[nx + i] = u_new;
    }
}

bool generic_hidden_ok(int H) { return H == 16 || H == 32 || H == 64 || H == 128; }

int generic_check(int F, int H, int L, int B, int nx, int radius, int hops) {
    if (F < 1 || F > kGenMaxF || !generic_hidden_ok(H) || L < 1 || L > kMaxL)
        return set_error(FLUXGNN_EUNSUP, "generic FluxGNN kernel: input_dim 1..%d, hidden_dim in {16,32,64,128}, 1..%d layers; "
                                         "got (%d, %d, %d)", kGenMaxF, kMaxL, F, H, L);
    if (B < 1 || nx < 1 || radius < 1 || hops < 1 || hops > radius || hops > kMaxHops)
        return set_error(FLUXGNN_EINVAL, "generic FluxGNN kernel: bad shape (B=%d nx=%d radius=%d hops=%d)", B, nx, radius, hops);
    if (kGenRows - 2 * (L * radius + hops) < 8)
        return set_error(FLUXGNN_EUNSUP, "receptive field L*radius+hops = %d cells does not fit a %d-cell tile", L * radius + hops, kGenRows);
    return FLUXGNN_OK;
}

int generic_launch(const void* packed, int F, int H, int L, const float* feats, const float* state, const float* x, int B,
                   int nx, int radius, int hops, float* flux_edges, float* face_flux, cudaStream_t stream) {
    GenericArgs a{};
    a.packed = (const float*)packed;
    a.feats = feats; a.state = state; a.x = x;
    a.flux_edges = flux_edges; a.face_flux = face_flux;
    a.F = F; a.L = L; a.B = B; a.nx = nx; a.radius = radius; a.hops = hops;
    a.halo = L * radius + hops;
    a.valid = kGenRows - 2 * a.halo;
    a.tiles_per_ic = (nx + a.valid - 1) / a.valid;
    const long long grid = (long long)B * a.tiles_per_ic;
    if (grid > 0x7fffffffLL) return set_error(FLUXGNN_EINVAL, "too many tiles");
    const size_t smem = (size_t)3 * kGenRows * H * sizeof(float);
#define FLUXGNN_GEN(HH)                                                                                                     \
    case HH:                                                                                                                \
        FLUXGNN_CUDA_OK(cudaFuncSetAttribute(generic_forward_kernel<HH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        generic_forward_kernel<HH><<<(unsigned)grid, kGenThreads, smem, stream>>>(a);                                       \
        break;
    switch (H) {
        FLUXGNN_GEN(16) FLUXGNN_GEN(32) FLUXGNN_GEN(64) FLUXGNN_GEN(128)
        default: return set_error(FLUXGNN_EUNSUP, "generic FluxGNN kernel: hidden_dim %d", H);
    }
#undef FLUXGNN_GEN
    FLUXGNN_CUDA_OK(cudaGetLastError());
    count_launch();
    return FLUXGNN_OK;
}

}  // namespace
}  // namespace fluxgnn

using namespace fluxgnn;

extern "C" {

size_t fluxgnn_generic_packed_bytes(int input_dim, int hidden, int num_layers) {
    if (input_dim < 1 || input_dim > kGenMaxF || !generic_hidden_ok(hidden) || num_layers < 1 || num_layers > kMaxL) return 0;
    return generic_floats(input_dim, hidden, num_layers) * sizeof(float);
}

int fluxgnn_generic_pack(const float* w_in, const float* b_in, const float* w_upd, const float* b_upd, const float* w_e1,
                         const float* b_e1, const float* w_e2, const float* b_e2, int input_dim, int hidden, int num_layers,
                         void* packed, void* stream) {
    if (fluxgnn_generic_packed_bytes(input_dim, hidden, num_layers) == 0)
        return set_error(FLUXGNN_EUNSUP, "generic FluxGNN kernel: unsupported architecture (%d, %d, %d)", input_dim, hidden, num_layers);
    if (!w_in || !b_in || !w_upd || !b_upd || !w_e1 || !b_e1 || !w_e2 || !b_e2 || !packed)
        return set_error(FLUXGNN_EINVAL, "null weight pointer");
    const size_t total = generic_floats(input_dim, hidden, num_layers);
    generic_pack_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(w_in, b_in, w_upd, b_upd, w_e1, b_e1, w_e2,
                                                                                         b_e2, input_dim, hidden, num_layers,
                                                                                         (float*)packed);
    FLUXGNN_CUDA_OK(cudaGetLastError());
    count_launch();
    return FLUXGNN_OK;
}

int fluxgnn_generic_forward_ring(const void* packed, int input_dim, int hidden, int num_layers, const float* feats,
                                 const float* state, const float* x, int B, int nx, int radius, int hops, float* flux_edges,
                                 float* face_flux, void* stream) {
    int rc = generic_check(input_dim, hidden, num_layers, B, nx, radius, hops);
    if (rc != FLUXGNN_OK) return rc;
    if (!packed || (!feats && (!state || !x))) return set_error(FLUXGNN_EINVAL, "generic_forward_ring: null pointer");
    if (!feats && input_dim != 4) return set_error(FLUXGNN_EINVAL, "generic_forward_ring: a [n,u,E] state + x needs input_dim 4");
    if (!flux_edges && !face_flux) return set_error(FLUXGNN_EINVAL, "generic_forward_ring: no output requested");
    return generic_launch(packed, input_dim, hidden, num_layers, feats, state, x, B, nx, radius, hops, flux_edges, face_flux,
                          (cudaStream_t)stream);
}

// workspace = [face flux B*nx][state ping-pong B*3*nx][field-solve scratch]
size_t fluxgnn_generic_workspace_bytes(int B, int nx) {
    if (B < 1 || nx < 1) return 0;
    return (size_t)B * 4 * nx * sizeof(float) + fluxgnn_poisson_workspace_bytes(B, nx);
}

int fluxgnn_generic_hybrid_rollout(const void* packed, int hidden, int num_layers, const float* state_in, float* state_out,
                                   const float* x, const double* gtab, int B, int nx, double length, int radius, float c,
                                   float dt, int steps, int record_every, float* traj, void* workspace, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    int rc = generic_check(4, hidden, num_layers, B, nx, radius, 1);
    if (rc != FLUXGNN_OK) return rc;
    if (!packed || !state_in || !state_out || !x || !workspace || state_in == state_out || steps < 1 || !(length > 0.0))
        return set_error(FLUXGNN_EINVAL, "generic_hybrid_rollout: bad argument");
    if (traj && record_every < 1) return set_error(FLUXGNN_EINVAL, "generic_hybrid_rollout: record_every must be >= 1");
    float* face = (float*)workspace;
    float* pong = face + (size_t)B * nx;
    void* fft_ws = pong + (size_t)B * 3 * nx;
    const size_t state_floats = (size_t)B * 3 * nx;
    const long long cells = (long long)B * nx;
    const unsigned blocks = (unsigned)((cells + 255) / 256 < 148 * 16 ? (cells + 255) / 256 : 148 * 16);
    const float* src = state_in;
    for (int t = 0; t < steps; ++t) {
        float* dst = ((steps - 1 - t) % 2 == 0) ? state_out : pong;
        rc = generic_launch(packed, 4, hidden, num_layers, nullptr, src, x, B, nx, radius, 1, nullptr, face, stream);
        if (rc != FLUXGNN_OK) return rc;
        generic_fv_kernel<<<blocks, 256, 0, stream>>>(src, face, dst, B, nx, c, dt);
        FLUXGNN_CUDA_OK(cudaGetLastError());
        count_launch();
        rc = launch_poisson_for_generic(dst, 3LL * nx, dst + 2 * (size_t)nx, 3LL * nx, gtab, B, nx, length, fft_ws, stream);
        if (rc != FLUXGNN_OK) return rc;
        if (traj && (t + 1) % record_every == 0)
            FLUXGNN_CUDA_OK(cudaMemcpyAsync(traj + (size_t)((t + 1) / record_every - 1) * state_floats, dst,
                                            state_floats * sizeof(float), cudaMemcpyDeviceToDevice, stream));
        src = dst;
    }
    return FLUXGNN_OK;
}

}  // extern "C"
