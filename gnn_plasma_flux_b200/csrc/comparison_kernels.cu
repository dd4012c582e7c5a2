// The reference's two comparison models as fused forward kernels:
//   PureGNN  scripts/training/train_pure_gnn.py:35-76   end-to-end GNN, Tanh, residual edge messages,
//                                                        3 outputs per node;  state' = state + delta
//            rolled out as in scripts/evaluation/benchmark_timing.py:129-143
//   PINN     scripts/training/train_pinn.py:36-61        dense Tanh MLP on the flattened state, residual
//            rolled out as in scripts/evaluation/benchmark_timing.py:186-189
// They are not on the hybrid hot path; the kernels are plain FP32-pipe code kept small: one CTA per
// IC with the activations in shared memory (PureGNN: a whole multi-step rollout per launch), and a
// tiled dense layer (PINN).  Weights are read K-major so that a warp's loads are coalesced.
#include <stdint.h>

#include "common.cuh"

namespace fluxgnn {

namespace {

constexpr int kCmpThreads = 256;
constexpr int kCmpMaxNx = 128;

// Packed PureGNN weights (floats), everything K-major ([k][n], n contiguous):
//   w_in [4][H], b_in [H], per layer { Wsrc [H][H], Wdst [H][H], b [H] }, W_o1 [H][H], b_o1 [H], W_o2 [H][3] (+pad), b_o2 [4]
__host__ __device__ inline size_t pure_gnn_floats(int H, int L) {
    return (size_t)4 * H + H + (size_t)L * (2 * H * H + H) + (size_t)H * H + H + (size_t)H * 4 + 4;
}

__global__ void pure_gnn_pack_kernel(const float* __restrict__ w_in, const float* __restrict__ b_in,
                                     const float* __restrict__ w_upd, const float* __restrict__ b_upd,
                                     const float* __restrict__ w_o1, const float* __restrict__ b_o1,
                                     const float* __restrict__ w_o2, const float* __restrict__ b_o2,
                                     int H, int L, float* __restrict__ packed) {
    const size_t total = pure_gnn_floats(H, L);
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        size_t o = idx;
        float v = 0.f;
        if (o < (size_t)4 * H) {
            v = w_in[(o % H) * 4 + o / H];                                  // [f][n] <- W_in[n][f]
        } else if ((o -= (size_t)4 * H) < (size_t)H) {
            v = b_in[o];
        } else {
            o -= H;
            const size_t per = (size_t)2 * H * H + H;
            if (o < per * L) {
                const int l = (int)(o / per);
                const size_t q = o % per;
                const float* W = w_upd + (size_t)l * H * 2 * H;             // nn.Linear [H][2H]: [:, :H] src, [:, H:] dst
                if (q < (size_t)H * H) v = W[(q % H) * 2 * H + q / H];                       // Wsrc^T [k][n]
                else if (q < (size_t)2 * H * H) v = W[((q - (size_t)H * H) % H) * 2 * H + H + (q - (size_t)H * H) / H];
                else v = b_upd[(size_t)l * H + (q - (size_t)2 * H * H)];
            } else {
                o -= per * L;
                if (o < (size_t)H * H) v = w_o1[(o % H) * H + o / H];
                else if ((o -= (size_t)H * H) < (size_t)H) v = b_o1[o];
                else if ((o -= H) < (size_t)H * 4) v = (o % 4 < 3) ? w_o2[(o % 4) * H + o / 4] : 0.f;   // [k][4] <- W_o2[c][k]
                else v = ((o - (size_t)H * 4) < 3) ? b_o2[o - (size_t)H * 4] : 0.f;
            }
        }
        packed[idx] = v;
    }
}

// out[r][n] = sum_k in[r][k] * Wt[k][n]  for the rows of this thread's row group; 8 rows per sweep.
template <int H>
__device__ __forceinline__ void dense_rows(const float* __restrict__ in, const float* __restrict__ Wt, float* __restrict__ out,
                                           int nx, int n, int rgroup, int ngroups) {
    for (int r0 = rgroup * 8; r0 < nx; r0 += ngroups * 8) {
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll 2
        for (int k = 0; k < H; k += 4) {                            // 128-bit broadcast reads of the activations
            const float w0 = __ldg(Wt + (size_t)(k + 0) * H + n), w1 = __ldg(Wt + (size_t)(k + 1) * H + n);
            const float w2 = __ldg(Wt + (size_t)(k + 2) * H + n), w3 = __ldg(Wt + (size_t)(k + 3) * H + n);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float4 a = *reinterpret_cast<const float4*>(in + (r0 + i) * H + k);
                acc[i] = fmaf(a.w, w3, fmaf(a.z, w2, fmaf(a.y, w1, fmaf(a.x, w0, acc[i]))));
            }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i)
            if (r0 + i < nx) out[(r0 + i) * H + n] = acc[i];
    }
}

// One CTA per IC; nx <= 128 (rows padded to a multiple of 8 in shared memory).
template <int H>
__global__ void __launch_bounds__(kCmpThreads) pure_gnn_rollout_kernel(const float* __restrict__ packed,
                                                                      const float* __restrict__ state_in,
                                                                      float* __restrict__ state_out,
                                                                      const float* __restrict__ x, int nx, int L, int steps,
                                                                      int delta_only) {
    extern __shared__ float cmp_smem[];
    const int rows = (nx + 7) & ~7;
    float* h = cmp_smem;                        // [rows][H]
    float* P = h + rows * H;                    // Wsrc h
    float* Q = P + rows * H;                    // Wdst h
    float* st = Q + rows * H;                   // [4][rows]: n, u, E, x
    const int tid = threadIdx.x, n = tid % H, rgroup = tid / H, ngroups = kCmpThreads / H;
    const float* src = state_in + (size_t)blockIdx.x * 3 * nx;
    for (int i = tid; i < 3 * nx; i += kCmpThreads) st[(i / nx) * rows + i % nx] = src[i];
    for (int i = tid; i < nx; i += kCmpThreads) st[3 * rows + i] = x[i];
    for (int i = tid; i < (rows - nx) * H; i += kCmpThreads) h[nx * H + i] = 0.f;    // padding rows stay finite
    __syncthreads();
    const float* w_in = packed;
    const float* b_in = w_in + 4 * H;
    const float* layers = b_in + H;
    const size_t per = (size_t)2 * H * H + H;
    const float* w_o1 = layers + per * L;
    const float* b_o1 = w_o1 + (size_t)H * H;
    const float* w_o2 = b_o1 + H;
    const float* b_o2 = w_o2 + (size_t)H * 4;

    for (int step = 0; step < steps; ++step) {
        // input layer: h = tanh(W_in [n,u,E,x] + b)            (train_pure_gnn.py:60)
        for (int r = rgroup; r < nx; r += ngroups) {
            float v = __ldg(b_in + n);
#pragma unroll
            for (int f = 0; f < 4; ++f) v = fmaf(__ldg(w_in + f * H + n), st[f * rows + r], v);
            h[r * H + n] = tanhf(v);
        }
        __syncthreads();
        // message passing: h_j += tanh(Wsrc h_{j-1} + Wdst h_j + b) + tanh(Wsrc h_{j+1} + Wdst h_j + b)   (:63-72)
        for (int l = 0; l < L; ++l) {
            const float* Wsrc = layers + per * l;
            dense_rows<H>(h, Wsrc, P, nx, n, rgroup, ngroups);
            dense_rows<H>(h, Wsrc + (size_t)H * H, Q, nx, n, rgroup, ngroups);
            __syncthreads();
            const float b = __ldg(Wsrc + (size_t)2 * H * H + n);
            for (int r = rgroup; r < nx; r += ngroups) {
                const int rm = (r == 0) ? nx - 1 : r - 1, rp = (r == nx - 1) ? 0 : r + 1;
                const float q = Q[r * H + n] + b;
                // index_add_ order: the edge from j-1 (first edge block), then the edge from j+1
                const float upd = tanhf(P[rm * H + n] + q) + tanhf(P[rp * H + n] + q);
                h[r * H + n] += upd;
            }
            __syncthreads();
        }
        // output MLP: delta = W_o2 tanh(W_o1 h + b_o1) + b_o2;  state += delta          (:75, benchmark_timing.py:138-142)
        dense_rows<H>(h, w_o1, P, nx, n, rgroup, ngroups);
        __syncthreads();
        for (int r = rgroup; r < nx; r += ngroups) P[r * H + n] = tanhf(P[r * H + n] + __ldg(b_o1 + n));
        __syncthreads();
        for (int w = tid; w < 3 * nx; w += kCmpThreads) {
            const int c = w / nx, r = w % nx;
            float acc = 0.f;
            for (int k = 0; k < H; ++k) acc = fmaf(P[r * H + k], __ldg(w_o2 + k * 4 + c), acc);
            const float delta = acc + __ldg(b_o2 + c);
            st[c * rows + r] = delta_only ? delta : st[c * rows + r] + delta;      // delta_only: PureGNN.forward's own output
        }
        __syncthreads();
    }
    float* dst = state_out + (size_t)blockIdx.x * 3 * nx;
    for (int i = tid; i < 3 * nx; i += kCmpThreads) dst[i] = st[(i / nx) * rows + i % nx];
}

// y[r][n] = act(sum_k in[r][k] W[n][k] + bias[n]) (+ residual[r][n]);  W in nn.Linear layout [N][K].
// Tile: 8 rows x 128 outputs per CTA, K in chunks of 32 staged (transposed) through shared memory.
__global__ void __launch_bounds__(128) dense_layer_kernel(const float* __restrict__ in, const float* __restrict__ W,
                                                          const float* __restrict__ bias, const float* __restrict__ residual,
                                                          float* __restrict__ out, int rows, int K, int N, int act) {
    __shared__ float Ws[32][129];
    __shared__ float Xs[8][32];
    const int tid = threadIdx.x, n0 = blockIdx.x * 128, r0 = blockIdx.y * 8;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int k0 = 0; k0 < K; k0 += 32) {
        for (int w = tid; w < 128 * 32; w += 128) {                 // coalesced along k
            const int nn = w >> 5, kk = w & 31;
            Ws[kk][nn] = (n0 + nn < N && k0 + kk < K) ? __ldg(W + (size_t)(n0 + nn) * K + k0 + kk) : 0.f;
        }
        for (int w = tid; w < 8 * 32; w += 128) {
            const int rr = w >> 5, kk = w & 31;
            Xs[rr][kk] = (r0 + rr < rows && k0 + kk < K) ? in[(size_t)(r0 + rr) * K + k0 + kk] : 0.f;
        }
        __syncthreads();
#pragma unroll 8
        for (int kk = 0; kk < 32; ++kk) {
            const float w = Ws[kk][tid];
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] = fmaf(Xs[i][kk], w, acc[i]);
        }
        __syncthreads();
    }
    const int n = n0 + tid;
    if (n >= N) return;
    const float b = bias ? __ldg(bias + n) : 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        if (r0 + i >= rows) break;
        float v = acc[i] + b;
        if (act == 1) v = tanhf(v);
        if (residual) v += residual[(size_t)(r0 + i) * N + n];
        out[(size_t)(r0 + i) * N + n] = v;
    }
}

}  // namespace

}  // namespace fluxgnn

using namespace fluxgnn;

extern "C" {

size_t fluxgnn_pure_gnn_packed_bytes(int hidden, int num_layers) {
    if ((hidden != 64 && hidden != 128) || num_layers < 1 || num_layers > 8) return 0;
    return pure_gnn_floats(hidden, num_layers) * sizeof(float);
}

int fluxgnn_pure_gnn_pack(const float* w_in, const float* b_in, const float* w_upd, const float* b_upd,
                          const float* w_o1, const float* b_o1, const float* w_o2, const float* b_o2,
                          int hidden, int num_layers, void* packed, void* stream) {
    if ((hidden != 64 && hidden != 128) || num_layers < 1 || num_layers > 8)
        return set_error(FLUXGNN_EUNSUP, "pure_gnn_pack: hidden must be 64 or 128 and num_layers 1..8 (got %d, %d)", hidden,
                         num_layers);
    if (!w_in || !b_in || !w_upd || !b_upd || !w_o1 || !b_o1 || !w_o2 || !b_o2 || !packed)
        return set_error(FLUXGNN_EINVAL, "pure_gnn_pack: null pointer");
    const size_t total = pure_gnn_floats(hidden, num_layers);
    pure_gnn_pack_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        w_in, b_in, w_upd, b_upd, w_o1, b_o1, w_o2, b_o2, hidden, num_layers, (float*)packed);
    FLUXGNN_CUDA_OK(cudaGetLastError());
    count_launch();
    return FLUXGNN_OK;
}

static int pure_gnn_launch(const void* packed, int hidden, int num_layers, const float* state_in, float* state_out,
                           const float* x, int B, int nx, int steps, int delta_only, void* stream) {
    if (!packed || !state_in || !state_out || !x) return set_error(FLUXGNN_EINVAL, "pure_gnn_rollout: null pointer");
    if ((hidden != 64 && hidden != 128) || num_layers < 1 || num_layers > 8)
        return set_error(FLUXGNN_EUNSUP, "pure_gnn_rollout: hidden must be 64 or 128 and num_layers 1..8");
    if (B < 1 || nx < 3 || steps < 1) return set_error(FLUXGNN_EINVAL, "pure_gnn_rollout: B=%d nx=%d steps=%d", B, nx, steps);
    if (nx > kCmpMaxNx) return set_error(FLUXGNN_EUNSUP, "pure_gnn_rollout: nx must be <= %d, got %d", kCmpMaxNx, nx);
    const int rows = (nx + 7) & ~7;
    const size_t smem = ((size_t)3 * rows * hidden + 4 * rows) * sizeof(float);
    if (hidden == 128) {
        FLUXGNN_CUDA_OK(cudaFuncSetAttribute(pure_gnn_rollout_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        pure_gnn_rollout_kernel<128><<<B, kCmpThreads, smem, (cudaStream_t)stream>>>((const float*)packed, state_in, state_out,
                                                                                   x, nx, num_layers, steps, delta_only);
    } else {
        FLUXGNN_CUDA_OK(cudaFuncSetAttribute(pure_gnn_rollout_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        pure_gnn_rollout_kernel<64><<<B, kCmpThreads, smem, (cudaStream_t)stream>>>((const float*)packed, state_in, state_out, x,
                                                                                  nx, num_layers, steps, delta_only);
    }
    FLUXGNN_CUDA_OK(cudaGetLastError());
    count_launch();
    return FLUXGNN_OK;
}

int fluxgnn_pure_gnn_rollout(const void* packed, int hidden, int num_layers, const float* state_in, float* state_out,
                             const float* x, int B, int nx, int steps, void* stream) {
    return pure_gnn_launch(packed, hidden, num_layers, state_in, state_out, x, B, nx, steps, 0, stream);
}

int fluxgnn_pure_gnn_delta(const void* packed, int hidden, int num_layers, const float* state, float* delta_out,
                           const float* x, int B, int nx, void* stream) {
    return pure_gnn_launch(packed, hidden, num_layers, state, delta_out, x, B, nx, 1, 1, stream);
}

int fluxgnn_dense_layer(const float* in, const float* weight, const float* bias, const float* residual, float* out,
                        int rows, int in_features, int out_features, int activation, void* stream) {
    if (!in || !weight || !out) return set_error(FLUXGNN_EINVAL, "dense_layer: null pointer");
    if (rows < 1 || in_features < 1 || out_features < 1 || (activation != 0 && activation != 1))
        return set_error(FLUXGNN_EINVAL, "dense_layer: rows=%d in=%d out=%d activation=%d", rows, in_features, out_features,
                         activation);
    if (out == in || out == residual) return set_error(FLUXGNN_EINVAL, "dense_layer: out must not alias in / residual");
    dim3 grid((unsigned)((out_features + 127) / 128), (unsigned)((rows + 7) / 8));
    dense_layer_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(in, weight, bias, residual, out, rows, in_features,
                                                               out_features, activation);
    FLUXGNN_CUDA_OK(cudaGetLastError());
    count_launch();
    return FLUXGNN_OK;
}

}  // extern "C"
