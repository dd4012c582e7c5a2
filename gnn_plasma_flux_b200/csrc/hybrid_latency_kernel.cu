// Latency mode of the fused hybrid step (sm_100a): ONE 128-row tile of whole ICs is computed by a thread-block
// cluster of 8 CTAs that split the OUTPUT FEATURES of every layer, for rollouts of so few ICs that the tile kernel
// would leave most SMs idle (the reference's own timing protocol is one IC of 64 cells,
// scripts/evaluation/benchmark_timing.py:63-72; BASELINE.json configs[0] is 20 ICs = 10 tiles).
//
//   src/flux_gnn.py:49        input MLP              every CTA computes all 128 features itself (no exchange)
//   src/flux_gnn.py:53-60     message passing        CTA r: the 16 output features of tx = 2r, 2r+1 (the tile kernel's thread
//                                                    columns, features tx + 16 j), both halves  Z = W[:, H:] h,  Y = W[:, :H] h;
//                                                    h' of its features is stored into the next-layer buffer of ALL 8 CTAs
//                                                    through distributed shared memory (st.shared::cluster), one cluster
//                                                    barrier per layer, activations double-buffered
//   src/flux_gnn.py:63-66     edge readout           per-CTA partial dot products, exchanged as 2 x 128 floats per CTA
//   src/hybrid_solver.py:45-58, src/baseline_solver.py:59-68   update and field solve: every CTA redundantly, CTA 0 writes
//
// A product thread (the first 128 of 256) owns TWO rows and the 8 features of one tx -- the weight operand is a warp-wide
// broadcast, so two rows per thread halve its shared-memory traffic --, and every sum runs in the tile kernel's order (k = 0..127 per
// accumulator; window sums hop by hop; the edge readout's chain over j, then the pair tx even + tx odd, then the
// tile kernel's butterfly tree across the CTAs), so the result is BIT-IDENTICAL to hybrid_tile_kernel -- the tests
// compare with torch.equal.  Weights: each CTA stages its 16 KiB slice of a layer (2 halves x 128 k x 16 columns) from
// the packed stream (L2-resident) into shared memory, the next layer's slice is prefetched into registers under the
// current product.
#include "common.cuh"
#include "hybrid_kernel.cuh"
#include "tile_common.cuh"

namespace fluxgnn {

namespace {

constexpr int kLatCluster = 8;
constexpr int kLatThreads = 256;
constexpr int kLatFeat = kH / kLatCluster;        // 16 output features per CTA
constexpr int kLatMaxR = 8;                       // neighbour rows kept in registers; larger radii are not dispatched here

struct __align__(128) LatSmem {
    float Hs[2][kH * kTileRows];                  // activations [buffer][feature][row], a complete copy in every CTA
    float Zs[kLatFeat * kTileRows];               // neighbour half of this CTA's features [local feature][row]
    float Wl[2][2][kH][kLatFeat];                 // weight slice [buffer][half: 0 = neighbour/col, 1 = self/row][k][column]
    float edge_all[kLatCluster][2][kTileRows];    // per source CTA: fwd / bwd partial sums (tx even + tx odd) of every row
    float pair[2][kTileRows];                     // tx odd -> tx even hand-over of the edge readout
    float sN[kTileRows], sU[kTileRows], sE[kTileRows], sX[kTileRows];
    float sF[kTileRows], sRho[kTileRows];
    double gtab[kTileRows];
    int rowIC[kTileRows];
    int rowCell[kTileRows];
    short prevRow[kTileRows], nextRow[kTileRows];
};
static_assert(sizeof(LatSmem) <= 227 * 1024, "latency tile does not fit shared memory");

__device__ __forceinline__ void st_cluster_f32(uint32_t addr, float v) {
    asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}

__global__ void __launch_bounds__(kLatThreads, 1) hybrid_latency_kernel(const HybridArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    LatSmem& S = *reinterpret_cast<LatSmem*>(smem_raw);
    const int tid = threadIdx.x;
    const bool gemm = tid < kTileRows;                     // product / epilogue threads: rows r0 and r0 + 64, thread column txl
    const int r0 = tid & 63, txl = (tid >> 6) & 1;
    const int row = tid & (kTileRows - 1);                 // the row a thread looks after outside the products
    const int crank = (int)cluster_ctarank();
    const int tx = 2 * crank + txl;                        // the tile kernel's thread column whose 8 features this thread owns
    const int nx = a.nx, radius = a.radius;
    const float inv_deg = 1.0f / (float)(2 * radius);
    const TileRows T{S.sN, S.sU, S.sE, S.sX, S.sF, S.sRho, S.gtab, S.rowIC, S.rowCell, S.prevRow, S.nextRow};
    const int first_tile = (int)blockIdx.x / kLatCluster, tile_stride = (int)gridDim.x / kLatCluster;
    const int used_rows = a.ics_per_tile * nx;
    const float* stream = a.packed + SmallParams::count;

    // shared::cluster addresses of the peers' activation buffers and edge tables
    uint32_t peer_hs[kLatCluster], peer_edge[kLatCluster];
#pragma unroll
    for (int p = 0; p < kLatCluster; ++p) {
        peer_hs[p] = cluster_map(smem_u32(&S.Hs[0][0]), p);
        peer_edge[p] = cluster_map(smem_u32(&S.edge_all[crank][0][0]), p);
    }
    for (int i = tid; i < nx; i += kLatThreads) S.gtab[i] = a.gtab[i];

    // this CTA's slice of a layer's weights: 1024 float4, four per thread
    float4 wreg[4];
    auto fetch_weights = [&](int layer) {
        const float* base = stream + (size_t)layer * kLayerFloats;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int idx = tid + kLatThreads * i, half = idx >> 9, rem = idx & 511, k = rem >> 2, piece = rem & 3;
            const int col = ((piece & 1) ? 64 : 0) + 4 * (2 * crank + (piece >> 1));
            wreg[i] = __ldg(reinterpret_cast<const float4*>(base + (size_t)half * kHalfFloats + k * kH + col));
        }
    };
    auto store_weights = [&](int buf) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int idx = tid + kLatThreads * i, half = idx >> 9, rem = idx & 511, k = rem >> 2, piece = rem & 3;
            *reinterpret_cast<float4*>(&S.Wl[buf][half][k][(piece >> 1) * 8 + (piece & 1) * 4]) = wreg[i];
        }
    };
    cluster_sync_all();                                    // every CTA of the cluster is running before remote stores start

    for (int tile = first_tile; tile < a.num_tiles; tile += tile_stride) {
        if (tid < kTileRows) tile_load_row(a, T, tile, true, tid, tid, 0, kTileRows);
        fetch_weights(0);
        __syncthreads();
        // neighbour rows of this thread's two rows, hop by hop (periodic inside the IC)
        int rp[2][kLatMaxR], rm[2][kLatMaxR];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            int p = r0 + 64 * q, m = r0 + 64 * q;
#pragma unroll
            for (int k = 0; k < kLatMaxR; ++k) {
                if (k < radius) { p = S.nextRow[p]; m = S.prevRow[m]; }
                rp[q][k] = p; rm[q][k] = m;
            }
        }

        for (int step = 0; step < a.steps; ++step) {
            // ---- input MLP (src/flux_gnn.py:49): all 128 features of this thread's row, 64 per thread ----
            {
                const float fn = S.sN[row], fu = S.sU[row], fe = S.sE[row], fx = S.sX[row];
                for (int f = (tid >> 7) * 64; f < (tid >> 7) * 64 + 64; ++f) {
                    float v = fmaf(__ldg(a.packed + SmallParams::w_in + 0 * kH + f), fn, __ldg(a.packed + SmallParams::b_in + f));
                    v = fmaf(__ldg(a.packed + SmallParams::w_in + 1 * kH + f), fu, v);
                    v = fmaf(__ldg(a.packed + SmallParams::w_in + 2 * kH + f), fe, v);
                    v = fmaf(__ldg(a.packed + SmallParams::w_in + 3 * kH + f), fx, v);
                    S.Hs[0][f * kTileRows + row] = fmaxf(v, 0.f);
                }
            }
            store_weights(0);                              // layer 0's slice (fetched before the step / under the last layer)
            __syncthreads();

            int cur = 0;
            for (int layer = 0; layer <= a.L; ++layer, cur ^= 1) {
                const int wb = layer & 1;
                // prefetch the next slice: the next layer's, or layer 0's for the next step
                if (layer < a.L) fetch_weights(layer + 1);
                else if (step + 1 < a.steps) fetch_weights(0);
                const float* bias = a.packed + (layer < a.L ? SmallParams::b_upd + layer * kH : SmallParams::b_e1);
                float accZ[2][8], accY[2][8];
                if (gemm) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float b = __ldg(bias + tx + 16 * j);
                        accZ[0][j] = 0.f; accZ[1][j] = 0.f; accY[0][j] = b; accY[1][j] = b;
                    }
                    const float* hrow = &S.Hs[cur][r0];
                    const float* wz = &S.Wl[wb][0][0][txl * 8];
                    const float* wy = &S.Wl[wb][1][0][txl * 8];
#pragma unroll 4
                    for (int k = 0; k < kH; ++k) {
                        const float h0 = hrow[k * kTileRows], h1 = hrow[k * kTileRows + 64];
                        const float4 z0 = *reinterpret_cast<const float4*>(wz + k * kLatFeat);
                        const float4 z1 = *reinterpret_cast<const float4*>(wz + k * kLatFeat + 4);
                        const float4 y0 = *reinterpret_cast<const float4*>(wy + k * kLatFeat);
                        const float4 y1 = *reinterpret_cast<const float4*>(wy + k * kLatFeat + 4);
                        const float wzv[8] = {z0.x, z0.y, z0.z, z0.w, z1.x, z1.y, z1.z, z1.w};
                        const float wyv[8] = {y0.x, y0.y, y0.z, y0.w, y1.x, y1.y, y1.z, y1.w};
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            accZ[0][j] = fmaf(h0, wzv[j], accZ[0][j]); accZ[1][j] = fmaf(h1, wzv[j], accZ[1][j]);
                            accY[0][j] = fmaf(h0, wyv[j], accY[0][j]); accY[1][j] = fmaf(h1, wyv[j], accY[1][j]);
                        }
                    }
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        S.Zs[(txl * 8 + j) * kTileRows + r0] = accZ[0][j];
                        S.Zs[(txl * 8 + j) * kTileRows + r0 + 64] = accZ[1][j];
                    }
                }
                if (layer < a.L) store_weights(wb ^ 1);    // nobody reads that buffer during this layer
                __syncthreads();                           // Z of this CTA's features complete

                if (layer < a.L) {
                    // ---- h' = relu(Y + mean_{|k|<=r, k!=0} Z_{i+k})  (src/flux_gnn.py:55-60) into buffer cur^1 of every CTA ----
                    if (gemm) {
#pragma unroll
                        for (int q = 0; q < 2; ++q) {
                            const uint32_t off0 = (uint32_t)(((cur ^ 1) * kH * kTileRows + r0 + 64 * q) * 4);
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                const float* zr = &S.Zs[(txl * 8 + j) * kTileRows];
                                float s = zr[rp[q][0]] + zr[rm[q][0]];
#pragma unroll
                                for (int k = 1; k < kLatMaxR; ++k)
                                    if (k < radius) { s += zr[rp[q][k]]; s += zr[rm[q][k]]; }
                                const float h = fmaxf(fmaf(s, inv_deg, accY[q][j]), 0.f);
                                const uint32_t off = off0 + (uint32_t)((tx + 16 * j) * kTileRows * 4);
#pragma unroll
                                for (int p = 0; p < kLatCluster; ++p) st_cluster_f32(peer_hs[p] + off, h);
                            }
                        }
                    }
                    cluster_sync_all();                    // every CTA holds the complete h'; Zs is free again
                } else {
                    // ---- edge readout (src/flux_gnn.py:63-66): accY = P + b1, Zs = Q; hop 1 only ----
                    float pf[2] = {0.f, 0.f}, pb[2] = {0.f, 0.f};
                    if (gemm) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float w2 = __ldg(a.packed + SmallParams::w_e2 + tx + 16 * j);
                            const float* zr = &S.Zs[(txl * 8 + j) * kTileRows];
#pragma unroll
                            for (int q = 0; q < 2; ++q) {
                                pf[q] = fmaf(w2, fmaxf(accY[q][j] + zr[rp[q][0]], 0.f), pf[q]);
                                pb[q] = fmaf(w2, fmaxf(accY[q][j] + zr[rm[q][0]], 0.f), pb[q]);
                            }
                        }
                        if (txl == 1) {
                            S.pair[0][r0] = pf[0]; S.pair[0][r0 + 64] = pf[1];
                            S.pair[1][r0] = pb[0]; S.pair[1][r0 + 64] = pb[1];
                        }
                    }
                    __syncthreads();
                    if (gemm && txl == 0) {
#pragma unroll
                        for (int q = 0; q < 2; ++q) {
                            const int r = r0 + 64 * q;
                            const float cf = pf[q] + S.pair[0][r], cb = pb[q] + S.pair[1][r];      // tx even + tx odd
#pragma unroll
                            for (int p = 0; p < kLatCluster; ++p) {
                                st_cluster_f32(peer_edge[p] + (uint32_t)(r * 4), cf);
                                st_cluster_f32(peer_edge[p] + (uint32_t)((kTileRows + r) * 4), cb);
                            }
                        }
                    }
                    cluster_sync_all();
                }
            }   // layers

            // ---- face flux (src/hybrid_solver.py:45-48): the tile kernel's summation tree over the thread columns ----
            if (tid < kTileRows) {
                const int j = tid;
                const float b2 = __ldg(a.packed + SmallParams::b_e2);
                auto tree = [&](int dir, int r) {
                    const float h0 = (S.edge_all[0][dir][r] + S.edge_all[1][dir][r]) + (S.edge_all[2][dir][r] + S.edge_all[3][dir][r]);
                    const float h1 = (S.edge_all[4][dir][r] + S.edge_all[5][dir][r]) + (S.edge_all[6][dir][r] + S.edge_all[7][dir][r]);
                    return (h0 + h1) + b2;
                };
                const float fwd = tree(0, j), bwd = tree(1, S.nextRow[j]);
                const float face = 0.5f * (fwd + bwd);
                if (crank == 0 && a.face_flux != nullptr && S.rowIC[j] >= 0)
                    a.face_flux[(size_t)S.rowIC[j] * nx + S.rowCell[j]] = face;
                S.sF[j] = face;
            }
            __syncthreads();
            // ---- finite-volume update (src/hybrid_solver.py:51-58), field solve (src/baseline_solver.py:59-68) ----
            float n_new = 0.f, u_new = 0.f;
            if (tid < kTileRows) tile_fv_update(a, T, tid, n_new, u_new);
            __syncthreads();                               // everyone has read the old n, u
            if (tid < kTileRows) tile_keep_row(T, tid, n_new, u_new);
            __syncthreads();
            {
                const int r = tid >> 1, half = tid & 1;
                double e = (r < used_rows) ? tile_field_partial(T, r, half, 2, nx) : 0.0;
                e += __shfl_xor_sync(0xffffffffu, e, 1);
                if (half == 0) S.sE[r] = (float)e;
            }
            __syncthreads();
            if (crank == 0 && tid < kTileRows) tile_write_out_row(a, T, tid, step);
            // the next step's first cluster barrier comes after its first product; edge_all is rewritten only after four more
        }   // steps
        __syncthreads();
    }       // tiles
    cluster_sync_all();                                    // shared memory must outlive the peers' remote stores
}

}  // namespace

// Clusters of 8 CTAs of this kernel the device runs at once (0: not launchable).
int hybrid_latency_max_clusters() {
    static int cached = -2;
    if (cached == -2) {
        int n = 0;
        cudaError_t e = cudaFuncSetAttribute(hybrid_latency_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(LatSmem));
        if (e == cudaSuccess) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(kLatCluster * 32);
            cfg.blockDim = dim3(kLatThreads);
            cfg.dynamicSmemBytes = sizeof(LatSmem);
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = kLatCluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
            cfg.attrs = attr; cfg.numAttrs = 1;
            e = cudaOccupancyMaxActiveClusters(&n, hybrid_latency_kernel, &cfg);
        }
        if (e != cudaSuccess) { (void)cudaGetLastError(); n = 0; }
        cached = n;
    }
    return cached;
}

bool hybrid_latency_supported(const HybridArgs& a) {
    return a.whole_ic && a.do_update && a.hops == 1 && a.flux_edges == nullptr && a.acts == nullptr && a.tc_parts == 0 &&
           a.tile_rows == kTileRows && a.radius <= kLatMaxR && !a.slab;
}

cudaError_t launch_hybrid_latency(const HybridArgs& a, int clusters, cudaStream_t stream) {
    cudaError_t e = cudaFuncSetAttribute(hybrid_latency_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(LatSmem));
    if (e != cudaSuccess) return e;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(clusters * kLatCluster));
    cfg.blockDim = dim3(kLatThreads);
    cfg.dynamicSmemBytes = sizeof(LatSmem);
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = kLatCluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, hybrid_latency_kernel, a);
}

}  // namespace fluxgnn
