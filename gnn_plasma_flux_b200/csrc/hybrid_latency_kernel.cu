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
//   src/flux_gnn.py:63-66     edge readout           per-CTA partial dot products, exchanged as 2 x 128 floats per CTA and hop
//                                                    (forward-only calls emit every hop's directed-edge fluxes)
//   src/hybrid_solver.py:45-58, src/baseline_solver.py:59-68   update and field solve: every CTA redundantly, CTA 0 writes
//
// A product thread (the first 128 of 256) owns 4 consecutive rows and 4 features (j = 0..3 or 4..7 of one tx) of both halves:
// 32 accumulators fed by three 128-bit shared-memory loads per k (the product is bound by the shared-memory return path,
// and a 4 x 4 tile needs the fewest operand bytes per multiply-add).  Every sum runs in the tile kernel's order (k = 0..127
// per accumulator; window sums hop by hop; the edge readout's chain over j -- handed from the j < 4 lane to its j >= 4
// neighbour by a shuffle --, then the pair tx even + tx odd, then the tile kernel's butterfly tree across the CTAs), so
// the result is BIT-IDENTICAL to hybrid_tile_kernel -- the tests compare with torch.equal.  Weights: each CTA stages its 16 KiB slice of a layer (2 halves x 128 k x 16 columns) from
// the packed stream (L2-resident) into shared memory, the next layer's slice is prefetched into registers under the
// current product.
#include <cstdio>

#include "common.cuh"
#include "hybrid_kernel.cuh"
#include "tile_common.cuh"

namespace fluxgnn {

namespace {

#ifdef FLUXGNN_LAT_TIMING
#define LAT_TICK(slot) do { if (tid == 0 && blockIdx.x == 0) { const long long n__ = clock64(); lat_t[slot] += n__ - lat_last; lat_last = n__; } } while (0)
#else
#define LAT_TICK(slot) do {} while (0)
#endif

#ifndef FLUXGNN_LAT_UNROLL
#define FLUXGNN_LAT_UNROLL 8          // clk per k-step of the product: unroll 4: 66, 8: 60, 16: 58 (floor of the shared-memory return path: 48)
#endif
constexpr int kLatUnroll = FLUXGNN_LAT_UNROLL;
constexpr int kLatCluster = 8;
constexpr int kLatThreads = 256;
constexpr int kLatFeat = kH / kLatCluster;        // 16 output features per CTA
constexpr int kLatMaxR = 4;                       // neighbour rows kept in registers; larger radii are not dispatched here

struct __align__(128) LatSmem {
    float Hs[2][kH * kTileRows];                  // activations [buffer][feature][row], a complete copy in every CTA
    float Zs[kLatFeat * kTileRows];               // neighbour half of this CTA's features [local feature][row]
    float Wl[2][2][kH][kLatFeat];                 // weight slice [buffer][half: 0 = neighbour/col, 1 = self/row][k][column]
    float edge_all[kLatCluster][kLatMaxR][2][kTileRows];   // per source CTA and hop: fwd / bwd partial sums (tx even + tx odd) of every row
    float sN[kTileRows], sU[kTileRows], sE[kTileRows], sX[kTileRows];
    float sF[kTileRows], sRho[kTileRows];
    double gtab[kTileRows];
    int rowIC[kTileRows];
    int rowCell[kTileRows];
    short prevRow[kTileRows], nextRow[kTileRows];
};
static_assert(sizeof(LatSmem) <= 227 * 1024, "latency tile does not fit shared memory");

__device__ __forceinline__ void st_cluster_f4(uint32_t addr, float4 v) {
    asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

__global__ void __launch_bounds__(kLatThreads, 1) hybrid_latency_kernel(const HybridArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    LatSmem& S = *reinterpret_cast<LatSmem*>(smem_raw);
    const int tid = threadIdx.x;
    const bool gemm = tid < kTileRows;                     // product / epilogue threads
    const int lane = tid & 31;
    const int rg = ((tid >> 5) & 3) * 8 + (lane >> 2);    // rows 4 rg .. 4 rg + 3
    const int txl = (lane >> 1) & 1, jh = lane & 1;        // thread column txl, features j = 4 jh .. 4 jh + 3 of it
    const int R0 = 4 * rg;
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
        peer_edge[p] = cluster_map(smem_u32(&S.edge_all[crank][0][0][0]), p);
    }
    if (a.do_update)
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
#ifdef FLUXGNN_LAT_TIMING
    long long lat_t[8] = {0, 0, 0, 0, 0, 0, 0, 0}, lat_last = clock64();
#endif

    for (int tile = first_tile; tile < a.num_tiles; tile += tile_stride) {
        if (tid < kTileRows) tile_load_row(a, T, tile, true, tid, tid, 0, kTileRows);
        fetch_weights(0);
        __syncthreads();
        // neighbour rows of this thread's four rows, hop by hop (periodic inside the IC)
        int rp[4][kLatMaxR], rm[4][kLatMaxR];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            int p = R0 + q, m = R0 + q;
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
                const float4* w4 = reinterpret_cast<const float4*>(a.packed + SmallParams::w_in);
                const float4* b4 = reinterpret_cast<const float4*>(a.packed + SmallParams::b_in);
                for (int g = (tid >> 7) * 16; g < (tid >> 7) * 16 + 16; ++g) {        // four features per iteration
                    const float4 w0 = __ldg(w4 + g), w1 = __ldg(w4 + kH / 4 + g), w2 = __ldg(w4 + 2 * (kH / 4) + g),
                                 w3 = __ldg(w4 + 3 * (kH / 4) + g), b = __ldg(b4 + g);
                    const float wv[4][4] = {{w0.x, w1.x, w2.x, w3.x}, {w0.y, w1.y, w2.y, w3.y}, {w0.z, w1.z, w2.z, w3.z},
                                            {w0.w, w1.w, w2.w, w3.w}};
                    const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        float v = fmaf(wv[i][0], fn, bv[i]);
                        v = fmaf(wv[i][1], fu, v);
                        v = fmaf(wv[i][2], fe, v);
                        v = fmaf(wv[i][3], fx, v);
                        S.Hs[0][(4 * g + i) * kTileRows + row] = fmaxf(v, 0.f);
                    }
                }
            }
            store_weights(0);                              // layer 0's slice (fetched before the step / under the last layer)
            __syncthreads();
            LAT_TICK(0);

            int cur = 0;
            for (int layer = 0; layer <= a.L; ++layer, cur ^= 1) {
                const int wb = layer & 1;
                // prefetch the next slice: the next layer's, or layer 0's for the next step
                if (layer < a.L) fetch_weights(layer + 1);
                else if (step + 1 < a.steps) fetch_weights(0);
                const float* bias = a.packed + (layer < a.L ? SmallParams::b_upd + layer * kH : SmallParams::b_e1);
                float accZ[4][4], accY[4][4];              // [row][feature jj]
                if (gemm) {
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) {
                        const float b = __ldg(bias + tx + 16 * (4 * jh + jj));
#pragma unroll
                        for (int q = 0; q < 4; ++q) { accZ[q][jj] = 0.f; accY[q][jj] = b; }
                    }
                    const float* hrow = &S.Hs[cur][R0];
                    const float* wz = &S.Wl[wb][0][0][txl * 8 + jh * 4];
                    const float* wy = &S.Wl[wb][1][0][txl * 8 + jh * 4];
#pragma unroll kLatUnroll
                    for (int k = 0; k < kH; ++k) {
                        const float4 h = *reinterpret_cast<const float4*>(hrow + k * kTileRows);
                        const float4 z = *reinterpret_cast<const float4*>(wz + k * kLatFeat);
                        const float4 y = *reinterpret_cast<const float4*>(wy + k * kLatFeat);
                        const float hv[4] = {h.x, h.y, h.z, h.w}, zv[4] = {z.x, z.y, z.z, z.w}, yv[4] = {y.x, y.y, y.z, y.w};
#pragma unroll
                        for (int q = 0; q < 4; ++q)
#pragma unroll
                            for (int jj = 0; jj < 4; ++jj) {
                                accZ[q][jj] = fmaf(hv[q], zv[jj], accZ[q][jj]);
                                accY[q][jj] = fmaf(hv[q], yv[jj], accY[q][jj]);
                            }
                    }
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj)
                        *reinterpret_cast<float4*>(&S.Zs[(txl * 8 + 4 * jh + jj) * kTileRows + R0]) =
                            make_float4(accZ[0][jj], accZ[1][jj], accZ[2][jj], accZ[3][jj]);
                }
                if (layer < a.L) store_weights(wb ^ 1);    // nobody reads that buffer during this layer
                __syncthreads();                           // Z of this CTA's features complete
                LAT_TICK(1);

                if (layer < a.L) {
                    // ---- h' = relu(Y + mean_{|k|<=r, k!=0} Z_{i+k})  (src/flux_gnn.py:55-60) into buffer cur^1 of every CTA ----
                    if (gemm) {
#pragma unroll
                        for (int jj = 0; jj < 4; ++jj) {
                            const float* zr = &S.Zs[(txl * 8 + 4 * jh + jj) * kTileRows];
                            float hq[4];
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                float s = zr[rp[q][0]] + zr[rm[q][0]];
#pragma unroll
                                for (int k = 1; k < kLatMaxR; ++k)
                                    if (k < radius) { s += zr[rp[q][k]]; s += zr[rm[q][k]]; }
                                hq[q] = fmaxf(fmaf(s, inv_deg, accY[q][jj]), 0.f);
                            }
                            const uint32_t off = (uint32_t)((((cur ^ 1) * kH + tx + 16 * (4 * jh + jj)) * kTileRows + R0) * 4);
                            const float4 hv = make_float4(hq[0], hq[1], hq[2], hq[3]);
#pragma unroll
                            for (int p = 0; p < kLatCluster; ++p) st_cluster_f4(peer_hs[p] + off, hv);
                        }
                    }
                    // (tried: slice to local shared memory first, then all eight warps copy it to the peers -- slower, 14.5 k + 8.0 k
                    //  of barrier instead of 13.0 k + 5.8 k clk per step: the pushes are bound by the SM's egress, not by issue)
                    LAT_TICK(2);
                    cluster_sync_all();                    // every CTA holds the complete h'; Zs is free again
                    LAT_TICK(3);
                } else {
                    // ---- edge readout (src/flux_gnn.py:63-66): accY = P + b1, Zs = Q; hop 1 only.  The tile kernel sums
                    // w2 relu(.) over j = 0..7 in one chain per (row, tx): the j < 4 lane runs its half and hands the sum to
                    // its j >= 4 neighbour, which continues the chain ----
                    if (gemm) {
                        float w2[4];
#pragma unroll
                        for (int jj = 0; jj < 4; ++jj) w2[jj] = __ldg(a.packed + SmallParams::w_e2 + tx + 16 * (4 * jh + jj));
#pragma unroll
                        for (int hop = 1; hop <= kLatMaxR; ++hop) {
                            if (hop > a.hops) break;
                            float pf[4], pb[4];
#pragma unroll
                            for (int pass = 0; pass < 2; ++pass) {
#pragma unroll
                                for (int q = 0; q < 4; ++q) {
                                    // pass 0: chains seeded with 0 (final for the j < 4 lanes); pass 1: the j >= 4 lanes redo theirs
                                    // seeded with the neighbour's result
                                    float sf = 0.f, sb = 0.f;
                                    if (pass == 1) {
                                        sf = __shfl_xor_sync(0xffffffffu, pf[q], 1);
                                        sb = __shfl_xor_sync(0xffffffffu, pb[q], 1);
                                    }
                                    if (pass == 0 || jh == 1) {
#pragma unroll
                                        for (int jj = 0; jj < 4; ++jj) {
                                            const float* zr = &S.Zs[(txl * 8 + 4 * jh + jj) * kTileRows];
                                            sf = fmaf(w2[jj], fmaxf(accY[q][jj] + zr[rp[q][hop - 1]], 0.f), sf);
                                            sb = fmaf(w2[jj], fmaxf(accY[q][jj] + zr[rm[q][hop - 1]], 0.f), sb);
                                        }
                                        pf[q] = sf; pb[q] = sb;
                                    }
                                }
                            }
                            // the complete sums of a thread column sit in its j >= 4 lane; tx even + tx odd, then to every CTA
                            float cf[4], cb[4];
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                cf[q] = pf[q] + __shfl_xor_sync(0xffffffffu, pf[q], 2);
                                cb[q] = pb[q] + __shfl_xor_sync(0xffffffffu, pb[q], 2);
                            }
                            if (txl == 0 && jh == 1) {
                                const float4 f4 = make_float4(cf[0], cf[1], cf[2], cf[3]), b4 = make_float4(cb[0], cb[1], cb[2], cb[3]);
                                const uint32_t off = (uint32_t)(((hop - 1) * 2 * kTileRows + R0) * 4);
#pragma unroll
                                for (int p = 0; p < kLatCluster; ++p) {
                                    st_cluster_f4(peer_edge[p] + off, f4);
                                    st_cluster_f4(peer_edge[p] + off + (uint32_t)(kTileRows * 4), b4);
                                }
                            }
                        }
                    }
                    LAT_TICK(4);
                    cluster_sync_all();
                    LAT_TICK(3);
                }
            }   // layers

            // ---- directed-edge fluxes and face flux (src/hybrid_solver.py:45-48): the tile kernel's summation tree over the
            // thread columns ----
            if (tid < kTileRows) {
                const int j = tid;
                const float b2 = __ldg(a.packed + SmallParams::b_e2);
                const int ic = S.rowIC[j], cell = S.rowCell[j];
                float face = 0.f;
                int jn = j;
                for (int hop = 1; hop <= a.hops; ++hop) {
                    jn = S.nextRow[jn];
                    auto tree = [&](int dir, int r) {
                        const float h0 = (S.edge_all[0][hop - 1][dir][r] + S.edge_all[1][hop - 1][dir][r]) +
                                         (S.edge_all[2][hop - 1][dir][r] + S.edge_all[3][hop - 1][dir][r]);
                        const float h1 = (S.edge_all[4][hop - 1][dir][r] + S.edge_all[5][hop - 1][dir][r]) +
                                         (S.edge_all[6][hop - 1][dir][r] + S.edge_all[7][hop - 1][dir][r]);
                        return (h0 + h1) + b2;
                    };
                    const float fwd = tree(0, j), bwd = tree(1, jn);
                    if (hop == 1) face = 0.5f * (fwd + bwd);
                    if (crank == 0 && a.flux_edges != nullptr && ic >= 0) {
                        float* fe = a.flux_edges + (size_t)ic * 2 * a.hops * nx + (size_t)(2 * (hop - 1)) * nx + cell;
                        fe[0] = fwd;
                        fe[nx] = bwd;
                    }
                }
                if (crank == 0 && a.face_flux != nullptr && ic >= 0) a.face_flux[(size_t)ic * nx + cell] = face;
                S.sF[j] = face;
            }
            if (!a.do_update) continue;                    // forward only (steps == 1)
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
            LAT_TICK(5);
            // the next step's first cluster barrier comes after its first product; edge_all is rewritten only after four more
        }   // steps
        __syncthreads();
    }       // tiles
    cluster_sync_all();                                    // shared memory must outlive the peers' remote stores
#ifdef FLUXGNN_LAT_TIMING
    if (tid == 0 && blockIdx.x == 0)
        printf("[latency timing] clk per step: input %lld, products %lld, node epilogue + push %lld, cluster barriers %lld, edge readout %lld, update + field %lld\n",
               lat_t[0] / a.steps, lat_t[1] / a.steps, lat_t[2] / a.steps, lat_t[3] / a.steps, lat_t[4] / a.steps, lat_t[5] / a.steps);
#endif
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
    return a.whole_ic && a.hops >= 1 && a.hops <= kLatMaxR && a.acts == nullptr && a.tc_parts == 0 &&
           a.tile_rows == kTileRows && a.radius <= kLatMaxR && !a.slab && (a.do_update || a.steps == 1);
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
