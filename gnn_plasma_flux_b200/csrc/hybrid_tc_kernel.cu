// Tensor-core variant of the fused hybrid step (sm_100a: tcgen05.mma + TMEM + bulk-copy TMA).
//
// Same tile structure, algebra and finite-volume / field-solve tail as the FP32-pipe
// kernel (hybrid_kernel_impl.cuh); only the five [128 x 128] x [128 x 256] contractions
// per step move to the 5th-generation tensor cores (src/flux_gnn.py:60,66 of the reference).
//
// The product is issued TRANSPOSED, D^T[n][i] = sum_k W[n][k] h[i][k]: the weight half is
// the UMMA "A" operand (M = 128 output features), the activations are the "B" operand
// (N = 128 tile rows), and the accumulator lands in TMEM with lane = feature, column = row.
// A thread that owns TMEM lane n therefore sees every row i of its feature in its own
// registers, so the +-r neighbour window of the message-passing mean needs no shuffles and
// no shared-memory round trip.  D_Z = W[:, H:] h lives in TMEM columns [0,128),
// D_Y = W[:, :H] h in [128,256).
//
// Precision modes (HybridArgs::tc_parts):
//   2  "tf32x3": h = h_hi + h_lo, W = W_hi + W_lo (each part a TF32 number);
//                D = W_hi h_hi + W_hi h_lo + W_lo h_hi, fp32 accumulation -> fp32-level accuracy
//   1  "tf32"  : D = tf32(W) tf32(h) (looser, documented tolerance)
//
// Warp roles (576 threads): warps 0-15 epilogue/compute (TMEM lane quadrant = warp % 4,
// 32-row chunk = warp / 4), warp 16 weight producer (cp.async.bulk of pre-swizzled 16 KiB
// operand images), warp 17 TMEM allocator + single-thread UMMA issuer.
#include "common.cuh"
#include "hybrid_kernel.cuh"
#include "tile_common.cuh"

namespace fluxgnn {

namespace {

constexpr int kTcStages = 5;
constexpr int kEpiThreads = 512;
constexpr int kProducerWarp = 16, kMmaWarp = 17;
constexpr int kTcThreads = kEpiThreads + 64;
constexpr uint32_t kTmemCols = 256;
constexpr uint32_t kColZ = 0, kColY = 128;

struct __align__(1024) TcSmem {
    float Bhi[4][kTileRows * 32];        // activations, hi part: 4 K-blocks of [128 rows][32 k], SW128 K-major
    float Blo[4][kTileRows * 32];        // activations, lo part (tf32x3 only)
    float Ws[kTcStages][kTcUnitFloats];  // streamed weight operand images
    float small[SmallParams::count];
    float sN[kTileRows], sU[kTileRows], sE[kTileRows], sX[kTileRows];
    float sF[kTileRows], sRho[kTileRows];
    float edge[2][kTileRows];            // per row: fwd (row j -> j+1) and bwd (row j -> j-1) dot products
    double gtab[kTileRows];
    int rowIC[kTileRows];
    int rowCell[kTileRows];
    short prevRow[kTileRows], nextRow[kTileRows];
    uint64_t full[kTcStages], empty[kTcStages];
    uint64_t act_ready, acc_ready;
    uint32_t tmem_base;
};
static_assert(sizeof(TcSmem) + 1024 <= 227 * 1024, "tensor tile does not fit shared memory");

struct Ring {
    int stage = 0;
    uint32_t phase = 0;
    __device__ __forceinline__ void advance() {
        if (++stage == kTcStages) { stage = 0; phase ^= 1; }
    }
};

}  // namespace

template <int R>
__global__ void __launch_bounds__(kTcThreads, 1) hybrid_tc_kernel(const HybridArgs a) {
    // declared 1024-byte aligned (128-byte swizzle); indexing the array itself keeps accesses in the shared space
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    TcSmem& S = *reinterpret_cast<TcSmem*>(smem_raw);
    if ((smem_u32(smem_raw) & 1023u) != 0) __trap();
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int nx = a.nx;
    const int parts = a.tc_parts;

    if (tid == 0) {
        for (int s = 0; s < kTcStages; ++s) {
            mbar_init(&S.full[s], 1);
            mbar_init(&S.empty[s], 1);
        }
        mbar_init(&S.act_ready, kEpiThreads);
        mbar_init(&S.acc_ready, 1);
        mbar_fence_init();
    }
    if (warp == kMmaWarp) tmem_alloc(&S.tmem_base, kTmemCols);
    for (int i = tid; i < SmallParams::count; i += kTcThreads) S.small[i] = a.packed[i];
    if (a.whole_ic && a.do_update)
        for (int i = tid; i < nx; i += kTcThreads) S.gtab[i] = a.gtab[i];
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = S.tmem_base;

    const int my_tiles = (a.num_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const int layers = a.L + 1;

    if (warp == kProducerWarp) {
        // ---------------- weight producer ------------------------------------------------
        if (lane == 0) {
            const float* stream = a.packed + SmallParams::count;
            Ring r;
            for (long long rep = 0; rep < (long long)my_tiles * a.steps; ++rep) {
                for (int layer = 0; layer < layers; ++layer) {
                    for (int u = 0; u < kTcUnitsPerLayer; ++u) {
                        if (parts == 1 && (u & 1)) continue;          // plain tf32: no lo units
                        mbar_wait(&S.empty[r.stage], r.phase ^ 1);
                        mbar_arrive_expect_tx(&S.full[r.stage], kTcUnitFloats * 4);
                        bulk_g2s(S.Ws[r.stage], stream + ((size_t)layer * kTcUnitsPerLayer + u) * kTcUnitFloats,
                                 kTcUnitFloats * 4, &S.full[r.stage]);
                        r.advance();
                    }
                }
            }
        }
    } else if (warp == kMmaWarp) {
        // ---------------- UMMA issuer ------------------------------------------------------
        // The whole warp walks the loop in uniform control flow, one elected lane issues (see
        // hybrid_tc16_kernel.cu: with a `lane == 0` branch around the loop every UTCHMMA was wrapped in
        // R2UR moves and an ELECT serialisation loop, ~136 clk of a lone warp per instruction).
        const uint32_t idesc = umma_idesc_tf32(128, 128);
        const bool leader = elect_one_lane();
        const uint64_t ws_desc0 = umma_desc_sw128(smem_u32(S.Ws[0]));
        const uint64_t bhi = umma_desc_sw128(smem_u32(S.Bhi)), blo = umma_desc_sw128(smem_u32(S.Blo));
        constexpr uint64_t kStageStep = (uint64_t)(kTcUnitFloats * 4) >> 4;     // address field counts 16-byte units
        constexpr uint64_t kKbStep = (uint64_t)(kTileRows * 128) >> 4;
        Ring r;
        uint32_t act_phase = 0;
        for (long long rep = 0; rep < (long long)my_tiles * a.steps; ++rep) {
            for (int layer = 0; layer < layers; ++layer) {
                mbar_wait(&S.act_ready, act_phase);
                act_phase ^= 1;
                tc_fence_after();
#pragma unroll
                for (int kb = 0; kb < 4; ++kb) {
#pragma unroll
                    for (int blk = 0; blk < 2; ++blk) {
                        const uint32_t d = tmem + (blk == 0 ? kColZ : kColY);
                        // hi weights x (hi [+ lo] activations)
                        mbar_wait(&S.full[r.stage], r.phase);
                        tc_fence_after();
                        uint64_t wd = ws_desc0 + r.stage * kStageStep;
                        if (leader) {
#pragma unroll
                            for (int ks = 0; ks < 4; ++ks) {            // 32 bytes of K per instruction = 2 address units
                                umma_tf32(d, wd + 2 * ks, bhi + kb * kKbStep + 2 * ks, idesc, (kb | ks) != 0);
                                if (parts == 2) umma_tf32(d, wd + 2 * ks, blo + kb * kKbStep + 2 * ks, idesc, 1);
                            }
                            umma_commit(&S.empty[r.stage]);
                        }
                        r.advance();
                        if (parts == 2) {
                            // lo weights x hi activations
                            mbar_wait(&S.full[r.stage], r.phase);
                            tc_fence_after();
                            wd = ws_desc0 + r.stage * kStageStep;
                            if (leader) {
#pragma unroll
                                for (int ks = 0; ks < 4; ++ks)
                                    umma_tf32(d, wd + 2 * ks, bhi + kb * kKbStep + 2 * ks, idesc, 1);
                                umma_commit(&S.empty[r.stage]);
                            }
                            r.advance();
                        }
                    }
                }
                if (leader) umma_commit(&S.acc_ready);
                __syncwarp();
            }
        }
    } else {
        // ---------------- epilogue / compute warps ---------------------------------------
        const int q = warp & 3, chunk = warp >> 2;            // TMEM lane quadrant, 32-row chunk
        const int i0 = 32 * chunk;
        const int n = 32 * q + lane;                          // this thread's feature = TMEM lane
        const uint32_t tlane = tmem + ((uint32_t)(32 * q) << 16);
        unsigned char* const bhi_bytes = reinterpret_cast<unsigned char*>(S.Bhi);
        unsigned char* const blo_bytes = reinterpret_cast<unsigned char*>(S.Blo);
        const float inv_deg = 1.0f / (float)(2 * R);
        const int seg = a.whole_ic ? nx : kTileRows;          // periodic segment inside the tile (multiple of 32)
        uint32_t acc_phase = 0;
        const TileRows T{S.sN, S.sU, S.sE, S.sX, S.sF, S.sRho, S.gtab, S.rowIC, S.rowCell, S.prevRow, S.nextRow};

        // activation element (row i0 + j, feature n): K-block q, 16-byte chunk (lane/4) ^ (j & 7)
        const uint32_t act_base = (uint32_t)(q * (kTileRows * 128) + i0 * 128 + ((lane & 3) << 2));
        uint32_t act_x[8];
#pragma unroll
        for (int v = 0; v < 8; ++v) act_x[v] = act_base + ((uint32_t)((lane >> 2) ^ v) << 4);
        auto store_act = [&](int j, float h) {                // j: row inside the chunk (compile-time after unrolling)
            const uint32_t off = act_x[j & 7] + (uint32_t)j * 128;
            const float hi = to_tf32(h);
            *reinterpret_cast<float*>(bhi_bytes + off) = hi;
            if (parts == 2) *reinterpret_cast<float*>(blo_bytes + off) = to_tf32(h - hi);
        };

        for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
            // ---- row bookkeeping + state load (as in the FP32-pipe kernel) ----------------
            if (tid < kTileRows) tile_load_row(a, T, tile, true, tid, tid, 0, kTileRows);
            named_sync(1, kEpiThreads);

            for (int step = 0; step < a.steps; ++step) {
                // ---- input MLP (src/flux_gnn.py:49): feature n, rows of this warpgroup's half ----
                {
                    const float w0 = S.small[SmallParams::w_in + 0 * kH + n];
                    const float w1 = S.small[SmallParams::w_in + 1 * kH + n];
                    const float w2 = S.small[SmallParams::w_in + 2 * kH + n];
                    const float w3 = S.small[SmallParams::w_in + 3 * kH + n];
                    const float b = S.small[SmallParams::b_in + n];
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const int i = i0 + j;
                        float v = fmaf(w0, S.sN[i], b);
                        v = fmaf(w1, S.sU[i], v);
                        v = fmaf(w2, S.sE[i], v);
                        v = fmaf(w3, S.sX[i], v);
                        store_act(j, fmaxf(v, 0.f));
                    }
                }
                tc_fence_before();
                fence_proxy_async();
                mbar_arrive(&S.act_ready);

                for (int layer = 0; layer < layers; ++layer) {
                    mbar_wait(&S.acc_ready, acc_phase);
                    acc_phase ^= 1;
                    tc_fence_after();
                    const bool is_edge = (layer == a.L);
                    const float bias = S.small[(is_edge ? SmallParams::b_e1 : SmallParams::b_upd + layer * kH) + n];
                    const float w_out = S.small[SmallParams::w_e2 + n];
                    {
                        const int seg0 = (i0 / seg) * seg;
                        const int cl = (i0 == seg0) ? i0 - 4 + seg : i0 - 4;
                        const int cr = (i0 + 32 == seg0 + seg) ? i0 + 32 - seg : i0 + 32;
                        float y[32], zc[32], zl[4], zr[4];
                        tmem_ld32(tlane + kColY + i0, y);
                        tmem_ld32(tlane + kColZ + i0, zc);
                        tmem_ld4(tlane + kColZ + cl, zl);
                        tmem_ld4(tlane + kColZ + cr, zr);
                        tc_wait_ld();
                        float zw[40];
#pragma unroll
                        for (int t = 0; t < 4; ++t) { zw[t] = zl[t]; zw[36 + t] = zr[t]; }
#pragma unroll
                        for (int t = 0; t < 32; ++t) zw[4 + t] = zc[t];
                        if (!is_edge) {
                            // h'_i = relu(Y_i + b + mean_{0<|k|<=R} Z_{i+k})   (src/flux_gnn.py:55-60)
#pragma unroll
                            for (int j = 0; j < 32; ++j) {
                                float s = zw[4 + j + 1] + zw[4 + j - 1];
#pragma unroll
                                for (int k = 2; k <= R; ++k) { s += zw[4 + j + k]; s += zw[4 + j - k]; }
                                store_act(j, fmaxf(fmaf(s, inv_deg, y[j] + bias), 0.f));
                            }
                        } else {
                            // edge readout (src/flux_gnn.py:63-66): this feature's term of the two dot products
                            // of row i, written feature-transposed so that one thread can sum a row:
                            //   fwd (row i, col i+1): w2[n] relu(P_i + b1 + Q_{i+1});  bwd (row i, col i-1)
                            float* fbuf = reinterpret_cast<float*>(S.Bhi);
                            float* gbuf = reinterpret_cast<float*>(S.Blo);
#pragma unroll
                            for (int j = 0; j < 32; ++j) {
                                const int i = i0 + j;
                                const float p = y[j] + bias;
                                fbuf[i * kH + (n ^ (i & 31))] = w_out * fmaxf(p + zw[4 + j + 1], 0.f);
                                gbuf[i * kH + (n ^ (i & 31))] = w_out * fmaxf(p + zw[4 + j - 1], 0.f);
                            }
                        }
                    }
                    if (!is_edge) {
                        tc_fence_before();
                        fence_proxy_async();
                        mbar_arrive(&S.act_ready);
                    }
                }

                // ---- reduce the edge terms over the 128 features: thread (row, which) ------------
                named_sync(1, kEpiThreads);
                if (tid < 2 * kTileRows) {
                    const int row = tid & (kTileRows - 1), which = tid >> 7;
                    const float* buf = reinterpret_cast<const float*>(which ? S.Blo : S.Bhi) + row * kH;
                    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll 8
                    for (int m = 0; m < kH; m += 4) {
                        s0 += buf[(m + 0) ^ (row & 31)];
                        s1 += buf[(m + 1) ^ (row & 31)];
                        s2 += buf[(m + 2) ^ (row & 31)];
                        s3 += buf[(m + 3) ^ (row & 31)];
                    }
                    S.edge[which][row] = (s0 + s1) + (s2 + s3);
                }
                named_sync(1, kEpiThreads);

                // ---- per row: face flux (src/hybrid_solver.py:45-48) ----------------------------
                float n_new = 0.f, u_new = 0.f;
                if (tid < kTileRows) {
                    const int j = tid, jn = S.nextRow[j];
                    const float b2 = S.small[SmallParams::b_e2];
                    const float fwd = S.edge[0][j] + b2;          // edge (row j, col j+1)
                    const float bwd = S.edge[1][jn] + b2;         // edge (row j+1, col j)
                    const float face = 0.5f * (fwd + bwd);
                    const int ic = S.rowIC[j], cell = S.rowCell[j];
                    if (a.flux_edges != nullptr && ic >= 0) {
                        float* fe = a.flux_edges + (size_t)ic * 2 * nx + cell;
                        fe[0] = fwd;
                        fe[nx] = bwd;
                    }
                    if (a.face_flux != nullptr && ic >= 0) a.face_flux[(size_t)ic * nx + cell] = face;
                    S.sF[j] = face;
                }
                if (!a.do_update) continue;
                named_sync(1, kEpiThreads);

                // ---- finite-volume update (src/hybrid_solver.py:51-58) ----
                if (tid < kTileRows) tile_fv_update(a, T, tid, n_new, u_new);
                if (!a.whole_ic) {
                    if (tid < kTileRows) tile_store_window_row(a, T, tid, n_new, u_new);
                    continue;
                }
                named_sync(1, kEpiThreads);
                if (tid < kTileRows) tile_keep_row(T, tid, n_new, u_new);
                named_sync(1, kEpiThreads);
                // ---- field solve: E = g (*) rho, four threads per row (src/baseline_solver.py:59-68) ----
                {
                    const int row = tid >> 2, part = tid & 3;
                    double e = tile_field_partial(T, row, part, 4, nx);
                    e += __shfl_xor_sync(0xffffffffu, e, 1);
                    e += __shfl_xor_sync(0xffffffffu, e, 2);
                    if (part == 0) S.sE[row] = (float)e;
                }
                named_sync(1, kEpiThreads);
                if (tid < kTileRows) tile_write_out_row(a, T, tid, step);
            }   // steps
            named_sync(1, kEpiThreads);
        }       // tiles
    }

    tc_fence_before();
    __syncthreads();
    if (warp == kMmaWarp) tmem_dealloc(tmem, kTmemCols);
}

template <int R>
static cudaError_t launch_tc_one(const HybridArgs& a, int grid, cudaStream_t stream) {
    const int smem = (int)sizeof(TcSmem);
    cudaError_t e = cudaFuncSetAttribute(hybrid_tc_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    hybrid_tc_kernel<R><<<grid, kTcThreads, smem, stream>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_hybrid_tc_tiles(const HybridArgs& a, int radius, int grid, cudaStream_t stream) {
    switch (radius) {
        case 1: return launch_tc_one<1>(a, grid, stream);
        case 2: return launch_tc_one<2>(a, grid, stream);
        case 3: return launch_tc_one<3>(a, grid, stream);
        case 4: return launch_tc_one<4>(a, grid, stream);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace fluxgnn
