// Fused hybrid step for sm_100a: FluxGNN on the periodic chain + finite-volume
// update + (for grids that fit one tile) the spectral field solve, all inside
// one persistent kernel with the activations resident in shared memory.
//
// What it replaces (paths under /root/reference):
//   src/graph_constructor.py:30-38   node features / ring edges  -> implicit
//   src/flux_gnn.py:49               input MLP                   -> input_layer()
//   src/flux_gnn.py:53-60            gather + index_add_ + mean + Linear + ReLU
//                                    -> two K=128 GEMM passes + window epilogue
//   src/flux_gnn.py:63-66            edge readout                -> split-W GEMM + edge epilogue
//   src/hybrid_solver.py:45-58       face flux, continuity and momentum update
//   src/baseline_solver.py:59-68     field solve (whole-IC tiles only)
//
// Algebra.  With W = [Wa | Wb] ([H][2H], nn.Linear layout) the reference's
//   h' = relu(W [h ; mean_nbr(h)] + b)
// is evaluated as  Y = Wa h,  Z = Wb h,  h'_i = relu(Y_i + mean_nbr(Z)_i + b)
// (the mean is linear, so it commutes with Wb).  Every layer, and the edge
// readout P = W1a h, Q = W1b h, is therefore the same [128 rows] x [K=128] x
// [N=256] product, done as two N=128 passes:
//   pass 0 -> Z (or Q) written to shared memory,
//   pass 1 -> Y (or P) kept in registers, combined with a +-r row window of Z.
//
// Tile = 128 rows (cells).  "Whole-IC" tiles hold floor(128/nx) complete ICs
// and wrap neighbours periodically inside each IC; "window" tiles (nx > 128)
// carry a halo of L*r+hops recomputed cells per side.  Activations live
// feature-major ([feature][row], row chunks XOR-swizzled) so that both the
// register-tile loads of the GEMM and the row-window reads are 128-bit and
// bank-conflict free.  Weights stream L2 -> shared memory in 8 KiB chunks with
// cp.async.bulk (TMA, SASS UBLKCP) issued by a producer warp through a 4-stage
// mbarrier ring; 8 consumer warps run an 8x8 register-tile FFMA kernel.
#pragma once

#include <cstdio>

#include "common.cuh"
#include "hybrid_kernel.cuh"
#include "tile_common.cuh"

namespace fluxgnn {

namespace {

#ifndef FLUXGNN_FFMA2
#define FLUXGNN_FFMA2 1          // packed fma.rn.f32x2 (SASS FFMA2) in the GEMM inner loop
#endif
#ifndef FLUXGNN_STAGES
#define FLUXGNN_STAGES 5
#endif
constexpr int kStages = FLUXGNN_STAGES;
constexpr int kConsumerWarps = 8;
constexpr int kConsumers = kConsumerWarps * 32;   // 256
constexpr int kThreads = kConsumers + 32;         // + producer warp

struct __align__(128) TileSmem {
    float Hs[kH * kTileRows];            // activations h   [feature][row], swizzled
    float Zs[kH * kTileRows];            // neighbour half  [feature][row], swizzled
    float Ws[kStages][kChunkFloats];     // streamed weight chunks [kChunkK k][128 permuted columns]
    float sN[kTileRows], sU[kTileRows], sE[kTileRows], sX[kTileRows];
    float sF[kTileRows], sRho[kTileRows];
    float edge[2][2][kMaxHops][kTileRows];   // [column half][fwd/bwd][hop][row] partial dot products
    double gtab[kTileRows];
    int rowIC[kTileRows];                // owning IC of the row's output, -1 = not owned
    int rowCell[kTileRows];
    short prevRow[kTileRows], nextRow[kTileRows];
    uint64_t full[kStages], empty[kStages];
    uint64_t skew;                       // split mode: group 1 starts once group 0 is kSkewChunks ahead
    uint64_t zready, zfree;              // cluster mode: the neighbours' Z is complete / they have read this CTA's Z
};
static_assert(sizeof(TileSmem) <= 227 * 1024, "tile does not fit shared memory");

// Split mode (whole-IC tiles with nx | 64): the two 64-row halves of a tile hold different ICs, so
// warps 0-3 (rows 0-63) and warps 4-7 (rows 64-127) never exchange data.  They run as two groups
// with their own named barrier, sharing only the weight ring, and group 1 is started a few chunks
// late: while one group is in an epilogue (little FMA work) the other is in its GEMM and has the
// FMA pipe of every scheduler to itself.
constexpr int kSkewChunks = (kStages >= 8) ? 3 : 2;

struct Pipe {
    int stage = 0;
    uint32_t phase = 0;
    int consumed = 0;                    // chunks this warp has finished (only compared with kSkewChunks)
#ifdef FLUXGNN_FFMA_TIMING
    long long t_wait = 0, t_gemm = 0, t_bar = 0;
#endif
    __device__ __forceinline__ void advance() {
        if (++stage == kStages) { stage = 0; phase ^= 1; }
    }
};

// Activation layout: feature-major [feature][row]; the 16-byte row chunk c of
// feature n sits at chunk position c ^ (n & 7).  A thread owns features
// n_j = tx + 16 j, so n & 7 = tx & 7 for all eight of them.
__device__ __forceinline__ int sw_off(int n, int row) {
    return n * kTileRows + ((((row >> 2) ^ (n & 7)) << 2) | (row & 3));
}

// The thread's j-th output feature.  The streamed weight chunks store column
// n = tx + 16 j at physical position 4 tx + j (j < 4) / 64 + 4 tx + (j - 4), so
// the eight weights a thread needs per k are two 128-bit loads (weight_column()
// in common.cuh is the inverse used by the packer).
__device__ __forceinline__ int col_of(int tx, int j) { return tx + 16 * j; }

// acc[i][j] += sum_k Hs[k][R0+i] * Wt[k][col_j]  over one K=128 half (streamed chunks).
// xo[s] = float offset of row chunk c0 under swizzle key s; chunk c1 = c0 ^ 1 is 4 floats
// above (s even) or below (s odd) because c0 is even.
__device__ __forceinline__ void gemm_pass(float (&acc)[8][8], TileSmem& S, Pipe& pipe,
                                          const int (&xo)[8], int tx, int lane, bool leads_skew) {
#if FLUXGNN_FFMA2
    // Packed accumulators: for row pair p = (2p, 2p+1) and column pair q = (2q, 2q+1)
    //   d[p][q] = {acc[2p][2q],   acc[2p+1][2q+1]}   (a pair) * (b pair)
    //   x[p][q] = {acc[2p][2q+1], acc[2p+1][2q]}     (a pair) * (swapped b pair)
    // so neither operand has to be broadcast into both halves of a register pair.
    float2 d[4][4], x[4][4];
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            d[p][q] = make_float2(acc[2 * p][2 * q], acc[2 * p + 1][2 * q + 1]);
            x[p][q] = make_float2(acc[2 * p][2 * q + 1], acc[2 * p + 1][2 * q]);
        }
#endif
#pragma unroll 1
#ifdef FLUXGNN_FFMA_TIMING
    const long long tg0 = clock64();
#endif
    for (int kc = 0; kc < kChunksPerHalf; ++kc) {
#ifdef FLUXGNN_FFMA_TIMING
        const long long tw0 = clock64();
        mbar_wait(&S.full[pipe.stage], pipe.phase);
        pipe.t_wait += clock64() - tw0;
#else
        mbar_wait(&S.full[pipe.stage], pipe.phase);
#endif
        const float* __restrict__ wst = S.Ws[pipe.stage];
        const float* __restrict__ hk = S.Hs + kc * kChunkK * kTileRows;
#pragma unroll
        for (int kk = 0; kk < kChunkK; ++kk) {
            const float* pa = hk + kk * kTileRows + xo[kk & 7];
            const float4 a0 = *reinterpret_cast<const float4*>(pa);
            const float4 a1 = *reinterpret_cast<const float4*>(pa + ((kk & 1) ? -4 : 4));
            const float4 b0 = *reinterpret_cast<const float4*>(wst + kk * kH + tx * 4);
            const float4 b1 = *reinterpret_cast<const float4*>(wst + kk * kH + 64 + tx * 4);
#if FLUXGNN_FFMA2
            const float2 A[4] = {make_float2(a0.x, a0.y), make_float2(a0.z, a0.w),
                                 make_float2(a1.x, a1.y), make_float2(a1.z, a1.w)};
            const float2 B[4] = {make_float2(b0.x, b0.y), make_float2(b0.z, b0.w),
                                 make_float2(b1.x, b1.y), make_float2(b1.z, b1.w)};
            const float2 Bs[4] = {make_float2(b0.y, b0.x), make_float2(b0.w, b0.z),
                                  make_float2(b1.y, b1.x), make_float2(b1.w, b1.z)};
#pragma unroll
            for (int p = 0; p < 4; ++p)
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    d[p][q] = __ffma2_rn(A[p], B[q], d[p][q]);
                    x[p][q] = __ffma2_rn(A[p], Bs[q], x[p][q]);
                }
#else
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
#endif
        }
        __syncwarp();
        if (lane == 0) {
            mbar_arrive(&S.empty[pipe.stage]);
            if (leads_skew && ++pipe.consumed == kSkewChunks) mbar_arrive(&S.skew);
        }
        pipe.advance();
    }
#ifdef FLUXGNN_FFMA_TIMING
    pipe.t_gemm += clock64() - tg0;
#endif
#if FLUXGNN_FFMA2
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            acc[2 * p][2 * q] = d[p][q].x;
            acc[2 * p + 1][2 * q + 1] = d[p][q].y;
            acc[2 * p][2 * q + 1] = x[p][q].x;
            acc[2 * p + 1][2 * q] = x[p][q].y;
        }
#endif
}

// 16-row window Z[n][R0-4 .. R0+11] (periodic inside the segment) of feature n
// zl / zr: shared::cluster address of the neighbouring CTA's Zs when the window's left / right chunk lies in its piece
// of the cluster window (0 otherwise); n = feature.
__device__ __forceinline__ void load_window(float (&v)[16], const float* zrow, int sw,
                                            int cl, int c0, int c1, int cr, uint32_t zl = 0, uint32_t zr = 0, int n = 0) {
    const float4 l = zl ? ld_cluster_f4(zl + (uint32_t)((n * kTileRows + ((cl ^ sw) << 2)) * 4))
                        : *reinterpret_cast<const float4*>(zrow + ((cl ^ sw) << 2));
    const float4 m0 = *reinterpret_cast<const float4*>(zrow + ((c0 ^ sw) << 2));
    const float4 m1 = *reinterpret_cast<const float4*>(zrow + ((c1 ^ sw) << 2));
    const float4 r = zr ? ld_cluster_f4(zr + (uint32_t)((n * kTileRows + ((cr ^ sw) << 2)) * 4))
                        : *reinterpret_cast<const float4*>(zrow + ((cr ^ sw) << 2));
    v[0] = l.x;  v[1] = l.y;  v[2] = l.z;  v[3] = l.w;
    v[4] = m0.x; v[5] = m0.y; v[6] = m0.z; v[7] = m0.w;
    v[8] = m1.x; v[9] = m1.y; v[10] = m1.z; v[11] = m1.w;
    v[12] = r.x; v[13] = r.y; v[14] = r.z; v[15] = r.w;
}

// Training forward: copy the group's rows of a feature-major activation buffer to the row-major
// global array dst[(ic*nx + cell)][128] (coalesced 512-byte rows; rows this tile does not own are skipped).
__device__ __forceinline__ void save_rows(const float* buf, float* dst, const TileSmem& S, int lt, int gthreads,
                                          int row0, int nrows, int nx) {
    for (int w = lt; w < nrows * 32; w += gthreads) {
        const int r = row0 + (w >> 5), fq = w & 31;
        const int ic = S.rowIC[r];
        if (ic < 0) continue;
        const size_t grow = (size_t)ic * nx + S.rowCell[r];
        const float4 v = make_float4(buf[sw_off(4 * fq, r)], buf[sw_off(4 * fq + 1, r)], buf[sw_off(4 * fq + 2, r)],
                                     buf[sw_off(4 * fq + 3, r)]);
        *reinterpret_cast<float4*>(dst + grow * kH + 4 * fq) = v;
    }
}

}  // namespace

// R = 1..4: radius known at compile time, 8-row groups never straddle an IC
//           (nx % 8 == 0 or window tiles) -> 128-bit window reads.
// R = 0   : any radius / any nx, neighbours found by walking prev/next tables.
// kSave: training forward -- also store h^0..h^L, P + b1 and Q row-major for the backward pass.
// A separate instantiation so that the inference kernel carries none of that code.
// kCluster: window / slab tiles whose CTAs form clusters that share one window (HybridArgs::cluster > 1); a separate
// instantiation so that the ordinary kernel carries none of the remote addressing or handshakes.
template <int R, bool kSave, bool kCluster = false>
__global__ void __launch_bounds__(kThreads, 1) hybrid_tile_kernel(const HybridArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    TileSmem& S = *reinterpret_cast<TileSmem*>(smem_raw);
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int nx = a.nx;

    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&S.full[s], 1);
            mbar_init(&S.empty[s], kConsumerWarps);
        }
        mbar_init(&S.skew, kConsumerWarps / 2);
        mbar_fence_init();
    }
    // cluster mode: this CTA is piece `crank` of a window of `csize` pieces; neighbours exist inside the window only
    const int csize = kCluster ? a.cluster : 1;
    const int crank = kCluster ? (int)cluster_ctarank() : 0;
    const bool hasL = kCluster && crank > 0, hasR = kCluster && crank < csize - 1;
    if (kCluster && tid == 0) {
        mbar_init(&S.zready, (hasL ? 1 : 0) + (hasR ? 1 : 0));
        mbar_init(&S.zfree, (hasL ? 1 : 0) + (hasR ? 1 : 0));
        mbar_fence_init();
    }
    if (a.whole_ic && a.do_update)
        for (int i = tid; i < nx; i += kThreads) S.gtab[i] = a.gtab[i];
    __syncthreads();
    if (kCluster) cluster_sync_all();                // nobody signals a barrier that is not initialised yet

    const int first_tile = (int)blockIdx.x / csize, tile_stride = (int)gridDim.x / csize;
    const int my_tiles = (a.num_tiles - first_tile + tile_stride - 1) / tile_stride;
    const int chunks_per_step = (a.L + 1) * 2 * kChunksPerHalf;

    if (warp == kConsumerWarps) {
        // ---------------- producer warp: stream the weights, forever in order ----------
        if (lane == 0) {
            const float* stream = a.packed + SmallParams::count;
            const long long total = (long long)my_tiles * a.steps * chunks_per_step;
            Pipe p;
            int c = 0;
            for (long long i = 0; i < total; ++i) {
                mbar_wait(&S.empty[p.stage], p.phase ^ 1);
                mbar_arrive_expect_tx(&S.full[p.stage], kChunkFloats * 4);
                bulk_g2s(S.Ws[p.stage], stream + (size_t)c * kChunkFloats, kChunkFloats * 4, &S.full[p.stage]);
                if (++c == chunks_per_step) c = 0;
                p.advance();
            }
        }
        return;
    }

    // ---------------- consumer warps --------------------------------------------------
    const int ty = (warp >> 1) * 4 + (lane >> 3);   // 0..15 : rows R0..R0+7
    const int tx = (warp & 1) * 8 + (lane & 7);     // 0..15 : columns col_of(tx, 0..7)
    const int R0 = ty * 8;
    const int sw = tx & 7;                          // swizzle key of all 8 of the thread's features
    const int c0 = ty * 2, c1 = ty * 2 + 1;
    int xo[8];
#pragma unroll
    for (int s = 0; s < 8; ++s) xo[s] = (c0 ^ s) << 2;
    const int radius = (R > 0) ? R : a.radius;
    const float inv_deg = 1.0f / (float)(2 * radius);
    int cl = 0, cr = 0;                             // chunk indices of the wrapped window ends
    if (R > 0) {
        const int seg = a.whole_ic ? nx : kTileRows;
        const int seg_start = (R0 / seg) * seg;
        int left = R0 - 4, right = R0 + 8;
        if (R0 == seg_start) left += seg;
        if (right == seg_start + seg) right -= seg;
        cl = (left & (kTileRows - 1)) >> 2;
        cr = (right & (kTileRows - 1)) >> 2;
    }
    const int used_rows = a.whole_ic ? a.ics_per_tile * nx : kTileRows;
    Pipe pipe;
    // group geometry: one group of 256 threads / 128 rows, or two groups of 128 threads / 64 rows
    const int grp = a.split ? (warp >> 2) : 0;
    const int gthreads = a.split ? kConsumers / 2 : kConsumers;
    const int lt = a.split ? (tid & (kConsumers / 2 - 1)) : tid;      // thread index inside the group
    const int nrows = a.split ? kTileRows / 2 : kTileRows;
    const int row0 = grp * (kTileRows / 2);
    const int myrow = (lt < nrows) ? row0 + lt : -1;                   // the row this thread looks after
    const int bar = 1 + grp;
    const bool leads_skew = a.split && grp == 0;
    const TileRows T{S.sN, S.sU, S.sE, S.sX, S.sF, S.sRho, S.gtab, S.rowIC, S.rowCell, S.prevRow, S.nextRow};
    // cluster mode: rows -4..-1 of the first row group are rows 120..123 of the left piece (4-row overlap), rows
    // 128..131 of the last row group are rows 4..7 of the right piece
    uint32_t zl = 0, zr_remote = 0, sig_l = 0, sig_r = 0, free_l = 0, free_r = 0;
    uint32_t zready_parity = 0, zfree_parity = 0;
    bool z_pending = false;                          // a Z generation of this CTA may still be read by a neighbour
    if constexpr (kCluster) {
        if (hasL) {
            if (ty == 0) { zl = cluster_map(smem_u32(S.Zs), crank - 1); cl = 30; }
            sig_l = cluster_map(smem_u32(&S.zready), crank - 1);
            free_l = cluster_map(smem_u32(&S.zfree), crank - 1);
        }
        if (hasR) {
            if (ty == 15) { zr_remote = cluster_map(smem_u32(S.Zs), crank + 1); cr = 1; }
            sig_r = cluster_map(smem_u32(&S.zready), crank + 1);
            free_r = cluster_map(smem_u32(&S.zfree), crank + 1);
        }
    }
    if (a.split && grp == 1) mbar_wait(&S.skew, 0);
#ifdef FLUXGNN_FFMA_TIMING
    const long long t_start = clock64();
#endif

    for (int tile = first_tile; tile < a.num_tiles; tile += tile_stride) {
        // ---- row bookkeeping + state load -----------------------------------------
        if (myrow >= 0) tile_load_row(a, T, tile, true, myrow, myrow, 0, kTileRows, crank);
        named_sync(bar, gthreads);

        for (int step = 0; step < a.steps; ++step) {
            float acc[8][8];

            // ---- input MLP: h0 = relu(W_in [n,u,E,x] + b_in)  (src/flux_gnn.py:49) ----
            {
                float fn[8], fu[8], fe[8], fx[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    fn[i] = S.sN[R0 + i]; fu[i] = S.sU[R0 + i]; fe[i] = S.sE[R0 + i]; fx[i] = S.sX[R0 + i];
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int n = col_of(tx, j);
                    const float w0 = __ldg(a.packed + SmallParams::w_in + 0 * kH + n);
                    const float w1 = __ldg(a.packed + SmallParams::w_in + 1 * kH + n);
                    const float w2 = __ldg(a.packed + SmallParams::w_in + 2 * kH + n);
                    const float w3 = __ldg(a.packed + SmallParams::w_in + 3 * kH + n);
                    const float b = __ldg(a.packed + SmallParams::b_in + n);
                    float h[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        float v = fmaf(w0, fn[i], b);
                        v = fmaf(w1, fu[i], v);
                        v = fmaf(w2, fe[i], v);
                        v = fmaf(w3, fx[i], v);
                        h[i] = fmaxf(v, 0.f);
                    }
                    float* hr = S.Hs + n * kTileRows;
                    *reinterpret_cast<float4*>(hr + ((c0 ^ sw) << 2)) = make_float4(h[0], h[1], h[2], h[3]);
                    *reinterpret_cast<float4*>(hr + ((c1 ^ sw) << 2)) = make_float4(h[4], h[5], h[6], h[7]);
                }
            }
            named_sync(bar, gthreads);
            if (kSave) save_rows(S.Hs, a.acts, S, lt, gthreads, row0, nrows, nx);

            for (int layer = 0; layer <= a.L; ++layer) {
                // ---- pass 0: Z = W[:, H:] h -> shared memory;  pass 1: Y = W[:, :H] h + b -> registers
                const float* bias = a.packed + (layer < a.L ? SmallParams::b_upd + layer * kH : SmallParams::b_e1);
#pragma unroll 1
                for (int pass = 0; pass < 2; ++pass) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float b = pass ? __ldg(bias + col_of(tx, j)) : 0.f;
#pragma unroll
                        for (int i = 0; i < 8; ++i) acc[i][j] = b;
                    }
                    gemm_pass(acc, S, pipe, xo, tx, lane, leads_skew);
                    if (pass == 0) {
                        if (kCluster && z_pending) {         // the neighbours have read the previous layer's edge rows
                            mbar_wait_cluster(&S.zfree, zfree_parity);
                            zfree_parity ^= 1;
                        }
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            float* zr = S.Zs + col_of(tx, j) * kTileRows;
                            *reinterpret_cast<float4*>(zr + ((c0 ^ sw) << 2)) = make_float4(acc[0][j], acc[1][j], acc[2][j], acc[3][j]);
                            *reinterpret_cast<float4*>(zr + ((c1 ^ sw) << 2)) = make_float4(acc[4][j], acc[5][j], acc[6][j], acc[7][j]);
                        }
                        // cluster mode: the left neighbour reads rows 4..7 (written by warps 0 and 1), the right one rows
                        // 120..123 (warps 6 and 7).  Tell them as soon as those rows are stored: the signal then travels
                        // under the whole second GEMM pass instead of being waited for.
                        if constexpr (kCluster) {
                            if (warp < 2) {
                                named_sync(3, 64);
                                if (tid == 0 && hasL) mbar_arrive_remote(sig_l);
                            } else if (warp >= kConsumerWarps - 2) {
                                named_sync(4, 64);
                                if (tid == (kConsumerWarps - 2) * 32 && hasR) mbar_arrive_remote(sig_r);
                            }
                        }
                    }
                }
                named_sync(bar, gthreads);      // Z complete; nobody reads Hs any more
                if constexpr (kCluster) {       // the neighbours' edge rows of Z (signalled one GEMM pass ago)
                    mbar_wait_cluster(&S.zready, zready_parity);
                    zready_parity ^= 1;
                    z_pending = true;
                }

                if (layer < a.L) {
                    // ---- node update: h' = relu(Y + mean_{|k|<=r, k!=0} Z_{i+k})  (src/flux_gnn.py:55-60)
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int n = col_of(tx, j);
                        const float* zr = S.Zs + n * kTileRows;
                        float h[8];
                        if (R > 0) {
                            float v[16];
                            load_window(v, zr, sw, cl, c0, c1, cr, kCluster ? zl : 0u, kCluster ? zr_remote : 0u, n);
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                const int c = 4 + i;
                                float s = v[c + 1] + v[c - 1];
#pragma unroll
                                for (int k = 2; k <= R; ++k) { s += v[c + k]; s += v[c - k]; }
                                h[i] = fmaxf(fmaf(s, inv_deg, acc[i][j]), 0.f);
                            }
                        } else {
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                int rp = R0 + i, rm = R0 + i;
                                float s = 0.f;
                                for (int k = 0; k < radius; ++k) {
                                    rp = S.nextRow[rp]; rm = S.prevRow[rm];
                                    s += S.Zs[sw_off(n, rp)];
                                    s += S.Zs[sw_off(n, rm)];
                                }
                                h[i] = fmaxf(fmaf(s, inv_deg, acc[i][j]), 0.f);
                            }
                        }
                        float* hr = S.Hs + n * kTileRows;
                        *reinterpret_cast<float4*>(hr + ((c0 ^ sw) << 2)) = make_float4(h[0], h[1], h[2], h[3]);
                        *reinterpret_cast<float4*>(hr + ((c1 ^ sw) << 2)) = make_float4(h[4], h[5], h[6], h[7]);
                    }
                    named_sync(bar, gthreads);  // h' complete; Z free
                    if (kCluster && tid == 0) {
                        if (hasL) mbar_arrive_remote(free_l);
                        if (hasR) mbar_arrive_remote(free_r);
                    }
                    if (kSave)
                        save_rows(S.Hs, a.acts + (size_t)(layer + 1) * a.acts_stride, S, lt, gthreads, row0, nrows, nx);
                } else {
                    // ---- edge readout (src/flux_gnn.py:63-66): acc = P + b1, Zs = Q ---------
                    //   fwd edge (row j, col j+k):  w2 . relu(P_j + Q_{j+k})
                    //   bwd edge (row j, col j-k):  w2 . relu(P_j + Q_{j-k})   (edge index j-k)
                    if (kSave) {                         // training forward: keep P + b1 and Q for the backward pass
                        save_rows(S.Zs, a.acts + (size_t)(a.L + 2) * a.acts_stride, S, lt, gthreads, row0, nrows, nx);
#pragma unroll
                        for (int j = 0; j < 8; ++j) {    // h^L is dead: park P in its buffer
                            float* hr = S.Hs + col_of(tx, j) * kTileRows;
                            *reinterpret_cast<float4*>(hr + ((c0 ^ sw) << 2)) = make_float4(acc[0][j], acc[1][j], acc[2][j], acc[3][j]);
                            *reinterpret_cast<float4*>(hr + ((c1 ^ sw) << 2)) = make_float4(acc[4][j], acc[5][j], acc[6][j], acc[7][j]);
                        }
                        named_sync(bar, gthreads);
                        save_rows(S.Hs, a.acts + (size_t)(a.L + 1) * a.acts_stride, S, lt, gthreads, row0, nrows, nx);
                    }
#pragma unroll
                    for (int hop = 1; hop <= kMaxHops; ++hop) {
                        if (hop <= a.hops) {
                            float pf[8], pb[8];
#pragma unroll
                            for (int i = 0; i < 8; ++i) { pf[i] = 0.f; pb[i] = 0.f; }
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                const int n = col_of(tx, j);
                                const float w2 = __ldg(a.packed + SmallParams::w_e2 + n);
                                const float* zr = S.Zs + n * kTileRows;
                                if (R > 0) {
                                    float v[16];
                                    load_window(v, zr, sw, cl, c0, c1, cr, kCluster ? zl : 0u, kCluster ? zr_remote : 0u, n);
#pragma unroll
                                    for (int i = 0; i < 8; ++i) {
                                        pf[i] = fmaf(w2, fmaxf(acc[i][j] + v[4 + i + hop], 0.f), pf[i]);
                                        pb[i] = fmaf(w2, fmaxf(acc[i][j] + v[4 + i - hop], 0.f), pb[i]);
                                    }
                                } else {
#pragma unroll
                                    for (int i = 0; i < 8; ++i) {
                                        int rp = R0 + i, rm = R0 + i;
                                        for (int k = 0; k < hop; ++k) { rp = S.nextRow[rp]; rm = S.prevRow[rm]; }
                                        pf[i] = fmaf(w2, fmaxf(acc[i][j] + S.Zs[sw_off(n, rp)], 0.f), pf[i]);
                                        pb[i] = fmaf(w2, fmaxf(acc[i][j] + S.Zs[sw_off(n, rm)], 0.f), pb[i]);
                                    }
                                }
                            }
                            // reduce over the 8 lanes that share these rows (lane bits 0..2)
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
#pragma unroll
                                for (int m = 1; m <= 4; m <<= 1) {
                                    pf[i] += __shfl_xor_sync(0xffffffffu, pf[i], m);
                                    pb[i] += __shfl_xor_sync(0xffffffffu, pb[i], m);
                                }
                            }
                            if ((lane & 7) == 0) {
#pragma unroll
                                for (int i = 0; i < 8; ++i) {
                                    S.edge[warp & 1][0][hop - 1][R0 + i] = pf[i];
                                    S.edge[warp & 1][1][hop - 1][R0 + i] = pb[i];
                                }
                            }
                        }
                    }
                    named_sync(bar, gthreads);
                    if (kCluster && tid == 0) {
                        if (hasL) mbar_arrive_remote(free_l);
                        if (hasR) mbar_arrive_remote(free_r);
                    }
                }
            }   // layers

            // ---- per-row: directed-edge fluxes, face flux (src/hybrid_solver.py:45-48) ----
            float n_new = 0.f, u_new = 0.f;
            if (myrow >= 0) {
                const int j = myrow;
                const float b2 = __ldg(a.packed + SmallParams::b_e2);
                const int ic = S.rowIC[j], cell = S.rowCell[j];
                float face = 0.f;
                int jn = j;
                for (int hop = 1; hop <= a.hops; ++hop) {
                    jn = S.nextRow[jn];
                    const float fwd = (S.edge[0][0][hop - 1][j] + S.edge[1][0][hop - 1][j]) + b2;
                    const float bwd = (S.edge[0][1][hop - 1][jn] + S.edge[1][1][hop - 1][jn]) + b2;
                    if (hop == 1) face = 0.5f * (fwd + bwd);
                    if (a.flux_edges != nullptr && ic >= 0) {
                        float* fe = a.flux_edges + (size_t)ic * 2 * a.hops * nx + (size_t)(2 * (hop - 1)) * nx + cell;
                        fe[0] = fwd;
                        fe[nx] = bwd;
                    }
                }
                if (a.face_flux != nullptr && ic >= 0) a.face_flux[(size_t)ic * nx + cell] = face;
                S.sF[j] = face;
            }
            if (!a.do_update) continue;          // forward only (steps == 1)
            named_sync(bar, gthreads);

            // ---- finite-volume update (src/hybrid_solver.py:51-58) ----
            if (myrow >= 0) tile_fv_update(a, T, myrow, n_new, u_new);
            if (!a.whole_ic) {
                // window tiles: E' comes from the separate field-solve kernel
                if (myrow >= 0) tile_store_window_row(a, T, myrow, n_new, u_new);
                continue;
            }
            named_sync(bar, gthreads);           // everyone has read the old n, u
            if (myrow >= 0) tile_keep_row(T, myrow, n_new, u_new);
            named_sync(bar, gthreads);
            // ---- field solve: E = g (*) rho, two threads per row (src/baseline_solver.py:59-68) ----
            {
                const int row = row0 + (lt >> 1), half = lt & 1;
                double e = (row < used_rows) ? tile_field_partial(T, row, half, 2, nx) : 0.0;
                e += __shfl_xor_sync(0xffffffffu, e, 1);
                if (half == 0) S.sE[row] = (float)e;
            }
            named_sync(bar, gthreads);
            // ---- write-out: last step and recorded steps ---------------------------------
            if (myrow >= 0) tile_write_out_row(a, T, myrow, step);
        }   // steps
        named_sync(bar, gthreads);   // state arrays are rewritten by the next tile's load
    }       // tiles
    if (kCluster && z_pending) mbar_wait_cluster(&S.zfree, zfree_parity);   // shared memory must outlive the neighbours' reads
#ifdef FLUXGNN_FFMA_TIMING
    if (blockIdx.x == 0 && lane == 0 && (warp & 3) == 0)
        printf("[ffma timing] warp %d (group %d): total %lld clk, in GEMM passes %lld, of which waiting for weights %lld\n",
               warp, grp, clock64() - t_start, pipe.t_gemm, pipe.t_wait);
#endif
}

// ---------------------------------------------------------------------------
// host side: one translation unit per radius instantiates this (hybrid_r*.cu)
// ---------------------------------------------------------------------------
template <int R, bool kSave>
cudaError_t launch_one(const HybridArgs& a, int grid, cudaStream_t stream) {
    // per device and per function; cheap and idempotent, so set it on every launch
    cudaError_t e = cudaFuncSetAttribute(hybrid_tile_kernel<R, kSave, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)sizeof(TileSmem));
    if (e != cudaSuccess) return e;
    hybrid_tile_kernel<R, kSave, false><<<grid, kThreads, sizeof(TileSmem), stream>>>(a);
    return cudaGetLastError();
}

// the cluster instantiation (inference, compile-time radius): grid = clusters x a.cluster CTAs
template <int R>
cudaError_t launch_one_cluster(const HybridArgs& a, int grid, cudaStream_t stream) {
    cudaError_t e = cudaFuncSetAttribute(hybrid_tile_kernel<R, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)sizeof(TileSmem));
    if (e != cudaSuccess) return e;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = sizeof(TileSmem);
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)a.cluster;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, hybrid_tile_kernel<R, false, true>, a);
}

// How many clusters of `csize` CTAs of this kernel the device holds at once (one CTA per SM by shared memory; clusters
// must sit inside a GPC, so this can be less than #SMs / csize).
template <int R>
cudaError_t max_clusters_one(int csize, int* out) {
    cudaError_t e = cudaFuncSetAttribute(hybrid_tile_kernel<R, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)sizeof(TileSmem));
    if (e != cudaSuccess) return e;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(csize * 64));
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = sizeof(TileSmem);
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)csize;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaOccupancyMaxActiveClusters(out, hybrid_tile_kernel<R, false, true>, &cfg);
}

}  // namespace fluxgnn
