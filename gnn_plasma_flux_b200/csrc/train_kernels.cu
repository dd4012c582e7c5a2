// Backward pass of FluxGNN.forward on the radius-r ring: gradients of the
// directed-edge fluxes w.r.t. every parameter and the node features, from activations the
// forward tile kernel saved.  What it differentiates (paths under /root/reference):
//   src/flux_gnn.py:49      h0 = relu(X W_in^T + b_in)
//   src/flux_gnn.py:53-60   h' = relu(Wa h + Wb mean_nbr(h) + b)          (W = [Wa | Wb])
//   src/flux_gnn.py:63-66   f(a->b) = w2 . relu(W1a h_a + W1b h_b + b1) + b2
// used by the training loop scripts/training/train_ablation.py:128-206 through autograd.
//
// All tensors are row-major [row][feature] with row = ic * nx + cell.  mean_nbr is symmetric,
// so with dpre = dh' * (h' > 0) and dZ = mean_nbr(dpre):
//   dh = dpre Wa + dZ Wb,   dWa = dpre^T h,   dWb = dZ^T h,   db = sum_rows dpre.
#include "common.cuh"
#include "train_kernels.cuh"

namespace fluxgnn {

namespace {

__device__ __forceinline__ float4 relu_mask(float4 g, float4 h) {
    return make_float4(h.x > 0.f ? g.x : 0.f, h.y > 0.f ? g.y : 0.f, h.z > 0.f ? g.z : 0.f, h.w > 0.f ? g.w : 0.f);
}

__device__ __forceinline__ void add4(float4& a, float4 b) { a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w; }

// one feature of the edge-readout backward: the four pre-activations row i takes part in
__device__ __forceinline__ void edge_acc(float& gp, float& gq, float& gw, float w, float s_f, float s_bm, float s_fm,
                                         float s_b, float gf_i, float gf_m, float gb_i, float gb_m) {
    gp += w * ((s_f > 0.f ? gf_i : 0.f) + (s_bm > 0.f ? gb_m : 0.f));
    gq += w * ((s_fm > 0.f ? gf_m : 0.f) + (s_b > 0.f ? gb_i : 0.f));
    gw += gf_i * fmaxf(s_f, 0.f) + gb_i * fmaxf(s_b, 0.f);
}

}  // namespace

// dpre = dH * (H' > 0);  dZ = (1/2r) sum_{k=1..r} (dpre_{i-k} + dpre_{i+k});  db += sum_rows dpre.
// One thread per (row, 4 features); blockDim = (32, 8): 32 feature quads x 8 rows.
__global__ void __launch_bounds__(256) bwd_mask_mean_kernel(const float* __restrict__ dH, const float* __restrict__ Hn,
                                                            float* __restrict__ dpre, float* __restrict__ dZ,
                                                            float* __restrict__ db, long long rows, int nx, int radius) {
    const int fq = threadIdx.x;                                   // feature quad 0..31
    float4 bsum = make_float4(0.f, 0.f, 0.f, 0.f);                // this thread's share of the bias gradient
    for (long long row = (long long)blockIdx.x * blockDim.y + threadIdx.y; row < rows;
         row += (long long)gridDim.x * blockDim.y) {
        const long long ic_base = (row / nx) * nx;
        const int cell = (int)(row - ic_base);
        const float4* g4 = reinterpret_cast<const float4*>(dH);
        const float4* h4 = reinterpret_cast<const float4*>(Hn);
        const float4 mine = relu_mask(g4[row * 32 + fq], h4[row * 32 + fq]);
        add4(bsum, mine);
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int k = 1; k <= radius; ++k) {
            const long long rp = ic_base + (cell + k) % nx;
            long long rm = cell - k;
            rm = ic_base + ((rm % nx) + nx) % nx;
            add4(acc, relu_mask(g4[rp * 32 + fq], h4[rp * 32 + fq]));
            add4(acc, relu_mask(g4[rm * 32 + fq], h4[rm * 32 + fq]));
        }
        const float inv = 1.0f / (float)(2 * radius);
        reinterpret_cast<float4*>(dpre)[row * 32 + fq] = mine;
        reinterpret_cast<float4*>(dZ)[row * 32 + fq] = make_float4(acc.x * inv, acc.y * inv, acc.z * inv, acc.w * inv);
    }
    // bias gradient: reduce the 8 row-lanes of the block, then one atomic per feature and block
    __shared__ float4 red[8][32];
    red[threadIdx.y][fq] = bsum;
    __syncthreads();
    if (threadIdx.y == 0) {
        float4 s = red[0][fq];
        for (int r = 1; r < 8; ++r) add4(s, red[r][fq]);
        atomicAdd(db + 4 * fq + 0, s.x);
        atomicAdd(db + 4 * fq + 1, s.y);
        atomicAdd(db + 4 * fq + 2, s.z);
        atomicAdd(db + 4 * fq + 3, s.w);
    }
}

// ---- the two GEMM shapes of the backward pass ----------------------------------------------------
// Both are 128 x 128 output tiles over K slabs of 16 held k-major in shared memory (As[k][m], Bs[k][n]), 256 threads
// with an 8 x 8 register tile each: thread (tm, tn) owns rows {4 tm .. 4 tm + 3} u {64 + 4 tm ..} and columns
// {4 tn ..} u {64 + 4 tn ..}, so every operand load is a conflict-free LDS.128 (a quarter warp reads one broadcast
// row chunk and eight consecutive column chunks).  Slabs are double-buffered: the next slab travels global ->
// registers while the current one is multiplied, one __syncthreads per slab.
constexpr int kSlab = 16;

struct SlabRegs {
    float4 a0, a1, b0, b1;
};

// Packed accumulators (fma.rn.f32x2, SASS FFMA2: one issue slot per two multiply-adds, which is what lets the operand
// loads issue in the shadow of the arithmetic): for row pair p and column pair q
//   d[p][q] = {acc[2p][2q],   acc[2p+1][2q+1]},   x[p][q] = {acc[2p][2q+1], acc[2p+1][2q]}
// (a pair) * (b pair) and (a pair) * (swapped b pair) -- the scheme of the forward kernel's gemm_pass.
struct PackedAcc {
    float2 d[4][4], x[4][4];
    __device__ __forceinline__ void clear() {
#pragma unroll
        for (int p = 0; p < 4; ++p)
#pragma unroll
            for (int q = 0; q < 4; ++q) d[p][q] = x[p][q] = make_float2(0.f, 0.f);
    }
    __device__ __forceinline__ void unpack(float (&acc)[8][8]) const {
#pragma unroll
        for (int p = 0; p < 4; ++p)
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                acc[2 * p][2 * q] = d[p][q].x;
                acc[2 * p + 1][2 * q + 1] = d[p][q].y;
                acc[2 * p][2 * q + 1] = x[p][q].x;
                acc[2 * p + 1][2 * q] = x[p][q].y;
            }
    }
};

__device__ __forceinline__ void mma_slab(PackedAcc& c, const float (*As)[128], const float (*Bs)[128], int tm, int tn) {
#pragma unroll
    for (int k = 0; k < kSlab; ++k) {
        const float4 a0 = *reinterpret_cast<const float4*>(&As[k][4 * tm]);
        const float4 a1 = *reinterpret_cast<const float4*>(&As[k][64 + 4 * tm]);
        const float4 b0 = *reinterpret_cast<const float4*>(&Bs[k][4 * tn]);
        const float4 b1 = *reinterpret_cast<const float4*>(&Bs[k][64 + 4 * tn]);
        const float2 A[4] = {make_float2(a0.x, a0.y), make_float2(a0.z, a0.w), make_float2(a1.x, a1.y), make_float2(a1.z, a1.w)};
        const float2 B[4] = {make_float2(b0.x, b0.y), make_float2(b0.z, b0.w), make_float2(b1.x, b1.y), make_float2(b1.z, b1.w)};
        const float2 Bx[4] = {make_float2(b0.y, b0.x), make_float2(b0.w, b0.z), make_float2(b1.y, b1.x), make_float2(b1.w, b1.z)};
#pragma unroll
        for (int p = 0; p < 4; ++p)
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                c.d[p][q] = __ffma2_rn(A[p], B[q], c.d[p][q]);
                c.x[p][q] = __ffma2_rn(A[p], Bx[q], c.x[p][q]);
            }
    }
}

// C[rows][128] = A1[rows][128] W1[128][ldw] + A2[rows][128] W2[128][ldw]   (W row-major, first 128 columns used)
// One block per 128 rows; the 16 K slabs run over A1/W1 then A2/W2.  A is row-major, so its slab is transposed on
// the way into shared memory (thread = one row, 8 consecutive k: 32 consecutive rows per warp -> conflict-free stores).
__global__ void __launch_bounds__(256, 2) bwd_gemm_nn_kernel(const float* __restrict__ A1, const float* __restrict__ W1,
                                                             const float* __restrict__ A2, const float* __restrict__ W2,
                                                             int ldw, float* __restrict__ C, long long rows) {
    __shared__ __align__(16) float As[2][kSlab][128];     // [k][row]
    __shared__ __align__(16) float Bs[2][kSlab][128];     // [k][col]
    const int tid = threadIdx.x;
    const int tm = tid >> 4, tn = tid & 15;
    const long long row0 = (long long)blockIdx.x * 128;
    const int lrow = tid & 127, lkq = (tid >> 7) * 8;             // A loader: row, first of 8 k
    const int wk = tid >> 4, wc = (tid & 15) * 8;                  // W loader: k, first of 8 columns
    const bool row_ok = row0 + lrow < rows;
    const int slabs = (A2 != nullptr ? 2 : 1) * (kH / kSlab);
    PackedAcc pacc;
    pacc.clear();
    SlabRegs g;
    auto fetch = [&](int sl) {
        const float* A = sl < kH / kSlab ? A1 : A2;
        const float* W = sl < kH / kSlab ? W1 : W2;
        const int k0 = (sl % (kH / kSlab)) * kSlab;
        g.a0 = g.a1 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row_ok) {
            g.a0 = *reinterpret_cast<const float4*>(A + (row0 + lrow) * kH + k0 + lkq);
            g.a1 = *reinterpret_cast<const float4*>(A + (row0 + lrow) * kH + k0 + lkq + 4);
        }
        g.b0 = *reinterpret_cast<const float4*>(W + (size_t)(k0 + wk) * ldw + wc);
        g.b1 = *reinterpret_cast<const float4*>(W + (size_t)(k0 + wk) * ldw + wc + 4);
    };
    auto stash = [&](int buf) {
        As[buf][lkq + 0][lrow] = g.a0.x; As[buf][lkq + 1][lrow] = g.a0.y; As[buf][lkq + 2][lrow] = g.a0.z; As[buf][lkq + 3][lrow] = g.a0.w;
        As[buf][lkq + 4][lrow] = g.a1.x; As[buf][lkq + 5][lrow] = g.a1.y; As[buf][lkq + 6][lrow] = g.a1.z; As[buf][lkq + 7][lrow] = g.a1.w;
        *reinterpret_cast<float4*>(&Bs[buf][wk][wc]) = g.b0;
        *reinterpret_cast<float4*>(&Bs[buf][wk][wc + 4]) = g.b1;
    };
    fetch(0);
    stash(0);
    __syncthreads();
    for (int sl = 0; sl < slabs; ++sl) {
        const int buf = sl & 1;
        if (sl + 1 < slabs) fetch(sl + 1);
        mma_slab(pacc, As[buf], Bs[buf], tm, tn);
        if (sl + 1 < slabs) stash(buf ^ 1);
        __syncthreads();
    }
    float acc[8][8];
    pacc.unpack(acc);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const long long r = row0 + (i < 4 ? 4 * tm + i : 64 + 4 * tm + (i - 4));
        if (r < rows) {
            *reinterpret_cast<float4*>(C + r * kH + 4 * tn) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
            *reinterpret_cast<float4*>(C + r * kH + 64 + 4 * tn) = make_float4(acc[i][4], acc[i][5], acc[i][6], acc[i][7]);
        }
    }
}

// dW[n][ldw-strided k] += sum_rows A[row][n] * Bm[row][k]   (both [rows][128]); split over row ranges.
// Persistent blocks: each walks its share of the rows in 16-row slabs (both operands are k-major as stored), keeps
// the full 128 x 128 partial product in registers and issues its atomicAdds once at the end.
__global__ void __launch_bounds__(256, 2) bwd_gemm_tn_kernel(const float* __restrict__ A, const float* __restrict__ Bm,
                                                             float* __restrict__ dW, int ldw, long long rows) {
    __shared__ __align__(16) float As[2][kSlab][128];     // [row in slab][n]
    __shared__ __align__(16) float Bs[2][kSlab][128];     // [row in slab][k]
    const int tid = threadIdx.x;
    const int tm = tid >> 4, tn = tid & 15;
    const long long per_block = ((rows + gridDim.x - 1) / gridDim.x + kSlab - 1) / kSlab * kSlab;
    const long long row_begin = (long long)blockIdx.x * per_block;
    const long long row_end = row_begin + per_block < rows ? row_begin + per_block : rows;
    if (row_begin >= rows) return;
    const int lr = tid >> 4, lc = (tid & 15) * 8;                  // loader: row in slab, first of 8 features
    PackedAcc pacc;
    pacc.clear();
    SlabRegs g;
    auto fetch = [&](long long r0) {
        g.a0 = g.a1 = g.b0 = g.b1 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r0 + lr < row_end) {
            g.a0 = *reinterpret_cast<const float4*>(A + (r0 + lr) * kH + lc);
            g.a1 = *reinterpret_cast<const float4*>(A + (r0 + lr) * kH + lc + 4);
            g.b0 = *reinterpret_cast<const float4*>(Bm + (r0 + lr) * kH + lc);
            g.b1 = *reinterpret_cast<const float4*>(Bm + (r0 + lr) * kH + lc + 4);
        }
    };
    auto stash = [&](int buf) {
        *reinterpret_cast<float4*>(&As[buf][lr][lc]) = g.a0;
        *reinterpret_cast<float4*>(&As[buf][lr][lc + 4]) = g.a1;
        *reinterpret_cast<float4*>(&Bs[buf][lr][lc]) = g.b0;
        *reinterpret_cast<float4*>(&Bs[buf][lr][lc + 4]) = g.b1;
    };
    fetch(row_begin);
    stash(0);
    __syncthreads();
    int buf = 0;
    for (long long r0 = row_begin; r0 < row_end; r0 += kSlab) {
        const bool more = r0 + kSlab < row_end;
        if (more) fetch(r0 + kSlab);
        mma_slab(pacc, As[buf], Bs[buf], tm, tn);
        if (more) stash(buf ^ 1);
        __syncthreads();
        buf ^= 1;
    }
    float acc[8][8];
    pacc.unpack(acc);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int n = i < 4 ? 4 * tm + i : 64 + 4 * tm + (i - 4);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int k = j < 4 ? 4 * tn + j : 64 + 4 * tn + (j - 4);
            atomicAdd(dW + (size_t)n * ldw + k, acc[i][j]);
        }
    }
}

// Edge readout backward.  P already contains b1.  For hop k (edge blocks [i -> i+k], [i+k -> i]):
//   fwd_k[i] = w2 . relu(P_i + Q_{i+k}) + b2,   bwd_k[i] = w2 . relu(P_{i+k} + Q_i) + b2
// dflux layout per IC: [2*hops][nx].  Outputs dP, dQ [rows][128]; accumulates dw2[128], db1[128], db2[1].
__global__ void __launch_bounds__(256) bwd_edge_kernel(const float* __restrict__ P, const float* __restrict__ Q,
                                                       const float* __restrict__ w2, const float* __restrict__ dflux,
                                                       float* __restrict__ dP, float* __restrict__ dQ,
                                                       float* __restrict__ dw2, float* __restrict__ db1,
                                                       float* __restrict__ db2, long long rows, int nx, int hops) {
    const int fq = threadIdx.x;
    float4 gw = make_float4(0.f, 0.f, 0.f, 0.f), gb1 = gw;       // accumulated over this thread's rows
    float gb2 = 0.f;
    for (long long row = (long long)blockIdx.x * blockDim.y + threadIdx.y; row < rows;
         row += (long long)gridDim.x * blockDim.y) {
        float4 gp = make_float4(0.f, 0.f, 0.f, 0.f), gq = gp;
        const long long ic = row / nx, ic_base = ic * nx;
        const int cell = (int)(row - ic_base);
        const float4* P4 = reinterpret_cast<const float4*>(P);
        const float4* Q4 = reinterpret_cast<const float4*>(Q);
        const float4 w = reinterpret_cast<const float4*>(w2)[fq];
        const float4 p0 = P4[row * 32 + fq], q0 = Q4[row * 32 + fq];
        const float* df = dflux + ic * 2 * hops * nx;
        for (int k = 1; k <= hops; ++k) {
            const int cp = (cell + k) % nx, cm = ((cell - k) % nx + nx) % nx;
            const long long rp = ic_base + cp, rm = ic_base + cm;
            const float4 pp = P4[rp * 32 + fq], qp = Q4[rp * 32 + fq];
            const float4 pm = P4[rm * 32 + fq], qm = Q4[rm * 32 + fq];
            const float* dfw = df + (size_t)(2 * (k - 1)) * nx;       // fwd block of hop k
            const float* dbw = dfw + nx;                              // bwd block of hop k
            const float gf_i = dfw[cell], gf_m = dfw[cm], gb_i = dbw[cell], gb_m = dbw[cm];
            // pre-activations this row takes part in
            const float4 s_f = make_float4(p0.x + qp.x, p0.y + qp.y, p0.z + qp.z, p0.w + qp.w);   // fwd_k[i]   : P_i + Q_{i+k}
            const float4 s_bm = make_float4(p0.x + qm.x, p0.y + qm.y, p0.z + qm.z, p0.w + qm.w);  // bwd_k[i-k] : P_i + Q_{i-k}
            const float4 s_fm = make_float4(pm.x + q0.x, pm.y + q0.y, pm.z + q0.z, pm.w + q0.w);  // fwd_k[i-k] : P_{i-k} + Q_i
            const float4 s_b = make_float4(pp.x + q0.x, pp.y + q0.y, pp.z + q0.z, pp.w + q0.w);   // bwd_k[i]   : P_{i+k} + Q_i
            edge_acc(gp.x, gq.x, gw.x, w.x, s_f.x, s_bm.x, s_fm.x, s_b.x, gf_i, gf_m, gb_i, gb_m);
            edge_acc(gp.y, gq.y, gw.y, w.y, s_f.y, s_bm.y, s_fm.y, s_b.y, gf_i, gf_m, gb_i, gb_m);
            edge_acc(gp.z, gq.z, gw.z, w.z, s_f.z, s_bm.z, s_fm.z, s_b.z, gf_i, gf_m, gb_i, gb_m);
            edge_acc(gp.w, gq.w, gw.w, w.w, s_f.w, s_bm.w, s_fm.w, s_b.w, gf_i, gf_m, gb_i, gb_m);
            if (fq == 0) gb2 += gf_i + gb_i;
        }
        reinterpret_cast<float4*>(dP)[row * 32 + fq] = gp;
        reinterpret_cast<float4*>(dQ)[row * 32 + fq] = gq;
        add4(gb1, gp);                                                // db1 = sum_rows dP
    }
    __shared__ float4 redw[8][32], redb[8][32];
    __shared__ float redb2[8];
    redw[threadIdx.y][fq] = gw;
    redb[threadIdx.y][fq] = gb1;
    if (fq == 0) redb2[threadIdx.y] = gb2;
    __syncthreads();
    if (threadIdx.y == 0) {
        float4 sw = redw[0][fq], sb = redb[0][fq];
        for (int r = 1; r < 8; ++r) { add4(sw, redw[r][fq]); add4(sb, redb[r][fq]); }
        atomicAdd(dw2 + 4 * fq + 0, sw.x); atomicAdd(dw2 + 4 * fq + 1, sw.y);
        atomicAdd(dw2 + 4 * fq + 2, sw.z); atomicAdd(dw2 + 4 * fq + 3, sw.w);
        atomicAdd(db1 + 4 * fq + 0, sb.x); atomicAdd(db1 + 4 * fq + 1, sb.y);
        atomicAdd(db1 + 4 * fq + 2, sb.z); atomicAdd(db1 + 4 * fq + 3, sb.w);
        if (fq == 0) {
            float s = 0.f;
            for (int r = 0; r < 8; ++r) s += redb2[r];
            atomicAdd(db2, s);
        }
    }
}

// Input layer backward: dpre0 = dH0 * (H0 > 0);  dfeat[row][f] = sum_n dpre0[n] W_in[n][f];
// dW_in[n][f] += dpre0[n] feat[f];  db_in[n] += dpre0[n].  feat = (n, u, E, x) of the row.
// One warp per row (lane owns features lane, lane+32, ...); 8 rows per block.
__global__ void __launch_bounds__(256) bwd_input_kernel(const float* __restrict__ dH0, const float* __restrict__ H0,
                                                        const float* __restrict__ w_in, const float* __restrict__ state,
                                                        const float* __restrict__ x, float* __restrict__ dstate,
                                                        float* __restrict__ dw_in, float* __restrict__ db_in,
                                                        long long rows, int nx) {
    const int lane = threadIdx.x, wy = threadIdx.y;
    __shared__ float sgw[8][kH][kF];
    __shared__ float sgb[8][kH];
    // per-thread accumulators over this warp's rows: features n = lane + 32 q
    float aw[4][4], ab[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        ab[q] = 0.f;
#pragma unroll
        for (int f = 0; f < 4; ++f) aw[q][f] = 0.f;
    }
    for (long long row = (long long)blockIdx.x * blockDim.y + wy; row < rows; row += (long long)gridDim.x * blockDim.y) {
        const long long ic = row / nx;
        const int cell = (int)(row - ic * nx);
        const float* st = state + ic * 3 * nx + cell;
        const float feat[4] = {st[0], st[nx], st[2 * (size_t)nx], x[cell]};
        float dfe[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int n = lane + 32 * q;
            const float g = H0[row * kH + n] > 0.f ? dH0[row * kH + n] : 0.f;
            ab[q] += g;
#pragma unroll
            for (int f = 0; f < 4; ++f) {
                dfe[f] = fmaf(g, w_in[n * kF + f], dfe[f]);
                aw[q][f] = fmaf(g, feat[f], aw[q][f]);
            }
        }
#pragma unroll
        for (int f = 0; f < 4; ++f)
            for (int o = 16; o > 0; o >>= 1) dfe[f] += __shfl_xor_sync(0xffffffffu, dfe[f], o);
        if (lane == 0 && dstate != nullptr) {
            float* ds = dstate + ic * 3 * nx + cell;
            ds[0] = dfe[0]; ds[nx] = dfe[1]; ds[2 * (size_t)nx] = dfe[2];      // x has no gradient
        }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        sgb[wy][lane + 32 * q] = ab[q];
#pragma unroll
        for (int f = 0; f < 4; ++f) sgw[wy][lane + 32 * q][f] = aw[q][f];
    }
    __syncthreads();
    const int t = wy * 32 + lane;                                          // 256 threads: 2 per feature n
    {
        const int n = t >> 1, f0 = (t & 1) * 2;
        float s0 = 0.f, s1 = 0.f, sb = 0.f;
        for (int r = 0; r < 8; ++r) { s0 += sgw[r][n][f0]; s1 += sgw[r][n][f0 + 1]; sb += sgb[r][n]; }
        atomicAdd(dw_in + n * kF + f0, s0);
        atomicAdd(dw_in + n * kF + f0 + 1, s1);
        if ((t & 1) == 0) atomicAdd(db_in + n, sb);
    }
}

// ---------------------------------------------------------------------------------------------
// Backward of the finite-volume part of one hybrid step (src/hybrid_solver.py:45-58 as the training
// rollout writes it, scripts/training/train_ablation.py:180-195):
//   F_i  = 0.5 (fwd_i + bwd_i)                       fwd_i = flux_edges[i], bwd_i = flux_edges[nx + i]
//   n'_i = n_i - c (F_i - F_{i-1})
//   u'_i = u_i - c (0.5 u_i^2 - 0.5 u_{i-1}^2) + dt E_i
//   E'   = field solve of n', DETACHED (train_ablation.py:198-200: it goes through numpy)
// step_bwd_flux_kernel: gradient w.r.t. the two hop-1 edge blocks from the gradients w.r.t. n' and
// (optionally) the face flux itself:  dF_i = gF_i - c (gn'_i - gn'_{i+1});  dfwd_i = dbwd_i = 0.5 dF_i.
__global__ void __launch_bounds__(256) step_bwd_flux_kernel(const float* __restrict__ g_out, const float* __restrict__ g_face,
                                                            float* __restrict__ dflux, long long cells, int nx, float c) {
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < cells; t += (long long)gridDim.x * blockDim.x) {
        const long long ic = t / nx;
        const int i = (int)(t - ic * nx);
        const int ip = (i + 1 == nx) ? 0 : i + 1;
        const float* gn = g_out + ic * 3 * nx;
        float dF = -c * (gn[i] - gn[ip]);
        if (g_face != nullptr) dF += g_face[t];
        float* d = dflux + ic * 2 * nx;
        d[i] = 0.5f * dF;
        d[nx + i] = 0.5f * dF;
    }
}

// step_bwd_direct_kernel: the paths that do not go through the network, ADDED to dstate (which already
// holds the network's input gradients):  dn_i += gn'_i;  du_i += gu'_i (1 - c u_i) + gu'_{i+1} c u_i;  dE_i += dt gu'_i.
__global__ void __launch_bounds__(256) step_bwd_direct_kernel(const float* __restrict__ g_out, const float* __restrict__ state,
                                                              float* __restrict__ dstate, long long cells, int nx, float c,
                                                              float dt) {
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < cells; t += (long long)gridDim.x * blockDim.x) {
        const long long ic = t / nx;
        const int i = (int)(t - ic * nx);
        const int ip = (i + 1 == nx) ? 0 : i + 1;
        const float* g = g_out + ic * 3 * nx;
        const float u = state[ic * 3 * nx + nx + i];
        float* d = dstate + ic * 3 * nx;
        const float gu = g[nx + i], gup = g[nx + ip];
        d[i] += g[i];
        d[nx + i] += gu * (1.0f - c * u) + gup * (c * u);
        d[2 * (size_t)nx + i] += dt * gu;
    }
}

}  // namespace fluxgnn
