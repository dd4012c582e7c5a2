// Tensor-core variant of the fused hybrid step with 16-bit operands and 256-row tiles
// (sm_100a: tcgen05.mma kind::f16 + TMEM + bulk-copy TMA).
//
// Shape (scripts/probes/umma_probe.cu, profiles/r1_p_umma_probe.txt): with a tight, warp-uniform issue
// loop one tcgen05.mma of M = 128 costs 128 clk at N = 256 and 64 clk at N = 128 for a 32-byte K step
// -- both at the pipe's peak (8.1 kflop/clk/SM for 16-bit operands, half of that for TF32) -- but
// 48 clk at N = 64 (66 %) and the same time at M = 64 (50 %).  So 16-bit operands (K = 16 per
// instruction) halve the tensor time of TF32, and N >= 128 keeps the pipe efficient.  The tile is 256
// rows, normally run as two independent N = 128 groups (kSplit below).
//
// The product is issued TRANSPOSED as in hybrid_tc_kernel.cu: D^T[n][i] = sum_k W[n][k] h[i][k];
// the accumulator has TMEM lane = feature, column = row, so the +-r window of the message-passing
// mean is register indexing.  D_Z = W[:, H:] h in TMEM columns [0,256), D_Y = W[:, :H] h in [256,512)
// (the whole tensor memory of the SM: one CTA per SM).
//
// Precision modes (HybridArgs::tc_parts / tc_format):
//   "fp16x3": h = h0 + h1, 256 W = w0 + w1 with every part an fp16 number (11 significant bits each,
//             22 together -- the same as the tf32x3 split);  D = w0 h0 + w0 h1 + w1 h0 accumulated in
//             fp32, then scaled by 2^-8 in the epilogue.  The power-of-two scale keeps w1 (~2^-11 |w0|)
//             in fp16's normal range; h1 falls into the subnormal range only for |h| < 0.125, where
//             the absolute error (<= 2^-25) is below the fp32 rounding noise of the sum.
//   "fp16"  : D = w0 h0 (one product, 2x the TF32 rate; same 11-bit operands as plain tf32)
//   "bf16"  : one product with bfloat16 operands (8 significant bits; loosest tolerance, widest range)
//
// Warp roles (576 threads): warps 0-15 epilogue (TMEM lane quadrant = warp % 4, 32-row chunk = warp / 4),
// warp 16 weight producer, warp 17 TMEM allocator + UMMA issuer.
//
// kSplit: the two 128-row halves of a CTA tile are independent LOGICAL tiles -- different ICs (nx <= 128)
// or two windows with their own halos (nx > 128) -- i.e. two GROUPS with their own barriers and their
// own columns of the accumulators.  The issuer alternates between them with N = 128 instructions and
// the 16 epilogue warps serve whichever group is ready, so while one half is in its epilogue (or in
// the finite-volume / field-solve tail) the tensor pipe works on the other.  N = 128 instructions are
// as efficient as N = 256 ones, so the split costs no tensor time.  Only very wide receptive fields
// (halo > 24 cells) keep one 256-row window per CTA (api.cu, plan_tiles).
//
// Epilogue arithmetic.  The epilogue warps are issue-bound, so their FP32 work is packed two rows per
// instruction where the register pairs line up: window sums from aligned pairs (FADD2), bias / unscale /
// mean as FFMA2, the hi/lo residual of the operand split as one mixed-precision FMA per element
// (fma.rn.f32.f16, SASS FHFMA).  Rollouts (no per-edge output) add the two terms of a face per feature
// before the warp transpose-sum, i.e. one feature reduction per face instead of two, and the
// finite-volume tail needs three barriers: a row reads its left neighbour's u before the first one and
// forms both of its faces itself.
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdlib.h>

#include "common.cuh"
#include "hybrid_kernel.cuh"
#include "tile_common.cuh"

namespace fluxgnn {

namespace {

constexpr int kRows = kTc16TileRows;              // 256 cells per tile (= UMMA N)
constexpr int kStages = 4;
constexpr int kEpiThreads = 512;
constexpr int kProducerWarp = 16, kMmaWarp = 17;
constexpr int kThreads = kEpiThreads + 64;
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kColZ = 0, kColY = 256;
constexpr int kActPartBytes = kRows * 128 * 2;    // one part (hi or lo) of the activations: 256 rows x 128 k x 2 B
constexpr float kUnscale = 1.0f / kTc16WeightScale;

struct __align__(1024) Smem {
    unsigned char act[2][kActPartBytes];          // [part]: 64 KiB of activations, the MN-major UMMA B operand
                                                  //   ([64-row block][k atom][1024 B], see umma_desc_b_mn)
    unsigned char Ws[kStages][kTc16UnitBytes];    // streamed weight operand images (UMMA A operand)
    float small[SmallParams::count];
    float sN[kRows], sU[kRows], sE[kRows], sX[kRows];
    float sF[kRows], sRho[kRows];
    float edgeP[4][2][kRows];                     // per feature quadrant: partial fwd / bwd dot products of each row
    double gtab[128];
    int rowIC[kRows];
    int rowCell[kRows];
    short prevRow[kRows], nextRow[kRows];
    uint64_t full[kStages], empty[kStages];
    uint64_t act_ready[2], acc_ready[2];          // per group (kSplit: the two 128-row halves)
    uint32_t tmem_base;
#ifdef FLUXGNN_TC_TIMING
    long long timing[16];
#endif
};
static_assert(sizeof(Smem) + 1024 <= 227 * 1024, "tensor tile does not fit shared memory");

#ifdef FLUXGNN_TC_TIMING
__device__ long long g_tc16_timing[16];
#define TC_TICK(slot)                                                      \
    do {                                                                   \
        if (tid == 0 && blockIdx.x == 0) {                                 \
            const long long now__ = clock64();                             \
            S.timing[slot] += now__ - tc_last__;                           \
            tc_last__ = now__;                                             \
        }                                                                  \
    } while (0)
#else
#define TC_TICK(slot) do {} while (0)
#endif

struct Ring {
    int stage = 0;
    uint32_t phase = 0;
    __device__ __forceinline__ void advance() {
        if (++stage == kStages) { stage = 0; phase ^= 1; }
    }
};

// kind::f16 instruction descriptor: fp32 accumulate, A (weights) K-major, B (activations) MN-major (bit 16),
// fmt 0 = fp16, 1 = bf16.
__device__ __forceinline__ uint32_t idesc_f16(int M, int N, int fmt) {
    return (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10) | (1u << 16) | ((uint32_t)(N >> 3) << 17) |
           ((uint32_t)(M >> 4) << 24);
}

// The activations are the MN-major B operand: for one k the rows (N) are contiguous, so the thread that owns
// feature k = TMEM lane k stores 8 consecutive rows of it with ONE 16-byte store (K-major needed 2-byte
// stores 128 bytes apart: 64 per 32 rows and part, and the MIO queue throttled the epilogue).
// Canonical MN-major SWIZZLE_128B layout: atoms of [8 k][64 rows] 16-bit = 1024 bytes (k row = 128 bytes, its
// eight 16-byte chunks XOR-swizzled with k % 8); atoms of consecutive k groups 1024 bytes apart (SBO), 64-row
// blocks kActRowBlockBytes apart (LBO).  [part][64-row block (4)][k atom (16)][1024 B].
constexpr int kActRowBlockBytes = 16 * 1024;      // one 64-row block: 128 k x 64 rows x 2 bytes
__device__ __forceinline__ uint64_t umma_desc_b_mn(uint32_t smem_addr) {
    return (uint64_t)((smem_addr & 0x3FFFF) >> 4) | ((uint64_t)(kActRowBlockBytes >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}

__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                         uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// 8 consecutive rows of one feature -> one 16-byte chunk of hi parts (and one of lo parts)
__device__ __forceinline__ void sts_u128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
template <bool kBf16>
__device__ __forceinline__ void pack_pair(float h0, float h1, uint32_t& hi, uint32_t& lo) {
    // hi = rn16(h), lo = rn16(h - hi).  The residual is one mixed-precision FMA per element
    // (fma.rn.f32.f16: -1 * hi + h, exact), so that a pair costs F2FP + 2 FHFMA + F2FP.
    float r0, r1;
    if constexpr (kBf16) {
        const __nv_bfloat162 a = __floats2bfloat162_rn(h0, h1);
        hi = *reinterpret_cast<const uint32_t*>(&a);
        const unsigned short a0 = (unsigned short)(hi & 0xffffu), a1 = (unsigned short)(hi >> 16), m1 = 0xbf80;
        asm("fma.rn.f32.bf16 %0, %1, %2, %3;" : "=f"(r0) : "h"(a0), "h"(m1), "f"(h0));
        asm("fma.rn.f32.bf16 %0, %1, %2, %3;" : "=f"(r1) : "h"(a1), "h"(m1), "f"(h1));
        const __nv_bfloat162 b = __floats2bfloat162_rn(r0, r1);
        lo = *reinterpret_cast<const uint32_t*>(&b);
    } else {
        const __half2 a = __floats2half2_rn(h0, h1);
        hi = *reinterpret_cast<const uint32_t*>(&a);
        const unsigned short a0 = (unsigned short)(hi & 0xffffu), a1 = (unsigned short)(hi >> 16), m1 = 0xbc00;
        asm("fma.rn.f32.f16 %0, %1, %2, %3;" : "=f"(r0) : "h"(a0), "h"(m1), "f"(h0));
        asm("fma.rn.f32.f16 %0, %1, %2, %3;" : "=f"(r1) : "h"(a1), "h"(m1), "f"(h1));
        const __half2 b = __floats2half2_rn(r0, r1);
        lo = *reinterpret_cast<const uint32_t*>(&b);
    }
}

__device__ __forceinline__ float2 dup2(float v) { return make_float2(v, v); }

// Sum over the 32 lanes of v[r] for every r, result for r = lane (butterfly transpose-reduce:
// 31 shuffles instead of 32 x 5).
__device__ __forceinline__ float lane_transpose_sum(float (&v)[32], int lane) {
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) {
        const bool upper = (lane & s) != 0;
#pragma unroll
        for (int k = 0; k < s; ++k) {
            const float send = upper ? v[k] : v[k + s];
            const float keep = upper ? v[k + s] : v[k];
            v[k] = keep + __shfl_xor_sync(0xffffffffu, send, s);
        }
    }
    return v[0];
}

}  // namespace

// PARTS: 2 = x3 split (hi + lo parts), 1 = one product.  kSplit: two independent 128-row groups (see above).
template <int R, int PARTS, bool kBf16, bool kSplit>
__global__ void __launch_bounds__(kThreads, 1) hybrid_tc16_kernel(const HybridArgs a) {
    constexpr int kGroups = kSplit ? 2 : 1;
    constexpr int kGroupRows = kRows / kGroups;              // rows of a group = UMMA N
    // 1024-byte aligned as declared (the 128-byte swizzle needs it; checked below).  Indexing the array
    // itself -- not a re-aligned generic pointer -- keeps every access in the shared address space (LDS/STS).
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    Smem& S = *reinterpret_cast<Smem*>(smem_raw);
    if ((smem_u32(smem_raw) & 1023u) != 0) __trap();
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int nx = a.nx;

#ifdef FLUXGNN_TC_TIMING
    if (tid < 16) S.timing[tid] = 0;
#endif
    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&S.full[s], 1);
            mbar_init(&S.empty[s], 1);
        }
        for (int g = 0; g < 2; ++g) {
            mbar_init(&S.act_ready[g], kEpiThreads);          // all 16 epilogue warps publish every group
            mbar_init(&S.acc_ready[g], 1);
        }
        mbar_fence_init();
    }
    if (warp == kMmaWarp) tmem_alloc(&S.tmem_base, kTmemCols);
    for (int i = tid; i < SmallParams::count; i += kThreads) S.small[i] = a.packed[i];
    if (a.whole_ic && a.do_update)
        for (int i = tid; i < nx; i += kThreads) S.gtab[i] = a.gtab[i];
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = S.tmem_base;

    // a.num_tiles counts LOGICAL tiles (one per group); a CTA tile holds kGroups of them
    const int cta_tiles = (a.num_tiles + kGroups - 1) / kGroups;
    const int my_tiles = (cta_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const int layers = a.L + 1;

    if (warp == kProducerWarp) {
        // ---------------- weight producer ------------------------------------------------
        if (lane == 0) {
            const unsigned char* stream = reinterpret_cast<const unsigned char*>(a.packed + SmallParams::count);
            Ring r;
            for (long long rep = 0; rep < (long long)my_tiles * a.steps; ++rep) {
                for (int lg = 0; lg < layers * kGroups; ++lg) {      // every group consumes the whole layer
                    const int layer = lg / kGroups;
                    for (int u = 0; u < kTc16UnitsPerLayer; ++u) {
                        if (PARTS == 1 && (u & 1)) continue;          // one product: no lo units
#ifdef FLUXGNN_TC_DEBUG_NOSTREAM                              // timing experiment: UMMAs on stale weights, no stream
                        continue;
#endif
                        mbar_wait(&S.empty[r.stage], r.phase ^ 1);
                        mbar_arrive_expect_tx(&S.full[r.stage], kTc16UnitBytes);
                        bulk_g2s(S.Ws[r.stage], stream + ((size_t)layer * kTc16UnitsPerLayer + u) * kTc16UnitBytes,
                                 kTc16UnitBytes, &S.full[r.stage]);
                        r.advance();
                    }
                }
            }
        }
    } else if (warp == kMmaWarp) {
        // ---------------- UMMA issuer ------------------------------------------------------
        // The WHOLE warp walks the loop in uniform control flow and one elected lane issues.  With a
        // `lane == 0` branch around everything the compiler cannot prove the descriptors warp-uniform
        // (ring stage, group) and wraps every UTCHMMA in R2UR moves and an ELECT / BRA.U.ANY
        // serialisation loop: ~15 dependent instructions of a lone warp per UMMA, ~136 clk against the
        // 64 clk the tensor pipe needs for an N = 128 instruction.
        const uint32_t idesc = idesc_f16(128, kGroupRows, kBf16 ? 1 : 0);
        const bool leader = elect_one_lane();
        const uint64_t ws_desc0 = umma_desc_sw128(smem_u32(S.Ws[0]));
        const uint64_t hi_desc0 = umma_desc_b_mn(smem_u32(S.act[0]));
        const uint64_t lo_desc0 = umma_desc_b_mn(smem_u32(S.act[1]));
        constexpr uint64_t kStageStep = kTc16UnitBytes >> 4;          // descriptor address field counts 16-byte units
        constexpr uint64_t kKbStep = (uint64_t)(8 * 1024) >> 4;        // 64 k = 8 atoms
        constexpr uint64_t kKsStep = (uint64_t)(2 * 1024) >> 4;        // 16 k = 2 atoms per instruction
        constexpr uint64_t kGroupStep = (uint64_t)((kGroupRows / 64) * kActRowBlockBytes) >> 4;
        Ring r;
        uint32_t act_phase = 0;
        for (long long rep = 0; rep < (long long)my_tiles * a.steps; ++rep) {
            for (int lg = 0; lg < layers * kGroups; ++lg) {
                const int grp = lg % kGroups;                         // kSplit: the halves alternate
                // the group's rows of the B operand and its columns of the accumulators
                const uint64_t bhi = hi_desc0 + grp * kGroupStep, blo = lo_desc0 + grp * kGroupStep;
                mbar_wait(&S.act_ready[grp], act_phase);
                if (grp == kGroups - 1) act_phase ^= 1;
                tc_fence_after();
#pragma unroll
                for (int kb = 0; kb < 2; ++kb) {
#pragma unroll
                    for (int blk = 0; blk < 2; ++blk) {
                        const uint32_t d = tmem + (blk == 0 ? kColZ : kColY) + grp * kGroupRows;
                        // hi weights x (hi [+ lo] activations)
#ifndef FLUXGNN_TC_DEBUG_NOSTREAM
                        mbar_wait(&S.full[r.stage], r.phase);
#endif
                        tc_fence_after();
                        uint64_t wd = ws_desc0 + r.stage * kStageStep;
                        if (leader) {
#pragma unroll
                            for (int ks = 0; ks < 4; ++ks) {          // 32 bytes of K per instruction = 2 address units
                                umma_f16(d, wd + 2 * ks, bhi + kb * kKbStep + ks * kKsStep, idesc, (kb | ks) != 0);
                                if (PARTS == 2) umma_f16(d, wd + 2 * ks, blo + kb * kKbStep + ks * kKsStep, idesc, 1);
                            }
                            umma_commit(&S.empty[r.stage]);
                        }
                        r.advance();
                        if (PARTS == 2) {
                            // lo weights x hi activations
#ifndef FLUXGNN_TC_DEBUG_NOSTREAM
                            mbar_wait(&S.full[r.stage], r.phase);
#endif
                            tc_fence_after();
                            wd = ws_desc0 + r.stage * kStageStep;
                            if (leader) {
#pragma unroll
                                for (int ks = 0; ks < 4; ++ks)
                                    umma_f16(d, wd + 2 * ks, bhi + kb * kKbStep + ks * kKsStep, idesc, 1);
                                umma_commit(&S.empty[r.stage]);
                            }
                            r.advance();
                        }
                    }
                }
                if (leader) umma_commit(&S.acc_ready[grp]);
                __syncwarp();
            }
        }
    } else {
        // ---------------- epilogue / compute warps ---------------------------------------
        // All 16 warps serve whichever group's accumulators are ready: per layer they finish group 0 while
        // the tensor pipe works on group 1, then group 1 while it works on group 0's next layer.  The
        // finite-volume tail of a group and the input layer of its NEXT step (or next tile) run right after
        // its edge readout, so that the issuer always has the other group's products to go on with.
        const int q = warp & 3;                               // TMEM lane quadrant = feature quadrant
        const int cw = warp >> 2;                             // 0..3: the warp's 32-row chunk(s) inside a group
        constexpr int kChunksPerWarp = kGroupRows / 128;      // 1 (two groups) or 2 (one 256-row group)
        constexpr int kRowThreads = kEpiThreads / kGroupRows; // threads per row in the field solve: 4 or 2
        const TileRows T{S.sN, S.sU, S.sE, S.sX, S.sF, S.sRho, S.gtab, S.rowIC, S.rowCell, S.prevRow, S.nextRow};
        const int n = 32 * q + lane;                          // this thread's feature = TMEM lane
        const uint32_t tlane = tmem + ((uint32_t)(32 * q) << 16);
        const float inv_deg = kUnscale / (float)(2 * R);
        const int seg = a.whole_ic ? nx : kGroupRows;         // periodic segment inside the group (multiple of 32)
        uint32_t acc_phase = 0;
        const bool fuse_faces = (a.flux_edges == nullptr);    // no per-edge output: one feature reduction per face

        // rows 8c .. 8c+7 (c = chunk inside a 64-row block) of feature n: one 16-byte chunk at
        //   (row block) * kActRowBlockBytes + (n / 8) * 1024 + (n % 8) * 128 + ((c ^ (n % 8)) << 4)
        const uint32_t act_feat = smem_u32(S.act[0]) + (uint32_t)((n >> 3) * 1024 + (n & 7) * 128);
        const uint32_t act_lo_off = (uint32_t)kActPartBytes;              // S.act[1] - S.act[0]
        // store_rows8: h[0..7] = rows i0 + 8c .. of this feature (i0 a multiple of 32, c = 0..3)
        auto store_rows8 = [&](int i0, int c, const float (&h)[8]) {
            const uint32_t chunk = (uint32_t)(((i0 >> 5) & 1) * 4 + c);
            const uint32_t addr = act_feat + (uint32_t)(i0 >> 6) * kActRowBlockBytes + ((chunk ^ (uint32_t)(n & 7)) << 4);
            uint32_t hi[4], lo[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) pack_pair<kBf16>(h[2 * t], h[2 * t + 1], hi[t], lo[t]);
            sts_u128(addr, hi[0], hi[1], hi[2], hi[3]);
            if (PARTS == 2) sts_u128(addr + act_lo_off, lo[0], lo[1], lo[2], lo[3]);
        };
        auto publish_activations = [&](int g) {               // generic-proxy stores -> visible to the UMMAs of group g
            tc_fence_before();
            fence_proxy_async();
            mbar_arrive(&S.act_ready[g]);
        };

        // ---- row bookkeeping + state load of group g's logical tile (tile_common.cuh) ----
        auto load_rows = [&](int g, int cta_tile) {
            const int tile = cta_tile * kGroups + g;
            if (tid < kGroupRows)
                tile_load_row(a, T, tile, tile < a.num_tiles, g * kGroupRows + tid, tid, g * kGroupRows, kGroupRows);
        };

        // ---- input MLP (src/flux_gnn.py:49): feature n, the warp's chunk(s) of group g ----
        auto input_layer = [&](int g) {
            const float2 w0 = dup2(S.small[SmallParams::w_in + 0 * kH + n]);
            const float2 w1 = dup2(S.small[SmallParams::w_in + 1 * kH + n]);
            const float2 w2 = dup2(S.small[SmallParams::w_in + 2 * kH + n]);
            const float2 w3 = dup2(S.small[SmallParams::w_in + 3 * kH + n]);
            const float2 b = dup2(S.small[SmallParams::b_in + n]);
#pragma unroll 1
            for (int half = 0; half < kChunksPerWarp; ++half) {
                const int i0 = g * kGroupRows + 32 * (cw + 4 * half);
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    float h[8];
#pragma unroll
                    for (int j = 0; j < 8; j += 4) {
                        const float4 vn = *reinterpret_cast<const float4*>(&S.sN[i0 + 8 * c + j]);
                        const float4 vu = *reinterpret_cast<const float4*>(&S.sU[i0 + 8 * c + j]);
                        const float4 ve = *reinterpret_cast<const float4*>(&S.sE[i0 + 8 * c + j]);
                        const float4 vx = *reinterpret_cast<const float4*>(&S.sX[i0 + 8 * c + j]);
                        // two cells per packed FMA, same order of operations per cell as the scalar chain
                        const float2 lo2 = __ffma2_rn(w3, make_float2(vx.x, vx.y), __ffma2_rn(w2, make_float2(ve.x, ve.y),
                                           __ffma2_rn(w1, make_float2(vu.x, vu.y), __ffma2_rn(w0, make_float2(vn.x, vn.y), b))));
                        const float2 hi2 = __ffma2_rn(w3, make_float2(vx.z, vx.w), __ffma2_rn(w2, make_float2(ve.z, ve.w),
                                           __ffma2_rn(w1, make_float2(vu.z, vu.w), __ffma2_rn(w0, make_float2(vn.z, vn.w), b))));
                        h[j + 0] = fmaxf(lo2.x, 0.f);
                        h[j + 1] = fmaxf(lo2.y, 0.f);
                        h[j + 2] = fmaxf(hi2.x, 0.f);
                        h[j + 3] = fmaxf(hi2.y, 0.f);
                    }
                    store_rows8(i0, c, h);
                }
            }
            publish_activations(g);
        };

        // ---- one layer's epilogue for group g: message passing (layer < L) or edge readout (layer == L) ----
        auto layer_epilogue = [&](int g, int layer) {
            const bool is_edge = (layer == a.L);
            const float bias = S.small[(is_edge ? SmallParams::b_e1 : SmallParams::b_upd + layer * kH) + n];
            const float w_out = S.small[SmallParams::w_e2 + n];
#pragma unroll 1
            for (int half = 0; half < kChunksPerWarp; ++half) {
                const int i0 = g * kGroupRows + 32 * (cw + 4 * half);
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
                if (!is_edge && R == 3) {
                    // h'_i = relu(Y_i + b + mean_{0<|k|<=3} Z_{i+k})   (src/flux_gnn.py:55-60), two rows per packed
                    // instruction.  Window w[0..39] = rows i0-4 .. i0+35 as aligned pairs P[m] = (w[2m], w[2m+1]);
                    // with hs[m] = w[2m] + w[2m+1] and G[k] = P[k-1] + P[k+1] (packed) the sums of the rows of pair m are
                    //   S(2m) = hs[m-1] + hs[m+1] + G[m-1].y,   S(2m+1) = hs[m-1] + hs[m+1] + G[m+1].x
                    float2 P[20], G[20];
                    float hs[20];
                    P[0] = make_float2(zl[0], zl[1]);
                    P[1] = make_float2(zl[2], zl[3]);
                    P[18] = make_float2(zr[0], zr[1]);
                    P[19] = make_float2(zr[2], zr[3]);
#pragma unroll
                    for (int m = 0; m < 16; ++m) P[2 + m] = make_float2(zc[2 * m], zc[2 * m + 1]);
#pragma unroll
                    for (int m = 1; m < 19; ++m) {
                        hs[m] = P[m].x + P[m].y;
                        G[m] = __fadd2_rn(P[m - 1], P[m + 1]);
                    }
                    const float2 us2 = dup2(kUnscale), bias2 = dup2(bias), inv2 = dup2(inv_deg);
#pragma unroll
                    for (int ch = 0; ch < 4; ++ch) {
                        float h[8];
#pragma unroll
                        for (int t = 0; t < 4; ++t) {
                            const int m = 4 * ch + t, mc = m + 2;
                            const float tm = hs[mc - 1] + hs[mc + 1];
                            const float2 sum = make_float2(tm + G[mc - 1].y, tm + G[mc + 1].x);
                            const float2 r = __ffma2_rn(sum, inv2, __ffma2_rn(make_float2(y[2 * m], y[2 * m + 1]), us2, bias2));
                            h[2 * t] = fmaxf(r.x, 0.f);
                            h[2 * t + 1] = fmaxf(r.y, 0.f);
                        }
                        store_rows8(i0, ch, h);
                    }
                } else if (!is_edge) {
                    // h'_i = relu(Y_i + b + mean_{0<|k|<=R} Z_{i+k})   (src/flux_gnn.py:55-60)
                    // the window sum shares the pair sums pz[t] = Z[t] + Z[t+1] between neighbouring rows
                    float pz[39];
                    if constexpr (R >= 2) {
#pragma unroll
                        for (int t = 0; t < 39; ++t) pz[t] = zw[t] + zw[t + 1];
                    }
                    auto window_sum = [&](int c) {                      // c = window centre in zw
                        if constexpr (R == 1) return zw[c - 1] + zw[c + 1];
                        else if constexpr (R == 2) return pz[c - 2] + pz[c + 1];
                        else if constexpr (R == 3) return (pz[c - 3] + pz[c + 2]) + (zw[c - 1] + zw[c + 1]);
                        else return (pz[c - 4] + pz[c + 3]) + (pz[c - 2] + pz[c + 1]);
                    };
                    const float2 us2 = dup2(kUnscale), bias2 = dup2(bias), inv2 = dup2(inv_deg);
#pragma unroll
                    for (int ch = 0; ch < 4; ++ch) {
                        float h[8];
#pragma unroll
                        for (int t = 0; t < 4; ++t) {                   // bias, unscale and mean: two rows per FFMA2
                            const int j = 8 * ch + 2 * t;
                            const float2 sum = make_float2(window_sum(4 + j), window_sum(5 + j));
                            const float2 r = __ffma2_rn(sum, inv2, __ffma2_rn(make_float2(y[j], y[j + 1]), us2, bias2));
                            h[2 * t] = fmaxf(r.x, 0.f);
                            h[2 * t + 1] = fmaxf(r.y, 0.f);
                        }
                        store_rows8(i0, ch, h);
                    }
                } else {
                    // edge readout (src/flux_gnn.py:63-66): this feature's term of the two dot products of
                    // row i -- fwd (row i, col i+1): w2[n] relu(P_i + b1 + Q_{i+1});  bwd (row i, col i-1) --
                    // summed over the warp's 32 features in registers, over the 4 quadrants in shared memory
                    float f[32];
                    if (fuse_faces) {
                        // Rollouts only need the face flux 0.5 (fwd_i + bwd_{i+1}) (src/hybrid_solver.py:45-48): add
                        // the two terms of face i + 1/2 per feature and reduce once.  P + b1 of row i0 + 32 (the
                        // periodic right neighbour of the chunk) comes from one more accumulator column.
                        float yr[4];
                        tmem_ld4(tlane + kColY + cr, yr);
                        tc_wait_ld();
                        const float2 us2 = dup2(kUnscale), bias2 = dup2(bias), w2 = dup2(w_out);
                        float yb[34];
#pragma unroll
                        for (int m = 0; m < 16; ++m) {
                            const float2 t = __ffma2_rn(make_float2(y[2 * m], y[2 * m + 1]), us2, bias2);
                            yb[2 * m] = t.x;
                            yb[2 * m + 1] = t.y;
                        }
                        yb[32] = fmaf(yr[0], kUnscale, bias);
#pragma unroll
                        for (int m = 0; m < 16; ++m) {
                            const int j = 2 * m;
                            const float2 fw = make_float2(fmaxf(fmaf(zw[4 + j + 1], kUnscale, yb[j]), 0.f),
                                                          fmaxf(fmaf(zw[4 + j + 2], kUnscale, yb[j + 1]), 0.f));
                            const float2 bw = make_float2(fmaxf(fmaf(zw[4 + j], kUnscale, yb[j + 1]), 0.f),
                                                          fmaxf(fmaf(zw[4 + j + 1], kUnscale, yb[j + 2]), 0.f));
                            const float2 t = __fmul2_rn(w2, __fadd2_rn(fw, bw));
                            f[j] = t.x;
                            f[j + 1] = t.y;
                        }
                        S.edgeP[q][0][i0 + lane] = lane_transpose_sum(f, lane);
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            y[j] = fmaf(y[j], kUnscale, bias);
                            f[j] = w_out * fmaxf(fmaf(zw[4 + j + 1], kUnscale, y[j]), 0.f);
                        }
                        S.edgeP[q][0][i0 + lane] = lane_transpose_sum(f, lane);
#pragma unroll
                        for (int j = 0; j < 32; ++j) f[j] = w_out * fmaxf(fmaf(zw[4 + j - 1], kUnscale, y[j]), 0.f);
                        S.edgeP[q][1][i0 + lane] = lane_transpose_sum(f, lane);
                    }
                }
            }
            if (!is_edge) publish_activations(g);
        };

        // ---- after the edge readout of group g: face flux, finite-volume update, field solve, write-out ----
        auto finish_step = [&](int g, int step) {
            const int myrow = (tid < kGroupRows) ? g * kGroupRows + tid : -1;
            // u of the left neighbour is read BEFORE the barrier: below every row overwrites its own state as
            // soon as it has its update, without a barrier between the reads and the writes
            int prow = 0;
            float u_prev = 0.f;
            if (a.do_update && myrow >= 0) {
                prow = S.prevRow[myrow];
                u_prev = S.sU[prow];
            }
            named_sync(1, kEpiThreads);                        // edgeP of this group complete
            if (myrow >= 0) {
                // face flux (src/hybrid_solver.py:45-48): edges (row j, col j+1) and (row j+1, col j)
                const int j = myrow;
                const float b2 = S.small[SmallParams::b_e2];
                auto edge_sum = [&](int dir, int row) {
                    return (S.edgeP[0][dir][row] + S.edgeP[1][dir][row]) + (S.edgeP[2][dir][row] + S.edgeP[3][dir][row]);
                };
                const int ic = S.rowIC[j], cell = S.rowCell[j];
                float face, face_prev = 0.f;
                if (fuse_faces) {                              // edgeP[.][0] = fwd_j + bwd_{j+1} without the biases
                    face = fmaf(0.5f, edge_sum(0, j), b2);
                    if (a.do_update) face_prev = fmaf(0.5f, edge_sum(0, prow), b2);   // same bits as row prow's own face
                } else {
                    const float fwd = edge_sum(0, j) + b2, bwd = edge_sum(1, S.nextRow[j]) + b2;
                    face = 0.5f * (fwd + bwd);
                    if (a.do_update) face_prev = 0.5f * ((edge_sum(0, prow) + b2) + (edge_sum(1, j) + b2));
                    if (ic >= 0) {
                        float* fe = a.flux_edges + (size_t)ic * 2 * nx + cell;
                        fe[0] = fwd;
                        fe[nx] = bwd;
                    }
                }
                if (a.face_flux != nullptr && ic >= 0) a.face_flux[(size_t)ic * nx + cell] = face;
                if (a.do_update) {
                    // finite-volume update (src/hybrid_solver.py:51-58)
                    float n_new, u_new;
                    tile_fv_update_values(a, S.sN[j], S.sU[j], u_prev, S.sE[j], face, face_prev, n_new, u_new);
                    if (a.whole_ic) tile_keep_row(T, j, n_new, u_new);
                    else tile_store_window_row(a, T, j, n_new, u_new);
                }
            }
            if (!a.do_update || !a.whole_ic) return;
            named_sync(1, kEpiThreads);                        // rho of every row
            // field solve: E = g (*) rho, kRowThreads threads per row (src/baseline_solver.py:59-68)
            {
                const int row = g * kGroupRows + tid / kRowThreads, part = tid % kRowThreads;
                double e = tile_field_partial(T, row, part, kRowThreads, nx);
#pragma unroll
                for (int m = 1; m < kRowThreads; m <<= 1) e += __shfl_xor_sync(0xffffffffu, e, m);
                if (part == 0) S.sE[row] = (float)e;
            }
            named_sync(1, kEpiThreads);                        // the new state of every row is in shared memory
            if (myrow >= 0) tile_write_out_row(a, T, myrow, step);
        };

#ifdef FLUXGNN_TC_TIMING
        long long tc_last__ = clock64();
#endif
        // prologue: first tile of both groups
        for (int g = 0; g < kGroups; ++g) load_rows(g, blockIdx.x);
        named_sync(1, kEpiThreads);
        for (int g = 0; g < kGroups; ++g) input_layer(g);
        TC_TICK(2);
        for (int cta_tile = blockIdx.x; cta_tile < cta_tiles; cta_tile += gridDim.x) {
            const int next_tile = cta_tile + (int)gridDim.x;
            for (int step = 0; step < a.steps; ++step) {
                for (int layer = 0; layer < layers; ++layer) {
                    for (int g = 0; g < kGroups; ++g) {
                        mbar_wait(&S.acc_ready[g], acc_phase);
                        tc_fence_after();
                        TC_TICK(3);
                        layer_epilogue(g, layer);
                        TC_TICK(layer < a.L ? 4 : 5);
                        if (layer < a.L) continue;
                        finish_step(g, step);
                        TC_TICK(8);
                        // group g moves on to its next step / next tile while the other group's products run
                        if (step + 1 < a.steps) {                // whole-IC tiles: finish_step ended on a barrier
                            input_layer(g);
                        } else if (next_tile < cta_tiles) {
                            named_sync(1, kEpiThreads);        // nobody reads this group's rows any more
                            load_rows(g, next_tile);
                            named_sync(1, kEpiThreads);
                            input_layer(g);
                        }
                        TC_TICK(2);
                    }
                    acc_phase ^= 1;
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
#ifdef FLUXGNN_TC_TIMING
    if (blockIdx.x == 0 && tid < 16) g_tc16_timing[tid] += S.timing[tid];
#endif
    if (warp == kMmaWarp) tmem_dealloc(tmem, kTmemCols);
}

#ifdef FLUXGNN_TC_TIMING
extern "C" int fluxgnn_debug_tc_timing(long long* out16, int reset) {
    if (out16 && cudaMemcpyFromSymbol(out16, g_tc16_timing, sizeof(g_tc16_timing)) != cudaSuccess) return -1;
    if (reset) {
        long long zero[16] = {0};
        if (cudaMemcpyToSymbol(g_tc16_timing, zero, sizeof(zero)) != cudaSuccess) return -1;
    }
    return 0;
}
#endif

template <int R, int PARTS, bool kBf16, bool kSplit>
static cudaError_t launch_one(const HybridArgs& a, int grid, cudaStream_t stream) {
    const int smem = (int)sizeof(Smem);
    cudaError_t e = cudaFuncSetAttribute(hybrid_tc16_kernel<R, PARTS, kBf16, kSplit>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    hybrid_tc16_kernel<R, PARTS, kBf16, kSplit><<<grid, kThreads, smem, stream>>>(a);
    return cudaGetLastError();
}

template <int R, bool kSplit>
static cudaError_t launch_mode(const HybridArgs& a, int grid, cudaStream_t stream) {
    if (a.tc_format == 1) return launch_one<R, 1, true, kSplit>(a, grid, stream);            // bf16
    if (a.tc_parts == 2) return launch_one<R, 2, false, kSplit>(a, grid, stream);             // fp16x3
    return launch_one<R, 1, false, kSplit>(a, grid, stream);                                  // fp16
}

// a.tc_group_rows (api.cu, plan_tiles): 128 = two independent groups per CTA tile, 256 = one group
template <int R>
static cudaError_t launch_radius(const HybridArgs& a, int grid, cudaStream_t stream) {
    return a.tc_group_rows == 128 ? launch_mode<R, true>(a, grid, stream) : launch_mode<R, false>(a, grid, stream);
}

cudaError_t launch_hybrid_tc16_tiles(const HybridArgs& a, int radius, int grid, cudaStream_t stream) {
    switch (radius) {
        case 1: return launch_radius<1>(a, grid, stream);
        case 2: return launch_radius<2>(a, grid, stream);
        case 3: return launch_radius<3>(a, grid, stream);
        case 4: return launch_radius<4>(a, grid, stream);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace fluxgnn
