// Classical step with the field solve as a parallel prefix sum ("scan solve"), sm_100a.
//
//   src/baseline_solver.py:80-94   upwind fluxes, viscous Laplacian, forward Euler   (fv_cell, bit-exact)
//   src/baseline_solver.py:59-68   E = Re ifft(i fft(n - 1) / k)                       (reconstructed, see below)
//
// The spectral operator is the zero-mean periodic antiderivative  dE/dx = -(rho - mean rho),  rho = n - 1.  On a
// long grid it can be evaluated without a transform: with the inclusive prefix C_j = sum_{i<=j} rho_i,
//   E_j = -dx (C_j - rho_j/2 - (j + 1/2) mean(rho) - mu) + (dx/24)(rho_{j+1} - rho_{j-1}),
//   mu  = S/2 - M1/N - S/(2N),   S = sum rho_i,   M1 = sum i rho_i        (makes mean(E) = 0),
// is the trapezoid rule between cell centres plus the first Euler-Maclaurin correction.  Per Fourier mode its
// multiplier differs from i/k by the factor  1 - r - kappa sin(kappa)/12,  r = (kappa/2) cot(kappa/2),  kappa = k dx,
// which is bounded by (1/16) (2 sin(kappa/2))^4 on (0, pi]: the deviation from the spectral field is therefore at most
//   || E_scan - E_spectral ||_inf  <=  rms(Delta^4 rho) * L / (32 sqrt 3)                (Cauchy-Schwarz over the modes)
// and the kernel EVALUATES that bound for every field it reconstructs (certificate).  A field whose bound exceeds
// tol * max|E| is reported through `flag` (first uncertified step, per IC); the host mirror then repeats those ICs with
// the FFT solve.  Smooth states on long grids pass with a margin set by fp32 rounding noise (bound ~3e-8 absolute);
// white noise or short grids do not, and never silently use this path.
//
// Because E is a function of n alone, the rollout keeps only (n, u) in HBM between steps: step s+1 reconstructs the field
// of its input from n and the per-segment sums the previous launch left behind, so a step moves 16 bytes per cell
// (read n, u; write n', u') instead of 24 + the transform passes.  Every grid is cut into contiguous segments, one
// persistent CTA each; a 128-thread CTA streams its segment in 1024-cell chunks through a 4-stage shared-memory ring filled by 1-D
// bulk copies (TMA, SASS UBLKCP) and carries the running prefix in a register.  No inter-CTA dependency inside a
// launch: segment prefixes come from the records of the previous launch.
#include <climits>
#include <cstring>

#include "common.cuh"
#include "field_kernels.cuh"

namespace fluxgnn {

namespace {

#ifndef FLUXGNN_SCAN_THREADS
#define FLUXGNN_SCAN_THREADS 128
#endif
// Measured at 2^24 cells x batch 8 (us per 2^24 cells): 512 threads x 2 CTAs per SM 59; 256 x 4 (64 registers: spills) 63;
// 256 x 3 52.6; 384 x 2 58; 64 x 12 52; 128 x 6 49.9; 128 x 5, four stages 47.3 (5.67 TB/s on the 16 B/cell that move):
// small barrier domains, 102 registers per thread, no spills.
constexpr int kScanThreads = FLUXGNN_SCAN_THREADS;
#ifndef FLUXGNN_SCAN_CTAS
#define FLUXGNN_SCAN_CTAS 5
#endif
constexpr int kScanCtasPerSm = FLUXGNN_SCAN_CTAS;
constexpr int kScanPer = 8;                                  // consecutive cells per thread
constexpr int kScanChunk = kScanThreads * kScanPer;          // 1024 cells
#ifndef FLUXGNN_SCAN_STAGES
#define FLUXGNN_SCAN_STAGES 4
#endif
constexpr int kScanStages = FLUXGNN_SCAN_STAGES;
constexpr int kScanHalo = 4;                                 // staged halo cells per side (16 bytes)
constexpr int kScanRow = kScanChunk + 2 * kScanHalo;         // floats per staged array

struct ScanRec {             // what one segment's CTA leaves for the next launch
    double S;                // sum of rho' over the segment (new state)
    double M1;               // sum of j * rho'_j, j = cell index inside the IC
    double D4;               // sum of (Delta^4 rho)^2 of the INPUT state (certificate of the field reconstructed here)
    float maxE;              // max |E| of the field reconstructed here
    int pad;
};
static_assert(sizeof(ScanRec) == 32, "record layout");

struct ScanArgs {
    const float* in;         // [B][3][nx]
    float* out;              // [B][3][nx]: n', u' planes written (modes 0, 1)
    float* e_out;            // [B][3][nx]-strided base of a state whose E plane receives the input state's field (modes 1, 2)
    float* flux_out;         // [B][nx] continuity flux n*u of the input state, or null
    const ScanRec* rec_in;   // [B][segs] from the previous launch (modes 1, 2)
    ScanRec* rec_out;        // [B][segs]
    int* flag;               // [B]: first uncertified step of every IC (INT_MAX = all certified)
    int step;                // index of the step whose input field the records in rec_in certify
    int B, nx, segs, seg_chunks;
    float c, dt, nu, dx2;
    double dx, length, tol;
};

struct ScanSmem {
    float nbuf[kScanStages][kScanRow];
    float ubuf[kScanStages][kScanRow];
    float wsum[2][kScanThreads / 32];
    double red[4][kScanThreads / 32];
    float redf[kScanThreads / 32];
    double bc[4];            // broadcast: P_seg, rbar, mu, (unused)
    uint64_t full[kScanStages];
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// MODE 0: first step of a rollout -- E is read from the input state, no reconstruction, no certificate.
// MODE 1: later steps            -- E reconstructed from n and rec_in; optionally stored to e_out (recorded states).
// MODE 2: materialise            -- E of the input state reconstructed and stored to e_out; no update.
// kPacked: the finite-volume update in packed fp32x2 arithmetic (fv_pair; needs fv_reciprocal(dx2) != 0), bit-identical
// to the scalar form.
template <int MODE, bool kPacked>
__global__ void __launch_bounds__(kScanThreads, kScanCtasPerSm) baseline_scan_kernel(const ScanArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    ScanSmem& S = *reinterpret_cast<ScanSmem*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ic = (int)blockIdx.x / a.segs, seg = (int)blockIdx.x - ic * a.segs;
    const int nx = a.nx;
    const int seg_begin = seg * a.seg_chunks * kScanChunk;              // cell indices fit an int (nx is one)
    int seg_end = (int)min((long long)seg_begin + (long long)a.seg_chunks * kScanChunk, (long long)nx);
    const int nchunks = (int)((seg_end - seg_begin + kScanChunk - 1) / kScanChunk);
    const float* pn = a.in + (size_t)ic * 3 * nx;
    const float* pu = pn + nx;
    const float* pe = pu + nx;

    if (tid == 0) {
        for (int s = 0; s < kScanStages; ++s) mbar_init(&S.full[s], 1);
        mbar_fence_init();
    }
    __syncthreads();

    // producer (thread 0): stage chunk k = cells [c0, c0 + len) plus 4 halo cells per side, periodic wrap applied here
    auto issue = [&](int k) {
        const int st = k % kScanStages;
        const int c0 = seg_begin + k * kScanChunk;
        const int len = (seg_end - c0 < kScanChunk) ? (seg_end - c0) : kScanChunk;
        const int left = (c0 == 0) ? nx - kScanHalo : c0 - kScanHalo;
        const int right = (c0 + len == nx) ? 0 : c0 + len;
        const uint32_t bytes = (uint32_t)(len + 2 * kScanHalo) * 4u;
        mbar_arrive_expect_tx(&S.full[st], (MODE == 2) ? bytes : 2 * bytes);
        bulk_g2s(&S.nbuf[st][0], pn + left, kScanHalo * 4, &S.full[st]);
        bulk_g2s(&S.nbuf[st][kScanHalo], pn + c0, (uint32_t)len * 4u, &S.full[st]);
        bulk_g2s(&S.nbuf[st][kScanHalo + len], pn + right, kScanHalo * 4, &S.full[st]);
        if (MODE != 2) {
            bulk_g2s(&S.ubuf[st][0], pu + left, kScanHalo * 4, &S.full[st]);
            bulk_g2s(&S.ubuf[st][kScanHalo], pu + c0, (uint32_t)len * 4u, &S.full[st]);
            bulk_g2s(&S.ubuf[st][kScanHalo + len], pu + right, kScanHalo * 4, &S.full[st]);
        }
    };
    if (tid == 0)
        for (int k = 0; k < kScanStages && k < nchunks; ++k) issue(k);

    // ---- prologue: sums of the previous launch's records -> segment prefix, mean, zero-mean constant, certificate ----
    double P = 0.0, rbar = 0.0, mu = 0.0;
    if (MODE != 0) {
        double s_all = 0.0, s_before = 0.0, m_all = 0.0, d_all = 0.0;
        float e_max = 0.f;
        const ScanRec* rec = a.rec_in + (size_t)ic * a.segs;
        for (int r = tid; r < a.segs; r += kScanThreads) {
            const ScanRec v = rec[r];
            s_all += v.S;
            if (r < seg) s_before += v.S;
            m_all += v.M1;
            d_all += v.D4;
            e_max = fmaxf(e_max, v.maxE);
        }
        s_all = warp_sum(s_all); s_before = warp_sum(s_before); m_all = warp_sum(m_all); d_all = warp_sum(d_all);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) e_max = fmaxf(e_max, __shfl_xor_sync(0xffffffffu, e_max, o));
        if (lane == 0) {
            S.red[0][warp] = s_all; S.red[1][warp] = s_before; S.red[2][warp] = m_all; S.red[3][warp] = d_all;
            S.redf[warp] = e_max;
        }
        __syncthreads();
        if (tid == 0) {
            double sa = 0.0, sb = 0.0, ma = 0.0, da = 0.0;
            float em = 0.f;
            for (int w = 0; w < kScanThreads / 32; ++w) {
                sa += S.red[0][w]; sb += S.red[1][w]; ma += S.red[2][w]; da += S.red[3][w];
                em = fmaxf(em, S.redf[w]);
            }
            const double N = (double)nx;
            S.bc[0] = sb;
            S.bc[1] = sa / N;
            S.bc[2] = 0.5 * sa - ma / N - sa / (2.0 * N);
            // certificate of the field the PREVIOUS launch reconstructed (its records carry D4 and max|E|)
            if (seg == 0 && a.step > 0) {
                const double bound = sqrt(da / N) * a.length * (1.0 / (32.0 * 1.7320508075688772));
                if (!(bound <= a.tol * (double)em)) atomicMin(a.flag + ic, a.step - 1);
            }
        }
        __syncthreads();
        P = S.bc[0]; rbar = S.bc[1]; mu = S.bc[2];
    }

    const float rdx2 = fv_reciprocal(a.dx2);
    const float dxf = (float)a.dx, dx24 = (float)(a.dx / 24.0), rbarf = (float)rbar;
    double run = P;                      // prefix of rho over the cells of this IC before the current chunk
    double accS = 0.0, accM = 0.0;
    float accD = 0.f, accE = 0.f;

    int st = 0;                          // ring stage and phase parity of chunk k (k % kScanStages, (k / kScanStages) & 1)
    uint32_t parity = 0;
    int j0 = seg_begin + tid * kScanPer;                        // first cell of this thread inside the IC
    for (int k = 0; k < nchunks; ++k, j0 += kScanChunk, st = (st + 1 == kScanStages) ? 0 : st + 1, parity ^= (st == 0)) {
        const bool active = j0 < seg_end;
        mbar_wait(&S.full[st], parity);

        float nv[kScanPer + 4], uv[kScanPer + 2];      // cells -2..9 of n, -1..8 of u relative to the thread's first cell
        if (active) {
            const float* sn = &S.nbuf[st][kScanHalo + tid * kScanPer];
            const float4 n0 = *reinterpret_cast<const float4*>(sn), n1 = *reinterpret_cast<const float4*>(sn + 4);
            const float2 nl = *reinterpret_cast<const float2*>(sn - 2), nr = *reinterpret_cast<const float2*>(sn + 8);
            nv[0] = nl.x; nv[1] = nl.y;
            nv[2] = n0.x; nv[3] = n0.y; nv[4] = n0.z; nv[5] = n0.w;
            nv[6] = n1.x; nv[7] = n1.y; nv[8] = n1.z; nv[9] = n1.w;
            nv[10] = nr.x; nv[11] = nr.y;
            if (MODE != 2) {
                const float* su = &S.ubuf[st][kScanHalo + tid * kScanPer];
                const float4 u0 = *reinterpret_cast<const float4*>(su), u1 = *reinterpret_cast<const float4*>(su + 4);
                uv[0] = su[-1];
                uv[1] = u0.x; uv[2] = u0.y; uv[3] = u0.z; uv[4] = u0.w;
                uv[5] = u1.x; uv[6] = u1.y; uv[7] = u1.z; uv[8] = u1.w;
                uv[9] = su[8];
            }
        } else {
#pragma unroll
            for (int i = 0; i < kScanPer + 4; ++i) nv[i] = 1.0f;
#pragma unroll
            for (int i = 0; i < kScanPer + 2; ++i) uv[i] = 0.f;
        }

        float rho[kScanPer + 4];
#pragma unroll
        for (int i = 0; i < kScanPer + 4; ++i) rho[i] = __fsub_rn(nv[i], 1.0f);      // :60, float32 like numpy

        // ---- block-wide exclusive prefix of rho: fp32 inside a warp (at most 256 cells), fp64 across warps and chunks ----
        float loc[kScanPer];
        float excl32 = 0.f;
        if (MODE != 0) {
            float t = 0.f;
#pragma unroll
            for (int i = 0; i < kScanPer; ++i) { t += rho[2 + i]; loc[i] = t; }      // inactive threads hold rho = 0
            float incl = t;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const float v = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += v;
            }
            if (lane == 31) S.wsum[k & 1][warp] = incl;
            excl32 = incl - t;
        }
        __syncthreads();                 // warp totals visible
        // Refill the stage of the PREVIOUS chunk.  A thread that has arrived here has issued every instruction of that
        // chunk's arithmetic, and instructions issue in order with their operands ready, so its shared-memory loads from
        // that stage have been performed.  Anything weaker races with the bulk copy of the async proxy: refilling THIS
        // chunk's stage after the barrier (its loads are merely in flight), or releasing it through mbarrier arrivals
        // issued behind the loads, both produced sporadic stale cells under scripts/stress_scan.py.
        if (tid == 0 && k >= 1 && k - 1 + kScanStages < nchunks) issue(k - 1 + kScanStages);
        double base = 0.0;
        if (MODE != 0) {
            // every warp scans the warp totals itself (one lane each, fp64): no second block barrier
            double w = (lane < kScanThreads / 32) ? (double)S.wsum[k & 1][lane] : 0.0;
#pragma unroll
            for (int o = 1; o < kScanThreads / 32; o <<= 1) {
                const double v = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += v;
            }
            const double total = __shfl_sync(0xffffffffu, w, kScanThreads / 32 - 1);
            const double before = __shfl_sync(0xffffffffu, w, (warp + 31) & 31);     // lane warp-1 (warp 0: unused)
            base = run + (warp ? before : 0.0) + (double)excl32;
            run += total;
        }
        if (!active) continue;

        float ev[kScanPer];
        if (MODE == 0) {
            const float4 e0 = *reinterpret_cast<const float4*>(pe + j0), e1 = *reinterpret_cast<const float4*>(pe + j0 + 4);
            ev[0] = e0.x; ev[1] = e0.y; ev[2] = e0.z; ev[3] = e0.w;
            ev[4] = e1.x; ev[5] = e1.y; ev[6] = e1.z; ev[7] = e1.w;
        } else {
            // E_j = -dx (C_j - rho_j/2 - (j + 1/2) rbar - mu) + (dx/24)(rho_{j+1} - rho_{j-1})
            const float t0 = (float)(-a.dx * (base - ((double)j0 + 0.5) * rbar - mu));
            float d4 = 0.f;
#pragma unroll
            for (int i = 0; i < kScanPer; ++i) {
                const float br = fmaf(-0.5f, rho[2 + i], loc[i]) - (float)i * rbarf;
                ev[i] = fmaf(-dxf, br, t0) + dx24 * (rho[3 + i] - rho[1 + i]);
                accE = fmaxf(accE, fabsf(ev[i]));
                const float q = fmaf(6.0f, rho[2 + i], fmaf(-4.0f, rho[1 + i] + rho[3 + i], rho[i] + rho[4 + i]));
                d4 = fmaf(q, q, d4);
            }
            accD += d4;
            if (a.e_out != nullptr) {
                float* po = a.e_out + (size_t)ic * 3 * nx + 2 * (size_t)nx + j0;
                *reinterpret_cast<float4*>(po) = make_float4(ev[0], ev[1], ev[2], ev[3]);
                *reinterpret_cast<float4*>(po + 4) = make_float4(ev[4], ev[5], ev[6], ev[7]);
            }
        }
        if (MODE == 2) continue;

        // ---- finite-volume update (bit-exact numpy order, field_kernels.cuh) and the sums of the new state ----
        float nn[kScanPer], un[kScanPer];
        float s1, m1;
        if (kPacked) {
            float2 s2 = make_float2(0.f, 0.f), m2 = make_float2(0.f, 0.f);
#pragma unroll
            for (int i = 0; i < kScanPer; i += 2) {
                float2 n2, u2;
                fv_pair(nv[1 + i], make_float2(nv[2 + i], nv[3 + i]), uv[i], make_float2(uv[1 + i], uv[2 + i]), uv[3 + i],
                        make_float2(ev[i], ev[i + 1]), a.c, a.dt, a.nu, a.dx2, rdx2, n2, u2);
                nn[i] = n2.x; nn[i + 1] = n2.y; un[i] = u2.x; un[i + 1] = u2.y;
                const float2 r2 = __fadd2_rn(n2, make_float2(-1.0f, -1.0f));
                s2 = __fadd2_rn(s2, r2);
                m2 = __ffma2_rn(make_float2((float)i, (float)(i + 1)), r2, m2);
            }
            s1 = s2.x + s2.y; m1 = m2.x + m2.y;
        } else {
            s1 = 0.f; m1 = 0.f;
#pragma unroll
            for (int i = 0; i < kScanPer; ++i) {
                const FvOut o = fv_cell(nv[1 + i], nv[2 + i], uv[i], uv[1 + i], uv[2 + i], ev[i], a.c, a.dt, a.nu, a.dx2, rdx2);
                nn[i] = o.n; un[i] = o.u;
                const float r = __fsub_rn(o.n, 1.0f);
                s1 += r;
                m1 = fmaf((float)i, r, m1);
            }
        }
        accS += (double)s1;
        accM = fma((double)j0, (double)s1, accM) + (double)m1;
        float* po = a.out + (size_t)ic * 3 * nx + j0;
        *reinterpret_cast<float4*>(po) = make_float4(nn[0], nn[1], nn[2], nn[3]);
        *reinterpret_cast<float4*>(po + 4) = make_float4(nn[4], nn[5], nn[6], nn[7]);
        *reinterpret_cast<float4*>(po + nx) = make_float4(un[0], un[1], un[2], un[3]);
        *reinterpret_cast<float4*>(po + nx + 4) = make_float4(un[4], un[5], un[6], un[7]);
        if (a.flux_out != nullptr) {                         // F_n = n u of the input state (:70-71, :84)
            float* pf = a.flux_out + (size_t)ic * nx + j0;
            *reinterpret_cast<float4*>(pf) = make_float4(__fmul_rn(nv[2], uv[1]), __fmul_rn(nv[3], uv[2]),
                                                         __fmul_rn(nv[4], uv[3]), __fmul_rn(nv[5], uv[4]));
            *reinterpret_cast<float4*>(pf + 4) = make_float4(__fmul_rn(nv[6], uv[5]), __fmul_rn(nv[7], uv[6]),
                                                             __fmul_rn(nv[8], uv[7]), __fmul_rn(nv[9], uv[8]));
        }
    }

    // ---- this segment's record ----
    accS = warp_sum(accS); accM = warp_sum(accM);
    double accD64 = warp_sum((double)accD);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) accE = fmaxf(accE, __shfl_xor_sync(0xffffffffu, accE, o));
    __syncthreads();                     // the prologue's use of S.red is over
    if (lane == 0) { S.red[0][warp] = accS; S.red[1][warp] = accM; S.red[2][warp] = accD64; S.redf[warp] = accE; }
    __syncthreads();
    if (tid == 0) {
        ScanRec r;
        r.S = 0.0; r.M1 = 0.0; r.D4 = 0.0; r.maxE = 0.f; r.pad = 0;
        for (int w = 0; w < kScanThreads / 32; ++w) {
            r.S += S.red[0][w]; r.M1 += S.red[1][w]; r.D4 += S.red[2][w];
            r.maxE = fmaxf(r.maxE, S.redf[w]);
        }
        if (MODE == 0) { r.D4 = 0.0; r.maxE = 1.0f; }
        a.rec_out[(size_t)ic * a.segs + seg] = r;
    }
}

// certificate of the last reconstructed field (the records of the materialise launch), and flag initialisation
__global__ void baseline_scan_flag_kernel(const ScanRec* rec, int B, int segs, int nx, double length, double tol, int step,
                                          int* flag, int init) {
    const int ic = blockIdx.x;
    if (init) { if (threadIdx.x == 0) flag[ic] = INT_MAX; return; }
    double d = 0.0;
    float em = 0.f;
    for (int r = threadIdx.x; r < segs; r += 32) {
        d += rec[(size_t)ic * segs + r].D4;
        em = fmaxf(em, rec[(size_t)ic * segs + r].maxE);
    }
    d = warp_sum(d);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) em = fmaxf(em, __shfl_xor_sync(0xffffffffu, em, o));
    if (threadIdx.x == 0) {
        const double bound = sqrt(d / (double)nx) * length * (1.0 / (32.0 * 1.7320508075688772));
        if (!(bound <= tol * (double)em)) atomicMin(flag + ic, step);
    }
}

// host mirror of fv_reciprocal()'s predicate: can dx2 be divided by through the Markstein sequence?
bool fv_reciprocal_host(float b) {
    unsigned bits;
    memcpy(&bits, &b, sizeof bits);
    return b > 0.f && b < 1.f && (bits & 0x7fffffu) != 0x7fffffu && (bits >> 23) != 0u;
}

void scan_geometry(int B, int nx, int sms, int* segs, int* seg_chunks) {
    const int nchunks = (nx + kScanChunk - 1) / kScanChunk;
    int target = (kScanCtasPerSm * sms) / B;
    if (target < 1) target = 1;
    if (target > nchunks) target = nchunks;
    const int sc = (nchunks + target - 1) / target;
    *seg_chunks = sc;
    *segs = (nchunks + sc - 1) / sc;
}

// ---------------------------------------------------------------------------------------------------------------
// The same solve for ONE SLAB of a domain-decomposed grid (gnn_plasma_flux_b200/domain.py, field_solve="scan"): the
// "distributed Poisson reduction" of BASELINE.json's north_star.  Rank r owns cells [r S, (r+1) S) of every IC.
//   slab_sums  : per-segment S = sum rho, M1 = sum j rho (j global)                          -> workspace
//   slab_msg   : the rank's 48-byte message per IC: totals, the certificate sums of the field it reconstructed last,
//                and its first / last two densities (the neighbours' stencil needs them)      -> all-gather, 48 B x B
//   slab_field : prefix over the ranks before this one + the segments before this CTA's, then the reconstruction
//                and certificate of scan_poisson's single-GPU kernel.  Plain loads (slabs of the hybrid solver start
//                at halo = L r + 1 cells, not 16-byte aligned), neighbours through warp shuffles.
// One collective of a few hundred bytes replaces the four all-to-alls (4 bytes per cell each) of the distributed FFT.
// ---------------------------------------------------------------------------------------------------------------
struct SlabMsg {
    double S, M1, D4;        // totals of this rank's slab: sum rho, sum j rho, sum (Delta^4 rho)^2 of its previous field
    float maxE;              // max |E| of its previous field
    float edge[4];           // n of its first two and last two cells
    float pad;
};
static_assert(sizeof(SlabMsg) == 48, "message layout");

struct SlabSeg { double S, M1, D4; float maxE; int pad; };      // per-segment scratch (sums now, certificate of the last field)

constexpr int kSlabThreads = 256;
constexpr int kSlabChunk = kSlabThreads * kScanPer;

__host__ __device__ inline void slab_geometry(int B, int S, int sms, int* segs, int* seg_chunks) {
    const int nchunks = (S + kSlabChunk - 1) / kSlabChunk;
    int target = (4 * sms) / B;
    if (target < 1) target = 1;
    if (target > nchunks) target = nchunks;
    const int sc = (nchunks + target - 1) / target;
    *seg_chunks = sc;
    *segs = (nchunks + sc - 1) / sc;
}

__device__ __forceinline__ void slab_load8(const float* row, long long j, bool vec, float (&v)[8]) {
    if (vec) {
        const float4 a = *reinterpret_cast<const float4*>(row + j), b = *reinterpret_cast<const float4*>(row + j + 4);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = row[j + i];
    }
}

__global__ void __launch_bounds__(kSlabThreads) slab_sums_kernel(const float* __restrict__ n, long long n_ld, int S, int segs,
                                                                 int seg_chunks, long long j_base, SlabSeg* __restrict__ seg_out) {
    __shared__ double red[2][kSlabThreads / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ic = (int)blockIdx.x / segs, seg = (int)blockIdx.x - ic * segs;
    const float* row = n + (size_t)ic * n_ld;
    const bool vec = ((reinterpret_cast<uintptr_t>(row) & 15) == 0) && (n_ld & 3) == 0;
    const long long begin = (long long)seg * seg_chunks * kSlabChunk;
    long long end = begin + (long long)seg_chunks * kSlabChunk;
    if (end > S) end = S;
    double accS = 0.0, accM = 0.0;
    for (long long j0 = begin + (long long)tid * kScanPer; j0 < end; j0 += kSlabChunk) {
        float v[8];
        slab_load8(row, j0, vec, v);
        float s1 = 0.f, m1 = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float r = __fsub_rn(v[i], 1.0f);
            s1 += r;
            m1 = fmaf((float)i, r, m1);
        }
        accS += (double)s1;
        accM = fma((double)(j_base + j0), (double)s1, accM) + (double)m1;
    }
    accS = warp_sum(accS); accM = warp_sum(accM);
    if (lane == 0) { red[0][warp] = accS; red[1][warp] = accM; }
    __syncthreads();
    if (tid == 0) {
        double a = 0.0, b = 0.0;
        for (int w = 0; w < kSlabThreads / 32; ++w) { a += red[0][w]; b += red[1][w]; }
        SlabSeg& o = seg_out[(size_t)ic * segs + seg];      // D4 / maxE keep the certificate of the previous field
        o.S = a; o.M1 = b;
    }
}

// one warp per IC; lane l sums the records l, l + 32, ..., then a fixed butterfly: deterministic, and 32 loads in flight
// instead of a serial chain of `segs` dependent ones (16 us at 74 segments)
// Peer memory (bases != nullptr): the message is also stored into slot `rank` of EVERY rank's gather buffer, which sits
// `offset` bytes into that rank's symmetric allocation -- the all-gather of the distributed solve is this store.
__global__ void slab_msg_kernel(const float* __restrict__ n, long long n_ld, int S, int segs, const SlabSeg* __restrict__ seg_rec,
                                SlabMsg* __restrict__ msg, void* const* __restrict__ bases, long long offset, int rank,
                                int world, int B) {
    const int ic = blockIdx.x, lane = threadIdx.x;
    double s = 0.0, m1 = 0.0, d4 = 0.0;
    float em = 0.f;
    for (int q = lane; q < segs; q += 32) {
        const SlabSeg r = seg_rec[(size_t)ic * segs + q];
        s += r.S; m1 += r.M1; d4 += r.D4; em = fmaxf(em, r.maxE);
    }
    s = warp_sum(s); m1 = warp_sum(m1); d4 = warp_sum(d4);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) em = fmaxf(em, __shfl_xor_sync(0xffffffffu, em, o));
    if (lane != 0) return;
    SlabMsg m;
    m.S = s; m.M1 = m1; m.D4 = d4; m.maxE = em; m.pad = 0.f;
    const float* row = n + (size_t)ic * n_ld;
    m.edge[0] = row[0]; m.edge[1] = row[1]; m.edge[2] = row[S - 2]; m.edge[3] = row[S - 1];
    msg[ic] = m;
    if (bases != nullptr)
        for (int p = 0; p < world; ++p)
            reinterpret_cast<SlabMsg*>(static_cast<char*>(bases[p]) + offset)[(size_t)rank * B + ic] = m;
}

struct SlabFieldArgs {
    const float* n; long long n_ld; float* E; long long e_ld;
    int B, S, segs, seg_chunks, rank, ranks, step;
    double dx, length, tol;
    const SlabMsg* msg_all;      // [ranks][B]
    SlabSeg* seg_rec;            // [B][segs]: S, M1 of this launch's density in; D4, maxE of the field reconstructed here out
    int* flag;
    float* E_left;               // slabs over peer memory (nullable): the E rows of the ring neighbours' next states, laid out like
    float* E_right;              //   E / e_ld.  The first / last `halo` cells of this slab's field are ALSO stored into the left
    int halo;                    //   neighbour's right ghost zone (its cells S .. S+halo-1) / the right one's left ghosts (-halo .. -1)
};

__global__ void __launch_bounds__(kSlabThreads) slab_field_kernel(const SlabFieldArgs a) {
    __shared__ float wsum[2][kSlabThreads / 32];
    __shared__ double redd[kSlabThreads / 32];
    __shared__ float redf[kSlabThreads / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ic = (int)blockIdx.x / a.segs, seg = (int)blockIdx.x - ic * a.segs;
    const int S = a.S;
    const float* row = a.n + (size_t)ic * a.n_ld;
    float* out = a.E + (size_t)ic * a.e_ld;
    const bool vec_in = ((reinterpret_cast<uintptr_t>(row) & 15) == 0) && (a.n_ld & 3) == 0;
    const bool vec_out = ((reinterpret_cast<uintptr_t>(out) & 15) == 0) && (a.e_ld & 3) == 0;
    const long long begin = (long long)seg * a.seg_chunks * kSlabChunk;
    long long end = begin + (long long)a.seg_chunks * kSlabChunk;
    if (end > S) end = S;

    // ---- prologue: totals over the ranks, prefix of the ranks and segments before (parallel loads, fixed reduction order) ----
    double s_tot = 0.0, m_tot = 0.0, d_tot = 0.0, P = 0.0;
    float e_max = 0.f;
    {
        double v_s = 0.0, v_m = 0.0, v_d = 0.0, v_p = 0.0;
        float v_e = 0.f;
        for (int g = tid; g < a.ranks; g += kSlabThreads) {
            const SlabMsg& m = a.msg_all[(size_t)g * a.B + ic];
            v_s += m.S; v_m += m.M1; v_d += m.D4; v_e = fmaxf(v_e, m.maxE);
            if (g < a.rank) v_p += m.S;
        }
        for (int q = tid; q < seg; q += kSlabThreads) v_p += a.seg_rec[(size_t)ic * a.segs + q].S;
        v_s = warp_sum(v_s); v_m = warp_sum(v_m); v_d = warp_sum(v_d); v_p = warp_sum(v_p);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v_e = fmaxf(v_e, __shfl_xor_sync(0xffffffffu, v_e, o));
        __shared__ double pro[4][kSlabThreads / 32];
        __shared__ float prof[kSlabThreads / 32];
        if (lane == 0) { pro[0][warp] = v_s; pro[1][warp] = v_m; pro[2][warp] = v_d; pro[3][warp] = v_p; prof[warp] = v_e; }
        __syncthreads();
#pragma unroll
        for (int w = 0; w < kSlabThreads / 32; ++w) {
            s_tot += pro[0][w]; m_tot += pro[1][w]; d_tot += pro[2][w]; P += pro[3][w];
            e_max = fmaxf(e_max, prof[w]);
        }
    }
    const double N = (double)S * (double)a.ranks;
    const double rbar = s_tot / N, mu = 0.5 * s_tot - m_tot / N - s_tot / (2.0 * N);
    if (seg == 0 && tid == 0 && a.step > 0) {              // certificate of the field of the previous call
        const double bound = sqrt(d_tot / N) * a.length * (1.0 / (32.0 * 1.7320508075688772));
        if (!(bound <= a.tol * (double)e_max)) atomicMin(a.flag, a.step - 1);
    }
    const SlabMsg& ml = a.msg_all[(size_t)((a.rank + a.ranks - 1) % a.ranks) * a.B + ic];
    const SlabMsg& mr = a.msg_all[(size_t)((a.rank + 1) % a.ranks) * a.B + ic];
    const float left2[2] = {ml.edge[2], ml.edge[3]}, right2[2] = {mr.edge[0], mr.edge[1]};
    auto fetch = [&](long long j) -> float {              // density of slab cell j in [-2, S+1]
        return j < 0 ? left2[j + 2] : (j >= S ? right2[j - S] : row[j]);
    };

    const float dxf = (float)a.dx, dx24 = (float)(a.dx / 24.0), rbarf = (float)rbar;
    const long long j_base = (long long)a.rank * S;
    double run = P;
    float accD = 0.f, accE = 0.f;
    int k = 0;
    for (long long c0 = begin; c0 < end; c0 += kSlabChunk, ++k) {
        const long long j0 = c0 + (long long)tid * kScanPer;
        const bool active = j0 < end;
        float v[8];
        if (active) slab_load8(row, j0, vec_in, v);
        else {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = 1.0f;
        }
        // neighbours: the adjacent lanes hold them, except at warp, chunk and slab ends
        float l0 = __shfl_up_sync(0xffffffffu, v[6], 1), l1 = __shfl_up_sync(0xffffffffu, v[7], 1);
        float r0 = __shfl_down_sync(0xffffffffu, v[0], 1), r1 = __shfl_down_sync(0xffffffffu, v[1], 1);
        if (active) {
            if (lane == 0) { l0 = fetch(j0 - 2); l1 = fetch(j0 - 1); }
            if (lane == 31 || j0 + 8 >= end) { r0 = fetch(j0 + 8); r1 = fetch(j0 + 9); }
        }
        float rho[12];
        rho[0] = __fsub_rn(l0, 1.0f); rho[1] = __fsub_rn(l1, 1.0f);
#pragma unroll
        for (int i = 0; i < 8; ++i) rho[2 + i] = __fsub_rn(v[i], 1.0f);
        rho[10] = __fsub_rn(r0, 1.0f); rho[11] = __fsub_rn(r1, 1.0f);
        float loc[8], t = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) { t += rho[2 + i]; loc[i] = t; }
        float incl = t;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const float u = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += u;
        }
        if (lane == 31) wsum[k & 1][warp] = incl;
        __syncthreads();
        double w = (lane < kSlabThreads / 32) ? (double)wsum[k & 1][lane] : 0.0;
#pragma unroll
        for (int o = 1; o < kSlabThreads / 32; o <<= 1) {
            const double u = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= o) w += u;
        }
        const double total = __shfl_sync(0xffffffffu, w, kSlabThreads / 32 - 1);
        const double before = __shfl_sync(0xffffffffu, w, (warp + 31) & 31);
        const double base = run + (warp ? before : 0.0) + (double)(incl - t);
        run += total;
        if (!active) continue;
        const float t0 = (float)(-a.dx * (base - ((double)(j_base + j0) + 0.5) * rbar - mu));
        float ev[8], d4 = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float br = fmaf(-0.5f, rho[2 + i], loc[i]) - (float)i * rbarf;
            ev[i] = fmaf(-dxf, br, t0) + dx24 * (rho[3 + i] - rho[1 + i]);
            accE = fmaxf(accE, fabsf(ev[i]));
            const float q = fmaf(6.0f, rho[2 + i], fmaf(-4.0f, rho[1 + i] + rho[3 + i], rho[i] + rho[4 + i]));
            d4 = fmaf(q, q, d4);
        }
        accD += d4;
        if (vec_out) {
            *reinterpret_cast<float4*>(out + j0) = make_float4(ev[0], ev[1], ev[2], ev[3]);
            *reinterpret_cast<float4*>(out + j0 + 4) = make_float4(ev[4], ev[5], ev[6], ev[7]);
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) out[j0 + i] = ev[i];
        }
        if (a.E_left != nullptr && j0 < a.halo) {              // the halo exchange of E' is this store
            float* pl = a.E_left + (size_t)ic * a.e_ld + S;
#pragma unroll
            for (int i = 0; i < 8; ++i)
                if (j0 + i < a.halo) pl[j0 + i] = ev[i];
        }
        if (a.E_right != nullptr && j0 + 8 > S - a.halo) {
            float* pr = a.E_right + (size_t)ic * a.e_ld - S;
#pragma unroll
            for (int i = 0; i < 8; ++i)
                if (j0 + i >= S - a.halo) pr[j0 + i] = ev[i];
        }
    }
    double d = warp_sum((double)accD);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) accE = fmaxf(accE, __shfl_xor_sync(0xffffffffu, accE, o));
    if (lane == 0) { redd[warp] = d; redf[warp] = accE; }
    __syncthreads();
    if (tid == 0) {
        double dd = 0.0;
        float em = 0.f;
        for (int w2 = 0; w2 < kSlabThreads / 32; ++w2) { dd += redd[w2]; em = fmaxf(em, redf[w2]); }
        SlabSeg& o = a.seg_rec[(size_t)ic * a.segs + seg];
        o.D4 = dd; o.maxE = em;
    }
}

// the certificate of the LAST field of a run: the messages of one more sums + all-gather round carry its sums
__global__ void slab_certify_kernel(const SlabMsg* __restrict__ msg_all, int B, int S, int ranks, double length, double tol,
                                    int step, int* flag) {
    const int ic = blockIdx.x * blockDim.x + threadIdx.x;
    if (ic >= B) return;
    double d = 0.0;
    float em = 0.f;
    for (int g = 0; g < ranks; ++g) {
        d += msg_all[(size_t)g * B + ic].D4;
        em = fmaxf(em, msg_all[(size_t)g * B + ic].maxE);
    }
    const double bound = sqrt(d / ((double)S * ranks)) * length * (1.0 / (32.0 * 1.7320508075688772));
    if (!(bound <= tol * (double)em)) atomicMin(flag, step);
}

}  // namespace

bool baseline_scan_supported(int B, int nx) {
    return nx >= 4096 && nx <= (1 << 30) && (nx % kScanPer) == 0 && B >= 1 && B <= 4096;      // cell indices are ints
}

size_t baseline_scan_workspace_bytes(int B, int nx, int sms) {
    int segs = 0, sc = 0;
    scan_geometry(B, nx, sms, &segs, &sc);
    return (size_t)B * 3 * nx * sizeof(float) + 2 * (size_t)B * segs * sizeof(ScanRec) + 256;
}

// T-step rollout; see fluxgnn_baseline_rollout_scan in include/fluxgnn.h.  Returns a CUDA error code (0 = ok) and the
// number of kernels launched through *launches.
cudaError_t launch_baseline_rollout_scan(const float* state_in, float* state_out, int B, int nx, double length, float c, float dt,
                                         float nu, float dx2, int steps, int record_every, float* traj, float* flux_n,
                                         double tol, void* workspace, int* flag, int sms, cudaStream_t stream, int* launches) {
    int segs = 0, sc = 0;
    scan_geometry(B, nx, sms, &segs, &sc);
    const size_t state_floats = (size_t)B * 3 * nx;
    float* tmp = (float*)workspace;
    ScanRec* rec[2];
    rec[0] = reinterpret_cast<ScanRec*>(reinterpret_cast<unsigned char*>(workspace) +
                                        ((state_floats * sizeof(float) + 255) / 256) * 256);
    rec[1] = rec[0] + (size_t)B * segs;
    const size_t smem = sizeof(ScanSmem);
    cudaError_t e;
    const bool packed = fv_reciprocal_host(dx2);
    auto k_first = packed ? baseline_scan_kernel<0, true> : baseline_scan_kernel<0, false>;
    auto k_step = packed ? baseline_scan_kernel<1, true> : baseline_scan_kernel<1, false>;
    auto k_final = baseline_scan_kernel<2, false>;
    if ((e = cudaFuncSetAttribute(k_first, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(k_step, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(k_final, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
    int nl = 0;
    baseline_scan_flag_kernel<<<B, 32, 0, stream>>>(nullptr, B, segs, nx, length, tol, 0, flag, 1);
    ++nl;
    ScanArgs a;
    a.B = B; a.nx = nx; a.segs = segs; a.seg_chunks = sc;
    a.c = c; a.dt = dt; a.nu = nu; a.dx2 = dx2;
    a.dx = length / (double)nx; a.length = length; a.tol = tol;
    a.flag = flag;
    const unsigned grid = (unsigned)(B * segs);
    const float* src = state_in;
    for (int t = 0; t < steps; ++t) {
        const bool recorded = traj != nullptr && (t + 1) % record_every == 0;
        // ping-pong between the workspace and state_out such that the last step lands in state_out; a recorded state is
        // written straight into its trajectory slot (the next step reads it from there and adds its field)
        float* dst = ((steps - 1 - t) % 2 == 0) ? state_out : tmp;
        if (recorded && t != steps - 1) dst = traj + (size_t)((t + 1) / record_every - 1) * state_floats;
        a.in = src; a.out = dst;
        a.flux_out = flux_n ? flux_n + (size_t)t * B * nx : nullptr;
        a.rec_in = rec[(t + 1) & 1]; a.rec_out = rec[t & 1];
        a.step = t;
        // the field of a recorded (or the caller's) input state is stored into that state's E plane
        const bool in_recorded = t > 0 && traj != nullptr && t % record_every == 0;
        a.e_out = in_recorded ? const_cast<float*>(src) : nullptr;
        if (t == 0) k_first<<<grid, kScanThreads, smem, stream>>>(a);
        else k_step<<<grid, kScanThreads, smem, stream>>>(a);
        ++nl;
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
        src = dst;
    }
    // materialise the field of the final state and certify it
    a.in = state_out; a.out = nullptr; a.e_out = state_out; a.flux_out = nullptr;
    a.rec_in = rec[(steps + 1) & 1]; a.rec_out = rec[steps & 1];
    a.step = steps;
    k_final<<<grid, kScanThreads, smem, stream>>>(a);
    ++nl;
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    baseline_scan_flag_kernel<<<B, 32, 0, stream>>>(rec[steps & 1], B, segs, nx, length, tol, steps, flag, 0);
    ++nl;
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    if (traj != nullptr && steps % record_every == 0) {
        e = cudaMemcpyAsync(traj + (size_t)(steps / record_every - 1) * state_floats, state_out, state_floats * sizeof(float),
                            cudaMemcpyDeviceToDevice, stream);
        if (e != cudaSuccess) return e;
    }
    *launches = nl;
    return cudaSuccess;
}

// ---- slab (domain-decomposed) entry points; see include/fluxgnn.h ----
bool scan_slab_supported(int B, int S) { return B >= 1 && B <= 4096 && S >= 64 && (S % kScanPer) == 0; }

size_t scan_slab_workspace_bytes(int B, int S, int sms) {
    int segs = 0, sc = 0;
    slab_geometry(B, S, sms, &segs, &sc);
    return (size_t)B * segs * sizeof(SlabSeg);
}

cudaError_t launch_scan_slab_sums(const float* n, long long n_ld, int B, int S, long long j_base, void* workspace, void* msg,
                                  int sms, cudaStream_t stream, void* const* bases, long long offset, int rank, int world) {
    int segs = 0, sc = 0;
    slab_geometry(B, S, sms, &segs, &sc);
    slab_sums_kernel<<<(unsigned)(B * segs), kSlabThreads, 0, stream>>>(n, n_ld, S, segs, sc, j_base, (SlabSeg*)workspace);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    slab_msg_kernel<<<(unsigned)B, 32, 0, stream>>>(n, n_ld, S, segs, (const SlabSeg*)workspace, (SlabMsg*)msg, bases, offset,
                                                    rank, world, B);
    return cudaGetLastError();
}

cudaError_t launch_scan_slab_field(const float* n, long long n_ld, float* E, long long e_ld, int B, int S, int rank, int ranks,
                                   double length, const void* msg_all, void* workspace, double tol, int step, int* flag, int sms,
                                   cudaStream_t stream, float* E_left, float* E_right, int halo) {
    SlabFieldArgs a;
    a.E_left = E_left; a.E_right = E_right; a.halo = halo;
    a.n = n; a.n_ld = n_ld; a.E = E; a.e_ld = e_ld;
    a.B = B; a.S = S; a.rank = rank; a.ranks = ranks; a.step = step;
    slab_geometry(B, S, sms, &a.segs, &a.seg_chunks);
    a.length = length; a.dx = length / ((double)S * ranks); a.tol = tol;
    a.msg_all = (const SlabMsg*)msg_all; a.seg_rec = (SlabSeg*)workspace; a.flag = flag;
    slab_field_kernel<<<(unsigned)(B * a.segs), kSlabThreads, 0, stream>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_scan_slab_certify(int B, int S, int ranks, double length, const void* msg_all, double tol, int step, int* flag,
                                     cudaStream_t stream) {
    slab_certify_kernel<<<(unsigned)((B + 63) / 64), 64, 0, stream>>>((const SlabMsg*)msg_all, B, S, ranks, length, tol, step, flag);
    return cudaGetLastError();
}

}  // namespace fluxgnn
