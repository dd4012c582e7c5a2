// C ABI of libfluxgnn.so (declared in include/fluxgnn.h): argument checking,
// tiling decisions and launch sequencing.  No allocation, no synchronisation,
// no CPU fallback.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdint>
#include <cstdlib>

#include "common.cuh"
#include "field_kernels.cuh"
#include "hybrid_kernel.cuh"
#include "train_kernels.cuh"

namespace fluxgnn {

static thread_local char g_err[512] = "";
static std::atomic<unsigned long long> g_launches{0};

int set_error(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

int cuda_fail(cudaError_t err, const char* what) {
    return set_error(FLUXGNN_ECUDA, "CUDA error %d (%s) in %s", (int)err, cudaGetErrorString(err), what);
}

void count_launch(int n) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }

static int sm_count(int* out) {
    int dev = 0;
    FLUXGNN_CUDA_OK(cudaGetDevice(&dev));
    FLUXGNN_CUDA_OK(cudaDeviceGetAttribute(out, cudaDevAttrMultiProcessorCount, dev));
    return FLUXGNN_OK;
}

// FP32-pipe kernel, window / slab tiles: a cluster of c CTAs shares ONE window of 128c - 4(c-1) rows (the pieces overlap
// by one 4-row chunk and read each other's edge rows of Z through distributed shared memory), so only the window's two
// outer ends carry the recomputed halo.  Pick the c in 1..4 that needs the fewest SM-tile-steps for this grid, counting
// the SMs the hardware can actually fill with clusters of that size.  FLUXGNN_CLUSTER=c forces it (1 = off; test hook).
static void choose_cluster(HybridArgs& a, int cells) {
    int sms = 0;
    if (sm_count(&sms) != FLUXGNN_OK) return;
    const char* force = getenv("FLUXGNN_CLUSTER");
    const int forced = (force != nullptr && force[0] >= '1' && force[0] <= '4') ? force[0] - '0' : 0;
    double best = 0.0;
    int best_c = 1;
    for (int c = 1; c <= 4; ++c) {
        if (forced && c != forced) continue;
        const int rows = kTileRows * c - 4 * (c - 1), valid = rows - 2 * a.halo;
        if (valid < 8) continue;
        const int eff_sms = c == 1 ? sms : hybrid_max_active_clusters(c) * c;
        if (eff_sms < 1) continue;
        const long long windows = (long long)a.B * ((cells + valid - 1) / valid);
        const long long slots = eff_sms / c;                                   // windows in flight
        const long long waves = (windows + slots - 1) / slots;                 // each wave costs one tile-step
        // measured on B200 (scripts/time_cluster_windows.py, profiles/r2_cluster_windows.md): at equal tile count a
        // clustered step is ~7 % slower (lock-step of the pieces, remote reads), so clusters must save more than that
        const double cost = (double)waves * (c == 1 ? 1.0 : 1.07);
        if (best == 0.0 || cost < best) {
            best = cost;
            best_c = c;
        }
    }
    if (best_c > 1) {
        a.cluster = best_c;
        a.valid = kTileRows * best_c - 4 * (best_c - 1) - 2 * a.halo;
        a.tiles_per_ic = (cells + a.valid - 1) / a.valid;
        a.num_tiles = (int)((long long)a.B * a.tiles_per_ic);
    }
}

// Tiling of [B][nx] cells into 128-row tiles (see hybrid_kernel.cu).
static int plan_tiles(HybridArgs& a, int* fast_radius) {
    const int nx = a.nx;
    if (a.tile_rows == 0) a.tile_rows = kTileRows;
    // height of a logical tile: 128; the 16-bit tensor kernel packs two 128-row groups (or one 256-row group) per CTA
    if (a.tile_rows == kTc16TileRows && a.tc_group_rows == 0) {
        // two independent 128-row groups overlap their epilogues with each other's products (1.36x per tile);
        // as windows they carry twice the halo, which pays up to a receptive field of 24 cells
        const char* nosplit = getenv("FLUXGNN_TC16_NO_SPLIT");         // test hook
        const bool forced_single = nosplit != nullptr && nosplit[0] == '1';
        const int halo = a.L * a.radius + a.hops;
        a.tc_group_rows = (!forced_single && (nx <= kTileRows || halo <= 24)) ? 128 : kTc16TileRows;
    }
    const int rows = a.tc_group_rows ? a.tc_group_rows : a.tile_rows;
    if (nx <= kTileRows) {                         // whole ICs (the API's workspace / gtab rules key on 128)
        a.whole_ic = 1;
        a.ics_per_tile = rows / nx;
        a.tiles_per_ic = 0;
        a.valid = a.halo = 0;
        const long long tiles = ((long long)a.B + a.ics_per_tile - 1) / a.ics_per_tile;
        if (tiles > 0x7fffffffLL) return set_error(FLUXGNN_EINVAL, "too many tiles");
        a.num_tiles = (int)tiles;
        *fast_radius = (nx % 8 == 0 && a.radius <= 4) ? a.radius : 0;
    } else {
        a.whole_ic = 0;
        a.ics_per_tile = 0;
        a.halo = a.L * a.radius + a.hops;
        a.valid = rows - 2 * a.halo;
        if (a.valid < 8)
            return set_error(FLUXGNN_EUNSUP, "receptive field L*radius+hops = %d cells does not fit a %d-cell tile", a.halo, rows);
        a.tiles_per_ic = (nx + a.valid - 1) / a.valid;
        const long long tiles = (long long)a.B * a.tiles_per_ic;
        if (tiles > 0x7fffffffLL) return set_error(FLUXGNN_EINVAL, "too many tiles");
        a.num_tiles = (int)tiles;
        *fast_radius = (a.radius <= 4) ? a.radius : 0;
        if (a.tile_rows == kTileRows && a.tc_group_rows == 0 && a.tc_parts == 0 && a.acts == nullptr && *fast_radius > 0)
            choose_cluster(a, nx);
        // Packed remainders: when the last window of an IC is short enough that several of them (each with its own halo
        // on both sides) fit one tile, the last windows of consecutive ICs share tiles -- BASELINE configs[2]
        // (1024 cells, radius 2: 9 windows of 110 cells + 34) takes 9.5 instead of 10 tiles per IC.  The arithmetic of
        // a row does not depend on where in which tile it is computed: results are bit-identical (test hook
        // FLUXGNN_NO_PACK=1 disables it).
        const char* nopack = getenv("FLUXGNN_NO_PACK");
        const int full = nx / a.valid, rem = nx - full * a.valid;
        if (a.cluster <= 1 && full >= 1 && rem > 0 && a.B > 1 && !(nopack != nullptr && nopack[0] == '1')) {
            const int seg = rem + 2 * a.halo, per_tile = rows / seg;
            if (per_tile >= 2) {
                a.pack_full = full;
                a.pack_per_tile = per_tile < a.B ? per_tile : a.B;
                a.pack_seg = seg;
                a.pack_rem = rem;
                a.tiles_per_ic = full;
                const long long packed_tiles = (long long)a.B * full + ((long long)a.B + a.pack_per_tile - 1) / a.pack_per_tile;
                if (packed_tiles > 0x7fffffffLL) return set_error(FLUXGNN_EINVAL, "too many tiles");
                a.num_tiles = (int)packed_tiles;
            }
        }
    }
    // test hook: walk the prev/next tables even where the 128-bit window path applies
    const char* force = getenv("FLUXGNN_FORCE_GENERIC");
    if (force != nullptr && force[0] == '1') *fast_radius = 0;
    // two skewed 64-row groups when the halves of a tile hold different ICs (test hook to disable)
    const char* nosplit = getenv("FLUXGNN_NO_SPLIT");
    a.split = (a.whole_ic && nx <= 64 && 64 % nx == 0 && !(nosplit != nullptr && nosplit[0] == '1')) ? 1 : 0;
    return FLUXGNN_OK;
}

static int check_model(const void* packed, int L, int B, int nx, int radius) {
    if (packed == nullptr) return set_error(FLUXGNN_EINVAL, "packed weights pointer is null");
    if (L < 1 || L > kMaxL) return set_error(FLUXGNN_EUNSUP, "num_layers must be in 1..%d, got %d", kMaxL, L);
    if (B < 1 || nx < 1) return set_error(FLUXGNN_EINVAL, "B and nx must be >= 1 (B=%d nx=%d)", B, nx);
    if (radius < 1) return set_error(FLUXGNN_EINVAL, "radius must be >= 1, got %d", radius);
    return FLUXGNN_OK;
}

// fast_radius >= 0: FP32-pipe kernel (0 = generic walk); < 0: tensor-core kernel of radius -fast_radius
static int launch_tiles(const HybridArgs& a, int fast_radius, cudaStream_t stream) {
    int sms = 0;
    int rc = sm_count(&sms);
    if (rc != FLUXGNN_OK) return rc;
    const int groups = a.tc_group_rows ? a.tile_rows / a.tc_group_rows : 1;       // logical tiles per CTA tile
    const int cta_tiles = (a.num_tiles + groups - 1) / groups;
    int grid = cta_tiles < sms ? cta_tiles : sms;
    if (a.cluster > 1) {                                                          // one window per cluster
        const int slots = hybrid_max_active_clusters(a.cluster);
        grid = (a.num_tiles < slots ? a.num_tiles : slots) * a.cluster;
    }
    if (fast_radius < 0 && a.tile_rows == kTc16TileRows)
        FLUXGNN_CUDA_OK(launch_hybrid_tc16_tiles(a, -fast_radius, grid, stream));
    else if (fast_radius < 0)
        FLUXGNN_CUDA_OK(launch_hybrid_tc_tiles(a, -fast_radius, grid, stream));
    else
        FLUXGNN_CUDA_OK(launch_hybrid_tiles(a, fast_radius, grid, stream));
    count_launch();
    return FLUXGNN_OK;
}

// Field solve dispatch: FFT for power-of-two grids of 256 cells and more, otherwise the
// direct circular convolution with the fp64 table (any nx up to kPoissonDirectMaxNx).
static int launch_poisson(const float* n, long long ns, float* E, long long es, const double* gtab,
                          int B, int nx, double length, void* fft_ws, cudaStream_t stream) {
    if (poisson_fft_supported(nx)) return launch_poisson_fft(n, ns, E, es, B, nx, length, fft_ws, stream);
    if (nx > kPoissonDirectMaxNx)
        return set_error(FLUXGNN_EUNSUP, "field solve for nx=%d: only powers of two are supported above %d cells",
                         nx, kPoissonDirectMaxNx);
    if (gtab == nullptr) return set_error(FLUXGNN_EINVAL, "field solve for nx=%d needs the fluxgnn_poisson_table()", nx);
    dim3 grid((unsigned)B, (unsigned)((nx + 255) / 256));
    poisson_direct_kernel<<<grid, 256, (size_t)nx * sizeof(float), stream>>>(n, ns, E, es, gtab, nx);
    FLUXGNN_CUDA_OK(cudaGetLastError());
    count_launch();
    return FLUXGNN_OK;
}

// generic_kernels.cu: the same field-solve dispatch for the generic-architecture rollout
int launch_poisson_for_generic(const float* n, long long ns, float* E, long long es, const double* gtab, int B, int nx,
                               double length, void* fft_ws, cudaStream_t stream) {
    return launch_poisson(n, ns, E, es, gtab, B, nx, length, fft_ws, stream);
}

}  // namespace fluxgnn

using namespace fluxgnn;

extern "C" {

int fluxgnn_abi_version(void) { return FLUXGNN_ABI_VERSION; }

/* measurement aid (scripts/time_cluster_windows.py): clusters of `csize` CTAs of the FP32-pipe tile kernel the device holds at once */
int fluxgnn_debug_max_clusters(int csize) { return hybrid_max_active_clusters(csize); }

const char* fluxgnn_last_error(void) { return g_err; }

unsigned long long fluxgnn_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

long long fluxgnn_ffma_probe(float* out, int blocks, int threads, int iters, int packed, void* stream) {
    if (out == nullptr || blocks < 1 || iters < 1 || threads < 32 || threads > 256 || threads % 32)
        return set_error(FLUXGNN_EINVAL, "ffma_probe: bad argument");
    if (packed)
        ffma2_probe_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(out, iters, 0.999f, 0.001f);
    else
        ffma_probe_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(out, iters, 0.999f, 0.001f);
    FLUXGNN_CUDA_OK(cudaGetLastError());
    count_launch();
    return (long long)blocks * threads * iters * kFfmaProbeFlopsPerIter;
}

size_t fluxgnn_packed_weight_bytes(int num_layers) {
    if (num_layers < 1 || num_layers > kMaxL) return 0;
    return packed_floats(num_layers) * sizeof(float);
}

int fluxgnn_pack_weights(const float* w_in, const float* b_in, const float* w_upd, const float* b_upd,
                         const float* w_e1, const float* b_e1, const float* w_e2, const float* b_e2,
                         int num_layers, void* packed, void* stream) {
    if (num_layers < 1 || num_layers > kMaxL)
        return set_error(FLUXGNN_EUNSUP, "num_layers must be in 1..%d, got %d", kMaxL, num_layers);
    if (!w_in || !b_in || !w_upd || !b_upd || !w_e1 || !b_e1 || !w_e2 || !b_e2 || !packed)
        return set_error(FLUXGNN_EINVAL, "null weight pointer");
    const size_t total = packed_floats(num_layers);
    const int blocks = (int)((total + 255) / 256);
    pack_weights_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(w_in, b_in, w_upd, b_upd, w_e1, b_e1, w_e2, b_e2,
                                                                   num_layers, (float*)packed);
    FLUXGNN_CUDA_OK(cudaGetLastError());
    count_launch();
    return FLUXGNN_OK;
}

size_t fluxgnn_packed_tc_weight_bytes(int num_layers) {
    if (num_layers < 1 || num_layers > kMaxL) return 0;
    return packed_tc_floats(num_layers) * sizeof(float);
}

int fluxgnn_pack_weights_tc(const float* w_in, const float* b_in, const float* w_upd, const float* b_upd,
                            const float* w_e1, const float* b_e1, const float* w_e2, const float* b_e2,
                            int num_layers, void* packed, void* stream) {
    if (num_layers < 1 || num_layers > kMaxL)
        return set_error(FLUXGNN_EUNSUP, "num_layers must be in 1..%d, got %d", kMaxL, num_layers);
    if (!w_in || !b_in || !w_upd || !b_upd || !w_e1 || !b_e1 || !w_e2 || !b_e2 || !packed)
        return set_error(FLUXGNN_EINVAL, "null weight pointer");
    const size_t total = packed_tc_floats(num_layers);
    const int blocks = (int)((total + 255) / 256);
    pack_weights_tc_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(w_in, b_in, w_upd, b_upd, w_e1, b_e1, w_e2, b_e2,
                                                                      num_layers, (float*)packed);
    FLUXGNN_CUDA_OK(cudaGetLastError());
    count_launch();
    return FLUXGNN_OK;
}

size_t fluxgnn_packed_tc16_weight_bytes(int num_layers) {
    if (num_layers < 1 || num_layers > kMaxL) return 0;
    return packed_tc16_bytes(num_layers);
}

int fluxgnn_pack_weights_tc16(const float* w_in, const float* b_in, const float* w_upd, const float* b_upd,
                              const float* w_e1, const float* b_e1, const float* w_e2, const float* b_e2,
                              int num_layers, int precision, void* packed, void* stream) {
    if (num_layers < 1 || num_layers > kMaxL)
        return set_error(FLUXGNN_EUNSUP, "num_layers must be in 1..%d, got %d", kMaxL, num_layers);
    if (precision != FLUXGNN_TC_FP16X3 && precision != FLUXGNN_TC_FP16 && precision != FLUXGNN_TC_BF16)
        return set_error(FLUXGNN_EINVAL, "pack_weights_tc16: precision must be FLUXGNN_TC_FP16X3, _FP16 or _BF16");
    if (!w_in || !b_in || !w_upd || !b_upd || !w_e1 || !b_e1 || !w_e2 || !b_e2 || !packed)
        return set_error(FLUXGNN_EINVAL, "null weight pointer");
    const size_t total = packed_tc16_bytes(num_layers) / 2;
    const int blocks = (int)((total + 255) / 256);
    pack_weights_tc16_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(w_in, b_in, w_upd, b_upd, w_e1, b_e1, w_e2, b_e2,
                                                                        num_layers, precision == FLUXGNN_TC_BF16 ? 1 : 0,
                                                                        (unsigned char*)packed);
    FLUXGNN_CUDA_OK(cudaGetLastError());
    count_launch();
    return FLUXGNN_OK;
}

int fluxgnn_poisson_table(int nx, double length, double* gtab, void* stream) {
    if (nx < 1 || !(length > 0.0) || gtab == nullptr)
        return set_error(FLUXGNN_EINVAL, "poisson_table: nx=%d length=%g gtab=%p", nx, length, (void*)gtab);
    if (nx > kPoissonDirectMaxNx)
        return set_error(FLUXGNN_EUNSUP, "poisson_table: the direct solve covers nx <= %d", kPoissonDirectMaxNx);
    poisson_table_kernel<<<nx, 128, 0, (cudaStream_t)stream>>>(nx, length, gtab);
    FLUXGNN_CUDA_OK(cudaGetLastError());
    count_launch();
    return FLUXGNN_OK;
}

int fluxgnn_poisson_uses_table(int nx) { return (nx >= 1 && !poisson_fft_supported(nx)) ? 1 : 0; }

size_t fluxgnn_poisson_workspace_bytes(int B, int nx) {
    if (B < 1 || nx < 1 || !poisson_fft_supported(nx)) return 0;
    return poisson_fft_workspace_bytes(B, nx);
}

int fluxgnn_poisson_spectral(const float* n, long long n_ic_stride, float* E, long long e_ic_stride,
                             const double* gtab, int B, int nx, double length, void* workspace, void* stream) {
    if (!n || !E || B < 1 || nx < 1 || n_ic_stride < nx || e_ic_stride < nx || !(length > 0.0))
        return set_error(FLUXGNN_EINVAL, "poisson_spectral: bad argument (B=%d nx=%d)", B, nx);
    return launch_poisson(n, n_ic_stride, E, e_ic_stride, gtab, B, nx, length, workspace, (cudaStream_t)stream);
}

}  // extern "C"

static bool tc_precision_known(int precision) {
    return precision == FLUXGNN_TC_TF32X3 || precision == FLUXGNN_TC_TF32 || precision == FLUXGNN_TC_FP16X3 ||
           precision == FLUXGNN_TC_FP16 || precision == FLUXGNN_TC_BF16;
}

// Tensor-path fields of the launch arguments; must run BEFORE plan_tiles (the tile height depends on it).
static void tc_configure(HybridArgs& a, int precision) {
    a.tc_parts = (precision == FLUXGNN_TC_TF32X3 || precision == FLUXGNN_TC_FP16X3) ? 2 : 1;
    a.tc_format = (precision == FLUXGNN_TC_BF16) ? 1 : 0;
    const bool wide = precision == FLUXGNN_TC_FP16X3 || precision == FLUXGNN_TC_FP16 || precision == FLUXGNN_TC_BF16;
    a.tile_rows = wide ? kTc16TileRows : kTileRows;
}

static int tc_shape_ok(int whole_ic, int nx, int radius) {
    if (radius > 4) return set_error(FLUXGNN_EUNSUP, "tensor path: radius must be <= 4, got %d", radius);
    if (whole_ic && nx != 32 && nx != 64 && nx != 128)
        return set_error(FLUXGNN_EUNSUP, "tensor path: whole-IC tiles need nx in {32, 64, 128}, got %d "
                                         "(use the fp32 entry point)", nx);
    return FLUXGNN_OK;
}

// Latency mode: with fewer tiles than cluster slots the tile kernel would use one SM per tile and leave the rest idle; a
// cluster of 8 CTAs per tile splits every layer's output features instead (bit-identical results).  FLUXGNN_LATENCY=0 / 1
// forces the choice (test hook).  Returns 0 when it launched, 1 when the caller should use the tile kernel, < 0 on error.
static int try_latency_mode(const HybridArgs& a, int precision, cudaStream_t stream) {
    if (precision != 0 || !hybrid_latency_supported(a)) return 1;
    const char* env = getenv("FLUXGNN_LATENCY");
    const bool forced = env != nullptr && env[0] == '1', off = env != nullptr && env[0] == '0';
    const int slots = off ? 0 : hybrid_latency_max_clusters();
    if (slots < 1 || !(forced || a.num_tiles <= 2 * slots)) return 1;
    FLUXGNN_CUDA_OK(launch_hybrid_latency(a, a.num_tiles < slots ? a.num_tiles : slots, stream));
    count_launch();
    return FLUXGNN_OK;
}

static int forward_ring_impl(int precision, const void* packed, int num_layers, const float* state, const float* x,
                             int B, int nx, int radius, int hops, float* flux_edges, float* face_flux,
                             void* stream, float* acts = nullptr) {
    int rc = check_model(packed, num_layers, B, nx, radius);
    if (rc != FLUXGNN_OK) return rc;
    if (!state || !x) return set_error(FLUXGNN_EINVAL, "forward_ring: null state or x");
    if (hops < 1 || hops > kMaxHops || hops > radius)
        return set_error(FLUXGNN_EINVAL, "forward_ring: hops must be in 1..min(radius,%d), got %d", kMaxHops, hops);
    if (!flux_edges && !face_flux) return set_error(FLUXGNN_EINVAL, "forward_ring: no output requested");
    HybridArgs a{};
    a.packed = (const float*)packed;
    a.state_in = state;
    a.x = x;
    a.flux_edges = flux_edges;
    a.face_flux = face_flux;
    a.B = B; a.nx = nx; a.radius = radius; a.L = num_layers; a.hops = hops;
    a.do_update = 0; a.steps = 1; a.record_every = 1;
    a.acts = acts;
    a.acts_stride = (long long)B * nx * kH;
    int fast = 0;
    if (precision != 0) tc_configure(a, precision);
    rc = plan_tiles(a, &fast);
    if (rc != FLUXGNN_OK) return rc;
    if (precision != 0) {
        if (hops != 1) return set_error(FLUXGNN_EUNSUP, "tensor path: forward emits hop 1 only");
        rc = tc_shape_ok(a.whole_ic, nx, radius);
        if (rc != FLUXGNN_OK) return rc;
        fast = -radius;
    }
    int lat = try_latency_mode(a, precision, (cudaStream_t)stream);
    if (lat != 1) return lat;                              // 0: launched, < 0: error
    return launch_tiles(a, fast, (cudaStream_t)stream);
}

extern "C" {

int fluxgnn_forward_ring(const void* packed, int num_layers, const float* state, const float* x,
                         int B, int nx, int radius, int hops, float* flux_edges, float* face_flux,
                         void* stream) {
    return forward_ring_impl(0, packed, num_layers, state, x, B, nx, radius, hops, flux_edges, face_flux, stream);
}

int fluxgnn_forward_ring_tc(const void* packed_tc, int num_layers, int precision, const float* state,
                            const float* x, int B, int nx, int radius, float* flux_edges, float* face_flux,
                            void* stream) {
    if (!tc_precision_known(precision))
        return set_error(FLUXGNN_EINVAL, "forward_ring_tc: precision must be one of FLUXGNN_TC_*, got %d", precision);
    return forward_ring_impl(precision, packed_tc, num_layers, state, x, B, nx, radius, 1, flux_edges, face_flux,
                             stream);
}

// workspace = [state ping-pong buffer][FFT scratch]; both absent for nx <= 128
size_t fluxgnn_hybrid_workspace_bytes(int B, int nx) {
    if (B < 1 || nx < 1 || nx <= kTileRows) return 0;
    return (size_t)B * 3 * nx * sizeof(float) + fluxgnn_poisson_workspace_bytes(B, nx);
}

// workspace = [state ping-pong buffer][FFT scratch][tile-major n, u and halo side arrays of the fused step]
size_t fluxgnn_baseline_workspace_bytes(int B, int nx) {
    if (B < 1 || nx < 1) return 0;
    return (size_t)B * 3 * nx * sizeof(float) + fluxgnn_poisson_workspace_bytes(B, nx) +
           baseline_fused_workspace_floats(B, nx) * sizeof(float);
}

}  // extern "C"

// precision: 0 = fp32 (FP32-pipe kernel), 1 = tf32x3, 2 = tf32 (tensor-core kernel)
static int hybrid_rollout_impl(int precision, const void* packed, int num_layers, const float* state_in,
                               float* state_out, const float* x, const double* gtab, int B, int nx, double length,
                               int radius, float c, float dt, int steps, int record_every, float* traj,
                               void* workspace, void* stream_, float* diag = nullptr, float* acts = nullptr,
                               float* face_flux = nullptr) {
    cudaStream_t stream = (cudaStream_t)stream_;
    int rc = check_model(packed, num_layers, B, nx, radius);
    if (rc != FLUXGNN_OK) return rc;
    if (!state_in || !state_out || !x) return set_error(FLUXGNN_EINVAL, "hybrid_rollout: null pointer");
    if (nx <= kTileRows && !gtab) return set_error(FLUXGNN_EINVAL, "hybrid_rollout: gtab required for nx <= %d", kTileRows);
    if (!(length > 0.0)) return set_error(FLUXGNN_EINVAL, "hybrid_rollout: length must be positive");
    if (state_in == state_out) return set_error(FLUXGNN_EINVAL, "hybrid_rollout: state_in and state_out alias");
    if (steps < 1) return set_error(FLUXGNN_EINVAL, "hybrid_rollout: steps must be >= 1, got %d", steps);
    if (traj && record_every < 1) return set_error(FLUXGNN_EINVAL, "hybrid_rollout: record_every must be >= 1");
    HybridArgs a{};
    a.packed = (const float*)packed;
    a.x = x;
    a.gtab = gtab;
    a.B = B; a.nx = nx; a.radius = radius; a.L = num_layers; a.hops = 1;
    a.do_update = 1;
    a.c = c; a.dt = dt;
    a.record_every = record_every < 1 ? 1 : record_every;
    a.acts = acts;                                        // training step: the kernel also saves its activations
    a.acts_stride = (long long)B * nx * kH;
    a.face_flux = face_flux;
    int fast = 0;
    if (precision != 0) tc_configure(a, precision);
    rc = plan_tiles(a, &fast);
    if (rc != FLUXGNN_OK) return rc;
    if (precision != 0) {
        rc = tc_shape_ok(a.whole_ic, nx, radius);
        if (rc != FLUXGNN_OK) return rc;
        fast = -radius;                                   // launch_tiles(): negative = tensor kernel
    }

    if (a.whole_ic) {
        // whole ICs per tile: the complete rollout is ONE persistent launch, state in shared memory
        a.state_in = state_in;
        a.state_out = state_out;
        a.traj = traj;
        a.diag = diag;
        a.steps = steps;
        int lat = try_latency_mode(a, precision, stream);
        if (lat != 1) return lat;                          // 0: launched, < 0: error
        return launch_tiles(a, fast, stream);
    }
    if (diag != nullptr)
        return set_error(FLUXGNN_EUNSUP, "in-kernel diagnostics need whole-IC tiles (nx <= %d), got nx=%d; reduce a "
                                         "recorded trajectory with fluxgnn_rollout_metrics instead", kTileRows, nx);
    // window tiles: per step  tile kernel (n', u')  ->  field-solve kernel (E')
    const size_t state_floats = (size_t)B * 3 * nx;
    const bool need_ws = steps > 1 || fluxgnn_poisson_workspace_bytes(B, nx) > 0;
    if (need_ws && workspace == nullptr)
        return set_error(FLUXGNN_EINVAL, "hybrid_rollout: fluxgnn_hybrid_workspace_bytes() of workspace required");
    void* fft_ws = workspace ? (void*)((float*)workspace + state_floats) : nullptr;
    const float* src = state_in;
    a.steps = 1;
    a.traj = nullptr;
    for (int t = 0; t < steps; ++t) {
        float* dst = ((steps - 1 - t) % 2 == 0) ? state_out : (float*)workspace;
        a.state_in = src;
        a.state_out = dst;
        rc = launch_tiles(a, fast, stream);
        if (rc != FLUXGNN_OK) return rc;
        rc = launch_poisson(dst, 3LL * nx, dst + 2 * (size_t)nx, 3LL * nx, gtab, B, nx, length, fft_ws, stream);
        if (rc != FLUXGNN_OK) return rc;
        if (traj && (t + 1) % record_every == 0) {
            FLUXGNN_CUDA_OK(cudaMemcpyAsync(traj + (size_t)((t + 1) / record_every - 1) * state_floats, dst,
                                            state_floats * sizeof(float), cudaMemcpyDeviceToDevice, stream));
        }
        src = dst;
    }
    return FLUXGNN_OK;
}

extern "C" {

int fluxgnn_hybrid_rollout(const void* packed, int num_layers, const float* state_in, float* state_out,
                           const float* x, const double* gtab, int B, int nx, double length, int radius,
                           float c, float dt, int steps, int record_every, float* traj, void* workspace,
                           void* stream) {
    return hybrid_rollout_impl(0, packed, num_layers, state_in, state_out, x, gtab, B, nx, length, radius, c, dt,
                               steps, record_every, traj, workspace, stream);
}

int fluxgnn_hybrid_rollout_diag(const void* packed, int num_layers, int precision, const float* state_in,
                                float* state_out, const float* x, const double* gtab, int B, int nx, double length,
                                int radius, float c, float dt, int steps, float* diag, void* stream) {
    if (precision != 0 && !tc_precision_known(precision))
        return set_error(FLUXGNN_EINVAL, "hybrid_rollout_diag: precision must be 0 (fp32 kernel) or one of FLUXGNN_TC_*");
    if (!diag) return set_error(FLUXGNN_EINVAL, "hybrid_rollout_diag: null diagnostics buffer");
    return hybrid_rollout_impl(precision, packed, num_layers, state_in, state_out, x, gtab, B, nx, length, radius, c, dt,
                               steps, 1, nullptr, nullptr, stream, diag);
}

int fluxgnn_hybrid_rollout_tc(const void* packed_tc, int num_layers, int precision, const float* state_in,
                              float* state_out, const float* x, const double* gtab, int B, int nx, double length,
                              int radius, float c, float dt, int steps, int record_every, float* traj,
                              void* workspace, void* stream) {
    if (!tc_precision_known(precision))
        return set_error(FLUXGNN_EINVAL, "hybrid_rollout_tc: precision must be one of FLUXGNN_TC_*, got %d", precision);
    return hybrid_rollout_impl(precision, packed_tc, num_layers, state_in, state_out, x, gtab, B, nx, length, radius,
                               c, dt, steps, record_every, traj, workspace, stream);
}

size_t fluxgnn_train_acts_bytes(int num_layers, int B, int nx) {
    if (num_layers < 1 || num_layers > kMaxL || B < 1 || nx < 1) return 0;
    return (size_t)(num_layers + 3) * B * nx * kH * sizeof(float);
}

int fluxgnn_forward_ring_train(const void* packed, int num_layers, const float* state, const float* x, int B, int nx,
                               int radius, int hops, float* flux_edges, float* acts, void* stream) {
    if (!acts || !flux_edges) return set_error(FLUXGNN_EINVAL, "forward_ring_train: flux_edges and acts are required");
    return forward_ring_impl(0, packed, num_layers, state, x, B, nx, radius, hops, flux_edges, nullptr, stream, acts);
}

size_t fluxgnn_backward_workspace_bytes(int B, int nx) {
    if (B < 1 || nx < 1) return 0;
    return (size_t)3 * B * nx * kH * sizeof(float);
}

int fluxgnn_backward_ring(const float* w_in, const float* w_upd, const float* w_e1, const float* w_e2, int num_layers,
                          const float* state, const float* x, const float* acts, const float* dflux, int B, int nx,
                          int radius, int hops, float* g_w_in, float* g_b_in, float* g_w_upd, float* g_b_upd,
                          float* g_w_e1, float* g_b_e1, float* g_w_e2, float* g_b_e2, float* dstate, void* workspace,
                          void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    const int L = num_layers;
    if (L < 1 || L > kMaxL) return set_error(FLUXGNN_EUNSUP, "num_layers must be in 1..%d, got %d", kMaxL, L);
    if (B < 1 || nx < 1 || radius < 1 || hops < 1 || hops > kMaxHops || hops > radius)
        return set_error(FLUXGNN_EINVAL, "backward_ring: bad shape (B=%d nx=%d radius=%d hops=%d)", B, nx, radius, hops);
    if (!w_in || !w_upd || !w_e1 || !w_e2 || !state || !x || !acts || !dflux || !g_w_in || !g_b_in || !g_w_upd ||
        !g_b_upd || !g_w_e1 || !g_b_e1 || !g_w_e2 || !g_b_e2 || !workspace)
        return set_error(FLUXGNN_EINVAL, "backward_ring: null pointer");
    const long long rows = (long long)B * nx;
    const size_t stride = (size_t)rows * kH;
    float* G0 = (float*)workspace;
    float* G1 = G0 + stride;
    float* G2 = G1 + stride;
    const float* Pa = acts + (size_t)(L + 1) * stride;
    const float* Qa = acts + (size_t)(L + 2) * stride;
    const dim3 blk(32, 8);
    int sms = 0;
    int rc_sm = sm_count(&sms);
    if (rc_sm != FLUXGNN_OK) return rc_sm;
    const long long row_groups = (rows + 7) / 8;                  // elementwise kernels: grid-stride, 8 rows per block pass
    const unsigned g8 = (unsigned)(row_groups < 16LL * sms ? row_groups : 16LL * sms);
    const unsigned g128 = (unsigned)((rows + 127) / 128);          // row-product GEMMs: one block per 128 rows
    const long long slabs = (rows + 255) / 256;                   // weight-gradient GEMMs: at most 2 blocks per SM
    const unsigned g256 = (unsigned)(slabs < 2LL * sms ? slabs : 2LL * sms);
    // edge readout
    bwd_edge_kernel<<<g8, blk, 0, stream>>>(Pa, Qa, w_e2, dflux, G0, G1, g_w_e2, g_b_e1, g_b_e2, rows, nx, hops);
    const float* HL = acts + (size_t)L * stride;
    bwd_gemm_tn_kernel<<<g256, 256, 0, stream>>>(G0, HL, g_w_e1, 2 * kH, rows);          // d W1[:, :H] = dP^T h
    bwd_gemm_tn_kernel<<<g256, 256, 0, stream>>>(G1, HL, g_w_e1 + kH, 2 * kH, rows);     // d W1[:, H:] = dQ^T h
    bwd_gemm_nn_kernel<<<g128, 256, 0, stream>>>(G0, w_e1, G1, w_e1 + kH, 2 * kH, G2, rows);
    count_launch(4);
    // message-passing layers, last to first
    for (int l = L - 1; l >= 0; --l) {
        const float* W = w_upd + (size_t)l * kH * 2 * kH;
        float* gW = g_w_upd + (size_t)l * kH * 2 * kH;
        const float* Hl = acts + (size_t)l * stride;
        bwd_mask_mean_kernel<<<g8, blk, 0, stream>>>(G2, acts + (size_t)(l + 1) * stride, G0, G1, g_b_upd + l * kH, rows,
                                                     nx, radius);
        bwd_gemm_tn_kernel<<<g256, 256, 0, stream>>>(G0, Hl, gW, 2 * kH, rows);           // d W[:, :H] = dpre^T h
        bwd_gemm_tn_kernel<<<g256, 256, 0, stream>>>(G1, Hl, gW + kH, 2 * kH, rows);      // d W[:, H:] = mean(dpre)^T h
        bwd_gemm_nn_kernel<<<g128, 256, 0, stream>>>(G0, W, G1, W + kH, 2 * kH, G2, rows);
        count_launch(4);
    }
    bwd_input_kernel<<<g8, blk, 0, stream>>>(G2, acts, w_in, state, x, dstate, g_w_in, g_b_in, rows, nx);
    count_launch();
    FLUXGNN_CUDA_OK(cudaGetLastError());
    return FLUXGNN_OK;
}

size_t fluxgnn_step_backward_workspace_bytes(int B, int nx) {
    if (B < 1 || nx < 1) return 0;
    return fluxgnn_backward_workspace_bytes(B, nx) + (size_t)2 * B * nx * sizeof(float);
}

int fluxgnn_hybrid_step_train(const void* packed, int num_layers, const float* state_in, float* state_out,
                              const float* x, const double* gtab, int B, int nx, double length, int radius, float c,
                              float dt, float* face_flux, float* acts, void* workspace, void* stream) {
    if (!acts) return set_error(FLUXGNN_EINVAL, "hybrid_step_train: acts is required");
    return hybrid_rollout_impl(0, packed, num_layers, state_in, state_out, x, gtab, B, nx, length, radius, c, dt, 1, 1,
                               nullptr, workspace, stream, nullptr, acts, face_flux);
}

int fluxgnn_hybrid_step_backward(const float* w_in, const float* w_upd, const float* w_e1, const float* w_e2,
                                 int num_layers, const float* state, const float* x, const float* acts,
                                 const float* g_state_out, const float* g_face, int B, int nx, int radius, float c, float dt,
                                 float* g_w_in, float* g_b_in, float* g_w_upd, float* g_b_upd, float* g_w_e1, float* g_b_e1,
                                 float* g_w_e2, float* g_b_e2, float* dstate, void* workspace, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!g_state_out || !dstate || !workspace || !state)
        return set_error(FLUXGNN_EINVAL, "hybrid_step_backward: null pointer");
    if (B < 1 || nx < 1) return set_error(FLUXGNN_EINVAL, "hybrid_step_backward: bad shape (B=%d nx=%d)", B, nx);
    int sms = 0;
    int rc = sm_count(&sms);
    if (rc != FLUXGNN_OK) return rc;
    const long long cells = (long long)B * nx;
    const long long blocks = (cells + 255) / 256;
    const unsigned grid = (unsigned)(blocks < 16LL * sms ? blocks : 16LL * sms);
    float* dflux = (float*)((char*)workspace + fluxgnn_backward_workspace_bytes(B, nx));
    step_bwd_flux_kernel<<<grid, 256, 0, stream>>>(g_state_out, g_face, dflux, cells, nx, c);
    FLUXGNN_CUDA_OK(cudaGetLastError());
    count_launch();
    rc = fluxgnn_backward_ring(w_in, w_upd, w_e1, w_e2, num_layers, state, x, acts, dflux, B, nx, radius, 1, g_w_in, g_b_in,
                               g_w_upd, g_b_upd, g_w_e1, g_b_e1, g_w_e2, g_b_e2, dstate, workspace, stream_);
    if (rc != FLUXGNN_OK) return rc;
    step_bwd_direct_kernel<<<grid, 256, 0, stream>>>(g_state_out, state, dstate, cells, nx, c, dt);
    FLUXGNN_CUDA_OK(cudaGetLastError());
    count_launch();
    return FLUXGNN_OK;
}

int fluxgnn_rollout_metrics(const float* pred, const float* truth, long long num_states, int nx, float* out,
                            void* stream) {
    if (!pred || !out || num_states < 1 || nx < 1 || num_states > 0x7fffffffLL)
        return set_error(FLUXGNN_EINVAL, "rollout_metrics: bad argument (states=%lld nx=%d)", num_states, nx);
    rollout_metrics_kernel<<<(unsigned)num_states, 256, 0, (cudaStream_t)stream>>>(pred, truth, nx, out);
    FLUXGNN_CUDA_OK(cudaGetLastError());
    count_launch();
    return FLUXGNN_OK;
}

int fluxgnn_hybrid_slab_step(const void* packed, int num_layers, int precision, const float* state_ext,
                             const float* x_ext, float* state_out, int B, int owned, int halo, int radius,
                             float c, float dt, void* stream) {
    return fluxgnn_hybrid_slab_step_ld(packed, num_layers, precision, state_ext, x_ext, state_out, owned, 0, B, owned,
                                       halo, radius, c, dt, stream);
}

static int hybrid_slab_step_impl(const void* packed, int num_layers, int precision, const float* state_ext,
                                 const float* x_ext, float* state_out, int out_ld, int out_off, int B, int owned,
                                 int halo, int radius, float c, float dt, float* left_out, float* right_out, void* stream) {
    int rc = check_model(packed, num_layers, B, owned, radius);
    if (rc != FLUXGNN_OK) return rc;
    if (!state_ext || !x_ext || !state_out) return set_error(FLUXGNN_EINVAL, "hybrid_slab_step: null pointer");
    if (precision != 0 && !tc_precision_known(precision))
        return set_error(FLUXGNN_EINVAL, "hybrid_slab_step: bad precision %d", precision);
    if (halo != num_layers * radius + 1)
        return set_error(FLUXGNN_EINVAL, "hybrid_slab_step: halo must be num_layers*radius+1 = %d, got %d",
                         num_layers * radius + 1, halo);
    if (owned < halo) return set_error(FLUXGNN_EINVAL, "hybrid_slab_step: a slab must own at least `halo` cells");
    HybridArgs a{};
    a.packed = (const float*)packed;
    a.state_in = state_ext;
    a.state_out = state_out;
    a.x = x_ext;
    a.B = B; a.nx = owned; a.radius = radius; a.L = num_layers; a.hops = 1;
    a.do_update = 1; a.steps = 1; a.record_every = 1;
    a.c = c; a.dt = dt;
    a.tile_rows = kTileRows;
    if (precision != 0) tc_configure(a, precision);
    int rows = a.tile_rows;
    if (a.tile_rows == kTc16TileRows) {                                  // as in plan_tiles
        const char* nosplit = getenv("FLUXGNN_TC16_NO_SPLIT");
        a.tc_group_rows = (!(nosplit != nullptr && nosplit[0] == '1') && halo <= 24) ? 128 : kTc16TileRows;
        rows = a.tc_group_rows;
    }
    // always window tiles: the receptive field is served by the ghost cells
    a.whole_ic = 0;
    a.halo = halo;
    a.valid = rows - 2 * halo;
    if (a.valid < 8)
        return set_error(FLUXGNN_EUNSUP, "receptive field of %d cells does not fit a %d-cell tile", halo, rows);
    a.tiles_per_ic = (owned + a.valid - 1) / a.valid;
    const long long tiles = (long long)B * a.tiles_per_ic;
    if (tiles > 0x7fffffffLL) return set_error(FLUXGNN_EINVAL, "too many tiles");
    a.num_tiles = (int)tiles;
    a.slab = 1;
    a.ld_in = owned + 2 * halo;
    if (precision == 0 && radius <= 4) choose_cluster(a, owned);
    if (out_off < 0 || out_ld < out_off + owned)
        return set_error(FLUXGNN_EINVAL, "hybrid_slab_step: output row of %d floats cannot hold %d cells at offset %d",
                         out_ld, owned, out_off);
    a.ld_out = out_ld;
    a.out_off = out_off;
    if ((left_out != nullptr || right_out != nullptr) && out_off < halo)
        return set_error(FLUXGNN_EINVAL, "hybrid_slab_step_peer: the output rows need %d ghost cells in front of cell 0", halo);
    a.peer_left = left_out;
    a.peer_right = right_out;
    int fast = (radius <= 4) ? radius : 0;
    if (precision != 0) {
        rc = tc_shape_ok(0, owned, radius);
        if (rc != FLUXGNN_OK) return rc;
        fast = -radius;
    }
    return launch_tiles(a, fast, (cudaStream_t)stream);
}

int fluxgnn_hybrid_slab_step_ld(const void* packed, int num_layers, int precision, const float* state_ext,
                                const float* x_ext, float* state_out, int out_ld, int out_off, int B, int owned,
                                int halo, int radius, float c, float dt, void* stream) {
    return hybrid_slab_step_impl(packed, num_layers, precision, state_ext, x_ext, state_out, out_ld, out_off, B, owned, halo,
                                 radius, c, dt, nullptr, nullptr, stream);
}

int fluxgnn_hybrid_slab_step_peer(const void* packed, int num_layers, int precision, const float* state_ext,
                                  const float* x_ext, float* state_out, int out_ld, int out_off, int B, int owned,
                                  int halo, int radius, float c, float dt, float* left_out, float* right_out, void* stream) {
    if (!left_out || !right_out) return set_error(FLUXGNN_EINVAL, "hybrid_slab_step_peer: null neighbour pointer");
    return hybrid_slab_step_impl(packed, num_layers, precision, state_ext, x_ext, state_out, out_ld, out_off, B, owned, halo,
                                 radius, c, dt, left_out, right_out, stream);
}

int fluxgnn_poisson_dist_pack(const float* n, long long ic_stride, int B, int S, void* z, void* stream) {
    if (!n || !z || B < 1 || S < 1 || ic_stride < S) return set_error(FLUXGNN_EINVAL, "poisson_dist_pack: bad argument");
    return launch_poisson_dist_pack(n, ic_stride, B, S, (float2*)z, 0, nullptr, (cudaStream_t)stream);
}

int fluxgnn_poisson_dist_unpack(const void* e, float* E, long long ic_stride, int B, int S, void* stream) {
    if (!e || !E || B < 1 || S < 1 || ic_stride < S) return set_error(FLUXGNN_EINVAL, "poisson_dist_unpack: bad argument");
    return launch_poisson_dist_pack(nullptr, ic_stride, B, S, (float2*)const_cast<void*>(e), 1, E, (cudaStream_t)stream);
}

int fluxgnn_poisson_dist_rank_dft(const void* in, void* out, int G, long long chunk, long long flat0, int S, int inverse,
                                  void* stream) {
    if (!in || !out || in == out || chunk < 1 || flat0 < 0)
        return set_error(FLUXGNN_EINVAL, "poisson_dist_rank_dft: bad argument (chunk=%lld)", chunk);
    return launch_poisson_rank_dft((const float2*)in, (float2*)out, G, chunk, flat0, S, inverse ? 1 : 0, (cudaStream_t)stream);
}

int fluxgnn_poisson_dist_local(void* y, void* scratch, int P, int S, int G, int rank, double length, void* stream) {
    if (!y || !(length > 0.0)) return set_error(FLUXGNN_EINVAL, "poisson_dist_local: bad argument");
    return launch_poisson_dist_local((float2*)y, (float2*)scratch, P, S, G, rank, length, (cudaStream_t)stream);
}

static int baseline_slab_step_impl(const float* state_ext, float* state_out, int out_ld, int out_off, float* flux_n,
                                   int B, int owned, int halo, float c, float dt, float nu, float dx2, float* left_out,
                                   float* right_out, void* stream) {
    if (!state_ext || !state_out || B < 1 || owned < 1 || halo < 1)
        return set_error(FLUXGNN_EINVAL, "baseline_slab_step: bad argument (B=%d owned=%d halo=%d)", B, owned, halo);
    if (out_off < 0 || out_ld < out_off + owned)
        return set_error(FLUXGNN_EINVAL, "baseline_slab_step: output row of %d floats cannot hold %d cells at offset %d",
                         out_ld, owned, out_off);
    int sms = 0;
    int rc = sm_count(&sms);
    if (rc != FLUXGNN_OK) return rc;
    const bool vec = ((owned | halo | out_ld | out_off) & 3) == 0 && (((uintptr_t)state_ext | (uintptr_t)state_out) & 15) == 0 &&
                     (flux_n == nullptr || ((uintptr_t)flux_n & 15) == 0);
    const long long cells = (long long)B * owned;
    long long blocks = ((vec ? cells / 4 : cells) + 255) / 256;
    if (blocks > (long long)sms * 64) blocks = (long long)sms * 64;
    if ((left_out != nullptr || right_out != nullptr) && (out_off < halo || owned < halo))
        return set_error(FLUXGNN_EINVAL, "baseline_slab_step_peer: the output rows need %d ghost cells in front of cell 0", halo);
    baseline_fv_slab_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(state_ext, state_out, flux_n, B, owned, halo,
                                                                              out_ld, out_off, vec ? 1 : 0, c, dt, nu, dx2,
                                                                              left_out, right_out);
    FLUXGNN_CUDA_OK(cudaGetLastError());
    count_launch();
    return FLUXGNN_OK;
}

int fluxgnn_baseline_slab_step(const float* state_ext, float* state_out, int out_ld, int out_off, float* flux_n,
                               int B, int owned, int halo, float c, float dt, float nu, float dx2, void* stream) {
    return baseline_slab_step_impl(state_ext, state_out, out_ld, out_off, flux_n, B, owned, halo, c, dt, nu, dx2, nullptr,
                                   nullptr, stream);
}

int fluxgnn_baseline_slab_step_peer(const float* state_ext, float* state_out, int out_ld, int out_off, float* flux_n,
                                    int B, int owned, int halo, float c, float dt, float nu, float dx2, float* left_out,
                                    float* right_out, void* stream) {
    if (!left_out || !right_out) return set_error(FLUXGNN_EINVAL, "baseline_slab_step_peer: null neighbour pointer");
    return baseline_slab_step_impl(state_ext, state_out, out_ld, out_off, flux_n, B, owned, halo, c, dt, nu, dx2, left_out,
                                   right_out, stream);
}

int fluxgnn_baseline_rollout(const float* state_in, float* state_out, const double* gtab, int B, int nx,
                             double length, float c, float dt, float nu, float dx2, int steps, int record_every,
                             float* traj, float* flux_n, void* workspace, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!state_in || !state_out || B < 1 || nx < 1 || !(length > 0.0))
        return set_error(FLUXGNN_EINVAL, "baseline_rollout: bad argument (B=%d nx=%d)", B, nx);
    if (state_in == state_out) return set_error(FLUXGNN_EINVAL, "baseline_rollout: state_in and state_out alias");
    if (steps < 1) return set_error(FLUXGNN_EINVAL, "baseline_rollout: steps must be >= 1, got %d", steps);
    if ((steps > 1 || fluxgnn_poisson_workspace_bytes(B, nx) > 0) && workspace == nullptr)
        return set_error(FLUXGNN_EINVAL, "baseline_rollout: fluxgnn_baseline_workspace_bytes() of workspace required");
    if (traj && record_every < 1) return set_error(FLUXGNN_EINVAL, "baseline_rollout: record_every must be >= 1");
    int sms = 0;
    int rc = sm_count(&sms);
    if (rc != FLUXGNN_OK) return rc;
    // Short grids with the direct field solve (the reference's default 64 cells): the whole rollout of an IC in one
    // persistent CTA, one launch instead of two per step; bit-identical (FLUXGNN_BASELINE_PERSIST=0: test hook).
    {
        const char* env = getenv("FLUXGNN_BASELINE_PERSIST");
        if (!poisson_fft_supported(nx) && nx <= kBaselineSmallMaxNx && gtab != nullptr && !(env != nullptr && env[0] == '0')) {
            const size_t smem = (size_t)nx * (sizeof(double) + 6 * sizeof(float));
            FLUXGNN_CUDA_OK(cudaFuncSetAttribute(baseline_small_rollout_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            const int threads = nx >= 256 ? 256 : ((nx + 31) / 32) * 32;
            baseline_small_rollout_kernel<<<(unsigned)B, threads, smem, stream>>>(state_in, state_out, traj, flux_n, gtab, B, nx, steps,
                                                                                 record_every < 1 ? 1 : record_every, c, dt, nu, dx2);
            FLUXGNN_CUDA_OK(cudaGetLastError());
            count_launch();
            return FLUXGNN_OK;
        }
    }
    const size_t state_floats = (size_t)B * 3 * nx;
    void* fft_ws = workspace ? (void*)((float*)workspace + state_floats) : nullptr;
    const long long cells = (long long)B * nx;
    long long blocks = (((nx & 3) == 0 ? cells / 4 : cells) + 255) / 256;
    if (blocks > (long long)sms * 64) blocks = (long long)sms * 64;
    // Opt-in (FLUXGNN_BASELINE_FUSE=1): after the first step the inverse column stages, the finite-volume update and
    // the forward column stages of consecutive steps run as one kernel (fft_poisson.cu, "fused classical step"): two
    // launches and 32 bytes per cell-update instead of four and 44, bit-identical results.  Measured on B200 it is
    // SLOWER than the four-kernel sequence (0.190 against 0.172 ms per step at 2^24 cells, profiles/r2_c5_fused.md):
    // one 512-thread CTA per SM serialises its phases, while the stand-alone finite-volume kernel streams at 93 % of
    // the HBM peak; so the default stays unfused.
    const char* fuse = getenv("FLUXGNN_BASELINE_FUSE");
    if (steps >= 3 && !traj && !flux_n && baseline_fused_supported(nx, dx2) && B <= 65535 &&
        fuse != nullptr && fuse[0] == '1') {
        float* tmp = (float*)workspace;                                    // natural-layout state after step 1
        float2* Y = (float2*)fft_ws;
        float* fused_ws = (float*)fft_ws + fluxgnn_poisson_workspace_bytes(B, nx) / sizeof(float);
        baseline_fv_kernel<<<(unsigned)blocks, 256, 0, stream>>>(state_in, tmp, nullptr, B, nx, c, dt, nu, dx2);
        FLUXGNN_CUDA_OK(cudaGetLastError());
        count_launch();
        rc = launch_baseline_to_tiles(tmp, fused_ws, 0, B, nx, stream);
        if (rc != FLUXGNN_OK) return rc;
        rc = launch_poisson_fft_cols(tmp, 3LL * nx, Y, nullptr, 0, B, nx, 0, stream);
        if (rc != FLUXGNN_OK) return rc;
        rc = launch_poisson_fft_rows(Y, B, nx, length, stream);
        if (rc != FLUXGNN_OK) return rc;
        for (int t = 1; t < steps; ++t) {
            rc = launch_baseline_fused_cols(Y, fused_ws, (t - 1) & 1, t == steps - 1 ? state_out : nullptr, B, nx, c, dt, nu,
                                            dx2, stream);
            if (rc != FLUXGNN_OK) return rc;
            rc = launch_poisson_fft_rows(Y, B, nx, length, stream);
            if (rc != FLUXGNN_OK) return rc;
        }
        return launch_poisson_fft_cols(nullptr, 0, Y, state_out + 2 * (size_t)nx, 3LL * nx, B, nx, 1, stream);
    }
    const float* src = state_in;
    for (int t = 0; t < steps; ++t) {
        float* dst = ((steps - 1 - t) % 2 == 0) ? state_out : (float*)workspace;
        baseline_fv_kernel<<<(unsigned)blocks, 256, 0, stream>>>(src, dst, flux_n ? flux_n + (size_t)t * cells : nullptr,
                                                                 B, nx, c, dt, nu, dx2);
        FLUXGNN_CUDA_OK(cudaGetLastError());
        count_launch();
        rc = launch_poisson(dst, 3LL * nx, dst + 2 * (size_t)nx, 3LL * nx, gtab, B, nx, length, fft_ws, stream);
        if (rc != FLUXGNN_OK) return rc;
        if (traj && (t + 1) % record_every == 0) {
            FLUXGNN_CUDA_OK(cudaMemcpyAsync(traj + (size_t)((t + 1) / record_every - 1) * state_floats, dst,
                                            state_floats * sizeof(float), cudaMemcpyDeviceToDevice, stream));
        }
        src = dst;
    }
    return FLUXGNN_OK;
}

int fluxgnn_baseline_scan_supported(int B, int nx) { return baseline_scan_supported(B, nx) ? 1 : 0; }

size_t fluxgnn_baseline_scan_workspace_bytes(int B, int nx) {
    int sms = 0;
    if (!baseline_scan_supported(B, nx) || sm_count(&sms) != FLUXGNN_OK) return 0;
    return baseline_scan_workspace_bytes(B, nx, sms);
}

int fluxgnn_baseline_rollout_scan(const float* state_in, float* state_out, int B, int nx, double length, float c, float dt,
                                  float nu, float dx2, int steps, int record_every, float* traj, float* flux_n,
                                  double cert_tol, void* workspace, int* first_uncertified, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!state_in || !state_out || !workspace || !first_uncertified || !(length > 0.0) || !(cert_tol > 0.0))
        return set_error(FLUXGNN_EINVAL, "baseline_rollout_scan: bad argument");
    if (state_in == state_out) return set_error(FLUXGNN_EINVAL, "baseline_rollout_scan: state_in and state_out alias");
    if (steps < 1) return set_error(FLUXGNN_EINVAL, "baseline_rollout_scan: steps must be >= 1, got %d", steps);
    if (traj && record_every < 1) return set_error(FLUXGNN_EINVAL, "baseline_rollout_scan: record_every must be >= 1");
    if (!baseline_scan_supported(B, nx))
        return set_error(FLUXGNN_EUNSUP, "baseline_rollout_scan: needs nx >= 4096 and nx %% 8 == 0 (B=%d nx=%d)", B, nx);
    int sms = 0;
    int rc = sm_count(&sms);
    if (rc != FLUXGNN_OK) return rc;
    int launches = 0;
    FLUXGNN_CUDA_OK(launch_baseline_rollout_scan(state_in, state_out, B, nx, length, c, dt, nu, dx2, steps, record_every, traj,
                                                 flux_n, cert_tol, workspace, first_uncertified, sms, stream, &launches));
    count_launch(launches);
    return FLUXGNN_OK;
}

int fluxgnn_scan_slab_supported(int B, int S) { return scan_slab_supported(B, S) ? 1 : 0; }

size_t fluxgnn_scan_slab_workspace_bytes(int B, int S) {
    int sms = 0;
    if (!scan_slab_supported(B, S) || sm_count(&sms) != FLUXGNN_OK) return 0;
    return scan_slab_workspace_bytes(B, S, sms);
}

int fluxgnn_scan_slab_sums(const float* n, long long n_ld, int B, int S, long long j_base, void* workspace, void* msg,
                           void* stream) {
    if (!n || !workspace || !msg || n_ld < S) return set_error(FLUXGNN_EINVAL, "scan_slab_sums: bad argument");
    if (!scan_slab_supported(B, S)) return set_error(FLUXGNN_EUNSUP, "scan_slab_sums: needs S %% 8 == 0, S >= 64 (B=%d S=%d)", B, S);
    int sms = 0;
    int rc = sm_count(&sms);
    if (rc != FLUXGNN_OK) return rc;
    FLUXGNN_CUDA_OK(launch_scan_slab_sums(n, n_ld, B, S, j_base, workspace, msg, sms, (cudaStream_t)stream));
    count_launch(2);
    return FLUXGNN_OK;
}

int fluxgnn_scan_slab_sums_peer(const float* n, long long n_ld, int B, int S, long long j_base, void* workspace, void* msg,
                                const void* peer_bases_dev, long long offset, int rank, int world, void* stream) {
    if (!n || !workspace || !msg || n_ld < S || !peer_bases_dev || offset < 0 || offset % 16 || world < 1 || rank < 0 ||
        rank >= world)
        return set_error(FLUXGNN_EINVAL, "scan_slab_sums_peer: bad argument");
    if (!scan_slab_supported(B, S)) return set_error(FLUXGNN_EUNSUP, "scan_slab_sums: needs S %% 8 == 0, S >= 64 (B=%d S=%d)", B, S);
    int sms = 0;
    int rc = sm_count(&sms);
    if (rc != FLUXGNN_OK) return rc;
    FLUXGNN_CUDA_OK(launch_scan_slab_sums(n, n_ld, B, S, j_base, workspace, msg, sms, (cudaStream_t)stream,
                                          (void* const*)peer_bases_dev, offset, rank, world));
    count_launch(2);
    return FLUXGNN_OK;
}

int fluxgnn_scan_slab_field(const float* n, long long n_ld, float* E, long long e_ld, int B, int S, int rank, int ranks,
                            double length, const void* msg_all, void* workspace, double cert_tol, int step,
                            int* first_uncertified, void* stream) {
    if (!n || !E || !msg_all || !workspace || !first_uncertified || n_ld < S || e_ld < S || ranks < 1 || rank < 0 ||
        rank >= ranks || !(length > 0.0) || !(cert_tol > 0.0) || step < 0)
        return set_error(FLUXGNN_EINVAL, "scan_slab_field: bad argument");
    if (!scan_slab_supported(B, S)) return set_error(FLUXGNN_EUNSUP, "scan_slab_field: needs S %% 8 == 0, S >= 64 (B=%d S=%d)", B, S);
    int sms = 0;
    int rc = sm_count(&sms);
    if (rc != FLUXGNN_OK) return rc;
    FLUXGNN_CUDA_OK(launch_scan_slab_field(n, n_ld, E, e_ld, B, S, rank, ranks, length, msg_all, workspace, cert_tol, step,
                                           first_uncertified, sms, (cudaStream_t)stream));
    count_launch();
    return FLUXGNN_OK;
}

int fluxgnn_scan_slab_field_peer(const float* n, long long n_ld, float* E, long long e_ld, int B, int S, int rank, int ranks,
                                 double length, const void* msg_all, void* workspace, double cert_tol, int step,
                                 int* first_uncertified, float* E_left, float* E_right, int halo, void* stream) {
    if (!n || !E || !msg_all || !workspace || !first_uncertified || n_ld < S || e_ld < S || ranks < 1 || rank < 0 ||
        rank >= ranks || !(length > 0.0) || !(cert_tol > 0.0) || step < 0 || !E_left || !E_right || halo < 1 || halo > S)
        return set_error(FLUXGNN_EINVAL, "scan_slab_field_peer: bad argument");
    if (!scan_slab_supported(B, S)) return set_error(FLUXGNN_EUNSUP, "scan_slab_field: needs S %% 8 == 0, S >= 64 (B=%d S=%d)", B, S);
    int sms = 0;
    int rc = sm_count(&sms);
    if (rc != FLUXGNN_OK) return rc;
    FLUXGNN_CUDA_OK(launch_scan_slab_field(n, n_ld, E, e_ld, B, S, rank, ranks, length, msg_all, workspace, cert_tol, step,
                                           first_uncertified, sms, (cudaStream_t)stream, E_left, E_right, halo));
    count_launch();
    return FLUXGNN_OK;
}

int fluxgnn_scan_slab_certify(int B, int S, int ranks, double length, const void* msg_all, double cert_tol, int step,
                              int* first_uncertified, void* stream) {
    if (!msg_all || !first_uncertified || B < 1 || S < 1 || ranks < 1 || !(length > 0.0) || !(cert_tol > 0.0))
        return set_error(FLUXGNN_EINVAL, "scan_slab_certify: bad argument");
    FLUXGNN_CUDA_OK(launch_scan_slab_certify(B, S, ranks, length, msg_all, cert_tol, step, first_uncertified,
                                             (cudaStream_t)stream));
    count_launch();
    return FLUXGNN_OK;
}

// Clusters of 8 CTAs the latency mode of fluxgnn_hybrid_rollout runs concurrently on the current device (0: unavailable).
int fluxgnn_peer_halo_push(const float* state_ext, float* left_ext, float* right_ext, int B, int owned, int halo,
                           int ch0, int ch1, void* stream) {
    if (!state_ext || !left_ext || !right_ext || B < 1 || owned < 1 || halo < 1 || halo > owned || ch0 < 0 || ch1 > 3 ||
        ch0 >= ch1)
        return set_error(FLUXGNN_EINVAL, "peer_halo_push: bad argument (B=%d owned=%d halo=%d channels %d..%d)", B, owned,
                         halo, ch0, ch1);
    const long long total = (long long)B * (ch1 - ch0) * 2 * halo;
    const unsigned grid = (unsigned)((total + 255) / 256 < 1024 ? (total + 255) / 256 : 1024);
    peer_halo_push_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(state_ext, left_ext, right_ext, B, owned, halo, ch0, ch1);
    FLUXGNN_CUDA_OK(cudaGetLastError());
    count_launch();
    return FLUXGNN_OK;
}

int fluxgnn_peer_allgather(const void* src, long long bytes, const void* peer_bases_dev, long long offset, int rank,
                           int world, void* stream) {
    if (!src || !peer_bases_dev || bytes < 16 || bytes % 16 || offset < 0 || offset % 16 || rank < 0 || world < 1 ||
        rank >= world || ((size_t)src & 15))
        return set_error(FLUXGNN_EINVAL, "peer_allgather: bad argument (bytes=%lld offset=%lld rank=%d world=%d)", bytes,
                         offset, rank, world);
    peer_allgather_kernel<<<(unsigned)world, 256, 0, (cudaStream_t)stream>>>((const uint4*)src, bytes,
                                                                              (void* const*)peer_bases_dev, offset, rank);
    FLUXGNN_CUDA_OK(cudaGetLastError());
    count_launch();
    return FLUXGNN_OK;
}

int fluxgnn_latency_cluster_slots(void) { return hybrid_latency_max_clusters(); }

}  // extern "C"
