// Launch interface of the fused hybrid tile kernel (hybrid_kernel.cu).
#pragma once

#include <cuda_runtime.h>
#include <stddef.h>

namespace fluxgnn {

constexpr int kTileRows = 128;     // cells per tile (= GEMM M)

struct HybridArgs {
    const float* packed;       // fluxgnn_pack_weights() output
    const float* state_in;     // [B][3][nx]
    float* state_out;          // [B][3][nx]            (update mode)
    const float* x;            // [nx] cell centres
    const double* gtab;        // [nx] field-solve kernel (whole-IC update mode)
    float* flux_edges;         // nullable [B][2*hops*nx]
    float* face_flux;          // nullable [B][nx]
    float* traj;               // nullable [steps/record_every][B][3][nx]
    float* diag;               // nullable [steps][B][4], whole-IC tiles: energy 0.5 mean(u^2+E^2), charge mean(n),
                               //   number of non-finite values, 0 -- after EVERY step, reduced inside the kernel
    int B, nx, radius, L, hops;
    int whole_ic;              // 1: tile = floor(128/nx) complete ICs; 0: window of one IC + halo
    int ics_per_tile;          // whole-IC tiles
    int tiles_per_ic, valid, halo;   // window tiles
    int num_tiles;
    int do_update;             // 0: forward only; 1: finite-volume update (+ field solve when whole_ic)
    int steps, record_every;   // steps > 1 only for whole-IC update mode
    float c, dt;               // float32(dt/dx), float32(dt)
    int tc_parts;              // tensor path only: 2 = x3 split (hi/lo parts, three products), 1 = one product
    int tc_format;             // 16-bit tensor path: 0 = fp16, 1 = bf16 operands
    int tile_rows;             // cells per CTA tile: 128 (kTileRows) or 256 (16-bit tensor path)
    int tc_group_rows;         // 16-bit tensor path: rows of one independent group = height of a LOGICAL tile
                               //   (128: two groups per CTA tile, 256: one); num_tiles, valid, tiles_per_ic and
                               //   ics_per_tile count logical tiles.  0 elsewhere (logical tile = CTA tile).
    float* acts;               // nullable (FP32-pipe kernel, training forward): saved activations, row-major
    long long acts_stride;     //   [L+3][B*nx][128]: h^0..h^L, then P + b1 and Q of the edge readout
    int split;                 // FP32-pipe kernel: two skewed 64-row groups per tile (whole-IC tiles, nx | 64)
    int slab;                  // 1: one slab of a domain-decomposed grid.  state_in and x are
    int ld_in;                 //    [..][ld_in] = nx owned cells + `halo` ghost cells per side, no
                               //    periodic wrap; nx counts the OWNED cells; outputs are [..][nx]
    int cluster;               // FP32-pipe kernel, window / slab tiles: CTAs per thread-block cluster (0/1 = none).  The CTAs of a
                               //    cluster hold consecutive, 4-row overlapping 128-row pieces of ONE window and read each
                               //    other's edge rows of Z through distributed shared memory, so only the cluster's two outer
                               //    ends carry a recomputed halo.  num_tiles / valid / tiles_per_ic count cluster windows.
    int pack_full;             // window tiles of a periodic grid whose last window is short: > 0 = full windows per IC; the
    int pack_per_tile;         //    first B * pack_full tiles are those, every later tile holds pack_per_tile LAST windows
    int pack_seg, pack_rem;    //    (of consecutive ICs) as segments of pack_seg = pack_rem + 2 * halo rows, pack_rem owned each
    float* peer_left;          // slabs over peer memory (nullable): the ring neighbours' NEXT extended states, laid out like
    float* peer_right;         //    state_out (ld_out, out_off).  n', u' of the first / last `halo` owned cells are ALSO stored
                               //    into the left neighbour's right ghost zone / the right neighbour's left ghost zone (NVLink)
    int ld_out, out_off;       // window / slab outputs: row length of state_out and offset of cell 0 in it
                               //    (0, 0 = [..][nx]; a slab that writes the interior of the next extended state
                               //    passes ld_out = nx + 2*halo, out_off = halo)
};

// fast_radius 1..4 selects the compile-time-radius window path; 0 the generic path.
cudaError_t launch_hybrid_tiles(const HybridArgs& a, int fast_radius, int grid, cudaStream_t stream);

int hybrid_max_active_clusters(int csize);    // 0 when clusters of that size cannot be launched

// Latency mode, hybrid_latency_kernel.cu: one whole-IC tile per cluster of 8 CTAs that split the output features; results
// bit-identical to launch_hybrid_tiles.  For rollouts of fewer tiles than the device has cluster slots.
int hybrid_latency_max_clusters();
bool hybrid_latency_supported(const HybridArgs& a);
cudaError_t launch_hybrid_latency(const HybridArgs& a, int clusters, cudaStream_t stream);

// Tensor-core (tcgen05) variant, hybrid_tc_kernel.cu: radius 1..4, a.hops == 1, segments of 32/64/128 rows.
cudaError_t launch_hybrid_tc_tiles(const HybridArgs& a, int radius, int grid, cudaStream_t stream);

// 16-bit tensor-core variant, hybrid_tc16_kernel.cu: 256-row tiles (a.tile_rows == 256), radius 1..4,
// a.hops == 1, whole-IC segments of 32/64/128 rows.
cudaError_t launch_hybrid_tc16_tiles(const HybridArgs& a, int radius, int grid, cudaStream_t stream);

}  // namespace fluxgnn
