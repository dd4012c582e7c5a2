// Per-row pieces of the fused step that all three tile kernels share (FP32-pipe, TF32 tensor, 16-bit
// tensor): which cell a tile row is, loading its state, the finite-volume update in numpy's fp32
// operation order, one row of the in-tile field solve and the write-out.  The kernels differ in how
// they compute the GNN face flux between these pieces, not in the pieces.
#pragma once

#include "hybrid_kernel.cuh"

namespace fluxgnn {

// Shared-memory arrays of a CTA tile (one entry per tile row), whatever struct they live in.
struct TileRows {
    float *sN, *sU, *sE, *sX;        // state (n, u, E) and position of the row's cell
    float *sF, *sRho;                // face flux F_{j+1/2}; rho = n' - 1 for the field solve
    const double* gtab;              // field-solve kernel g (whole-IC tiles)
    int *rowIC, *rowCell;            // owning IC of the row's OUTPUT (-1: not owned) and its cell
    short *prevRow, *nextRow;        // periodic neighbours inside the IC (whole-IC tiles) / the window
};

// Row j of the CTA tile is row jl of logical tile `tile`, whose `rows` rows start at row0.
//   whole-IC tiles: the logical tile holds a.ics_per_tile complete ICs, neighbours wrap inside each IC
//                   (src/graph_constructor.py:34-38);
//   window tiles  : rows [halo, halo + valid) of window t of one IC are owned, the rest is recomputed halo;
//   slabs         : as windows, with ghost cells instead of the periodic wrap.
constexpr int kClusterStep = 124;      // rows between the pieces of a cluster window: 128 minus a 4-row (one 16-byte chunk) overlap

__device__ __forceinline__ void tile_load_row(const HybridArgs& a, const TileRows& T, int tile, bool tile_ok, int j, int jl,
                                              int row0, int rows, int crank = 0) {
    const int nx = a.nx;
    int ic, cell, prev = row0 + ((jl - 1) & (rows - 1)), next = row0 + ((jl + 1) & (rows - 1));
    bool live, owned;
    int src = -1, ld = nx;                  // source index / row length when they differ from (cell, nx)
    if (a.whole_ic) {
        const int slot = jl / nx;
        cell = jl - slot * nx;
        ic = tile * a.ics_per_tile + slot;
        live = tile_ok && (slot < a.ics_per_tile) && (ic < a.B);
        owned = live;
        if (slot < a.ics_per_tile) {
            prev = (cell == 0) ? j + nx - 1 : j - 1;
            next = (cell == nx - 1) ? j - nx + 1 : j + 1;
        }
    } else {
        // packed remainders (api.cu plan_tiles): the short last windows of pack_per_tile consecutive ICs share one tile,
        // each a segment with its own halo on both sides; rows next to a junction see the other segment's cells, which is
        // harmless because they are halo rows whose results are dropped
        const long long full_tiles = (long long)a.B * a.pack_full;
        if (a.pack_full > 0 && tile >= full_tiles) {
            const int seg = jl / a.pack_seg, w = jl - seg * a.pack_seg;
            const long long lic = (long long)(tile - full_tiles) * a.pack_per_tile + seg;
            ic = (int)lic;
            live = tile_ok && seg < a.pack_per_tile && lic < a.B;
            const long long gcell = (long long)a.pack_full * a.valid - a.halo + w;
            cell = (int)(((gcell % nx) + nx) % nx);
            owned = live && w >= a.halo && w < a.halo + a.pack_rem;
        } else {
            ic = tile / a.tiles_per_ic;
            const int t = tile - ic * a.tiles_per_ic;
            // piece `crank` of a cluster window starts kClusterStep rows after its left neighbour; it owns the rows
            // 4-hops .. 127-hops (first piece: from the halo; last piece: up to the halo), so the owned rows tile the window
            const int csize = a.cluster > 1 ? a.cluster : 1;
            const int wrow = crank * kClusterStep + jl;                     // row inside the window
            const long long gcell = (long long)t * a.valid - a.halo + wrow;
            cell = (int)(((gcell % nx) + nx) % nx);
            live = tile_ok;
            // (a row's outputs need the edge partial sums of the `hops` rows after it and the face flux of the row before it)
            const int lo = (crank == 0) ? a.halo : 4 - a.hops, hi = (crank == csize - 1) ? rows - a.halo : rows - a.hops;
            owned = tile_ok && (jl >= lo) && (jl < hi) && ((long long)t * a.valid + (wrow - a.halo) < nx);
            if (a.slab) {                       // ghost cells instead of the periodic wrap
                long long s = gcell + a.halo;
                s = s < 0 ? 0 : (s >= a.ld_in ? a.ld_in - 1 : s);
                src = (int)s;
                ld = a.ld_in;
                cell = (int)(gcell < 0 ? 0 : (gcell >= nx ? nx - 1 : gcell));
            }
        }
    }
    T.rowIC[j] = owned ? ic : -1;
    T.rowCell[j] = cell;
    T.prevRow[j] = (short)prev;
    T.nextRow[j] = (short)next;
    float vn = 0.f, vu = 0.f, ve = 0.f, vx = 0.f;
    if (live) {
        if (src < 0) src = cell;
        const float* st = a.state_in + (size_t)ic * 3 * ld + src;
        vn = __ldg(st);
        vu = __ldg(st + ld);
        ve = __ldg(st + 2 * (size_t)ld);
        vx = __ldg(a.x + src);
    }
    T.sN[j] = vn; T.sU[j] = vu; T.sE[j] = ve; T.sX[j] = vx;
}

// Finite-volume update of row j, numpy's fp32 operation order (src/hybrid_solver.py:51-58; no viscosity):
//   n' = n - c (F_j - F_{j-1}),   u' = u - c (u_j^2/2 - u_{j-1}^2/2) + dt E
__device__ __forceinline__ void tile_fv_update_values(const HybridArgs& a, float n, float u, float up, float e, float face,
                                                      float face_prev, float& n_new, float& u_new) {
    n_new = __fsub_rn(n, __fmul_rn(a.c, __fsub_rn(face, face_prev)));
    const float fu = __fmul_rn(__fmul_rn(0.5f, u), u);
    const float fup = __fmul_rn(__fmul_rn(0.5f, up), up);
    const float u_adv = __fsub_rn(u, __fmul_rn(a.c, __fsub_rn(fu, fup)));
    u_new = __fadd_rn(u_adv, __fmul_rn(a.dt, e));
}
__device__ __forceinline__ void tile_fv_update(const HybridArgs& a, const TileRows& T, int j, float& n_new, float& u_new) {
    const int p = T.prevRow[j];
    tile_fv_update_values(a, T.sN[j], T.sU[j], T.sU[p], T.sE[j], T.sF[j], T.sF[p], n_new, u_new);
}

// Window tiles: n', u' of an owned row go straight to state_out (E' comes from the field-solve kernel).
__device__ __forceinline__ void tile_store_window_row(const HybridArgs& a, const TileRows& T, int j, float n_new, float u_new) {
    if (T.rowIC[j] < 0) return;
    const int ld = a.ld_out ? a.ld_out : a.nx;
    float* so = a.state_out + (size_t)T.rowIC[j] * 3 * ld + a.out_off + T.rowCell[j];
    so[0] = n_new;
    so[ld] = u_new;
    // slabs over peer memory: the halo exchange is this store (the neighbours' ghost zones of their next state)
    const int cell = T.rowCell[j];
    if (a.peer_left != nullptr && cell < a.halo) {
        float* pl = a.peer_left + (size_t)T.rowIC[j] * 3 * ld + a.out_off + a.nx + cell;
        pl[0] = n_new;
        pl[ld] = u_new;
    }
    if (a.peer_right != nullptr && cell >= a.nx - a.halo) {
        float* pr = a.peer_right + (size_t)T.rowIC[j] * 3 * ld + a.out_off - a.nx + cell;
        pr[0] = n_new;
        pr[ld] = u_new;
    }
}

// Whole-IC tiles: the new state stays in shared memory; rho = n' - n0 (src/baseline_solver.py:60).
__device__ __forceinline__ void tile_keep_row(const TileRows& T, int j, float n_new, float u_new) {
    T.sN[j] = n_new;
    T.sU[j] = u_new;
    T.sRho[j] = __fsub_rn(n_new, 1.0f);
}

// One thread's share (cells part, part + parts, ...) of E_row = sum_i g[(cell - i) mod nx] rho_i, accumulated in
// fp64 (src/baseline_solver.py:59-68 as a circular convolution).
__device__ __forceinline__ double tile_field_partial(const TileRows& T, int row, int part, int parts, int nx) {
    const int cell = T.rowCell[row], base = row - cell;
    double e = 0.0;
    for (int i = part; i < nx; i += parts) {
        int d = cell - i;
        if (d < 0) d += nx;
        e = fma(T.gtab[d], (double)T.sRho[base + i], e);
    }
    return e;
}

// Whole-IC tiles: final state after the last step, and the recorded steps of the trajectory.
__device__ __forceinline__ void tile_write_out_row(const HybridArgs& a, const TileRows& T, int j, int step) {
    if (T.rowIC[j] < 0) return;
    const int nx = a.nx;
    const size_t off = (size_t)T.rowIC[j] * 3 * nx + T.rowCell[j];
    if (step == a.steps - 1) {
        a.state_out[off] = T.sN[j];
        a.state_out[off + nx] = T.sU[j];
        a.state_out[off + 2 * (size_t)nx] = T.sE[j];
    }
    if (a.traj != nullptr && (step + 1) % a.record_every == 0) {
        float* tr = a.traj + (size_t)((step + 1) / a.record_every - 1) * a.B * 3 * nx + off;
        tr[0] = T.sN[j];
        tr[nx] = T.sU[j];
        tr[2 * (size_t)nx] = T.sE[j];
    }
    // Per-step diagnostics of the IC whose first cell this row is (scripts/evaluation/evaluate_all.py:134-141,
    // evaluate_long_rollout.py:53-66): one thread walks the IC's rows in order, fp64 sums -> deterministic.
    if (a.diag != nullptr && T.rowCell[j] == 0) {
        double e = 0.0, c = 0.0;
        int bad = 0;
        for (int i = 0; i < nx; ++i) {
            const float n = T.sN[j + i], u = T.sU[j + i], f = T.sE[j + i];
            e += (double)u * u + (double)f * f;
            c += n;
            bad += (isfinite(n) ? 0 : 1) + (isfinite(u) ? 0 : 1) + (isfinite(f) ? 0 : 1);
        }
        float* d = a.diag + ((size_t)step * a.B + T.rowIC[j]) * 4;
        d[0] = (float)(0.5 * e / nx);
        d[1] = (float)(c / nx);
        d[2] = (float)bad;
        d[3] = 0.f;
    }
}

}  // namespace fluxgnn
