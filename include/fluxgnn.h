/*
 * fluxgnn.h -- C ABI of libfluxgnn.so, the sm_100a implementation of the hybrid
 * rollout hot path of shanedirksen/gnn-plasma-flux.
 *
 * The reference is pure Python and has no FFI of its own; every entry point
 * below names the reference code (path:line under /root/reference) whose
 * arithmetic it replaces.  INTEGRATION.md shows the ctypes binding a
 * maintainer of the reference would add.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, no C++ / torch types.
 *   - Every pointer is a DEVICE pointer owned by the caller unless its name
 *     starts with `host_`.  Arrays are row-major contiguous float32 unless
 *     stated.  The library never allocates caller-visible memory and never
 *     synchronises; every call is asynchronous on `stream` (a cudaStream_t
 *     passed as void*; NULL = the legacy default stream).
 *   - Return value: 0 on success, a negative FLUXGNN_E* code on failure.
 *     fluxgnn_last_error() returns a thread-local message for the last failure.
 *   - There is no CPU fallback: without a CUDA device every compute entry
 *     point fails with FLUXGNN_ECUDA.
 *
 * State layout: state[B][3][nx] = (n, u, E) per initial condition ("IC"),
 * the batched form of the reference's numpy [3, nx] state
 * (src/hybrid_solver.py:34-35, src/baseline_solver.py:80-81).
 */
#ifndef FLUXGNN_H_
#define FLUXGNN_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FLUXGNN_ABI_VERSION 4

#define FLUXGNN_OK        0
#define FLUXGNN_EINVAL   -1   /* bad argument (shape, null pointer, unsupported size) */
#define FLUXGNN_ECUDA    -2   /* CUDA runtime error (message has the CUDA string)    */
#define FLUXGNN_EUNSUP   -3   /* valid request this build does not implement          */

/* Architecture limits of this build (src/config.py:19-23 uses F=4, H=128, L=4). */
#define FLUXGNN_INPUT_DIM   4
#define FLUXGNN_HIDDEN    128
#define FLUXGNN_MAX_LAYERS  8
#define FLUXGNN_MAX_HOPS    4    /* edge blocks the forward entry point can emit */

int fluxgnn_abi_version(void);
const char* fluxgnn_last_error(void);

/* Number of kernels this library has launched since it was loaded (all threads). */
unsigned long long fluxgnn_launch_count(void);

/* Measurement aid (bench.py): launches `blocks` x `threads` (<= 256) threads that each run `iters`
 * iterations of 128 FMAs on registers (16 independent chains; packed != 0 uses the
 * two-wide fma.rn.f32x2 / FFMA2 form the GEMM uses); out[blocks*threads] receives a
 * checksum.  Returns the number of FLOPs the launch executes (negative = error).
 * Time it with CUDA events to get the FP32-pipe roofline of the device. */
long long fluxgnn_ffma_probe(float* out, int blocks, int threads, int iters, int packed, void* stream);

/* ---- weights: FluxGNN.state_dict() -> streaming layout -------------------
 * Replaces the parameter container of src/flux_gnn.py:11-38.  Inputs are the
 * tensors of the state_dict, unchanged (nn.Linear layout [out, in]):
 *   w_in[H][F], b_in[H]                       input_mlp.0      (:17-20)
 *   w_upd[l][H][2H], b_upd[l][H]  l<L         update_mlps.l.0  (:23-31), contiguous over l
 *   w_e1[H][2H], b_e1[H]                      edge_mlp.0       (:34-36)
 *   w_e2[H], b_e2[1]                          edge_mlp.2       (:37)
 * `packed` receives fluxgnn_packed_weight_bytes(L) bytes: the small vectors,
 * then for every layer the two [H(k)][H(n)] halves of the weight matrix in the
 * 16-row chunks the kernel streams through shared memory with bulk copies. */
size_t fluxgnn_packed_weight_bytes(int num_layers);
int fluxgnn_pack_weights(const float* w_in, const float* b_in,
                         const float* w_upd, const float* b_upd,
                         const float* w_e1, const float* b_e1,
                         const float* w_e2, const float* b_e2,
                         int num_layers, void* packed, void* stream);

/* ---- field solve ------------------------------------------------------------
 * src/baseline_solver.py:59-68:  E = Re ifft(i * fft(n - 1) / k),  E_hat(0) = 0,
 * k = 2*pi*fftfreq(nx, length/nx) (:26); the Nyquist bin drops out of Re().
 * Two exact realisations of that operator:
 *   - power-of-two nx in 2^8..2^25: FFT (one CTA per IC up to 2^14 cells, a
 *     transpose-free four-step transform above; needs
 *     fluxgnn_poisson_workspace_bytes(B, nx) bytes of scratch, 0 up to 2^14);
 *   - any other nx <= 12288 (and the in-kernel solve of whole-IC tiles,
 *     nx <= 128): circular convolution E = g (*) (n - 1) with
 *     g = Re ifft(i/k), accumulated in fp64.  fluxgnn_poisson_table() fills
 *     gtab[nx] (float64, device); fluxgnn_poisson_uses_table(nx) says whether a
 *     standalone solve at this nx reads it (gtab may be NULL otherwise).
 * `n` and `E` have the given strides (in floats) between ICs so that a channel
 * of a [B][3][nx] state can be passed directly. */
int fluxgnn_poisson_uses_table(int nx);
int fluxgnn_poisson_table(int nx, double length, double* gtab, void* stream);
size_t fluxgnn_poisson_workspace_bytes(int B, int nx);
int fluxgnn_poisson_spectral(const float* n, long long n_ic_stride,
                             float* E, long long e_ic_stride,
                             const double* gtab, int B, int nx, double length,
                             void* workspace, void* stream);

/* ---- FluxGNN.forward on the radius-r ring ----------------------------------
 * Replaces build_chain_graph + FluxGNN.forward (src/graph_constructor.py:30-38,
 * src/flux_gnn.py:40-67) for the periodic chain, with no edge_index tensor:
 * edge blocks are, for hop k = 1..hops, [i -> i+k] then [i+k -> i], nx edges
 * each (k = 1 is exactly the reference's edge order).
 *   state[B][3][nx], x[nx] (cell centres, float32)
 *   flux_edges[B][2*hops*nx]  (nullable)    per directed edge, as FluxGNN.forward
 *   face_flux[B][nx]          (nullable)    0.5*(fwd+bwd) of hop 1 (src/hybrid_solver.py:45-48)
 * radius >= 1; hops in 1..FLUXGNN_MAX_HOPS, hops <= radius. */
int fluxgnn_forward_ring(const void* packed, int num_layers,
                         const float* state, const float* x,
                         int B, int nx, int radius, int hops,
                         float* flux_edges, float* face_flux, void* stream);

/* ---- HybridSolver.step / .run ------------------------------------------------
 * Replaces src/hybrid_solver.py:34-73 for a batch of ICs.  One call advances
 * `steps` time steps:  GNN face flux -> n' = n - (dt/dx)(F_i - F_{i-1});
 * u' = u - (dt/dx)(u_i^2/2 - u_{i-1}^2/2) + dt*E (no viscosity, as the
 * reference) -> E' = solve_poisson(n').
 *   state_in[B][3][nx] -> state_out[B][3][nx]   (may not alias)
 *   traj: nullable; when given, the state after every `record_every`-th step
 *         is stored at traj[(t/record_every)-1][B][3][nx], t = 1..steps.
 *   workspace: fluxgnn_hybrid_workspace_bytes(B, nx) bytes (0 for nx <= 128).
 *   gtab: required for nx <= 128 and wherever fluxgnn_poisson_uses_table(nx).
 *   c = float32(dt/dx) and dt = float32(dt), the scalars numpy uses
 *   (src/hybrid_solver.py:52,57-58). */
size_t fluxgnn_hybrid_workspace_bytes(int B, int nx);
int fluxgnn_hybrid_rollout(const void* packed, int num_layers,
                           const float* state_in, float* state_out,
                           const float* x, const double* gtab,
                           int B, int nx, double length, int radius, float c, float dt,
                           int steps, int record_every, float* traj,
                           void* workspace, void* stream);

/* ---- tensor-core variant of the same step (tcgen05 / TMEM) ---------------------
 * The five 128x128x256 contractions of src/flux_gnn.py:60,66 run on the tensor cores;
 * everything else (window mean, ReLU, edge readout, finite-volume update, field solve)
 * is unchanged.  precision:
 *   FLUXGNN_TC_TF32X3  h and W are split into two TF32 parts and three products are
 *                      accumulated in fp32: fp32-level accuracy (same parity gates as
 *                      fluxgnn_hybrid_rollout).
 *   FLUXGNN_TC_TF32    one TF32 product: ~3e-4 relative flux error (tolerances and measured
 *                      values in DESIGN.md section 6).
 * packed_tc comes from fluxgnn_pack_weights_tc (same inputs as fluxgnn_pack_weights;
 * the stream holds pre-swizzled UMMA operand images, hi and lo parts).
 * Supported: radius <= 4; nx in {32, 64, 128} or nx > 128; other shapes -> FLUXGNN_EUNSUP
 * (the caller decides to use the fp32 entry point; there is no silent fallback).
 *
 * 16-bit operand modes (tcgen05 kind::f16, 256-cell tiles; twice the TF32 rate per product):
 *   FLUXGNN_TC_FP16X3  h and 2^8 W are split into two fp16 parts each (22 significant bits
 *                      together, as in the TF32 split), three products accumulated in fp32 and
 *                      rescaled: fp32-level accuracy, same parity gates as fluxgnn_hybrid_rollout.
 *   FLUXGNN_TC_FP16    one fp16 product (11-bit operands like plain TF32; activations beyond
 *                      65504 overflow).
 *   FLUXGNN_TC_BF16    one bfloat16 product (8-bit operands): loosest tolerance, fp32 range.
 * Their packed_tc comes from fluxgnn_pack_weights_tc16(..., precision, ...): the fp16 image
 * serves FP16X3 and FP16, the bf16 image serves BF16.
 * Range of the fp16 modes: the image stores 2^8 W, so the update- and edge-layer weights must stay
 * below 65504 / 2^8 = 255.875 in magnitude, and hidden activations below 65504; beyond that the
 * operands are inf and the outputs non-finite (fluxgnn_hybrid_rollout_diag / fluxgnn_rollout_metrics
 * count them).  The packing call is asynchronous and cannot report it: the Python mirror checks the
 * weights before packing (FluxGNN.packed_weights) and raises; C callers check max|W| themselves. */
#define FLUXGNN_TC_TF32X3 1
#define FLUXGNN_TC_TF32   2
#define FLUXGNN_TC_FP16X3 3
#define FLUXGNN_TC_FP16   4
#define FLUXGNN_TC_BF16   5
size_t fluxgnn_packed_tc16_weight_bytes(int num_layers);
int fluxgnn_pack_weights_tc16(const float* w_in, const float* b_in,
                              const float* w_upd, const float* b_upd,
                              const float* w_e1, const float* b_e1,
                              const float* w_e2, const float* b_e2,
                              int num_layers, int precision, void* packed_tc, void* stream);
size_t fluxgnn_packed_tc_weight_bytes(int num_layers);
int fluxgnn_pack_weights_tc(const float* w_in, const float* b_in,
                            const float* w_upd, const float* b_upd,
                            const float* w_e1, const float* b_e1,
                            const float* w_e2, const float* b_e2,
                            int num_layers, void* packed_tc, void* stream);
int fluxgnn_forward_ring_tc(const void* packed_tc, int num_layers, int precision,
                            const float* state, const float* x,
                            int B, int nx, int radius,
                            float* flux_edges /* [B][2*nx], hop 1 */, float* face_flux, void* stream);
int fluxgnn_hybrid_rollout_tc(const void* packed_tc, int num_layers, int precision,
                              const float* state_in, float* state_out,
                              const float* x, const double* gtab,
                              int B, int nx, double length, int radius, float c, float dt,
                              int steps, int record_every, float* traj,
                              void* workspace, void* stream);

/* ---- FluxGNN of any size on the ring --------------------------------------------------------------
 * src/flux_gnn.py:11-67 for architectures other than (4, 128, L): input_dim 1..16, hidden_dim in
 * {16, 32, 64, 128}, 1..8 layers -- e.g. FluxGNN(4, 64, 3) of examples/smoke_test.py:50-56 and the class
 * default (2, 32, 2).  Plain FP32-pipe kernels, one CTA per 128-cell window; same outputs and edge order
 * as fluxgnn_forward_ring.  Weights: the state_dict tensors as for fluxgnn_pack_weights, packed K-major by
 * fluxgnn_generic_pack into fluxgnn_generic_packed_bytes(F, H, L) bytes.
 *   feats[B][nx][F]  node features as FluxGNN.forward receives them (src/graph_constructor.py:30-32), or
 *   feats = NULL and state[B][3][nx] + x[nx] for input_dim = 4 (features n, u, E, x).
 * fluxgnn_generic_hybrid_rollout: HybridSolver.step / .run (src/hybrid_solver.py:34-73) with such a model
 * (input_dim 4): per step the forward kernel, the finite-volume update and the field solve;
 * workspace: fluxgnn_generic_workspace_bytes(B, nx); gtab as for fluxgnn_hybrid_rollout. */
size_t fluxgnn_generic_packed_bytes(int input_dim, int hidden, int num_layers);
int fluxgnn_generic_pack(const float* w_in, const float* b_in, const float* w_upd, const float* b_upd,
                         const float* w_e1, const float* b_e1, const float* w_e2, const float* b_e2,
                         int input_dim, int hidden, int num_layers, void* packed, void* stream);
int fluxgnn_generic_forward_ring(const void* packed, int input_dim, int hidden, int num_layers,
                                 const float* feats, const float* state, const float* x,
                                 int B, int nx, int radius, int hops,
                                 float* flux_edges, float* face_flux, void* stream);
size_t fluxgnn_generic_workspace_bytes(int B, int nx);
int fluxgnn_generic_hybrid_rollout(const void* packed, int hidden, int num_layers,
                                   const float* state_in, float* state_out,
                                   const float* x, const double* gtab,
                                   int B, int nx, double length, int radius, float c, float dt,
                                   int steps, int record_every, float* traj,
                                   void* workspace, void* stream);

/* ---- rollout with in-kernel diagnostics -----------------------------------
 * fluxgnn_hybrid_rollout for grids of up to 128 cells that also reduces, inside the persistent kernel
 * and after EVERY step, what scripts/evaluation/evaluate_all.py:134-141 and
 * evaluate_long_rollout.py:53-66 compute from host copies of the trajectory:
 *   diag[step][ic] = { 0.5*mean(u^2 + E^2), mean(n), number of non-finite values, 0 }   (4 floats)
 * so a 1000-step sweep over 65536 ICs returns 1 GB of diagnostics instead of 800 GB of states.
 * precision 0 = fp32 kernel (packed from fluxgnn_pack_weights), FLUXGNN_TC_* = tensor kernels. */
int fluxgnn_hybrid_rollout_diag(const void* packed, int num_layers, int precision,
                                const float* state_in, float* state_out,
                                const float* x, const double* gtab,
                                int B, int nx, double length, int radius, float c, float dt,
                                int steps, float* diag /* [steps][B][4] */, void* stream);

/* ---- training: FluxGNN.forward with saved activations and its backward pass --
 * What scripts/training/train_ablation.py:128-206 needs from the model: the edge fluxes of
 * src/flux_gnn.py:40-67 on the ring, differentiable w.r.t. every parameter and the node features.
 *   fluxgnn_forward_ring_train  = fluxgnn_forward_ring (fp32 kernel) that also stores, row-major
 *       [row = ic*nx + cell][128], h^0..h^L, (P + b1) and Q of the edge readout into
 *       acts[(L+3)][B*nx][128]  (fluxgnn_train_acts_bytes).
 *   fluxgnn_backward_ring: given dflux[B][2*hops*nx] (gradient w.r.t. flux_edges) ADDS the
 *       parameter gradients into g_* (nn.Linear layouts, the caller zeroes them) and writes
 *       dstate[B][3][nx] (gradient w.r.t. n, u, E; nullable).  Weights are the raw state_dict tensors
 *       (w_upd contiguous over layers).  workspace: fluxgnn_backward_workspace_bytes(B, nx). */
size_t fluxgnn_train_acts_bytes(int num_layers, int B, int nx);
int fluxgnn_forward_ring_train(const void* packed, int num_layers, const float* state, const float* x,
                               int B, int nx, int radius, int hops, float* flux_edges, float* acts,
                               void* stream);
size_t fluxgnn_backward_workspace_bytes(int B, int nx);
int fluxgnn_backward_ring(const float* w_in, const float* w_upd, const float* w_e1, const float* w_e2,
                          int num_layers, const float* state, const float* x, const float* acts,
                          const float* dflux, int B, int nx, int radius, int hops,
                          float* g_w_in, float* g_b_in, float* g_w_upd, float* g_b_upd,
                          float* g_w_e1, float* g_b_e1, float* g_w_e2, float* g_b_e2,
                          float* dstate, void* workspace, void* stream);

/* ---- one differentiable hybrid step (the training rollout) -----------
 * The body of the reference's multi-step training rollout, scripts/training/train_ablation.py:172-206:
 * per step  model(build_chain_graph(state)) -> face flux -> n', u' in torch ops (differentiable) ->
 * field solve through numpy (DETACHED, :198-200).  Here the forward is ONE launch of the fused tile
 * kernel in its activation-saving instantiation (plus the field-solve kernel for nx > 128), and the
 * backward is fluxgnn_backward_ring bracketed by the two finite-volume adjoint kernels.
 *   fluxgnn_hybrid_step_train: as fluxgnn_hybrid_rollout with steps = 1, and also writes
 *       face_flux[B][nx] (nullable; F_pred of train_ablation.py:126) and acts (fluxgnn_train_acts_bytes).
 *       workspace: fluxgnn_hybrid_workspace_bytes(B, nx) (may be NULL when that is 0).
 *   fluxgnn_hybrid_step_backward: g_state_out[B][3][nx] = gradient w.r.t. the step's output (the E
 *       channel is ignored: E' is detached), g_face[B][nx] = gradient w.r.t. face_flux (nullable).
 *       ADDS the parameter gradients into g_* (as fluxgnn_backward_ring) and WRITES dstate[B][3][nx],
 *       the gradient w.r.t. (n, u, E) of state_in through the network and through the update formulas.
 *       workspace: fluxgnn_step_backward_workspace_bytes(B, nx). */
int fluxgnn_hybrid_step_train(const void* packed, int num_layers, const float* state_in, float* state_out,
                              const float* x, const double* gtab, int B, int nx, double length, int radius,
                              float c, float dt, float* face_flux, float* acts, void* workspace, void* stream);
size_t fluxgnn_step_backward_workspace_bytes(int B, int nx);
int fluxgnn_hybrid_step_backward(const float* w_in, const float* w_upd, const float* w_e1, const float* w_e2,
                                 int num_layers, const float* state, const float* x, const float* acts,
                                 const float* g_state_out, const float* g_face, int B, int nx, int radius,
                                 float c, float dt, float* g_w_in, float* g_b_in, float* g_w_upd, float* g_b_upd,
                                 float* g_w_e1, float* g_b_e1, float* g_w_e2, float* g_b_e2, float* dstate,
                                 void* workspace, void* stream);

/* ---- rollout diagnostics on the device ---------------------------
 * The metrics every evaluation script of the reference computes on the host after
 * copying whole trajectories back (scripts/evaluation/evaluate_all.py:118-159,
 * evaluate_long_rollout.py:38-66).  pred / truth: num_states stored states [3][nx]
 * (any leading [T][B] shape flattened); truth nullable.  out[num_states][8]:
 *   0..2 mean squared error per channel (n, u, E); 3 energy 0.5*mean(u^2+E^2);
 *   4 charge mean(n); 5 count of non-finite values; 6..7 zero. */
int fluxgnn_rollout_metrics(const float* pred, const float* truth, long long num_states, int nx,
                            float* out, void* stream);

/* ---- one slab of a domain-decomposed grid (single large grid) -------
 * The same hybrid step for `owned` consecutive cells of a longer periodic grid held
 * by another rank.  The caller supplies ghost cells instead of the periodic wrap:
 *   state_ext[B][3][owned + 2*halo]  cells -halo .. owned+halo-1 of this slab
 *   x_ext[owned + 2*halo]            their GLOBAL cell centres (float32)
 *   state_out[B][3][owned]           n' and u' are written; E' is NOT (the field solve
 *                                    is global: gather n' and call fluxgnn_poisson_spectral)
 * halo must equal num_layers*radius + 1, the receptive field of one step.
 * precision: 0 = fp32 kernel (packed from fluxgnn_pack_weights), FLUXGNN_TC_* = tensor
 * kernel (packed from fluxgnn_pack_weights_tc). */
int fluxgnn_hybrid_slab_step(const void* packed, int num_layers, int precision,
                             const float* state_ext, const float* x_ext, float* state_out,
                             int B, int owned, int halo, int radius, float c, float dt, void* stream);

/* The same slab step writing into a row-strided output: n' and u' of cell i go to
 * state_out[ic][0..1][out_off + i] of an array [B][3][out_ld].  With out_ld = owned + 2*halo and
 * out_off = halo the slab writes the interior of the NEXT extended state directly (the ghost cells
 * are then filled by the neighbours' halo stores / sends): no concatenation on the step path. */
int fluxgnn_hybrid_slab_step_ld(const void* packed, int num_layers, int precision,
                                const float* state_ext, const float* x_ext, float* state_out,
                                int out_ld, int out_off, int B, int owned, int halo, int radius,
                                float c, float dt, void* stream);

/* ---- peer-memory exchange of a domain-decomposed grid (halo exchange + distributed field solve) ----
 * For ranks whose extended states sit in SYMMETRIC memory (every rank has its peers' allocations mapped over
 * NVLink; gnn_plasma_flux_b200.domain.SymmetricMemoryFabric): the exchange is plain stores into the peers'
 * memory, ordered by the fabric's signal-pad barrier -- no NCCL call on the step path.
 *   fluxgnn_peer_halo_push: stores channels ch0..ch1-1 of this rank's first / last `halo` owned cells into the
 *       right ghosts of left_ext / the left ghosts of right_ext (the ring neighbours' extended states
 *       [B][3][owned + 2*halo], peer pointers).
 *   fluxgnn_peer_allgather: copies `bytes` (multiple of 16) from src into slot `rank` of every rank's gather
 *       buffer, which sits `offset` bytes into that rank's symmetric allocation; peer_bases_dev is a DEVICE array
 *       of the `world` allocation base pointers (torch's _SymmetricMemory.buffer_ptrs_dev). */
/* The same exchange FUSED into the producing kernels (the steady state of the peer-memory step): the slab kernels
 * store n', u' of their first / last `halo` owned cells also into left_out / right_out -- the ring neighbours' NEXT
 * extended states, laid out like state_out (out_ld, out_off >= halo) -- and the field kernel of the prefix-sum solve
 * stores E' of those cells into E_left / E_right (the E rows of the neighbours' next states, laid out like E / e_ld).
 * Otherwise identical to fluxgnn_hybrid_slab_step_ld, fluxgnn_baseline_slab_step and fluxgnn_scan_slab_field. */
int fluxgnn_hybrid_slab_step_peer(const void* packed, int num_layers, int precision,
                                  const float* state_ext, const float* x_ext, float* state_out,
                                  int out_ld, int out_off, int B, int owned, int halo, int radius,
                                  float c, float dt, float* left_out, float* right_out, void* stream);
int fluxgnn_baseline_slab_step_peer(const float* state_ext, float* state_out, int out_ld, int out_off, float* flux_n,
                                    int B, int owned, int halo, float c, float dt, float nu, float dx2,
                                    float* left_out, float* right_out, void* stream);
int fluxgnn_scan_slab_field_peer(const float* n, long long n_ld, float* E, long long e_ld, int B, int S, int rank,
                                 int ranks, double length, const void* msg_all, void* workspace, double cert_tol,
                                 int step, int* first_uncertified, float* E_left, float* E_right, int halo,
                                 void* stream);
/* fluxgnn_scan_slab_sums whose message kernel also stores the 48-byte message of every IC into slot `rank` of every
 * rank's gather buffer (`offset` bytes into the symmetric allocations whose bases peer_bases_dev lists): the all-gather
 * of the distributed prefix-sum solve fused into its producer. */
int fluxgnn_scan_slab_sums_peer(const float* n, long long n_ld, int B, int S, long long j_base, void* workspace,
                                void* msg, const void* peer_bases_dev, long long offset, int rank, int world,
                                void* stream);
int fluxgnn_peer_halo_push(const float* state_ext, float* left_ext, float* right_ext, int B, int owned, int halo,
                           int ch0, int ch1, void* stream);
int fluxgnn_peer_allgather(const void* src, long long bytes, const void* peer_bases_dev, long long offset, int rank,
                           int world, void* stream);

/* ---- one slab of the classical solver (baseline-only domain decomposition) ------
 * src/baseline_solver.py:80-94 (upwind fluxes, viscous Laplacian, forward Euler) for `owned`
 * consecutive cells of a longer periodic grid:  state_ext[B][3][owned + 2*halo] carries `halo` >= 1
 * ghost cells per side (the stencil reads one; halo = 4 keeps 128-bit loads aligned).  n', u' are
 * written as in fluxgnn_hybrid_slab_step_ld; E' comes from the distributed field solve.
 * flux_n: nullable [B][owned], the continuity flux F_n. */
int fluxgnn_baseline_slab_step(const float* state_ext, float* state_out, int out_ld, int out_off,
                               float* flux_n, int B, int owned, int halo,
                               float c, float dt, float nu, float dx2, void* stream);

/* ---- distributed field solve ("distributed FFT via all-to-all") --------------------------
 * The operator of src/baseline_solver.py:59-68 for ONE periodic grid of nx = G*S cells whose rank r
 * (r < G, G and S powers of two, S >= 256, G <= 16) holds cells [r*S, (r+1)*S) of B density rows.
 * Two rows (ICs 2p, 2p+1) travel as one complex signal z = (n_a - 1) + i (n_b - 1): the multiplier
 * i/k is then diagonal and the result is E_a + i E_b.  P = ceil(B/2).  The nx-point transform is
 * decimated in frequency over the rank index (element jr*S + jl, bin kr + G*kl).  Per step and rank:
 *   1. fluxgnn_poisson_dist_pack      n[B][S] (row stride ic_stride) -> z[P][S]
 *   2. all-to-all (the caller's; NCCL): flat z cut into G equal chunks, chunk q -> rank q
 *   3. fluxgnn_poisson_dist_rank_dft  in[G][chunk] (chunk of every rank) -> out[G][chunk] (for every rank),
 *                                     flat0 = rank*chunk is the flat index of the chunk's first element
 *   4. all-to-all -> y[P][S], the rank's own share of the spectrum's input
 *   5. fluxgnn_poisson_dist_local     y in place: FFT_S, diagonal multiplier of bins rank + G*kl, inverse FFT_S
 *                                     (scratch: P*S complex numbers, needed above S = 2^14)
 *   6. all-to-all, fluxgnn_poisson_dist_rank_dft(inverse = 1), all-to-all -> e[P][S] = E_a + i E_b
 *   7. fluxgnn_poisson_dist_unpack    e -> E[B][S] (row stride ic_stride, e.g. the E channel of an extended state)
 * With G = 1 steps 2-4 and 6 are the identity and the result equals fluxgnn_poisson_spectral to rounding. */
int fluxgnn_poisson_dist_pack(const float* n, long long ic_stride, int B, int S, void* z, void* stream);
int fluxgnn_poisson_dist_unpack(const void* e, float* E, long long ic_stride, int B, int S, void* stream);
int fluxgnn_poisson_dist_rank_dft(const void* in, void* out, int G, long long chunk, long long flat0, int S,
                                  int inverse, void* stream);
int fluxgnn_poisson_dist_local(void* y, void* scratch, int P, int S, int G, int rank, double length, void* stream);

/* ---- BaselineSolver.step / .run ------------------------------------------------
 * Replaces src/baseline_solver.py:70-118: upwind continuity flux n*u,
 * left-differenced u^2/2, viscous Laplacian, forward Euler, field solve.
 *   flux_n: nullable [steps][B][nx], the continuity flux F_n of every step
 *           (`return_flux` / `record_flux`, :84,:99-100,:109-111).
 *   traj:   as above.  The kernel divides by dx2 = float32(dx*dx) exactly as
 *           numpy does (:78).
 *   workspace: fluxgnn_baseline_workspace_bytes(B, nx) bytes, required when
 *           steps > 1 or the field solve needs scratch. */
size_t fluxgnn_baseline_workspace_bytes(int B, int nx);
int fluxgnn_baseline_rollout(const float* state_in, float* state_out,
                             const double* gtab, int B, int nx, double length,
                             float c, float dt, float nu, float dx2,
                             int steps, int record_every, float* traj,
                             float* flux_n, void* workspace, void* stream);

/* Latency mode of fluxgnn_hybrid_rollout and fluxgnn_forward_ring (fp32 kernel, whole-IC tiles, radius <= 4): when a call has at most twice as
 * many 128-row tiles as the device has cluster slots, every tile is computed by a cluster of 8 CTAs that split the
 * output features of each layer (csrc/hybrid_latency_kernel.cu) -- bit-identical results, ~3x lower time per step for the
 * reference's own timing protocol of one 64-cell IC (scripts/evaluation/benchmark_timing.py:63-72).
 * FLUXGNN_LATENCY=0 / 1 in the environment forces the choice. */
int fluxgnn_latency_cluster_slots(void);

/* ---- classical rollout with the field solve as a certified prefix sum ("scan solve") -----------
 * The same step as fluxgnn_baseline_rollout (src/baseline_solver.py:80-101) for long grids
 * (fluxgnn_baseline_scan_supported: nx >= 4096, nx % 8 == 0).  The operator of
 * src/baseline_solver.py:59-68 is the zero-mean periodic antiderivative of -(n - 1 - mean); it is
 * evaluated as trapezoid prefix sum + first Euler-Maclaurin term, and for every field so obtained
 * the kernel evaluates the rigorous bound
 *     max|E_scan - E_spectral| <= rms(4th difference of n) * length / (32 sqrt 3).
 * first_uncertified (DEVICE int[B]) receives, per IC, the index of the first step whose input field had
 * bound > cert_tol * max|E| (index `steps` = the field of the final state), INT_MAX if every field of
 * that IC was certified.  The caller reads it after the stream has finished and repeats the ICs that
 * are not INT_MAX with fluxgnn_baseline_rollout (BaselineSolver.rollout(field_solve="auto")
 * does exactly that).  Between steps only n and u live in HBM (16 B per cell-update); E is written
 * for recorded states and the final state.  state_in's E is used for the first step as given.
 *   workspace: fluxgnn_baseline_scan_workspace_bytes(B, nx) bytes. */
int fluxgnn_baseline_scan_supported(int B, int nx);
size_t fluxgnn_baseline_scan_workspace_bytes(int B, int nx);
int fluxgnn_baseline_rollout_scan(const float* state_in, float* state_out, int B, int nx, double length,
                                  float c, float dt, float nu, float dx2,
                                  int steps, int record_every, float* traj, float* flux_n,
                                  double cert_tol, void* workspace, int* first_uncertified,
                                  void* stream);

/* ---- the scan solve on one slab of a domain-decomposed grid (north_star's "distributed
 * Poisson reduction") ------------------------------------------------------------------------
 * Rank r of `ranks` owns cells [r S, (r+1) S) of every IC (S % 8 == 0).  Per field solve:
 *   1. fluxgnn_scan_slab_sums: the slab's sums into `msg` (B records of FLUXGNN_SCAN_MSG_BYTES);
 *   2. the caller all-gathers the messages -> msg_all[ranks][B];
 *   3. fluxgnn_scan_slab_field: E rows of the slab, the operator of src/baseline_solver.py:59-68 as
 *      in fluxgnn_baseline_rollout_scan, certificate included: *first_uncertified (device int, set
 *      to INT_MAX by the caller before step 0) receives step-1 if the field of the PREVIOUS call
 *      had bound > cert_tol * max|E| (its sums travel in the messages); after the last step one
 *      more sums + all-gather + fluxgnn_scan_slab_certify closes the record.
 *   n / E: [B] rows, row strides n_ld / e_ld floats (views into the extended slab buffers).
 *   workspace: fluxgnn_scan_slab_workspace_bytes(B, S) bytes, ZEROED by the caller once, then kept
 *   across calls (it carries the certificate sums from one call to the next). */
#define FLUXGNN_SCAN_MSG_BYTES 48
int fluxgnn_scan_slab_supported(int B, int S);
size_t fluxgnn_scan_slab_workspace_bytes(int B, int S);
int fluxgnn_scan_slab_sums(const float* n, long long n_ld, int B, int S, long long j_base,
                           void* workspace, void* msg, void* stream);
int fluxgnn_scan_slab_field(const float* n, long long n_ld, float* E, long long e_ld, int B, int S,
                            int rank, int ranks, double length, const void* msg_all, void* workspace,
                            double cert_tol, int step, int* first_uncertified, void* stream);
int fluxgnn_scan_slab_certify(int B, int S, int ranks, double length, const void* msg_all,
                              double cert_tol, int step, int* first_uncertified, void* stream);

/* ---- the reference's comparison models ---------------------------------------
 * PureGNN (scripts/training/train_pure_gnn.py:35-76), rolled out as in
 * scripts/evaluation/benchmark_timing.py:129-143: per step  state += PureGNN([n,u,E,x], ring edges).
 * Weights: input_mlp.0 [H][4], update_mlps.l.0 stacked [L][H][2H] / [L][H], output_mlp.0 [H][H],
 * output_mlp.2 [3][H] / [3] (nn.Linear layouts), H in {64, 128}; nx <= 128; one launch per rollout.
 * PINN (scripts/training/train_pinn.py:36-61) is a chain of fluxgnn_dense_layer calls:
 *   out[r][n] = act(sum_k in[r][k] weight[n][k] + bias[n]) (+ residual[r][n]),  activation 0 = none, 1 = tanh. */
size_t fluxgnn_pure_gnn_packed_bytes(int hidden, int num_layers);
int fluxgnn_pure_gnn_pack(const float* w_in, const float* b_in, const float* w_upd, const float* b_upd,
                          const float* w_o1, const float* b_o1, const float* w_o2, const float* b_o2,
                          int hidden, int num_layers, void* packed, void* stream);
int fluxgnn_pure_gnn_rollout(const void* packed, int hidden, int num_layers,
                             const float* state_in /* [B][3][nx] */, float* state_out,
                             const float* x, int B, int nx, int steps, void* stream);
/* PureGNN.forward itself (train_pure_gnn.py:57-76): delta_out[B][3][nx] = the model output, NOT added to the state. */
int fluxgnn_pure_gnn_delta(const void* packed, int hidden, int num_layers,
                           const float* state /* [B][3][nx] */, float* delta_out,
                           const float* x, int B, int nx, void* stream);
int fluxgnn_dense_layer(const float* in, const float* weight, const float* bias,
                        const float* residual /* nullable */, float* out,
                        int rows, int in_features, int out_features, int activation, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FLUXGNN_H_ */
