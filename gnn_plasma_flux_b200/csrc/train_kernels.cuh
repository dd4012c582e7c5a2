// Backward kernels of train_kernels.cu (declarations for api.cu).
#pragma once

#include <cuda_runtime.h>

namespace fluxgnn {

__global__ void bwd_mask_mean_kernel(const float* dH, const float* Hn, float* dpre, float* dZ, float* db,
                                     long long rows, int nx, int radius);
__global__ void bwd_gemm_nn_kernel(const float* A1, const float* W1, const float* A2, const float* W2, int ldw,
                                   float* C, long long rows);
__global__ void bwd_gemm_tn_kernel(const float* A, const float* Bm, float* dW, int ldw, long long rows);
__global__ void bwd_edge_kernel(const float* P, const float* Q, const float* w2, const float* dflux, float* dP,
                                float* dQ, float* dw2, float* db1, float* db2, long long rows, int nx, int hops);
__global__ void bwd_input_kernel(const float* dH0, const float* H0, const float* w_in, const float* state,
                                 const float* x, float* dstate, float* dw_in, float* db_in, long long rows, int nx);
__global__ void step_bwd_flux_kernel(const float* g_out, const float* g_face, float* dflux, long long cells, int nx,
                                     float c);
__global__ void step_bwd_direct_kernel(const float* g_out, const float* state, float* dstate, long long cells, int nx,
                                       float c, float dt);

}  // namespace fluxgnn
