// hybrid_tile_kernel<3, false, true>: inference, compile-time stencil radius 3, clustered window tiles
#include "hybrid_kernel_impl.cuh"

namespace fluxgnn {
template cudaError_t launch_one_cluster<3>(const HybridArgs&, int, cudaStream_t);
}
