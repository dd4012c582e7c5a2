// hybrid_tile_kernel<4, false, true>: inference, compile-time stencil radius 4, clustered window tiles
#include "hybrid_kernel_impl.cuh"

namespace fluxgnn {
template cudaError_t launch_one_cluster<4>(const HybridArgs&, int, cudaStream_t);
}
