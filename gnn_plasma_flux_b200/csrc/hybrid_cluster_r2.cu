// hybrid_tile_kernel<2, false, true>: inference, compile-time stencil radius 2, clustered window tiles
#include "hybrid_kernel_impl.cuh"

namespace fluxgnn {
template cudaError_t launch_one_cluster<2>(const HybridArgs&, int, cudaStream_t);
}
