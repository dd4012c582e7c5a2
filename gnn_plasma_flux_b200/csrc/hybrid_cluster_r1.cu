// hybrid_tile_kernel<1, false, true>: inference, compile-time stencil radius 1, clustered window tiles
#include "hybrid_kernel_impl.cuh"

namespace fluxgnn {
template cudaError_t launch_one_cluster<1>(const HybridArgs&, int, cudaStream_t);
template cudaError_t max_clusters_one<1>(int, int*);
}
