// hybrid_tile_kernel<2, false>: inference, compile-time stencil radius 2
#include "hybrid_kernel_impl.cuh"

namespace fluxgnn {
template cudaError_t launch_one<2, false>(const HybridArgs&, int, cudaStream_t);
}
