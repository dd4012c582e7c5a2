// hybrid_tile_kernel<3, false>: inference, compile-time stencil radius 3
#include "hybrid_kernel_impl.cuh"

namespace fluxgnn {
template cudaError_t launch_one<3, false>(const HybridArgs&, int, cudaStream_t);
}
