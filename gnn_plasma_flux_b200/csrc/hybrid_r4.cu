// hybrid_tile_kernel<4, false>: inference, compile-time stencil radius 4
#include "hybrid_kernel_impl.cuh"

namespace fluxgnn {
template cudaError_t launch_one<4, false>(const HybridArgs&, int, cudaStream_t);
}
