// hybrid_tile_kernel<1, false>: inference, compile-time stencil radius 1
#include "hybrid_kernel_impl.cuh"

namespace fluxgnn {
template cudaError_t launch_one<1, false>(const HybridArgs&, int, cudaStream_t);
}
