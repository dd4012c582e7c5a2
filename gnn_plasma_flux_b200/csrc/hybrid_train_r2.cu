// hybrid_tile_kernel<2, true>: training forward (activations saved), stencil radius 2
#include "hybrid_kernel_impl.cuh"

namespace fluxgnn {
template cudaError_t launch_one<2, true>(const HybridArgs&, int, cudaStream_t);
}
