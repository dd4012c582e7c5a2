// hybrid_tile_kernel<3, true>: training forward (activations saved), stencil radius 3
#include "hybrid_kernel_impl.cuh"

namespace fluxgnn {
template cudaError_t launch_one<3, true>(const HybridArgs&, int, cudaStream_t);
}
