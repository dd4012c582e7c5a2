// hybrid_tile_kernel<4, true>: training forward (activations saved), stencil radius 4
#include "hybrid_kernel_impl.cuh"

namespace fluxgnn {
template cudaError_t launch_one<4, true>(const HybridArgs&, int, cudaStream_t);
}
