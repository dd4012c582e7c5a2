// hybrid_tile_kernel<1, true>: training forward (activations saved), stencil radius 1
#include "hybrid_kernel_impl.cuh"

namespace fluxgnn {
template cudaError_t launch_one<1, true>(const HybridArgs&, int, cudaStream_t);
}
