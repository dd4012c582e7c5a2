// hybrid_tile_kernel<0, true>: training forward (activations saved), generic neighbour walk
#include "hybrid_kernel_impl.cuh"

namespace fluxgnn {
template cudaError_t launch_one<0, true>(const HybridArgs&, int, cudaStream_t);
}
