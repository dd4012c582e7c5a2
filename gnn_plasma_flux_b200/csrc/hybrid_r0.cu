// hybrid_tile_kernel<0, false>: inference, generic neighbour walk (any radius, any nx)
#include "hybrid_kernel_impl.cuh"

namespace fluxgnn {
template cudaError_t launch_one<0, false>(const HybridArgs&, int, cudaStream_t);
}
