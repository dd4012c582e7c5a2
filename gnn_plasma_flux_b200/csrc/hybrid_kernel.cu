// Radius dispatch of the fused hybrid tile kernel; the kernel itself is in
// hybrid_kernel_impl.cuh and is instantiated once per radius in hybrid_r*.cu.
#include "common.cuh"
#include "hybrid_kernel.cuh"

namespace fluxgnn {

template <int R, bool kSave>
cudaError_t launch_one(const HybridArgs& a, int grid, cudaStream_t stream);

template <bool kSave>
static cudaError_t dispatch(const HybridArgs& a, int fast_radius, int grid, cudaStream_t stream) {
    switch (fast_radius) {
        case 1: return launch_one<1, kSave>(a, grid, stream);
        case 2: return launch_one<2, kSave>(a, grid, stream);
        case 3: return launch_one<3, kSave>(a, grid, stream);
        case 4: return launch_one<4, kSave>(a, grid, stream);
        default: return launch_one<0, kSave>(a, grid, stream);
    }
}

template <int R>
cudaError_t max_clusters_one(int csize, int* out);
template <int R>
cudaError_t launch_one_cluster(const HybridArgs& a, int grid, cudaStream_t stream);

// Clusters of `csize` CTAs the device runs concurrently (every radius instantiation has the same footprint); cached.
int hybrid_max_active_clusters(int csize) {
    static int cache[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    if (csize < 2 || csize > 8) return 0;
    if (cache[csize] == 0) {
        int n = 0;
        if (max_clusters_one<1>(csize, &n) != cudaSuccess || n < 1) {
            (void)cudaGetLastError();
            n = -1;
        }
        cache[csize] = n;
    }
    return cache[csize] > 0 ? cache[csize] : 0;
}

cudaError_t launch_hybrid_tiles(const HybridArgs& a, int fast_radius, int grid, cudaStream_t stream) {
    if (a.cluster > 1) {
        switch (fast_radius) {
            case 1: return launch_one_cluster<1>(a, grid, stream);
            case 2: return launch_one_cluster<2>(a, grid, stream);
            case 3: return launch_one_cluster<3>(a, grid, stream);
            case 4: return launch_one_cluster<4>(a, grid, stream);
            default: return cudaErrorInvalidValue;          // api.cu only clusters compile-time radii
        }
    }
    return a.acts != nullptr ? dispatch<true>(a, fast_radius, grid, stream) : dispatch<false>(a, fast_radius, grid, stream);
}

}  // namespace fluxgnn
