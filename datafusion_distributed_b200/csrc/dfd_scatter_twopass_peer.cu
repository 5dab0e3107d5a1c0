// dfd_scatter_twopass_peer.cu — two-pass k_scatter (K1 tiling) instantiations, peer-store (fused exchange) mode (see dfd_launch.cuh).
#include "dfd_launch.cuh"

namespace dfd {
int launch_scatter_twopass_peer(const ScatterParams& sp, int width, bool fast, int sm_count, size_t smem, cudaStream_t stream) {
    return launch_scatter_impl<true, 0>(sp, width, fast, sm_count, smem, stream);
}
}  // namespace dfd
