// dfd_scatter_twopass_peer.cu — k_scatter instantiations, peer-store (fused exchange) mode (see dfd_launch.cuh).
#include "dfd_launch.cuh"

namespace dfd {
int launch_scatter_twopass_peer(const ScatterParams& sp, int width, bool fast, int sm_count, size_t smem, cudaStream_t stream) {
    return launch_scatter_impl<true, false>(sp, width, fast, sm_count, smem, stream);
}
}  // namespace dfd
