// dfd_scatter_follow_peer.cu — follow-up k_scatter on the single-pass tiling instantiations, peer-store (fused exchange) mode (see dfd_launch.cuh).
#include "dfd_launch.cuh"

namespace dfd {
int launch_scatter_follow_peer(const ScatterParams& sp, int width, bool fast, int sm_count, size_t smem, cudaStream_t stream) {
    return launch_scatter_impl<true, 2>(sp, width, fast, sm_count, smem, stream);
}
}  // namespace dfd
