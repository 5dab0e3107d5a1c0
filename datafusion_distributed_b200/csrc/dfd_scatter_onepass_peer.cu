// dfd_scatter_onepass_peer.cu — single-pass k_scatter_onepass instantiations, peer-store (fused exchange) mode (see dfd_launch.cuh).
#include "dfd_launch.cuh"

namespace dfd {
int launch_scatter_onepass_peer(const ScatterParams& sp, int width, bool fast, int sm_count, size_t smem, cudaStream_t stream) {
    return launch_scatter_impl<true, 1>(sp, width, fast, sm_count, smem, stream);
}
}  // namespace dfd
