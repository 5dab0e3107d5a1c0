// dfd_scatter_twopass_local.cu — two-pass k_scatter (K1 tiling) instantiations, local mode (see dfd_launch.cuh).
#include "dfd_launch.cuh"

namespace dfd {
int launch_scatter_twopass_local(const ScatterParams& sp, int width, bool fast, int sm_count, size_t smem, cudaStream_t stream) {
    return launch_scatter_impl<false, 0>(sp, width, fast, sm_count, smem, stream);
}
}  // namespace dfd
