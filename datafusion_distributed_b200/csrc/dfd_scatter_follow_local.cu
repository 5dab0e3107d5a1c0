// dfd_scatter_follow_local.cu — follow-up k_scatter on the single-pass tiling instantiations, local mode (see dfd_launch.cuh).
#include "dfd_launch.cuh"

namespace dfd {
int launch_scatter_follow_local(const ScatterParams& sp, int width, bool fast, int sm_count, size_t smem, cudaStream_t stream) {
    return launch_scatter_impl<false, 2>(sp, width, fast, sm_count, smem, stream);
}
}  // namespace dfd
