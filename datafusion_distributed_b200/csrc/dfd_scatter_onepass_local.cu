// dfd_scatter_onepass_local.cu — single-pass k_scatter_onepass instantiations, local mode (see dfd_launch.cuh).
#include "dfd_launch.cuh"

namespace dfd {
int launch_scatter_onepass_local(const ScatterParams& sp, int width, bool fast, int sm_count, size_t smem, cudaStream_t stream) {
    return launch_scatter_impl<false, 1>(sp, width, fast, sm_count, smem, stream);
}
}  // namespace dfd
