// dfd_launch.cuh — tile geometry and the template dispatch of the scatter kernels.
// The instantiations are spread over several translation units (dfd_scatter_*.cu) so that they compile in
// parallel; dfd_api.cu only sees the declarations at the bottom of dfd_internal.h.
#pragma once
#include <cuda_runtime.h>

#include <cstdlib>
#include <type_traits>

#include "dfd_internal.h"
#include "dfd_kernels.cuh"

namespace dfd {

// Tile geometry of K1/K2 (rows per CTA = THREADS * K).
#ifndef DFD_TILE_THREADS
#define DFD_TILE_THREADS 256
#endif
#ifndef DFD_TILE_K
#define DFD_TILE_K 6
#endif
#ifndef DFD_TILE_MIN_CTAS
#define DFD_TILE_MIN_CTAS 6
#endif
// single-pass kernel: ring depth (tiles in flight per CTA) and resident CTAs per SM
#ifndef DFD_ONEPASS_NB
#define DFD_ONEPASS_NB 3
#endif
#ifndef DFD_ONEPASS_MIN_CTAS
#define DFD_ONEPASS_MIN_CTAS 4
#endif
constexpr int ONEPASS_NB = DFD_ONEPASS_NB;
constexpr int ONEPASS_MIN_CTAS = DFD_ONEPASS_MIN_CTAS;
constexpr int TILE_THREADS = DFD_TILE_THREADS;
constexpr int TILE_K = DFD_TILE_K;
constexpr int TILE_MIN_CTAS = DFD_TILE_MIN_CTAS;
// aligned write-out (see k_scatter): used when N <= ALIGNED_MAX_N; each run wastes < 62 virtual slots
constexpr uint32_t ALIGNED_MAX_N = 16;
constexpr int TILE_KV = TILE_K + (62 * (int)ALIGNED_MAX_N + TILE_THREADS - 1) / TILE_THREADS;
constexpr int TILE_ROWS = TILE_THREADS * TILE_K;


template <bool FAST, typename V, bool PEER, int KV, bool ONEPASS>
static int launch_scatter_kv(const ScatterParams& sp, int sm_count, size_t smem, cudaStream_t stream) {
    cudaError_t e;
    if constexpr (ONEPASS) {
        if constexpr (std::is_same<V, BitColumn>::value) {
            return set_error(DFD_ERR_INTERNAL, "bit-packed columns take the two-pass k_scatter");
        } else {
            auto kern = k_scatter_onepass<TILE_THREADS, TILE_K, KV, ONEPASS_NB, ONEPASS_MIN_CTAS, FAST, V, PEER>;
            smem = onepass_smem_bytes<TILE_THREADS, TILE_K, ONEPASS_NB>(sp.N, (int)sizeof(V), PEER, KV != TILE_K);
            if (smem > 227 * 1024) return set_error(DFD_ERR_UNSUPPORTED, "single-pass kernel needs %zu B of shared memory per CTA", smem);
            // (static per instantiation: the attribute and the occupancy are properties of the kernel + smem size)
            static thread_local size_t cfg_smem = 0;
            static thread_local int cfg_per_sm = 0;
            if (cfg_smem != smem) {
                if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess)
                    return cuda_error(e, "cudaFuncSetAttribute(k_scatter_onepass)");
                if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&cfg_per_sm, kern, TILE_THREADS + 32, smem)) != cudaSuccess)
                    return cuda_error(e, "cudaOccupancyMaxActiveBlocksPerMultiprocessor");
                if (cfg_per_sm < 1) cfg_per_sm = 1;
                cfg_smem = smem;
            }
            int64_t grid = (int64_t)cfg_per_sm * sm_count;
            if (grid > sp.n_tiles) grid = sp.n_tiles;
            kern<<<(unsigned)grid, TILE_THREADS + 32, smem, stream>>>(sp);
        }
    } else {
        auto kern = k_scatter<TILE_THREADS, TILE_K, KV, TILE_MIN_CTAS, FAST, V, PEER>;
        if (smem > 48 * 1024) {
            if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess)
                return cuda_error(e, "cudaFuncSetAttribute(k_scatter)");
        }
        kern<<<(unsigned)sp.n_tiles, TILE_THREADS, smem, stream>>>(sp);
    }
    e = cudaGetLastError();
    return e == cudaSuccess ? DFD_OK : cuda_error(e, "k_scatter");
}

template <bool FAST, typename V, bool PEER, bool ONEPASS>
static int launch_scatter_t(const ScatterParams& sp, int sm_count, size_t smem, cudaStream_t stream) {
    if (use_aligned(sp.N, PEER)) return launch_scatter_kv<FAST, V, PEER, TILE_KV, ONEPASS>(sp, sm_count, smem, stream);
    return launch_scatter_kv<FAST, V, PEER, TILE_K, ONEPASS>(sp, sm_count, smem, stream);
}

template <bool FAST, bool PEER, bool ONEPASS>
static int launch_scatter_w(const ScatterParams& sp, int width, int sm_count, size_t smem, cudaStream_t stream) {
    switch (width) {
        case 8: return launch_scatter_t<FAST, uint64_t, PEER, ONEPASS>(sp, sm_count, smem, stream);
        case 4: return launch_scatter_t<FAST, uint32_t, PEER, ONEPASS>(sp, sm_count, smem, stream);
        case 2: return launch_scatter_t<FAST, uint16_t, PEER, ONEPASS>(sp, sm_count, smem, stream);
        case 1: return launch_scatter_t<FAST, uint8_t, PEER, ONEPASS>(sp, sm_count, smem, stream);
        case 16: return launch_scatter_t<FAST, uint4, PEER, ONEPASS>(sp, sm_count, smem, stream);
        default:
            if constexpr (PEER || ONEPASS) {
                return set_error(DFD_ERR_INTERNAL, "bit-packed columns take the two-pass local k_scatter instantiation");
            } else {
                return launch_scatter_t<FAST, BitColumn, false, false>(sp, sm_count, smem, stream);
            }
    }
}

// one definition per translation unit (dfd_scatter_*.cu)
template <bool PEER, bool ONEPASS>
int launch_scatter_impl(const ScatterParams& sp, int width, bool fast, int sm_count, size_t smem, cudaStream_t stream) {
    return fast ? launch_scatter_w<true, PEER, ONEPASS>(sp, width, sm_count, smem, stream)
                : launch_scatter_w<false, PEER, ONEPASS>(sp, width, sm_count, smem, stream);
}

}  // namespace dfd
