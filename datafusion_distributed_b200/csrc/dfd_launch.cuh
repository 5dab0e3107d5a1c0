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
#define DFD_ONEPASS_NB 2
#endif
#ifndef DFD_ONEPASS_MIN_CTAS
#define DFD_ONEPASS_MIN_CTAS 4
#endif
#ifndef DFD_ONEPASS_K
#define DFD_ONEPASS_K 10  // rows per consumer thread per tile (tile = 256 x K rows): larger tiles amortise ranking / look-back
#endif
constexpr int ONEPASS_K = DFD_ONEPASS_K;
constexpr int FOLLOW_MIN_CTAS = 4;  // follow-up k_scatter launches on the single-pass tiling
constexpr int ONEPASS_NB = DFD_ONEPASS_NB;
constexpr int ONEPASS_MIN_CTAS = DFD_ONEPASS_MIN_CTAS;
constexpr int TILE_THREADS = DFD_TILE_THREADS;
constexpr int TILE_K = DFD_TILE_K;
constexpr int TILE_MIN_CTAS = DFD_TILE_MIN_CTAS;
// aligned write-out (see k_scatter): used when N <= ALIGNED_MAX_N; each run wastes < 62 virtual slots
constexpr uint32_t ALIGNED_MAX_N = 16;
constexpr int TILE_KV = TILE_K + (62 * (int)ALIGNED_MAX_N + TILE_THREADS - 1) / TILE_THREADS;
constexpr int TILE_ROWS = TILE_THREADS * TILE_K;
constexpr int ONEPASS_KV = ONEPASS_K + (62 * (int)ALIGNED_MAX_N + TILE_THREADS - 1) / TILE_THREADS;
constexpr int ONEPASS_ROWS = TILE_THREADS * ONEPASS_K;


// MODE: 0 = two-pass k_scatter on the K1 tiling (TILE_K); 1 = single-pass k_scatter_onepass (ONEPASS_K);
//       2 = "follow-up" k_scatter on the SINGLE-PASS tiling: further column-width groups / bit columns of a single-pass
//           call, driven by the per-tile counts and cursors the single-pass launch left in hist_out / base_out
template <bool FAST, typename V, bool PEER, bool ALIGNED, int MODE>
static int launch_scatter_kv(const ScatterParams& sp, int sm_count, size_t smem, cudaStream_t stream) {
    cudaError_t e;
    if constexpr (MODE == 1) {
        if constexpr (std::is_same<V, BitColumn>::value) {
            return set_error(DFD_ERR_INTERNAL, "bit-packed columns take the two-pass k_scatter");
        } else {
            constexpr int KV = ALIGNED ? ONEPASS_KV : ONEPASS_K;
            auto kern = k_scatter_onepass<TILE_THREADS, ONEPASS_K, KV, ONEPASS_NB, ONEPASS_MIN_CTAS, FAST, V, PEER>;
            smem = onepass_smem_bytes<TILE_THREADS, ONEPASS_K, ONEPASS_NB>(sp.N, (int)sizeof(V), PEER, ALIGNED);
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
        constexpr int K = MODE == 2 ? ONEPASS_K : TILE_K;
        constexpr int KV = ALIGNED ? (MODE == 2 ? ONEPASS_KV : TILE_KV) : K;
        constexpr int CTAS = MODE == 2 ? FOLLOW_MIN_CTAS : TILE_MIN_CTAS;
        auto kern = k_scatter<TILE_THREADS, K, KV, CTAS, FAST, V, PEER>;
        if (MODE == 2) smem = scatter_smem_bytes<TILE_THREADS, ONEPASS_K>(sp.N, sp.stage_width, PEER, ALIGNED);
        if (smem > 227 * 1024) return set_error(DFD_ERR_UNSUPPORTED, "k_scatter needs %zu B of shared memory per CTA", smem);
        if (smem > 48 * 1024) {
            if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess)
                return cuda_error(e, "cudaFuncSetAttribute(k_scatter)");
        }
        kern<<<(unsigned)sp.n_tiles, TILE_THREADS, smem, stream>>>(sp);
    }
    e = cudaGetLastError();
    return e == cudaSuccess ? DFD_OK : cuda_error(e, "k_scatter");
}

template <bool FAST, typename V, bool PEER, int MODE>
static int launch_scatter_t(const ScatterParams& sp, int sm_count, size_t smem, cudaStream_t stream) {
    if (use_aligned(sp.N, PEER)) return launch_scatter_kv<FAST, V, PEER, true, MODE>(sp, sm_count, smem, stream);
    return launch_scatter_kv<FAST, V, PEER, false, MODE>(sp, sm_count, smem, stream);
}

template <bool FAST, bool PEER, int MODE>
static int launch_scatter_w(const ScatterParams& sp, int width, int sm_count, size_t smem, cudaStream_t stream) {
    switch (width) {
        case 8: return launch_scatter_t<FAST, uint64_t, PEER, MODE>(sp, sm_count, smem, stream);
        case 4: return launch_scatter_t<FAST, uint32_t, PEER, MODE>(sp, sm_count, smem, stream);
        case 2: return launch_scatter_t<FAST, uint16_t, PEER, MODE>(sp, sm_count, smem, stream);
        case 1: return launch_scatter_t<FAST, uint8_t, PEER, MODE>(sp, sm_count, smem, stream);
        case 16: return launch_scatter_t<FAST, uint4, PEER, MODE>(sp, sm_count, smem, stream);
        default:
            if constexpr (PEER || MODE == 1) {
                return set_error(DFD_ERR_INTERNAL, "bit-packed columns take the local k_scatter instantiation");
            } else {
                return launch_scatter_t<FAST, BitColumn, false, MODE>(sp, sm_count, smem, stream);
            }
    }
}

// one definition per translation unit (dfd_scatter_*.cu)
template <bool PEER, int MODE>
int launch_scatter_impl(const ScatterParams& sp, int width, bool fast, int sm_count, size_t smem, cudaStream_t stream) {
    return fast ? launch_scatter_w<true, PEER, MODE>(sp, width, sm_count, smem, stream)
                : launch_scatter_w<false, PEER, MODE>(sp, width, sm_count, smem, stream);
}

}  // namespace dfd
