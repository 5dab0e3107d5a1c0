// dfd_kernels.cuh — sm_100a kernels of the hash-repartition hot path.
//
// Replaces the CPU inner loop of DataFusion's RepartitionExec(Hash) that the
// reference runs on every producer worker (src/worker/impl_execute_task.rs:77-86):
//   create_hashes -> `hash % N` -> per-destination index vectors -> take per column
// with three device passes over Arrow columnar buffers:
//   K1 k_tile_hist     hash(keys) -> destination -> per-tile radix histogram
//   K1b k_scan_tiles                   exclusive scans -> per-(tile,destination) write cursors
//   K2 k_scatter       fused hash -> stable rank (warp match/ballot) -> shared-memory
//                      staging of each column in destination order -> coalesced run writes
// Integer / byte work bounded by HBM bandwidth; no tensor cores.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <type_traits>

#include "dfd_hash.cuh"
#include "dfd_types.cuh"

namespace dfd {

// ---------------------------------------------------------------------------
// small block-scan helper: exclusive scan of one value per thread
// ---------------------------------------------------------------------------
template <int THREADS>
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* s_warp /*[THREADS/32 + 1]*/,
                                                         uint32_t& block_total) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += t;
    }
    if (lane == 31) s_warp[w] = inc;
    __syncthreads();
    if (w == 0) {
        uint32_t x = lane < THREADS / 32 ? s_warp[lane] : 0;
        uint32_t xi = x;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, xi, d);
            if (lane >= d) xi += t;
        }
        if (lane < THREADS / 32) s_warp[lane] = xi - x;
        if (lane == 31) s_warp[THREADS / 32] = xi;
    }
    __syncthreads();
    uint32_t res = s_warp[w] + inc - v;
    block_total = s_warp[THREADS / 32];
    __syncthreads();
    return res;
}

// Lanes of the warp holding the same destination id `d` (d < 2^nbits), from
// nbits ballots (fully unrolled per bit count; nbits is warp-uniform).
template <int NB>
__device__ __forceinline__ unsigned peers_of_t(uint32_t d) {
    unsigned peers = 0xffffffffu;
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        unsigned bit = (d >> b) & 1u;
        unsigned bal = __ballot_sync(0xffffffffu, bit);
        peers &= bal ^ (bit - 1u);  // bit ? bal : ~bal
    }
    return peers;
}

__device__ __forceinline__ unsigned peers_of(uint32_t d, int nbits) {
    switch (nbits) {
        case 1: return peers_of_t<1>(d);
        case 2: return peers_of_t<2>(d);
        case 3: return peers_of_t<3>(d);
        case 4: return peers_of_t<4>(d);
        case 5: return peers_of_t<5>(d);
        case 6: return peers_of_t<6>(d);
        case 7: return peers_of_t<7>(d);
        case 8: return peers_of_t<8>(d);
        default: return __match_any_sync(0xffffffffu, d);
    }
}

// ---------------------------------------------------------------------------
// K0 (debug / parity): destination id per row
// ---------------------------------------------------------------------------
__global__ void k_partition_ids(KeySet keys, HashState st, ModN mod, int64_t n_rows, uint32_t* __restrict__ dest) {
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_rows; r += (int64_t)gridDim.x * blockDim.x)
        dest[r] = mod_n(row_hash<false>(keys, r, st), mod);
}

// ---------------------------------------------------------------------------
// K1: per-tile destination histogram.  Tile t covers rows [t*T, (t+1)*T) —
// the same tiling K2 uses.  hist is destination-major ([N][n_tiles]) so the
// tile scan reads contiguously.
//   NF > 0 : N <= 4*NF.  Each thread counts its K rows into NF packed u64
//            accumulators (four 16-bit fields each), one xor-shuffle tree per
//            warp adds them up and lanes 0..N-1 publish the fields: ~6
//            integer instructions per row instead of a ballot cascade.
//   NF == 0: any N.  Warp-aggregated shared-memory atomics (ballot peers).
// ---------------------------------------------------------------------------
template <int THREADS, int K, bool FAST_I64, int NF>
__global__ void __launch_bounds__(THREADS) k_tile_hist(KeySet keys, HashState st, ModN mod, int64_t n_rows,
                                                        int64_t n_tiles, uint32_t N, uint32_t* __restrict__ hist) {
    constexpr int T = THREADS * K;
    static_assert(K * 32 < 65536, "16-bit packed counters");
    extern __shared__ uint32_t s_hist[];
    const int lane = threadIdx.x & 31;
    const int nbits = 32 - __clz(N);  // ids 0..N (N = "no row")
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        for (uint32_t p = threadIdx.x; p < N; p += THREADS) s_hist[p] = 0;
        __syncthreads();
        const int64_t row0 = tile * T;
        const int tile_rows = (int)((n_rows - row0) < T ? (n_rows - row0) : T);
        uint32_t d[K];
        if (tile_rows == T) {  // full tile (all but the last): no per-row bounds predicates
#pragma unroll
            for (int j = 0; j < K; ++j) d[j] = mod_n(row_hash<FAST_I64>(keys, row0 + j * THREADS + (int)threadIdx.x, st), mod);
        } else {
#pragma unroll
            for (int j = 0; j < K; ++j) {
                int t = j * THREADS + (int)threadIdx.x;
                d[j] = t < tile_rows ? mod_n(row_hash<FAST_I64>(keys, row0 + t, st), mod) : N;
            }
        }
        if constexpr (NF > 0) {
            unsigned long long acc[NF];
#pragma unroll
            for (int q = 0; q < NF; ++q) acc[q] = 0;
#pragma unroll
            for (int j = 0; j < K; ++j) {
                unsigned long long inc = d[j] < N ? 1ULL << ((d[j] & 3u) * 16u) : 0ULL;
#pragma unroll
                for (int q = 0; q < NF; ++q) acc[q] += ((d[j] >> 2) == (uint32_t)q) ? inc : 0ULL;
            }
#pragma unroll
            for (int q = 0; q < NF; ++q) {
#pragma unroll
                for (int sh = 16; sh >= 1; sh >>= 1) acc[q] += __shfl_xor_sync(0xffffffffu, acc[q], sh);
            }
            if ((uint32_t)lane < N) {
                unsigned long long a = acc[0];
#pragma unroll
                for (int q = 1; q < NF; ++q) a = ((lane >> 2) == q) ? acc[q] : a;
                uint32_t c = (uint32_t)(a >> ((lane & 3) * 16)) & 0xffffu;
                if (c) atomicAdd(&s_hist[lane], c);
            }
        } else {
#pragma unroll
            for (int j = 0; j < K; ++j) {
                unsigned peers = peers_of(d[j], nbits);
                if (d[j] < N && (peers & ((1u << lane) - 1)) == 0) atomicAdd(&s_hist[d[j]], __popc(peers));
            }
        }
        __syncthreads();
        for (uint32_t p = threadIdx.x; p < N; p += THREADS) hist[(int64_t)p * n_tiles + tile] = s_hist[p];
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------
// K1b: one block per destination: exclusive scan of its tile counts
// (tile_base, relative to the destination's start), then the LAST block to
// finish turns the N totals into part_starts[N+1].  `done` is a zeroed
// counter the kernel resets for the next call.
// ---------------------------------------------------------------------------
template <int THREADS>
__device__ __forceinline__ unsigned long long block_exclusive_scan_u64(unsigned long long v, unsigned long long* s_warp,
                                                                       unsigned long long& total) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    unsigned long long inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        unsigned long long t = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += t;
    }
    if (lane == 31) s_warp[w] = inc;
    __syncthreads();
    if (w == 0) {
        unsigned long long x = lane < THREADS / 32 ? s_warp[lane] : 0;
        unsigned long long xi = x;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            unsigned long long t = __shfl_up_sync(0xffffffffu, xi, d);
            if (lane >= d) xi += t;
        }
        if (lane < THREADS / 32) s_warp[lane] = xi - x;
        if (lane == 31) s_warp[32] = xi;
    }
    __syncthreads();
    unsigned long long res = s_warp[w] + inc - v;
    total = s_warp[32];
    __syncthreads();
    return res;
}

template <int THREADS>
__global__ void __launch_bounds__(THREADS) k_scan_tiles(const uint32_t* __restrict__ hist, uint32_t* __restrict__ tile_base,
                                                         int64_t* __restrict__ totals, int64_t* __restrict__ part_starts,
                                                         unsigned* __restrict__ done, int64_t n_tiles, uint32_t N) {
    constexpr int W = THREADS / 32;
    __shared__ unsigned long long s_warp[33];
    __shared__ unsigned long long s_wsum[W];
    __shared__ bool s_last;
    const uint32_t p = blockIdx.x;
    const uint32_t* h = hist + (int64_t)p * n_tiles;
    uint32_t* b = tile_base + (int64_t)p * n_tiles;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    // warp w owns the contiguous tiles [lo, hi); every access is a coalesced 32-wide row
    const int64_t per = ((n_tiles + W - 1) / W + 31) & ~(int64_t)31;
    const int64_t lo = (int64_t)w * per;
    const int64_t hi = lo + per < n_tiles ? lo + per : n_tiles;
    unsigned long long sum = 0;
    for (int64_t i = lo + lane; i < hi; i += 32) sum += h[i];
#pragma unroll
    for (int sh = 16; sh >= 1; sh >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, sh);
    if (lane == 0) s_wsum[w] = sum;
    __syncthreads();
    unsigned long long total = 0, base = 0;
#pragma unroll
    for (int ww = 0; ww < W; ++ww) {
        unsigned long long v = s_wsum[ww];
        if (ww < w) base += v;
        total += v;
    }
    unsigned long long run = base;
    for (int64_t i0 = lo; i0 < hi; i0 += 32) {
        const int64_t i = i0 + lane;
        const uint32_t v = i < hi ? h[i] : 0;
        uint32_t inc = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= d) inc += t;
        }
        if (i < hi) b[i] = (uint32_t)(run + inc - v);
        run += __shfl_sync(0xffffffffu, inc, 31);
    }
    if (threadIdx.x == 0) {
        totals[p] = (int64_t)total;
        __threadfence();
        s_last = atomicAdd(done, 1u) == N - 1;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    // last block: part_starts[q] = sum(totals[0..q)), part_starts[N] = n_rows
    unsigned long long carry = 0;
    for (uint32_t q0 = 0; q0 < N; q0 += THREADS) {
        uint32_t q = q0 + threadIdx.x;
        unsigned long long v = q < N ? (unsigned long long)((volatile int64_t*)totals)[q] : 0;
        unsigned long long tot;
        unsigned long long ex = block_exclusive_scan_u64<THREADS>(v, s_warp, tot);
        if (q < N) part_starts[q] = (int64_t)(carry + ex);
        carry += tot;
    }
    if (threadIdx.x == 0) {
        part_starts[N] = (int64_t)carry;
        *done = 0;
    }
}

// ---------------------------------------------------------------------------
// K2: fused hash -> stable rank -> staged scatter of every column.
//
// One CTA owns one tile of T = THREADS*K rows.  Warp w owns the contiguous
// rows [w*32K, (w+1)*32K) of the tile and walks them in K rounds of 32, so
// every global load is one fully coalesced 32-lane request.  Ranks come from
// __match_any_sync + running per-warp counters in shared memory (stable: rank
// order == row order), a block scan turns them into positions in a tile-local
// staging buffer sorted by destination, and each column is then (a) scattered
// into the staging buffer and (b) streamed out so that every destination's
// run is written as consecutive, coalesced global stores.
// ---------------------------------------------------------------------------
template <typename V>
struct StageIO {
    static __device__ __forceinline__ V ld(const void* base, int64_t i) { return ((const V*)base)[i]; }
};

constexpr uint32_t SLOT_NONE = 0xffffffffu;

template <int THREADS, int K, int KV, typename V, int CHUNK = K>
__device__ __forceinline__ void scatter_fixed_column(const PayloadCol& c, void* stage_raw, int64_t row0, int tile_rows,
                                                     const uint32_t (&ps)[K], const uint32_t (&slot)[KV], const int64_t* delta,
                                                     int t0, void* const* out_base /* per destination (peer mode) or nullptr */) {
    static_assert(K % CHUNK == 0, "CHUNK must divide K");
    V* stage = (V*)stage_raw;
    const V* in = (const V*)c.in + (c.in_offset + row0);  // tile-relative indexing below is 32-bit
    V* out = (V*)c.out;                                    // local mode: delta[] holds absolute output rows
    V v[CHUNK];
#pragma unroll
    for (int j = 0; j < CHUNK; ++j) {
        int t = t0 + j * 32;
        if (t < tile_rows) v[j] = in[t];
    }
    __syncthreads();  // staging buffer free (previous column fully written out)
#pragma unroll
    for (int ch = 0; ch < K / CHUNK; ++ch) {
        if (ch > 0) {
#pragma unroll
            for (int j = 0; j < CHUNK; ++j) {
                int t = t0 + (ch * CHUNK + j) * 32;
                if (t < tile_rows) v[j] = in[t];
            }
        }
#pragma unroll
        for (int j = 0; j < CHUNK; ++j) {
            int t = t0 + (ch * CHUNK + j) * 32;
            if (t < tile_rows) stage[ps[ch * CHUNK + j]] = v[j];
        }
    }
    __syncthreads();  // staging buffer holds the tile in destination order
#pragma unroll
    for (int k = 0; k < KV; ++k) {
        if (slot[k] != SLOT_NONE) {
            const uint32_t i = slot[k] & 0xffffu, p = slot[k] >> 16;
            V* o = out_base ? (V*)out_base[p] : out;  // peer mode: the owner rank's receive window (NVLink store)
            o[(int64_t)i + delta[p]] = stage[i];
        }
    }
}

// bit column (boolean values or a validity bitmap): staged as one byte per row,
// written back with warp-aggregated atomicOr on 32-bit output words.
template <int THREADS, int K, int KV>
__device__ __forceinline__ void scatter_bit_column(const PayloadCol& c, void* stage_raw, int64_t row0, int tile_rows,
                                                   const uint32_t (&ps)[K], const uint32_t (&slot)[KV], const int64_t* delta, int t0) {
    uint8_t* stage = (uint8_t*)stage_raw;
    const uint8_t* in = (const uint8_t*)c.in;
    unsigned* out = (unsigned*)c.out;
    const int lane = threadIdx.x & 31;
    uint8_t v[K];
#pragma unroll
    for (int j = 0; j < K; ++j) {
        int t = t0 + j * 32;
        v[j] = (t < tile_rows) ? (uint8_t)bit_is_set(in, row0 + t + c.in_offset) : 0;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < K; ++j) {
        int t = t0 + j * 32;
        if (t < tile_rows) stage[ps[j]] = v[j];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < KV; ++k) {
        const bool active = slot[k] != SLOT_NONE;
        const uint32_t i = slot[k] & 0xffffu;
        int64_t d = active ? (int64_t)i + delta[slot[k] >> 16] : -1;
        unsigned bit = (active && stage[i]) ? (1u << (d & 31)) : 0u;
        int64_t word = active ? (d >> 5) : -1;
        unsigned peers = __match_any_sync(0xffffffffu, word);
        unsigned merged = __reduce_or_sync(peers, bit);
        if (active && merged && (peers & ((1u << lane) - 1)) == 0) atomicOr(out + word, merged);
    }
}

struct BitColumn {};  // tag: bit-packed column (boolean values / validity bitmap)

// ---- decoupled look-back descriptors (single-pass mode) --------------------
// One 64-bit word per (destination, tile): high half = (call epoch << 2) | state, low half = rows.
// Status and value travel in ONE word, so relaxed loads/stores suffice (Merrill & Garland's
// single-word trick); a word whose epoch is not the current call's reads as "not published", so
// the table never needs clearing between calls.
constexpr uint32_t LB_AGG = 1u;     // value = rows of this tile for the destination
constexpr uint32_t LB_PREFIX = 2u;  // value = rows of tiles 0..this for the destination (inclusive)

__device__ __forceinline__ unsigned long long lb_pack(uint32_t epoch, uint32_t state, uint32_t value) {
    return ((unsigned long long)((epoch << 2) | state) << 32) | value;
}
__device__ __forceinline__ void lb_store(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long lb_load(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// first output row of destination p's region (single-pass mode without an explicit dest_base[])
template <bool PEER>
__device__ __forceinline__ int64_t region_base_of(const ScatterParams& P, uint32_t p) {
    if (P.dest_base) return P.dest_base[p];
    if (PEER) return ((int64_t)(p % P.parts_per_rank) * P.world + P.rank) * P.region_stride;
    return (int64_t)p * P.region_stride;
}

// One instantiation per element type V: a launch moves all columns of one
// width (the host groups them), so the hot instantiation (8-byte values)
// carries no code or registers for the other widths.
// KV: write-out iterations per thread.  KV == K: staging slots are written out
// linearly.  KV > K ("aligned" mode, small N): the write-out walks a virtual slot
// space in which every destination's run is shifted so that each warp's 32 rows
// start on a 32-row (256 B for 8-byte values) boundary of the OUTPUT buffer —
// full-line stores to HBM and full-size write packets over NVLink.
// ONEPASS: no K1/K1b.  CTAs take tile tickets in launch order, count their own
// destinations while ranking (phase 1), publish the counts and resolve their
// write cursors by decoupled look-back over the predecessors' descriptors
// (warp w looks back for destinations w, w+W, ... 32 predecessors at a time):
// every row is hashed once and the key column is read once.  Order is stable
// (cursor = sum over lower tiles).  Destinations live in fixed regions
// (dest_base / region_stride); a tile that would overflow a region sets
// overflow_out and writes nothing (the host re-runs with exact regions).
template <int THREADS, int K, int KV, int MIN_CTAS, bool FAST_I64, typename V, bool PEER, bool ONEPASS>
__global__ void __launch_bounds__(THREADS, MIN_CTAS) k_scatter(const __grid_constant__ ScatterParams P) {
    constexpr int T = THREADS * K;
    constexpr int W = THREADS / 32;
    extern __shared__ __align__(16) unsigned char smem[];
    const uint32_t N = P.N;
    // layout: stage | delta[N] | warp_cnt[W][N] | tile_start[N+1] | scan scratch | misc[2]
    unsigned char* stage = smem;
    const uint32_t off_delta = ((uint32_t)T * (uint32_t)P.stage_width + 15u) & ~15u;
    const uint32_t off_wc = off_delta + N * 8u;
    const uint32_t off_ts = off_wc + (uint32_t)W * N * 4u;
    const uint32_t off_scan = off_ts + (N + 1u) * 4u;
    const uint32_t off_misc = off_scan + (uint32_t)(W + 1) * 4u;          // [0] tile ticket, [1] tile overflow
    const uint32_t off_ob = (off_misc + 2u * 4u + 7u) & ~7u;               // peer mode only: per-destination bases
    const uint32_t off_vs = off_ob + N * 8u;                                // aligned mode only: virtual run starts
#define OUT_BASE ((void**)(smem + off_ob))
#define VSTART ((uint32_t*)(smem + off_vs))
    if (!ONEPASS && P.abort_flag && *P.abort_flag) return;  // a window / region overflowed: write nothing
#define DELTA ((int64_t*)(smem + off_delta))
#define WARP_CNT ((uint32_t*)(smem + off_wc))
#define TILE_START ((uint32_t*)(smem + off_ts))
#define S_SCAN ((uint32_t*)(smem + off_scan))
#define S_MISC ((uint32_t*)(smem + off_misc))

    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int64_t tile = blockIdx.x;
    if constexpr (ONEPASS) {
        if (threadIdx.x == 0) {
            S_MISC[0] = atomicAdd(P.lb_ticket, 1u);
            S_MISC[1] = 0;
        }
        __syncthreads();
        tile = S_MISC[0];
    }
    const int64_t row0 = tile * T;
    const int tile_rows = (int)((P.n_rows - row0) < T ? (P.n_rows - row0) : T);
    const int t0 = w * (K * 32) + lane;  // this thread's first tile-relative row; rows t0 + 32*j

    // ---- two-pass mode: tile_start / delta from the K1 histogram (independent of phase 1).
    // delta[p] maps a staging slot i to its absolute output row: out_row = i + delta[p]
    if constexpr (!ONEPASS) {
        uint32_t carry = 0;
        for (uint32_t p0 = 0; p0 < N; p0 += THREADS) {
            uint32_t p = p0 + threadIdx.x;
            uint32_t c = p < N ? P.hist[(int64_t)p * P.n_tiles + tile] : 0;
            uint32_t tot;
            uint32_t ex = block_exclusive_scan<THREADS>(c, S_SCAN, tot);
            if (p < N) {
                uint32_t ts = carry + ex;
                TILE_START[p] = ts;
                DELTA[p] = region_base_of<PEER>(P, p) + (int64_t)P.tile_base[(int64_t)p * P.n_tiles + tile] - (int64_t)ts;
            }
            carry += tot;
        }
        if (threadIdx.x == 0) TILE_START[N] = carry;
    }

    // ---- phase 1: destination + stable rank of every row of the tile
    uint32_t* wc = WARP_CNT + (uint32_t)w * N;
    for (uint32_t p = lane; p < N; p += 32) wc[p] = 0;
    __syncwarp();
    const int nbits = 32 - __clz(N);
    uint32_t pos[K];  // first: (dest << 16 | rank) ; later: staging position
#pragma unroll
    for (int j = 0; j < K; ++j) {
        int t = t0 + j * 32;
        bool valid = t < tile_rows;
        uint32_t d = valid ? mod_n(row_hash<FAST_I64>(P.keys, row0 + t, P.st), P.mod) : N;
        unsigned peers = peers_of(d, nbits);
        uint32_t rank = __popc(peers & ((1u << lane) - 1));
        uint32_t base = valid ? wc[d] : 0;
        __syncwarp();
        if (valid && rank == 0) wc[d] = base + __popc(peers);
        __syncwarp();
        pos[j] = (d << 16) | (base + rank);
    }
    __syncthreads();
    if constexpr (ONEPASS) {
        // tile counts = sum of the warps' counts; publish them, then turn warp_cnt into staging bases
        uint32_t carry = 0;
        for (uint32_t p0 = 0; p0 < N; p0 += THREADS) {
            uint32_t p = p0 + threadIdx.x;
            uint32_t c = 0;
            if (p < N) {
#pragma unroll
                for (int ww = 0; ww < W; ++ww) c += WARP_CNT[(uint32_t)ww * N + p];
                // tile 0 has no predecessor: its aggregate IS its inclusive prefix
                lb_store(P.lb_desc + (int64_t)p * P.n_tiles + tile, lb_pack(P.lb_epoch, tile == 0 ? LB_PREFIX : LB_AGG, c));
            }
            uint32_t tot;
            uint32_t ex = block_exclusive_scan<THREADS>(c, S_SCAN, tot);
            if (p < N) {
                uint32_t run = carry + ex;
                TILE_START[p] = run;
#pragma unroll
                for (int ww = 0; ww < W; ++ww) {
                    uint32_t cc = WARP_CNT[(uint32_t)ww * N + p];
                    WARP_CNT[(uint32_t)ww * N + p] = run;
                    run += cc;
                }
            }
            carry += tot;
        }
        if (threadIdx.x == 0) TILE_START[N] = carry;
    } else {
        // warp_cnt[w][p] -> staging base of (warp w, destination p)
        for (uint32_t p = threadIdx.x; p < N; p += THREADS) {
            uint32_t run = TILE_START[p];
#pragma unroll
            for (int ww = 0; ww < W; ++ww) {
                uint32_t c = WARP_CNT[(uint32_t)ww * N + p];
                WARP_CNT[(uint32_t)ww * N + p] = run;
                run += c;
            }
        }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < K; ++j) {
        uint32_t d = pos[j] >> 16;
        pos[j] = d < N ? wc[d] + (pos[j] & 0xffffu) : 0;
    }
    if constexpr (ONEPASS) {
        // ---- decoupled look-back: exclusive prefix of every destination over the lower tiles
        for (uint32_t p = (uint32_t)w; p < N; p += W) {
            const uint32_t cnt = TILE_START[p + 1] - TILE_START[p];
            uint32_t excl = 0;
            if (tile > 0) {
                const unsigned long long* d = P.lb_desc + (int64_t)p * P.n_tiles;
                int64_t j = tile - 1;  // lane 0 looks at the nearest predecessor
                for (;;) {
                    const int64_t idx = j - lane;
                    uint32_t st = LB_PREFIX, val = 0;
                    if (idx >= 0) {
                        unsigned long long v;
                        uint32_t hi;
                        do {
                            v = lb_load(d + idx);
                            hi = (uint32_t)(v >> 32);
                        } while ((hi >> 2) != P.lb_epoch || (hi & 3u) == 0u);
                        st = hi & 3u;
                        val = (uint32_t)v;
                    }
                    const unsigned pm = __ballot_sync(0xffffffffu, st == LB_PREFIX);
                    const int first = __ffs(pm) - 1;  // nearest predecessor with an inclusive prefix (-1: none)
                    excl += __reduce_add_sync(0xffffffffu, (first < 0 || lane <= first) ? val : 0u);
                    if (pm) break;
                    j -= 32;
                }
                if (lane == 0) lb_store(P.lb_desc + (int64_t)p * P.n_tiles + tile, lb_pack(P.lb_epoch, LB_PREFIX, excl + cnt));
            }
            if (lane == 0) {
                const int64_t cap = P.dest_cap ? P.dest_cap[p] : P.region_stride;
                if ((int64_t)excl + (int64_t)cnt > cap) S_MISC[1] = 1;
                DELTA[p] = region_base_of<PEER>(P, p) + (int64_t)excl - (int64_t)TILE_START[p];
                if (P.hist_out) {
                    P.hist_out[(int64_t)p * P.n_tiles + tile] = cnt;
                    P.base_out[(int64_t)p * P.n_tiles + tile] = excl;
                }
                if (tile == P.n_tiles - 1) P.totals_out[p] = (int64_t)excl + (int64_t)cnt;
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            if (S_MISC[1]) *P.overflow_out = 1;
            // every CTA has its ticket once `n_tiles` CTAs are past this point: the last one re-arms the counters
            if (atomicAdd(P.lb_ticket + 1, 1u) == (unsigned)(P.n_tiles - 1)) {
                P.lb_ticket[0] = 0;
                P.lb_ticket[1] = 0;
            }
        }
        if (S_MISC[1]) return;  // a region is too small: this tile writes nothing
    }
    // ---- which staging slot (and destination) each of this thread's write-out iterations handles
    uint32_t slot[KV];
    if constexpr (KV == K) {
#pragma unroll
        for (int k = 0; k < K; ++k) {
            uint32_t i = k * THREADS + threadIdx.x;
            uint32_t lo = 0, hi = N;  // last p with tile_start[p] <= i
            while (hi - lo > 1) {
                uint32_t mid = (lo + hi) >> 1;
                if (TILE_START[mid] <= i) lo = mid; else hi = mid;
            }
            slot[k] = i < (uint32_t)tile_rows ? (i | (lo << 16)) : SLOT_NONE;
        }
    } else {
        // virtual run of destination p: [VSTART[p], VSTART[p+1]) = m_p leading pad + its rows, rounded up to 32,
        // where m_p = (first output row of the run) mod 32
        {
            uint32_t carry = 0;
            for (uint32_t p0 = 0; p0 < N; p0 += THREADS) {
                uint32_t p = p0 + threadIdx.x;
                uint32_t len = 0;
                if (p < N) {
                    uint32_t ts = TILE_START[p], cnt = TILE_START[p + 1] - ts;
                    uint32_t m = (uint32_t)((int64_t)ts + DELTA[p]) & 31u;
                    len = cnt ? (m + cnt + 31u) & ~31u : 0u;
                }
                uint32_t tot;
                uint32_t ex = block_exclusive_scan<THREADS>(len, S_SCAN, tot);
                if (p < N) VSTART[p] = carry + ex;
                carry += tot;
            }
            if (threadIdx.x == 0) VSTART[N] = carry;
            __syncthreads();
        }
        const uint32_t vtotal = VSTART[N];
#pragma unroll
        for (int k = 0; k < KV; ++k) {
            uint32_t vs = k * THREADS + threadIdx.x;
            slot[k] = SLOT_NONE;
            if (vs < vtotal) {
                uint32_t lo = 0, hi = N;  // last p with VSTART[p] <= vs
                while (hi - lo > 1) {
                    uint32_t mid = (lo + hi) >> 1;
                    if (VSTART[mid] <= vs) lo = mid; else hi = mid;
                }
                uint32_t ts = TILE_START[lo], cnt = TILE_START[lo + 1] - ts;
                uint32_t m = (uint32_t)((int64_t)ts + DELTA[lo]) & 31u;
                uint32_t off = vs - VSTART[lo] - m;  // wraps for the leading pad
                if (off < cnt) slot[k] = (ts + off) | (lo << 16);
            }
        }
    }

    // ---- phase 2: every column through the staging buffer
#pragma unroll 1
    for (int c = 0; c < P.n_cols; ++c) {
        const PayloadCol& col = P.cols[c];
        if constexpr (std::is_same<V, BitColumn>::value) {
            scatter_bit_column<THREADS, K, KV>(col, stage, row0, tile_rows, pos, slot, DELTA, t0);
        } else {
            if (PEER) {
                // (the previous column's write-out reads OUT_BASE: the barrier inside
                //  scatter_fixed_column orders this rewrite after it only for the staging
                //  buffer, so fence explicitly)
                __syncthreads();
                for (uint32_t p = threadIdx.x; p < N; p += THREADS)
                    OUT_BASE[p] = (char*)P.peer_base[p / P.parts_per_rank] + (size_t)col.out;
            }
            scatter_fixed_column<THREADS, K, KV, V, (sizeof(V) == 16 && K % 2 == 0 ? K / 2 : K)>(col, stage, row0, tile_rows, pos, slot, DELTA, t0,
                                                                                           PEER ? OUT_BASE : nullptr);
        }
    }
#undef OUT_BASE
#undef VSTART
#undef DELTA
#undef WARP_CNT
#undef TILE_START
#undef S_SCAN
#undef S_MISC
}

// ---------------------------------------------------------------------------
// K4: variable-width payload columns (Utf8 / LargeUtf8 / Binary).
// K2 scatters an iota column, giving src[j] = input row of output row j.  Per
// var-width column: gather the string lengths in output order, exclusive-scan
// them into the output offsets (3-phase device scan), then copy the bytes.
// ---------------------------------------------------------------------------
constexpr int VAR_BLOCK = 256;
constexpr int VAR_ITEMS = 8;  // rows per thread in the scan kernels (block = 2048 rows)

__global__ void k_iota_u32(uint32_t* __restrict__ out, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) out[i] = (uint32_t)i;
}

template <typename OFF>
__device__ __forceinline__ unsigned long long var_len(const OFF* __restrict__ in_off, int64_t in_offset, uint32_t src) {
    const int64_t j = (int64_t)src + in_offset;
    return (unsigned long long)(in_off[j + 1] - in_off[j]);
}

// phase a: per-block sum of the gathered lengths
template <typename OFF>
__global__ void __launch_bounds__(VAR_BLOCK) k_var_block_sums(const OFF* __restrict__ in_off, int64_t in_offset,
                                                               const uint32_t* __restrict__ src, int64_t n,
                                                               unsigned long long* __restrict__ block_sums) {
    __shared__ unsigned long long s_warp[VAR_BLOCK / 32];
    const int64_t base = (int64_t)blockIdx.x * (VAR_BLOCK * VAR_ITEMS) + (int64_t)threadIdx.x * VAR_ITEMS;
    unsigned long long sum = 0;
#pragma unroll
    for (int k = 0; k < VAR_ITEMS; ++k)
        if (base + k < n) sum += var_len(in_off, in_offset, src[base + k]);
#pragma unroll
    for (int sh = 16; sh >= 1; sh >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, sh);
    if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t = 0;
        for (int w = 0; w < VAR_BLOCK / 32; ++w) t += s_warp[w];
        block_sums[blockIdx.x] = t;
    }
}

// phase b: exclusive scan of the block sums in place (single CTA), total -> block_sums[n_blocks]
__global__ void __launch_bounds__(1024) k_var_scan_block_sums(unsigned long long* __restrict__ block_sums, int64_t n_blocks) {
    __shared__ unsigned long long s_warp[33];
    unsigned long long carry = 0;
    for (int64_t b0 = 0; b0 < n_blocks; b0 += 1024) {
        int64_t b = b0 + threadIdx.x;
        unsigned long long v = b < n_blocks ? block_sums[b] : 0, tot;
        unsigned long long ex = block_exclusive_scan_u64<1024>(v, s_warp, tot);
        if (b < n_blocks) block_sums[b] = carry + ex;
        carry += tot;
    }
    if (threadIdx.x == 0) block_sums[n_blocks] = carry;
}

// phase c: output offsets = block base + block-local exclusive scan of the lengths
template <typename OFF>
__global__ void __launch_bounds__(VAR_BLOCK) k_var_write_offsets(const OFF* __restrict__ in_off, int64_t in_offset,
                                                                  const uint32_t* __restrict__ src, int64_t n,
                                                                  const unsigned long long* __restrict__ block_sums,
                                                                  OFF* __restrict__ out_off) {
    __shared__ unsigned long long s_warp[33];
    const int64_t base = (int64_t)blockIdx.x * (VAR_BLOCK * VAR_ITEMS) + (int64_t)threadIdx.x * VAR_ITEMS;
    unsigned long long len[VAR_ITEMS], sum = 0;
#pragma unroll
    for (int k = 0; k < VAR_ITEMS; ++k) {
        len[k] = base + k < n ? var_len(in_off, in_offset, src[base + k]) : 0;
        sum += len[k];
    }
    unsigned long long tot;
    unsigned long long run = block_sums[blockIdx.x] + block_exclusive_scan_u64<VAR_BLOCK>(sum, s_warp, tot);
#pragma unroll
    for (int k = 0; k < VAR_ITEMS; ++k) {
        if (base + k < n) out_off[base + k] = (OFF)run;
        run += len[k];
    }
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) out_off[n] = (OFF)block_sums[gridDim.x];
}

// bytes of output row j <- bytes of input row src[j]; one thread per row
template <typename OFF>
__global__ void __launch_bounds__(256) k_var_copy_bytes(const OFF* __restrict__ in_off, int64_t in_offset,
                                                         const uint8_t* __restrict__ in_data, const uint32_t* __restrict__ src,
                                                         const OFF* __restrict__ out_off, uint8_t* __restrict__ out_data, int64_t n) {
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = (int64_t)src[j] + in_offset;
        const uint8_t* s = in_data + in_off[r];
        const int64_t len = (int64_t)(in_off[r + 1] - in_off[r]);
        uint8_t* d = out_data + out_off[j];
        int64_t i = 0;
        if ((((uintptr_t)s ^ (uintptr_t)d) & 7) == 0) {  // co-aligned: byte head, 8-byte body
            for (; i < len && ((uintptr_t)(s + i) & 7); ++i) d[i] = s[i];
            for (; i + 8 <= len; i += 8) *(uint64_t*)(d + i) = *(const uint64_t*)(s + i);
        }
        for (; i < len; ++i) d[i] = s[i];
    }
}

// must mirror the offsets computed at the top of k_scatter (the peer / aligned tables are last,
// so launches that do not use them simply do not allocate them)
// ---------------------------------------------------------------------------
// Exchange helpers for bit-packed and variable-width columns (NCCL mode): bitmaps travel as one
// byte per row, strings as (lengths, bytes); the receiver rebuilds bitmaps and offsets.
// ---------------------------------------------------------------------------
__global__ void k_bits_to_bytes(const uint8_t* __restrict__ bits, int64_t bit_offset, int64_t n, uint8_t* __restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = bit_is_set(bits, i + bit_offset) ? 1 : 0;
}

// out bitmap words are fully written (n rounded up to 32 rows per warp): no pre-zeroing, no atomics
__global__ void k_bytes_to_bits(const uint8_t* __restrict__ in, int64_t n, unsigned* __restrict__ out_words) {
    const int64_t n32 = (n + 31) & ~(int64_t)31;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n32; i += (int64_t)gridDim.x * blockDim.x) {
        unsigned b = __ballot_sync(0xffffffffu, i < n && in[i] != 0);
        if ((threadIdx.x & 31) == 0) out_words[i >> 5] = b;
    }
}

template <typename OFF>
__global__ void k_offsets_to_lengths(const OFF* __restrict__ off, int64_t n, OFF* __restrict__ len) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) len[i] = off[i + 1] - off[i];
}

// bytes[g] / first[g] of every destination's run of a destination-sorted var-width column
template <typename OFF>
__global__ void k_var_dest_bytes(const OFF* __restrict__ off, const int64_t* __restrict__ part_starts, uint32_t N,
                                 int64_t* __restrict__ bytes, int64_t* __restrict__ first) {
    for (uint32_t g = blockIdx.x * blockDim.x + threadIdx.x; g < N; g += gridDim.x * blockDim.x) {
        const int64_t a = (int64_t)off[part_starts[g]], b = (int64_t)off[part_starts[g + 1]];
        bytes[g] = b - a;
        first[g] = a;
    }
}

// lengths -> exclusive offsets, same 3-phase scan as K4 (phase b is k_var_scan_block_sums)
template <typename OFF>
__global__ void __launch_bounds__(VAR_BLOCK) k_len_block_sums(const OFF* __restrict__ len, int64_t n, unsigned long long* __restrict__ block_sums) {
    __shared__ unsigned long long s_warp[VAR_BLOCK / 32];
    const int64_t base = (int64_t)blockIdx.x * (VAR_BLOCK * VAR_ITEMS) + (int64_t)threadIdx.x * VAR_ITEMS;
    unsigned long long sum = 0;
#pragma unroll
    for (int k = 0; k < VAR_ITEMS; ++k)
        if (base + k < n) sum += (unsigned long long)len[base + k];
#pragma unroll
    for (int sh = 16; sh >= 1; sh >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, sh);
    if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t = 0;
        for (int w = 0; w < VAR_BLOCK / 32; ++w) t += s_warp[w];
        block_sums[blockIdx.x] = t;
    }
}

template <typename OFF>
__global__ void __launch_bounds__(VAR_BLOCK) k_len_write_offsets(const OFF* __restrict__ len, int64_t n,
                                                                  const unsigned long long* __restrict__ block_sums, OFF* __restrict__ out_off) {
    __shared__ unsigned long long s_warp[33];
    const int64_t base = (int64_t)blockIdx.x * (VAR_BLOCK * VAR_ITEMS) + (int64_t)threadIdx.x * VAR_ITEMS;
    unsigned long long l[VAR_ITEMS], sum = 0;
#pragma unroll
    for (int k = 0; k < VAR_ITEMS; ++k) {
        l[k] = base + k < n ? (unsigned long long)len[base + k] : 0;
        sum += l[k];
    }
    unsigned long long tot;
    unsigned long long run = block_sums[blockIdx.x] + block_exclusive_scan_u64<VAR_BLOCK>(sum, s_warp, tot);
#pragma unroll
    for (int k = 0; k < VAR_ITEMS; ++k) {
        if (base + k < n) out_off[base + k] = (OFF)run;
        run += l[k];
    }
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) out_off[n] = (OFF)block_sums[gridDim.x];
}

template <int THREADS, int K>
inline size_t scatter_smem_bytes(uint32_t N, int stage_width, bool peer, bool aligned) {
    size_t off = ((size_t)THREADS * K * stage_width + 15) & ~(size_t)15;
    off += (size_t)N * 8;
    off += (size_t)(THREADS / 32) * N * 4;
    off += (size_t)(N + 1) * 4;
    off += (size_t)(THREADS / 32 + 1) * 4;
    off += 2 * 4;  // misc: tile ticket, tile overflow (single-pass mode)
    off = (off + 7) & ~(size_t)7;
    if (peer || aligned) off += (size_t)N * 8;  // per-destination output bases (peer mode)
    if (aligned) off += (size_t)(N + 1) * 4;    // virtual run starts (aligned mode)
    return off;
}

}  // namespace dfd
