// dfd_types.cuh — plain structs shared by the kernels and the host-side launch code.
#pragma once
#include <cstdint>

#include "dfd_hash.cuh"

namespace dfd {

constexpr int MAX_COLS_PER_LAUNCH = 24;
constexpr uint32_t MAX_PARTITIONS = 4096;
constexpr int MAX_RANKS = 16;

struct PayloadCol {
    const void* in;          // values (fixed) or bitmap (bool / validity pass)
    void* out;               // local mode: output buffer; peer mode: byte offset into every rank's receive window
    int64_t in_offset;       // Arrow logical offset of the input (rows)
    int32_t width;           // bytes; 0 => bit column (bool values or validity)
    int32_t pad;
};

struct ScatterParams {
    KeySet keys;
    HashState st;
    ModN mod;
    int64_t n_rows;
    int64_t n_tiles;
    const uint32_t* hist;        // [N][n_tiles] per-tile destination counts (K1)
    const uint32_t* tile_base;   // [N][n_tiles] exclusive scan of hist along tiles (rows < 2^32 per call)
    const int64_t* dest_base;    // [N] first output row of destination p for THIS producer:
                                 //   local mode: part_starts[p]; peer mode: row inside the owner's receive window
    PayloadCol cols[MAX_COLS_PER_LAUNCH];
    int32_t n_cols;
    uint32_t N;
    int32_t stage_width;         // widest staged element (bytes)
    uint32_t parts_per_rank;     // peer mode: destination p lives on rank p / parts_per_rank
    void* peer_base[MAX_RANKS];  // peer mode: every rank's receive window (CUDA-IPC mapped, NVLink)
    const int32_t* abort_flag;   // peer mode: non-zero => a receive window would overflow; do nothing
    // ---- single-pass mode (k_scatter<..., ONEPASS>): no K1/K1b; tile cursors by decoupled look-back ----
    unsigned long long* lb_desc; // [N][n_tiles] look-back descriptors: (epoch<<2 | state) << 32 | rows
    unsigned* lb_ticket;         // [0] next tile ticket, [1] finished CTAs (both reset by the last CTA)
    uint32_t lb_epoch;           // call epoch (30 bits): descriptors of older calls read as "not published"
    int32_t rank;                // peer mode: this producer's task index
    int32_t world;               // peer mode: number of workers
    int64_t region_stride;       // rows per destination region when dest_base == nullptr:
                                 //   local: region p starts at row p*stride; peer: (p % parts_per_rank)*world + rank
    const int64_t* dest_cap;     // [N] rows each region can hold (nullptr: region_stride)
    int64_t* totals_out;         // [N] rows per destination (written by the last tile)
    int32_t* overflow_out;       // set to 1 when a region is too small (that tile writes nothing)
    uint32_t* hist_out;          // optional [N][n_tiles]: per-tile counts / cursors for follow-up launches
    uint32_t* base_out;          //   (other column widths) that run the two-pass k_scatter code path
    // peer mode: "window free" flags in THIS worker's window header, ready_flags[o] >= ready_epoch once consumer o has
    // finished reading the previous shuffle's rows (checked by every CTA before its first store to a peer)
    const unsigned long long* ready_flags;
    unsigned long long ready_epoch;
    // two-pass mode, non-trivial keys (strings, several keys, nullable keys): K1 leaves every row's destination here so that
    // the K2 launches (one per column width) do not hash the keys again
    const uint16_t* dest_cache;
};

// Header at the start of every worker's receive window (peer-memory flags of the single-pass exchange; no NCCL on the
// critical path).  ready[o]  : written by consumer o into every PRODUCER's header: "my window may be overwritten, epoch e"
//                   done[r]   : written by producer r into every CONSUMER's header: "my rows of shuffle e have landed"
//                   counts[r][q], overflow[r] : what producer r sent to this consumer's partition q / whether it overflowed
constexpr uint32_t XCHG_MAX_P = 256;
constexpr uint32_t XCHG_META_MAX = 2048;  // push transport: int64 metadata entries per producer ((1 + var columns) x N)
constexpr size_t XCHG_HEADER_BYTES = 320 * 1024;
struct ExchangeHeader {
    unsigned long long ready[MAX_RANKS];
    unsigned long long done[MAX_RANKS];
    int overflow[MAX_RANKS];
    long long counts[MAX_RANKS][XCHG_MAX_P];
    // push transport (all column kinds): meta_flag[r] = e once producer r's row / byte counts of shuffle e are in meta[r][]
    // (it is sent only after r's stream has finished reading r's own window, so it doubles as the "window free" signal)
    unsigned long long meta_flag[MAX_RANKS];
    long long meta[MAX_RANKS][XCHG_META_MAX];
};
static_assert(sizeof(ExchangeHeader) <= XCHG_HEADER_BYTES, "exchange header must fit its reservation");

}  // namespace dfd
