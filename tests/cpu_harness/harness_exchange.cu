// TEST INFRASTRUCTURE (CPU suite only): what csrc/dfd_exchange.cu needs besides the stand-in runtime to run its push
// transport (every column kind; shuffle / coalesce / broadcast routes; back-pressure rounds) with THREADS as workers:
//   * the PartitionJob stages and the small conversion launches, restated with the CPU oracle;
//   * CPU emulations of the exchange kernels that path launches (k_slice_rows, k_xchg_allgather_meta, k_push_runs,
//     k_xchg_done_barrier), registered by name with the stand-in runtime's cudaLaunchKernel.
// and to run the single-pass exchange (fixed-width non-null schemas) with its overflow -> exact two-pass re-run:
//   * PartitionJob::run_onepass<PEER> / run_scatter<PEER> as row loops that store into the owners' windows;
//   * k_xchg_signal_ready, k_xchg_publish_wait, k_exchange_plan.
// The NCCL-mode transport needs nothing more here: its sends / receives go through the stand-in NCCL's mailboxes (fake_nccl.cpp).
#include <chrono>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "dfd_b200.h"
#include "dfd_internal.h"

int harness_partition(dfd_partitioner* p, const dfd_column* in, int n_cols, int64_t n, const dfd_column* out);
int harness_destinations(dfd_partitioner* p, const dfd_column* in, int n_cols, int64_t n, std::vector<uint64_t>* hashes);

namespace {
inline bool bit(const uint8_t* b, int64_t i) { return (b[i >> 3] >> (i & 7)) & 1; }
}  // namespace

// ---- PartitionJob: local form (push transport) and peer form (single-pass / two-pass fused exchange) ---------------------
namespace {
thread_local std::vector<uint32_t> t_dest;  // destination of every row of the job this worker (= thread) is running
bool wait_flags(const unsigned long long* flags, int world, unsigned long long epoch);

int destinations_of(dfd::PartitionJob& job) {
    std::vector<dfd_column> in;
    for (const dfd::PartitionJob::VarCol& v : job.var_cols) in.push_back(v.in);
    std::vector<uint64_t> h;
    if (int rc = harness_destinations(job.p, in.data(), (int)in.size(), job.n_rows, &h)) return rc;
    t_dest.resize((size_t)job.n_rows);
    for (int64_t i = 0; i < job.n_rows; ++i) t_dest[(size_t)i] = (uint32_t)(h[(size_t)i] % job.p->N);
    return DFD_OK;
}
}  // namespace

int dfd::PartitionJob::prepare(Partitioner* part, const dfd_column* in_cols, int n_cols, int64_t rows, const dfd_column* out_cols, bool peer_mode, cudaStream_t st) {
    p = part;
    stream = st;
    peer = peer_mode;
    n_rows = rows;
    var_cols.clear();
    for (int i = 0; i < n_cols; ++i) {
        if (peer_mode && (in_cols[i].kind != DFD_COL_FIXED || in_cols[i].validity)) return set_error(DFD_ERR_UNSUPPORTED, "column %d: peer stores move fixed-width non-null columns", i);
        var_cols.push_back(VarCol{in_cols[i], out_cols[i]});  // (all columns: the stand-in needs nothing else)
    }
    int rc = p->ctx->scratch.ensure(sizeof(int64_t) * (size_t)p->N + 256, p->ctx->device);
    if (rc) return rc;
    d_totals = (int64_t*)p->ctx->scratch.ptr;
    return DFD_OK;
}

// K1 + K1b: per-destination totals (the exchange all-gathers them in the two-pass fused path)
int dfd::PartitionJob::run_hist_scan() {
    if (int rc = destinations_of(*this)) return rc;
    memset(d_totals, 0, sizeof(int64_t) * (size_t)p->N);
    for (uint32_t g : t_dest) d_totals[g]++;
    return DFD_OK;
}

// K2.  Local: dense destination-sorted output + part_starts.  Peer: row k of destination g goes to row dest_base[g] + k of
// the window slot of g's owner (out.values is the column's byte offset inside every slot); nothing is written once the plan
// kernel has raised the abort flag.
int dfd::PartitionJob::run_scatter(const int64_t* dest_base, void* const* peer_base, int, uint32_t parts_per_rank, const int32_t* abort_flag) {
    std::vector<dfd_column> in, out;
    for (const VarCol& v : var_cols) { in.push_back(v.in); out.push_back(v.out); }
    if (!peer_base) return harness_partition(p, in.data(), (int)in.size(), n_rows, out.data());  // writes p->d_part_starts like K1b
    if (abort_flag && *abort_flag) return DFD_OK;
    std::vector<int64_t> cursor(p->N, 0);
    for (int64_t i = 0; i < n_rows; ++i) {
        const uint32_t g = t_dest[(size_t)i];
        const int64_t row = dest_base[g] + cursor[g]++;
        char* base = (char*)peer_base[g / parts_per_rank];
        for (size_t c = 0; c < in.size(); ++c) {
            const size_t w = (size_t)in[c].width;
            memcpy(base + (size_t)out[c].values + (size_t)row * w, (const char*)in[c].values + (size_t)(i + in[c].offset) * w, w);
        }
    }
    return DFD_OK;
}

// k_scatter_onepass<PEER>: every (partition q of consumer o, producer `rank`) pair owns rows [(q T + rank) sub_cap, + sub_cap)
// of every column of o's slot; rows beyond a sub-window are dropped and flagged (the counts stay exact), and no store is
// issued before every consumer has released its window for this epoch.
int dfd::PartitionJob::run_onepass(const OnePassLayout& L) {
    if (!L.peer_base) return set_error(DFD_ERR_UNSUPPORTED, "harness: only the peer form of the single-pass scatter is emulated");
    if (L.ready_flags && !wait_flags(L.ready_flags, L.world, L.ready_epoch)) return set_error(DFD_ERR_INTERNAL, "harness: a consumer never released its window");
    if (int rc = destinations_of(*this)) return rc;
    const uint32_t N = p->N, P = L.parts_per_rank;
    std::vector<int64_t> count(N, 0);
    bool overflow = false;
    for (int64_t i = 0; i < n_rows; ++i) {
        const uint32_t g = t_dest[(size_t)i];
        const int64_t k = count[g]++;
        if (k >= L.region_stride) { overflow = true; continue; }
        const int64_t row = ((int64_t)(g % P) * L.world + L.rank) * L.region_stride + k;
        char* base = (char*)L.peer_base[g / P];
        for (const VarCol& v : var_cols) {
            const size_t w = (size_t)v.in.width;
            memcpy(base + (size_t)v.out.values + (size_t)row * w, (const char*)v.in.values + (size_t)(i + v.in.offset) * w, w);
        }
    }
    memcpy(L.d_totals, count.data(), sizeof(int64_t) * (size_t)N);
    if (overflow) *L.d_overflow = 1;
    return DFD_OK;
}

int dfd::launch_bits_to_bytes(const uint8_t* bits, int64_t bit_offset, int64_t n, uint8_t* out, cudaStream_t) {
    for (int64_t i = 0; i < n; ++i) out[i] = bit(bits, i + bit_offset) ? 1 : 0;
    return DFD_OK;
}
int dfd::launch_offsets_to_lengths(const void* off, int ow, int64_t n, void* len, cudaStream_t) {
    for (int64_t i = 0; i < n; ++i) {
        if (ow == 8) ((int64_t*)len)[i] = ((const int64_t*)off)[i + 1] - ((const int64_t*)off)[i];
        else ((int32_t*)len)[i] = ((const int32_t*)off)[i + 1] - ((const int32_t*)off)[i];
    }
    return DFD_OK;
}
int dfd::launch_var_dest_bytes(const void* off, int ow, const int64_t* part_starts, uint32_t N, int64_t* bytes, int64_t* first, cudaStream_t) {
    for (uint32_t g = 0; g < N; ++g) {
        const int64_t a = ow == 8 ? ((const int64_t*)off)[part_starts[g]] : (int64_t)((const int32_t*)off)[part_starts[g]];
        const int64_t b = ow == 8 ? ((const int64_t*)off)[part_starts[g + 1]] : (int64_t)((const int32_t*)off)[part_starts[g + 1]];
        bytes[g] = b - a;
        first[g] = a;
    }
    return DFD_OK;
}

// ---- CPU emulations of the exchange kernels of the push transport ------------------------------------------------------
namespace {
struct FakeDim3 { unsigned x, y, z; };
typedef int (*HarnessKernel)(FakeDim3 grid, FakeDim3 block, void** args);
extern "C" void harness_register_kernel(const char* name_fragment, HarnessKernel fn);

template <typename T>
T arg(void** args, int i) { T v; memcpy(&v, args[i], sizeof v); return v; }

inline void store_release(unsigned long long* p, unsigned long long v) { __atomic_store_n(p, v, __ATOMIC_RELEASE); }
inline unsigned long long load_acquire(const unsigned long long* p) { return __atomic_load_n(p, __ATOMIC_ACQUIRE); }

// wait until flags[t] >= epoch for every t < world; bounded like the kernels' clock64 bound (HARNESS_FLAG_TIMEOUT_MS, default 20 s)
bool wait_flags(const unsigned long long* flags, int world, unsigned long long epoch) {
    static const long timeout_ms = getenv("HARNESS_FLAG_TIMEOUT_MS") ? atol(getenv("HARNESS_FLAG_TIMEOUT_MS")) : 20000;
    const auto t0 = std::chrono::steady_clock::now();
    for (int t = 0; t < world; ++t)
        while (load_acquire(&flags[t]) < epoch) {
            if (std::chrono::steady_clock::now() - t0 > std::chrono::milliseconds(timeout_ms)) return false;
            __builtin_ia32_pause();
        }
    return true;
}

// k_slice_rows(const int64_t* starts, uint32_t n, int64_t* rows)
int cpu_slice_rows(FakeDim3, FakeDim3, void** a) {
    const int64_t* starts = arg<const int64_t*>(a, 0);
    const uint32_t n = arg<uint32_t>(a, 1);
    int64_t* rows = arg<int64_t*>(a, 2);
    for (uint32_t g = 0; g < n; ++g) rows[g] = starts[g + 1] - starts[g];
    return 0;
}

// k_xchg_allgather_meta(local, peer_hdr, rank, world, epoch, my_meta, n_meta, timed_out)
int cpu_allgather_meta(FakeDim3, FakeDim3, void** a) {
    dfd::ExchangeHeader* local = arg<dfd::ExchangeHeader*>(a, 0);
    dfd::ExchangeHeader* const* peer = arg<dfd::ExchangeHeader* const*>(a, 1);
    const int rank = arg<int>(a, 2), world = arg<int>(a, 3);
    const unsigned long long epoch = arg<unsigned long long>(a, 4);
    const int64_t* my_meta = arg<const int64_t*>(a, 5);
    const uint32_t n_meta = arg<uint32_t>(a, 6);
    int32_t* timed_out = arg<int32_t*>(a, 7);
    for (int o = 0; o < world; ++o)
        for (uint32_t i = 0; i < n_meta; ++i) peer[o]->meta[rank][i] = my_meta[i];
    for (int o = 0; o < world; ++o) store_release(&peer[o]->meta_flag[rank], epoch);
    if (!wait_flags(local->meta_flag, world, epoch)) *timed_out = 1;
    return 0;
}

// k_xchg_done_barrier(local, peer_hdr, rank, world, epoch, timed_out)
int cpu_done_barrier(FakeDim3, FakeDim3, void** a) {
    dfd::ExchangeHeader* local = arg<dfd::ExchangeHeader*>(a, 0);
    dfd::ExchangeHeader* const* peer = arg<dfd::ExchangeHeader* const*>(a, 1);
    const int rank = arg<int>(a, 2), world = arg<int>(a, 3);
    const unsigned long long epoch = arg<unsigned long long>(a, 4);
    int32_t* timed_out = arg<int32_t*>(a, 5);
    for (int o = 0; o < world; ++o) store_release(&peer[o]->done[rank], epoch);
    if (!wait_flags(local->done, world, epoch)) *timed_out = 1;
    return 0;
}

// the run descriptor of k_push_runs — a COPY of the private struct in csrc/dfd_exchange.cu (same field order and types;
// a mismatch shows up at once as garbage segments in every test of tests/test_exchange_cpu_harness.py)
struct PushRun {
    const char* src;
    char* dst;
    long long n;
    long long a, b;
    int kind;
    int first_block;
};
enum { RUN_BYTES = 0, RUN_BITS = 1, RUN_OFF32 = 2, RUN_OFF64 = 3, RUN_ONES = 4 };

// k_push_runs(const PushRun* runs, int n_runs)
int cpu_push_runs(FakeDim3, FakeDim3, void** a) {
    const PushRun* runs = arg<const PushRun*>(a, 0);
    const int n_runs = arg<int>(a, 1);
    for (int i = 0; i < n_runs; ++i) {
        const PushRun& r = runs[i];
        switch (r.kind) {
            case RUN_BYTES: if (r.n) memmove(r.dst, r.src, (size_t)r.n); break;
            case RUN_BITS: {
                const size_t words = (size_t)((r.n + 31) / 32);
                memset(r.dst, 0, words * 4);
                for (long long k = 0; k < r.n; ++k)
                    if (bit((const uint8_t*)r.src, r.a + k)) r.dst[k >> 3] = (char)(r.dst[k >> 3] | (1 << (k & 7)));
                break;
            }
            case RUN_ONES: memset(r.dst, 0xff, (size_t)((r.n + 31) / 32) * 4); break;
            case RUN_OFF32: for (long long k = 0; k < r.n; ++k) ((int*)r.dst)[k] = (int)((long long)((const int*)r.src)[k] - r.a + r.b); break;
            default: for (long long k = 0; k < r.n; ++k) ((long long*)r.dst)[k] = ((const long long*)r.src)[k] - r.a + r.b; break;
        }
    }
    return 0;
}

// k_xchg_signal_ready(peer_hdr, rank, world, epoch): "my window is free" into every producer's header
int cpu_signal_ready(FakeDim3, FakeDim3, void** a) {
    dfd::ExchangeHeader* const* peer = arg<dfd::ExchangeHeader* const*>(a, 0);
    const int rank = arg<int>(a, 1), world = arg<int>(a, 2);
    const unsigned long long epoch = arg<unsigned long long>(a, 3);
    for (int o = 0; o < world; ++o) store_release(&peer[o]->ready[rank], epoch);
    return 0;
}

// k_xchg_publish_wait(local, peer_hdr, rank, world, P, epoch, totals, overflow, timed_out)
int cpu_publish_wait(FakeDim3, FakeDim3, void** a) {
    dfd::ExchangeHeader* local = arg<dfd::ExchangeHeader*>(a, 0);
    dfd::ExchangeHeader* const* peer = arg<dfd::ExchangeHeader* const*>(a, 1);
    const int rank = arg<int>(a, 2), world = arg<int>(a, 3);
    const uint32_t P = arg<uint32_t>(a, 4);
    const unsigned long long epoch = arg<unsigned long long>(a, 5);
    const int64_t* totals = arg<const int64_t*>(a, 6);
    const int32_t* overflow = arg<const int32_t*>(a, 7);
    int32_t* timed_out = arg<int32_t*>(a, 8);
    for (uint32_t g = 0; g < P * (uint32_t)world; ++g) peer[g / P]->counts[rank][g % P] = totals[g];
    for (int o = 0; o < world; ++o) peer[o]->overflow[rank] = *overflow;
    for (int o = 0; o < world; ++o) store_release(&peer[o]->done[rank], epoch);
    if (!wait_flags(local->done, world, epoch)) *timed_out = 1;
    return 0;
}

// k_exchange_plan(counts[T][N], world, P, rank, capacity_rows, dest_base[N], my_part_starts[P+1], abort_flag)
int cpu_exchange_plan(FakeDim3, FakeDim3, void** a) {
    const int64_t* counts = arg<const int64_t*>(a, 0);
    const int world = arg<int>(a, 1);
    const uint32_t P = arg<uint32_t>(a, 2);
    const int rank = arg<int>(a, 3);
    const int64_t capacity_rows = arg<int64_t>(a, 4);
    int64_t* dest_base = arg<int64_t*>(a, 5);
    int64_t* my_part_starts = arg<int64_t*>(a, 6);
    int32_t* abort_flag = arg<int32_t*>(a, 7);
    const uint32_t N = P * (uint32_t)world;
    int overflow = 0;
    for (int o = 0; o < world; ++o) {
        int64_t run = 0;
        for (uint32_t q = 0; q < P; ++q) {
            const uint32_t g = (uint32_t)o * P + q;
            if (o == rank) my_part_starts[q] = run;
            int64_t before_me = 0, tot = 0;
            for (int r = 0; r < world; ++r) {
                const int64_t c = counts[(int64_t)r * N + g];
                if (r < rank) before_me += c;
                tot += c;
            }
            dest_base[g] = run + before_me;
            run += tot;
        }
        if (o == rank) my_part_starts[P] = run;
        if (run > capacity_rows) overflow = 1;
    }
    *abort_flag = overflow;
    return 0;
}

__attribute__((constructor)) void register_exchange_kernels() {
    harness_register_kernel("k_xchg_signal_ready", cpu_signal_ready);
    harness_register_kernel("k_xchg_publish_wait", cpu_publish_wait);
    harness_register_kernel("k_exchange_plan", cpu_exchange_plan);
    harness_register_kernel("k_slice_rows", cpu_slice_rows);
    harness_register_kernel("k_xchg_allgather_meta", cpu_allgather_meta);
    harness_register_kernel("k_xchg_done_barrier", cpu_done_barrier);
    harness_register_kernel("k_push_runs", cpu_push_runs);
}
}  // namespace
