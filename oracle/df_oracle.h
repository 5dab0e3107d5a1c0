/*
 * oracle/df_oracle.h — CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 *
 * A plain-C restatement of the arithmetic on datafusion-distributed's
 * hash-repartition shuffle path.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it; the product
 * (datafusion_distributed_b200/) never links, imports or falls back to it.
 *
 * PARITY STATUS: **parity unpinned**.  The arithmetic restated here lives in
 * third-party crates that are NOT under /root/reference and cannot be built in
 * this image (no cargo/rustc, no vendored registry):
 *   - ahash 0.8.12          (reference Cargo.lock:32-33)   fallback hasher
 *   - datafusion-common 53.0.0 (Cargo.lock:1960-1985)       hash_utils::{create_hashes,combine_hashes}
 *   - datafusion-physical-plan 53.0.0 (Cargo.lock:2572-2574) repartition::{BatchPartitioner,RepartitionExec}
 *   - arrow-select 58.1.0   (Cargo.lock:429-430)            take
 * The reference holds no golden vector that pins partition placement
 * (SURVEY.md §8c); this oracle is pinned only against (a) an independent
 * pure-Python big-integer restatement (oracle/oracle_py.py) and (b) the
 * self-consistency vectors recorded in SURVEY.md §8(c).  The four places where
 * a real-ahash check could disagree are isolated as named constants/functions
 * below (ORC_PI2, orc_state_with_seeds, orc_hasher_from_state, orc_combine_hashes).
 *
 * Reference call sites of the restated dependency (the path this follows):
 *   src/execution_plans/network_shuffle.rs:126-134  (RepartitionExec scaled to P*T)
 *   src/worker/impl_execute_task.rs:77-86           (plan.execute(partition) on the producer)
 *   src/execution_plans/network_shuffle.rs:213-238  (consumer: off = P*task_index, stream off+partition)
 *   src/execution_plans/benchmarks/shuffle_bench.rs:203-211 (Hash([id], partitions*consumer_tasks))
 */
#ifndef DF_ORACLE_H
#define DF_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- ahash 0.8.12 fallback hasher (fallback_hash.rs / random_state.rs) ---- */

typedef struct {
    uint64_t k0, k1, k2, k3; /* RandomState keys after with_seeds() */
} orc_random_state;

/* RandomState::with_seeds(a,b,c,d): k_i = seed_i ^ PI2[i]  (random_state.rs) */
orc_random_state orc_state_with_seeds(uint64_t s0, uint64_t s1, uint64_t s2, uint64_t s3);

/* DataFusion REPARTITION_RANDOM_STATE = with_seeds(0,0,0,0) (repartition/mod.rs) */
orc_random_state orc_repartition_random_state(void);

/* hash_one for the integer widths: value is zero-extended from `width` bytes
 * (Hasher::write_{u8,u16,u32,u64} all call update(x as u64)). */
uint64_t orc_hash_one_u64(const orc_random_state* st, uint64_t x);
/* write_u128 -> large_update */
uint64_t orc_hash_one_u128(const orc_random_state* st, uint64_t lo, uint64_t hi);
/* impl Hash for str: write(bytes); write_u8(0xff) */
uint64_t orc_hash_one_str(const orc_random_state* st, const uint8_t* p, size_t len);
/* impl Hash for [u8]: write_usize(len); write(bytes) */
uint64_t orc_hash_one_bytes(const orc_random_state* st, const uint8_t* p, size_t len);

/* datafusion-common hash_utils::combine_hashes(l, r) */
uint64_t orc_combine_hashes(uint64_t l, uint64_t r);

/* ---- Arrow-shaped column descriptors (no Arrow dependency) ---- */

enum {
    ORC_FIXED = 0,   /* primitive, `width` bytes per value (1,2,4,8,16)          */
    ORC_BOOL = 1,    /* bit-packed boolean values                                 */
    ORC_UTF8 = 2,    /* int32 offsets + data (Utf8; hashed as str)                */
    ORC_LARGE_UTF8 = 3, /* int64 offsets                                          */
    ORC_BINARY = 4,  /* int32 offsets, hashed as [u8]                             */
    ORC_INTERVAL_DAY_TIME = 5,       /* 8 B {days: i32, milliseconds: i32}: derived Hash, one write_i32 per field   */
    ORC_INTERVAL_MONTH_DAY_NANO = 6, /* 16 B {months: i32, days: i32, nanoseconds: i64}: write_i32, write_i32, write_i64 */
};

typedef struct {
    int32_t kind;
    int32_t width;            /* ORC_FIXED only */
    const void* values;       /* fixed: values; bool: bitmap; utf8/binary: data bytes */
    const void* offsets;      /* utf8/binary */
    const uint8_t* validity;  /* Arrow validity bitmap (LSB first) or NULL */
    int64_t offset;           /* Arrow logical offset of the array */
} orc_column;

/* create_hashes(arrays, random_state, hashes_buffer): hashes[] must be zeroed
 * by the caller (BatchPartitioner does hash_buffer.resize(n, 0)).  Column 0
 * overwrites, column j>=1 combines; null rows leave the hash untouched. */
void orc_create_hashes(const orc_column* cols, int n_cols, int64_t n_rows,
                       const orc_random_state* st, uint64_t* hashes);

/* BatchPartitioner::partition (Hash arm): dest = hash % num_partitions;
 * row indices appended to their destination in row order.
 *   counts[num_partitions]  (out) rows per destination
 *   indices[n_rows]         (out) row indices grouped by destination, stable
 *   starts[num_partitions+1](out) start of each destination's run in indices */
void orc_partition_indices(const uint64_t* hashes, int64_t n_rows, uint32_t num_partitions,
                           int64_t* counts, uint32_t* indices, int64_t* starts);

/* arrow-select take for one fixed-width column: out[i] = values[idx[i]] */
void orc_take_fixed(const void* values, int width, const uint32_t* idx, int64_t n, void* out);

/* ---- Whole-operator restatement used as the timed CPU baseline ----
 *
 * RepartitionExec(Hash(keys, N)) over a table of `n_cols` fixed-width,
 * non-null columns presented as consecutive `batch_size`-row batches that are
 * dealt round-robin to `n_threads` input partitions (one worker thread per
 * input partition, like RepartitionExec's pull_from_input tasks).  Per batch:
 * create_hashes -> index vectors -> take per (partition, column) -> append to
 * that partition's coalescer (LimitedBatchCoalescer, target batch_size rows).
 * Output: for every destination the concatenation (in thread-major, then row
 * order) of everything sent to it, written into out_cols[c] at
 * out_starts[p] (rows).  Returns 0 on success.                              */
int orc_repartition_table(const void* const* cols, const int32_t* widths, int n_cols,
                          int64_t n_rows, const int32_t* key_cols, int n_keys,
                          uint32_t num_partitions, int64_t batch_size, int n_threads,
                          void* const* out_cols, int64_t* out_counts /*[N]*/,
                          int64_t* out_starts /*[N+1]*/);

/* ---- The same operator as a long-running worker runs it (timed CPU arm) ----
 * Persistent thread pool (threads created once; reference workers: tokio multi-thread runtime + mimalloc,
 * benchmarks/cdk/bin/worker.rs:32), per-thread reusable hash / index / take buffers and per-destination
 * coalescer batches; a completed output batch is handed to a consumer that drops it (its memory is reused).
 * Per batch the arithmetic and the copies are exactly those of orc_repartition_table.
 *   use_threads : input partitions == threads used for this call (<= pool size; 0 = all)
 *   out_counts[N], *out_batches (output batches emitted), *out_checksum (observable side effect of the copies) */
typedef struct orc_pool orc_pool;
orc_pool* orc_pool_create(int n_threads);
void orc_pool_destroy(orc_pool* pool);
int orc_pool_threads(const orc_pool* pool);
int orc_repartition_stream(orc_pool* pool, const void* const* cols, const int32_t* widths, int n_cols, int64_t n_rows,
                           const int32_t* key_cols, int n_keys, uint32_t num_partitions, int64_t batch_size, int use_threads,
                           int64_t* out_counts, int64_t* out_batches, uint64_t* out_checksum);

/* Lightweight version for big parity checks: destination id per row only. */
void orc_partition_ids(const orc_column* key_cols, int n_keys, int64_t n_rows,
                       uint32_t num_partitions, uint32_t* dest /*[n_rows]*/);

#ifdef __cplusplus
}
#endif
#endif
