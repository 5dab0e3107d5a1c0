"""oracle/oracle_py.py — CPU ORACLE, pure-Python big-integer restatement.

TEST INFRASTRUCTURE, NOT PRODUCT CODE: only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline leg may import this.  PARITY STATUS: "parity unpinned"
(see oracle/df_oracle.h) — the upstream crates (ahash 0.8.12,
datafusion-common/physical-plan 53.0.0) are not on disk and cannot be built
here.  This file exists as an INDEPENDENT second restatement (Python ints have
exact 128-bit products) used to pin the C oracle and to generate
tests/golden/*.json (see tests/golden/make_golden.py).

What is restated, and from where:
  ahash 0.8.12 src/random_state.rs   PI2, RandomState::with_seeds
  ahash 0.8.12 src/fallback_hash.rs  AHasher::{from_random_state,update,large_update,write,finish}
  ahash 0.8.12 src/operations.rs     folded_multiply, read_small
  datafusion-common 53 hash_utils.rs create_hashes, combine_hashes, HashValue impls
  datafusion-physical-plan 53 repartition/mod.rs  BatchPartitioner::partition (Hash arm),
                                                  REPARTITION_RANDOM_STATE
Reference call sites: src/execution_plans/network_shuffle.rs:126-134,213-238;
src/worker/impl_execute_task.rs:77-86.
"""
from __future__ import annotations

M64 = (1 << 64) - 1
PI2 = (0x452821E638D01377, 0xBE5466CF34E90C6C, 0xC0AC29B7C97C50DD, 0x3F84D5B5B5470917)
MULTIPLE = 6364136223846793005
ROT = 23


def rotl(x: int, r: int) -> int:
    r &= 63
    return ((x << r) | (x >> (64 - r))) & M64 if r else x


def folded_multiply(s: int, by: int) -> int:
    r = s * by
    return (r & M64) ^ (r >> 64)


def with_seeds(s0=0, s1=0, s2=0, s3=0):
    return (s0 ^ PI2[0], s1 ^ PI2[1], s2 ^ PI2[2], s3 ^ PI2[3])


REPARTITION_RANDOM_STATE = with_seeds(0, 0, 0, 0)


class AHasher:
    def __init__(self, st=REPARTITION_RANDOM_STATE):
        k0, k1, k2, k3 = st
        self.buffer, self.pad, self.extra = k1, k0, (k2, k3)

    def update(self, x: int):
        self.buffer = folded_multiply((x ^ self.buffer) & M64, MULTIPLE)

    def large_update(self, lo: int, hi: int):
        combined = folded_multiply(lo ^ self.extra[0], hi ^ self.extra[1])
        self.buffer = rotl(((self.buffer + self.pad) & M64) ^ combined, ROT)

    def write(self, data: bytes):
        n = len(data)
        self.buffer = ((self.buffer + n) * MULTIPLE) & M64
        le = lambda b: int.from_bytes(b, "little")
        if n > 8:
            if n > 16:
                self.large_update(le(data[n - 16:n - 8]), le(data[n - 8:]))
                while len(data) > 16:
                    self.large_update(le(data[:8]), le(data[8:16]))
                    data = data[16:]
            else:
                self.large_update(le(data[:8]), le(data[n - 8:]))
        else:
            if n >= 2:
                if n >= 4:
                    a, b = le(data[:4]), le(data[n - 4:])
                else:
                    a, b = le(data[:2]), data[n - 1]
            elif n > 0:
                a = b = data[0]
            else:
                a = b = 0
            self.large_update(a, b)

    def finish(self) -> int:
        return rotl(folded_multiply(self.buffer, self.pad), self.buffer & 63)


def hash_one_int(x: int, width: int = 8, st=REPARTITION_RANDOM_STATE) -> int:
    """hash_one of an integer of `width` bytes (two's complement, zero-extended)."""
    h = AHasher(st)
    x &= (1 << (8 * width)) - 1
    if width == 16:
        h.large_update(x & M64, x >> 64)
    else:
        h.update(x)
    return h.finish()


def hash_one_interval_day_time(days: int, millis: int, st=REPARTITION_RANDOM_STATE) -> int:
    """arrow-buffer IntervalDayTime {days: i32, milliseconds: i32} #[derive(Hash)]: write_i32 per field
    (DataFusion hash_utils: hash_value!(.., IntervalDayTime, IntervalMonthDayNano) -> state.hash_one(self))."""
    h = AHasher(st)
    h.update(days & 0xFFFFFFFF)
    h.update(millis & 0xFFFFFFFF)
    return h.finish()


def hash_one_interval_month_day_nano(months: int, days: int, nanos: int, st=REPARTITION_RANDOM_STATE) -> int:
    """IntervalMonthDayNano {months: i32, days: i32, nanoseconds: i64} #[derive(Hash)]: write_i32, write_i32, write_i64."""
    h = AHasher(st)
    h.update(months & 0xFFFFFFFF)
    h.update(days & 0xFFFFFFFF)
    h.update(nanos & M64)
    return h.finish()


def hash_one_str(s: bytes, st=REPARTITION_RANDOM_STATE) -> int:
    h = AHasher(st)
    h.write(s)
    h.update(0xFF)
    return h.finish()


def hash_one_bytes(s: bytes, st=REPARTITION_RANDOM_STATE) -> int:
    h = AHasher(st)
    h.update(len(s))
    h.write(s)
    return h.finish()


def combine_hashes(l: int, r: int) -> int:
    return (((17 * 37 + l) & M64) * 37 + r) & M64


def create_hashes(columns, n_rows: int, st=REPARTITION_RANDOM_STATE):
    """columns: list of (kind, width, values) with values a list; None == null.
    kind in {"int", "str", "bytes", "bool"}."""
    hashes = [0] * n_rows
    for ci, (kind, width, values) in enumerate(columns):
        for i in range(n_rows):
            v = values[i]
            if v is None:
                continue
            if kind == "int":
                hv = hash_one_int(int(v), width, st)
            elif kind == "bool":
                hv = hash_one_int(1 if v else 0, 1, st)
            elif kind == "str":
                hv = hash_one_str(bytes(v), st)
            elif kind == "bytes":
                hv = hash_one_bytes(bytes(v), st)
            else:
                raise ValueError(kind)
            hashes[i] = combine_hashes(hv, hashes[i]) if ci >= 1 else hv
    return hashes


def partition(hashes, num_partitions: int):
    """BatchPartitioner Hash arm: per-destination row index lists in row order."""
    out = [[] for _ in range(num_partitions)]
    for i, h in enumerate(hashes):
        out[h % num_partitions].append(i)
    return out
