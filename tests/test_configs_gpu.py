"""BASELINE.json configs[2..4] as parity-test cases (synthetic stand-ins per SURVEY.md §8d,
at sizes the oracle finishes in seconds): the CUDA partitioner vs the CPU oracle, bit-exact and
in order.  cfg-3: TPC-H q1 post-partial-aggregate shuffle (Utf8 x2 keys, aggregate-state columns).
cfg-4: TPC-H q5 join shuffles (Int64 keys, Decimal128 payload, 1- and 2-key).  cfg-5: ClickBench
GROUP BY UserID, SearchPhrase (Int64 + Utf8 keys, Zipf skew)."""
import numpy as np
import pyarrow as pa
import pytest

import datafusion_distributed_b200 as dfd
from oracle import oracle as orc
from tests.util import expected_partitions

pytestmark = pytest.mark.gpu


def check(ctx, arrays, key_cols, N):
    n = len(arrays[0])
    cols = [dfd.DeviceColumn.from_arrow(ctx, a) for a in arrays]
    part = dfd.HashPartitioner(ctx, dfd.Partitioning.Hash(key_cols, N))
    outs, starts = part.partition(cols, n)
    dest = orc.partition_ids([arrays[k] for k in key_cols], n, N)
    order, ref_starts = expected_partitions(dest, N)
    assert np.array_equal(starts, ref_starts)
    idx = pa.array(order)
    for c, arr in enumerate(arrays):
        assert outs[c].to_arrow(ctx, 0, n).equals(arr.take(idx)), (c, arr.type)
    return np.diff(starts)


def dec128(rng, n):
    raw = np.zeros((n, 2), dtype=np.int64)
    raw[:, 0] = rng.integers(0, 10**9, n)
    return pa.Array.from_buffers(pa.decimal128(15, 2), n, [None, pa.py_buffer(raw.tobytes())])


def test_cfg3_tpch_q1_group_by_shuffle(ctx):
    """Hash([l_returnflag, l_linestatus], 12): 4 groups x input partitions rows, 10 state columns."""
    rng = np.random.Generator(np.random.PCG64(3))
    groups = [("A", "F"), ("N", "F"), ("N", "O"), ("R", "F")]
    n = 4 * 24
    rf = pa.array([groups[i % 4][0] for i in range(n)], type=pa.string())
    ls = pa.array([groups[i % 4][1] for i in range(n)], type=pa.string())
    state = [dec128(rng, n) for _ in range(4)] + [pa.array(rng.integers(0, 10**6, n, dtype=np.int64)) for _ in range(3)] + \
            [pa.array(rng.standard_normal(n)) for _ in range(3)]
    counts = check(ctx, [rf, ls] + state, [0, 1], 12)
    assert (counts > 0).sum() <= 4  # same-key rows meet in one destination


def test_cfg4_tpch_q5_join_shuffles(ctx):
    rng = np.random.Generator(np.random.PCG64(5))
    n = 600_000  # lineitem stand-in (SF10 has 59 986 052 rows; same key pattern)
    i = np.arange(n, dtype=np.int64)
    l_orderkey = pa.array((i // 8) * 32 + i % 8)
    l_suppkey = pa.array(rng.integers(1, 100_001, n, dtype=np.int64))
    c_nationkey = pa.array(rng.integers(0, 25, n, dtype=np.int64))
    cols = [l_orderkey, l_suppkey, dec128(rng, n), dec128(rng, n), c_nationkey]
    check(ctx, cols, [0], 48)        # Hash([l_orderkey], 6 x 8)
    check(ctx, cols, [1, 4], 48)     # Hash([l_suppkey, c_nationkey], 48)
    m = 15_000  # customer stand-in
    check(ctx, [pa.array(np.arange(1, m + 1, dtype=np.int64)), pa.array(rng.integers(0, 25, m, dtype=np.int64))], [0], 48)


def test_cfg5_clickbench_skewed_utf8_keys(ctx):
    rng = np.random.Generator(np.random.PCG64(29))
    n = 200_000
    uid = (rng.zipf(1.1, n) % 170_000).astype(np.int64)
    phrase_id = rng.zipf(1.1, n) % 60_000
    lens = 5 + (phrase_id * 7919) % 56
    phrases = np.array(["" if rng.random() < 0.7 else "p%d" % pid + "x" * int(ln) for pid, ln in zip(phrase_id, lens)], dtype=object)
    arrays = [pa.array(uid), pa.array(phrases.tolist(), type=pa.string()), pa.array(rng.integers(0, 1 << 40, n, dtype=np.int64))]
    counts = check(ctx, arrays, [0, 1], 6)
    assert counts.max() / counts.mean() < 3.0
