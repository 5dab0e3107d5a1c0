"""Property-based CPU tests (hypothesis), in the spirit of the reference's set-equality comparator
(src/test_utils/property_based.rs:15-119): invariants of the oracle and of the exchange arithmetic that
hold for ANY input, not just the seeded fixtures."""
import numpy as np
import pyarrow as pa
from hypothesis import given, settings
from hypothesis import strategies as st

import datafusion_distributed_b200 as dfd
from oracle import oracle as orc
from oracle import oracle_py as op

i64 = st.integers(min_value=-(2**63), max_value=2**63 - 1)


@settings(max_examples=60, deadline=None)
@given(st.lists(st.one_of(st.none(), i64), min_size=0, max_size=200), st.integers(min_value=1, max_value=64))
def test_partitions_are_a_stable_permutation_of_the_input(keys, N):
    n = len(keys)
    k = pa.array(keys, type=pa.int64())
    dest = orc.partition_ids([k], n, N) if n else np.zeros(0, dtype=np.uint32)
    hashes = orc.create_hashes([k], n) if n else np.zeros(0, dtype=np.uint64)
    counts, indices, starts = orc.partition_indices(hashes, N)
    assert counts.sum() == n and starts[-1] == n
    assert sorted(indices.tolist()) == list(range(n))                      # nothing lost, nothing duplicated
    for p in range(N):
        seg = indices[starts[p]:starts[p + 1]]
        assert (dest[seg] == p).all() and (np.diff(seg.astype(np.int64)) > 0).all()   # right destination, input order kept


@settings(max_examples=60, deadline=None)
@given(st.lists(i64, min_size=1, max_size=100), st.integers(min_value=1, max_value=12), st.integers(min_value=1, max_value=8))
def test_task_and_local_partition_split_is_consistent(keys, P, T):
    """Invariant (vi): with N = P*T, consumer task = g // P, local partition = g % P == h % P."""
    k = np.array(keys, dtype=np.int64)
    g = orc.partition_ids([k], len(k), P * T)
    assert np.array_equal(g % P, orc.partition_ids([k], len(k), P))
    assert ((g // P) < T).all()


@settings(max_examples=40, deadline=None)
@given(st.lists(st.one_of(st.none(), st.binary(min_size=0, max_size=40)), min_size=1, max_size=60))
def test_c_and_python_restatements_agree_on_strings(values):
    n = len(values)
    b = pa.array(values, type=pa.binary())
    s = pa.Array.from_buffers(pa.string(), n, b.buffers(), null_count=b.null_count)
    assert orc.create_hashes([s], n).tolist() == op.create_hashes([("str", 0, values)], n)
    assert orc.create_hashes([b], n).tolist() == op.create_hashes([("bytes", 0, values)], n)


@settings(max_examples=80, deadline=None)
@given(st.integers(min_value=1, max_value=8), st.integers(min_value=1, max_value=6), st.data())
def test_exchange_plan_is_a_consistent_all_to_all(T, P, data):
    N = T * P
    counts = np.array(data.draw(st.lists(st.lists(st.integers(min_value=0, max_value=1000), min_size=N, max_size=N), min_size=T, max_size=T)),
                      dtype=np.int64)
    plans = [dfd.exchange_plan(counts, P, r) for r in range(T)]
    # what rank r sends for destination g is exactly what the owner of g expects from r, at the place dest_base points to
    for r in range(T):
        assert plans[r]["send_start"][-1] + counts[r, -1] == counts[r].sum()
        for g in range(N):
            o, q = divmod(g, P)
            assert plans[r]["dest_base"][g] == plans[o]["recv_start"][q, r]
    # receive segments tile each worker's buffer without gaps or overlap
    for o in range(T):
        segs = sorted((int(plans[o]["recv_start"][q, r]), int(counts[r, o * P + q])) for q in range(P) for r in range(T))
        pos = 0
        for start, ln in segs:
            assert start == pos
            pos += ln
        assert pos == plans[o]["recv_rows"] == plans[o]["part_starts"][-1]
    assert sum(p["recv_rows"] for p in plans) == counts.sum()
