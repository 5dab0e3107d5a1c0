"""CPU tests of the multi-worker path (no GPU): the exchange arithmetic
(`dfd_exchange_plan`, product host logic behind the C ABI) driven by a real
world_size-2 `gloo` run, with the CPU oracle standing in for the partition
kernel, checked against the oracle's single-node result."""
import os
import socket
import uuid

import numpy as np
import pytest

import datafusion_distributed_b200 as dfd
from oracle import oracle as orc
from tests.util import cfg2_columns


def test_plan_matches_bruteforce():
    rng = np.random.Generator(np.random.PCG64(0))
    for T, P in [(1, 8), (2, 4), (4, 2), (8, 1), (3, 5)]:
        N = T * P
        counts = rng.integers(0, 50, (T, N), dtype=np.int64)
        counts[rng.random((T, N)) < 0.2] = 0
        for rank in range(T):
            plan = dfd.exchange_plan(counts, P, rank)
            assert np.array_equal(plan["send_start"], np.concatenate([[0], np.cumsum(counts[rank])[:-1]]))
            run = 0
            for q in range(P):
                assert plan["part_starts"][q] == run
                for r in range(T):
                    assert plan["recv_start"][q, r] == run
                    run += counts[r, rank * P + q]
            assert plan["part_starts"][P] == run == plan["recv_rows"]
            # dest_base[g] == the owner's recv_start for (g % P, me)
            for g in range(N):
                owner = dfd.exchange_plan(counts, P, g // P)
                assert plan["dest_base"][g] == owner["recv_start"][g % P, rank]


def test_plan_rejects_bad_arguments():
    with pytest.raises(dfd.DfdError):
        dfd.exchange_plan(np.array([[1, -1]], dtype=np.int64), 2, 0)
    with pytest.raises(dfd.DfdError):
        dfd.exchange_plan(np.zeros((2, 4), dtype=np.int64), 2, 2)


def test_network_shuffle_exec_surface():
    ex = dfd.NetworkShuffleExec.try_new(dfd.Partitioning.Hash([0], 4), uuid.uuid4(), 1, task_count=2, input_task_count=2)
    assert ex.name() == "NetworkShuffleExec"
    assert ex.output_partitioning().partition_count == 4          # advertised: Hash(keys, P)
    assert ex.input_stage_plan().partition_count == 8             # producer scaled to P * task_count
    assert len(ex.input_stage.tasks) == 2
    with pytest.raises(ValueError):
        dfd.NetworkShuffleExec.try_new(dfd.Partitioning.Hash([], 4), uuid.uuid4(), 1, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, P, n_rows, ret):
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        N = P * world
        cols = cfg2_columns(n_rows, 3)
        lo, hi = rank * n_rows // world, (rank + 1) * n_rows // world  # contiguous row range per producer task
        local = [c[lo:hi] for c in cols]
        # producer: local partition (CPU oracle stands in for K1/K2 on this GPU-less box)
        parts, counts, starts = orc.repartition_table(local, [0], N, 8192, 1)
        all_counts = [torch.zeros(N, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(all_counts, torch.from_numpy(counts.copy()))
        cm = torch.stack(all_counts).numpy()
        plan = dfd.exchange_plan(cm, P, rank)  # product host logic
        recv = [np.zeros(plan["recv_rows"], dtype=np.int64) for _ in cols]
        for c in range(len(cols)):
            for peer in range(world):  # exchange (gloo point-to-point in place of ncclSend/ncclRecv)
                for q in range(P):
                    g_out = peer * P + q
                    seg = parts[c][plan["send_start"][g_out]: plan["send_start"][g_out] + cm[rank, g_out]]
                    g_in = rank * P + q
                    n_in = int(cm[peer, g_in])
                    dst = recv[c][plan["recv_start"][q, peer]: plan["recv_start"][q, peer] + n_in]
                    if peer == rank:
                        dst[:] = seg
                    else:
                        ops = []
                        t_out = torch.from_numpy(seg.copy())
                        t_in = torch.from_numpy(dst)
                        if rank < peer:
                            if len(seg): dist.send(t_out, peer)
                            if n_in: dist.recv(t_in, peer)
                        else:
                            if n_in: dist.recv(t_in, peer)
                            if len(seg): dist.send(t_out, peer)
        # consumer check: my P partitions == the single-node oracle's global partitions rank*P+q, exact order
        ref, rc, rs = orc.repartition_table(cols, [0], N, 8192, 1)
        for q in range(P):
            g = rank * P + q
            a, b = plan["part_starts"][q], plan["part_starts"][q + 1]
            for c in range(len(cols)):
                assert np.array_equal(recv[c][a:b], ref[c][rs[g]:rs[g + 1]]), (rank, q, c)
        ret.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        ret.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("P", [1, 4])
def test_two_worker_shuffle_over_gloo(built, P):
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, P, 40_000, ret)) for r in range(2)]
    for p in procs:
        p.start()
    results = [ret.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(results) == [(0, "ok"), (1, "ok")], results


def test_coalesce_task_groups_match_the_reference(built):
    """NetworkCoalesceExec's task grouping (reference: src/execution_plans/network_coalesce.rs:264-289 and its unit tests
    :300-420): contiguous groups, the first `input % consumers` groups get one extra task, max_len sizes the output."""
    import datafusion_distributed_b200 as dfd

    for input_tasks, consumers in [(9, 3), (9, 2), (3, 3), (3, 5), (1, 1), (7, 3), (8, 8), (16, 4), (5, 1), (0, 2)]:
        base, extra = divmod(input_tasks, consumers)
        start = 0
        covered = []
        for t in range(consumers):
            want_len = base + (1 if t < extra else 0)
            s, l, m = dfd.task_group(input_tasks, t, consumers)
            assert (s, l) == (start, want_len), (input_tasks, consumers, t)
            assert m == base + (1 if extra else 0)
            covered += list(range(s, s + l))
            start += want_len
        assert covered == list(range(input_tasks))
    assert dfd.task_group(4, 0, 0) == (0, 0, 0)
    node = dfd.NetworkCoalesceExec.try_new(3, __import__("uuid").uuid4(), 1, 2, 9)  # 9 input tasks x 3 partitions -> 2 consumers
    assert node.output_partition_count() == 3 * 5
    import pytest as _pt
    with _pt.raises(ValueError):
        dfd.NetworkCoalesceExec.try_new(3, __import__("uuid").uuid4(), 1, 0, 9)


def _route_source(route, P, T, Cn, consumer, segment=None):
    import ctypes as C

    from datafusion_distributed_b200 import _native as nv

    prod, sl, nseg = C.c_int(-7), C.c_uint32(0), C.c_uint32(0)
    if segment is None:
        nv.check(nv.lib().dfd_route_segment_source(route, P, T, Cn, consumer, 0, None, None, C.byref(nseg)))
        return nseg.value
    nv.check(nv.lib().dfd_route_segment_source(route, P, T, Cn, consumer, segment, C.byref(prod), C.byref(sl), C.byref(nseg)))
    return prod.value, sl.value


def test_routing_table_is_the_reference_execute_arithmetic(built):
    """dfd_route_segment_source (the table every worker derives before pushing) against a restatement of the three
    `execute(partition, ctx)` bodies of the reference: which input task, and which of its partitions, every consumer
    partition stream reads (network_shuffle.rs:219-231, network_coalesce.rs:186-226, network_broadcast.rs:230-241)."""
    SHUFFLE, COALESCE, BROADCAST = 0, 1, 2
    for T in (1, 2, 3, 4, 8):
        for P in (1, 3, 6):
            # --- NetworkShuffleExec::execute: off = P * task_index; partition off + p from EVERY input task (select_all)
            for task_index in range(T):
                assert _route_source(SHUFFLE, P, T, T, task_index) == P * T
                off = P * task_index
                for p in range(P):
                    streams = {(input_task, off + p) for input_task in range(T)}
                    got = {_route_source(SHUFFLE, P, T, T, task_index, p * T + r) for r in range(T)}
                    assert got == streams
                    for r in range(T):  # producer order inside a partition == input task order
                        assert _route_source(SHUFFLE, P, T, T, task_index, p * T + r) == (r, off + p)
            # --- NetworkBroadcastExec::execute: the same loop over input tasks; the producer's BroadcastExec serves this
            # consumer's replica of its partition p under index off + p, i.e. producer slice p
            for Cn in range(1, T + 1):
                for task_index in range(T):
                    n = _route_source(BROADCAST, P, T, Cn, task_index)
                    assert n == (P * T if task_index < Cn else 0)
                    for s in range(n):
                        assert _route_source(BROADCAST, P, T, Cn, task_index, s) == (s % T, s // T)
            # --- NetworkCoalesceExec::execute
            for Cn in range(1, T + 1):
                base, extra = divmod(T, Cn)
                max_len = base + (1 if extra else 0)
                seen = []
                for task_index in range(Cn):
                    length = base + (1 if task_index < extra else 0)
                    start = task_index * base + min(task_index, extra)
                    partition_count = max_len * P  # what the node advertises: partitions_per_task x the largest group
                    assert _route_source(COALESCE, P, T, Cn, task_index) == partition_count
                    for partition in range(partition_count):
                        input_task_offset, target_partition = divmod(partition, P)
                        want = (-1, target_partition) if input_task_offset >= length else (start + input_task_offset, target_partition)
                        assert _route_source(COALESCE, P, T, Cn, task_index, partition) == want
                        if want[0] >= 0:
                            seen.append(want)
                assert sorted(seen) == [(r, g) for r in range(T) for g in range(P)]  # every producer partition read exactly once
                for idle in range(Cn, T):
                    assert _route_source(COALESCE, P, T, Cn, idle) == 0


def test_routing_table_rejects_bad_arguments(built):
    import ctypes as C

    import datafusion_distributed_b200 as dfd
    from datafusion_distributed_b200 import _native as nv

    f = nv.lib().dfd_route_segment_source
    n = C.c_uint32(0)
    assert f(0, 4, 2, 1, 0, 0, None, None, C.byref(n)) != 0   # shuffle: consumers == workers
    assert f(3, 4, 2, 2, 0, 0, None, None, C.byref(n)) != 0   # unknown route
    assert f(1, 0, 2, 2, 0, 0, None, None, C.byref(n)) != 0   # no partitions
    assert f(1, 4, 2, 3, 0, 0, None, None, C.byref(n)) != 0   # more consumers than workers
    p = C.c_int(0)
    assert f(2, 4, 2, 2, 0, 8, C.byref(p), None, None) != 0   # segment out of range
    assert "out of range" in dfd.last_error() if hasattr(dfd, "last_error") else True
