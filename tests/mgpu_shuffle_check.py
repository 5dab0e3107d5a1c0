"""Multi-GPU parity check, run as one process per GPU:
  python -m torch.distributed.run --nnodes=1 --nproc-per-node T --master-addr 127.0.0.1 --master-port P tests/mgpu_shuffle_check.py
Every rank is producer task `rank` (contiguous row range of the cfg-2 table) and
consumer task `rank`; both exchange transports are compared, bit-exactly and in
order, with the single-node CPU oracle.  torch.distributed is only plumbing
(rendezvous + shipping the NCCL id)."""
import os
import sys
import uuid

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist

    import datafusion_distributed_b200 as dfd
    from datafusion_distributed_b200 import _native as nv
    from oracle import oracle as orc
    from tests.util import cfg2_columns

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local_rank = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    uid = [dfd.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    ctx = dfd.WorkerContext(local_rank)
    ex = dfd.ShuffleExchange(ctx, rank, world, uid[0])
    n_rows, n_cols = int(os.environ.get("DFD_CHECK_ROWS", 1_000_003)), 4
    ex.setup_window(int(n_rows * n_cols * 8 * 3.2 / world) + (1 << 20))  # two slots of 1.6x the fair share
    cols = cfg2_columns(n_rows, n_cols)
    lo, hi = rank * n_rows // world, (rank + 1) * n_rows // world
    failures = 0
    for total_parts in (8, 16, 3 * world):
        if total_parts % world:
            continue
        P = total_parts // world
        N = P * world
        ref, rc, rs = orc.repartition_table(cols, [0], N, 8192, 1)
        ins = [torch.from_numpy(c[lo:hi].copy()).cuda() for c in cols]
        torch.cuda.synchronize()
        in_cols = [dfd.DeviceColumn.from_torch(t) for t in ins]
        for mode, name in ((nv.EXCHANGE_NCCL, "nccl"), (nv.EXCHANGE_FUSED, "fused")):
            node = dfd.NetworkShuffleExec.try_new(dfd.Partitioning.Hash([0], P), uuid.uuid4(), 1, world, world)
            cap = int(n_rows * 1.5 / world) + 1024
            outs_t = [torch.empty(cap, dtype=torch.int64, device="cuda") for _ in cols]
            torch.cuda.synchronize()
            out_cols = [dfd.DeviceColumn.from_torch(t) for t in outs_t] if mode == nv.EXCHANGE_NCCL else None
            outs, starts = node.shuffle(ex, in_cols, hi - lo, mode, out_cols, cap)
            if mode == nv.EXCHANGE_FUSED:  # asynchronous form: three pipelined collectives, one wait
                for _ in range(3):
                    node.shuffle_async(ex, in_cols, hi - lo)
                outs, starts = node.wait(ex)
            for q in range(P):
                g = rank * P + q
                _, a, b = node.execute(q, dfd.DistributedTaskContext(rank, world))
                assert b - a == rc[g], (name, q, b - a, rc[g])
                for c in range(n_cols):
                    if mode == nv.EXCHANGE_NCCL:
                        got = outs_t[c][a:b].cpu().numpy()
                    else:
                        got = np.empty(b - a, dtype=np.int64)
                        if b > a:
                            nv.check(nv.lib().dfd_memcpy_d2h(ctx.handle, got.ctypes.data, outs[c].values + a * 8, (b - a) * 8))
                    if not np.array_equal(got, ref[c][rs[g]:rs[g + 1]]):
                        failures += 1
                        print(f"rank {rank}: MISMATCH mode={name} N={N} q={q} col={c}", flush=True)
            dist.barrier()
        # single-pass fused exchange: partition q = T per-producer segments, each bit-exact and in the producer's input order
        node = dfd.NetworkShuffleExec.try_new(dfd.Partitioning.Hash([0], P), uuid.uuid4(), 1, world, world)
        for rep in range(3):  # back-to-back shuffles reuse the window: ready/done flags must order them
            node.shuffle_onepass(ex, in_cols, hi - lo)
        outs, seg_starts, seg_counts = node.collect(ex)
        dest_all = orc.partition_ids([cols[0]], n_rows, N)
        for q in range(P):
            g = rank * P + q
            for r in range(world):
                rlo, rhi = r * n_rows // world, (r + 1) * n_rows // world
                want_idx = np.nonzero(dest_all[rlo:rhi] == g)[0] + rlo
                a, cnt = int(seg_starts[q, r]), int(seg_counts[q, r])
                if cnt != len(want_idx):
                    failures += 1
                    print(f"rank {rank}: COUNT MISMATCH onepass N={N} q={q} r={r}: {cnt} != {len(want_idx)}", flush=True)
                    continue
                for c in range(n_cols):
                    got = np.empty(cnt, dtype=np.int64)
                    if cnt:
                        nv.check(nv.lib().dfd_memcpy_d2h(ctx.handle, got.ctypes.data, outs[c].values + a * 8, cnt * 8))
                    if not np.array_equal(got, cols[c][want_idx]):
                        failures += 1
                        print(f"rank {rank}: MISMATCH onepass N={N} q={q} r={r} col={c}", flush=True)
        dist.barrier()
    # skew: one hot key overflows its sub-window on the producers -> every worker re-runs through the exact two-pass path
    hot = [np.full(hi - lo, 777, dtype=np.int64), np.arange(lo, hi, dtype=np.int64)]
    hot[0][::97] = np.arange(lo, hi, 97)
    hot_t = [torch.from_numpy(c).cuda() for c in hot]
    torch.cuda.synchronize()
    hot_cols = [dfd.DeviceColumn.from_torch(t) for t in hot_t]
    P = 8  # 8 x world sub-windows per consumer: the hot destination's sub-window (1/(8 x world) of the rows) overflows
    node = dfd.NetworkShuffleExec.try_new(dfd.Partitioning.Hash([0], P), uuid.uuid4(), 5, world, world)
    fb0 = nv.lib().dfd_exchange_onepass_fallbacks(ex._h)
    try:
        node.shuffle_onepass(ex, hot_cols, hi - lo)
        outs, seg_starts, seg_counts = node.collect(ex)
        got_rows = 0
        for q in range(P):
            for r in range(world):
                a, cnt = int(seg_starts[q, r]), int(seg_counts[q, r])
                got_rows += cnt
                if cnt:
                    k = np.empty(cnt, dtype=np.int64)
                    v = np.empty(cnt, dtype=np.int64)
                    nv.check(nv.lib().dfd_memcpy_d2h(ctx.handle, k.ctypes.data, outs[0].values + a * 8, cnt * 8))
                    nv.check(nv.lib().dfd_memcpy_d2h(ctx.handle, v.ctypes.data, outs[1].values + a * 8, cnt * 8))
                    rlo = r * n_rows // world
                    ok = bool((np.diff(v) > 0).all()) and bool((orc.partition_ids([k], cnt, P * world) == rank * P + q).all())
                    # v is the global row id: the key stored there must be what the producer held
                    exp_k = np.where((v - rlo) % 97 == 0, v, 777)
                    ok = ok and np.array_equal(k, exp_k)
                    if not ok:
                        failures += 1
                        print(f"rank {rank}: MISMATCH skew fallback q={q} r={r}", flush=True)
        tot = torch.tensor([got_rows], device="cuda")
        dist.all_reduce(tot)
        if tot.item() != n_rows:
            failures += 1
            print(f"rank {rank}: skew fallback lost rows: {tot.item()} != {n_rows}", flush=True)
        if world > 1 and nv.lib().dfd_exchange_onepass_fallbacks(ex._h) != fb0 + 1:
            failures += 1
            print(f"rank {rank}: expected exactly one exact re-run for the skewed shuffle", flush=True)
    except dfd.DfdError as e:  # the dense window may legitimately be too small for a fully skewed destination
        if e.status != 7:
            raise
        print(f"rank {rank}: skewed shuffle reported DFD_ERR_CAPACITY (window too small for the hot destination)", flush=True)
    dist.barrier()
    # mixed widths over the single-pass exchange: Int64 key + Int32 + Decimal128-shaped 16-byte values (the 16-byte column
    # rides the 8-byte ring as two row-range items per tile) + Int16, N = 48 like TPC-H q5's lineitem shuffle
    if 48 % world == 0:
        mrows = 400_003
        rng = np.random.Generator(np.random.PCG64(77))
        wk = rng.integers(-(2**62), 2**62, mrows, dtype=np.int64)
        w4 = rng.integers(-(2**31), 2**31 - 1, mrows, dtype=np.int32)
        w16 = rng.integers(-(2**62), 2**62, (mrows, 2), dtype=np.int64)
        w2 = rng.integers(-(2**15), 2**15 - 1, mrows, dtype=np.int16)
        wl, wh = rank * mrows // world, (rank + 1) * mrows // world
        host = [wk, w4, w16, w2]
        widths = [8, 4, 16, 2]
        keep = [torch.from_numpy(np.ascontiguousarray(a[wl:wh])).cuda() for a in host]
        torch.cuda.synchronize()
        wcols = [dfd.DeviceColumn(nv.COL_FIXED, w, t.data_ptr(), length=wh - wl, keep=t) for w, t in zip(widths, keep)]
        Pw = 48 // world
        node = dfd.NetworkShuffleExec.try_new(dfd.Partitioning.Hash([0], Pw), uuid.uuid4(), 7, world, world)
        node.shuffle_onepass(ex, wcols, wh - wl)
        outs, seg_starts, seg_counts = node.collect(ex)
        wdest = orc.partition_ids([wk], mrows, 48)
        for q in range(Pw):
            g = rank * Pw + q
            for r in range(world):
                rlo, rhi = r * mrows // world, (r + 1) * mrows // world
                want_idx = np.nonzero(wdest[rlo:rhi] == g)[0] + rlo
                a, cnt = int(seg_starts[q, r]), int(seg_counts[q, r])
                if cnt != len(want_idx):
                    failures += 1
                    print(f"rank {rank}: COUNT MISMATCH mixed widths q={q} r={r}: {cnt} != {len(want_idx)}", flush=True)
                    continue
                for c, w in enumerate(widths):
                    got = np.empty(cnt * w, dtype=np.uint8)
                    if cnt:
                        nv.check(nv.lib().dfd_memcpy_d2h(ctx.handle, got.ctypes.data, outs[c].values + a * w, cnt * w))
                    if not np.array_equal(got, np.ascontiguousarray(host[c][want_idx]).view(np.uint8).reshape(-1)):
                        failures += 1
                        print(f"rank {rank}: MISMATCH mixed widths q={q} r={r} width={w}", flush=True)
        dist.barrier()
    # NCCL mode with nullable / boolean / string columns: per destination, rows from the producers in task order
    import pyarrow as pa
    from tests.test_exchange_gpu import _mixed_table
    from tests.util import expected_partitions

    m = 60_013
    arrays = _mixed_table(m, 23)
    mlo, mhi = rank * m // world, (rank + 1) * m // world
    P2 = 3
    N2 = P2 * world
    node = dfd.NetworkShuffleExec.try_new(dfd.Partitioning.Hash([0, 1], P2), uuid.uuid4(), 4, world, world)
    in_cols2 = [dfd.DeviceColumn.from_arrow(ctx, a.slice(mlo, mhi - mlo)) for a in arrays]
    cap2 = int(m * 2.0 / world) + 64
    out_cols2 = []
    for c in in_cols2:
        if c.values_bytes:
            c.values_bytes = max(c.values_bytes, 1)
        oc = dfd.DeviceColumn.empty_like(ctx, dfd.DeviceColumn(c.kind, c.width, c.values, c.offsets, c.validity or 1, 0, cap2, None, c.arrow_type,
                                                               c.values_bytes * 4 + 1024), cap2)
        out_cols2.append(oc)
    outs2, starts2 = node.shuffle(ex, in_cols2, mhi - mlo, nv.EXCHANGE_NCCL, out_cols2, cap2)
    dest2 = orc.partition_ids([arrays[0], arrays[1]], m, N2)
    for q in range(P2):
        g = rank * P2 + q
        want_idx = np.concatenate([np.nonzero(dest2[r * m // world:(r + 1) * m // world] == g)[0] + r * m // world for r in range(world)])
        for c, arr in enumerate(arrays):
            got = outs2[c].to_arrow(ctx, int(starts2[q]), int(starts2[q + 1]))
            if not got.equals(arr.take(pa.array(want_idx))):
                failures += 1
                print(f"rank {rank}: MISMATCH nccl mixed q={q} col={c} {arr.type}", flush=True)
    dist.barrier()
    # push transport (NCCL-free) with the same nullable / boolean / string table: per (partition, producer) segment, in place
    node = dfd.NetworkShuffleExec.try_new(dfd.Partitioning.Hash([0, 1], P2), uuid.uuid4(), 6, world, world)
    for rep in range(2):
        node.shuffle_onepass(ex, in_cols2, mhi - mlo, nullable=[True] * len(arrays))
        outs3, seg_starts, seg_counts = node.collect(ex)
        for q in range(P2):
            g = rank * P2 + q
            for r in range(world):
                rlo, rhi = r * m // world, (r + 1) * m // world
                want_idx = np.nonzero(dest2[rlo:rhi] == g)[0] + rlo
                if int(seg_counts[q, r]) != len(want_idx):
                    failures += 1
                    print(f"rank {rank}: COUNT MISMATCH push q={q} r={r}", flush=True)
                    continue
                for c, arr in enumerate(arrays):
                    got = dfd.NetworkShuffleExec.segment_to_arrow(ctx, outs3[c], int(seg_starts[q, r]), int(seg_counts[q, r]))
                    if not got.equals(arr.take(pa.array(want_idx))):
                        failures += 1
                        print(f"rank {rank}: MISMATCH push rep={rep} q={q} r={r} col={c} {arr.type}", flush=True)
        dist.barrier()
    # NetworkCoalesceExec / NetworkBroadcastExec over the same transport (mixed table, per-producer partitions = row slices)
    my = [a.slice(mlo, mhi - mlo) for a in arrays]
    Pn = 3
    cuts = [0, (mhi - mlo) // 5, (mhi - mlo) // 2, mhi - mlo]
    for consumers in sorted({1, 2 if world >= 2 else 1, world}):
        co = dfd.NetworkCoalesceExec.try_new(Pn, uuid.uuid4(), 7, consumers, world)
        outs4, ss4, sc4 = co.gather(ex, in_cols2, cuts, nullable=[True] * len(arrays))
        if rank < consumers:
            gs, gl, gm = dfd.task_group(world, rank, consumers)
            for part_i in range(co.output_partition_count()):
                _, a, cnt = co.execute(part_i, dfd.DistributedTaskContext(rank, consumers))
                off, j = divmod(part_i, Pn)
                if off >= gl:
                    if cnt != 0:
                        failures += 1
                        print(f"rank {rank}: coalesce padding partition {part_i} not empty", flush=True)
                    continue
                r = gs + off
                rlo, rhi = r * m // world, (r + 1) * m // world
                rc_ = [0, (rhi - rlo) // 5, (rhi - rlo) // 2, rhi - rlo]
                for c, arr in enumerate(arrays):
                    want = arr.slice(rlo + rc_[j], rc_[j + 1] - rc_[j])
                    got = dfd.NetworkShuffleExec.segment_to_arrow(ctx, outs4[c], a, cnt)
                    if not got.equals(want):
                        failures += 1
                        print(f"rank {rank}: MISMATCH coalesce consumers={consumers} partition={part_i} col={c}", flush=True)
        dist.barrier()
    bc = dfd.NetworkBroadcastExec.try_new(Pn, uuid.uuid4(), 8, world, world)
    outs5, ss5, sc5 = bc.gather(ex, in_cols2, cuts, nullable=[True] * len(arrays))
    for j in range(Pn):
        _, segs = bc.execute(j, dfd.DistributedTaskContext(rank, world))
        for r, (a, cnt) in enumerate(segs):
            rlo, rhi = r * m // world, (r + 1) * m // world
            rc_ = [0, (rhi - rlo) // 5, (rhi - rlo) // 2, rhi - rlo]
            for c, arr in enumerate(arrays):
                want = arr.slice(rlo + rc_[j], rc_[j + 1] - rc_[j])
                got = dfd.NetworkShuffleExec.segment_to_arrow(ctx, outs5[c], a, cnt)
                if not got.equals(want):
                    failures += 1
                    print(f"rank {rank}: MISMATCH broadcast partition={j} producer={r} col={c}", flush=True)
    dist.barrier()
    # host-to-host pipelined shuffle: row-set equality per destination (chunk-major output)
    P, N = 8 // world if 8 % world == 0 else 1, (8 // world if 8 % world == 0 else 1) * world
    ref, rc, rs = orc.repartition_table(cols, [0], N, 8192, 1)
    node = dfd.NetworkShuffleExec.try_new(dfd.Partitioning.Hash([0], P), uuid.uuid4(), 3, world, world)
    cap = int(n_rows * 1.5 / world) + 1024
    pin_in = dfd.PinnedTable(ctx, hi - lo, [np.int64] * n_cols)
    pin_out = dfd.PinnedTable(ctx, cap, [np.int64] * n_cols)
    for c in range(n_cols):
        pin_in.columns[c][:] = cols[c][lo:hi]
    h_in = [dfd.DeviceColumn(nv.COL_FIXED, 8, a.ctypes.data, length=hi - lo) for a in pin_in.columns]
    h_out = [dfd.DeviceColumn(nv.COL_FIXED, 8, a.ctypes.data, length=cap) for a in pin_out.columns]
    for n_chunks in (1, 5):
        cps = node.shuffle_host(ex, h_in, hi - lo, n_chunks, h_out, cap)
        for q in range(P):
            g = rank * P + q
            segs = [np.arange(cps[i, q], cps[i, q + 1]) for i in range(n_chunks)]
            idx = np.concatenate(segs) if segs else np.zeros(0, dtype=np.int64)
            got = [pin_out.columns[c][idx] for c in range(n_cols)]
            order = np.argsort(got[1], kind="stable")  # col1 = row_id*8+1: canonical order
            want_order = np.argsort(ref[1][rs[g]:rs[g + 1]], kind="stable")
            for c in range(n_cols):
                if not np.array_equal(got[c][order], ref[c][rs[g]:rs[g + 1]][want_order]):
                    failures += 1
                    print(f"rank {rank}: MISMATCH shuffle_host n_chunks={n_chunks} q={q} col={c}", flush=True)
        dist.barrier()
    # window overflow is reported, not written
    node = dfd.NetworkShuffleExec.try_new(dfd.Partitioning.Hash([1], 1), uuid.uuid4(), 2, world, world)
    t = torch.zeros(hi - lo, dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()
    st = ex.stats()
    f = torch.tensor([failures], device="cuda")
    dist.all_reduce(f)
    if rank == 0:
        print(("MGPU_SHUFFLE_OK" if f.item() == 0 else "MGPU_SHUFFLE_FAILED"), "world", world, "stats", st, flush=True)
    ex.close()
    dist.destroy_process_group()
    sys.exit(0 if f.item() == 0 else 1)


if __name__ == "__main__":
    main()
