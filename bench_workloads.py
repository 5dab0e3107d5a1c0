"""bench_workloads.py — the other BASELINE.json configs as bench lines (`bench.py --workload cfg3|cfg4|cfg5`).

  cfg3  TPC-H SF1 q1: the post-partial-aggregate shuffle GROUP BY (l_returnflag, l_linestatus): a handful of rows per
        producer, Utf8 keys, 10 aggregate-state columns -> latency bound: reported as microseconds per shuffle.
        Plan shape: /root/reference tests/tpch_plans_test.rs:29-32.
  cfg4  TPC-H SF10 q5: the lineitem side of the partitioned join on l_orderkey at SF10 cardinality
        (59 986 052 rows x {l_orderkey, l_suppkey: Int64; l_extendedprice, l_discount: Decimal128}), Hash N = 48
        (target_partitions 6 x 8 tasks).  Plan shape: tests/tpch_plans_test.rs:223-226.
  cfg5  ClickBench q16/q17-style GROUP BY (UserID, SearchPhrase): 100 M rows, UserID Zipf(1.1) over 17 M ids,
        SearchPhrase 70 % empty / Zipf over 6 M phrases of 5-60 bytes; keys Int64 + Utf8, skewed destinations.
        Plan shape: tests/clickbench_plans_test.rs:379-398.
The data sets are synthetic stand-ins with the published cardinalities (tpchgen / hits.parquet are not available
offline).  One process per GPU; every rank checks bit-parity against the CPU oracle on a seeded slice
before the timed region, exactly like the cfg2 line.  Timing: CUDA events on the library stream, max over ranks."""
from __future__ import annotations

import json
import os
import sys
import time
import uuid

import numpy as np

NVLINK_PEAK_GBS = 770.0


# ------------------------------------------------------------------------------------------------ data ----

_LINES = np.repeat(np.arange(7), np.arange(1, 8))  # 1..7 lines per order, 28 lines per 7 orders


def cfg4_columns(lo: int, hi: int, seed: int = 5):
    """Rows [lo, hi) of the lineitem stand-in: l_orderkey follows TPC-H's sparse key pattern
    (orderkey = (o / 8) * 32 + o % 8, 1-7 lines per order), l_suppkey uniform in [1, 100 000], two Decimal128 columns."""
    i = np.arange(lo, hi, dtype=np.int64)
    order = (i // 28) * 7 + _LINES[i % 28]
    orderkey = (order // 8) * 32 + order % 8
    rng = np.random.Generator(np.random.PCG64([seed, lo]))
    supp = rng.integers(1, 100_001, hi - lo, dtype=np.int64)
    price = np.zeros((hi - lo, 2), dtype=np.int64)
    price[:, 0] = rng.integers(90_000, 10_500_000, hi - lo, dtype=np.int64)  # Decimal128(15,2) as (lo, hi) limbs
    disc = np.zeros((hi - lo, 2), dtype=np.int64)
    disc[:, 0] = rng.integers(0, 11, hi - lo, dtype=np.int64)
    return [orderkey, supp, price, disc]


def cfg5_columns(lo: int, hi: int, seed: int = 29):
    """Rows [lo, hi) of the hits stand-in as pyarrow arrays: UserID Int64 (Zipf 1.1 over 17 M ids), SearchPhrase Utf8
    (70 % empty, else Zipf over 6 M phrases whose text is a pure function of the phrase id), c Int64 (a partial count)."""
    import pyarrow as pa

    n = hi - lo
    rng = np.random.Generator(np.random.PCG64([seed, lo]))

    def zipf(m, size, s=1.1):  # truncated power law over 1..m by inverse CDF (vectorised; numpy's rejection sampler is 10x slower)
        u = rng.random(size)
        return np.minimum((1.0 + u * (float(m) ** (1.0 - s) - 1.0)) ** (1.0 / (1.0 - s)), float(m)).astype(np.int64)

    uid = (zipf(17_000_000, n) - 1) * 2_654_435_761 % (1 << 40)
    nonempty = rng.random(n) >= 0.7
    pid = np.where(nonempty, zipf(6_000_000, n) - 1, -1).astype(np.int64)
    lens = np.where(pid >= 0, 5 + (pid * 40_503) % 56, 0).astype(np.int64)
    offsets = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(lens, out=offsets[1:])
    total = int(offsets[-1])
    within = np.arange(total, dtype=np.int64) - np.repeat(offsets[:-1], lens)
    base = np.repeat(pid * 7_919, lens)
    data = ((base + within * 31) % 26 + 97).astype(np.uint8)
    phrase = pa.Array.from_buffers(pa.string(), n, [None, pa.py_buffer(offsets.astype(np.int32).tobytes()), pa.py_buffer(data.tobytes())])
    cnt = np.ones(n, dtype=np.int64)
    return [pa.array(uid), phrase, pa.array(cnt)]


def cfg3_columns(rank: int, input_partitions: int = 6, seed: int = 3):
    """One producer's partial-aggregate output for q1: (l_returnflag, l_linestatus) groups A/F, N/F, N/O, R/F per input
    partition; sum/avg/count states as 4 x Decimal128, 4 x Int64, 2 x Float64."""
    import pyarrow as pa

    groups = [("A", "F"), ("N", "F"), ("N", "O"), ("R", "F")] * input_partitions
    n = len(groups)
    rng = np.random.Generator(np.random.PCG64([seed, rank]))
    cols = [pa.array([g[0] for g in groups], type=pa.string()), pa.array([g[1] for g in groups], type=pa.string())]
    for _ in range(4):
        raw = np.zeros((n, 2), dtype=np.int64)
        raw[:, 0] = rng.integers(0, 1 << 50, n)
        cols.append(pa.Array.from_buffers(pa.decimal128(38, 4), n, [None, pa.py_buffer(raw.tobytes())]))
    for _ in range(4):
        cols.append(pa.array(rng.integers(0, 1 << 40, n, dtype=np.int64)))
    for _ in range(2):
        cols.append(pa.array(rng.standard_normal(n)))
    return cols


# ---------------------------------------------------------------------------------------------- helpers ----

def _dist_setup(torch, dfd, world):
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local_rank)
    uid = [None]
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        uid = [dfd.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
    ctx = dfd.WorkerContext(local_rank)
    ex = dfd.ShuffleExchange(ctx, rank, world, uid[0])
    return dist, rank, local_rank, ctx, ex


def _allreduce(torch, dist, world, value, op="sum", dtype=None):
    t = torch.tensor([value], dtype=dtype or torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op={"sum": dist.ReduceOp.SUM, "max": dist.ReduceOp.MAX}[op])
    return t.item()


def _fixed16(torch, nv, dfd, arr2):
    """(n, 2) int64 numpy -> device Decimal128-shaped column (16-byte fixed values)."""
    t = torch.from_numpy(arr2).cuda()
    return dfd.DeviceColumn(nv.COL_FIXED, 16, t.data_ptr(), length=arr2.shape[0], keep=t), t


def _emit(line):
    print(json.dumps(line))


def _timed(ctx, dist, torch, world, fn, steps):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ctx.synchronize()
    ctx.timer_start()
    out = None
    for _ in range(steps):
        out = fn()
    ms = ctx.timer_stop()
    return _allreduce(torch, dist, world, ms, "max") / steps, out


# ------------------------------------------------------------------------------------------------ cfg4 ----

def run_cfg4(args, torch, dfd, world):
    from datafusion_distributed_b200 import _native as nv
    from oracle import oracle as orc

    dist, rank, local_rank, ctx, ex = _dist_setup(torch, dfd, world)
    n_total = args.rows if args.rows != (1 << 26) else 59_986_052
    N = 48
    P = N // world
    lo, hi = rank * n_total // world, (rank + 1) * n_total // world
    n = hi - lo
    widths = [8, 8, 16, 16]
    row_bytes = sum(widths)

    def upload(cols):
        keep, dcols = [], []
        for c in cols:
            if c.ndim == 2:
                dc, t = _fixed16(torch, nv, dfd, c)
            else:
                t = torch.from_numpy(c).cuda()
                dc = dfd.DeviceColumn.from_torch(t)
            keep.append(t)
            dcols.append(dc)
        return dcols, keep

    # ---- bit-parity on a seeded slice (every rank; per (partition, producer) segment, values and order)
    n_chk = args.parity_rows
    chk = cfg4_columns(0, n_chk)
    clo, chi = rank * n_chk // world, (rank + 1) * n_chk // world
    ccols, _k = upload([c[clo:chi] for c in chk])
    dest = orc.partition_ids([chk[0]], n_chk, N)
    bad = 0
    if world == 1:
        part = dfd.HashPartitioner(ctx, dfd.Partitioning.Hash([0], N))
        outs, starts, counts = part.partition_onepass(ccols, n_chk)
        segs = {(q, 0): (int(starts[q]), int(counts[q])) for q in range(N)}
    else:
        ex.setup_window(int(n * row_bytes * 1.3) + (8 << 20))
        node = dfd.NetworkShuffleExec.try_new(dfd.Partitioning.Hash([0], P), uuid.uuid4(), 4, world, world)
        node.shuffle_onepass(ex, ccols, chi - clo)
        outs, ss, sc = node.collect(ex)
        segs = {(q, r): (int(ss[q, r]), int(sc[q, r])) for q in range(P) for r in range(world)}
    for (q, r), (a, cnt) in segs.items():
        g = rank * P + q if world > 1 else q
        rlo, rhi = r * n_chk // world, (r + 1) * n_chk // world
        want = np.nonzero(dest[rlo:rhi] == g)[0] + rlo
        if cnt != len(want):
            bad += 1
            continue
        for c, w in enumerate(widths):
            got = np.empty(cnt * w, dtype=np.uint8)
            if cnt:
                nv.check(nv.lib().dfd_memcpy_d2h(ctx.handle, got.ctypes.data, outs[c].values + a * w, cnt * w))
            if not np.array_equal(got, np.ascontiguousarray(chk[c][want]).view(np.uint8).reshape(-1)):
                bad += 1
    bad = int(_allreduce(torch, dist, world, bad, "sum"))
    if bad:
        if rank == 0:
            _emit({"workload": "cfg4", "n_gpus": world, "parity_checked": False, "parity_mismatching_segments": bad})
        sys.exit(3)
    del ccols, _k, outs

    # ---- timed: the full lineitem stand-in
    cols = cfg4_columns(lo, hi)
    dcols, keep = upload(cols)
    del cols
    torch.cuda.synchronize()
    if world == 1:
        rr = part.default_region_rows(n)
        outs_t = [torch.empty(N * rr * (w // 8), dtype=torch.int64, device="cuda") for w in widths]
        out_cols = [dfd.DeviceColumn(nv.COL_FIXED, w, t.data_ptr(), length=N * rr, keep=t) for w, t in zip(widths, outs_t)]
        step = lambda: part.partition_onepass(dcols, n, rr, out_cols, sync=False)
        fin = lambda: int(part.collect()[1].sum())
    else:
        step = lambda: node.shuffle_onepass(ex, dcols, n)
        fin = lambda: int(node.collect(ex)[2].sum())
    for _ in range(max(args.warmup, 3)):
        step()
    fin()
    ms, _ = _timed(ctx, dist, torch, world, step, args.steps)
    got = fin()
    assert int(_allreduce(torch, dist, world, got, "sum")) == n_total
    reruns = ctx.metrics()["onepass_reruns"] if world == 1 else int(nv.lib().dfd_exchange_onepass_fallbacks(ex._h))
    if rank == 0:
        rows_s = n_total / (ms / 1e3)
        if world == 1:
            import bench as B

            peak, src = B.measured_peaks()
            ach = 2.0 * row_bytes * n_total / (ms / 1e3) / 1e9
            roof = {"bound": "hbm", "kernel": "k_scatter_onepass (one launch, mixed 8- and 16-byte columns)", "achieved": ach, "peak": peak,
                    "unit": "GB/s", "frac": ach / peak, "peak_source": src, "traffic": None, "algorithmic_bytes_per_row": 2 * row_bytes}
        else:
            alg = row_bytes * n * (world - 1) / world
            ach = alg / (ms / 1e3) / 1e9
            roof = {"bound": "nvlink", "kernel": "k_scatter_onepass<PEER> + k_scatter<PEER> (peer stores)", "achieved": ach, "peak": NVLINK_PEAK_GBS,
                    "unit": "GB/s", "frac": ach / NVLINK_PEAK_GBS, "peak_source": "measured peer copy per direction (B200_PROFILING.md)",
                    "traffic": None, "algorithmic_bytes_per_gpu_per_direction": alg}
        _emit({"metric": "shuffle rows/sec (TPC-H SF10 q5 lineitem repartition on l_orderkey, Hash N=48)", "value": rows_s, "unit": "rows/s",
               "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True,
               "scaling": "strong", "vs_baseline": None, "dtype": "int64/decimal128", "data": "synthetic",
               "config": {"workload": "cfg4: TPC-H SF10 q5 stand-in, lineitem 59 986 052 rows x {l_orderkey, l_suppkey: Int64; l_extendedprice, "
                                      "l_discount: Decimal128}, Hash([l_orderkey], 48), device-resident", "rows": n_total, "num_partitions": N,
                          "partitions_per_task": P}, "parity_checked": True, "parity_rows": n_chk, "exact_reruns": int(reruns), "roofline": roof,
               "gpu_launches": int(ctx.metrics()["kernel_launches"])})
    ex.close()
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------ cfg5 ----

def run_cfg5(args, torch, dfd, world):
    import pyarrow as pa

    from datafusion_distributed_b200 import _native as nv
    from oracle import oracle as orc

    dist, rank, local_rank, ctx, ex = _dist_setup(torch, dfd, world)
    n_total = args.rows if args.rows != (1 << 26) else 100_000_000
    P = 3
    N = P * world
    lo, hi = rank * n_total // world, (rank + 1) * n_total // world
    n = hi - lo
    arrays = cfg5_columns(lo, hi)
    str_bytes = arrays[1].buffers()[2].size
    dcols = [dfd.DeviceColumn.from_arrow(ctx, a) for a in arrays]
    row_bytes_avg = 8 + 8 + 4 + str_bytes / max(n, 1)
    # (every worker must pass the same window size: derive it from the largest producer)
    win = int(_allreduce(torch, dist, world, int(n * row_bytes_avg * 2.2) + (64 << 20), "max"))
    ex.setup_window((win + 4095) // 4096 * 4096)  # skewed destinations: generous window
    node = dfd.NetworkShuffleExec.try_new(dfd.Partitioning.Hash([0, 1], P), uuid.uuid4(), 5, world, world)
    nullable = [False, False, False]

    # ---- parity on a seeded slice: per (partition, producer) segment, in place from the window
    n_chk = min(args.parity_rows, 1 << 20)
    chk = cfg5_columns(0, n_chk, seed=31)
    clo, chi = rank * n_chk // world, (rank + 1) * n_chk // world
    ccols = [dfd.DeviceColumn.from_arrow(ctx, a.slice(clo, chi - clo)) for a in chk]
    dest = orc.partition_ids([chk[0], chk[1]], n_chk, N)
    node.shuffle_onepass(ex, ccols, chi - clo, nullable)
    outs, ss, sc = node.collect(ex)
    bad = 0
    for q in range(P):
        g = rank * P + q
        for r in range(world):
            rlo, rhi = r * n_chk // world, (r + 1) * n_chk // world
            want = np.nonzero(dest[rlo:rhi] == g)[0] + rlo
            if int(sc[q, r]) != len(want):
                bad += 1
                continue
            for c, arr in enumerate(chk):
                got = dfd.NetworkShuffleExec.segment_to_arrow(ctx, outs[c], int(ss[q, r]), int(sc[q, r]))
                if not got.equals(arr.take(pa.array(want))):
                    bad += 1
    bad = int(_allreduce(torch, dist, world, bad, "sum"))
    if bad:
        if rank == 0:
            _emit({"workload": "cfg5", "n_gpus": world, "parity_checked": False, "parity_mismatching_segments": bad})
        sys.exit(3)

    # ---- timed: push transport (NCCL-free) and, for comparison, the NCCL send/recv transport
    def push_step():
        node.shuffle_onepass(ex, dcols, n, nullable)
        return node.collect(ex)

    for _ in range(max(args.warmup, 3)):
        push_step()
    st0 = ex.stats()
    ms_push, res = _timed(ctx, dist, torch, world, push_step, args.steps)
    st1 = ex.stats()
    counts = res[2].sum(axis=1)  # rows of my P partitions
    part_rows = torch.zeros(N, dtype=torch.int64, device="cuda")
    part_rows[rank * P:(rank + 1) * P] = torch.from_numpy(counts).cuda()
    if world > 1:
        dist.all_reduce(part_rows)
    assert int(part_rows.sum().item()) == n_total
    cap_rows = int(part_rows.view(world, P).sum(dim=1).max().item() * 1.05) + 1024
    tot_bytes = torch.tensor([str_bytes], dtype=torch.int64, device="cuda")
    if world > 1:
        dist.all_reduce(tot_bytes)
    out_cols = []
    for c in dcols:
        proto = dfd.DeviceColumn(c.kind, c.width, c.values, c.offsets, 0, 0, cap_rows, None, c.arrow_type, int(tot_bytes.item() // world * 2) + 1024)
        out_cols.append(dfd.DeviceColumn.empty_like(ctx, proto, cap_rows))

    def nccl_step():
        return node.shuffle(ex, dcols, n, nv.EXCHANGE_NCCL, out_cols, cap_rows)

    ms_nccl = None
    try:
        for _ in range(2):
            nccl_step()
        ms_nccl, _ = _timed(ctx, dist, torch, world, nccl_step, max(2, args.steps // 2))
    except dfd.DfdError as e:  # capacity of the caller-provided NCCL-mode buffers
        ms_nccl = None
        if rank == 0:
            print(f"# NCCL-mode comparison skipped: {e}", file=sys.stderr)
    sent = (st1["bytes_sent"] - st0["bytes_sent"]) / args.steps
    sent_max = _allreduce(torch, dist, world, sent, "max")
    if rank == 0:
        pr = part_rows.cpu().numpy()
        ach = sent_max * (world - 1) / max(world, 1) / (ms_push / 1e3) / 1e9 if world > 1 else None
        _emit({"metric": f"shuffle rows/sec (ClickBench GROUP BY UserID, SearchPhrase stand-in, {n_total / 1e6:g}M rows, Int64 + Utf8 keys, skewed)",
               "value": n_total / (ms_push / 1e3), "unit": "rows/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
               "ms_per_step": ms_push, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "int64/utf8", "data": "synthetic",
               "config": {"workload": f"cfg5: {n_total / 1e6:g}M rows (the config names 100M on 8 GPUs = 12.5M per GPU; --rows scales it to the GPUs used), UserID Int64 Zipf(1.1) over 17M ids, SearchPhrase Utf8 (70% empty, Zipf over 6M phrases of "
                                      "5-60 B), c Int64; Hash([UserID, SearchPhrase], 3 x tasks), device-resident", "rows": n_total,
                          "num_partitions": N, "partitions_per_task": P, "exchange": "push (local partition -> flag all-gather -> k_push_runs peer stores)"},
               "parity_checked": True, "parity_rows": n_chk,
               "skew": {"max_partition_rows": int(pr.max()), "mean_partition_rows": float(pr.mean()), "max_over_mean": float(pr.max() / pr.mean())},
               "roofline": {"bound": "nvlink" if world > 1 else "hbm", "kernel": "k_push_runs (+ local K1/K1b/K2/K4)", "achieved": ach,
                            "peak": NVLINK_PEAK_GBS, "unit": "GB/s", "frac": (ach / NVLINK_PEAK_GBS) if ach else None,
                            "peak_source": "measured peer copy per direction (B200_PROFILING.md)", "traffic": None,
                            "bytes_pushed_per_gpu_per_step_max": sent_max},
               "nccl_mode_ms_per_step": ms_nccl, "push_over_nccl": (ms_nccl / ms_push) if ms_nccl else None,
               "gpu_launches": int(ctx.metrics()["kernel_launches"])})
    ex.close()
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------ cfg3 ----

def run_cfg3(args, torch, dfd, world):
    import pyarrow as pa

    from datafusion_distributed_b200 import _native as nv
    from oracle import oracle as orc

    dist, rank, local_rank, ctx, ex = _dist_setup(torch, dfd, world)
    P = 3
    N = P * world
    arrays = cfg3_columns(rank)
    n = len(arrays[0])
    dcols = [dfd.DeviceColumn.from_arrow(ctx, a) for a in arrays]
    ex.setup_window(4 << 20)
    node = dfd.NetworkShuffleExec.try_new(dfd.Partitioning.Hash([0, 1], P), uuid.uuid4(), 3, world, world)
    nullable = [False] * len(arrays)
    # parity: every rank knows every producer's rows (pure function of the rank)
    everyone = [cfg3_columns(r) for r in range(world)]
    node.shuffle_onepass(ex, dcols, n, nullable)
    outs, ss, sc = node.collect(ex)
    bad = 0
    for q in range(P):
        g = rank * P + q
        for r in range(world):
            d = orc.partition_ids([everyone[r][0], everyone[r][1]], n, N)
            want = np.nonzero(d == g)[0]
            if int(sc[q, r]) != len(want):
                bad += 1
                continue
            for c, arr in enumerate(everyone[r]):
                got = dfd.NetworkShuffleExec.segment_to_arrow(ctx, outs[c], int(ss[q, r]), int(sc[q, r]))
                if not got.equals(arr.take(pa.array(want))):
                    bad += 1
    bad = int(_allreduce(torch, dist, world, bad, "sum"))
    if bad:
        if rank == 0:
            _emit({"workload": "cfg3", "n_gpus": world, "parity_checked": False, "parity_mismatching_segments": bad})
        sys.exit(3)

    def push_step():
        node.shuffle_onepass(ex, dcols, n, nullable)
        return node.collect(ex)

    for _ in range(20):
        push_step()
    steps = max(args.steps, 200)
    ms_push, _ = _timed(ctx, dist, torch, world, push_step, steps)
    cap = n * world + 64
    out_cols = [dfd.DeviceColumn.empty_like(ctx, dfd.DeviceColumn(c.kind, c.width, c.values, c.offsets, 0, 0, cap, None, c.arrow_type, 4096), cap)
                for c in dcols]
    nccl_step = lambda: node.shuffle(ex, dcols, n, nv.EXCHANGE_NCCL, out_cols, cap)
    for _ in range(20):
        nccl_step()
    ms_nccl, _ = _timed(ctx, dist, torch, world, nccl_step, steps)
    if rank == 0:
        _emit({"metric": "shuffle latency (TPC-H q1 post-partial-aggregate repartition, Utf8 keys)", "value": ms_push * 1e3, "unit": "us/shuffle",
               "n_gpus": world, "steps": steps, "warmup": 20, "ms_per_step": ms_push, "higher_is_better": False, "scaling": "weak",
               "vs_baseline": None, "dtype": "utf8/decimal128/int64/float64", "data": "synthetic",
               "config": {"workload": f"cfg3: TPC-H q1 stand-in, {n} partial-aggregate rows per producer task (4 groups x 6 input partitions), keys "
                                      "(l_returnflag, l_linestatus) Utf8, 10 state columns; Hash(keys, 3 x tasks)", "rows_per_task": n,
                          "num_partitions": N, "exchange": "push (NCCL-free)"},
               "parity_checked": True, "parity_rows": n * world,
               "roofline": {"bound": "latency", "kernel": "k_xchg_allgather_meta + k_push_runs + k_xchg_done_barrier", "achieved": None, "peak": None,
                            "unit": None, "frac": None, "traffic": None,
                            "note": "a few hundred bytes per shuffle: bounded by kernel-launch + NVLink flag round trips + one host sync, not bandwidth"},
               "nccl_mode_us_per_shuffle": ms_nccl * 1e3, "push_over_nccl": ms_nccl / ms_push,
               "gpu_launches": int(ctx.metrics()["kernel_launches"])})
    ex.close()
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------- agg ----

def run_agg(args, torch, dfd, world):
    """PartialReduce ahead of the shuffle: the producer's Partial-aggregate output (group key + 4 states) enters
    from HOST memory once, is repartitioned (K1/K1b/K2), reduced per destination (dfd_partial_reduce_device) and — at world > 1 —
    exchanged (pre-partitioned shuffle) on the device; only the REDUCED rows come back to the host.  e2e = wall clock over the
    whole table, H2D of every chunk and D2H of every reduced chunk inside the timed region."""
    from datafusion_distributed_b200 import _native as nv

    dist, rank, local_rank, ctx, ex = _dist_setup(torch, dfd, world)
    n_total = args.rows
    groups = 1 << 16
    P = 8 // world if 8 % world == 0 else 1
    N = P * world
    lo, hi = rank * n_total // world, (rank + 1) * n_total // world
    n = hi - lo
    chunk = 1 << 22
    ops = [-1, nv.AGG_SUM_I64, nv.AGG_SUM_I64, nv.AGG_MIN_I64, nv.AGG_MAX_I64]
    pt = dfd.PinnedTable(ctx, n, [np.int64] * 5)
    rng = np.random.Generator(np.random.PCG64([77, rank]))
    for a in range(0, n, chunk):
        b = min(a + chunk, n)
        pt.columns[0][a:b] = rng.integers(0, groups, b - a, dtype=np.int64) * 2_654_435_761
        for j in range(1, 5):
            pt.columns[j][a:b] = rng.integers(-(1 << 40), 1 << 40, b - a, dtype=np.int64)
    lib_stream = torch.cuda.ExternalStream(ctx.stream_ptr())
    d_in = [torch.empty(chunk, dtype=torch.int64, device="cuda") for _ in range(5)]
    d_part = [torch.empty(chunk, dtype=torch.int64, device="cuda") for _ in range(5)]
    d_red = [torch.empty(chunk, dtype=torch.int64, device="cuda") for _ in range(5)]
    h_red = dfd.PinnedTable(ctx, chunk, [np.int64] * 5)
    h_red_t = [torch.from_numpy(c) for c in h_red.columns]
    h_in_t = [torch.from_numpy(c) for c in pt.columns]
    part = dfd.HashPartitioner(ctx, dfd.Partitioning.Hash([0], N))
    red = dfd.PartialReduceExec(ctx, [0], ops)
    node = dfd.NetworkShuffleExec.try_new(dfd.Partitioning.Hash([0], P), uuid.uuid4(), 9, world, world)
    if world > 1:
        ex.setup_window(64 << 20)
    in_cols = [dfd.DeviceColumn.from_torch(t) for t in d_in]
    part_cols = [dfd.DeviceColumn.from_torch(t) for t in d_part]
    red_cols = [dfd.DeviceColumn.from_torch(t) for t in d_red]

    dbg = os.environ.get("DFD_BENCH_DEBUG")

    def one_pass():
        rows_out = h2d = d2h = 0
        for a in range(0, n, chunk):
            if dbg:
                print(f"[agg rank {rank}] chunk at row {a} t={time.perf_counter():.3f}", file=sys.stderr, flush=True)
            m = min(chunk, n - a)
            with torch.cuda.stream(lib_stream):
                for j in range(5):
                    d_in[j][:m].copy_(h_in_t[j][a:a + m], non_blocking=True)
            h2d += m * 40
            part.partition(in_cols, m, part_cols, sync=False)
            _, out_starts = red.reduce(part_cols, m, part.part_starts_device_ptr(), N, red_cols)
            g = int(out_starts[-1])
            if world > 1:
                wcols, ss, sc = node.shuffle_partitioned(ex, red_cols, out_starts)
                g = int(sc.sum())
                for q in range(P):  # the consumer reads its segments from the window: copy them to the host
                    for r in range(world):
                        cnt = int(sc[q, r])
                        if cnt:
                            for j in range(5):
                                nv.check(nv.lib().dfd_memcpy_d2h(ctx.handle, h_red.columns[j].ctypes.data, wcols[j].values + int(ss[q, r]) * 8, cnt * 8))
            else:
                with torch.cuda.stream(lib_stream):
                    for j in range(5):
                        h_red_t[j][:g].copy_(d_red[j][:g], non_blocking=True)
                lib_stream.synchronize()
            rows_out += g
            d2h += g * 40
        return rows_out, h2d, d2h

    one_pass()
    times = []
    for _ in range(max(2, args.steps // 3)):
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        rows_out, h2d, d2h = one_pass()
        times.append(_allreduce(torch, dist, world, time.perf_counter() - t0, "max"))
    sec = sum(times) / len(times)
    tot_out = int(_allreduce(torch, dist, world, rows_out, "sum"))
    h2d_all = int(_allreduce(torch, dist, world, h2d, "sum"))  # (collectives: every rank, not only the printing one)
    d2h_all = int(_allreduce(torch, dist, world, d2h, "sum"))
    if rank == 0:
        _emit({"metric": "shuffle rows/sec, end to end with device-side PartialReduce (group key + 4 aggregate states, 65 536 groups)",
               "value": n_total / sec, "unit": "rows/s", "n_gpus": world, "steps": len(times), "warmup": 1, "ms_per_step": sec * 1e3,
               "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "int64", "data": "synthetic",
               "config": {"workload": f"agg: {n_total} Partial-aggregate rows x (key Int64 + sum, count, min, max Int64), Hash([key], {N}), chunks of "
                                      f"{chunk} rows: H2D -> k_tile_hist/k_scan_tiles/k_scatter -> PartialReduce -> "
                                      + ("pre-partitioned exchange -> " if world > 1 else "") + "D2H of the reduced rows only", "rows": n_total},
               "e2e": {"value": n_total / sec, "unit": "rows/s", "h2d_bytes_per_step": h2d_all, "d2h_bytes_per_step": d2h_all, "rows_after_reduce": tot_out,
                       "reduction": n_total / max(tot_out, 1)},
               "gpu_launches": int(ctx.metrics()["kernel_launches"])})
    ex.close()
    if world > 1:
        dist.destroy_process_group()


# --------------------------------------------------------------------------------------------- fixture ----

def fixture_table(n: int, seed: int = 3):
    """`n` random rows of the reference's own shuffle-bench schema (src/execution_plans/benchmarks/fixture.rs:13-33):
    id Int64, metric Float64, flag Boolean?, label Utf8?, category Dictionary<Int32, Utf8>?, raw UInt8, ts Timestamp(ns),
    count Int32, tags List<Utf8?>? — built with vectorised numpy / pyarrow kernels (the reference uses arrow's
    create_random_batch)."""
    import pyarrow as pa
    import pyarrow.compute as pc

    rng = np.random.Generator(np.random.PCG64(seed))
    pool = pa.array([f"label-{i:05d}-" + "x" * (i % 23) for i in range(4096)], type=pa.string())
    tagpool = pa.array([f"t{i}" + "y" * (i % 7) for i in range(512)], type=pa.string())

    def mask(p):
        return pa.array(rng.random(n) < p)

    label = pc.if_else(mask(0.1), pa.scalar(None, pa.string()), pc.take(pool, pa.array(rng.integers(0, 4096, n, dtype=np.int32))))
    flag = pc.if_else(mask(0.1), pa.scalar(None, pa.bool_()), pa.array(rng.random(n) < 0.5))
    cat_idx = pc.if_else(mask(0.1), pa.scalar(None, pa.int32()), pa.array(rng.integers(0, 16, n, dtype=np.int32)))
    category = pa.DictionaryArray.from_arrays(cat_idx, pa.array([f"category-{i}" for i in range(16)], type=pa.string()))
    lens = rng.integers(0, 4, n, dtype=np.int32)
    offs = np.zeros(n + 1, dtype=np.int32)
    np.cumsum(lens, out=offs[1:])
    ne = int(offs[-1])
    elems = pc.if_else(pa.array(rng.random(ne) < 0.1), pa.scalar(None, pa.string()), pc.take(tagpool, pa.array(rng.integers(0, 512, ne, dtype=np.int32))))
    tags = pa.ListArray.from_arrays(pa.array(offs), elems, mask=pa.array(rng.random(n) < 0.1))
    schema = pa.schema([pa.field("id", pa.int64(), False), pa.field("metric", pa.float64(), False), ("flag", pa.bool_()), ("label", pa.string()),
                        ("category", pa.dictionary(pa.int32(), pa.string())), pa.field("raw", pa.uint8(), False),
                        pa.field("ts", pa.timestamp("ns"), False), pa.field("count", pa.int32(), False), ("tags", pa.list_(pa.string()))])
    cols = [pa.array(rng.integers(-(2**62), 2**62, n, dtype=np.int64)), pa.array(rng.standard_normal(n)), flag, label, category,
            pa.array(rng.integers(0, 256, n).astype(np.uint8)), pa.array(rng.integers(0, 2**60, n, dtype=np.int64)).cast(pa.timestamp("ns")),
            pa.array(rng.integers(-(2**31), 2**31 - 1, n).astype(np.int32)), tags]
    return pa.Table.from_arrays(cols, schema=schema)


def run_fixture(args, torch, dfd, world):
    """The reference's own bench schema (9 columns: nullable booleans / strings, a dictionary, a List<Utf8>) in the reference's
    own batch size (8192 rows) through the host operator, HOST batches in, HOST batches out: what a worker's
    RepartitionExec(Hash) sees.  Every batch is appended to the open device chunk (bit-granular bitmap concatenation, string
    offsets re-based), so the kernels still run on 1 Mi-row chunks.  Parity: a 100 000-row prefix against the oracle's
    partition ids + pyarrow take, values and order.  CPU arm beside it: the same batches through oracle partition ids +
    arrow `take` per destination (what RepartitionExec does with arrow-rs), one thread per input partition."""
    import threading
    from concurrent.futures import ThreadPoolExecutor

    import pyarrow as pa

    from oracle import oracle as orc
    from tests.util import expected_partitions

    dist, rank, local_rank, ctx, ex = _dist_setup(torch, dfd, world)
    N = 8
    n = args.rows if args.rows != (1 << 26) else 1 << 22
    table = fixture_table(n)
    batches = table.to_batches(max_chunksize=8192)

    def through_operator(bs, check=None):
        op = dfd.RepartitionExec(ctx, table.schema, dfd.Partitioning.Hash([0], N), chunk_rows=1 << 20, pipeline_depth=3, pinned_pool_chunks=6)
        readers = [op.execute(p) for p in range(N)]
        got = [[] for _ in range(N)]

        def consume(p):
            for rb in readers[p]:
                got[p].append(rb if check else rb.num_rows)

        ths = [threading.Thread(target=consume, args=(p,)) for p in range(N)]
        t0 = time.perf_counter()
        for t in ths:
            t.start()
        for b in bs:
            op.push_batch(b)
        op.finish()
        for t in ths:
            t.join()
        dt = time.perf_counter() - t0
        st = op.stats()
        del readers
        op.close()
        return dt, st, got

    # ---- parity on a prefix
    m = min(n, 100_000)
    pre = table.slice(0, m)
    _, _, got = through_operator(pre.to_batches(max_chunksize=8192), check=True)
    dest = orc.partition_ids([pre.column(0).combine_chunks()], m, N)
    order, starts = expected_partitions(dest, N)
    bad = 0
    for p in range(N):
        want = pre.take(pa.array(order[starts[p]:starts[p + 1]]))
        have = pa.Table.from_batches(got[p], schema=table.schema) if got[p] else table.schema.empty_table()
        for name in table.column_names:
            a, b = have.column(name).combine_chunks(), want.column(name).combine_chunks()
            if pa.types.is_dictionary(a.type):
                a, b = a.dictionary_decode(), b.dictionary_decode()
            if not a.equals(b):
                bad += 1
    if bad:
        _emit({"workload": "fixture", "parity_checked": False, "mismatching_columns": bad})
        sys.exit(3)

    # ---- timed: the whole table in 8192-row batches
    through_operator(batches)
    times, st = [], None
    for _ in range(max(2, args.steps // 3)):
        dt, st, got = through_operator(batches)
        assert sum(sum(g) for g in got) == n
        times.append(dt)
    sec = sum(times) / len(times)

    # ---- CPU arm: partition ids + arrow take per destination, input partitions on a thread pool
    threads = min(32, os.cpu_count() or 1)
    sample = batches[: max(1, min(len(batches), (1 << 21) // 8192))]

    def cpu_one(b):
        d = orc.partition_ids([b.column(0)], b.num_rows, N)
        o, s = expected_partitions(d, N)
        return sum(b.take(pa.array(o[s[p]:s[p + 1]])).num_rows for p in range(N) if s[p + 1] > s[p])

    with ThreadPoolExecutor(threads) as tp:
        list(tp.map(cpu_one, sample[:threads]))
        t0 = time.perf_counter()
        rows_cpu = sum(tp.map(cpu_one, sample))
        cpu_sec = time.perf_counter() - t0
    if rank == 0:
        _emit({"metric": "shuffle rows/sec, end to end (reference bench schema, 8192-row batches)", "value": n / sec, "unit": "rows/s", "n_gpus": world,
               "steps": len(times), "warmup": 1, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
               "dtype": "mixed (i64/f64/bool/utf8/dictionary/u8/timestamp/i32/list<utf8>)", "data": "synthetic",
               "config": {"workload": f"fixture: {n} rows of the reference's 9-column bench schema (fixture.rs:13-33), {len(batches)} input batches of 8192 "
                                      f"rows, Hash([id], {N}), chunk_rows 1 Mi, host batches in / host batches out", "rows": n},
               "parity_checked": True, "parity_rows": m,
               "e2e": {"value": n / sec, "unit": "rows/s", "h2d_bytes_per_step": int(st["bytes_h2d"]), "d2h_bytes_per_step": int(st["bytes_d2h"]),
                       "output_batches": int(sum(len(g) for g in got)), "input_batches": len(batches),
                       "operator": {"pinned_chunks": int(st["pinned_chunks"]), "pinned_chunks_allocated": int(st["pinned_chunks_allocated"]),
                                    "pinned_chunks_reused": int(st["pinned_chunks_reused"]), "push_ms": st["ns_push"] / 1e6,
                                    "wait_d2h_ms": st["ns_wait_d2h"] / 1e6, "wait_pool_ms": st["ns_wait_pool"] / 1e6}},
               "cpu_baseline": {"value": rows_cpu / cpu_sec, "unit": "rows/s", "cores": threads, "kind": "port",
                                "sample": f"{rows_cpu} rows ({len(sample)} batches): oracle partition ids + pyarrow take per destination, "
                                          f"{threads} threads over input batches"},
               "gpu_launches": int(ctx.metrics()["kernel_launches"])})
    ex.close()
    if world > 1:
        dist.destroy_process_group()


def run(args, torch, dfd, world):
    {"cfg3": run_cfg3, "cfg4": run_cfg4, "cfg5": run_cfg5, "agg": run_agg, "fixture": run_fixture}[args.workload](args, torch, dfd, world)


# ------------------------------------------------------------------------------ CPU reference arm (cfg4) ----

def run_reference_cfg4(args):
    """`--impl reference --workload cfg4`: the oracle port of RepartitionExec(Hash([l_orderkey], 48)) over the same lineitem
    stand-in on the host cores (persistent pool, thread sweep) — the CPU side of the north star's '>= 10x on SF10 q5'."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    from oracle import oracle as orc

    n_total = args.rows if args.rows != (1 << 26) else 59_986_052
    cols = cfg4_columns(0, n_total)
    cols = [c if c.ndim == 1 else np.ascontiguousarray(c).view(np.dtype("V16")).reshape(-1) for c in cols]
    cores = os.cpu_count() or 1
    pool = orc.WorkerPool(cores)
    sample = [c[:1 << 24] for c in cols]
    sweep = {}
    for t in sorted({t for t in (1, 8, 16, 32, 64, 96, 128, 192, 256, cores) if t <= cores}):
        pool.repartition(sample, [0], 48, 8192, t)
        t0 = time.perf_counter()
        pool.repartition(sample, [0], 48, 8192, t)
        sweep[t] = (1 << 24) / (time.perf_counter() - t0)
    best = max(sweep, key=sweep.get)
    times = []
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        counts, _ = pool.repartition(cols, [0], 48, 8192, best)
        if i >= args.warmup:
            times.append(time.perf_counter() - t0)
    pool.close()
    ms = 1e3 * sum(times) / len(times)
    v = n_total / (ms / 1e3)
    _emit({"impl": "reference", "metric": "shuffle rows/sec (TPC-H SF10 q5 lineitem repartition on l_orderkey, Hash N=48)", "value": v, "unit": "rows/s",
           "n_gpus": args.gpus, "steps": len(times), "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
           "vs_baseline": None, "dtype": "int64/decimal128", "data": "synthetic",
           "config": {"workload": "cfg4: lineitem stand-in, Hash([l_orderkey], 48), batch 8192", "rows_per_step": n_total},
           "cpu_baseline": {"value": v, "unit": "rows/s", "cores": best, "host_cores": cores, "kind": "port",
                            "threads_swept": {str(k): round(x) for k, x in sweep.items()},
                            "sample": "full table per step; oracle port of RepartitionExec(Hash) + coalescer on a persistent thread pool"},
           "e2e": {"value": v, "unit": "rows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
