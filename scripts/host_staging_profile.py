"""Where the host operator's producer thread spends its time on the reference's bench schema (8192-row batches), measured on
the CPU harness (tests/cpu_harness): the operator's own host logic (batch coalescing, bitmap / offset staging, list splitting,
slicing the output) separated from the stand-in kernels, the copies and the allocations of the stand-in runtime.

    python scripts/host_staging_profile.py [col,col,...]      # default: all nine columns of fixture.rs:13-33

Test infrastructure: it loads the harness library built under /tmp, never the product library."""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench_workloads as W  # noqa: E402
import tests.test_exec_cpu_harness as H  # noqa: E402


def main():
    tmp = "/tmp/dfd_host_profile"
    os.makedirs(tmp, exist_ok=True)
    lib = C.CDLL(H._build_harness(tmp))
    ns, ctx = H._make_namespace(lib), H._Ctx(lib)
    for f in ("harness_kernel_ns", "harness_copy_ns", "harness_copy_bytes", "harness_copy_calls", "harness_alloc_ns"):
        getattr(lib, f).restype = C.c_uint64
    n = 1 << 20
    table = W.fixture_table(n)
    if len(sys.argv) > 1:
        table = table.select(sys.argv[1].split(","))
    batches = table.to_batches(max_chunksize=8192)
    snap = lambda: (lib.harness_kernel_ns(), lib.harness_copy_ns(), lib.harness_copy_bytes(), lib.harness_copy_calls(), lib.harness_alloc_ns())  # noqa: E731
    for _ in range(3):
        ex = ns.RepartitionExec(ctx, table.schema, ns.Partitioning.Hash([0], 8), chunk_rows=1 << 20, pipeline_depth=3, pinned_pool_chunks=6)
        a = snap()
        t0 = time.perf_counter()
        for b in batches:
            ex.push_batch(b)
        ex.finish()
        wall = time.perf_counter() - t0
        st, b_ = ex.stats(), snap()
        for p in range(8):
            ex.execute(p).read_all()
        ex.close()
    k, c, nb, calls, al = (y - x for x, y in zip(a, b_))
    host = (st["ns_push"] - k - c - al) / 1e6
    print(f"columns {table.column_names}")
    print(f"{len(batches)} batches of 8192 rows: push {st['ns_push'] / 1e6:.1f} ms (wall {wall * 1e3:.1f} ms) = host logic {host:.1f} ms "
          f"({host * 1e3 / len(batches):.1f} us per batch) + stand-in kernels {k / 1e6:.1f} ms + allocations {al / 1e6:.1f} ms + copies "
          f"{c / 1e6:.1f} ms ({nb / 1e6:.0f} MB in {calls} calls)")
    ctx.close()


if __name__ == "__main__":
    main()
