"""Host mirror of the reference's operator surface for the shuffle path.

`RepartitionExec` ≙ DataFusion's `RepartitionExec(Partitioning::Hash)` as the
reference builds it (src/execution_plans/network_shuffle.rs:126-134) and runs
it on every producer task (src/worker/impl_execute_task.rs:77-86):
`execute(partition)` returns the stream of that destination's record batches.
Host Arrow batches in, host Arrow batches out; the work happens on the GPU.
"""
from __future__ import annotations

import ctypes as C
from typing import Iterable, Optional

from . import _native as nv
from .device import WorkerContext
from .partitioner import Partitioning


class PinnedTable:
    """Column buffers in pinned host memory (what an Arrow allocator plugged into
    `dfd_host_alloc` gives the upstream operator), exposed as pyarrow arrays."""

    def __init__(self, ctx: WorkerContext, n_rows: int, dtypes):
        import numpy as np

        self.ctx = ctx
        self.n_rows = n_rows
        self._ptrs = []
        self.columns = []
        ctx._adopt(self)
        for dt in dtypes:
            dt = np.dtype(dt)
            nbytes = max(n_rows * dt.itemsize, 16)
            p = C.c_void_p()
            nv.check(nv.lib().dfd_host_alloc(ctx.handle, nbytes, C.byref(p)))
            self._ptrs.append(p)
            buf = (C.c_char * nbytes).from_address(p.value)
            self.columns.append(np.frombuffer(buf, dtype=dt, count=n_rows))

    def record_batches(self, names, batch_rows: int):
        """Zero-copy pyarrow RecordBatches over the pinned buffers."""
        import pyarrow as pa

        out = []
        for lo in range(0, self.n_rows, batch_rows):
            hi = min(lo + batch_rows, self.n_rows)
            arrays = []
            for col in self.columns:
                t = pa.from_numpy_dtype(col.dtype)
                buf = pa.foreign_buffer(col.ctypes.data + lo * col.dtype.itemsize, (hi - lo) * col.dtype.itemsize, base=self)
                arrays.append(pa.Array.from_buffers(t, hi - lo, [None, buf]))
            out.append(pa.RecordBatch.from_arrays(arrays, names=list(names)))
        return out

    def close(self):
        for p in self._ptrs:
            if p and self.ctx.handle:
                nv.lib().dfd_host_free(self.ctx.handle, p)
        self._ptrs = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class RepartitionExec:
    """`RepartitionExec::try_new(input, Partitioning::Hash(exprs, n))` on one GPU worker."""

    def __init__(self, ctx: WorkerContext, schema, partitioning: Partitioning, chunk_rows: int = 0,
                 pipeline_depth: int = 0, pinned_pool_chunks: int = 0, max_pinned_chunks: int = 0):
        self.ctx = ctx
        self.schema = schema
        self.partitioning = partitioning
        cs = nv.ArrowSchemaStruct()
        schema._export_to_c(C.addressof(cs))
        keys = (C.c_int32 * len(partitioning.key_cols))(*partitioning.key_cols)
        opts = nv.DfdExecOptions(chunk_rows, pipeline_depth, pinned_pool_chunks, max_pinned_chunks, 0)
        self._h = C.c_void_p()
        try:
            nv.check(nv.lib().dfd_repartition_exec_create(ctx.handle, C.byref(cs), keys, len(partitioning.key_cols),
                                                          partitioning.partition_count, C.byref(opts), C.byref(self._h)))
        finally:
            if cs.release:  # the operator only borrows the schema
                C.CFUNCTYPE(None, C.c_void_p)(cs.release)(C.addressof(cs))
        ctx._adopt(self)

    def name(self) -> str:
        return "RepartitionExec"

    def output_partitioning(self) -> Partitioning:
        return self.partitioning

    def push_batch(self, batch):
        """Feed one input RecordBatch (≙ one item of the child plan's stream)."""
        ca = nv.ArrowArrayStruct()
        batch._export_to_c(C.addressof(ca))
        nv.check(nv.lib().dfd_repartition_exec_push(self._h, C.byref(ca)))

    def finish(self):
        nv.check(nv.lib().dfd_repartition_exec_finish(self._h))

    def abort(self, message: str):
        """The producer's input failed: every partition stream ends with an error carrying `message`
        (≙ RepartitionExec forwarding an input error to all of its output partitions)."""
        nv.check(nv.lib().dfd_repartition_exec_abort(self._h, message.encode()))

    def run(self, reader):
        """Pull a pyarrow RecordBatchReader (≙ child.execute()) to exhaustion."""
        cs = nv.ArrowArrayStreamStruct()
        reader._export_to_c(C.addressof(cs))
        nv.check(nv.lib().dfd_repartition_exec_run(self._h, C.byref(cs)))

    def execute(self, partition: int):
        """≙ ExecutionPlan::execute(partition, ctx): a RecordBatchReader of that destination."""
        import pyarrow as pa

        cs = nv.ArrowArrayStreamStruct()
        nv.check(nv.lib().dfd_repartition_exec_execute(self._h, partition, C.byref(cs)))
        return pa.RecordBatchReader._import_from_c(C.addressof(cs))

    def stats(self) -> dict:
        st = nv.DfdExecStats()
        nv.check(nv.lib().dfd_repartition_exec_stats(self._h, C.byref(st)))
        return {k: getattr(st, k) for k, _ in st._fields_}

    def close(self):
        if self._h and self.ctx.handle:
            nv.lib().dfd_repartition_exec_destroy(self._h)
        self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
