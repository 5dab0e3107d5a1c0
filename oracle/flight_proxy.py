"""oracle/flight_proxy.py — CPU + Arrow-Flight stand-in for the reference's whole shuffle path.

TEST / BASELINE INFRASTRUCTURE, NOT PRODUCT CODE (only bench.py's `--impl reference` leg and tests
import it).  The real path cannot be built here (Rust), so this restates its *shape* with the pieces
that exist in the image:

  producer task  : oracle port of RepartitionExec(Hash(keys, P*T_cons)) over its row range
                   (one thread per input partition, 8192-row batches)          impl_execute_task.rs:77-86
  server half    : pyarrow.flight server per producer; one `do_get` per (consumer, producer) streams
                   the consumer's P partitions as Arrow IPC record batches (LZ4_FRAME like the
                   reference's default `compression = "lz4"`, or none)         impl_execute_task.rs:171-207
  client half    : per consumer task, one Flight stream per producer, batches decoded into fresh
                   buffers                                                      worker_connection_pool.rs:239-380
  NetworkShuffleExec::execute : off = P*task_index, partitions off..off+P from every producer
                                                                                network_shuffle.rs:213-238
Transport is localhost gRPC (the reference's tests/benches use in-process workers too:
src/test_utils/localhost.rs:20-74).  tonic/arrow-flight (Rust) vs grpc++/pyarrow (C++) differ in
constant factors; the stand-in is labelled "port" everywhere it is reported.
"""
from __future__ import annotations

import threading
import time
from typing import List, Optional

import numpy as np
import pyarrow as pa
import pyarrow.flight as fl

from . import oracle as orc


class _Producer(fl.FlightServerBase):
    """One producer worker: serves its repartitioned output, `do_get(ticket = b"start,end")`."""

    def __init__(self, schema: pa.Schema, compression: Optional[str]):
        super().__init__("grpc://127.0.0.1:0")
        self.schema = schema
        self.parts: List[List[pa.RecordBatch]] = []
        self.options = pa.ipc.IpcWriteOptions(compression=compression)

    def set_output(self, parts):
        self.parts = parts

    def do_get(self, context, ticket):
        start, end = (int(x) for x in ticket.ticket.decode().split(","))
        batches = [b for p in range(start, end) for b in self.parts[p]]
        reader = pa.RecordBatchReader.from_batches(self.schema, batches)
        return fl.RecordBatchStream(reader, options=self.options)


class FlightShuffleProxy:
    """T_prod producers -> T_cons consumers, P partitions per consumer, Hash([key_col], P*T_cons)."""

    def __init__(self, names, producers: int, consumers: int, partitions_per_consumer: int, compression: Optional[str] = "lz4",
                 batch_size: int = 8192):
        self.names = list(names)
        self.T_prod, self.T_cons, self.P = producers, consumers, partitions_per_consumer
        self.N = self.P * self.T_cons
        self.batch_size = batch_size
        self.schema = pa.schema([(n, pa.int64()) for n in self.names])
        self.servers = [_Producer(self.schema, compression) for _ in range(producers)]
        self.threads = [threading.Thread(target=s.serve, daemon=True) for s in self.servers]
        for t in self.threads:
            t.start()
        self.ports = [s.port for s in self.servers]

    def close(self):
        for s in self.servers:
            s.shutdown()

    # -- producer side ---------------------------------------------------------------------
    def _repartition(self, cols, threads: int):
        """RepartitionExec(Hash) on one producer's rows -> per destination, 8192-row record batches."""
        outs, counts, starts = orc.repartition_table(cols, [0], self.N, self.batch_size, threads, materialize=True)
        parts = []
        for p in range(self.N):
            lo, hi = int(starts[p]), int(starts[p + 1])
            bl = []
            for a in range(lo, hi, self.batch_size):
                b = min(a + self.batch_size, hi)
                bl.append(pa.RecordBatch.from_arrays([pa.array(c[a:b]) for c in outs], names=self.names))
            parts.append(bl)
        return parts

    # -- one shuffle -------------------------------------------------------------------------
    def run(self, producer_cols: List[List[np.ndarray]], threads_per_producer: int = 1):
        """producer_cols[r] = producer r's columns.  Returns (seconds, rows_received, per-consumer tables).
        `self.last_phases` = (repartition seconds, exchange seconds): the real reference overlaps the two
        (consumers poll while producers are still partitioning), so callers may charge max() instead of the sum."""
        t0 = time.perf_counter()
        # stage N: every producer repartitions its rows (concurrently, like T worker processes)
        results = [None] * self.T_prod

        def produce(r):
            results[r] = self._repartition(producer_cols[r], threads_per_producer)

        pt = [threading.Thread(target=produce, args=(r,)) for r in range(self.T_prod)]
        for t in pt:
            t.start()
        for t in pt:
            t.join()
        for r in range(self.T_prod):
            self.servers[r].set_output(results[r])
        t1 = time.perf_counter()
        # stage N+1: every consumer task pulls its partition range from every producer
        rows = [0] * self.T_cons

        def pull(ci, r):  # one Flight stream per (consumer, producer) pair, all pairs concurrently (select_all)
            off = self.P * ci
            client = fl.connect(f"grpc://127.0.0.1:{self.ports[r]}")
            t = client.do_get(fl.Ticket(f"{off},{off + self.P}".encode())).read_all()
            client.close()
            with lock:
                rows[ci] += t.num_rows
                tables[ci][r] = t

        lock = threading.Lock()
        tables = [[None] * self.T_prod for _ in range(self.T_cons)]
        ct = [threading.Thread(target=pull, args=(ci, r)) for ci in range(self.T_cons) for r in range(self.T_prod)]
        for t in ct:
            t.start()
        for t in ct:
            t.join()
        t2 = time.perf_counter()
        self.last_phases = (t1 - t0, t2 - t1)
        return t2 - t0, sum(rows), tables
