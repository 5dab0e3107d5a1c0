"""Shared helpers for the parity tests (test infrastructure)."""
from __future__ import annotations

import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "hash_vectors.json")


def golden():
    with open(GOLDEN) as f:
        return json.load(f)


def cfg2_columns(n_rows: int, n_cols: int = 8, seed: int = 42):
    """SURVEY.md §8(d) cfg-2 shape: col0 = uniform i64 key, cols j>=1 = row_id*8+j."""
    rng = np.random.Generator(np.random.PCG64(seed))
    key = rng.integers(-(2**63), 2**63 - 1, n_rows, dtype=np.int64, endpoint=True)
    return [key] + [np.arange(n_rows, dtype=np.int64) * 8 + j for j in range(1, n_cols)]


def expected_partitions(dest: np.ndarray, num_partitions: int):
    """Stable per-destination row index lists from destination ids."""
    order = np.argsort(dest, kind="stable")
    counts = np.bincount(dest, minlength=num_partitions).astype(np.int64)
    starts = np.zeros(num_partitions + 1, dtype=np.int64)
    np.cumsum(counts, out=starts[1:])
    return order, starts
