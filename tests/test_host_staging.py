"""CPU tests of the host operator's staging arithmetic (datafusion_distributed_b200/csrc/dfd_host_staging.h — the exact
functions dfd_exec.cu calls), compiled with g++ and driven with pyarrow arrays: bit-granular bitmap concatenation,
Utf8View <-> offsets + bytes, List<Utf8> rows -> the hidden lengths / bytes / validity columns."""
import ctypes as C
import os
import random
import subprocess

import numpy as np
import pyarrow as pa
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def shim(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("staging") / "libhost_staging_shim.so")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-Wall", "-Wextra", "-Werror", "-shared", "-fPIC", "-I",
                           os.path.join(ROOT, "datafusion_distributed_b200", "csrc"), os.path.join(ROOT, "tests", "cpp", "host_staging_shim.cpp"), "-o", out])
    lib = C.CDLL(out)
    lib.t_view_offsets.restype = C.c_int64
    lib.t_split_list_rows.restype = C.c_int64
    return lib


def ptr(a):
    return C.c_void_p(a.ctypes.data)


def bits_of(buf, lo, n):
    a = np.frombuffer(buf, dtype=np.uint8)
    return np.unpackbits(a, bitorder="little")[lo:lo + n]


def test_append_bits_concatenates_any_alignment(shim):
    rnd = random.Random(7)
    for _ in range(300):
        pieces, want = [], []
        dst = np.full(600, 0xAA, dtype=np.uint8)  # dirty destination: the function must not rely on zeroed memory past `at`
        at = 0
        for _ in range(rnd.randint(1, 6)):
            n = rnd.choice([0, 1, 7, 8, 9, 63, 64, 65, rnd.randint(0, 700)])
            lo = rnd.randint(0, 70)
            if rnd.random() < 0.25:
                src, bits = None, np.ones(n, dtype=np.uint8)
            else:
                src = np.frombuffer(rnd.randbytes((lo + n + 7) // 8 + 2), dtype=np.uint8).copy()
                bits = np.unpackbits(src, bitorder="little")[lo:lo + n]
            shim.t_append_bits(ptr(dst), C.c_int64(at), ptr(src) if src is not None else None, C.c_int64(lo), C.c_int64(n))
            pieces.append(src)
            want.append(bits)
            at += n
        want = np.concatenate(want) if want else np.zeros(0, dtype=np.uint8)
        got = np.unpackbits(dst, bitorder="little")
        assert np.array_equal(got[:at], want)
        if at % 8:  # the rest of the last byte is left zero so the next append continues cleanly
            assert not got[at:(at + 7) // 8 * 8].any()


def random_strings(rnd, n, null_p=0.1):
    out = []
    for _ in range(n):
        if rnd.random() < null_p:
            out.append(None)
        else:
            out.append("".join(rnd.choice("abcdefghijklmnopqrstuvwxyz0123456789") for _ in range(rnd.choice([0, 1, 4, 11, 12, 13, 20, 40]))))
    return out


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_views_to_offsets_and_bytes_and_back(shim, seed):
    rnd = random.Random(seed)
    values = random_strings(rnd, 5000)
    arr = pa.array(values, type=pa.string_view())
    # several variadic data buffers: concatenate two arrays with their own buffers
    arr = pa.concat_arrays([arr, pa.array(random_strings(rnd, 3000), type=pa.string_view())])
    values = arr.to_pylist()
    for lo, n in [(0, len(arr)), (13, 1000), (4999, 700), (len(arr) - 1, 1), (17, 0)]:
        sl = arr.slice(lo, n)
        bufs = sl.buffers()  # [validity, views, data...]
        views = np.frombuffer(bufs[1], dtype=np.uint8)
        valid = np.frombuffer(bufs[0], dtype=np.uint8) if bufs[0] is not None else None
        data_ptrs = (C.c_void_p * max(1, len(bufs) - 2))(*[b.address for b in bufs[2:]])
        off = np.empty(n + 1, dtype=np.int32)
        total = shim.t_view_offsets(ptr(views), ptr(valid) if valid is not None else None, C.c_int64(sl.offset), C.c_int64(n), ptr(off))
        want = [v.encode() if v is not None else b"" for v in values[lo:lo + n]]
        assert total == sum(len(w) for w in want)
        assert off[0] == 0 and np.array_equal(np.diff(off), [len(w) for w in want])
        out = np.zeros(total + 16, dtype=np.uint8)
        shim.t_view_bytes(ptr(views), data_ptrs, C.c_int64(sl.offset), C.c_int64(n), ptr(off), ptr(out))
        assert out[:total].tobytes() == b"".join(want)
        # and back: views over ONE data buffer must read as the same strings through pyarrow
        v2 = np.zeros(16 * max(n, 1), dtype=np.uint8)
        shim.t_build_views(ptr(off), ptr(out), C.c_int64(n), ptr(v2))
        rebuilt = pa.Array.from_buffers(pa.string_view(), n, [None, pa.py_buffer(v2.tobytes()), pa.py_buffer(out.tobytes())])
        rebuilt.validate(full=True)
        assert rebuilt.to_pylist() == [w.decode() for w in want]


def test_view_offsets_refuse_more_than_2gib(shim):
    views = np.zeros(16 * 3, dtype=np.uint8)
    views.view(np.int32)[0::4] = 0x7FFFFFF0 // 3 + 100  # three long views: the running total passes 2^31 - 1 at the third
    off = np.empty(4, dtype=np.int32)
    assert shim.t_view_offsets(ptr(views), None, C.c_int64(0), C.c_int64(2), ptr(off)) > 0
    assert shim.t_view_offsets(ptr(views), None, C.c_int64(0), C.c_int64(3), ptr(off)) == -1


@pytest.mark.parametrize("seed,elem_nulls", [(5, True), (6, False)])
def test_list_rows_split_into_lengths_bytes_validity(shim, seed, elem_nulls):
    rnd = random.Random(seed)
    rows = []
    for _ in range(4000):
        if rnd.random() < 0.1:
            rows.append(None)
        else:
            rows.append(random_strings(rnd, rnd.randint(0, 4), 0.15 if elem_nulls else 0.0))
    arr = pa.array(rows, type=pa.list_(pa.string()))
    for lo, n in [(0, len(arr)), (7, 1234), (3999, 1), (100, 0)]:
        sl = arr.slice(lo, n)
        child = sl.values  # NOT sliced by pyarrow's .values: the list offsets address it from element 0 of the child
        loff = np.frombuffer(sl.buffers()[1], dtype=np.int32)
        cb = child.buffers()
        coff = np.frombuffer(cb[1], dtype=np.int32)[child.offset:]
        cvalid = np.frombuffer(cb[0], dtype=np.uint8) if cb[0] is not None else None
        e0, e1 = int(loff[sl.offset]), int(loff[sl.offset + n])
        ne = e1 - e0
        len_off, bytes_off, valid_off = (np.empty(n + 1, dtype=np.int32) for _ in range(3))
        lengths, valid_bytes = np.empty(ne + 4, dtype=np.int32), np.empty(ne + 16, dtype=np.uint8)
        got_ne = shim.t_split_list_rows(ptr(loff), ptr(coff), ptr(cvalid) if cvalid is not None else None, C.c_int64(child.offset), C.c_int64(sl.offset),
                                        C.c_int64(n), ptr(len_off), ptr(bytes_off), ptr(lengths), ptr(valid_off), ptr(valid_bytes))
        assert got_ne == ne
        py = rows[lo:lo + n]
        elems = [e for r in py if r is not None for e in r]
        assert ne == len(elems)
        per_row = [0 if r is None else len(r) for r in py]
        assert np.array_equal(np.diff(len_off), [4 * k for k in per_row]) and len_off[0] == 0
        assert np.array_equal(np.diff(valid_off), per_row) and valid_off[0] == 0
        assert np.array_equal(lengths[:ne], [0 if e is None else len(e) for e in elems])
        assert np.array_equal(valid_bytes[:ne], [0 if e is None else 1 for e in elems])
        assert np.array_equal(np.diff(bytes_off), [0 if r is None else sum(len(e) for e in r if e is not None) for r in py]) and bytes_off[0] == 0
        data = np.frombuffer(cb[2], dtype=np.uint8) if cb[2] is not None else np.zeros(0, dtype=np.uint8)
        start = int(coff[e0])
        assert data[start:start + int(bytes_off[n])].tobytes() == "".join(e for e in elems if e is not None).encode()


def test_list_rows_with_a_sliced_child_array(shim):
    """The child of a ListArray may itself carry an array offset (children[0]->offset != 0 in the C Data export): child
    offsets and child validity are addressed relative to it."""
    rnd = random.Random(11)
    pool = pa.array(random_strings(rnd, 300, 0.2), type=pa.string())
    k = 37
    child = pool.slice(k)  # offset 37 into shared buffers
    sizes = [rnd.randint(0, 3) for _ in range(80)]
    offs = np.zeros(len(sizes) + 1, dtype=np.int32)
    np.cumsum(sizes, out=offs[1:])
    arr = pa.ListArray.from_arrays(pa.array(offs), child)
    assert arr.values.offset == k
    rows = arr.to_pylist()
    lo, n = 5, 60
    loff = np.frombuffer(arr.buffers()[1], dtype=np.int32)
    cb = arr.values.buffers()
    coff = np.frombuffer(cb[1], dtype=np.int32)[k:]
    cvalid = np.frombuffer(cb[0], dtype=np.uint8)
    e0, e1 = int(loff[lo]), int(loff[lo + n])
    ne = e1 - e0
    len_off, bytes_off, valid_off = (np.empty(n + 1, dtype=np.int32) for _ in range(3))
    lengths, valid_bytes = np.empty(ne + 4, dtype=np.int32), np.empty(ne + 16, dtype=np.uint8)
    assert shim.t_split_list_rows(ptr(loff), ptr(coff), ptr(cvalid), C.c_int64(k), C.c_int64(lo), C.c_int64(n), ptr(len_off), ptr(bytes_off),
                                  ptr(lengths), ptr(valid_off), ptr(valid_bytes)) == ne
    elems = [e for r in rows[lo:lo + n] for e in r]
    assert np.array_equal(lengths[:ne], [0 if e is None else len(e) for e in elems])
    assert np.array_equal(valid_bytes[:ne], [0 if e is None else 1 for e in elems])
    data = np.frombuffer(cb[2], dtype=np.uint8)
    start = int(coff[e0])
    assert data[start:start + int(bytes_off[n])].tobytes() == "".join(e for e in elems if e is not None).encode()


def test_list_of_primitives_rows_split(shim):
    """List<Int64> rows -> element counts (through the lengths column's offsets), the rows' byte ranges of the child values and one
    validity byte per element, with a sliced list and a sliced child."""
    rnd = random.Random(13)
    pool = pa.array([None if rnd.random() < 0.2 else rnd.getrandbits(40) for _ in range(500)], type=pa.int64())
    k = 11
    child = pool.slice(k)
    sizes = [rnd.randint(0, 4) for _ in range(120)]
    offs = np.zeros(len(sizes) + 1, dtype=np.int32)
    np.cumsum(sizes, out=offs[1:])
    arr = pa.ListArray.from_arrays(pa.array(offs), child)
    rows = arr.to_pylist()
    lo, n, w = 9, 100, 8
    loff = np.frombuffer(arr.buffers()[1], dtype=np.int32)
    cb = arr.values.buffers()
    cvalid = np.frombuffer(cb[0], dtype=np.uint8)
    e0, e1 = int(loff[lo]), int(loff[lo + n])
    ne = e1 - e0
    len_off, bytes_off, valid_off = (np.empty(n + 1, dtype=np.int32) for _ in range(3))
    lengths, valid_bytes = np.empty(ne + 4, dtype=np.int32), np.empty(ne + 16, dtype=np.uint8)
    shim.t_split_list_rows_fixed.restype = C.c_int64
    assert shim.t_split_list_rows_fixed(ptr(loff), C.c_int32(w), ptr(cvalid), C.c_int64(k), C.c_int64(lo), C.c_int64(n), ptr(len_off), ptr(bytes_off),
                                        ptr(lengths), ptr(valid_off), ptr(valid_bytes)) == ne
    per_row = [len(r) for r in rows[lo:lo + n]]
    assert np.array_equal(np.diff(len_off), [4 * c for c in per_row]) and np.array_equal(np.diff(bytes_off), [w * c for c in per_row])
    assert np.array_equal(np.diff(valid_off), per_row) and (lengths[:ne] == w).all()
    elems = [e for r in rows[lo:lo + n] for e in r]
    assert np.array_equal(valid_bytes[:ne], [0 if e is None else 1 for e in elems])
    values = np.frombuffer(cb[1], dtype=np.int64)[k + e0:k + e1]
    assert [int(v) for v, e in zip(values, elems) if e is not None] == [e for e in elems if e is not None]
