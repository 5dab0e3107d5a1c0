"""CPU tests of the drop-in boundary: the C-ABI library builds for sm_100a,
loads, exports every symbol include/dfd_b200.h declares, and fails loudly
(no CPU fallback) when there is no CUDA device."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "dfd_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dfd_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_what_python_binds(built):
    from datafusion_distributed_b200 import _native as nv

    assert sorted(nv.SIGNATURES) == declared_symbols()


def test_library_exports_every_declared_symbol(built):
    from datafusion_distributed_b200 import _native as nv

    L = nv.lib()
    for name in declared_symbols():
        assert hasattr(L, name), name
    assert L.dfd_abi_version() == 1
    assert L.dfd_status_name(6) == b"DFD_ERR_UNSUPPORTED"


def test_library_is_sm100a_and_has_the_kernels(built):
    from datafusion_distributed_b200 import LIB_PATH

    out = subprocess.run(["cuobjdump", "-lelf", LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out, out
    sass = subprocess.run(["cuobjdump", "-sass", LIB_PATH], capture_output=True, text=True).stdout
    for k in ("k_scatter", "k_tile_hist", "k_scan_tiles", "k_partition_ids"):
        assert k in sass, k
    assert "VOTE" in sass  # warp-ballot ranking is in the scatter kernel


def test_header_compiles_as_plain_c(tmp_path):
    src = tmp_path / "t.c"
    src.write_text('#include "dfd_b200.h"\nint main(void){ struct ArrowArray a; (void)a; return DFD_OK; }\n')
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-c", str(src),
                           "-o", str(tmp_path / "t.o")])


def test_no_cpu_fallback_without_gpu(built):
    """On a box without CUDA the context constructor must raise; with CUDA it must work."""
    import datafusion_distributed_b200 as dfd
    from datafusion_distributed_b200 import _native as nv

    n = C.c_int(-1)
    st = nv.lib().dfd_device_count(C.byref(n))
    if st == 0 and n.value > 0:
        pytest.skip("GPU present")
    with pytest.raises(dfd.DfdError) as ei:
        dfd.WorkerContext(0)
    assert ei.value.status == 3  # DFD_ERR_CUDA
    assert "no CPU fallback" in str(ei.value)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "datafusion_distributed_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                assert "oracle" not in open(os.path.join(dp, f)).read().lower(), os.path.join(dp, f)
