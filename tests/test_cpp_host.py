"""The C++ host mirror (include/dfd_b200.hpp): compiles with plain g++ against the C ABI; on a GPU
box the compiled self-test also runs the kernels and compares with the C oracle."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build_selftest(tmp_path):
    import datafusion_distributed_b200 as dfd
    from oracle import oracle as orc

    libdir = os.path.dirname(dfd.LIB_PATH)
    orc.build()
    exe = str(tmp_path / "test_host_mirror")
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "oracle"),
           os.path.join(ROOT, "tests", "cpp", "test_host_mirror.cpp"), "-o", exe, "-L", libdir, "-L", os.path.join(ROOT, "oracle"),
           "-ldfd_b200", "-ldf_oracle", f"-Wl,-rpath,{libdir}", f"-Wl,-rpath,{os.path.join(ROOT, 'oracle')}"]
    subprocess.check_call(cmd)
    return exe


def test_cpp_host_mirror_compiles_and_surface_checks(built, tmp_path):
    exe = build_selftest(tmp_path)
    out = subprocess.run([exe, "--no-gpu"], capture_output=True, text=True, timeout=60)
    assert "CPP_HOST_MIRROR_SURFACE_OK" in out.stdout, out.stdout + out.stderr


@pytest.mark.gpu
def test_cpp_host_mirror_against_oracle_on_gpu(built, tmp_path):
    exe = build_selftest(tmp_path)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert "CPP_HOST_MIRROR_OK" in out.stdout, out.stdout + out.stderr
