"""The C-ABI shared library: loads, exports every declared symbol, fails loudly without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from xbitinfo_b200 import _native as nat
from conftest import ROOT, _have_gpu


def declared_functions():
    text = open(os.path.join(ROOT, "include", "xbitinfo_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(xbi_[a-z_0-9]+)\s*\(", text)))


def test_library_is_built_in_tree():
    assert os.path.exists(nat.LIB_PATH), "run `python -m xbitinfo_b200.csrc.build`"
    assert nat.lib().xbi_abi_version() == 1


def test_every_declared_symbol_is_exported():
    lib = C.CDLL(nat.LIB_PATH)
    names = declared_functions()
    assert len(names) >= 10
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/xbitinfo_b200.h but not exported"
    assert set(names) == set(nat.EXPORTED_SYMBOLS)


def test_workspace_size_and_counters():
    assert 1024 <= nat.lib().xbi_workspace_bytes() <= 1 << 16
    assert nat.kernel_launches() >= 0
    assert nat.fast_path_launches() <= nat.kernel_launches()


def test_argument_errors_do_not_need_a_gpu():
    x = np.zeros(16, dtype="float32")
    lib = nat.lib()
    rc = lib.xbi_bitround_host(x.ctypes.data_as(C.c_void_p), x.ctypes.data_as(C.c_void_p), nat.F32, 16, 24, 0)
    assert rc == nat.ERR_ARG and b"keepbits too large" in lib.xbi_last_error()
    rc = lib.xbi_bitround_host(x.ctypes.data_as(C.c_void_p), x.ctypes.data_as(C.c_void_p), nat.U32, 16, 3, 0)
    assert rc == nat.ERR_ARG
    with pytest.raises(ValueError):
        nat.check(rc)
    mi = np.empty(32)
    rc = lib.xbi_bitinformation_host(x.ctypes.data_as(C.c_void_p), 99, 1, 16, 1, 0, 0, 0, 0.99,
                                     mi.ctypes.data_as(C.c_void_p), None, 0)
    assert rc == nat.ERR_ARG
    rc = lib.xbi_bitinformation_host(x.ctypes.data_as(C.c_void_p), nat.U32, 1, 16, 1, nat.MASK_NAN, 0, 0, 0.99,
                                     mi.ctypes.data_as(C.c_void_p), None, 0)
    assert rc == nat.ERR_ARG


def test_chunk_grid_argument_errors_do_not_need_a_gpu():
    lib = nat.lib()
    i64 = nat.int64_array
    rc = lib.xbi_count_pairs_chunked(None, nat.F32, 0, i64([]), i64([]), 0, 0, 0, None, None)
    assert rc == nat.ERR_ARG and b"ndim" in lib.xbi_last_error()
    rc = lib.xbi_count_pairs_chunked(None, nat.F32, 2, i64([4, 4]), i64([2, 2]), 2, 0, 0, None, None)
    assert rc == nat.ERR_ARG and b"axis" in lib.xbi_last_error()
    rc = lib.xbi_count_pairs_chunked(None, nat.F32, 2, i64([4, 4]), i64([2, 0]), 1, 0, 0, None, None)
    assert rc == nat.ERR_ARG and b"chunk extents" in lib.xbi_last_error()
    rc = lib.xbi_count_pairs_chunked(None, nat.F32, 6, i64([2] * 6), i64([1] * 6), 0, 0, 0, None, None)
    assert rc == nat.ERR_UNSUPPORTED
    with pytest.raises(NotImplementedError):
        nat.check(rc)
    rc = lib.xbi_bitround_chunked(None, nat.U32, 1, i64([4]), i64([2]), None, None)
    assert rc == nat.ERR_ARG and b"float" in lib.xbi_last_error()
    rc = lib.xbi_mutual_information_batched(None, 65, 1, 0, 0.99, None, None)
    assert rc == nat.ERR_ARG
    dummy = C.c_void_p(8)  # never dereferenced: planning fails first
    rc = lib.xbi_count_pairs_chunked_v(None, nat.F32, 1, i64([6]), i64([2]), i64([0, 4, 5]), dummy, 0, 0, 0, None, None)
    assert rc == nat.ERR_ARG and b"from 0 to the extent" in lib.xbi_last_error()
    rc = lib.xbi_count_pairs_chunked_v(None, nat.F32, 1, i64([6]), i64([2]), i64([0, 6, 6]), dummy, 0, 0, 0, None, None)
    assert rc == nat.ERR_ARG and b"ascending" in lib.xbi_last_error()
    rc = lib.xbi_bitround_chunked_v(None, nat.F32, 1, i64([6]), None, None, None, None, None)
    assert rc == nat.ERR_ARG
    # nothing to do: no chunks, no pointers needed
    assert lib.xbi_count_pairs_chunked(None, nat.F32, 2, i64([0, 4]), i64([2, 2]), 1, 0, 0, None, None) == 0
    assert lib.xbi_mutual_information_batched(None, 32, 0, 0, 0.99, None, None) == 0


@pytest.mark.skipif(_have_gpu(), reason="only meaningful on a box without a GPU")
def test_product_path_fails_loudly_without_a_gpu():
    from xbitinfo_b200 import engine

    x = np.zeros((4, 8), dtype="float32")
    with pytest.raises(nat.XbiError, match="no CUDA device"):
        engine.bitinformation(x, 1)
    with pytest.raises(nat.XbiError, match="no CUDA device"):
        engine.bitround(x, 3)


def test_product_package_does_not_import_the_oracle():
    import importlib
    import sys

    for m in list(sys.modules):
        if m.startswith("xbitinfo_b200"):
            importlib.import_module(m)
    pkg = os.path.join(ROOT, "xbitinfo_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f


def test_header_is_plain_c_and_links(tmp_path):
    """include/xbitinfo_b200.h compiles as C99 with warnings as errors and a C program links against the
    shared library: the boundary a cgo / JNI / C binding would use."""
    import shutil
    import subprocess

    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    exe = tmp_path / "abi_consumer"
    libdir = os.path.dirname(nat.LIB_PATH)
    cmd = [gcc, "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "c", "abi_consumer.c"), "-o", str(exe), "-L", libdir, "-lxbitinfo_b200",
           f"-Wl,-rpath,{libdir}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0 and "abi consumer ok" in r.stdout, (r.returncode, r.stdout, r.stderr)


def test_host_copy_pool(tmp_path):
    """The thread pool behind the pageable-memory route of the host entry points, on CPU: compiled from
    tests/c/hostpool_check.cu against the object the library is linked from."""
    import shutil
    import subprocess

    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    csrc = os.path.dirname(nat.LIB_PATH)
    obj = os.path.join(csrc, "build", "xbi_hostpool.o")
    if not os.path.exists(nvcc) or not os.path.exists(obj):
        pytest.skip("nvcc or the built object is not available")
    exe = tmp_path / "hostpool_check"
    cmd = [nvcc, "-O2", "-std=c++17", "-I", csrc, os.path.join(ROOT, "tests", "c", "hostpool_check.cu"), obj,
           "-o", str(exe), "-Xcompiler", "-pthread", "-lpthread", "-Wno-deprecated-gpu-targets"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    for threads in ("1", "3", "8"):
        r = subprocess.run([str(exe)], capture_output=True, text=True, env=dict(os.environ, XBI_HOST_THREADS=threads),
                           timeout=300)
        assert r.returncode == 0 and f"host pool ok ({threads} threads)" in r.stdout, (threads, r.returncode, r.stdout, r.stderr)
