"""-m "not gpu": the C-ABI shared library loads and exports every symbol include/cutfemx_b200.h
declares; host-only entry points behave; without a CUDA device the library fails LOUDLY (there is
no CPU fallback in the product)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from cutfemx_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "cutfemx_b200.h")


def _declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cfx_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_list_agree():
    assert sorted(_lib.SYMBOLS) == _declared_symbols()


def test_library_exports_every_declared_symbol(built_lib):
    L = C.CDLL(built_lib)
    missing = [s for s in _declared_symbols() if not hasattr(L, s)]
    assert not missing, missing
    # and nothing outside the cfx_ namespace leaks as a dynamic text symbol of ours
    out = subprocess.run(["nm", "-D", "--defined-only", built_lib], capture_output=True, text=True).stdout
    ours = [ln.split()[-1] for ln in out.splitlines() if " T " in ln and "cfx" in ln.split()[-1].lower()]
    extern_c = [s for s in ours if not s.startswith("_Z")]
    assert sorted(extern_c) == _declared_symbols()


def test_library_is_sm100a_only(built_lib):
    out = subprocess.run(["cuobjdump", "--list-elf", built_lib], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_library_does_not_link_the_oracle(built_lib):
    out = subprocess.run(["ldd", built_lib], capture_output=True, text=True).stdout
    assert "oracle" not in out
    for root, _, files in os.walk(os.path.join(ROOT, "cutfemx_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(root, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "liboracle" not in txt, f


def test_version_and_host_only_calls(built_lib):
    L = _lib.lib()
    assert L.cfx_version() == 100
    n = C.c_int()
    assert L.cfx_simplex_rule(3, 4, C.byref(n), None, None, 0) == 0
    assert n.value == 14
    assert L.cfx_simplex_rule(2, -1, C.byref(n), None, None, 0) != 0  # cut.cpp:164-168: order must be >= 0
    assert b"order" in L.cfx_last_error(None)
    p = np.zeros(14 * 3)
    w = np.zeros(14)
    assert L.cfx_simplex_rule(3, 4, C.byref(n), C.c_void_p(p.ctypes.data), C.c_void_p(w.ctypes.data), 3) != 0
    assert L.cfx_pattern_sizes(None, None, None) != 0
    assert L.cfx_list_size(None) == 0 and L.cfx_launch_count(None) == 0


@pytest.mark.parametrize("dim", [1, 2, 3])
@pytest.mark.parametrize("order", [0, 1, 2, 3, 4, 5, 6, 7])
def test_builtin_rules_match_oracle_tables(built_lib, dim, order):
    """The product's sub-simplex tables (csrc/quadrature.cu) and the oracle's (oracle/rules.py) were
    written independently; same points/weights up to ordering => pointwise-comparable rules."""
    from oracle import rules as R

    L = _lib.lib()
    n = C.c_int()
    assert L.cfx_simplex_rule(dim, order, C.byref(n), None, None, 0) == 0
    p = np.zeros(n.value * dim)
    w = np.zeros(n.value)
    assert L.cfx_simplex_rule(dim, order, C.byref(n), C.c_void_p(p.ctypes.data), C.c_void_p(w.ctypes.data),
                              n.value) == 0
    po, wo = R.simplex_rule(dim, order)
    po = np.asarray(po).reshape(-1, dim)
    assert wo.size == n.value
    np.testing.assert_allclose(p.reshape(-1, dim), po, rtol=0, atol=4e-15)
    np.testing.assert_allclose(w, wo, rtol=0, atol=1e-15)


def test_no_cpu_fallback():
    """Without a CUDA device every compute entry point is unreachable: ctx creation fails with a
    message (status CFX_ERR_CUDA), and the Python mirror raises CfxError."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present: the failure path is exercised on the CPU box")
    L = _lib.lib()
    h = C.c_void_p()
    rc = L.cfx_ctx_create(0, None, C.byref(h))
    assert rc == -2 and not h.value
    assert b"no CUDA device" in L.cfx_last_error(None) and b"no CPU fallback" in L.cfx_last_error(None)
    import cutfemx_b200 as cfx
    from cutfemx_b200 import mesh as M

    mesh = M.create_rectangle(2, 2)
    V = M.functionspace(mesh, 1)
    phi = M.Function(V, "phi").interpolate(lambda x, y, z: x - 0.1)
    with pytest.raises(cfx.CfxError):
        cfx.cut(phi)


def test_host_patch_compiles_and_links(built_lib, tmp_path):
    """host_patch/cutfemx_gpu_seams.cpp -- the C++ a CutFEMx maintainer adds next to cut/cut.cpp (SURVEY.md section 7
    step 2) -- compiles against include/cutfemx_b200.h and the DOLFINx / CutCells stand-in headers, and every library
    call it makes resolves against libcutfemx_b200.so (-Wl,--no-undefined)."""
    import shutil
    import subprocess

    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("g++ not available")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = tmp_path / "host_patch_check.so"
    cmd = [gxx, "-std=c++20", "-Wall", "-Wextra", "-Werror", "-DCUTFEMX_HOST_PATCH_STUBS", "-I", os.path.join(root, "include"),
           "-I", os.path.join(root, "host_patch", "stubs"), "-fPIC", "-shared",
           os.path.join(root, "host_patch", "cutfemx_gpu_seams.cpp"), "-o", str(out),
           "-L", os.path.dirname(built_lib), "-lcutfemx_b200", "-Wl,--no-undefined"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert out.exists()
