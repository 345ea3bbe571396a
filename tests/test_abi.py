"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports
every symbol include/dpgo_b200.h declares, and refuses to compute without a
GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from dpgo_ros_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "dpgo_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(dpgo_b200_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_header_symbol():
    L = capi.lib()
    syms = header_symbols()
    assert len(syms) >= 45
    for s in syms:
        assert hasattr(L, s), f"libdpgo_b200.so does not export {s}"
    assert sorted(capi.SYMBOLS) == syms


def test_version_and_struct_sizes():
    L = capi.lib()
    assert b"sm_100a" in L.dpgo_b200_version()
    # the ctypes mirrors must match the C structs (layout: ints and doubles, natural alignment)
    assert C.sizeof(capi.Params) == 144
    assert C.sizeof(capi.OptResult) == 64
    assert C.sizeof(capi.Status) == 32


def test_bad_parameters_are_rejected_before_any_cuda_call():
    L = capi.lib()
    h = C.c_void_p()
    assert L.dpgo_b200_agent_create(0, C.byref(capi.make_params(d=2)), 0, C.byref(h)) == -1
    assert L.dpgo_b200_agent_create(0, C.byref(capi.make_params(r=9)), 0, C.byref(h)) == -1
    assert L.dpgo_b200_agent_create(3, C.byref(capi.make_params(num_robots=2)), 0, C.byref(h)) == -1
    assert L.dpgo_b200_agent_create(0, C.byref(capi.make_params(cost_type=2)), 0, C.byref(h)) == -1
    assert b"L2 and GNC_TLS" in L.dpgo_b200_last_error()


@pytest.mark.skipif(capi.lib().dpgo_b200_device_count() > 0, reason="only meaningful without a GPU")
def test_no_cpu_fallback_without_gpu():
    L = capi.lib()
    h = C.c_void_p()
    rc = L.dpgo_b200_agent_create(0, C.byref(capi.make_params()), 0, C.byref(h))
    assert rc == -3  # DPGO_B200_ERR_CUDA
    M = np.zeros((5, 8), order="F")
    out = np.zeros_like(M, order="F")
    dp = C.POINTER(C.c_double)
    assert L.dpgo_b200_manifold_project(0, 5, 2, M.ctypes.data_as(dp), out.ctypes.data_as(dp)) == -3
    t = C.c_void_p()
    assert L.dpgo_b200_team_create(0, C.byref(t)) == -3


def test_library_is_built_for_sm_100a_only():
    """Every cubin in the shared library targets sm_100a (write for B200 only; no multi-arch fallback), and the hot
    kernel carries the TMA bulk copy of the preconditioner slab (UBLKCP in SASS)."""
    import shutil
    import subprocess
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not available")
    out = subprocess.run(["cuobjdump", "--list-elf", capi.LIB_PATH], capture_output=True, text=True, timeout=120).stdout
    elfs = [l for l in out.splitlines() if l.startswith("ELF file")]
    assert len(elfs) >= 20
    assert all("sm_100a" in l for l in elfs), [l for l in elfs if "sm_100a" not in l]
    sass = subprocess.run(["cuobjdump", "-sass", "-fun", "_ZN4dpgo10k_team_runILi5ELi1ELb0EEEvNS_7TeamDevENS_7RunArgsE",
                           capi.LIB_PATH], capture_output=True, text=True, timeout=300).stdout
    assert sass.count("UBLKCP") >= 1 and sass.count("DFMA") > 500
    assert "ATOM" not in sass.replace("ATOMIC", "")   # no atomics on data in the hot kernel
