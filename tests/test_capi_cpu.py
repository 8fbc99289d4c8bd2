"""CPU tests of the C-ABI shared library: it loads, exports exactly what include/amps_b200.h declares,
its struct layout matches the ctypes binding, and it refuses to run without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "amps_b200.h")


def declared_symbols():
    txt = open(HEADER).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"AMPS_B200_API\s+[\w\s\*]+?\b(amps_\w+)\s*\(", txt)))


@pytest.fixture(scope="module")
def capi():
    from gr_amps_b200 import capi as c
    if not os.path.exists(c.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    return c


def test_header_and_binding_agree(capi):
    syms = declared_symbols()
    assert len(syms) >= 20
    assert sorted(capi.EXPORTS) == syms


def test_library_exports_every_declared_symbol(capi):
    L = capi.lib()
    for name in declared_symbols():
        assert hasattr(L, name), name
    out = subprocess.check_output(["nm", "-D", "--defined-only", capi.LIB_PATH], text=True)
    exported = {line.split()[-1] for line in out.splitlines() if " T " in line}
    assert set(declared_symbols()) <= exported
    # -fvisibility=hidden: nothing but the C ABI (and toolchain symbols) leaks out
    leaked = [s for s in exported if not s.startswith("amps_") and "cuda" not in s.lower() and not s.startswith("_")]
    assert leaked == []


def test_abi_struct_sizes(capi):
    b, w = C.c_size_t(0), C.c_size_t(0)
    assert capi.lib().amps_b200_abi_sizes(C.byref(b), C.byref(w)) == 0
    assert (b.value, w.value) == (C.sizeof(capi.Burst), C.sizeof(capi.ReccWords)) == (5208, 1804)
    assert capi.Burst.symbols.offset == 24 and capi.Burst.decoded.offset == 3400


def test_no_cpu_fallback(capi):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    assert capi.lib().amps_b200_device_count() == 0
    with pytest.raises(capi.AmpsError) as e:
        capi.ReccIq(max_samples=38400)
    assert e.value.status == -2          # AMPS_E_NODEVICE
    with pytest.raises(capi.AmpsError):
        capi.ReccDecode()


def test_strerror(capi):
    L = capi.lib()
    assert L.amps_b200_strerror(0) == b"ok"
    assert b"sm_100" in L.amps_b200_strerror(-2)
    assert L.amps_b200_version() >= 100
