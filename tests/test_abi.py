"""The C-ABI library loads and exports every symbol include/wsb200.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "wsb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ws_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported_and_bound():
    import wsb200
    lib = wsb200.load()
    syms = declared_symbols()
    assert len(syms) >= 55
    for s in syms:
        assert hasattr(lib, s), f"libwsb200.so does not export {s}"
        assert s in wsb200._lib.SIGNATURES, f"python binding misses {s}"
    assert set(wsb200._lib.SIGNATURES) == set(syms)
    assert lib.ws_abi_version() == 2


def test_no_cpu_fallback():
    """Without a device every context creation fails loudly (skipped where a GPU exists)."""
    import wsb200
    n = ctypes.c_int()
    rc = wsb200.load().ws_device_count(ctypes.byref(n))
    if rc == 0 and n.value > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(wsb200.WsError) as e:
        wsb200.SMCState(10)
    assert e.value.code == -2 and "no CPU fallback" in str(e.value)


def test_product_never_imports_the_oracle():
    """no import / include / dlopen of anything under oracle/ from the product package"""
    pkg = os.path.join(ROOT, "weightedsampling.jl_b200")
    pat = re.compile(r"(^\s*(from|import)\s+oracle\b)|(#include\s*[\"<][^\n]*oracle)|(libws_oracle)|(oracle/[A-Za-z_]+\.(py|c|so))", re.M)
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")) or f == "Makefile":
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                assert not pat.search(txt), f


def _build_abi_smoke(tmp_path):
    """tests/host/abi_smoke.c: plain C99 against include/wsb200.h and libwsb200.so only"""
    import subprocess
    exe = os.path.join(str(tmp_path), "abi_smoke")
    libdir = os.path.join(ROOT, "weightedsampling.jl_b200", "lib")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-I" + os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "host", "abi_smoke.c"), "-L" + libdir, "-lwsb200", "-lm",
                    "-Wl,-rpath," + libdir, "-o", exe], check=True)
    return exe


def test_header_is_self_sufficient_for_a_plain_c_consumer(tmp_path):
    """the header compiles as C99 with -Wall -Wextra -Werror and links against the library: everything a binding
    needs (types, enums, prototypes) is in include/wsb200.h.  Without a GPU the program must fail loudly."""
    import subprocess
    exe = _build_abi_smoke(tmp_path)
    fixture = os.path.join(ROOT, "tests", "golden", "abi_smoke.bin")
    p = subprocess.run([exe, fixture], capture_output=True, text=True)
    if p.returncode != 0:
        assert p.returncode == 2 and "no CUDA device" in p.stderr, (p.returncode, p.stdout, p.stderr)


@pytest.mark.gpu
def test_plain_c_consumer_reproduces_the_oracle(tmp_path):
    """LGSSM + Resample every step + a RW move through the C ABI from C, on replayed streams, against the numbers the
    oracle's hand-built program wrote (tests/golden/make_golden.py: abi_smoke_fixture)."""
    import subprocess
    exe = _build_abi_smoke(tmp_path)
    p = subprocess.run([exe, os.path.join(ROOT, "tests", "golden", "abi_smoke.bin")], capture_output=True, text=True)
    print(p.stdout, p.stderr)
    assert p.returncode == 0, (p.stdout, p.stderr)
    assert "mismatching values 0" in p.stdout
