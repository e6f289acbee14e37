"""The C-ABI library loads, exports every symbol include/*.h declares, and refuses to run without a
CUDA device (no CPU fallback).  No compute calls here."""
import ctypes as C
import pathlib
import re

import pytest

ROOT = pathlib.Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def native():
    from bls_bn254_b200 import _native
    _native.build()
    return _native


def test_header_symbols_are_exported(native):
    hdr = (ROOT / "include" / "bls_bn254_b200.h").read_text()
    declared = sorted(set(re.findall(r"\b(bn254_[a-z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 30
    lib = C.CDLL(str(native.SO))
    for name in declared:
        assert hasattr(lib, name), "header declares %s but the library does not export it" % name
    assert sorted(native.EXPORTED_SYMBOLS) == declared


def test_every_header_entry_cites_the_reference():
    hdr = (ROOT / "include" / "bls_bn254_b200.h").read_text()
    for cite in ("g1.rs:910-928", "g2.rs:919-936", "pairings.rs:760-802", "pairings.rs:499-514", "error.rs:5-10", "g1.rs:283-328"):
        assert cite in hdr


def test_no_cpu_fallback(native):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import bls_bn254_b200 as b
    assert native.lib().bn254_device_count() == 0
    with pytest.raises(b.Bn254Error) as e:
        b.Context([0])
    assert e.value.code == -20


def test_product_does_not_touch_the_oracle(native):
    """the package never imports, includes, links or loads anything under oracle/ or tests/"""
    import subprocess
    pkg = ROOT / "bls_bn254_b200"
    for f in pkg.glob("*.py"):
        text = f.read_text()
        assert not re.search(r"^\s*(from|import)\s+(oracle|tests)\b", text, re.M), f
        assert "liborc" not in text and "libhostsim" not in text, f
    for f in (pkg / "csrc").glob("*.cu*"):
        for inc in re.findall(r'#include\s+"([^"]+)"', f.read_text()):
            assert "oracle" not in inc and "tests" not in inc and "orc_" not in inc, (f, inc)
    needed = subprocess.run(["readelf", "-d", str(native.SO)], capture_output=True, text=True).stdout
    assert "liborc" not in needed and "hostsim" not in needed


def test_rust_ffi_block_matches_the_header():
    """shim/src/ffi.rs (the extern "C" block of the Rust shim) is generated from include/bls_bn254_b200.h and must not drift"""
    import subprocess
    import sys
    out = subprocess.run([sys.executable, str(ROOT / "tools" / "gen_rust_ffi.py"), "--check"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    ffi = (ROOT / "shim" / "src" / "ffi.rs").read_text()
    hdr = (ROOT / "include" / "bls_bn254_b200.h").read_text()
    for name in set(re.findall(r"\b(bn254_[a-z0-9_]+)\s*\(", hdr)):
        assert "pub fn %s(" % name in ffi, name


def test_executed_op_counts_are_in_sync_with_the_kernel_sources():
    """profiles/executed_ops.json (bench.py's frac_executed) is what the host build of the device headers counts today"""
    import subprocess
    import sys
    out = subprocess.run([sys.executable, str(ROOT / "tools" / "count_ops.py"), "--check"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
