"""The C++ host side (include/bls_bn254_b200.hpp) over the C ABI: compiles as C++11 against the shared library; its test
program (tests/cpp/test_api.cpp, written like the reference's own unit tests) skips without a GPU and passes on one."""
import json
import pathlib
import subprocess

import pytest

ROOT = pathlib.Path(__file__).resolve().parent.parent


def _build(tmp):
    from bls_bn254_b200 import _native
    _native.build()
    exe = tmp / "test_api"
    libdir = _native.SO.parent
    cmd = ["g++", "-std=c++11", "-O1", "-Wall", "-Wextra", "-I", str(ROOT / "include"), str(ROOT / "tests" / "cpp" / "test_api.cpp"),
           "-o", str(exe), "-L", str(libdir), "-lbls_bn254_b200", "-Wl,-rpath," + str(libdir)]
    out = subprocess.run(cmd, capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    k = json.loads((ROOT / "tests" / "golden" / "reference_kats.json").read_text())
    lines = []
    hx = lambda s: s.encode().hex() or "-"
    for v in k["g1_hash"]:
        lines.append("g1_hash %s %s %s" % (hx(v["msg"]), hx(k["g1_hash_dst"]), v["p_x"] + v["p_y"]))
    for v in k["g1_encode"]:
        lines.append("g1_encode %s %s %s" % (hx(v["msg"]), hx(k["g1_encode_dst"]), v["p_x"] + v["p_y"]))
    for v in k["g2_hash"]:
        lines.append("g2_hash %s %s %s" % (hx(v["msg"]), hx(k["g2_hash_dst"]), v["x_c1"] + v["x_c0"] + v["y_c1"] + v["y_c0"]))
    for v in k["g2_encode"]:
        lines.append("g2_encode %s %s %s" % (hx(v["msg"]), hx(k["g2_encode_dst"]), v["x_c1"] + v["x_c0"] + v["y_c1"] + v["y_c0"]))
    lines.append("gt_generator " + "".join(k["gt_generator"]))
    katfile = tmp / "kats.txt"
    katfile.write_text("\n".join(lines) + "\n")
    return exe, katfile


def test_cpp_host_side_compiles_and_refuses_without_a_gpu(tmp_path):
    import torch
    exe, katfile = _build(tmp_path)
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: see the gpu-marked test")
    out = subprocess.run([str(exe), str(katfile)], capture_output=True, text=True, timeout=120)
    assert out.returncode == 77, out.stdout + out.stderr       # Context() threw NoDevice: no CPU path
    assert "no CUDA device" in out.stdout


@pytest.mark.gpu
def test_cpp_host_side_on_gpu(tmp_path):
    exe, katfile = _build(tmp_path)
    out = subprocess.run([str(exe), str(katfile)], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "all checks passed" in out.stdout
