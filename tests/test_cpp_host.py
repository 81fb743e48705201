"""The C++ host mirror (ballxworld_b200/host/bxw_host.hpp) of the reference's Rust interfaces: it must compile and link against
the C ABI on any machine (CPU test), and its parity tests (tests/cpp/test_host.cpp: the reference's own RLE tests, generator,
mesher, batched region handlers and edits against the oracle) must pass on a GPU."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "tests", "cpp", "_build")
BIN = os.path.join(BUILD, "test_host")


def build_binary(name="test_host"):
    from ballxworld_b200.build import build as build_cuda

    build_cuda()
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-s"])
    os.makedirs(BUILD, exist_ok=True)
    src = os.path.join(ROOT, "tests", "cpp", name + ".cpp")
    hdr = os.path.join(ROOT, "ballxworld_b200", "host", "bxw_host.hpp")
    out = os.path.join(BUILD, name)
    if os.path.exists(out) and os.path.getmtime(out) >= max(os.path.getmtime(src), os.path.getmtime(hdr)):
        return out
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"),
           "-I" + os.path.join(ROOT, "ballxworld_b200", "host"), "-I" + os.path.join(ROOT, "oracle"), src, "-o", out,
           "-L" + os.path.join(ROOT, "ballxworld_b200", "csrc"), "-L" + os.path.join(ROOT, "oracle"), "-lbxw_cuda", "-lbxw_oracle",
           "-Wl,-rpath,$ORIGIN/../../../ballxworld_b200/csrc", "-Wl,-rpath,$ORIGIN/../../../oracle"]
    subprocess.check_call(cmd)
    return out


def test_cpp_host_mirror_compiles_and_links():
    assert os.path.exists(build_binary())


def test_cpp_host_logic_without_a_gpu():
    """Coordinate rules, registry, save blobs; and on a machine without a CUDA device the context must refuse to exist."""
    import torch

    args = [build_binary("test_host_cpu")] + ([] if torch.cuda.is_available() else ["--expect-no-gpu"])
    r = subprocess.run(args, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "test result: ok" in r.stdout, r.stdout[-3000:] + r.stderr[-2000:]


@pytest.mark.gpu
def test_cpp_host_mirror_parity():
    r = subprocess.run([build_binary()], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "test result: ok" in r.stdout, r.stdout[-3000:] + r.stderr[-2000:]
