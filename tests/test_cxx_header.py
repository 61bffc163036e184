"""The C++ host header (include/ra/cuda.hh) against the reference's known-answer tests (tests/cxx/kat.cc).

Linked twice: (1) against oracle/libracu_oracle.so - the racu_* ABI over host memory executed by the CPU oracle - which
checks the header's flattening, slicing, agreement and error behaviour without a GPU; (2) against ra_ra_b200/libracu.so,
run with -m gpu on the device (the binary is built here by build() and travels with the snapshot)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CXX = os.path.join(ROOT, "tests", "cxx")
BUILD = os.path.join(CXX, "build")


def build_kat(which):
    os.makedirs(BUILD, exist_ok=True)
    out = os.path.join(BUILD, f"kat_{which}")
    libdir, lib = {"oracle": (os.path.join(ROOT, "oracle"), "racu_oracle"), "cuda": (os.path.join(ROOT, "ra_ra_b200"), "racu")}[which]
    deps = [os.path.join(CXX, "kat.cc"), os.path.join(CXX, "harness.hh"), os.path.join(ROOT, "include", "ra", "cuda.hh"),
            os.path.join(ROOT, "include", "racu.h"), os.path.join(libdir, f"lib{lib}.so")]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        subprocess.run(["g++", "-std=c++23", "-O1", "-Wall", "-Wextra", "-I", os.path.join(ROOT, "include"), "-o", out,
                        os.path.join(CXX, "kat.cc"), "-L", libdir, f"-l{lib}", f"-Wl,-rpath,$ORIGIN/../../../{os.path.relpath(libdir, ROOT)}"],
                       check=True)
    return out


def build_bench(which="cuda"):
    """tests/cxx/bench_header.cc: the header timed end to end (run by bench.py on the GPU box)."""
    os.makedirs(BUILD, exist_ok=True)
    out = os.path.join(BUILD, f"bench_header_{which}")
    libdir, lib = {"oracle": (os.path.join(ROOT, "oracle"), "racu_oracle"), "cuda": (os.path.join(ROOT, "ra_ra_b200"), "racu")}[which]
    deps = [os.path.join(CXX, "bench_header.cc"), os.path.join(ROOT, "include", "ra", "cuda.hh"), os.path.join(ROOT, "include", "racu.h"),
            os.path.join(libdir, f"lib{lib}.so")]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        subprocess.run(["g++", "-std=c++23", "-O2", "-Wall", "-Wextra", "-I", os.path.join(ROOT, "include"), "-o", out,
                        os.path.join(CXX, "bench_header.cc"), "-L", libdir, f"-l{lib}", f"-Wl,-rpath,$ORIGIN/../../../{os.path.relpath(libdir, ROOT)}"],
                       check=True)
    return out


def test_header_kat_on_oracle():
    """With RACU_WARM_JIT=1 also warms the run-time specialiser's disk cache (ra_ra_b200/jit_cache, which travels with the tree) for the GPU
    run of the same binary: the stand-in hands every descriptor to the product's racu_prepare (oracle/racu_shim.cc, RACU_SHIM_PREPARE)."""
    exe = build_kat("oracle")
    env = dict(os.environ)
    lib = os.path.join(ROOT, "ra_ra_b200", "libracu.so")
    if os.path.exists(lib) and os.environ.get("RACU_WARM_JIT"):
        env["RACU_SHIM_PREPARE"] = lib
    r = subprocess.run([exe], capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-1000:]
    assert " 0 failed" in r.stdout


def test_header_kat_cuda_binary_builds():
    from ra_ra_b200 import build
    build.build_lib()
    assert os.path.exists(build_kat("cuda"))


@pytest.mark.gpu
def test_header_kat_on_cuda():
    exe = os.path.join(BUILD, "kat_cuda")
    if not os.path.exists(exe):
        exe = build_kat("cuda")
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-1000:]
    assert " 0 failed" in r.stdout
