"""Builds ra_ra_b200/libracu.so (hand-written sm_100a kernels + C ABI) in-tree with nvcc. No JIT, no torch extension."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libracu.so")
STATIC_PARTS = 4  # the static program table is compiled in this many parallel translation units
SOURCES = ["ply_kernels.cu"] + [f"static_kernels.cu#{k}" for k in range(STATIC_PARTS)] + [ "stencil.cu", "transpose.cu", "gemm.cu", "gemm_tc5.cu", "plan.cc", "api.cc", "special.cc", "comm.cc", "jit.cc"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-std=c++20", "-O3", "-lineinfo", "-fmad=false",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-Wall", "-Xcompiler", "-Wno-unknown-pragmas", "--expt-relaxed-constexpr"]


def _headers():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".h")]
    hs.append(os.path.join(HERE, "..", "include", "racu.h"))
    return hs


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


# Sources the run-time specialiser (jit.cc) hands to NVRTC, embedded in the library as text. The last four stand in for
# the host headers kcore.h / racu.h include: NVRTC has the fixed-width types and math built in, nothing else is used.
JIT_HEADERS = ["static_device.h", "kbody.h", "kcore.h", "static_progs.h", "../../include/racu.h"]
JIT_STUBS = {
    "stdint.h": """#pragma once
typedef signed char int8_t; typedef unsigned char uint8_t; typedef short int16_t; typedef unsigned short uint16_t;
typedef int int32_t; typedef unsigned int uint32_t; typedef long long int64_t; typedef unsigned long long uint64_t;
typedef unsigned long uintptr_t; typedef long intptr_t;
""",
    "stddef.h": "#pragma once\n",
    "math.h": "#pragma once\n",
    "type_traits": """#pragma once
namespace std {
template <class A, class B> struct is_same { static constexpr bool value = false; };
template <class A> struct is_same<A, A> { static constexpr bool value = true; };
template <class A, class B> inline constexpr bool is_same_v = is_same<A, B>::value;
template <bool C, class A, class B> struct conditional { using type = A; };
template <class A, class B> struct conditional<false, A, B> { using type = B; };
template <bool C, class A, class B> using conditional_t = typename conditional<C, A, B>::type;
template <class T, T v> struct integral_constant { static constexpr T value = v; };
}
""",
}


def _write_jit_sources(bdir):
    parts = []
    for name in JIT_HEADERS:
        parts.append((name, open(os.path.join(CSRC, name)).read()))
    parts += list(JIT_STUBS.items())
    out = []
    for name, text in parts:
        assert ')RACUSRC"' not in text
        # string literals are capped at 64 KiB by some compilers: concatenate adjacent raw literals of <= 16 KiB
        chunks = [text[i:i + 16000] for i in range(0, len(text), 16000)] or [""]
        out.append('{"%s",\n%s},\n' % (name, "\n".join('R"RACUSRC(%s)RACUSRC"' % c for c in chunks)))
    path = os.path.join(bdir, "jit_sources.inc")
    new = "".join(out)
    if not os.path.exists(path) or open(path).read() != new:
        open(path, "w").write(new)
    return path


def build_lib(force=False, verbose=False):
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    bdir = os.path.join(CSRC, "build")
    os.makedirs(bdir, exist_ok=True)
    jit_inc = _write_jit_sources(bdir)
    hdrs = _headers()
    jobs = []
    for s in SOURCES:
        name, _, part = s.partition("#")
        src = os.path.join(CSRC, name)
        obj = os.path.join(bdir, s.replace("#", "_") + ".o")
        if force or _stale(obj, [src] + hdrs + ([jit_inc] if name == "jit.cc" else [])):
            if name.endswith(".cc"):  # host-only translation units
                cmd = [os.environ.get("CXX", "g++"), "-O2", "-std=c++20", "-fPIC", "-Wall", "-Wno-unknown-pragmas",
                       "-I", os.path.join(os.path.dirname(os.path.dirname(nvcc)), "include"), "-c", src, "-o", obj]
            else:
                defs = [f"-DSP_PART={part}", f"-DSP_NPART={STATIC_PARTS}"] if part else []
                cmd = [nvcc] + NVCC_FLAGS + defs + ["-c", src, "-o", obj]
            jobs.append(cmd)
    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed: " + " ".join(cmd) + "\n" + r.stdout + r.stderr)
        if verbose and (r.stdout or r.stderr):
            sys.stderr.write(r.stdout + r.stderr)
    with ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        list(ex.map(run, jobs))
    objs = [os.path.join(bdir, s.replace("#", "_") + ".o") for s in SOURCES]
    if force or jobs or _stale(OUT, objs):
        run([nvcc, "-shared", "-o", OUT] + objs + ["-ldl", "-gencode", "arch=compute_100a,code=sm_100a"])
    return OUT


if __name__ == "__main__":
    print(build_lib(force="--force" in sys.argv, verbose=True))
