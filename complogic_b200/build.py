"""In-tree build of libcomplogic_cuda.so (host C++ + embedded sm_100a cubin).

    python -m complogic_b200.build [--force]

nvcc cross-compiles the kernels for sm_100a without a GPU; the resulting cubin is embedded in the
shared library as a byte array, so the .so is the only artefact that has to travel to the GPU box.
The library has no link-time dependency on CUDA: libcuda.so.1 is dlopen'ed when a context is made.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
BUILD = os.path.join(CSRC, "_build")
SO = os.path.join(HERE, "libcomplogic_cuda.so")
CUDA_HOME = os.environ.get("CUDA_HOME", "/usr/local/cuda")

HOST_SOURCES = ["compiler.cpp", "lower.cpp", "json.cpp", "spec.cpp", "backend.cpp", "capi.cpp"]
HEADERS = ["host.hpp", "flat.hpp", "device.hpp", os.path.join("..", "..", "include", "complogic_cuda.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17"]


def _tool(name: str) -> str:
    p = os.path.join(CUDA_HOME, "bin", name)
    return p if os.path.exists(p) else (shutil.which(name) or name)


def _newer(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _gen_dispatch() -> None:
    """PTX text of the 256-entry LOP3 jump table (see kernels.cu, lut3): one file per (K, uniform?)."""
    for K in (1, 2, 4):
        for uni in (True, False):
            path = os.path.join(BUILD, f"lut_dispatch_k{K}{'u' if uni else 'd'}.inc")
            u = ".uni" if uni else ""
            lines = ['"{\\n"', '"TBL_%=: .branchtargets ' + ", ".join(f"L{n}_%=" for n in range(256)) + ';\\n"',
                     f'"brx.idx{u} %{4 * K}, TBL_%=;\\n"']
            for n in range(256):
                body = " ".join(f"lop3.b32 %{k}, %{K + k}, %{2 * K + k}, %{3 * K + k}, {n};" for k in range(K))
                lines.append(f'"L{n}_%=: {body} bra{u} DONE_%=;\\n"')
            lines += ['"DONE_%=:\\n"', '"}\\n"']
            text = "\n".join(lines) + "\n"
            if not os.path.exists(path) or open(path).read() != text:
                with open(path, "w") as f:
                    f.write(text)
    # row kernel (clg_coop_rows, K = 4): ONE warp-uniform table per pass, applied in place to the lanes whose
    # table index (%13) equals the pass (%14); operand a / the result live in %0..%3
    path = os.path.join(BUILD, "lut_dispatch_k4p.inc")
    lines = ['"{\\n"', '".reg .pred q;\\n"', '"setp.eq.u32 q, %13, %14;\\n"',
             '"TBL_%=: .branchtargets ' + ", ".join(f"L{n}_%=" for n in range(256)) + ';\\n"', '"brx.idx.uni %12, TBL_%=;\\n"']
    for n in range(256):
        body = " ".join(f"@q lop3.b32 %{k}, %{k}, %{4 + k}, %{8 + k}, {n};" for k in range(4))
        lines.append(f'"L{n}_%=: {body} bra.uni DONE_%=;\\n"')
    lines += ['"DONE_%=:\\n"', '"}\\n"']
    text = "\n".join(lines) + "\n"
    if not os.path.exists(path) or open(path).read() != text:
        with open(path, "w") as f:
            f.write(text)
    # ... and the form for rows with a single table: no predicate (ptxas turns a predicated in-place LOP3 into
    # LOP3 + SEL, twice the ALU work; lanes without an instruction compute a value nobody reads)
    path = os.path.join(BUILD, "lut_dispatch_k4r.inc")
    lines = ['"{\\n"', '"TBL_%=: .branchtargets ' + ", ".join(f"L{n}_%=" for n in range(256)) + ';\\n"', '"brx.idx.uni %12, TBL_%=;\\n"']
    for n in range(256):
        body = " ".join(f"lop3.b32 %{k}, %{k}, %{4 + k}, %{8 + k}, {n};" for k in range(4))
        lines.append(f'"L{n}_%=: {body} bra.uni DONE_%=;\\n"')
    lines += ['"DONE_%=:\\n"', '"}\\n"']
    text = "\n".join(lines) + "\n"
    if not os.path.exists(path) or open(path).read() != text:
        with open(path, "w") as f:
            f.write(text)


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(BUILD, exist_ok=True)
    _gen_dispatch()
    cu = os.path.join(CSRC, "kernels.cu")
    cubin = os.path.join(BUILD, "kernels.cubin")
    inc = os.path.join(BUILD, "kernels_cubin.inc")
    if force or _newer(cubin, [cu, os.path.join(CSRC, "flat.hpp"), os.path.abspath(__file__)]):
        cmd = [_tool("nvcc"), "-cubin", *NVCC_FLAGS, "-I" + BUILD, "-Xptxas", "-v", "-o", cubin, cu]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or r.returncode:
            sys.stderr.write(r.stdout + r.stderr)
        if r.returncode:
            raise RuntimeError("nvcc failed")
        with open(os.path.join(BUILD, "ptxas_info.txt"), "w") as f:
            f.write(r.stderr)
    if force or _newer(inc, [cubin]):
        with open(inc, "w") as f:
            subprocess.check_call([_tool("bin2c"), "--name", "clg_cubin", "--const", "--static", cubin], stdout=f)
    srcs = [os.path.join(CSRC, s) for s in HOST_SOURCES]
    deps = srcs + [os.path.join(CSRC, h) for h in HEADERS] + [inc]
    if force or _newer(SO, deps):
        objs = []
        procs = []
        for s in srcs:
            o = os.path.join(BUILD, os.path.basename(s) + ".o")
            objs.append(o)
            if force or _newer(o, [s, inc] + [os.path.join(CSRC, h) for h in HEADERS]):
                procs.append(subprocess.Popen(["g++", "-O2", "-std=c++17", "-fPIC", "-Wall", "-Wextra", "-fvisibility=hidden",
                                               "-I" + os.path.join(CUDA_HOME, "include"), "-I" + BUILD, "-c", s, "-o", o]))
        if any(p.wait() for p in procs):
            raise RuntimeError("g++ failed")
        subprocess.check_call(["g++", "-shared", "-o", SO, *objs, "-ldl", "-lpthread"])
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
