"""The C-ABI library loads and exports every symbol include/complogic_cuda.h declares; without a GPU it fails
loudly instead of falling back (no compute calls here)."""
import ctypes as C
import os
import re

import pytest

import complogic_b200 as host
from complogic_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "complogic_cuda.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(clg_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported_and_bound():
    syms = header_symbols()
    assert len(syms) >= 55
    L = _lib.lib()
    for s in syms:
        assert hasattr(L, s), f"{s} is declared in the header but not exported by the library"
        assert s in _lib.PROTOTYPES, f"{s} has no ctypes prototype"
    assert set(_lib.PROTOTYPES) <= set(syms)
    assert L.clg_abi_version() == 1


def test_struct_layouts_match_the_header():
    assert C.sizeof(_lib.ClgOp) == 32 and host.OP_DTYPE.itemsize == 32
    assert C.sizeof(_lib.ClgGate) == 8 + 13 * 8
    assert C.sizeof(_lib.ClgFlatInfo) == 16 + 18 * 4


def test_no_library_dependency_on_cuda_at_link_time():
    import subprocess
    out = subprocess.run(["readelf", "-d", _lib.SO_PATH], capture_output=True, text=True).stdout
    needed = re.findall(r"NEEDED.*\[(.*?)\]", out)
    assert not any("cuda" in n or "nvrtc" in n or "torch" in n for n in needed), needed


def test_the_embedded_kernels_are_sm100a():
    import subprocess
    cubin = os.path.join(ROOT, "complogic_b200", "csrc", "_build", "kernels.cubin")
    if not os.path.exists(cubin):
        pytest.skip("cubin not in tree (library was built elsewhere)")
    out = subprocess.run(["cuobjdump", "-elf", cubin], capture_output=True, text=True).stdout
    assert "sm_100a" in out or "EF_CUDA_SM100" in out or "sm_100" in out
    sass = subprocess.run(["cuobjdump", "-sass", "-fun", "clg_lane_k4", cubin], capture_output=True, text=True).stdout
    assert "UBLKCP" in sass, "the lane kernel must stage its instruction stream with cp.async.bulk (TMA bulk copy)"
    assert "LOP3.LUT" in sass and "LDS.128" in sass and "SYNCS" in sass


@pytest.mark.skipif(os.path.exists("/dev/nvidiactl"), reason="a GPU is present")
def test_without_a_gpu_the_product_fails_loudly():
    with pytest.raises(host.ClgError) as e:
        host.Context(0)
    assert e.value.code == _lib.CLG_E_NO_DRIVER
    assert "no CPU execution path" in str(e.value)
    c = host.Compiler(2)
    sim = c.compile([host.And(0, 1, c.alloc())])
    with pytest.raises(host.ClgError):
        sim.run([True, True])  # Simulation.run is a device operation; there is no fallback


def test_shard_ranges_tile_the_batch():
    for nv in (0, 1, 31, 32, 33, 127, 128, 129, 1000, 4097, 1 << 20, (1 << 20) + 77):
        words = (nv + 31) // 32
        for n in (1, 2, 3, 4, 8):
            prev = 0
            for s in range(n):
                w0, w1 = host.shard_range(nv, n, s)
                assert w0 == prev and w0 <= w1 <= words and (w0 % 4 == 0 or w0 == words)
                prev = w1
            assert prev == words
    with pytest.raises(host.ClgError):
        host.shard_range(100, 2, 2)


def test_spec_ptx_assembles_for_sm100a(tmp_path):
    """The generated specialised kernels are checked with ptxas on the build host (no GPU needed)."""
    import shutil
    import subprocess
    from complogic_b200 import circuits

    if not shutil.which("ptxas"):
        pytest.skip("ptxas not installed")
    for circ, k in ((circuits.ripple_adder(32), 4), (circuits.array_multiplier(16), 2), (circuits.array_multiplier(8), 1)):
        fp = host.FlatProgram.levelize(circ.sim, circ.n_immediates, circ.out_regs)
        ptx = tmp_path / f"{circ.name}.ptx"
        ptx.write_text(fp.spec_ptx(k))
        cubin = tmp_path / f"{circ.name}.cubin"
        r = subprocess.run(["ptxas", "-arch=sm_100a", "-v", "-o", str(cubin), str(ptx)], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        assert "0 bytes spill stores" in r.stderr
        sass = subprocess.run(["cuobjdump", "-sass", str(cubin)], capture_output=True, text=True).stdout
        n_lop3 = sass.count("LOP3.LUT")
        assert n_lop3 >= fp.info["n_luts"] * k  # one LOP3 per LUT per word (+ inverted stores)
        assert (".128" in sass) if k == 4 else True
    # one kernel out of the middle of a chunked chain (live registers reloaded from / stored to the scratch buffer)
    dag = circuits.random_dag(levels=1500, width=48, n_immediates=64, window=2, seed=31)
    fp = host.FlatProgram.levelize(dag.sim, 64, dag.out_regs)
    text, n = fp.spec_chunk_ptx(4, 1)
    assert n >= 3 and ".entry clg_spec_chunk" in text
    (tmp_path / "chunk.ptx").write_text(text)
    r = subprocess.run(["ptxas", "-arch=sm_100a", "-v", "-o", str(tmp_path / "c.cubin"), str(tmp_path / "chunk.ptx")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    # the fused multi-step form of a stateful program: carried registers live in registers across the step loop
    c = host.Compiler(2)
    sim = c.compile([host.DLatch(0, 1, c.alloc())])
    fp = host.FlatProgram.levelize(sim, 2, None, flags=host.LV_STATEFUL)
    for k in (1, 4):
        ptx = tmp_path / f"dlatch_steps{k}.ptx"
        ptx.write_text(fp.spec_ptx(k, steps=True))
        r = subprocess.run(["ptxas", "-arch=sm_100a", "-o", str(tmp_path / "d.cubin"), str(ptx)], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr


def test_header_is_plain_c(tmp_path):
    """The boundary must be bindable from C (and so from Rust/Go/... FFI): the header compiles as strict C99."""
    import subprocess
    src = tmp_path / "t.c"
    src.write_text('#include "complogic_cuda.h"\nint main(void) { clg_op op = {CLG_OP_NAND, 0, 0, 1, 2}; '
                   'return (int)(op.c - 2) + (int)sizeof(clg_gate) - 112 + (int)sizeof(clg_op) - 32; }\n')
    exe = tmp_path / "t"
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"),
                        str(src), "-o", str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert subprocess.run([str(exe)]).returncode == 0


def _build_demo(tmp_path):
    import subprocess
    exe = tmp_path / "adder_demo"
    libdir = os.path.join(ROOT, "complogic_b200")
    r = subprocess.run(["g++", "-std=c++17", "-Wall", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "adder_demo.cpp"),
                        "-L", libdir, "-lcomplogic_cuda", f"-Wl,-rpath,{libdir}", "-o", str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_cpp_host_links_against_the_abi(tmp_path):
    """examples/adder_demo.cpp drives the whole path from C++ through the C ABI only; without a GPU it must stop at
    clg_ctx_create with the no-driver error (after compiling and levelizing on the host)."""
    import subprocess
    exe = _build_demo(tmp_path)
    r = subprocess.run([str(exe), "8", "100"], capture_output=True, text=True)
    assert "adder8: 169 ops (152 NAND), 169 registers" in r.stdout
    assert "levelized: 16 LOP3 for 152 NAND" in r.stdout
    if not os.path.exists("/dev/nvidiactl"):
        assert r.returncode == 1 and "no CPU execution path" in r.stderr


@pytest.mark.gpu
def test_cpp_host_runs_on_the_gpu(tmp_path):
    import subprocess
    exe = _build_demo(tmp_path)
    r = subprocess.run([str(exe), "32", "200000"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "200000 vectors on the GPU" in r.stdout and ": 0 mismatches" in r.stdout


def test_product_never_touches_the_oracle():
    """oracle/ is test infrastructure: nothing under the product tree may import, link or execute it (comments may
    mention it), the shared library must not depend on it, and bench.py may only use it for the CPU baseline legs."""
    import re
    import subprocess
    pat = re.compile(r"^\s*(from|import)\s+oracle\b|liborc|clg_oracle\.c|\borc_[a-z_]+\(", re.M)
    for top in ("complogic_b200", "include", "complogic-cuda", "examples", "tools"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, top)):
            if "_build" in dirpath or "__pycache__" in dirpath:
                continue
            for name in files:
                if not name.endswith((".py", ".cpp", ".hpp", ".h", ".cu", ".rs", ".sh", ".toml")):
                    continue
                text = open(os.path.join(dirpath, name), errors="replace").read()
                code = "\n".join(l for l in text.splitlines() if not l.lstrip().startswith(("#", "//", "*", "/*")))
                assert not pat.search(code), f"{top}/{name} references the oracle"
    so = os.path.join(ROOT, "complogic_b200", "libcomplogic_cuda.so")
    needed = subprocess.run(["readelf", "-d", so], capture_output=True, text=True).stdout
    assert "liborc" not in needed and "libcuda" not in needed  # (libcuda is dlopen'ed at run time)
    bench = open(os.path.join(ROOT, "bench.py")).read()
    uses = [m.start() for m in re.finditer(r"from oracle import|import oracle", bench)]
    assert uses, "bench.py times the CPU baseline with the oracle"
    for pos in uses:  # each use sits inside a CPU-baseline / reference-arm block
        context = bench[max(0, pos - 1500):pos]
        assert "cpu" in context.lower() or "reference" in context.lower()


def test_rust_ffi_is_generated_from_the_header():
    """complogic-cuda/src/ffi.rs cannot be compiled here (no rustc), so it is GENERATED from include/complogic_cuda.h
    (tools/gen_ffi.py) and checked: the committed text is what the generator produces from the current header, it declares
    exactly the functions the header declares and the ctypes table binds, with the same number of arguments, and every
    `clg_*` function the safe wrapper (lib.rs) calls is declared."""
    import subprocess
    import sys
    assert subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gen_ffi.py"), "--check"]).returncode == 0, \
        "complogic-cuda/src/ffi.rs is stale: run python tools/gen_ffi.py"
    text = open(os.path.join(ROOT, "complogic-cuda", "src", "ffi.rs")).read()
    fns = dict(re.findall(r"pub fn (clg_\w+)\(([^)]*)\)", text))
    assert set(fns) == set(_lib.PROTOTYPES), set(fns) ^ set(_lib.PROTOTYPES)
    for name, args in fns.items():
        n_args = 0 if not args.strip() else args.count(":")
        assert n_args == len(_lib.PROTOTYPES[name][1]), name
    used = set(re.findall(r"\b(clg_[a-z_0-9]+)\(", open(os.path.join(ROOT, "complogic-cuda", "src", "lib.rs")).read()))
    assert used <= set(fns), used - set(fns)
    for struct, ct in (("clg_op", _lib.ClgOp), ("clg_gate", _lib.ClgGate), ("clg_flat_info", _lib.ClgFlatInfo)):
        body = re.search(r"pub struct %s \{(.*?)\}" % struct, text, flags=re.S).group(1)
        assert [f for f in re.findall(r"pub (\w+):", body)] == [f[0] for f in ct._fields_], struct
