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
    assert C.sizeof(_lib.ClgFlatInfo) == 16 + 12 * 4


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
