"""Host-side mirror of the reference's public API for the NAND-VM path, over the C ABI.

Same names, argument meaning and error behaviour as complogic/src/{compile,gates,simulation}.rs so
that callers (and the reference's own tests, restated in tests/) switch over unchanged:

    compiler = Compiler(2); out = compiler.alloc()
    sim = compiler.compile([And(0, 1, out)])      # Compiler::compile, compile.rs:93
    sim.run([True, True]); sim.registers[out]      # Simulation::run, simulation.rs:16 -- ON THE GPU

and the batched entry point the backend adds:

    cs = CudaSimulation.from_simulation(sim, n_immediates=2, out_regs=[out])
    out = cs.run_batch(imm_sliced, n_vectors)      # bit-sliced [rows][stride] uint32

Everything that executes a program goes through libcomplogic_cuda.so onto a CUDA device; there is no
CPU execution path here (the CPU oracle lives in oracle/ and is test infrastructure only).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, fields
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from ._lib import (ClgError, ClgFlatInfo, ClgGate, ClgOp, ReferencePanic, check, lib, MODE_AUTO, MODE_COOP, MODE_LANE,
                   MODE_SPEC, MODE_NAMES, LV_STATEFUL, LV_NO_FUSE, LV_NO_LANE_STREAM, LV_NO_PARTS, LV_NO_ROWS)

OP_DTYPE = np.dtype([("tag", "<u4"), ("reserved", "<u4"), ("a", "<u8"), ("b", "<u8"), ("c", "<u8")])
assert OP_DTYPE.itemsize == C.sizeof(ClgOp) == 32

# ---- Op (compile.rs:7-16) ----------------------------------------------------------------------------
Op = Tuple  # ("Nand", a, b, out) | ("Set", id, val)


def Nand_op(a: int, b: int, out: int) -> Op:
    return ("Nand", a, b, out)


def Set_op(id: int, val: bool) -> Op:
    return ("Set", id, bool(val))


def ops_to_array(ops) -> np.ndarray:
    if isinstance(ops, np.ndarray):
        assert ops.dtype == OP_DTYPE
        return ops
    arr = np.zeros(len(ops), dtype=OP_DTYPE)
    for i, op in enumerate(ops):
        arr[i] = (0, 0, op[1], op[2], op[3]) if op[0] == "Nand" else (1, 0, op[1], 1 if op[2] else 0, 0)
    return arr


def array_to_ops(arr: np.ndarray) -> List[Op]:
    return [("Nand", a, b, c) if tag == 0 else ("Set", a, bool(b)) for tag, _, a, b, c in arr.tolist()]


def _copy_ops(ptr, n) -> np.ndarray:
    if not n:
        return np.zeros(0, dtype=OP_DTYPE)
    return np.frombuffer(C.string_at(ptr, n * 32), dtype=OP_DTYPE).copy()


# ---- Incrementer (compile.rs:18-49) ---------------------------------------------------------------------
class Incrementer:
    def __init__(self, val: int = 0):
        self.val = val

    @staticmethod
    def set(val: int) -> "Incrementer":
        return Incrementer(val)

    def next(self) -> int:
        cur = self.val
        self.val += 1
        return cur

    def skip(self, count: int) -> None:
        self.val += count


# ---- gates (gates.rs:5-190) ---------------------------------------------------------------------------------
@dataclass(frozen=True)
class _Gate:
    KIND = -1

    def to_c(self) -> ClgGate:
        g = ClgGate()
        g.kind = self.KIND
        for i, f in enumerate(fields(self)):
            g.f[i] = getattr(self, f.name)
        return g

    def to_json(self) -> str:
        """serde_json's text for `Gate` (gates.rs:103-117), e.g. {"And":{"a":0,"b":1,"out":2}}"""
        g, out = self.to_c(), C.c_void_p()
        check(lib().clg_gate_to_json(C.byref(g), C.byref(out)))
        try:
            return C.cast(out, C.c_char_p).value.decode()
        finally:
            lib().clg_free(out)

    @staticmethod
    def from_json(text: str) -> "_Gate":
        g = ClgGate()
        check(lib().clg_gate_from_json(text.encode(), C.byref(g)))
        cls = _GATE_KINDS[g.kind]
        return cls(*[int(g.f[i]) for i in range(len(fields(cls)))])

    def create(self, incrementer: Incrementer) -> List[Op]:
        """Gate::create(&self, &mut Incrementer) -> Ops (gates.rs:193)."""
        g = self.to_c()
        inc, ptr, n = C.c_uint64(incrementer.val), C.c_void_p(), C.c_size_t()
        check(lib().clg_gate_create(C.byref(g), C.byref(inc), C.byref(ptr), C.byref(n)))
        incrementer.val = inc.value
        ops = array_to_ops(_copy_ops(ptr, n.value))
        lib().clg_free(ptr)
        return ops


def _gate(name, kind, *names):
    cls = type(name, (_Gate,), {"__annotations__": {n: int for n in names}, "KIND": kind})
    return dataclass(frozen=True)(cls)


Nand = _gate("Nand", 0, "a", "b", "out")
Not = _gate("Not", 1, "a", "out")
And = _gate("And", 2, "a", "b", "out")
Or = _gate("Or", 3, "a", "b", "out")
Nor = _gate("Nor", 4, "a", "b", "out")
Xor = _gate("Xor", 5, "a", "b", "out")
RSLatch = _gate("RSLatch", 6, "s", "r", "q")
RSLatchTest = _gate("RSLatchTest", 7, "s", "r", "q")
DLatch = _gate("DLatch", 8, "d", "e", "q")
HalfAdder = _gate("HalfAdder", 9, "a", "b", "s", "c")
FullAdder = _gate("FullAdder", 10, "a", "b", "cin", "s", "cout")
FourBitAdder = _gate("FourBitAdder", 11, "a1", "a2", "a3", "a4", "b1", "b2", "b3", "b4", "s1", "s2", "s3", "s4", "cout")
Gate = _Gate
_GATE_KINDS = {c.KIND: c for c in (Nand, Not, And, Or, Nor, Xor, RSLatch, RSLatchTest, DLatch, HalfAdder, FullAdder, FourBitAdder)}


# ---- device context ------------------------------------------------------------------------------------------
class Context:
    """clg_ctx: one CUDA device (its primary context), one host thread at a time."""

    _default = {}

    def __init__(self, device: int = 0):
        self.device = device
        h = C.c_void_p()
        check(lib().clg_ctx_create(device, C.byref(h)))
        self._h = h

    @classmethod
    def default(cls, device: Optional[int] = None) -> "Context":
        if device is None:
            import os
            device = int(os.environ.get("LOCAL_RANK", "0"))
        if device not in cls._default:
            cls._default[device] = Context(device)
        return cls._default[device]

    @property
    def name(self) -> str:
        buf = C.create_string_buffer(128)
        check(lib().clg_ctx_device_name(self._h, buf, 128))
        return buf.value.decode()

    @property
    def sm_count(self) -> int:
        return lib().clg_ctx_sm_count(self._h)

    def synchronize(self) -> None:
        check(lib().clg_ctx_synchronize(self._h))

    def measure_lop3_peak(self) -> float:
        """32-bit LOP3 lane-ops per second of a pure-LOP3 kernel on this device (measured roofline denominator)."""
        v = C.c_double()
        check(lib().clg_measure_lop3_peak(self._h, C.byref(v)))
        return v.value

    def fill_random(self, dptr: int, n_rows: int, stride: int, seed: int, stream: int = 0) -> None:
        check(lib().clg_fill_random(self._h, dptr, n_rows, stride, seed, stream))

    def slice_bools(self, bools_dptr, sliced_dptr, n_vectors, n_rows, stride, stream=0):
        check(lib().clg_slice_bools(self._h, bools_dptr, sliced_dptr, n_vectors, n_rows, stride, stream))

    def unslice_bools(self, sliced_dptr, bools_dptr, n_vectors, n_rows, stride, stream=0):
        check(lib().clg_unslice_bools(self._h, sliced_dptr, bools_dptr, n_vectors, n_rows, stride, stream))

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                lib().clg_ctx_destroy(self._h)
                self._h = None
        except Exception:
            pass


# ---- Simulation (simulation.rs:5-36) ---------------------------------------------------------------------------
class Simulation:
    """`Simulation { pub ops, pub registers }`. run() executes on the GPU (1 lane, carried registers)."""

    def __init__(self, ops: Sequence[Op] = (), registers: Sequence[bool] = (), _ops_array: Optional[np.ndarray] = None):
        self._ops_arr = _ops_array if _ops_array is not None else ops_to_array(list(ops))
        self._ops_list: Optional[List[Op]] = None if _ops_array is not None else list(ops)
        self.registers: List[bool] = [bool(r) for r in registers]
        self._runner = None  # (n_immediates, n_registers, handles...)
        self.mode = MODE_AUTO

    @property
    def ops(self) -> List[Op]:
        if self._ops_list is None:
            self._ops_list = array_to_ops(self._ops_arr)
        return self._ops_list

    @property
    def ops_array(self) -> np.ndarray:
        return self._ops_arr

    @property
    def nand_count(self) -> int:
        return int((self._ops_arr["tag"] == 0).sum())

    def _c_sim(self):
        h = C.c_void_p()
        regs = np.array(self.registers, dtype=np.uint8)
        check(lib().clg_simulation_new(self._ops_arr.ctypes.data, len(self._ops_arr), regs.ctypes.data, len(regs), C.byref(h)))
        return h

    def run(self, immediates: Sequence[bool], ctx: Optional[Context] = None) -> None:
        """Simulation::run(&mut self, immediates: &[bool]) -- simulation.rs:16-30, evaluated on the device."""
        imm = np.array([1 if x else 0 for x in immediates], dtype=np.uint8)
        key = (len(imm), len(self.registers), self.mode)
        if self._runner is None or self._runner[0] != key:
            self._drop_runner()
            ctx = ctx or Context.default()
            sim_h = self._c_sim()
            r = C.c_void_p()
            try:
                check(lib().clg_sim_runner_create(ctx._h, sim_h, len(imm), self.mode, C.byref(r)))
            except Exception:
                lib().clg_simulation_destroy(sim_h)
                raise
            self._runner = (key, ctx, sim_h, r)
        _, _, sim_h, r = self._runner
        n = C.c_size_t()
        regs_ptr = lib().clg_simulation_registers(sim_h, C.byref(n))
        view = (C.c_uint8 * n.value).from_address(regs_ptr) if n.value else []
        for i, v in enumerate(self.registers):  # registers are public and writable between runs
            view[i] = 1 if v else 0
        check(lib().clg_sim_runner_run(r, imm.ctypes.data, len(imm)))
        self.registers = [bool(x) for x in view]

    def register(self, id: int) -> bool:
        """Simulation::register (simulation.rs:33-35)."""
        if id >= len(self.registers):
            raise ReferencePanic(_lib.CLG_E_BAD_REGISTER, "register index out of range (simulation.rs:34)")
        return self.registers[id]

    def to_json(self) -> str:
        h, out = self._c_sim(), C.c_void_p()
        try:
            check(lib().clg_simulation_to_json(h, C.byref(out)))
            s = C.string_at(out).decode()
            lib().clg_free(out)
            return s
        finally:
            lib().clg_simulation_destroy(h)

    @staticmethod
    def from_json(text: str) -> "Simulation":
        h = C.c_void_p()
        check(lib().clg_simulation_from_json(text.encode(), C.byref(h)))
        try:
            return _sim_from_handle(h)
        finally:
            lib().clg_simulation_destroy(h)

    def _drop_runner(self):
        if self._runner is not None:
            _, _, sim_h, r = self._runner
            lib().clg_sim_runner_destroy(r)
            lib().clg_simulation_destroy(sim_h)
            self._runner = None

    def __del__(self):
        try:
            self._drop_runner()
        except Exception:
            pass


def _sim_from_handle(h) -> Simulation:
    n = C.c_size_t()
    p = lib().clg_simulation_ops(h, C.byref(n))
    ops = _copy_ops(p, n.value)
    m = C.c_size_t()
    rp = lib().clg_simulation_registers(h, C.byref(m))
    regs = [bool(b) for b in C.string_at(rp, m.value)] if m.value else []
    return Simulation(registers=regs, _ops_array=ops)


# ---- Compiler (compile.rs:51-203) ---------------------------------------------------------------------------------
class Compiler:
    def __init__(self, immediate_count: int):
        h = C.c_void_p()
        check(lib().clg_compiler_new(immediate_count, C.byref(h)))
        self._h = h
        self.incrementer = _IncrementerView(self)

    # the three public fields of the reference struct
    @property
    def immediate_count(self) -> int:
        return lib().clg_compiler_immediate_count(self._h)

    @immediate_count.setter
    def immediate_count(self, v: int) -> None:
        lib().clg_compiler_set_immediate_count(self._h, v)

    @property
    def ops(self) -> List[Op]:
        n = C.c_size_t()
        p = lib().clg_compiler_ops(self._h, C.byref(n))
        return array_to_ops(_copy_ops(p, n.value))

    @property
    def last_dot(self) -> str:
        """Text the reference writes to ./graph.dot on every compile (compile.rs:150); kept in memory here."""
        return lib().clg_compiler_graph_dot(self._h).decode()

    def reset_ops(self) -> None:
        lib().clg_compiler_reset_ops(self._h)

    def reset_incrementer(self) -> None:
        lib().clg_compiler_reset_incrementer(self._h)

    def alloc(self) -> int:
        return lib().clg_compiler_alloc(self._h)

    def to_json(self) -> str:
        """serde_json's text for the derives at compile.rs:18,51: {"immediate_count":..,"ops":[..],"incrementer":{"val":..}}"""
        out = C.c_void_p()
        check(lib().clg_compiler_to_json(self._h, C.byref(out)))
        try:
            return C.cast(out, C.c_char_p).value.decode()
        finally:
            lib().clg_free(out)

    @staticmethod
    def from_json(text: str) -> "Compiler":
        h = C.c_void_p()
        check(lib().clg_compiler_from_json(text.encode(), C.byref(h)))
        c = Compiler.__new__(Compiler)
        c._h = h
        c.incrementer = _IncrementerView(c)
        return c

    def compile(self, gates: Sequence[_Gate]) -> Simulation:
        arr = (ClgGate * max(1, len(gates)))(*[g.to_c() for g in gates])
        out = C.c_void_p()
        check(lib().clg_compiler_compile(self._h, arr, len(gates), C.byref(out)))
        try:
            return _sim_from_handle(out)
        finally:
            lib().clg_simulation_destroy(out)

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                lib().clg_compiler_destroy(self._h)
                self._h = None
        except Exception:
            pass


class _IncrementerView:
    """`compiler.incrementer` as a live view of the C-side field (val / next / skip)."""

    def __init__(self, c: Compiler):
        self._c = c

    @property
    def val(self) -> int:
        return lib().clg_compiler_incrementer(self._c._h)

    @val.setter
    def val(self, v: int) -> None:
        lib().clg_compiler_set_incrementer(self._c._h, v)

    def next(self) -> int:
        return self._c.alloc()

    def skip(self, count: int) -> None:
        self.val = self.val + count


# ---- the new pass + the device path ------------------------------------------------------------------------------------
class FlatProgram:
    """clg_flat: the flat, level-ordered, register-allocated device buffer of one program."""

    def __init__(self, handle):
        self._h = handle
        info = ClgFlatInfo()
        check(lib().clg_flat_get_info(self._h, C.byref(info)))
        self.info = {f[0]: getattr(info, f[0]) for f in ClgFlatInfo._fields_}

    @staticmethod
    def levelize(sim: Simulation, n_immediates: int, out_regs: Optional[Sequence[int]] = None, flags: int = 0) -> "FlatProgram":
        sh = sim._c_sim()
        try:
            h = C.c_void_p()
            if out_regs is None:
                check(lib().clg_levelize(sh, n_immediates, None, 0, flags, C.byref(h)))
            else:
                regs = np.ascontiguousarray(out_regs, dtype=np.uint64)
                if len(regs) == 0:
                    raise ValueError("out_regs must name at least one register (None = all)")
                check(lib().clg_levelize(sh, n_immediates, regs.ctypes.data, len(regs), flags, C.byref(h)))
            return FlatProgram(h)
        finally:
            lib().clg_simulation_destroy(sh)

    def to_bytes(self) -> bytes:
        p, n = C.c_void_p(), C.c_size_t()
        check(lib().clg_flat_bytes(self._h, C.byref(p), C.byref(n)))
        return C.string_at(p, n.value)

    @staticmethod
    def from_bytes(data: bytes) -> "FlatProgram":
        h = C.c_void_p()
        check(lib().clg_flat_from_bytes(data, len(data), C.byref(h)))
        return FlatProgram(h)

    def state_regs(self) -> np.ndarray:
        p, n = C.c_void_p(), C.c_size_t()
        check(lib().clg_flat_state_regs(self._h, C.byref(p), C.byref(n)))
        return np.frombuffer(C.string_at(p, n.value * 8), dtype=np.uint64).copy() if n.value else np.zeros(0, np.uint64)

    def spec_ptx(self, k: int = 4, steps: bool = False) -> str:
        out = C.c_void_p()
        check((lib().clg_flat_spec_steps_ptx if steps else lib().clg_flat_spec_ptx)(self._h, k, C.byref(out)))
        s = C.string_at(out).decode()
        lib().clg_free(out)
        return s

    def spec_chain_info(self, k: int = 1) -> dict:
        """Shape of the specialised chain at k words per thread: chunks, distinct kernels to assemble, launches per run."""
        a, b, c = C.c_uint32(), C.c_uint32(), C.c_uint32()
        check(lib().clg_flat_spec_chain_info(self._h, k, C.byref(a), C.byref(b), C.byref(c)))
        return {"chunks": a.value, "kernels": b.value, "launches": c.value}

    def spec_batch_chain_info(self, k: int = 1) -> dict:
        """Shape of the second specialised chain (the one big batches run): chunks, kernels, launches, registers handed over."""
        a, b, c, d = C.c_uint32(), C.c_uint32(), C.c_uint32(), C.c_uint32()
        check(lib().clg_flat_spec_batch_chain_info(self._h, k, C.byref(a), C.byref(b), C.byref(c), C.byref(d)))
        return {"chunks": a.value, "kernels": b.value, "launches": c.value, "handed_over": d.value}

    def spec_chunk_ptx(self, k: int, chunk: int, with_bases: bool = False, second: bool = False):
        """PTX of one kernel of the chunked specialised form and the length of the chain (and, with_bases, the rows
        (imm, out, state) the kernel's pointers are advanced by). second: of the chain big batches run."""
        out, n, bases = C.c_void_p(), C.c_uint32(), (C.c_uint32 * 3)()
        fn = lib().clg_flat_spec_batch_chunk_ptx if second else lib().clg_flat_spec_chunk_ptx
        check(fn(self._h, k, chunk, C.byref(n), bases, C.byref(out)))
        s = C.string_at(out).decode()
        lib().clg_free(out)
        return (s, n.value, tuple(bases)) if with_bases else (s, n.value)

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                lib().clg_flat_destroy(self._h)
                self._h = None
        except Exception:
            pass


def round_stride(n_vectors: int) -> int:
    """words per row for n_vectors lanes, rounded up to 4 (16-byte rows)."""
    return ((n_vectors + 31) // 32 + 3) // 4 * 4


class CudaSimulation:
    """A program resident on one device: the batched `Simulation::run`."""

    def __init__(self, flat: FlatProgram, ctx: Optional[Context] = None, mode: int = MODE_AUTO):
        self.flat = flat
        self.ctx = ctx or Context.default()
        h = C.c_void_p()
        check(lib().clg_exec_create(self.ctx._h, flat._h, mode, C.byref(h)))
        self._h = h
        self.info = flat.info

    @staticmethod
    def from_simulation(sim: Simulation, n_immediates: int, out_regs: Optional[Sequence[int]] = None, flags: int = 0,
                        ctx: Optional[Context] = None, mode: int = MODE_AUTO) -> "CudaSimulation":
        return CudaSimulation(FlatProgram.levelize(sim, n_immediates, out_regs, flags), ctx, mode)

    @property
    def mode(self) -> int:
        return lib().clg_exec_mode(self._h)

    @property
    def mode_name(self) -> str:
        return MODE_NAMES[self.mode]

    @property
    def words_per_thread(self) -> int:
        return lib().clg_exec_words_per_thread(self._h)

    @property
    def launches_per_run(self) -> int:
        return lib().clg_exec_launches_per_run(self._h)

    @property
    def kernel_name(self) -> str:
        """Name of the kernel the most recent run launched ("" before any run)."""
        return lib().clg_exec_kernel_name(self._h).decode()

    @property
    def ptx(self) -> str:
        return lib().clg_exec_ptx(self._h).decode()

    def prepare(self, n_vectors: int) -> None:
        """clg_exec_prepare: choose the mode for this batch size, build its plan and tune it now instead of in the first run."""
        check(lib().clg_exec_prepare(self._h, n_vectors))

    def run_device(self, imm_dptr: int, out_dptr: int, n_vectors: int, stride: int, state_dptr: int = 0, stream: int = 0) -> None:
        """clg_exec_run: device pointers, asynchronous on `stream`."""
        check(lib().clg_exec_run(self._h, imm_dptr, out_dptr, state_dptr or None, n_vectors, stride, stream or None))

    def run_steps_device(self, imm_dptr: int, out_dptr: int, state_dptr: int, n_vectors: int, stride: int, n_steps: int,
                         stream: int = 0) -> None:
        """clg_exec_run_steps: n_steps run() calls per lane, imm [steps][n_in][stride], out [steps][n_out][stride]."""
        check(lib().clg_exec_run_steps(self._h, imm_dptr or None, out_dptr or None, state_dptr or None, n_vectors, stride,
                                       n_steps, stream or None))

    def run_batch(self, imm: np.ndarray, n_vectors: int, state: Optional[np.ndarray] = None) -> np.ndarray:
        """Host arrays in, host array out (H2D + kernel + D2H inside): imm [n_inputs][stride] uint32."""
        imm = np.ascontiguousarray(imm, dtype=np.uint32)
        n_in, n_out, n_state = self.info["n_inputs"], self.info["n_outputs"], self.info["n_state"]
        if imm.ndim != 2 or imm.shape[0] != n_in:
            raise ValueError(f"imm must be [{n_in}][stride]")
        stride = imm.shape[1] if n_in else round_stride(n_vectors)
        out = np.zeros((n_out, stride), dtype=np.uint32)
        sp = None
        if n_state:
            if state is None or state.shape != (n_state, stride) or state.dtype != np.uint32 or not state.flags.c_contiguous:
                raise ValueError(f"state must be a C-contiguous uint32 [{n_state}][{stride}] array (updated in place)")
            sp = state.ctypes.data
        check(lib().clg_exec_run_host(self._h, imm.ctypes.data if n_in else None, out.ctypes.data, sp, n_vectors, stride))
        return out

    def run_bools(self, imm: np.ndarray) -> np.ndarray:
        """`&[bool]`-shaped batch: imm [n_vectors][n_inputs] of 0/1 bytes -> [n_vectors][n_outputs]."""
        imm = np.ascontiguousarray(imm, dtype=np.uint8)
        n = imm.shape[0]
        out = np.zeros((n, self.info["n_outputs"]), dtype=np.uint8)
        check(lib().clg_exec_run_bools_host(self._h, imm.ctypes.data, out.ctypes.data, n))
        return out

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                lib().clg_exec_destroy(self._h)
                self._h = None
        except Exception:
            pass


def shard_range(n_vectors: int, n_shards: int, shard: int):
    """clg_shard_range: the word range [w0, w1) of every row that `shard` of `n_shards` owns."""
    w0, w1 = C.c_size_t(), C.c_size_t()
    check(lib().clg_shard_range(n_vectors, n_shards, shard, C.byref(w0), C.byref(w1)))
    return w0.value, w1.value


def run_sharded(flat: FlatProgram, imm: np.ndarray, n_vectors: int, devices: Sequence[int], mode: int = MODE_AUTO):
    """clg_run_sharded_host: one process, one host thread + context per device, no collective."""
    imm = np.ascontiguousarray(imm, dtype=np.uint32)
    stride = imm.shape[1]
    out = np.zeros((flat.info["n_outputs"], stride), dtype=np.uint32)
    dev = (C.c_int * len(devices))(*devices)
    secs = (C.c_double * len(devices))()
    check(lib().clg_run_sharded_host(flat._h, mode, dev, len(devices), imm.ctypes.data, out.ctypes.data, n_vectors, stride, secs))
    return out, list(secs)


class ShardedSimulation:
    """clg_sharded_*: the persistent single-process multi-GPU executor. Contexts, executors and staging buffers are made
    once (one worker thread per device); every run starts the devices together. Stateful programs are allowed."""

    def __init__(self, flat: FlatProgram, devices: Sequence[int], mode: int = MODE_AUTO):
        self.flat, self.devices = flat, list(devices)
        dev = (C.c_int * len(self.devices))(*self.devices)
        h = C.c_void_p()
        check(lib().clg_sharded_create(flat._h, mode, dev, len(self.devices), C.byref(h)))
        self._h = h
        self.last_seconds: List[float] = []

    def close(self):
        if self._h:
            lib().clg_sharded_destroy(self._h)
            self._h = None

    def __del__(self):
        self.close()

    def run(self, imm: np.ndarray, n_vectors: int, state: Optional[np.ndarray] = None, steps: int = 1) -> np.ndarray:
        """imm [n_inputs][stride] (steps == 1) or [steps][n_inputs][stride]; state [n_state][stride] is updated in place.
        Returns out [n_outputs][stride] or [steps][n_outputs][stride]."""
        imm = np.ascontiguousarray(imm, dtype=np.uint32)
        stride = imm.shape[-1]
        n_out = self.flat.info["n_outputs"]
        out = np.zeros((steps, n_out, stride) if imm.ndim == 3 else (n_out, stride), dtype=np.uint32)
        secs = (C.c_double * len(self.devices))()
        sp = state.ctypes.data if state is not None else None
        if imm.ndim == 3:
            assert imm.shape[0] == steps
            check(lib().clg_sharded_run_steps_host(self._h, imm.ctypes.data, out.ctypes.data, sp, n_vectors, stride, steps, secs))
        else:
            check(lib().clg_sharded_run_host(self._h, imm.ctypes.data, out.ctypes.data, sp, n_vectors, stride, secs))
        self.last_seconds = list(secs)
        return out
