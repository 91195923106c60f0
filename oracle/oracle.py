"""ctypes front end of oracle/clg_oracle.c -- TEST INFRASTRUCTURE ONLY.

Mirrors the reference's Rust API closely enough that the golden tests read like the reference's
own unit tests (complogic/src/{simulation,compile,gates}.rs `mod tests`):

    compiler = Compiler(2); out = compiler.alloc()
    sim = compiler.compile([And(0, 1, out)]); sim.run([True, True]); sim.registers[out]

Parity status: pinned by the reference's 14 known-answer tests and graph.dot (the Rust reference
cannot be executed in this image: no cargo/rustc) -- tests/test_oracle_golden.py.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, fields
from typing import List, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "clg_oracle.c")
_SO = os.path.join(_HERE, "_build", "liborc.so")


def build(force: bool = False) -> str:
    """gcc the C restatement into oracle/_build/liborc.so (git-ignored, travels with gpurun)."""
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(_SRC):
        os.makedirs(os.path.dirname(_SO), exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-Wall", "-shared", "-fPIC", "-pthread", "-o", _SO, _SRC])
    return _SO


class _OrcOp(C.Structure):
    _fields_ = [("tag", C.c_uint64), ("a", C.c_uint64), ("b", C.c_uint64), ("c", C.c_uint64)]


OP_DTYPE = np.dtype([("tag", "<u8"), ("a", "<u8"), ("b", "<u8"), ("c", "<u8")])
assert OP_DTYPE.itemsize == 32 == C.sizeof(_OrcOp)

_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        u64p, u32p, u8p, vp = (C.POINTER(C.c_uint64), C.POINTER(C.c_uint32), C.POINTER(C.c_uint8), C.c_void_p)
        L.orc_run.argtypes = [vp, C.c_size_t, vp, C.c_size_t, vp, C.c_size_t]
        L.orc_run.restype = C.c_int
        L.orc_gate_create.argtypes = [C.c_int, u64p, u64p, C.POINTER(vp)]
        L.orc_gate_create.restype = C.c_size_t
        L.orc_free.argtypes = [vp]
        L.orc_compile.argtypes = [C.c_uint64, C.c_uint64, C.POINTER(C.c_int), u64p, C.c_size_t,
                                  C.POINTER(vp), C.POINTER(C.c_size_t), u64p, C.POINTER(C.c_char_p)]
        L.orc_compile.restype = C.c_int
        batch = [vp, C.c_size_t, C.c_size_t, vp, C.c_size_t, C.c_size_t, C.c_size_t, C.c_size_t,
                 vp, C.c_size_t, vp, C.c_int]
        L.orc_run_batch_scalar.argtypes = batch + [C.c_int]
        L.orc_run_batch_scalar.restype = C.c_int
        L.orc_run_batch_sliced.argtypes = batch
        L.orc_run_batch_sliced.restype = C.c_int
        L.orc_now.restype = C.c_double
        _lib = L
    return _lib


# ---- Op (compile.rs:7-16) -------------------------------------------------------------------
Op = Tuple  # ("Nand", a, b, out) | ("Set", id, val)


def Nand_op(a: int, b: int, out: int) -> Op:
    return ("Nand", a, b, out)


def Set_op(id: int, val: bool) -> Op:
    return ("Set", id, bool(val))


def ops_to_array(ops: Sequence[Op]) -> np.ndarray:
    arr = np.zeros(len(ops), dtype=OP_DTYPE)
    for i, op in enumerate(ops):
        if op[0] == "Nand":
            arr[i] = (0, op[1], op[2], op[3])
        else:
            arr[i] = (1, op[1], 1 if op[2] else 0, 0)
    return arr


def array_to_ops(arr: np.ndarray) -> List[Op]:
    out = []
    for tag, a, b, c in arr.tolist():
        out.append(("Nand", a, b, c) if tag == 0 else ("Set", a, bool(b)))
    return out


def _take_ops(ptr, n) -> np.ndarray:
    if not n:
        return np.zeros(0, dtype=OP_DTYPE)
    buf = C.string_at(ptr, n * 32)
    lib().orc_free(ptr)
    return np.frombuffer(buf, dtype=OP_DTYPE).copy()


# ---- gates (gates.rs:5-117) -----------------------------------------------------------------
@dataclass(frozen=True)
class _Gate:
    KIND = -1

    def fields13(self) -> List[int]:
        v = [getattr(self, f.name) for f in fields(self)]
        return v + [0] * (13 - len(v))

    def create(self, incrementer: "Incrementer") -> List[Op]:
        """Gate::create (gates.rs:193)."""
        f = (C.c_uint64 * 13)(*self.fields13())
        inc = C.c_uint64(incrementer.val)
        ptr = C.c_void_p()
        n = lib().orc_gate_create(self.KIND, f, C.byref(inc), C.byref(ptr))
        incrementer.val = inc.value
        return array_to_ops(_take_ops(ptr, n))


def _gate(kind, *names):
    cls = dataclass(frozen=True)(type("G", (_Gate,), {"__annotations__": {n: int for n in names}, "KIND": kind}))
    return cls


Nand = _gate(0, "a", "b", "out"); Nand.__name__ = "Nand"
Not = _gate(1, "a", "out"); Not.__name__ = "Not"
And = _gate(2, "a", "b", "out"); And.__name__ = "And"
Or = _gate(3, "a", "b", "out"); Or.__name__ = "Or"
Nor = _gate(4, "a", "b", "out"); Nor.__name__ = "Nor"
Xor = _gate(5, "a", "b", "out"); Xor.__name__ = "Xor"
RSLatch = _gate(6, "s", "r", "q"); RSLatch.__name__ = "RSLatch"
RSLatchTest = _gate(7, "s", "r", "q"); RSLatchTest.__name__ = "RSLatchTest"
DLatch = _gate(8, "d", "e", "q"); DLatch.__name__ = "DLatch"
HalfAdder = _gate(9, "a", "b", "s", "c"); HalfAdder.__name__ = "HalfAdder"
FullAdder = _gate(10, "a", "b", "cin", "s", "cout"); FullAdder.__name__ = "FullAdder"
FourBitAdder = _gate(11, "a1", "a2", "a3", "a4", "b1", "b2", "b3", "b4", "s1", "s2", "s3", "s4", "cout")
FourBitAdder.__name__ = "FourBitAdder"


# ---- Incrementer / Compiler (compile.rs:18-203) ------------------------------------------------
class Incrementer:
    def __init__(self, val: int = 0):
        self.val = val

    def next(self) -> int:
        cur = self.val
        self.val += 1
        return cur

    def skip(self, count: int) -> None:
        self.val += count


class ReferencePanic(Exception):
    """Raised where the Rust reference would panic (index out of bounds)."""


class Simulation:
    """simulation.rs:5-36 -- public `ops` and `registers`."""

    def __init__(self, ops: Sequence[Op] = (), registers: Sequence[bool] = (), _ops_array=None):
        self._ops_arr = _ops_array if _ops_array is not None else ops_to_array(list(ops))
        self._ops_list = None if _ops_array is not None else list(ops)
        self.registers = [bool(r) for r in registers]

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

    def run(self, immediates: Sequence[bool]) -> None:
        ops = self._ops_arr
        regs = np.array(self.registers, dtype=np.uint8)
        imm = np.array([1 if x else 0 for x in immediates], dtype=np.uint8)
        rc = lib().orc_run(ops.ctypes.data, len(ops), regs.ctypes.data, len(regs), imm.ctypes.data, len(imm))
        if rc != 0:
            raise ReferencePanic("index out of bounds (simulation.rs:20-26)")
        self.registers = [bool(x) for x in regs]

    def register(self, id: int) -> bool:
        if id >= len(self.registers):
            raise ReferencePanic("index out of bounds (simulation.rs:34)")
        return self.registers[id]


class Compiler:
    def __init__(self, immediate_count: int):
        self.immediate_count = immediate_count
        self.ops: List[Op] = []
        self.incrementer = Incrementer(immediate_count)
        self.last_dot: str | None = None

    def reset_ops(self) -> None:
        self.ops = [Set_op(i, False) for i in range(self.immediate_count)]

    def reset_incrementer(self) -> None:
        self.incrementer = Incrementer(self.immediate_count)

    def alloc(self) -> int:
        return self.incrementer.next()

    def compile(self, gates: Sequence[_Gate]) -> Simulation:
        kinds = (C.c_int * max(1, len(gates)))(*[g.KIND for g in gates])
        flat: List[int] = []
        for g in gates:
            flat += g.fields13()
        fl = (C.c_uint64 * max(1, len(flat)))(*flat)
        ptr, n, nreg, dot = C.c_void_p(), C.c_size_t(), C.c_uint64(), C.c_char_p()
        rc = lib().orc_compile(self.immediate_count, self.incrementer.val, kinds, fl, len(gates),
                               C.byref(ptr), C.byref(n), C.byref(nreg), C.byref(dot))
        if rc != 0:
            raise ReferencePanic("graph index out of bounds (compile.rs:137)")
        if dot.value is not None:
            self.last_dot = dot.value.decode()
            lib().orc_free(C.cast(dot, C.c_void_p))
        return Simulation(registers=[False] * nreg.value, _ops_array=_take_ops(ptr, n.value))


def _as_orc_ops(ops) -> np.ndarray:
    """Accept tuple ops, the oracle's array, or the product's 32-byte array (same bytes: u32 tag + u32 zero)."""
    if hasattr(ops, "ops_array"):  # a Simulation (oracle's or the product mirror's)
        ops = ops.ops_array
    if not isinstance(ops, np.ndarray):
        return ops_to_array(ops)
    if ops.dtype != OP_DTYPE:
        assert ops.dtype.itemsize == 32
        ops = np.ascontiguousarray(ops).view(OP_DTYPE)
    return np.ascontiguousarray(ops)


# ---- batch drivers (bit-sliced [row][stride] uint32) -----------------------------
def _batch(fn, ops, n_regs, imm, n_vectors, out_regs, steps, all_steps, *extra):
    ops = _as_orc_ops(ops)
    imm = np.ascontiguousarray(imm, dtype=np.uint32)
    if imm.ndim == 2:
        imm = imm[None]
    assert imm.ndim == 3 and imm.shape[0] == steps
    _, n_imm, stride = imm.shape
    assert stride * 32 >= n_vectors
    oregs = np.ascontiguousarray(out_regs, dtype=np.uint64)
    out = np.zeros(((steps if all_steps else 1), len(oregs), stride), dtype=np.uint32)
    rc = fn(ops.ctypes.data, len(ops), n_regs, imm.ctypes.data, n_imm, stride, n_vectors, steps,
            oregs.ctypes.data, len(oregs), out.ctypes.data, 1 if all_steps else 0, *extra)
    if rc != 0:
        raise ReferencePanic("index out of bounds")
    return out if all_steps else out[0]


def run_batch_scalar(ops, n_regs, imm, n_vectors, out_regs, steps=1, all_steps=False, threads=1):
    """One reference-shaped run() per vector (oracle tier L0)."""
    return _batch(lib().orc_run_batch_scalar, ops, n_regs, imm, n_vectors, out_regs, steps, all_steps, threads)


def run_batch_sliced(ops, n_regs, imm, n_vectors, out_regs, steps=1, all_steps=False):
    """64-lane bit-sliced CPU evaluation of the same op list (oracle tier L1)."""
    return _batch(lib().orc_run_batch_sliced, ops, n_regs, imm, n_vectors, out_regs, steps, all_steps)
