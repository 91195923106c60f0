"""ctypes binding of libcomplogic_cuda.so (include/complogic_cuda.h). No torch types cross it."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(HERE, "libcomplogic_cuda.so")

(CLG_OK, CLG_E_INVALID, CLG_E_BAD_REGISTER, CLG_E_NOMEM, CLG_E_CUDA, CLG_E_NO_DRIVER, CLG_E_UNSUPPORTED, CLG_E_PTX,
 CLG_E_FORMAT) = range(9)
MODE_AUTO, MODE_LANE, MODE_COOP, MODE_SPEC = range(4)
MODE_NAMES = {MODE_AUTO: "auto", MODE_LANE: "lane", MODE_COOP: "coop", MODE_SPEC: "spec"}
LV_STATEFUL, LV_NO_FUSE, LV_NO_LANE_STREAM, LV_NO_PARTS, LV_NO_ROWS = 1, 2, 4, 8, 16


class ClgOp(C.Structure):
    _fields_ = [("tag", C.c_uint32), ("reserved", C.c_uint32), ("a", C.c_uint64), ("b", C.c_uint64), ("c", C.c_uint64)]


class ClgGate(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("reserved", C.c_uint32), ("f", C.c_uint64 * 13)]


class ClgFlatInfo(C.Structure):
    _fields_ = [("ref_nand_count", C.c_uint64), ("ref_set_count", C.c_uint64), ("n_inputs", C.c_uint32),
                ("n_outputs", C.c_uint32), ("n_state", C.c_uint32), ("n_luts", C.c_uint32), ("n_levels", C.c_uint32),
                ("max_level_width", C.c_uint32), ("level_slots", C.c_uint32), ("lane_slots", C.c_uint32),
                ("level_instrs", C.c_uint32), ("lane_instrs", C.c_uint32), ("n_parts", C.c_uint32), ("reserved", C.c_uint32),
                ("rows_segments", C.c_uint32), ("rows_slots", C.c_uint32), ("rows_instrs", C.c_uint32),
                ("rows_from_reg", C.c_uint32), ("rows_no_store", C.c_uint32), ("rows_warps", C.c_uint32)]


vp, sz, u64, u32, i32 = C.c_void_p, C.c_size_t, C.c_uint64, C.c_uint32, C.c_int
P = C.POINTER

# name -> (restype, argtypes); every symbol include/complogic_cuda.h declares
PROTOTYPES = {
    "clg_last_error": (C.c_char_p, []),
    "clg_abi_version": (u32, []),
    "clg_free": (None, [vp]),
    "clg_gate_create": (i32, [P(ClgGate), P(u64), P(vp), P(sz)]),
    "clg_simulation_new": (i32, [vp, sz, vp, sz, P(vp)]),
    "clg_simulation_destroy": (None, [vp]),
    "clg_simulation_ops": (vp, [vp, P(sz)]),
    "clg_simulation_registers": (vp, [vp, P(sz)]),
    "clg_simulation_register": (i32, [vp, sz, P(i32)]),
    "clg_simulation_nand_count": (u64, [vp]),
    "clg_compiler_new": (i32, [u64, P(vp)]),
    "clg_compiler_destroy": (None, [vp]),
    "clg_compiler_reset_ops": (None, [vp]),
    "clg_compiler_reset_incrementer": (None, [vp]),
    "clg_compiler_alloc": (u64, [vp]),
    "clg_compiler_immediate_count": (u64, [vp]),
    "clg_compiler_set_immediate_count": (None, [vp, u64]),
    "clg_compiler_incrementer": (u64, [vp]),
    "clg_compiler_set_incrementer": (None, [vp, u64]),
    "clg_compiler_ops": (vp, [vp, P(sz)]),
    "clg_compiler_compile": (i32, [vp, vp, sz, P(vp)]),
    "clg_compiler_graph_dot": (C.c_char_p, [vp]),
    "clg_levelize": (i32, [vp, sz, vp, sz, u32, P(vp)]),
    "clg_flat_destroy": (None, [vp]),
    "clg_flat_bytes": (i32, [vp, P(vp), P(sz)]),
    "clg_flat_from_bytes": (i32, [vp, sz, P(vp)]),
    "clg_flat_get_info": (i32, [vp, P(ClgFlatInfo)]),
    "clg_flat_state_regs": (i32, [vp, P(vp), P(sz)]),
    "clg_flat_spec_ptx": (i32, [vp, i32, P(vp)]),
    "clg_flat_spec_steps_ptx": (i32, [vp, i32, P(vp)]),
    "clg_flat_spec_chunk_ptx": (i32, [vp, i32, u32, P(u32), P(u32), P(vp)]),
    "clg_flat_spec_batch_chunk_ptx": (i32, [vp, i32, u32, P(u32), P(u32), P(vp)]),
    "clg_flat_spec_chain_info": (i32, [vp, i32, P(u32), P(u32), P(u32)]),
    "clg_flat_spec_batch_chain_info": (i32, [vp, i32, P(u32), P(u32), P(u32), P(u32)]),
    "clg_ctx_create": (i32, [i32, P(vp)]),
    "clg_ctx_destroy": (None, [vp]),
    "clg_ctx_device_name": (i32, [vp, C.c_char_p, sz]),
    "clg_ctx_sm_count": (i32, [vp]),
    "clg_ctx_synchronize": (i32, [vp]),
    "clg_dev_alloc": (i32, [vp, sz, P(vp)]),
    "clg_dev_free": (i32, [vp, vp]),
    "clg_dev_memset": (i32, [vp, vp, i32, sz, vp]),
    "clg_memcpy_h2d": (i32, [vp, vp, vp, sz, vp]),
    "clg_memcpy_d2h": (i32, [vp, vp, vp, sz, vp]),
    "clg_host_alloc_pinned": (i32, [vp, sz, P(vp)]),
    "clg_host_free_pinned": (i32, [vp, vp]),
    "clg_exec_create": (i32, [vp, vp, i32, P(vp)]),
    "clg_exec_destroy": (None, [vp]),
    "clg_exec_mode": (i32, [vp]),
    "clg_exec_words_per_thread": (i32, [vp]),
    "clg_exec_launches_per_run": (i32, [vp]),
    "clg_exec_kernel_name": (C.c_char_p, [vp]),
    "clg_exec_prepare": (C.c_int, [vp, C.c_size_t]),
    "clg_exec_ptx": (C.c_char_p, [vp]),
    "clg_exec_run": (i32, [vp, vp, vp, vp, sz, sz, vp]),
    "clg_exec_run_steps": (i32, [vp, vp, vp, vp, sz, sz, sz, vp]),
    "clg_exec_run_host": (i32, [vp, vp, vp, vp, sz, sz]),
    "clg_exec_run_bools_host": (i32, [vp, vp, vp, sz]),
    "clg_slice_bools": (i32, [vp, vp, vp, sz, sz, sz, vp]),
    "clg_unslice_bools": (i32, [vp, vp, vp, sz, sz, sz, vp]),
    "clg_fill_random": (i32, [vp, vp, sz, sz, u64, vp]),
    "clg_measure_lop3_peak": (i32, [vp, P(C.c_double)]),
    "clg_sim_runner_create": (i32, [vp, vp, sz, i32, P(vp)]),
    "clg_sim_runner_destroy": (None, [vp]),
    "clg_sim_runner_run": (i32, [vp, vp, sz]),
    "clg_shard_range": (i32, [sz, i32, i32, P(sz), P(sz)]),
    "clg_run_sharded_host": (i32, [vp, i32, P(i32), i32, vp, vp, sz, sz, P(C.c_double)]),
    "clg_sharded_create": (i32, [vp, i32, P(i32), i32, P(vp)]),
    "clg_sharded_destroy": (None, [vp]),
    "clg_sharded_n_devices": (i32, [vp]),
    "clg_sharded_run_host": (i32, [vp, vp, vp, vp, sz, sz, P(C.c_double)]),
    "clg_sharded_run_steps_host": (i32, [vp, vp, vp, vp, sz, sz, sz, P(C.c_double)]),
    "clg_simulation_to_json": (i32, [vp, P(vp)]),
    "clg_simulation_from_json": (i32, [C.c_char_p, P(vp)]),
    "clg_compiler_to_json": (i32, [vp, P(vp)]),
    "clg_compiler_from_json": (i32, [C.c_char_p, P(vp)]),
    "clg_gate_to_json": (i32, [P(ClgGate), P(vp)]),
    "clg_gate_from_json": (i32, [C.c_char_p, P(ClgGate)]),
}

_lib = None


class BackendMissing(RuntimeError):
    """libcomplogic_cuda.so is not built. There is deliberately no fallback."""


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise BackendMissing(f"{SO_PATH} is missing: run `python -m complogic_b200.build` (there is no CPU fallback)")
        L = C.CDLL(SO_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


class ClgError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"[clg status {code}] {msg}")
        self.code = code


class ReferencePanic(ClgError):
    """CLG_E_BAD_REGISTER: the reference would have panicked (index out of bounds)."""


def check(rc: int) -> None:
    if rc != CLG_OK:
        msg = lib().clg_last_error().decode(errors="replace")
        raise (ReferencePanic if rc == CLG_E_BAD_REGISTER else ClgError)(rc, msg)
