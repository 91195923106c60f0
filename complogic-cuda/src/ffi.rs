//! Raw bindings of `include/complogic_cuda.h` (cudarc-style: plain `extern "C"`, no bindgen).
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_void};

pub const CLG_OK: c_int = 0;
pub const CLG_E_BAD_REGISTER: c_int = 2;

pub const CLG_OP_NAND: u32 = 0;
pub const CLG_OP_SET: u32 = 1;

/// `enum Op` flattened: tag 0 = Nand(a, b, out = c), tag 1 = Set(id = a, val = b != 0).
#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct clg_op {
  pub tag: u32,
  pub reserved: u32,
  pub a: u64,
  pub b: u64,
  pub c: u64,
}

#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct clg_flat_info {
  pub ref_nand_count: u64,
  pub ref_set_count: u64,
  pub n_inputs: u32,
  pub n_outputs: u32,
  pub n_state: u32,
  pub n_luts: u32,
  pub n_levels: u32,
  pub max_level_width: u32,
  pub level_slots: u32,
  pub lane_slots: u32,
  pub level_instrs: u32,
  pub lane_instrs: u32,
}

pub enum clg_simulation {}
pub enum clg_flat {}
pub enum clg_ctx {}
pub enum clg_exec {}
pub enum clg_sim_runner {}

pub const CLG_LV_STATEFUL: u32 = 1;
pub const CLG_MODE_AUTO: c_int = 0;

extern "C" {
  pub fn clg_last_error() -> *const c_char;
  pub fn clg_simulation_new(ops: *const clg_op, n_ops: usize, registers: *const u8, n_registers: usize,
                            out: *mut *mut clg_simulation) -> c_int;
  pub fn clg_simulation_destroy(sim: *mut clg_simulation);
  pub fn clg_simulation_registers(sim: *mut clg_simulation, n: *mut usize) -> *mut u8;
  pub fn clg_levelize(sim: *const clg_simulation, n_immediates: usize, out_regs: *const u64, n_out: usize,
                      flags: u32, out: *mut *mut clg_flat) -> c_int;
  pub fn clg_flat_destroy(f: *mut clg_flat);
  pub fn clg_flat_get_info(f: *const clg_flat, info: *mut clg_flat_info) -> c_int;
  pub fn clg_ctx_create(device: c_int, out: *mut *mut clg_ctx) -> c_int;
  pub fn clg_ctx_destroy(ctx: *mut clg_ctx);
  pub fn clg_exec_create(ctx: *mut clg_ctx, flat: *const clg_flat, mode: c_int, out: *mut *mut clg_exec) -> c_int;
  pub fn clg_exec_destroy(ex: *mut clg_exec);
  pub fn clg_exec_run(ex: *mut clg_exec, imm: *const u32, out: *mut u32, state: *mut u32, n_vectors: usize,
                      stride: usize, stream: *mut c_void) -> c_int;
  pub fn clg_exec_run_steps(ex: *mut clg_exec, imm: *const u32, out: *mut u32, state: *mut u32, n_vectors: usize,
                            stride: usize, n_steps: usize, stream: *mut c_void) -> c_int;
  pub fn clg_exec_run_host(ex: *mut clg_exec, imm: *const u32, out: *mut u32, state: *mut u32, n_vectors: usize,
                           stride: usize) -> c_int;
  pub fn clg_exec_run_bools_host(ex: *mut clg_exec, imm: *const u8, out: *mut u8, n_vectors: usize) -> c_int;
  pub fn clg_sim_runner_create(ctx: *mut clg_ctx, sim: *mut clg_simulation, n_immediates: usize, mode: c_int,
                               out: *mut *mut clg_sim_runner) -> c_int;
  pub fn clg_sim_runner_destroy(r: *mut clg_sim_runner);
  pub fn clg_sim_runner_run(r: *mut clg_sim_runner, immediates: *const u8, n_immediates: usize) -> c_int;
  pub fn clg_run_sharded_host(flat: *const clg_flat, mode: c_int, devices: *const c_int, n_devices: c_int,
                              imm: *const u32, out: *mut u32, n_vectors: usize, stride: usize,
                              seconds_out: *mut f64) -> c_int;
}
