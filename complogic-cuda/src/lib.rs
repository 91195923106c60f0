//! `complogic-cuda`: run a compiled complogic `Simulation` on a B200 through libcomplogic_cuda.so.
//!
//! ```ignore
//! let mut compiler = Compiler::new(2);
//! let out = compiler.alloc();
//! let sim = compiler.compile(vec![&Gate::from(And { a: 0, b: 1, out })]);   // unchanged reference code
//! let cuda = CudaSimulation::from_simulation(&sim, 2, &[out], 0)?;          // levelize + upload
//! let result = cuda.run_batch(&imm_sliced, n_vectors)?;                      // Simulation::run x n_vectors
//! ```
//! NOT COMPILED in the build image of the backend (no Rust toolchain there); see INTEGRATION.md.
pub mod ffi;

use complogic::{Op, Simulation};
use ffi::*;
use std::ffi::CStr;
use std::ptr;

#[derive(Debug)]
pub struct Error {
  pub code: i32,
  pub message: String,
}

fn check(rc: i32) -> Result<(), Error> {
  if rc == CLG_OK {
    return Ok(());
  }
  let message = unsafe { CStr::from_ptr(clg_last_error()) }.to_string_lossy().into_owned();
  Err(Error { code: rc, message })
}

fn to_c_ops(ops: &[Op]) -> Vec<clg_op> {
  ops
    .iter()
    .map(|op| match *op {
      Op::Nand(a, b, out) => clg_op { tag: CLG_OP_NAND, reserved: 0, a: a as u64, b: b as u64, c: out as u64 },
      Op::Set(id, val) => clg_op { tag: CLG_OP_SET, reserved: 0, a: id as u64, b: val as u64, c: 0 },
    })
    .collect()
}

/// Words per row for `n_vectors` lanes (rows are 16-byte aligned).
pub fn round_stride(n_vectors: usize) -> usize {
  ((n_vectors + 31) / 32 + 3) / 4 * 4
}

/// A program resident on one GPU: the batched `Simulation::run` (simulation.rs:16-30).
pub struct CudaSimulation {
  ctx: *mut clg_ctx,
  flat: *mut clg_flat,
  exec: *mut clg_exec,
  pub info: clg_flat_info,
}

impl CudaSimulation {
  /// `n_immediates` = length of the immediates slice every lane's `run` receives; `out_regs` = registers to read back
  /// (empty = all of `Simulation::registers`).
  pub fn from_simulation(sim: &Simulation, n_immediates: usize, out_regs: &[usize], device: i32) -> Result<Self, Error> {
    let ops = to_c_ops(&sim.ops);
    let regs: Vec<u8> = sim.registers.iter().map(|&b| b as u8).collect();
    let outs: Vec<u64> = out_regs.iter().map(|&r| r as u64).collect();
    unsafe {
      let mut csim = ptr::null_mut();
      check(clg_simulation_new(ops.as_ptr(), ops.len(), regs.as_ptr(), regs.len(), &mut csim))?;
      let mut flat = ptr::null_mut();
      let rc = clg_levelize(csim, n_immediates, if outs.is_empty() { ptr::null() } else { outs.as_ptr() }, outs.len(), 0, &mut flat);
      clg_simulation_destroy(csim);
      check(rc)?;
      let mut ctx = ptr::null_mut();
      check(clg_ctx_create(device, &mut ctx))?;
      let mut exec = ptr::null_mut();
      check(clg_exec_create(ctx, flat, CLG_MODE_AUTO, &mut exec))?;
      let mut info = clg_flat_info::default();
      check(clg_flat_get_info(flat, &mut info))?;
      Ok(Self { ctx, flat, exec, info })
    }
  }

  /// Bit-sliced batch: `imm` is `[n_immediates][stride]` words (vector v = bit v%32 of word v/32); returns
  /// `[n_outputs][stride]`. Lane v is a fresh all-false `Simulation` followed by `run(imm_v)`.
  pub fn run_batch(&self, imm: &[u32], n_vectors: usize) -> Result<Vec<u32>, Error> {
    let stride = round_stride(n_vectors);
    assert_eq!(imm.len(), self.info.n_inputs as usize * stride);
    let mut out = vec![0u32; self.info.n_outputs as usize * stride];
    check(unsafe { clg_exec_run_host(self.exec, imm.as_ptr(), out.as_mut_ptr(), ptr::null_mut(), n_vectors, stride) })?;
    Ok(out)
  }

  /// `&[bool]`-shaped batch: one immediates slice per vector, as `Simulation::run` takes it.
  pub fn run_bools(&self, immediates: &[Vec<bool>]) -> Result<Vec<Vec<bool>>, Error> {
    let (n, ni, no) = (immediates.len(), self.info.n_inputs as usize, self.info.n_outputs as usize);
    let flat: Vec<u8> = immediates.iter().flat_map(|v| v.iter().map(|&b| b as u8)).collect();
    assert_eq!(flat.len(), n * ni);
    let mut out = vec![0u8; n * no];
    check(unsafe { clg_exec_run_bools_host(self.exec, flat.as_ptr(), out.as_mut_ptr(), n) })?;
    Ok(out.chunks(no.max(1)).map(|c| c.iter().map(|&b| b != 0).collect()).collect())
  }
}

impl Drop for CudaSimulation {
  fn drop(&mut self) {
    unsafe {
      clg_exec_destroy(self.exec);
      clg_flat_destroy(self.flat);
      clg_ctx_destroy(self.ctx);
    }
  }
}

/// Drop-in for a single `Simulation::run` with the reference's carried-register semantics, on the device.
pub fn run_on_gpu(sim: &mut Simulation, immediates: &[bool], device: i32) -> Result<(), Error> {
  let ops = to_c_ops(&sim.ops);
  let regs: Vec<u8> = sim.registers.iter().map(|&b| b as u8).collect();
  let imm: Vec<u8> = immediates.iter().map(|&b| b as u8).collect();
  unsafe {
    let (mut csim, mut ctx, mut runner) = (ptr::null_mut(), ptr::null_mut(), ptr::null_mut());
    check(clg_simulation_new(ops.as_ptr(), ops.len(), regs.as_ptr(), regs.len(), &mut csim))?;
    check(clg_ctx_create(device, &mut ctx))?;
    check(clg_sim_runner_create(ctx, csim, imm.len(), CLG_MODE_AUTO, &mut runner))?;
    let rc = clg_sim_runner_run(runner, imm.as_ptr(), imm.len());
    if rc == CLG_OK {
      let mut n = 0usize;
      let p = clg_simulation_registers(csim, &mut n);
      for (i, r) in sim.registers.iter_mut().enumerate().take(n) {
        *r = *p.add(i) != 0;
      }
    }
    clg_sim_runner_destroy(runner);
    clg_ctx_destroy(ctx);
    clg_simulation_destroy(csim);
    check(rc)
  }
}
