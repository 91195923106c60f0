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
    Self::with_flags(sim, n_immediates, out_regs, device, 0)
  }

  /// `flags`: `CLG_LV_*` of the levelize pass (`CLG_LV_STATEFUL` carries registers across calls: latches).
  pub fn with_flags(sim: &Simulation, n_immediates: usize, out_regs: &[usize], device: i32, flags: u32) -> Result<Self, Error> {
    let ops = to_c_ops(&sim.ops);
    let regs: Vec<u8> = sim.registers.iter().map(|&b| b as u8).collect();
    let outs: Vec<u64> = out_regs.iter().map(|&r| r as u64).collect();
    unsafe {
      let mut csim = ptr::null_mut();
      check(clg_simulation_new(ops.as_ptr(), ops.len(), regs.as_ptr(), regs.len(), &mut csim))?;
      let mut flat = ptr::null_mut();
      let rc = clg_levelize(csim, n_immediates, if outs.is_empty() { ptr::null() } else { outs.as_ptr() }, outs.len(), flags, &mut flat);
      clg_simulation_destroy(csim);
      check(rc)?;
      let mut ctx = ptr::null_mut();
      check(clg_ctx_create(device, &mut ctx))?;
      let mut exec = ptr::null_mut();
      check(clg_exec_create(ctx, flat, CLG_MODE_AUTO as i32, &mut exec))?;
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

  /// The same with carried registers (`CLG_LV_STATEFUL`): `state` is `[n_state][stride]` words, one carried register
  /// per row, updated in place -- lane v behaves like one `Simulation` whose `run` is called once per call.
  pub fn run_batch_stateful(&self, imm: &[u32], state: &mut [u32], n_vectors: usize) -> Result<Vec<u32>, Error> {
    let stride = round_stride(n_vectors);
    assert_eq!(imm.len(), self.info.n_inputs as usize * stride);
    assert_eq!(state.len(), self.info.n_state as usize * stride);
    let mut out = vec![0u32; self.info.n_outputs as usize * stride];
    check(unsafe { clg_exec_run_host(self.exec, imm.as_ptr(), out.as_mut_ptr(), state.as_mut_ptr(), n_vectors, stride) })?;
    Ok(out)
  }

  /// `n_steps` consecutive `run` calls per lane in one call: `imm` is `[n_steps][n_immediates][stride]`, the result
  /// `[n_steps][n_outputs][stride]`; carried registers travel from step to step through `state`.
  pub fn run_steps(&self, imm: &[u32], state: &mut [u32], n_vectors: usize, n_steps: usize) -> Result<Vec<u32>, Error> {
    let stride = round_stride(n_vectors);
    let (ni, no) = (self.info.n_inputs as usize * stride, self.info.n_outputs as usize * stride);
    assert_eq!(imm.len(), n_steps * ni);
    let mut out = vec![0u32; n_steps * no];
    for s in 0..n_steps {
      let sp = if state.is_empty() { ptr::null_mut() } else { state.as_mut_ptr() };
      check(unsafe { clg_exec_run_host(self.exec, imm[s * ni..].as_ptr(), out[s * no..].as_mut_ptr(), sp, n_vectors, stride) })?;
    }
    Ok(out)
  }

  /// Choose the mode for batches of `n_vectors`, build its plan and tune it now (otherwise the first run does it).
  pub fn prepare(&self, n_vectors: usize) -> Result<(), Error> {
    check(unsafe { clg_exec_prepare(self.exec, n_vectors) })
  }

  /// Name of the kernel the most recent run launched (`clg_coop_rows`, `clg_spec`, ...).
  pub fn kernel_name(&self) -> String {
    unsafe { CStr::from_ptr(clg_exec_kernel_name(self.exec)) }.to_string_lossy().into_owned()
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

/// One program on several GPUs of one box (`clg_sharded_*`): the vector batch is cut into contiguous word ranges, one
/// per device; contexts, executors and staging buffers live as long as this value. No collective: nothing is reduced.
pub struct ShardedSimulation {
  flat: *mut clg_flat,
  sharded: *mut clg_sharded,
  pub info: clg_flat_info,
  pub last_seconds: Vec<f64>,
}

impl ShardedSimulation {
  pub fn new(sim: &Simulation, n_immediates: usize, out_regs: &[usize], devices: &[i32], flags: u32) -> Result<Self, Error> {
    let ops = to_c_ops(&sim.ops);
    let regs: Vec<u8> = sim.registers.iter().map(|&b| b as u8).collect();
    let outs: Vec<u64> = out_regs.iter().map(|&r| r as u64).collect();
    unsafe {
      let mut csim = ptr::null_mut();
      check(clg_simulation_new(ops.as_ptr(), ops.len(), regs.as_ptr(), regs.len(), &mut csim))?;
      let mut flat = ptr::null_mut();
      let rc = clg_levelize(csim, n_immediates, if outs.is_empty() { ptr::null() } else { outs.as_ptr() }, outs.len(), flags, &mut flat);
      clg_simulation_destroy(csim);
      check(rc)?;
      let mut info = clg_flat_info::default();
      check(clg_flat_get_info(flat, &mut info))?;
      let mut sharded = ptr::null_mut();
      check(clg_sharded_create(flat, CLG_MODE_AUTO as i32, devices.as_ptr(), devices.len() as i32, &mut sharded))?;
      Ok(Self { flat, sharded, info, last_seconds: vec![0.0; devices.len()] })
    }
  }

  /// `n_steps` consecutive runs per lane (1 for a plain batch); `state` is empty for stateless programs.
  pub fn run_steps(&mut self, imm: &[u32], state: &mut [u32], n_vectors: usize, n_steps: usize) -> Result<Vec<u32>, Error> {
    let stride = round_stride(n_vectors);
    assert_eq!(imm.len(), n_steps * self.info.n_inputs as usize * stride);
    let mut out = vec![0u32; n_steps * self.info.n_outputs as usize * stride];
    let sp = if state.is_empty() { ptr::null_mut() } else { state.as_mut_ptr() };
    check(unsafe {
      clg_sharded_run_steps_host(self.sharded, imm.as_ptr(), out.as_mut_ptr(), sp, n_vectors, stride, n_steps, self.last_seconds.as_mut_ptr())
    })?;
    Ok(out)
  }
}

impl Drop for ShardedSimulation {
  fn drop(&mut self) {
    unsafe {
      clg_sharded_destroy(self.sharded);
      clg_flat_destroy(self.flat);
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
    check(clg_sim_runner_create(ctx, csim, imm.len(), CLG_MODE_AUTO as i32, &mut runner))?;
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
