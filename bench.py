#!/usr/bin/env python
"""bench.py -- the hot path's headline metric on B200: NAND gate-evals/s (gates x vectors / s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload dag1m|mul16|adder32]

One "step" is one pass of the hot path (`Simulation::run` for every vector of the batch,
complogic/src/simulation.rs:16-30) over one batch of synthetic random input vectors.
Default workload = the configuration BASELINE.json's target is quoted on ("1M-gate circuit x 2^24
vectors", configs[3]): the synthetic random levelized NAND DAG, 1 048 576 gates, depth 256, 2^24 vectors
PER GPU (weak scaling: every rank evaluates its own shard, no data-path collective).

Prints ONE JSON line (rank 0). Contract keys: metric/value/unit/n_gpus/steps/warmup/ms_per_step/
higher_is_better/scaling/vs_baseline/dtype/data/config/clocks/e2e/gpu_launches + roofline + cpu_baseline.
`--impl reference` times the reference's CPU VM (the oracle's literal restatement, all host threads) on a
bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "NAND gate-evals/sec (gates x vectors / s)"
UNIT = "gate-evals/s"
L2_BYTES = 126 * 1024 * 1024


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), float(d.get("sm_max_mhz", 1965.0)), "measured"
    return 6650.0, 1965.0, "fallback"


def build_workload(name: str, api=None):
    """-> (Circuit, default log2 vectors per GPU, seed)"""
    from complogic_b200 import circuits

    if name == "dag1m":
        return circuits.random_dag(256, 4096, 1024, 2, 1, api=api), 24, 0xDA6
    if name == "dag_thin":
        return circuits.random_dag(1500, 48, 64, 2, 31, api=api), 23, 0xDA9
    if name == "dag64k":
        return circuits.random_dag(64, 1024, 256, 2, 3, api=api), 23, 0xDA8
    if name == "dag1m_kinf":
        return circuits.random_dag(256, 4096, 1024, 0, 1, api=api), 24, 0xDA7
    if name == "mul16":
        return circuits.array_multiplier(16, api=api), 26, 0xC0FFEF
    if name == "adder32":
        return circuits.ripple_adder(32, api=api), 24, 0xC0FFEE
    if name == "mul64x128":
        # config 5: ONE input vector through 128 flattened 64x64 multipliers (1.08e7 gates) -- latency, not throughput
        return circuits.flatten_instances(circuits.array_multiplier(64, api=api), 128, api=api), 0, 0x5EED
    if name == "mul64":
        return circuits.array_multiplier(64, api=api), 22, 0xC0FFF0
    if name in ("mul64x4", "mul64x8", "mul64x12"):
        m = int(name[6:])
        return circuits.flatten_instances(circuits.array_multiplier(64, api=api), m, api=api), 24 if m == 12 else 21, 0xC0FFF1
    if name == "mul32x48":
        return circuits.flatten_instances(circuits.array_multiplier(32, api=api), 48, api=api), 24, 0xC0FFF2
    if name == "adder8":
        return circuits.ripple_adder(8, api=api), 17, 0xC0FFED
    if name == "latch1500":
        return circuits.latch_chain(1500, api=api), 24, 0x1A7C4
    raise SystemExit(f"unknown workload {name}")


WORKLOAD_DESC = {
    "dag1m": "synthetic random levelized NAND DAG, 1 048 576 gates (256 levels x 4096), 1024 immediates, window K=2, "
             "4096 outputs (BASELINE configs[3])",
    "dag_thin": "a long thin random DAG: 72 000 gates (1500 levels x 48), 64 immediates -- ~90 live values, 39 k instructions",
    "dag64k": "a smaller random levelized DAG: 65 536 gates (64 levels x 1024), 256 immediates, window K=2 -- narrow levels, "
              "several cooperative CTAs per SM",
    "dag1m_kinf": "the same random levelized DAG with no locality window (K = infinity: inputs drawn from ALL earlier levels); "
                  "most of its 1 048 576 gates cannot reach the 4096 outputs and are eliminated by the pass",
    "mul16": "16x16 array multiplier, 4 896 NANDs with the reference's gate expansions (BASELINE configs[2])",
    "adder32": "32-bit ripple-carry adder, 608 NANDs (BASELINE configs[1])",
    "adder8": "8-bit ripple-carry adder, 152 NANDs (BASELINE configs[0])",
    "mul64": "64x64 array multiplier, 84 096 NANDs (one instance of config 5's circuit), batched",
    "mul64x4": "64x64 array multiplier x 4 instances flattened, batched (49 408 instructions)",
    "latch1500": "a SEQUENTIAL circuit: 1 500 D latches in a row (24 000 NANDs), 6 000 carried registers per lane rewritten "
                 "in place every step; one step per launch on 2^24 lanes: bound by HBM (13 502 rows of 2 MB per step)",
    "mul64x12": "a STRUCTURED 1M-gate circuit: 64x64 array multiplier x 12 instances flattened = 1 009 152 NANDs, 2^24 vectors; "
                "runs as 2 launches of two specialised kernels (the halves of an instance, each below the instruction-cache cliff; the twelve instances side by side)",
    "mul64x8": "64x64 array multiplier x 8 instances flattened = 672 768 NANDs, batched: a program too long for one "
               "specialised kernel (98 816 instructions), run as a chain of chunk kernels",
    "mul32x48": "a STRUCTURED 1M-gate circuit of narrower parts: 32x32 array multiplier x 48 instances flattened = 986 112 NANDs, "
                "2^24 vectors; one instance per kernel (50 KB of SASS: it stays in the instruction cache), all 48 side by side in one launch",
    "mul64x128": "64x64 array multiplier (84 096 NANDs) x 128 instances flattened into one program = 10 764 288 gates, "
                 "ONE input vector (BASELINE configs[4]; M = 128 instances gives the ~10M gates the config names, see SURVEY "
                 "8d). --mode coop: the warp-cooperative intra-level mode, one CTA per part; auto: the instances are equal "
                 "independent chunks, so ONE launch of one specialised kernel with an instance per blockIdx.y",
}


def bind_to_gpu_numa_node(device: int):
    """One process per GPU: run this process on the CPUs next to its GPU, so that the pinned host buffers of the e2e
    leg (first touched, and therefore placed, by this process) sit on the GPU's own NUMA node. Without it eight ranks
    share whatever node the launcher started them on and the copies queue behind one memory controller."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(device)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if m >> b & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return {"cpus": len(cpus), "first_cpu": min(cpus)}
    except Exception as e:  # no NVML, no permission: run unbound
        return {"error": str(e)[:80]}
    return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(device), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0: float, t1: float):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        for ts, line in self.rows:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            inside = t0 - 0.05 <= ts <= t1 + 0.15
            try:
                if inside:
                    sm.append(float(parts[0]))
                    power.append(float(parts[2]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            if inside:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        if not sm:  # region shorter than the sampling period: fall back to every sample taken
            for ts, line in self.rows:
                try:
                    sm.append(float(line.split(",")[0]))
                except ValueError:
                    pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "power_w_max": max(power) if power else None}


def cpu_reference_throughput(circ, seed: int, threads: int, target_s: float):
    """The reference VM on the host cores: literal restatement of simulation.rs:16-30 (32-byte ops, one byte per
    register, bounds checks, one run() per vector), `threads` pthreads each on its own slice of vectors."""
    from oracle import oracle
    from tests.helpers import random_sliced, stride_for

    gates = circ.sim.nand_count
    R = len(circ.sim.registers)
    ops = oracle._as_orc_ops(circ.sim)
    # calibrate on one word per thread, then size the sample for ~target_s
    nv = 32 * threads
    imm = random_sliced(circ.n_immediates, stride_for(nv), seed)
    t0 = time.perf_counter()
    oracle.run_batch_scalar(ops, R, imm, nv, circ.out_regs, threads=threads)
    dt = max(time.perf_counter() - t0, 1e-6)
    scale = max(1, int(target_s / dt))
    nv = 32 * threads * min(scale, 1 << 16)
    imm = random_sliced(circ.n_immediates, stride_for(nv), seed)
    t0 = time.perf_counter()
    oracle.run_batch_scalar(ops, R, imm, nv, circ.out_regs, threads=threads)
    dt = time.perf_counter() - t0
    return gates * nv / dt, nv, dt


def run_reference(args):
    """--impl reference: rank 0 only; every step is a bounded sample of the same workload on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle
    from tests.helpers import random_sliced, stride_for

    oracle.build()
    circ, lg, seed = build_workload(args.workload, api=None)
    threads = os.cpu_count() or 1
    gates = circ.sim.nand_count
    R = len(circ.sim.registers)
    ops = oracle._as_orc_ops(circ.sim)
    # size a step for ~ (budget / (steps + warmup)) seconds
    per_step = max(0.5, min(20.0, 120.0 / (args.steps + args.warmup)))
    _, nv_cal, dt_cal = cpu_reference_throughput(circ, seed, threads, 1.0)
    nv = max(32 * threads, int(nv_cal * per_step / max(dt_cal, 1e-6)) // (32 * threads) * (32 * threads))
    imm = random_sliced(circ.n_immediates, stride_for(nv), seed)
    times = []
    for s in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        oracle.run_batch_scalar(ops, R, imm, nv, circ.out_regs, threads=threads)
        if s >= args.warmup:
            times.append(time.perf_counter() - t0)
    ms = 1e3 * sum(times) / len(times)
    value = gates * nv / (ms * 1e-3)
    sample = f"{nv} vectors per step ({threads} threads x {nv // threads} lanes) of {args.workload}, {gates} NANDs"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic",
        "config": {"workload": args.workload, "description": WORKLOAD_DESC[args.workload], "gates": gates,
                   "vectors_per_step": nv, "note": "reference CPU VM: C restatement of simulation.rs:16-30 (the Rust "
                   "reference cannot be built in this image: no cargo/rustc), all host threads"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def parity_sample(circ, imm_h, out_h, nv, n_groups=16, seed=12345):
    """Compare sampled word columns of a benched output with the oracle VM (bit-sliced tier, validated against the
    scalar VM in tests/): groups of 4 words -- the first, the last, and random ones -- of every output row."""
    from oracle import oracle

    oracle.build()
    words = (nv + 31) // 32
    groups = max(1, words // 4)
    rng = np.random.default_rng(seed)
    pick = sorted({0, groups - 1, *[int(x) for x in rng.integers(0, groups, size=max(0, n_groups - 2))]})
    cols = np.concatenate([np.arange(4 * g, min(words, 4 * g + 4)) for g in pick])
    sub = np.ascontiguousarray(imm_h[:, cols])
    n_sub = len(cols) * 32
    want = oracle.run_batch_sliced(circ.sim, len(circ.sim.registers), sub, n_sub, circ.out_regs)
    got = np.ascontiguousarray(out_h[:, cols])
    if cols[-1] == words - 1 and nv % 32:  # lanes past the batch are not defined
        m = np.uint32((1 << (nv % 32)) - 1)
        want[:, -1] &= m
        got[:, -1] &= m
    bad = int((want != got).sum())
    return {"words": int(len(cols)), "rows": int(got.shape[0]), "vectors": int(n_sub), "groups": [int(g) for g in pick], "mismatching_words": bad,
            "ok": bad == 0, "oracle": "oracle.run_batch_sliced (C restatement of simulation.rs:16-30, bit-sliced tier)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="dag1m", choices=list(WORKLOAD_DESC))
    ap.add_argument("--log2-vectors", type=int, default=None, help="vectors per GPU (log2); default = the config's")
    ap.add_argument("--mode", default="auto", choices=["auto", "lane", "coop", "spec"])
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--extras", action="store_true", help="(default at N=1) also time the other BASELINE configs, device-resident")
    ap.add_argument("--no-extras", action="store_true", help="skip the other BASELINE configs")
    ap.add_argument("--no-parity-sample", action="store_true", help="skip the oracle check of sampled output columns")
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3  # timing rule: >= 3 warm-up steps
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    import complogic_b200 as host

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local)
    numa = bind_to_gpu_numa_node(local) if world > 1 else None
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    hbm_peak, sm_max_mhz, peak_kind = load_peaks()
    ctx = host.Context(local)
    mode = {"auto": host.MODE_AUTO, "lane": host.MODE_LANE, "coop": host.MODE_COOP, "spec": host.MODE_SPEC}[args.mode]

    def prepare(workload, lg2):
        circ, lg_default, seed = build_workload(workload)
        lg2 = lg_default if lg2 is None else lg2
        t0 = time.perf_counter()
        cs = host.CudaSimulation.from_simulation(circ.sim, circ.n_immediates, circ.out_regs, ctx=ctx, mode=mode,
                                                 flags=host.LV_STATEFUL if circ.stateful else 0)
        compile_s = time.perf_counter() - t0
        nv = 1 << lg2
        stride = host.round_stride(nv)
        n_in, n_out, n_state = cs.info["n_inputs"], cs.info["n_outputs"], cs.info["n_state"]
        io_bytes = (n_in + n_out + 2 * n_state) * stride * 4  # carried registers are read and written every step
        state = torch.zeros((max(1, n_state), stride), dtype=torch.int32, device=dev) if n_state else None
        # inputs larger than L2, or rotate enough buffer sets that a step never finds its rows in L2
        sets = 1 if n_in * stride * 4 > 2 * L2_BYTES else max(2, int(np.ceil(3 * L2_BYTES / max(1, io_bytes))))
        bufs = []
        for b in range(sets):
            imm = torch.empty((n_in, stride), dtype=torch.int32, device=dev)
            out = torch.zeros((n_out, stride), dtype=torch.int32, device=dev)
            ctx.fill_random(imm.data_ptr(), n_in, stride, seed + 1000 * rank + b, torch.cuda.current_stream().cuda_stream)
            bufs.append((imm, out))
        torch.cuda.synchronize()
        return dict(circ=circ, cs=cs, nv=nv, stride=stride, bufs=bufs, seed=seed, compile_s=compile_s, io_bytes=io_bytes, state=state,
                    gates=cs.info["ref_nand_count"], name=workload)

    def timed_steps(w, steps, warmup, sample_clocks=False):
        cs, bufs, nv, stride = w["cs"], w["bufs"], w["nv"], w["stride"]
        state_ptr = w["state"].data_ptr() if w["state"] is not None else 0
        stream = torch.cuda.current_stream().cuda_stream
        for s in range(warmup):
            imm, out = bufs[s % len(bufs)]
            cs.run_device(imm.data_ptr(), out.data_ptr(), nv, stride, state_dptr=state_ptr, stream=stream)
        barrier()
        sampler = ClockSampler(local) if sample_clocks else None
        if sampler:
            time.sleep(0.25)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.time()
        e0.record()
        for s in range(steps):
            imm, out = bufs[s % len(bufs)]
            cs.run_device(imm.data_ptr(), out.data_ptr(), nv, stride, state_dptr=state_ptr, stream=stream)
        e1.record()
        torch.cuda.synchronize()
        t1 = time.time()
        clocks = sampler.stop(t0, t1) if sampler else None
        barrier()
        ms = max_over_ranks(e0.elapsed_time(e1) / steps)
        return ms, clocks

    w = prepare(args.workload, args.log2_vectors)
    cs = w["cs"]
    ms, clocks = timed_steps(w, args.steps, args.warmup, sample_clocks=True)
    # what the timed launches ran as (the host-buffer e2e leg below works on column chunks and may pick another mode)
    run_mode, run_k, run_launches, run_kernel = cs.mode_name, cs.words_per_thread, cs.launches_per_run, cs.kernel_name
    total_vectors = w["nv"] * world
    value = w["gates"] * total_vectors / (ms * 1e-3)

    # ---- roofline of the dominant (only) kernel -----------------------------------------------------------
    lop3_measured = ctx.measure_lop3_peak() if rank == 0 else 0.0  # lane-ops/s of a pure-LOP3 kernel, this device
    lop3_nominal = 148 * 64 * sm_max_mhz * 1e6
    lane_ops = w["gates"] * w["nv"] / 32.0  # algorithmic: 1 LOP3 lane-op per 32 gate-evals
    issued_lane_ops = cs.info["n_luts"] * w["nv"] / 32.0  # what the kernel really issues after LUT3 fusion
    per_gpu_s = ms * 1e-3
    hbm_achieved = w["io_bytes"] / per_gpu_s / 1e9
    gates_per_byte = w["gates"] * w["nv"] / w["io_bytes"]
    io_bound = gates_per_byte < 91.0 or (hbm_achieved / hbm_peak) > (lane_ops / per_gpu_s / lop3_nominal)
    if io_bound:
        roofline = {"bound": "hbm", "achieved": hbm_achieved, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_achieved / hbm_peak,
                    "traffic": None, "peak_source": f"MEASURED_PEAKS.json ({peak_kind})"}
    else:
        roofline = {"bound": "lop3 (integer ALU pipe; not a tensor-core path)", "achieved": lane_ops / per_gpu_s / 1e12,
                    "peak": lop3_nominal / 1e12, "unit": "T lane-op/s", "frac": lane_ops / per_gpu_s / lop3_nominal,
                    "traffic": None,
                    "peak_source": f"nominal 148 SM x 64 LOP3 lanes/clk x {sm_max_mhz:.0f} MHz; measured pure-LOP3 kernel: "
                                   f"{lop3_measured / 1e12:.2f} T lane-op/s"}
    try:  # measured DRAM traffic of this kernel from the committed ncu capture, scaled to this batch size
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            tr = json.load(f).get(args.workload)
        if tr and tr["kernel"] == run_kernel:
            roofline["traffic"] = tr["dram_bytes"] * (w["nv"] / tr["profiled_vectors"])
            roofline["traffic_source"] = f"ncu capture at {tr['profiled_vectors']} vectors, scaled by batch size"
            if "lsu_data_pipe_pct_of_peak" in tr:  # the pipe that actually bounds the shared-memory interpreter
                roofline["ncu_lsu_data_pipe_frac"] = tr["lsu_data_pipe_pct_of_peak"] / 100.0
                roofline["ncu_alu_pipe_frac"] = tr["alu_pipe_pct_active"] / 100.0
    except Exception:
        pass
    roofline.update({
        "kernel": run_kernel,
        "algorithmic_lane_ops_per_launch": lane_ops, "issued_lop3_lane_ops_per_launch": issued_lane_ops,
        "frac_of_measured_lop3": (lane_ops / per_gpu_s / lop3_measured) if lop3_measured else None,
        "frac_issued_of_nominal_lop3": issued_lane_ops / per_gpu_s / lop3_nominal,
        "hbm_GBps": hbm_achieved, "hbm_frac": hbm_achieved / hbm_peak, "algorithmic_bytes_per_launch": w["io_bytes"],
        "gates_per_io_byte": gates_per_byte,
        # an interpreter that keeps the register stack in shared memory moves (loads + stores) 16-byte slots through a
        # 128 B/clk/SM data path (measured: tools/microbench.cu, profiles/r02_microbench_*.txt): LUTs/clk/SM <= 8 / accesses per LUT
        "smem_interpreter_bound_frac": None if run_mode == "spec" else 1.0 / 6.0,
    })

    # ---- e2e: the C-ABI call with HOST buffers, H2D + kernel + D2H inside the timed region ----------------------
    e2e = None
    if not args.no_e2e:
        n_in, n_out, stride, nv = cs.info["n_inputs"], cs.info["n_outputs"], w["stride"], w["nv"]
        h_imm = torch.empty((n_in, stride), dtype=torch.int32, pin_memory=True)
        h_out = torch.empty((n_out, stride), dtype=torch.int32, pin_memory=True)
        h_imm.copy_(w["bufs"][0][0])
        n_state = cs.info["n_state"]  # a sequential circuit's carried registers travel with every call as well
        h_state = torch.zeros((n_state, stride), dtype=torch.int32, pin_memory=True) if n_state else None
        torch.cuda.synchronize()
        d_imm, d_out = w["bufs"][0]
        stream = torch.cuda.current_stream().cuda_stream
        lib = host._lib.lib()

        def e2e_step():
            # the public host-buffer call: column chunks pipelined over H2D / kernel / D2H streams inside
            host._lib.check(lib.clg_exec_run_host(cs._h, h_imm.data_ptr(), h_out.data_ptr(), h_state.data_ptr() if n_state else None, nv, stride))

        for _ in range(2):
            e2e_step()
        barrier()
        e2e_steps = max(2, min(args.steps, 5))
        t0 = time.perf_counter()  # the call is synchronous (returns with the results on the host): wall clock
        for _ in range(e2e_steps):
            e2e_step()
        torch.cuda.synchronize()
        e_ms = 1e3 * (time.perf_counter() - t0) / e2e_steps
        barrier()
        e_ms = max_over_ranks(e_ms)
        checksum = int(h_out.view(torch.int64).sum().item()) & 0xFFFFFFFFFFFFFFFF  # the result is really on the host
        # ---- what was benched is verified: sampled word columns of the e2e output (first, last, random groups of 4
        # words, every output row) against the oracle VM on the same input columns -- outside the timed region
        parity = None
        if rank == 0 and not args.no_parity_sample and not n_state:
            parity = parity_sample(w["circ"], h_imm.numpy().view(np.uint32), h_out.numpy().view(np.uint32), nv, n_groups=16)
            if not parity["ok"]:
                print(json.dumps({"error": "benched output differs from the oracle VM", "parity_sample": parity}), flush=True)
                raise SystemExit(3)
        # ---- host-transfer ceiling: the same bytes over PCIe with plain copies and NO kernel, all ranks at once
        # (H2D and D2H on their own streams, pinned buffers): what any host-buffer API could reach on this box
        cp_in, cp_out = torch.cuda.Stream(), torch.cuda.Stream()
        def copy_step():
            with torch.cuda.stream(cp_in):
                d_imm.copy_(h_imm, non_blocking=True)
                if n_state:
                    w["state"].copy_(h_state, non_blocking=True)
            with torch.cuda.stream(cp_out):
                h_out.copy_(d_out, non_blocking=True)
                if n_state:
                    h_state.copy_(w["state"], non_blocking=True)
            cp_in.synchronize()
            cp_out.synchronize()
        copy_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            copy_step()
        c_ms = 1e3 * (time.perf_counter() - t0) / e2e_steps
        barrier()
        c_ms = max_over_ranks(c_ms)
        e2e = {"value": w["gates"] * total_vectors / (e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": (n_in + n_state) * stride * 4,
               "d2h_bytes_per_step": (n_out + n_state) * stride * 4, "ms_per_step": e_ms, "steps": e2e_steps,
               "api": "clg_exec_run_host (pinned host buffers; chunked H2D | kernel | D2H pipeline on three streams)",
               "out_checksum": f"{checksum:016x}", "numa_binding": numa, "parity_sample": parity,
               "copy_ceiling": {"ms_per_step": c_ms, "note": "same H2D + D2H bytes with plain pinned-memory copies on two streams, no kernel, "
                                "all ranks concurrently", "GBps_per_gpu": ((n_in + n_out + 2 * n_state) * stride * 4) / (c_ms * 1e-3) / 1e9},
               "frac_of_copy_ceiling": c_ms / e_ms, "kernel_ms": ms, "bound": "kernel" if ms > c_ms else "host copies"}
        del h_imm, h_out

    main_gates, main_compile_s = w["gates"], w["compile_s"]
    l2_note = ("inputs larger than 2x L2 (no flush needed)" if len(w["bufs"]) == 1
               else f"rotating {len(w['bufs'])} buffer sets, together larger than 3x L2")
    def config_record(w, ms_x):
        """One BASELINE config, device-resident: its own roofline (HBM or ISSUED LOP3, whichever is closer to its peak;
        gate-evals stay counted in reference NANDs, which for fused programs exceeds what is issued -- reported as a note)."""
        cs_x = w["cs"]
        nv_total = w["nv"] * world
        hbm = w["io_bytes"] / (ms_x * 1e-3) / 1e9
        issued = cs_x.info["n_luts"] * w["nv"] / 32 / (ms_x * 1e-3) / lop3_nominal
        bound = "hbm" if hbm / hbm_peak >= issued else "lop3-issue"
        rec = {"ms_per_step": ms_x, "value": w["gates"] * nv_total / (ms_x * 1e-3), "unit": UNIT, "vectors_per_gpu": w["nv"], "total_vectors": nv_total,
               "mode": cs_x.mode_name, "kernel": cs_x.kernel_name, "K": cs_x.words_per_thread, "launches_per_step": cs_x.launches_per_run,
               "gates": w["gates"], "n_luts": cs_x.info["n_luts"],
               "roofline": {"bound": bound, "frac": hbm / hbm_peak if bound == "hbm" else issued,
                            "achieved": hbm if bound == "hbm" else issued * lop3_nominal / 1e12, "peak": hbm_peak if bound == "hbm" else lop3_nominal / 1e12,
                            "unit": "GB/s" if bound == "hbm" else "T lane-op/s (LOP3 actually issued)", "hbm_frac": hbm / hbm_peak, "lop3_issued_frac": issued,
                            "note_nand_equivalent_lop3_frac": w["gates"] * w["nv"] / 32 / (ms_x * 1e-3) / lop3_nominal}}
        if cs_x.mode_name != "spec":
            # an interpreter keeps the register stack in shared memory: what bounds it is the LSU data pipe (3-4 slot
            # accesses per LUT through 128 B/clk/SM), not LOP3 issue; the issued-LOP3 fraction is reported for comparison only
            rec["roofline"]["note"] = (f"interpreted ({cs_x.kernel_name}, {cs_x.words_per_thread} word(s) per slot): bound by shared-memory "
                                       "gathers (DESIGN.md 5), not by LOP3 issue or HBM")
        if w["nv"] == 1:
            rec["roofline"]["note"] = "ONE input vector: a latency measurement (a dependent chain per instance), no roofline applies"
        return rec

    extras = None
    strong = None
    if not args.no_extras and (world > 1 or args.extras or args.workload == "dag1m"):
        extras = {}
        if world == 1:
            # every BASELINE config in the driver's line: configs[1] adder32 x 2^24, configs[2] mul16 x 2^26, the structured
            # 1M-gate circuit, configs[4] one vector through 128 flattened multipliers, and configs[3] without its locality window
            plan = [("adder32", None, 20), ("mul16", None, 20), ("mul64x12", None, 5), ("mul32x48", None, 5), ("mul64x128", None, 20), ("dag1m_kinf", 22, 5)]
        else:
            # configs[2] "at 1/2/4/8 B200": 2^26 vectors in total, sharded over the ranks
            # (strong: a 20-microsecond kernel per GPU at N = 8) and, as "mul16_weak", 2^26 vectors per GPU
            plan = [("mul16", 26 - int(np.log2(world)), 20), ("mul16", 26, 20)]
        for name, lg2, steps_x in plan:
            if name == args.workload:
                continue
            del w
            torch.cuda.empty_cache()
            key = name if name not in extras else name + "_weak"
            try:
                w = prepare(name, lg2)
                ms_x, _ = timed_steps(w, steps_x, 3)
                extras[key] = config_record(w, ms_x)
            except Exception as e:  # a config that cannot run must not take the headline down with it
                extras[key] = {"error": str(e)[:200]}
                w = {"bufs": []}
        if world == 1:
            # layout helpers (HBM-bound): vector-major `&[bool]` bytes <-> bit-sliced words, and the `&[bool]`-shaped entry
            # point end to end (host bytes in, host bytes out: transposes + kernel + copies, PCIe moves a byte per bool)
            try:
                del w
                torch.cuda.empty_cache()
                tr = {}
                for nv_t, rows_t in ((1 << 22, 64), (1 << 20, 1024), (1 << 22, 33)):
                    stride_t = host.round_stride(nv_t)
                    bools = torch.randint(0, 2, (nv_t, rows_t), dtype=torch.uint8, device=dev)
                    sliced = torch.zeros((rows_t, stride_t), dtype=torch.int32, device=dev)
                    st = torch.cuda.current_stream().cuda_stream
                    for nm, fn in (("slice", lambda: ctx.slice_bools(bools.data_ptr(), sliced.data_ptr(), nv_t, rows_t, stride_t, st)),
                                   ("unslice", lambda: ctx.unslice_bools(sliced.data_ptr(), bools.data_ptr(), nv_t, rows_t, stride_t, st))):
                        for _ in range(3):
                            fn()
                        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        e0.record()
                        for _ in range(20):
                            fn()
                        e1.record()
                        torch.cuda.synchronize()
                        ms_t = e0.elapsed_time(e1) / 20
                        gbs = (nv_t * rows_t + rows_t * stride_t * 4) / (ms_t * 1e-3) / 1e9
                        tr[f"{nm}_{rows_t}rows"] = {"ms": ms_t, "GBps": gbs, "hbm_frac": gbs / hbm_peak, "vectors": nv_t, "rows": rows_t}
                    del bools, sliced
                extras["transpose"] = dict(tr, roofline={"bound": "hbm", "unit": "GB/s", "peak": hbm_peak,
                                                         "frac": min(v["hbm_frac"] for v in tr.values() if v["rows"] % 16 == 0),
                                                         "note": "frac = the slowest of the four shapes with rows % 16 == 0 (TMA-tiled kernels); the "
                                                                 "*_33rows entries are the staged kernels for any row count"})
                circ_b, _, _ = build_workload("adder32")
                cs_b = host.CudaSimulation.from_simulation(circ_b.sim, circ_b.n_immediates, circ_b.out_regs, ctx=ctx, mode=mode)
                nv_b = 1 << 22
                imm_b = torch.randint(0, 2, (nv_b, 65), dtype=torch.uint8).pin_memory()
                out_b = torch.zeros((nv_b, 33), dtype=torch.uint8).pin_memory()
                lib_b = host._lib.lib()
                def bools_step():
                    host._lib.check(lib_b.clg_exec_run_bools_host(cs_b._h, imm_b.data_ptr(), out_b.data_ptr(), nv_b))
                bools_step()
                t0 = time.perf_counter()
                for _ in range(3):
                    bools_step()
                b_ms = 1e3 * (time.perf_counter() - t0) / 3
                a = imm_b[:1000, :32].numpy().astype(np.uint64)
                b = imm_b[:1000, 32:64].numpy().astype(np.uint64)
                wts = (np.uint64(1) << np.arange(32, dtype=np.uint64))
                want = (a * wts).sum(1) + (b * wts).sum(1) + imm_b[:1000, 64].numpy().astype(np.uint64)
                got = (out_b[:1000].numpy().astype(np.uint64) * (np.uint64(1) << np.arange(33, dtype=np.uint64))).sum(1)
                extras["bools_api"] = {"workload": "adder32 through clg_exec_run_bools_host (one &[bool] per vector, host bytes in and out)",
                                       "vectors": nv_b, "ms_per_call": b_ms, "value": circ_b.sim.nand_count * nv_b / (b_ms * 1e-3), "unit": UNIT,
                                       "host_bytes_per_call": nv_b * (65 + 33), "GBps_over_pcie": nv_b * 98 / (b_ms * 1e-3) / 1e9,
                                       "checked": bool((want == got).all())}
                del imm_b, out_b
                w = {"bufs": []}
            except Exception as e:
                extras["transpose_error"] = str(e)[:200]
                w = {"bufs": []}
        if world > 1 and args.workload == "dag1m":
            # strong scaling of the headline config: 2^24 vectors IN TOTAL, sharded over the ranks (configs[3])
            del w
            torch.cuda.empty_cache()
            w = prepare("dag1m", 24 - int(np.log2(world)))
            ms_s, _ = timed_steps(w, max(3, args.steps), 3)
            strong = {"scaling": "strong", "total_vectors": w["nv"] * world, "vectors_per_gpu": w["nv"], "ms_per_step": ms_s,
                      "value": w["gates"] * w["nv"] * world / (ms_s * 1e-3), "unit": UNIT, "kernel": w["cs"].kernel_name}

    # ---- the product's own multi-GPU call: ONE process (rank 0) drives all N GPUs through clg_sharded_* (a worker
    # thread, context, executor and staging buffers per device, made once; per-GPU result buffers copied back by the
    # host). 2^24 vectors in total, host buffers in and out; the other ranks wait at the barrier with idle GPUs.
    sharded = None
    if world > 1 and args.workload == "dag1m" and not args.no_e2e:
        del w
        try:  # the e2e leg's pinned host buffers (10.7 GB per rank) are no longer needed
            del h_imm, h_out, d_imm, d_out
        except NameError:
            pass
        import gc
        gc.collect()
        torch.cuda.empty_cache()
        barrier()
        if rank == 0:
            try:
                circ_s, _, seed_s = build_workload("dag1m")
                flat_s = host.FlatProgram.levelize(circ_s.sim, circ_s.n_immediates, circ_s.out_regs)
                sh = host.ShardedSimulation(flat_s, list(range(world)), mode)
                nv_s = 1 << 24
                stride_s = host.round_stride(nv_s)
                n_in_s, n_out_s = flat_s.info["n_inputs"], flat_s.info["n_outputs"]
                hs_imm = torch.empty((n_in_s, stride_s), dtype=torch.int32, pin_memory=True)
                hs_out = torch.empty((n_out_s, stride_s), dtype=torch.int32, pin_memory=True)
                d_tmp = torch.empty((n_in_s, stride_s), dtype=torch.int32, device=dev)
                ctx.fill_random(d_tmp.data_ptr(), n_in_s, stride_s, seed_s + 77, torch.cuda.current_stream().cuda_stream)
                hs_imm.copy_(d_tmp)
                torch.cuda.synchronize()
                del d_tmp
                lib_s = host._lib.lib()
                secs = (__import__("ctypes").c_double * world)()

                def sharded_call():
                    host._lib.check(lib_s.clg_sharded_run_host(sh._h, hs_imm.data_ptr(), hs_out.data_ptr(), None, nv_s, stride_s, secs))

                sharded_call()  # (first call: plans, staging buffers)
                t0 = time.perf_counter()
                for _ in range(3):
                    sharded_call()
                s_ms = 1e3 * (time.perf_counter() - t0) / 3
                par_s = parity_sample(circ_s, hs_imm.numpy().view(np.uint32), hs_out.numpy().view(np.uint32), nv_s, n_groups=2 * world)
                sharded = {"api": "clg_sharded_run_host (one process, one worker thread + context + executor per device, no collective)",
                           "devices": world, "total_vectors": nv_s, "ms_per_call": s_ms, "value": circ_s.sim.nand_count * nv_s / (s_ms * 1e-3),
                           "unit": UNIT, "per_device_seconds": [round(x, 4) for x in secs], "host_bytes_per_call": (n_in_s + n_out_s) * stride_s * 4,
                           "parity_sample": {k: par_s[k] for k in ("words", "rows", "mismatching_words", "ok")}}
                sh.close()
                del hs_imm, hs_out
                if not par_s["ok"]:
                    print(json.dumps({"error": "sharded output differs from the oracle VM", "parity_sample": par_s}), flush=True)
                    raise SystemExit(3)
            except SystemExit:
                raise
            except Exception as e:
                sharded = {"error": str(e)[:200]}
        barrier()
        w = {"bufs": []}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import oracle

        oracle.build()
        circ, _, seed = build_workload(args.workload)
        threads = os.cpu_count() or 1
        v, nv_s, dt = cpu_reference_throughput(circ, seed, threads, args.cpu_seconds)
        v1, nv1, dt1 = cpu_reference_throughput(circ, seed, 1, min(3.0, args.cpu_seconds))
        # extra row: the oracle's bit-sliced CPU evaluator (same op list, 64 lanes per u64 word, one thread), so the gain
        # from bit-slicing and the gain from the GPU can be read apart
        from tests.helpers import random_sliced, stride_for
        nv_b = 64 * max(1, int(2e8 // max(1, circ.sim.nand_count)))
        imm_b = random_sliced(circ.n_immediates, stride_for(nv_b), seed)
        tb = time.perf_counter()
        oracle.run_batch_sliced(circ.sim, len(circ.sim.registers), imm_b, nv_b, circ.out_regs)
        tb = time.perf_counter() - tb
        cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
               "bitsliced_1thread_value": circ.sim.nand_count * nv_b / tb, "bitsliced_sample": f"{nv_b} vectors in {tb:.2f} s",
               "sample": f"{nv_s} vectors of {args.workload} in {dt:.1f} s on {threads} threads (scalar VM restating "
                         f"simulation.rs:16-30; 32-byte ops, byte registers, bounds checks)",
               "single_thread_value": v1, "single_thread_sample": f"{nv1} vectors in {dt1:.1f} s"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32", "data": "synthetic",
            "config": {"workload": args.workload, "description": WORKLOAD_DESC[args.workload], "gates": main_gates,
                       "vectors_per_gpu": 1 << (args.log2_vectors or build_workload(args.workload)[1]), "total_vectors": total_vectors,
                       "mode": run_mode, "words_per_thread_or_cta": run_k, "fused_lop3_instructions": cs.info["n_luts"],
                       "levels": cs.info["n_levels"], "slots": cs.info["level_slots"] if run_mode == "coop" else cs.info["lane_slots"],
                       "l2": l2_note, "sharding": "independent vector shards per rank, no collective",
                       "bit_slicing": "one u32 lane word carries 32 circuit instances; NAND over 32 instances is one LOP3",
                       "compile_seconds": main_compile_s},
            "clocks": clocks, "e2e": e2e, "gpu_launches": args.steps * run_launches,
            "roofline": roofline, "cpu_baseline": cpu,
        }
        if sharded:
            line["sharded_single_process"] = sharded
        if extras:
            line["other_configs"] = extras
        if strong:
            line["strong_scaling"] = strong
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
