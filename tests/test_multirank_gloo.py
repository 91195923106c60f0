"""The N>1 path on the CPU: world_size-2 `gloo` processes shard a batch exactly as the multi-GPU drivers do
(clg_shard_range: contiguous 4-word-aligned ranges, no data-path collective), evaluate their shard, and the
gathered result must equal the oracle on the whole batch. Also checks bench.py's max-over-ranks timing rule."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import complogic_b200 as host
from complogic_b200 import circuits
from oracle import oracle
from tests import flat_interp
from tests.helpers import mask_tail, random_sliced, stride_for


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, nv, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        circ = circuits.array_multiplier(8)
        stride = stride_for(nv)
        imm = random_sliced(16, stride, seed=99)  # every rank regenerates the same batch, uses only its shard
        fp = host.FlatProgram.levelize(circ.sim, 16, circ.out_regs)
        w0, w1 = host.shard_range(nv, world, rank)
        flat = flat_interp.Flat(fp.to_bytes())
        sw = max(4, (w1 - w0 + 3) // 4 * 4)
        shard = np.zeros((16, sw), np.uint32)
        shard[:, :w1 - w0] = imm[:, w0:w1]
        out = flat_interp.run(flat, flat_interp.STREAM_LEVEL, shard)[:, :w1 - w0]
        full = torch.zeros((16, stride), dtype=torch.int64)
        full[:, w0:w1] = torch.from_numpy(out.astype(np.int64))
        gathered = [torch.zeros_like(full) for _ in range(world)]
        dist.all_gather(gathered, full)  # test plumbing only: the product never exchanges data between shards
        t = torch.tensor([1.0 + rank], dtype=torch.float64)  # pretend per-rank step time
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            total = sum(gathered).numpy().astype(np.uint32)
            want = oracle.run_batch_sliced(circ.sim, len(circ.sim.registers), imm, nv, circ.out_regs)
            q.put((bool(np.array_equal(mask_tail(total, nv), mask_tail(want, nv))), float(t.item()), (w0, w1)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("nv", [1000, 4096 + 37])
def test_two_ranks_shard_without_exchange(nv):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, nv, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    ok, tmax, _ = q.get(timeout=10)
    assert ok and tmax == 2.0  # value = units of all ranks / MAX over ranks of the step time
