"""complogic_b200 -- B200-native backend for complogic's NAND-VM hot path."""
from .api import (And, Compiler, Context, CudaSimulation, DLatch, FlatProgram, FourBitAdder, FullAdder, Gate, HalfAdder,
                  Incrementer, Nand, Nand_op, Nor, Not, Op, Or, RSLatch, RSLatchTest, Set_op, Simulation, Xor,
                  array_to_ops, ops_to_array, round_stride, run_sharded, shard_range, ShardedSimulation, OP_DTYPE)
from ._lib import (ClgError, ReferencePanic, BackendMissing, MODE_AUTO, MODE_COOP, MODE_LANE, MODE_SPEC, MODE_NAMES,
                   LV_STATEFUL, LV_NO_FUSE, LV_NO_LANE_STREAM, LV_NO_PARTS, LV_NO_ROWS)
