#!/bin/bash
# tile shapes of the TMA-tiled transposes (CLG_TRANSPOSE_TMA = index into {32x1024, 64x1024, 64x512, 128x512, 64x256,
# 128x256, 32x512}; default: 64x512, 32x1024 below 48 rows) against the LDG/STG kernels (CLG_TRANSPOSE_NO_TMA)
echo "== default"; python tools/bench_transpose.py 2>&1
for t in ${TMA_SHAPES:-0 6 4}; do echo "== CLG_TRANSPOSE_TMA=$t"; CLG_TRANSPOSE_TMA=$t python tools/bench_transpose.py 2>&1 | grep -v "rows=   17\|rows=   33"; done
echo "== CLG_TRANSPOSE_NO_TMA=1"; CLG_TRANSPOSE_NO_TMA=1 python tools/bench_transpose.py 2>&1 | grep -v "rows=   17\|rows=   33"
