#!/usr/bin/env bash
# round 2, GPU call 18: pipeline timelines (tracing build) of the split-K forward kernel after the L2 exchange, both math modes
set -u
mkdir -p gpurun_out/r02
O=gpurun_out/r02
export TNB_LIB_DIR=$PWD/nnet-asr_b200/lib_trace
for m in 3x bf16; do
  echo "=== $m fwd 1024x2048x2048"; DBG_OP=fwd DBG_SPLIT=1 python tools/dbg_timeline.py $m N N 1024 2048 2048
  echo "=== $m dx"; DBG_OP=dx DBG_SPLIT=1 python tools/dbg_timeline.py $m N T 1024 2048 2048
  echo "=== $m grid"; DBG_MATH=$m python tools/dbg_grid.py N N 1024 2048 2048
done > $O/timeline_c18.txt 2>&1
cat $O/timeline_c18.txt
