#!/usr/bin/env bash
# round 2, GPU call 19: pipeline timeline of the fused update GEMM (tracing build), both math modes; plain wall of the three ops
set -u
mkdir -p gpurun_out/r02
O=gpurun_out/r02
python tools/prof_gemm.py all 1024 2048 2048 200 3x
export TNB_LIB_DIR=$PWD/nnet-asr_b200/lib_trace
for m in 3x bf16; do
  echo "=== $m upd 2048x2048x1024"; DBG_OP=upd python tools/dbg_timeline.py $m T N 2048 2048 1024
  echo "=== $m grid upd-shaped TN gemm"; DBG_MATH=$m python tools/dbg_grid.py T N 2048 2048 1024
done > $O/timeline_c19.txt 2>&1
cat $O/timeline_c19.txt
