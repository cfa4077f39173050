#!/usr/bin/env bash
# round 2: ncu --set full captures of the dominant kernels (each only after the same command has run cleanly without ncu)
set -u
mkdir -p gpurun_out/r02/ncu
O=gpurun_out/r02/ncu
cap() { # name, kernel regex, skip, count, cmd...
  name=$1; rx=$2; skip=$3; cnt=$4; shift 4
  "$@" > $O/$name.plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c $cnt -o $O/$name "$@" > $O/$name.ncu.log 2>&1
  echo "$name rc=$?"
}
cap fwd_3x gemm_tcgen05 3 1 python tools/prof_gemm.py fwd 1024 2048 2048 6 3x
cap upd_3x gemm_tcgen05 3 1 python tools/prof_gemm.py upd 1024 2048 2048 6 3x
cap dx_3x gemm_tcgen05 3 1 python tools/prof_gemm.py dx 1024 2048 2048 6 3x
cap fwd_bf16 gemm_tcgen05 3 1 python tools/prof_gemm.py fwd 1024 2048 2048 6 bf16
MATH=3xtf32 MODE=layer cap batch_3x gemm_multi 3 1 python tools/dbg/batch_timeline.py
MATH=bf16 MODE=layer cap batch_bf16 gemm_multi 3 1 python tools/dbg/batch_timeline.py
PEER_BENCH_WORLDS=8 PEER_BENCH_ITERS=1 cap peer_virtual dp_peer_update_virtual 2 1 python tools/peer_virtual_bench.py
ls -la $O
