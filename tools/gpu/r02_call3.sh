#!/usr/bin/env bash
set -u
mkdir -p gpurun_out/r02
O=gpurun_out/r02
for m in 3xtf32 bf16; do for mode in layer upd dx; do MATH=$m MODE=$mode timeout 120 python tools/dbg/batch_timeline.py > $O/timeline_${m}_${mode}.txt 2>&1; echo "== $m $mode rc=$?"; cat $O/timeline_${m}_${mode}.txt; done; done
timeout 120 python tools/dbg/batch_vs_unbatched.py > $O/batch_vs_unbatched.txt 2>&1; echo "rc=$?"; cat $O/batch_vs_unbatched.txt
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "gemm_batch" > $O/pytest_batch.log 2>&1; echo "batch rc=$?"; tail -5 $O/pytest_batch.log
