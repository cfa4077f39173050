#!/usr/bin/env bash
# round 2, GPU call 17: split-K accumulator exchange through L2 vs through DSMEM — parity, per-GEMM times, bench
set -u
mkdir -p gpurun_out/r02
O=gpurun_out/r02
python -m pytest tests/test_gpu_kernels.py tests/test_gpu_network.py tests/test_gpu_bf16.py -m gpu -q -x > $O/pytest_c17.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_c17.log
for m in 3x bf16; do
  for x in l2 dsmem; do
    for w in fwd dx; do
      echo -n "$m $x: "; TNB_GEMM_XCHG=$x python tools/prof_gemm.py $w 1024 2048 2048 200 $m
    done
  done
done 2>&1 | tee $O/xchg_ab.txt
for x in l2 dsmem; do
  TNB_GEMM_XCHG=$x python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-extras > $O/bench_xchg_${x}_3x.json 2> $O/bench_xchg_${x}_3x.err
  TNB_GEMM_XCHG=$x python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-extras --math bf16 > $O/bench_xchg_${x}_bf16.json 2> $O/bench_xchg_${x}_bf16.err
  python - <<P
import json
for m in ("3x","bf16"):
    d=json.load(open("$O/bench_xchg_${x}_%s.json"%m)); print("$x", m, "value %.0f ms %.4f gemm_ms %.4f"%(d["value"], d["ms_per_step"], d["roofline"]["gemm_ms_per_step"]))
P
done
