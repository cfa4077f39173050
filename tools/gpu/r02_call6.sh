#!/usr/bin/env bash
set -u
mkdir -p gpurun_out/r02
O=gpurun_out/r02
timeout 120 python bench.py --steps 3 --warmup 3 --windows 1 --no-cpu-baseline --no-extras > $O/bench_plain_short.json 2> $O/bench_plain_short.err &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_batch_3x.csv python bench.py --steps 3 --warmup 3 --windows 1 --no-cpu-baseline --no-extras > $O/ncu_bench.log 2>&1
echo "ncu rc=$?"
TNB_GEMM_DEBUG=1 timeout 120 python bench.py --steps 1 --warmup 3 --windows 1 --no-cpu-baseline --no-extras 2>&1 | grep "gemm batch" | sort | uniq -c | head -20
