#!/usr/bin/env bash
set -u
mkdir -p gpurun_out/r02
O=gpurun_out/r02
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_network.py tests/test_gpu_bf16.py -m gpu -q -x > $O/pytest6.log 2>&1; echo "pytest rc=$?"; tail -5 $O/pytest6.log
timeout 200 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-extras > $O/bench5_3x.json 2> $O/bench5_3x.err; python -c "import json;d=json.load(open('$O/bench5_3x.json'));print('3x ms/step',d['ms_per_step'],'launches/step',d['gpu_launches']/50)"
timeout 120 python bench.py --steps 3 --warmup 3 --windows 1 --no-cpu-baseline --no-extras > $O/bench_plain_short2.json 2> $O/bench_plain_short2.err &&
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $O/launches_r02_3x.csv python bench.py --steps 3 --warmup 3 --windows 1 --no-cpu-baseline --no-extras > $O/ncu_bench2.log 2>&1
echo "ncu rc=$?"
