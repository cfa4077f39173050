#!/usr/bin/env bash
# round 2, GPU call 16: the whole GPU suite (new full-size D/E drop-in tests and the virtual-rank whole-step test), default bench line
set -u
mkdir -p gpurun_out/r02
O=gpurun_out/r02
python -m pytest tests -m gpu -q -rxXs > $O/pytest_gpu_c16.log 2>&1; echo "pytest rc=$?"; tail -6 $O/pytest_gpu_c16.log
python bench.py > $O/bench_n1_c16.json 2> $O/bench_n1_c16.err; echo "bench rc=$?"; cat $O/bench_n1_c16.json
