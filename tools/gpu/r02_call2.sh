#!/usr/bin/env bash
# round 2, GPU call 2: first run of the batched persistent GEMM kernel (tnb_gemm_batch)
set -u
mkdir -p gpurun_out/r02
O=gpurun_out/r02
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "gemm_batch" > $O/pytest_batch.log 2>&1; echo "batch rc=$?"; tail -15 $O/pytest_batch.log
timeout 600 python -m pytest tests/test_gpu_network.py -m gpu -q -x -k "config_c" > $O/pytest_config_c.log 2>&1; echo "config_c rc=$?"; tail -15 $O/pytest_config_c.log
timeout 900 python -m pytest tests -m gpu -q -rxX > $O/pytest_gpu2.log 2>&1; echo "pytest rc=$?"; tail -6 $O/pytest_gpu2.log
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-extras > $O/bench2_3x.json 2> $O/bench2_3x.err; cat $O/bench2_3x.json
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-extras --math bf16 > $O/bench2_bf16.json 2> $O/bench2_bf16.err; cat $O/bench2_bf16.json
TNB_GEMM_BATCH=0 timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-extras > $O/bench2_3x_nobatch.json 2> $O/bench2_3x_nobatch.err; cat $O/bench2_3x_nobatch.json
