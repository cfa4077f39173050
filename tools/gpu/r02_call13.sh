#!/usr/bin/env bash
set -u
export MASTER_ADDR=127.0.0.1
mkdir -p gpurun_out/r02
O=gpurun_out/r02
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "scatter or peer or objective" > $O/pytest7.log 2>&1; echo "pytest rc=$?"; tail -5 $O/pytest7.log
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_cli.py -m gpu -q -x -rxXs -k "two_gpu or multi" > $O/pytest_2gpu_b.log 2>&1; echo "pytest 2gpu rc=$?"; tail -5 $O/pytest_2gpu_b.log
for push in 1 0; do
TNB_DP_PUSH=$push timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2957$push bench.py --gpus 2 --steps 50 --warmup 5 --no-extras > $O/bench_n2_push$push.json 2> $O/bench_n2_push$push.err; echo "bench push=$push rc=$?"
python -c "import json;d=json.load(open('$O/bench_n2_push$push.json'));print('push=$push weak value %.0f ms %.4f e2e %.0f'%(d['value'],d['ms_per_step'],d['e2e']['value']))"
done
