#!/usr/bin/env bash
set -u
export MASTER_ADDR=127.0.0.1
mkdir -p gpurun_out/r02
O=gpurun_out/r02
nvidia-smi -L
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_cli.py -m gpu -q -x -rxXs -k "two_gpu or multi or data_parallel" > $O/pytest_2gpu.log 2>&1; echo "pytest rc=$?"; tail -8 $O/pytest_2gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 2 --steps 50 --warmup 5 > $O/bench_n2.json 2> $O/bench_n2.err; echo "bench rc=$?"; cat $O/bench_n2.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('weak value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'])
for k in ('strong_scaling','bf16'):
    if k in d: print(k,d[k]['value'],d[k]['ms_per_step'],d[k].get('bunch_per_gpu'))
"
tail -3 $O/bench_n2.err
