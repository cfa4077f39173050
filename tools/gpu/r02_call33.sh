#!/usr/bin/env bash
# round 2, GPU call 33 (N GPUs): the default bench line exactly as the driver launches it (extras on)
set -u
export MASTER_ADDR=127.0.0.1
N=${1:-4}
mkdir -p gpurun_out/r02
O=gpurun_out/r02
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29591 bench.py --gpus $N --steps 50 --warmup 5 > $O/bench_n${N}_final.json 2> $O/bench_n${N}_final.err; echo "bench rc=$?"
python -c "
import json;d=json.load(open('$O/bench_n${N}_final.json'))
print('N=$N value %.0f ms %.4f e2e %.0f' % (d['value'], d['ms_per_step'], d['e2e']['value']))
for k in ('bf16','strong_scaling'):
    if k in d: print(k, {x:d[k][x] for x in d[k] if x in ('value','ms_per_step','global_bunch','bunch_per_gpu')})
" || tail -5 $O/bench_n${N}_final.err
