#!/usr/bin/env bash
# round 2, GPU call 29 (2 GPUs): the default bench line exactly as the driver launches it (extras on), and the reference arm's rank handling
set -u
export MASTER_ADDR=127.0.0.1
N=${1:-2}
mkdir -p gpurun_out/r02
O=gpurun_out/r02
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29591 bench.py --gpus $N --steps 50 --warmup 5 > $O/bench_n${N}_final.json 2> $O/bench_n${N}_final.err; echo "bench rc=$?"
python -c "
import json;d=json.load(open('$O/bench_n${N}_final.json'))
print('N=$N value %.0f ms %.4f e2e %.0f' % (d['value'], d['ms_per_step'], d['e2e']['value']))
for k in ('bf16','strong'):
    if k in d: print(k, {x:d[k][x] for x in d[k] if x in ('value','ms_per_step','global_bunch','bunch_per_gpu')})
" || tail -5 $O/bench_n${N}_final.err
TNB_DP_PUSH=2 DP_EQUIV_MODES=peer timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29577 tools/dp_equivalence.py > $O/dp_equiv_final.log 2>&1; echo "dp_equivalence rc=$? $(grep -c DP_EQUIV_OK $O/dp_equiv_final.log)"; grep "dp ok" $O/dp_equiv_final.log | head
