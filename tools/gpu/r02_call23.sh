#!/usr/bin/env bash
# round 2, GPU call 23 (N GPUs): copy-engine pushes on one stream per destination vs one stream; timeline
set -u
export MASTER_ADDR=127.0.0.1
N=${1:-4}
mkdir -p gpurun_out/r02
O=gpurun_out/r02
run() { name=$1; shift
  env "$@" timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29591 bench.py --gpus $N --steps 50 --warmup 5 --windows 3 --no-extras > $O/bench_n${N}_$name.json 2> $O/bench_n${N}_$name.err
  python -c "import json;d=json.load(open('$O/bench_n${N}_$name.json'));print('N=$N $name value %.0f ms %.4f e2e %.0f gemm_ms %.4f xent %.12f'%(d['value'],d['ms_per_step'],d['e2e']['value'],d['roofline']['gemm_ms_per_step'],d['final_stats']['xent_per_frame']))" || tail -3 $O/bench_n${N}_$name.err
}
run fan_t TNB_DP_TRACE=1
grep "dp trace\] r0" $O/bench_n${N}_fan_t.err | tail -16
run fan
run nofan TNB_DP_PUSH_STREAMS=0
run push1 TNB_DP_PUSH=1
