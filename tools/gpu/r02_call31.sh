#!/usr/bin/env bash
# round 2, GPU call 31 (2 GPUs): where does the end-to-end loop lose 0.12 ms per bunch with several ranks? (timing experiments)
set -u
export MASTER_ADDR=127.0.0.1
N=${1:-2}
mkdir -p gpurun_out/r02
O=gpurun_out/r02
run() { name=$1; shift
  env "$@" timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29591 bench.py --gpus $N --steps 50 --warmup 5 --windows 1 --no-extras > $O/bench_n${N}_$name.json 2> $O/bench_n${N}_$name.err
  python -c "import json;d=json.load(open('$O/bench_n${N}_$name.json'));print('N=$N $name ms %.4f e2e ms %.4f host %s'%(d['ms_per_step'],1000*d['e2e']['windows_s'][0]/d['steps'],d['e2e']['host_ms_per_step']))" || tail -3 $O/bench_n${N}_$name.err
}
run e2e_base
run e2e_noh2d TNH_E2E_DEBUG=1
run e2e_nod2h TNH_E2E_DEBUG=2
run e2e_none TNH_E2E_DEBUG=3
run e2e_copystats TNH_E2E_DEBUG=4
