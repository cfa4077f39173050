#!/usr/bin/env bash
# round 2, GPU call 21 (N GPUs): copy-engine gradient push (TNB_DP_PUSH=2) against the GEMM-epilogue push (1) and the pull (0):
# equivalence tests first, then the bench with the peer-kernel timeline
set -u
export MASTER_ADDR=127.0.0.1
N=${1:-4}
mkdir -p gpurun_out/r02
O=gpurun_out/r02
python -m pytest tests/test_gpu_multi.py tests/test_gpu_cli.py -m gpu -q -x -k "two_gpu or two_gpus" > $O/pytest_multi_c21.log 2>&1; echo "pytest multi rc=$?"; tail -3 $O/pytest_multi_c21.log
for p in 1 0; do
  TNB_DP_PUSH=$p DP_EQUIV_MODES=peer timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29577 tools/dp_equivalence.py > $O/dp_equiv_push$p.log 2>&1; echo "dp_equivalence push=$p world=$N rc=$? $(grep -c DP_EQUIV_OK $O/dp_equiv_push$p.log)"
done
run() { name=$1; shift
  env "$@" timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29591 bench.py --gpus $N --steps 50 --warmup 5 --windows 3 --no-extras > $O/bench_n${N}_$name.json 2> $O/bench_n${N}_$name.err
  python -c "import json;d=json.load(open('$O/bench_n${N}_$name.json'));print('N=$N $name value %.0f ms %.4f e2e %.0f gemm_ms %.4f xent %.12f'%(d['value'],d['ms_per_step'],d['e2e']['value'],d['roofline']['gemm_ms_per_step'],d['final_stats']['xent_per_frame']))" || tail -3 $O/bench_n${N}_$name.err
}
run push2 TNB_DP_PUSH=2 TNB_DP_TRACE=1
grep "dp trace\] r0" $O/bench_n${N}_push2.err | tail -15
run push1 TNB_DP_PUSH=1
run push0 TNB_DP_PUSH=0
run push2_defer16 TNB_DP_PUSH=2 TNB_DP_DEFER=1:6
run push2_defer00 TNB_DP_PUSH=2 TNB_DP_DEFER=0:0
