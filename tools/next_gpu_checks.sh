#!/usr/bin/env bash
# What a new GPU session should run first.  Each block writes into gpurun_out/next/.
#
#   1 GPU :  gpurun --timeout 900 -- 'bash tools/next_gpu_checks.sh one'
#   N GPUs:  gpurun --gpus N --timeout 600 -- 'bash tools/next_gpu_checks.sh many N'      (N = 2, 4 or 8)
set -u
O=gpurun_out/next
mkdir -p $O
export MASTER_ADDR=127.0.0.1
case "${1:-one}" in
  one)
    # the whole GPU suite, the smoke entry point, the default bench line (3xTF32 headline + bf16 sub-object + CPU baseline), and
    # the launch list of one bunch under ncu (only after the same command has run cleanly without it)
    python -m pytest tests -m gpu -q -rxXs > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest_gpu.log
    python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/smoke.log
    python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"
    python bench.py --steps 2 --warmup 3 --windows 1 --no-extras --no-cpu-baseline > $O/bench_small.json 2> $O/bench_small.err &&
    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv \
        --log-file $O/launches.csv python bench.py --steps 2 --warmup 3 --windows 1 --no-extras --no-cpu-baseline > $O/ncu_launches.log 2>&1
    ;;
  many)
    N=${2:-2}
    # N-rank equivalence (every schedule, both math modes), the 2-GPU tests of the suite, the default bench line as the driver launches
    # it, and the same with the data-parallel step's timeline on stderr (TNB_DP_TRACE=1: peer-memory kernels and copy-engine pushes)
    python -m pytest tests/test_gpu_multi.py tests/test_gpu_cli.py -m gpu -q -k "two_gpu or two_gpus" > $O/pytest_multi.log 2>&1; tail -2 $O/pytest_multi.log
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus $N --steps 50 --warmup 5 \
        > $O/bench_n$N.json 2> $O/bench_n$N.err; echo "bench rc=$?"
    TNB_DP_TRACE=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29572 bench.py --gpus $N --steps 50 \
        --warmup 5 --windows 3 --no-extras > $O/bench_n${N}_trace.json 2> $O/bench_n${N}_trace.err; grep "dp trace\] r0" $O/bench_n${N}_trace.err | tail -24
    ;;
esac
