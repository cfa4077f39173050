#!/usr/bin/env bash
# What the next GPU session should run first, as three gpurun calls.  Each block writes into gpurun_out/.
#
#   1 GPU :  gpurun --timeout 600 -- 'bash tools/next_gpu_checks.sh one'
#   2 GPUs:  gpurun --gpus 2 --timeout 600 -- 'bash tools/next_gpu_checks.sh two'
#   4 GPUs:  gpurun --gpus 4 --timeout 400 -- 'bash tools/next_gpu_checks.sh four'
set -u
mkdir -p gpurun_out
case "${1:-one}" in
  one)
    # GPU goldens that need the reference TNetCu (learning-rate factors), then the whole GPU suite incl. the tests added after
    # round 1's last full run, then an ncu capture of the peer-memory kernel driven by virtual ranks on one device
    python tests/golden/make_golden.py --impl gpu --only opt_ --out gpurun_out/golden > gpurun_out/make_golden_opt.log 2>&1
    cp gpurun_out/golden/*.npz tests/golden/ 2>/dev/null
    python -m pytest tests -m gpu -q -rxX > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
    tail -5 gpurun_out/pytest_gpu.log
    ncu --set full --clock-control none --import-source on -k regex:dp_peer_update_kernel -c 8 -o gpurun_out/peer_virtual \
        python -m pytest tests/test_gpu_kernels.py -q -k "peer and 100-260-8" > gpurun_out/ncu_peer.log 2>&1
    python bench.py --steps 50 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
    ;;
  two)
    export MASTER_ADDR=127.0.0.1
    python -m pytest tests/test_gpu_multi.py -m gpu -q > gpurun_out/pytest_multi.log 2>&1; tail -3 gpurun_out/pytest_multi.log
    python tools/symm_probe.py > /dev/null 2>&1 || true
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 tools/symm_probe.py \
        > gpurun_out/symm_probe_n2.txt 2>&1
    python tools/dp_sweep.py --gpus 2 --modes peer,allreduce --ctas 12,20,32 --out gpurun_out/dp_sweep_n2.jsonl | tee gpurun_out/dp_sweep_n2.txt
    python tools/dp_sweep.py --gpus 2 --math bf16 --modes peer,allreduce --ctas 20 --out gpurun_out/dp_sweep_n2_bf16.jsonl | tee gpurun_out/dp_sweep_n2_bf16.txt
    ;;
  four)
    export MASTER_ADDR=127.0.0.1
    DP_EQUIV_MODES=peer python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29572 \
        tools/dp_equivalence.py > gpurun_out/dp_equiv_n4.log 2>&1; grep -E "dp ok|DP_EQUIV" gpurun_out/dp_equiv_n4.log
    python tools/dp_sweep.py --gpus 4 --modes peer,allreduce --ctas 20 --out gpurun_out/dp_sweep_n4.jsonl | tee gpurun_out/dp_sweep_n4.txt
    ;;
esac
