#!/usr/bin/env bash
# round 2, GPU call 1: pending goldens, baseline of the GPU suite / bench, measured TF32 peak, ncu of the HBM kernels and of the peer kernel
set -u
mkdir -p gpurun_out/r02
O=gpurun_out/r02
python tests/golden/make_golden.py --impl gpu --only opt_ --out $O/golden > $O/make_golden_opt.log 2>&1; echo "golden rc=$?"
cp $O/golden/*.npz tests/golden/ 2>/dev/null
python -m pytest tests -m gpu -q -rxX > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest_gpu.log
python tools/tf32_peak.py > $O/tf32_peak.json 2> $O/tf32_peak.err; cat $O/tf32_peak.json
python bench.py --steps 50 --warmup 5 --no-cpu-baseline > $O/bench_n1_base.json 2> $O/bench_n1_base.err; cat $O/bench_n1_base.json
python bench.py --steps 50 --warmup 5 --no-cpu-baseline --math bf16 > $O/bench_n1_bf16_base.json 2> $O/bench_n1_bf16_base.err; cat $O/bench_n1_bf16_base.json
TNB_BENCH_ITERS_SCALE=0.1 python tools/bench_kernels.py > $O/hbm_kernels_plain.txt 2>&1 &&
TNB_BENCH_ITERS_SCALE=0.1 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,sm__warps_active.avg.pct_of_peak_sustained_active \
  --clock-control none --csv --log-file $O/hbm_kernels_ncu.csv python tools/bench_kernels.py > $O/hbm_kernels_ncu.log 2>&1
echo "ncu hbm rc=$?"
python tools/peer_virtual_bench.py > $O/peer_virtual_bench.txt 2>&1 &&
PEER_BENCH_WORLDS=8 PEER_BENCH_ITERS=1 ncu --set full --clock-control none --import-source on -k regex:dp_peer_update_virtual -s 2 -c 1 -o $O/peer_virtual \
    python tools/peer_virtual_bench.py > $O/ncu_peer.log 2>&1
echo "ncu peer rc=$?"; cat $O/peer_virtual_bench.txt
