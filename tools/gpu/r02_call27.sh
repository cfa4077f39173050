#!/usr/bin/env bash
# round 2, GPU call 27 (1 GPU): whole GPU suite, smoke, default bench line, launch list of one bunch, ncu --set full of the split-K forward / dX kernels (L2 exchange)
set -u
mkdir -p gpurun_out/r02/ncu2
O=gpurun_out/r02
python -m pytest tests -m gpu -q -rxXs > $O/pytest_gpu_c27.log 2>&1; echo "pytest rc=$?"; tail -5 $O/pytest_gpu_c27.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_c27.log 2>&1; echo "smoke rc=$?"; tail -1 $O/smoke_c27.log
python bench.py > $O/bench_n1_c27.json 2> $O/bench_n1_c27.err; echo "bench rc=$?"; python -c "
import json;d=json.load(open('$O/bench_n1_c27.json'));print('value %.0f ms %.4f e2e %.0f frac %.4f bf16 %.0f ms %.4f frac %.4f cpu %s'%(d['value'],d['ms_per_step'],d['e2e']['value'],d['roofline']['frac'],d['bf16']['value'],d['bf16']['ms_per_step'],d['bf16']['roofline']['frac'],d.get('cpu_baseline',{}).get('value')))"
python bench.py --steps 2 --warmup 3 --windows 1 --no-extras --no-cpu-baseline > $O/bench_small_c27.json 2> $O/bench_small_c27.err &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file $O/launches_c27.csv python bench.py --steps 2 --warmup 3 --windows 1 --no-extras --no-cpu-baseline > $O/ncu_launches_c27.log 2>&1
echo "ncu launches rc=$?"
cap() { name=$1; rx=$2; skip=$3; cnt=$4; shift 4
  "$@" > $O/ncu2/$name.plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c $cnt -o $O/ncu2/$name "$@" > $O/ncu2/$name.ncu.log 2>&1
  echo "$name rc=$?"
}
cap fwd_3x gemm_tcgen05 3 1 python tools/prof_gemm.py fwd 1024 2048 2048 6 3x
cap dx_3x gemm_tcgen05 3 1 python tools/prof_gemm.py dx 1024 2048 2048 6 3x
