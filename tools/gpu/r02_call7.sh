#!/usr/bin/env bash
set -u
mkdir -p gpurun_out/r02
O=gpurun_out/r02
run() { # name, env...
  name=$1; shift
  env "$@" timeout 120 python bench.py --steps 50 --warmup 5 --windows 3 --no-cpu-baseline --no-extras $EXTRA > $O/pol_$name.json 2> $O/pol_$name.err
  python -c "import json;d=json.load(open('$O/pol_$name.json'));print('$name ms/step %.4f gemm %.4f launches/step %d'%(d['ms_per_step'],d['roofline']['gemm_ms_per_step'],d['gpu_launches']/50))"
}
EXTRA=""
run 3x_off TNB_GEMM_BATCH=0
run 3x_pol1 TNB_BATCH_POLICY=1
run 3x_pol2_148 TNB_BATCH_POLICY=2 TNB_BATCH_FLUSH_TILES=148
run 3x_pol2_96 TNB_BATCH_POLICY=2 TNB_BATCH_FLUSH_TILES=96
run 3x_pol2_74 TNB_BATCH_POLICY=2 TNB_BATCH_FLUSH_TILES=74
run 3x_pol2_220 TNB_BATCH_POLICY=2 TNB_BATCH_FLUSH_TILES=220
EXTRA="--math bf16"
run bf_off TNB_GEMM_BATCH=0
run bf_pol1 TNB_BATCH_POLICY=1
run bf_pol2_148 TNB_BATCH_POLICY=2 TNB_BATCH_FLUSH_TILES=148
run bf_pol2_74 TNB_BATCH_POLICY=2 TNB_BATCH_FLUSH_TILES=74
run bf_pol2_220 TNB_BATCH_POLICY=2 TNB_BATCH_FLUSH_TILES=220
