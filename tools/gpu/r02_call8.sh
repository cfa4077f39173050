#!/usr/bin/env bash
set -u
mkdir -p gpurun_out/r02
O=gpurun_out/r02
run() { name=$1; shift
  env "$@" timeout 120 python bench.py --steps 50 --warmup 5 --windows 3 --no-cpu-baseline --no-extras $EXTRA > $O/pol_$name.json 2> $O/pol_$name.err
  python -c "import json;d=json.load(open('$O/pol_$name.json'));print('$name ms/step %.4f gemm %.4f launches/step %d'%(d['ms_per_step'],d['roofline']['gemm_ms_per_step'],d['gpu_launches']/50))"
}
EXTRA=""
run 3x_pol2_1 TNB_BATCH_POLICY=2 TNB_BATCH_FLUSH_TILES=1
run 3x_pol2_1_p64 TNB_BATCH_POLICY=2 TNB_BATCH_FLUSH_TILES=1 TNB_BATCH_PAIRS=64
EXTRA="--math bf16"
run bf_pol2_1 TNB_BATCH_POLICY=2 TNB_BATCH_FLUSH_TILES=1
run bf_pol1_p64 TNB_BATCH_POLICY=1 TNB_BATCH_PAIRS=64
