#!/usr/bin/env bash
set -u
mkdir -p gpurun_out/r02
O=gpurun_out/r02
for m in 3xtf32 bf16; do for mode in layer upd; do MATH=$m MODE=$mode timeout 120 python tools/dbg/batch_timeline.py > $O/timeline2_${m}_${mode}.txt 2>&1; echo "== $m $mode rc=$?"; cat $O/timeline2_${m}_${mode}.txt | cut -c1-330; done; done
timeout 120 python tools/dbg/batch_vs_unbatched.py > $O/batch_vs_unbatched2.txt 2>&1; echo "rc=$?"; grep -E "math|W 0|eout 1" $O/batch_vs_unbatched2.txt
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_bf16.py tests/test_gpu_network.py -m gpu -q -x > $O/pytest3.log 2>&1; echo "pytest rc=$?"; tail -5 $O/pytest3.log
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-extras > $O/bench3_3x.json 2> $O/bench3_3x.err; python -c "import json;d=json.load(open('$O/bench3_3x.json'));print('3x ms/step',d['ms_per_step'],d['roofline']['gemm_ms_per_step'])"
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-extras --math bf16 > $O/bench3_bf16.json 2> $O/bench3_bf16.err; python -c "import json;d=json.load(open('$O/bench3_bf16.json'));print('bf16 ms/step',d['ms_per_step'],d['roofline']['gemm_ms_per_step'])"
