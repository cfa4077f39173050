#!/usr/bin/env bash
set -u
mkdir -p gpurun_out/r02
O=gpurun_out/r02
MATH=3xtf32 MODE=upd timeout 40 python tools/dbg/batch_timeline.py > $O/timeline3_3xtf32_upd.txt 2>&1; rc=$?; echo "== 3xtf32 upd rc=$rc"; cat $O/timeline3_3xtf32_upd.txt | cut -c1-330
if [ $rc -ne 0 ]; then echo "3xTF32 batch kernel failed: stopping"; exit 1; fi
for m in 3xtf32 bf16; do MATH=$m MODE=layer timeout 40 python tools/dbg/batch_timeline.py > $O/timeline3_${m}_layer.txt 2>&1; echo "== $m layer rc=$?"; cat $O/timeline3_${m}_layer.txt | cut -c1-330; done
timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_bf16.py tests/test_gpu_network.py -m gpu -q -x > $O/pytest3.log 2>&1; echo "pytest rc=$?"; tail -5 $O/pytest3.log
timeout 120 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-extras > $O/bench3_3x.json 2> $O/bench3_3x.err; python -c "import json;d=json.load(open('$O/bench3_3x.json'));print('3x ms/step',d['ms_per_step'],d['roofline']['gemm_ms_per_step'],d['e2e']['host_ms_per_step'])"
timeout 120 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-extras --math bf16 > $O/bench3_bf16.json 2> $O/bench3_bf16.err; python -c "import json;d=json.load(open('$O/bench3_bf16.json'));print('bf16 ms/step',d['ms_per_step'],d['roofline']['gemm_ms_per_step'],d['e2e']['host_ms_per_step'])"
