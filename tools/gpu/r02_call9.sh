#!/usr/bin/env bash
set -u
mkdir -p gpurun_out/r02
O=gpurun_out/r02
timeout 900 python -m pytest tests -m gpu -q -x -rxXs > $O/pytest4.log 2>&1; echo "pytest rc=$?"; tail -8 $O/pytest4.log
timeout 200 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-extras > $O/bench4_3x.json 2> $O/bench4_3x.err; python -c "import json;d=json.load(open('$O/bench4_3x.json'));print('3x ms/step',d['ms_per_step'],'launches/step',d['gpu_launches']/50,'e2e',d['e2e']['value'])"
timeout 200 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-extras --math bf16 > $O/bench4_bf16.json 2> $O/bench4_bf16.err; python -c "import json;d=json.load(open('$O/bench4_bf16.json'));print('bf16 ms/step',d['ms_per_step'],'launches/step',d['gpu_launches']/50)"
