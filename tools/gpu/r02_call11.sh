#!/usr/bin/env bash
set -u
mkdir -p gpurun_out/r02
O=gpurun_out/r02
timeout 600 python -m pytest tests/test_gpu_network.py tests/test_gpu_cli.py tests/test_gpu_kernels.py -m gpu -q -x -k "rbm or Rbm or trbm or gemm" > $O/pytest5.log 2>&1; echo "pytest rc=$?"; tail -5 $O/pytest5.log
for c in A B D E; do
  timeout 300 python bench.py --config $c --steps 50 --warmup 5 --no-extras > $O/bench_cfg$c.json 2> $O/bench_cfg$c.err; echo "cfg $c rc=$?"
  python -c "import json;d=json.load(open('$O/bench_cfg$c.json'));print('$c value %.0f frames/s  ms/step %.4f  roof %.4f  cpu %s  e2e %.0f'%(d['value'],d['ms_per_step'],d['roofline']['frac'],d.get('cpu_baseline',{}).get('value'),d['e2e']['value']))" || tail -5 $O/bench_cfg$c.err
done
