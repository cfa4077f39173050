#!/usr/bin/env python
"""Sweep the data-parallel schedule knobs with bench.py on N GPUs of this box and print one table line per setting.

  python tools/dp_sweep.py --gpus 2                         # default grid: schedules x TNB_DP_PEER_CTAS x TNB_DP_DEFER
  python tools/dp_sweep.py --gpus 8 --math bf16 --ctas 20 --defer 1:5 --modes peer,allreduce

Every setting is one `torchrun bench.py --gpus N --no-cpu-baseline` run (about 25 s each, mostly start-up), so keep the grid small
on a metered box.  The table goes to stdout and, with --out, the raw bench lines to a JSON-lines file (profiles/ material)."""
import argparse
import itertools
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=2)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--math", default="3xtf32")
    ap.add_argument("--modes", default="peer,allreduce", help="comma list of TNB_DP_MODE values")
    ap.add_argument("--ctas", default="12,20,32", help="comma list of TNB_DP_PEER_CTAS (peer schedule only)")
    ap.add_argument("--defer", default="1:5", help="comma list of TNB_DP_DEFER begin:end windows")
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    rows = []
    port = 29600
    for mode, defer in itertools.product(a.modes.split(","), a.defer.split(",")):
        for ctas in (a.ctas.split(",") if mode == "peer" else [""]):
            env = dict(os.environ, MASTER_ADDR="127.0.0.1", TNB_DP_MODE=mode, TNB_DP_DEFER=defer)
            if ctas:
                env["TNB_DP_PEER_CTAS"] = ctas
            port += 1
            cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(a.gpus), "--master-addr", "127.0.0.1",
                   "--master-port", str(port), os.path.join(ROOT, "bench.py"), "--gpus", str(a.gpus), "--steps", str(a.steps), "--warmup",
                   str(a.warmup), "--math", a.math, "--no-cpu-baseline"]
            r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env, timeout=600)
            line = [l for l in r.stdout.splitlines() if l.startswith("{")]
            if r.returncode != 0 or not line:
                print("%-10s defer %-5s ctas %-3s FAILED rc=%d\n%s" % (mode, defer, ctas, r.returncode, r.stderr[-1500:]), flush=True)
                continue
            d = json.loads(line[-1])
            d["sweep"] = {"mode": mode, "defer": defer, "peer_ctas": ctas}
            rows.append(d)
            print("%-10s defer %-5s ctas %-3s  %.3f ms/bunch  %.0f frames/s  e2e %.0f  xent/frame %.9f" %
                  (mode, defer, ctas, d["ms_per_step"], d["value"], d["e2e"]["value"], d["final_stats"]["xent_per_frame"]), flush=True)
    if a.out:
        with open(a.out, "w") as f:
            for d in rows:
                f.write(json.dumps(d) + "\n")


if __name__ == "__main__":
    main()
