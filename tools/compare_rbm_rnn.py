#!/usr/bin/env python
"""BASELINE configs[3] and [4] at full size, reference GPU binaries (oracle/_ref/TRbmCu, TRecurrentCu) next to the drop-ins on
the same files and flags: RBM CD-1 (Gaussian-Bernoulli 429 -> 2048, bunch 128) and the simple recurrent layer (351 + 1024 ->
1024 -> 135, BPTT 20).  Prints each binary's report line and wall time of the whole process."""
import importlib.util, os, re, subprocess, sys, tempfile, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "nnet-asr_b200", "python"))
spec = importlib.util.spec_from_file_location("make_golden", os.path.join(ROOT, "tests", "golden", "make_golden.py"))
MG = importlib.util.module_from_spec(spec); spec.loader.exec_module(MG)
REF, BIN = os.path.join(ROOT, "oracle", "_ref"), os.path.join(ROOT, "nnet-asr_b200", "bin")
which = sys.argv[1] if len(sys.argv) > 1 else "both"

def timed(fn, *a, **k):
    t = time.perf_counter(); r = fn(*a, **k); return r, time.perf_counter() - t

if which in ("rbm", "both"):
    cfg = dict(raw_dim=39, ctx=5, nhid=2048, vistype="gauss", hidtype="bern", n_utt=int(os.environ.get("RBM_UTTS", "200")), n_frames=1000, bunch=128, cache=16384,
               lr=0.001, mmt=0.5, wc=2e-4, seed=11)
    out = {}
    for tag, exe in (("reference TRbmCu", os.path.join(REF, "TRbmCu")), ("this repo bin/TRbmCu", os.path.join(BIN, "TRbmCu"))):
        with tempfile.TemporaryDirectory() as d:
            (rep, LF, txt), wall = timed(MG.run_rbm, "rbm_full", cfg, d, exe=exe, save=False)
        out[tag] = (rep, wall, LF)
        fin = re.search(r"FINISHED \(\s*([0-9.eE+-]+)s", txt)
        loop = float(fin.group(1)) if fin else float("nan")
        print("RBM 429->2048 CD-1  %-22s %s  process wall %.1f s, training loop %.2f s -> %.0f frames/s" % (tag, rep, wall, loop, rep["frames"] / loop), flush=True)
    a, b = out["reference TRbmCu"], out["this repo bin/TRbmCu"]
    print("  relative difference of the final weights: %.2e (max |dW| / max |W|), Mse ratio %.6f, speed-up %.1fx" % (
        np.abs(a[2][3] - b[2][3]).max() / np.abs(a[2][3]).max(), b[0]["err"] / a[0]["err"], a[1] / b[1]), flush=True)
if which in ("rnn", "both"):
    cfg = dict(raw_dim=39, ctx=4, nhid=1024, n_out=135, n_utt=int(os.environ.get("RNN_UTTS", "8")), n_frames=500, bptt=20, lr=0.0005, seed=21)
    out = {}
    for tag, exe in (("reference TRecurrentCu", os.path.join(REF, "TRecurrentCu")), ("this repo bin/TRecurrentCu", os.path.join(BIN, "TRecurrentCu"))):
        with tempfile.TemporaryDirectory() as d:
            (rep, LF, txt), wall = timed(MG.run_rnn, "rnn_full", cfg, d, exe=exe, save=False)
        out[tag] = (rep, wall, LF)
        fin = re.search(r"FINISHED \(\s*([0-9.eE+-]+)s", txt)
        loop = float(fin.group(1)) if fin else float("nan")
        print("RNN 351+1024->1024->135 BPTT 20  %-26s %s  process wall %.1f s, training loop %.2f s -> %.0f frames/s" % (tag, rep, wall, loop, rep["frames"] / loop), flush=True)
    a, b = out["reference TRecurrentCu"], out["this repo bin/TRecurrentCu"]
    print("  relative difference of the final recurrent weights: %.2e, Xent ratio %.6f, speed-up %.1fx" % (
        np.abs(a[2][0][1] - b[2][0][1]).max() / np.abs(a[2][0][1]).max(), b[0]["err"] / a[0]["err"], a[1] / b[1]), flush=True)
