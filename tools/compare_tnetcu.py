#!/usr/bin/env python
"""Same files, same flags, same B200: the unmodified reference GPU trainer (oracle/_ref/TNetCu: CuBaseLib + legacy cuBLAS SGEMM,
compiled for sm_100 by oracle/build_ref.sh) next to the drop-in nnet-asr_b200/bin/TNetCu.  Prints each binary's own report and
[FPS] line (TNetCu.cc:471: frames / wall time of the training loop, file reading included) and the ratio.
usage: compare_tnetcu.py [config]   config = A (351-1024-135, bunch 256) | B (351-2048-135) | C (429-2048x6-3000, bunch 1024)"""
import os, re, subprocess, sys, tempfile, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "nnet-asr_b200", "python"))
from tnet_b200 import formats as F

CFG = {
    "A": dict(ctx=4, hidden=[1024], n_out=135, bunch=256, cache=12800, n_utt=400, n_frames=1000, lr=0.008, mmt=0.0),
    "B": dict(ctx=4, hidden=[2048], n_out=135, bunch=256, cache=16384, n_utt=400, n_frames=1000, lr=0.008, mmt=0.5),
    "C": dict(ctx=5, hidden=[2048] * 6, n_out=3000, bunch=1024, cache=131072, n_utt=768, n_frames=1024, lr=0.008, mmt=0.5),
}
name = sys.argv[1] if len(sys.argv) > 1 else "C"
c = CFG[name]
rng = np.random.default_rng(20240607)
d = tempfile.mkdtemp(prefix="tnetcu_cmp_")
utts = F.gen_utterances(c["n_utt"], c["n_frames"], 39, c["n_out"], rng, vary_len=False)
paths = F.write_dataset(d, utts, c["n_out"], c["ctx"])
dims = [39 * (2 * c["ctx"] + 1)] + c["hidden"] + [c["n_out"]]
init = os.path.join(d, "init.nnet")
with open(init, "w") as f:
    for l in range(len(dims) - 1):
        nin, nout = dims[l], dims[l + 1]
        last = l == len(dims) - 2
        f.write("<biasedlinearity> %d %d\nm %d %d\n" % (nout, nin, nout, nin)); f.flush()
        (0.1 * rng.standard_normal(nout * nin)).astype(np.float32).tofile(f, sep=" ", format="%.6g")
        f.write("\nv %d\n" % nout); f.flush()
        (np.zeros(nout, np.float32) if last else (rng.random(nout) / 5.0 - 4.1).astype(np.float32)).tofile(f, sep=" ", format="%.6g")
        f.write("\n<%s> %d %d\n" % ("softmax" if last else "sigmoid", nout, nout))
frames = c["n_utt"] * c["n_frames"]
print("config %s: dims %s, bunch %d, %d frames" % (name, dims, c["bunch"], frames), flush=True)
res = {}
for tag, exe in (("reference TNetCu (cuBLAS fp32)", os.path.join(ROOT, "oracle", "_ref", "TNetCu")), ("this repo bin/TNetCu (3xTF32)", os.path.join(ROOT, "nnet-asr_b200", "bin", "TNetCu")),
                 ("this repo bin/TNetCu --MATH=bf16", os.path.join(ROOT, "nnet-asr_b200", "bin", "TNetCu"))):
    cmd = [exe, "-H", init, "-I", paths["mlf"], "-L", "*/", "-X", "lab", "-S", paths["scp"], "-m", paths["labelmap"], "-n", repr(c["lr"]),
           "--TARGETMMF=" + os.path.join(d, "out.nnet"), "--BUNCHSIZE=%d" % c["bunch"], "--CACHESIZE=%d" % c["cache"], "--RANDOMIZE=TRUE", "--SEED=123",
           "--FEATURETRANSFORM=" + paths["transform"], "--STARTFRMEXT=%d" % c["ctx"], "--ENDFRMEXT=%d" % c["ctx"], "--MOMENTUM=%g" % c["mmt"]]
    if "bf16" in tag:
        cmd.append("--MATH=bf16")
    if "this repo" in tag and os.environ.get("TNETCU_EXTRA"):
        cmd += os.environ["TNETCU_EXTRA"].split()
    t = time.perf_counter()
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    wall = time.perf_counter() - t
    if r.returncode != 0:
        print(tag, "FAILED:\n", r.stdout[-1500:]); continue
    fps = re.search(r"\[\s*FPS:\s*([0-9.eE+]+)", r.stdout)
    rep = re.search(r"Xent:\S+ frames:\d+ err/frm:\S+ correct\[\S+%\]", r.stdout)
    res[tag] = float(fps.group(1)) if fps else float("nan")
    print("%-34s FPS %-10s %s   (process wall %.1f s incl. reading/writing the text network)" % (tag, fps.group(1) if fps else "?", rep.group(0) if rep else "", wall), flush=True)
    if os.environ.get("TNETCU_EXTRA") and "this repo" in tag:
        print(r.stdout[-600:])
k = list(res)
for t in k[1:]:
    print("%s / %s = %.1fx" % (t, k[0], res[t] / res[k[0]]))
