#!/usr/bin/env python
"""Generate golden vectors by running the UNMODIFIED reference binaries built by
oracle/build_ref.sh (oracle/_ref/{TNet,TNetCu,TRbmCu,TRecurrentCu}) on small seeded synthetic sets.

  python tests/golden/make_golden.py --impl cpu            # here (no GPU): reference CPU trainer TNet
  python tests/golden/make_golden.py --impl gpu            # on a B200 via gpurun: reference TNetCu/TRbmCu/TRecurrentCu

Each case is stored as tests/golden/<impl>_<case>.npz: the inputs (features, labels, initial network,
flags) and the reference's outputs (final network as written to its 6-digit text file, Xent/Mse,
frames, correct).  tests/test_oracle_golden.py replays the inputs through oracle/tnet_oracle.c and
compares; the -m gpu parity tests replay them through the CUDA path.

Does not read /root/reference at run time (only the prebuilt binaries in oracle/_ref).
"""
import argparse
import os
import re
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "nnet-asr_b200", "python"))
from tnet_b200 import formats as F  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref")
OUT = os.path.join(ROOT, "tests", "golden")

# name -> dict(raw_dim, ctx, hidden[], n_out, n_utt, n_frames, bunch, cache, lr, mmt, wc, gdf, seed, randomize)
MLP_CASES = {
    # CPU-comparable update rule: TNetCu --GRADDIVFRM=F --MOMENTUM=0  ==  CPU TNet  
    "mlp_small": dict(raw_dim=13, ctx=2, hidden=[32], n_out=10, n_utt=24, n_frames=90, bunch=64, cache=512,
                      lr=0.01, mmt=0.0, wc=0.0, gdf=False, seed=123, randomize=True),
    "mlp_norand_wc": dict(raw_dim=13, ctx=2, hidden=[24, 24], n_out=12, n_utt=16, n_frames=80, bunch=32, cache=256,
                          lr=0.005, mmt=0.0, wc=1e-4, gdf=False, seed=7, randomize=False),
    # GPU-only rule: momentum + 1/N + L2; wide output (>256 -> serial softmax/check_class/colsum paths)
    "mlp_mmt_wide": dict(raw_dim=13, ctx=3, hidden=[48], n_out=300, n_utt=24, n_frames=120, bunch=128, cache=1024,
                         lr=0.5, mmt=0.5, wc=1e-5, gdf=True, seed=99, randomize=True),
    # bunch > 512 rows (AddColSum serial path for every layer)
    "mlp_bigbunch": dict(raw_dim=13, ctx=2, hidden=[40], n_out=20, n_utt=30, n_frames=140, bunch=640, cache=1920,
                         lr=1.0, mmt=0.9, wc=0.0, gdf=True, seed=5, randomize=True),
}
GPU_ONLY = {"mlp_mmt_wide", "mlp_bigbunch"}
# option cases (fixtures *_opt_*): per-layer learning-rate factors with a frozen first layer (the backpropagation stopper moves up,
# cuNetwork.cc:80-135) and cross-validation (forward + objective only, nothing written, TNetCu.cc:437-466 / TNet.cc:344)
OPT_CASES = {
    "opt_lrfactors": dict(raw_dim=13, ctx=2, hidden=[24, 20], n_out=12, n_utt=16, n_frames=80, bunch=32, cache=256,
                          lr=0.01, mmt=0.0, wc=0.0, gdf=False, seed=71, randomize=True, factors="0:1:0.5"),
    "opt_cv": dict(raw_dim=13, ctx=2, hidden=[24], n_out=12, n_utt=16, n_frames=80, bunch=32, cache=256,
                   lr=0.01, mmt=0.0, wc=0.0, gdf=False, seed=72, randomize=False, cv=True),
}


def _b(v):
    return "TRUE" if v else "FALSE"


def parse_report(txt):
    m = re.search(r"Xent:(\S+) frames:(\d+) err/frm:(\S+) correct\[(\S+)%\]", txt)
    if m:
        return dict(err=float(m.group(1)), frames=int(m.group(2)), correct_pct=float(m.group(4)))
    m = re.search(r"Mse:(\S+) frames:(\d+) err/frm:(\S+)", txt)
    if m:
        return dict(err=float(m.group(1)), frames=int(m.group(2)), correct_pct=float("nan"))
    raise RuntimeError("no report line in:\n" + txt[-2000:])


def pack_layers(prefix, layers, out):
    k = 0
    for L in layers:
        if L[0] == "affine":
            out["%s_Wt%d" % (prefix, k)] = L[1]
            out["%s_b%d" % (prefix, k)] = L[2]
            k += 1


def run_mlp(case, cfg, impl, workdir, exe=None, save=True, extra=None):
    rng = np.random.default_rng(cfg["seed"] + 1000)
    utts = F.gen_utterances(cfg["n_utt"], cfg["n_frames"], cfg["raw_dim"], cfg["n_out"], rng)
    paths = F.write_dataset(workdir, utts, cfg["n_out"], cfg["ctx"])
    dims = [cfg["raw_dim"] * (2 * cfg["ctx"] + 1)] + cfg["hidden"] + [cfg["n_out"]]
    layers = F.gen_mlp_init(dims, rng)
    init = os.path.join(workdir, "init.nnet")
    F.write_mlp(init, layers)
    layers = F.read_mlp(init)  # exactly what the binaries parse
    final = os.path.join(workdir, "final.nnet")
    if exe is None:
        exe = os.path.join(REF, "TNet" if impl == "cpu" else "TNetCu")
    cmd = [exe, "-H", init, "-I", paths["mlf"], "-L", "*/", "-X", "lab", "-S", paths["scp"], "-m", paths["labelmap"],
           "-n", repr(cfg["lr"]), "--TARGETMMF=" + final, "--BUNCHSIZE=%d" % cfg["bunch"], "--CACHESIZE=%d" % cfg["cache"],
           "--RANDOMIZE=" + _b(cfg["randomize"]), "--SEED=%d" % cfg["seed"], "--FEATURETRANSFORM=" + paths["transform"],
           "--STARTFRMEXT=%d" % cfg["ctx"], "--ENDFRMEXT=%d" % cfg["ctx"], "--WEIGHTCOST=" + repr(cfg["wc"])]
    if impl == "cpu":
        cmd += ["--THREADS=1"]
    else:
        cmd += ["--MOMENTUM=" + repr(cfg["mmt"]), "--GRADDIVFRM=" + _b(cfg["gdf"])]
    if cfg.get("factors"):
        cmd += ["--LEARNRATEFACTORS=" + cfg["factors"]]
    if cfg.get("cv"):
        cmd += ["--CROSSVALIDATE=TRUE"]
    if extra:
        cmd += list(extra)      # flags of the drop-in that the reference does not have (--GPUS, --LOADER, ...)
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        raise RuntimeError("reference failed:\n" + res.stdout[-3000:])
    rep = parse_report(res.stdout)
    if cfg.get("cv"):
        assert not os.path.exists(final), "a cross-validation run must not write a network"
        out_layers = layers          # nothing is trained: the fixture's "final" network is the initial one
    else:
        out_layers = F.read_mlp(final)
    if not save:
        return rep, out_layers, res.stdout
    names = list(utts.keys())
    data = dict(
        lr_factors=np.array([float(v) for v in cfg["factors"].split(":")] if cfg.get("factors") else [], dtype=np.float64),
        cv=np.int64(1 if cfg.get("cv") else 0),
        feats=np.concatenate([utts[n][0] for n in names]), labels=np.concatenate([utts[n][1] for n in names]),
        lengths=np.array([utts[n][0].shape[0] for n in names], dtype=np.int32),
        dims=np.array(dims, dtype=np.int32),
        cfg=np.array([cfg["ctx"], cfg["bunch"], cfg["cache"], cfg["seed"], int(cfg["randomize"]), int(cfg["gdf"])], dtype=np.int64),
        hyper=np.array([cfg["lr"], cfg["mmt"], cfg["wc"]], dtype=np.float64),
        ref_err=np.float64(rep["err"]), ref_frames=np.int64(rep["frames"]), ref_correct_pct=np.float64(rep["correct_pct"]),
    )
    pack_layers("init", layers, data)
    pack_layers("final", out_layers, data)
    np.savez_compressed(os.path.join(OUT, "%s_%s.npz" % (impl, case)), **data)
    print("%s %s: %s" % (impl, case, rep))


# --------------------------------------------------------------------------- networks with the OffsetGemm layers
# first layer: ("shared", bo) = <sharedlinearity> with one instance per spliced frame (2*ctx+1 instances of raw_dim -> bo), or
# ("discrete", [(in_i, out_i), ...]) = <discretelinearity>; then <sigmoid>, <biasedlinearity> to n_out, <softmax>.
NET_CASES = {
    # CPU-comparable rule (TNetLib/SharedLinearity.cc:222-258: W -= lr/K * dW, no momentum, no 1/N, no L2)
    "net_shared": dict(raw_dim=13, ctx=2, first=("shared", 8), n_out=10, n_utt=20, n_frames=90, bunch=64, cache=512,
                       lr=0.02, mmt=0.0, wc=0.0, gdf=False, seed=61, randomize=True),
    "net_shared_mmt": dict(raw_dim=13, ctx=3, first=("shared", 6), n_out=12, n_utt=20, n_frames=100, bunch=128, cache=768,
                           lr=0.8, mmt=0.5, wc=1e-4, gdf=True, seed=62, randomize=True),
    "net_discrete": dict(raw_dim=13, ctx=2, first=("discrete", [(26, 12), (39, 20)]), n_out=10, n_utt=20, n_frames=100, bunch=64,
                         cache=512, lr=0.6, mmt=0.5, wc=1e-4, gdf=True, seed=63, randomize=True),
}
NET_GPU_ONLY = {"net_shared_mmt", "net_discrete"}
LAST_FINAL_TEXT = None


def build_net(cfg, rng):
    nin = cfg["raw_dim"] * (2 * cfg["ctx"] + 1)
    kind, arg = cfg["first"]
    if kind == "shared":
        K = 2 * cfg["ctx"] + 1
        first = ("shared", K, (0.1 * rng.standard_normal((arg, nin // K))).astype(np.float32), rng.uniform(-4.1, -3.9, arg).astype(np.float32))
        nhid = arg * K
    else:
        assert sum(b[0] for b in arg) == nin
        nhid = sum(b[1] for b in arg)
        first = ("discrete", [(0.1 * rng.standard_normal((o, i))).astype(np.float32) for i, o in arg], rng.uniform(-4.1, -3.9, nhid).astype(np.float32))
    top = ("affine", (0.1 * rng.standard_normal((cfg["n_out"], nhid))).astype(np.float32), np.zeros(cfg["n_out"], np.float32))
    return [first, ("sigmoid", nhid), top, ("softmax", cfg["n_out"])]


def run_net(case, cfg, impl, workdir, exe=None, save=True):
    """Like run_mlp for an arbitrary layer list; the fixture keeps the initial and final network FILES as text."""
    rng = np.random.default_rng(cfg["seed"] + 1000)
    utts = F.gen_utterances(cfg["n_utt"], cfg["n_frames"], cfg["raw_dim"], cfg["n_out"], rng)
    paths = F.write_dataset(workdir, utts, cfg["n_out"], cfg["ctx"])
    init = os.path.join(workdir, "init.nnet")
    F.write_mlp(init, build_net(cfg, rng))
    final = os.path.join(workdir, "final.nnet")
    if exe is None:
        exe = os.path.join(REF, "TNet" if impl == "cpu" else "TNetCu")
    cmd = [exe, "-H", init, "-I", paths["mlf"], "-L", "*/", "-X", "lab", "-S", paths["scp"], "-m", paths["labelmap"],
           "-n", repr(cfg["lr"]), "--TARGETMMF=" + final, "--BUNCHSIZE=%d" % cfg["bunch"], "--CACHESIZE=%d" % cfg["cache"],
           "--RANDOMIZE=" + _b(cfg["randomize"]), "--SEED=%d" % cfg["seed"], "--FEATURETRANSFORM=" + paths["transform"],
           "--STARTFRMEXT=%d" % cfg["ctx"], "--ENDFRMEXT=%d" % cfg["ctx"], "--WEIGHTCOST=" + repr(cfg["wc"])]
    if impl == "cpu":
        cmd += ["--THREADS=1"]
    else:
        cmd += ["--MOMENTUM=" + repr(cfg["mmt"]), "--GRADDIVFRM=" + _b(cfg["gdf"])]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        raise RuntimeError("reference failed:\n" + res.stdout[-3000:])
    rep = parse_report(res.stdout)
    global LAST_FINAL_TEXT
    LAST_FINAL_TEXT = open(final).read()       # the written network file as text (tests compare its layout with the reference's)
    if not save:
        return rep, F.read_mlp(final), res.stdout
    names = list(utts.keys())
    np.savez_compressed(
        os.path.join(OUT, "%s_%s.npz" % (impl, case)),
        feats=np.concatenate([utts[n][0] for n in names]), labels=np.concatenate([utts[n][1] for n in names]),
        lengths=np.array([utts[n][0].shape[0] for n in names], dtype=np.int32),
        dims=np.array([cfg["raw_dim"] * (2 * cfg["ctx"] + 1), cfg["n_out"]], dtype=np.int32),
        cfg=np.array([cfg["ctx"], cfg["bunch"], cfg["cache"], cfg["seed"], int(cfg["randomize"]), int(cfg["gdf"])], dtype=np.int64),
        hyper=np.array([cfg["lr"], cfg["mmt"], cfg["wc"]], dtype=np.float64),
        init_net=np.frombuffer(open(init, "rb").read(), dtype=np.uint8), final_net=np.frombuffer(open(final, "rb").read(), dtype=np.uint8),
        ref_err=np.float64(rep["err"]), ref_frames=np.int64(rep["frames"]), ref_correct_pct=np.float64(rep["correct_pct"]))
    print("%s %s: %s" % (impl, case, rep))


# --------------------------------------------------------------------------- RBM (TRbmCu) / recurrent (TRecurrentCu)
RBM_CASES = {
    "rbm_gb": dict(raw_dim=13, ctx=1, nhid=32, vistype="gauss", hidtype="bern", n_utt=12, n_frames=100, bunch=32, cache=256,
                   lr=0.001, mmt=0.5, wc=2e-4, seed=11),
    "rbm_bb": dict(raw_dim=13, ctx=1, nhid=24, vistype="bern", hidtype="bern", n_utt=12, n_frames=100, bunch=32, cache=256,
                   lr=0.1, mmt=0.5, wc=2e-4, seed=12),
    # <rbmsparse> (cuRbmSparse.cc): the file carries the sparsity cost; large enough here for the penalty to move the weights
    "rbm_sparse_bb": dict(raw_dim=13, ctx=1, nhid=24, vistype="bern", hidtype="bern", n_utt=12, n_frames=100, bunch=32, cache=256,
                          lr=0.1, mmt=0.5, wc=2e-4, seed=13, sparse_cost=0.01),
}
RNN_CASES = {
    "rnn_small": dict(raw_dim=13, ctx=1, nhid=20, n_out=8, n_utt=6, n_frames=40, bptt=4, lr=0.05, seed=21),
}


def run_rbm(case, cfg, workdir, exe=None, save=True):
    rng = np.random.default_rng(cfg["seed"] + 1000)
    utts = F.gen_utterances(cfg["n_utt"], cfg["n_frames"], cfg["raw_dim"], 4, rng)
    if cfg["vistype"] == "bern":  # visible probabilities in (0,1)
        utts = {k: ((1.0 / (1.0 + np.exp(-v[0]))).astype(np.float32), v[1]) for k, v in utts.items()}
    paths = F.write_dataset(workdir, utts, 4, cfg["ctx"])
    nvis = cfg["raw_dim"] * (2 * cfg["ctx"] + 1)
    # tools/init/gen_rbm_init.py:39-76 : W ~ 0.1*N(0,1) (scaled), biases 0 / small
    Wt = (0.1 * rng.standard_normal((cfg["nhid"], nvis))).astype(np.float32)
    vb = np.zeros(nvis, np.float32)
    hb = (rng.random(cfg["nhid"]) / 5.0 - 0.1).astype(np.float32)
    init = os.path.join(workdir, "init.rbm")
    cost = cfg.get("sparse_cost")
    if cost is None:
        F.write_mlp(init, [("rbm", cfg["vistype"], cfg["hidtype"], Wt, vb, hb)])
    else:
        F.write_mlp(init, [("rbmsparse", cfg["vistype"], cfg["hidtype"], Wt, vb, hb, cost)])
    L0 = F.read_mlp(init)[0]
    final = os.path.join(workdir, "final.rbm")
    cmd = [exe or os.path.join(REF, "TRbmCu"), "-H", init, "-S", paths["scp"], "-n", repr(cfg["lr"]), "--TARGETMMF=" + final,
           "--BUNCHSIZE=%d" % cfg["bunch"], "--CACHESIZE=%d" % cfg["cache"], "--RANDOMIZE=TRUE", "--SEED=%d" % cfg["seed"],
           "--FEATURETRANSFORM=" + paths["transform"], "--STARTFRMEXT=%d" % cfg["ctx"], "--ENDFRMEXT=%d" % cfg["ctx"],
           "--MOMENTUM=" + repr(cfg["mmt"]), "--WEIGHTCOST=" + repr(cfg["wc"])]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        raise RuntimeError("reference failed:\n" + res.stdout[-3000:])
    rep = parse_report(res.stdout)
    LF = F.read_mlp(final)[0]
    if not save:
        return rep, LF, res.stdout
    names = list(utts.keys())
    np.savez_compressed(
        os.path.join(OUT, "gpu_%s.npz" % case),
        feats=np.concatenate([utts[n][0] for n in names]), lengths=np.array([utts[n][0].shape[0] for n in names], np.int32),
        cfg=np.array([cfg["ctx"], cfg["bunch"], cfg["cache"], cfg["seed"], int(cfg["vistype"] == "gauss"), int(cfg["hidtype"] == "gauss")], np.int64),
        hyper=np.array([cfg["lr"], cfg["mmt"], cfg["wc"]], np.float64),
        init_Wt=L0[3], init_vb=L0[4], init_hb=L0[5], final_Wt=LF[3], final_vb=LF[4], final_hb=LF[5],
        sparse_cost=np.float64(-1.0 if cost is None else cost),
        ref_err=np.float64(rep["err"]), ref_frames=np.int64(rep["frames"]))
    print("gpu %s: %s" % (case, rep))


def run_rnn(case, cfg, workdir, exe=None, save=True):
    rng = np.random.default_rng(cfg["seed"] + 1000)
    utts = F.gen_utterances(cfg["n_utt"], cfg["n_frames"], cfg["raw_dim"], cfg["n_out"], rng)
    paths = F.write_dataset(workdir, utts, cfg["n_out"], cfg["ctx"])
    nin = cfg["raw_dim"] * (2 * cfg["ctx"] + 1)
    H = cfg["nhid"]
    # tools/init/gen_recurrent_init.py:31-47 style: small gaussian weights
    Wr = (0.1 * rng.standard_normal((H, nin + H))).astype(np.float32)
    br = np.zeros(H, np.float32)
    Wo = (0.1 * rng.standard_normal((cfg["n_out"], H))).astype(np.float32)
    bo = np.zeros(cfg["n_out"], np.float32)
    init = os.path.join(workdir, "init.rnn")
    F.write_mlp(init, [("recurrent", Wr, br, nin), ("affine", Wo, bo), ("softmax", cfg["n_out"])])
    L = F.read_mlp(init)
    final = os.path.join(workdir, "final.rnn")
    cmd = [exe or os.path.join(REF, "TRecurrentCu"), "-H", init, "-I", paths["mlf"], "-L", "*/", "-X", "lab", "-S", paths["scp"],
           "-m", paths["labelmap"], "-n", repr(cfg["lr"]), "--TARGETMMF=" + final, "--BPTT=%d" % cfg["bptt"],
           "--FEATURETRANSFORM=" + paths["transform"], "--STARTFRMEXT=%d" % cfg["ctx"], "--ENDFRMEXT=%d" % cfg["ctx"]]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        raise RuntimeError("reference failed:\n" + res.stdout[-3000:])
    rep = parse_report(res.stdout)
    LF = F.read_mlp(final)
    if not save:
        return rep, LF, res.stdout
    names = list(utts.keys())
    np.savez_compressed(
        os.path.join(OUT, "gpu_%s.npz" % case),
        feats=np.concatenate([utts[n][0] for n in names]), labels=np.concatenate([utts[n][1] for n in names]),
        lengths=np.array([utts[n][0].shape[0] for n in names], np.int32),
        cfg=np.array([cfg["ctx"], cfg["bptt"], nin, H, cfg["n_out"]], np.int64), hyper=np.array([cfg["lr"]], np.float64),
        init_Wr=L[0][1], init_br=L[0][2], init_Wo=L[1][1], init_bo=L[1][2],
        final_Wr=LF[0][1], final_br=LF[0][2], final_Wo=LF[1][1], final_bo=LF[1][2],
        ref_err=np.float64(rep["err"]), ref_frames=np.int64(rep["frames"]), ref_correct_pct=np.float64(rep["correct_pct"]))
    print("gpu %s: %s" % (case, rep))


# forward-only tool: name -> dict(raw_dim, ctx, hidden[], n_out, n_utt, n_frames, seed, log)
FEACAT_CASES = {
    "feacat_post": dict(raw_dim=13, ctx=3, hidden=[48, 40], n_out=300, n_utt=6, n_frames=70, seed=31, log=False),
    "feacat_logpost": dict(raw_dim=13, ctx=2, hidden=[32], n_out=20, n_utt=5, n_frames=50, seed=32, log=True),
}


def run_feacat(case, cfg, workdir, exe=None, save=True):
    """reference CPU TFeaCat (src/TFeaCat.cc) / our TFeaCatCu on the same files: per-utterance output features"""
    rng = np.random.default_rng(cfg["seed"] + 3000)
    utts = F.gen_utterances(cfg["n_utt"], cfg["n_frames"], cfg["raw_dim"], cfg["n_out"], rng)
    paths = F.write_dataset(workdir, utts, cfg["n_out"], cfg["ctx"])
    dims = [cfg["raw_dim"] * (2 * cfg["ctx"] + 1)] + cfg["hidden"] + [cfg["n_out"]]
    layers = F.gen_mlp_init(dims, rng, negbias=False)
    init = os.path.join(workdir, "init.nnet")
    F.write_mlp(init, layers)
    layers = F.read_mlp(init)
    outdir = os.path.join(workdir, "out")
    os.makedirs(outdir, exist_ok=True)
    if exe is None:
        exe = os.path.join(REF, "TFeaCat")
    cmd = [exe, "-H", init, "-S", paths["scp"], "-l", outdir, "-y", "post", "--FEATURETRANSFORM=" + paths["transform"],
           "--STARTFRMEXT=%d" % cfg["ctx"], "--ENDFRMEXT=%d" % cfg["ctx"], "--LOGPOSTERIOR=" + _b(cfg["log"])]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        raise RuntimeError("TFeaCat failed:\n" + res.stdout[-3000:])
    names = list(utts.keys())
    outs = [F.read_htk(os.path.join(outdir, os.path.splitext(os.path.basename(f))[0] + ".post"))[0] for f in paths["files"]]
    if not save:
        return outs, res.stdout
    data = dict(feats=np.concatenate([utts[n][0] for n in names]), lengths=np.array([utts[n][0].shape[0] for n in names], dtype=np.int32),
                dims=np.array(dims, dtype=np.int32), cfg=np.array([cfg["ctx"], int(cfg["log"])], dtype=np.int64),
                ref_out=np.concatenate(outs))
    pack_layers("init", layers, data)
    np.savez_compressed(os.path.join(OUT, "cpu_%s.npz" % case), **data)
    print("cpu %s: %d utterances, output %s" % (case, len(outs), data["ref_out"].shape))


NORM_CASES = {"norm_splice": dict(raw_dim=13, ctx=3, n_utt=9, n_frames=80, seed=41)}


def run_norm(case, cfg, workdir, exe=None, save=True):
    """reference CPU TNorm (src/TNorm.cc) / our TNormCu: global mean/variance of the spliced features -> <bias> + <window>"""
    rng = np.random.default_rng(cfg["seed"] + 4000)
    utts = F.gen_utterances(cfg["n_utt"], cfg["n_frames"], cfg["raw_dim"], 4, rng)
    utts = {k: ((v[0] * rng.uniform(0.2, 3.0, cfg["raw_dim"]) + rng.uniform(-2, 2, cfg["raw_dim"])).astype(np.float32), v[1]) for k, v in utts.items()}
    paths = F.write_dataset(workdir, utts, 4, cfg["ctx"])
    target = os.path.join(workdir, "norm.transf")
    cmd = [exe or os.path.join(REF, "TNorm"), "-H", paths["transform"], "-S", paths["scp"], "--TARGETMMF=" + target,
           "--STARTFRMEXT=%d" % cfg["ctx"], "--ENDFRMEXT=%d" % cfg["ctx"]]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        raise RuntimeError("TNorm failed:\n" + res.stdout[-3000:])
    tk = open(target).read().split()
    assert tk[0] == "<bias>" and tk[3] == "v"
    dim = int(tk[1])
    bias = np.array(tk[5:5 + dim], dtype=np.float64)
    w0 = 5 + dim
    assert tk[w0] == "<window>" and tk[w0 + 3] == "v"
    window = np.array(tk[w0 + 5:w0 + 5 + dim], dtype=np.float64)
    m = re.search(r"frames: (\d+)", res.stdout)
    frames = int(m.group(1)) if m else -1
    if not save:
        return bias, window, frames, res.stdout
    names = list(utts.keys())
    np.savez_compressed(os.path.join(OUT, "cpu_%s.npz" % case), feats=np.concatenate([utts[n][0] for n in names]),
                        lengths=np.array([utts[n][0].shape[0] for n in names], dtype=np.int32), cfg=np.array([cfg["ctx"]], dtype=np.int64),
                        ref_bias=bias, ref_window=window, ref_frames=np.int64(frames))
    print("cpu %s: dim %d frames %d" % (case, dim, frames))


def main():
    global OUT
    ap = argparse.ArgumentParser()
    ap.add_argument("--impl", choices=["cpu", "gpu"], required=True)
    ap.add_argument("--out", default=OUT)
    ap.add_argument("--only", default=None, help="regular expression: generate only the cases whose name matches")
    a = ap.parse_args()
    OUT = a.out
    os.makedirs(OUT, exist_ok=True)

    def want(case):
        return a.only is None or re.search(a.only, case) is not None

    for case, cfg in MLP_CASES.items():
        if (a.impl == "cpu" and case in GPU_ONLY) or not want(case):
            continue
        with tempfile.TemporaryDirectory() as d:
            run_mlp(case, cfg, a.impl, d)
    for case, cfg in OPT_CASES.items():
        if want(case) and not (a.impl == "cpu" and cfg.get("factors")):   # the CPU trainer advertises LEARNRATEFACTORS but rejects it
            with tempfile.TemporaryDirectory() as d:
                run_mlp(case, cfg, a.impl, d)
    for case, cfg in NET_CASES.items():
        if (a.impl == "cpu" and case in NET_GPU_ONLY) or not want(case):
            continue
        with tempfile.TemporaryDirectory() as d:
            run_net(case, cfg, a.impl, d)
    if a.impl == "cpu":
        for case, cfg in NORM_CASES.items():
            if want(case):
                with tempfile.TemporaryDirectory() as d:
                    run_norm(case, cfg, d)
        for case, cfg in FEACAT_CASES.items():
            if want(case):
                with tempfile.TemporaryDirectory() as d:
                    run_feacat(case, cfg, d)
    if a.impl == "gpu":
        for case, cfg in RBM_CASES.items():
            if want(case):
                with tempfile.TemporaryDirectory() as d:
                    run_rbm(case, cfg, d)
        for case, cfg in RNN_CASES.items():
            if want(case):
                with tempfile.TemporaryDirectory() as d:
                    run_rnn(case, cfg, d)


if __name__ == "__main__":
    main()
