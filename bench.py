#!/usr/bin/env python
"""bench.py — training frames/s of TNet's bunch hot path (forward, softmax+xent, backward, SGD update) on B200.

  python bench.py --gpus N --steps K --warmup W            this repo's CUDA path (one rank per GPU under torchrun)
  python bench.py --impl reference --gpus N --steps K ...   the reference's own CPU trainer (oracle/_ref/TNet) on host cores

Workload (BASELINE.json configs[2], the configuration the metric is quoted on): deep DNN 429-2048x6-3000 sigmoid/softmax,
bunch 1024 PER GPU, synthetic N(0,1) features and uniform labels, random-init weights (W~0.1*N(0,1), hidden bias
U[-4.1,-3.9]), learning rate 0.008 / momentum 0.5 / weightcost 1e-6, GRADDIVFRM on.  A step = one bunch: CuCache::GetBunch
row window, forward, fused softmax+cross-entropy+accuracy, backward, momentum/L2 update.  Data parallel: every rank owns
its rows of the (N x 1024)-frame global bunch, per-layer [dW|db] is summed over NVLink peer memory (csrc/peer.cu), N in the update rule is the global
frame count ("scaling": "weak").

One JSON line on stdout (rank 0).  `value` = frames/s with the training set resident in HBM; `e2e` = the same through
the host-buffer entry points (pinned host features + int labels copied H2D every step, statistics read back D2H every
step, one submission kept in flight (three with several ranks) so that the copy of the next bunch overlaps the current step); `roofline` = algorithmic GEMM flops / CUDA-event time of the tcgen05 GEMM launches of a second pass over the same K steps (an
event pair around every launch; separate from the pass `value` is timed on, where consecutive GEMMs overlap by PDL);
`cpu_baseline` = the reference CPU trainer on a bounded sample of the same workload (rank 0, N=1 only).
"""
import argparse
import ctypes as C
import json
import os
import re
import shutil
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "nnet-asr_b200", "python"))

DIMS = [429, 2048, 2048, 2048, 2048, 2048, 2048, 3000]
BUNCH = 1024
RAW_DIM, CTX = 39, 5
LR, MMT, WC = 0.008, 0.5, 1e-6
RESIDENT_BUNCHES = 16
WORKLOAD = "deep DNN 429-2048x6-3000 sigmoid/softmax, bunch 1024 per GPU (BASELINE configs[2])"


def flops_per_frame(dims):
    """6*sum(in*out) - 2*in_1*out_1 (no dX for the first trainable layer)."""
    s = sum(a * b for a, b in zip(dims[:-1], dims[1:]))
    return 6 * s - 2 * dims[0] * dims[1]


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(bf16_sustained=d.get("bf16_tflops_sustained", 1388.8), bf16_burst=d.get("bf16_tflops", 1648.3),
                    hbm=d.get("hbm_gbs", 6552.6), source="measured (MEASURED_PEAKS.json)")
    return dict(bf16_sustained=1400.0, bf16_burst=1590.0, hbm=6650.0, source="fallback (B200_PROFILING.md)")


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.f = tempfile.NamedTemporaryFile(mode="w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "20"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        out = dict(sm_mhz=None, sm_max_mhz=None, reasons=[])
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        rows = [l.strip().split(", ") for l in open(self.f.name) if l.strip()]
        os.unlink(self.f.name)
        rows = [r for r in rows if len(r) >= 9 and r[0].strip() == str(self.idx)]
        if not rows:
            return out
        sm = sorted(float(r[1]) for r in rows if r[1].replace(".", "").isdigit())
        if sm:
            out["sm_mhz"] = sm[len(sm) // 2]
            out["sm_max_mhz"] = float(rows[0][2])
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for k, n in enumerate(names):
            if any(r[5 + k].strip().lower() == "active" for r in rows):
                out["reasons"].append(n)
        out["power_w_max"] = max((float(r[3]) for r in rows if re.match(r"^[0-9.]+$", r[3])), default=None)
        return out


# ------------------------------------------------------------------------------------------------ reference CPU arm
def write_reference_dataset(d, n_frames, utt_len=256, seed=20240607):
    """Files the reference trainers read: HTK features (raw 39-dim), <expand> +-5 transform, MLF, label map, SCP and the
    429-2048x6-3000 network (tools/init/gen_mlp_init.py --gauss --negbias statistics)."""
    from tnet_b200 import formats as F
    rng = np.random.default_rng(seed)
    utts = F.gen_utterances(max(1, n_frames // utt_len), utt_len, RAW_DIM, DIMS[-1], rng, vary_len=False)
    paths = F.write_dataset(d, utts, DIMS[-1], CTX)
    net = os.path.join(d, "init.nnet")
    with open(net, "w") as f:           # numpy's C-level text writer: ~300 MB of weights in a few seconds
        for l in range(len(DIMS) - 1):
            nin, nout = DIMS[l], DIMS[l + 1]
            last = l == len(DIMS) - 2
            f.write("<biasedlinearity> %d %d\nm %d %d\n" % (nout, nin, nout, nin))
            f.flush()
            (0.1 * rng.standard_normal(nout * nin)).astype(np.float32).tofile(f, sep=" ", format="%.6g")
            f.write("\nv %d\n" % nout)
            f.flush()
            b = np.zeros(nout, np.float32) if last else (rng.random(nout) / 5.0 - 4.1).astype(np.float32)
            b.tofile(f, sep=" ", format="%.6g")
            f.write("\n<%s> %d %d\n" % ("softmax" if last else "sigmoid", nout, nout))
    paths["net"] = net
    return paths


def reference_cachesize(threads, utt_len):
    """The CPU trainer gives every worker thread its own cache of (CACHESIZE / threads / (BUNCHSIZE / threads)) * (BUNCHSIZE / threads)
    rows (TNetLib/Platform.h:159-160) and refuses segments longer than half of it (TNetLib/Cache.cc:66-72): size the total so that
    each thread's share holds four of its bunch slices AND four utterances, whatever the core count of the box."""
    per_thread_bunch = max(1, BUNCH // threads)
    per_thread = max(4 * per_thread_bunch, 4 * utt_len)
    per_thread = ((per_thread + per_thread_bunch - 1) // per_thread_bunch) * per_thread_bunch
    return per_thread * threads


def run_reference_tnet(paths, d, threads, n_utts_scp=None, utt_len=256):
    exe = os.path.join(ROOT, "oracle", "_ref", "TNet")
    cachesize = reference_cachesize(threads, utt_len)
    scp = paths["scp"]
    if n_utts_scp is not None:
        scp = os.path.join(d, "sub_%d.scp" % n_utts_scp)
        open(scp, "w").write("\n".join(paths["files"][:n_utts_scp]) + "\n")
    cmd = [exe, "-H", paths["net"], "-I", paths["mlf"], "-L", "*/", "-X", "lab", "-S", scp, "-m", paths["labelmap"],
           "-n", repr(LR / BUNCH),      # the CPU trainer has no 1/N: the schedulers divide lr by the bunch size
           "--TARGETMMF=" + os.path.join(d, "out.nnet"), "--BUNCHSIZE=%d" % BUNCH, "--CACHESIZE=%d" % cachesize,
           "--RANDOMIZE=TRUE", "--SEED=123", "--FEATURETRANSFORM=" + paths["transform"], "--STARTFRMEXT=%d" % CTX,
           "--ENDFRMEXT=%d" % CTX, "--WEIGHTCOST=%g" % WC, "--THREADS=%d" % threads]
    env = dict(os.environ, OPENBLAS_NUM_THREADS="1", OMP_NUM_THREADS="1")
    t0 = time.perf_counter()
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env)
    wall = time.perf_counter() - t0
    if res.returncode != 0:
        raise RuntimeError("reference TNet failed:\n" + res.stdout[-2000:])
    m = re.search(r"frames:(\d+)", res.stdout)
    return wall, int(m.group(1)) if m else 0


def cpu_reference_throughput(bunches_a, bunches_b, threads=None):
    """frames/s of the reference CPU trainer on config C, from the SLOPE between two runs of different length so that the
    fixed cost of parsing/writing the 28M-weight text network (inside TNet's own timer) drops out."""
    exe = os.path.join(ROOT, "oracle", "_ref", "TNet")
    if not os.path.exists(exe):
        raise RuntimeError("oracle/_ref/TNet is not built (run __graft_entry__.build() where /root/reference exists)")
    ncpu = os.cpu_count() or 1
    if threads is None and os.environ.get("TNB_REF_THREADS"):
        threads = int(os.environ["TNB_REF_THREADS"])      # testing: e.g. 32 threads on a smaller box
    if threads is None:
        threads = 1
        while threads * 2 <= min(ncpu, 32) and BUNCH % (threads * 2) == 0:
            threads *= 2
    utt_len = 256
    d = tempfile.mkdtemp(prefix="tnet_cpu_")
    try:
        need = max(bunches_a, bunches_b) * BUNCH
        paths = write_reference_dataset(d, need, utt_len)
        per = BUNCH // utt_len
        ta, fa = run_reference_tnet(paths, d, threads, bunches_a * per, utt_len)
        tb, fb = run_reference_tnet(paths, d, threads, bunches_b * per, utt_len)
    finally:
        shutil.rmtree(d, ignore_errors=True)
    if fb <= fa or tb <= ta:
        raise RuntimeError("reference runs did not scale: %s frames in %.1fs vs %s in %.1fs" % (fa, ta, fb, tb))
    fps = (fb - fa) / (tb - ta)
    return fps, threads, "reference TNet --THREADS=%d, slope between %d and %d frames of the workload (%.1fs, %.1fs wall)" % (
        threads, fa, fb, ta, tb)


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    steps = max(1, min(args.steps, 12))     # bounded sample: each CPU step is one bunch (1024 frames in config C)
    try:
        # the small configs train a bunch in a few milliseconds on the host: longer runs, so that every worker thread sees data
        fps, threads, sample = cpu_reference_throughput(2, 2 + steps) if args.config == "C" else cpu_reference_throughput(40, 40 + 20 * steps)
    except Exception as e:  # the oracle always exists; a failure here is an error, not "unavailable"
        print(json.dumps({"impl": "reference", "error": str(e)[:300]}))
        return 1
    line = {
        "impl": "reference", "metric": "training_frames_per_sec", "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": 2, "ms_per_step": 1000.0 * BUNCH / fps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "bunch": BUNCH, "note": "CPU reference trains ONE %d-frame bunch per step on host cores" % BUNCH},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": threads, "kind": "reference", "sample": sample},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------ the other BASELINE configs
CONFIGS = {
    # name: dims, bunch, splice context, momentum, workload text (BASELINE.json configs[0..4])
    "A": ([351, 1024, 135], 256, 4, 0.0, "examples/01test-shape MLP3 351-1024-135, bunch 256 (BASELINE configs[0])"),
    "B": ([351, 2048, 135], 256, 4, 0.5, "MLP3 newbob TIMIT-shape 351-2048-135, bunch 256, momentum (BASELINE configs[1])"),
    "C": (None, None, None, None, None),
}


def select_config(name):
    global DIMS, BUNCH, CTX, MMT, WORKLOAD
    if name in ("A", "B"):
        DIMS, BUNCH, CTX, MMT, WORKLOAD = CONFIGS[name]


def bench_rbm_rnn(args):
    """BASELINE configs[3] (D: TRbmCu CD-1, Gaussian-Bernoulli 429 x 2048, bunch 128) and configs[4] (E: TRecurrentCu 351+1024 -> 1024 -> 135,
    BPTT 20): frames/s through the C handles of libtnetb200_host.so, the oracle port timed beside it on a bounded sample."""
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    import torch
    from tnet_b200 import abi, host
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    torch.cuda.set_device(0)
    host.select_gpu(0)
    host.set_math({"3xtf32": abi.MATH_3XTF32, "tf32": abi.MATH_TF32, "bf16": abi.MATH_BF16}[args.math])
    L, ctx = abi.lib(), host.ctx_handle()
    stream = torch.cuda.ExternalStream(host_stream(L, ctx), device=torch.device("cuda", 0))
    r = np.random.default_rng(20240607)
    pk = measured_peaks()
    sampler = ClockSampler(0)
    sampler.start()

    def windows(fn, n):
        out = []
        for _ in range(n):
            host.sync(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            fn()
            e1.record(stream)
            host.sync(); torch.cuda.synchronize()
            out.append(e0.elapsed_time(e1))
        return out

    if args.config == "D":
        nvis, nhid, bunch = 429, 2048, 128
        Wt = (0.1 * r.standard_normal((nhid, nvis))).astype(np.float32)
        rbm = host.Rbm(Wt, np.zeros(nvis, np.float32), np.zeros(nhid, np.float32), True, False, bunch, 0.001, 0.5, 2e-4)
        steps = max(args.steps, 50)
        X = r.standard_normal((steps * bunch, nvis)).astype(np.float32)
        cache = host.Cache(steps * bunch, bunch)

        def fill():
            cache.add(X, np.zeros((X.shape[0], 1), np.float32))
        for _ in range(max(1, args.warmup // 3)):
            fill(); rbm.cd1_from_cache(cache)
        ms = []
        l0 = 0
        for _ in range(args.windows):
            fill()                                   # host -> device upload of the window's frames: outside the timed region
            l0 = host.launches()
            ms += windows(lambda: rbm.cd1_from_cache(cache), 1)
            l0 = host.launches() - l0
        t = sorted(ms)[len(ms) // 2]
        frames = steps * bunch
        value = frames / (t / 1000.0)
        # end to end: every bunch uploaded from host memory inside the timed region (tnh_rbm_cd1_bunch: upload, step, synchronise)
        t0 = time.perf_counter()
        for b in range(steps):
            rbm.cd1(X[b * bunch:(b + 1) * bunch])
        e2e = frames / (time.perf_counter() - t0)
        fpf = 10 * nvis * nhid
        # oracle port on a bounded sample
        orbm = O.Rbm(Wt, np.zeros(nvis, np.float32), np.zeros(nhid, np.float32), True, False, 0.001, 0.5, 2e-4)
        z = [r.integers(129, 2 ** 31, (bunch, nhid)).astype(np.uint32) for _ in range(4)]
        t0 = time.perf_counter()
        nb = 0
        while time.perf_counter() - t0 < 10.0:
            orbm.cd1(X[:bunch], z); nb += 1
        cpu = nb * bunch / (time.perf_counter() - t0)
        work = "TRbmCu CD-1, Gaussian-Bernoulli RBM 429 x 2048, bunch 128, lr 0.001 / momentum 0.5 / weightcost 2e-4 (BASELINE configs[3])"
        cfg = {"workload": work, "bunch": bunch, "flops_per_frame": fpf, "windows_ms": ms, "gemm_math": args.math,
               "timed_region": "K x (propagate, sample, reconstruct, propagate, fused CD-1 update, MSE) from a device-resident cache"}
        e2e_obj = {"value": e2e, "unit": "frames/s", "h2d_bytes_per_step": bunch * nvis * 4, "d2h_bytes_per_step": 0}
        sample = "oracle port (oracle/tnet_oracle.c, OpenMP GEMM), %d CD-1 bunches of 128 frames" % nb
    else:
        nin, H, nout, bptt, T = 351, 1024, 135, 20, 1000
        Wr = (0.1 * r.standard_normal((H, nin + H))).astype(np.float32)
        Wo = (0.1 * r.standard_normal((nout, H))).astype(np.float32)
        layers = [("recurrent", Wr, np.zeros(H, np.float32), nin), ("affine", Wo, np.zeros(nout, np.float32)), ("softmax", nout)]
        rnn = host.Rnn(layers, bptt, 0.001)
        X = r.standard_normal((T, nin)).astype(np.float32)
        lab = r.integers(0, nout, T).astype(np.int32)
        rnn.train_utterance(X[:200], lab[:200])
        ms = []
        l0 = 0
        for _ in range(min(3, args.windows)):
            l0 = host.launches()
            ms += windows(lambda: rnn.train_utterance(X, lab), 1)
            l0 = host.launches() - l0
        t = sorted(ms)[len(ms) // 2]
        frames, steps, bunch = T, T, 1
        value = frames / (t / 1000.0)
        e2e = value                                  # train_utterance already takes host buffers (1.4 MB upload inside the timed call)
        fpf = 2 * (nin + H) * H * (2 + bptt) + 4 * H * nout   # forward GEMV + BPTT chain GEMVs + rank-1 updates + output layer
        # oracle port on a bounded sample: frames of the same layer through orc_rnn_propagate / orc_rnn_update
        ornn = O.Rnn(Wr, np.zeros(H, np.float32), nin, bptt, 0.001)
        t0 = time.perf_counter()
        nf = 0
        while time.perf_counter() - t0 < 10.0 and nf < T:
            ornn.propagate(X[nf]); ornn.update(np.full(H, 1e-3, np.float32)); nf += 1
        cpu = nf / (time.perf_counter() - t0)
        work = "TRecurrentCu simple recurrent layer 351+1024 -> 1024 -> 135 softmax, BPTT 20, one 1000-frame utterance (BASELINE configs[4])"
        cfg = {"workload": work, "bunch": 1, "flops_per_frame": fpf, "windows_ms": ms, "gemm_math": "fp32 (GEMV / rank-1 kernels)",
               "timed_region": "one utterance, frame by frame (forward, cross-entropy, truncated BPTT update), per-frame sequence replayed as a CUDA graph"}
        e2e_obj = {"value": e2e, "unit": "frames/s", "h2d_bytes_per_step": (nin + 1) * 4, "d2h_bytes_per_step": 0}
        sample = "oracle port (recurrent layer only: orc_rnn_propagate + orc_rnn_update), %d frames" % nf
    clocks = sampler.stop()
    achieved = fpf * value / 1e12
    line = {"metric": "training_frames_per_sec", "value": value, "unit": "frames/s", "n_gpus": 1, "steps": steps, "warmup": args.warmup,
            "ms_per_step": t / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if args.config == "E" else {"3xtf32": "tf32x3", "tf32": "tf32", "bf16": "bf16"}[args.math], "data": "synthetic",
            "config": cfg, "clocks": clocks, "e2e": e2e_obj, "gpu_launches": int(l0),
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": pk["bf16_sustained"], "unit": "TFLOP/s", "frac": achieved / pk["bf16_sustained"],
                         "traffic": None, "peak_source": pk["source"] + " (bf16 sustained)",
                         "note": "whole step (every kernel) as algorithmic flops per second; config %s is %s" % (
                             args.config, "launch/latency-bound at bunch 128 (a 429 x 2048 layer)" if args.config == "D" else
                             "frame-serial by construction (the weights change every frame): frames/s is the figure of merit")},
            "cpu_baseline": {"value": cpu, "unit": "frames/s", "cores": os.cpu_count() if args.config == "D" else 1, "kind": "port", "sample": sample}}
    os.write(json_fd, (json.dumps(line) + "\n").encode())
    return 0


# ------------------------------------------------------------------------------------------------ this repo's arm
def ncu_traffic(math):
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the dominant GEMM from the committed ncu --set full capture
    (profiles/ncu_traffic.json, written by tools/summarize_ncu.py from the .ncu-rep of the same bench command); None without one."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if not os.path.exists(p):
        return None, None
    try:
        d = json.load(open(p)).get(math)
        return (d["dram_bytes_per_launch"], d.get("source")) if d else (None, None)
    except Exception:
        return None, None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--math", default="3xtf32", choices=["3xtf32", "tf32", "bf16"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: 1024 frames per GPU (global bunch 1024*N); strong: the 1024-frame bunch partitioned over the GPUs")
    ap.add_argument("--windows", type=int, default=5, help="timed windows of --steps bunches; the median is reported")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the bf16-mode and other-scaling sub-objects and the peak probe")
    ap.add_argument("--config", default="C", choices=["A", "B", "C", "D", "E"],
                    help="BASELINE.json configs[0..4]; C (the configuration the metric is quoted on) is the default and the only multi-GPU one")
    args = ap.parse_args()
    select_config(args.config)
    if args.impl == "reference":
        return reference_arm(args)
    if args.config in ("D", "E"):
        args.warmup = max(args.warmup, 3)
        return bench_rbm_rnn(args)
    args.warmup = max(args.warmup, 3)
    args.windows = max(args.windows, 1)
    # stdout carries exactly ONE line (the JSON): libraries that print there (NCCL's version banner does) go to stderr instead
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    import torch
    import torch.distributed as dist
    from tnet_b200 import abi, host

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        args.gpus = world
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local))
    host.select_gpu(local)
    MATH = {"3xtf32": abi.MATH_3XTF32, "tf32": abi.MATH_TF32, "bf16": abi.MATH_BF16}
    L, H = abi.lib(), host.hlib()
    ctx = host.ctx_handle()

    # ---- data-parallel communicator of the library (NCCL id exchanged through torch.distributed) ----
    if world > 1:
        idbuf = (C.c_ubyte * 128)()
        if rank == 0:
            abi.check(L.tnb_comm_unique_id(idbuf))
        t = torch.tensor(list(bytes(idbuf)), dtype=torch.uint8, device="cuda")
        dist.broadcast(t, 0)
        idbuf = (C.c_ubyte * 128)(*t.cpu().tolist())
        abi.check(L.tnb_comm_init(ctx, idbuf, C.c_int(rank), C.c_int(world)))

    stream = torch.cuda.ExternalStream(host_stream(L, ctx), device=torch.device("cuda", local))
    dp_state = {"schedule": "none"}

    def barrier():
        host.sync()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(x):
        if world == 1:
            return x
        tt = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item())

    def make_net(math, scaling):
        """network + resident synthetic set for one (math mode, scaling) arm.  weak: every rank draws its own 1024-frame bunches;
        strong: ONE global set (same seed everywhere), rank g owns rows [g*B/G, (g+1)*B/G) of every 1024-frame bunch
        (the reference CPU trainer's bunchsize_/num_thr, TNetLib/Platform.h:159)."""
        host.set_math(MATH[math])
        net = host.Net(dims=DIMS, seed=1)
        net.set_hyper(LR, mmt=MMT, wc=WC, gdf=True)
        if world > 1:
            dp_state["schedule"] = os.environ.get("TNB_DP_MODE", "peer")
            ok = 1
            try:
                net.set_data_parallel(world)
            except abi.TnbError as e:
                ok = 0
                sys.stderr.write("[bench] rank %d: data-parallel set-up (%s) failed: %s\n" % (rank, dp_state["schedule"], e))
            tt = torch.tensor([ok], dtype=torch.int32, device="cuda")
            dist.all_reduce(tt, op=dist.ReduceOp.MIN)
            if int(tt.item()) == 0:
                if dp_state["schedule"] != "peer":
                    raise SystemExit("data-parallel set-up failed")
                # the peer-memory schedule needs CUDA IPC between the ranks' processes; where that is not available the NCCL
                # schedule still is — say so loudly and in the JSON line instead of losing the measurement
                sys.stderr.write("[bench] falling back to TNB_DP_MODE=allreduce (NCCL) on all ranks\n")
                os.environ["TNB_DP_MODE"] = "allreduce"
                dp_state["schedule"] = "allreduce (peer-memory set-up failed)"
                net = host.Net(dims=DIMS, seed=1)
                net.set_hyper(LR, mmt=MMT, wc=WC, gdf=True)
                net.set_data_parallel(world)
        if scaling == "weak" or world == 1:
            bunch = BUNCH
            rng = np.random.default_rng(20240607 + rank)
            X = rng.standard_normal((RESIDENT_BUNCHES * bunch, DIMS[0])).astype(np.float32)
            lab = rng.integers(0, DIMS[-1], RESIDENT_BUNCHES * bunch).astype(np.int32)
        else:
            if BUNCH % world:
                raise SystemExit("strong scaling needs the bunch (%d) to divide by the GPU count" % BUNCH)
            bunch = BUNCH // world
            rng = np.random.default_rng(20240607)
            Xg = rng.standard_normal((RESIDENT_BUNCHES, BUNCH, DIMS[0])).astype(np.float32)
            labg = rng.integers(0, DIMS[-1], (RESIDENT_BUNCHES, BUNCH)).astype(np.int32)
            X = np.ascontiguousarray(Xg[:, rank * bunch:(rank + 1) * bunch]).reshape(-1, DIMS[0])
            lab = np.ascontiguousarray(labg[:, rank * bunch:(rank + 1) * bunch]).reshape(-1)
        net.load_resident(X, lab)
        return net, bunch, X, lab

    def timed_windows(net, bunch, n_windows):
        """n_windows x (args.steps bunches bracketed by barrier + synchronize, CUDA events on the library's stream, max over ranks)"""
        net.train_resident(bunch, 0, args.warmup)
        barrier()
        out, first, launches = [], args.warmup, 0
        for _ in range(n_windows):
            l0 = host.launches()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            net.train_resident(bunch, first, args.steps)
            e1.record(stream)
            barrier()
            out.append(allmax(e0.elapsed_time(e1)))
            launches = host.launches() - l0
            first += args.steps
        return out, launches, first

    def median(v):
        v = sorted(v)
        return v[len(v) // 2] if len(v) % 2 else 0.5 * (v[len(v) // 2 - 1] + v[len(v) // 2])

    def gemm_pass(net, bunch, first):
        # roofline pass: K steps with a CUDA event pair around every GEMM launch (tnb_ctx_profile_*).  Kept out of the timed windows
        # because an event record between two kernels removes their programmatic-dependent-launch overlap.
        abi.check(L.tnb_ctx_profile_begin(ctx))
        net.train_resident(bunch, first, args.steps)
        barrier()
        gms, gl, gfl = C.c_double(), C.c_ulonglong(), C.c_double()
        abi.check(L.tnb_ctx_profile_end(ctx, C.byref(gms), C.byref(gl), C.byref(gfl)))
        return gms.value, int(gl.value), gfl.value

    # ================================================================== headline arm
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()             # clocks are sampled from the warm-up to the end of the timed windows
    net, bunch, X, lab = make_net(args.math, args.scaling)
    win_ms, launches, first = timed_windows(net, bunch, args.windows)
    clocks = sampler.stop() if rank == 0 else {}
    ms = median(win_ms)
    if world > 1 and os.environ.get("TNB_DP_TRACE"):
        # debugging: %globaltimer stamps of the last peer-memory update kernels of this rank (csrc/peer.cu), one line per launch
        buf, seq = (C.c_longlong * 384)(), (C.c_uint * 2)()
        if L.tnb_peer_trace_read(ctx, buf, seq) == 0:
            n_upd = len(DIMS) - 1
            ev = []
            for s_ in range(seq[0] - 3 * n_upd + 1, seq[0] + 1):
                r_ = [buf[(s_ % 64) * 4 + k] for k in range(4)]
                ev.append((r_[0], "kernel entry | wait ready %6.1f | own rows %6.1f | wait done %6.1f | exit %10.1f", r_))
            for s_ in range(max(0, seq[1] - 3 * n_upd), seq[1]):
                r_ = [buf[256 + (s_ % 64) * 2 + k] for k in range(2)]
                ev.append((r_[0], "push start   | copies %6.1f", r_))
            ev.sort(key=lambda e_: e_[0])
            t0 = ev[0][0]
            txt = ["[dp trace] rank %d: peer-memory kernels and copy-engine pushes of the last 3 bunches, us" % rank]
            for t_, fmt, r_ in ev:
                if len(r_) == 4:
                    line = fmt % ((r_[1] - r_[0]) / 1e3, (r_[2] - r_[1]) / 1e3, (r_[3] - r_[2]) / 1e3, (r_[3] - t0) / 1e3)
                else:
                    line = fmt % ((r_[1] - r_[0]) / 1e3)
                txt.append("[dp trace] r%d %9.1f %s" % (rank, (t_ - t0) / 1e3, line))
            sys.stderr.write("\n".join(txt) + "\n")
    gms, gl, gfl = gemm_pass(net, bunch, first)
    frames = args.steps * bunch * world
    value = frames / (ms / 1000.0)

    # ---- end to end: pinned host buffers in, statistics out, every step ----
    xb = torch.empty((bunch, DIMS[0]), dtype=torch.float32).pin_memory()
    lb = torch.empty((bunch,), dtype=torch.int32).pin_memory()
    xb.copy_(torch.from_numpy(X[:bunch]))
    lb.copy_(torch.from_numpy(lab[:bunch]))
    xp, lp = C.cast(xb.data_ptr(), C.POINTER(C.c_float)), C.cast(lb.data_ptr(), C.POINTER(C.c_int))

    # One submission is kept in flight: submit bunch k+1 (H2D on the copy stream, step behind it), then collect bunch k (blocks
    # until ITS statistics have been copied back).  Every step's input crosses PCIe inside the timed region and every step's
    # result is read by the host; the copy of bunch k+1 overlaps the step of bunch k, as a loader thread would arrange it.
    host_t = {"submit": 0.0, "collect": 0.0}   # host time inside the two calls (diagnostic: enqueue cost vs waiting for the GPU)

    # submissions kept in flight behind the one being collected: one is enough on a single GPU; with several ranks a host thread
    # that is late by a fraction of a millisecond would otherwise stall its GPU and, at the next exchange, every other rank's
    E2E_DEPTH = 1 if world == 1 else 3

    def e2e_run(n):
        host_t["submit"] = host_t["collect"] = 0.0
        ahead = min(E2E_DEPTH, n)
        for _ in range(ahead):
            net.submit_bunch_labels(xp, lp, bunch)
        st_ = None
        for _ in range(n - ahead):
            ta = time.perf_counter()
            net.submit_bunch_labels(xp, lp, bunch)
            tb = time.perf_counter()
            st_ = net.collect()
            host_t["submit"] += tb - ta
            host_t["collect"] += time.perf_counter() - tb
        for _ in range(ahead):
            st_ = net.collect()
        return st_

    e2e_run(3)
    e2e_s, st = [], None
    for _ in range(min(3, args.windows)):
        barrier()
        t0 = time.perf_counter()
        st = e2e_run(args.steps)
        barrier()
        e2e_s.append(allmax(time.perf_counter() - t0))
    e2e_value = frames / median(e2e_s)
    net.close()

    # ================================================================== the other arms, reported inside the same line
    extras = {}
    if not args.no_extras:
        def side_arm(math, scaling):
            n2, b2, _, _ = make_net(math, scaling)
            w2, l2, f2 = timed_windows(n2, b2, max(3, min(args.windows, 5)))
            g2 = gemm_pass(n2, b2, f2)
            n2.close()
            return b2, median(w2), w2, l2, g2

        def try_side_arm(key, math, scaling):
            # a sub-object that fails must not lose the headline measurement: it is dropped (on every rank) and named on stderr
            res, ok = None, 1
            try:
                res = side_arm(math, scaling)
            except Exception as e:
                ok = 0
                sys.stderr.write("[bench] rank %d: the %s arm failed and is left out of the line: %s\n" % (rank, key, e))
            if world > 1:
                tt = torch.tensor([ok], dtype=torch.int32, device="cuda")
                dist.all_reduce(tt, op=dist.ReduceOp.MIN)
                ok = int(tt.item())
            if ok:
                extras[key] = res
        if args.math == "3xtf32":
            try_side_arm("bf16", "bf16", args.scaling)
        if world > 1:
            other = "strong" if args.scaling == "weak" else "weak"
            try_side_arm(other, args.math, other)
        host.set_math(MATH[args.math])

    # ---- measured dense peaks of THIS box (library GEMM as a yardstick; nothing in the product calls it) ----
    peaks_live = None
    if rank == 0 and not args.no_extras:
        try:
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            from tf32_peak import probe_dense_peak
            peaks_live = {"tf32": probe_dense_peak("tf32", sustain_s=1.5), "bf16": probe_dense_peak("bf16", sustain_s=1.5)}
        except Exception as e:  # the probe is a yardstick: its failure must not lose the measurement
            sys.stderr.write("[bench] dense-peak probe failed: %s\n" % e)

    if rank == 0:
        pk = measured_peaks()
        fpf = flops_per_frame(DIMS)

        def roof(math, gms_, gl_, gfl_, ms_step, bunch_):
            gemm_tflops = (gfl_ / (gms_ / 1000.0)) / 1e12 if gms_ > 0 else 0.0
            passes = 3 if math == "3xtf32" else 1
            traffic, tsrc = ncu_traffic(math)
            r = {"bound": "tensor", "achieved": gemm_tflops, "peak": pk["bf16_sustained"], "unit": "TFLOP/s",
                 "frac": gemm_tflops / pk["bf16_sustained"], "traffic": traffic, "traffic_source": tsrc,
                 "kernel": "tcgen05 GEMM launches of K steps, one CUDA event pair per launch (no PDL overlap in this pass)",
                 "gemm_ms_per_step": gms_ / args.steps, "gemm_launches": gl_, "issued_tflops": gemm_tflops * passes,
                 "peak_source": pk["source"] + " (bf16 sustained)",
                 # whole step (every kernel, PDL overlap included) as algorithmic flops per second, per GPU
                 "step_tflops": fpf * bunch_ / (ms_step / 1000.0) / 1e12,
                 "step_frac": fpf * bunch_ / (ms_step / 1000.0) / 1e12 / pk["bf16_sustained"]}
            if peaks_live and math != "bf16":
                tpk = peaks_live["tf32"]
                r["tf32_peak_measured"] = {"burst_tflops": tpk["burst_tflops"], "sustained_tflops": tpk["sustained_tflops"],
                                           "how": "torch.matmul fp32 %d^3 with TF32 allowed: best of 10 and back to back 1.5 s, this run" % tpk["n"]}
                # issued tf32 flops against the MEASURED tf32 peak: the burst figure for the kernel timed alone in the event pass
                r["frac_issued_of_measured_tf32_burst"] = gemm_tflops * passes / tpk["burst_tflops"]
                r["step_frac_issued_of_measured_tf32_sustained"] = r["step_tflops"] * passes / tpk["sustained_tflops"]
            if peaks_live and math == "bf16":
                r["bf16_peak_this_run"] = {"burst_tflops": peaks_live["bf16"]["burst_tflops"], "sustained_tflops": peaks_live["bf16"]["sustained_tflops"]}
            return r

        line = {
            "metric": "training_frames_per_sec", "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": args.scaling if world > 1 else "weak",
            "vs_baseline": None, "dtype": {"3xtf32": "tf32x3", "tf32": "tf32", "bf16": "bf16"}[args.math], "data": "synthetic",
            "config": {"workload": WORKLOAD, "dims": DIMS, "bunch_per_gpu": bunch, "global_bunch": bunch * world,
                       "parallelism": "dp%d" % world, "dp_schedule": dp_state["schedule"],
                       "dp_gradient_transport": (None if world == 1 or not str(dp_state["schedule"]).startswith("peer") else
                                                 {"0": "owner pulls (peer loads)", "1": "GEMM-epilogue peer stores, every layer",
                                                  "2": "fp32, copy engines for the upper layers + GEMM-epilogue peer stores for the %s lowest"
                                                       % os.environ.get("TNB_DP_PUSH_TAIL", "2")}.get(os.environ.get("TNB_DP_PUSH", "2"), "?")),
                       "learn_rate": LR, "momentum": MMT, "weightcost": WC,
                       "l2_note": ("no L2 flush: weights+corrections (224 MB) and activations exceed the 126 MB L2 every step" if args.config == "C" else
                                   "no L2 flush: the model (2-4 MB) is L2-resident by nature; the 16 resident bunches rotate"),
                       "timed_region": "K x (row window of the resident set, int labels -> one-hot, forward, softmax+xent+accuracy, backward, "
                                       "update); the cache shuffle and the splice run once per cache fill and are outside it (~5 us per bunch amortised)",
                       "windows_ms": win_ms, "windows": "median of %d windows of %d bunches" % (len(win_ms), args.steps),
                       "scaling_note": "weak = 1024 frames per GPU; strong = the 1024-frame bunch partitioned, rows [g*B/G,(g+1)*B/G) per GPU. "
                                       "The 85 % efficiency target is read on the weak curve: at a fixed 1024-frame global bunch the 112 MB fp32 "
                                       "gradient exchange per bunch exceeds the per-GPU compute at 8 GPUs by construction (SURVEY 8e); the "
                                       "strong curve is reported beside it",
                       "flops_per_frame": fpf, "gemm_math": args.math},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": bunch * DIMS[0] * 4 + bunch * 4,
                    "d2h_bytes_per_step": 24, "windows_s": e2e_s, "submissions_in_flight": E2E_DEPTH,
                    # rank 0's host time per step inside the two calls: enqueueing a step vs waiting for the previous one's statistics
                    "host_ms_per_step": {"submit": 1000.0 * host_t["submit"] / max(1, args.steps - E2E_DEPTH),
                                         "collect_wait": 1000.0 * host_t["collect"] / max(1, args.steps - E2E_DEPTH)}},
            "gpu_launches": int(launches),
            "roofline": roof(args.math, gms, gl, gfl, ms / args.steps, bunch),
            "final_stats": {"xent_per_frame": st[0] / max(1, st[1]), "frames": st[1]},
        }
        for k, (b2, m2, w2, l2, g2) in extras.items():
            math2 = "bf16" if k == "bf16" else args.math
            obj = {"value": args.steps * b2 * world / (m2 / 1000.0), "unit": "frames/s", "ms_per_step": m2 / args.steps,
                   "bunch_per_gpu": b2, "global_bunch": b2 * world, "windows_ms": w2, "gpu_launches": int(l2),
                   "roofline": roof(math2, g2[0], g2[1], g2[2], m2 / args.steps, b2)}
            if k == "bf16":
                obj["dtype"] = "bf16"
                obj["note"] = "north_star's bf16 mode, reported separately: bf16 operands (resident twins), fp32 accumulation, fp32 master weights"
                line["bf16"] = obj
            else:
                obj["scaling"] = k
                line[k + "_scaling"] = obj
        if world == 1 and not args.no_cpu_baseline:
            try:
                fps, threads, sample = cpu_reference_throughput(2, 10) if args.config == "C" else cpu_reference_throughput(40, 240)
                line["cpu_baseline"] = {"value": fps, "unit": "frames/s", "cores": threads, "kind": "reference", "sample": sample}
            except Exception as e:
                line["cpu_baseline"] = {"value": None, "unit": "frames/s", "cores": 0, "kind": "reference", "sample": "failed: %s" % str(e)[:200]}
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        barrier()
        abi.check(L.tnb_comm_destroy(ctx))
        dist.destroy_process_group()
    return 0


def host_stream(L, ctx):
    s = C.c_void_p()
    from tnet_b200 import abi
    abi.check(L.tnb_ctx_stream(ctx, C.byref(s)))
    return s.value


if __name__ == "__main__":
    sys.exit(main())
