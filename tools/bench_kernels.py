#!/usr/bin/env python
"""Achieved HBM GB/s of the memory-bound kernels of the hot path at config-C sizes, against the measured copy bandwidth of
MEASURED_PEAKS.json (north_star: "achieved HBM GB/s for the elementwise, softmax and shuffle kernels").  Each kernel is run
back to back over buffers larger than L2 where the shape allows it (the cache gather) or with the rotating-buffer trick
(NBUF copies of the operands, > 126 MB in total) so that every launch reads from DRAM."""
import ctypes as C, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "nnet-asr_b200", "python"))
from tnet_b200 import abi

L = abi.lib()
ctx = abi.Context(0)
pk = 6552.6
p = os.path.join(ROOT, "MEASURED_PEAKS.json")
if os.path.exists(p):
    pk = json.load(open(p)).get("hbm_gbs", pk)
r = np.random.default_rng(0)
out = []


ITERS_SCALE = float(os.environ.get("TNB_BENCH_ITERS_SCALE", "1"))   # < 1 for short runs under ncu


def timeit(name, nbytes, fns, iters=40):
    """fns: list of closures over rotating buffer sets"""
    iters = max(1, int(iters * ITERS_SCALE))
    for f in fns:
        f()
    ctx.sync()
    t = time.perf_counter()
    for i in range(iters):
        fns[i % len(fns)]()
    ctx.sync()
    us = (time.perf_counter() - t) / iters * 1e6
    gbs = nbytes / us / 1e3
    out.append((name, nbytes / 1e6, us, gbs, gbs / pk))
    print("%-58s %8.1f MB  %7.1f us  %7.0f GB/s  %5.1f %% of %.0f" % (name, nbytes / 1e6, us, gbs, 100 * gbs / pk, pk), flush=True)


rows, nout, H = 1024, 3000, 2048
NB = 6
# fused softmax + cross-entropy + accuracy: read a, t; write y, err
sets = []
for _ in range(NB):
    A = abi.DMat.from_numpy(ctx, r.standard_normal((rows, nout)).astype(np.float32))
    T = abi.DMat(ctx, rows, nout); Y = abi.DMat(ctx, rows, nout); E = abi.DMat(ctx, rows, nout)
    lab = abi.DMat.from_numpy(ctx, r.integers(0, nout, (1, rows)).astype(np.int32))
    abi.check(L.tnb_onehot(ctx.h, T.p(), lab.p(C.c_int), T.dim))
    sets.append((A, T, Y, E, lab))
st = abi.DStats(ctx)
timeit("softmax+xent+accuracy 1024x3000 (tnb_softmax_xent)", 4 * rows * nout * 4,
       [lambda s=s: abi.check(L.tnb_softmax_xent(ctx.h, s[0].p(), s[1].p(), s[2].p(), s[3].p(), s[0].dim, st.p())) for s in sets])
timeit("one-hot targets 1024x3000 (tnb_onehot)", rows * nout * 4,
       [lambda s=s: abi.check(L.tnb_onehot(ctx.h, s[1].p(), s[4].p(C.c_int), s[1].dim)) for s in sets])
# elementwise on a hidden layer
sets = [(abi.DMat.from_numpy(ctx, r.random((4 * rows, H)).astype(np.float32)), abi.DMat(ctx, 4 * rows, H), abi.DMat(ctx, 4 * rows, H)) for _ in range(NB)]
timeit("sigmoid 4096x2048 (tnb_sigmoid)", 2 * 4 * rows * H * 4, [lambda s=s: abi.check(L.tnb_sigmoid(ctx.h, s[1].p(), s[0].p(), s[0].dim)) for s in sets])
timeit("diff-sigmoid 4096x2048 (tnb_diff_sigmoid)", 3 * 4 * rows * H * 4,
       [lambda s=s: abi.check(L.tnb_diff_sigmoid(ctx.h, s[2].p(), s[1].p(), s[0].p(), s[0].dim)) for s in sets])
vec = abi.DMat(ctx, 1, H)
timeit("column sums 4096x2048 (tnb_add_col_sum)", 4 * rows * H * 4, [lambda s=s: abi.check(L.tnb_add_col_sum(ctx.h, C.c_float(1.0), s[0].p(), C.c_float(0.0), vec.p(), s[0].dim)) for s in sets])
d16 = [abi.DMat16(ctx, 4 * rows, H) for _ in range(NB)]
timeit("fp32 -> bf16 twin 4096x2048 (tnb_to_bf16)", 4 * rows * H * 6, [lambda s=s, d=d: abi.check(L.tnb_to_bf16(ctx.h, d.p(), C.c_int(d.stride), s[0].p(), s[0].dim)) for s, d in zip(sets, d16)])
# SGD update from a summed gradient: read G, corr, W; write corr, W
sets = [(abi.DMat.from_numpy(ctx, (0.01 * r.standard_normal((H, H))).astype(np.float32)), abi.DMat(ctx, H, H), abi.DMat(ctx, H, H)) for _ in range(NB)]
timeit("SGD update 2048x2048 (tnb_sgd_update)", 5 * H * H * 4,
       [lambda s=s: abi.check(L.tnb_sgd_update(ctx.h, s[0].p(), s[1].p(), s[2].p(), s[1].dim, None, None, None, C.c_float(0.008), C.c_float(0.5),
                                               C.c_float(1e-6), C.c_int(1), C.c_int(1024))) for s in sets])
del sets, d16
# CuCache::Randomize: gather 131072 rows of 429 features and of 3000 targets (one permutation)
cache = 131072
perm = abi.DMat.from_numpy(ctx, r.permutation(cache).astype(np.int32).reshape(1, -1))
for cols, name in ((429, "features"), (3000, "targets")):
    src = abi.DMat(ctx, cache, cols); dst = abi.DMat(ctx, cache, cols)
    timeit("cache shuffle %d x %d %s (tnb_randomize)" % (cache, cols, name), 2 * cache * cols * 4,
           [lambda: abi.check(L.tnb_randomize(ctx.h, dst.p(), src.p(), perm.p(C.c_int), dst.dim, src.dim))], iters=10)
    del src, dst
# splice: [T x 39] -> [T x 429] with 11 clamped offsets
Tn = 262144
offs = abi.DMat.from_numpy(ctx, np.arange(-5, 6, dtype=np.int32).reshape(1, -1))
x = abi.DMat.from_numpy(ctx, r.standard_normal((Tn, 39)).astype(np.float32)); y = abi.DMat(ctx, Tn, 429)
timeit("splice %d x 39 -> x 429 (tnb_expand)" % Tn, Tn * (39 + 429) * 4, [lambda: abi.check(L.tnb_expand(ctx.h, y.p(), x.p(), offs.p(C.c_int), y.dim, x.dim))], iters=10)
print(json.dumps([{"kernel": n, "MB": mb, "us": us, "GBps": g, "frac_of_hbm_peak": f} for n, mb, us, g, f in out]))
