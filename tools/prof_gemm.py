#!/usr/bin/env python
"""Run one of the three fused CuBiasedLinearity GEMMs of config C a few times (a target for `ncu --set full -k regex:gemm_tcgen05`
and for wall-clock A/B comparisons).  usage: prof_gemm.py {fwd|dx|upd|all} [rows nin nout] [iters] [math]"""
import ctypes as C, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "nnet-asr_b200", "python"))
from tnet_b200 import abi

a = sys.argv[1:]
which = a[0] if a else "all"
rows, nin, nout = (int(a[1]), int(a[2]), int(a[3])) if len(a) >= 4 else (1024, 2048, 2048)
iters = int(a[4]) if len(a) >= 5 else 20
math = {"3x": abi.MATH_3XTF32, "tf32": abi.MATH_TF32, "bf16": getattr(abi, "MATH_BF16", 3)}[a[5] if len(a) >= 6 else "3x"]
ctx = abi.Context(0, math)
L = abi.lib()
r = np.random.default_rng(0)
X = abi.DMat.from_numpy(ctx, r.random((rows, nin)).astype(np.float32))
E = abi.DMat.from_numpy(ctx, (r.standard_normal((rows, nout)) * 0.01).astype(np.float32))
W = abi.DMat.from_numpy(ctx, (0.1 * r.standard_normal((nin, nout))).astype(np.float32))
cW = abi.DMat(ctx, nin, nout)
b = abi.DMat.from_numpy(ctx, np.zeros(nout, np.float32))
cb = abi.DMat(ctx, 1, nout)
Y = abi.DMat(ctx, rows, nout)
Ep = abi.DMat(ctx, rows, nin)


def fwd():
    abi.check(L.tnb_affine_fwd(ctx.h, X.p(), X.dim, W.p(), W.dim, b.p(), Y.p(), Y.dim, C.c_int(abi.ACT_SIGMOID)))


def dx():
    abi.check(L.tnb_affine_bwd_dx(ctx.h, E.p(), E.dim, W.p(), W.dim, X.p(), X.dim, Ep.p(), Ep.dim))


def upd():
    abi.check(L.tnb_affine_update(ctx.h, X.p(), X.dim, E.p(), E.dim, W.p(), W.dim, b.p(), cW.p(), cb.p(), C.c_float(0.008),
                                  C.c_float(0.5), C.c_float(1e-6), C.c_int(1), C.c_int(0)))


for name, fn in (("fwd", fwd), ("dx", dx), ("upd", upd)):
    if which not in (name, "all"):
        continue
    for _ in range(3):
        fn()
    ctx.sync()
    t = time.perf_counter()
    for _ in range(iters):
        fn()
    ctx.sync()
    us = (time.perf_counter() - t) / iters * 1e6
    fl = 2.0 * rows * nin * nout
    print("%-4s rows=%d nin=%d nout=%d: %.1f us/call back-to-back, %.1f TFLOP/s algorithmic" % (name, rows, nin, nout, us, fl / us / 1e6))
