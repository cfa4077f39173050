#!/usr/bin/env python
"""Timeline of ONE tnb_gemm_batch launch shaped like a hidden layer of config C's backward pass (dX 1024x2048x2048 mandatory +
weight-update tiles 2048x2048x1024 from the pool), from the kernel's own clock64() stamps (TNB_BATCH_TRACE=1).
   MATH=bf16|3xtf32  python tools/dbg/batch_timeline.py"""
import ctypes as C, os, sys
import numpy as np
os.environ["TNB_BATCH_TRACE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "nnet-asr_b200", "python"))
from tnet_b200 import abi

L = abi.lib()
bf = os.environ.get("MATH", "3xtf32") == "bf16"
ctx = abi.Context(0, abi.MATH_BF16 if bf else abi.MATH_3XTF32)
r = np.random.default_rng(0)
rows, H = 1024, 2048
mode = os.environ.get("MODE", "layer")   # layer: dX + pool updates ; upd: updates only ; dx: dX only
E = abi.DMat.from_numpy(ctx, (0.1 * r.standard_normal((rows, H))).astype(np.float32))
W = abi.DMat.from_numpy(ctx, (0.1 * r.standard_normal((H, H))).astype(np.float32))
Yp = abi.DMat.from_numpy(ctx, r.random((rows, H)).astype(np.float32))
Ep = abi.DMat(ctx, rows, H)
X2 = abi.DMat.from_numpy(ctx, r.standard_normal((rows, H)).astype(np.float32))
W2 = abi.DMat.from_numpy(ctx, (0.1 * r.standard_normal((H, H))).astype(np.float32))
c2 = abi.DMat(ctx, H, H)
W3 = abi.DMat.from_numpy(ctx, (0.1 * r.standard_normal((H, H))).astype(np.float32))
c3 = abi.DMat(ctx, H, H)
jobs = (abi.GemmJob * 3)()
abi.check(L.tnb_job_affine_bwd_dx(C.byref(jobs[0]), E.p(), E.dim, W.p(), W.dim, Yp.p(), Yp.dim, Ep.p(), Ep.dim))
abi.check(L.tnb_job_affine_update(C.byref(jobs[1]), X2.p(), X2.dim, E.p(), E.dim, W2.p(), W2.dim, c2.p(), C.c_float(0.008), C.c_float(0.5),
                                  C.c_float(1e-6), C.c_int(1), C.c_int(0)))
abi.check(L.tnb_job_affine_update(C.byref(jobs[2]), X2.p(), X2.dim, E.p(), E.dim, W3.p(), W3.dim, c3.p(), C.c_float(0.008), C.c_float(0.5),
                                  C.c_float(1e-6), C.c_int(1), C.c_int(0)))
if bf:
    tw = {k: abi.DMat16.from_fp32(ctx, m) for k, m in (("E", E), ("W", W), ("X2", X2), ("W2", W2), ("W3", W3))}
    Ep16 = abi.DMat16(ctx, rows, H)
    abi.check(L.tnb_job_set_twins(C.byref(jobs[0]), tw["E"].p(), C.c_int(tw["E"].stride), tw["W"].p(), C.c_int(tw["W"].stride), Ep16.p(), C.c_int(Ep16.stride), None, C.c_int(0)))
    abi.check(L.tnb_job_set_twins(C.byref(jobs[1]), tw["X2"].p(), C.c_int(tw["X2"].stride), tw["E"].p(), C.c_int(tw["E"].stride), None, C.c_int(0), tw["W2"].p(), C.c_int(tw["W2"].stride)))
    abi.check(L.tnb_job_set_twins(C.byref(jobs[2]), tw["X2"].p(), C.c_int(tw["X2"].stride), tw["E"].p(), C.c_int(tw["E"].stride), None, C.c_int(0), tw["W3"].p(), C.c_int(tw["W3"].stride)))


def launch():
    pool = (abi.GemmJob * 2)(jobs[1], jobs[2])
    if mode == "layer":
        abi.check(L.tnb_gemm_batch(ctx.h, C.byref(jobs[0]), C.c_int(1), pool, C.c_int(2)))
    elif mode == "upd":
        abi.check(L.tnb_gemm_batch(ctx.h, None, C.c_int(0), pool, C.c_int(1)))
    else:
        abi.check(L.tnb_gemm_batch(ctx.h, C.byref(jobs[0]), C.c_int(1), None, C.c_int(0)))
    return pool


for _ in range(3):
    pool = launch()
ctx.sync()
import time
t = time.perf_counter()
for _ in range(20):
    launch()
ctx.sync()
print("mode %s %s: %.1f us per launch (host wall over 20 back-to-back launches); pool left %d + %d tiles" % (
    mode, "bf16" if bf else "3xtf32", (time.perf_counter() - t) / 20 * 1e6, pool[0].tile_count, pool[1].tile_count))
buf = np.zeros((128, 64), np.int64)
abi.check(L.tnb_gemm_batch_trace_read(ctx.h, buf.ctypes.data_as(C.POINTER(C.c_longlong))))
print("pair: set-up | per tile [TMA first..last | MMA first operands .. last issue | accumulator complete .. epilogue done] | exit   (cycles from entry)")
for pr in range(128):
    t0 = buf[pr, 0]
    if t0 == 0:
        continue
    if pr not in (0, 1, 31, 32, 33, 50, 73) and pr % 16:
        continue
    s = "pair %3d: %6d |" % (pr, buf[pr, 1] - t0)
    for i in range(8):
        if buf[pr, 8 + 6 * i] == 0:
            break
        s += " [%6d..%6d | %6d..%6d | %6d..%6d]" % tuple(buf[pr, k + 6 * i] - t0 for k in (12, 13, 8, 9, 10, 11))
    s += " | %6d" % (buf[pr, 2] - t0)
    print(s)
ex = buf[:, 2] - buf[:, 0]
ex = ex[buf[:, 0] != 0]
print("exit - entry over %d pairs: min %d median %d max %d cycles" % (len(ex), ex.min(), np.median(ex), ex.max()))
