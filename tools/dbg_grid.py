#!/usr/bin/env python
"""Grid-level timing of one GEMM launch (needs a `make TRACE=1` build): per-CTA %globaltimer at entry / epilogue start / exit.
Shows how much of the kernel's duration is launch skew, mainloop, epilogue and tail.  usage: dbg_grid.py [ta tb M N K [epi]]"""
import ctypes as C, os, sys, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "nnet-asr_b200", "python"))
import torch
from tnet_b200 import abi
import os as _os
MATH = {"3x": abi.MATH_3XTF32, "tf32": abi.MATH_TF32, "bf16": abi.MATH_BF16}[_os.environ.get("DBG_MATH", "3x")]
ctx = abi.Context(0, MATH)
r = np.random.default_rng(0)
a = sys.argv[1:]
ta, tb, M, N, K = (a[0], a[1], int(a[2]), int(a[3]), int(a[4])) if len(a) >= 5 else ("N", "N", 1024, 2048, 2048)
A = r.standard_normal((K, M) if ta == "T" else (M, K)).astype(np.float32); B = r.standard_normal((N, K) if tb == "T" else (K, N)).astype(np.float32)
dA, dB, dC = abi.DMat.from_numpy(ctx, A), abi.DMat.from_numpy(ctx, B), abi.DMat(ctx, M, N)
for _ in range(3):
    abi.gemm(ctx, ta, tb, 1.0, dA, dB, 0.0, dC)
ctx.sync()
import time
t = time.perf_counter()
for _ in range(20):
    abi.gemm(ctx, ta, tb, 1.0, dA, dB, 0.0, dC)
ctx.sync()
wall = (time.perf_counter() - t) / 20 * 1e6
ts = np.zeros(4 * 1024, np.int64)
best = None
for tag in ("cg1_3x", "cg1_1x", "cg2", "split", "bf16_a", "bf16_b"):   # one trace buffer per instantiation unit: take the latest
    t_ = np.zeros(4 * 1024, np.int64)
    getattr(abi.lib(), "tnb_dbg_read_cta_" + tag)(t_.ctypes.data_as(C.c_void_p))
    if best is None or t_.max() > best.max():
        best = t_
ts = best
ts = ts.reshape(1024, 4)
ts = ts[ts[:, 0] > 0]
t0 = ts[:, 0].min()
ent, epi, ext = (ts[:, 0] - t0) / 1e3, (ts[:, 1] - t0) / 1e3, (ts[:, 2] - t0) / 1e3
print("%s%s M=%d N=%d K=%d: %d CTAs on %d SMs, back-to-back wall %.1f us/launch" % (ta, tb, M, N, K, len(ts), len(set(ts[:, 3])), wall))
pc = lambda x: "min %.1f p50 %.1f p90 %.1f max %.1f" % (x.min(), np.percentile(x, 50), np.percentile(x, 90), x.max())
print("entry  (us after first CTA): " + pc(ent))
print("epilogue start             : " + pc(epi))
print("exit                       : " + pc(ext))
print("per-CTA mainloop (entry->epilogue) " + pc(epi - ent) + "; epilogue+teardown " + pc(ext - epi))
