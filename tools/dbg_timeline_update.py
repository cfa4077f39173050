import ctypes as C, os, sys, time, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "nnet-asr_b200", "python"))
from tnet_b200 import abi
L = abi.lib()
ctx = abi.Context(0, abi.MATH_3XTF32)
r = np.random.default_rng(0)
rows, nin, nout = 1024, 2048, 2048
X = r.standard_normal((rows, nin)).astype(np.float32); E = (0.1 * r.standard_normal((rows, nout))).astype(np.float32)
W = (0.1 * r.standard_normal((nin, nout))).astype(np.float32); b = np.zeros(nout, np.float32)
dX, dE, dW, db, dcW, dcb = [abi.DMat.from_numpy(ctx, a) for a in (X, E, W, b, np.zeros_like(W), np.zeros_like(b))]
def call():
    abi.check(L.tnb_affine_update(ctx.h, dX.p(), dX.dim, dE.p(), dE.dim, dW.p(), dW.dim, db.p(), dcW.p(), dcb.p(),
                                  C.c_float(0.01), C.c_float(0.5), C.c_float(1e-6), C.c_int(1), C.c_int(0)))
for _ in range(3): call()
ctx.sync(); t = time.perf_counter()
for _ in range(20): call()
ctx.sync(); print("affine_update wall: %.1f us each (GEMM + colsum + bias kernels)" % ((time.perf_counter() - t) / 20 * 1e6))
ts = np.zeros(8 * 256, np.int64); L.tnb_dbg_read_ts(ts.ctypes.data_as(C.c_void_p)); ts = ts.reshape(8, 256)
nkb = 32; e = ts[1]
print("entry->setup %d, mainloop %d, last commit->epilogue start %d, epilogue %d, end->final %d, TOTAL %d cycles" % (e[1]-e[0], ts[5,nkb-1]-ts[0,0], e[2]-ts[5,nkb-1], e[3]-e[2], e[4]-e[3], e[4]-e[0]))
