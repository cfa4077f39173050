import ctypes as C, os, sys, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "nnet-asr_b200", "python"))
from tnet_b200 import abi
ctx = abi.Context(0, abi.MATH_3XTF32 if len(sys.argv) < 2 or sys.argv[1] == "3x" else abi.MATH_TF32)
r = np.random.default_rng(0)
ta, tb, M, N, K = (sys.argv[2], sys.argv[3], int(sys.argv[4]), int(sys.argv[5]), int(sys.argv[6])) if len(sys.argv) > 6 else ("N", "N", 1024, 2048, 2048)
A = r.standard_normal((K, M) if ta == "T" else (M, K)).astype(np.float32); B = r.standard_normal((N, K) if tb == "T" else (K, N)).astype(np.float32)
dA, dB, dC = abi.DMat.from_numpy(ctx, A), abi.DMat.from_numpy(ctx, B), abi.DMat(ctx, M, N)
import time
for _ in range(3):
    abi.gemm(ctx, ta, tb, 1.0, dA, dB, 0.0, dC)
ctx.sync()
t = time.perf_counter()
for _ in range(20):
    abi.gemm(ctx, ta, tb, 1.0, dA, dB, 0.0, dC)
ctx.sync()
print("kernel wall (20 back-to-back): %.1f us each" % ((time.perf_counter() - t) / 20 * 1e6))
ts = np.zeros(8 * 256, np.int64)
abi.lib().tnb_dbg_read_ts(ts.ctypes.data_as(C.c_void_p))
nkb = min(64, (K + 31) // 32)
ts = ts.reshape(8, 256)[:, :nkb]
t0 = ts[0, 0]
e=ts[1]; print("entry->setup %d, setup->first empty_seen %d, last commit->epilogue start %d, epilogue %d, epilogue end->final sync %d, TOTAL entry->final %d cycles" % (e[1]-e[0], ts[0,0]-e[1], e[2]-ts[5,nkb-1], e[3]-e[2], e[4]-e[3], e[4]-e[0]))
names = ["prod:empty_seen", "-", "conv:full_seen", "conv:arrived", "mma:conv_seen", "mma:committed"]
print("kb  " + "  ".join("%15s" % n for i, n in enumerate(names) if i != 1))
for kb in list(range(0, 6)) + list(range(nkb - 3, nkb)):
    print("%2d  " % kb + "  ".join("%15d" % (ts[i, kb] - t0) for i in range(6) if i != 1))
d = np.diff(ts[5, 8:nkb-2]); print("steady-state cycles per k-block (mma commit to commit): mean %.0f min %d max %d" % (d.mean(), d.min(), d.max()))
print("full_seen - empty_seen (TMA latency incl. issue):", (ts[2, 8:nkb-2] - ts[0, 8:nkb-2]).mean())
print("arrived - full_seen (conversion):", (ts[3, 8:nkb-2] - ts[2, 8:nkb-2]).mean())
print("mma conv_seen - conv arrived:", (ts[4, 8:nkb-2] - ts[3, 8:nkb-2]).mean())
print("mma committed - conv_seen (issue):", (ts[5, 8:nkb-2] - ts[4, 8:nkb-2]).mean())
