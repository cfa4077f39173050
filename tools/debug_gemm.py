#!/usr/bin/env python
"""GEMM bring-up helper (run on a B200): one case per subprocess so a trapping kernel cannot poison the next case.
Prints max error against numpy float64 and, when wrong, a coarse map of which 32x32 blocks of C are off."""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "nnet-asr_b200", "python"))

CASES = [
    # ta tb M N K math
    ("N", "T", 128, 128, 32, 1), ("N", "T", 128, 128, 64, 1), ("N", "T", 128, 128, 256, 1),
    ("N", "N", 128, 128, 32, 1), ("N", "N", 128, 128, 256, 1),
    ("T", "N", 128, 128, 32, 1), ("T", "N", 128, 128, 256, 1), ("T", "T", 128, 128, 64, 1),
    ("N", "T", 128, 64, 64, 1), ("N", "N", 128, 256 * 148 * 2, 64, 1),
    ("N", "T", 128, 128, 256, 0), ("N", "N", 128, 128, 256, 0), ("T", "N", 128, 128, 256, 0),
    ("N", "N", 300, 200, 100, 0), ("N", "N", 1024, 2048, 2048, 0), ("T", "N", 2048, 2048, 1024, 0),
]


def one(ta, tb, M, N, K, math):
    from tnet_b200 import abi
    ctx = abi.Context(0, math)
    r = np.random.default_rng(1)
    A = r.standard_normal((K, M) if ta == "T" else (M, K)).astype(np.float32)
    B = r.standard_normal((N, K) if tb == "T" else (K, N)).astype(np.float32)
    dA, dB, dC = abi.DMat.from_numpy(ctx, A), abi.DMat.from_numpy(ctx, B), abi.DMat(ctx, M, N)
    abi.gemm(ctx, ta, tb, 1.0, dA, dB, 0.0, dC)
    got = dC.download().astype(np.float64)
    ref = (A.T if ta == "T" else A).astype(np.float64) @ (B.T if tb == "T" else B).astype(np.float64)
    err = np.abs(got - ref)
    tol = (3e-2 if math == 1 else 1e-4) * np.sqrt(K)
    print("%s%s M=%d N=%d K=%d math=%d  max_err=%.3e  frac_bad=%.4f" % (ta, tb, M, N, K, math, err.max(), (err > tol).mean()))
    if err.max() > tol and M <= 512 and N <= 512:
        if M % 32 == 0 and N % 32 == 0:
            bm = (err > tol).reshape(M // 32, 32, N // 32, 32)
            print((bm.mean(axis=(1, 3)) > 0).astype(int))
        print("got[0,:8]", got[0, :8], "\nref[0,:8]", ref[0, :8])


if __name__ == "__main__":
    if len(sys.argv) > 1:
        a = sys.argv[1:]
        one(a[0], a[1], int(a[2]), int(a[3]), int(a[4]), int(a[5]))
    else:
        for c in CASES:
            res = subprocess.run([sys.executable, __file__] + [str(x) for x in c], stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                                 text=True, timeout=300)
            print(res.stdout.strip()[-1500:] or ("rc=%d (no output)" % res.returncode), flush=True)
