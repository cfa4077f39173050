#!/usr/bin/env python
"""Hardware probe (run on a B200): does tcgen05.mma.kind::tf32 truncate or round the low 13 mantissa bits of fp32
operands read from shared memory?  Decides whether the 3xTF32 converter must rewrite the `hi` operand."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "nnet-asr_b200", "python"))
from tnet_b200 import abi

ctx = abi.Context(0, abi.MATH_TF32)
r = np.random.default_rng(0)
M, N, K = 128, 128, 64
A = r.standard_normal((M, K)).astype(np.float32)
B = r.standard_normal((N, K)).astype(np.float32)   # used as B^T: both operands K-major


def run(a, b):
    dA, dB, dC = abi.DMat.from_numpy(ctx, a), abi.DMat.from_numpy(ctx, b), abi.DMat(ctx, M, N)
    abi.gemm(ctx, "N", "T", 1.0, dA, dB, 0.0, dC)
    return dC.download()


def trunc(x):
    return (x.view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)


def rna(x):
    u = x.view(np.uint32).astype(np.uint64) + 0x1000
    return (u & 0xFFFFE000).astype(np.uint32).view(np.float32)


raw = run(A, B)
t = run(trunc(A), trunc(B))
n = run(rna(A), rna(B))
print("raw == truncated inputs :", np.array_equal(raw, t), float(np.abs(raw - t).max()))
print("raw == rounded inputs   :", np.array_equal(raw, n), float(np.abs(raw - n).max()))
ref = A.astype(np.float64) @ B.astype(np.float64).T
print("max abs err raw/trunc/rna vs fp64:", float(np.abs(raw - ref).max()), float(np.abs(t - ref).max()), float(np.abs(n - ref).max()))
