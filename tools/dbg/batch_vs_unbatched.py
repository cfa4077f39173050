#!/usr/bin/env python
"""Debug aid: the batched backward schedule (tnb_gemm_batch) against the one-launch-per-GEMM schedule on the same inputs, layer by
layer and weight by weight, in both math modes.  Differences beyond summation-order noise point at the batch path."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "nnet-asr_b200", "python"))
from tnet_b200 import abi, host, formats as F

dims = [int(x) for x in os.environ.get("DIMS", "429,320,256,3000").split(",")]
bunch = int(os.environ.get("BUNCH", "256"))
steps = int(os.environ.get("STEPS", "2"))
for math in (abi.MATH_BF16, abi.MATH_3XTF32):
    host.set_math(math)
    r = np.random.default_rng(7)
    layers = F.gen_mlp_init(dims, r)
    X = r.standard_normal((bunch, dims[0])).astype(np.float32)
    T = np.zeros((bunch, dims[-1]), np.float32)
    T[np.arange(bunch), r.integers(0, dims[-1], bunch)] = 1
    nets = []
    for batching in (True, False):
        n = host.Net(layers)
        n.set_batching(batching)
        n.set_hyper(0.1, mmt=0.5, wc=1e-4, gdf=True)
        nets.append(n)
    for step in range(steps):
        for n in nets:
            n.train_bunch(X, T)
        print("math %d step %d" % (math, step))
        for i in range(len(layers)):
            if layers[i][0] == "affine" and i + 1 < len(layers) and layers[i + 1][0] == "sigmoid":
                continue
            a, b = nets[0].layer_out(i, bunch), nets[1].layer_out(i, bunch)
            print("  out %d: max |diff| %.3e of max %.3e" % (i, np.abs(a - b).max(), np.abs(b).max()))
        for i in range(1, len(layers)):
            kind = layers[i][0]
            if kind == "softmax" or (kind == "affine" and layers[i - 1][0] == "sigmoid"):
                continue
            a, b = nets[0].layer_eout(i, bunch), nets[1].layer_eout(i, bunch)
            print("  eout %d: max |diff| %.3e of max %.3e" % (i, np.abs(a - b).max(), np.abs(b).max()))
        for i in range(0, len(layers), 2):
            (Wa, ba), (Wb, bb) = nets[0].get_affine_raw(i), nets[1].get_affine_raw(i)
            d = np.abs(Wa - Wb)
            bad_rows = np.where(d.max(axis=0) > 1e-6)[0]   # Wt is [nout x nin]: axis 0 -> per input row of W
            print("  W %d: max |diff| %.3e (|dW| %.3e), bias %.3e; input rows off: %s" % (
                i, d.max(), np.abs(Wb - layers[i][1]).max(), np.abs(ba - bb).max(),
                ("%d..%d (%d)" % (bad_rows.min(), bad_rows.max(), len(bad_rows))) if len(bad_rows) else "none"))
