#!/usr/bin/env python
"""Debug aid: the weight-update job (TN, EPI_UPDATE) through tnb_gemm_batch against tnb_affine_update(_bf16) for a list of shapes."""
import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "nnet-asr_b200", "python"))
from tnet_b200 import abi
L = abi.lib()
for math in (abi.MATH_BF16, abi.MATH_3XTF32):
    ctx = abi.Context(0, math)
    bf = math == abi.MATH_BF16
    for rows, nin, nout in [(256, 429, 320), (256, 512, 320), (256, 429, 512), (1000, 429, 320), (256, 320, 256), (256, 256, 3000)]:
        r = np.random.default_rng(rows + nin + nout)
        X = r.standard_normal((rows, nin)).astype(np.float32)
        E = (0.1 * r.standard_normal((rows, nout))).astype(np.float32)
        W0 = (0.1 * r.standard_normal((nin, nout))).astype(np.float32)
        dX_, dE_ = abi.DMat.from_numpy(ctx, X), abi.DMat.from_numpy(ctx, E)
        res = []
        for which in ("single", "batch"):
            W, cW = abi.DMat.from_numpy(ctx, W0), abi.DMat(ctx, nin, nout)
            if bf:
                X16, E16, W16 = abi.DMat16.from_fp32(ctx, dX_), abi.DMat16.from_fp32(ctx, dE_), abi.DMat16.from_fp32(ctx, W)
            if which == "single":
                if bf:
                    abi.check(L.tnb_affine_update_bf16(ctx.h, X16.p(), C.c_int(X16.stride), dX_.dim, E16.p(), C.c_int(E16.stride), dE_.p(), dE_.dim,
                                                       W.p(), W.dim, W16.p(), C.c_int(W16.stride), None, cW.p(), None, C.c_float(0.1), C.c_float(0.5),
                                                       C.c_float(1e-4), C.c_int(1), C.c_int(0)))
                else:
                    abi.check(L.tnb_affine_update(ctx.h, dX_.p(), dX_.dim, dE_.p(), dE_.dim, W.p(), W.dim, None, cW.p(), None, C.c_float(0.1),
                                                  C.c_float(0.5), C.c_float(1e-4), C.c_int(1), C.c_int(0)))
            else:
                job = abi.GemmJob()
                abi.check(L.tnb_job_affine_update(C.byref(job), dX_.p(), dX_.dim, dE_.p(), dE_.dim, W.p(), W.dim, cW.p(), C.c_float(0.1), C.c_float(0.5),
                                                  C.c_float(1e-4), C.c_int(1), C.c_int(0)))
                if bf:
                    abi.check(L.tnb_job_set_twins(C.byref(job), X16.p(), C.c_int(X16.stride), E16.p(), C.c_int(E16.stride), None, C.c_int(0), W16.p(), C.c_int(W16.stride)))
                print("   batch_ok", L.tnb_gemm_batch_ok(ctx.h, C.byref(job)), "tiles", job.tile_count)
                abi.check(L.tnb_gemm_batch(ctx.h, C.byref(job), C.c_int(1), None, C.c_int(0)))
            ctx.sync()
            res.append((W.download(), cW.download()))
        dW = np.abs(res[0][0] - res[1][0]); dc = np.abs(res[0][1] - res[1][1])
        bad = np.argwhere(dc > 1e-5 * np.abs(res[0][1]).max())
        print("math %d rows %d nin %d nout %d: |W diff| %.3e  |corr diff| %.3e of %.3e ; bad corr elements %d, rows %s cols %s" % (
            math, rows, nin, nout, dW.max(), dc.max(), np.abs(res[0][1]).max(), len(bad),
            (bad[:, 0].min(), bad[:, 0].max()) if len(bad) else "-", (bad[:, 1].min(), bad[:, 1].max()) if len(bad) else "-"))
    ctx.close()
