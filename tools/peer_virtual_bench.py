#!/usr/bin/env python
"""dp_peer_update_kernel at config-C layer size (2048 x 2048) with 2 / 4 / 8 virtual ranks on ONE GPU, as one cooperative grid
(tnb_dp_peer_update_virtual).  The 'peer' loads and stores hit local HBM here, so this shows the kernel's own structure (loads in
flight, CTAs per rank) — the NVLink numbers come from the multi-GPU bench.  Used for the ncu capture of this kernel."""
import ctypes as C, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "nnet-asr_b200", "python"))
from tnet_b200 import abi

L = abi.lib()
ctx = abi.Context(0)
rows = cols = int(os.environ.get("PEER_BENCH_DIM", "2048"))
iters = int(os.environ.get("PEER_BENCH_ITERS", "10"))
r = np.random.default_rng(0)
for world in [int(w) for w in os.environ.get("PEER_BENCH_WORLDS", "2,4,8").split(",")]:
    for ctas in [int(c) for c in os.environ.get("PEER_BENCH_CTAS", "18").split(",")]:
        W0 = (0.1 * r.standard_normal((rows, cols))).astype(np.float32)
        Wd = [abi.DMat.from_numpy(ctx, W0) for _ in range(world)]
        Kd = [abi.DMat(ctx, rows, cols) for _ in range(world)]
        bd = [abi.DMat(ctx, 1, cols) for _ in range(world)]
        kbd = [abi.DMat(ctx, 1, cols) for _ in range(world)]
        Gd = [abi.DMat.from_numpy(ctx, (0.01 * r.standard_normal((rows + 1, cols))).astype(np.float32)) for _ in range(world)]
        flags = [abi.DMat(ctx, 1, 64, np.uint32) for _ in range(world)]
        fl = (C.POINTER(C.c_uint) * world)(*[f.p(C.c_uint) for f in flags])
        jobs = (abi.PeerJob * world)()
        for k in range(world):
            for q in range(world):
                jobs[k].G[q] = Gd[q].ptr.value
                jobs[k].W[q] = Wd[q].ptr.value
            jobs[k].corrW, jobs[k].bias, jobs[k].corrb = Kd[k].ptr.value, bd[k].ptr.value, kbd[k].ptr.value
            jobs[k].dW = abi.MatrixDim(rows, cols, Wd[k].stride)
            jobs[k].rows_pad, jobs[k].lr, jobs[k].mmt, jobs[k].wc, jobs[k].grad_div_frm, jobs[k].n_frames = rows, 0.008, 0.5, 1e-6, 1, 1024 * world
        seq = 0
        for _ in range(2):
            seq += 1
            abi.check(L.tnb_dp_peer_update_virtual(ctx.h, jobs, C.c_int(world), fl, C.c_uint(seq), C.c_int(ctas)))
        ctx.sync()
        t = time.perf_counter()
        for _ in range(iters):
            seq += 1
            abi.check(L.tnb_dp_peer_update_virtual(ctx.h, jobs, C.c_int(world), fl, C.c_uint(seq), C.c_int(ctas)))
        ctx.sync()
        us = (time.perf_counter() - t) / iters * 1e6
        # per rank: reads world gradient blocks + momentum + weights of its block, writes momentum + world weight blocks
        per_rank = (rows // world) * cols * 4 * (world + 2 + 1 + world)
        print("world %d, %d CTAs per rank: %.1f us per layer (host wall, synchronous launches), %.0f GB/s over all %d virtual ranks"
              % (world, ctas, us, world * per_rank / us / 1e3, world), flush=True)
        for m in Wd + Kd + bd + kbd + Gd + flags:
            m.free()
