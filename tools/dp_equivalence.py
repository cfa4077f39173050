#!/usr/bin/env python
"""Run under torchrun with WORLD_SIZE ranks (one per GPU): the data-parallel CUDA path (rows of every bunch split over the
ranks, per-layer exchange of [dW;db] — peer-memory kernel, NCCL all-reduce or NCCL reduce-scatter/all-gather — and update with
the global frame count) must reproduce the single-GPU run of the full bunch.  Rank 0 prints DP_EQUIV_OK.
DP_EQUIV_MODES=peer,allreduce,shard selects the schedules (default: all three)."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "nnet-asr_b200", "python"))
import torch
import torch.distributed as dist
from tnet_b200 import abi, formats as F, host

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
host.select_gpu(local)
L = abi.lib()
ctx = host.ctx_handle()
idbuf = (C.c_ubyte * 128)()
if rank == 0:
    abi.check(L.tnb_comm_unique_id(idbuf))
t = torch.tensor(list(bytes(idbuf)), dtype=torch.uint8, device="cuda")
dist.broadcast(t, 0)
idbuf = (C.c_ubyte * 128)(*t.cpu().tolist())
abi.check(L.tnb_comm_init(ctx, idbuf, C.c_int(rank), C.c_int(world)))

r = np.random.default_rng(9)
dims, B, steps = [429, 512, 256, 300], 256 * world, 3
layers = F.gen_mlp_init(dims, r)
X = r.standard_normal((steps, B, dims[0])).astype(np.float32)
lab = r.integers(0, dims[-1], (steps, B)).astype(np.int32)
T = np.zeros((steps, B, dims[-1]), np.float32)
for s in range(steps):
    T[s, np.arange(B), lab[s]] = 1
modes = os.environ.get("DP_EQUIV_MODES", "allreduce,shard,peer").split(",")
for mode in modes:      # the data-parallel schedules of CuNetwork::Backpropagate (TNB_DP_MODE); "peer" = csrc/peer.cu, no NCCL on the data path
    for mname, math in (("3xtf32", abi.MATH_3XTF32), ("bf16", abi.MATH_BF16)):
        os.environ["TNB_DP_MODE"] = mode
        host.set_math(math)
        net = host.Net(layers)
        net.set_hyper(0.5, mmt=0.5, wc=1e-4, gdf=True)
        net.set_data_parallel(world)
        rows = slice(rank * B // world, (rank + 1) * B // world)
        for s in range(steps):
            net.train_bunch(X[s, rows], T[s, rows])
        e, fr, co = net.stats()
        tot = torch.tensor([e, fr, co], dtype=torch.float64, device="cuda")
        dist.all_reduce(tot)
        got = net.get_layers()
        ok = True
        if rank == 0:
            ref = host.Net(layers)
            ref.set_hyper(0.5, mmt=0.5, wc=1e-4, gdf=True)
            for s in range(steps):
                ref.train_bunch(X[s], T[s])
            re_, rfr, rco = ref.stats()
            want = ref.get_layers()
            for a, b in zip(got, want):
                if a[0] == "affine":
                    np.testing.assert_allclose(a[1], b[1], rtol=2e-4, atol=2e-5 * np.abs(b[1]).max())
                    np.testing.assert_allclose(a[2], b[2], rtol=2e-4, atol=2e-5 * max(1e-2, np.abs(b[2]).max()))
            assert int(tot[1].item()) == rfr and abs(tot[0].item() - re_) <= 1e-4 * abs(re_) and abs(int(tot[2].item()) - rco) <= 2, (tot, re_, rfr, rco)
            print("dp ok: schedule=%s math=%s world=%d xent=%.4f frames=%d" % (mode, mname, world, tot[0].item(), int(tot[1].item())), flush=True)
        # every rank must hold the same weights
        chk = torch.tensor([float(np.abs(got[0][1]).sum())], dtype=torch.float64, device="cuda")
        lo, hi = chk.clone(), chk.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        assert lo.item() == hi.item(), "ranks diverged"
host.set_math(abi.MATH_3XTF32)
if rank == 0:
    print("DP_EQUIV_OK world=%d" % world)
abi.check(L.tnb_comm_destroy(ctx))
dist.destroy_process_group()
