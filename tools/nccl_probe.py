#!/usr/bin/env python
"""Under torchrun: time NCCL all-reduce / reduce-scatter / all-gather of one layer's and the whole model's gradient (fp32 and bf16)
with nothing else running — the communication roofline of the data-parallel step on this box."""
import os, torch, torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
def timeit(fn, n=20):
    for _ in range(5): fn()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
for dt in (torch.float32, torch.bfloat16):
    for elems in (2048 * 2048, 27994112):
        x = torch.ones(elems // world * world, dtype=dt, device="cuda")
        shard = torch.empty(elems // world, dtype=dt, device="cuda")
        t_ar = timeit(lambda: dist.all_reduce(x))
        t_rs = timeit(lambda: dist.reduce_scatter_tensor(shard, x))
        t_ag = timeit(lambda: dist.all_gather_into_tensor(x, shard))
        if rank == 0:
            mb = x.numel() * x.element_size() / 1e6
            print("world %d %s %.1f MB: all-reduce %.0f us (algbw %.0f GB/s), reduce-scatter %.0f us, all-gather %.0f us" % (world, str(dt)[6:], mb, t_ar, mb / t_ar * 1e3, t_rs, t_ag), flush=True)
dist.destroy_process_group()
