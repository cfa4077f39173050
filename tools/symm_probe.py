#!/usr/bin/env python
"""Probe (run under torchrun, one rank per GPU) of what torch's symmetric-memory plumbing offers on this box, for the multicast
variant of the peer-memory update kernel planned in DESIGN.md 10.1(a): peer pointers of a symmetric allocation, the NVSwitch
multicast pointer (multimem.ld_reduce / multimem.st target) and the signal pad.  Prints one line per rank; changes nothing.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 tools/symm_probe.py
"""
import os

import torch
import torch.distributed as dist


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    info = {"rank": rank, "world": world, "torch": torch.__version__}
    try:
        import torch.distributed._symmetric_memory as symm_mem
        n = 64 << 20  # bytes
        t = symm_mem.empty(n // 4, dtype=torch.float32, device=torch.device("cuda", local))
        hdl = symm_mem.rendezvous(t, dist.group.WORLD.group_name)
        for name in ("buffer_ptrs", "multicast_ptr", "signal_pad_ptrs", "buffer_size", "signal_pad_size", "rank", "world_size"):
            try:
                v = getattr(hdl, name)
                info[name] = [hex(p) for p in v] if isinstance(v, (list, tuple)) else (hex(v) if "ptr" in name else v)
            except Exception as e:  # the attribute set differs between torch versions
                info[name] = "n/a (%s)" % type(e).__name__
        try:
            info["has_multicast_support"] = bool(type(hdl).has_multicast_support(torch.device("cuda", local).type, local))
        except Exception as e:
            info["has_multicast_support"] = "n/a (%s: %s)" % (type(e).__name__, str(e)[:80])
        # sanity: a value written by every rank into its own buffer is visible through the peers' mapped tensors
        t.fill_(float(rank + 1))
        hdl.barrier()
        peer = (rank + 1) % world
        seen = float(hdl.get_buffer(peer, (4,), torch.float32)[0].item())
        info["peer_read_ok"] = seen == float(peer + 1)
        hdl.barrier()
    except Exception as e:
        info["error"] = "%s: %s" % (type(e).__name__, str(e)[:300])
    print(info, flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
