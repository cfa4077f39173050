#!/usr/bin/env python
"""Dense tensor-core peaks of this box, measured the way MEASURED_PEAKS.json's bf16 figure was (torch.matmul N^3, CUDA events):
best of 10 launches (burst) and back to back for a few seconds (sustained), for TF32 (fp32 operands, allow_tf32) and bf16.
bench.py runs the same probe (probe_dense_peak) so that the 3xTF32 roofline is divided by a measured TF32 peak, not by an
assumed bf16/2.  Library GEMM used as a yardstick only; nothing in the product calls it."""
import json
import sys
import time

import torch


def probe_dense_peak(dtype="tf32", n=8192, burst_iters=10, sustain_s=2.0, device="cuda"):
    torch.backends.cuda.matmul.allow_tf32 = True
    dt = torch.float32 if dtype == "tf32" else torch.bfloat16
    a = torch.randn(n, n, device=device, dtype=dt)
    b = torch.randn(n, n, device=device, dtype=dt)
    c = torch.empty(n, n, device=device, dtype=dt)
    for _ in range(3):
        torch.matmul(a, b, out=c)
    torch.cuda.synchronize()
    flops = 2.0 * n ** 3
    best = 0.0
    for _ in range(burst_iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b, out=c)
        e1.record()
        e1.synchronize()
        best = max(best, flops / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    # sustained: back-to-back launches for sustain_s seconds
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iters = 0
    t0 = time.perf_counter()
    e0.record()
    while True:
        for _ in range(20):
            torch.matmul(a, b, out=c)
        iters += 20
        torch.cuda.synchronize()
        if time.perf_counter() - t0 > sustain_s:
            break
    e1.record()
    e1.synchronize()
    sustained = flops * iters / (e0.elapsed_time(e1) * 1e-3) / 1e12
    del a, b, c
    return dict(dtype=dtype, n=n, burst_tflops=best, sustained_tflops=sustained)


if __name__ == "__main__":
    out = [probe_dense_peak("tf32"), probe_dense_peak("bf16")]
    json.dump(out, sys.stdout)
    print()
