#!/usr/bin/env python
"""Does the H2D rate depend on WHICH page-locked allocation a copy reads?  16 separate pinned tensors of
the C2 reference matrix's size against 16 slices of one pinned slab; per buffer the median of 10 timed
copies (CUDA events), alone on the machine.

    python tools/exp_pinned_slots.py
"""
import json
import sys

import torch


def rate(dst, src, reps=10):
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        dst.copy_(src, non_blocking=True)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return round(src.numel() / (ts[len(ts) // 2] * 1e-3) / 1e9, 1)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 3_263_900
    dev = torch.device("cuda", 0)
    dst = torch.empty(n, dtype=torch.uint8, device=dev)
    base = torch.randint(0, 3, (n,), dtype=torch.uint8)
    sep = [base.clone().pin_memory() for _ in range(16)]
    slab = torch.empty(16 * n, dtype=torch.uint8).pin_memory()
    slices = [slab[i * n:(i + 1) * n] for i in range(16)]
    for s in slices:
        s.copy_(base)
    out = {"bytes": n, "separate_GBps": [rate(dst, s) for s in sep], "slab_slices_GBps": [rate(dst, s) for s in slices]}
    # the same slices once the CPU's caches hold something else (1 GiB written by the CPU)
    junk = torch.empty(1 << 30, dtype=torch.uint8)
    junk.fill_(1)
    out["slab_slices_after_cpu_cache_flush_GBps"] = [rate(dst, s) for s in slices]
    for s in slices:
        s.copy_(base)
    out["slab_slices_rewritten_GBps"] = [rate(dst, s) for s in slices]
    # back to back on one stream: 16 copies, one event pair
    for name, bufs in (("separate", sep), ("slab_slices", slices)):
        ts = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for b in bufs:
                dst.copy_(b, non_blocking=True)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        out[f"{name}_16_back_to_back_GBps"] = round(16 * n / (ts[2] * 1e-3) / 1e9, 1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
