#!/usr/bin/env python
"""A STREAM of independent C2 passes (one chromosome after the other) through ONE graph:
serialized launches vs SSTAR_B200_FLAG_INDEPENDENT (consecutive one-launch kernels overlap), with
one or two CTAs per SM.  The ring of input copies exceeds the L2 (ring x 4.9 MB), so every pass
reads its rows from HBM.  Every pass's tables are compared with a single flushed pass.

    python tools/exp_stream.py [ring] [replays]
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ring = int(sys.argv[1]) if len(sys.argv) > 1 else 48
    replays = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    import torch
    import __graft_entry__ as entry
    entry.build()
    import bench
    from sstar_b200 import _cabi
    from sstar_b200.engine import get_engine

    dev = torch.device("cuda", 0)
    eng = get_engine(0)
    wk = bench.make_workload(os.environ.get("WL", "c2"), 1)
    case = bench.ShardCase(torch, eng, dev, wk, use_graph=True)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    flush.fill_(1)
    case.step_device()
    torch.cuda.synchronize()
    want_s, want_n = case.score_d.clone(), case.nsnp_d.clone()
    W, T = case.W, case.T
    units = W * T
    d = case.d
    cases = []
    for _ in range(ring):
        cases.append((d["ref"].clone(), d["tgt"].clone(), d["pos"].clone(), d["ws"].clone(), d["we"].clone(),
                      (torch.empty((W, T), dtype=torch.int64, device=dev), torch.empty((W, T), dtype=torch.int32, device=dev))))
    out = {"ring": ring, "replays": replays, "ring_bytes": ring * (case.V * (case.R + case.T) + 4 * case.V)}

    def run(tag, plan):
        for c in cases:
            c[5][0].zero_(); c[5][1].zero_()
        for _ in range(3):
            plan.run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        flush.fill_(0)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(replays):
            plan.run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / (replays * ring)
        ok = all(torch.equal(c[5][0], want_s) and torch.equal(c[5][1], want_n) for c in cases)
        out[tag] = {"us_per_pass": ms * 1e3, "G_units_per_s": units / ms / 1e6, "all_tables_equal": ok,
                    "launches": plan.launches}
        print(tag, out[tag], flush=True)

    # serialized: the same graph without the flag
    torch_graph = eng.plan_device_stream(cases, *bench.PARAMS, max_win_variants=case.hint, flags=0)
    # (plan_device_stream always sets the flag: build the serialized one by hand)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for c in cases:
            eng.score_device(*c[:5], *bench.PARAMS, max_win_variants=case.hint, out=c[5], flags=0)

    class P:
        launches = ring

        def run(self):
            g.replay()
    run("serialized", P())
    run("independent_pair" + os.environ.get("SSTAR_B200_FUSED_PAIR", "1"), torch_graph)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
