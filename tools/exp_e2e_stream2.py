#!/usr/bin/env python
"""A STREAM of C2 chromosomes end to end through the host entry (sstar_b200_score_windows_host with
SSTAR_B200_FLAG_INDEPENDENT): ring of pinned inputs / scratch buffers / pinned tables, one sync at the end.

    python tools/exp_e2e_stream2.py [ring] [steps]
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ring = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 400
    import torch
    import __graft_entry__ as entry
    entry.build()
    import bench
    from sstar_b200 import _cabi
    from sstar_b200.engine import get_engine

    dev = torch.device("cuda", 0)
    eng = get_engine(0)
    wk = bench.make_workload("c2", 1)
    case = bench.ShardCase(torch, eng, dev, wk, use_graph=False)
    case.step_e2e()
    want_s, want_n = case.score_p.clone(), case.nsnp_p.clone()
    W, T = case.W, case.T
    keys = ("ref", "tgt", "pos", "ws", "we")
    pins = [{k: case.pin[k].clone().pin_memory() for k in keys} for _ in range(ring)]
    outs = [(torch.zeros((W, T), dtype=torch.int64).pin_memory(), torch.zeros((W, T), dtype=torch.int32).pin_memory())
            for _ in range(ring)]
    res = {}
    for rep in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(steps):
            s = i % ring
            p = pins[s]
            eng.score_pinned(p["ref"], p["tgt"], p["pos"], p["ws"], p["we"], outs[s][0], outs[s][1], *bench.PARAMS,
                             max_win_variants=case.hint, sync=False, flags=_cabi.FLAG_INDEPENDENT, scratch_slot=s + 1)
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        ok = all(torch.equal(o[0], want_s) and torch.equal(o[1], want_n) for o in outs)
        res = {"us_per_step": dt / steps * 1e6, "enqueue_us_per_step": (t1 - t0) / steps * 1e6,
               "GBps_h2d": case.V * (case.R + case.T + 4) / (dt / steps) / 1e9, "tables_ok": ok, "launches": eng.launches}
        print(res, flush=True)
        for o in outs:
            o[0].zero_(); o[1].zero_()
    print(json.dumps(res))


if __name__ == "__main__":
    main()
