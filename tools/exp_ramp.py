#!/usr/bin/env python
"""How the bench's two stream regions depend on the number of steps K (the driver runs K = 20):
total time of K device-resident passes (CUDA events) and of K host-entry calls (wall clock), five
repetitions each, so that the fixed cost (ramp + drain) and the slope can be read apart.

    python tools/exp_ramp.py [workload]
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import __graft_entry__ as entry
    entry.build()
    import bench
    from sstar_b200.engine import get_engine

    dev = torch.device("cuda", 0)
    eng = get_engine(0)
    wk = bench.make_workload(sys.argv[1] if len(sys.argv) > 1 else "c2", 1)
    case = bench.ShardCase(torch, eng, dev, wk)
    sync = lambda: torch.cuda.synchronize(dev)
    for _ in range(3):
        case.step_device(); case.step_e2e()
    case.make_stream()
    case.run_stream(2 * case.ring)
    case.make_e2e_stream(int(os.environ.get("RING", "4")))
    case.run_e2e_stream(8)
    if os.environ.get("FLUSH_HOST"):
        junk = torch.empty(1 << 30, dtype=torch.uint8)
        junk.fill_(1)
        print("host caches flushed (1 GiB written)", flush=True)
    out = {"device_stream_us": {}, "e2e_stream_us": {}, "e2e_enqueue_us": {}}
    for k in [int(x) for x in os.environ.get("KS", "1,2,5,10,20,40,80,200,1000").split(",")]:
        case.timed_stream(k, sync)
        out["device_stream_us"][k] = sorted(round(case.timed_stream(k, sync)[0] * 1e3, 1) for _ in range(5))
        reps, enq = [], []
        for _ in range(5):
            sync()
            t0 = time.perf_counter()
            from sstar_b200 import _cabi
            n = len(case.e2e_slots)
            for i in range(k):
                p, (sp, npn) = case.e2e_slots[i % n]
                eng.score_pinned(p["ref"], p["tgt"], p["pos"], p["ws"], p["we"], sp, npn, *bench.PARAMS,
                                 max_win_variants=case.hint, sync=False, flags=_cabi.FLAG_INDEPENDENT, scratch_slot=1 + i % n)
            t1 = time.perf_counter()
            sync()
            t2 = time.perf_counter()
            reps.append(round((t2 - t0) * 1e6, 1)); enq.append(round((t1 - t0) * 1e6, 1))
        out["e2e_stream_us"][k] = sorted(reps)
        out["e2e_enqueue_us"][k] = sorted(enq)
        print(k, out["device_stream_us"][k], out["e2e_stream_us"][k], out["e2e_enqueue_us"][k], flush=True)
    # completion time of every call of one burst of 48 (an event behind each call on the caller's stream)
    from sstar_b200 import _cabi
    n = len(case.e2e_slots)
    for rep in range(2):
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(49)]
        sync()
        evs[0].record()
        for i in range(48):
            p, (sp, npn) = case.e2e_slots[i % n]
            eng.score_pinned(p["ref"], p["tgt"], p["pos"], p["ws"], p["we"], sp, npn, *bench.PARAMS,
                             max_win_variants=case.hint, sync=False, flags=_cabi.FLAG_INDEPENDENT, scratch_slot=1 + i % n)
            evs[i + 1].record()
        sync()
        done = [evs[0].elapsed_time(e) * 1e3 for e in evs[1:]]
        out[f"burst48_call_us_{rep}"] = [round(b - a, 1) for a, b in zip([0.0] + done[:-1], done)]
        print("per-call completion deltas (us):", out[f"burst48_call_us_{rep}"], flush=True)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
