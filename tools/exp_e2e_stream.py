#!/usr/bin/env python
"""How fast can a STREAM of C2 chromosomes go end to end (pinned host rows -> scores in pinned host
tables) when consecutive calls overlap: H2D copies of call k+1 on a copy stream under the kernel
and the D2H copies of call k.  Plain torch copies + the device entry point (upper-bound probe for a
pipelined host entry).

    python tools/exp_e2e_stream.py [ring] [steps]
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
    case.step_device()
    torch.cuda.synchronize()
    want = case.score_d.cpu()
    W, T = case.W, case.T
    keys = ("ref", "tgt", "pos", "ws", "we")
    pins = [{k: case.pin[k].clone().pin_memory() for k in keys} for _ in range(ring)]
    devs = [{k: torch.empty_like(case.d[k]) for k in keys} for _ in range(ring)]
    outs_d = [(torch.empty((W, T), dtype=torch.int64, device=dev), torch.empty((W, T), dtype=torch.int32, device=dev))
              for _ in range(ring)]
    outs_p = [(torch.empty((W, T), dtype=torch.int64).pin_memory(), torch.empty((W, T), dtype=torch.int32).pin_memory())
              for _ in range(ring)]
    copy_s, comp_s, back_s = torch.cuda.Stream(dev), torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    ev_in = [torch.cuda.Event() for _ in range(ring)]
    ev_done = [torch.cuda.Event() for _ in range(ring)]   # the kernel of slot i has consumed its inputs
    ev_out = [torch.cuda.Event() for _ in range(ring)]
    res = {}
    for mode in ("one_buffer", "split"):
        for rep in range(2):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for i in range(steps):
                s = i % ring
                with torch.cuda.stream(copy_s):
                    if i >= ring:
                        copy_s.wait_event(ev_done[s])
                    for k in keys:
                        devs[s][k].copy_(pins[s][k], non_blocking=True)
                    ev_in[s].record(copy_s)
                with torch.cuda.stream(comp_s):
                    comp_s.wait_event(ev_in[s])
                    if i >= ring:
                        comp_s.wait_event(ev_out[s])
                    d = devs[s]
                    eng.score_device(d["ref"], d["tgt"], d["pos"], d["ws"], d["we"], *bench.PARAMS,
                                     max_win_variants=case.hint, out=outs_d[s], stream=comp_s,
                                     flags=_cabi.FLAG_INDEPENDENT)
                    ev_done[s].record(comp_s)
                bs = back_s if mode == "split" else comp_s
                with torch.cuda.stream(bs):
                    bs.wait_event(ev_done[s])
                    outs_p[s][0].copy_(outs_d[s][0], non_blocking=True)
                    outs_p[s][1].copy_(outs_d[s][1], non_blocking=True)
                    ev_out[s].record(bs)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
        ok = all(torch.equal(o[0], want) for o in outs_p)
        res[mode] = {"us_per_step": dt / steps * 1e6, "GBps_h2d": case.V * (case.R + case.T + 4) / (dt / steps) / 1e9,
                     "tables_ok": ok}
        print(mode, res[mode], flush=True)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
