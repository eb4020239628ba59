"""Experiment: end-to-end host call variants.  python tools/exp_e2e.py [workload]"""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from sstar_b200.engine import get_engine

name = sys.argv[1] if len(sys.argv) > 1 else "c2"
wk = bench.make_workload(name, 1)
st = bench.workload_stats(wk)
dev = torch.device("cuda", 0)
eng = get_engine(0)
pin = {k: torch.from_numpy(np.ascontiguousarray(wk[k])).pin_memory() for k in ("ref", "tgt", "pos", "ws", "we")}
W, T = st["W"], st["T"]
score_p = torch.empty((W, T), dtype=torch.int64).pin_memory()
nsnp_p = torch.empty((W, T), dtype=torch.int32).pin_memory()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
ref_out = None
for label, kw in [("staged", dict(staged=True)), ("pipelined (auto)", dict()), ("pipelined, copy engine inputs", dict(flags=64)),
                  ("pipelined, zero-copy inputs", dict(flags=32)), ("pipelined+copy_out", dict(flags=4)),
                  ("pipelined+single_chunk", dict(flags=8))]:
    ts = []
    for i in range(25):
        flush.fill_(i & 1)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        eng.score_pinned(pin["ref"], pin["tgt"], pin["pos"], pin["ws"], pin["we"], score_p, nsnp_p,
                         5000, 5, -10000, max_win_variants=st["max_window_variants"], sync=True, **kw)
        ts.append(time.perf_counter() - t0)
    out = (score_p.clone(), nsnp_p.clone())
    if ref_out is None:
        ref_out = out
    same = bool(torch.equal(out[0], ref_out[0]) and torch.equal(out[1], ref_out[1]))
    ts = sorted(ts[5:])
    print(f"{name} {label:34s} median {ts[len(ts)//2]*1e6:9.1f} us  min {ts[0]*1e6:9.1f} us  same={same}")
