"""Scoring pass (launch pipeline) against the rate of missing target calls (-1): even a handful per window group
flags the groups and sends their windows through the general body, so low rates cost the body's overhead on
plain units, high rates the per-class DP.    python tools/exp_missing_rates.py [c3-shard|c4-shard]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from sstar_b200.engine import get_engine
from oracle import c_oracle
name = sys.argv[1] if len(sys.argv) > 1 else "c3-shard"
wk = bench.make_workload(name, 1); st = bench.workload_stats(wk)
dev = torch.device("cuda", 0); eng = get_engine(0)
d = {k: torch.from_numpy(np.ascontiguousarray(wk[k])).to(dev) for k in ("ref", "tgt", "pos", "ws", "we")}
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
sel = np.unique(np.linspace(0, st["W"] - 1, 16).astype(int))
base = None
for rate in (0.0, 5e-6, 5e-5, 5e-4, 5e-3):
    torch.manual_seed(1)
    tgt = d["tgt"].clone()
    if rate:
        tgt[torch.rand(tgt.shape, device=dev) < rate] = -1
    run = lambda: eng.score_device(d["ref"], tgt, d["pos"], d["ws"], d["we"], 5000, 5, -10000, max_win_variants=st["max_window_variants"])
    for _ in range(3):
        out = run()
    ts = []
    for i in range(10):
        flush.fill_(i & 1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = run(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    t = float(np.median(ts))
    base = base or t
    so, no = c_oracle.score_windows(wk["ref"], tgt.cpu().numpy(), wk["pos"], wk["ws"][sel], wk["we"][sel])
    ok = np.array_equal(out[0].cpu().numpy()[sel], so) and np.array_equal(out[1].cpu().numpy()[sel], no)
    print(f"{name} missing rate {rate:g}: {t:.1f} us per pass ({t / base:.2f}x clean), bit-exact on {len(sel)} windows: {ok}", flush=True)
