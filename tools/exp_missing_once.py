"""A few passes over a workload with missing target calls (profiling aid for the signed / general-genotype bodies):
    python tools/exp_missing_once.py [workload] [rate, default 0.005]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from sstar_b200.engine import get_engine
name = sys.argv[1] if len(sys.argv) > 1 else "c3-shard"
wk = bench.make_workload(name, 1); st = bench.workload_stats(wk)
dev = torch.device("cuda", 0); eng = get_engine(0)
d = {k: torch.from_numpy(np.ascontiguousarray(wk[k])).to(dev) for k in ("ref", "tgt", "pos", "ws", "we")}
torch.manual_seed(1)
tgt_m = d["tgt"].clone()
rate = float(sys.argv[2]) if len(sys.argv) > 2 else 0.005
tgt_m[torch.rand(tgt_m.shape, device=dev) < rate] = -1
for _ in range(5):
    out = eng.score_device(d["ref"], tgt_m, d["pos"], d["ws"], d["we"], 5000, 5, -10000, max_win_variants=st["max_window_variants"])
torch.cuda.synchronize()
print("done", int((out[0] != np.iinfo(np.int64).min).sum()))
