"""How slow is the all-pairwise route (what windows with missing / multi-allelic calls take)?"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from sstar_b200.engine import get_engine
name = sys.argv[1] if len(sys.argv) > 1 else "c2"
wk = bench.make_workload(name, 1); st = bench.workload_stats(wk)
dev = torch.device("cuda", 0); eng = get_engine(0)
d = {k: torch.from_numpy(np.ascontiguousarray(wk[k])).to(dev) for k in ("ref", "tgt", "pos", "ws", "we")}
res = {}
for algo in ("auto", "pairwise"):
    for _ in range(3):
        out = eng.score_device(d["ref"], d["tgt"], d["pos"], d["ws"], d["we"], 5000, 5, -10000, algo=algo,
                               max_win_variants=st["max_window_variants"])
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    torch.cuda.synchronize(); ev[0].record()
    for _ in range(10):
        out = eng.score_device(d["ref"], d["tgt"], d["pos"], d["ws"], d["we"], 5000, 5, -10000, algo=algo,
                               max_win_variants=st["max_window_variants"])
    ev[1].record(); torch.cuda.synchronize()
    res[algo] = (ev[0].elapsed_time(ev[1]) / 10, out[0].clone(), out[1].clone())
print(f"{name}: auto {res['auto'][0]*1e3:.1f} us, pairwise {res['pairwise'][0]*1e3:.1f} us per pass "
      f"({res['pairwise'][0]/res['auto'][0]:.1f}x); identical: "
      f"{bool(torch.equal(res['auto'][1], res['pairwise'][1]) and torch.equal(res['auto'][2], res['pairwise'][2]))}")

# the same chromosome with 0.5 % of the target calls missing (-1): the general per-class DP
tgt_m = d["tgt"].clone()
mask = torch.rand(tgt_m.shape, device=dev) < 0.005
tgt_m[mask] = -1
for _ in range(3):
    out = eng.score_device(d["ref"], tgt_m, d["pos"], d["ws"], d["we"], 5000, 5, -10000, max_win_variants=st["max_window_variants"])
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
torch.cuda.synchronize(); ev[0].record()
for _ in range(10):
    out = eng.score_device(d["ref"], tgt_m, d["pos"], d["ws"], d["we"], 5000, 5, -10000, max_win_variants=st["max_window_variants"])
ev[1].record(); torch.cuda.synchronize()
t_gen = ev[0].elapsed_time(ev[1]) / 10
from oracle import c_oracle
sel = np.unique(np.linspace(0, st["W"] - 1, 24).astype(int))
so, no = c_oracle.score_windows(wk["ref"], tgt_m.cpu().numpy(), wk["pos"], wk["ws"][sel], wk["we"][sel])
ok = np.array_equal(out[0].cpu().numpy()[sel], so) and np.array_equal(out[1].cpu().numpy()[sel], no)
print(f"{name} with 0.5% missing target calls: {t_gen*1e3:.1f} us per pass ({t_gen/res['auto'][0]:.2f}x the clean data), "
      f"bit-exact vs oracle on {len(sel)} windows: {ok}")
