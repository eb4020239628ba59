"""Cost of recovering the S* SNP sets (scores + backpointers + ragged fill) next to plain scoring."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from sstar_b200.engine import get_engine
name = sys.argv[1] if len(sys.argv) > 1 else "c2"
wk = bench.make_workload(name, 1); st = bench.workload_stats(wk)
dev = torch.device("cuda", 0); eng = get_engine(0)
d = {k: torch.from_numpy(np.ascontiguousarray(wk[k])).to(dev) for k in ("ref", "tgt", "pos", "ws", "we")}
hint = st["max_window_variants"]
def timed(fn, reps=10):
    for _ in range(3): out = fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): out = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps, out
t_score, o1 = timed(lambda: eng.score_device(d["ref"], d["tgt"], d["pos"], d["ws"], d["we"], max_win_variants=hint))
t_pair, _ = timed(lambda: eng.score_device(d["ref"], d["tgt"], d["pos"], d["ws"], d["we"], algo="pairwise", max_win_variants=hint))
t_sets, o2 = timed(lambda: eng.snp_sets_device(d["ref"], d["tgt"], d["pos"], d["ws"], d["we"], max_win_variants=hint))
same = bool(torch.equal(o1[0], o2[0]) and torch.equal(o1[1], o2[1]))
units = st["W"] * wk["tgt"].shape[1]
print(f"{name}: {units} units; score {t_score*1e6:.1f} us, pairwise score {t_pair*1e6:.1f} us, "
      f"scores + SNP sets {t_sets*1e6:.1f} us (wall, incl. the host read of the total; "
      f"{int(o2[2][-1])} chain sites, {units/t_sets/1e6:.1f} M units/s); scores identical: {same}")
