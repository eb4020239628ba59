import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from sstar_b200.engine import get_engine
wk = bench.make_workload("c2", 1); st = bench.workload_stats(wk)
dev = torch.device("cuda", 0); eng = get_engine(0)
pin = {k: torch.from_numpy(np.ascontiguousarray(wk[k])).pin_memory() for k in ("ref", "tgt", "pos", "ws", "we")}
d = {k: torch.empty_like(v, device=dev) for k, v in pin.items()}
W, T, R = st["W"], st["T"], st["R"]
score_d = torch.empty((W, T), dtype=torch.int64, device=dev); nsnp_d = torch.empty((W, T), dtype=torch.int32, device=dev)
score_p = torch.empty((W, T), dtype=torch.int64).pin_memory(); nsnp_p = torch.empty((W, T), dtype=torch.int32).pin_memory()
s1 = torch.cuda.current_stream(); s2 = torch.cuda.Stream()
hint = st["max_window_variants"]
def run(evs):
    evs[0].record(s1)
    s2.wait_stream(s1)
    for k in ("pos", "ws", "we"): d[k].copy_(pin[k], non_blocking=True)
    evs[1].record(s1)
    with torch.cuda.stream(s2):
        d["ref"].copy_(pin["ref"], non_blocking=True); evs[5].record(s2)
        d["tgt"].copy_(pin["tgt"], non_blocking=True); evs[2].record(s2)
    s1.wait_stream(s2)
    eng.pack_device(d["ref"], d["tgt"], W=W, max_win_variants=hint); evs[3].record(s1)
    eng.score_packed_device(R, d["tgt"], d["pos"], d["ws"], d["we"], 5000, 5, -10000, max_win_variants=hint, out=(score_d, nsnp_d)); evs[4].record(s1)
    score_p.copy_(score_d, non_blocking=True); nsnp_p.copy_(nsnp_d, non_blocking=True); evs[6].record(s1)
for it in range(6):
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(7)]
    torch.cuda.synchronize(); t0 = time.perf_counter(); run(evs); torch.cuda.synchronize(); wall = time.perf_counter() - t0
    if it >= 3:
        f = lambda a, b: evs[a].elapsed_time(evs[b]) * 1e3
        print(f"wall {wall*1e6:.0f} us | small copies done +{f(0,1):.0f} | ref copy done +{f(0,5):.0f} | tgt copy done +{f(0,2):.0f} | pack done +{f(0,3):.0f} | score done +{f(0,4):.0f} | D2H done +{f(0,6):.0f}")
