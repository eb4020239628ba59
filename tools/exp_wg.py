"""Device-resident pass time vs windows per score CTA (SSTAR_B200_WG experiment hook), graph replay, L2 flushed."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from sstar_b200.engine import get_engine
name = sys.argv[1] if len(sys.argv) > 1 else "c2"
wk = bench.make_workload(name, 0)
dev = torch.device("cuda", 0); eng = get_engine(0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for wg in (0, 1, 2, 3, 4, 8, 16):
    if wg: os.environ["SSTAR_B200_WG"] = str(wg)
    else: os.environ.pop("SSTAR_B200_WG", None)
    case = bench.DeviceCase(torch, eng, dev, wk)
    for _ in range(20): case.step_device()
    tot = case.timed(case.step_device, 300, flush, lambda: torch.cuda.synchronize())
    sc = case.timed(case.step_score, 100, flush, lambda: torch.cuda.synchronize())
    print(f"{name} wg={wg:2d} (0 = heuristic): pass {tot/300*1e3:.2f} us, score+pairwise alone {sc/100*1e3:.2f} us")
