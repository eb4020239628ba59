"""How much of the C2 pass is cold-cache cost?  Same graph replay with and without the L2 flush between steps."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from sstar_b200.engine import get_engine
name = sys.argv[1] if len(sys.argv) > 1 else "c2"
wk = bench.make_workload(name, 1)
dev = torch.device("cuda", 0); eng = get_engine(0)
case = bench.DeviceCase(torch, eng, dev, wk)
big = torch.empty(256 << 20, dtype=torch.uint8, device=dev); tiny = torch.empty(1, dtype=torch.uint8, device=dev)
for _ in range(20): case.step_device()
sync = lambda: torch.cuda.synchronize()
for label, fl in (("L2 flushed", big), ("warm (no flush)", tiny), ("L2 flushed", big)):
    t = case.timed(case.step_device, 300, fl, sync)
    print(f"{name} {label:16s}: pass {t/300*1e3:.2f} us")
