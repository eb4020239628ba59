"""In-place (zero-copy) pack rate vs the number of pack CTAs walking the matrix front to back."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from sstar_b200.engine import get_engine
name = sys.argv[1] if len(sys.argv) > 1 else "c2"
wk = bench.make_workload(name, 0); st = bench.workload_stats(wk)
dev = torch.device("cuda", 0); eng = get_engine(0)
pin = {k: torch.from_numpy(np.ascontiguousarray(wk[k])).pin_memory() for k in ("ref", "tgt")}
d = {k: v.to(dev) for k, v in pin.items()}
s = torch.cuda.current_stream()
ws_ref = None
for ctas in [int(x) for x in os.environ.get("CTAS", "0,16,32,64,96,128,192,256,384").split(",")]:
    os.environ["SSTAR_B200_PACK_CTAS"] = str(ctas)
    for label, src in (("PCIe", pin), ("HBM", d)):
        ts = []
        for it in range(30):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); e0.record(s)
            ws = eng.pack_device(src["ref"], src["tgt"], W=st["W"], max_win_variants=st["max_window_variants"])
            e1.record(s); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) * 1e3)
        nb = src["ref"].numel() + src["tgt"].numel()
        if ws_ref is None: ws_ref = ws[:4 << 20].clone()
        same = bool(torch.equal(ws[256:4 << 20], ws_ref[256:]))
        print(f"{name} pack CTAs {ctas:4d} {label:5s}: median {np.median(ts[5:]):7.1f} us ({nb/np.median(ts[5:])/1e3:6.1f} GB/s) same={same}")
