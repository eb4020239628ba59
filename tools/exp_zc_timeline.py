"""Where the in-place (zero-copy) host entry spends its time on C2: enqueue, pack over PCIe, DP, sync."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from sstar_b200.engine import get_engine
name = sys.argv[1] if len(sys.argv) > 1 else "c2"
wk = bench.make_workload(name, 1); st = bench.workload_stats(wk)
dev = torch.device("cuda", 0); eng = get_engine(0)
pin = {k: torch.from_numpy(np.ascontiguousarray(wk[k])).pin_memory() for k in ("ref", "tgt", "pos", "ws", "we")}
W, T, R = st["W"], st["T"], st["R"]
score_p = torch.empty((W, T), dtype=torch.int64).pin_memory(); nsnp_p = torch.empty((W, T), dtype=torch.int32).pin_memory()
hint = st["max_window_variants"]
s = torch.cuda.current_stream()
rows = []
for it in range(40):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); time.sleep(0.0005)
    t0 = time.perf_counter(); e0.record(s)
    eng.score_pinned(pin["ref"], pin["tgt"], pin["pos"], pin["ws"], pin["we"], score_p, nsnp_p, max_win_variants=hint, sync=False)
    t1 = time.perf_counter(); e1.record(s)
    s.synchronize(); t2 = time.perf_counter()
    rows.append(((t1 - t0) * 1e6, (t2 - t0) * 1e6, e0.elapsed_time(e1) * 1e3))
rows = np.array(rows[8:])
print(f"{name} host entry: enqueue returns after {np.median(rows[:,0]):.1f} us, synced after {np.median(rows[:,1]):.1f} us, "
      f"GPU span (event to event) {np.median(rows[:,2]):.1f} us")
# the pack kernel alone, reading the pinned matrices in place vs from HBM
d = {k: v.to(dev) for k, v in pin.items()}
for label, src in (("in place over PCIe", pin), ("from HBM", d)):
    ts = []
    for it in range(30):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record(s)
        eng.pack_device(src["ref"], src["tgt"], W=W, max_win_variants=hint)
        e1.record(s); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) * 1e3)
    nb = src["ref"].numel() + src["tgt"].numel()
    print(f"  pack alone {label}: {np.median(ts[5:]):.1f} us ({nb/np.median(ts[5:])/1e3:.1f} GB/s)")
# plain copies of the same bytes
ts = []
for it in range(30):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record(s)
    d["ref"].copy_(pin["ref"], non_blocking=True); d["tgt"].copy_(pin["tgt"], non_blocking=True)
    e1.record(s); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) * 1e3)
print(f"  two cudaMemcpyAsync of the matrices: {np.median(ts[5:]):.1f} us ({nb/np.median(ts[5:])/1e3:.1f} GB/s)")
