"""Measure the device TSV emitter: python tools/exp_emit.py [W] [T]"""
import os, sys, time, tempfile
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sstar_b200.engine import get_engine
from sstar_b200.feature_table import FeatureTable

W = int(sys.argv[1]) if len(sys.argv) > 1 else 3041
T = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
rng = np.random.default_rng(1)
score = rng.integers(-10_000, 600_000, size=(W, T)).astype(np.int64)
score[rng.random((W, T)) < 0.3] = np.iinfo(np.int64).min
nsnp = rng.integers(0, 400, size=(W, T)).astype(np.int32)
ws = (1 + 10_000 * np.arange(W)).astype(np.int32)
we = ws + 49_999
labels = [f"tsk_{500 + t // 2}_{t % 2 + 1}" for t in range(T)]
eng = get_engine(0)
dev = torch.device("cuda", 0)
d = [torch.from_numpy(x).to(dev) for x in (score, nsnp, ws, we)]
for _ in range(3):
    out, total, status = eng.format_tsv_device(*d, "1", labels)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
ts = []
for _ in range(10):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out, total, status = eng.format_tsv_device(*d, "1", labels)   # includes the 16-byte readback sync
    ts.append(time.perf_counter() - t0)
dt = sorted(ts)[len(ts) // 2]
rows = W * T
print(f"device emitter: {rows} rows, {total/1e6:.1f} MB of text, {dt*1e3:.3f} ms per table "
      f"({rows/dt/1e9:.2f} G rows/s, {(total + 12*rows)/dt/1e9:.1f} GB/s read+written), status {status}")
t0 = time.perf_counter(); host = out[:total].cpu().numpy().tobytes(); t1 = time.perf_counter()
print(f"D2H of the text: {(t1-t0)*1e3:.1f} ms")
n = max(1, min(W, 200_000 // T))
tab = FeatureTable("1", ws[:n], we[:n], labels, score[:n], nsnp[:n])
with tempfile.TemporaryDirectory() as td:
    t0 = time.perf_counter(); tab.to_tsv(os.path.join(td, "py.tsv")); t1 = time.perf_counter()
    ref = open(os.path.join(td, "py.tsv"), "rb").read()
hdr = len(ref.split(b"\n", 1)[0]) + 1
assert host[: len(ref) - hdr] == ref[hdr:], "device text differs from the Python emitter"
print(f"python emitter: {n*T} rows in {(t1-t0)*1e3:.1f} ms ({n*T/(t1-t0)/1e6:.2f} M rows/s); first {n*T} rows identical")
