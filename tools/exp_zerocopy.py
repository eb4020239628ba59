"""Experiment: does the pack kernel stream pinned HOST memory (zero-copy over PCIe) as fast as
cudaMemcpyAsync moves it?  python tools/exp_zerocopy.py [workload]"""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from sstar_b200.engine import get_engine

name = sys.argv[1] if len(sys.argv) > 1 else "c2"
wk = bench.make_workload(name, 1)
dev = torch.device("cuda", 0)
eng = get_engine(0)
ref_p = torch.from_numpy(wk["ref"]).pin_memory()
tgt_p = torch.from_numpy(wk["tgt"]).pin_memory()
ref_d, tgt_d = ref_p.to(dev), tgt_p.to(dev)
W = len(wk["ws"])
nbytes = ref_p.numel() + tgt_p.numel()

def timed(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    for a, b in evs:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    ms = sorted(a.elapsed_time(b) for a, b in evs)
    return ms[len(ms) // 2]

def copy():
    ref_d.copy_(ref_p, non_blocking=True); tgt_d.copy_(tgt_p, non_blocking=True)

t_copy = timed(copy)
t_pack_dev = timed(lambda: eng.pack_device(ref_d, tgt_d, W=W))
t_pack_host = timed(lambda: eng.pack_device(ref_p, tgt_p, W=W))
# correctness of the zero-copy pack: same planes
ws1 = eng.pack_device(ref_d, tgt_d, W=W).clone()
ws2 = eng.pack_device(ref_p, tgt_p, W=W)
torch.cuda.synchronize()
same = bool(torch.equal(ws1, ws2))
print(f"{name}: {nbytes/1e6:.1f} MB  memcpy {t_copy*1e3:.1f} us ({nbytes/t_copy/1e6:.1f} GB/s)  "
      f"pack(device) {t_pack_dev*1e3:.1f} us  pack(pinned host, zero-copy) {t_pack_host*1e3:.1f} us "
      f"({nbytes/t_pack_host/1e6:.1f} GB/s)  identical planes: {same}")
