"""Phase timeline of the one-launch kernel on C2 (debug build with -DSSTAR_B200_FUSED_TIMING):
%globaltimer stamps of thread 0 of every CTA at the phase edges.  python tools/exp_fused_timing.py [build]"""
import ctypes, os, subprocess, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
LIB = os.path.join(ROOT, "sstar_b200", "csrc", "libsstar_b200_timing.so")
if len(sys.argv) > 1 and sys.argv[1] == "build":
    csrc = os.path.join(ROOT, "sstar_b200", "csrc")
    subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler",
                           "-fPIC", "-shared", "-DSSTAR_B200_FUSED_TIMING", "-I", os.path.join(ROOT, "include"), "-o", LIB,
                           os.path.join(csrc, "sstar_b200.cu"), os.path.join(csrc, "vcf_scan.cpp"), "-lz", "-lpthread"])
    sys.exit(0)
import torch
from sstar_b200 import _cabi
_cabi.LIB_PATH = LIB
from sstar_b200.engine import get_engine, max_window_variants
from sstar_b200.synth import regular_windows, synth_chromosome
ref, tgt, pos = synth_chromosome(10_000_000, 100, 50, False, 12346)
ws, we = regular_windows(pos, 50_000, 10_000)
eng = get_engine(0); dev = eng.tdev
d = [torch.from_numpy(np.ascontiguousarray(x)).to(dev) for x in (ref, tgt, pos, ws, we)]
hint = max_window_variants(pos, ws, we)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for i in range(5):
    flush.fill_(i & 1)
    eng.score_device(*d, max_win_variants=hint)
torch.cuda.synchronize()
out = np.zeros((4096, 8), dtype=np.uint64)
assert eng.lib.sstar_b200_debug_fused_stamps(out.ctypes.data_as(ctypes.c_void_p)) == 0
n = int((out[:, 0] != 0).sum())
st = out[:n].astype(np.int64)
GHZ = 1.965
names = ["entry", "list scanned", "matrices landed", "packed", "-", "DP done (thread 0)", "exit", "bounds done"]
life = (st[:, 6] - st[:, 0]) / GHZ / 1e3
print(f"{n} CTAs; CTA life (SM cycles / {GHZ} GHz): min {life.min():.2f} median {np.median(life):.2f} max {life.max():.2f} us")
for a, b in ((0, 1), (1, 7), (7, 2), (2, 3), (3, 5), (5, 6)):
    dlt = (st[:, b] - st[:, a]) / GHZ / 1e3
    print(f"  {names[a]:>18s} -> {names[b]:<18s}: median {np.median(dlt):5.2f} max {dlt.max():5.2f} us")
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
for label, fn in (("direct launch", lambda: eng.score_device(*d, max_win_variants=hint)),):
    ts = []
    for i in range(30):
        flush.fill_(i & 1)
        ev[0].record(); fn(); ev[1].record()
        torch.cuda.synchronize()
        ts.append(ev[0].elapsed_time(ev[1]) * 1e3)
    print(f"{label}: median {np.median(ts):.2f} us, min {min(ts):.2f} us between events (L2 flushed)")
one = torch.zeros(32, device=dev)
for label, fn, fl in (("tiny torch kernel (x.add_), flushed", lambda: one.add_(1), True), ("tiny torch kernel, not flushed", lambda: one.add_(1), False),
                      ("fused kernel, NOT flushed", lambda: eng.score_device(*d, max_win_variants=hint), False),
                      ("fused kernel, flushed, 2 back to back (2nd - 1st)", None, True)):
    ts = []
    for i in range(30):
        if fl:
            flush.fill_(i & 1)
        if fn is None:
            e3 = torch.cuda.Event(enable_timing=True)
            ev[0].record(); eng.score_device(*d, max_win_variants=hint); ev[1].record(); eng.score_device(*d, max_win_variants=hint); e3.record()
            torch.cuda.synchronize()
            ts.append(ev[1].elapsed_time(e3) * 1e3)
        else:
            ev[0].record(); fn(); ev[1].record()
            torch.cuda.synchronize()
            ts.append(ev[0].elapsed_time(ev[1]) * 1e3)
    print(f"{label}: median {np.median(ts):.2f} us, min {min(ts):.2f} us")
# does the flush's own dirty data sit in front of the timed kernel?  write-only flush vs write + read of another buffer
other = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
other.fill_(3)
for label, pre in (("flush = 256 MiB write", lambda i: flush.fill_(i & 1)),
                   ("flush = 256 MiB write, then 256 MiB read of another buffer", lambda i: (flush.fill_(i & 1), other.view(torch.int64).sum())),
                   ("flush = 256 MiB write, then 20 us of idle-ish work (tiny kernels)", lambda i: (flush.fill_(i & 1), [one.add_(1) for _ in range(8)]))):
    for name, fn in (("fused kernel", lambda: eng.score_device(*d, max_win_variants=hint)), ("tiny kernel", lambda: one.add_(1))):
        ts = []
        for i in range(40):
            pre(i)
            ev[0].record(); fn(); ev[1].record()
            torch.cuda.synchronize()
            ts.append(ev[0].elapsed_time(ev[1]) * 1e3)
        print(f"{label} | {name}: median {np.median(ts):.2f} us, min {min(ts):.2f} us")
