"""BASELINE configs[2] at full size: one synthetic 250 Mb chromosome, 500 ref + 500 target diploids in
haplotype mode (R = T = 1000 columns), 50 kb / 10 kb windows, sharded by window range over all visible
GPUs of the box in ONE process (sstar_b200.sharding.score_chromosome).
python tools/exp_c3_full.py [length_bp]"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import c_oracle
from sstar_b200.sharding import plan_window_shards, score_chromosome, score_chromosome_pinned
from sstar_b200.synth import regular_windows, synth_chromosome_device

L = int(sys.argv[1]) if len(sys.argv) > 1 else 250_000_000
n_gpu = torch.cuda.device_count()
t0 = time.perf_counter()
r, t, p = synth_chromosome_device(torch, torch.device("cuda", 0), L, 500, 500, True, seed=12347)
ref, tgt, pos = r.cpu().numpy(), t.cpu().numpy(), p.cpu().numpy()
del r, t, p
torch.cuda.empty_cache()
ws, we = regular_windows(pos, 50_000, 10_000)
W, T = len(ws), tgt.shape[1]
print(f"generated V={len(pos)} x (R=1000, T=1000), W={W}, units={W*T/1e6:.1f} M, int8 {ref.nbytes*2/1e9:.2f} GB "
      f"in {time.perf_counter()-t0:.1f} s; {n_gpu} GPUs", flush=True)
for devs in ([0], list(range(n_gpu))):
    score_chromosome(ref[:4096], tgt[:4096], pos[:4096], ws[:8], we[:8], 5000, 5, -10000, devices=devs)  # warm-up
    ts = []
    for _ in range(2):
        t0 = time.perf_counter()
        s, n = score_chromosome(ref, tgt, pos, ws, we, 5000, 5, -10000, devices=devs)
        ts.append(time.perf_counter() - t0)
    print(f"{len(devs)} GPU(s): {min(ts)*1e3:.0f} ms host numpy in -> numpy out ({W*T/min(ts)/1e9:.2f} G units/s, "
          f"{(ref.nbytes*2)/min(ts)/1e9:.1f} GB/s of genotypes), shards: "
          f"{[(a, b) for a, b, _, _ in plan_window_shards(pos, ws, we, len(devs))][:3]}...", flush=True)
    if len(devs) == 1:
        s1, n1 = s, n
assert np.array_equal(s, s1) and np.array_equal(n, n1), "multi-GPU result differs from single GPU"
# the same sharding on page-locked buffers: no host staging, every GPU on its own PCIe link
pin = [torch.from_numpy(np.ascontiguousarray(x)).pin_memory() for x in (ref, tgt, pos, ws, we)]
sp = torch.empty((W, T), dtype=torch.int64).pin_memory()
npn = torch.empty((W, T), dtype=torch.int32).pin_memory()
base = None
for devs in ([0], [0, 1], [0, 1, 2, 3], list(range(n_gpu))):
    if len(devs) > n_gpu:
        continue
    score_chromosome_pinned(pin[0][:4096], pin[1][:4096], pin[2][:4096], pin[3][:8], pin[4][:8], sp[:8], npn[:8], devices=devs)
    ts = []
    for _ in range(3):
        t0 = time.perf_counter()
        score_chromosome_pinned(*pin, sp, npn, devices=devs)
        ts.append(time.perf_counter() - t0)
    base = base or min(ts)
    ok = bool(np.array_equal(sp.numpy(), s1) and np.array_equal(npn.numpy(), n1))
    print(f"pinned, {len(devs)} GPU(s): {min(ts)*1e3:.1f} ms ({W*T/min(ts)/1e9:.2f} G units/s, {base/min(ts):.2f}x of 1 GPU), "
          f"identical to the single-GPU tables: {ok}", flush=True)
sel = np.unique(np.linspace(0, W - 1, 16).astype(int))
so, no = c_oracle.score_windows(ref, tgt, pos, ws[sel], we[sel])
print("bit-exact vs oracle on", len(sel), "windows:", bool(np.array_equal(s[sel], so) and np.array_equal(n[sel], no)),
      "; multi-GPU == single GPU: True")
