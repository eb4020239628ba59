"""C5 (BASELINE.json configs[4]): 10,000 synthetic 50 kb replicates scored in one replicate-batched
call (the simulate / train path).  python tools/exp_c5.py [n_rep] [nref] [ntgt]"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import c_oracle
from sstar_b200.engine import get_engine
from sstar_b200.synth import synth_chromosome

n_rep = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000
nref = int(sys.argv[2]) if len(sys.argv) > 2 else 5
ntgt = int(sys.argv[3]) if len(sys.argv) > 3 else 10
L = 50_000
# one long synthetic chromosome cut into n_rep independent 50 kb "replicates" (positions folded to [1, L])
ref, tgt, pos = synth_chromosome(L * n_rep, nref, ntgt, False, seed=5)
rep_id = (pos.astype(np.int64) - 1) // L
off = np.searchsorted(rep_id, np.arange(n_rep + 1)).astype(np.int32)
pos_local = ((pos.astype(np.int64) - 1) % L + 1).astype(np.int32)
eng = get_engine(0); dev = torch.device("cuda", 0)
d = [torch.from_numpy(np.ascontiguousarray(x)).to(dev) for x in (ref, tgt, pos_local, off)]
hint = int(np.diff(off).max())
for _ in range(3):
    s, n = eng.score_replicates_device(d[0], d[1], d[2], d[3], 1, L, max_win_variants=hint)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
torch.cuda.synchronize(); ev[0].record()
for _ in range(20):
    s, n = eng.score_replicates_device(d[0], d[1], d[2], d[3], 1, L, max_win_variants=hint)
ev[1].record(); torch.cuda.synchronize()
ms = ev[0].elapsed_time(ev[1]) / 20
sel = np.unique(np.linspace(0, n_rep - 1, 200).astype(int))
ok = True
sh, nh = s.cpu().numpy(), n.cpu().numpy()
for r in sel:
    a, b = off[r], off[r + 1]
    so, no = c_oracle.score_windows(ref[a:b], tgt[a:b], pos_local[a:b], [1], [L])
    ok &= np.array_equal(sh[r], so[0]) and np.array_equal(nh[r], no[0])
t0 = time.perf_counter()
for r in sel[:50]:
    a, b = off[r], off[r + 1]
    c_oracle.score_windows(ref[a:b], tgt[a:b], pos_local[a:b], [1], [L])
cpu_us = (time.perf_counter() - t0) / (50 * ntgt) * 1e6
print(f"C5: {n_rep} replicates x {ntgt} columns = {n_rep*ntgt} units, V = {len(pos)} ({np.diff(off).mean():.0f} variants "
      f"per replicate, nan fraction {np.mean(sh == np.iinfo(np.int64).min):.2f}); {ms*1e3:.1f} us per batch "
      f"({n_rep*ntgt/(ms*1e-3)/1e6:.0f} M units/s device-resident); bit-exact vs oracle on {len(sel)} replicates: {ok}; "
      f"scalar C oracle {cpu_us:.1f} us/unit")
