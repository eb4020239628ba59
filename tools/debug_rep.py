import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from sstar_b200.engine import get_engine
from sstar_b200.simulate_batch import build_feature_input
from oracle import c_oracle

def align(x, a=256): return (x + a - 1) // a * a
phased = True
rng = np.random.default_rng(5)
nref, ntgt, ploidy, L = 5, 10, 2, 50_000
reps = []
for r in range(40):
    nsite = int(rng.integers(0 if r else 3, 260)) if r % 9 else 3
    pos = np.sort(rng.choice(L + 1, size=max(nsite, 1), replace=False))
    freq = rng.random(len(pos))[:, None] ** 3
    gts = (rng.random((len(pos), (nref + ntgt) * ploidy)) < freq).astype(np.int8)
    gts[:, : nref * ploidy] &= (rng.random((len(pos), 1)) < 0.5)
    reps.append((pos, gts))
refs, tgts, poss, off = [], [], [], [0]
for pos, gts in reps:
    r, t, p, s, e = build_feature_input(pos, gts, nref, ntgt, ploidy, L, phased)
    refs.append(r); tgts.append(t); poss.append(p.astype(np.int32)); off.append(off[-1] + len(p))
ref = np.concatenate(refs); tgt = np.concatenate(tgts); pos = np.concatenate(poss); off = np.asarray(off, np.int32)
eng = get_engine(0)
score, nsnp = eng.score_replicates_host(ref, tgt, pos, off, 1, L)
torch.cuda.synchronize()
V, T = tgt.shape; G = (V + 31) // 32; Tp = (T + 31) // 32 * 32; N = len(reps)
raw = eng._dev["ws"].cpu().numpy()
o = 256
o_abs = o; o = align(o + G * 4); o_exo = o; o = align(o + G * 4); o_car = o; o = align(o + G * Tp * 4); o_two = o; o = align(o + G * Tp * 4)
o_lo = o; o = align(o + N * 4); o_hi = o
lo = raw[o_lo:o_lo + 4 * N].view(np.int32); hi = raw[o_hi:o_hi + 4 * N].view(np.int32)
absent = ref.astype(np.int32).sum(1) == 0
for i in range(N):
    so, no = c_oracle.score_windows(refs[i], tgts[i], poss[i], [1], [L])
    if not (np.array_equal(nsnp[i], no[0]) and np.array_equal(score[i], so[0])):
        print("rep", i, "off", off[i], off[i + 1], "gpu lo/hi", lo[i], hi[i], "nsnp gpu", nsnp[i][:4], "exp", no[0][:4],
              "Vr exp", int((~absent[off[i]:off[i + 1]]).sum()), "len", off[i + 1] - off[i], "first pos", poss[i][:3])
print("done; mismatching lo/hi:", [(i, lo[i], hi[i], off[i], off[i + 1]) for i in range(N) if lo[i] != off[i] or hi[i] != off[i + 1]])
