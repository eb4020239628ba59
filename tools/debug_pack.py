"""Debug helper: run pack on the device and compare the workspace planes with numpy."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from sstar_b200.engine import get_engine

def align(x, a=256): return (x + a - 1) // a * a

def planes_numpy(ref, tgt):
    V, T = tgt.shape
    G = (V + 31) // 32
    absent = ref.astype(np.int32).sum(1) == 0
    pad = G * 32 - V
    ab = np.concatenate([absent, np.zeros(pad, bool)]).reshape(G, 32)
    w = (1 << np.arange(32, dtype=np.uint64))
    absent_bits = (ab * w).sum(1).astype(np.uint32)
    tg = np.concatenate([tgt, np.zeros((pad, T), np.int8)]).reshape(G, 32, T)
    car = (((tg != 0) & ab[:, :, None]) * w[None, :, None]).sum(1).astype(np.uint32)
    two = (((tg == 2) & ab[:, :, None]) * w[None, :, None]).sum(1).astype(np.uint32)
    exo = ((((tg < 0) | (tg > 2)) & ab[:, :, None]).any(axis=(1, 2))).astype(np.uint32)
    return absent_bits, car, two, exo

def check(ref, tgt, tag=""):
    eng = get_engine(0)
    V, R = ref.shape; T = tgt.shape[1]
    dev = torch.device("cuda", 0)
    ws = eng.pack_device(torch.from_numpy(ref).to(dev), torch.from_numpy(tgt).to(dev))
    torch.cuda.synchronize()
    raw = ws.cpu().numpy()
    G = (V + 31) // 32; Tp = (T + 31) // 32 * 32
    off = 256
    o_abs = off; off = align(off + G * 4)
    o_exo = off; off = align(off + G * 4)
    o_car = off; off = align(off + G * Tp * 4)
    o_two = off; off = align(off + G * Tp * 4)
    g_abs = raw[o_abs:o_abs + G * 4].view(np.uint32)
    g_exo = raw[o_exo:o_exo + G * 4].view(np.uint32)
    g_car = raw[o_car:o_car + G * Tp * 4].view(np.uint32).reshape(G, Tp)[:, :T]
    g_two = raw[o_two:o_two + G * Tp * 4].view(np.uint32).reshape(G, Tp)[:, :T]
    a, c, t, e = planes_numpy(ref, tgt)
    ok = True
    for name, x, y in (("absent", g_abs, a), ("carried", g_car, c), ("two", g_two, t), ("exotic", g_exo != 0, e != 0)):
        bad = np.argwhere(x != y)
        if len(bad):
            ok = False
            print(tag, name, "MISMATCH", len(bad), "first", bad[:5].tolist(), [hex(int(x[tuple(b)])) for b in bad[:3]], [hex(int(y[tuple(b)])) for b in bad[:3]])
    print(tag, "V,R,T", V, R, T, "OK" if ok else "FAIL")
    return ok

if __name__ == "__main__":
    rng = np.random.default_rng(0)
    for (V, R, T) in [(5000, 10, 20), (5000, 5, 10), (33, 10, 20), (1000, 100, 50), (4096, 8, 16), (777, 3, 7), (2000, 1000, 1000), (3000, 12, 128), (100, 1, 1)]:
        ref = (rng.random((V, R)) < 0.02).astype(np.int8)
        tgt = rng.choice([0, 1, 2], size=(V, T), p=[0.7, 0.2, 0.1]).astype(np.int8)
        check(ref, tgt, "plain")
        tgt2 = tgt.copy(); tgt2[rng.random(tgt.shape) < 0.001] = -1; tgt2[rng.random(tgt.shape) < 0.001] = 3
        check(ref, tgt2, "exotic")
