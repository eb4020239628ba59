"""The kernels beside the scoring pass -- result emitter (N1), ingest filters (N2), match numerators (N4), SNP sets --
each at a size that fills the GPU, timed with CUDA events (median of 10 after warm-up) against the bytes the step must
move (algorithmic, stated per line) and the measured HBM peak.  One JSON line per step.
    python tools/exp_aux_kernels.py [out.json]"""
import json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from sstar_b200.engine import get_engine
eng = get_engine(0); dev = eng.tdev
peak, _ = bench.measured_peak()
rng = np.random.default_rng(1)
lines = []


ONCE = os.environ.get("AUX_ONCE") == "1"  # profiling aid: every step exactly once, no timing loops


def timed(fn, reps=10):
    if ONCE:
        return float("nan")
    for _ in range(3):
        fn()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


def report(step, kernels, ms, nbytes, what, **extra):
    rec = {"step": step, "kernels": kernels, "ms": ms, "algorithmic_bytes": int(nbytes), "bytes_are": what,
           "achieved_GBs": nbytes / (ms * 1e-3) / 1e9, "frac_of_hbm_peak": nbytes / (ms * 1e-3) / 1e9 / peak, **extra}
    lines.append(rec)
    print(json.dumps(rec), flush=True)


# ---- N1 result emitter: a C3-shard sized table (3,041 windows x 1,000 columns)
W, T = 3041, 1000
score = rng.integers(-10000, 400000, size=(W, T)).astype(np.int64)
score[rng.random((W, T)) < 0.6] = np.iinfo(np.int64).min
nsnp = rng.integers(0, 400, size=(W, T)).astype(np.int32)
ws = (1 + 10_000 * np.arange(W)).astype(np.int32); we = (ws + 49_999).astype(np.int32)
labels = [f"tsk_{i // 2}_{i % 2 + 1}" for i in range(T)]
d = [torch.from_numpy(x).to(dev) for x in (score, nsnp, ws, we)]
out, total, status = eng.format_tsv_device(*d, "1", labels)
ms = timed(lambda: eng.format_tsv_device(*d, "1", labels))
report("emit (sstar_b200_format_tsv)", ["emit_len_kernel", "emit_scan_kernel", "emit_format_kernel"], ms, 2 * 12 * W * T + total,
       "both tables read twice (length pass, format pass) + the text written; incl. the 16-byte readback sync",
       rows=W * T, text_bytes=int(total), rows_per_s=W * T / (ms * 1e-3))

# ---- N2 ingest filters: raw calls [V, N, 2] of 150 samples
V, N = 400_000, 150
gt = (rng.random((V, N, 2)) < 0.05).astype(np.int8)
gt[rng.random((V, N, 2)) < 0.005] = -1
pos = np.sort(rng.choice(np.arange(1, 40 * V), size=V, replace=False)).astype(np.int32)
gt_d, pos_d = torch.from_numpy(gt).to(dev), torch.from_numpy(pos).to(dev)
rc, tc = list(range(100)), list(range(100, 150))
r, t, p, shared = eng.ingest_device(gt_d, pos_d, rc, tc, False)
ms = timed(lambda: eng.ingest_device(gt_d, pos_d, rc, tc, False))
kept = int(p.shape[0])
report("ingest (sstar_b200_ingest_ex)", ["ingest_flag_kernel", "ingest_scan_kernel", "ingest_write_kernel"], ms,
       2 * V * N * 2 + kept * (N + 4) + V * 5,
       "raw calls read twice (flag pass, write pass) + keep flags + the compacted matrices and positions written; incl. the "
       "16-byte readback sync", variants=V, kept=kept)

# ---- N4 match numerators: tracts of ~300 variants against 10 sources
K, S = 20_000, 10
lo = rng.integers(0, V - 400, size=K).astype(np.int32); hi = (lo + rng.integers(100, 400, size=K)).astype(np.int32)
tcol = rng.integers(0, 100, size=K).astype(np.int32); hap = rng.integers(-1, 2, size=K).astype(np.int32)
src = np.arange(100, 100 + S, dtype=np.int32)
dd = [torch.from_numpy(x).to(dev) for x in (lo, hi, tcol, hap, src)]
outn = torch.zeros((K, S), dtype=torch.int64, device=dev)
import ctypes
def match():
    rcode = eng.lib.sstar_b200_match_numerators(0, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream), gt_d.data_ptr(), V, N,
                                                dd[0].data_ptr(), dd[1].data_ptr(), dd[2].data_ptr(), dd[3].data_ptr(), K,
                                                dd[4].data_ptr(), S, 2, outn.data_ptr())
    assert rcode == 0
ms = timed(match)
rows = int((hi - lo).sum())
report("match (sstar_b200_match_numerators)", ["match_kernel"], ms, rows * (32 + 32) + K * S * 8,
       "per tract row: the 32-byte sector of the target column + the sector(s) of the 10 adjacent source columns; numerators written",
       tracts=K, sources=S, tract_rows=rows, pairs_per_s=K * S / (ms * 1e-3))
# parity of a sample against numpy
if ONCE:
    match()
g = gt.astype(np.int32)
for k in rng.choice(K, 50, replace=False):
    sl = g[lo[k]:hi[k]]
    tg = sl[:, tcol[k]]
    x = np.where((tg < 0).any(1), -1, (tg > 0).sum(1)) if hap[k] < 0 else np.where(tg[:, hap[k]] >= 0, tg[:, hap[k]] * 2, -1)
    for j, sc in enumerate(src):
        sg = sl[:, sc]
        y = np.where((sg < 0).any(1), -1, (sg > 0).sum(1))
        ok = (x >= 0) & (y >= 0)
        assert int(outn[k, j]) == int((2 - np.abs(x[ok] - y[ok])).sum()), (k, j)

# ---- SNP sets on the C3 shard
wk = bench.make_workload("c3-shard", 1); st = bench.workload_stats(wk)
dw = {k: torch.from_numpy(np.ascontiguousarray(wk[k])).to(dev) for k in ("ref", "tgt", "pos", "ws", "we")}
hint = st["max_window_variants"]
f_sets = lambda: eng.snp_sets_device(dw["ref"], dw["tgt"], dw["pos"], dw["ws"], dw["we"], max_win_variants=hint)
f_score = lambda: eng.score_device(dw["ref"], dw["tgt"], dw["pos"], dw["ws"], dw["we"], max_win_variants=hint)
o = f_sets()
ms_sets, ms_score = timed(f_sets, 5), timed(f_score, 5)
ab = bench.algorithmic_bytes(st)["pipeline"] + 8 * st["units"] + 4 * int(o[2][-1])
report("SNP sets (sstar_b200_snp_set_offsets + _fill), C3 shard", ["prep_kernel", "snp_count_kernel", "snp_set_kernel<false>", "scan_*",
       "snp_set_kernel<true>"], ms_sets, ab, "the scoring pass's bytes + the offsets table + the chain positions; incl. the host read of "
       "the total between the two calls", units=st["units"], chain_sites=int(o[2][-1]), scores_alone_ms=ms_score,
       ratio_to_scores_alone=ms_sets / ms_score)
if len(sys.argv) > 1:
    json.dump(lines, open(sys.argv[1], "w"), indent=1)
