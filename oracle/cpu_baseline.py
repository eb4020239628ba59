"""CPU baseline leg for bench.py -- TEST INFRASTRUCTURE (see oracle/__init__.py).

When the UNMODIFIED reference is installed next to the repo (``baseline/_ref``, made by
``pip install --no-deps --target baseline/_ref <reference>``; it is git-ignored and travels to the
GPU box) the workers run the reference's own job -- ``FeatureVectorPreprocessor.run`` ->
``Sstar.compute`` (sstar/feature_vector_preprocessor.py:71, sstar/sstar.py:43) -- on the window
slices ``WindowDataGenerator`` would hand it (``kind = "reference"``).  Otherwise (``kind =
"port"``) they time the numpy port of the reference recipe (``oracle.sstar_port.compute_float``: float64 pair
matrices, ``np.union1d``, Python double loop -- the reference's cost model, sstar/sstar.py:156-217)
over a pool of worker processes, one task per window as the reference's ``mp_manager`` does
(sstar/mp_manager.py:162-163), but WITHOUT its Manager-proxy pickling and 5 s exit tail
(sstar/mp_manager.py:238-242): workers are forked after the arrays are in place and the pool is
warm before the clock starts.  That makes this a generous stand-in for the reference's CPU path.
"""
from __future__ import annotations

import multiprocessing as mp
import os
import time

import numpy as np

from .sstar_port import compute_float, float_scores_to_int

_G = {}
_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load_reference_job():
    """The reference's ``FeatureVectorPreprocessor`` class from ``baseline/_ref`` (or the read-only
    source tree when it is mounted), or None.  Only numpy is needed by these two modules."""
    import importlib
    import sys
    for root in (os.path.join(_ROOT, "baseline", "_ref"), "/root/reference"):
        if not os.path.isfile(os.path.join(root, "sstar", "sstar.py")):
            continue
        sys.path.insert(0, root)
        try:
            mod = importlib.import_module("sstar.feature_vector_preprocessor")
            return mod.FeatureVectorPreprocessor, root
        except Exception:
            for k in [k for k in sys.modules if k == "sstar" or k.startswith("sstar.")]:
                del sys.modules[k]
        finally:
            sys.path.remove(root)
    return None, None


def _init(ref, tgt, pos, ws, we, params, job=None, phased=False):
    _G.update(ref=ref, tgt=tgt, pos=pos, ws=ws, we=we, params=params, job=job, phased=phased)


def _score_window(w):
    pos = _G["pos"]
    inside = (pos >= _G["ws"][w]) & (pos <= _G["we"][w])  # window_data_generator.py:113
    b, mx, p = _G["params"]
    if _G.get("job") is not None:  # the reference's own job object, called as its mp_worker calls it
        rows = _G["job"].run(chr_name="1", start=int(_G["ws"][w]), end=int(_G["we"][w]), ploidy=2,
                             is_phased=_G["phased"], ref_gts=_G["ref"][inside], tgt_gts=_G["tgt"][inside],
                             pos=pos[inside])
        return (w, float_scores_to_int([r["S*_score"] for r in rows]),
                np.asarray([r["Region_ind_SNP_number"] for r in rows], np.int32))
    out = compute_float(ref_gts=_G["ref"][inside], tgt_gts=_G["tgt"][inside], pos=pos[inside],
                        match_bonus=b, max_mismatch=mx, mismatch_penalty=p)
    return w, float_scores_to_int(out["sstar"]), np.asarray(out["region_ind_SNP_number"], np.int32)


class CpuBaseline:
    """Pool of forked workers holding one chromosome; ``run(windows)`` scores those windows."""

    def __init__(self, ref, tgt, pos, ws, we, params=(5000, 5, -10000), nproc=None, phased=False,
                 use_reference=True):
        self.nproc = int(nproc or os.cpu_count() or 1)
        self.T = tgt.shape[1]
        job, self.kind, self.source = None, "port", "oracle/sstar_port.py (numpy port of the reference recipe)"
        if use_reference:
            cls, root = load_reference_job()
            if cls is not None:
                ntgt = self.T // 2 if phased else self.T
                job = cls(ref_samples=[f"r{i}" for i in range(max(1, ref.shape[1] // (2 if phased else 1)))],
                          tgt_samples=[f"t{i}" for i in range(ntgt)], match_bonus=params[0],
                          max_mismatch=params[1], mismatch_penalty=params[2])
                self.kind = "reference"
                self.source = f"{os.path.relpath(root, _ROOT) if root.startswith(_ROOT) else root}: " \
                              "sstar.feature_vector_preprocessor.FeatureVectorPreprocessor.run -> sstar.sstar.Sstar.compute"
        _init(ref, tgt, pos, ws, we, params, job, phased)  # in the parent too (nproc == 1 path, calibration)
        self.pool = None
        if self.nproc > 1:
            ctx = mp.get_context("fork")
            self.pool = ctx.Pool(self.nproc, initializer=_init, initargs=(ref, tgt, pos, ws, we, params, job, phased))
            self.pool.map(_score_window, list(range(min(len(ws), self.nproc))))  # warm the workers

    def per_window_seconds(self, k=6):
        t0 = time.perf_counter()
        W = len(_G["ws"])
        idx = np.linspace(0, W - 1, num=min(k, W)).astype(int)
        for w in idx:
            _score_window(int(w))
        return (time.perf_counter() - t0) / max(len(idx), 1)

    def run(self, windows):
        """-> (seconds, {w: (score[T], nsnp[T])}) for the given window indices."""
        windows = [int(w) for w in windows]
        t0 = time.perf_counter()
        if self.pool is None:
            res = [_score_window(w) for w in windows]
        else:
            chunk = max(1, len(windows) // (self.nproc * 8))
            res = self.pool.map(_score_window, windows, chunksize=chunk)
        dt = time.perf_counter() - t0
        return dt, {w: (s, n) for w, s, n in res}

    def close(self):
        if self.pool is not None:
            self.pool.close()
            self.pool.join()
            self.pool = None
