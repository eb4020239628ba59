"""CPU baseline leg for bench.py -- TEST INFRASTRUCTURE (see oracle/__init__.py).

Times the numpy port of the reference recipe (``oracle.sstar_port.compute_float``: float64 pair
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


def _init(ref, tgt, pos, ws, we, params):
    _G.update(ref=ref, tgt=tgt, pos=pos, ws=ws, we=we, params=params)


def _score_window(w):
    pos = _G["pos"]
    inside = (pos >= _G["ws"][w]) & (pos <= _G["we"][w])  # window_data_generator.py:113
    b, mx, p = _G["params"]
    out = compute_float(ref_gts=_G["ref"][inside], tgt_gts=_G["tgt"][inside], pos=pos[inside],
                        match_bonus=b, max_mismatch=mx, mismatch_penalty=p)
    return w, float_scores_to_int(out["sstar"]), np.asarray(out["region_ind_SNP_number"], np.int32)


class CpuBaseline:
    """Pool of forked workers holding one chromosome; ``run(windows)`` scores those windows."""

    def __init__(self, ref, tgt, pos, ws, we, params=(5000, 5, -10000), nproc=None):
        self.nproc = int(nproc or os.cpu_count() or 1)
        self.T = tgt.shape[1]
        _init(ref, tgt, pos, ws, we, params)  # in the parent too (nproc == 1 path, calibration)
        self.pool = None
        if self.nproc > 1:
            ctx = mp.get_context("fork")
            self.pool = ctx.Pool(self.nproc, initializer=_init, initargs=(ref, tgt, pos, ws, we, params))
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
