"""ctypes loader for the C oracle (oracle/sstar_oracle.c) -- TEST INFRASTRUCTURE."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build() -> str:
    path = os.path.join(_HERE, "libsstar_oracle.so")
    src = os.path.join(_HERE, "sstar_oracle.c")
    if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return path


def _lib():
    global _LIB
    if _LIB is None:
        lib = ctypes.CDLL(build())
        lib.sstar_oracle_score_windows.restype = ctypes.c_int
        lib.sstar_oracle_score_windows.argtypes = [
            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32,
            ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32,
            ctypes.c_int32, ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p]
        lib.sstar_oracle_snp_sets.restype = ctypes.c_int
        lib.sstar_oracle_snp_sets.argtypes = lib.sstar_oracle_score_windows.argtypes + [
            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64]
        _LIB = lib
    return _LIB


def score_windows(ref_gts, tgt_gts, pos, win_start, win_end, match_bonus=5000, max_mismatch=5,
                  mismatch_penalty=-10000):
    """(score[W,T] int64 with INT64_MIN = NaN, nsnp[W,T] int32) via the scalar C oracle."""
    ref = np.ascontiguousarray(ref_gts, dtype=np.int8)
    tgt = np.ascontiguousarray(tgt_gts, dtype=np.int8)
    p = np.ascontiguousarray(pos, dtype=np.int32)
    ws = np.ascontiguousarray(win_start, dtype=np.int32)
    we = np.ascontiguousarray(win_end, dtype=np.int32)
    V, R = ref.shape if ref.ndim == 2 else (0, 0)
    T = tgt.shape[1]
    W = len(ws)
    score = np.empty((W, T), dtype=np.int64)
    nsnp = np.empty((W, T), dtype=np.int32)
    rc = _lib().sstar_oracle_score_windows(
        ref.ctypes.data, tgt.ctypes.data, p.ctypes.data, V, R, T, ws.ctypes.data, we.ctypes.data,
        W, int(match_bonus), int(max_mismatch), int(mismatch_penalty), score.ctypes.data,
        nsnp.ctypes.data)
    if rc != 0:
        raise MemoryError("C oracle allocation failed")
    return score, nsnp


def snp_sets(ref_gts, tgt_gts, pos, win_start, win_end, match_bonus=5000, max_mismatch=5,
             mismatch_penalty=-10000):
    """(score, nsnp, set_off int64 [W*T+1], set_pos int32 [total]): scores plus the positions on
    every unit's best chain, via the scalar C oracle."""
    ref = np.ascontiguousarray(ref_gts, dtype=np.int8)
    tgt = np.ascontiguousarray(tgt_gts, dtype=np.int8)
    p = np.ascontiguousarray(pos, dtype=np.int32)
    ws = np.ascontiguousarray(win_start, dtype=np.int32)
    we = np.ascontiguousarray(win_end, dtype=np.int32)
    V, R = ref.shape if ref.ndim == 2 else (0, 0)
    T = tgt.shape[1]
    W = len(ws)
    score = np.empty((W, T), dtype=np.int64)
    nsnp = np.empty((W, T), dtype=np.int32)
    off = np.zeros(W * T + 1, dtype=np.int64)
    args = [ref.ctypes.data, tgt.ctypes.data, p.ctypes.data, V, R, T, ws.ctypes.data, we.ctypes.data,
            W, int(match_bonus), int(max_mismatch), int(mismatch_penalty), score.ctypes.data,
            nsnp.ctypes.data, off.ctypes.data]
    if _lib().sstar_oracle_snp_sets(*args, None, 0) != 0:
        raise MemoryError("C oracle allocation failed")
    sets = np.empty(max(int(off[-1]), 1), dtype=np.int32)
    if _lib().sstar_oracle_snp_sets(*args, sets.ctypes.data, int(off[-1])) != 0:
        raise MemoryError("C oracle allocation failed")
    return score, nsnp, off, sets[:int(off[-1])]
