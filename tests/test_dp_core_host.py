"""The device DP core (sstar_b200/csrc/dp_core.cuh) compiled for the host and driven through a
scalar emulation of score_group_kernel (tests/emul/emul_score.cpp), against the C oracle.

Runs without a GPU: it pins the arithmetic of the kernels (list DP, waiting-site ring, walk DP,
int32/int64 switch, group logic) before any GPU time is spent.  Bit-exact."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from oracle import c_oracle
from sstar_b200.synth import regular_windows, synth_chromosome

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "emul", "emul_score.cpp")
LIB = os.path.join(HERE, "emul", "libemul_score.so")
CORE = os.path.join(os.path.dirname(HERE), "sstar_b200", "csrc", "dp_core.cuh")


@pytest.fixture(scope="module")
def emul():
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(SRC), os.path.getmtime(CORE)):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", LIB, SRC])
    lib = ctypes.CDLL(LIB)
    lib.emul_score_windows.restype = ctypes.c_int
    lib.emul_score_windows.argtypes = [ctypes.c_void_p] * 3 + [ctypes.c_int64, ctypes.c_int32, ctypes.c_int32,
                                                               ctypes.c_void_p, ctypes.c_void_p] + \
        [ctypes.c_int32] * 7 + [ctypes.c_void_p] * 4

    def run(ref, tgt, pos, ws, we, params=(5000, 5, -10000), wg=8, L=64, words_cap=64):
        ref = np.ascontiguousarray(ref, dtype=np.int8)
        tgt = np.ascontiguousarray(tgt, dtype=np.int8)
        pos = np.ascontiguousarray(pos, dtype=np.int32)
        ws = np.ascontiguousarray(ws, dtype=np.int32)
        we = np.ascontiguousarray(we, dtype=np.int32)
        V, R = ref.shape
        T = tgt.shape[1]
        W = len(ws)
        score = np.empty((W, T), np.int64)
        nsnp = np.empty((W, T), np.int32)
        slow = np.zeros(W, np.uint8)
        path = np.zeros((W, T), np.int32)
        rc = lib.emul_score_windows(ref.ctypes.data, tgt.ctypes.data, pos.ctypes.data, V, R, T, ws.ctypes.data,
                                    we.ctypes.data, W, *params, wg, L, words_cap, score.ctypes.data,
                                    nsnp.ctypes.data, slow.ctypes.data, path.ctypes.data)
        assert rc == 0
        return score, nsnp, slow.astype(bool), path
    return run


def check(emul, ref, tgt, pos, ws, we, params=(5000, 5, -10000), **kw):
    s, n, slow, path = emul(ref, tgt, pos, ws, we, params, **kw)
    so, no = c_oracle.score_windows(ref, tgt, pos, ws, we, *params)
    keep = ~slow
    assert np.array_equal(s[keep], so[keep])
    assert np.array_equal(n[keep], no[keep])
    return path[keep], slow


@pytest.mark.parametrize("wg", [1, 3, 8, 16])
@pytest.mark.parametrize("shape", [(100, 50, False, 1_500_000), (5, 10, False, 500_000), (20, 33, True, 800_000)])
def test_list_path_matches_oracle(emul, wg, shape):
    n_ref, n_tgt, phased, length = shape
    ref, tgt, pos = synth_chromosome(length, n_ref, n_tgt, phased, seed=500 + n_ref)
    ws, we = regular_windows(pos, 50_000, 10_000)
    path, slow = check(emul, ref, tgt, pos, ws, we, wg=wg, L=512)
    assert not slow.any() and (path[path > 0] == 1).all()


def test_list_overflow_falls_back_per_column(emul):
    ref, tgt, pos = synth_chromosome(600_000, 5, 10, False, seed=9)
    ws, we = regular_windows(pos, 50_000, 10_000)
    path, _ = check(emul, ref, tgt, pos, ws, we, wg=8, L=24)
    assert (path == 2).any() and (path == 1).any()


def test_crowded_positions_exercise_the_ring(emul):
    # many carried sites within 9 bp of each other: several sites wait at once
    rng = np.random.default_rng(4)
    V = 4000
    pos = np.cumsum(rng.choice([1, 1, 2, 3, 9, 10, 11, 40], size=V)).astype(np.int32)
    tgt = rng.choice([0, 1, 2], size=(V, 7), p=[0.2, 0.5, 0.3]).astype(np.int8)
    ref = np.zeros((V, 2), np.int8)
    ref[rng.random(V) < 0.15, 0] = 1
    ws = np.arange(1, int(pos[-1]), 300, dtype=np.int32)
    we = ws + 1499
    for params in [(5000, 5, -10000), (1, 1, 0), (7, 0, -3), (123, 4, -7)]:
        path, _ = check(emul, ref, tgt, pos, ws, we, params, wg=8, L=1024)
        assert (path == 1).mean() > 0.9
        check(emul, ref, tgt, pos, ws, we, params, wg=4, L=8)       # per-column overflow -> walk DP
        check(emul, ref, tgt, pos, ws, we, params, wg=8, words_cap=2)  # group too wide -> walk DP


def test_missing_data_takes_the_general_dp(emul):
    # negative dosages (missing calls) and signed reference sums: the per-class DP, not the slow list
    ref, tgt, pos = synth_chromosome(500_000, 10, 24, False, seed=77)
    rng = np.random.default_rng(5)
    tgt = tgt.copy()
    miss = rng.random(tgt.shape) < 0.01
    tgt[miss] = rng.choice([-2, -1], size=int(miss.sum())).astype(np.int8)
    ref = ref.copy()
    ref[rng.random(ref.shape) < 0.002] = -1
    ws, we = regular_windows(pos, 50_000, 10_000)
    for params in [(5000, 5, -10000), (5000, 0, -10000), (5000, 2, -10000), (1, 1, 0), (123, 3, -7)]:
        for wg in (1, 8):
            path, slow = check(emul, ref, tgt, pos, ws, we, params, wg=wg, L=512)
            # -1 / -2 only: the genotype bytes come from the three planes, the classes live in registers
            assert not slow.any() and (path == 6).mean() > 0.8
    # a sprinkle of multi-allelic dosages on top: those sub-groups read the bytes (per-class tables)
    tgt3 = tgt.copy()
    tgt3[rng.random(tgt.shape) < 0.0005] = 3
    path, slow = check(emul, ref, tgt3, pos, ws, we, wg=8, L=512)
    assert not slow.any() and (path == 5).any() and (path == 6).any()


def test_small_signed_dp_crowding_and_every_mismatch_bound(emul):
    # sites within 9 bp of each other (several waiting at once) holding -2 / -1 / 1 / 2, every max_mismatch
    # that changes which classes may follow one another (differences are 1, 2, 3 and 4)
    rng = np.random.default_rng(16)
    V = 3000
    pos = np.cumsum(rng.choice([1, 2, 3, 9, 10, 11, 40], size=V)).astype(np.int32)
    ref = np.zeros((V, 2), np.int8)
    ref[rng.random(V) < 0.1, 0] = 1
    ws = np.arange(1, int(pos[-1]), 400, dtype=np.int32)
    we = ws + 1999
    tgt = rng.choice([0, -2, -1, 1, 2], size=(V, 6), p=[0.4, 0.1, 0.15, 0.2, 0.15]).astype(np.int8)
    for mm in range(0, 6):
        for bonus, pen in [(5000, -10000), (7, -3), (0, 5)]:
            path, slow = check(emul, ref, tgt, pos, ws, we, (bonus, mm, pen), wg=4, L=2048)
            assert not slow.any() and (path == 6).all()


def test_general_dp_many_classes_and_crowding(emul):
    rng = np.random.default_rng(6)
    V = 3000
    pos = np.cumsum(rng.choice([1, 2, 3, 9, 10, 11, 40], size=V)).astype(np.int32)
    ref = np.zeros((V, 2), np.int8)
    ref[rng.random(V) < 0.1, 0] = 1
    ws = np.arange(1, int(pos[-1]), 400, dtype=np.int32)
    we = ws + 1999
    # up to 7 distinct genotype values: still the general DP, with several sites waiting at once
    tgt = rng.choice([0, -2, -1, 1, 2, 3, 4, 6], size=(V, 6), p=[0.3, 0.1, 0.1, 0.2, 0.15, 0.05, 0.05, 0.05]).astype(np.int8)
    for params in [(5000, 5, -10000), (7, 1, -3), (50, 8, -20)]:
        path, slow = check(emul, ref, tgt, pos, ws, we, params, wg=4, L=2048)
        assert not slow.any() and (path == 5).all()
    # more than 8 distinct values in one unit: that window is handed to the pairwise kernel
    tgt2 = tgt.copy()
    tgt2[:, 0] = rng.integers(-60, 60, size=V).astype(np.int8)
    path, slow = check(emul, ref, tgt2, pos, ws, we, wg=4, L=2048)
    assert slow.any()
    # a column that overflows its list in general mode also routes its windows there
    check(emul, ref, tgt, pos, ws, we, wg=4, L=16)


def test_int64_switch_and_odd_windows(emul):
    rng = np.random.default_rng(8)
    V = 3000
    pos = np.cumsum(rng.integers(1, 60, size=V)).astype(np.int32)
    tgt = rng.choice([0, 1, 2], size=(V, 5), p=[0.5, 0.3, 0.2]).astype(np.int8)
    ref = (rng.random((V, 3)) < 0.2).astype(np.int8)
    # non-monotone, empty, nested and out-of-range windows in one group
    ws = np.array([5000, 1, 70000, 300, 300, 10_000_000, -5, 40000, 20000], dtype=np.int32)
    we = np.array([9000, 90000, 70001, 200, 5000, 10_000_100, 2_000_000_000, 45000, 20000], dtype=np.int32)
    check(emul, ref, tgt, pos, ws, we, wg=4)
    path, _ = check(emul, ref, tgt, pos, ws, we, (2_000_000_000, 5, -2_000_000_000), wg=4)
    assert (path == 4).any()
    mono_s = np.sort(ws[[0, 2, 4, 7, 8]])
    check(emul, ref, tgt, pos, mono_s, mono_s + 4000, (2_000_000_000, 5, -2_000_000_000), wg=4)


def test_replicate_style_disjoint_windows(emul):
    ref, tgt, pos = synth_chromosome(800_000, 5, 10, False, seed=21)
    ws = np.arange(1, 800_000, 50_000, dtype=np.int32)
    we = ws + 49_999
    path, _ = check(emul, ref, tgt, pos, ws, we, wg=8, L=512, words_cap=64)
    assert (path[path > 0] == 1).any()


# ---- the fused kernel's unit paths (fused_kernels.cuh), scalar emulation ----------------------
@pytest.fixture(scope="module")
def emul_fused(emul):
    lib = ctypes.CDLL(LIB)
    lib.emul_fused_windows.restype = ctypes.c_int
    lib.emul_fused_windows.argtypes = [ctypes.c_void_p] * 3 + [ctypes.c_int64, ctypes.c_int32, ctypes.c_int32,
                                                               ctypes.c_void_p, ctypes.c_void_p] + \
        [ctypes.c_int32] * 5 + [ctypes.c_void_p] * 3

    def run(ref, tgt, pos, ws, we, params=(5000, 5, -10000), force_stream=0):
        ref = np.ascontiguousarray(ref, dtype=np.int8)
        tgt = np.ascontiguousarray(tgt, dtype=np.int8)
        pos = np.ascontiguousarray(pos, dtype=np.int32)
        ws = np.ascontiguousarray(ws, dtype=np.int32)
        we = np.ascontiguousarray(we, dtype=np.int32)
        V, R = ref.shape
        T, W = tgt.shape[1], len(ws)
        score = np.empty((W, T), np.int64)
        nsnp = np.empty((W, T), np.int32)
        path = np.zeros((W, T), np.int32)
        assert lib.emul_fused_windows(ref.ctypes.data, tgt.ctypes.data, pos.ctypes.data, V, R, T, ws.ctypes.data,
                                      we.ctypes.data, W, *params, force_stream, score.ctypes.data, nsnp.ctypes.data,
                                      path.ctypes.data) == 0
        so, no = c_oracle.score_windows(ref, tgt, pos, ws, we, *params)
        assert np.array_equal(score, so) and np.array_equal(nsnp, no)
        return path
    return run


@pytest.mark.parametrize("params", [(5000, 5, -10000), (5000, 0, -10000), (1, 1, 0), (123, 4, -7), (5000, 300, -10000),
                                    (2 ** 27, 2, -(2 ** 27))])
def test_fused_unit_paths_match_oracle(emul_fused, params):
    rng = np.random.default_rng(abs(hash(params)) % 2 ** 31)
    ref, tgt, pos = synth_chromosome(500_000, 12, 9, False, seed=41)
    ws, we = regular_windows(pos, 50_000, 10_000)
    assert set(np.unique(emul_fused(ref, tgt, pos, ws, we, params))) <= {0, 1, 2}
    assert set(np.unique(emul_fused(ref, tgt, pos, ws, we, params, force_stream=1))) <= {0, 4}
    # missing calls / multi-allelic dosages: a few genotype values -> the general walk
    few = tgt.copy()
    few[rng.random(few.shape) < 0.03] = -1
    few[rng.random(few.shape) < 0.01] = 3
    p = emul_fused(ref, few, pos, ws, we, params)
    assert 3 in p and 4 not in p
    emul_fused(ref, few, pos, ws, we, params, force_stream=1)
    # many genotype values per unit -> the class tables, incl. the int8 extremes and crowded sites
    many = tgt.copy().astype(np.int16)
    hit = rng.random(many.shape) < 0.4
    many[hit] = rng.integers(-128, 128, size=int(hit.sum()))
    many = many.astype(np.int8)
    dense_pos = np.sort(rng.choice(np.arange(1, 4 * len(pos)), size=len(pos), replace=False)).astype(np.int32)
    dws, dwe = regular_windows(dense_pos, 500, 100)
    p = emul_fused(np.zeros_like(ref), many, dense_pos, dws[:200], dwe[:200], params)
    assert 4 in p
    # clean genotypes on crowded positions (several sites inside 9 bp): the parked-site ring of the register walk
    p = emul_fused(np.zeros_like(ref), tgt, dense_pos, dws[:300], dwe[:300], params)
    assert set(np.unique(p)) <= {0, 1, 2}


# ---- the O(n) chain DP with backpointers (SNP sets) against the oracle's literal chains ----------
@pytest.mark.parametrize("params", [(5000, 5, -10000), (5000, 0, -10000), (1, 1, 0), (0, 3, 0), (123, 4, -7), (10, 1, -10),
                                    (2 ** 27, 2, -(2 ** 27))])
def test_chain_backpointers_match_oracle(emul, params):
    lib = ctypes.CDLL(LIB)
    lib.emul_chain_windows.restype = ctypes.c_int
    lib.emul_chain_windows.argtypes = [ctypes.c_void_p] * 3 + [ctypes.c_int64, ctypes.c_int32, ctypes.c_int32,
                                                               ctypes.c_void_p, ctypes.c_void_p] + \
        [ctypes.c_int32] * 4 + [ctypes.c_void_p] * 4
    CAP = 128
    rng = np.random.default_rng(abs(hash(params)) % 2 ** 31)

    def run(ref, tgt, pos, ws, we):
        ref = np.ascontiguousarray(ref, dtype=np.int8); tgt = np.ascontiguousarray(tgt, dtype=np.int8)
        pos = np.ascontiguousarray(pos, dtype=np.int32)
        ws = np.ascontiguousarray(ws, dtype=np.int32); we = np.ascontiguousarray(we, dtype=np.int32)
        V, R = ref.shape
        T, W = tgt.shape[1], len(ws)
        score = np.empty(W * T, np.int64); ln = np.empty(W * T, np.int32)
        sp = np.zeros(W * T * CAP, np.int32); fast = np.zeros(W * T, np.uint8)
        assert lib.emul_chain_windows(ref.ctypes.data, tgt.ctypes.data, pos.ctypes.data, V, R, T, ws.ctypes.data, we.ctypes.data,
                                      W, *params, score.ctypes.data, ln.ctypes.data, sp.ctypes.data, fast.ctypes.data) == 0
        so, no, off, sets = c_oracle.snp_sets(ref, tgt, pos, ws, we, *params)
        nfast = 0
        for u in range(W * T):
            if not fast[u]:
                continue
            nfast += 1
            assert score[u] == so.reshape(-1)[u], u
            want = sets[off[u]:off[u + 1]]
            assert ln[u] == len(want), (u, ln[u], len(want))
            assert np.array_equal(sp[u * CAP:u * CAP + ln[u]], want), u
        return nfast, W * T

    ref, tgt, pos = synth_chromosome(500_000, 12, 9, False, seed=43)
    ws, we = regular_windows(pos, 50_000, 10_000)
    nf, tot = run(ref, tgt, pos, ws, we)
    assert nf == tot
    # tie-heavy: equally spaced sites, few genotype classes, every site reference-absent
    V = 400
    tpos = (np.arange(V) * 10 + 1).astype(np.int32)
    ttgt = rng.integers(0, 3, size=(V, 7)).astype(np.int8)
    tws = np.arange(1, 3500, 250, dtype=np.int32)
    nf, tot = run(np.zeros((V, 2), np.int8), ttgt, tpos, tws, tws + 999)
    assert nf == tot
    # crowded positions and long units (some beyond the cap are skipped, the rest must agree)
    dpos = np.sort(rng.choice(np.arange(1, 3000), size=900, replace=False)).astype(np.int32)
    dtgt = (rng.random((900, 5)) < 0.5).astype(np.int8) * rng.integers(1, 3, size=(900, 5)).astype(np.int8)
    nf, tot = run(np.zeros((900, 2), np.int8), dtgt, dpos, [1, 1, 500, 1500], [400, 3000, 900, 1700])
    assert 0 < nf < tot
