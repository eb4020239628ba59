"""CPU restatement of the S* operator -- TEST INFRASTRUCTURE (see oracle/__init__.py).

Follows ``sstar/sstar.py`` of the reference (paths relative to /root/reference):

* ``Sstar.compute``                    sstar/sstar.py:43-106  -> :func:`compute_float`
* ``_create_matrices``                 sstar/sstar.py:156-181 -> :func:`_focal_sites`, :func:`_pair_matrix`
* ``_calc_ind_sstar``                  sstar/sstar.py:183-217 -> :func:`_pair_matrix`, :func:`_chain_float`
* sample loop                          sstar/sstar.py:219-244 -> loop in :func:`compute_float`
* window slicing                       sstar/generators/window_data_generator.py:106-129 -> :func:`score_windows`

Two restatements are kept:

``compute_float``   same numeric recipe as the reference (float64 matrices holding
                    integers, ``np.union1d`` for the SNP count, a Python double loop
                    for the chain) -- this is the variant timed as the CPU baseline,
                    because it has the reference's cost model.
``compute_int``     the S-SPEC of SURVEY.md section 8 in pure integer arithmetic, used to
                    cross-check ``compute_float`` and to define the NaN sentinel
                    (``NAN_SENTINEL`` = INT64_MIN) that the C-ABI uses.
"""
from __future__ import annotations

import numpy as np

NAN_SENTINEL = np.iinfo(np.int64).min
MIN_PAIR_DISTANCE = 10  # sstar/sstar.py:195 -- pairs closer than 10 bp never link


def _need(kwargs, *names):
    """Required-kwarg check with the reference's message (sstar/sstar.py:36-40)."""
    absent = [n for n in names if n not in kwargs]
    if absent:
        raise ValueError(f"Missing required arguments: {absent}")
    return [kwargs[n] for n in names]


def _focal_sites(tgt_col: np.ndarray, absent_pos: np.ndarray):
    """Carried reference-absent sites of one target column (sstar/sstar.py:162-167)."""
    keep = tgt_col != 0
    return absent_pos[keep], tgt_col[keep]


def _pair_matrix(p: np.ndarray, g: np.ndarray, bonus, max_mm, penalty) -> np.ndarray:
    """n x n pair-score matrix, float64, as sstar/sstar.py:174-179 and :195-203."""
    pf = p.astype(np.float64)
    gf = g.astype(np.float64)
    dist = np.abs(pf.reshape(-1, 1) - pf.reshape(1, -1))
    gdist = np.abs(gf.reshape(-1, 1) - gf.reshape(1, -1))
    near = dist < MIN_PAIR_DISTANCE
    far = np.logical_not(near)
    s = np.array(dist, copy=True)
    s[near] = -np.inf
    s[far & (gdist == 0)] += bonus
    s[far & (gdist > 0) & (gdist <= max_mm)] = penalty
    s[far & (gdist > max_mm)] = -np.inf
    return s


def _chain_float(s: np.ndarray) -> float:
    """Max-plus chain over the lower triangle (sstar/sstar.py:205-217)."""
    n = s.shape[0]
    ending_at = np.full(n, -np.inf, dtype=np.float64)
    for j in range(n):
        top = -np.inf
        for i in range(j):
            top = max(top, ending_at[i] + s[j, i], s[j, i])
        ending_at[j] = top
    if np.isneginf(ending_at).all():
        return np.nan
    return float(np.nanmax(ending_at))


def compute_float(**kwargs) -> dict:
    """Reference-recipe S* for one window; returns the reference's dict layout."""
    ref_gts, tgt_gts, pos = _need(kwargs, "ref_gts", "tgt_gts", "pos")
    bonus = kwargs.get("match_bonus", 5000)
    max_mm = kwargs.get("max_mismatch", 5)
    penalty = kwargs.get("mismatch_penalty", -10000)

    ref_rowsum = np.sum(ref_gts, axis=1)  # signed: sstar/sstar.py:87-88
    absent = ref_rowsum == 0
    poly_pos = pos[ref_rowsum != 0]
    absent_pos = pos[absent]
    absent_gt = tgt_gts[absent]

    scores, counts = [], []
    for col in range(absent_gt.shape[1]):
        p, g = _focal_sites(absent_gt[:, col], absent_pos)
        counts.append(int(np.union1d(poly_pos, p).size))  # sstar/sstar.py:171
        if p.size <= 2:  # sstar/sstar.py:191-193
            scores.append(np.nan)
            continue
        scores.append(_chain_float(_pair_matrix(p, g, bonus, max_mm, penalty)))
    return {"sstar": scores, "region_ind_SNP_number": counts}


def chain_int(p, g, bonus: int, max_mm: int, penalty: int):
    """Integer S-SPEC chain for one column. Returns (score or None, argmax chain).

    The optional SNP set follows SURVEY section 8 S-SPEC item 6 (first i wins, extension
    preferred over a fresh start, first j wins); it is not part of the reference
    output and only pinned by the orphaned legacy fixtures.
    """
    n = len(p)
    if n <= 2:
        return None, []
    best = [None] * n  # None == -inf
    back = [-1] * n
    for j in range(n):
        for i in range(j):
            d = int(p[j]) - int(p[i])
            q = abs(int(g[j]) - int(g[i]))
            if d < MIN_PAIR_DISTANCE or q > max_mm:
                continue
            s = d + bonus if q == 0 else penalty
            cand = s
            link = -1
            if best[i] is not None and best[i] + s >= s:
                cand = best[i] + s
                link = i
            if best[j] is None or cand > best[j]:
                best[j] = cand
                back[j] = link if link >= 0 else -(i + 2)  # fresh start at i
    top, top_j = None, -1
    for j in range(n):
        if best[j] is not None and (top is None or best[j] > top):
            top, top_j = best[j], j
    if top is None:
        return None, []
    chain = []
    j = top_j
    while True:
        chain.append(j)
        b = back[j]
        if b >= 0:
            j = b
        else:
            chain.append(-b - 2)
            break
    chain.reverse()
    return top, chain


def compute_int(ref_gts, tgt_gts, pos, match_bonus=5000, max_mismatch=5,
                mismatch_penalty=-10000):
    """Integer S-SPEC for one window -> (score[T] int64 with NAN_SENTINEL, nsnp[T] int32)."""
    ref_gts = np.asarray(ref_gts)
    tgt_gts = np.asarray(tgt_gts)
    pos = np.asarray(pos)
    rowsum = ref_gts.astype(np.int64).sum(axis=1)
    absent = rowsum == 0
    poly = set(int(x) for x in pos[~absent])
    T = tgt_gts.shape[1]
    score = np.full(T, NAN_SENTINEL, dtype=np.int64)
    nsnp = np.zeros(T, dtype=np.int32)
    apos = pos[absent]
    agt = tgt_gts[absent]
    for t in range(T):
        col = agt[:, t]
        keep = col != 0
        p = [int(x) for x in apos[keep]]
        g = [int(x) for x in col[keep]]
        nsnp[t] = len(poly.union(p))
        s, _ = chain_int(p, g, int(match_bonus), int(max_mismatch), int(mismatch_penalty))
        if s is not None:
            score[t] = s
    return score, nsnp


def float_scores_to_int(scores) -> np.ndarray:
    """Map the reference's float list (NaN = no score) onto the int64 sentinel form."""
    a = np.asarray(scores, dtype=np.float64)
    out = np.full(a.shape, NAN_SENTINEL, dtype=np.int64)
    ok = ~np.isnan(a)
    out[ok] = a[ok].astype(np.int64)
    assert np.all(a[ok] == out[ok]), "reference score is not an exact integer"
    return out


def score_windows(ref_gts, tgt_gts, pos, win_start, win_end, match_bonus=5000,
                  max_mismatch=5, mismatch_penalty=-10000, impl="float"):
    """All windows of one chromosome -> (score[W,T] int64, nsnp[W,T] int32).

    Window slicing is the reference's closed-interval boolean mask over every
    variant (sstar/generators/window_data_generator.py:113-116).
    """
    ref_gts = np.asarray(ref_gts)
    tgt_gts = np.asarray(tgt_gts)
    pos = np.asarray(pos)
    W, T = len(win_start), tgt_gts.shape[1]
    score = np.full((W, T), NAN_SENTINEL, dtype=np.int64)
    nsnp = np.zeros((W, T), dtype=np.int32)
    for w in range(W):
        inside = (pos >= win_start[w]) & (pos <= win_end[w])
        args = dict(ref_gts=ref_gts[inside], tgt_gts=tgt_gts[inside], pos=pos[inside],
                    match_bonus=match_bonus, max_mismatch=max_mismatch,
                    mismatch_penalty=mismatch_penalty)
        if impl == "float":
            out = compute_float(**args)
            score[w] = float_scores_to_int(out["sstar"])
            nsnp[w] = out["region_ind_SNP_number"]
        else:
            score[w], nsnp[w] = compute_int(**args)
    return score, nsnp


def snp_sets_windows(ref_gts, tgt_gts, pos, win_start, win_end, match_bonus=5000, max_mismatch=5,
                     mismatch_penalty=-10000):
    """Per unit (w, t), in that order: (score or None, [positions on the best chain]) -- the legacy
    columns S*_SNP_number / S*_SNPs (reference fixtures tests/exp_results/test.score.*.exp.results.tsv),
    rule of SURVEY section 8a item 6.  Pure Python: small cases only."""
    ref_gts = np.asarray(ref_gts)
    tgt_gts = np.asarray(tgt_gts)
    pos = np.asarray(pos)
    out = []
    for w in range(len(win_start)):
        inside = (pos >= win_start[w]) & (pos <= win_end[w])
        rg, tg, pp = ref_gts[inside], tgt_gts[inside], pos[inside]
        absent = rg.astype(np.int64).sum(axis=1) == 0
        apos, agt = pp[absent], tg[absent]
        for t in range(tgt_gts.shape[1]):
            keep = agt[:, t] != 0
            p = [int(x) for x in apos[keep]]
            g = [int(x) for x in agt[keep, t]]
            s, chain = chain_int(p, g, int(match_bonus), int(max_mismatch), int(mismatch_penalty))
            out.append((s, [p[k] for k in chain]))
    return out
