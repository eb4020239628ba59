"""CPU restatement of the reference's tract match rate -- TEST INFRASTRUCTURE.

Follows sstar/match.py: ``calc_match_pct`` (:30-73), ``_dosage_for_position_index`` (:88-162, with
scikit-allel's ``to_n_alt`` = number of alleles > 0 and ``is_missing`` = any allele < 0),
``_positions_in_bed_interval`` (:408-427) and the tract x source loop of ``run_match`` (:497-545).
scikit-allel is absent here, so ``run_match`` itself cannot be executed; parity is pinned on the
known answers of the reference's tests (tests/test_match.py:67-77, 187-262, 315-455), restated in
tests/test_match_port.py.
"""
from __future__ import annotations

import numpy as np


def calc_match_pct(x, y, denominator, P=2):
    if denominator == 0:
        return "NA"
    x = np.atleast_1d(np.asarray(x, dtype=float))
    y = np.atleast_1d(np.asarray(y, dtype=float))
    ok = (x >= 0) & (y >= 0)
    score_sum = np.sum(1.0 - np.abs(x[ok] - y[ok]) / float(P))
    return float(score_sum / denominator)


def dosage(gt, pos, ind_index, positions, hap_index=None, P=2):
    """gt int [V, N, 2]; positions absent from ``pos`` and missing genotypes give -1."""
    gt = np.asarray(gt)
    index = {int(p): i for i, p in enumerate(pos)}
    out = np.full(len(positions), -1.0)
    for j, p in enumerate(positions):
        i = index.get(int(p))
        if i is None:
            continue
        a = gt[i, ind_index]
        if hap_index is None:
            out[j] = -1.0 if (a < 0).any() else float((a > 0).sum())
        else:
            out[j] = float(a[hap_index]) * P if a[hap_index] >= 0 else -1.0
    return out


def tract_rate(gt, pos, tgt_index, hap_index, src_indices, start, end, P=2):
    pos = np.asarray(pos)
    region = pos[np.searchsorted(pos, start, side="right"):np.searchsorted(pos, end, side="right")]
    positions = region.astype(int).tolist()
    if len(positions) == 0:
        return "NA"
    tgt_dos = dosage(gt, pos, tgt_index, positions, hap_index, P)
    rates = [calc_match_pct(tgt_dos, dosage(gt, pos, s, positions, None, P), len(positions), P) for s in src_indices]
    return float(np.mean(rates))
