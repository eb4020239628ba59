"""Seam S3: drop-in for ``sstar.sstar.Sstar`` (reference sstar/sstar.py:26-106).

Same name, same keyword contract, same return layout and the same error for a
missing argument -- but the arithmetic runs on the GPU through the C-ABI
(``sstar_b200_score_windows_ex``); there is no CPU path.
"""
from __future__ import annotations

from typing import Any

import numpy as np

from .engine import NAN_SENTINEL, get_engine


class Sstar:
    """S* statistic for one window (all target columns), computed on the B200."""

    @staticmethod
    def _require(kwargs: dict[str, Any], *names: str) -> tuple[Any, ...]:
        missing = [n for n in names if n not in kwargs]
        if missing:
            raise ValueError(f"Missing required arguments: {missing}")
        return tuple(kwargs[n] for n in names)

    @staticmethod
    def compute(**kwargs) -> dict[str, Any]:
        """``ref_gts[Vw,R]``, ``tgt_gts[Vw,T]``, ``pos[Vw]`` (+ optional ``match_bonus``,
        ``max_mismatch``, ``mismatch_penalty``) -> ``{"sstar": [...], "region_ind_SNP_number": [...]}``.

        ``pos`` must be strictly ascending (the reference's own inputs always are:
        sstar/utils/utils.py:140-164 rejects duplicates, VCF / msprime order is sorted).
        Extra keyword ``device`` selects the GPU (default 0), ``algo`` the kernel family.
        """
        ref_gts, tgt_gts, pos = Sstar._require(kwargs, "ref_gts", "tgt_gts", "pos")
        bonus = kwargs.get("match_bonus", 5000)
        max_mm = kwargs.get("max_mismatch", 5)
        penalty = kwargs.get("mismatch_penalty", -10000)

        ref = np.asarray(ref_gts)
        tgt = np.asarray(tgt_gts)
        p = np.asarray(pos)
        if tgt.ndim != 2:
            raise ValueError("tgt_gts must be 2-D (n_sites, n_samples)")
        T = tgt.shape[1]
        if T == 0:
            return {"sstar": [], "region_ind_SNP_number": []}
        if ref.ndim == 2 and ref.shape[1] == 0:  # empty panel: every signed row sum is 0
            ref = np.zeros((ref.shape[0], 1), dtype=np.int8)
        lo = int(p[0]) if p.size else 0
        hi = int(p[-1]) if p.size else 0
        eng = get_engine(int(kwargs.get("device", 0)))
        score, nsnp = eng.score_host(ref, tgt, p, [lo], [hi], int(bonus), int(max_mm), int(penalty),
                                     algo=kwargs.get("algo", "auto"))
        scores = [np.nan if s == NAN_SENTINEL else float(s) for s in score[0].tolist()]
        return {"sstar": scores, "region_ind_SNP_number": [int(x) for x in nsnp[0].tolist()]}
