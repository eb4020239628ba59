"""S* SNP sets on the host side: the legacy score table with S*_SNP_number / S*_SNPs.

The current reference's feature table stops at ``Region_ind_SNP_number``
(sstar/feature_vector_preprocessor.py:167-181); the older ``sstar score`` output -- kept by the
reference only as fixtures (tests/exp_results/test.score.{unphased,phased}.exp.results.tsv) --
also listed the sites on the best chain.  ``ScoreEngine.snp_sets_device`` recovers them on the GPU;
this module turns its ragged result into that table.
"""
from __future__ import annotations

import numpy as np

from .engine import NAN_SENTINEL, get_engine

LEGACY_HEADER = "chrom\tstart\tend\tsample\tS*_score\tregion_ind_SNP_number\tS*_SNP_number\tS*_SNPs"


def legacy_score_table(chrom, win_start, win_end, labels, score, nsnp, set_off, set_pos) -> str:
    """Rows in the legacy order (sample-major, windows ascending), 0-based window starts."""
    score = np.asarray(score)
    nsnp = np.asarray(nsnp)
    W, T = score.shape
    if len(labels) != T:
        raise ValueError(f"{len(labels)} labels for {T} columns")
    lines = [LEGACY_HEADER]
    for t in range(T):
        for w in range(W):
            u = w * T + t
            head = f"{chrom}\t{int(win_start[w]) - 1}\t{int(win_end[w])}\t{labels[t]}"
            if score[w, t] == NAN_SENTINEL:
                lines.append(f"{head}\tNA\t{int(nsnp[w, t])}\tNA\tNA")
            else:
                chain = set_pos[int(set_off[u]):int(set_off[u + 1])]
                lines.append(f"{head}\t{int(score[w, t])}\t{int(nsnp[w, t])}\t{len(chain)}\t"
                             + ",".join(str(int(x)) for x in chain))
    return "\n".join(lines) + "\n"


def score_with_snp_sets(chrom, ref_gts, tgt_gts, pos, win_start, win_end, labels, match_bonus=5000,
                        max_mismatch=5, mismatch_penalty=-10000, device: int = 0) -> str:
    """numpy in, legacy table text out (scoring and backtracking on the GPU)."""
    out = get_engine(device).snp_sets_host(ref_gts, tgt_gts, pos, win_start, win_end, match_bonus,
                                           max_mismatch, mismatch_penalty)
    return legacy_score_table(chrom, win_start, win_end, labels, *out)
