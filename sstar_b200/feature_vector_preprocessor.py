"""The scheduler "job" (reference sstar/feature_vector_preprocessor.py:26-183).

``run`` keeps the reference's per-window contract (one window in, ``T`` row dicts out, through the
monkeypatchable ``Sstar.compute`` seam).  ``run_chromosome`` is the batched form the GPU wants:
every window of a chromosome in one device call, optionally sharded over several GPUs.
"""
from __future__ import annotations

from typing import Any

import numpy as np

from .feature_table import FeatureTable, sample_labels
from .sstar import Sstar


class FeatureVectorPreprocessor:
    def __init__(self, ref_samples: list[str], tgt_samples: list[str], match_bonus: int,
                 max_mismatch: int, mismatch_penalty: int):
        if len(ref_samples) == 0:
            raise ValueError("No reference sample is provided.")
        if len(tgt_samples) == 0:
            raise ValueError("No target sample is provided.")
        self.match_bonus = match_bonus
        self.max_mismatch = max_mismatch
        self.mismatch_penalty = mismatch_penalty
        self.samples = {"ref": ref_samples, "tgt": tgt_samples}

    # ---- reference contract: one window -> T dicts
    def run(self, chr_name: str, start: int, end: int, ploidy: int, is_phased: bool,
            ref_gts: np.ndarray, tgt_gts: np.ndarray, pos: np.ndarray) -> list[dict[str, Any]]:
        res = {"chr_name": chr_name, "start": start, "end": end}
        res.update(Sstar.compute(ref_gts=ref_gts, tgt_gts=tgt_gts, pos=pos,
                                 match_bonus=self.match_bonus, max_mismatch=self.max_mismatch,
                                 mismatch_penalty=self.mismatch_penalty))
        return self._fmt_res(res=res, ploidy=ploidy, is_phased=is_phased)

    def _fmt_res(self, res: dict[str, Any], ploidy: int, is_phased: bool) -> list[dict[str, Any]]:
        labels = sample_labels(self.samples["tgt"], ploidy, is_phased)
        return [{"Chromosome": res["chr_name"], "Start": res["start"], "End": res["end"],
                 "Sample": lab, "S*_score": res["sstar"][i],
                 "Region_ind_SNP_number": res["region_ind_SNP_number"][i]}
                for i, lab in enumerate(labels)]

    # ---- batched GPU contract: one chromosome -> FeatureTable
    def run_chromosome(self, chr_name: str, ploidy: int, is_phased: bool, ref_gts: np.ndarray,
                       tgt_gts: np.ndarray, pos: np.ndarray, win_start, win_end, devices=None,
                       algo: str = "auto") -> FeatureTable:
        from .sharding import score_chromosome
        labels = sample_labels(self.samples["tgt"], ploidy, is_phased)
        if np.asarray(tgt_gts).shape[1] != len(labels):
            raise ValueError(f"tgt_gts has {np.asarray(tgt_gts).shape[1]} columns but {len(labels)} "
                             "sample labels follow from the target list / ploidy / phasing")
        score, nsnp = score_chromosome(ref_gts, tgt_gts, pos, win_start, win_end, self.match_bonus,
                                       self.max_mismatch, self.mismatch_penalty, devices=devices, algo=algo)
        return FeatureTable(chr_name, win_start, win_end, labels, score, nsnp)
