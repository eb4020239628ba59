"""Window tiling (reference sstar/generators/window_data_generator.py:30-151).

Same constructor and the same per-window dicts from ``get()``, but nothing is sliced up front:
the whole chromosome stays as ONE pair of genotype matrices plus a window list, which is what the
GPU path consumes (``chromosome()``); ``get()`` cuts the reference's per-window views lazily for
callers that still want them (the reference copies every window eagerly with an O(W*V) mask).
"""
from __future__ import annotations

import numpy as np

from ..utils import read_data, split_genome
from .generic_generator import GenericGenerator


class WindowDataGenerator(GenericGenerator):
    def __init__(self, vcf_file: str, chr_name: str, ref_ind_file: str, tgt_ind_file: str, win_len: int,
                 win_step: int, anc_allele_file: str = None, ploidy: int = 2, is_phased: bool = True):
        if win_len <= 0:
            raise ValueError("win_len must be greater than 0.")
        if win_step < 0:
            raise ValueError("win_step must be non-negative.")
        if ploidy <= 0:
            raise ValueError("ploidy must be greater than 0.")
        self.ploidy = ploidy
        self.is_phased = is_phased
        self.chr_name = chr_name

        ref_data, _, tgt_data, _ = read_data(vcf_file, ref_ind_file, tgt_ind_file, anc_allele_file,
                                             is_phased, chr_name=chr_name)
        if chr_name not in tgt_data:
            raise ValueError(f"{chr_name} is not present in the VCF file.")
        self.pos = tgt_data[chr_name]["POS"]
        self.ref_gts = ref_data[chr_name]["GT"]
        self.tgt_gts = tgt_data[chr_name]["GT"]
        self.windows = split_genome(pos=self.pos, chr_name=chr_name, polymorphism_size=win_len,
                                    step_size=win_step)

    def chromosome(self) -> dict:
        """The batched view the GPU scheduler consumes: whole matrices + window coordinates."""
        starts = np.asarray([w[1][0] for w in self.windows], dtype=np.int64)
        ends = np.asarray([w[1][1] for w in self.windows], dtype=np.int64)
        return {"chr_name": self.chr_name, "ploidy": self.ploidy, "is_phased": self.is_phased,
                "ref_gts": self.ref_gts, "tgt_gts": self.tgt_gts, "pos": self.pos,
                "win_start": starts, "win_end": ends}

    def get(self):
        pos = self.pos
        for chr_name, (start, end) in self.windows:
            idx = (pos >= start) & (pos <= end)
            yield {"chr_name": chr_name, "start": start, "end": end, "ploidy": self.ploidy,
                   "is_phased": self.is_phased, "ref_gts": self.ref_gts[idx],
                   "tgt_gts": self.tgt_gts[idx], "pos": pos[idx]}

    def __len__(self):
        return len(self.windows)
