"""Source match rates for inferred tracts (reference sstar/match.py, the ``sstar2 match`` step).

Same functions, arguments, errors and output file as the reference; the tract x source loops of
``run_match`` (:497-545) run on the GPU as integer numerators (``sstar_b200_match_numerators``)
and the host forms the rates exactly as the reference does: ``(N / P) / count`` per source,
``float(np.mean(...))`` over the sources.  For ``ploidy = 2`` every term of the reference's float
sum is a multiple of 1/2, so the rates are bit-identical; for a ploidy that is not a power of two
they agree to rounding.
"""
from __future__ import annotations

import os
from typing import Optional, Union

import numpy as np

from .utils import parse_ind_file
from .utils.utils import read_geno_data


def calc_match_pct(x, y, denominator: int, P: int = 2) -> Union[float, str]:
    """Window-normalised dosage match rate [match.py:30-73]; -1 encodes a missing dosage."""
    if denominator == 0:
        return "NA"
    x = np.atleast_1d(np.asarray(x, dtype=float))
    y = np.atleast_1d(np.asarray(y, dtype=float))
    ok = (x >= 0) & (y >= 0)
    return float(np.sum(1.0 - np.abs(x[ok] - y[ok]) / float(P)) / denominator)


def _base_sample_name(sample: str, samples: list) -> tuple:
    """Tract label -> (base sample, haplotype index or None) [match.py:262-296]."""
    if sample in samples:
        return sample, None
    base, sep, suffix = sample.rpartition("_")
    if sep and suffix in {"1", "2"} and base in samples:
        return base, int(suffix) - 1
    raise ValueError(f"Target sample {sample} is not present in the target list.")


def _resolve_chrom(chrom: object, data: dict) -> str:
    """BED chromosome label -> VCF chromosome key, with or without ``chr`` [match.py:299-337]."""
    chrom = str(chrom)
    if chrom in data:
        return chrom
    alt_chrom = chrom[3:] if chrom.startswith("chr") else f"chr{chrom}"
    if alt_chrom in data:
        return alt_chrom
    available = ", ".join(str(c) for c in data.keys())
    raise ValueError(f"Chromosome {chrom} is not present in the VCF data. Available chromosomes: {available}")


def _validate_disjoint_samples(tgt_samples: list, src_samples: list) -> None:
    overlap = set(tgt_samples) & set(src_samples)
    if overlap:
        raise ValueError(f"Target and source sample lists must not overlap: {', '.join(sorted(overlap))}")


def _validate_sorted_positions(data: dict) -> None:
    for chrom, chrom_data in data.items():
        pos = np.asarray(chrom_data["POS"])
        if np.any(pos[:-1] > pos[1:]):
            raise ValueError(f"VCF positions must be sorted within chromosome {chrom}.")


def _positions_in_bed_interval(pos: np.ndarray, start: int, end: int) -> np.ndarray:
    """VCF positions with ``start < POS <= end`` [match.py:408-427]."""
    left = np.searchsorted(pos, start, side="right")
    right = np.searchsorted(pos, end, side="right")
    return pos[left:right]


def run_match(vcf_file: str, tgt_ind_file: str, src_ind_file: str, tract_file: str, output_file: str,
              ploidy: int = 2, device: Optional[int] = None) -> None:
    """Aggregated source match rates for inferred tracts [match.py:430-565]."""
    import pandas as pd
    from .engine import get_engine
    from .sharding import visible_devices
    tgt_samples = parse_ind_file(tgt_ind_file)
    src_samples = parse_ind_file(src_ind_file)
    _validate_disjoint_samples(tgt_samples, src_samples)
    data, returned_samples = read_geno_data(vcf_file, tgt_samples + src_samples, None, filter_missing=False)
    sample_to_index = {str(s): i for i, s in enumerate(returned_samples)}
    src_indices = [sample_to_index[s] for s in src_samples]
    _validate_sorted_positions(data)

    tracts = pd.read_csv(tract_file, sep="\t", header=None)
    if tracts.shape[1] < 4:
        raise ValueError("The tract file must contain at least four columns: chrom, start, end, and sample.")
    tracts = tracts.iloc[:, :4]
    tracts.columns = ["chrom", "start", "end", "sample"]

    # resolve every tract on the host (labels, chromosomes, variant row ranges), group by chromosome
    K = len(tracts)
    chrom_of, lo, hi, tgt_col, hap = [None] * K, np.zeros(K, np.int32), np.zeros(K, np.int32), \
        np.zeros(K, np.int32), np.full(K, -1, np.int32)
    for k, tract in enumerate(tracts.itertuples(index=False)):
        base_sample, hap_index = _base_sample_name(str(tract.sample), tgt_samples)
        tgt_col[k] = sample_to_index[base_sample]
        hap[k] = -1 if hap_index is None else hap_index
        c = _resolve_chrom(tract.chrom, data)
        chrom_of[k] = c
        pos = np.asarray(data[c]["POS"])
        lo[k] = np.searchsorted(pos, int(tract.start), side="right")
        hi[k] = np.searchsorted(pos, int(tract.end), side="right")
    rates: list = ["NA"] * K
    eng = get_engine(visible_devices()[0] if device is None else device)
    for c in dict.fromkeys(chrom_of):
        sel = np.asarray([k for k in range(K) if chrom_of[k] == c and hi[k] > lo[k]], dtype=np.int64)
        if sel.size == 0:
            continue
        numer = eng.match_numerators(data[c]["GT"], lo[sel], hi[sel], tgt_col[sel], hap[sel], src_indices, ploidy)
        for j, k in enumerate(sel):
            count = int(hi[k] - lo[k])
            source_rates = [float((float(n) / float(ploidy)) / count) for n in numer[j]]
            rates[k] = float(np.mean(source_rates))

    rows = [{"chrom": t.chrom, "start": int(t.start), "end": int(t.end), "sample": str(t.sample),
             "match_rate": rates[k]} for k, t in enumerate(tracts.itertuples(index=False))]
    output_dir = os.path.dirname(output_file)
    if output_dir:
        os.makedirs(output_dir, exist_ok=True)
    pd.DataFrame(rows, columns=["chrom", "start", "end", "sample", "match_rate"]).to_csv(
        output_file, sep="\t", index=False, header=False)
