"""CPU restatement of the tiling / ingest / emit steps around S* -- TEST INFRASTRUCTURE.

Follows (paths relative to /root/reference):

* ``split_genome`` window branch          sstar/utils/utils.py:426-488     -> :func:`split_windows`
* ``read_geno_data`` / ``read_data``      sstar/utils/utils.py:49-111,225-308 -> :func:`load_populations`
* ``align_population_data_by_position``   sstar/utils/utils.py:167-222     -> inside :func:`load_populations`
* ``FeatureVectorPreprocessor._fmt_res``  sstar/feature_vector_preprocessor.py:139-183 -> :func:`sample_labels`
* ``preprocess`` tail                     sstar/preprocess.py:114-122      -> :func:`rows_to_tsv`

scikit-allel 1.3.13 (un-vendored, absent here) is replaced by a plain-text VCF
reader restating the behaviours the reference relies on:
selected samples come back in VCF header order, ``GT`` is int8 ``[V, N, 2]`` with
-1 for a missing allele, a call is missing when any allele is < 0, hom-alt means
all alleles equal and > 0.  Behaviour for multi-allelic ALT / partially missing
calls / ind-list order != VCF order is NOT pinned by any reference test.
"""
from __future__ import annotations

import gzip
import numpy as np

from .sstar_port import NAN_SENTINEL


def split_windows(pos, win_len: int, win_step: int):
    """1-based closed windows: list of (start, end)."""
    if win_step is None or win_step <= 0 or win_len <= 0:
        raise ValueError("`step_size` and `polymorphism_size` must be positive integers.")
    if len(pos) == 0:
        raise ValueError("`pos` array must not be empty.")
    first, last = int(pos[0]), int(pos[-1])
    start = max(1, (first + win_step) // win_step * win_step - win_len + 1)
    out = []
    while last >= start:
        out.append((start, start + win_len - 1))
        start += win_step
    return out


def read_ind_list(path):
    with open(path) as fh:
        names = [ln.rstrip() for ln in fh.readlines()]
    if not names:
        raise Exception(f"No sample is found in {path}! Please check your data.")
    return names


def _parse_gt(field: str):
    gt = field.split(":", 1)[0]
    sep = "|" if "|" in gt else "/"
    parts = gt.split(sep)
    a = [-1, -1]
    for i, x in enumerate(parts[:2]):
        a[i] = -1 if x in (".", "") else int(x)
    return a


def read_vcf_text(path, samples, chrom=None):
    """Minimal text VCF reader -> dict(chrom -> dict(POS, REF, ALT, GT[V,N,2])), names."""
    opener = gzip.open if str(path).endswith(".gz") else open
    names, cols = None, None
    rows = {}
    with opener(path, "rt") as fh:
        for ln in fh:
            if ln.startswith("##"):
                continue
            f = ln.rstrip("\n").split("\t")
            if ln.startswith("#CHROM"):
                header = f[9:]
                want = set(samples)
                cols = [i for i, s in enumerate(header) if s in want]
                names = [header[i] for i in cols]
                continue
            if chrom is not None and f[0] != chrom:
                continue
            rec = rows.setdefault(f[0], {"POS": [], "REF": [], "ALT": [], "GT": []})
            rec["POS"].append(int(f[1]))
            rec["REF"].append(f[3])
            rec["ALT"].append(f[4].split(",")[0])
            rec["GT"].append([_parse_gt(f[9 + i]) for i in cols])
    out = {}
    for c, rec in rows.items():
        out[c] = {
            "POS": np.asarray(rec["POS"], dtype=np.int32),
            "REF": np.asarray(rec["REF"], dtype=object),
            "ALT": np.asarray(rec["ALT"], dtype=object),
            "GT": np.asarray(rec["GT"], dtype=np.int8).reshape(len(rec["POS"]), len(cols), 2),
        }
    return out, names


def _take(d, keep):
    return {k: v[keep] for k, v in d.items()}


def _one_population(path, ind_names, chrom):
    data, found = read_vcf_text(path, ind_names, chrom)
    if not data:
        raise ValueError(f"{chrom} is not present in the VCF file.")
    for c in list(data):
        p = data[c]["POS"]
        if len(np.unique(p)) != len(p):
            u, cnt = np.unique(p, return_counts=True)
            dup = ", ".join(str(x) for x in u[cnt > 1])
            raise ValueError(f"Duplicate variant positions found on chromosome {c}: {dup}")
        all_missing = (data[c]["GT"] < 0).any(axis=2).sum(axis=1) == len(found)
        data[c] = _take(data[c], ~all_missing)
    return data


def load_populations(vcf_file, ref_ind_file, tgt_ind_file, is_phased, chr_name=None):
    """-> (ref_gt[V,R], tgt_gt[V,T], pos[V]) per chromosome dict, ref_names, tgt_names."""
    ref_names = read_ind_list(ref_ind_file)
    tgt_names = read_ind_list(tgt_ind_file)
    ref = _one_population(vcf_file, ref_names, chr_name)
    tgt = _one_population(vcf_file, tgt_names, chr_name)
    out = {}
    for c in tgt:
        if c not in ref:
            raise ValueError(f"Chromosome {c} is missing from the reference data.")
        common = np.intersect1d(ref[c]["POS"], tgt[c]["POS"])
        if len(common) == 0:
            raise ValueError("No shared variant positions between reference and target "
                             f"on chromosome {c} after filtering.")
        r = _take(ref[c], np.isin(ref[c]["POS"], common))
        t = _take(tgt[c], np.isin(tgt[c]["POS"], common))
        if not np.array_equal(r["POS"], t["POS"]):
            raise ValueError("Reference and target variants are not aligned "
                             f"on chromosome {c} after filtering.")

        def hom_alt(gt):
            return (gt[:, :, 0] > 0) & (gt[:, :, 0] == gt[:, :, 1])

        fixed = (hom_alt(r["GT"]).sum(axis=1) == len(ref_names)) & \
                (hom_alt(t["GT"]).sum(axis=1) == len(tgt_names))
        r, t = _take(r, ~fixed), _take(t, ~fixed)
        if is_phased:
            rg = r["GT"].reshape(r["GT"].shape[0], -1)
            tg = t["GT"].reshape(t["GT"].shape[0], -1)
        else:
            rg = r["GT"].sum(axis=2)
            tg = t["GT"].sum(axis=2)
        out[c] = (rg, tg, t["POS"])
    return out, ref_names, tgt_names


def sample_labels(tgt_names, ploidy: int, is_phased: bool):
    if is_phased:
        return [f"{tgt_names[i // ploidy]}_{i % ploidy + 1}" for i in range(len(tgt_names) * ploidy)]
    return list(tgt_names)


def rows_to_tsv(path, chr_name, windows, labels, score, nsnp):
    """Emit exactly as the reference does: list of dicts -> pandas ``to_csv``."""
    import pandas as pd

    rows = []
    for w, (s, e) in enumerate(windows):
        for t, lab in enumerate(labels):
            v = score[w, t]
            rows.append({"Chromosome": chr_name, "Start": s, "End": e, "Sample": lab,
                         "S*_score": np.nan if v == NAN_SENTINEL else float(v),
                         "Region_ind_SNP_number": int(nsnp[w, t])})
    rows.sort(key=lambda x: (x["Chromosome"], x["Start"], x["End"]))
    pd.DataFrame(rows).to_csv(path, sep="\t", index=False, na_rep="NA")
