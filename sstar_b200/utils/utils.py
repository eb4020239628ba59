"""Host-side ingest and tiling helpers with the reference's names and error behaviour.

Mirror of ``sstar/utils/utils.py`` (reference paths in brackets).  scikit-allel is not a
dependency here: the VCF is parsed once by the native scanner behind :func:`_scan_vcf`
(``csrc/vcf_scan.cpp``) and both populations are cut from that single pass (the reference reads the file twice, utils.py:267,273).

What is reproduced (and pinned by the reference's golden TSVs):
  * samples come back in VCF header order, labels come from the ind-file order  [utils.py:80,86]
  * a call is missing when any allele is < 0; rows where a whole population is
    missing are dropped per population, then positions are intersected          [utils.py:105-107,203-214]
  * duplicate positions are an error                                             [utils.py:140-164]
  * sites hom-alt in every reference AND every target individual are dropped     [utils.py:282-291]
  * phased: [V, N, ploidy] -> [V, N*ploidy]; unphased: sum over the ploidy axis  [utils.py:293-306]
  * windows: 1-based closed intervals                                            [utils.py:479-488]
"""
from __future__ import annotations

from typing import Any

import numpy as np


def parse_ind_file(filename: str) -> list[str]:
    """One sample name per line [utils.py:25-46]."""
    with open(filename, "r") as fh:
        samples = [line.rstrip() for line in fh]
    if len(samples) == 0:
        raise Exception(f"No sample is found in {filename}! Please check your data.")
    return samples


def _scan_vcf(path: str, chr_name: str | None, need_alleles: bool = True, samples=None):
    """One streaming pass over a text VCF (plain, gzip or BGZF) -> kept samples (header order) +
    per-chromosome columns, through the native multi-threaded scanner (``sstar_b200_vcf_scan_ex``,
    csrc/vcf_scan.cpp).  ``samples``: names to keep (the ref + tgt lists; what
    ``allel.read_vcf(samples=...)`` does, utils.py:80) -- None keeps every column.

    ``REF`` / ``ALT`` are only materialised as Python strings when ``need_alleles`` (they are used
    by the ancestral-allele polarisation alone)."""
    import ctypes
    from .. import _cabi
    lib = _cabi.load()
    with open(str(path), "rb"):  # same FileNotFoundError / PermissionError as the reference's reader
        pass
    handle = ctypes.c_void_p()
    names, n_names = None, 0
    if samples is not None:
        uniq = sorted({str(x) for x in samples})
        names, n_names = (ctypes.c_char_p * max(len(uniq), 1))(*[u.encode() for u in uniq]), len(uniq)
    rc = lib.sstar_b200_vcf_scan_ex(str(path).encode(), None if chr_name is None else str(chr_name).encode(), names,
                                    n_names, 0, ctypes.byref(handle))
    if rc != 0:
        raise ValueError(lib.sstar_b200_vcf_last_error().decode("utf-8", "replace"))
    owner = _ScanOwner(lib, handle)  # POS / GT below are VIEWS of the scanner's arrays; it is freed with the last of them

    def view(addr, ctype, count, dtype):
        raw = (ctype * count).from_address(addr)
        raw._owner = owner
        return np.frombuffer(raw, dtype=dtype)

    n = lib.sstar_b200_vcf_n_samples(handle)
    header = [lib.sstar_b200_vcf_sample(handle, i).decode() for i in range(n)]
    out = {}
    for c in range(lib.sstar_b200_vcf_n_chroms(handle)):
        name = lib.sstar_b200_vcf_chrom(handle, c).decode()
        v = lib.sstar_b200_vcf_n_variants(handle, c)
        pos = view(lib.sstar_b200_vcf_pos(handle, c), ctypes.c_int32, v, np.int32) if v else np.zeros(0, dtype=np.int32)
        gt = view(lib.sstar_b200_vcf_gt(handle, c), ctypes.c_int8, v * n * 2, np.int8).reshape(v, n, 2) \
            if v and n else np.zeros((v, n, 2), dtype=np.int8)
        rec = {"POS": pos, "REF": None, "ALT": None, "GT": gt}
        if need_alleles:
            for key, which in (("REF", 0), ("ALT", 1)):
                off_p = ctypes.c_void_p()
                base = lib.sstar_b200_vcf_alleles(handle, c, which, ctypes.byref(off_p))
                off = np.ctypeslib.as_array(ctypes.cast(off_p, ctypes.POINTER(ctypes.c_int32)), shape=(v + 1,))
                raw = ctypes.string_at(base, int(off[v])) if v else b""
                rec[key] = np.asarray([raw[off[i]:off[i + 1]].decode() for i in range(v)], dtype=object)
        out[name] = rec
    return header, out


class _ScanOwner:
    """Keeps a native scan alive while numpy views of its arrays exist (no copy of the calls)."""

    def __init__(self, lib, handle):
        self._lib, self._handle = lib, handle

    def __del__(self):
        try:
            self._lib.sstar_b200_vcf_free(self._handle)
        except Exception:
            pass


def filter_data(data: dict, c: str, index: np.ndarray) -> dict:
    """Keep the rows selected by ``index`` on chromosome ``c`` [utils.py:114-137]."""
    index = np.asarray(index)
    if index.dtype == bool and index.all():  # nothing to drop (the usual case): the arrays stay as they are
        return data
    for key in ("POS", "REF", "ALT", "GT"):
        if data[c][key] is not None:  # REF / ALT are only loaded for the ancestral-allele step
            data[c][key] = data[c][key][index]
    return data


def validate_unique_variant_positions(data: dict, c: str) -> None:
    """[utils.py:140-164]"""
    pos = data[c]["POS"]
    uniq, counts = np.unique(pos, return_counts=True)
    if len(uniq) != len(pos):
        text = ", ".join(str(p) for p in uniq[counts > 1])
        raise ValueError(f"Duplicate variant positions found on chromosome {c}: {text}")


def _take_columns(gt: np.ndarray, cols) -> np.ndarray:
    """gt[:, cols, :]: a view when the columns are one run (a population listed as the header has it), else one
    gather of whole calls (the two alleles of a diploid call move as one 16-bit element)."""
    cols = list(cols)
    if cols and cols == list(range(cols[0], cols[0] + len(cols))):
        return gt[:, cols[0]:cols[0] + len(cols), :]
    if gt.shape[2] == 2 and gt.dtype == np.int8 and gt.flags.c_contiguous and cols:
        calls = gt.view(np.int16).reshape(gt.shape[0], gt.shape[1])
        return np.take(calls, cols, axis=1).view(np.int8).reshape(gt.shape[0], len(cols), 2)
    return gt[:, cols, :]


def _select_population(header, scanned, ind, filter_missing, anc_allele, chr_name):
    want = set(ind)
    cols = [i for i, s in enumerate(header) if s in want]
    samples = np.asarray([header[i] for i in cols])
    if not scanned:
        raise ValueError(f"{chr_name} is not present in the VCF file.")
    data: dict[str, dict[str, Any]] = {}
    for c, rec in scanned.items():
        data[c] = {"POS": rec["POS"], "REF": rec["REF"], "ALT": rec["ALT"], "GT": _take_columns(rec["GT"], cols)}
        validate_unique_variant_positions(data, c)
        if filter_missing:
            missing_calls = _any_allele_missing(data[c]["GT"]).sum(axis=1)
            data = filter_data(data, c, ~(missing_calls == len(samples)))
        if anc_allele is not None:
            data = check_anc_allele(data, anc_allele, c)
    return data, samples


def read_geno_data(vcf: str, ind: list[str], anc_allele_file: str, filter_missing: bool,
                   chr_name: str = None) -> tuple[dict, np.ndarray]:
    """Genotypes of the listed samples [utils.py:49-111]."""
    header, scanned = _scan_vcf(vcf, chr_name, need_alleles=anc_allele_file is not None, samples=ind)
    anc = read_anc_allele(anc_allele_file) if anc_allele_file is not None else None
    return _select_population(header, scanned, ind, filter_missing, anc, chr_name)


def align_population_data_by_position(ref_data: dict, tgt_data: dict) -> tuple[dict, dict]:
    """Keep shared positions only [utils.py:167-222]."""
    for c in tgt_data.keys():
        if c not in ref_data:
            raise ValueError(f"Chromosome {c} is missing from the reference data.")
        common = np.intersect1d(ref_data[c]["POS"], tgt_data[c]["POS"])
        if len(common) == 0:
            raise ValueError("No shared variant positions between reference and target "
                             f"on chromosome {c} after filtering.")
        ref_data = filter_data(ref_data, c, np.isin(ref_data[c]["POS"], common))
        tgt_data = filter_data(tgt_data, c, np.isin(tgt_data[c]["POS"], common))
        if not np.array_equal(ref_data[c]["POS"], tgt_data[c]["POS"]):
            raise ValueError("Reference and target variants are not aligned "
                             f"on chromosome {c} after filtering.")
    return ref_data, tgt_data


# The three per-call tests below walk the ploidy axis (length 2) with slices: a numpy reduction over an inner
# axis that short costs ten times the element-wise form (C2: 219 of the route's 390 ms were such reductions).
def _any_allele_missing(gt: np.ndarray) -> np.ndarray:
    """[V, N] bool: a call with a missing allele (allel's is_missing, utils.py:106)."""
    out = gt[:, :, 0] < 0
    for k in range(1, gt.shape[2]):
        out |= gt[:, :, k] < 0
    return out


def _is_hom_alt(gt: np.ndarray) -> np.ndarray:
    first = gt[:, :, 0]
    out = first > 0
    for k in range(1, gt.shape[2]):
        out &= gt[:, :, k] == first
    return out


def _dosage(gt: np.ndarray) -> np.ndarray:
    """[V, N] int8: allele sum of every call (np.sum over the ploidy axis, utils.py:305-306)."""
    acc = gt[:, :, 0].astype(np.int16)
    for k in range(1, gt.shape[2]):
        acc += gt[:, :, k]
    return acc.astype(np.int8)


def read_data(vcf_file: str, ref_ind_file: str, tgt_ind_file: str, anc_allele_file: str,
              is_phased: bool, chr_name: str = None) -> tuple[dict, list, dict, list]:
    """Reference + target genotype tables ready for tiling [utils.py:225-308].

    Genotype matrices come back as int8 (the reference returns int8 when phased and int64
    dosages when unphased; the values are identical).
    """
    ref_data = ref_samples = tgt_data = tgt_samples = None
    header = scanned = None
    anc = read_anc_allele(anc_allele_file) if anc_allele_file is not None else None
    if ref_ind_file is not None or tgt_ind_file is not None:
        listed = [x for f in (ref_ind_file, tgt_ind_file) if f is not None for x in parse_ind_file(f)]
        header, scanned = _scan_vcf(vcf_file, chr_name, need_alleles=anc is not None, samples=listed)

    def copy_scan():
        return {c: dict(rec) for c, rec in scanned.items()}

    if ref_ind_file is not None:
        ref_samples = parse_ind_file(ref_ind_file)
        ref_data, _ = _select_population(header, copy_scan(), ref_samples, True, anc, chr_name)
    if tgt_ind_file is not None:
        tgt_samples = parse_ind_file(tgt_ind_file)
        tgt_data, _ = _select_population(header, copy_scan(), tgt_samples, True, anc, chr_name)

    chr_names: list[str] = []
    if ref_ind_file is not None and tgt_ind_file is not None:
        ref_data, tgt_data = align_population_data_by_position(ref_data, tgt_data)
        chr_names = list(tgt_data.keys())
        for c in chr_names:
            ref_fixed = _is_hom_alt(ref_data[c]["GT"]).sum(axis=1) == len(ref_samples)
            tgt_fixed = _is_hom_alt(tgt_data[c]["GT"]).sum(axis=1) == len(tgt_samples)
            keep = ~(ref_fixed & tgt_fixed)
            ref_data = filter_data(ref_data, c, keep)
            tgt_data = filter_data(tgt_data, c, keep)

    for c in chr_names:
        for data in (ref_data, tgt_data):
            gt = data[c]["GT"]
            v, n, ploidy = gt.shape
            if is_phased:
                data[c]["GT"] = np.ascontiguousarray(gt.reshape(v, n * ploidy))
            else:
                data[c]["GT"] = _dosage(gt)
    return ref_data, ref_samples, tgt_data, tgt_samples


def read_anc_allele(anc_allele_file: str) -> dict:
    """BED-like ancestral allele table: chrom, start, end(1-based pos), allele [utils.py:338-360]."""
    table: dict[str, dict[int, str]] = {}
    with open(anc_allele_file, "r") as fh:
        for line in fh:
            e = line.rstrip().split()
            if not e:
                continue
            table.setdefault(e[0], {})[int(e[2])] = e[3]
    if not table:
        raise Exception("No ancestral allele is found! Please check your data.")
    return table


def check_anc_allele(data: dict, anc_allele: dict, c: str) -> dict:
    """Polarise genotypes by the ancestral allele [utils.py:363-423]: keep sites whose ancestral
    allele is REF, flip (abs(g - 1)) those where it is ALT, drop the rest and unknown sites."""
    pos = data[c]["POS"]
    anc_c = anc_allele[c]
    known = np.fromiter((int(p) in anc_c for p in pos), dtype=bool, count=len(pos))
    is_ref = np.fromiter((anc_c.get(int(p)) == r for p, r in zip(pos, data[c]["REF"])), dtype=bool,
                         count=len(pos))
    is_alt = np.fromiter((anc_c.get(int(p)) == a for p, a in zip(pos, data[c]["ALT"])), dtype=bool,
                         count=len(pos))
    keep = known & (is_ref | is_alt)
    flip = is_alt[keep]
    data = filter_data(data, c, keep)
    gt = np.array(data[c]["GT"], copy=True)
    gt[flip] = np.abs(gt[flip] - 1)
    data[c]["GT"] = gt
    return data


def split_genome(pos: np.ndarray, chr_name: str, polymorphism_size: int, step_size: int = None,
                 window_based: bool = True, random_polymorphisms: bool = False,
                 seed: int = None) -> list[tuple]:
    """Window list ``[(chr_name, [start, end]), ...]`` [utils.py:426-520].

    Window mode: ``start0 = max(1, (pos[0] + step) // step * step - size + 1)``, windows of
    ``size`` bp every ``step`` bp while ``pos[-1] >= start``; empty windows are emitted too.
    """
    if (step_size is not None and step_size <= 0) or polymorphism_size <= 0:
        raise ValueError("`step_size` and `polymorphism_size` must be positive integers.")
    if len(pos) == 0:
        raise ValueError("`pos` array must not be empty.")
    out: list[tuple] = []
    if window_based:
        first, last = int(pos[0]), int(pos[-1])
        start = max(1, (first + step_size) // step_size * step_size - polymorphism_size + 1)
        while last >= start:
            out.append((chr_name, [start, start + polymorphism_size - 1]))
            start += step_size
        return out
    n = len(pos)
    if random_polymorphisms:
        if seed is not None:
            np.random.seed(seed)
        if n < polymorphism_size:
            raise ValueError("No windows could be created with the given number of polymorphisms.")
        picked = np.random.choice(n, size=polymorphism_size, replace=False)
        out.append((chr_name, np.sort(picked).tolist()))
        return out
    i = 0
    while i + polymorphism_size <= n:
        out.append((chr_name, list(range(i, i + polymorphism_size))))
        i += step_size
    if not out:
        raise ValueError("No windows could be created with the given number of polymorphisms and step size.")
    if out[-1][1][1] != n - 1:
        out.append((chr_name, list(range(n - polymorphism_size, n))))
    return out
