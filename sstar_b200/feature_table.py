"""Columnar result table + the byte-exact TSV emitter.

The reference builds one dict per (window, sample) row and lets pandas print them
(sstar/feature_vector_preprocessor.py:139-183, sstar/preprocess.py:117-122).  Here the results stay
as the two ``[W, T]`` arrays the GPU produced; rows / dicts are materialised only on request.

Text format reproduced from ``DataFrame.to_csv(sep="\\t", index=False, na_rep="NA")``:
float score column printed as ``51470.0`` (``repr`` of an integral float), NaN as ``NA``,
integer columns plain, ``\\n`` line ends, header ``Chromosome .. Region_ind_SNP_number``.
"""
from __future__ import annotations

import os

import numpy as np

from ._cabi import NAN_SENTINEL

COLUMNS = ["Chromosome", "Start", "End", "Sample", "S*_score", "Region_ind_SNP_number"]


def sample_labels(tgt_samples, ploidy: int, is_phased: bool) -> list[str]:
    """Row labels, sstar/feature_vector_preprocessor.py:162-178."""
    if is_phased:
        return [f"{tgt_samples[i // ploidy]}_{i % ploidy + 1}" for i in range(len(tgt_samples) * ploidy)]
    return [tgt_samples[i] for i in range(len(tgt_samples))]


class FeatureTable:
    def __init__(self, chr_name, win_start, win_end, labels, score, nsnp, replicate=None):
        self.chr_name = chr_name
        self.win_start = np.asarray(win_start)
        self.win_end = np.asarray(win_end)
        self.labels = list(labels)
        self.score = np.asarray(score)
        self.nsnp = np.asarray(nsnp)
        self.replicate = replicate
        W, T = self.score.shape
        if len(self.labels) != T or self.win_start.shape[0] != W or self.nsnp.shape != (W, T):
            raise ValueError("inconsistent feature table shapes")

    def __len__(self):
        return self.score.size

    def rows(self) -> list[dict]:
        """The reference's row dicts, in (window, sample) order."""
        out = []
        for w in range(self.score.shape[0]):
            s, e = int(self.win_start[w]), int(self.win_end[w])
            sc, ns = self.score[w].tolist(), self.nsnp[w].tolist()
            for t, lab in enumerate(self.labels):
                row = {"Chromosome": self.chr_name, "Start": s, "End": e, "Sample": lab,
                       "S*_score": np.nan if sc[t] == NAN_SENTINEL else float(sc[t]),
                       "Region_ind_SNP_number": int(ns[t])}
                if self.replicate is not None:
                    row["Replicate"] = self.replicate[w]
                out.append(row)
        return out

    def _int32_table(self) -> bool:
        i32 = np.iinfo(np.int32)
        ok = lambda a: a.size == 0 or (a.min() >= i32.min and a.max() <= i32.max)
        rep = np.asarray(self.replicate) if self.replicate is not None else np.zeros(0, dtype=np.int64)
        return (ok(self.win_start) and ok(self.win_end) and np.issubdtype(rep.dtype, np.integer) and ok(rep)
                and len(str(self.chr_name).encode()) < 64)

    def tsv_body(self, device=None, col_order=None) -> bytes:
        """The rows as TSV text (no header), byte-identical to the reference's
        ``pd.DataFrame(rows).to_csv(sep='\\t', na_rep='NA')``.

        ``device`` (a CUDA device index) formats the rows with the device emitter
        (``sstar_b200_format_tsv``); ``None`` formats them in Python.  ``col_order`` permutes the
        rows inside each window (the simulate path sorts rows by sample name)."""
        W, T = self.score.shape
        chrom = str(self.chr_name)
        order = list(range(T)) if col_order is None else [int(c) for c in col_order]
        if device is not None and W > 0 and self._int32_table():
            from .engine import get_engine
            body = get_engine(device).format_tsv_host(self.score, self.nsnp, self.win_start, self.win_end, chrom,
                                                      self.labels, self.replicate, col_order)
            if body is not None:
                return bytes(body)
        parts = []
        for w in range(W):
            head = f"{chrom}\t{int(self.win_start[w])}\t{int(self.win_end[w])}\t"
            tail = "\n" if self.replicate is None else f"\t{self.replicate[w]}\n"
            sc, ns = self.score[w].tolist(), self.nsnp[w].tolist()
            parts.extend(
                f"{head}{self.labels[t]}\tNA\t{ns[t]}{tail}" if sc[t] == NAN_SENTINEL
                else f"{head}{self.labels[t]}\t{float(sc[t])!r}\t{ns[t]}{tail}"
                for t in order)
        return "".join(parts).encode()

    def to_tsv(self, path, header: bool = True, mode: str = "w", device=None, col_order=None) -> None:
        """Write the table (see :meth:`tsv_body`), creating the output directory like the reference
        does (sstar/preprocess.py:119-121)."""
        out_dir = os.path.dirname(str(path))
        if out_dir:
            os.makedirs(out_dir, exist_ok=True)
        cols = COLUMNS + (["Replicate"] if self.replicate is not None else [])
        with open(path, mode + "b") as fh:
            if header:
                fh.write(("\t".join(cols) + "\n").encode())
            fh.write(self.tsv_body(device=device, col_order=col_order))
