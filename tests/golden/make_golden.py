"""Generate the golden fixtures in this directory FROM THE REFERENCE ITSELF.

Run once in the build container (the reference tree is not present on the GPU box):

    python tests/golden/make_golden.py [/root/reference]

It imports the reference's own ``sstar.sstar.Sstar`` and
``sstar.feature_vector_preprocessor.FeatureVectorPreprocessor`` (both need only
numpy), runs them on seeded inputs and freezes inputs + outputs:

* ``sstar_fuzz.npz``      -- 600 random single-window cases through ``Sstar.compute``
                             (sstar/sstar.py:43), incl. missing (<0) and multi-allelic
                             (>2) genotypes, sites closer than 10 bp, degenerate
                             parameters (max_mismatch=0, penalty=0, bonus=1 ...).
* ``test0_windows.npz``   -- every 50 kb / 10 kb window of tests/exp_results/test.0.vcf
                             (50 ref + 50 tgt diploids, msprime) through
                             ``FeatureVectorPreprocessor.run``, phased and unphased.
* ``ref_fixtures/``       -- verbatim copies of the reference's own TEST DATA (not
                             sources): the VCFs, ind lists and expected TSVs that pin
                             this path (tests/test_preprocess.py, tests/test_infer.py).

No reference source code is copied; only data and outputs.
"""
import os
import shutil
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
sys.path.insert(0, REF)
sys.path.insert(1, REPO)

from sstar.sstar import Sstar  # noqa: E402  (the reference)
from sstar.feature_vector_preprocessor import FeatureVectorPreprocessor  # noqa: E402

from oracle.sstar_port import float_scores_to_int  # noqa: E402
from oracle.tiling_port import load_populations, split_windows  # noqa: E402

PARAM_SETS = [(5000, 5, -10000), (5000, 5, -10000), (5000, 0, -10000), (1, 1, 0),
              (123, 4, -7), (0, 2, -1), (5000, 1, -10000), (70000, 3, -250000)]


def random_case(rng, idx):
    kind = idx % 6
    V = int(rng.integers(0, 70)) if idx % 11 else int(rng.integers(60, 260))
    R = int(rng.integers(1, 7))
    T = int(rng.integers(1, 9))
    if kind == 0:    # phased haplotypes
        tvals, tp = [0, 1], [0.6, 0.4]
    elif kind == 1:  # unphased dosages
        tvals, tp = [0, 1, 2], [0.55, 0.3, 0.15]
    elif kind == 2:  # unphased with missing calls (negative dosages)
        tvals, tp = [-2, -1, 0, 1, 2], [0.05, 0.1, 0.45, 0.25, 0.15]
    elif kind == 3:  # multi-allelic / higher ploidy
        tvals, tp = [0, 1, 2, 3, 4, 7], [0.4, 0.2, 0.15, 0.1, 0.1, 0.05]
    elif kind == 4:  # dense carriers
        tvals, tp = [0, 1, 2], [0.1, 0.5, 0.4]
    else:            # phased with missing
        tvals, tp = [-1, 0, 1], [0.1, 0.5, 0.4]
    tgt = rng.choice(tvals, size=(V, T), p=tp).astype(np.int8)
    # reference panel: mostly absent, sometimes cancelling signed sums (+1, -1)
    ref = np.zeros((V, R), dtype=np.int8)
    poly = rng.random(V) < 0.3
    ref[poly] = rng.choice([0, 1, 2], size=(int(poly.sum()), R), p=[0.5, 0.3, 0.2])
    if R >= 2 and V:
        cancel = rng.random(V) < 0.05
        ref[cancel, 0], ref[cancel, 1] = 1, -1
        ref[cancel, 2:] = 0
    # positions: strictly ascending, with clusters closer than 10 bp
    gaps = rng.choice([1, 3, 9, 10, 11, 200, 1500, 9000], size=V,
                      p=[0.05, 0.05, 0.08, 0.08, 0.08, 0.26, 0.25, 0.15])
    pos = (np.cumsum(gaps) + int(rng.integers(0, 5))).astype(np.int32) if V else np.zeros(0, np.int32)
    if V and idx % 7 == 0:
        pos = pos - pos[0]  # a site at position 0, as the reference's KATs have
    return ref, tgt, pos, PARAM_SETS[idx % len(PARAM_SETS)]


def make_fuzz(n_cases=600, seed=20261019):
    rng = np.random.default_rng(seed)
    meta, refs, tgts, poss, scores, nsnps = [], [], [], [], [], []
    for c in range(n_cases):
        ref, tgt, pos, (b, mx, p) = random_case(rng, c)
        out = Sstar.compute(ref_gts=ref, tgt_gts=tgt, pos=pos, match_bonus=b,
                            max_mismatch=mx, mismatch_penalty=p)
        meta.append([ref.shape[0], ref.shape[1], tgt.shape[1], b, mx, p])
        refs.append(ref.ravel()); tgts.append(tgt.ravel()); poss.append(pos)
        scores.append(float_scores_to_int(out["sstar"]))
        nsnps.append(np.asarray(out["region_ind_SNP_number"], dtype=np.int32))
    np.savez_compressed(
        os.path.join(HERE, "sstar_fuzz.npz"),
        meta=np.asarray(meta, dtype=np.int64), ref=np.concatenate(refs), tgt=np.concatenate(tgts),
        pos=np.concatenate(poss), score=np.concatenate(scores), nsnp=np.concatenate(nsnps))
    print("sstar_fuzz.npz:", n_cases, "cases,", sum(len(s) for s in scores), "units")


def copy_fixtures():
    dst = os.path.join(HERE, "ref_fixtures")
    os.makedirs(dst, exist_ok=True)
    for rel in ["tests/data/test.vcf", "tests/data/test.ref.ind.list", "tests/data/test.tgt.ind.list",
                "tests/exp_results/test.sstar.unphased.expected.tsv",
                "tests/exp_results/test.sstar.phased.expected.tsv",
                "tests/exp_results/test.infer.expected.features.tsv",
                "tests/exp_results/test.score.unphased.exp.results.tsv",
                "tests/exp_results/test.score.phased.exp.results.tsv",
                "tests/exp_results/test.0.vcf", "tests/exp_results/test.0.ref.ind.list",
                "tests/exp_results/test.0.tgt.ind.list"]:
        shutil.copyfile(os.path.join(REF, rel), os.path.join(dst, os.path.basename(rel)))
        os.chmod(os.path.join(dst, os.path.basename(rel)), 0o644)
    print("ref_fixtures/: copied")


def make_test0():
    fx = os.path.join(HERE, "ref_fixtures")
    out = {}
    for phased in (False, True):
        data, ref_names, tgt_names = load_populations(
            os.path.join(fx, "test.0.vcf"), os.path.join(fx, "test.0.ref.ind.list"),
            os.path.join(fx, "test.0.tgt.ind.list"), phased, "1")
        rg, tg, pos = data["1"]
        wins = split_windows(pos, 50000, 10000)
        job = FeatureVectorPreprocessor(ref_names, tgt_names, 5000, 5, -10000)
        sc, ns = [], []
        for (s, e) in wins:
            m = (pos >= s) & (pos <= e)
            rows = job.run(chr_name="1", start=s, end=e, ploidy=2, is_phased=phased,
                           ref_gts=rg[m], tgt_gts=tg[m], pos=pos[m])
            sc.append(float_scores_to_int([r["S*_score"] for r in rows]))
            ns.append([r["Region_ind_SNP_number"] for r in rows])
        tag = "phased" if phased else "unphased"
        out[f"{tag}_score"] = np.asarray(sc, dtype=np.int64)
        out[f"{tag}_nsnp"] = np.asarray(ns, dtype=np.int32)
        out[f"{tag}_win"] = np.asarray(wins, dtype=np.int64)
        print("test0", tag, out[f"{tag}_score"].shape)
    np.savez_compressed(os.path.join(HERE, "test0_windows.npz"), **out)


if __name__ == "__main__":
    copy_fixtures()
    make_fuzz()
    make_test0()
