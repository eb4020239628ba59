"""GPU tests of the drop-in seams S1 / S2 and the replicate-batched path, against the reference's
golden TSVs (tests/test_preprocess.py:25-88, tests/test_infer.py:26-60) and the oracle."""
import os

import numpy as np
import pandas as pd
import pytest

from conftest import FIXTURES
from oracle import c_oracle

pytestmark = pytest.mark.gpu
FX = FIXTURES


@pytest.fixture(scope="module", autouse=True)
def _built():
    import __graft_entry__ as g
    g.build()


@pytest.mark.parametrize("phased,step,expected", [
    (False, 10000, "test.sstar.unphased.expected.tsv"),
    (True, 10000, "test.sstar.phased.expected.tsv"),
    (True, 50000, "test.infer.expected.features.tsv"),
])
def test_preprocess_golden_tsv(tmp_path, phased, step, expected):
    from sstar_b200.preprocess import preprocess
    out = tmp_path / "nested" / "check.tsv"
    preprocess(vcf_file=os.path.join(FX, "test.vcf"), chr_name="21",
               ref_ind_file=os.path.join(FX, "test.ref.ind.list"),
               tgt_ind_file=os.path.join(FX, "test.tgt.ind.list"), win_len=50000, win_step=step,
               match_bonus=5000, max_mismatch=5, mismatch_penalty=-10000, output_file=str(out),
               nprocess=1, is_phased=phased)
    # the reference's own assertion ...
    pd.testing.assert_frame_equal(pd.read_csv(out, sep="\t"), pd.read_csv(os.path.join(FX, expected), sep="\t"),
                                  check_dtype=False, check_exact=False, rtol=1e-6, atol=1e-8)
    # ... and the stronger one: identical bytes
    assert open(out).read() == open(os.path.join(FX, expected)).read()


def test_mp_manager_seam_returns_reference_rows():
    from sstar_b200.feature_vector_preprocessor import FeatureVectorPreprocessor
    from sstar_b200.generators import WindowDataGenerator
    from sstar_b200.mp_manager import mp_manager
    from sstar_b200.utils import parse_ind_file
    gen = WindowDataGenerator(os.path.join(FX, "test.vcf"), "21", os.path.join(FX, "test.ref.ind.list"),
                              os.path.join(FX, "test.tgt.ind.list"), 50000, 10000, is_phased=False)
    job = FeatureVectorPreprocessor(parse_ind_file(os.path.join(FX, "test.ref.ind.list")),
                                    parse_ind_file(os.path.join(FX, "test.tgt.ind.list")), 5000, 5, -10000)
    rows = mp_manager(job=job, data_generator=gen, nprocess=4)
    want = pd.read_csv(os.path.join(FX, "test.sstar.unphased.expected.tsv"), sep="\t")
    got = pd.DataFrame(rows)
    got["Chromosome"] = got["Chromosome"].astype(int)
    pd.testing.assert_frame_equal(got, want, check_dtype=False)
    # window-by-window through job.run (the reference's per-task contract) gives the same rows
    rows2 = []
    for params in gen.get():
        rows2.extend(job.run(**params))
    pd.testing.assert_frame_equal(pd.DataFrame(rows2).assign(Chromosome=lambda d: d["Chromosome"].astype(int)),
                                  want, check_dtype=False)


def test_sharded_over_available_gpus_equals_single():
    import torch
    from sstar_b200.sharding import score_chromosome
    from sstar_b200.synth import regular_windows, synth_chromosome
    ref, tgt, pos = synth_chromosome(2_000_000, 10, 12, True, seed=21)
    ws, we = regular_windows(pos, 50_000, 10_000)
    so, no = c_oracle.score_windows(ref, tgt, pos, ws, we)
    n = torch.cuda.device_count()
    for devs in ([0], list(range(n)), [0] * 3):  # the last: 3 shards queued on one device
        if devs == [0] * 3:
            from sstar_b200.sharding import plan_window_shards
            from sstar_b200.engine import get_engine
            eng = get_engine(0)
            parts = [eng.score_host(ref[k0:k1], tgt[k0:k1], pos[k0:k1], ws[w0:w1], we[w0:w1])
                     for (w0, w1, k0, k1) in plan_window_shards(pos, ws, we, 3)]
            s = np.concatenate([p[0] for p in parts]); c = np.concatenate([p[1] for p in parts])
        else:
            s, c = score_chromosome(ref, tgt, pos, ws, we, 5000, 5, -10000, devices=devs)
        assert np.array_equal(s, so) and np.array_equal(c, no), devs


@pytest.mark.parametrize("phased", [True, False])
def test_replicate_batch_matches_per_replicate_oracle(phased):
    from sstar_b200.simulate_batch import build_feature_input, score_replicate_batch
    rng = np.random.default_rng(5)
    nref, ntgt, ploidy, L = 5, 10, 2, 50_000
    reps = []
    for r in range(40):
        nsite = int(rng.integers(0 if r else 3, 260)) if r % 9 else 3
        pos = np.sort(rng.choice(L + 1, size=max(nsite, 1), replace=False))
        freq = rng.random(len(pos))[:, None] ** 3
        gts = (rng.random((len(pos), (nref + ntgt) * ploidy)) < freq).astype(np.int8)
        gts[:, : nref * ploidy] &= (rng.random((len(pos), 1)) < 0.5)
        reps.append((pos, gts))
    table = score_replicate_batch(reps, list(range(100, 140)), nref, ntgt, ploidy, L, phased, 5000, 5, -10000)
    assert table.score.shape == (40, ntgt * (ploidy if phased else 1))
    for i, (pos, gts) in enumerate(reps):
        r, t, p, s, e = build_feature_input(pos, gts, nref, ntgt, ploidy, L, phased)
        so, no = c_oracle.score_windows(r, t, p, [s], [e])
        assert np.array_equal(table.score[i], so[0]) and np.array_equal(table.nsnp[i], no[0]), i
    rows = table.rows()
    assert rows[0]["Replicate"] == 100 and rows[-1]["Replicate"] == 139 and rows[0]["Sample"].startswith("tsk_5")


# ---- result emitter: device-formatted TSV == the Python emitter == pandas
def _table(W, T, seed, replicate, long_labels=False):
    from sstar_b200.feature_table import FeatureTable
    rng = np.random.default_rng(seed)
    score = rng.integers(-10_000, 600_000, size=(W, T)).astype(np.int64)
    score[rng.random((W, T)) < 0.25] = np.iinfo(np.int64).min          # NaN
    score[rng.random((W, T)) < 0.05] = rng.integers(-9, 10, size=1)[0]  # tiny values, zero, negatives
    if W and T:
        score[0, 0] = 9_999_999_999_999_998                             # longest score the device emitter prints
        score[-1, -1] = -9_999_999_999_999_998
    nsnp = rng.integers(0, 3000, size=(W, T)).astype(np.int32)
    ws = np.sort(rng.integers(-60_000, 250_000_000, size=W)).astype(np.int64)
    labels = [f"{'NA' if t % 3 else 'sample_with_a_long_name_'}{t}_{t % 2 + 1}" if long_labels else f"tsk_{t}"
              for t in range(T)]
    rep = list(rng.integers(0, 100_000, size=W)) if replicate else None
    return FeatureTable("chr21" if long_labels else 1, ws, ws + 49_999, labels, score, nsnp, replicate=rep)


@pytest.mark.parametrize("W,T,replicate,long_labels", [(1, 1, False, False), (37, 50, False, True),
                                                       (5, 700, True, False), (300, 257, True, True)])
def test_device_emitter_matches_python_and_pandas(tmp_path, W, T, replicate, long_labels):
    table = _table(W, T, W * 1000 + T, replicate, long_labels)
    order = None if not replicate else list(np.argsort(np.array(table.labels, dtype=object), kind="stable"))
    a, b = tmp_path / "dev.tsv", tmp_path / "py.tsv"
    table.to_tsv(a, device=0, col_order=order)
    table.to_tsv(b, col_order=order)
    got = open(a, "rb").read()
    assert got == open(b, "rb").read()
    if order is None:  # pandas prints rows in table order
        assert got.decode() == pd.DataFrame(table.rows()).to_csv(sep="\t", index=False, na_rep="NA")


def test_device_emitter_defers_huge_scores_to_the_host(tmp_path):
    table = _table(3, 4, 5, False)
    table.score[1, 1] = 10 ** 16   # pandas prints 1e+16 here
    table.to_tsv(tmp_path / "x.tsv", device=0)
    assert open(tmp_path / "x.tsv").read() == pd.DataFrame(table.rows()).to_csv(sep="\t", index=False, na_rep="NA")
