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
    # the same shards on page-locked tensors, result blocks stored in place
    from sstar_b200.sharding import score_chromosome_pinned
    pin = [torch.from_numpy(np.ascontiguousarray(x)).pin_memory() for x in (ref, tgt, pos, ws, we)]
    sp = torch.zeros(so.shape, dtype=torch.int64).pin_memory()
    cp = torch.zeros(no.shape, dtype=torch.int32).pin_memory()
    for devs in ([0], list(range(n)), [0, 0, 0]):
        sp.zero_(); cp.zero_()
        score_chromosome_pinned(*pin, sp, cp, devices=devs)
        assert np.array_equal(sp.numpy(), so) and np.array_equal(cp.numpy(), no), devs


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


# ---- ingest filters on the device vs the numpy mirror of read_data
def _write_case_vcf(tmp_path, seed, n_ind=14, V=900, anc=False):
    """A VCF with every filter case: rows missing in a whole population, partially missing calls,
    sites fixed in both populations, multi-allelic calls; optionally an ancestral-allele table."""
    rng = np.random.default_rng(seed)
    pos = np.sort(rng.choice(np.arange(1, 400_000), size=V, replace=False))
    g = (rng.random((V, n_ind, 2)) < rng.random((V, 1, 1)) ** 2).astype(int)
    g[rng.random((V, n_ind, 2)) < 0.03] = -1                      # scattered missing alleles
    g[rng.random(V) < 0.05, : n_ind // 2] = -1                    # whole first half (ref pop) missing
    g[rng.random(V) < 0.05, n_ind // 2:] = -1                     # whole second half missing
    g[rng.random(V) < 0.06] = 1                                   # fixed alt everywhere
    g[rng.random((V, n_ind, 2)) < 0.004] = 2                      # second alt allele
    names = [f"s{i}" for i in range(n_ind)]
    bases = np.array(list("ACGT"))
    ref = bases[rng.integers(0, 4, V)]
    alt = bases[(np.searchsorted(bases, ref) + rng.integers(1, 4, V)) % 4]
    path = tmp_path / f"case{seed}.vcf"
    with open(path, "w") as fh:
        fh.write("##fileformat=VCFv4.2\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t" + "\t".join(names) + "\n")
        for k in range(V):
            calls = ["|".join("." if a < 0 else str(a) for a in g[k, i]) for i in range(n_ind)]
            fh.write(f"9\t{pos[k]}\t.\t{ref[k]}\t{alt[k]},G\t.\t.\t.\tGT\t" + "\t".join(calls) + "\n")
    # ind lists in an order different from the VCF header
    (tmp_path / "ref.list").write_text("\n".join(reversed(names[: n_ind // 2])) + "\n")
    (tmp_path / "tgt.list").write_text("\n".join(names[n_ind // 2:]) + "\n")
    anc_path = None
    if anc:
        anc_path = tmp_path / "anc.bed"
        with open(anc_path, "w") as fh:
            for k in range(V):
                u = rng.random()
                if u < 0.1:
                    continue                                      # unknown site -> dropped
                allele = ref[k] if u < 0.6 else alt[k] if u < 0.9 else "N"
                fh.write(f"9\t{pos[k] - 1}\t{pos[k]}\t{allele}\n")
    return str(path), str(tmp_path / "ref.list"), str(tmp_path / "tgt.list"), None if anc_path is None else str(anc_path)


@pytest.mark.parametrize("phased", [True, False])
@pytest.mark.parametrize("anc", [False, True])
def test_device_ingest_matches_read_data(tmp_path, phased, anc):
    from sstar_b200.engine import get_engine
    from sstar_b200.preprocess import _anc_modes
    from sstar_b200.utils import parse_ind_file, read_data
    from sstar_b200.utils.utils import _scan_vcf, read_anc_allele
    vcf, rl, tl, anc_file = _write_case_vcf(tmp_path, 40 + int(anc), anc=anc)
    ref_data, _, tgt_data, _ = read_data(vcf, rl, tl, anc_file, phased, chr_name="9")
    header, scanned = _scan_vcf(vcf, "9", need_alleles=anc)
    rec = scanned["9"]
    rcols = [i for i, s in enumerate(header) if s in set(parse_ind_file(rl))]
    tcols = [i for i, s in enumerate(header) if s in set(parse_ind_file(tl))]
    mode = _anc_modes(rec, read_anc_allele(anc_file)["9"]) if anc else None
    r, t, p, shared = get_engine(0).ingest_device(rec["GT"], rec["POS"], rcols, tcols, phased, mode)
    assert np.array_equal(p.cpu().numpy(), tgt_data["9"]["POS"]) and shared >= len(p) > 100
    assert np.array_equal(r.cpu().numpy(), ref_data["9"]["GT"])
    assert np.array_equal(t.cpu().numpy(), tgt_data["9"]["GT"])


@pytest.mark.parametrize("phased", [True, False])
def test_preprocess_device_route_equals_host_route(tmp_path, monkeypatch, phased):
    from sstar_b200.preprocess import preprocess
    vcf, rl, tl, anc_file = _write_case_vcf(tmp_path, 7, n_ind=20, V=1500, anc=True)
    kw = dict(vcf_file=vcf, chr_name="9", ref_ind_file=rl, tgt_ind_file=tl, win_len=50000, win_step=10000,
              match_bonus=5000, max_mismatch=5, mismatch_penalty=-10000, nprocess=2, is_phased=phased,
              anc_allele_file=anc_file)
    preprocess(output_file=str(tmp_path / "dev.tsv"), **kw)
    monkeypatch.setenv("SSTAR_B200_HOST_INGEST", "1")
    preprocess(output_file=str(tmp_path / "host.tsv"), **kw)
    a, b = open(tmp_path / "dev.tsv", "rb").read(), open(tmp_path / "host.tsv", "rb").read()
    assert a == b and a.count(b"\n") > 100


def test_preprocess_device_route_errors(tmp_path):
    from sstar_b200.preprocess import preprocess
    vcf, rl, tl, _ = _write_case_vcf(tmp_path, 3, n_ind=6, V=50)
    kw = dict(ref_ind_file=rl, tgt_ind_file=tl, win_len=50000, win_step=10000, match_bonus=5000, max_mismatch=5,
              mismatch_penalty=-10000, output_file=str(tmp_path / "o.tsv"))
    with pytest.raises(ValueError, match="is not present in the VCF file"):
        preprocess(vcf_file=vcf, chr_name="10", **kw)
    dup = tmp_path / "dup.vcf"
    lines = open(vcf).read().splitlines()
    dup.write_text("\n".join(lines + [lines[-1]]) + "\n")
    with pytest.raises(ValueError, match="Duplicate variant positions found on chromosome 9"):
        preprocess(vcf_file=str(dup), chr_name="9", **kw)
    (tmp_path / "none.list").write_text("nobody\n")
    with pytest.raises(ValueError, match="No shared variant positions"):
        preprocess(vcf_file=vcf, chr_name="9", **{**kw, "ref_ind_file": str(tmp_path / "none.list")})


# ---- simulate(): replicates scored in one device call, rows printed by
# the device emitter, against the reference's logic restated with row dicts + the oracle
@pytest.mark.parametrize("is_phased,is_shuffled", [(False, False), (True, True)])
def test_simulate_device_path_matches_reference_logic(tmp_path, monkeypatch, is_phased, is_shuffled):
    import test_simulate_host as host
    import sstar_b200.msprime_simulator as ms
    from sstar_b200.simulate import simulate
    monkeypatch.setattr(ms.MsprimeSimulator, "simulate_replicate",
                        lambda self, rep=None, seed=None: host._fake_replicate(rep, seed, self.nref, self.ntgt,
                                                                               self.ploidy, self.seq_len))
    kw = dict(demo_model_file="unused.yaml", nrep=9, nref=3, ntgt=12, ref_id="Ref", tgt_id="Tgt", ploidy=2,
              seq_len=50_000, mut_rate=1e-8, rec_rate=1e-8, match_bonus=5000, max_mismatch=5,
              mismatch_penalty=-10000, output_dir=str(tmp_path / "out"), output_prefix="t", nfeature=400,
              is_phased=is_phased, is_shuffled=is_shuffled, keep_sim_data=False, seed=77, nprocess=1)
    simulate(**kw)
    got = open(tmp_path / "out" / "t.training.features.tsv").read()

    class OracleSim(ms.MsprimeSimulator):   # scoring by the oracle, one replicate at a time
        def score_batch(self, replicates, reps):
            from sstar_b200.feature_table import FeatureTable, sample_labels
            from sstar_b200.simulate_batch import build_feature_input, sample_names
            sc, ns = [], []
            for pos, gts in replicates:
                r, t, p, s, e = build_feature_input(pos, gts, self.nref, self.ntgt, self.ploidy, self.seq_len, self.is_phased)
                so, no = c_oracle.score_windows(r, t, p, [s], [e])
                sc.append(so[0]); ns.append(no[0])
            labels = sample_labels(sample_names(self.nref, self.ntgt)[1], self.ploidy, self.is_phased)
            return FeatureTable("1", [1] * len(reps), [self.seq_len] * len(reps), labels, np.stack(sc), np.stack(ns),
                                replicate=list(reps))
    sim = OracleSim("unused.yaml", 3, 12, "Ref", "Tgt", 2, 50_000, 1e-8, 1e-8, "t", kw["output_dir"], is_phased,
                    5000, 5, -10000)
    assert got == host._reference_logic(sim, 9, 400, 77, is_shuffled)


def test_many_replicates_one_call():
    # thousands of replicates in one batch (BASELINE config 5 shape): positions restart in every
    # replicate, so windows of different replicates must never share a list
    from sstar_b200.engine import get_engine
    from sstar_b200.synth import synth_chromosome
    L, n_rep = 50_000, 3000
    ref, tgt, pos = synth_chromosome(L * n_rep, 5, 10, False, seed=9)
    rep_id = (pos.astype(np.int64) - 1) // L
    off = np.searchsorted(rep_id, np.arange(n_rep + 1)).astype(np.int32)
    pos_local = ((pos.astype(np.int64) - 1) % L + 1).astype(np.int32)
    s, n = get_engine(0).score_replicates_host(ref, tgt, pos_local, off, 1, L)
    assert s.shape == (n_rep, 10)
    for r in np.unique(np.linspace(0, n_rep - 1, 150).astype(int)):
        a, b = off[r], off[r + 1]
        so, no = c_oracle.score_windows(ref[a:b], tgt[a:b], pos_local[a:b], [1], [L])
        assert np.array_equal(s[r], so[0]) and np.array_equal(n[r], no[0]), r


# ---- sstar2 infer (N3): feature table -> predictions -> tracts
def test_infer_end_to_end_golden(tmp_path):
    # reference tests/test_infer.py:26-60 (paths of the config rewritten to the fixtures)
    import yaml
    from sstar_b200.infer import infer
    cfg = yaml.safe_load(open(os.path.join(FX, "test.config.yaml")))
    pre = cfg["preprocessing"]
    pre["vcf_file"] = os.path.join(FX, "test.vcf")
    pre["ref_ind_file"] = os.path.join(FX, "test.ref.ind.list")
    pre["tgt_ind_file"] = os.path.join(FX, "test.tgt.ind.list")
    (tmp_path / "cfg.yaml").write_text(yaml.safe_dump(cfg))
    feat, pred, tract = tmp_path / "f.tsv", tmp_path / "p.tsv", tmp_path / "t.bed"
    infer(model=os.path.join(FX, "test.qr.model.onnx"), config=str(tmp_path / "cfg.yaml"), feat_file=feat,
          pred_file=pred, tract_file=tract, match_bonus=5000, max_mismatch=5, mismatch_penalty=-10000)
    assert open(feat).read() == open(os.path.join(FX, "test.infer.expected.features.tsv")).read()
    pd.testing.assert_frame_equal(pd.read_csv(pred, sep="\t"),
                                  pd.read_csv(os.path.join(FX, "test.infer.expected.pred.tsv"), sep="\t"),
                                  check_dtype=False, check_exact=False, rtol=1e-6, atol=1e-8)
    assert open(tract).read() == open(os.path.join(FX, "test.infer.expected.inferred.tracts.bed")).read()


@pytest.mark.parametrize("W,T", [(1, 8), (300, 37), (19, 600)])
def test_device_prediction_tables_match_the_pandas_route(tmp_path, W, T):
    # prediction column, (Sample, Chromosome, Start, End) order and BED selection on the device ==
    # the file-based mirror of the reference's pandas code, byte for byte
    from sstar_b200 import quantile_regression as qr
    from sstar_b200.feature_table import FeatureTable
    rng = np.random.default_rng(W + T)
    score = rng.integers(20_000, 120_000, size=(W, T)).astype(np.int64)
    score[rng.random((W, T)) < 0.2] = np.iinfo(np.int64).min
    nsnp = rng.integers(0, 420, size=(W, T)).astype(np.int32)
    ws = (1 + 10_000 * np.arange(W)).astype(np.int64)
    labels = [f"ind{t // 2 + 1}_{t % 2 + 1}" for t in range(T)]   # ind10_1 sorts before ind2_1
    table = FeatureTable(21, ws, ws + 49_999, labels, score, nsnp)
    model = os.path.join(FX, "test.qr.model.onnx")
    table.to_tsv(tmp_path / "feat.tsv")
    qr.infer(str(tmp_path / "feat.tsv"), model, str(tmp_path / "host.pred.tsv"), str(tmp_path / "host.bed"))
    qr.infer_table_device(table, model, str(tmp_path / "dev.pred.tsv"), str(tmp_path / "dev.bed"))
    assert open(tmp_path / "dev.pred.tsv").read() == open(tmp_path / "host.pred.tsv").read()
    assert open(tmp_path / "dev.bed").read() == open(tmp_path / "host.bed").read()
    assert open(tmp_path / "dev.bed").read().count("\n") > 0 or W * T < 10


def test_reference_ingest_cases_through_the_device_route(tmp_path):
    # the reference's own read_data cases (tests/utils/test_utils.py:155-250) through preprocess()
    import test_host_logic as hl
    from sstar_b200.preprocess import preprocess
    kw = dict(win_len=1000, win_step=1000, match_bonus=5000, max_mismatch=5, mismatch_penalty=-10000)
    rows = ["1\t100\t.\tA\tG\t.\tPASS\t.\tGT\t0|0\t0|1\t0|1\t0|0", "1\t200\t.\tA\tG\t.\tPASS\t.\tGT\t.|.\t.|.\t0|1\t0|0",
            "1\t300\t.\tA\tG\t.\tPASS\t.\tGT\t0|1\t0|0\t.|.\t.|.", "1\t400\t.\tA\tG\t.\tPASS\t.\tGT\t0|1\t0|0\t0|1\t0|0",
            "1\t500\t.\tA\tG\t.\tPASS\t.\tGT\t1|1\t1|1\t1|1\t1|1"]
    vcf, r, t = hl._write(tmp_path, "m.vcf", ["ref1", "ref2", "tgt1", "tgt2"], rows, ["ref1", "ref2"], ["tgt1", "tgt2"])
    preprocess(vcf_file=vcf, chr_name="1", ref_ind_file=r, tgt_ind_file=t, output_file=str(tmp_path / "o.tsv"), **kw)
    got = pd.read_csv(tmp_path / "o.tsv", sep="\t")
    assert got["Region_ind_SNP_number"].tolist() == [2, 2, 2, 2] and got["S*_score"].isna().all()   # sites 100 and 400 survive
    vcf, r, t = hl._write(tmp_path, "n.vcf", ["ref1", "tgt1"], ["1\t100\t.\tA\tG\t.\tPASS\t.\tGT\t0|1\t.|.",
                                                                 "1\t200\t.\tA\tG\t.\tPASS\t.\tGT\t.|.\t0|1"], ["ref1"], ["tgt1"])
    with pytest.raises(ValueError, match="No shared variant positions.*chromosome 1"):
        preprocess(vcf_file=vcf, chr_name="1", ref_ind_file=r, tgt_ind_file=t, output_file=str(tmp_path / "o2.tsv"), **kw)


# ---- sstar2 match (N4): tract x source match rates
_MATCH_COLS = ["chrom", "start", "end", "sample", "match_rate"]


def _match_files(tmp_path, tract_text, src="src1\nsrc2\n"):
    # the reference's own mini input (tests/test_match.py:43-64)
    (tmp_path / "mini.vcf").write_text(
        "##fileformat=VCFv4.2\n##FORMAT=<ID=GT,Number=1,Type=String,Description=\"Genotype\">\n"
        "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\tind1\tsrc1\tsrc2\n"
        "21\t100\t.\tA\tG\t.\tPASS\t.\tGT\t0|0\t0|0\t1|1\n"
        "21\t200\t.\tA\tG\t.\tPASS\t.\tGT\t0|1\t1|1\t0|1\n"
        "21\t300\t.\tA\tG\t.\tPASS\t.\tGT\t1|1\t1|1\t.|.\n")
    (tmp_path / "tgt.ind.list").write_text("ind1\n")
    (tmp_path / "src.ind.list").write_text(src)
    (tmp_path / "tract.bed").write_text(tract_text)
    return dict(vcf_file=str(tmp_path / "mini.vcf"), tgt_ind_file=str(tmp_path / "tgt.ind.list"),
                src_ind_file=str(tmp_path / "src.ind.list"), tract_file=str(tmp_path / "tract.bed"),
                output_file=str(tmp_path / "nested" / "match.tsv"), ploidy=2)


def test_run_match_reference_cases(tmp_path):
    # tests/test_match.py:315-455
    from sstar_b200.match import run_match
    kw = _match_files(tmp_path, "21\t0\t250\tind1_1\n")
    run_match(**kw)
    df = pd.read_csv(kw["output_file"], sep="\t", header=None, names=_MATCH_COLS)
    assert df["sample"].tolist() == ["ind1_1"] and df["match_rate"].tolist() == pytest.approx([0.375])
    kw = _match_files(tmp_path, "21\t0\t250\tind1\n")
    run_match(**kw)
    assert pd.read_csv(kw["output_file"], sep="\t", header=None, names=_MATCH_COLS)["match_rate"].tolist() == \
        pytest.approx([0.625])
    kw = _match_files(tmp_path, "21\t0\t250\tind1_1\textra\n21\t0\t250\tind1_2\textra\nchr21\t400\t500\tind1_1\tx\n")
    run_match(**kw)
    df = pd.read_csv(kw["output_file"], sep="\t", header=None, names=_MATCH_COLS, keep_default_na=False)
    assert df["sample"].tolist() == ["ind1_1", "ind1_2", "ind1_1"] and df["chrom"].tolist() == ["21", "21", "chr21"]
    assert df["match_rate"].tolist() == ["0.375", "0.625", "NA"]
    with pytest.raises(ValueError, match="tract file must contain at least four columns"):
        run_match(**_match_files(tmp_path, "21\t0\t250\n"))
    with pytest.raises(ValueError, match="Target and source sample lists must not overlap"):
        run_match(**_match_files(tmp_path, "21\t0\t250\tind1_1\n", src="ind1\nsrc1\n"))


def test_run_match_random_against_oracle(tmp_path):
    from oracle import match_port
    from sstar_b200.match import run_match
    rng = np.random.default_rng(12)
    n_tgt, n_src, V = 6, 9, 700
    names = [f"t{i}" for i in range(n_tgt)] + [f"s{i}" for i in range(n_src)]
    order = rng.permutation(len(names))           # VCF column order differs from the list order
    pos = {"3": np.sort(rng.choice(np.arange(1, 200_000), V, replace=False)),
           "chr4": np.sort(rng.choice(np.arange(1, 200_000), V // 2, replace=False))}
    gts = {}
    with open(tmp_path / "r.vcf", "w") as fh:
        fh.write("##fileformat=VCFv4.2\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t" +
                 "\t".join(names[i] for i in order) + "\n")
        for c, p in pos.items():
            g = rng.choice([0, 1, 2, -1], size=(len(p), len(names), 2), p=[0.55, 0.35, 0.03, 0.07])
            gts[c] = g
            for k in range(len(p)):
                calls = ["|".join("." if a < 0 else str(a) for a in g[k, j]) for j in range(len(names))]
                fh.write(f"{c}\t{p[k]}\t.\tA\tG,T\t.\t.\t.\tGT\t" + "\t".join(calls) + "\n")
    (tmp_path / "t.list").write_text("\n".join(names[:n_tgt]) + "\n")
    (tmp_path / "s.list").write_text("\n".join(names[n_tgt:]) + "\n")
    tracts = []
    for _ in range(120):
        c = rng.choice(["3", "chr3", "4", "chr4"])
        a = int(rng.integers(0, 190_000))
        b = a + int(rng.integers(0, 30_000))
        lab = f"t{rng.integers(0, n_tgt)}" + rng.choice(["", "_1", "_2"])
        tracts.append((c, a, b, lab))
    (tmp_path / "tr.bed").write_text("".join(f"{c}\t{a}\t{b}\t{lab}\n" for c, a, b, lab in tracts))
    run_match(str(tmp_path / "r.vcf"), str(tmp_path / "t.list"), str(tmp_path / "s.list"), str(tmp_path / "tr.bed"),
              str(tmp_path / "out.tsv"), ploidy=2)
    got = pd.read_csv(tmp_path / "out.tsv", sep="\t", header=None, names=_MATCH_COLS, keep_default_na=False)
    col = {names[i]: j for j, i in enumerate(order)}
    want = []
    for c, a, b, lab in tracts:
        key = c if c in gts else (c[3:] if c.startswith("chr") else "chr" + c)
        base, _, suf = lab.partition("_")
        r = match_port.tract_rate(gts[key], pos[key], col[base], None if not suf else int(suf) - 1,
                                  [col[s] for s in names[n_tgt:]], a, b)
        want.append("NA" if r == "NA" else repr(r))
    assert got["match_rate"].tolist() == want          # bit-identical rates (ploidy 2)
    assert "NA" in want and len(set(want)) > 50


# ---- multi-GPU product routes.  Listing a device several times ("0,0,0") runs the same sharded code
# on a one-GPU box (shards queue on the device's worker thread); with more GPUs present the real
# devices are used as well.
def _device_lists():
    import torch
    n = torch.cuda.device_count()
    return ["0,0,0"] + ([",".join(str(i) for i in range(n))] if n > 1 else [])


@pytest.mark.parametrize("phased", [True, False])
def test_preprocess_sharded_device_route_prints_the_same_bytes(tmp_path, monkeypatch, phased):
    """preprocess() with several GPUs stays device-resident (ingest per raw row range, halo rows by
    peer copy, one text block per GPU) and prints what one GPU prints (and what the numpy-ingest
    route prints)."""
    from sstar_b200.preprocess import preprocess
    vcf, rl, tl, anc_file = _write_case_vcf(tmp_path, 11, n_ind=16, V=5000, anc=True)
    kw = dict(vcf_file=vcf, chr_name="9", ref_ind_file=rl, tgt_ind_file=tl, win_len=50000, win_step=10000,
              match_bonus=5000, max_mismatch=5, mismatch_penalty=-10000, is_phased=phased, anc_allele_file=anc_file)
    monkeypatch.setenv("SSTAR_B200_MIN_SHARD_ROWS", "256")
    monkeypatch.setenv("SSTAR_GPUS", "1")
    preprocess(output_file=str(tmp_path / "one.tsv"), **kw)
    one = open(tmp_path / "one.tsv", "rb").read()
    assert one.count(b"\n") > 300
    for devs in _device_lists():
        monkeypatch.setenv("SSTAR_GPUS", devs)
        preprocess(output_file=str(tmp_path / "many.tsv"), **kw)
        assert open(tmp_path / "many.tsv", "rb").read() == one, devs
    monkeypatch.setenv("SSTAR_B200_HOST_INGEST", "1")
    preprocess(output_file=str(tmp_path / "host.tsv"), **kw)
    assert open(tmp_path / "host.tsv", "rb").read() == one


def test_preprocess_chromosomes_one_scan_many_gpus(tmp_path, monkeypatch):
    """Several chromosomes of one VCF: scanned once, whole chromosomes assigned to GPUs by cost
    (or each sharded when there are fewer chromosomes than GPUs); same bytes as preprocess() per
    chromosome."""
    from sstar_b200.preprocess import preprocess, preprocess_chromosomes
    parts, names = [], None
    for i, (c, V) in enumerate((("3", 700), ("5", 1500), ("8", 400), ("11", 900))):
        vcf, rl, tl, _ = _write_case_vcf(tmp_path, 20 + i, n_ind=12, V=V)
        lines = open(vcf).read().splitlines()
        head = [l for l in lines if l.startswith("#")]
        parts += [c + l[1:] for l in lines if not l.startswith("#")]
    allvcf = tmp_path / "all.vcf"
    allvcf.write_text("\n".join(head + parts) + "\n")
    kw = dict(ref_ind_file=rl, tgt_ind_file=tl, win_len=50000, win_step=10000, match_bonus=5000, max_mismatch=5,
              mismatch_penalty=-10000, is_phased=False)
    chroms = ["3", "5", "8", "11"]
    monkeypatch.setenv("SSTAR_GPUS", "1")
    for c in chroms:
        preprocess(vcf_file=str(allvcf), chr_name=c, output_file=str(tmp_path / f"want.{c}.tsv"), **kw)
    monkeypatch.setenv("SSTAR_B200_MIN_SHARD_ROWS", "128")
    for devs in ["1", "0,0"] + _device_lists()[1:] + ["0,0,0,0,0,0"]:   # the last: fewer chromosomes than "GPUs"
        monkeypatch.setenv("SSTAR_GPUS", devs)
        out = preprocess_chromosomes(str(allvcf), chroms, output_files=str(tmp_path / "got.{chr}.tsv"), **kw)
        for c in chroms:
            assert open(out[c], "rb").read() == open(tmp_path / f"want.{c}.tsv", "rb").read(), (devs, c)
    with pytest.raises(ValueError, match="12 is not present in the VCF file"):
        preprocess_chromosomes(str(allvcf), ["3", "12"], output_files=str(tmp_path / "x.{chr}.tsv"), **kw)


def test_listed_sample_absent_from_vcf(tmp_path, monkeypatch):
    """An ind file naming a sample the VCF does not hold: the reference compares the hom-alt counts of
    its fixed-site filter with len(samples) of the LIST (utils.py:282-291), so no fixed site is ever
    dropped -- on the device route as on the numpy route.  A missing TARGET sample ends in
    SystemExit (the reference's worker fails while formatting: feature_vector_preprocessor.py:174-181)."""
    from sstar_b200.preprocess import preprocess
    vcf, rl, tl, _ = _write_case_vcf(tmp_path, 31, n_ind=12, V=1200)
    kw = dict(vcf_file=vcf, chr_name="9", tgt_ind_file=tl, win_len=50000, win_step=10000, match_bonus=5000,
              max_mismatch=5, mismatch_penalty=-10000, is_phased=False)
    rl2 = tmp_path / "ref_plus.list"
    rl2.write_text(open(rl).read() + "ghost\n")
    monkeypatch.setenv("SSTAR_GPUS", "1")
    preprocess(ref_ind_file=rl, output_file=str(tmp_path / "plain.tsv"), **kw)
    preprocess(ref_ind_file=str(rl2), output_file=str(tmp_path / "dev.tsv"), **kw)
    monkeypatch.setenv("SSTAR_B200_HOST_INGEST", "1")
    preprocess(ref_ind_file=str(rl2), output_file=str(tmp_path / "host.tsv"), **kw)
    monkeypatch.delenv("SSTAR_B200_HOST_INGEST")
    dev, host, plain = (open(tmp_path / f, "rb").read() for f in ("dev.tsv", "host.tsv", "plain.tsv"))
    assert dev == host and dev != plain   # the fixed sites stay in: Region_ind_SNP_number differs
    tl2 = tmp_path / "tgt_plus.list"
    tl2.write_text(open(tl).read() + "ghost\n")
    with pytest.raises(SystemExit):
        preprocess(ref_ind_file=rl, output_file=str(tmp_path / "x.tsv"), **{**kw, "tgt_ind_file": str(tl2)})


def test_replicate_batch_over_gpus_equals_one_gpu():
    """simulate()'s batches spread over SSTAR_GPUS as contiguous replicate ranges
    (sstar/simulate.py:147-176 fans them over worker processes)."""
    import torch
    import test_simulate_host as host
    from sstar_b200.simulate_batch import score_replicate_batch
    nref, ntgt, ploidy, L = 4, 9, 2, 50_000
    reps = [host._fake_replicate(r, 1000 + r, nref, ntgt, ploidy, L) for r in range(57)]
    ids = list(range(200, 257))
    for phased in (False, True):
        one = score_replicate_batch(reps, ids, nref, ntgt, ploidy, L, phased, 5000, 5, -10000, device=0)
        for devs in ([0, 0, 0], list(range(torch.cuda.device_count()))):
            many = score_replicate_batch(reps, ids, nref, ntgt, ploidy, L, phased, 5000, 5, -10000, devices=devs)
            assert np.array_equal(many.score, one.score) and np.array_equal(many.nsnp, one.nsnp), devs
            assert many.replicate == one.replicate
