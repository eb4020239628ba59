"""CPU tests of the host-side mirror: ingest, tiling, row formatting, TSV emit, sharding plan.
(The scores fed to the emitter here come from the oracle -- no GPU involved.)"""
import os

import numpy as np
import pytest

from conftest import FIXTURES
from oracle import c_oracle, tiling_port
from oracle.sstar_port import NAN_SENTINEL

FX = FIXTURES


def _gen(phased, step=10000, vcf="test.vcf", chrom="21", pre="test"):
    from sstar_b200.generators import WindowDataGenerator
    return WindowDataGenerator(os.path.join(FX, vcf), chrom, os.path.join(FX, f"{pre}.ref.ind.list"),
                               os.path.join(FX, f"{pre}.tgt.ind.list"), 50000, step, is_phased=phased)


@pytest.mark.parametrize("phased", [False, True])
@pytest.mark.parametrize("vcf,chrom,pre", [("test.vcf", "21", "test"), ("test.0.vcf", "1", "test.0")])
def test_ingest_matches_oracle_port(phased, vcf, chrom, pre):
    g = _gen(phased, vcf=vcf, chrom=chrom, pre=pre)
    data, _, _ = tiling_port.load_populations(os.path.join(FX, vcf), os.path.join(FX, f"{pre}.ref.ind.list"),
                                              os.path.join(FX, f"{pre}.tgt.ind.list"), phased, chrom)
    rg, tg, pos = data[chrom]
    assert np.array_equal(g.ref_gts, rg) and np.array_equal(g.tgt_gts, tg) and np.array_equal(g.pos, pos)
    assert [tuple(w[1]) for w in g.windows] == tiling_port.split_windows(pos, 50000, 10000)


def test_generator_get_matches_manual_slicing():
    # reference tests/generators/test_window_data_generator.py:85-113
    g = _gen(True, step=50000, vcf="test.0.vcf", chrom="1", pre="test.0")
    got = list(g.get())
    assert len(got) == len(g)
    for d, (c, (s, e)) in zip(got, g.windows):
        idx = (g.pos >= s) & (g.pos <= e)
        assert d["chr_name"] == c and d["start"] == s and d["end"] == e
        assert np.array_equal(d["pos"], g.pos[idx]) and np.array_equal(d["tgt_gts"], g.tgt_gts[idx])
    pos = np.array([1, 10, 20, 21, 30])
    assert np.array_equal(pos[(pos >= 10) & (pos <= 20)], [10, 20])


def test_split_genome_reference_cases():
    from sstar_b200.utils import split_genome
    pos = np.array([261, 592, 619, 693, 1008, 1125, 1307, 1424, 1436, 1491, 1743, 1853, 1917, 2075])
    assert split_genome(pos=pos, chr_name="1", polymorphism_size=50000, step_size=10000) == [("1", [1, 50000])]
    assert split_genome(np.array([1, 2, 3, 4, 5]), "1", 3, 2)[0] == ("1", [1, 3])
    poly = split_genome(pos, "1", 10, 1, window_based=False)
    # (the reference's tail check compares element [1] of the last index list, so it appends a
    #  duplicate of the last window; its own test zips against 5 expected windows)
    assert [p[1][0] for p in poly[:5]] == [0, 1, 2, 3, 4] and poly[0][1] == list(range(10))
    assert split_genome(pos, "1", 10, window_based=False, random_polymorphisms=True, seed=12345) == \
        [("1", [0, 3, 6, 7, 8, 9, 10, 11, 12, 13])]
    for bad in (dict(step_size=-1), dict(step_size=0)):
        with pytest.raises(ValueError):
            split_genome(pos, "1", 50000, **bad)
    with pytest.raises(ValueError):
        split_genome(np.array([]), "1", 50000, 10000)


def test_read_data_errors(tmp_path):
    from sstar_b200.utils import read_data
    vcf = tmp_path / "dup.vcf"
    vcf.write_text("##fileformat=VCFv4.2\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\tr1\tt1\n"
                   "1\t100\t.\tA\tC\t.\t.\t.\tGT\t0|0\t0|1\n1\t100\t.\tA\tG\t.\t.\t.\tGT\t0|0\t1|1\n")
    (tmp_path / "r").write_text("r1\n")
    (tmp_path / "t").write_text("t1\n")
    with pytest.raises(ValueError, match="Duplicate variant positions found on chromosome 1: 100"):
        read_data(str(vcf), str(tmp_path / "r"), str(tmp_path / "t"), None, True, chr_name="1")
    with pytest.raises(ValueError, match="2 is not present in the VCF file"):
        read_data(os.path.join(FX, "test.vcf"), os.path.join(FX, "test.ref.ind.list"),
                  os.path.join(FX, "test.tgt.ind.list"), None, True, chr_name="2")
    # whole-population-missing rows are dropped per population, then positions are intersected
    vcf2 = tmp_path / "miss.vcf"
    vcf2.write_text("##fileformat=VCFv4.2\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\tr1\tt1\n"
                    "1\t100\t.\tA\tC\t.\t.\t.\tGT\t.|.\t0|1\n1\t200\t.\tA\tG\t.\t.\t.\tGT\t0|0\t1|1\n"
                    "1\t300\t.\tA\tG\t.\t.\t.\tGT\t0|1\t.|.\n1\t400\t.\tA\tG\t.\t.\t.\tGT\t1|1\t1|1\n")
    ref, _, tgt, _ = read_data(str(vcf2), str(tmp_path / "r"), str(tmp_path / "t"), None, True, chr_name="1")
    assert ref["1"]["POS"].tolist() == [200] and tgt["1"]["GT"].tolist() == [[1, 1]]


@pytest.mark.parametrize("phased,step,expected", [
    (False, 10000, "test.sstar.unphased.expected.tsv"),
    (True, 10000, "test.sstar.phased.expected.tsv"),
    (True, 50000, "test.infer.expected.features.tsv"),
])
def test_emitter_is_byte_identical_to_reference_tsv(tmp_path, phased, step, expected):
    from sstar_b200.feature_table import FeatureTable, sample_labels
    from sstar_b200.utils import parse_ind_file
    g = _gen(phased, step)
    c = g.chromosome()
    s, n = c_oracle.score_windows(c["ref_gts"], c["tgt_gts"], c["pos"], c["win_start"], c["win_end"])
    labels = sample_labels(parse_ind_file(os.path.join(FX, "test.tgt.ind.list")), 2, phased)
    table = FeatureTable("21", c["win_start"], c["win_end"], labels, s, n)
    out = tmp_path / "sub" / "o.tsv"
    table.to_tsv(out)
    assert open(out).read() == open(os.path.join(FX, expected)).read()
    rows = table.rows()
    assert len(rows) == len(table) and set(rows[0]) == {"Chromosome", "Start", "End", "Sample", "S*_score",
                                                         "Region_ind_SNP_number"}


def test_fmt_res_matches_reference_row_format(monkeypatch):
    # reference tests/test_feature_vector_preprocessor.py:25-70 (Sstar.compute is the seam)
    import sstar_b200.feature_vector_preprocessor as fvp
    job = fvp.FeatureVectorPreprocessor(["ref1"], ["tgt1"], 123, 4, -7)
    seen = {}

    def fake(**kw):
        seen.update(kw)
        return {"sstar": np.array([42.0]), "region_ind_SNP_number": np.array([1000])}

    monkeypatch.setattr("sstar_b200.feature_vector_preprocessor.Sstar.compute", fake)
    out = job.run(chr_name="chr1", start=10, end=20, ploidy=2, is_phased=False, ref_gts=np.array([[0], [1]]),
                  tgt_gts=np.array([[1], [1]]), pos=np.array([11, 19]))
    assert set(seen) == {"ref_gts", "tgt_gts", "pos", "match_bonus", "max_mismatch", "mismatch_penalty"}
    assert out == [{"Chromosome": "chr1", "Start": 10, "End": 20, "Sample": "tgt1", "S*_score": 42.0,
                    "Region_ind_SNP_number": 1000}]
    with pytest.raises(ValueError, match="No reference sample"):
        fvp.FeatureVectorPreprocessor([], ["a"], 1, 1, -1)
    with pytest.raises(ValueError, match="No target sample"):
        fvp.FeatureVectorPreprocessor(["a"], [], 1, 1, -1)


def test_scheduler_generic_jobs():
    # reference tests/test_mp_manager.py: SimpleJob echo and FailureJob -> "error"
    from sstar_b200.generators import RandomNumberGenerator
    from sstar_b200.mp_manager import mp_manager

    class Echo:
        def run(self, rep, seed):
            return [(rep, seed)]

    class Boom:
        def run(self, rep, seed):
            raise Exception("boom")

    gen = RandomNumberGenerator(nrep=5, seed=123)
    res = mp_manager(job=Echo(), data_generator=gen, nprocess=2)
    assert sorted(res) == sorted(zip(gen.rep_list, gen.seed_list))
    assert mp_manager(job=Boom(), data_generator=gen, nprocess=2) == "error"


def test_random_number_generator_seeds():
    # reference tests/generators/test_random_number_generator.py:27-38
    from sstar_b200.generators import RandomNumberGenerator
    np.random.seed(123)
    want = np.random.randint(1, 2 ** 31, 10).tolist()
    g = RandomNumberGenerator(nrep=4, start_rep=6, seed=123)
    assert g.seed_list == want[6:10] and g.rep_list == [6, 7, 8, 9] and len(g) == 4
    with pytest.raises(ValueError):
        RandomNumberGenerator(nrep=0)


def test_preprocess_argument_errors(tmp_path):
    from sstar_b200.preprocess import preprocess
    with pytest.raises(ValueError, match="Number of processes"):
        preprocess("x.vcf", "1", "r", "t", 50000, 10000, str(tmp_path / "o"), 5000, 5, -10000, nprocess=0)


def test_shard_plan_covers_every_window():
    from sstar_b200.sharding import plan_window_shards
    from sstar_b200.synth import regular_windows, synth_chromosome
    ref, tgt, pos = synth_chromosome(1_200_000, 6, 5, False, seed=3)
    ws, we = regular_windows(pos, 50_000, 10_000)
    for n in (1, 2, 3, 8, 500):
        shards = plan_window_shards(pos, ws, we, n)
        assert shards[0][0] == 0 and shards[-1][1] == len(ws)
        assert all(a[1] == b[0] for a, b in zip(shards, shards[1:]))
        so, no = c_oracle.score_windows(ref, tgt, pos, ws, we)
        for (w0, w1, k0, k1) in shards:  # a shard scored from ITS variant rows equals the full result
            s, c = c_oracle.score_windows(ref[k0:k1], tgt[k0:k1], pos[k0:k1], ws[w0:w1], we[w0:w1])
            assert np.array_equal(s, so[w0:w1]) and np.array_equal(c, no[w0:w1])


def test_simulate_tiler_matches_port():
    from sstar_b200.simulate_batch import build_feature_input, sample_names
    rng = np.random.default_rng(0)
    nref, ntgt, ploidy, L = 3, 4, 2, 5000
    pos = np.sort(rng.choice(L + 1, size=60, replace=False))
    pos[0] = 0  # msprime can place a site at position 0: it falls outside [1, seq_len]
    gts = (rng.random((60, (nref + ntgt) * ploidy)) < 0.3).astype(np.int8)
    gts[5] = 1  # fixed-alt in both populations: dropped
    for phased in (True, False):
        r, t, p, s, e = build_feature_input(pos, gts, nref, ntgt, ploidy, L, phased)
        assert (s, e) == (1, L) and 0 not in p and pos[5] not in p
        assert r.shape == (len(p), nref * (ploidy if phased else 1))
        assert t.shape == (len(p), ntgt * (ploidy if phased else 1))
    assert sample_names(2, 2) == (["tsk_0", "tsk_1"], ["tsk_2", "tsk_3"])


# ---- the reference's own ingest cases, verbatim inputs (tests/utils/test_utils.py:155-250,
# tests/generators/test_window_data_generator.py:116-150)
_HDR = "##fileformat=VCFv4.2\n##FORMAT=<ID=GT,Number=1,Type=String,Description=\"Genotype\">\n" \
       "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t"


def _write(tmp_path, name, samples, rows, ref, tgt):
    (tmp_path / name).write_text(_HDR + "\t".join(samples) + "\n" + "\n".join(rows) + "\n")
    (tmp_path / "ref.ind.list").write_text("\n".join(ref) + "\n")
    (tmp_path / "tgt.ind.list").write_text("\n".join(tgt) + "\n")
    return str(tmp_path / name), str(tmp_path / "ref.ind.list"), str(tmp_path / "tgt.ind.list")


def test_reference_case_alignment_after_missing_filtering(tmp_path):
    from sstar_b200.utils import read_data
    rows = ["1\t100\t.\tA\tG\t.\tPASS\t.\tGT\t0|0\t0|1\t0|1\t0|0", "1\t200\t.\tA\tG\t.\tPASS\t.\tGT\t.|.\t.|.\t0|1\t0|0",
            "1\t300\t.\tA\tG\t.\tPASS\t.\tGT\t0|1\t0|0\t.|.\t.|.", "1\t400\t.\tA\tG\t.\tPASS\t.\tGT\t0|1\t0|0\t0|1\t0|0",
            "1\t500\t.\tA\tG\t.\tPASS\t.\tGT\t1|1\t1|1\t1|1\t1|1"]
    vcf, r, t = _write(tmp_path, "m.vcf", ["ref1", "ref2", "tgt1", "tgt2"], rows, ["ref1", "ref2"], ["tgt1", "tgt2"])
    ref_data, _, tgt_data, _ = read_data(vcf_file=vcf, ref_ind_file=r, tgt_ind_file=t, anc_allele_file=None, is_phased=True)
    assert np.array_equal(ref_data["1"]["POS"], [100, 400]) and np.array_equal(tgt_data["1"]["POS"], [100, 400])
    assert ref_data["1"]["GT"].shape[0] == tgt_data["1"]["GT"].shape[0] == 2


def test_reference_case_duplicates_and_no_shared_positions(tmp_path):
    from sstar_b200.utils import read_data
    vcf, r, t = _write(tmp_path, "d.vcf", ["ref1", "tgt1"], ["1\t100\t.\tA\tG\t.\tPASS\t.\tGT\t0|1\t0|0",
                                                              "1\t100\t.\tA\tT\t.\tPASS\t.\tGT\t0|0\t0|1"], ["ref1"], ["tgt1"])
    with pytest.raises(ValueError, match="Duplicate variant positions.*100"):
        read_data(vcf_file=vcf, ref_ind_file=r, tgt_ind_file=t, anc_allele_file=None, is_phased=True)
    vcf, r, t = _write(tmp_path, "n.vcf", ["ref1", "tgt1"], ["1\t100\t.\tA\tG\t.\tPASS\t.\tGT\t0|1\t.|.",
                                                              "1\t200\t.\tA\tG\t.\tPASS\t.\tGT\t.|.\t0|1"], ["ref1"], ["tgt1"])
    with pytest.raises(ValueError, match="No shared variant positions.*chromosome 1"):
        read_data(vcf_file=vcf, ref_ind_file=r, tgt_ind_file=t, anc_allele_file=None, is_phased=True)


def test_reference_case_generator_scopes_alignment_to_requested_chromosome(tmp_path):
    from sstar_b200.generators import WindowDataGenerator
    rows = ["1\t100\t.\tA\tG\t.\tPASS\t.\tGT\t0|1\t0|0", "1\t200\t.\tA\tG\t.\tPASS\t.\tGT\t0|0\t0|1",
            "2\t100\t.\tA\tG\t.\tPASS\t.\tGT\t0|1\t.|.", "2\t200\t.\tA\tG\t.\tPASS\t.\tGT\t.|.\t0|1"]
    vcf, r, t = _write(tmp_path, "c.vcf", ["ref1", "tgt1"], rows, ["ref1"], ["tgt1"])
    gen = WindowDataGenerator(vcf_file=vcf, chr_name="1", ref_ind_file=r, tgt_ind_file=t, win_len=1000, win_step=1000)
    windows = list(gen.get())
    assert len(windows) == 1 and windows[0]["chr_name"] == "1" and np.array_equal(windows[0]["pos"], [100, 200])
    pos = np.array([1, 10, 20, 21, 30])   # closed interval boundaries (test_window_data_generator.py:107-113)
    assert np.array_equal(pos[(pos >= 10) & (pos <= 20)], [10, 20])


@pytest.mark.parametrize("phased", [False, True])
def test_legacy_score_table_format(phased):
    """The legacy table writer (host side of the SNP-set extension) on chains from the oracle:
    identical to the reference's legacy fixture but for the one unit whose score changed since."""
    from oracle import c_oracle
    from sstar_b200.snp_sets import legacy_score_table
    data, _, tgt_names = tiling_port.load_populations(
        os.path.join(FIXTURES, "test.vcf"), os.path.join(FIXTURES, "test.ref.ind.list"),
        os.path.join(FIXTURES, "test.tgt.ind.list"), phased, "21")
    rg, tg, pos = data["21"]
    wins = tiling_port.split_windows(pos, 50000, 10000)
    ws = np.array([w[0] for w in wins], dtype=np.int32)
    we = np.array([w[1] for w in wins], dtype=np.int32)
    labels = [f"{n}_{h}" for n in tgt_names for h in (1, 2)] if phased else list(tgt_names)
    text = legacy_score_table("21", ws, we, labels, *c_oracle.snp_sets(rg, tg, pos, ws, we))
    tag = "phased" if phased else "unphased"
    with open(os.path.join(FIXTURES, f"test.score.{tag}.exp.results.tsv")) as f:
        want = f.read().splitlines()
    got = text.splitlines()
    diff = [(g, w) for g, w in zip(got, want) if g != w]
    assert len(got) == len(want)
    assert len(diff) == (0 if phased else 1)
    if diff:
        assert diff[0][1] == "21\t30000\t80000\tind1\tNA\t5\tNA\tNA"


def test_plan_replicate_shards_cover_and_balance():
    from sstar_b200.sharding import plan_replicate_shards
    rng = np.random.default_rng(3)
    off = np.concatenate([[0], np.cumsum(rng.integers(50, 200, size=1000))])
    for n in (1, 2, 3, 8):
        sh = plan_replicate_shards(off, n)
        assert len(sh) == n and sh[0][0] == 0 and sh[-1][1] == 1000
        assert all(a[1] == b[0] for a, b in zip(sh[:-1], sh[1:]))
        work = [off[b] - off[a] for a, b in sh]
        assert max(work) < 1.15 * (off[-1] / n) + 400
    assert plan_replicate_shards([0], 4) == []
    assert plan_replicate_shards([0, 5, 9], 8) == [(0, 1), (1, 2)]


def test_plan_chromosome_assignment_is_greedy_lpt():
    from sstar_b200.sharding import plan_chromosome_assignment
    costs = {f"chr{i}": c for i, c in enumerate([248, 242, 198, 190, 181, 171, 159, 145, 138, 133, 135, 133, 114, 107,
                                                   102, 90, 83, 80, 58, 64, 46, 50], start=1)}
    plan = plan_chromosome_assignment(costs, 8)
    assert sorted(c for g in plan for c in g) == sorted(costs)
    loads = [sum(costs[c] for c in g) for g in plan]
    assert max(loads) <= 1.12 * sum(costs.values()) / 8
    assert plan_chromosome_assignment({"a": 1}, 4) == [["a"], [], [], []]


def test_call_level_helpers_equal_the_numpy_reductions():
    """The slice forms of the per-call tests (missing allele, hom-alt, dosage) and the column selection
    against the plain numpy expressions they replace (sstar/utils/utils.py:106,282-291,305-306)."""
    from sstar_b200.utils import utils as u
    rng = np.random.default_rng(5)
    gt = rng.integers(-1, 4, (500, 23, 2)).astype(np.int8)
    assert np.array_equal(u._any_allele_missing(gt), (gt < 0).any(axis=2))
    first = gt[:, :, 0]
    assert np.array_equal(u._is_hom_alt(gt), (first > 0) & (gt == first[:, :, None]).all(axis=2))
    assert np.array_equal(u._dosage(gt), gt.sum(axis=2, dtype=np.int16).astype(np.int8))
    for cols in ([3, 1, 17, 22, 0], [4, 5, 6], [0], [], list(range(23))):
        got = u._take_columns(gt, cols)
        assert got.shape == (500, len(cols), 2) and np.array_equal(got, gt[:, cols, :])
    # a filter that drops nothing leaves the arrays alone; one that drops rows copies
    data = {"1": {"POS": np.arange(500), "REF": None, "ALT": None, "GT": gt}}
    assert u.filter_data(data, "1", np.ones(500, bool))["1"]["GT"] is gt
    keep = np.arange(500) % 3 != 0
    assert np.array_equal(u.filter_data(data, "1", keep)["1"]["GT"], gt[keep])
