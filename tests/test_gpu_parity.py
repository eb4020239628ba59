"""GPU parity tests: the CUDA path (through the C-ABI) vs the oracle and the golden fixtures.

Bar: bit-exact (integer scores, NaN-ness, SNP counts)."""
import os

import numpy as np
import pytest

from conftest import FIXTURES
from oracle import c_oracle, tiling_port
from oracle.sstar_port import NAN_SENTINEL

pytestmark = pytest.mark.gpu

ALGOS = ["auto", "pairwise"]


@pytest.fixture(scope="module")
def eng():
    import __graft_entry__ as g
    g.build()
    from sstar_b200.engine import get_engine
    return get_engine(0)


# ---- the reference's own KATs (tests/test_sstar.py:25-104), through the drop-in Sstar.compute
def test_kat_all_zero_returns_nan():
    from sstar_b200.sstar import Sstar
    out = Sstar.compute(ref_gts=np.zeros((3, 2), dtype=int), tgt_gts=np.zeros((3, 2), dtype=int),
                        pos=np.array([10, 50, 100], dtype=int))["sstar"]
    assert out == [np.nan, np.nan]


def test_kat_two_matching_loci_far_apart():
    from sstar_b200.sstar import Sstar
    tgt = np.array([[1, 1], [2, 0], [1, 1]], dtype=int)
    out = Sstar.compute(ref_gts=np.zeros((3, 2), dtype=int), tgt_gts=tgt,
                        pos=np.array([0, 50, 100], dtype=int), match_bonus=5000)["sstar"]
    assert len(out) == 2 and out[0] == 5100 and np.isnan(out[1])


def test_kat_close_positions_or_too_few_loci():
    from sstar_b200.sstar import Sstar
    out = Sstar.compute(ref_gts=np.zeros((3, 2), dtype=int), tgt_gts=np.ones((3, 2), dtype=int),
                        pos=np.array([0, 5, 8], dtype=int))["sstar"]
    assert len(out) == 2 and np.all(np.isnan(out))
    out = Sstar.compute(ref_gts=np.zeros((2, 2), dtype=int), tgt_gts=np.array([[1, 0], [0, 1]]),
                        pos=np.array([0, 100], dtype=int))["sstar"]
    assert len(out) == 2 and np.all(np.isnan(out))


def test_kat_missing_params():
    from sstar_b200.sstar import Sstar
    with pytest.raises(ValueError, match="Missing required arguments"):
        Sstar.compute()


# ---- golden vectors produced by the reference itself
@pytest.mark.parametrize("algo", ALGOS)
def test_fuzz_golden(eng, fuzz_cases, algo):
    for i, case in enumerate(fuzz_cases):
        b, mx, p = case["params"]
        pos = case["pos"]
        if len(pos) == 0:
            ws, we = [0], [0]
        else:
            ws, we = [int(pos[0])], [int(pos[-1])]
        s, n = eng.score_host(case["ref"], case["tgt"], pos, ws, we, b, mx, p, algo=algo)
        assert np.array_equal(s[0], case["score"]), (i, algo, s[0], case["score"])
        assert np.array_equal(n[0], case["nsnp"]), (i, algo)


def test_fuzz_golden_batched_as_replicates(eng, fuzz_cases):
    # all cases with the same shape/params in ONE replicate-batched call (the simulate path)
    groups = {}
    for case in fuzz_cases:
        key = (case["ref"].shape[1], case["tgt"].shape[1], case["params"])
        groups.setdefault(key, []).append(case)
    for (R, T, (b, mx, p)), cases in groups.items():
        ref = np.concatenate([c["ref"] for c in cases])
        tgt = np.concatenate([c["tgt"] for c in cases])
        pos = np.concatenate([c["pos"] for c in cases])
        off = np.cumsum([0] + [len(c["pos"]) for c in cases])
        s, n = eng.score_replicates_host(ref, tgt, pos, off, -5, 2 ** 31 - 1, b, mx, p)
        assert np.array_equal(s, np.stack([c["score"] for c in cases]))
        assert np.array_equal(n, np.stack([c["nsnp"] for c in cases]))


@pytest.mark.parametrize("algo", ALGOS)
@pytest.mark.parametrize("phased", [False, True])
def test_test0_vcf_windows(eng, test0_golden, phased, algo):
    tag = "phased" if phased else "unphased"
    data, _, _ = tiling_port.load_populations(
        os.path.join(FIXTURES, "test.0.vcf"), os.path.join(FIXTURES, "test.0.ref.ind.list"),
        os.path.join(FIXTURES, "test.0.tgt.ind.list"), phased, "1")
    rg, tg, pos = data["1"]
    wins = test0_golden[f"{tag}_win"]
    s, n = eng.score_host(rg, tg, pos, wins[:, 0], wins[:, 1], algo=algo)
    assert np.array_equal(s, test0_golden[f"{tag}_score"])
    assert np.array_equal(n, test0_golden[f"{tag}_nsnp"])


# ---- seeded synthetic chromosomes vs the C oracle
@pytest.mark.parametrize("algo", ALGOS)
@pytest.mark.parametrize("n_ref,n_tgt,phased,length", [
    (100, 50, False, 1_500_000),   # C2 shape
    (5, 10, False, 600_000),       # tutorial shape: small panel, large n
    (20, 33, True, 800_000),       # odd T (66 columns)
    (3, 70, True, 300_000),        # T > 128 columns
])
def test_synthetic_vs_oracle(eng, algo, n_ref, n_tgt, phased, length):
    _synthetic_vs_oracle(eng, algo, n_ref, n_tgt, phased, length)


def test_c4_panel_shape_vs_oracle(eng):
    """BASELINE configs[3]'s panel: 100 reference + 2,000 target diploids, unphased dosages
    (R = 100, T = 2,000), more than 200 windows of 50 kb / 10 kb, every unit against the C oracle."""
    n_win = _synthetic_vs_oracle(eng, "auto", 100, 2000, False, 2_200_000)
    assert n_win >= 200


def _synthetic_vs_oracle(eng, algo, n_ref, n_tgt, phased, length):
    from sstar_b200.synth import regular_windows, synth_chromosome
    ref, tgt, pos = synth_chromosome(length, n_ref, n_tgt, phased, seed=1000 + n_ref)
    ws, we = regular_windows(pos, 50_000, 10_000)
    s, n = eng.score_host(ref, tgt, pos, ws, we, algo=algo)
    so, no = c_oracle.score_windows(ref, tgt, pos, ws, we)
    assert np.array_equal(s, so)
    assert np.array_equal(n, no)
    return len(ws)


@pytest.mark.parametrize("params", [(5000, 5, -10000), (5000, 0, -10000), (1, 1, 0), (123, 4, -7),
                                    (2_000_000_000, 5, -2_000_000_000)])
def test_parameters_and_missing_data(eng, params):
    from sstar_b200.synth import regular_windows, synth_chromosome
    ref, tgt, pos = synth_chromosome(500_000, 10, 24, False, seed=77)
    rng = np.random.default_rng(5)
    tgt = tgt.copy()
    miss = rng.random(tgt.shape) < 0.01          # sprinkle missing calls: negative dosages
    tgt[miss] = rng.choice([-2, -1], size=int(miss.sum())).astype(np.int8)
    ref = ref.copy()
    ref[rng.random(ref.shape) < 0.002] = -1      # signed sums may cancel to 0
    ws, we = regular_windows(pos, 50_000, 10_000)
    so, no = c_oracle.score_windows(ref, tgt, pos, ws, we, *params)
    for algo in ALGOS:
        s, n = eng.score_host(ref, tgt, pos, ws, we, *params, algo=algo)
        assert np.array_equal(s, so), algo
        assert np.array_equal(n, no), algo


def test_edge_windows(eng):
    # empty windows, windows past the data, single-variant windows, ragged overlap
    pos = np.array([5, 14, 15, 24, 25, 40, 1000, 1010, 1020, 5000], dtype=np.int32)
    tgt = np.array([[1, 2, 0]] * 10, dtype=np.int8)
    tgt[3, 0] = 0
    ref = np.zeros((10, 2), dtype=np.int8)
    ref[6] = [1, 0]
    ws = np.array([1, 1, 6, 41, 1000, 6000, 15, -50], dtype=np.int32)
    we = np.array([5000, 4, 40, 999, 1020, 7000, 15, 2_000_000_000], dtype=np.int32)
    so, no = c_oracle.score_windows(ref, tgt, pos, ws, we)
    for algo in ALGOS:
        s, n = eng.score_host(ref, tgt, pos, ws, we, algo=algo)
        assert np.array_equal(s, so) and np.array_equal(n, no), algo


def test_long_window_many_variants(eng):
    # one window holding thousands of carried sites (pairwise lists spill to the global arena)
    rng = np.random.default_rng(3)
    V = 6000
    pos = np.cumsum(rng.integers(1, 40, size=V)).astype(np.int32)
    tgt = rng.choice([0, 1, 2], size=(V, 5), p=[0.3, 0.4, 0.3]).astype(np.int8)
    tgt[:, 4] = rng.choice([0, 1, 2, 3, -1], size=V).astype(np.int8)  # exotic column -> pairwise path
    ref = np.zeros((V, 3), dtype=np.int8)
    ref[rng.random(V) < 0.2, 1] = 1
    ws = np.array([1, 1000], dtype=np.int32)
    we = np.array([int(pos[-1]), 60000], dtype=np.int32)
    so, no = c_oracle.score_windows(ref, tgt, pos, ws, we)
    for algo in ALGOS:
        s, n = eng.score_host(ref, tgt, pos, ws, we, algo=algo)
        assert np.array_equal(s, so) and np.array_equal(n, no), algo


def test_rejects_unsorted_positions(eng):
    with pytest.raises(ValueError, match="strictly ascending"):
        eng.score_host(np.zeros((3, 1), np.int8), np.ones((3, 1), np.int8), [10, 5, 20], [1], [30])


def test_window_bounds_every_range_length(eng):
    # exercises the 32-ary warp search at every range length around its pivots (1..200 variants
    # left and right of a window edge), checked through Region_ind_SNP_number and the scores
    rng = np.random.default_rng(11)
    V = 1500
    pos = np.cumsum(rng.integers(1, 30, size=V)).astype(np.int32)
    tgt = rng.choice([0, 1, 2], size=(V, 3), p=[0.5, 0.3, 0.2]).astype(np.int8)
    ref = (rng.random((V, 2)) < 0.3).astype(np.int8)
    ws, we = [], []
    for k in range(0, 400):
        ws.append(int(pos[k])); we.append(int(pos[min(V - 1, k + (k * 7) % 230)]))
        ws.append(int(pos[0]) - 3); we.append(int(pos[k]) + (k % 3))
        ws.append(int(pos[V - 1 - k])); we.append(int(pos[-1]) + 5)
    so, no = c_oracle.score_windows(ref, tgt, pos, ws, we)
    s, n = eng.score_host(ref, tgt, pos, ws, we)
    assert np.array_equal(n, no) and np.array_equal(s, so)


def test_host_entry_zero_copy_and_staged_agree(eng):
    # the host entry point on mapped (pinned) buffers streams them over PCIe from inside the
    # kernels; forced staging and pageable buffers take the copy path: same tables either way
    import torch
    from sstar_b200.synth import regular_windows, synth_chromosome
    ref, tgt, pos = synth_chromosome(700_000, 12, 40, False, seed=31)
    ws, we = regular_windows(pos, 50_000, 10_000)
    so, no = c_oracle.score_windows(ref, tgt, pos, ws, we)
    W, T = so.shape
    hint = int(max(np.searchsorted(pos, we, "right") - np.searchsorted(pos, ws, "left")))
    from sstar_b200 import _cabi
    for pinned, staged, flags in [(True, False, 0), (True, True, 0), (False, False, 0),
                                  (True, False, _cabi.FLAG_ZERO_COPY_INPUTS), (True, False, _cabi.FLAG_COPY_INPUTS),
                                  (True, False, _cabi.FLAG_COPY_OUTPUTS | _cabi.FLAG_SINGLE_CHUNK)]:
        mk = (lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()) if pinned else \
             (lambda a: torch.from_numpy(np.ascontiguousarray(a)))
        outs = torch.zeros((W, T), dtype=torch.int64), torch.zeros((W, T), dtype=torch.int32)
        if pinned:
            outs = tuple(o.pin_memory() for o in outs)
        eng.score_pinned(mk(ref), mk(tgt), mk(pos), mk(ws), mk(we), outs[0], outs[1], 5000, 5, -10000,
                         max_win_variants=hint, staged=staged, flags=flags)
        assert np.array_equal(outs[0].numpy(), so), (pinned, staged, flags)
        assert np.array_equal(outs[1].numpy(), no), (pinned, staged, flags)


def test_large_shape_properties(eng):
    # a C3-like shard (R = T = 1000 haplotype columns, ~34k variants, ~800 windows): too big for the
    # oracle as a whole, so check it through properties that do not depend on size, plus samples
    import torch
    from sstar_b200.synth import regular_windows, synth_chromosome_device
    dev = torch.device("cuda", 0)
    r, t, p = synth_chromosome_device(torch, dev, 8_000_000, 500, 500, True, seed=99)
    ref, tgt, pos = r.cpu().numpy(), t.cpu().numpy(), p.cpu().numpy()
    ws, we = regular_windows(pos, 50_000, 10_000)
    W = len(ws)
    s, n = eng.score_host(ref, tgt, pos, ws, we)
    assert s.shape == (W, 1000) and W > 700
    # (a) any split of the window list gives the same rows
    h = W // 2 + 3
    s1, n1 = eng.score_host(ref, tgt, pos, ws[:h], we[:h])
    s2, n2 = eng.score_host(ref, tgt, pos, ws[h:], we[h:])
    assert np.array_equal(np.concatenate([s1, s2]), s) and np.array_equal(np.concatenate([n1, n2]), n)
    # (b) target columns are independent of each other
    sc, nc = eng.score_host(ref, tgt[:, 5::7], pos, ws, we)
    assert np.array_equal(sc, s[:, 5::7]) and np.array_equal(nc, n[:, 5::7])
    # (c) the literal pairwise recurrence agrees on a window sample, (d) so does the oracle
    rng = np.random.default_rng(0)
    sel = np.sort(rng.choice(W, size=40, replace=False))
    sp, np_ = eng.score_host(ref, tgt, pos, ws[sel], we[sel], algo="pairwise")
    assert np.array_equal(sp, s[sel]) and np.array_equal(np_, n[sel])
    so, no = c_oracle.score_windows(ref, tgt, pos, ws[sel[:12]], we[sel[:12]])
    assert np.array_equal(so, s[sel[:12]]) and np.array_equal(no, n[sel[:12]])
    # (e) shifting every position by a constant shifts nothing but the coordinates
    s3, n3 = eng.score_host(ref, tgt, pos + 1_000_003, ws + 1_000_003, we + 1_000_003)
    assert np.array_equal(s3, s) and np.array_equal(n3, n)
    # (f) scores are either NaN or at least the worst single link, counts bound the carried sites
    valid = s != NAN_SENTINEL
    assert (s[valid] >= -10000 * 200).all() and (n >= 0).all()


@pytest.mark.parametrize("n_ref,n_tgt,phased,length", [(3, 2, True, 300_000),      # T = 4
                                                        (25, 32, True, 900_000),    # T = 64, R = 50 (R % 4 != 0)
                                                        (8, 260, False, 700_000),   # T = 260: two item batches
                                                        (50, 500, True, 400_000)])  # T = 1000
def test_wide_and_narrow_panels_with_odd_genotypes(eng, n_ref, n_tgt, phased, length):
    # column counts from 4 to 1000, reference widths that are not multiples of 4, tail rows
    # (V % 32 != 0), missing calls and multi-allelic values, through the pipelined and the staged
    # host entry
    from sstar_b200 import _cabi
    from sstar_b200.synth import regular_windows, synth_chromosome
    ref, tgt, pos = synth_chromosome(length, n_ref, n_tgt, phased, seed=4000 + n_tgt)
    assert tgt.shape[1] % 4 == 0
    rng = np.random.default_rng(n_tgt)
    tgt = tgt.copy()
    m = rng.random(tgt.shape) < 0.003
    tgt[m] = rng.choice([-2, -1, 3], size=int(m.sum())).astype(np.int8)
    ref = ref.copy()
    ref[rng.random(ref.shape) < 0.001] = -1
    ws, we = regular_windows(pos, 50_000, 10_000)
    so, no = c_oracle.score_windows(ref, tgt, pos, ws, we)
    for flags in (0, _cabi.FLAG_STAGED_COPIES):
        s, n = eng.score_host(ref, tgt, pos, ws, we, flags=flags)
        assert np.array_equal(s, so) and np.array_equal(n, no), flags


def test_degenerate_shapes(eng):
    # no windows, no variants, one variant, one column, one reference column
    z8 = lambda r, c: np.zeros((r, c), dtype=np.int8)
    s, n = eng.score_host(z8(5, 2), np.ones((5, 3), np.int8), [1, 20, 40, 60, 80], [], [])
    assert s.shape == (0, 3) and n.shape == (0, 3)
    s, n = eng.score_host(z8(0, 2), z8(0, 3), [], [1, 50], [100, 60])
    assert (s == NAN_SENTINEL).all() and (n == 0).all() and s.shape == (2, 3)
    s, n = eng.score_host(z8(1, 1), np.ones((1, 1), np.int8), [7], [1, 7, 8], [10, 7, 9])
    assert (s == NAN_SENTINEL).all() and n[:, 0].tolist() == [1, 1, 0]
    pos = np.arange(10, 400, 30)
    tgt = np.ones((len(pos), 1), np.int8)
    so, no = c_oracle.score_windows(z8(len(pos), 1), tgt, pos, [1], [1000])
    s, n = eng.score_host(z8(len(pos), 1), tgt, pos, [1], [1000])
    assert np.array_equal(s, so) and np.array_equal(n, no) and s[0, 0] == (pos[-1] - pos[0]) + 5000 * (len(pos) - 1)


def test_understated_window_hint_is_safe(eng):
    """options.max_win_variants far below the truth: the pairwise arena still holds a slab of the
    largest window, so the results do not change (only the pairwise kernel's parallelism does)."""
    import torch
    rng = np.random.default_rng(11)
    V = 4000
    pos = np.cumsum(rng.integers(1, 30, size=V)).astype(np.int32)
    tgt = rng.choice([0, 1, 2], size=(V, 6), p=[0.3, 0.4, 0.3]).astype(np.int8)
    ref = np.zeros((V, 3), dtype=np.int8)
    ref[rng.random(V) < 0.2, 1] = 1
    ws = np.array([1, 500], dtype=np.int32)
    we = np.array([int(pos[-1]), 40000], dtype=np.int32)
    so, no = c_oracle.score_windows(ref, tgt, pos, ws, we)
    dev = eng.tdev
    d = [torch.from_numpy(x).to(dev) for x in (ref, tgt, pos, ws, we)]
    for algo in ALGOS:
        s, n = eng.score_device(*d, algo=algo, max_win_variants=1)
        assert np.array_equal(s.cpu().numpy(), so) and np.array_equal(n.cpu().numpy(), no), algo
    got = eng.snp_sets_device(*d, max_win_variants=1)
    want = c_oracle.snp_sets(ref, tgt, pos, ws, we)
    for a, b in zip(got, want):
        assert np.array_equal(a.cpu().numpy(), b)


def test_random_shapes_through_every_entry(eng):
    """Seeded random chromosomes and window layouts (overlapping, disjoint, nested, empty, unsorted
    ends), odd panel widths, random parameters, occasional missing / multi-allelic calls: the host
    entry (in place and staged), the device entry and the SNP-set entry against the C oracle."""
    import torch
    from sstar_b200 import _cabi
    rng = np.random.default_rng(20240607)
    for case in range(24):
        V = int(rng.choice([0, 1, 2, 31, 32, 33, 500, 3000]))
        R = int(rng.integers(1, 70))
        T = int(rng.integers(1, 140))
        pos = np.cumsum(rng.integers(1, 60, size=V)).astype(np.int32)
        ref = (rng.random((V, R)) < 0.02).astype(np.int8)
        tgt = rng.choice([0, 1, 2], size=(V, T), p=[0.6, 0.3, 0.1]).astype(np.int8)
        if case % 3 == 0 and V:
            m = rng.random(tgt.shape) < 0.01
            tgt[m] = rng.choice([-1, 3, 7], size=int(m.sum())).astype(np.int8)
            ref[rng.random(ref.shape) < 0.005] = -1
        span = int(pos[-1]) if V else 1000
        W = int(rng.choice([1, 2, 7, 40, 300]))
        ws = rng.integers(-50, span + 50, size=W).astype(np.int32)
        we = (ws + rng.integers(0, max(2, span // 3), size=W)).astype(np.int32)
        if case % 2 == 0:  # the usual layout: sorted starts
            order = np.argsort(ws, kind="stable")
            ws, we = ws[order], we[order]
        params = [(5000, 5, -10000), (1, 1, 0), (77, 0, -3), (2_000_000_000, 5, -2_000_000_000)][case % 4]
        so, no = c_oracle.score_windows(ref, tgt, pos, ws, we, *params)
        for flags in (0, _cabi.FLAG_STAGED_COPIES, _cabi.FLAG_COPY_INPUTS):
            s, n = eng.score_host(ref, tgt, pos, ws, we, *params, flags=flags)
            assert np.array_equal(s, so) and np.array_equal(n, no), (case, flags, V, R, T, W)
        if V:
            d = [torch.from_numpy(x).to(eng.tdev) for x in (ref, tgt, pos, ws, we)]
            s, n = eng.score_device(*d, *params)
            assert np.array_equal(s.cpu().numpy(), so) and np.array_equal(n.cpu().numpy(), no), (case, "device")
            got = eng.snp_sets_device(*d, *params)
            want = c_oracle.snp_sets(ref, tgt, pos, ws, we, *params)
            for a, b in zip(got, want):
                assert np.array_equal(a.cpu().numpy(), b), (case, "snp sets")


def test_many_small_windows_in_place(eng):
    # 70,000 windows over 20,000 variants: the in-place host entry's fetch CTAs hit their cap and
    # the bounds CTAs run in several waves behind them
    rng = np.random.default_rng(8)
    V, T = 20_000, 12
    pos = np.cumsum(rng.integers(1, 30, size=V)).astype(np.int32)
    ref = (rng.random((V, 6)) < 0.05).astype(np.int8)
    tgt = rng.choice([0, 1, 2], size=(V, T), p=[0.5, 0.3, 0.2]).astype(np.int8)
    ws = np.sort(rng.integers(0, int(pos[-1]), size=70_000)).astype(np.int32)
    we = (ws + rng.integers(0, 400, size=ws.size)).astype(np.int32)
    so, no = c_oracle.score_windows(ref, tgt, pos, ws, we)
    for flags in (0, 1):
        s, n = eng.score_host(ref, tgt, pos, ws, we, flags=flags)
        assert np.array_equal(s, so) and np.array_equal(n, no), flags


def test_rows_too_wide_to_stage(eng):
    # 7,000 genotype bytes per variant row: a 32-row tile does not fit shared memory, so the pack
    # kernel takes its byte path straight from global memory; T = 4,000 is 32 column chunks
    rng = np.random.default_rng(12)
    V, R, T = 330, 3000, 4000
    pos = np.cumsum(rng.integers(20, 400, size=V)).astype(np.int32)
    ref = (rng.random((V, R)) < 0.0004).astype(np.int8)
    ref[rng.random(V) < 0.1, 7] = -1            # signed sums: a -1 and a +1 cancel
    tgt = rng.choice([0, 1, 2], size=(V, T), p=[0.7, 0.2, 0.1]).astype(np.int8)
    tgt[rng.random((V, T)) < 0.0005] = -1       # a few missing calls
    ws = np.arange(1, int(pos[-1]), 10_000, dtype=np.int32)
    we = (ws + 49_999).astype(np.int32)
    so, no = c_oracle.score_windows(ref, tgt, pos, ws, we)
    for flags in (0, 1):
        s, n = eng.score_host(ref, tgt, pos, ws, we, flags=flags)
        assert np.array_equal(s, so) and np.array_equal(n, no), flags


@pytest.mark.parametrize("R,T", [(60, 4000), (16, 2512), (1200, 1600),   # T % 16 == 0: row segments go through TMA
                                 (60, 4004), (33, 3001), (7, 2406),        # staged with their own 16-byte phases
                                 (2000, 1000)])                            # reference too wide for chunks: whole rows, 96 KB tiles
def test_wide_panels_take_column_chunks(eng, R, T):
    # 32 rows of R + T bytes exceed a third of an SM's shared memory: the pack stages one column
    # chunk of the target rows per CTA, every chunk redoing the row sums
    rng = np.random.default_rng(R + T)
    V = 1000 + R % 7          # tail rows (V % 32 != 0)
    pos = np.cumsum(rng.integers(20, 400, size=V)).astype(np.int32)
    ref = (rng.random((V, R)) < 0.01).astype(np.int8)
    ref[rng.random(V) < 0.1, R // 2] = -1
    tgt = rng.choice([0, 1, 2], size=(V, T), p=[0.7, 0.2, 0.1]).astype(np.int8)
    tgt[rng.random((V, T)) < 0.0005] = rng.choice([-1, 3]).astype(np.int8)
    ws = np.arange(1, int(pos[-1]), 10_000, dtype=np.int32)
    we = (ws + 49_999).astype(np.int32)
    so, no = c_oracle.score_windows(ref, tgt, pos, ws, we)
    for flags in (0, 1):
        s, n = eng.score_host(ref, tgt, pos, ws, we, flags=flags)
        assert np.array_equal(s, so) and np.array_equal(n, no), flags


def test_host_threads_share_the_kernels_shared_memory_limits(eng):
    """cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is process-wide: a call from another host
    thread that needs little shared memory must not lower the limit under a later large launch
    (the multi-GPU routes run one host thread per device next to the caller's own thread)."""
    from concurrent.futures import ThreadPoolExecutor
    rng = np.random.default_rng(5)

    def case(V, R, T):
        pos = np.sort(rng.choice(np.arange(1, 40 * V), size=V, replace=False)).astype(np.int32)
        ref = (rng.random((V, R)) < 0.02).astype(np.int8)
        tgt = (rng.random((V, T)) < 0.05).astype(np.int8)
        ws = np.arange(1, int(pos[-1]), 10_000, dtype=np.int32)
        return ref, tgt, pos, ws, (ws + 49_999).astype(np.int32)

    wide, small = case(3000, 100, 1900), case(300, 4, 6)   # 64 KB pack tiles vs a few hundred bytes
    want_w, want_s = c_oracle.score_windows(*wide), c_oracle.score_windows(*small)
    with ThreadPoolExecutor(max_workers=1) as other:
        for _ in range(2):
            got = eng.score_host(*wide)
            assert np.array_equal(got[0], want_w[0]) and np.array_equal(got[1], want_w[1])
            got = other.submit(eng.score_host, *small).result()
            assert np.array_equal(got[0], want_s[0]) and np.array_equal(got[1], want_s[1])


@pytest.mark.parametrize("n_ref,n_tgt,phased,length,values,rate", [
    (10, 24, False, 600_000, [-2, -1], 0.01),          # T = 24: -1 / -2 only, entries from the three planes
    (10, 25, False, 600_000, [-2, -1], 0.01),          # T = 25: ragged last quad of every row
    (50, 500, True, 400_000, [-1], 0.005),             # T = 1000 haplotypes, missing alleles (the C3 shape)
    (50, 500, True, 400_000, [-2, -1, 3], 0.004),      # ... with multi-allelic dosages: groups of level 2 among level 1
    (8, 130, False, 500_000, [-2, -1, -3, 4], 0.01),   # T = 130 (two column chunks of the score grid), |g| > 2 both signs
    (8, 64, False, 500_000, [-128, 127, -2], 0.01),    # the int8 extremes
    (50, 500, True, 400_000, [-1], 0.0003),            # sparse missing calls: a few odd blocks per warp step (the ballots)
    (20, 150, False, 500_000, [-2, -1, 5, -7], 0.0004),  # ... with |g| > 2 among them, T = 150 (ragged last quad: 150 % 4 = 2)
])
def test_missing_calls_through_the_launch_pipeline(eng, n_ref, n_tgt, phased, length, values, rate):
    """Missing calls (-1 / -2) and other int8 genotypes through pack -> score launches (no one-launch kernel):
    sub-groups whose other genotypes all are -1 / -2 build their entries from the planes (|g| == 2, sign) and
    keep the classes in registers, any other byte switches the sub-group to the bytes of the matrix; clean
    groups, flagged groups and both kinds of flagged groups next to each other.  Bit-exact vs the oracle for
    several mismatch bounds, both algorithms; some rows cluster the missing calls so that whole 32-row groups
    stay clean while their neighbours are flagged."""
    from sstar_b200 import _cabi
    from sstar_b200.synth import regular_windows, synth_chromosome
    ref, tgt, pos = synth_chromosome(length, n_ref, n_tgt, phased, seed=9100 + n_tgt)
    rng = np.random.default_rng(n_tgt + len(values))
    tgt = tgt.copy()
    m = rng.random(tgt.shape) < rate
    m[(np.arange(len(pos)) // 32) % 3 == 0] = False  # every third 32-row group keeps clean genotypes
    tgt[m] = rng.choice(values, size=int(m.sum())).astype(np.int8)
    ref = ref.copy()
    ref[rng.random(ref.shape) < 0.001] = -1
    ws, we = regular_windows(pos, 50_000, 10_000)
    for params in [(5000, 5, -10000), (5000, 1, -10000), (77, 3, -5), (5000, 0, -10000)]:
        so, no = c_oracle.score_windows(ref, tgt, pos, ws, we, *params)
        for algo in ALGOS:
            s, n = eng.score_host(ref, tgt, pos, ws, we, *params, algo=algo, flags=_cabi.FLAG_NO_FUSED)
            assert np.array_equal(n, no), (params, algo)
            assert np.array_equal(s, so), (params, algo)
    # the workspace is reused: clean data right after flagged data must not see stale sign words
    ref2, tgt2, pos2 = synth_chromosome(length, n_ref, n_tgt, phased, seed=9100 + n_tgt)
    so, no = c_oracle.score_windows(ref2, tgt2, pos2, ws, we)
    s, n = eng.score_host(ref2, tgt2, pos2, ws, we, flags=_cabi.FLAG_NO_FUSED)
    assert np.array_equal(s, so) and np.array_equal(n, no)
