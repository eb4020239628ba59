"""GPU parity of the one-launch path (csrc/fused_kernels.cuh) through the C-ABI device entry
points: forced on every shape it accepts, against the C oracle and the golden fixtures, and
against the pack -> score launches (which must give the same tables).  Bit-exact."""
import numpy as np
import pytest

from oracle import c_oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    import __graft_entry__ as g
    g.build()
    from sstar_b200.engine import get_engine
    return get_engine(0)


def _score(eng, ref, tgt, pos, ws, we, params=(5000, 5, -10000), hint=None, fused=True, offset=0):
    """Device entry point; ``offset`` rows of padding in front make the base pointers unaligned."""
    import torch
    from sstar_b200 import _cabi
    from sstar_b200.engine import max_window_variants
    dev = eng.tdev
    ref = np.ascontiguousarray(ref, dtype=np.int8)
    tgt = np.ascontiguousarray(tgt, dtype=np.int8)
    pos = np.ascontiguousarray(pos, dtype=np.int32)
    ws = np.ascontiguousarray(ws, dtype=np.int32)
    we = np.ascontiguousarray(we, dtype=np.int32)

    def up(a):
        if offset == 0 or a.ndim == 1:
            return torch.from_numpy(a).to(dev)
        pad = torch.zeros((offset,) + a.shape[1:], dtype=torch.int8, device=dev)
        return torch.cat([pad, torch.from_numpy(a).to(dev)])[offset:]

    if hint is None:
        hint = max_window_variants(pos, ws, we)
    flags = _cabi.FLAG_FORCE_FUSED if fused else _cabi.FLAG_NO_FUSED
    s, n = eng.score_device(up(ref), up(tgt), torch.from_numpy(pos).to(dev), torch.from_numpy(ws).to(dev),
                            torch.from_numpy(we).to(dev), *params, max_win_variants=hint, flags=flags)
    torch.cuda.synchronize()
    assert eng.launches == (1 if fused and tgt.shape[1] <= 128 and len(ws) else eng.launches)
    return s.cpu().numpy(), n.cpu().numpy()


def _check(eng, ref, tgt, pos, ws, we, params=(5000, 5, -10000), **kw):
    so, no = c_oracle.score_windows(ref, tgt, pos, ws, we, *params)
    s, n = _score(eng, ref, tgt, pos, ws, we, params, **kw)
    assert np.array_equal(s, so), np.argwhere(s != so)[:5]
    assert np.array_equal(n, no), np.argwhere(n != no)[:5]
    return so


def test_fuzz_golden_single_windows(eng, fuzz_cases):
    """The reference-generated single-window cases (tests/golden/make_golden.py), each a fused call."""
    for c in fuzz_cases:
        if len(c["pos"]) == 0:
            continue
        ws = np.array([int(c["pos"].min())], dtype=np.int32)
        we = np.array([int(c["pos"].max())], dtype=np.int32)
        s, n = _score(eng, c["ref"], c["tgt"], c["pos"], ws, we, c["params"])
        assert np.array_equal(s[0], c["score"]) and np.array_equal(n[0], c["nsnp"])


@pytest.mark.parametrize("n_ref,n_tgt,phased,length", [
    (100, 50, False, 1_500_000),   # C2 shape: 10 windows scored at a time by a 512-thread CTA
    (5, 10, False, 600_000),       # tutorial shape: narrow reference panel (unaligned 5-byte rows)
    (20, 33, True, 800_000),       # 66 columns: 7 windows at a time
    (3, 64, True, 300_000),        # T = 128: the widest panel the fused kernel takes
    (40, 16, True, 700_000),       # T = 32 exactly
    (300, 60, False, 400_000),     # wide reference panel: the tile limit shrinks the home rows per CTA
])
def test_synthetic_shapes(eng, n_ref, n_tgt, phased, length):
    from sstar_b200.synth import regular_windows, synth_chromosome
    ref, tgt, pos = synth_chromosome(length, n_ref, n_tgt, phased, seed=2000 + n_ref)
    ws, we = regular_windows(pos, 50_000, 10_000)
    so = _check(eng, ref, tgt, pos, ws, we)
    s2, n2 = _score(eng, ref, tgt, pos, ws, we, fused=False)   # the pack -> score launches agree
    assert np.array_equal(s2, so)
    _check(eng, ref, tgt, pos, ws, we, offset=3)               # unaligned bases: spans staged by vector loads
    _check(eng, ref[:-5], tgt[:-5], pos[:-5], ws, we)          # V not a multiple of 32: ragged last group


@pytest.mark.parametrize("params", [(5000, 5, -10000), (5000, 0, -10000), (1, 1, 0), (123, 4, -7),
                                    (2 ** 27, 2, -(2 ** 27))])
def test_parameters_missing_calls_and_many_genotype_values(eng, params):
    from sstar_b200.synth import regular_windows, synth_chromosome
    rng = np.random.default_rng(7)
    ref, tgt, pos = synth_chromosome(500_000, 12, 9, False, seed=41)
    ws, we = regular_windows(pos, 50_000, 10_000)
    _check(eng, ref, tgt, pos, ws, we, params)
    few = tgt.copy()
    few[rng.random(few.shape) < 0.03] = -1   # missing calls: negative dosages (utils.py:305-306)
    few[rng.random(few.shape) < 0.01] = 3
    miss_ref = ref.copy()
    miss_ref[rng.random(ref.shape) < 0.01] = -1
    miss_ref[rng.random(ref.shape) < 0.01] = 1  # +1 / -1 can cancel in the signed row sum (sstar.py:87)
    _check(eng, miss_ref, few, pos, ws, we, params)
    # more than 8 genotype values per unit: the class-table streaming DP inside the same kernel
    many = tgt.astype(np.int16)
    hit = rng.random(many.shape) < 0.4
    many[hit] = rng.integers(-128, 128, size=int(hit.sum()))
    dense = np.sort(rng.choice(np.arange(1, 4 * len(pos)), size=len(pos), replace=False)).astype(np.int32)
    dws, dwe = regular_windows(dense, 500, 100)
    _check(eng, np.zeros_like(ref), many.astype(np.int8), dense, dws[:300], dwe[:300], params)


def test_windows_larger_than_the_tile_are_streamed(eng):
    """An understated size hint (or a tile cut by the shared-memory limit) is safe: the CTA streams
    such windows from global memory through the class-table DP."""
    from sstar_b200.synth import regular_windows, synth_chromosome
    ref, tgt, pos = synth_chromosome(600_000, 30, 20, False, seed=77)
    ws, we = regular_windows(pos, 50_000, 10_000)
    for hint in (1, 8, 40, 150):
        _check(eng, ref, tgt, pos, ws, we, hint=hint)
    # one window over everything next to normal ones
    ws2 = np.concatenate([ws[:5], [1], ws[5:9]]).astype(np.int32)
    we2 = np.concatenate([we[:5], [int(pos[-1])], we[5:9]]).astype(np.int32)
    _check(eng, ref, tgt, pos, ws2, we2, hint=64)
    rng = np.random.default_rng(3)
    tgt_m = tgt.copy()
    tgt_m[rng.random(tgt.shape) < 0.02] = -2
    _check(eng, ref, tgt_m, pos, ws2, we2, hint=16)


def test_edge_windows_and_tiny_inputs(eng):
    rng = np.random.default_rng(11)
    pos = np.array([5, 100, 104, 113, 2000, 2001, 2010, 2019, 2028, 5000, 5011, 90_000], dtype=np.int32)
    ref = np.zeros((12, 3), dtype=np.int8)
    ref[4, 1] = 1
    tgt = rng.integers(0, 3, size=(12, 5)).astype(np.int8)
    ws = np.array([1, 1, 50, 2000, 2001, 6000, 95_000, -50, 5011, 90_000, 1, 113], dtype=np.int32)
    we = np.array([100_000, 4, 120, 2028, 2028, 80_000, 99_000, 5, 5011, 90_000, 1, 104], dtype=np.int32)  # last: end < start
    _check(eng, ref, tgt, pos, ws, we)
    _check(eng, ref, tgt, pos, ws[::-1].copy(), we[::-1].copy())            # unordered windows
    _check(eng, ref, tgt, pos, np.repeat(ws, 3), np.repeat(we, 3))          # duplicates
    for V in (1, 2, 3, 31, 32, 33):
        _check(eng, ref[:V] if V <= 12 else np.zeros((V, 3), np.int8),
               tgt[:V] if V <= 12 else rng.integers(0, 3, size=(V, 5)).astype(np.int8),
               pos[:V] if V <= 12 else (np.arange(V, dtype=np.int32) * 37 + 1), ws[:4], we[:4])


def test_columns_beyond_the_fused_limit_take_the_launch_pipeline(eng):
    from sstar_b200.synth import regular_windows, synth_chromosome
    ref, tgt, pos = synth_chromosome(300_000, 3, 70, True, seed=5)   # T = 140 > 128
    ws, we = regular_windows(pos, 50_000, 10_000)
    so, no = c_oracle.score_windows(ref, tgt, pos, ws, we)
    s, n = _score(eng, ref, tgt, pos, ws, we, fused=True)
    assert eng.launches > 1 and np.array_equal(s, so) and np.array_equal(n, no)


def test_replicate_batches_take_the_fused_kernel(eng):
    """sstar_b200_score_replicates: disjoint windows, positions restart per replicate."""
    import torch
    rng = np.random.default_rng(19)
    n_rep, L = 300, 50_000
    refs, tgts, poss, off = [], [], [], [0]
    for r in range(n_rep):
        v = int(rng.integers(0, 220)) if r % 17 else 0            # some replicates are empty
        p = np.sort(rng.choice(np.arange(0, L + 200), size=v, replace=False)).astype(np.int32)  # 0 and > L fall outside
        refs.append((rng.random((v, 6)) < 0.1).astype(np.int8))
        tgts.append(rng.integers(0, 3, size=(v, 11)).astype(np.int8) * (rng.random((v, 11)) < 0.5))
        poss.append(p)
        off.append(off[-1] + v)
    ref, tgt, pos = np.concatenate(refs), np.concatenate(tgts).astype(np.int8), np.concatenate(poss)
    dev = eng.tdev
    d = [torch.from_numpy(np.ascontiguousarray(x)).to(dev) for x in (ref, tgt, pos, np.asarray(off, np.int32))]
    hint = int(np.diff(off).max())
    from sstar_b200 import _cabi
    got = {}
    for name, flags in (("fused", 0), ("launches", _cabi.FLAG_NO_FUSED)):
        s, n = eng.score_replicates_device(d[0], d[1], d[2], d[3], 1, L, max_win_variants=hint, flags=flags)
        torch.cuda.synchronize()
        got[name] = (s.cpu().numpy(), n.cpu().numpy(), eng.launches)
    assert got["fused"][2] == 1 and got["launches"][2] > 1
    for r in range(n_rep):
        a, b = off[r], off[r + 1]
        so, no = c_oracle.score_windows(ref[a:b], tgt[a:b], pos[a:b], [1], [L])
        for name in got:
            assert np.array_equal(got[name][0][r], so[0]) and np.array_equal(got[name][1][r], no[0]), (name, r)


def test_many_windows_per_cta_and_several_list_rounds(eng):
    """Dense window lists: more windows than a CTA scores at a time (several batches) and more than
    one round of the window-list scan (W > 1,024), in shuffled order with duplicates."""
    from sstar_b200.synth import regular_windows, synth_chromosome
    ref, tgt, pos = synth_chromosome(1_200_000, 30, 25, False, seed=91)
    ws, we = regular_windows(pos, 50_000, 500)          # ~2,300 windows, ~3 rows apart
    assert len(ws) > 2 * 1024
    _check(eng, ref, tgt, pos, ws, we)
    rng = np.random.default_rng(2)
    perm = rng.permutation(len(ws))
    perm = np.concatenate([perm, perm[:300]])
    _check(eng, ref, tgt, pos, ws[perm], we[perm])
    # an understated hint with this many windows: every round leaves windows to the streaming fallback
    _check(eng, ref[:4000], tgt[:4000], pos[:4000], ws[:1200], we[:1200], hint=60)


def test_random_shapes_windows_and_genotypes(eng):
    """Fuzz of the one-launch path: panel shapes, row counts, position densities, window lists (any
    order, any length), genotype alphabets, parameters, hints and base alignments drawn at random;
    every case twice (a race would show as a difference between the runs) and against the oracle."""
    rng = np.random.default_rng(20260101)
    for case in range(60):
        R = int(rng.choice([1, 2, 3, 5, 16, 17, 40, 100, 129]))
        T = int(rng.choice([1, 2, 7, 10, 31, 32, 33, 50, 64, 96, 127, 128]))
        V = int(rng.choice([1, 5, 31, 33, 200, 777, 3000]))
        span = int(rng.choice([V * 3, V * 40, V * 900]))
        pos = np.sort(rng.choice(np.arange(1, span + 2), size=V, replace=False)).astype(np.int32)
        ref = (rng.random((V, R)) < rng.choice([0.0, 0.02, 0.3])).astype(np.int8)
        alphabet = [np.array([0, 1]), np.array([0, 1, 2]), np.array([0, 0, 0, 1, 2, -1]), np.arange(-3, 6)][int(rng.integers(0, 4))]
        tgt = rng.choice(alphabet, size=(V, T)).astype(np.int8)
        tgt[rng.random((V, T)) < 0.5] = 0
        if rng.random() < 0.3:
            ref[rng.random((V, R)) < 0.02] = -1
        W = int(rng.choice([1, 3, 40, 300]))
        ws = rng.integers(-10, span + 10, size=W).astype(np.int32)
        we = (ws + rng.choice([0, 9, 50, span // 20 + 1, span // 3 + 1, span], size=W)).astype(np.int32)
        if rng.random() < 0.4:
            order = np.argsort(ws, kind="stable")
            ws, we = ws[order], we[order]
        params = [(5000, 5, -10000), (1, 1, 0), (7, 0, -3), (100, 2, -50)][int(rng.integers(0, 4))]
        hint = None if rng.random() < 0.6 else int(rng.choice([1, 5, 64]))
        off = int(rng.choice([0, 0, 1, 3]))
        so, no = c_oracle.score_windows(ref, tgt, pos, ws, we, *params)
        s1, n1 = _score(eng, ref, tgt, pos, ws, we, params, hint=hint, offset=off)
        s2, n2 = _score(eng, ref, tgt, pos, ws, we, params, hint=hint, offset=off)
        assert np.array_equal(s1, s2) and np.array_equal(n1, n2), f"case {case}: two runs differ"
        assert np.array_equal(s1, so) and np.array_equal(n1, no), f"case {case}: R={R} T={T} V={V} W={W} hint={hint}"


def test_stream_of_independent_calls(eng):
    """SSTAR_B200_FLAG_INDEPENDENT: a stream of DIFFERENT chromosomes (own buffers each) scored back to
    back -- direct launches, then one graph (ScoreEngine.plan_device_stream) replayed twice; consecutive
    one-launch kernels overlap.  Every table against the oracle; an unflagged call and a copy behind the
    stream keep full stream order."""
    import torch
    from oracle import tiling_port
    from sstar_b200 import _cabi
    from sstar_b200.engine import max_window_variants
    from sstar_b200.synth import synth_chromosome
    dev = eng.tdev
    params = (5000, 5, -10000)
    chroms, want = [], []
    for i, (length, n_ref, n_tgt, phased) in enumerate([(900_000, 30, 20, False), (400_000, 100, 50, False),
                                                         (1_500_000, 10, 25, True), (300_000, 5, 3, False),
                                                         (2_000_000, 60, 40, False), (50_000, 4, 64, True)]):
        ref, tgt, pos = synth_chromosome(length=length, n_ref=n_ref, n_tgt=n_tgt, phased=phased, seed=100 + i)
        if i == 2:
            tgt = tgt.copy()
            tgt[::97, ::3] = -1  # missing calls: the general walk inside the one-launch kernel
        wins = tiling_port.split_windows(pos, 50_000, 10_000)
        ws = np.asarray([w[0] for w in wins], dtype=np.int32)
        we = np.asarray([w[1] for w in wins], dtype=np.int32)
        chroms.append((ref, tgt, pos.astype(np.int32), ws, we))
        want.append(c_oracle.score_windows(ref, tgt, pos, ws, we, *params))
    hint = max(max_window_variants(c[2], c[3], c[4]) for c in chroms)
    cases = []
    for ref, tgt, pos, ws, we in chroms:
        d = [torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in (ref, tgt, pos, ws, we)]
        out = (torch.zeros((len(ws), tgt.shape[1]), dtype=torch.int64, device=dev),
               torch.zeros((len(ws), tgt.shape[1]), dtype=torch.int32, device=dev))
        cases.append(tuple(d) + (out,))

    def check(tag):
        torch.cuda.synchronize()
        for i, c in enumerate(cases):
            assert np.array_equal(c[5][0].cpu().numpy(), want[i][0]), f"{tag}: chromosome {i} scores"
            assert np.array_equal(c[5][1].cpu().numpy(), want[i][1]), f"{tag}: chromosome {i} SNP counts"
            c[5][0].zero_(); c[5][1].zero_()

    # direct launches, three rounds back to back
    for _ in range(3):
        for c in cases:
            eng.score_device(*c[:5], *params, max_win_variants=hint, out=c[5], flags=_cabi.FLAG_INDEPENDENT)
            assert eng.launches == 1
    check("direct")
    # an unflagged call that CONSUMES a flagged call's output position: a device copy of its table right
    # behind the stream must see the finished table (full stream order for everything unflagged)
    for c in cases:
        eng.score_device(*c[:5], *params, max_win_variants=hint, out=c[5], flags=_cabi.FLAG_INDEPENDENT)
    snap = [c[5][0].clone() for c in cases]
    torch.cuda.synchronize()
    for i, s in enumerate(snap):
        assert np.array_equal(s.cpu().numpy(), want[i][0]), f"copy behind the stream: chromosome {i}"
    check("before the graph")
    plan = eng.plan_device_stream(cases, *params, max_win_variants=hint)
    assert plan.launches == len(cases)
    for c in cases:
        c[5][0].zero_(); c[5][1].zero_()
    plan.run()
    plan.run()
    check("graph")
    plan.run()
    check("graph again")


def test_host_stream_of_independent_calls(eng):
    """The host entry with SSTAR_B200_FLAG_INDEPENDENT: different chromosomes back to back, TWO scratch
    buffers cycled (so every scratch comes round again while earlier calls may still be running: the
    library orders its copies behind the previous call that used it), one synchronize at the end; then
    an unflagged call on a scratch the stream used.  Every table against the oracle."""
    import torch
    from oracle import tiling_port
    from sstar_b200 import _cabi
    from sstar_b200.engine import max_window_variants
    from sstar_b200.synth import synth_chromosome
    params = (5000, 5, -10000)
    jobs = []
    for i, (length, n_ref, n_tgt, phased) in enumerate([(2_000_000, 100, 50, False), (300_000, 30, 20, False),
                                                         (2_000_000, 100, 50, False), (1_000_000, 10, 25, True),
                                                         (2_000_000, 100, 50, False), (600_000, 8, 64, True),
                                                         (2_000_000, 100, 50, False)]):
        ref, tgt, pos = synth_chromosome(length=length, n_ref=n_ref, n_tgt=n_tgt, phased=phased, seed=300 + i)
        wins = tiling_port.split_windows(pos, 50_000, 10_000)
        ws = np.asarray([w[0] for w in wins], dtype=np.int32)
        we = np.asarray([w[1] for w in wins], dtype=np.int32)
        pos = pos.astype(np.int32)
        want = c_oracle.score_windows(ref, tgt, pos, ws, we, *params)
        pin = [torch.from_numpy(np.ascontiguousarray(a)).pin_memory() for a in (ref, tgt, pos, ws, we)]
        out = (torch.zeros((len(ws), tgt.shape[1]), dtype=torch.int64).pin_memory(),
               torch.zeros((len(ws), tgt.shape[1]), dtype=torch.int32).pin_memory())
        jobs.append((pin, out, want, max_window_variants(pos, ws, we)))
    for rounds in range(2):
        for i, (pin, out, _want, hint) in enumerate(jobs):
            eng.score_pinned(*pin, out[0], out[1], *params, max_win_variants=hint, sync=False,
                             flags=_cabi.FLAG_INDEPENDENT, scratch_slot=1 + i % 2)
        torch.cuda.synchronize()
        for i, (_pin, out, want, _hint) in enumerate(jobs):
            assert np.array_equal(out[0].numpy(), want[0]), f"round {rounds}: chromosome {i} scores"
            assert np.array_equal(out[1].numpy(), want[1]), f"round {rounds}: chromosome {i} SNP counts"
            out[0].zero_(); out[1].zero_()
    # stream, then an unflagged call on the same scratch without a synchronize in between
    pin, out, want, hint = jobs[0]
    eng.score_pinned(*pin, out[0], out[1], *params, max_win_variants=hint, sync=False, flags=_cabi.FLAG_INDEPENDENT,
                     scratch_slot=1)
    pin1, out1, want1, hint1 = jobs[1]
    eng.score_pinned(*pin1, out1[0], out1[1], *params, max_win_variants=hint1, sync=True, scratch_slot=1)
    torch.cuda.synchronize()
    assert np.array_equal(out[0].numpy(), want[0]) and np.array_equal(out1[0].numpy(), want1[0])
    assert np.array_equal(out[1].numpy(), want[1]) and np.array_equal(out1[1].numpy(), want1[1])
