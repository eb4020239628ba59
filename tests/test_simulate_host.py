"""Host logic of the simulate path (reference sstar/simulate.py:140-195): batching, the
(Replicate, Chromosome, Start, End, Sample) sort with Sample as a string, truncation to nfeature,
the seeded shuffle and the TSV -- against a plain restatement with row dicts and pandas.
CPU only: replicates come from a fake source and are scored by the oracle."""
import os
import random

import numpy as np
import pandas as pd
import pytest

from oracle import c_oracle


def _fake_replicate(rep, seed, nref=3, ntgt=12, ploidy=2, seq_len=50_000):
    rng = np.random.default_rng(1000 + (seed or 0) % 1000 + rep)
    nsite = int(rng.integers(40, 160))
    pos = np.sort(rng.choice(seq_len + 1, size=nsite, replace=False))
    freq = rng.random(nsite)[:, None] ** 2
    gts = (rng.random((nsite, (nref + ntgt) * ploidy)) < freq).astype(np.int8)
    gts[:, : nref * ploidy] &= (rng.random((nsite, 1)) < 0.4)
    return pos.tolist(), gts


@pytest.fixture
def patched(monkeypatch):
    import sstar_b200.msprime_simulator as ms
    from sstar_b200.feature_table import FeatureTable, sample_labels
    from sstar_b200.simulate_batch import build_feature_input, sample_names

    def simulate_replicate(self, rep=None, seed=None):
        return _fake_replicate(rep, seed, self.nref, self.ntgt, self.ploidy, self.seq_len)

    def score_batch(self, replicates, reps):
        scores, counts = [], []
        for pos, gts in replicates:
            r, t, p, s, e = build_feature_input(pos, gts, self.nref, self.ntgt, self.ploidy, self.seq_len, self.is_phased)
            so, no = c_oracle.score_windows(r, t, p, [s], [e], self.match_bonus, self.max_mismatch, self.mismatch_penalty)
            scores.append(so[0]); counts.append(no[0])
        labels = sample_labels(sample_names(self.nref, self.ntgt)[1], self.ploidy, self.is_phased)
        n = len(replicates)
        return FeatureTable("1", [1] * n, [self.seq_len] * n, labels, np.stack(scores), np.stack(counts),
                            replicate=list(reps))

    monkeypatch.setattr(ms.MsprimeSimulator, "simulate_replicate", simulate_replicate)
    monkeypatch.setattr(ms.MsprimeSimulator, "score_batch", score_batch)
    return ms


def _reference_logic(sim, nrep, nfeature, seed, is_shuffled):
    """sstar/simulate.py:140-195 restated with row dicts."""
    from sstar_b200.generators import RandomNumberGenerator
    total, start_rep = [], 0
    key = lambda x: (x["Replicate"], x["Chromosome"], x["Start"], x["End"], x["Sample"])
    while len(total) < nfeature:
        feats = []
        for p in RandomNumberGenerator(start_rep=start_rep, nrep=nrep, seed=seed).get():
            feats.extend(sim.run(**p))
        feats.sort(key=key)
        total.extend(feats)
        start_rep += nrep
    total = total[:nfeature]
    if is_shuffled:
        random.seed(seed)
        random.shuffle(total)
    else:
        total.sort(key=key)
    return pd.DataFrame(total).to_csv(sep="\t", index=False, na_rep="NA")


@pytest.mark.parametrize("is_phased,is_shuffled,nfeature", [(False, False, 100), (True, False, 333),
                                                            (False, True, 250), (True, True, 97)])
def test_simulate_matches_reference_logic(tmp_path, patched, capsys, is_phased, is_shuffled, nfeature):
    from sstar_b200.simulate import simulate
    kw = dict(demo_model_file="unused.yaml", nrep=7, nref=3, ntgt=12, ref_id="Ref", tgt_id="Tgt", ploidy=2,
              seq_len=50_000, mut_rate=1e-8, rec_rate=1e-8, match_bonus=5000, max_mismatch=5,
              mismatch_penalty=-10000, output_dir=str(tmp_path / "out"), output_prefix="t", nfeature=nfeature,
              is_phased=is_phased, is_shuffled=is_shuffled, keep_sim_data=False, seed=913, nprocess=1)
    simulate(**kw)
    got = open(tmp_path / "out" / "t.training.features.tsv").read()
    sim = patched.MsprimeSimulator(kw["demo_model_file"], 3, 12, "Ref", "Tgt", 2, 50_000, 1e-8, 1e-8, "t",
                                   kw["output_dir"], is_phased, 5000, 5, -10000)
    assert got == _reference_logic(sim, 7, nfeature, 913, is_shuffled)
    assert got.count("\n") == nfeature + 1
    assert "Number of obtained features:" in capsys.readouterr().out


def test_simulate_argument_errors(tmp_path, patched):
    from sstar_b200.simulate import simulate
    with pytest.raises(ValueError, match="nfeature must be positive"):
        simulate("x", 1, 1, 1, "a", "b", 2, 10, 0, 0, 1, 1, -1, str(tmp_path), "p", 0, True, False, False, 1, 1)
