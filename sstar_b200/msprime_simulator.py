"""The per-replicate simulation job (reference sstar/msprime_simulator.py:30-277).

Same constructor and ``run(rep, seed)`` contract (a list of per-sample row dicts tagged with
``Replicate``).  The coalescent simulation itself is msprime's (imported lazily: it is an
un-vendored dependency of the reference); scoring goes through the replicate-batched device call.
``simulate_replicate`` is the seam tests and other replicate sources override.
"""
from __future__ import annotations

import os

from .simulate_batch import score_replicate_batch


class MsprimeSimulator:
    def __init__(self, demo_model_file: str, nref: int, ntgt: int, ref_id: str, tgt_id: str, ploidy: int,
                 seq_len: int, mut_rate: float, rec_rate: float, output_prefix: str, output_dir: str,
                 is_phased: bool, match_bonus: int, max_mismatch: int, mismatch_penalty: int,
                 keep_sim_data: bool = False):
        self.demo_model_file = demo_model_file
        self.nref = nref
        self.ntgt = ntgt
        self.ref_id = ref_id
        self.tgt_id = tgt_id
        self.ploidy = ploidy
        self.seq_len = seq_len
        self.mut_rate = mut_rate
        self.rec_rate = rec_rate
        self.output_prefix = output_prefix
        self.output_dir = output_dir
        self.test_tgt_id = self.tgt_id
        self.is_phased = is_phased
        self.match_bonus = match_bonus
        self.max_mismatch = max_mismatch
        self.mismatch_penalty = mismatch_penalty
        self.keep_sim_data = keep_sim_data

    def simulate_replicate(self, rep: int = None, seed: int = None):
        """One replicate -> (positions, genotype_matrix [sites, (nref+ntgt)*ploidy]) exactly as
        ``[int(site.position) for site in ts.sites()]`` and ``ts.genotype_matrix()`` give them
        [msprime_simulator.py:145-166,207-208]."""
        try:
            import demes
            import msprime
        except ImportError as e:  # pragma: no cover - msprime is not installed in the build image
            raise ImportError("msprime and demes are required to simulate replicates "
                              "(they are dependencies of the reference, not of sstar_b200)") from e
        output_dir = self.output_dir if rep is None else os.path.join(self.output_dir, str(rep))
        output_prefix = self.output_prefix if rep is None else f"{self.output_prefix}.{rep}"
        demography = msprime.Demography.from_demes(demes.load(self.demo_model_file))
        samples = [msprime.SampleSet(self.nref, ploidy=self.ploidy, population=self.ref_id),
                   msprime.SampleSet(self.ntgt, ploidy=self.ploidy, population=self.test_tgt_id)]
        ts = msprime.sim_ancestry(recombination_rate=self.rec_rate, sequence_length=self.seq_len,
                                  samples=samples, demography=demography, random_seed=seed)
        ts = msprime.sim_mutations(ts, rate=self.mut_rate, random_seed=seed, model=msprime.BinaryMutationModel())
        if self.keep_sim_data:
            os.makedirs(output_dir, exist_ok=True)
            ts.dump(os.path.join(output_dir, f"{output_prefix}.ts"))
            if seed is not None:
                with open(os.path.join(output_dir, f"{output_prefix}.seedmsprime"), "w") as o:
                    o.write(f"{seed}\n")
        return [int(site.position) for site in ts.sites()], ts.genotype_matrix()

    def score_batch(self, replicates, reps):
        """Several replicates in ONE device call -> FeatureTable with a Replicate column."""
        return score_replicate_batch(replicates, reps, self.nref, self.ntgt, self.ploidy, self.seq_len,
                                     self.is_phased, self.match_bonus, self.max_mismatch, self.mismatch_penalty)

    def run(self, rep: int = None, seed: int = None) -> list[dict[str, object]]:
        """The reference's per-task contract [msprime_simulator.py:116-192]."""
        table = self.score_batch([self.simulate_replicate(rep=rep, seed=seed)], [rep])
        return table.rows()
