"""Replicate-batched scoring for the simulate / train path.

The reference scores each msprime replicate separately inside a worker
(``MsprimeSimulator.run`` -> ``_build_feature_input_from_ts`` -> ``FeatureVectorPreprocessor.run``,
sstar/msprime_simulator.py:179-256) and gathers ``nrep`` replicates per batch through
``mp_manager`` (sstar/simulate.py:147-176).  Here a batch of replicates is concatenated along the
variant axis and scored in ONE device call (``sstar_b200_score_replicates``).

msprime / tskit are not dependencies: a replicate is ``(positions, genotype_matrix)`` exactly as
``[int(site.position) for site in ts.sites()]`` and ``ts.genotype_matrix()`` give them.
"""
from __future__ import annotations

import numpy as np

from .feature_table import FeatureTable, sample_labels


def build_feature_input(pos, gts, nref: int, ntgt: int, ploidy: int, seq_len: int, is_phased: bool):
    """One replicate -> (ref_gts, tgt_gts, pos, start, end), as msprime_simulator.py:194-256:
    drop sites fixed-alt in both populations, keep the single window ``split_genome(pos, "1",
    seq_len, seq_len)[0]`` (= [1, seq_len] -- a site at position 0 falls outside), and sum
    haplotype pairs when unphased."""
    pos = np.asarray(pos)
    gts = np.asarray(gts)
    nref_hap = nref * ploidy
    ref_h, tgt_h = gts[:, :nref_hap], gts[:, nref_hap:]
    fixed = (ref_h.sum(axis=1) == ref_h.shape[1]) & (tgt_h.sum(axis=1) == tgt_h.shape[1])
    keep = ~fixed
    pos, ref_h, tgt_h = pos[keep], ref_h[keep], tgt_h[keep]
    if len(pos) == 0:
        raise ValueError("`pos` array must not be empty.")
    start = max(1, (int(pos[0]) + seq_len) // seq_len * seq_len - seq_len + 1)
    end = start + seq_len - 1
    inside = (pos >= start) & (pos <= end)
    pos, ref_h, tgt_h = pos[inside], ref_h[inside], tgt_h[inside]
    if is_phased:
        return ref_h.astype(np.int8), tgt_h.astype(np.int8), pos, start, end
    ref = ref_h.reshape(len(pos), nref, ploidy).sum(axis=2).astype(np.int8)
    tgt = tgt_h.reshape(len(pos), ntgt, ploidy).sum(axis=2).astype(np.int8)
    return ref, tgt, pos, start, end


def sample_names(nref: int, ntgt: int, identifier: str = "tsk_"):
    """msprime_simulator.py:258-277."""
    return ([f"{identifier}{i}" for i in range(nref)],
            [f"{identifier}{i}" for i in range(nref, nref + ntgt)])


def score_replicate_batch(replicates, reps, nref: int, ntgt: int, ploidy: int, seq_len: int,
                          is_phased: bool, match_bonus: int, max_mismatch: int,
                          mismatch_penalty: int, device: int = None, algo: str = "auto",
                          devices=None) -> FeatureTable:
    """``replicates``: list of ``(positions, genotype_matrix)``; ``reps``: their replicate numbers.
    Returns a FeatureTable with one "window" per replicate and a ``Replicate`` column.

    The reference fans the replicates of a batch over its worker processes
    (sstar/simulate.py:147-176); here they go over the GPUs ``SSTAR_GPUS`` names (or ``devices``)
    as contiguous replicate ranges of equal variant work, one device call and one host thread per
    GPU, results concatenated in replicate order.  ``device=k`` pins the batch to one GPU."""
    from .engine import get_engine
    from .sharding import _worker, plan_replicate_shards, visible_devices
    refs, tgts, poss, starts, ends, off = [], [], [], [], [], [0]
    for pos, gts in replicates:
        r, t, p, s, e = build_feature_input(pos, gts, nref, ntgt, ploidy, seq_len, is_phased)
        refs.append(r); tgts.append(t); poss.append(p.astype(np.int32)); starts.append(s); ends.append(e)
        off.append(off[-1] + len(p))
    if len(set(starts)) != 1 or len(set(ends)) != 1:
        # windows differ between replicates (first site beyond seq_len): fall back to per-window bounds
        raise ValueError("replicates do not share the window [1, seq_len]")
    devs = [int(device)] if device is not None else visible_devices(devices)
    ref_all, tgt_all, pos_all = np.concatenate(refs), np.concatenate(tgts), np.concatenate(poss)
    off = np.asarray(off, dtype=np.int64)
    shards = plan_replicate_shards(off, len(devs)) if len(devs) > 1 else [(0, len(replicates))]

    def work(dev, r0, r1):
        a, b = int(off[r0]), int(off[r1])
        return get_engine(dev).score_replicates_host(ref_all[a:b], tgt_all[a:b], pos_all[a:b],
                                                     (off[r0:r1 + 1] - off[r0]).astype(np.int32), starts[0], ends[0],
                                                     match_bonus, max_mismatch, mismatch_penalty, algo=algo)

    if len(shards) == 1:
        score, nsnp = work(devs[0], *shards[0])
    else:
        parts = [f.result() for f in [_worker(dev).submit(work, dev, r0, r1) for dev, (r0, r1) in zip(devs, shards)]]
        score = np.concatenate([p[0] for p in parts])
        nsnp = np.concatenate([p[1] for p in parts])
    _, tgt_names = sample_names(nref, ntgt)
    labels = sample_labels(tgt_names, ploidy, is_phased)
    n = len(replicates)
    return FeatureTable("1", [starts[0]] * n, [ends[0]] * n, labels, score, nsnp, replicate=list(reps))
