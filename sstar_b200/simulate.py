"""``simulate(...)`` with the reference's signature (sstar/simulate.py:30-195): batches of ``nrep``
replicates until ``nfeature`` rows exist, then ``<prefix>.training.features.tsv``.

The reference sends every replicate through ``mp_manager`` as its own task (simulation + scoring
in the worker).  Here the replicates of a batch are simulated on the host (``nprocess`` worker
processes, msprime is CPU code) and scored together in ONE device call; the rows are printed by
the device emitter in the reference's order: sorted by (Replicate, Chromosome, Start, End,
Sample) -- Sample compares as a string -- truncated to ``nfeature`` rows and optionally shuffled
with ``random.seed(seed); random.shuffle(rows)``.
"""
from __future__ import annotations

import multiprocessing
import os
import random
import shutil

import numpy as np

from .feature_table import COLUMNS
from .generators import RandomNumberGenerator
from .msprime_simulator import MsprimeSimulator


def _simulate_one(args):
    simulator, params = args
    return simulator.simulate_replicate(**params)


def _table_lines(table) -> list[bytes]:
    """Rows of one batch as TSV lines, replicates ascending, samples in string order."""
    order = sorted(range(len(table.labels)), key=lambda t: table.labels[t])
    try:
        import torch
        device = 0 if torch.cuda.is_available() else None
    except Exception:  # pragma: no cover
        device = None
    body = table.tsv_body(device=device, col_order=order)
    return body.split(b"\n")[:-1] if body else []


def simulate(demo_model_file: str, nrep: int, nref: int, ntgt: int, ref_id: str, tgt_id: str, ploidy: int,
             seq_len: int, mut_rate: float, rec_rate: float, match_bonus: int, max_mismatch: int,
             mismatch_penalty: int, output_dir: str, output_prefix: str, nfeature: int, is_phased: bool,
             is_shuffled: bool, keep_sim_data: bool, seed: int, nprocess: int) -> None:
    if nfeature <= 0:
        raise ValueError("nfeature must be positive.")
    if output_dir:
        os.makedirs(output_dir, exist_ok=True)
    simulator = MsprimeSimulator(demo_model_file=demo_model_file, nref=nref, ntgt=ntgt, ref_id=ref_id,
                                 tgt_id=tgt_id, ploidy=ploidy, seq_len=seq_len, mut_rate=mut_rate,
                                 rec_rate=rec_rate, output_prefix=output_prefix, output_dir=output_dir,
                                 is_phased=is_phased, match_bonus=match_bonus, max_mismatch=max_mismatch,
                                 mismatch_penalty=mismatch_penalty, keep_sim_data=keep_sim_data)
    lines: list[bytes] = []
    start_rep = 0
    batch_num = 0
    is_stopped = False
    while not is_stopped:
        generator = RandomNumberGenerator(start_rep=start_rep, nrep=nrep, seed=seed)
        tasks = list(generator.get())
        try:
            if nprocess > 1 and len(tasks) > 1:
                with multiprocessing.get_context("fork").Pool(min(nprocess, len(tasks))) as pool:
                    replicates = pool.map(_simulate_one, [(simulator, p) for p in tasks])
            else:
                replicates = [simulator.simulate_replicate(**p) for p in tasks]
            table = simulator.score_batch(replicates, [p["rep"] for p in tasks])
        except (ImportError, ValueError):
            raise
        except Exception as e:
            print(f"sstar_b200: simulation batch failed ({e})")
            raise SystemExit("Some errors occurred, stopping the program ...")
        if not keep_sim_data:
            for i in range(start_rep, start_rep + nrep):
                shutil.rmtree(os.path.join(output_dir, f"{i}"), ignore_errors=True)
        lines.extend(_table_lines(table))
        is_stopped = len(lines) >= nfeature
        print(f"Number of obtained features: {len(lines)} in batch {batch_num}")
        start_rep += nrep
        batch_num += 1

    lines = lines[:nfeature]
    if is_shuffled:
        random.seed(seed)
        random.shuffle(lines)  # the same swaps the reference applies to its list of row dicts
    output_file = os.path.join(output_dir, f"{output_prefix}.training.features.tsv")
    with open(output_file, "wb") as fh:
        fh.write(("\t".join(COLUMNS + ["Replicate"]) + "\n").encode())
        if lines:
            fh.write(b"\n".join(lines) + b"\n")
