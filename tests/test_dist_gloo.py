"""world_size-2 gloo test (CPU) of the N>1 host logic: window-range sharding, per-rank scoring of
its own variant rows (the oracle stands in for the device here), host gather in window order, and
the max-over-ranks / sum-of-units reduction bench.py uses."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import c_oracle
    from sstar_b200.sharding import plan_window_shards
    from sstar_b200.synth import regular_windows, synth_chromosome
    ref, tgt, pos = synth_chromosome(900_000, 8, 6, False, seed=11)
    ws, we = regular_windows(pos, 50_000, 10_000)
    shards = plan_window_shards(pos, ws, we, world)
    w0, w1, k0, k1 = shards[rank]
    s, n = c_oracle.score_windows(ref[k0:k1], tgt[k0:k1], pos[k0:k1], ws[w0:w1], we[w0:w1])
    gathered = [None] * world
    dist.all_gather_object(gathered, (w0, w1, s, n))
    # timing reduction as in bench.py: MAX over ranks of the step time, SUM of the units
    t = torch.tensor([10.0 + rank, float(s.size)], dtype=torch.float64)
    tmax, usum = t[:1].clone(), t[1:].clone()
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    dist.all_reduce(usum, op=dist.ReduceOp.SUM)
    if rank == 0:
        score = np.concatenate([g[2] for g in sorted(gathered, key=lambda g: g[0])])
        nsnp = np.concatenate([g[3] for g in sorted(gathered, key=lambda g: g[0])])
        so, no = c_oracle.score_windows(ref, tgt, pos, ws, we)
        ok = np.array_equal(score, so) and np.array_equal(nsnp, no)
        ok = ok and float(tmax) == 10.0 + world - 1 and int(usum) == so.size
        with open(out_path, "w") as fh:
            fh.write("ok" if ok else "mismatch")
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_and_gather(tmp_path):
    out = tmp_path / "result.txt"
    mp.spawn(_worker, args=(2, _free_port(), str(out)), nprocs=2, join=True)
    assert out.read_text() == "ok"


def _table_worker(rank, world, port, out_path):
    """The bench's gather: every rank stores its block into ONE shared host table (no message)."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import c_oracle
    from sstar_b200.sharding import SharedHostTable, plan_window_shards
    from sstar_b200.synth import regular_windows, synth_chromosome
    ref, tgt, pos = synth_chromosome(700_000, 8, 5, False, seed=12)
    ws, we = regular_windows(pos, 50_000, 10_000)
    W, T = len(ws), tgt.shape[1]
    table = SharedHostTable(W, T) if rank == 0 else None
    box = [table.handle if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    if rank != 0:
        table = SharedHostTable(W, T, handle=box[0])
    w0, w1, k0, k1 = plan_window_shards(pos, ws, we, world)[rank]
    s, n = c_oracle.score_windows(ref[k0:k1], tgt[k0:k1], pos[k0:k1], ws[w0:w1], we[w0:w1])
    table.score[w0:w1], table.nsnp[w0:w1] = s, n
    dist.barrier()
    if rank == 0:
        so, no = c_oracle.score_windows(ref, tgt, pos, ws, we)
        ok = np.array_equal(table.score, so) and np.array_equal(table.nsnp, no)
        with open(out_path, "w") as fh:
            fh.write("ok" if ok else "mismatch")
    dist.barrier()
    table.close()
    dist.destroy_process_group()


def test_shared_host_table_gathers_without_messages(tmp_path):
    out = tmp_path / "result.txt"
    mp.spawn(_table_worker, args=(2, _free_port(), str(out)), nprocs=2, join=True)
    assert out.read_text() == "ok"


def test_shared_host_table_named_file_fallback():
    from sstar_b200.sharding import SharedHostTable
    a = SharedHostTable(5, 3, kind="shm")
    b = SharedHostTable(5, 3, handle=a.handle)
    a.score[:] = 11
    b.nsnp[:] = 4
    assert (b.score == 11).all() and (a.nsnp == 4).all()
    path = a.handle[1]
    b.close()
    a.close()
    assert not os.path.exists(path)
    with pytest.raises(ValueError):
        SharedHostTable(6, 3, handle=("shm", path, 0, 1))
