"""Multi-GPU sharding of one chromosome by window range.

Windows are independent, so there is no collective on the data path: each GPU receives the
variant rows its window range touches (its own step range plus a halo of ``win_len - win_step``
bp), scores them, and the host concatenates the ``[W_g, T]`` blocks in window order -- the
equivalent of the reference's gather + stable sort (sstar/mp_manager.py:172-179,
sstar/preprocess.py:117).
"""
from __future__ import annotations

import os

import numpy as np


def visible_devices(devices=None) -> list[int]:
    """``devices`` argument > ``SSTAR_GPUS`` env ("4" or "0,2,3") > [0]."""
    if devices is None:
        env = os.environ.get("SSTAR_GPUS", "").strip()
        if not env:
            return [0]
        devices = [int(x) for x in env.split(",")] if "," in env else int(env)
    out = list(range(devices)) if isinstance(devices, int) else [int(d) for d in devices]
    if not out:
        raise ValueError("no GPU selected: SSTAR_GPUS / devices must be a positive count (\"4\") or a list of device "
                         "indices (\"0,2,3\")")
    return out


def plan_window_shards(pos, win_start, win_end, n_shards: int):
    """Contiguous window ranges with ~equal (variants-in-window) work.

    -> list of (w0, w1, k0, k1): windows [w0, w1) need variant rows [k0, k1).
    Shards may be fewer than ``n_shards`` when there are few windows; empty shards are dropped.
    """
    pos = np.asarray(pos)
    ws = np.asarray(win_start, dtype=np.int64)
    we = np.asarray(win_end, dtype=np.int64)
    W = ws.shape[0]
    if W == 0:
        return []
    lo = np.searchsorted(pos, ws, side="left")
    hi = np.searchsorted(pos, we, side="right")
    hi = np.maximum(hi, lo)
    work = np.cumsum((hi - lo) + 8)  # +8: fixed cost per window
    n_shards = max(1, min(int(n_shards), W))
    cuts = [0]
    for s in range(1, n_shards):
        cuts.append(int(np.searchsorted(work, work[-1] * s / n_shards, side="left")) + 1)
    cuts.append(W)
    shards = []
    for a, b in zip(cuts[:-1], cuts[1:]):
        a, b = min(a, W), min(b, W)
        if b <= a:
            continue
        k0, k1 = int(lo[a:b].min()), int(hi[a:b].max())
        shards.append((a, b, k0, max(k1, k0)))
    return shards


def score_chromosome(ref_gts, tgt_gts, pos, win_start, win_end, match_bonus, max_mismatch,
                     mismatch_penalty, devices=None, algo="auto"):
    """All windows of one chromosome on one or several GPUs -> (score[W,T], nsnp[W,T])."""
    from .engine import as_int8, check_positions, get_engine
    devs = visible_devices(devices)
    ref = as_int8(ref_gts, "ref_gts")
    tgt = as_int8(tgt_gts, "tgt_gts")
    p = check_positions(pos)
    ws = np.asarray(win_start, dtype=np.int64)
    we = np.asarray(win_end, dtype=np.int64)
    if ws.size and (ws.min() < -(2 ** 31) or we.max() > 2 ** 31 - 1):
        raise ValueError("window coordinates do not fit int32")
    if len(devs) == 1:
        return get_engine(devs[0]).score_host(ref, tgt, p, ws, we, match_bonus, max_mismatch,
                                              mismatch_penalty, algo=algo)
    W, T = ws.shape[0], tgt.shape[1]
    score = np.empty((W, T), dtype=np.int64)
    nsnp = np.empty((W, T), dtype=np.int32)

    # one host thread per GPU: staging a shard into pinned memory, the pipelined device call and the
    # copy of its result block all overlap across devices (torch copies and ctypes calls release the GIL)
    def work(dev, shard):
        w0, w1, k0, k1 = shard
        stream, s_pin, n_pin = get_engine(dev).score_host_async(ref[k0:k1], tgt[k0:k1], p[k0:k1], ws[w0:w1],
                                                                we[w0:w1], match_bonus, max_mismatch,
                                                                mismatch_penalty, algo=algo)
        stream.synchronize()
        score[w0:w1], nsnp[w0:w1] = s_pin.numpy(), n_pin.numpy()  # host gather, window order, one copy

    shards = plan_window_shards(p, ws, we, len(devs))
    futures = [_worker(dev).submit(work, dev, shard) for dev, shard in zip(devs, shards)]
    for f in futures:
        f.result()
    return score, nsnp


def score_chromosome_pinned(ref_p, tgt_p, pos_p, ws_p, we_p, score_p, nsnp_p, match_bonus=5000, max_mismatch=5,
                            mismatch_penalty=-10000, devices=None, algo="auto"):
    """The same sharding on PAGE-LOCKED torch CPU tensors (inputs and the two [W, T] output
    tables): every GPU pulls its row range and stores its result block in place over its own PCIe
    link, with no staging copy on the host -- the route that scales with the number of GPUs.
    Returns when every shard is done."""
    devs = visible_devices(devices)
    pos = pos_p.numpy()
    ws, we = ws_p.numpy(), we_p.numpy()
    lo = np.searchsorted(pos, ws.astype(np.int64), side="left")
    hi = np.searchsorted(pos, we.astype(np.int64), side="right")

    def work(dev, shard):
        w0, w1, k0, k1 = shard
        hint = int(max(0, (hi[w0:w1] - lo[w0:w1]).max())) if w1 > w0 else 0
        eng = get_engine_for(dev)
        with eng.torch.cuda.device(eng.tdev):
            eng.score_pinned(ref_p[k0:k1], tgt_p[k0:k1], pos_p[k0:k1], ws_p[w0:w1], we_p[w0:w1], score_p[w0:w1],
                             nsnp_p[w0:w1], match_bonus, max_mismatch, mismatch_penalty, algo=algo,
                             max_win_variants=hint, sync=True)

    shards = plan_window_shards(pos, ws, we, len(devs))
    futures = [_worker(dev).submit(work, dev, shard) for dev, shard in zip(devs, shards)]
    for f in futures:
        f.result()


def get_engine_for(dev: int):
    from .engine import get_engine
    return get_engine(dev)


_WORKERS: dict = {}


def _worker(dev: int):
    """A persistent single-thread executor per device (the C-ABI keeps per-thread, per-device state)."""
    from concurrent.futures import ThreadPoolExecutor
    ex = _WORKERS.get(dev)
    if ex is None:
        ex = _WORKERS[dev] = ThreadPoolExecutor(max_workers=1, thread_name_prefix=f"sstar-gpu{dev}")
    return ex


# ---------------------------------------------------------------------------------------------
# One process per GPU (torchrun / any launcher): the host gather
# ---------------------------------------------------------------------------------------------
class SharedHostTable:
    """The ``[W, T]`` result tables (int64 scores, int32 SNP counts) of one chromosome in SHARED,
    page-locked host memory: every rank's GPU stores its ``[W_g, T]`` block straight into its window
    range, so the gather of sstar/mp_manager.py:172-179 + the sort of sstar/preprocess.py:117 is
    nothing but the ranks' own device-to-host traffic -- no queue, no pickle, no concatenation.

    Rank 0 creates it (``SharedHostTable(W, T)``) and sends ``table.handle`` to the other ranks by
    whatever channel the launcher offers; they attach with ``SharedHostTable(W, T, handle=...)``.
    The memory is an anonymous ``memfd`` (shmem: pinnable, not bounded by the size of /dev/shm),
    reached by the other processes through ``/proc/<pid>/fd/<fd>``.  ``register()`` page-locks and
    maps it for the calling process' CUDA context, after which the C-ABI host entry writes it in
    place."""

    def __init__(self, W: int, T: int, handle=None, kind: str = "memfd"):
        import mmap
        self.W, self.T = int(W), int(T)
        n = self.W * self.T
        self._off_nsnp = (n * 8 + 4095) // 4096 * 4096
        self.nbytes = max(self._off_nsnp + n * 4, 4096)
        self._unlink = None
        if handle is None:
            if kind == "memfd" and hasattr(os, "memfd_create"):
                self._fd = os.memfd_create("sstar_b200_table")
                self.handle = ("memfd", os.getpid(), self._fd, self.nbytes)
            else:  # a named tmpfs file, for hosts where another process' /proc/<pid>/fd is off limits
                import tempfile
                self._fd, path = tempfile.mkstemp(prefix="sstar_b200_table_", dir="/dev/shm")
                self._unlink = path
                self.handle = ("shm", path, 0, self.nbytes)
            os.ftruncate(self._fd, self.nbytes)
        else:
            kind, a, b, nbytes = handle
            if int(nbytes) != self.nbytes:
                raise ValueError("shared table handle does not match the table shape")
            self._fd = os.open(f"/proc/{int(a)}/fd/{int(b)}" if kind == "memfd" else str(a), os.O_RDWR)
            self.handle = tuple(handle)
        self._mm = mmap.mmap(self._fd, self.nbytes)
        raw = np.frombuffer(self._mm, dtype=np.uint8)
        self.score = raw[:n * 8].view(np.int64).reshape(self.W, self.T)
        self.nsnp = raw[self._off_nsnp:self._off_nsnp + n * 4].view(np.int32).reshape(self.W, self.T)
        self._raw = raw
        self._registered = None

    def register(self, torch):
        """Page-lock + map the table for this process' CUDA context; returns torch views
        ``(score [W,T] int64, nsnp [W,T] int32)`` the host entry point can write in place."""
        t = torch.from_numpy(self._raw)
        if self._registered is None:
            rc = torch.cuda.cudart().cudaHostRegister(t.data_ptr(), self.nbytes, 3)  # portable | mapped
            if int(rc) != 0:
                raise RuntimeError(f"cudaHostRegister of the shared result table failed ({rc})")
            self._registered = (torch, t.data_ptr())
        n = self.W * self.T
        return (t[:n * 8].view(torch.int64).view(self.W, self.T),
                t[self._off_nsnp:self._off_nsnp + n * 4].view(torch.int32).view(self.W, self.T))

    def close(self):
        if self._registered is not None:
            torch, ptr = self._registered
            torch.cuda.cudart().cudaHostUnregister(ptr)
            self._registered = None
        self.score = self.nsnp = self._raw = None
        try:
            self._mm.close()
        except BufferError:  # a caller still holds a view: the mapping goes with the process
            pass
        os.close(self._fd)
        if self._unlink:
            try:
                os.unlink(self._unlink)
            except OSError:
                pass


def plan_replicate_shards(rep_offset, n_shards: int):
    """Contiguous replicate ranges with ~equal variant counts (sstar/simulate.py:147-176 fans the
    replicates of a batch over its workers; here over GPUs) -> list of (r0, r1)."""
    off = np.asarray(rep_offset, dtype=np.int64)
    n = off.shape[0] - 1
    if n <= 0:
        return []
    n_shards = max(1, min(int(n_shards), n))
    work = (off[1:] - off[0]) + 8 * np.arange(1, n + 1)  # +8: fixed cost per replicate
    cuts = [0]
    for s in range(1, n_shards):
        cuts.append(int(np.searchsorted(work, work[-1] * s / n_shards, side="left")) + 1)
    cuts.append(n)
    return [(a, b) for a, b in zip(cuts[:-1], cuts[1:]) if min(b, n) > min(a, n)]


def plan_chromosome_assignment(costs, n_gpus: int):
    """Whole chromosomes -> GPUs, greedily by cost (``W_chr * T``), largest
    first onto the least loaded GPU.  ``costs``: {name: cost}.  -> list (one per GPU) of names, in
    the order that GPU should run them.  The reference loops ``preprocess()`` over chromosomes
    (sstar/preprocess.py:90-112 is per ``chr_name``); independent chromosomes need no exchange."""
    n_gpus = max(1, int(n_gpus))
    load = [0.0] * n_gpus
    plan = [[] for _ in range(n_gpus)]
    for name, c in sorted(costs.items(), key=lambda kv: (-float(kv[1]), str(kv[0]))):
        g = min(range(n_gpus), key=lambda i: (load[i], i))
        plan[g].append(name)
        load[g] += float(c)
    return plan
