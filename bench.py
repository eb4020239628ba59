#!/usr/bin/env python
"""bench.py -- S* (window x individual) scores/sec on B200, with the reference's CPU path beside it.

    python bench.py --gpus N --steps K --warmup W              # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...    # the reference's own CPU code

Workload (BASELINE.json configs[1], "C2"): a synthetic msprime-style chromosome of N x 10 Mb (N = the
number of GPUs; N = 1 is exactly configs[1]), 100 reference + 50 target diploids, unphased dosages,
50 kb windows / 10 kb step.  One "step" is one pass of the hot path over the chromosome: pack ->
window bounds -> chain DP for every (window, target column) unit.

With N > 1 (torchrun, one rank per GPU) the chromosome goes through the product's partition
(``sstar_b200.sharding``): ``plan_window_shards`` cuts contiguous window ranges, every rank owns
the variant rows its range touches (own step range + the ``win_len - win_step`` halo) in its own
page-locked allocation, scores them on its GPU, and stores its ``[W_g, T]`` block into ONE result
table in shared page-locked host memory (``SharedHostTable``) -- the reference's gather + sort
(sstar/mp_manager.py:172-179, sstar/preprocess.py:117).  No collective on the data path: NCCL only
carries the barriers and the max / sum of the timings.  Rank 0 then checks the gathered table against
the oracle, bit for bit, on every window (C2 sizes) -- in particular those straddling the shard cuts.

value  : units/s with every rank's rows resident in HBM (CUDA events around each step, max over ranks).
e2e    : units/s through ScoreEngine.score_pinned (one C-ABI call per rank: its rows from pinned host
         memory, kernels, its block of the shared host table written), wall clock, max over ranks.
strong : the product's partitions at FIXED total work -- BASELINE configs[2] (one 250 Mb chromosome,
         R = T = 1000 haplotype columns, 24.3 M units, by window range), configs[4] (10,000 replicates
         of 50 kb, by replicate range) and configs[3] (22 autosomes, T = 2,000, whole chromosomes by
         cost; device-resident only) over the N ranks, oracle-checked in the run.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "S* (window x individual) scores/sec"
UNIT = "scores/s"
PARAMS = (5000, 5, -10000)
SEG_LEN = 10_000_000
WORKLOADS = {
    # name: (length bp per GPU, n_ref, n_tgt, phased, win_len, win_step, description)
    "c2": (SEG_LEN, 100, 50, False, 50_000, 10_000,
           "C2: synthetic 10 Mb msprime-style chromosome, 100 ref + 50 target diploids (unphased), "
           "50 kb windows / 10 kb step"),
    "c2-small": (1_000_000, 100, 50, False, 50_000, 10_000, "C2 shape at 1 Mb (debug)"),
    "c2-x4": (40_000_000, 100, 50, False, 50_000, 10_000, "C2 shape at 40 Mb: 20 MB of genotypes (host-entry threshold check)"),
    # large shapes (generated on the device; not the headline line -- roofline evidence)
    "c3-shard": (31_250_000, 500, 500, True, 50_000, 10_000,
                 "C3 shard: 1/8 of a 250 Mb chromosome, 500 ref + 500 target diploids, haplotype mode "
                 "(R = T = 1000 columns), 50 kb / 10 kb"),
    "c4-shard": (30_000_000, 100, 2000, False, 50_000, 10_000,
                 "C4 shard: 30 Mb, 100 ref + 2000 target diploids (unphased), 50 kb / 10 kb"),
    # wide panels (debug): one 32-row tile is 163 KB (one pack CTA per SM) / does not fit at all (byte path)
    "wide5k": (30_000_000, 50, 2500, True, 50_000, 10_000, "30 Mb, 50 ref + 2500 target diploids, haplotype mode (T = 5000)"),
    "wide8k": (30_000_000, 50, 4000, True, 50_000, 10_000, "30 Mb, 50 ref + 4000 target diploids, haplotype mode (T = 8000)"),
}
LARGE = {"c3-shard", "c4-shard", "wide5k", "wide8k"}
C3 = dict(length=250_000_000, n_ref=500, n_tgt=500, seed=12347,
          desc="C3 (BASELINE configs[2]): one synthetic 250 Mb chromosome, 500 ref + 500 target diploids, haplotype "
               "mode (R = T = 1000 columns), 50 kb / 10 kb, sharded by window range over the ranks")
# GRCh38 autosome lengths (Mb), chr1 .. chr22: 2.88 Gb
C4_MB = [248.96, 242.19, 198.30, 190.21, 181.54, 170.81, 159.35, 145.14, 138.39, 133.80, 135.09, 133.28, 114.36, 107.04, 101.99,
         90.34, 83.26, 80.37, 58.62, 64.44, 46.71, 50.82]
C4 = dict(n_ref=100, n_tgt=2000, seed=12349,
          desc="C4 (BASELINE configs[3]): synthetic 22-autosome genome (GRCh38 lengths, 2.88 Gb), 100 ref + 2,000 target "
               "diploids (unphased; R is not given by BASELINE: 100 as in configs[1]), 50 kb / 10 kb, whole chromosomes assigned "
               "to the ranks greedily by length (sstar_b200.sharding.plan_chromosome_assignment)")
C5 = dict(n_rep=10_000, seq_len=50_000, n_ref=5, n_tgt=10, seed=5,
          desc="C5 (BASELINE configs[4]): 10,000 synthetic 50 kb replicates, 5 ref + 10 target diploids (unphased, "
               "the tutorial's shape), one window [1, 50000] each, contiguous replicate ranges over the ranks")


# ------------------------------------------------------------------------------------------
# workloads
# ------------------------------------------------------------------------------------------
def make_workload(name, world=1):
    """The FULL chromosome of a run on ``world`` ranks (every rank builds it from the same seeds):
    ``world`` independent segments of the named shape laid end to end.  world = 1 is the shape itself."""
    from sstar_b200.synth import regular_windows, synth_chromosome, synth_chromosome_device
    length, n_ref, n_tgt, phased, wl, wstep, desc = WORKLOADS[name]
    seeds = [12345 + 1 + 1000 * r for r in range(world)]
    if name in LARGE:
        import torch
        if world != 1:
            raise SystemExit(f"workload {name} is a single-GPU shape (use the default workload with --gpus N)")
        dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
        r, t, p = synth_chromosome_device(torch, dev, length, n_ref, n_tgt, phased, seeds[0])
        ref, tgt, pos = r.cpu().numpy(), t.cpu().numpy(), p.cpu().numpy()
        del r, t, p
        torch.cuda.empty_cache()
    else:
        segs = [synth_chromosome(length, n_ref, n_tgt, phased, s) for s in seeds]
        ref = np.concatenate([s[0] for s in segs])
        tgt = np.concatenate([s[1] for s in segs])
        pos = np.concatenate([s[2].astype(np.int64) + i * length for i, s in enumerate(segs)]).astype(np.int32)
    ws, we = regular_windows(pos, wl, wstep)
    if world > 1:
        desc = f"{world} x 10 Mb = one {world * length // 1_000_000} Mb chromosome of the C2 shape ({desc}), sharded by " \
               f"window range over {world} GPUs"
    return dict(ref=ref, tgt=tgt, pos=pos, ws=ws, we=we, seed=seeds[0], seeds=seeds, desc=desc, name=name,
                phased=phased, length=length * world)


def workload_stats(wk):
    """Shape + work statistics; n (carried ref-absent sites per unit) is exact for T <= 64 columns
    and estimated from a 64-column sample beyond that."""
    pos, ws, we = wk["pos"], wk["ws"], wk["we"]
    absent = wk["ref"].sum(axis=1, dtype=np.int32) == 0
    lo = np.searchsorted(pos, ws, side="left")
    hi = np.searchsorted(pos, we, side="right")
    T = wk["tgt"].shape[1]
    cols = np.arange(T) if T <= 64 else np.unique(np.linspace(0, T - 1, 64).astype(int))
    car = ((wk["tgt"][:, cols] != 0) & absent[:, None]).astype(np.int64)
    csum = np.concatenate([np.zeros((1, car.shape[1]), np.int64), np.cumsum(car, axis=0)])
    n = csum[hi] - csum[lo]
    scale = T / len(cols)
    return dict(V=int(pos.size), R=int(wk["ref"].shape[1]), T=int(T), W=int(ws.size),
                units=int(ws.size * T), sum_n=int(n.sum() * scale),
                sum_pairs=int((n * (n - 1) // 2).sum() * scale), max_n=int(n.max()),
                n_exact=bool(T <= 64), absent_fraction=float(absent.mean()),
                mean_window_variants=float((hi - lo).mean()), max_window_variants=int((hi - lo).max()))


def shared_config(wk, st, world):
    """The ``config`` object -- the SAME dict in both arms (ours / --impl reference)."""
    return {"workload": wk["desc"], **st, "seed": wk["seed"], "params": list(PARAMS), "n_gpus": world,
            "partition": "one rank: the whole chromosome" if world == 1 else
                         f"sstar_b200.sharding.plan_window_shards: {world} contiguous window ranges, each rank holds its "
                         "rows + the win_len - win_step halo; result blocks gathered in one shared host table",
            "l2": "GPU arm, value: inputs larger than the L2 -- consecutive steps read DISTINCT device copies of the rows "
                  "(a ring of copies of 256 MB or more together; no flush needed); single_pass, e2e and the large "
                  "shapes (at_scale, strong): L2 flushed between timed iterations (256 MiB device write); CPU arm: "
                  "not applicable",
            "timing": "GPU arm, value: exactly K steps between two CUDA events on the launching stream, barrier + device "
                      "synchronize on both sides, max over ranks; the steps are independent calls "
                      "(SSTAR_B200_FLAG_INDEPENDENT) replayed from one CUDA graph per ring, so consecutive one-launch "
                      "kernels overlap; e2e: wall clock from the first of K host-buffer calls to ONE stream synchronize "
                      "after the last (consecutive calls overlap transfer and scoring), max over ranks; both K-step "
                      "regions are run three times, each behind untimed warm-up steps of the same kind, and the median "
                      "is reported (all three listed); before the e2e regions the CPU's caches are flushed (1 GiB "
                      "written elsewhere), so the page-locked rows come from DRAM; CPU arm: wall clock"}


def algorithmic_bytes(st):
    """every int8 genotype read once + pos + 12 B per unit written."""
    pack = st["V"] * (st["R"] + st["T"])
    return dict(pipeline=pack + 4 * st["V"] + 12 * st["units"], pack=pack,
                score=4 * st["V"] + 12 * st["units"])


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def recorded_traffic(workload, kernel, key="dram_bytes"):
    """Per-launch counters from the committed ncu --set full capture (profiles/traffic.json), if any:
    ``dram_bytes`` = dram__bytes_read.sum + dram__bytes_write.sum, ``warp_inst`` = smsp__inst_executed.sum.
    (ncu cannot run inside the timed program; the file names the capture each number comes from.)"""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
            rec = json.load(fh).get(workload, {}).get(kernel)
        if isinstance(rec, dict):
            return rec.get(key)
        return rec if key == "dram_bytes" else None
    except Exception:
        return None


def issue_peak_ginst():
    """Warp instructions per second the SMs can issue: SMs x 4 schedulers x max SM clock."""
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            mhz = float(json.load(fh)["sm_max_mhz"])
    except Exception:
        mhz = 1965.0
    try:
        import torch
        sms = torch.cuda.get_device_properties(0).multi_processor_count
    except Exception:
        sms = 148
    return sms * 4 * mhz * 1e6 / 1e9, sms, mhz


class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons through NVML while the timed regions run."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown",
               0x4: "sw_power_cap", 0x80: "hw_power_brake_slowdown"}

    def __init__(self, index, period=float(os.environ.get("SSTAR_BENCH_CLOCK_PERIOD", "0.01"))):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.marks = [], []
        self.stop_flag = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.max_mhz = None

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        while not self.stop_flag.is_set():
            try:
                mhz = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                try:
                    rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    rs = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.samples.append((time.perf_counter(), mhz, rs))
            except Exception:
                pass
            time.sleep(self.period)

    def mark(self):
        self.marks.append(time.perf_counter())

    def summary(self):
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        sel = self.samples
        if len(self.marks) >= 2:
            inside = [s for s in self.samples if self.marks[0] <= s[0] <= self.marks[-1]]
            sel = inside or self.samples
        bits = 0
        for _, _, rs in sel:
            bits |= rs
        return {"sm_mhz": statistics.median(m for _, m, _ in sel), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(v for k, v in self.REASONS.items() if bits & k), "samples": len(sel)}


def dist_setup():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return world, rank, local


# ------------------------------------------------------------------------------------------
# the reference arm: the reference's own CPU code on the box's host cores
# ------------------------------------------------------------------------------------------
def cpu_arm(wk, st, cores, passes_cap=10, check=None):
    """Times the reference's job (``FeatureVectorPreprocessor.run`` -> ``Sstar.compute`` from
    baseline/_ref when installed, else the oracle's numpy port) over a pre-forked pool of ``cores``
    workers on a bounded, evenly spread window sample.  -> (units/s, description dict)."""
    from oracle.cpu_baseline import CpuBaseline
    W, T = st["W"], st["T"]
    base = CpuBaseline(wk["ref"], wk["tgt"], wk["pos"], wk["ws"], wk["we"], PARAMS, cores, phased=wk["phased"])
    per_win = base.per_window_seconds()
    n_win = int(min(W, max(cores, 20.0 / max(per_win, 1e-6))))  # ~20 core-seconds per pass at most
    sample = np.linspace(0, W - 1, num=n_win).astype(int)
    base.run(sample[: max(cores, n_win // 8)])
    passes = int(min(passes_cap, max(1, round(12.0 / max(per_win * n_win, 1e-6)))))  # ~12 core-seconds of work
    dts, res = [], {}
    for _ in range(passes):
        dt, res = base.run(sample)
        dts.append(dt)
    base.close()
    if check is not None:
        check(res)
    dt = statistics.mean(dts)
    info = {"value": n_win * T / dt, "unit": UNIT, "cores": cores, "kind": base.kind,
            "sample": f"{n_win} of {W} windows x {T} columns of the {wk['name']} chromosome, mean of {passes} passes "
                      f"({dt:.2f} s wall each); {base.source}; one task per window as mp_manager does, but the pool is "
                      "pre-forked and warm (no Manager-proxy pickling, no 5 s worker-exit tail)",
            "in_process_us_per_unit": 1e6 * per_win / T, "seconds_per_pass": dt, "passes": passes}
    return info


def run_reference(args, world, rank):
    """The reference arm: rank 0 alone runs and prints; the other ranks exit without work."""
    if rank != 0:
        return
    wk = make_workload(args.workload, world)
    st = workload_stats(wk)
    cores = os.cpu_count() or 1
    from oracle.cpu_baseline import CpuBaseline
    base = CpuBaseline(wk["ref"], wk["tgt"], wk["pos"], wk["ws"], wk["we"], PARAMS, cores, phased=wk["phased"])
    per_win = base.per_window_seconds()
    # bounded sample: ~20 s of CPU work per step at most, the whole run near a minute
    work = min(20.0, 60.0 * cores / max(args.steps + args.warmup, 1))
    n_win = int(min(st["W"], max(cores, work / max(per_win, 1e-6))))
    sample = np.linspace(0, st["W"] - 1, num=n_win).astype(int)
    for _ in range(args.warmup):
        base.run(sample[: max(cores, len(sample) // 8)])
    secs = []
    for _ in range(args.steps):
        dt, _ = base.run(sample)
        secs.append(dt)
    base.close()
    units = n_win * st["T"]
    value = units / statistics.mean(secs)
    sample_desc = f"{n_win} of {st['W']} windows x {st['T']} columns of the {wk['name']} chromosome per step"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * statistics.mean(secs),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": shared_config(wk, st, world),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": base.kind,
                         "sample": sample_desc + f"; {base.source}; pool pre-forked and warm, no mp_manager 5 s tail"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------
# one rank's shard on its GPU
# ------------------------------------------------------------------------------------------
class ShardCase:
    """The rows / windows one rank owns: page-locked host copies (its own allocation), device copies,
    the captured pass and the timing helpers.  ``out_p`` are this rank's blocks of the host table."""

    def __init__(self, torch, eng, dev, arrays, out_p=None, use_graph=True, flags=0):
        self.torch, self.eng, self.dev, self.flags = torch, eng, dev, flags
        pos, ws, we = arrays["pos"], arrays["ws"], arrays["we"]
        lo = np.searchsorted(pos, ws.astype(np.int64), side="left")
        hi = np.searchsorted(pos, we.astype(np.int64), side="right")
        self.hint = int(max(0, (hi - lo).max())) if len(ws) else 0
        self.V, self.R, self.T, self.W = len(pos), arrays["ref"].shape[1], arrays["tgt"].shape[1], len(ws)
        W, T = self.W, self.T
        self.pin = {k: torch.from_numpy(np.ascontiguousarray(arrays[k])).pin_memory()
                    for k in ("ref", "tgt", "pos", "ws", "we")}
        self.d = {k: v.to(dev) for k, v in self.pin.items()}
        self.score_d = torch.empty((W, T), dtype=torch.int64, device=dev)
        self.nsnp_d = torch.empty((W, T), dtype=torch.int32, device=dev)
        if out_p is None:
            out_p = (torch.empty((W, T), dtype=torch.int64).pin_memory(),
                     torch.empty((W, T), dtype=torch.int32).pin_memory())
        self.score_p, self.nsnp_p = out_p
        # the whole pass (pack -> bounds -> DP) captured once: ONE host call per step
        self.plan = None
        if use_graph:
            d = self.d
            self.plan = eng.plan_device(d["ref"], d["tgt"], d["pos"], d["ws"], d["we"], *PARAMS,
                                        max_win_variants=self.hint, out=(self.score_d, self.nsnp_d), flags=flags)

    def step_device(self):
        if self.plan is not None:
            self.plan.run()
            return self.plan.launches
        d = self.d
        self.eng.score_device(d["ref"], d["tgt"], d["pos"], d["ws"], d["we"], *PARAMS,
                              max_win_variants=self.hint, out=(self.score_d, self.nsnp_d), flags=self.flags)
        return self.eng.launches

    def step_pack(self):
        self.eng.pack_device(self.d["ref"], self.d["tgt"], W=self.W, max_win_variants=self.hint)

    def step_score(self):
        d = self.d
        self.eng.score_packed_device(self.R, d["tgt"], d["pos"], d["ws"], d["we"], *PARAMS,
                                     max_win_variants=self.hint, out=(self.score_d, self.nsnp_d))

    def step_e2e(self, staged=False):
        p = self.pin
        self.eng.score_pinned(p["ref"], p["tgt"], p["pos"], p["ws"], p["we"], self.score_p, self.nsnp_p,
                              *PARAMS, max_win_variants=self.hint, sync=True, staged=staged)

    def make_stream(self, min_bytes=256 << 20, max_ring=64):
        """A ring of DISTINCT device copies of this rank's rows (together larger than the L2) and ONE graph
        that scores them one after the other (``ScoreEngine.plan_device_stream``: the calls carry
        SSTAR_B200_FLAG_INDEPENDENT, so consecutive one-launch kernels overlap).  One step = one pass over
        one copy; consecutive steps read different copies, so every step's rows come from HBM."""
        torch = self.torch
        nbytes = self.V * (self.R + self.T + 4) + 8 * self.W
        self.ring = int(min(max_ring, max(2, -(-min_bytes // max(nbytes, 1)))))
        self.ring_bytes = self.ring * nbytes
        self.ring_cases = []
        for _ in range(self.ring):
            out = (torch.empty_like(self.score_d), torch.empty_like(self.nsnp_d))
            self.ring_cases.append(tuple(self.d[k].clone() for k in ("ref", "tgt", "pos", "ws", "we")) + (out,))
        self._stream_plans = {}

    def _stream_plan(self, n):
        if n not in self._stream_plans:
            self._stream_plans[n] = self.eng.plan_device_stream(self.ring_cases[:n], *PARAMS, max_win_variants=self.hint,
                                                                flags=self.flags)
        return self._stream_plans[n]

    def run_stream(self, steps):
        """Enqueue exactly ``steps`` passes (whole rings, then the remainder); returns the launches enqueued."""
        full, rest = divmod(steps, self.ring)
        n = 0
        if full:
            plan = self._stream_plan(self.ring)
            for _ in range(full):
                plan.run()
            n += full * plan.launches
        if rest:
            plan = self._stream_plan(rest)
            plan.run()
            n += plan.launches
        return n

    def timed_stream(self, steps, barrier, warm=0):
        """EXACTLY ``steps`` passes between two CUDA events on the launching stream, a barrier + device
        synchronize on both sides.  No flush: the ring of inputs is larger than the L2.  ``warm`` untimed
        passes of the same stream go first (the region then starts with this kernel's code and the graph's
        launch state in place, as the warm-up steps are meant to leave them)."""
        torch = self.torch
        self._stream_plan(self.ring)
        if steps % self.ring:
            self._stream_plan(steps % self.ring)
        if warm > 0:
            self.run_stream(warm)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        n = self.run_stream(steps)
        e1.record()
        barrier()
        return e0.elapsed_time(e1), n

    def stream_tables_equal(self, used):
        """Every ring copy scored in the timed region holds the tables of ``step_device`` (same rows)."""
        t = self.torch
        return all(bool(t.equal(c[5][0], self.score_d) and t.equal(c[5][1], self.nsnp_d)) for c in self.ring_cases[:used])

    def make_e2e_stream(self, slots=4):
        """A ring of page-locked copies of this rank's rows, each with its own device scratch and its own
        page-locked result tables (slot 0 writes this rank's block of the shared host table)."""
        torch = self.torch
        self.e2e_slots = [(self.pin, (self.score_p, self.nsnp_p))]
        for _ in range(slots - 1):
            pin = {k: v.clone().pin_memory() for k, v in self.pin.items()}
            self.e2e_slots.append((pin, (torch.zeros_like(self.score_p).pin_memory(), torch.zeros_like(self.nsnp_p).pin_memory())))

    def run_e2e_stream(self, steps):
        """``steps`` host-entry calls back to back (SSTAR_B200_FLAG_INDEPENDENT: the copies of call k+1 run
        under the scoring of call k), ONE stream synchronize at the end."""
        from sstar_b200 import _cabi
        n = len(self.e2e_slots)
        for i in range(steps):
            p, (sp, npn) = self.e2e_slots[i % n]
            self.eng.score_pinned(p["ref"], p["tgt"], p["pos"], p["ws"], p["we"], sp, npn, *PARAMS,
                                  max_win_variants=self.hint, sync=False, flags=self.flags | _cabi.FLAG_INDEPENDENT,
                                  scratch_slot=1 + i % n)
        self.torch.cuda.synchronize(self.dev)

    def wall_e2e_stream(self, steps, barrier):
        barrier()
        t0 = time.perf_counter()
        self.run_e2e_stream(steps)
        return time.perf_counter() - t0

    def e2e_slots_equal(self):
        t = self.torch
        return all(bool(t.equal(o[0], self.score_p) and t.equal(o[1], self.nsnp_p)) for _p, o in self.e2e_slots[1:])

    def timed(self, fn, steps, flush, barrier):
        """K steps, L2 flushed before each, CUDA events around each step on the launching stream."""
        torch = self.torch
        evs = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(steps)]
        barrier()
        for i in range(steps):
            flush.fill_(i & 1)
            evs[i][0].record()
            fn()
            evs[i][1].record()
        barrier()
        return sum(e[0].elapsed_time(e[1]) for e in evs)

    def wall(self, fn, steps, flush):
        torch = self.torch
        tot = 0.0
        for i in range(steps):
            flush.fill_(i & 1)
            torch.cuda.synchronize(self.dev)
            t0 = time.perf_counter()
            fn()
            tot += time.perf_counter() - t0
        return tot

    def device_equals_host_block(self):
        """The device-resident tables of the last ``step_device`` == this rank's block of the host
        table written by the last ``step_e2e`` (bit for bit)."""
        return bool(self.torch.equal(self.score_d.cpu(), self.score_p) and
                    self.torch.equal(self.nsnp_d.cpu(), self.nsnp_p))


_HOST_JUNK = {}


def host_cache_flush(torch, nbytes=1 << 30):
    """Put the page-locked host buffers where a pipeline's inputs are -- in DRAM, not in the CPU's caches.
    The bench itself has just WRITTEN them (clones of the rows, table resets), and on this pool's hosts DMA
    reads of lines that sit modified in the CPU's caches run at 13-40 GB/s instead of 50, for as long as
    nothing evicts them (``tools/exp_pinned_slots.py``: 16 slices of one page-locked slab read at 49.8 ...
    13.6 GB/s in the order they were written, all at 49 GB/s once the CPU has written 1 GiB elsewhere).  The
    host-side twin of the L2 flush: outside every timed region, before the e2e ones."""
    if "buf" not in _HOST_JUNK:
        while True:  # (a host short of memory gets a smaller sweep rather than a failed run)
            try:
                _HOST_JUNK["buf"], _HOST_JUNK["n"] = torch.empty(nbytes, dtype=torch.uint8), 0
                break
            except (RuntimeError, MemoryError):
                if nbytes <= (64 << 20):
                    raise
                nbytes //= 2
    _HOST_JUNK["n"] += 1
    _HOST_JUNK["buf"].fill_(_HOST_JUNK["n"] & 1)


def oracle_check(wk, score, nsnp, sel, what):
    """Rows ``sel`` of the gathered host table vs the scalar C oracle, bit-exact, or refuse to report."""
    from oracle import c_oracle
    sel = np.asarray(sel, dtype=np.int64)
    so, no = c_oracle.score_windows(wk["ref"], wk["tgt"], wk["pos"], wk["ws"][sel], wk["we"][sel], *PARAMS)
    ok = bool(np.array_equal(score[sel], so) and np.array_equal(nsnp[sel], no))
    if not ok:
        raise SystemExit(f"bench: GPU result differs from the oracle ({what}) -- refusing to report a number")
    return int(so.size)


def cut_windows(shards, W, radius=6):
    """Windows within ``radius`` of every shard cut (their variants straddle two ranks' row ranges)."""
    sel = set()
    for (w0, _w1, _k0, _k1) in shards[1:]:
        sel.update(range(max(0, w0 - radius), min(W, w0 + radius)))
    return sorted(sel)


def roofline_block(name, st, k, tot_ms, pack_ms, score_ms, launches, stream=None, world=1):
    """Pass-level and per-kernel HBM roofline (SURVEY.md section 8d bytes; measured peak).  PER GPU: a launch
    processes one rank's share of the job, so the whole job's algorithmic bytes are divided by the ranks (the
    halo rows a shard re-reads are not counted) and set against ONE GPU's measured peak."""
    ab = {key: v / world for key, v in algorithmic_bytes(st).items()}
    peak, peak_src = measured_peak()

    def kern(kname, ms, nbytes, note=None):
        out = {"kernel": kname, "ms": ms / k, "algorithmic_bytes": nbytes,
               "achieved": nbytes / (ms / k * 1e-3) / 1e9, "frac": nbytes / (ms / k * 1e-3) / 1e9 / peak,
               "traffic": recorded_traffic(name, kname)}
        winst = recorded_traffic(name, kname, "warp_inst")
        if winst:  # issue-slot view: executed warp instructions (ncu) over the live-measured duration
            ipeak, sms, mhz = issue_peak_ginst()
            out["issue"] = {"warp_inst": winst, "achieved": winst / (ms / k * 1e-3) / 1e9, "peak": ipeak,
                            "unit": "G warp-inst/s", "frac": winst / (ms / k * 1e-3) / 1e9 / ipeak,
                            "peak_source": f"{sms} SMs x 4 schedulers x {mhz:.0f} MHz",
                            "warp_inst_source": "profiles/traffic.json (ncu --set full capture of this workload)"}
        if kname == "score_group_kernel":
            # integer operations of the REFERENCE formulation -- 8 per pair
            # evaluation of the O(n^2) recurrence plus one mask test per window variant and unit --
            # over the time this kernel takes (the prefix-max DP executes fewer; see "issue").
            ops = (8 * st["sum_pairs"] + st["units"] * st["mean_window_variants"]) / world
            ipeak, sms, mhz = issue_peak_ginst()
            apeak = sms * 128 * mhz * 1e6 / 1e12
            out["alu_reference_formulation"] = {
                "int_ops": int(ops), "achieved": ops / (ms / k * 1e-3) / 1e12, "peak": apeak, "unit": "T int-op/s",
                "frac": ops / (ms / k * 1e-3) / 1e12 / apeak, "peak_source": f"{sms} SMs x 128 lanes x {mhz:.0f} MHz",
                "exact": bool(st["n_exact"])}
        if note:
            out["note"] = note
        return out

    pipe = ab["pipeline"] / (tot_ms / k * 1e-3) / 1e9
    fused = launches == 1
    kname = "fused_score_pair_kernel" if stream else "fused_score_kernel"
    return {
        "bound": "hbm",
        "kernel": (f"one pass = ONE launch of {kname} (window bounds + TMA row staging + pack + chain DP per "
                   "CTA of a few windows; csrc/fused_kernels.cuh)" +
                   (f"; average launch duration = the timed region / its {stream['launches']} launches, which overlap "
                    f"(independent calls, two CTAs per SM): ring of {stream['ring']} input copies = "
                    f"{stream['ring_bytes'] / 1e6:.0f} MB" if stream else "") if fused else
                   "one pass = prep_kernel (pack + window bounds) -> score_group_kernel (-> score_pairwise_kernel only "
                   "when a window was routed to it)"),
        "achieved": pipe, "peak": peak, "unit": "GB/s", "frac": pipe / peak,
        "traffic": ((recorded_traffic(name, kname) or recorded_traffic(name, "fused_score_kernel")) if fused else
                    sum(filter(None, (recorded_traffic(name, x) for x in ("prep_kernel", "score_group_kernel")))) or None),
        "traffic_source": "profiles/traffic.json (per-launch dram bytes of one ncu --set full capture; null when the "
                          "workload has no capture)",
        "peak_source": peak_src, "algorithmic_bytes": ab["pipeline"], "per_gpu": True, "n_gpus": world,
        "bytes_per_unit": ab["pipeline"] * world / st["units"], "pass_ms": tot_ms / k, "launches_per_pass": launches,
        "pass_ms_repeats": (stream or {}).get("repeats_ms_per_step"),
        "kernels_note": ("the pass is the fused kernel (its roofline is the pass-level one above); the two kernels below "
                         "are the launch pipeline's, which larger inputs take, timed alone on this input") if fused else None,
        "kernels": [
            kern("prep_kernel", pack_ms, ab["pack"], "HBM bound: every int8 genotype read once; timed alone"),
            kern("score_group_kernel", score_ms, ab["score"],
                 "integer-ALU / latency bound, not HBM bound (its bytes are pos + the two result tables); "
                 "timed alone")],
    }


def bind_near_gpu(torch, local):
    """Multi-rank runs: keep this rank's threads (and so the pinned host buffers it first touches) on the
    NUMA node its GPU hangs off; eight ranks reading host memory in place otherwise cross the
    socket interconnect at random.  Returns a short description for the JSON line (None = unchanged)."""
    try:
        p = torch.cuda.get_device_properties(local)
        bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bdf}/numa_node") as fh:
            node = int(fh.read())
        if node < 0:
            return None
        with open(f"/sys/devices/system/node/node{node}/cpulist") as fh:
            cpus = set()
            for part in fh.read().strip().split(","):
                a, _, b = part.partition("-")
                cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return f"numa node {node} ({len(cpus)} cpus)"
    except Exception:
        return None


class Ranks:
    """The plumbing a multi-rank run needs (barrier, max / min / sum over ranks, broadcast of a small
    object); trivial on one rank.  NCCL carries nothing else."""

    def __init__(self, torch, dev, world, rank, dist):
        self.torch, self.dev, self.world, self.rank, self.dist = torch, dev, world, rank, dist

    def barrier(self):
        self.torch.cuda.synchronize(self.dev)
        if self.dist is not None:
            self.dist.barrier()
            self.torch.cuda.synchronize(self.dev)

    def reduce(self, values, op="max"):
        t = self.torch.tensor([float(v) for v in values], dtype=self.torch.float64, device=self.dev)
        if self.dist is not None:
            self.dist.all_reduce(t, op={"max": self.dist.ReduceOp.MAX, "min": self.dist.ReduceOp.MIN,
                                        "sum": self.dist.ReduceOp.SUM}[op])
        return [float(x) for x in t.tolist()]

    def bcast(self, obj):
        if self.dist is None:
            return obj
        box = [obj]
        self.dist.broadcast_object_list(box, src=0)
        return box[0]

    def gather(self, obj):
        if self.dist is None:
            return [obj]
        box = [None] * self.world
        self.dist.all_gather_object(box, obj)
        return box

    def shared_table(self, W, T):
        from sstar_b200.sharding import SharedHostTable
        for kind in ("memfd", "shm"):
            table = SharedHostTable(W, T, kind=kind) if self.rank == 0 else None
            handle = self.bcast(table.handle if table is not None else None)
            ok = 1.0
            if table is None:
                try:
                    table = SharedHostTable(W, T, handle=handle)
                except OSError:
                    ok = 0.0
            if self.reduce([ok], "min")[0] == 1.0:
                break
            if table is not None:
                table.close()
        else:
            raise SystemExit("bench: the ranks cannot share a host table")
        views = table.register(self.torch)
        self.barrier()
        return table, views


# ------------------------------------------------------------------------------------------
# strong scaling blocks: fixed total work over the N ranks
# ------------------------------------------------------------------------------------------
def strong_c3(torch, eng, dev, rk, flush, length, steps, e2e_steps, use_graph, flags=0):
    """BASELINE configs[2]: the whole 250 Mb chromosome over the ranks, by window range."""
    from sstar_b200.sharding import plan_window_shards
    from sstar_b200.synth import regular_windows, synth_chromosome_device
    t0 = time.perf_counter()
    # every rank generates the same chromosome on its own GPU (same seed, same device code), keeps
    # its own rows; rank 0 also keeps a host copy for the oracle
    r, t, p = synth_chromosome_device(torch, dev, length, C3["n_ref"], C3["n_tgt"], True, seed=C3["seed"])
    pos = p.cpu().numpy()
    ws, we = regular_windows(pos, 50_000, 10_000)
    W, T, R, V = len(ws), int(t.shape[1]), int(r.shape[1]), len(pos)
    shards = plan_window_shards(pos, ws, we, rk.world)
    if len(shards) != rk.world:
        raise SystemExit("bench: fewer window shards than ranks")
    w0, w1, k0, k1 = shards[rk.rank]
    arrays = dict(ref=r[k0:k1].cpu().numpy(), tgt=t[k0:k1].cpu().numpy(), pos=pos[k0:k1], ws=ws[w0:w1], we=we[w0:w1])
    full = dict(ref=r.cpu().numpy(), tgt=t.cpu().numpy(), pos=pos, ws=ws, we=we) if rk.rank == 0 else None
    del r, t, p
    torch.cuda.empty_cache()
    # the ranks must hold slices of ONE chromosome: their checksums against rank 0's copy
    sums = rk.gather((k0, k1, int(arrays["ref"].sum(dtype=np.int64)), int(arrays["tgt"].sum(dtype=np.int64)),
                      int(arrays["pos"].sum(dtype=np.int64))))
    if rk.rank == 0:
        for (a, b, sr, stg, sp) in sums:
            if (int(full["ref"][a:b].sum(dtype=np.int64)), int(full["tgt"][a:b].sum(dtype=np.int64)),
                    int(full["pos"][a:b].sum(dtype=np.int64))) != (sr, stg, sp):
                raise SystemExit("bench: the ranks generated different chromosomes")
    table, (tab_score, tab_nsnp) = rk.shared_table(W, T)
    case = ShardCase(torch, eng, dev, arrays, out_p=(tab_score[w0:w1], tab_nsnp[w0:w1]), use_graph=use_graph, flags=flags)
    gen_s = time.perf_counter() - t0
    for _ in range(2):
        flush.fill_(1)
        case.step_device()
        case.step_e2e()
    dev_ms = case.timed(case.step_device, steps, flush, rk.barrier)
    rk.barrier()
    if rk.rank == 0:
        table.score[:] = 0  # the timed calls below must refill the whole table
        table.nsnp[:] = -1
    host_cache_flush(torch)
    rk.barrier()
    e2e_s = case.wall(case.step_e2e, e2e_steps, flush)
    rk.barrier()
    same = case.device_equals_host_block()
    dev_ms, e2e_ms = rk.reduce([dev_ms, e2e_s * 1e3], "max")
    same = rk.reduce([1.0 if same else 0.0], "min")[0] == 1.0
    out = None
    if rk.rank == 0:
        if not same:
            raise SystemExit("bench: C3 device-resident tables differ from the gathered host table")
        sel = sorted(set(cut_windows(shards, W)) | set(np.unique(np.linspace(0, W - 1, 512).astype(int)).tolist()))
        t1 = time.perf_counter()
        checked = oracle_check(full, table.score, table.nsnp, sel, "C3 strong block")
        units = W * T
        h2d = V * (R + T) + 4 * V + 8 * W
        out = {"workload": C3["desc"], "V": V, "R": R, "T": T, "W": W, "units": units, "n_gpus": rk.world,
               "length_bp": length, "shards": [[a, b] for a, b, _, _ in shards],
               "value": units / (dev_ms / steps * 1e-3), "unit": UNIT, "ms_per_step": dev_ms / steps, "steps": steps,
               "e2e": {"value": units / (e2e_ms / e2e_steps * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms / e2e_steps,
                       "steps": e2e_steps, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(12 * units),
                       "api": "per rank: ScoreEngine.score_pinned on its own page-locked rows -> its block of the "
                              "shared host table (chunked H2D overlapped with pack + scoring, per-chunk D2H)"},
               "parity": {"checked_units": checked, "windows": len(sel), "bit_exact_vs_oracle": True,
                          "device_tables_equal_host_table": True, "oracle_s": time.perf_counter() - t1},
               "setup_s": gen_s}
    rk.barrier()
    del case
    table.close()
    return out


def strong_c4(torch, eng, dev, rk, flush, scale, steps):
    """BASELINE configs[3]: 22 autosomes over the ranks, whole chromosomes by cost.  Device-resident only
    (27 GB of genotypes at full size: every rank generates ITS chromosomes on its GPU and keeps them there);
    ``scale`` shortens every chromosome (debug)."""
    from oracle import c_oracle
    from sstar_b200.engine import max_window_variants
    from sstar_b200.sharding import plan_chromosome_assignment
    from sstar_b200.synth import regular_windows, synth_chromosome_device
    t0 = time.perf_counter()
    lengths = {f"chr{i + 1}": int(mb * 1e6 * scale) for i, mb in enumerate(C4_MB)}
    plan = plan_chromosome_assignment(lengths, rk.world)
    mine = plan[rk.rank]
    cases, units, V_tot, W_tot = [], 0, 0, 0
    for name in mine:
        r, t, p = synth_chromosome_device(torch, dev, lengths[name], C4["n_ref"], C4["n_tgt"], False,
                                          seed=C4["seed"] + int(name[3:]))
        pos = p.cpu().numpy()
        ws, we = regular_windows(pos, 50_000, 10_000)
        ws_d, we_d = torch.from_numpy(ws).to(dev), torch.from_numpy(we).to(dev)
        out = (torch.empty((len(ws), t.shape[1]), dtype=torch.int64, device=dev),
               torch.empty((len(ws), t.shape[1]), dtype=torch.int32, device=dev))
        hint = max_window_variants(pos, ws, we)
        plan_c = eng.plan_device(r, t, p, ws_d, we_d, *PARAMS, max_win_variants=hint, out=out)
        cases.append(dict(name=name, ref=r, tgt=t, pos=pos, ws=ws, we=we, out=out, plan=plan_c))
        units += len(ws) * int(t.shape[1]); V_tot += len(pos); W_tot += len(ws)
    setup_s = time.perf_counter() - t0

    def step():
        for c in cases:
            c["plan"].run()

    for _ in range(2):
        step()
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(steps)]
    rk.barrier()
    for i in range(steps):
        flush.fill_(i & 1)
        evs[i][0].record()
        step()
        evs[i][1].record()
    rk.barrier()
    ms = sum(e[0].elapsed_time(e[1]) for e in evs)
    # parity: 24 spread windows of this rank's first and last chromosome against the C oracle
    ok, checked = True, 0
    for c in ([cases[0], cases[-1]] if len(cases) > 1 else cases):
        sel = np.unique(np.linspace(0, len(c["ws"]) - 1, 24).astype(int))
        lo = int(np.searchsorted(c["pos"], c["ws"][sel].min(), side="left"))
        hi = int(np.searchsorted(c["pos"], c["we"][sel].max(), side="right"))
        for w in sel:
            a = int(np.searchsorted(c["pos"], c["ws"][w], side="left")); b = int(np.searchsorted(c["pos"], c["we"][w], side="right"))
            so, no = c_oracle.score_windows(c["ref"][a:b].cpu().numpy(), c["tgt"][a:b].cpu().numpy(), c["pos"][a:b],
                                            c["ws"][w:w + 1], c["we"][w:w + 1], *PARAMS)
            ok = ok and bool(np.array_equal(c["out"][0][w].cpu().numpy(), so[0]) and np.array_equal(c["out"][1][w].cpu().numpy(), no[0]))
            checked += so.size
    ms, = rk.reduce([ms], "max")
    units_t, v_t, w_t, checked_t = rk.reduce([units, V_tot, W_tot, checked], "sum")
    ok = rk.reduce([1.0 if ok else 0.0], "min")[0] == 1.0
    loads = rk.gather({"rank": rk.rank, "chromosomes": mine, "units": units})
    out = None
    if rk.rank == 0:
        if not ok:
            raise SystemExit("bench: GPU result differs from the oracle (C4 strong block) -- refusing to report a number")
        out = {"workload": C4["desc"] + ("" if scale == 1.0 else f"; every chromosome at {scale:g} of its length"),
               "V": int(v_t), "R": C4["n_ref"], "T": C4["n_tgt"], "W": int(w_t), "units": int(units_t), "n_gpus": rk.world,
               "assignment": loads, "value": units_t / (ms / steps * 1e-3), "unit": UNIT, "ms_per_step": ms / steps, "steps": steps,
               "hbm_bytes": int(v_t * (C4["n_ref"] + C4["n_tgt"])),
               "e2e": None, "e2e_note": "not measured: 27 GB of int8 genotypes would be 0.5 s of PCIe per pass on one GPU",
               "parity": {"checked_units": int(checked_t), "bit_exact_vs_oracle": True,
                          "windows": "24 spread windows of every rank's first and last chromosome"},
               "setup_s": setup_s}
    rk.barrier()
    del cases
    torch.cuda.empty_cache()
    return out


def strong_c5(torch, eng, dev, rk, flush, steps, e2e_steps):
    """BASELINE configs[4]: 10,000 replicates of 50 kb in contiguous replicate ranges over the ranks."""
    from oracle import c_oracle
    from sstar_b200.sharding import plan_replicate_shards
    from sstar_b200.synth import synth_chromosome
    n_rep, L = C5["n_rep"], C5["seq_len"]
    # one long synthetic chromosome cut into n_rep independent 50 kb "replicates" (positions folded to [1, L])
    ref, tgt, pos = synth_chromosome(L * n_rep, C5["n_ref"], C5["n_tgt"], False, seed=C5["seed"])
    rep_id = (pos.astype(np.int64) - 1) // L
    off = np.searchsorted(rep_id, np.arange(n_rep + 1)).astype(np.int32)
    pos_local = ((pos.astype(np.int64) - 1) % L + 1).astype(np.int32)
    T, R = tgt.shape[1], ref.shape[1]
    shards = plan_replicate_shards(off, rk.world)
    if len(shards) != rk.world:
        raise SystemExit("bench: fewer replicate shards than ranks")
    r0, r1 = shards[rk.rank]
    a, b = int(off[r0]), int(off[r1])
    my_off = (off[r0:r1 + 1] - off[r0]).astype(np.int32)
    hint = int(np.diff(my_off).max())
    pin = [torch.from_numpy(np.ascontiguousarray(x)).pin_memory() for x in (ref[a:b], tgt[a:b], pos_local[a:b], my_off)]
    d = [x.to(dev) for x in pin]
    table, (tab_score, tab_nsnp) = rk.shared_table(n_rep, T)
    out_d = (torch.empty((r1 - r0, T), dtype=torch.int64, device=dev), torch.empty((r1 - r0, T), dtype=torch.int32, device=dev))
    plan = eng.plan_replicates_device(d[0], d[1], d[2], d[3], 1, L, *PARAMS, max_win_variants=hint, out=out_d)

    def step_e2e():
        eng.score_replicates_pinned(pin[0], pin[1], pin[2], pin[3], tab_score[r0:r1], tab_nsnp[r0:r1], 1, L, *PARAMS,
                                    max_win_variants=hint, sync=True)

    for _ in range(3):
        plan.run()
        step_e2e()
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(steps)]
    rk.barrier()
    for i in range(steps):
        flush.fill_(i & 1)
        evs[i][0].record()
        plan.run()
        evs[i][1].record()
    rk.barrier()
    dev_ms = sum(e[0].elapsed_time(e[1]) for e in evs)
    if rk.rank == 0:
        table.score[:] = 0
        table.nsnp[:] = -1
    host_cache_flush(torch)
    rk.barrier()
    e2e_s = 0.0
    for i in range(e2e_steps):
        flush.fill_(i & 1)
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        step_e2e()
        e2e_s += time.perf_counter() - t0
    rk.barrier()
    same = bool(torch.equal(out_d[0].cpu(), tab_score[r0:r1]) and torch.equal(out_d[1].cpu(), tab_nsnp[r0:r1]))
    dev_ms, e2e_ms = rk.reduce([dev_ms, e2e_s * 1e3], "max")
    same = rk.reduce([1.0 if same else 0.0], "min")[0] == 1.0
    out = None
    if rk.rank == 0:
        if not same:
            raise SystemExit("bench: C5 device-resident tables differ from the gathered host table")
        sel = sorted(set(np.unique(np.linspace(0, n_rep - 1, 400).astype(int)).tolist()) |
                     {r for s in shards for r in (s[0], s[1] - 1)})
        for r in sel:
            x, y = off[r], off[r + 1]
            so, no = c_oracle.score_windows(ref[x:y], tgt[x:y], pos_local[x:y], [1], [L], *PARAMS)
            if not (np.array_equal(table.score[r], so[0]) and np.array_equal(table.nsnp[r], no[0])):
                raise SystemExit("bench: GPU result differs from the oracle (C5 strong block) -- refusing to report a number")
        units = n_rep * T
        V = len(pos)
        out = {"workload": C5["desc"], "replicates": n_rep, "V": V, "R": R, "T": T, "units": units, "n_gpus": rk.world,
               "shards": [list(s) for s in shards],
               "value": units / (dev_ms / steps * 1e-3), "unit": UNIT, "ms_per_step": dev_ms / steps, "steps": steps,
               "launches_per_step": plan.launches,
               "e2e": {"value": units / (e2e_ms / e2e_steps * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms / e2e_steps,
                       "steps": e2e_steps, "h2d_bytes_per_step": int(V * (R + T) + 4 * V + 4 * (n_rep + 1)),
                       "d2h_bytes_per_step": int(12 * units),
                       "api": "per rank: ScoreEngine.score_replicates_pinned (H2D of its replicates, "
                              "sstar_b200_score_replicates, D2H into its block of the shared host table)"},
               "parity": {"checked_units": len(sel) * T, "replicates": len(sel), "bit_exact_vs_oracle": True,
                          "device_tables_equal_host_table": True}}
    rk.barrier()
    table.close()
    return out


# ------------------------------------------------------------------------------------------
def run_ours(args, world, rank, local):
    import torch
    import __graft_entry__ as entry
    use_dist = world > 1
    numa = bind_near_gpu(torch, local) if (use_dist and not args.no_numa) else None
    if use_dist:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        if rank == 0:
            entry.build()
        dist.barrier()
    else:
        dist = None
        entry.build()
    from sstar_b200.engine import get_engine
    from sstar_b200.sharding import plan_window_shards

    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    eng = get_engine(local)
    rk = Ranks(torch, dev, world, rank, dist)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    wk = make_workload(args.workload, world)
    st = workload_stats(wk)
    V, R, T, W = st["V"], st["R"], st["T"], st["W"]
    k = args.steps
    large = wk["name"] in LARGE

    # ---- the partition: this rank's window range and the variant rows it touches (halo included)
    shards = plan_window_shards(wk["pos"], wk["ws"], wk["we"], world)
    if len(shards) != world:
        raise SystemExit("bench: fewer window shards than ranks")
    w0, w1, k0, k1 = shards[rank]
    arrays = dict(ref=wk["ref"][k0:k1], tgt=wk["tgt"][k0:k1], pos=wk["pos"][k0:k1], ws=wk["ws"][w0:w1], we=wk["we"][w0:w1])
    table, (tab_score, tab_nsnp) = rk.shared_table(W, T)
    from sstar_b200 import _cabi
    flags = _cabi.FLAG_NO_FUSED if args.no_fused else 0
    case = ShardCase(torch, eng, dev, arrays, out_p=(tab_score[w0:w1], tab_nsnp[w0:w1]), use_graph=not args.no_graph,
                     flags=flags)
    barrier = rk.barrier

    if args.only_device:  # profiling aid: nothing but the device-resident pass is launched
        if (not large) and (not args.no_stream) and (not args.no_graph):
            case.make_stream()  # the headline's launches: the stream of independent one-launch passes
            case.run_stream(max(args.warmup, 3) + k)
        else:
            for i in range(max(args.warmup, 3) + k):
                flush.fill_(i & 1)
                case.step_device()
        torch.cuda.synchronize(dev)
        if rank == 0:
            print(json.dumps({"only_device": True, "workload": wk["name"], "passes": max(args.warmup, 3) + k}))
        return

    for _ in range(max(args.warmup, 3)):
        flush.fill_(1)
        launches_per_step = case.step_device()
        case.step_pack()
        case.step_score()
        case.step_e2e()
        case.step_e2e(staged=True)

    # the headline region: small shapes (the one-launch kernel) are timed as a STREAM of passes over a ring
    # of distinct device copies larger than the L2; shapes that exceed the L2 by themselves keep one
    # flushed, event-bracketed pass per step
    use_stream = (not large) and (not args.no_stream) and (not args.no_graph)
    if use_stream:
        case.make_stream()
        case.run_stream(min(k, 2 * case.ring))  # warm-up of the stream graphs (captures them, too)
        torch.cuda.synchronize(dev)

    sampler = ClockSampler(local)
    sampler.start()
    # ---- region A: device-resident rows, the whole pass per step
    sampler.mark()
    single_k = k if not use_stream else min(k, 200)
    single_ms = case.timed(case.step_device, single_k, flush, barrier)
    stream_launches = stream_reps_ms = None
    if use_stream:
        # the K-pass region three times (each behind its own warm-up passes), the median reported, all listed
        reps = sorted(case.timed_stream(k, barrier, warm=max(args.warmup, 3)) for _ in range(3))
        tot_ms, stream_launches = reps[1]
        stream_reps_ms = [r[0] for r in reps]
        if not case.stream_tables_equal(min(k, case.ring)):
            raise SystemExit("bench: a pass of the timed stream differs from the single pass")
    else:
        tot_ms = single_ms
    sampler.mark()
    # ---- per-kernel split for the roofline: the pack launch alone, the DP launches alone
    pack_ms = case.timed(case.step_pack, k, flush, barrier)
    score_ms = case.timed(case.step_score, k, flush, barrier)
    # ---- region B: end to end from pinned host buffers, wall clock per step (includes the sync)
    barrier()
    host_cache_flush(torch)
    staged_s = case.wall(lambda: case.step_e2e(staged=True), k, flush)  # plain H2D / D2H copies, for comparison
    barrier()
    if rank == 0:
        table.score[:] = 0  # the timed calls below must refill every rank's block
        table.nsnp[:] = -1
    host_cache_flush(torch)  # the rows and the table out of the CPU's caches (see host_cache_flush)
    barrier()
    sampler.mark()
    e2e_single_k = k if not use_stream else min(k, 200)
    e2e_single_s = case.wall(case.step_e2e, e2e_single_k, flush)
    barrier()
    if use_stream:
        # the e2e region: K host-entry calls as a stream (pinned rows -> copy engine -> kernels -> tables stored
        # in place in pinned host memory), every step's H2D and D2H traffic inside, one synchronize at the end
        case.make_e2e_stream()
        case.run_e2e_stream(2 * len(case.e2e_slots))
        if rank == 0:
            table.score[:] = 0
            table.nsnp[:] = -1
        host_cache_flush(torch)  # the slots were written (cloned) a moment ago
        barrier()
        # three repetitions of the K-call region, the median reported and all three listed (wall-clock regions that
        # the host drives see the host's hiccups: on the pool's VMs single repetitions scatter by tens of percent)
        e2e_reps = sorted(case.wall_e2e_stream(k, barrier) for _ in range(3))
        e2e_s = e2e_reps[1]
        barrier()
        if not case.e2e_slots_equal():
            raise SystemExit("bench: a slot of the e2e stream differs from the shared host table's block")
    else:
        e2e_s = e2e_single_s
        e2e_reps = [e2e_single_s]
    sampler.mark()
    sampler.stop_flag.set()
    sampler.join(timeout=2)

    same = case.device_equals_host_block()
    case_ring = (case.ring, case.ring_bytes) if use_stream else None
    tot_ms, e2e_ms, pack_ms, score_ms, staged_ms, single_ms, e2e_single_ms = rk.reduce(
        [tot_ms, e2e_s * 1e3, pack_ms, score_ms, staged_s * 1e3, single_ms, e2e_single_s * 1e3], "max")
    same = rk.reduce([1.0 if same else 0.0], "min")[0] == 1.0
    per_rank = rk.gather({"rank": rank, "windows": [w0, w1], "rows": [k0, k1], "units": (w1 - w0) * T,
                          "hbm_bytes": (k1 - k0) * (R + T)})
    total_units = float(W * T)

    parity = cpu = at_scale = None
    if rank == 0:
        # ---- parity in the same run (outside the timed regions): the GATHERED host table vs the C
        # oracle, bit-exact -- every window, or the cut windows + 48 spread ones for the large shapes
        if not same:
            raise SystemExit("bench: device-resident tables differ from the gathered host table")
        cuts = cut_windows(shards, W)
        sel = np.arange(W) if not large else np.unique(np.linspace(0, W - 1, 48).astype(int))
        checked = oracle_check(wk, table.score, table.nsnp, sel, "gathered host table")
        parity = {"checked_units": checked, "bit_exact_vs_oracle": True, "windows_checked": int(len(sel)),
                  "windows_straddling_shard_cuts": len(cuts), "device_tables_equal_host_table": True}
        if world == 1 and not args.no_cpu_baseline and not large:
            def same_as_gpu(res):  # the CPU-scored units must equal the GPU's
                for w, (s_, n_) in res.items():
                    assert np.array_equal(s_, table.score[w]) and np.array_equal(n_, table.nsnp[w]), "CPU arm != GPU"
            cpu = cpu_arm(wk, st, os.cpu_count() or 1, check=same_as_gpu)
    if world == 1 and args.at_scale != "none" and args.at_scale != wk["name"]:
        # the same kernels on a shape that fills the machine (C2 is a 5.6 MB problem: launch bound)
        bwk = make_workload(args.at_scale, 1)
        bst = workload_stats(bwk)
        big = ShardCase(torch, eng, dev, bwk, use_graph=not args.no_graph, flags=flags)
        kb = 20
        for _ in range(3):
            flush.fill_(1)
            nl = big.step_device(); big.step_pack(); big.step_score(); big.step_e2e()
        b_tot = big.timed(big.step_device, kb, flush, barrier)
        b_pack = big.timed(big.step_pack, kb, flush, barrier)
        b_score = big.timed(big.step_score, kb, flush, barrier)
        host_cache_flush(torch)
        b_e2e = big.wall(big.step_e2e, 5, flush)
        if not big.device_equals_host_block():
            raise SystemExit("bench: at_scale device-resident tables differ from the host tables")
        bsel = np.unique(np.linspace(0, bst["W"] - 1, 48).astype(int))
        # the same shape with missing calls (-1) sprinkled over the target panel: the signed body of the score
        # kernel and the warp-cooperative redo of the pack kernel; device-resident pass, oracle-checked
        missing = {}
        for rate in (5e-5, 5e-3):
            g = torch.Generator(device=dev)
            g.manual_seed(7)
            tgt_m = big.d["tgt"].clone()
            tgt_m[torch.rand(tgt_m.shape, device=dev, generator=g) < rate] = -1
            run = lambda: eng.score_device(big.d["ref"], tgt_m, big.d["pos"], big.d["ws"], big.d["we"], *PARAMS,
                                           max_win_variants=big.hint, flags=flags)
            for _ in range(3):
                out_m = run()
            t_m = big.timed(run, 10, flush, barrier) / 10
            mwk = dict(bwk, tgt=tgt_m.cpu().numpy())
            msel = bsel[::3]
            n_m = oracle_check(mwk, out_m[0].cpu().numpy(), out_m[1].cpu().numpy(), msel, "at_scale, missing calls")
            missing[f"{rate:g}"] = {"ms_per_step": t_m, "vs_clean": t_m / (b_tot / kb), "checked_units": n_m, "bit_exact_vs_oracle": True}
            del tgt_m, out_m, mwk
        at_scale = {"workload": bwk["desc"], **{x: bst[x] for x in ("V", "R", "T", "W", "units")},
                    "value": bst["units"] / (b_tot / kb * 1e-3), "unit": UNIT, "ms_per_step": b_tot / kb,
                    "e2e": {"value": bst["units"] / (b_e2e / 5), "unit": UNIT, "ms_per_step": 1e3 * b_e2e / 5},
                    "roofline": roofline_block(bwk["name"], bst, kb, b_tot, b_pack, b_score, nl),
                    "parity": {"checked_units": oracle_check(bwk, big.score_p.numpy(), big.nsnp_p.numpy(), bsel, "at_scale"),
                               "bit_exact_vs_oracle": True},
                    "missing_calls": {"what": "share of the target calls set to -1 -> device-resident pass (un-graphed call, L2 flushed)",
                                      **missing}}
        del big, bwk

    strong = {}
    if not large and args.strong != "none":
        del case
        torch.cuda.empty_cache()
        if args.strong in ("all", "c3"):
            strong["c3"] = strong_c3(torch, eng, dev, rk, flush, args.c3_length, 10, 5, not args.no_graph, flags)
        if args.strong in ("all", "c5"):
            strong["c5"] = strong_c5(torch, eng, dev, rk, flush, 20, 10)
        if args.strong in ("all", "c4"):
            strong["c4"] = strong_c4(torch, eng, dev, rk, flush, args.c4_scale, 5)

    if rank == 0:
        line = {
            "metric": METRIC, "value": total_units / (tot_ms / k * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": k, "warmup": max(args.warmup, 3), "ms_per_step": tot_ms / k, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int32",
            "data": "synthetic",
            "config": shared_config(wk, st, world),
            "notes": {"arithmetic": "int8 genotypes -> bit-planes -> int32 max-plus DP (int64 when a window's score "
                                    "bound exceeds 2^28) -> int64 scores, int32 SNP counts; bit-exact",
                      "rank0_cpu_binding": numa,
                      "launch": ("direct launches" if args.no_graph else
                                 "a stream of independent passes: one CUDA graph per ring of input copies "
                                 "(ScoreEngine.plan_device_stream)" if use_stream else
                                 "CUDA graph replay of the pass (ScoreEngine.plan_device)") +
                                ("; one-launch kernel disabled (--no-fused)" if args.no_fused else ""),
                      "collectives_in_timed_region": "none (NCCL: barriers and the max / sum of the timings only)"},
            "ranks": per_rank,
            "e2e": {"value": total_units / (e2e_ms / k * 1e-3), "unit": UNIT,
                    "h2d_bytes_per_step": int(sum(p["hbm_bytes"] + 4 * (p["rows"][1] - p["rows"][0]) +
                                                  8 * (p["windows"][1] - p["windows"][0]) for p in per_rank)),
                    "d2h_bytes_per_step": int(12 * W * T), "ms_per_step": e2e_ms / k,
                    "host_caches": "the CPU writes 1 GiB elsewhere before every e2e region (outside it): the page-locked rows "
                                   "are then read from DRAM, not from lines the bench's own set-up left modified in the "
                                   "CPU's caches (13-40 instead of 50 GB/s on this pool's hosts, tools/exp_pinned_slots.py)",
                    "repeats_ms_per_step": [1e3 * x / k for x in e2e_reps],
                    "api": ("per rank K calls of ScoreEngine.score_pinned -> sstar_b200_score_windows_host with "
                            "SSTAR_B200_FLAG_INDEPENDENT, back to back over a ring of 4 page-locked copies of its rows (own "
                            "device scratch and result tables each; slot 0 stores into the shared host table): every "
                            "call's rows go H2D through the copy engine under the previous call's scoring, its tables are "
                            "stored in place in pinned host memory; wall clock from the first call to ONE stream "
                            "synchronize after the last; the gather is part of it; the K-call region is run three times, "
                            "the median is reported (this rank's three: repeats_ms_per_step)" if use_stream else
                            "per rank ScoreEngine.score_pinned -> sstar_b200_score_windows_host on its own pinned rows: "
                            + ("matrices below 32 MB are packed in place over PCIe (the pack CTAs' tile loads are the "
                               "H2D transfer), " if (k1 - k0) * (R + T) < (32 << 20) else
                               "matrices go H2D in row chunks on a copy stream overlapped with pack + scoring, ")
                            + "result blocks stored in place into the shared host table (D2H copies from 8 MB); wall "
                              "clock incl. the stream sync; the gather is part of it (there is nothing left to do after)"),
                    "single_call": {"ms_per_step": e2e_single_ms / e2e_single_k,
                                    "value": total_units / (e2e_single_ms / e2e_single_k * 1e-3), "unit": UNIT,
                                    "steps": e2e_single_k,
                                    "note": "ONE call at a time, wall clock incl. its own stream synchronize (inputs below "
                                            "32 MB are packed in place over PCIe, tables stored in place)"},
                    "staged_copies_ms_per_step": staged_ms / k},
            "gpu_launches": int(stream_launches if use_stream else launches_per_step * k),
            "kernels_per_step": int(launches_per_step),
            "single_pass": {"ms_per_step": single_ms / single_k, "value": total_units / (single_ms / single_k * 1e-3),
                            "unit": UNIT, "steps": single_k,
                            "note": "ONE pass at a time: L2 flushed (256 MiB device write) before it, CUDA events around "
                                    "the one graph replay -- the latency of a lone call, event-to-event floor of an "
                                    "empty kernel after a flush (about 6 us) included"},
            "roofline": roofline_block(wk["name"], st, k, tot_ms, pack_ms, score_ms, int(launches_per_step),
                                       stream={"launches": int(stream_launches), "ring": case_ring[0],
                                               "ring_bytes": case_ring[1],
                                               "repeats_ms_per_step": [x / k for x in stream_reps_ms]} if use_stream else None,
                                       world=world),
            "cpu_baseline": cpu,
            "clocks": sampler.summary(),
            "parity": parity,
        }
        if world > 1:
            line["roofline"]["note"] = "pass-level figures are the aggregate over the ranks (sum of the ranks' bytes over " \
                                       "the slowest rank's time); peak is one GPU's -- divide by n_gpus for a per-GPU fraction"
            line["roofline"]["frac_per_gpu"] = line["roofline"]["frac"] / world
        if at_scale:
            line["at_scale"] = at_scale
        if strong:
            line["strong"] = strong
        print(json.dumps(line), flush=True)
    table.close()
    if use_dist:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--at-scale", default="c3-shard", choices=["none"] + sorted(LARGE),
                    help="N=1 only: also time the same kernels on a shape that fills the GPU")
    ap.add_argument("--strong", default="all", choices=["all", "c3", "c4", "c5", "none"],
                    help="strong-scaling blocks: whole C3 / C4 genome / C5 over the N ranks")
    ap.add_argument("--c4-scale", type=float, default=1.0, help="length factor of the C4 block's chromosomes (debug)")
    ap.add_argument("--c3-length", type=int, default=C3["length"], help="length of the strong block's chromosome (debug)")
    ap.add_argument("--no-numa", action="store_true", help="multi-rank runs: do not bind ranks to their GPU's NUMA node")
    ap.add_argument("--only-device", action="store_true", help="profiling aid: run only the device-resident passes")
    ap.add_argument("--no-fused", action="store_true", help="headline workload: take the pack -> score launches instead of "
                                                            "the one-launch kernel (comparison)")
    ap.add_argument("--no-stream", action="store_true", help="headline: one flushed, event-bracketed pass per step (the "
                                                             "round-1 method) instead of the stream of independent passes")
    ap.add_argument("--no-graph", action="store_true", help="launch the kernels directly instead of replaying the CUDA graph")
    args = ap.parse_args()
    world, rank, local = dist_setup()
    if args.impl == "reference":
        run_reference(args, world, rank)
        return
    run_ours(args, world, rank, local)


if __name__ == "__main__":
    main()
