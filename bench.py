#!/usr/bin/env python
"""bench.py -- S* (window x individual) scores/sec on B200, with the CPU path beside it.

    python bench.py --gpus N --steps K --warmup W              # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...    # the reference's CPU path (oracle port)

Workload (BASELINE.json configs[1], "C2"): one synthetic 10 Mb msprime-style chromosome,
100 reference + 50 target diploids, unphased dosages, 50 kb windows / 10 kb step.  One "step" is one
pass of the hot path over that chromosome: pack -> window bounds -> chain DP for every
(window, target column) unit.  With N > 1 (torchrun, one rank per GPU) every rank scores its own
chromosome of the same shape (different seed): weak scaling, no collective on the data path.

value  : units/s with inputs resident in HBM (CUDA events around each step, max over ranks).
e2e    : units/s through ScoreEngine.score_pinned (one C-ABI call: H2D of the int8 matrices from
         pinned host memory, kernels, D2H of both result tables), wall clock, max over ranks.
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
WORKLOADS = {
    # name: (length bp, n_ref, n_tgt, phased, win_len, win_step, description)
    "c2": (10_000_000, 100, 50, False, 50_000, 10_000,
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


def make_workload(name, rank):
    from sstar_b200.synth import regular_windows, synth_chromosome, synth_chromosome_device
    length, n_ref, n_tgt, phased, wl, wstep, desc = WORKLOADS[name]
    seed = 12345 + 1 + 1000 * rank
    if name in LARGE:
        import torch
        dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
        r, t, p = synth_chromosome_device(torch, dev, length, n_ref, n_tgt, phased, seed)
        ref, tgt, pos = r.cpu().numpy(), t.cpu().numpy(), p.cpu().numpy()
        del r, t, p
        torch.cuda.empty_cache()
    else:
        ref, tgt, pos = synth_chromosome(length, n_ref, n_tgt, phased, seed)
    ws, we = regular_windows(pos, wl, wstep)
    return dict(ref=ref, tgt=tgt, pos=pos, ws=ws, we=we, seed=seed, desc=desc, name=name)


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
    ``dram_bytes`` = dram__bytes_read.sum + dram__bytes_write.sum, ``warp_inst`` = smsp__inst_executed.sum."""
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

    def __init__(self, index, period=0.01):
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


def dist_setup(n_gpus):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return world, rank, local


def run_reference(args, world, rank):
    """The reference's CPU path (numpy port of its recipe) on all host cores, rank 0 only."""
    if rank != 0:
        return
    from oracle.cpu_baseline import CpuBaseline
    wk = make_workload(args.workload, 0)
    st = workload_stats(wk)
    cores = os.cpu_count() or 1
    base = CpuBaseline(wk["ref"], wk["tgt"], wk["pos"], wk["ws"], wk["we"], PARAMS, cores)
    per_win = base.per_window_seconds()
    # bounded sample: ~20 s of CPU work per step, at most the whole chromosome
    work = min(20.0, 60.0 * cores / max(args.steps + args.warmup, 1))  # keep the whole run near a minute
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
        "config": {"workload": wk["desc"], **{k: st[k] for k in ("V", "R", "T", "W", "units")},
                   "seed": wk["seed"], "params": list(PARAMS)},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": sample_desc + "; pool pre-forked and warm, no mp_manager 5 s tail"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


class DeviceCase:
    """One workload resident on one GPU: the captured pass and the timing helpers."""

    def __init__(self, torch, eng, dev, wk, use_graph=True):
        self.torch, self.eng, self.dev, self.wk = torch, eng, dev, wk
        self.st = st = workload_stats(wk)
        self.hint = st["max_window_variants"]
        W, T = st["W"], st["T"]
        self.pin = {k: torch.from_numpy(np.ascontiguousarray(wk[k])).pin_memory()
                    for k in ("ref", "tgt", "pos", "ws", "we")}
        self.d = {k: v.to(dev) for k, v in self.pin.items()}
        self.score_d = torch.empty((W, T), dtype=torch.int64, device=dev)
        self.nsnp_d = torch.empty((W, T), dtype=torch.int32, device=dev)
        self.score_p = torch.empty((W, T), dtype=torch.int64).pin_memory()
        self.nsnp_p = torch.empty((W, T), dtype=torch.int32).pin_memory()
        # the whole pass (pack -> bounds -> DP) captured once: ONE host call per step
        self.plan = None
        if use_graph:
            d = self.d
            self.plan = eng.plan_device(d["ref"], d["tgt"], d["pos"], d["ws"], d["we"], *PARAMS,
                                        max_win_variants=self.hint, out=(self.score_d, self.nsnp_d))

    def step_device(self):
        if self.plan is not None:
            self.plan.run()
            return self.plan.launches
        d = self.d
        self.eng.score_device(d["ref"], d["tgt"], d["pos"], d["ws"], d["we"], *PARAMS,
                              max_win_variants=self.hint, out=(self.score_d, self.nsnp_d))
        return self.eng.launches

    def step_pack(self):
        self.eng.pack_device(self.d["ref"], self.d["tgt"], W=self.st["W"], max_win_variants=self.hint)

    def step_score(self):
        d = self.d
        self.eng.score_packed_device(self.st["R"], d["tgt"], d["pos"], d["ws"], d["we"], *PARAMS,
                                     max_win_variants=self.hint, out=(self.score_d, self.nsnp_d))

    def step_e2e(self, staged=False):
        p = self.pin
        self.eng.score_pinned(p["ref"], p["tgt"], p["pos"], p["ws"], p["we"], self.score_p, self.nsnp_p,
                              *PARAMS, max_win_variants=self.hint, sync=True, staged=staged)

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

    def check_parity(self, large):
        """GPU vs the C oracle, bit-exact (all windows, or 48 spread ones for the large shapes)."""
        from oracle import c_oracle
        wk, W = self.wk, self.st["W"]
        sel = np.arange(W) if not large else np.unique(np.linspace(0, W - 1, 48).astype(int))
        so, no = c_oracle.score_windows(wk["ref"], wk["tgt"], wk["pos"], wk["ws"][sel], wk["we"][sel], *PARAMS)
        ok = bool(np.array_equal(self.score_p.numpy()[sel], so) and np.array_equal(self.nsnp_p.numpy()[sel], no)
                  and np.array_equal(self.score_d.cpu().numpy()[sel], so)
                  and np.array_equal(self.nsnp_d.cpu().numpy()[sel], no))
        if not ok:
            raise SystemExit("bench: GPU result differs from the oracle -- refusing to report a number")
        return {"checked_units": int(so.size), "bit_exact_vs_oracle": ok}


def roofline_block(name, st, k, tot_ms, pack_ms, score_ms, launches):
    """Pass-level and per-kernel HBM roofline (SURVEY.md section 8d bytes; measured peak)."""
    ab = algorithmic_bytes(st)
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
                            "peak_source": f"{sms} SMs x 4 schedulers x {mhz:.0f} MHz"}
        if kname == "score_group_kernel":
            # integer operations of the REFERENCE formulation -- 8 per pair
            # evaluation of the O(n^2) recurrence plus one mask test per window variant and unit --
            # over the time this kernel takes (the prefix-max DP executes fewer; see "issue").
            ops = 8 * st["sum_pairs"] + st["units"] * st["mean_window_variants"]
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
    return {
        "bound": "hbm",
        "kernel": "one pass = prep_kernel (pack + window bounds) -> score_group_kernel -> score_pairwise_kernel",
        "achieved": pipe, "peak": peak, "unit": "GB/s", "frac": pipe / peak,
        "traffic": sum(filter(None, (recorded_traffic(name, x) for x in ("prep_kernel", "score_group_kernel")))) or None,
        "peak_source": peak_src, "algorithmic_bytes": ab["pipeline"],
        "bytes_per_unit": ab["pipeline"] / st["units"], "pass_ms": tot_ms / k, "launches_per_pass": launches,
        "kernels": [
            kern("prep_kernel", pack_ms, ab["pack"], "HBM bound: every int8 genotype read once; timed alone"),
            kern("score_group_kernel", score_ms, ab["score"],
                 "integer-ALU / latency bound, not HBM bound (its bytes are pos + the two result tables); "
                 "timed alone together with the chained pairwise launch")],
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


def run_ours(args, world, rank, local):
    import torch
    import __graft_entry__ as entry
    if rank == 0:
        entry.build()
    use_dist = world > 1
    numa = bind_near_gpu(torch, local) if (use_dist and not args.no_numa) else None
    if use_dist:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        dist.barrier()
    else:
        dist = None
    from sstar_b200.engine import get_engine

    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    eng = get_engine(local)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    case = DeviceCase(torch, eng, dev, make_workload(args.workload, rank), use_graph=not args.no_graph)
    wk, st = case.wk, case.st
    V, R, T, W = st["V"], st["R"], st["T"], st["W"]
    k = args.steps

    def barrier():
        torch.cuda.synchronize(dev)
        if use_dist:
            dist.barrier()
            torch.cuda.synchronize(dev)

    if args.only_device:  # profiling aid: nothing but the device-resident pass is launched
        for i in range(max(args.warmup, 3) + k):
            flush.fill_(i & 1)
            case.step_device()
        torch.cuda.synchronize(dev)
        print(json.dumps({"only_device": True, "workload": wk["name"], "passes": max(args.warmup, 3) + k}))
        return

    for _ in range(max(args.warmup, 3)):
        flush.fill_(1)
        launches_per_step = case.step_device()
        case.step_pack()
        case.step_score()
        case.step_e2e()
        case.step_e2e(staged=True)

    sampler = ClockSampler(local)
    sampler.start()
    # ---- region A: device-resident inputs, the whole pass per step
    sampler.mark()
    tot_ms = case.timed(case.step_device, k, flush, barrier)
    sampler.mark()
    # ---- per-kernel split for the roofline: the pack launch alone, the DP launches alone
    pack_ms = case.timed(case.step_pack, k, flush, barrier)
    score_ms = case.timed(case.step_score, k, flush, barrier)
    # ---- region B: end to end from pinned host buffers, wall clock per step (includes the sync)
    barrier()
    staged_s = case.wall(lambda: case.step_e2e(staged=True), k, flush)  # plain H2D / D2H copies, for comparison
    barrier()
    sampler.mark()
    e2e_s = case.wall(case.step_e2e, k, flush)
    barrier()
    sampler.mark()
    sampler.stop_flag.set()
    sampler.join(timeout=2)

    times = torch.tensor([tot_ms, e2e_s * 1e3, pack_ms, score_ms, staged_s * 1e3], dtype=torch.float64, device=dev)
    units_t = torch.tensor([float(st["units"])], dtype=torch.float64, device=dev)
    if use_dist:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
        dist.all_reduce(units_t, op=dist.ReduceOp.SUM)
    tot_ms, e2e_ms, pack_ms, score_ms, staged_ms = (float(x) for x in times.tolist())
    total_units = float(units_t.item())

    parity = cpu = at_scale = None
    if rank == 0:
        # ---- parity in the same run (outside the timed regions): GPU vs C oracle, bit-exact
        parity = case.check_parity(wk["name"] in LARGE)
        if world == 1 and not args.no_cpu_baseline and wk["name"] not in LARGE:
            from oracle.cpu_baseline import CpuBaseline
            cores = os.cpu_count() or 1
            base = CpuBaseline(wk["ref"], wk["tgt"], wk["pos"], wk["ws"], wk["we"], PARAMS, cores)
            per_win = base.per_window_seconds()
            n_win = int(min(W, max(cores, 20.0 / max(per_win, 1e-6))))
            sample = np.linspace(0, W - 1, num=n_win).astype(int)
            base.run(sample[: max(cores, n_win // 8)])
            passes = int(min(10, max(1, round(12.0 / max(per_win * n_win, 1e-6)))))  # ~12 core-seconds of work
            dts = []
            for _ in range(passes):
                dt, res = base.run(sample)
                dts.append(dt)
            base.close()
            for w, (s_, n_) in res.items():  # the CPU-scored units must equal the GPU's
                assert np.array_equal(s_, case.score_p.numpy()[w]) and np.array_equal(n_, case.nsnp_p.numpy()[w]), \
                    "CPU port != GPU"
            dt = statistics.mean(dts)
            cpu = {"value": n_win * T / dt, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": f"{n_win} of {W} windows x {T} columns of the {wk['name']} chromosome, mean of {passes} "
                             f"passes ({dt:.2f} s wall each); numpy port of the reference recipe, pool pre-forked and "
                             "warm, no mp_manager 5 s tail",
                   "in_process_us_per_unit": 1e6 * per_win / T}
        if world == 1 and args.at_scale != "none" and args.at_scale != wk["name"]:
            # the same kernels on a shape that fills the machine (C2 is a 5.6 MB problem: launch bound)
            big = DeviceCase(torch, eng, dev, make_workload(args.at_scale, 0), use_graph=not args.no_graph)
            kb = 20
            for _ in range(3):
                flush.fill_(1)
                nl = big.step_device(); big.step_pack(); big.step_score(); big.step_e2e()
            b_tot = big.timed(big.step_device, kb, flush, barrier)
            b_pack = big.timed(big.step_pack, kb, flush, barrier)
            b_score = big.timed(big.step_score, kb, flush, barrier)
            b_e2e = big.wall(big.step_e2e, 5, flush)
            at_scale = {"workload": big.wk["desc"], **{x: big.st[x] for x in ("V", "R", "T", "W", "units")},
                        "value": big.st["units"] / (b_tot / kb * 1e-3), "unit": UNIT, "ms_per_step": b_tot / kb,
                        "e2e": {"value": big.st["units"] / (b_e2e / 5), "unit": UNIT, "ms_per_step": 1e3 * b_e2e / 5},
                        "roofline": roofline_block(big.wk["name"], big.st, kb, b_tot, b_pack, b_score, nl),
                        "parity": big.check_parity(True)}

    if rank == 0:
        line = {
            "metric": METRIC, "value": total_units / (tot_ms / k * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": k, "warmup": max(args.warmup, 3), "ms_per_step": tot_ms / k, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int32",
            "data": "synthetic",
            "config": {"workload": wk["desc"], **st, "seed": wk["seed"], "params": list(PARAMS),
                       "arithmetic": "int8 genotypes -> bit-planes -> int32 max-plus DP (int64 when a window's score "
                                     "bound exceeds 2^28) -> int64 scores, int32 SNP counts; bit-exact",
                       "per_rank": "every rank scores its own chromosome of this shape",
                       "rank0_cpu_binding": numa,
                       "l2": "flushed between timed iterations (256 MiB device write)",
                       "timing": "CUDA events on the launching stream around each step, summed over K steps",
                       "launch": "direct launches" if args.no_graph
                                 else "CUDA graph replay of the pass (ScoreEngine.plan_device)"},
            "e2e": {"value": total_units / (e2e_ms / k * 1e-3), "unit": UNIT,
                    "h2d_bytes_per_step": int(V * (R + T) + 4 * V + 8 * W),
                    "d2h_bytes_per_step": int(12 * W * T), "ms_per_step": e2e_ms / k,
                    "api": "ScoreEngine.score_pinned -> sstar_b200_score_windows_host on pinned host buffers: "
                           + ("matrices below 32 MB are packed in place over PCIe (the pack CTAs' tile loads are the "
                              "H2D transfer), " if V * (R + T) < (32 << 20) else
                              "matrices go H2D in row chunks on a copy stream overlapped with pack + scoring, ")
                           + "result tables stored in place into host memory (D2H copies from 8 MB); wall clock "
                             "incl. the stream sync",
                    "staged_copies_ms_per_step": staged_ms / k},
            "gpu_launches": int(launches_per_step * k),
            "kernels_per_step": int(launches_per_step),
            "roofline": roofline_block(wk["name"], st, k, tot_ms, pack_ms, score_ms, int(launches_per_step)),
            "cpu_baseline": cpu,
            "clocks": sampler.summary(),
            "parity": parity,
        }
        if at_scale:
            line["at_scale"] = at_scale
        print(json.dumps(line), flush=True)
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
    ap.add_argument("--no-numa", action="store_true", help="multi-rank runs: do not bind ranks to their GPU's NUMA node")
    ap.add_argument("--only-device", action="store_true", help="profiling aid: run only the device-resident passes")
    ap.add_argument("--no-graph", action="store_true", help="launch the kernels directly instead of replaying the CUDA graph")
    args = ap.parse_args()
    world, rank, local = dist_setup(args.gpus)
    if args.impl == "reference":
        run_reference(args, world, rank)
        return
    run_ours(args, world, rank, local)


if __name__ == "__main__":
    main()
