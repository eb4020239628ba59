"""CPU-only checks of bench.py's contract: the reference arm runs here (no GPU) and prints the line the
driver parses, both arms share one ``config`` object, the algorithmic bytes are SURVEY.md section 8(d)'s."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import REPO

sys.path.insert(0, REPO)


@pytest.fixture(scope="module")
def bench():
    import bench as b
    return b


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=300, cwd=REPO)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["higher_is_better"] is True and line["vs_baseline"] is None
    assert line["unit"] == "scores/s" and line["value"] > 0 and line["steps"] == 1
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["cpu_baseline"]["value"] == line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["gpu_launches"] == 0 and line["config"]["workload"].startswith("C2")


def test_reference_arm_other_ranks_exit_without_work():
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, timeout=120, cwd=REPO, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_both_arms_share_one_config_object(bench):
    from sstar_b200.synth import regular_windows, synth_chromosome
    ref, tgt, pos = synth_chromosome(300_000, 6, 4, False, 3)
    ws, we = regular_windows(pos, 50_000, 10_000)
    wk = dict(ref=ref, tgt=tgt, pos=pos, ws=ws, we=we, seed=3, desc="tiny", name="c2", phased=False)
    st = bench.workload_stats(wk)
    cfg = bench.shared_config(wk, st, 1)
    assert cfg == bench.shared_config(wk, st, 1) and json.loads(json.dumps(cfg)) == cfg
    assert cfg["workload"] == "tiny" and cfg["n_gpus"] == 1 and not ({"model", "global_batch", "seq_len"} & set(cfg))
    # V (R + T) + 4 V + 12 W T
    ab = bench.algorithmic_bytes(st)
    V, R, T, W = st["V"], st["R"], st["T"], st["W"]
    assert ab["pipeline"] == V * (R + T) + 4 * V + 12 * W * T and ab["pack"] + ab["score"] == ab["pipeline"]
    # the exact work statistics against a direct count
    absent = ref.sum(axis=1) == 0
    n = [[int(((tgt[(pos >= a) & (pos <= b), t] != 0) & absent[(pos >= a) & (pos <= b)]).sum()) for t in range(T)]
         for a, b in zip(ws, we)]
    assert st["sum_n"] == int(np.sum(n)) and st["max_n"] == int(np.max(n))


def test_host_cache_flush_is_a_plain_host_write(bench):
    import torch
    bench._HOST_JUNK.clear()
    bench.host_cache_flush(torch, nbytes=1 << 16)
    first = int(bench._HOST_JUNK["buf"][0])
    bench.host_cache_flush(torch, nbytes=1 << 16)
    assert bench._HOST_JUNK["buf"].numel() == 1 << 16 and int(bench._HOST_JUNK["buf"][0]) != first
    bench._HOST_JUNK.clear()


def test_build_reports_the_reference_install():
    import __graft_entry__ as g
    assert g.install_reference() in ("present", "installed", "absent")
