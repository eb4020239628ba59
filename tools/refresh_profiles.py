"""Turn the last tools/gpu_profile.sh capture (gpurun_out/) into the committed summaries under profiles/."""
import collections, csv, io, json, os, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.chdir(ROOT)
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
out = {}
for wl in ("c2", "c3-shard", "c4-shard"):
    rep = f"gpurun_out/prof_{tag}_{wl}.ncu-rep"
    for tool, name in (("ncu_summary.py", f"profiles/{tag}_ncu_full_{wl}.txt"),):
        open(name, "w").write(subprocess.run([sys.executable, "tools/" + tool, rep], capture_output=True, text=True).stdout)
    for k in ("score_group", "prep_kernel"):
        short = "score_group" if k == "score_group" else "prep"
        open(f"profiles/{tag}_ncu_lines_{short}_{wl}.txt", "w").write(
            subprocess.run([sys.executable, "tools/ncu_lines.py", rep, k, "1.0"], capture_output=True, text=True).stdout)
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw))); hdr, units = rows[0], rows[1]
    ix = {k: hdr.index(k) for k in ("Kernel Name", "dram__bytes_read.sum", "dram__bytes_write.sum",
                                    "smsp__inst_executed.sum", "gpu__time_duration.sum")}
    tob = lambda v, u: float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
    out[wl] = {}
    for r in rows[2:]:
        i_t = ix["gpu__time_duration.sum"]
        out[wl][r[ix["Kernel Name"]].split("(")[0]] = {
            "dram_bytes": int(tob(r[ix["dram__bytes_read.sum"]], units[ix["dram__bytes_read.sum"]]) +
                              tob(r[ix["dram__bytes_write.sum"]], units[ix["dram__bytes_write.sum"]])),
            "warp_inst": int(float(r[ix["smsp__inst_executed.sum"]])),
            "ncu_us": float(r[i_t]) if units[i_t] == "us" else float(r[i_t]) / 1e3}
json.dump(out, open("profiles/traffic.json", "w"), indent=1)
shutil.copy(f"gpurun_out/launches_{tag}_c2.csv", f"profiles/{tag}_launches_c2.csv")
for src, dst in (("bench_default.json", f"{tag}_bench_default_n1.json"), ("bench_c3-shard.json", f"{tag}_bench_c3-shard.json"),
                 ("bench_c4-shard.json", f"{tag}_bench_c4-shard.json")):
    if os.path.exists("gpurun_out/" + src):
        shutil.copy("gpurun_out/" + src, "profiles/" + dst)
print(json.dumps(out, indent=1))
rows = [r for r in csv.reader(open(f"profiles/{tag}_launches_c2.csv")) if len(r) > 5]
hdr = rows[0]; ki, vi, gi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size")
d = collections.OrderedDict()
for r in rows[1:]:
    try: d.setdefault((r[ki][:40], r[gi]), []).append(float(r[vi]))
    except ValueError: pass
for k, v in d.items(): print(k, len(v), 'median', sorted(v)[len(v) // 2], 'mean', round(sum(v) / len(v)))
for f in (f"{tag}_bench_default_n1.json", f"{tag}_bench_c3-shard.json", f"{tag}_bench_c4-shard.json"):
    x = json.loads(open("profiles/" + f).read().strip().splitlines()[-1]); r = x["roofline"]
    print(f, "pass us %.1f value %.3e frac %.3f e2e ms %.3f" % (x["ms_per_step"] * 1e3, x["value"], r["frac"], x["e2e"]["ms_per_step"]),
          [(k["kernel"], round(k["ms"] * 1e3, 1), round(k["frac"], 3), round(k.get("issue", {}).get("frac", 0), 3)) for k in r["kernels"]])
