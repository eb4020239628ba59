"""On the GPU box, after tools/gpu_profile_r02.sh captured the .ncu-rep files: write the text summaries
(gpurun_out/profiles_r02/) and traffic.json, then drop the reports (they exceed what gpurun copies back)."""
import csv, io, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.chdir(ROOT)
out_dir = "gpurun_out/profiles_r02"
os.makedirs(out_dir, exist_ok=True)
run = lambda *a: subprocess.run([sys.executable, *a], capture_output=True, text=True).stdout
tob = lambda v, u: float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
traffic = {}
for wl in ("c2", "c3-shard", "c4-shard", "aux"):
    rep = f"gpurun_out/prof_r02_{wl}.ncu-rep"
    if not os.path.exists(rep):
        continue
    open(f"{out_dir}/r02_ncu_full_{wl}.txt", "w").write(run("tools/ncu_summary.py", rep))
    if wl != "aux":
        for k, short in (("fused_score", "fused"), ("score_group", "score_group"), ("prep_kernel", "prep")):
            txt = run("tools/ncu_lines.py", rep, k, "1.0")
            if "total warp-inst 0," not in txt and txt.strip():
                open(f"{out_dir}/r02_ncu_lines_{short}_{wl}.txt", "w").write(txt)
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    if len(rows) < 3:
        continue
    hdr, units = rows[0], rows[1]
    ix = {k: hdr.index(k) for k in ("Kernel Name", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
                                    "gpu__time_duration.sum")}
    traffic[wl] = {}
    for r in rows[2:]:
        i_t = ix["gpu__time_duration.sum"]
        name = r[ix["Kernel Name"]].split("(")[0]
        if name in traffic[wl]:
            continue  # first launch of each kernel
        traffic[wl][name] = {
            "dram_bytes": int(tob(r[ix["dram__bytes_read.sum"]], units[ix["dram__bytes_read.sum"]]) +
                              tob(r[ix["dram__bytes_write.sum"]], units[ix["dram__bytes_write.sum"]])),
            "warp_inst": int(float(r[ix["smsp__inst_executed.sum"]])),
            "ncu_us": float(r[i_t]) if units[i_t] == "us" else float(r[i_t]) / 1e3}
    os.remove(rep)
json.dump(traffic, open(f"{out_dir}/traffic.json", "w"), indent=1)
print(json.dumps(traffic, indent=1))
