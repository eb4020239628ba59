"""Executed warp-instructions and stall samples per CUDA source line of one kernel of an .ncu-rep
(needs -lineinfo and --import-source on):  python tools/ncu_lines.py file.ncu-rep <kernel-regex> [min%]"""
import csv, io, subprocess, sys, collections

rep, pat = sys.argv[1], sys.argv[2]
min_pct = float(sys.argv[3]) if len(sys.argv) > 3 else 0.5
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass",
                      "--kernel-name", "regex:" + pat], capture_output=True, text=True).stdout
fname, hdr = None, None
agg = collections.OrderedDict()
seen_kernel = 0
for r in csv.reader(io.StringIO(raw)):
    if not r:
        continue
    if r[0] == "Kernel Name":
        seen_kernel += 1
        if seen_kernel > 1:
            break  # first matching launch only
        continue
    if r[0] == "File Name":
        fname = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        ie, ts, sm = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
        continue
    if hdr is None or len(r) != len(hdr):
        continue
    if r[0] != "":
        cur = (fname, int(r[0]), r[1].strip())
        agg.setdefault(cur, [0, 0, 0])
        continue
    try:
        a = agg[cur]
        a[0] += int(r[ie]); a[1] += int(r[ts]); a[2] += int(r[sm])
    except (ValueError, KeyError):
        pass
tot = sum(a[0] for a in agg.values()) or 1
tsm = sum(a[2] for a in agg.values()) or 1
print(f"total warp-inst {tot}, samples {tsm}")
for (f, ln, src), a in agg.items():
    if a[0] * 100.0 / tot >= min_pct or a[2] * 100.0 / tsm >= min_pct:
        eff = a[1] / a[0] if a[0] else 0
        print(f"{a[0]*100.0/tot:5.1f}% inst {a[2]*100.0/tsm:5.1f}% samp  thr/inst {eff:4.1f} | {f}:{ln:<4d} {src[:100]}")
