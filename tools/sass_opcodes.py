"""Per-kernel SASS evidence of the built library: architecture, registers / stack / static shared memory
(cuobjdump -res-usage) and counts of the opcodes that show which hardware paths the code takes --
UBLKCP (cp.async.bulk: the TMA engine's 1-D bulk copy), UTMALDG (cp.async.bulk.tensor), SYNCS
(mbarrier), REDUX (warp reduce), IDP (dp4a), PRMT (byte permute), POPC / FLO (bit scans), LDS / STS,
LDG / STG, ATOM / RED, SHFL, VOTE, BAR, MEMBAR, ACQBULK / griddepcontrol (PDL), no HMMA / UTCMMA
(this path is integer work, no tensor cores).
    python tools/sass_opcodes.py > profiles/r02_sass_opcodes.txt"""
import collections, re, subprocess, sys
lib = sys.argv[1] if len(sys.argv) > 1 else "sstar_b200/csrc/libsstar_b200.so"
res = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True).stdout
usage = {}
for m in re.finditer(r"Function (\S+):\s*\n\s*REG:(\d+) STACK:(\d+) SHARED:(\d+)", res):
    usage[m.group(1)] = (int(m.group(2)), int(m.group(3)), int(m.group(4)))
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
arch = sorted(set(re.findall(r"arch = (sm_\w+)", sass)))
want = ["UBLKCP", "UTMALDG", "SYNCS", "REDUX", "IDP", "PRMT", "POPC", "FLO", "LOP3", "IMNMX", "VIMNMX", "LDS", "STS", "LDG", "STG",
        "LDL", "STL", "ATOMS", "ATOMG", "RED", "SHFL", "VOTE", "BAR", "MEMBAR", "ACQBULK", "HMMA", "UTCHMMA", "IMMA"]
print(f"library: {lib}\narch: {', '.join(arch)}\n")
dem = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip() or n
for part in re.split(r"\n\s*Function : ", sass)[1:]:
    name = part.split("\n", 1)[0].strip()
    ops = collections.Counter()
    total = 0
    for line in part.split("\n"):
        m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m:
            total += 1
            ops[m.group(1).split(".")[0]] += 1
    reg, stack, shared = usage.get(name, (None, None, None))
    short = re.sub(r"\(.*", "", dem(name))
    print(f"{short}\n    instructions {total}, registers {reg}, stack {stack} B, static shared {shared} B")
    hits = [f"{o} x{ops[o]}" for o in want if ops[o]]
    print("    " + ", ".join(hits))
    pdl = len(re.findall(r"ACQBULK|PDL|DEPBAR", part))
    print()
