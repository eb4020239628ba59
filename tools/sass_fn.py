"""Dump the SASS of one kernel of the built library: python tools/sass_fn.py <substring> [out]"""
import subprocess, sys, re
lib = "sstar_b200/csrc/libsstar_b200.so"
sub = sys.argv[1]
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
parts = re.split(r"\n\s*Function : ", txt)
for p in parts[1:]:
    name = p.split("\n", 1)[0]
    if sub in name:
        lines = [l for l in p.split("\n") if re.search(r"/\*[0-9a-f]{4}\*/", l)]
        ins = [re.sub(r"\s*/\* 0x[0-9a-f]+ \*/\s*$", "", l).strip() for l in lines]
        out = "\n".join(ins)
        if len(sys.argv) > 2:
            open(sys.argv[2], "w").write(out)
        print(name, len(ins), "instructions")
