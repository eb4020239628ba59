#!/usr/bin/env python
"""The native VCF scanner on a C2-shaped file (20 MB, 32,639 variants x 150 diploids) stored as plain text,
as one gzip stream and as BGZF (bgzip), against the number of worker threads.  Host code only.

    python tools/exp_vcf_scan.py
"""
import ctypes
import gzip
import os
import struct
import sys
import tempfile
import time
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def write_bgzf(path, data, chunk=65280):
    with open(path, "wb") as fh:
        for i in list(range(0, len(data), chunk)) + [len(data)]:
            raw = data[i:i + chunk] if i < len(data) else b""
            co = zlib.compressobj(6, zlib.DEFLATED, -15)
            comp = co.compress(raw) + co.flush()
            fh.write(struct.pack("<BBBBIBBH", 0x1F, 0x8B, 8, 4, 0, 0, 0xFF, 6) + b"BC" +
                     struct.pack("<HH", 2, len(comp) + 25) + comp + struct.pack("<II", zlib.crc32(raw), len(raw)))


def main():
    import __graft_entry__ as entry
    entry.build()
    from sstar_b200 import _cabi
    from sstar_b200.synth import synth_chromosome, write_vcf
    lib = _cabi.load()
    ref, tgt, pos = synth_chromosome(10_000_000, 100, 50, True, seed=12346)
    with tempfile.TemporaryDirectory() as td:
        plain = os.path.join(td, "c2.vcf")
        write_vcf(plain, "1", pos, np.concatenate([ref, tgt], axis=1))
        data = open(plain, "rb").read()
        std = os.path.join(td, "c2.std.vcf.gz")
        with gzip.open(std, "wb", compresslevel=6) as fh:
            fh.write(data)
        bgz = os.path.join(td, "c2.bgz.vcf.gz")
        write_bgzf(bgz, data)
        print(f"{len(data) / 1e6:.1f} MB of text, {len(pos)} variants; host threads available: {os.cpu_count()}")
        for label, path in (("plain text", plain), ("one gzip stream", std), ("BGZF", bgz)):
            row = []
            for nt in (1, 2, 4, 8, 16):
                ts = []
                for _ in range(5):
                    h = ctypes.c_void_p()
                    t0 = time.perf_counter()
                    rc = lib.sstar_b200_vcf_scan_ex(path.encode(), b"1", None, 0, nt, ctypes.byref(h))
                    ts.append(time.perf_counter() - t0)
                    assert rc == 0 and lib.sstar_b200_vcf_n_variants(h, 0) == len(pos)
                    lib.sstar_b200_vcf_free(h)
                row.append(f"{nt}: {min(ts) * 1e3:.1f}")
            print(f"{label:16s} ({os.path.getsize(path) / 1e6:5.1f} MB)  ms by threads  " + "  ".join(row))


if __name__ == "__main__":
    main()
