"""Time the whole drop-in call preprocess() (VCF file in, TSV file out) on a C2-shaped VCF:
device route (native scan -> GPU ingest -> GPU scoring -> GPU emitter) vs the numpy ingest route.
python tools/exp_preprocess.py [length_bp] [n_ref] [n_tgt]"""
import os, sys, tempfile, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sstar_b200.preprocess import preprocess
from sstar_b200.synth import synth_chromosome, write_vcf

L = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
n_ref = int(sys.argv[2]) if len(sys.argv) > 2 else 100
n_tgt = int(sys.argv[3]) if len(sys.argv) > 3 else 50
with tempfile.TemporaryDirectory() as td:
    ref, tgt, pos = synth_chromosome(L, n_ref, n_tgt, True, seed=12346)
    vcf = os.path.join(td, "c2.vcf")
    names = write_vcf(vcf, "1", pos, np.concatenate([ref, tgt], axis=1))
    open(os.path.join(td, "ref.list"), "w").write("\n".join(names[:n_ref]) + "\n")
    open(os.path.join(td, "tgt.list"), "w").write("\n".join(names[n_ref:]) + "\n")
    kw = dict(vcf_file=vcf, chr_name="1", ref_ind_file=os.path.join(td, "ref.list"),
              tgt_ind_file=os.path.join(td, "tgt.list"), win_len=50000, win_step=10000, match_bonus=5000,
              max_mismatch=5, mismatch_penalty=-10000, nprocess=1, is_phased=False)
    print(f"VCF: {os.path.getsize(vcf)/1e6:.1f} MB, {len(pos)} variants, {n_ref}+{n_tgt} diploids")
    for label, env in (("device route", "0"), ("numpy-ingest route", "1")):
        os.environ["SSTAR_B200_HOST_INGEST"] = env
        out = os.path.join(td, f"out{env}.tsv")
        ts = []
        for _ in range(4):
            t0 = time.perf_counter(); preprocess(output_file=out, **kw); ts.append(time.perf_counter() - t0)
        rows = sum(1 for _ in open(out)) - 1
        print(f"{label:20s}: {min(ts[1:])*1e3:8.1f} ms per call (best of 3 after warm-up), {rows} rows, "
              f"{rows/min(ts[1:])/1e6:.2f} M rows/s end to end (file to file)")
    assert open(os.path.join(td, "out0.tsv"), "rb").read() == open(os.path.join(td, "out1.tsv"), "rb").read()
    print("identical bytes from both routes")
