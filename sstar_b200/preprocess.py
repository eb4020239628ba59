"""Seam S1: ``preprocess(...)`` with the reference's signature (sstar/preprocess.py:28-122).

VCF + sample lists in, feature TSV out -- byte-identical to the reference's table.  The scoring
runs on the GPU(s); ``nprocess`` is validated like the reference does and otherwise ignored.
"""
from __future__ import annotations

from .feature_vector_preprocessor import FeatureVectorPreprocessor
from .generators import WindowDataGenerator
from .utils import parse_ind_file


def preprocess(vcf_file: str, chr_name: str, ref_ind_file: str, tgt_ind_file: str, win_len: int,
               win_step: int, output_file: str, match_bonus: int, max_mismatch: int,
               mismatch_penalty: int, nprocess: int = 1, ploidy: int = 2, is_phased: bool = True,
               anc_allele_file: str = None) -> None:
    if nprocess <= 0:
        raise ValueError("Number of processes must be greater than 0.")

    generator = WindowDataGenerator(vcf_file=vcf_file, ref_ind_file=ref_ind_file,
                                    tgt_ind_file=tgt_ind_file, anc_allele_file=anc_allele_file,
                                    ploidy=ploidy, is_phased=is_phased, chr_name=chr_name,
                                    win_len=win_len, win_step=win_step)
    preprocessor = FeatureVectorPreprocessor(ref_samples=parse_ind_file(ref_ind_file),
                                             tgt_samples=parse_ind_file(tgt_ind_file),
                                             match_bonus=match_bonus, max_mismatch=max_mismatch,
                                             mismatch_penalty=mismatch_penalty)
    try:
        table = preprocessor.run_chromosome(**generator.chromosome())
    except Exception as e:
        print(f"sstar_b200: scoring failed ({e})")
        raise SystemExit("Some errors occurred, stopping the program ...")
    # windows are already in (Chromosome, Start, End) order == the reference's stable sort;
    # the rows are printed by the device emitter (same bytes as pandas' to_csv)
    from .sharding import visible_devices
    table.to_tsv(output_file, device=visible_devices()[0])
