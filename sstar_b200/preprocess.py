"""Seam S1: ``preprocess(...)`` with the reference's signature (sstar/preprocess.py:28-122).

VCF + sample lists in, feature TSV out -- byte-identical to the reference's table.  ``nprocess`` is
validated like the reference does and otherwise ignored.

On one GPU the whole call is device-resident between the text parser and the file write:
native VCF scan (host, threads) -> raw calls to the GPU -> ingest filters (``sstar_b200_ingest``)
-> pack / window bounds / chain DP (``sstar_b200_score_windows``) -> TSV text
(``sstar_b200_format_tsv``) -> one D2H copy of the text.  With several GPUs (``SSTAR_GPUS``) or
``SSTAR_B200_HOST_INGEST=1`` the filters run in numpy (``utils.read_data``) and the windows are
sharded over the devices; both routes print the same bytes.
"""
from __future__ import annotations

import os

import numpy as np

from .feature_table import COLUMNS, FeatureTable, sample_labels
from .feature_vector_preprocessor import FeatureVectorPreprocessor
from .generators import WindowDataGenerator
from .utils import parse_ind_file, split_genome
from .utils.utils import _scan_vcf, read_anc_allele, validate_unique_variant_positions


def _anc_modes(rec: dict, anc_c: dict) -> np.ndarray:
    """Per variant: 0 keep, 1 flip, 2 drop -- the decision of check_anc_allele (utils.py:363-423).
    Vectorised: the table's positions are matched with one sorted search, the allele strings
    compared as arrays."""
    pos = np.asarray(rec["POS"], dtype=np.int64)
    mode = np.full(len(pos), 2, dtype=np.int8)
    if not anc_c or len(pos) == 0:
        return mode
    keys = np.fromiter(anc_c.keys(), dtype=np.int64, count=len(anc_c))
    order = np.argsort(keys, kind="stable")
    keys = keys[order]
    alleles = np.asarray(list(anc_c.values()), dtype=object)[order]
    at = np.minimum(np.searchsorted(keys, pos), len(keys) - 1)
    known = keys[at] == pos
    anc = alleles[at]
    is_ref = known & (anc == np.asarray(rec["REF"], dtype=object))
    is_alt = known & (anc == np.asarray(rec["ALT"], dtype=object))
    mode[is_alt] = 1
    mode[is_ref] = 0  # (REF wins when both match, as the loop's if / elif did)
    return mode


def _preprocess_device(device, vcf_file, chr_name, ref_ind_file, tgt_ind_file, win_len, win_step, output_file,
                       match_bonus, max_mismatch, mismatch_penalty, is_phased, anc_allele_file):
    from .engine import check_positions, get_engine, max_window_variants
    if win_len <= 0:
        raise ValueError("win_len must be greater than 0.")
    if win_step < 0:
        raise ValueError("win_step must be non-negative.")
    ref_samples, tgt_samples = parse_ind_file(ref_ind_file), parse_ind_file(tgt_ind_file)
    anc = read_anc_allele(anc_allele_file) if anc_allele_file is not None else None
    header, scanned = _scan_vcf(vcf_file, chr_name, need_alleles=anc is not None, samples=ref_samples + tgt_samples)
    if not scanned or chr_name not in scanned:
        raise ValueError(f"{chr_name} is not present in the VCF file.")
    rec = scanned[chr_name]
    validate_unique_variant_positions({chr_name: rec}, chr_name)
    want_r, want_t = set(ref_samples), set(tgt_samples)
    ref_cols = [i for i, s in enumerate(header) if s in want_r]
    tgt_cols = [i for i, s in enumerate(header) if s in want_t]
    mode = _anc_modes(rec, anc[chr_name]) if anc is not None else None
    no_shared = ValueError("No shared variant positions between reference and target "
                           f"on chromosome {chr_name} after filtering.")
    if not ref_cols or not tgt_cols:
        raise no_shared
    eng = get_engine(device)
    torch = eng.torch
    # the fixed-site filter compares with the LISTED sample counts, as the reference does (utils.py:282-291)
    ref_d, tgt_d, pos_d, shared = eng.ingest_device(rec["GT"], rec["POS"], ref_cols, tgt_cols, is_phased, mode,
                                                    n_ref_listed=len(ref_samples), n_tgt_listed=len(tgt_samples))
    if shared == 0:
        raise no_shared
    pos = check_positions(pos_d.cpu().numpy())  # the window search and the DP need ascending positions
    windows = split_genome(pos=pos, chr_name=chr_name, polymorphism_size=win_len, step_size=win_step)
    ws = np.asarray([w[1][0] for w in windows], dtype=np.int64)
    we = np.asarray([w[1][1] for w in windows], dtype=np.int64)
    if ws.size and (ws.min() < -(2 ** 31) or we.max() > 2 ** 31 - 1):
        raise ValueError("window coordinates do not fit int32")
    labels = sample_labels(tgt_samples, 2, is_phased)
    if tgt_d.shape[1] != len(labels):
        # a listed target sample the VCF does not hold: the reference's worker fails in _fmt_res
        # (feature_vector_preprocessor.py:174-181), mp_manager returns "error" (preprocess.py:114-115)
        raise RuntimeError(f"tgt_gts has {tgt_d.shape[1]} columns but {len(labels)} sample labels follow from the "
                           "target list / ploidy / phasing")
    with torch.cuda.device(eng.tdev):
        ws_d = torch.from_numpy(ws.astype(np.int32)).to(eng.tdev)
        we_d = torch.from_numpy(we.astype(np.int32)).to(eng.tdev)
        score_d, nsnp_d = eng.score_device(ref_d, tgt_d, pos_d, ws_d, we_d, match_bonus, max_mismatch,
                                           mismatch_penalty, max_win_variants=max_window_variants(pos, ws, we))
        text_d, total, status = eng.format_tsv_device(score_d, nsnp_d, ws_d, we_d, str(chr_name), labels)
        if status & 1:
            raise RuntimeError("TSV emitter: output buffer too small (internal sizing error)")
        if status:  # a score the device emitter does not print: format on the host
            FeatureTable(chr_name, ws, we, labels, score_d.cpu().numpy(), nsnp_d.cpu().numpy()).to_tsv(output_file)
            return
        pin = eng._pbuf("tsv_pin", total, torch.uint8)[:total]
        pin.copy_(text_d[:total], non_blocking=True)
        torch.cuda.current_stream(eng.tdev).synchronize()
    out_dir = os.path.dirname(str(output_file))
    if out_dir:
        os.makedirs(out_dir, exist_ok=True)
    with open(output_file, "wb") as fh:
        fh.write(("\t".join(COLUMNS) + "\n").encode())
        fh.write(memoryview(pin.numpy()))


def preprocess(vcf_file: str, chr_name: str, ref_ind_file: str, tgt_ind_file: str, win_len: int,
               win_step: int, output_file: str, match_bonus: int, max_mismatch: int,
               mismatch_penalty: int, nprocess: int = 1, ploidy: int = 2, is_phased: bool = True,
               anc_allele_file: str = None) -> None:
    if nprocess <= 0:
        raise ValueError("Number of processes must be greater than 0.")
    from .sharding import visible_devices
    devices = visible_devices()
    if len(devices) == 1 and ploidy == 2 and os.environ.get("SSTAR_B200_HOST_INGEST", "") != "1":
        try:
            _preprocess_device(devices[0], vcf_file, chr_name, ref_ind_file, tgt_ind_file, win_len, win_step,
                               output_file, match_bonus, max_mismatch, mismatch_penalty, is_phased, anc_allele_file)
        except (ValueError, KeyError, OSError):
            raise  # the reference's own input errors
        except Exception as e:
            print(f"sstar_b200: scoring failed ({e})")
            raise SystemExit("Some errors occurred, stopping the program ...")
        return

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
    table.to_tsv(output_file, device=devices[0])
