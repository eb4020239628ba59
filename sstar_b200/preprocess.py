"""Seam S1: ``preprocess(...)`` with the reference's signature (sstar/preprocess.py:28-122).

VCF + sample lists in, feature TSV out -- byte-identical to the reference's table.  ``nprocess`` is
validated like the reference does and otherwise ignored.

The call is device-resident between the text parser and the file write, on one GPU or several
(``SSTAR_GPUS``): native VCF scan (host, threads; only the listed samples of the wanted chromosome
are kept) -> raw calls to the GPU(s) -> ingest filters (``sstar_b200_ingest_ex``) -> pack / window
bounds / chain DP (``sstar_b200_score_windows``) -> TSV text (``sstar_b200_format_tsv``) -> one D2H
copy of the text per GPU (see ``_score_scanned`` for the multi-GPU layout).  ``ploidy != 2`` or
``SSTAR_B200_HOST_INGEST=1`` run the filters in numpy (``utils.read_data``) instead; both routes
print the same bytes.
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


def _scan_inputs(vcf_file, chr_name, ref_ind_file, tgt_ind_file, anc_allele_file):
    """Host half of the device route: sample lists, ONE streaming pass over the VCF keeping the
    listed samples of the wanted chromosome(s), column indices of the two populations."""
    ref_samples, tgt_samples = parse_ind_file(ref_ind_file), parse_ind_file(tgt_ind_file)
    anc = read_anc_allele(anc_allele_file) if anc_allele_file is not None else None
    header, scanned = _scan_vcf(vcf_file, chr_name, need_alleles=anc is not None, samples=ref_samples + tgt_samples)
    want_r, want_t = set(ref_samples), set(tgt_samples)
    ref_cols = [i for i, s in enumerate(header) if s in want_r]
    tgt_cols = [i for i, s in enumerate(header) if s in want_t]
    return ref_samples, tgt_samples, anc, scanned, ref_cols, tgt_cols


def _score_scanned(devices, rec, chr_name, ref_samples, tgt_samples, ref_cols, tgt_cols, anc, win_len, win_step,
                   output_file, match_bonus, max_mismatch, mismatch_penalty, is_phased):
    """Raw calls of one chromosome -> feature TSV, device-resident on every GPU of ``devices``.

    One GPU: ingest filters -> pack / bounds / chain DP -> TSV text -> one D2H copy of the text.
    Several GPUs: the raw rows are cut into equal ranges and filtered where they land
    (``sstar_b200_ingest`` per GPU); the kept positions (4 bytes a row) come back for the window list
    and ``plan_window_shards``; every GPU then assembles the kept rows its window range touches --
    its own plus the ``win_len - win_step`` halo, fetched from the neighbour's memory with a peer
    copy over NVLink -- scores them and prints its block of rows; the host writes the blocks in
    window order (the reference's gather + stable sort, sstar/mp_manager.py:172-179,
    sstar/preprocess.py:117).  No host-side filtering on any route."""
    from .engine import check_positions, get_engine, max_window_variants
    from .sharding import _worker, plan_window_shards
    validate_unique_variant_positions({chr_name: rec}, chr_name)
    mode = _anc_modes(rec, anc[chr_name]) if anc is not None else None
    no_shared = ValueError("No shared variant positions between reference and target "
                           f"on chromosome {chr_name} after filtering.")
    if not ref_cols or not tgt_cols:
        raise no_shared
    V = len(rec["POS"])
    min_rows = int(os.environ.get("SSTAR_B200_MIN_SHARD_ROWS", "4096"))  # tiny inputs stay on one GPU
    devs = list(devices)[:max(1, min(len(devices), V // max(min_rows, 1)))]
    n = len(devs)
    cuts = [V * i // n for i in range(n + 1)]

    # the fixed-site filter compares with the LISTED sample counts, as the reference does (utils.py:282-291)
    def ingest(dev, a, b):
        return get_engine(dev).ingest_device(rec["GT"][a:b], rec["POS"][a:b], ref_cols, tgt_cols, is_phased,
                                             None if mode is None else mode[a:b], n_ref_listed=len(ref_samples),
                                             n_tgt_listed=len(tgt_samples))

    # every device call runs on its device's worker thread (one engine per GPU, not thread-safe)
    parts = [f.result() for f in [_worker(d).submit(ingest, d, a, b) for d, a, b in zip(devs, cuts[:-1], cuts[1:])]]
    if sum(p[3] for p in parts) == 0:
        raise no_shared
    pos_parts = [p[2].cpu().numpy() for p in parts]
    pos = check_positions(np.concatenate(pos_parts))  # the window search and the DP need ascending positions
    windows = split_genome(pos=pos, chr_name=chr_name, polymorphism_size=win_len, step_size=win_step)
    ws = np.asarray([w[1][0] for w in windows], dtype=np.int64)
    we = np.asarray([w[1][1] for w in windows], dtype=np.int64)
    if ws.size and (ws.min() < -(2 ** 31) or we.max() > 2 ** 31 - 1):
        raise ValueError("window coordinates do not fit int32")
    labels = sample_labels(tgt_samples, 2, is_phased)
    if parts[0][1].shape[1] != len(labels):
        # a listed target sample the VCF does not hold: the reference's worker fails in _fmt_res
        # (feature_vector_preprocessor.py:174-181), mp_manager returns "error" (preprocess.py:114-115)
        raise RuntimeError(f"tgt_gts has {parts[0][1].shape[1]} columns but {len(labels)} sample labels follow from "
                           "the target list / ploidy / phasing")
    own = np.concatenate([[0], np.cumsum([len(x) for x in pos_parts])])  # kept-row ranges the GPUs hold
    shards = plan_window_shards(pos, ws, we, n) if n > 1 else [(0, len(ws), 0, len(pos))]

    def score(dev, shard):
        w0, w1, k0, k1 = shard
        eng = get_engine(dev)
        torch = eng.torch
        with torch.cuda.device(eng.tdev):
            pieces = []
            for g in range(n):  # own rows + halo rows held by the neighbours (peer copies)
                a, b = max(k0, int(own[g])), min(k1, int(own[g + 1]))
                if b > a:
                    sl = slice(a - int(own[g]), b - int(own[g]))
                    pieces.append(tuple(t[sl].to(eng.tdev, non_blocking=True) for t in parts[g][:3]))
            ref_d, tgt_d, pos_d = (pieces[0] if len(pieces) == 1 else
                                   tuple(torch.cat([p[i] for p in pieces]) for i in range(3)))
            ref_d, tgt_d, pos_d = ref_d.contiguous(), tgt_d.contiguous(), pos_d.contiguous()
            ws_d = torch.from_numpy(ws[w0:w1].astype(np.int32)).to(eng.tdev)
            we_d = torch.from_numpy(we[w0:w1].astype(np.int32)).to(eng.tdev)
            score_d, nsnp_d = eng.score_device(ref_d, tgt_d, pos_d, ws_d, we_d, match_bonus, max_mismatch,
                                               mismatch_penalty,
                                               max_win_variants=max_window_variants(pos[k0:k1], ws[w0:w1], we[w0:w1]))
            text_d, total, status = eng.format_tsv_device(score_d, nsnp_d, ws_d, we_d, str(chr_name), labels)
            if status & 1:
                raise RuntimeError("TSV emitter: output buffer too small (internal sizing error)")
            if status:  # a score the device emitter does not print: this block is formatted on the host
                return FeatureTable(chr_name, ws[w0:w1], we[w0:w1], labels, score_d.cpu().numpy(),
                                    nsnp_d.cpu().numpy()).tsv_body()
            pin = torch.empty(max(total, 1), dtype=torch.uint8, pin_memory=True)[:total]  # (torch caches these)
            pin.copy_(text_d[:total], non_blocking=True)
            torch.cuda.current_stream(eng.tdev).synchronize()
            return memoryview(pin.numpy())

    blocks = [f.result() for f in [_worker(d).submit(score, d, sh) for d, sh in zip(devs, shards)]]
    out_dir = os.path.dirname(str(output_file))
    if out_dir:
        os.makedirs(out_dir, exist_ok=True)
    with open(output_file, "wb") as fh:
        fh.write(("\t".join(COLUMNS) + "\n").encode())
        for blk in blocks:
            fh.write(blk)


def _preprocess_device(devices, vcf_file, chr_name, ref_ind_file, tgt_ind_file, win_len, win_step, output_file,
                       match_bonus, max_mismatch, mismatch_penalty, is_phased, anc_allele_file):
    if win_len <= 0:
        raise ValueError("win_len must be greater than 0.")
    if win_step < 0:
        raise ValueError("win_step must be non-negative.")
    ref_samples, tgt_samples, anc, scanned, ref_cols, tgt_cols = _scan_inputs(vcf_file, chr_name, ref_ind_file,
                                                                             tgt_ind_file, anc_allele_file)
    if not scanned or chr_name not in scanned:
        raise ValueError(f"{chr_name} is not present in the VCF file.")
    _score_scanned(devices, scanned[chr_name], chr_name, ref_samples, tgt_samples, ref_cols, tgt_cols, anc, win_len,
                   win_step, output_file, match_bonus, max_mismatch, mismatch_penalty, is_phased)


def preprocess_chromosomes(vcf_file: str, chr_names, ref_ind_file: str, tgt_ind_file: str, win_len: int, win_step: int,
                           output_files, match_bonus: int, max_mismatch: int, mismatch_penalty: int,
                           is_phased: bool = True, anc_allele_file: str = None, devices=None) -> dict:
    """Several chromosomes of one VCF (the reference loops ``preprocess()`` per ``chr_name``,
    reading the file twice each time: sstar/preprocess.py:90-112, sstar/utils/utils.py:267,273).
    Here the file is scanned ONCE for all of them and whole chromosomes go to the GPUs greedily by
    cost (``plan_chromosome_assignment``, cost = variants x target columns ~ W_chr * T), every GPU
    running its chromosomes device-resident; with fewer chromosomes than GPUs each chromosome is
    sharded by window range over all of them instead.  ``output_files``: {chr_name: path} or a
    ``str.format`` pattern with ``{chr}``.  Returns {chr_name: output path}."""
    from .sharding import _worker, plan_chromosome_assignment, visible_devices
    if win_len <= 0:
        raise ValueError("win_len must be greater than 0.")
    if win_step < 0:
        raise ValueError("win_step must be non-negative.")
    devs = visible_devices(devices)
    names = [str(c) for c in chr_names]
    paths = {c: (output_files.format(chr=c) if isinstance(output_files, str) else output_files[c]) for c in names}
    ref_samples, tgt_samples, anc, scanned, ref_cols, tgt_cols = _scan_inputs(
        vcf_file, names[0] if len(names) == 1 else None, ref_ind_file, tgt_ind_file, anc_allele_file)
    for c in names:
        if c not in scanned:
            raise ValueError(f"{c} is not present in the VCF file.")

    def one(c, dev_list):
        _score_scanned(dev_list, scanned[c], c, ref_samples, tgt_samples, ref_cols, tgt_cols, anc, win_len, win_step,
                       paths[c], match_bonus, max_mismatch, mismatch_penalty, is_phased)

    try:
        if len(devs) == 1 or len(names) < len(devs):
            for c in names:
                one(c, devs)
        else:
            plan = plan_chromosome_assignment({c: len(scanned[c]["POS"]) * max(len(tgt_cols), 1) for c in names}, len(devs))
            # one host thread per GPU walks its list; a second pool, so the per-device workers stay
            # free for the calls _score_scanned makes
            from concurrent.futures import ThreadPoolExecutor
            with ThreadPoolExecutor(max_workers=len(devs)) as pool:
                futs = [pool.submit(lambda d=d, cs=cs: [one(c, [d]) for c in cs]) for d, cs in zip(devs, plan)]
                for f in futs:
                    f.result()
    except (ValueError, KeyError, OSError):
        raise
    except Exception as e:
        print(f"sstar_b200: scoring failed ({e})")
        raise SystemExit("Some errors occurred, stopping the program ...")
    return paths


def preprocess(vcf_file: str, chr_name: str, ref_ind_file: str, tgt_ind_file: str, win_len: int,
               win_step: int, output_file: str, match_bonus: int, max_mismatch: int,
               mismatch_penalty: int, nprocess: int = 1, ploidy: int = 2, is_phased: bool = True,
               anc_allele_file: str = None) -> None:
    if nprocess <= 0:
        raise ValueError("Number of processes must be greater than 0.")
    from .sharding import visible_devices
    devices = visible_devices()
    if ploidy == 2 and os.environ.get("SSTAR_B200_HOST_INGEST", "") != "1":
        try:
            _preprocess_device(devices, vcf_file, chr_name, ref_ind_file, tgt_ind_file, win_len, win_step,
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
