"""ctypes binding of ``libsstar_b200.so`` (C-ABI declared in ``include/sstar_b200.h``).

There is deliberately no fallback: if the CUDA library has not been built this module
raises, and every compute entry point needs a CUDA device.
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libsstar_b200.so")

NAN_SENTINEL = -(2 ** 63)  # SSTAR_B200_NAN
ALGO_AUTO = 0
ALGO_PAIRWISE = 1
ALGOS = {"auto": ALGO_AUTO, "pairwise": ALGO_PAIRWISE}


class Options(ctypes.Structure):
    _fields_ = [("algo", ctypes.c_int32), ("max_win_variants", ctypes.c_int32),
                ("flags", ctypes.c_int32), ("reserved", ctypes.c_int32 * 5)]


class TableOptions(ctypes.Structure):
    _fields_ = [("mode", ctypes.c_int32), ("sample_major", ctypes.c_int32), ("d_pred_value", ctypes.c_void_p),
                ("d_pred_text", ctypes.c_void_p), ("d_pred_off", ctypes.c_void_p), ("n_pred", ctypes.c_int32),
                ("max_pred_len", ctypes.c_int32)]


FLAG_STAGED_COPIES = 1
FLAG_COPY_OUTPUTS = 4
FLAG_SINGLE_CHUNK = 8
FLAG_ZERO_COPY_INPUTS = 32
FLAG_COPY_INPUTS = 64
FLAG_NO_FUSED = 128
FLAG_FORCE_FUSED = 256
FLAG_INDEPENDENT = 512


class SstarB200Error(RuntimeError):
    pass


_c = ctypes
_vp, _i32, _i64, _sz = _c.c_void_p, _c.c_int32, _c.c_int64, _c.c_size_t
_OPT = _c.POINTER(Options)

# name -> (restype, argtypes); mirrors include/sstar_b200.h one to one
SIGNATURES = {
    "sstar_b200_version": (_c.c_int, []),
    "sstar_b200_last_error": (_c.c_char_p, []),
    "sstar_b200_device_count": (_c.c_int, []),
    "sstar_b200_last_launch_count": (_c.c_int, []),
    "sstar_b200_workspace_bytes": (_sz, [_i64, _i32, _i32, _i32, _i32]),
    "sstar_b200_score_windows": (_c.c_int, [_c.c_int, _vp, _vp, _vp, _vp, _i64, _i32, _i32, _vp, _vp,
                                            _i32, _i32, _i32, _i32, _vp, _vp, _vp, _sz]),
    "sstar_b200_score_windows_ex": (_c.c_int, [_c.c_int, _vp, _vp, _vp, _vp, _i64, _i32, _i32, _vp, _vp,
                                               _i32, _i32, _i32, _i32, _vp, _vp, _vp, _sz, _OPT]),
    "sstar_b200_score_replicates": (_c.c_int, [_c.c_int, _vp, _vp, _vp, _vp, _i64, _i32, _i32, _vp, _i32,
                                               _i32, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _sz, _OPT]),
    "sstar_b200_host_scratch_bytes": (_sz, [_i64, _i32, _i32, _i32, _i32]),
    "sstar_b200_score_windows_host": (_c.c_int, [_c.c_int, _vp, _vp, _vp, _vp, _i64, _i32, _i32, _vp, _vp,
                                                 _i32, _i32, _i32, _i32, _vp, _vp, _vp, _sz, _OPT]),
    "sstar_b200_pack": (_c.c_int, [_c.c_int, _vp, _vp, _vp, _i64, _i32, _i32, _vp, _sz]),
    "sstar_b200_score_packed": (_c.c_int, [_c.c_int, _vp, _vp, _vp, _i64, _i32, _i32, _vp, _vp, _i32,
                                           _i32, _i32, _i32, _vp, _vp, _vp, _sz, _OPT]),
    "sstar_b200_vcf_scan": (_c.c_int, [_c.c_char_p, _c.c_char_p, _c.c_int, _c.POINTER(_vp)]),
    "sstar_b200_vcf_scan_ex": (_c.c_int, [_c.c_char_p, _c.c_char_p, _c.POINTER(_c.c_char_p), _i32, _c.c_int,
                                          _c.POINTER(_vp)]),
    "sstar_b200_vcf_free": (None, [_vp]),
    "sstar_b200_vcf_last_error": (_c.c_char_p, []),
    "sstar_b200_vcf_n_samples": (_i32, [_vp]),
    "sstar_b200_vcf_sample": (_c.c_char_p, [_vp, _i32]),
    "sstar_b200_vcf_n_chroms": (_i32, [_vp]),
    "sstar_b200_vcf_chrom": (_c.c_char_p, [_vp, _i32]),
    "sstar_b200_vcf_n_variants": (_i64, [_vp, _i32]),
    "sstar_b200_vcf_pos": (_vp, [_vp, _i32]),
    "sstar_b200_vcf_gt": (_vp, [_vp, _i32]),
    "sstar_b200_vcf_alleles": (_vp, [_vp, _i32, _i32, _c.POINTER(_vp)]),
    "sstar_b200_ingest_workspace_bytes": (_sz, [_i64]),
    "sstar_b200_ingest": (_c.c_int, [_c.c_int, _vp, _vp, _vp, _vp, _i64, _i32, _vp, _i32, _vp, _i32, _i32,
                                     _vp, _vp, _vp, _vp, _sz]),
    "sstar_b200_ingest_ex": (_c.c_int, [_c.c_int, _vp, _vp, _vp, _vp, _i64, _i32, _vp, _i32, _i32, _vp, _i32, _i32,
                                        _i32, _vp, _vp, _vp, _vp, _sz]),
    "sstar_b200_match_numerators": (_c.c_int, [_c.c_int, _vp, _vp, _i64, _i32, _vp, _vp, _vp, _vp, _i32, _vp, _i32,
                                               _i32, _vp]),
    "sstar_b200_snp_set_workspace_bytes": (_sz, [_i64, _i32, _i32, _i32, _i32]),
    "sstar_b200_snp_set_offsets": (_c.c_int, [_c.c_int, _vp, _vp, _vp, _vp, _i64, _i32, _i32, _vp, _vp, _i32, _i32,
                                              _i32, _i32, _vp, _vp, _vp, _vp, _sz, _c.POINTER(Options)]),
    "sstar_b200_snp_set_fill": (_c.c_int, [_c.c_int, _vp, _vp, _vp, _i64, _i32, _i32, _i32, _i32, _i32, _i32, _vp,
                                           _vp, _i64, _vp, _sz, _c.POINTER(Options)]),
    "sstar_b200_tsv_max_bytes": (_sz, [_i32, _i32, _i32, _i32, _i32]),
    "sstar_b200_tsv_workspace_bytes": (_sz, [_i32, _i32]),
    "sstar_b200_format_table": (_c.c_int, [_c.c_int, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _c.c_char_p, _vp, _vp,
                                           _i32, _vp, _vp, _c.POINTER(TableOptions), _vp, _sz, _vp, _sz]),
    "sstar_b200_format_tsv": (_c.c_int, [_c.c_int, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _c.c_char_p, _vp, _vp,
                                         _i32, _vp, _vp, _vp, _sz, _vp, _sz]),
}

_lib = None


def load():
    """Load the shared library (once) and attach the prototypes."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SstarB200Error(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a). sstar_b200 has no CPU fallback.")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        msg = load().sstar_b200_last_error().decode("utf-8", "replace")
        raise SstarB200Error(f"sstar_b200 error {rc}: {msg}")
