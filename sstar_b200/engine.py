"""Device-side plumbing for the S* engine: PyTorch owns the buffers, the C-ABI does the work.

``ScoreEngine`` is the host mirror of what the reference spreads over
``WindowDataGenerator`` (slicing), ``mp_manager`` (fan-out / gather) and
``FeatureVectorPreprocessor.run`` -> ``Sstar.compute`` (scoring): one call scores every
(window, target column) unit of a chromosome.

PyTorch is used for device memory, pinned host memory and streams only.
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _cabi
from ._cabi import NAN_SENTINEL, SstarB200Error


def as_int8(a, name: str) -> np.ndarray:
    """Genotype matrix -> C-contiguous int8, refusing values the int8 layout cannot hold.

    The reference carries int8 (phased, scikit-allel) or int64 dosages (``np.sum`` over the
    ploidy axis, sstar/utils/utils.py:305-306); both are tiny integers.
    """
    a = np.asarray(a)
    if a.ndim != 2:
        raise ValueError(f"{name} must be 2-D (n_sites, n_samples), got shape {a.shape}")
    if a.dtype == np.int8:
        return np.ascontiguousarray(a)
    if a.dtype == np.bool_:
        return np.ascontiguousarray(a, dtype=np.int8)
    if not np.issubdtype(a.dtype, np.integer):
        if a.size and not np.all(np.equal(np.mod(a, 1), 0)):
            raise ValueError(f"{name} holds non-integer genotypes")
    if a.size and (a.min() < -128 or a.max() > 127):
        raise ValueError(f"{name} holds genotypes outside the int8 range")
    return np.ascontiguousarray(a, dtype=np.int8)


def check_positions(pos) -> np.ndarray:
    p = np.asarray(pos)
    if p.ndim != 1:
        raise ValueError("pos must be 1-D")
    if p.size and (p.min() < -(2 ** 31) or p.max() > 2 ** 31 - 1):
        raise ValueError("pos does not fit int32")
    p = np.ascontiguousarray(p, dtype=np.int32)
    if p.size > 1 and not np.all(p[1:] > p[:-1]):
        raise ValueError("pos must be strictly ascending (unique, sorted variant positions)")
    return p


def max_window_variants(pos: np.ndarray, win_start, win_end) -> int:
    if len(win_start) == 0 or len(pos) == 0:
        return 0
    lo = np.searchsorted(pos, np.asarray(win_start, dtype=np.int64), side="left")
    hi = np.searchsorted(pos, np.asarray(win_end, dtype=np.int64), side="right")
    return int(max(0, (hi - lo).max()))


class ScoreEngine:
    """One engine per GPU. Not thread-safe; use one per host thread / process."""

    def __init__(self, device: int = 0):
        import torch

        self.torch = torch
        self.lib = _cabi.load()
        if not torch.cuda.is_available():
            raise SstarB200Error("sstar_b200 needs a CUDA device (there is no CPU fallback)")
        if device >= torch.cuda.device_count():
            raise SstarB200Error(f"CUDA device {device} not present")
        self.device = int(device)
        self.tdev = torch.device("cuda", self.device)
        self._dev = {}
        self._pin = {}
        self._pinned_calls = {}
        self.launches = 0  # kernels enqueued by the last call

    # ------------------------------------------------------------------ buffers
    def _dbuf(self, key, nbytes_elems, dtype):
        t = self._dev.get(key)
        if t is None or t.numel() < nbytes_elems or t.dtype != dtype:
            t = self.torch.empty(max(int(nbytes_elems), 1), dtype=dtype, device=self.tdev)
            self._dev[key] = t
        return t

    def _pbuf(self, key, n, dtype):
        t = self._pin.get(key)
        if t is None or t.numel() < n or t.dtype != dtype:
            t = self.torch.empty(max(int(n), 1), dtype=dtype, pin_memory=True)
            self._pin[key] = t
        return t

    def workspace(self, V, R, T, W, max_win_variants=0):
        need = self.lib.sstar_b200_workspace_bytes(int(V), int(R), int(T), int(W), int(max_win_variants))
        if need == 0:
            raise SstarB200Error("invalid shapes for workspace sizing")
        return self._dbuf("ws", need, self.torch.uint8)

    def _stream_ptr(self, stream=None):
        s = stream if stream is not None else self.torch.cuda.current_stream(self.tdev)
        return ctypes.c_void_p(s.cuda_stream)

    # ------------------------------------------------------------------ device API
    def score_device(self, ref_d, tgt_d, pos_d, ws_d, we_d, match_bonus=5000, max_mismatch=5,
                     mismatch_penalty=-10000, algo="auto", max_win_variants=0, out=None, stream=None, flags=0):
        """All inputs are device tensors (int8 [V,R], int8 [V,T], int32 [V], int32 [W] x2)."""
        torch = self.torch
        V, R = ref_d.shape
        T = tgt_d.shape[1]
        W = ws_d.shape[0]
        assert ref_d.dtype == torch.int8 and tgt_d.dtype == torch.int8 and pos_d.dtype == torch.int32
        assert ref_d.is_contiguous() and tgt_d.is_contiguous() and tgt_d.shape[0] == V
        if out is None:
            score = torch.empty((W, T), dtype=torch.int64, device=self.tdev)
            nsnp = torch.empty((W, T), dtype=torch.int32, device=self.tdev)
        else:
            score, nsnp = out
        ws = self.workspace(V, R, T, W, max_win_variants)
        opts = _cabi.Options(algo=_cabi.ALGOS[algo], max_win_variants=int(max_win_variants), flags=int(flags))
        rc = self.lib.sstar_b200_score_windows_ex(
            self.device, self._stream_ptr(stream), ref_d.data_ptr(), tgt_d.data_ptr(), pos_d.data_ptr(),
            V, R, T, ws_d.data_ptr(), we_d.data_ptr(), W, int(match_bonus), int(max_mismatch),
            int(mismatch_penalty), score.data_ptr(), nsnp.data_ptr(), ws.data_ptr(), ws.numel(),
            ctypes.byref(opts))
        _cabi.check(rc)
        self.launches = self.lib.sstar_b200_last_launch_count()
        return score, nsnp

    def snp_sets_device(self, ref_d, tgt_d, pos_d, ws_d, we_d, match_bonus=5000, max_mismatch=5,
                        mismatch_penalty=-10000, max_win_variants=0):
        """Scores plus the sites on every unit's best chain (the legacy S*_SNP_number / S*_SNPs columns).

        Device tensors in; returns device tensors (score [W,T] int64, nsnp [W,T] int32,
        set_off [W*T+1] int64, set_pos [total] int32): unit u = w*T + t owns
        set_pos[set_off[u]:set_off[u+1]], positions ascending.  One host read (the total) sits
        between the two C-ABI calls."""
        torch = self.torch
        V, R = ref_d.shape
        T = tgt_d.shape[1]
        W = ws_d.shape[0]
        assert ref_d.dtype == torch.int8 and tgt_d.dtype == torch.int8 and pos_d.dtype == torch.int32
        assert ref_d.is_contiguous() and tgt_d.is_contiguous() and tgt_d.shape[0] == V
        score = torch.empty((W, T), dtype=torch.int64, device=self.tdev)
        nsnp = torch.empty((W, T), dtype=torch.int32, device=self.tdev)
        set_off = torch.empty(W * T + 1, dtype=torch.int64, device=self.tdev)
        need = self.lib.sstar_b200_snp_set_workspace_bytes(int(V), int(R), int(T), int(W), int(max_win_variants))
        if need == 0:
            raise SstarB200Error("invalid shapes for workspace sizing")
        ws = self._dbuf("ws", need, torch.uint8)
        opts = _cabi.Options(algo=_cabi.ALGOS["auto"], max_win_variants=int(max_win_variants))
        params = (int(match_bonus), int(max_mismatch), int(mismatch_penalty))
        rc = self.lib.sstar_b200_snp_set_offsets(
            self.device, self._stream_ptr(), ref_d.data_ptr(), tgt_d.data_ptr(), pos_d.data_ptr(), V, R, T,
            ws_d.data_ptr(), we_d.data_ptr(), W, *params, score.data_ptr(), nsnp.data_ptr(), set_off.data_ptr(),
            ws.data_ptr(), ws.numel(), ctypes.byref(opts))
        _cabi.check(rc)
        launches = self.lib.sstar_b200_last_launch_count()
        total = int(set_off[-1].item())
        set_pos = torch.empty(max(total, 1), dtype=torch.int32, device=self.tdev)
        rc = self.lib.sstar_b200_snp_set_fill(
            self.device, self._stream_ptr(), tgt_d.data_ptr(), pos_d.data_ptr(), V, R, T, W, *params,
            set_off.data_ptr(), set_pos.data_ptr(), total, ws.data_ptr(), ws.numel(), ctypes.byref(opts))
        _cabi.check(rc)
        self.launches = launches + self.lib.sstar_b200_last_launch_count()
        return score, nsnp, set_off, set_pos[:total]

    def snp_sets_host(self, ref_gts, tgt_gts, pos, win_start, win_end, match_bonus=5000, max_mismatch=5,
                      mismatch_penalty=-10000):
        """numpy in, numpy out: (score, nsnp, set_off, set_pos) as in ``snp_sets_device``."""
        torch = self.torch
        ref = as_int8(ref_gts, "ref_gts")
        tgt = as_int8(tgt_gts, "tgt_gts")
        p = check_positions(pos)
        ws = np.ascontiguousarray(win_start, dtype=np.int32)
        we = np.ascontiguousarray(win_end, dtype=np.int32)
        if ref.shape[0] != tgt.shape[0] or ref.shape[0] != p.shape[0]:
            raise SstarB200Error("ref_gts, tgt_gts and pos must have the same number of rows")
        dev = lambda a: torch.from_numpy(a).to(self.tdev)
        out = self.snp_sets_device(dev(ref), dev(tgt), dev(p), dev(ws), dev(we), match_bonus, max_mismatch,
                                   mismatch_penalty, max_win_variants=max_window_variants(p, ws, we))
        return tuple(t.cpu().numpy() for t in out)

    def pack_device(self, ref_d, tgt_d, W=1, max_win_variants=0, stream=None):
        V, R = ref_d.shape
        T = tgt_d.shape[1]
        ws = self.workspace(V, R, T, W, max_win_variants)
        rc = self.lib.sstar_b200_pack(self.device, self._stream_ptr(stream), ref_d.data_ptr(),
                                      tgt_d.data_ptr(), V, R, T, ws.data_ptr(), ws.numel())
        _cabi.check(rc)
        self.launches = self.lib.sstar_b200_last_launch_count()
        return ws

    def score_packed_device(self, R, tgt_d, pos_d, ws_d, we_d, match_bonus=5000, max_mismatch=5,
                            mismatch_penalty=-10000, algo="auto", max_win_variants=0, out=None,
                            stream=None):
        torch = self.torch
        V, T = tgt_d.shape
        W = ws_d.shape[0]
        if out is None:
            score = torch.empty((W, T), dtype=torch.int64, device=self.tdev)
            nsnp = torch.empty((W, T), dtype=torch.int32, device=self.tdev)
        else:
            score, nsnp = out
        ws = self.workspace(V, R, T, W, max_win_variants)
        opts = _cabi.Options(algo=_cabi.ALGOS[algo], max_win_variants=int(max_win_variants))
        rc = self.lib.sstar_b200_score_packed(
            self.device, self._stream_ptr(stream), tgt_d.data_ptr(), pos_d.data_ptr(), V, R, T,
            ws_d.data_ptr(), we_d.data_ptr(), W, int(match_bonus), int(max_mismatch),
            int(mismatch_penalty), score.data_ptr(), nsnp.data_ptr(), ws.data_ptr(), ws.numel(),
            ctypes.byref(opts))
        _cabi.check(rc)
        self.launches = self.lib.sstar_b200_last_launch_count()
        return score, nsnp

    def score_replicates_device(self, ref_d, tgt_d, pos_d, rep_off_d, win_start, win_end,
                                match_bonus=5000, max_mismatch=5, mismatch_penalty=-10000,
                                algo="auto", max_win_variants=0, stream=None, out=None, flags=0):
        torch = self.torch
        V, R = ref_d.shape
        T = tgt_d.shape[1]
        N = rep_off_d.shape[0] - 1
        if out is None:
            score = torch.empty((N, T), dtype=torch.int64, device=self.tdev)
            nsnp = torch.empty((N, T), dtype=torch.int32, device=self.tdev)
        else:
            score, nsnp = out
        ws = self.workspace(V, R, T, N, max_win_variants)
        opts = _cabi.Options(algo=_cabi.ALGOS[algo], max_win_variants=int(max_win_variants), flags=int(flags))
        rc = self.lib.sstar_b200_score_replicates(
            self.device, self._stream_ptr(stream), ref_d.data_ptr(), tgt_d.data_ptr(), pos_d.data_ptr(),
            V, R, T, rep_off_d.data_ptr(), N, int(win_start), int(win_end), int(match_bonus),
            int(max_mismatch), int(mismatch_penalty), score.data_ptr(), nsnp.data_ptr(), ws.data_ptr(),
            ws.numel(), ctypes.byref(opts))
        _cabi.check(rc)
        self.launches = self.lib.sstar_b200_last_launch_count()
        return score, nsnp

    def plan_device(self, ref_d, tgt_d, pos_d, ws_d, we_d, match_bonus=5000, max_mismatch=5,
                    mismatch_penalty=-10000, algo="auto", max_win_variants=0, out=None, flags=0):
        """Capture one whole scoring pass over fixed device buffers into a CUDA graph.

        Returns a :class:`ScorePlan`; ``plan.run()`` replays pack -> bounds -> DP with ONE host
        call (the three launches keep their programmatic-dependent-launch edges inside the graph),
        which is what small chromosomes need: at C2 size the kernels take ~20 us and the host-side
        launch path would otherwise dominate.  Refill the input tensors in place between runs.
        """
        torch = self.torch
        W, T = ws_d.shape[0], tgt_d.shape[1]
        if out is None:
            out = (torch.empty((W, T), dtype=torch.int64, device=self.tdev),
                   torch.empty((W, T), dtype=torch.int32, device=self.tdev))
        kw = dict(match_bonus=match_bonus, max_mismatch=max_mismatch, mismatch_penalty=mismatch_penalty,
                  algo=algo, max_win_variants=max_win_variants, out=out, flags=flags)
        with torch.cuda.device(self.tdev):
            side = torch.cuda.Stream(self.tdev)
            side.wait_stream(torch.cuda.current_stream(self.tdev))
            with torch.cuda.stream(side):  # warm-up: sizes the workspace, sets kernel attributes
                self.score_device(ref_d, tgt_d, pos_d, ws_d, we_d, **kw)
            torch.cuda.current_stream(self.tdev).wait_stream(side)
            side.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                self.score_device(ref_d, tgt_d, pos_d, ws_d, we_d, **kw)
            launches = self.launches
        return ScorePlan(graph, out[0], out[1], launches, (ref_d, tgt_d, pos_d, ws_d, we_d, self._dev.get("ws")))

    def plan_device_stream(self, cases, match_bonus=5000, max_mismatch=5, mismatch_penalty=-10000, algo="auto",
                           max_win_variants=0, flags=0):
        """Capture a STREAM of independent passes -- one chromosome after the other, the loop of
        sstar/preprocess.py:90-112 -- into ONE graph.  ``cases``: a list of
        ``(ref_d, tgt_d, pos_d, ws_d, we_d, (score, nsnp))`` over DISTINCT buffers.  Every call carries
        ``SSTAR_B200_FLAG_INDEPENDENT``: consecutive one-launch kernels overlap (the next chromosome's
        CTAs stage their tiles while this one's finish their DP).  Returns a :class:`ScorePlan` whose
        ``outs`` lists the result tables in order."""
        torch = self.torch
        kw = dict(match_bonus=match_bonus, max_mismatch=max_mismatch, mismatch_penalty=mismatch_penalty,
                  algo=algo, max_win_variants=max_win_variants, flags=int(flags) | _cabi.FLAG_INDEPENDENT)
        with torch.cuda.device(self.tdev):
            side = torch.cuda.Stream(self.tdev)
            side.wait_stream(torch.cuda.current_stream(self.tdev))
            with torch.cuda.stream(side):  # warm-up: sizes the workspace, sets kernel attributes
                for c in cases[:1]:
                    self.score_device(*c[:5], out=c[5], **kw)
            torch.cuda.current_stream(self.tdev).wait_stream(side)
            side.synchronize()
            graph = torch.cuda.CUDAGraph()
            launches = 0
            with torch.cuda.graph(graph):
                for c in cases:
                    self.score_device(*c[:5], out=c[5], **kw)
                    launches += self.launches
        plan = ScorePlan(graph, cases[0][5][0], cases[0][5][1], launches, (cases, self._dev.get("ws")))
        plan.outs = [c[5] for c in cases]
        return plan

    def plan_replicates_device(self, ref_d, tgt_d, pos_d, rep_off_d, win_start, win_end, match_bonus=5000,
                               max_mismatch=5, mismatch_penalty=-10000, algo="auto", max_win_variants=0, out=None,
                               flags=0):
        """:meth:`plan_device` for a replicate batch (``sstar_b200_score_replicates``)."""
        torch = self.torch
        N, T = rep_off_d.shape[0] - 1, tgt_d.shape[1]
        if out is None:
            out = (torch.empty((N, T), dtype=torch.int64, device=self.tdev),
                   torch.empty((N, T), dtype=torch.int32, device=self.tdev))
        args = (ref_d, tgt_d, pos_d, rep_off_d, win_start, win_end, match_bonus, max_mismatch, mismatch_penalty)
        kw = dict(algo=algo, max_win_variants=max_win_variants, out=out, flags=flags)
        with torch.cuda.device(self.tdev):
            side = torch.cuda.Stream(self.tdev)
            side.wait_stream(torch.cuda.current_stream(self.tdev))
            with torch.cuda.stream(side):
                self.score_replicates_device(*args, **kw)
            torch.cuda.current_stream(self.tdev).wait_stream(side)
            side.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                self.score_replicates_device(*args, **kw)
            launches = self.launches
        return ScorePlan(graph, out[0], out[1], launches, (ref_d, tgt_d, pos_d, rep_off_d, self._dev.get("ws")))

    # ------------------------------------------------------------------ ingest filters
    def ingest_device(self, gt, pos, ref_cols, tgt_cols, is_phased: bool, mode=None, stream=None,
                      n_ref_listed=None, n_tgt_listed=None):
        """Raw calls ``gt`` int8 [V, N, 2] (numpy or device tensor) -> device tensors
        ``(ref_gts [V', R], tgt_gts [V', T], pos [V'], n_shared)`` after the reference's ingest
        filters (see include/sstar_b200.h, sstar_b200_ingest)."""
        torch = self.torch
        with torch.cuda.device(self.tdev):
            up = lambda a, dt: a if isinstance(a, torch.Tensor) else \
                torch.from_numpy(np.ascontiguousarray(a, dtype=dt)).to(self.tdev)
            gt_d, pos_d = up(gt, np.int8), up(pos, np.int32)
            V, N = int(gt_d.shape[0]), int(gt_d.shape[1])
            rc_d, tc_d = up(ref_cols, np.int32), up(tgt_cols, np.int32)
            mode_d = None if mode is None else up(mode, np.int8)
            n_ref, n_tgt = int(rc_d.numel()), int(tc_d.numel())
            R, T = (n_ref * 2, n_tgt * 2) if is_phased else (n_ref, n_tgt)
            ref_o = torch.empty((max(V, 1), R), dtype=torch.int8, device=self.tdev)
            tgt_o = torch.empty((max(V, 1), T), dtype=torch.int8, device=self.tdev)
            pos_o = torch.empty(max(V, 1), dtype=torch.int32, device=self.tdev)
            ws = self._dbuf("ingest_ws", self.lib.sstar_b200_ingest_workspace_bytes(V), torch.uint8)
            s = stream if stream is not None else torch.cuda.current_stream(self.tdev)
            rc = self.lib.sstar_b200_ingest_ex(
                self.device, ctypes.c_void_p(s.cuda_stream), gt_d.data_ptr(), pos_d.data_ptr(),
                mode_d.data_ptr() if mode_d is not None else None, V, N, rc_d.data_ptr(), n_ref,
                n_ref if n_ref_listed is None else int(n_ref_listed), tc_d.data_ptr(), n_tgt,
                n_tgt if n_tgt_listed is None else int(n_tgt_listed), int(bool(is_phased)), ref_o.data_ptr(),
                tgt_o.data_ptr(), pos_o.data_ptr(), ws.data_ptr(), ws.numel())
            _cabi.check(rc)
            self.launches = self.lib.sstar_b200_last_launch_count()
            with torch.cuda.stream(s):
                head_t = ws[:16].to("cpu", non_blocking=False)  # a copy on `s`: waits for the kernels enqueued on it
            head = head_t.numpy().view(np.int64)
        kept, shared = int(head[0]), int(head[1])
        return ref_o[:kept], tgt_o[:kept], pos_o[:kept], shared

    # ------------------------------------------------------------------ match rates
    def match_numerators(self, gt, tract_lo, tract_hi, tract_tgt, tract_hap, src_cols, ploidy: int = 2):
        """``gt`` int8 [V, N, 2]; per tract a variant row range, a target column and a haplotype
        index (-1 = diploid).  Returns int64 [K, S]: sum over sites called in both samples of
        ``ploidy - |x - y|`` (see include/sstar_b200.h, sstar_b200_match_numerators)."""
        torch = self.torch
        with torch.cuda.device(self.tdev):
            up = lambda a, dt: a if isinstance(a, torch.Tensor) else \
                torch.from_numpy(np.ascontiguousarray(a, dtype=dt)).to(self.tdev)
            gt_d = up(gt, np.int8)
            V, N = int(gt_d.shape[0]), int(gt_d.shape[1])
            d = [up(x, np.int32) for x in (tract_lo, tract_hi, tract_tgt, tract_hap, src_cols)]
            K, S = int(d[0].numel()), int(d[4].numel())
            out = torch.zeros((K, S), dtype=torch.int64, device=self.tdev)
            rc = self.lib.sstar_b200_match_numerators(
                self.device, self._stream_ptr(), gt_d.data_ptr(), V, N, d[0].data_ptr(), d[1].data_ptr(),
                d[2].data_ptr(), d[3].data_ptr(), K, d[4].data_ptr(), S, int(ploidy), out.data_ptr())
            _cabi.check(rc)
            self.launches = self.lib.sstar_b200_last_launch_count()
            return out.cpu().numpy()

    # ------------------------------------------------------------------ result emitter
    def format_tsv_device(self, score_d, nsnp_d, ws_d, we_d, chrom: str, labels, replicate_d=None,
                          col_order_d=None, stream=None, predictions=None, bed=False, sample_major=False):
        """Device tables -> TSV text on the device (no header line).  Returns ``(text_d, total,
        status)``: a uint8 device tensor whose first ``total`` bytes are the rows; ``status != 0``
        means the table must be formatted on the host (see include/sstar_b200.h).

        ``predictions``: float64 array, entry v = the model's prediction for
        ``Region_ind_SNP_number == v`` -- adds the ``Predicted_S*_score`` column, or with
        ``bed=True`` prints the BED rows of the units whose score exceeds it.  ``sample_major``
        orders the rows by (sample, window) instead of (window, sample)."""
        torch = self.torch
        W, T = score_d.shape
        enc = [str(x).encode() for x in labels]
        if len(enc) != T:
            raise ValueError("one label per column is required")
        off = np.zeros(T + 1, dtype=np.int32)
        off[1:] = np.cumsum([len(b) for b in enc])
        max_len = int(max((len(b) for b in enc), default=0))
        with torch.cuda.device(self.tdev):
            lab_d = torch.from_numpy(np.frombuffer(b"".join(enc) + b"\0", dtype=np.uint8).copy()).to(self.tdev)
            off_d = torch.from_numpy(off).to(self.tdev)
            topts, keep, extra = None, [], 0
            if predictions is not None or sample_major or bed:
                topts = _cabi.TableOptions(mode=1 if bed else 0, sample_major=int(bool(sample_major)))
                if predictions is not None:
                    pv = np.ascontiguousarray(predictions, dtype=np.float64)
                    ptxt = [repr(float(x)).encode() for x in pv]  # what DataFrame.to_csv prints for a float
                    poff = np.zeros(len(ptxt) + 1, dtype=np.int32)
                    poff[1:] = np.cumsum([len(b) for b in ptxt])
                    keep = [torch.from_numpy(pv).to(self.tdev),
                            torch.from_numpy(np.frombuffer(b"".join(ptxt) + b"\0", dtype=np.uint8).copy()).to(self.tdev),
                            torch.from_numpy(poff).to(self.tdev)]
                    topts.d_pred_value, topts.d_pred_text, topts.d_pred_off = (k.data_ptr() for k in keep)
                    topts.n_pred = len(pv)
                    topts.max_pred_len = int(max(len(b) for b in ptxt))
                    extra = W * T * (1 + topts.max_pred_len)
            cap = self.lib.sstar_b200_tsv_max_bytes(W, T, len(chrom.encode()), max_len, int(replicate_d is not None)) + extra
            out = self._dbuf("tsv_out", cap, torch.uint8)
            ws = self._dbuf("tsv_ws", self.lib.sstar_b200_tsv_workspace_bytes(W, T), torch.uint8)
            s = stream if stream is not None else torch.cuda.current_stream(self.tdev)
            rc = self.lib.sstar_b200_format_table(
                self.device, ctypes.c_void_p(s.cuda_stream), score_d.data_ptr(), nsnp_d.data_ptr(), ws_d.data_ptr(),
                we_d.data_ptr(), W, T, chrom.encode(), lab_d.data_ptr(), off_d.data_ptr(), max_len,
                replicate_d.data_ptr() if replicate_d is not None else None,
                col_order_d.data_ptr() if col_order_d is not None else None,
                ctypes.byref(topts) if topts is not None else None,
                out.data_ptr(), out.numel(), ws.data_ptr(), ws.numel())
            _cabi.check(rc)
            self.launches = self.lib.sstar_b200_last_launch_count()
            with torch.cuda.stream(s):
                head = ws[:16].to("cpu", non_blocking=False).numpy()  # a copy on `s`: waits for its kernels
        total = int(head[:8].view(np.int64)[0])
        status = int(head[8:12].view(np.uint32)[0])
        return out, total, status

    def format_tsv_host(self, score, nsnp, win_start, win_end, chrom, labels, replicate=None, col_order=None):
        """numpy tables -> TSV text (rows only, a bytes-like view of a pinned buffer) through the
        device emitter, or ``None`` when the table holds a score the device emitter does not print
        (|score| >= 1e16)."""
        torch = self.torch
        with torch.cuda.device(self.tdev):
            up = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a, dtype=dt)).to(self.tdev)
            W = np.asarray(score).shape[0]
            if W == 0:
                return b""
            text_d, total, status = self.format_tsv_device(
                up(score, np.int64), up(nsnp, np.int32), up(win_start, np.int32), up(win_end, np.int32), str(chrom),
                labels, None if replicate is None else up(replicate, np.int32),
                None if col_order is None else up(col_order, np.int32))
            if status & 1:
                raise SstarB200Error("TSV emitter: output buffer too small (internal sizing error)")
            if status:
                return None
            pin = self._pbuf("tsv_pin", total, torch.uint8)[:total]  # page-locked: D2H at PCIe rate
            pin.copy_(text_d[:total], non_blocking=True)
            torch.cuda.current_stream(self.tdev).synchronize()
            return memoryview(pin.numpy())  # valid until the next emitter call on this engine

    # ------------------------------------------------------------------ host API
    def _h2d(self, key, arr: np.ndarray, tdtype):
        """numpy -> pinned staging -> device (async on the current stream)."""
        torch = self.torch
        n = arr.size
        pin = self._pbuf("pin_" + key, n, tdtype)[:n]
        if n:
            pin.copy_(torch.from_numpy(arr.reshape(-1)))
        dev = self._dbuf("dev_" + key, n, tdtype)[:n]
        dev.copy_(pin, non_blocking=True)
        return dev

    def score_pinned(self, ref_p, tgt_p, pos_p, ws_p, we_p, score_p, nsnp_p, match_bonus=5000,
                     max_mismatch=5, mismatch_penalty=-10000, algo="auto", max_win_variants=0,
                     stream=None, sync=True, staged=False, flags=0, scratch_slot=0):
        """End-to-end call on HOST tensors: one C-ABI call.  With page-locked (pinned) tensors the
        kernels read the inputs and write both result tables in place over PCIe (zero-copy);
        ``staged=True`` (or pageable tensors) goes through H2D / D2H copies instead.

        ref_p int8 [V,R], tgt_p int8 [V,T], pos_p int32 [V], ws_p/we_p int32 [W],
        score_p int64 [W,T] and nsnp_p int32 [W,T] (outputs, written when the stream has synced).

        A stream of independent calls: ``flags=FLAG_INDEPENDENT, sync=False`` with ``scratch_slot``
        cycling through a few values (own device scratch each) and distinct host tensors per slot;
        the transfers of consecutive calls then overlap the scoring (one ``synchronize`` at the end).
        """
        V, R = ref_p.shape
        T = tgt_p.shape[1]
        W = ws_p.shape[0]
        # per-shape bookkeeping is cached: this call sits inside the end-to-end time
        key = (V, R, T, W, int(max_win_variants), algo, bool(staged), int(flags))
        cached = self._pinned_calls.get(key)
        if cached is None:
            need = self.lib.sstar_b200_host_scratch_bytes(V, R, T, W, int(max_win_variants))
            if need == 0:
                raise SstarB200Error("invalid shapes for scratch sizing")
            opts = _cabi.Options(algo=_cabi.ALGOS[algo], max_win_variants=int(max_win_variants),
                                 flags=(_cabi.FLAG_STAGED_COPIES if staged else 0) | int(flags))
            if len(self._pinned_calls) > 64:
                self._pinned_calls.clear()
            cached = self._pinned_calls[key] = (need, opts)
        need, opts = cached
        scratch = self._dbuf("host_scratch" if not scratch_slot else f"host_scratch{int(scratch_slot)}", need,
                             self.torch.uint8)
        s = stream if stream is not None else self.torch.cuda.current_stream(self.tdev)
        rc = self.lib.sstar_b200_score_windows_host(
            self.device, ctypes.c_void_p(s.cuda_stream), ref_p.data_ptr(), tgt_p.data_ptr(),
            pos_p.data_ptr(), V, R, T, ws_p.data_ptr(), we_p.data_ptr(), W, int(match_bonus),
            int(max_mismatch), int(mismatch_penalty), score_p.data_ptr(), nsnp_p.data_ptr(),
            scratch.data_ptr(), scratch.numel(), ctypes.byref(opts))
        _cabi.check(rc)
        self.launches = self.lib.sstar_b200_last_launch_count()
        if sync:
            s.synchronize()

    def _pin_copy(self, key, arr: np.ndarray, tdtype, shape):
        n = arr.size
        pin = self._pbuf("pin_" + key, n, tdtype)[:n]
        if n:
            pin.copy_(self.torch.from_numpy(arr.reshape(-1)))
        return pin.view(*shape)

    def score_host_async(self, ref_gts, tgt_gts, pos, win_start, win_end, match_bonus=5000,
                         max_mismatch=5, mismatch_penalty=-10000, algo="auto", flags=0):
        """Enqueue one chromosome (numpy in) and return a handle for :meth:`wait`.
        One call may be in flight per engine (the pinned staging buffers are reused)."""
        torch = self.torch
        ref = as_int8(ref_gts, "ref_gts")
        tgt = as_int8(tgt_gts, "tgt_gts")
        p = check_positions(pos)
        V, R = ref.shape
        T = tgt.shape[1]
        if tgt.shape[0] != V or p.shape[0] != V:
            raise ValueError("ref_gts, tgt_gts and pos must have the same number of sites")
        if R < 1 or T < 1:
            raise ValueError("at least one reference and one target column are required")
        ws = np.ascontiguousarray(win_start, dtype=np.int32)
        we = np.ascontiguousarray(win_end, dtype=np.int32)
        W = ws.shape[0]
        hint = max_window_variants(p, ws, we)
        score_h = self._pbuf("pin_score", W * T, torch.int64)[:W * T].view(W, T)
        nsnp_h = self._pbuf("pin_nsnp", W * T, torch.int32)[:W * T].view(W, T)
        with torch.cuda.device(self.tdev):
            stream = torch.cuda.current_stream(self.tdev)
            self.score_pinned(self._pin_copy("ref", ref, torch.int8, (V, R)),
                              self._pin_copy("tgt", tgt, torch.int8, (V, T)),
                              self._pin_copy("pos", p, torch.int32, (V,)),
                              self._pin_copy("ws", ws, torch.int32, (W,)),
                              self._pin_copy("we", we, torch.int32, (W,)),
                              score_h, nsnp_h, match_bonus, max_mismatch, mismatch_penalty, algo=algo,
                              max_win_variants=hint, stream=stream, sync=False, flags=flags)
        return (stream, score_h, nsnp_h)

    def wait(self, handle):
        stream, score_h, nsnp_h = handle
        stream.synchronize()
        return score_h.numpy().copy(), nsnp_h.numpy().copy()

    def score_host(self, ref_gts, tgt_gts, pos, win_start, win_end, match_bonus=5000, max_mismatch=5,
                   mismatch_penalty=-10000, algo="auto", flags=0):
        """numpy in, numpy out: (score[W,T] int64 with NAN_SENTINEL, nsnp[W,T] int32)."""
        return self.wait(self.score_host_async(ref_gts, tgt_gts, pos, win_start, win_end, match_bonus,
                                               max_mismatch, mismatch_penalty, algo=algo, flags=flags))

    def score_replicates_pinned(self, ref_p, tgt_p, pos_p, off_p, score_p, nsnp_p, win_start, win_end,
                                match_bonus=5000, max_mismatch=5, mismatch_penalty=-10000, algo="auto",
                                max_win_variants=0, sync=True):
        """A replicate batch end to end on PAGE-LOCKED host tensors: asynchronous H2D of the
        concatenated replicates, one device call, D2H of the two ``[Nrep, T]`` tables into
        ``score_p`` / ``nsnp_p`` (page-locked or registered host memory)."""
        torch = self.torch
        V, R = ref_p.shape
        T = tgt_p.shape[1]
        N = off_p.shape[0] - 1
        with torch.cuda.device(self.tdev):
            d = {}
            for key, src, dt in (("ref", ref_p, torch.int8), ("tgt", tgt_p, torch.int8), ("pos", pos_p, torch.int32),
                                 ("off", off_p, torch.int32)):
                n = src.numel()
                d[key] = self._dbuf("dev_" + key, n, dt)[:n]
                if n:
                    d[key].copy_(src.reshape(-1), non_blocking=True)
            out = (self._dbuf("rep_score", N * T, torch.int64)[:N * T].view(N, T),
                   self._dbuf("rep_nsnp", N * T, torch.int32)[:N * T].view(N, T))
            self.score_replicates_device(d["ref"].view(V, R), d["tgt"].view(V, T), d["pos"], d["off"], win_start,
                                         win_end, match_bonus, max_mismatch, mismatch_penalty, algo=algo,
                                         max_win_variants=max_win_variants, out=out)
            score_p.copy_(out[0], non_blocking=True)
            nsnp_p.copy_(out[1], non_blocking=True)
            if sync:
                torch.cuda.current_stream(self.tdev).synchronize()

    def score_replicates_host(self, ref_gts, tgt_gts, pos, rep_offset, win_start, win_end,
                              match_bonus=5000, max_mismatch=5, mismatch_penalty=-10000, algo="auto"):
        """Concatenated replicates (simulate path): outputs are [Nrep, T]."""
        torch = self.torch
        ref = as_int8(ref_gts, "ref_gts")
        tgt = as_int8(tgt_gts, "tgt_gts")
        p = np.ascontiguousarray(pos, dtype=np.int32)
        off = np.ascontiguousarray(rep_offset, dtype=np.int32)
        V, R = ref.shape
        T = tgt.shape[1]
        N = off.shape[0] - 1
        if off[0] != 0 or off[-1] != V or np.any(np.diff(off) < 0):
            raise ValueError("rep_offset must rise from 0 to V")
        for r in range(N):
            seg = p[off[r]:off[r + 1]]
            if seg.size > 1 and not np.all(seg[1:] > seg[:-1]):
                raise ValueError(f"positions of replicate {r} are not strictly ascending")
        hint = int(np.diff(off).max()) if N else 0
        with torch.cuda.device(self.tdev):
            ref_d = self._h2d("ref", ref, torch.int8).view(V, R)
            tgt_d = self._h2d("tgt", tgt, torch.int8).view(V, T)
            pos_d = self._h2d("pos", p, torch.int32)
            off_d = self._h2d("off", off, torch.int32)
            score_d, nsnp_d = self.score_replicates_device(
                ref_d, tgt_d, pos_d, off_d, win_start, win_end, match_bonus, max_mismatch,
                mismatch_penalty, algo=algo, max_win_variants=hint)
            score = score_d.cpu().numpy()
            nsnp = nsnp_d.cpu().numpy()
        return score, nsnp


class ScorePlan:
    """A captured scoring pass (see :meth:`ScoreEngine.plan_device`)."""

    def __init__(self, graph, score, nsnp, launches, keepalive):
        self.graph, self.score, self.nsnp, self.launches = graph, score, nsnp, launches
        self._keepalive = keepalive  # the graph holds raw pointers into these tensors

    def run(self):
        """Enqueue the pass on the current stream; results land in ``self.score`` / ``self.nsnp``."""
        self.graph.replay()
        return self.score, self.nsnp


_ENGINES = {}


def get_engine(device: int = 0) -> ScoreEngine:
    eng = _ENGINES.get(device)
    if eng is None:
        eng = _ENGINES[device] = ScoreEngine(device)
    return eng


def scores_to_float(score: np.ndarray) -> np.ndarray:
    """int64 with NAN_SENTINEL -> float64 with NaN (what the reference returns)."""
    out = score.astype(np.float64)
    out[score == NAN_SENTINEL] = np.nan
    return out
