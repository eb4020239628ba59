"""Inference half of the reference's quantile-regression model (sstar/quantile_regression.py:71-135).

The trained model is an ONNX ``TreeEnsembleRegressor`` over ONE integer feature
(``Region_ind_SNP_number``).  onnx / onnxruntime are not dependencies here: the file is read with a
minimal protobuf wire-format walker and the ensemble is evaluated directly -- once per distinct
feature value, which turns the model into a lookup table.  Training (GradientBoostingRegressor ->
ONNX, :32-67) is out of scope.

Numerics: leaf weights and thresholds are the float32 values stored in the file, the feature is
compared as float32 (``BRANCH_LEQ``: x <= threshold takes the true branch), and the tree outputs
are summed in float32 in tree order, then the base value is added in float32 and the result
widened to float64 -- what onnxruntime returns to the reference (a float32 prediction), so a
``S*_score > prediction`` comparison next to a rounding boundary falls the same way.  The
reference's own test compares predictions with rtol = 1e-6 (tests/test_infer.py:50-57).
"""
from __future__ import annotations

import os
import struct

import numpy as np

PRED_COLUMN = "Predicted_S*_score"


def _varint(b: bytes, i: int):
    r = s = 0
    while True:
        c = b[i]
        i += 1
        r |= (c & 0x7F) << s
        s += 7
        if c < 0x80:
            return r, i


def _fields(b: bytes):
    """(field number, wire type, value) triples of one protobuf message."""
    i, out = 0, []
    while i < len(b):
        key, i = _varint(b, i)
        f, w = key >> 3, key & 7
        if w == 0:
            v, i = _varint(b, i)
        elif w == 1:
            v, i = b[i:i + 8], i + 8
        elif w == 2:
            n, i = _varint(b, i)
            v, i = b[i:i + n], i + n
        elif w == 5:
            v, i = b[i:i + 4], i + 4
        else:
            raise ValueError(f"unsupported protobuf wire type {w}")
        out.append((f, w, v))
    return out


def _attributes(node: bytes) -> dict:
    """NodeProto.attribute (field 5) -> {name: floats | ints | strings}."""
    attrs = {}
    for f, _, v in _fields(node):
        if f != 5:
            continue
        name, floats, ints, strs, single = None, [], [], [], None
        for a, w, x in _fields(v):
            if a == 1:
                name = x.decode()
            elif a == 7:  # floats (packed or not)
                floats += list(struct.unpack(f"<{len(x) // 4}f", x))
            elif a == 8:  # ints (packed or not)
                if w == 2:
                    i = 0
                    while i < len(x):
                        val, i = _varint(x, i)
                        ints.append(val)
                else:
                    ints.append(x)
            elif a == 9:
                strs.append(x.decode())
            elif a == 4:
                single = x.decode()
            elif a == 3:
                single = x
            elif a == 2:
                single = struct.unpack("<f", x)[0]
        attrs[name] = floats or ints or strs or single
    return attrs


class TreeEnsemble:
    """ai.onnx.ml TreeEnsembleRegressor with one target, SUM aggregation and no post-transform."""

    def __init__(self, attrs: dict):
        if attrs.get("post_transform", "NONE") != "NONE" or attrs.get("aggregate_function", "SUM") != "SUM":
            raise ValueError("only SUM aggregation without post-transform is supported")
        if int(attrs.get("n_targets", 1)) != 1:
            raise ValueError("only single-target ensembles are supported")
        tree = np.asarray(attrs["nodes_treeids"], dtype=np.int64)
        node = np.asarray(attrs["nodes_nodeids"], dtype=np.int64)
        self.base = float(np.float32(attrs.get("base_values", [0.0])[0]))
        self.trees = []
        weights = {}
        for t, n, w in zip(attrs["target_treeids"], attrs["target_nodeids"], attrs["target_weights"]):
            weights[(int(t), int(n))] = weights.get((int(t), int(n)), 0.0) + float(np.float32(w))
        for t in sorted(set(tree.tolist())):
            sel = np.flatnonzero(tree == t)
            nodes = {}
            for i in sel:
                mode = attrs["nodes_modes"][i]
                if mode not in ("BRANCH_LEQ", "LEAF"):
                    raise ValueError(f"unsupported node mode {mode}")
                if int(attrs["nodes_featureids"][i]) != 0:
                    raise ValueError("the model must use a single feature")
                nodes[int(node[i])] = (mode == "LEAF", np.float32(attrs["nodes_values"][i]),
                                       int(attrs["nodes_truenodeids"][i]), int(attrs["nodes_falsenodeids"][i]),
                                       weights.get((int(t), int(node[i])), 0.0))
            self.trees.append(nodes)

    @classmethod
    def from_onnx(cls, path: str) -> "TreeEnsemble":
        with open(path, "rb") as fh:
            model = _fields(fh.read())
        graph = next(v for f, _, v in model if f == 7)
        for f, _, v in _fields(graph):
            if f == 1 and any(ff == 4 and x == b"TreeEnsembleRegressor" for ff, _, x in _fields(v)):
                return cls(_attributes(v))
        raise ValueError(f"{path} holds no TreeEnsembleRegressor node")

    def predict_one(self, x) -> float:
        """One prediction as onnxruntime's TreeEnsembleRegressor makes it for a float32 input
        (sstar/quantile_regression.py:106-118 feeds float32 and widens the float32 result): the leaf
        weights are summed in float32, in tree order, the base value is added in float32."""
        x = np.float32(x)
        total = np.float32(0.0)
        for nodes in self.trees:
            n = 0
            while True:
                leaf, thr, tr, fa, w = nodes[n]
                if leaf:
                    break
                n = tr if x <= thr else fa
            total = np.float32(total + np.float32(w))
        return float(np.float32(total + np.float32(self.base)))

    def lookup_table(self, max_value: int) -> np.ndarray:
        """Predictions for the feature values 0 .. max_value (float64)."""
        return np.asarray([self.predict_one(v) for v in range(int(max_value) + 1)], dtype=np.float64)

    def predict(self, x) -> np.ndarray:
        x = np.asarray(x)
        uniq, inv = np.unique(x, return_inverse=True)
        return np.asarray([self.predict_one(v) for v in uniq], dtype=np.float64)[inv].reshape(x.shape)


def infer(data: str, model: str, output: str, bed_file: str) -> None:
    """File-based step with the reference's signature and pandas semantics
    (sstar/quantile_regression.py:71-135): predictions for rows with a score, the full table sorted by
    (Sample, Chromosome, Start, End), and the BED of regions whose score exceeds the prediction."""
    import pandas as pd
    df = pd.read_csv(data, sep="\t")
    mask = df["S*_score"].notna() & df["Region_ind_SNP_number"].notna()
    df[PRED_COLUMN] = np.nan
    ens = TreeEnsemble.from_onnx(model)
    df.loc[mask, PRED_COLUMN] = ens.predict(df.loc[mask, "Region_ind_SNP_number"].to_numpy(dtype=np.float32))
    df.sort_values(by=["Sample", "Chromosome", "Start", "End"]).to_csv(output, sep="\t", index=False, na_rep="NA")
    bed = df.loc[df["S*_score"] > df[PRED_COLUMN], ["Chromosome", "Start", "End", "Sample"]].copy()
    bed["Start"] = bed["Start"] - 1
    bed.sort_values(by=["Sample", "Chromosome", "Start", "End"]).to_csv(bed_file, sep="\t", index=False, header=False)


def infer_table_device(table, model: str, output: str, bed_file: str, device: int = 0) -> None:
    """The same two files straight from a scored :class:`FeatureTable`, printed by the device emitter
    (sample-major row order, prediction column from the lookup table, BED rows selected on the
    device).  One chromosome per table, so (Sample, Chromosome, Start, End) order is (sample label
    as a string, window)."""
    from .engine import get_engine
    from .feature_table import COLUMNS
    eng = get_engine(device)
    torch = eng.torch
    ens = TreeEnsemble.from_onnx(model)
    lut = ens.lookup_table(int(table.nsnp.max()) if table.nsnp.size else 0)
    order = np.asarray(sorted(range(len(table.labels)), key=lambda t: table.labels[t]), dtype=np.int32)
    with torch.cuda.device(eng.tdev):
        up = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a, dtype=dt)).to(eng.tdev)
        d = (up(table.score, np.int64), up(table.nsnp, np.int32), up(table.win_start, np.int32), up(table.win_end, np.int32))
        order_d = up(order, np.int32)
        for path, bed in ((output, False), (bed_file, True)):
            text_d, total, status = eng.format_tsv_device(*d, str(table.chr_name), table.labels, col_order_d=order_d,
                                                          predictions=lut, bed=bed, sample_major=True)
            if status:
                raise RuntimeError(f"device emitter status {status}: format this table on the host (infer())")
            body = text_d[:total].cpu().numpy().tobytes()
            out_dir = os.path.dirname(str(path))
            if out_dir:
                os.makedirs(out_dir, exist_ok=True)
            with open(path, "wb") as fh:
                if not bed:
                    fh.write(("\t".join(COLUMNS + [PRED_COLUMN]) + "\n").encode())
                fh.write(body)
