"""The match-rate oracle and the host helpers of sstar_b200.match against the known answers of the
reference's own tests (tests/test_match.py).  CPU only."""
import numpy as np
import pytest

from oracle import match_port


@pytest.mark.parametrize("impl", ["oracle", "product"])
def test_calc_match_pct_known_answers(impl):
    # tests/test_match.py:67-77
    if impl == "oracle":
        f = match_port.calc_match_pct
    else:
        from sstar_b200.match import calc_match_pct as f
    assert f([0, 1, 2], [0, 1, 2], denominator=3, P=2) == 1.0
    assert f([0], [1], denominator=1, P=2) == 0.5
    assert f([0], [2], denominator=1, P=2) == 0.0
    assert f([0, 1, 2], [0, 2, 2], denominator=3, P=2) == pytest.approx((1.0 + 0.5 + 1.0) / 3)
    assert f([0, -1], [0, 2], denominator=2, P=2) == 0.5
    assert f([-1], [-1], denominator=1, P=2) == 0.0
    assert f([], [], denominator=0, P=2) == "NA"


def test_oracle_match_known_answers():
    # tests/test_match.py:187-262 (match, target haplotypes) and :315-352, :415-433 (run_match inputs)
    pos = np.array([100, 200, 300])
    gt = np.array([[[0, 0], [0, 0]], [[0, 1], [1, 1]], [[1, 1], [1, 1]]])
    assert match_port.tract_rate(gt, pos, 0, None, [1], 0, 300) == pytest.approx((1.0 + 0.5 + 1.0) / 3)
    gt2 = np.array([[[0, 1], [0, 0]], [[1, 0], [1, 1]]])
    assert match_port.tract_rate(gt2, [100, 200], 0, 0, [1], 0, 200) == 1.0
    assert match_port.tract_rate(gt2, [100, 200], 0, 1, [1], 0, 200) == 0.0
    mini = np.array([[[0, 0], [0, 0], [1, 1]], [[0, 1], [1, 1], [0, 1]], [[1, 1], [1, 1], [-1, -1]]])
    assert match_port.tract_rate(mini, pos, 0, 0, [1, 2], 0, 250) == pytest.approx(0.375)
    assert match_port.tract_rate(mini, pos, 0, 1, [1, 2], 0, 250) == pytest.approx(0.625)
    assert match_port.tract_rate(mini, pos, 0, None, [1, 2], 0, 250) == pytest.approx(0.625)
    assert match_port.tract_rate(mini, pos, 0, 0, [1, 2], 400, 500) == "NA"
    assert match_port.dosage(mini, pos, 2, [100, 300, 999]).tolist() == [2.0, -1.0, -1.0]


def test_host_helpers_known_answers():
    # tests/test_match.py:264-312
    from sstar_b200.match import (_base_sample_name, _positions_in_bed_interval, _resolve_chrom,
                                  _validate_disjoint_samples, _validate_sorted_positions)
    samples = ["ind1", "ind2"]
    assert _base_sample_name("ind1", samples) == ("ind1", None)
    assert _base_sample_name("ind1_1", samples) == ("ind1", 0)
    assert _base_sample_name("ind1_2", samples) == ("ind1", 1)
    with pytest.raises(ValueError):
        _base_sample_name("missing_1", samples)
    data = {"21": {}, "chr22": {}}
    assert _resolve_chrom("21", data) == "21" and _resolve_chrom("chr21", data) == "21"
    assert _resolve_chrom("22", data) == "chr22"
    with pytest.raises(ValueError):
        _resolve_chrom("23", data)
    _validate_disjoint_samples(["ind1", "ind2"], ["src1", "src2"])
    with pytest.raises(ValueError, match="Target and source sample lists must not overlap"):
        _validate_disjoint_samples(["ind1", "ind2"], ["src1", "ind1"])
    _validate_sorted_positions({"1": {"POS": np.array([100, 200, 300])}})
    with pytest.raises(ValueError, match="VCF positions must be sorted within chromosome 1"):
        _validate_sorted_positions({"1": {"POS": np.array([100, 300, 200])}})
    np.testing.assert_array_equal(_positions_in_bed_interval(np.array([100, 200, 300, 400]), 100, 300), [200, 300])
