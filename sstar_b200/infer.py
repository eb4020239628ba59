"""``infer(...)`` with the reference's signature (sstar/infer.py:28-93): feature table -> predicted
threshold -> tracts whose S* exceeds it.

Only the ``preprocessing`` block of the YAML config is read (plain ``yaml.safe_load``); the
reference's pydantic models and duplicate-key loader are configuration plumbing outside the hot
path and stay with the reference.
"""
from __future__ import annotations

import yaml

from . import quantile_regression as qr
from .preprocess import preprocess


def infer(model: str, config: str, feat_file: str, pred_file: str, tract_file: str, match_bonus: int,
          max_mismatch: int, mismatch_penalty: int) -> None:
    try:
        with open(config, "r") as f:
            config_dict = yaml.safe_load(f)
    except FileNotFoundError:
        raise FileNotFoundError(f"Configuration file '{config}' not found.")
    except yaml.YAMLError as e:
        raise ValueError(f"Error parsing YAML configuration file '{config}': {e}")
    preprocess(output_file=str(feat_file), match_bonus=match_bonus, max_mismatch=max_mismatch,
               mismatch_penalty=mismatch_penalty, **config_dict["preprocessing"])
    qr.infer(data=str(feat_file), model=model, output=str(pred_file), bed_file=str(tract_file))
