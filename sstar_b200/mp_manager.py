"""Seam S2: the scheduler (reference sstar/mp_manager.py:98-188).

Same call, same return convention (a flat ``list`` of row dicts, or the string ``"error"``), but
the fan-out is over GPUs instead of worker processes: a ``FeatureVectorPreprocessor`` job fed by a
``WindowDataGenerator`` is scored as ONE batched device call per GPU (windows sharded by range,
``SSTAR_GPUS`` selects the devices) -- no queues, no pickling, no 5 s worker-exit tail
(mp_manager.py:238-242).  Any other job / generator pair keeps the reference's task-by-task
semantics in-process.  ``nprocess`` is accepted for signature compatibility and ignored.
"""
from __future__ import annotations

from .feature_vector_preprocessor import FeatureVectorPreprocessor
from .generators import GenericGenerator


def mp_manager(job, data_generator: GenericGenerator, nprocess: int = 1, **kwargs):
    try:
        if isinstance(job, FeatureVectorPreprocessor) and hasattr(data_generator, "chromosome"):
            table = job.run_chromosome(**data_generator.chromosome(), **kwargs)
            return table.rows()
        res = []
        for params in data_generator.get():
            items = job.run(**params, **kwargs)
            if items is None:
                continue
            res.extend(items)
        return res
    except Exception as e:  # the reference's worker puts ("error", str(e)) on the out queue
        print(f"sstar_b200 scheduler: job did not complete successfully ({e}). Initiating shutdown.")
        print("All workers are terminated.")
        return "error"
