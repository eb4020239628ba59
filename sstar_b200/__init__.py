"""sstar_b200 -- B200-native S* scoring engine behind the Python seams of xin-huang/sstar.

Only what the hot path needs lives here:

* ``csrc/``                       CUDA kernels + the C-ABI (``include/sstar_b200.h``)
* ``_cabi``                       ctypes loader (no fallback)
* ``engine.ScoreEngine``          device buffers (PyTorch owns them) + H2D/D2H + the C-ABI calls
* ``sstar.Sstar``                 seam S3, ``Sstar.compute(**kwargs)``           (sstar/sstar.py:43)
* ``feature_vector_preprocessor`` the mp "job" + row formatting                   (feature_vector_preprocessor.py:34)
* ``generators``                  window tiling                                  (generators/window_data_generator.py:30)
* ``mp_manager``                  seam S2, GPU-sharded scheduler                 (mp_manager.py:98)
* ``preprocess``                  seam S1, ``preprocess(vcf_file, ...)`` -> TSV  (preprocess.py:28)
* ``simulate`` / ``msprime_simulator`` / ``simulate_batch``  the train-path tiler, replicate-batched (simulate.py:30)
"""
__version__ = "0.1.0"

from ._cabi import NAN_SENTINEL, SstarB200Error  # noqa: F401
