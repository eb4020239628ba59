"""CPU oracle for the S* hot path -- TEST INFRASTRUCTURE, NOT PRODUCT.

Everything under ``oracle/`` is a CPU restatement of the reference algorithm
(xin-huang/sstar, ``sstar2`` 2.0.0) used only as the *checker*:

* ``tests/``                       -- parity tests compare the CUDA path to it,
* ``__graft_entry__.smoke()``      -- one small GPU invocation checked against it,
* ``bench.py``                     -- the ``cpu_baseline`` leg and ``--impl reference``.

Nothing in ``sstar_b200/`` may import, call, link or execute anything from this
package; the product path fails loudly when the CUDA library is missing.

Parity status: PINNED.  ``oracle.sstar_port`` reproduces every golden vector the
reference holds for this path (``tests/test_sstar.py`` KATs, the 20 + 40 row
golden TSVs of ``tests/test_preprocess.py``, the 8 rows of ``tests/test_infer.py``)
and additionally matches outputs of the reference's own ``Sstar.compute`` /
``FeatureVectorPreprocessor.run`` imported in the build container and frozen as
fixtures under ``tests/golden/`` (script: ``tests/golden/make_golden.py``).
"""
