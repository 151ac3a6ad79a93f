"""TEST INFRASTRUCTURE ONLY -- CPU oracle for morna's hot path.

A CPU restatement of the reference algorithm (commanderson/morna, morna.py) used
as the *checker* for the CUDA path.  Only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it.
Nothing under ``morna_b200/`` imports this package: the product path has no CPU
fallback and fails loudly when the CUDA library is missing.

Parity pin: the oracle is checked against the golden neighbour lists embedded in
the reference's own unit tests (morna.py:1176-1187, 1267-1278, 1312-1323) and the
fixture ``tests/tiny_intropolis.tsv`` -- see ``tests/test_oracle_golden.py``.
"""
