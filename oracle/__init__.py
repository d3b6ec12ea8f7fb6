"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's unet-only hot path.

Nothing in ``openglottal_b200`` imports this package. Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``
may use it, and only as the checker or the timed CPU baseline -- never as the product path.

Parity pinning: the reference repository ships no tests and no golden vectors (and its
weights are absent from the snapshot), so the oracle is pinned against outputs of the
reference code itself, run in the build container by ``tests/golden/make_golden.py`` and
committed under ``tests/golden/`` ("Oracle").
"""
