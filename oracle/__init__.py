"""CPU oracle for the tiny-faces hot path -- TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is product code.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it, and only as the checker or the CPU baseline.

Parity pinning: the reference ships no golden vectors for this path (its two
tests need never-committed .mat files), so the oracle is
pinned against the reference *itself*: ``oracle/make_golden.py`` imports
``/root/reference/tinyfaces`` in the dev container, runs it on seeded
synthetic inputs and commits the inputs/outputs under ``tests/golden/``;
``tests/test_oracle_golden.py`` checks every oracle function against them.
"""
