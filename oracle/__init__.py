"""CPU oracle for the RePlay sequential-recommender hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in ``replay_b200`` may import this package; only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs use it, and there only as
the checker (or as the timed CPU baseline), never as the product.

The reference (sb-ai-lab/RePlay @ b4e051e8) is pure Python/PyTorch on this path, so the oracle is a plain-torch
restatement (fp32 or fp64 on CPU) of the reference's algorithm; every function cites the reference file:line it follows.

Parity pinning: the reference's own tests hold no arithmetic golden vectors for this path, so the
oracle is pinned against the reference ITSELF, imported in the build container through ``oracle/shim`` (stub ``polars`` /
``lightning``): ``oracle/gen_golden.py`` runs the real reference classes on seeded inputs and commits inputs, weights and
outputs under ``tests/golden/``; ``tests/test_oracle_golden.py`` asserts the restatement reproduces them.
"""
