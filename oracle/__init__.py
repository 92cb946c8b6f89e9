"""CPU oracle for the Poseidon-BN254 / poll-tree hot path of rhysbalevicius/infimum.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import, link or execute it, and
there only as the checker (or as the timed CPU baseline), never as the thing
shipped.  The product path (``infimum_b200``) is CUDA-only and raises when its
extension is missing.

Parity status: PINNED.  ``tests/test_oracle_golden.py`` checks this oracle
against every vector the reference's own tests hold for the path (see
``tests/golden/reference_vectors.json`` and ``tests/golden/extract_reference_vectors.py``).
The Rust reference itself cannot be built here (no cargo/rustc, ark-ff 0.4.2 /
ark-bn254 0.4.0 are not vendored), so ``oracle/_ref`` does not exist; the CPU
baseline kind is "port".
"""
