"""CPU oracle for the mycelium FEA hot path -- TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs may import it, and there only as the checker or the timed CPU baseline.
The product package (``mycelium_fea_project_b200``) never imports this package and
raises if its CUDA library is missing.

Parity status: PINNED.  ``fea_oracle`` is checked (tests/test_oracle_golden.py)
against the reference's committed goldens ``results/test_{I,X,t,y}`` (all four
output CSVs, bit-equal) and ``results/sim_20251117_181147`` (active cascade
bit-equal, force-displacement to 1e-12 relative), and -- when ``/root/reference``
is present -- against the reference module itself, imported unmodified
(``oracle.ref_shim``).  Intermediate quantities (K_e, CSR, BC sets, U) are pinned
by ``tests/golden/*.npz`` generated from the imported reference by
``tests/golden/make_golden.py``.
"""
