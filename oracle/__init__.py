"""CPU oracle for the hybrid-rollout hot path of gnn-plasma-flux.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it, and there only as the
checker (or as the timed CPU baseline), never as the thing shipped.  The
product package ``gnn_plasma_flux_b200`` never imports this package and fails
loudly when its CUDA library is missing.

Parity status: PINNED against the live reference.  ``oracle/make_golden.py``
imports the unmodified reference from ``/root/reference`` (possible only in the
build container), checks every function of ``oracle/ref_port.py`` against it on
seeded inputs (bit-exact for the fp32 port) and freezes the reference's own
outputs as fixtures under ``tests/golden/``.  The reference ships no golden
vectors or known-answer tests of its own (its ``examples/smoke_test.py`` checks
shapes only), so these generated fixtures are the pin.
"""
