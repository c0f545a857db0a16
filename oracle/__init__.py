"""CPU oracle for the MoMA criterion step -- TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` may import it, and there only as the checker / the CPU
arm being timed.  The product path (``moma_b200``) never imports this package
and fails loudly when its CUDA library is missing.

Parity pin: every function is checked against golden vectors produced by the
*unmodified* reference modules (``/root/reference/MoMA``, ``learning``) run in
the build container -- see ``tests/golden/make_golden.py`` (generator) and
``tests/test_oracle_golden.py`` (the pin).
"""
