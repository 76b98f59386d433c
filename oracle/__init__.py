"""CPU oracle for the 3D N^2-queens annealing hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it, and there only as the
checker or as the timed CPU baseline -- never as the thing shipped.  The
product path (``monte_carlo_collective_b200``) fails loudly when its CUDA
library is missing; it never falls back to this code.

Parity status: PINNED.  ``oracle/queens_numpy.py`` is checked against golden
vectors produced by running the real reference (``/root/reference``) in the
authoring container with ``oracle/gen_golden.py`` (committed under
``tests/golden/``), plus the known-answer values of SURVEY.md section 8(c).
"""
