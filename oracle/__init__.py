"""CPU oracle for the turboPy particle hot path -- TEST INFRASTRUCTURE ONLY.

This package restates, in plain numpy, the
arithmetic of the reference's hot path so that the CUDA kernels can be checked
bit-for-bit:

* ``oracle.boris``   -- ``BorisPush.push``            (turbopy/computetools.py:407-441)
* ``oracle.gather``  -- linear field gather, formulas A / B / C
                        (turbopy/core.py:711-755, turbopy/computetools.py:459-481
                        -> scipy.interpolate.interp1d -> numpy.interp)
* ``oracle.fd``      -- ``FiniteDifference`` stencils (turbopy/computetools.py:104-138,189-223)
* ``oracle.moments`` -- kinetic-energy / momentum sums (no reference function:
                        PARITY UNPINNED for that one, see its docstring)
* ``oracle.grid``    -- ``Grid`` coordinates          (turbopy/core.py:610-673,709)

Pinning: every function here is checked against the *imported, unmodified*
reference (``/root/reference`` + the shims in ``oracle/shims``) by
``oracle/make_golden.py`` in the build container, and the resulting vectors are
committed under ``tests/golden``; ``tests/test_oracle_*.py`` re-check the
oracle against those vectors and against the reference's own known-answer
tests (tests/test_computetools.py:15-62, tests/test_core.py:164-170).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this package.  The product
(``turbopy_b200``) never does: it has no CPU fallback and raises if its CUDA
library is missing.
"""
