"""CPU oracle for the BeNeRF render-and-image-formation path.

TEST INFRASTRUCTURE ONLY.  This package is a torch-CPU fp32 restatement of the
reference algorithm (WU-CVGL/BeNeRF @ 72cab91) with every random draw turned
into an explicit input.  It exists so that the CUDA engine in ``benerf_b200``
can be checked on machines where ``/root/reference`` is absent.  Only
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it; the product package
never does (tests/test_library.py greps for that).

Parity pin: the oracle is pinned against outputs of the *unmodified* reference
imported from ``/root/reference`` (tools/make_golden.py -> tests/golden/*.npz;
tests/test_oracle_golden.py re-checks the committed fixtures everywhere and,
where the reference tree is present, re-runs it live).  The reference ships no
tests or golden vectors of its own.

Each function cites the reference file:line it follows.
"""
from . import pose, rays, encode, mlp, composite, resample, render, image_formation, events  # noqa: F401
