"""TEST INFRASTRUCTURE ONLY -- the CPU oracle for the VL-CABS similarity path.

Nothing under ``oracle/`` is part of the shipped product.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import it, and there only as the checker / the CPU baseline, never
as the thing measured or shipped.  The product (``radzero_b200``) calls hand-written
sm_100a CUDA through the C-ABI in ``include/rz_b200.h`` and raises if that library
is missing.

Parity pinning: the reference (deepnoid-ai/RadZero) has no tests and no golden
vectors of its own.  The oracle is therefore pinned by
executing the reference's own ``exp/cxr_pt/model/losses.py`` and ``F.interpolate``
call sites in the build container (``tests/test_oracle_vs_reference.py``, skipped
where ``/root/reference`` is absent) and by the golden fixtures generated from the
reference by ``tests/golden/make_golden.py`` (committed, checked everywhere).

``oracle/align.py`` restates the AlignTransformer forward, whose
body is third-party code (transformers ``Dinov2Encoder``, pinned 4.39.0 by the reference, not
vendored): it is pinned against the installed transformers module and the committed golden
vectors ``tests/golden/align_golden.npz`` (``tests/test_oracle_align.py``).
"""
from .vlcabs import *  # noqa: F401,F403
