"""CPU oracle for the TomoSAR2Height dual-topology hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in ``tomosar2height_b200/`` may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` do, and there only as the checker.

Parity status: the reference ships no tests and no golden vectors for this
path, and its arithmetic lives in un-vendored third-party code
(torch_scatter 2.1.x, ATen).  The oracle is pinned two ways instead:

* ``tests/golden/make_golden.py`` imports the real reference model from
  ``/root/reference`` (with ``sys.modules`` stubs for the absent IO packages and
  a naive loop implementation of torch_scatter's documented CPU rules), runs
  forward+backward, and stores the results as fixtures under ``tests/golden/``.
  ``tests/test_oracle_golden.py`` checks this oracle against them.
* the explicit bilinear restatements are checked against ATen's own
  ``F.grid_sample`` / ``F.interpolate`` (the reference's actual dependency).

The torch_scatter semantics themselves (first-index tie rule, empty -> 0 /
arg = N) are *recalled* from upstream 2.1.x, anchored only by the reference's
single printed vector (pointnet.py:114-123) => for the scatter ops parity is
"pinned to the reference's call sites, unpinned against torch_scatter binaries".
"""
from .ops import (  # noqa: F401
    cell_index,
    segment_max,
    segment_mean,
    bilinear_sample_points,
    upsample_bilinear_align,
)
from .model import oracle_forward, oracle_loss, synth_state_dict, reference_param_shapes, Selections  # noqa: F401
