"""TEST INFRASTRUCTURE ONLY -- a CPU restatement of the handful of `ocnn`
(ocnn-pytorch 2.2.x, third party, un-vendored, un-pinned: reference
requirements.txt:1) symbols that the OctFusion hot path touches.  It exists so
that the *reference's own* modules under /root/reference can be imported and run
as the oracle in this container, and so that `oracle/` has an octree container.

parity unpinned at the `ocnn` boundary: nothing under /root/reference fixes the
Morton bit order or the Octree field layout; the semantics below are taken from
how the reference *uses* each symbol (call sites cited per function).
The product package (`octfusion_b200`) never imports this.
"""
from . import octree, nn, utils, modules, dataset  # noqa: F401
