"""TEST INFRASTRUCTURE.  Import the UNMODIFIED reference (read-only /root/reference) under
the `ocnn` shim.  Only usable in the build container: /root/reference does not exist on the
GPU box, and nothing under `-m gpu`, smoke() or bench.py calls this.

`skimage` and `trimesh` are imported at module top by reference
models/networks/diffusion_networks/ldm_diffusion_util.py:11-12 but never used on the U-Net
path, so they are stubbed.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get('OCTFUSION_REFERENCE', '/root/reference')
_SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'ocnn_shim')


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, 'models', 'networks'))


def _stub(name, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def ensure_shim():
    if _SHIM not in sys.path:
        sys.path.insert(0, _SHIM)
    import ocnn  # noqa: F401  (the shim)
    return sys.modules['ocnn']


def load():
    """Returns a namespace with the reference's own modules (unchanged code)."""
    if not available():
        raise RuntimeError('reference tree not present at %s' % REFERENCE_ROOT)
    ensure_shim()
    sk = _stub('skimage')
    sk.measure = _stub('skimage.measure')
    _stub('trimesh')
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    from models.networks import modules as ref_modules
    from models.networks.diffusion_networks import graph_unet_union, graph_unet_hr, graph_unet_lr
    from models.networks.diffusion_networks import ldm_diffusion_util
    from models.networks.dualoctree_networks import dual_octree
    ns = types.SimpleNamespace(modules=ref_modules, union=graph_unet_union, hr=graph_unet_hr,
                               lr=graph_unet_lr, util=ldm_diffusion_util, dual_octree=dual_octree)
    return ns


def load_util():
    """reference utils/util_dualoctree.py (split <-> octree helpers of the stage-1 -> stage-2 handoff).  Its module top
    imports plotting / mesh packages that the handoff functions never touch: stubbed."""
    load()
    mpl = _stub('matplotlib', use=lambda *a, **k: None)
    mpl.pyplot = _stub('matplotlib.pyplot')
    _stub('plyfile', PlyData=object, PlyElement=object)
    import importlib
    return importlib.import_module('utils.util_dualoctree')
