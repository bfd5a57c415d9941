class InputFeature:  # VAE-only (dual_octree.py:345); outside the U-Net hot path
    def __init__(self, *a, **k):
        pass

    def __call__(self, octree):
        raise NotImplementedError('ocnn shim: InputFeature is outside the hot path')
