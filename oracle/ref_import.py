"""TEST INFRASTRUCTURE -- import the UNMODIFIED reference (read-only at /root/reference) in
the build container so that golden vectors can be generated from it.

Nothing in tests/, bench.py or the package imports this module: /root/reference does not
exist on the GPU box. ``make_goldens.py`` is its only user.

Recipe: empty stub modules for the reference's unbuilt C++ corner
pooling extensions and for optional packages this image lacks, plus the ``numpy.int`` alias the
reference still uses (pipeline.py:161,162,168).
"""
import os
import sys
import types
import warnings

import numpy as np

REFERENCE_ROOT = os.environ.get('OKP_REFERENCE_ROOT', '/root/reference')


def load():
    """Returns (perception.pipeline, perception.utils.camera_utils, perception.datasets.video)."""
    if not os.path.isdir(REFERENCE_ROOT):
        raise RuntimeError(f"reference not mounted at {REFERENCE_ROOT}")
    for name in ['top_pool', 'bottom_pool', 'left_pool', 'right_pool', 'timm', 'h5py',
                 'skvideo', 'skvideo.io', 'albumentations', 'hud']:
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules['skvideo'].io = sys.modules['skvideo.io']
    if not hasattr(np, 'int'):
        np.int = int
    sys.dont_write_bytecode = True          # the mount is read-only
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        import perception.pipeline as pipeline
        from perception.utils import camera_utils
        from perception.datasets import video
    return pipeline, camera_utils, video
