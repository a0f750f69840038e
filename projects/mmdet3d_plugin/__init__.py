"""Drop-in plugin package: importing it registers the same ``type`` strings the reference's
``projects/mmdet3d_plugin/__init__.py:1-9`` registers (configs set ``plugin_dir='projects/mmdet3d_plugin/'`` and the
launcher ``importlib.import_module``s it, reference ``tools/test.py:138-150``).  The classes are the B200-native
implementations in ``focalformer3d_b200.model``."""
from focalformer3d_b200.model import (FocalFormer3D, FocalEncoder, FocalDecoder, TransFusionBBoxCoder,  # noqa: F401
                                      SparseEncoder, SECOND, SECONDFPN, HardSimpleVFE, build_model)
from focalformer3d_b200.config import DETECTORS, NECKS, HEADS, BBOX_CODERS  # noqa: F401
