"""xtag_clip_b200 -- B200 (sm_100a) implementation of XTag-CLIP's data-parallel hot path.

Host-side mirror of the reference's Python API for that path (same names, argument meaning, error
behaviour); all arithmetic runs in hand-written CUDA kernels reached through the C ABI of
``libxtag_b200.so`` (``include/xtag_b200.h``).  There is no CPU fallback: without the built
library or without a CUDA device the ops raise.

    from xtag_clip_b200 import ClipLoss, create_loss, AsymmetricLoss, TagHead, l2_normalize
"""
from .loss import ClipLoss, gather_features, create_loss                  # noqa: F401
from .siglip import SigLipLoss                                             # noqa: F401
from .asymmetric_loss import AsymmetricLoss                               # noqa: F401
from .tag_head import TagHead, l2_normalize, cross_attention, patch_reference_model   # noqa: F401
from .fusion_head import FusionHead, DQNCOSLoss, fusion_scores                        # noqa: F401

__version__ = "0.1.0"
