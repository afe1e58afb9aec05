"""occformer_b200 -- B200 (sm_100a) native hot path of OccFormer behind the reference's plugin surface.

Registry names / constructor kwargs / forward signatures / state_dict keys follow
projects/mmdet3d_plugin of the reference; all arithmetic runs in csrc/libocc_b200.so (C ABI, ctypes).
"""
from .registry import ATTENTION, BACKBONES, HEADS, NECKS, POSITIONAL_ENCODING  # noqa: F401
from .encoder import DualpathTransformerBlock, OccupancyEncoder  # noqa: F401
from .view_transformer import ViewTransformerLiftSplatShootVoxel, bev_pool  # noqa: F401
from .neck import MSDeformAttnPixelDecoder3D, MultiScaleDeformableAttention3D  # noqa: F401
from .head import (Mask2FormerNuscOccHead, Mask2FormerNuscPanopticOccHead, Mask2FormerOccHead,  # noqa: F401
                   SinePositionalEncoding3D)

__version__ = "0.1.0"
