"""mulit_view_object_detection_b200 -- B200 (sm_100a) implementation of the multi-view fusion hot
path of juliuserbach/mulit_view_object_detection: unproject -> fuse across views -> project ->
PyramidROIAlign -> per-class NMS, behind the reference's own layer signatures.

Importing this package loads ``libmvfusion.so`` and raises ImportError when it has not been
built -- there is no CPU or pure-PyTorch fallback.
"""
from .config import FusionConfig
from . import _lib
from ._lib import LIB_PATH, launch_count, version
from .layers import (unproj_feat, unproj_feat_notebook, grid_reas, convlstm, convlstm_step, ConvLSTMTensorCore, Conv3dTensorCore, unet_fuse, unproject_unet_fuse, unproject_ident_fuse,
                     proj_grid, depth_sampling, depth_sampling_conv3d, proj_grid_depth_sampling, PyramidROIAlign, refine_detections_graph,
                     DetectionLayer, ProposalLayer, non_max_suppression, unproject_fuse,
                     unproject_fuse_project, fusion_neck, prepare_params, view_reduce, channel_mean, HostPipeline, set_weights, weights, reused_lay)

__all__ = [
    "FusionConfig", "LIB_PATH", "launch_count", "version",
    "unproj_feat", "unproj_feat_notebook", "grid_reas", "convlstm", "convlstm_step", "ConvLSTMTensorCore", "Conv3dTensorCore",
    "unet_fuse", "unproject_unet_fuse", "unproject_ident_fuse", "proj_grid", "depth_sampling", "depth_sampling_conv3d", "proj_grid_depth_sampling", "PyramidROIAlign", "refine_detections_graph",
    "DetectionLayer", "ProposalLayer", "non_max_suppression", "unproject_fuse",
    "unproject_fuse_project", "fusion_neck", "prepare_params", "view_reduce", "channel_mean", "HostPipeline", "set_weights", "weights", "reused_lay",
]
