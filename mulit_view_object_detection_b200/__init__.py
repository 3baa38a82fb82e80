"""mulit_view_object_detection_b200 -- B200 (sm_100a) implementation of the multi-view fusion hot
path of juliuserbach/mulit_view_object_detection: unproject -> fuse across views -> project ->
PyramidROIAlign -> per-class NMS, behind the reference's own layer signatures.

``libmvfusion.so`` is loaded on the first access to any layer / binding attribute of this package and
that access raises ImportError when the library has not been built -- there is no CPU or pure-PyTorch
fallback.  The two pure-host modules (``config``: the attribute names of the reference's Config;
``synthetic``: the seeded input generator) import without the library, so that the CPU baseline arm
of bench.py never maps it.
"""
import importlib
import os

from .config import FusionConfig

LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libmvfusion.so")

_LIB_NAMES = ("launch_count", "version")
_LAYER_NAMES = (
    "unproj_feat", "unproj_feat_notebook", "grid_reas", "convlstm", "convlstm_step", "ConvLSTMTensorCore", "Conv3dTensorCore",
    "unet_fuse", "unproject_unet_fuse", "unproject_ident_fuse", "proj_grid", "depth_sampling", "depth_sampling_conv3d",
    "proj_grid_depth_sampling", "PyramidROIAlign", "refine_detections_graph", "DetectionLayer", "ProposalLayer",
    "non_max_suppression", "unproject_fuse", "unproject_fuse_project", "fusion_neck", "prepare_params", "view_reduce",
    "channel_mean", "HostPipeline", "set_weights", "weights", "reused_lay",
)

__all__ = ["FusionConfig", "LIB_PATH", "MaskRCNN", *_LIB_NAMES, *_LAYER_NAMES]


def __getattr__(name):
    # PEP 562: the first use of the product API loads the CUDA library (and fails loudly without it)
    if name in _LIB_NAMES:
        return getattr(importlib.import_module("._lib", __name__), name)
    if name in _LAYER_NAMES:
        return getattr(importlib.import_module(".layers", __name__), name)
    if name == "MaskRCNN":
        return importlib.import_module(".model", __name__).MaskRCNN
    if name in ("_lib", "layers", "dist", "weights_io", "synthetic", "config", "model", "model_host"):
        return importlib.import_module("." + name, __name__)
    raise AttributeError("module %r has no attribute %r" % (__name__, name))
