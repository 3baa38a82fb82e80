"""ctypes binding of libmvfusion.so (include/mvfusion.h).

There is NO fallback: if the shared library is missing the import raises, and every wrapper
raises if it is handed a tensor that is not a contiguous CUDA tensor.  Build the library with
``python -c "import __graft_entry__ as g; g.build()"`` or ``make -C mulit_view_object_detection_b200/csrc``.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmvfusion.so")

MVF_OK = 0
MVF_EINVAL, MVF_ENULL, MVF_EALIGN, MVF_ECUDA, MVF_EUNSUPPORTED, MVF_EWORKSPACE = -1, -2, -3, -4, -5, -6
FUSE_NONE, FUSE_SUM, FUSE_MEAN, FUSE_MAX = 0, 1, 2, 3
FLAG_RELU_IN, FLAG_RELU_OUT, FLAG_WORLD_GRID, FLAG_PRESPLIT = 1, 2, 4, 8
CONV_S1, CONV_S2, DECONV_S2 = 0, 1, 2
MAX_VIEWS, MAX_DIM, MAX_SAMPLES, MAX_NMS_BOXES, MAX_CLASSES = 32, 192, 64, 8192, 256


class MvfGrid(C.Structure):
    _fields_ = [("nvox", C.c_int32), ("nvox_z", C.c_int32),
                ("vmin", C.c_double), ("vmax", C.c_double), ("vsize", C.c_double),
                ("vmin_z", C.c_double), ("vmax_z", C.c_double), ("vsize_z", C.c_double)]


def grid_from_config(cfg):
    """MvfGrid from the attributes the reference layers read off ``config``
    (mrcnn/model_multi.py:157-160, :267, :294-296)."""
    return MvfGrid(int(cfg.nvox), int(cfg.nvox_z), float(cfg.vmin), float(cfg.vmax), float(cfg.vsize),
                   float(cfg.vmin_z), float(cfg.vmax_z), float(cfg.vsize_z))


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "libmvfusion.so not found at %s -- the CUDA library is the product and there is no "
            "CPU fallback; build it with `python -c 'import __graft_entry__ as g; g.build()'`" % LIB_PATH)
    return C.CDLL(LIB_PATH)


lib = _load()

_p, _i, _ll, _d, _f, _sz = C.c_void_p, C.c_int, C.c_longlong, C.c_double, C.c_float, C.c_size_t
_G = C.POINTER(MvfGrid)

_SIGS = {
    "mvf_unproject_fuse": (_i, [_p, _p, _p, _p, _G, _i, _i, _i, _i, _i, _i, _i, _i, _i, _d, _i, _i,
                                _p, _p, _p, _p, _p, _p, _p]),
    "mvf_unproject_fuse_tc_supported": (_i, [_i, _i, _i, _i]),
    "mvf_unproject_fuse_tc_workspace_bytes": (_sz, [_i, _i, _i, _i, _i]),
    "mvf_unproject_fuse_tc": (_i, [_p, _p, _p, _p, _G, _i, _i, _i, _i, _i, _i, _i, _i, _i, _d, _i, _i,
                                   _p, _p, _p, _p, _sz, _p]),
    "mvf_unproject_split_f16": (_i, [_p, _p, _p, _p, _G, _i, _i, _i, _i, _i, _i, _i, _i, _i, _p, _p, _sz, _p]),
    "mvf_view_reduce": (_i, [_p, _i, _i, _ll, _i, _i, _i, _p, _p, _p, _p]),
    "mvf_channel_mean": (_i, [_p, _i, _i, _ll, _i, _p, _p]),
    "mvf_ident_fuse": (_i, [_p, _p, _p, _p, _p, _i, _i, _ll, _i, _i, _p, _p]),
    "mvf_conv3d_wsplit_bytes": (_sz, [_i, _i, _i, _i]),
    "mvf_conv3d_prepare": (_i, [_p, _i, _i, _i, _i, _i, _i, _i, _p, _p]),
    "mvf_conv3d_tc_workspace_bytes": (_sz, [_i, _i, _i, _i, _i, _i, _i, _i, _i, _i]),
    "mvf_conv3d_tc": (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _p, _p, _sz, _p, _p]),
    "mvf_ident_wsplit_bytes": (_sz, [_i, _i, _i]),
    "mvf_ident_prepare": (_i, [_p, _i, _i, _i, _p, _p]),
    "mvf_ident_tc_workspace_bytes": (_sz, [_i, _i, _i, _i, _i, _i, _i]),
    "mvf_ident_fuse_tc": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _p, _p, _sz, _p]),
    "mvf_convlstm_step": (_i, [_p, _p, _p, _p, _p, _f, _i, _i, _i, _i, _i, _i, _i, _p, _p, _p]),
    "mvf_convlstm_wsplit_bytes": (_sz, [_i, _i]),
    "mvf_convlstm_prepare": (_i, [_p, _i, _i, _p, _p]),
    "mvf_convlstm_tc_workspace_bytes": (_sz, [_i, _i, _i, _i, _i, _i]),
    "mvf_convlstm_step_tc": (_i, [_p, _p, _p, _p, _p, _f, _i, _i, _i, _i, _i, _i, _i, _p, _p, _p, _sz, _p]),
    "mvf_convlstm_step_tc_slab": (_i, [_p, _p, _p, _p, _p, _f, _i, _i, _i, _i, _i, _i, _i, _i, _i, _p, _p, _p, _sz, _p, _p]),
    "mvf_project_rays": (_i, [_p, _p, _p, _p, _p, _G, _i, _i, _i, _i, _i, _i, _i, _d, _i, _i, _p, _p, _p, _p]),
    "mvf_project_depth_collapse": (_i, [_p, _p, _p, _p, _p, _G, _i, _i, _i, _i, _i, _i, _i, _d, _i, _i,
                                        _p, _f, _f, _f, _p, _p]),
    "mvf_depth_collapse": (_i, [_p, _i, _i, _ll, _i, _p, _f, _f, _f, _i, _p, _p]),
    "mvf_pyramid_roi_align": (_i, [_p, C.POINTER(_p), C.POINTER(_i), C.POINTER(_i), _i, _i, _i, _i, _i, _i, _i,
                                   _p, _p, _p]),
    "mvf_nms_workspace_bytes": (_sz, [_i, _i]),
    "mvf_nms": (_i, [_p, _p, _p, _i, _i, _f, _i, _i, _p, _p, _p, _sz, _p]),
    "mvf_refine_detections_workspace_bytes": (_sz, [_i, _i]),
    "mvf_refine_detections": (_i, [_p, _p, _p, _p, C.POINTER(_f), _i, _i, _i, _f, _f, _i, _p, _p, _p, _p, _sz, _p]),
    "mvf_proposals_workspace_bytes": (_sz, [_i, _i, _i]),
    "mvf_proposals": (_i, [_p, _p, _p, C.POINTER(_f), _i, _i, _i, _i, _f, _p, _p, _p, _sz, _p]),
    "mvf_pipeline_host_workspace_bytes": (_sz, [_G, _i, _i, _i, _i, _i, _i, _i, _i]),
    "mvf_host_aux_create": (_i, [C.POINTER(_p)]),
    "mvf_host_aux_destroy": (_i, [_p]),
    "mvf_unproject_fuse_project_host": (_i, [_p, _p, _p, _G, _i, _i, _i, _i, _i, _i, _i, _i, _i, _p, _p,
                                             _i, _i, _i, _p, _p, _sz, _p, _p]),
    "mvf_unproject_fuse_project_workspace_bytes": (_sz, [_i, _i, _i, _i, _i]),
    "mvf_unproject_fuse_project": (_i, [_p, _p, _p, _G, _i, _i, _i, _i, _i, _i, _i, _i, _i, _p, _p,
                                        _i, _i, _i, _p, _p, _p, _sz, _p]),
    "mvf_fusion_neck_level_host": (_i, [_p, _p, _p, _G, _i, _i, _i, _i, _i, _i, _i, _i, _i, _p, _p,
                                        _i, _i, _i, _p, _f, _f, _f, _p, _p, _sz, _p, _p]),
    "mvf_error_string": (C.c_char_p, [_i]),
    "mvf_version": (C.c_char_p, []),
    "mvf_launch_count": (C.c_ulonglong, []),
}
for _name, (_res, _args) in _SIGS.items():
    _fn = getattr(lib, _name)          # AttributeError here = header/library mismatch: fail loudly
    _fn.restype = _res
    _fn.argtypes = _args

EXPORTED = tuple(_SIGS)


def check(rc, what):
    """Map an MVF_E* return code to the exception the reference layer would have raised."""
    if rc == MVF_OK:
        return
    msg = "%s: %s (code %d)" % (what, lib.mvf_error_string(rc).decode(), rc)
    if rc in (MVF_EINVAL, MVF_ENULL, MVF_EALIGN, MVF_EUNSUPPORTED, MVF_EWORKSPACE):
        raise ValueError(msg)
    raise RuntimeError(msg)


def launch_count():
    return int(lib.mvf_launch_count())


def version():
    return lib.mvf_version().decode()
