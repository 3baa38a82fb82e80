"""CPU: the C-ABI library loads and exports every symbol include/mvfusion.h declares, and the
Python binding declares a signature for each (no compute calls without a GPU)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "mvfusion.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mvf_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_entry_points():
    names = _declared()
    for required in ("mvf_unproject_fuse", "mvf_project_rays", "mvf_project_depth_collapse", "mvf_view_reduce",
                     "mvf_convlstm_step", "mvf_ident_fuse", "mvf_pyramid_roi_align", "mvf_nms",
                     "mvf_refine_detections", "mvf_proposals", "mvf_unproject_fuse_project_host"):
        assert required in names


def test_library_exports_every_declared_symbol():
    from mulit_view_object_detection_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH)
    dll = ctypes.CDLL(_lib.LIB_PATH)
    for name in _declared():
        assert hasattr(dll, name), "libmvfusion.so does not export %s" % name
    assert set(_declared()) == set(_lib.EXPORTED), "Python binding and header disagree"


def test_no_gpu_calls_are_pure():
    from mulit_view_object_detection_b200 import _lib
    assert _lib.version().startswith("mvfusion")
    assert _lib.lib.mvf_error_string(_lib.MVF_EALIGN).decode().startswith("channel count")
    assert _lib.lib.mvf_nms_workspace_bytes(1, 1000) > 1000 * 32 * 4
    assert _lib.lib.mvf_nms_workspace_bytes(0, 5) == 0
    g = _lib.MvfGrid(64, 64, -2.5, 2.5, 5 / 64, 1.0, 10.0, 9 / 64)
    n = _lib.lib.mvf_pipeline_host_workspace_bytes(ctypes.byref(g), 1, 8, 40, 40, 256, 40, 40, 20)
    assert n >= 4 * (8 * 40 * 40 * 256 + 64 ** 3 * 256 + 20 * 40 * 40 * 256)


def test_argument_validation_without_gpu():
    """Bad arguments are rejected before any CUDA call, so this runs on a CPU-only host."""
    import numpy as np
    from mulit_view_object_detection_b200 import _lib
    g = _lib.MvfGrid(8, 8, -2.0, 2.0, 0.5, 1.0, 9.0, 1.0)
    buf = np.zeros(64, np.float32)
    p = ctypes.c_void_p(buf.ctypes.data)
    call = lambda **kw: _lib.lib.mvf_unproject_fuse(
        kw.get("feats", p), p, None, p, ctypes.byref(g), 1, kw.get("V", 2), 4, 4, kw.get("C", 4), 64, 64,
        kw.get("mode", 1), 0, 0.0, 0, 0, None, None, kw.get("out", p), None, None, None, None)
    assert call(feats=None) == _lib.MVF_ENULL
    assert call(C=6) == _lib.MVF_EALIGN
    assert call(mode=9) == _lib.MVF_EINVAL
    assert call(V=_lib.MAX_VIEWS + 1) == _lib.MVF_EUNSUPPORTED
    g_bad = _lib.MvfGrid(8, 8, -2.0, 2.0, 0.3, 1.0, 9.0, 1.0)       # tf.range would yield 14 centres, not 8
    assert _lib.lib.mvf_unproject_fuse(p, p, None, p, ctypes.byref(g_bad), 1, 2, 4, 4, 4, 64, 64, 1, 0, 0.0, 0, 0,
                                       None, None, p, None, None, None, None) == _lib.MVF_EINVAL


def test_layers_refuse_cpu_tensors():
    import pytest
    import torch
    import mulit_view_object_detection_b200 as m
    cfg = m.FusionConfig(nvox=8, nvox_z=8)
    with pytest.raises(ValueError, match="CUDA"):
        m.unproj_feat([torch.zeros(1, 2, 4, 4, 4), torch.zeros(1, 2, 3, 4), torch.zeros(1, 3, 3)], cfg)
    with pytest.raises(ValueError, match="CUDA"):
        m.non_max_suppression(torch.zeros(4, 4), torch.zeros(4), 10, 0.5)
