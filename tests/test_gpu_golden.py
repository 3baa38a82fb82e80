"""GPU: the CUDA kernels (through the C-ABI / layer mirror) against the golden fixtures produced by
the reference's own Python (tests/golden/make_golden.py)."""
import numpy as np
import pytest

from helpers import to_dev, close
from test_golden import load, cfg_from
from mulit_view_object_detection_b200.config import FusionConfig

pytestmark = pytest.mark.gpu


def _m():
    import mulit_view_object_detection_b200 as m
    return m


@pytest.mark.parametrize("name", ["fusion_a", "fusion_b"])
def test_fusion_pipeline_vs_reference_fixture(name):
    m = _m()
    d = load(name)
    cfg = cfg_from(d)
    dev = to_dev(d["feats"], d["Rcam"], d["Kmat"])
    per_view = m.unproj_feat(dev, cfg)
    close(per_view.cpu().numpy(), d["per_view"])                            # 1e-5 relative (FMA blend)
    rays, fused = m.unproject_fuse_project(*dev, cfg, int(d["proj_size"]), mode="sum")
    close(fused.cpu().numpy(), d["summed"])
    close(rays.cpu().numpy(), d["rays"])
    # voxel indices bit-exact: projecting the reference's own grid must reproduce its rays exactly
    rays2 = m.proj_grid(to_dev(d["summed"]) + dev[1:], cfg, int(d["proj_size"]))
    assert np.array_equal(rays2.cpu().numpy(), d["rays"])


def test_refine_and_detection_layer_vs_reference_fixture():
    m = _m()
    d = load("refine")
    cfg = FusionConfig(NUM_CLASSES=7, DETECTION_MIN_CONFIDENCE=0.3, DETECTION_MAX_INSTANCES=20)
    det = m.refine_detections_graph(*to_dev(d["rois"], d["probs"], d["deltas"], d["window"]), cfg)
    assert np.array_equal(det.cpu().numpy(), d["det"])
    cfg0 = FusionConfig(NUM_CLASSES=7, DETECTION_MIN_CONFIDENCE=0, DETECTION_MAX_INSTANCES=20)
    det0 = m.refine_detections_graph(*to_dev(d["rois"], d["probs"], d["deltas"], d["window"]), cfg0)
    assert np.array_equal(det0.cpu().numpy(), d["det_noconf"])
    d = load("detection_layer")
    cfgb = FusionConfig(NUM_CLASSES=7, DETECTION_MIN_CONFIDENCE=0.3, DETECTION_MAX_INSTANCES=20, IMAGES_PER_GPU=2,
                        IMAGE_SHAPE=np.array([96, 128, 3]))
    out = m.DetectionLayer(cfgb)(to_dev(d["rois"], d["probs"], d["deltas"]) + [d["image_meta"]])
    assert np.array_equal(out.cpu().numpy(), d["out"])


@pytest.mark.parametrize("name", ["roi_align_7x7", "roi_align_3x5"])
def test_roi_align_vs_reference_fixture(name):
    m = _m()
    d = load(name)
    layer = m.PyramidROIAlign(tuple(int(v) for v in d["pool"]))
    out = layer([to_dev(d["boxes"])[0], d["image_meta"]] + to_dev(d["P2"], d["P3"], d["P4"], d["P5"]))
    assert np.array_equal(out.cpu().numpy(), d["out"])


def test_proposals_vs_reference_fixture():
    m = _m()
    d = load("proposals")
    cfg = FusionConfig(PRE_NMS_LIMIT=int(d["pre_nms_limit"]), IMAGES_PER_GPU=2)
    out = m.ProposalLayer(int(d["proposal_count"]), float(d["nms_threshold"]), cfg)(to_dev(d["probs"], d["bbox"], d["anchors"]))
    assert np.array_equal(out.cpu().numpy(), d["out"])


def test_nms_vs_reference_numpy_nms():
    m = _m()
    d = load("nms_utils")
    for thr in (0.3, 0.5, 0.7):
        gold = d["keep_%02d" % int(thr * 10)]
        keep, count = m.non_max_suppression(*to_dev(d["boxes"], d["scores"]), 300, thr)
        assert int(count) == gold.shape[0]
        assert np.array_equal(keep.cpu().numpy()[:gold.shape[0]], gold)


def test_convlstm_vs_reference_cell():
    m = _m()
    d = load("convlstm")
    dW, db = to_dev(d["W"], d["b"])
    h = c = None
    for t in range(d["x"].shape[1]):
        h, c = m.convlstm_step(to_dev(d["x"][:, t])[0], h, c, dW, db)
        close(h.cpu().numpy(), d["h"][:, t], rtol=1e-5, atol=2e-6)
        close(c.cpu().numpy(), d["c"][:, t], rtol=1e-5, atol=2e-6)


@pytest.mark.parametrize("mode", ["add", "ident", "conv3d", "conv3d_tc"])
def test_grid_reas_vs_reference_fixture(mode):
    """The CUDA grid_reas against what the reference's own grid_reas computed (tests/golden/grid_reas_*.npz).  The tiny channel
    counts of the first fixtures take the exact-fp32 kernels for 'ident' and are rejected by the tensor-core U-Net; the
    'conv3d_tc' fixture has 32-channel sources and runs the U-Net on the tensor cores."""
    m = _m()
    from test_golden import named_weights, _dummy_depth
    from mulit_view_object_detection_b200 import weights_io as wio
    d = load("grid_reas_" + mode)
    tiny = mode == "conv3d"
    mode = mode.split("_")[0]
    V, F = int(d["V"]), int(d["F"])
    cfg = FusionConfig(GRID_REAS=mode, NUM_VIEWS=V, nvox=d["grids"].shape[2], nvox_z=d["grids"].shape[4], TOP_DOWN_PYRAMID_SIZE=F)
    named = named_weights(d)
    named.update({"grid_reas_depth_PG4" + s: w for s, w in _dummy_depth(cfg, F).items()})
    params = wio.fusion_params_from_keras(named, cfg, levels=(4,))
    if tiny:
        with pytest.raises(ValueError):            # 12 -> 8 channels: below the 32-channel K chunk of the tensor-core path
            m.grid_reas(to_dev(d["grids"])[0], "grid_reas_P4", cfg, params=m.prepare_params(params)["grid_reas_P4"])
        return
    out = m.grid_reas(to_dev(d["grids"])[0], "grid_reas_P4", cfg, params=m.prepare_params(params)["grid_reas_P4"])
    # four chained convolutions with un-normalised random weights (gain > 1 per layer): 1e-5 relative, floor 2e-6 of the output scale
    close(out.cpu().numpy(), d["out"], rtol=1e-5, atol=2e-6 * max(1.0, float(np.abs(d["out"]).max())))


def test_depth_sampling_vs_reference_fixture():
    m = _m()
    from test_golden import named_weights
    from mulit_view_object_detection_b200 import weights_io as wio
    d = load("depth_sampling_add")
    S, F = int(d["S"]), int(d["F"])
    cfg = FusionConfig(GRID_REAS="add", samples=S, TOP_DOWN_PYRAMID_SIZE=F, NUM_VIEWS=1)
    p = wio.fusion_params_from_keras(named_weights(d), cfg, levels=(4,))["grid_reas_depth_PG4"]
    out = m.depth_sampling(to_dev(d["x"])[0], cfg, "grid_reas_depth_PG4", params=p)
    close(out.cpu().numpy(), d["out"], rtol=1e-5, atol=1e-6)


def test_convlstm_sequence_vs_reference_convrnn3d():
    """convlstm() over the view axis (zero initial state, last output) against the reference's ConvRNN3D.call fixture."""
    m = _m()
    d = load("convlstm_sequence")
    dW, db = to_dev(d["W"], d["b"])
    out = m.convlstm(to_dev(d["x"])[0], "golden_convlstm", kernel=(3, 3, 3), filters=d["x"].shape[-1], params={"W": dW, "b": db})
    close(out.cpu().numpy(), d["out"], rtol=1e-5, atol=2e-6)
