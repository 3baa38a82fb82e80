"""CPU: the oracle against the golden fixtures in tests/golden/*.npz, which were produced by
executing the REFERENCE'S OWN Python (mrcnn/model_multi.py, mrcnn/recurrent.py, mrcnn/utils.py)
over the eager NumPy tf stand-in (tests/golden/make_golden.py, tests/golden/tf1_shim.py)."""
import os

import numpy as np
import pytest

import oracle
from mulit_view_object_detection_b200.config import FusionConfig

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return np.load(os.path.join(GOLD, name + ".npz"))


def cfg_from(d, **extra):
    kw = {k[4:]: d[k].item() if d[k].ndim == 0 else d[k] for k in d.files if k.startswith("cfg_")}
    kw.update(extra)
    return FusionConfig(**kw)


@pytest.mark.parametrize("name", ["fusion_a", "fusion_b"])
def test_unproj_feat_sum_proj_grid_bit_exact(name):
    d = load(name)
    cfg = cfg_from(d)
    per_view = oracle.unproj_feat(d["feats"], d["Rcam"], d["Kmat"], cfg)
    assert per_view.shape == d["per_view"].shape
    assert np.array_equal(per_view, d["per_view"])                 # reference unproj_feat, bit for bit
    summed = oracle.fuse_views(per_view, "sum")
    assert np.array_equal(summed, d["summed"])                     # K.sum(axis=1)
    rays = oracle.proj_grid(summed, d["Rcam"], d["Kmat"], cfg, int(d["proj_size"]))
    assert np.array_equal(rays, d["rays"])                         # reference proj_grid + nearest3
    assert (d["rays"] != 0).any() and (d["per_view"] != 0).mean() > 0.2


def test_box_helpers_bit_exact():
    d = load("boxes")
    applied = oracle.apply_box_deltas(d["boxes"], d["deltas"])
    assert np.array_equal(applied, d["applied"])
    assert np.array_equal(oracle.clip_boxes(applied, d["window"]), d["clipped"])
    # the reference's own NumPy twin (mrcnn/utils.py apply_box_deltas, float64) agrees to fp32 accuracy
    np.testing.assert_allclose(applied, d["utils_applied_f64"], rtol=2e-6, atol=1e-6)


def test_nms_against_reference_numpy_nms():
    """mrcnn/utils.py:381-415 run unmodified: same strict '>' rule; scores are distinct so the
    tie-break difference (utils: higher index first) does not matter."""
    d = load("nms_utils")
    for thr in (0.3, 0.5, 0.7):
        keep = oracle.non_max_suppression(d["boxes"], d["scores"], 10 ** 6, thr)
        assert np.array_equal(keep, d["keep_%02d" % int(thr * 10)])
    iou = oracle.iou_tf(d["boxes"][0], d["boxes"])
    np.testing.assert_allclose(iou, d["iou_row0"], rtol=1e-6, atol=1e-7)


def test_refine_detections_bit_exact():
    d = load("refine")
    cfg = FusionConfig(NUM_CLASSES=7, DETECTION_MIN_CONFIDENCE=0.3, DETECTION_MAX_INSTANCES=20)
    det, keep = oracle.refine_detections(d["rois"], d["probs"], d["deltas"], d["window"], cfg)
    assert np.array_equal(det, d["det"])
    assert (d["det"][:, 5] > 0).sum() >= 5
    cfg0 = FusionConfig(NUM_CLASSES=7, DETECTION_MIN_CONFIDENCE=0, DETECTION_MAX_INSTANCES=20)
    det0, _ = oracle.refine_detections(d["rois"], d["probs"], d["deltas"], d["window"], cfg0)
    assert np.array_equal(det0, d["det_noconf"])


def test_detection_layer_bit_exact():
    d = load("detection_layer")
    cfg = FusionConfig(NUM_CLASSES=7, DETECTION_MIN_CONFIDENCE=0.3, DETECTION_MAX_INSTANCES=20, IMAGES_PER_GPU=2,
                       IMAGE_SHAPE=np.array([96, 128, 3]))
    out = oracle.detection_layer(d["rois"], d["probs"], d["deltas"], d["image_meta"], cfg)
    assert np.array_equal(out, d["out"])


@pytest.mark.parametrize("name", ["roi_align_7x7", "roi_align_3x5"])
def test_pyramid_roi_align_bit_exact(name):
    d = load(name)
    maps = [d["P2"], d["P3"], d["P4"], d["P5"]]
    out = oracle.pyramid_roi_align(d["boxes"], d["image_meta"][0, 4:7], maps, tuple(int(v) for v in d["pool"]))
    assert np.array_equal(out, d["out"])
    assert len(np.unique(oracle.roi_levels(d["boxes"], d["image_meta"][0, 4:7]))) >= 3


def test_proposal_layer_bit_exact():
    d = load("proposals")
    cfg = FusionConfig(PRE_NMS_LIMIT=int(d["pre_nms_limit"]), IMAGES_PER_GPU=2)
    out = oracle.proposal_layer(d["probs"], d["bbox"], d["anchors"], int(d["proposal_count"]),
                                float(d["nms_threshold"]), cfg)
    assert np.array_equal(out, d["out"])


def test_convlstm_cell_matches_reference_cell():
    """ConvLSTMCell.call executed from mrcnn/recurrent.py: gate order j,i,f,o, forget bias 1."""
    d = load("convlstm")
    B, V = d["x"].shape[:2]
    F = d["W"].shape[-1] // 4
    c = np.zeros(d["x"].shape[:1] + d["x"].shape[2:5] + (F,), np.float32)
    h = np.zeros_like(c)
    for t in range(V):
        h, c = oracle.convlstm_cell_step(d["x"][:, t], c, h, d["W"], d["b"])
        np.testing.assert_allclose(h, d["h"][:, t], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(c, d["c"][:, t], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(oracle.convlstm(d["x"], d["W"], d["b"]), d["h"][:, -1], rtol=1e-5, atol=1e-6)


def test_pose_helpers_match_reference():
    """mrcnn/utils.py:1175-1218 quat2rot / vec2rot run unmodified vs the synthetic generator's own maths."""
    from mulit_view_object_detection_b200 import synthetic as syn
    d = load("poses")
    for vec, R in zip(d["vec"], d["vec_R"]):
        np.testing.assert_allclose(syn.look_at_rotation(vec[:3], vec[3:6], vec[6:]), R, rtol=1e-12, atol=1e-12)
    for q, R in zip(d["quat"], d["quat_R"]):
        assert abs(np.linalg.det(R) - 1) < 1e-9 and np.allclose(R @ R.T, np.eye(3), atol=1e-9)


# ---- the reference's grid_reas / depth_sampling WIRING (model_multi.py:394-488), executed with functional layer stand-ins
def named_weights(d):
    """{keras layer name: get_weights() list} as stored by make_golden.py (keys w__<layer>__<i>)."""
    named = {}
    for k in d.files:
        if k.startswith("w__"):
            _, layer, i = k.split("__")
            named.setdefault(layer, {})[int(i)] = d[k]
    return {k: [v[i] for i in sorted(v)] for k, v in named.items()}


@pytest.mark.parametrize("mode", ["add", "ident", "conv3d", "conv3d_tc"])
def test_grid_reas_wiring_and_layer_names(mode):
    """Layer names -> parameters through weights_io (the names the reference gives its Keras layers), then the oracle's
    grid_reas must reproduce what the reference's own grid_reas computed: view-major channel order, ReLU placement, the
    deconv-first skip concat, the BN layer names without underscore."""
    from mulit_view_object_detection_b200 import weights_io as wio
    d = load("grid_reas_" + mode)
    mode = mode.split("_")[0]
    V, F = int(d["V"]), int(d["F"])
    cfg = FusionConfig(GRID_REAS=mode, NUM_VIEWS=V, nvox=d["grids"].shape[2], nvox_z=d["grids"].shape[4], TOP_DOWN_PYRAMID_SIZE=F)
    named = named_weights(d)
    named.update({"grid_reas_depth_PG4" + s: w for s, w in _dummy_depth(cfg, F).items()})
    params = wio.fusion_params_from_keras(named, cfg, levels=(4,))["grid_reas_P4"]
    out = oracle.grid_reas(d["grids"], "grid_reas_P4", cfg, params)
    assert out.shape == d["out"].shape
    np.testing.assert_allclose(out, d["out"], rtol=2e-6, atol=2e-7)


def _dummy_depth(cfg, F):
    S = int(cfg.samples)
    one = [np.ones(1, np.float32), np.zeros(1, np.float32), np.zeros(1, np.float32), np.ones(1, np.float32)]
    if cfg.GRID_REAS == "conv3d":
        out = {}
        for i, (cin, cout) in enumerate(((F * S, 8), (8, F)), 1):
            out["_DepthwiseConv_%d" % i] = [np.ones((1, 1, cin, 1), np.float32), np.zeros(cin, np.float32)]
            out["2DConv_%d" % i] = [np.zeros((1, 1, cin, cout), np.float32), np.zeros(cout, np.float32)]
        return out
    return {"2DConv": [np.zeros((1, 1, S, 1), np.float32), np.zeros(1, np.float32)], "bn_deconv": one}


@pytest.mark.parametrize("mode", ["add", "conv3d"])
def test_depth_sampling_wiring_and_layer_names(mode):
    from mulit_view_object_detection_b200 import weights_io as wio
    d = load("depth_sampling_" + mode)
    S, F = int(d["S"]), int(d["F"])
    cfg = FusionConfig(GRID_REAS=mode, samples=S, TOP_DOWN_PYRAMID_SIZE=F, NUM_VIEWS=1)
    named = named_weights(d)
    if mode == "conv3d":                      # the grid_reas layers are not part of this fixture: placeholders for the mapper
        for suf, shp, n in (("_1", (3, 3, 3, d["x"].shape[-1], 2 * F), 2 * F), ("_2", (3, 3, 3, 2 * F, 4 * F), 4 * F),
                            ("_deconv_1", (3, 3, 3, 2 * F, 4 * F), 2 * F), ("_deconv_2", (3, 3, 3, F, 4 * F), F)):
            named["grid_reas_P4_3D_conv" + suf] = [np.zeros(shp, np.float32), np.zeros(n, np.float32)]
    p = wio.fusion_params_from_keras(named, cfg, levels=(4,))["grid_reas_depth_PG4"]
    if mode == "conv3d":
        out = oracle.depth_sampling_conv3d(d["x"], p)
    else:
        out = oracle.depth_sampling(d["x"], p["weight"], p["bias"], p.get("bn"))
    assert out.shape == d["out"].shape
    np.testing.assert_allclose(out, d["out"], rtol=2e-6, atol=2e-7)


def test_convlstm_over_views_matches_reference_convrnn3d():
    """ConvRNN3D.call + get_initial_state executed from the reference (recurrent.py:143-173,230-300): zero initial states with the
    input's channel count, the cell applied per view, the LAST output returned -- what convlstm() (model_multi.py:109-123) builds."""
    d = load("convlstm_sequence")
    out = oracle.convlstm(d["x"], d["W"], d["b"])
    assert out.shape == d["out"].shape
    np.testing.assert_allclose(out, d["out"], rtol=2e-6, atol=2e-7)
