"""GPU tests of the full ``model_multi`` inference row (SURVEY.md section 8(f) rank 4, config c4 at test size).

* the two heads (PyramidROIAlign kernel + dense layers) against the outputs of the reference's own ``fpn_classifier_graph`` /
  ``build_fpn_mask_graph`` (tests/golden/model_graphs.npz);
* ``MaskRCNN.predict`` stage by stage against oracle/model.py: every stage of the oracle is fed the PRODUCT's upstream
  tensors, so the discrete stages (top-k, NMS, per-class refinement) are compared on identical inputs -- proposals and
  detections must then agree exactly (same kernels as tests/test_gpu_heads.py), the dense stages within fp32-vs-float64
  accumulation error (2e-4 of the tensor maximum), the fusion neck within the path's 1e-5;
* ``detect`` end to end (molding, anchors, unmolding)."""
import numpy as np
import pytest

import oracle
from mulit_view_object_detection_b200.config import FusionConfig
from mulit_view_object_detection_b200 import model_host as MH
from mulit_view_object_detection_b200 import synthetic as syn
from test_model_host import G, graph_cfg, graph_params, rel

pytestmark = pytest.mark.gpu


def test_heads_match_reference_graph_builders():
    import torch
    import mulit_view_object_detection_b200 as m
    cfg = graph_cfg()
    net = m.MaskRCNN("inference", cfg, params=graph_params(cfg))
    maps = [torch.from_numpy(G["map%d" % i]).cuda() for i in range(4)]
    rois = torch.from_numpy(G["rois"]).cuda()
    logits, probs, bbox = net.fpn_classifier_graph(rois, maps, G["meta"])
    assert rel(logits.cpu().numpy(), G["cls_logits"]) <= 2e-4
    assert rel(probs.cpu().numpy(), G["cls_probs"]) <= 2e-4
    assert rel(bbox.cpu().numpy(), G["cls_bbox"]) <= 2e-4
    mask = net.build_fpn_mask_graph(rois[:, :6].contiguous(), maps, G["meta"])
    assert tuple(mask.shape) == G["mask"].shape and rel(mask.cpu().numpy(), G["mask"]) <= 2e-4
    P = net.build_resnet_fpn(torch.from_numpy(G["images"]).cuda())
    for i, o in enumerate(P):
        assert rel(o.cpu().numpy(), G["P%d" % (i + 2)]) <= 2e-4, i


def small_model_cfg(B=2, V=2):
    return FusionConfig(IMAGE_SHAPE=np.array([128, 128, 3]), NUM_VIEWS=V, IMAGES_PER_GPU=B, TOP_DOWN_PYRAMID_SIZE=64, NUM_CLASSES=5,
                        BACKBONE="resnet50", FPN_CLASSIF_FC_LAYERS_SIZE=64, nvox=16, nvox_z=16, samples=6, GRID_REAS="add",
                        PRE_NMS_LIMIT=300, POST_NMS_ROIS_INFERENCE=60, DETECTION_MAX_INSTANCES=12, DETECTION_MIN_CONFIDENCE=0.0,
                        IMAGE_MIN_DIM=128, IMAGE_MAX_DIM=128)


def small_params(cfg, seed):
    """Random parameters whose classifier never answers 'background' (bias of class 0 pushed down): with untrained weights every
    ROI would otherwise be class 0 and DetectionLayer / the mask head would only ever see zero rows."""
    p = MH.randomize(MH.init_params(cfg, seed=seed), seed=seed + 1)
    b = np.array(p["mrcnn_class_logits"]["bias"], np.float32)
    b[0] -= 20.0
    p["mrcnn_class_logits"] = {"kernel": p["mrcnn_class_logits"]["kernel"], "bias": b}
    return p


def test_predict_stage_by_stage_vs_oracle():
    import mulit_view_object_detection_b200 as m
    B, V = 2, 2
    cfg = small_model_cfg(B, V)
    params = small_params(cfg, 11)
    net = m.MaskRCNN("inference", cfg, params=params)
    rng = np.random.default_rng(13)
    images = rng.normal(0, 50, (B, V, 128, 128, 3)).astype(np.float32)
    _, Rcam, Kmat = syn.make_scene(cfg, B, V, 8, 8, 4, seed=14, image_hw=(128, 128))
    meta = np.stack([m.weights_io.compose_image_meta(0, (128, 128, 3), (128, 128, 3), (0, 0, 128, 128), 1.0,
                                                     np.zeros(cfg.NUM_CLASSES, np.int32)) for _ in range(B)]).astype(np.float32)
    anchors = np.broadcast_to(net.get_anchors((128, 128, 3)), (B,) + net.get_anchors((128, 128, 3)).shape).copy()
    res, feats = net.predict([images, meta, anchors, Rcam, Kmat], return_features=True)
    det, mclass, mbbox, mmask, rois, rclass, rbbox = (t.cpu().numpy() for t in res)
    assert det.shape == (B, 12, 6) and mmask.shape == (B, 12, 28, 28, 5) and rois.shape == (B, 60, 4)
    given = {"P": [p.cpu().numpy() for p in feats["P"]], "maps": [p.cpu().numpy() for p in feats["maps"]], "rpn_class": rclass,
             "rpn_bbox": rbbox, "rpn_rois": rois, "mrcnn_class": mclass, "mrcnn_bbox": mbbox, "detections": det}
    # each oracle stage on the product's upstream tensors
    o = oracle.model.predict(images, meta, anchors, Rcam, Kmat, params, cfg, given=given)
    for i in range(5):
        assert rel(given["P"][i], o["P"][i]) <= 2e-4, ("P", i)                 # backbone + FPN: fp32 cuDNN vs float64
    for i in range(5):
        a, b = given["maps"][i], o["maps"][i]
        assert a.shape == b.shape
        np.testing.assert_allclose(a, b, rtol=1e-5, atol=1e-5 * max(1.0, float(np.abs(b).max())))   # neck on the product's P maps
    assert not given["maps"][0].any() and not given["maps"][1].any()          # PG2 / PG3 are zeros (model_multi.py:2406-2410)
    assert given["maps"][2].any()
    assert rel(rclass, o["rpn_class"]) <= 2e-4 and rel(rbbox, o["rpn_bbox"]) <= 2e-4
    np.testing.assert_allclose(rois, o["rpn_rois"], rtol=0, atol=1e-6)         # ProposalLayer on identical inputs
    assert rel(mclass, o["mrcnn_class"]) <= 2e-4 and rel(mbbox, o["mrcnn_bbox"]) <= 2e-4
    np.testing.assert_allclose(det, o["detections"], rtol=0, atol=1e-6)        # DetectionLayer on identical inputs
    assert (det[..., 4] > 0).any()
    assert rel(mmask, o["mrcnn_mask"]) <= 2e-4


def test_detect_end_to_end():
    import mulit_view_object_detection_b200 as m
    B, V = 1, 3
    cfg = small_model_cfg(B, V)
    net = m.MaskRCNN("inference", cfg, params=small_params(cfg, 21))
    rng = np.random.default_rng(23)
    scenes = [[rng.integers(0, 255, (96, 128, 3)).astype(np.uint8) for _ in range(V)] for _ in range(B)]    # padded to 128x128
    _, Rcam, Kmat = syn.make_scene(cfg, B, V, 8, 8, 4, seed=24, image_hw=(128, 128))
    out = net.detect(scenes, Rcam, Kmat)
    assert len(out) == B
    r = out[0]
    n = r["rois"].shape[0]
    assert n > 0
    assert r["class_ids"].shape == (n,) and r["scores"].shape == (n,) and r["masks"].shape == (96, 128, n)
    assert r["rois"].dtype == np.int32 and (r["rois"][:, 2] > r["rois"][:, 0]).all() and (r["rois"][:, 3] > r["rois"][:, 1]).all()
    assert (r["class_ids"] > 0).all() and r["masks"].dtype == bool
    again = net.detect(scenes, Rcam, Kmat)[0]
    assert np.array_equal(again["rois"], r["rois"]) and np.array_equal(again["masks"], r["masks"])           # deterministic


@pytest.mark.parametrize("mode", ["conv3d", "ident", "lstm3d"])
def test_model_neck_modes_vs_oracle(mode):
    """The shipped configs use GRID_REAS='conv3d' (samples/interior/interior_multi.py:391,419): the model's fusion neck in the
    tensor-core modes against oracle.fusion_neck fed the PRODUCT's pyramid (tolerance of the 3 x fp16 split: rtol 1e-5 plus 1e-5 of
    the map's maximum), and the stages behind it on identical inputs as in the 'add' test."""
    import mulit_view_object_detection_b200 as m
    B, V = 1, 2
    cfg = small_model_cfg(B, V)
    cfg.GRID_REAS = mode
    params = small_params(cfg, 31)
    net = m.MaskRCNN("inference", cfg, params=params)
    rng = np.random.default_rng(33)
    images = rng.normal(0, 50, (B, V, 128, 128, 3)).astype(np.float32)
    _, Rcam, Kmat = syn.make_scene(cfg, B, V, 8, 8, 4, seed=34, image_hw=(128, 128))
    meta = np.stack([m.weights_io.compose_image_meta(0, (128, 128, 3), (128, 128, 3), (0, 0, 128, 128), 1.0,
                                                     np.zeros(cfg.NUM_CLASSES, np.int32)) for _ in range(B)]).astype(np.float32)
    a = net.get_anchors((128, 128, 3))
    anchors = np.broadcast_to(a, (B,) + a.shape).copy()
    res, feats = net.predict([images, meta, anchors, Rcam, Kmat], return_features=True)
    det, mclass, mbbox, mmask, rois, rclass, rbbox = (t.cpu().numpy() for t in res)
    P = [p.cpu().numpy() for p in feats["P"]]
    maps = [p.cpu().numpy() for p in feats["maps"]]
    o_maps = oracle.fusion_neck(P, Rcam, Kmat, cfg, params)
    for lvl, (g, o) in enumerate(zip(maps, o_maps)):
        assert g.shape == o.shape
        np.testing.assert_allclose(g, o, rtol=1e-5, atol=1e-5 * max(1.0, float(np.abs(o).max())), err_msg="PG%d" % (lvl + 2))
    assert maps[2].any() and not maps[0].any()
    given = {"P": P, "maps": maps, "rpn_class": rclass, "rpn_bbox": rbbox, "rpn_rois": rois, "mrcnn_class": mclass, "mrcnn_bbox": mbbox,
             "detections": det}
    o = oracle.model.predict(images, meta, anchors, Rcam, Kmat, params, cfg, given=given)
    np.testing.assert_allclose(rois, o["rpn_rois"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(det, o["detections"], rtol=0, atol=1e-6)
    assert rel(mmask, o["mrcnn_mask"]) <= 2e-4
