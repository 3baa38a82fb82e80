"""CPU tests of the full-model row: the host helpers against the reference's own NumPy
(fixture model_graphs.npz, written by tests/golden/make_golden.py from mrcnn/utils.py and model_multi.py run unmodified), and
the ORACLE's dense graphs (oracle/model.py) against the outputs of the reference's graph builders stored in the same fixture.
The product's dense graphs are device-agnostic torch code, so backbone + FPN + RPN are also checked here on the CPU; the
heads need the ROIAlign kernel and are checked on the GPU (tests/test_gpu_model.py)."""
import os

import numpy as np
import pytest

import oracle
from mulit_view_object_detection_b200.config import FusionConfig
from mulit_view_object_detection_b200 import model_host as MH

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "model_graphs.npz"))


def graph_cfg():
    cfg = FusionConfig(IMAGE_SHAPE=np.array([64, 64, 3]), NUM_VIEWS=2, IMAGES_PER_GPU=1, TOP_DOWN_PYRAMID_SIZE=16, NUM_CLASSES=5,
                       BACKBONE="resnet50", POOL_SIZE=3, MASK_POOL_SIZE=4, FPN_CLASSIF_FC_LAYERS_SIZE=32)
    return cfg


def graph_params(cfg):
    p = MH.randomize(MH.init_params(cfg, seed=5), seed=6)
    assert abs(MH.checksum(p) - float(G["weights_checksum"])) <= 1e-9 * abs(float(G["weights_checksum"])), "initialiser drifted: regenerate the fixture"
    return p


def rel(a, b):
    return float(np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64)).max() / (np.abs(b).max() + 1e-30))


def test_host_helpers_match_reference_numpy():
    cfg = graph_cfg()
    shapes = MH.compute_backbone_shapes(cfg, (128, 192, 3))
    assert np.array_equal(shapes, G["backbone_shapes"])
    a = MH.generate_pyramid_anchors((32, 64, 128, 256, 512), [0.5, 1, 2], shapes, [4, 8, 16, 32, 64], 1)
    assert np.array_equal(a.astype(np.float64), G["anchors"])                      # bit-exact (same NumPy expressions)
    assert np.array_equal(MH.norm_boxes(a, (128, 192)), G["anchors_norm"])
    assert np.array_equal(MH.denorm_boxes(MH.norm_boxes(a[:64], (128, 192)), (128, 192)), G["denorm"])


def test_resize_and_mold():
    img = (np.arange(48 * 64 * 3) % 251).reshape(48, 64, 3).astype(np.uint8)
    out, window, scale, padding, crop = MH.resize_image(img, min_dim=64, max_dim=64, mode="square")
    assert out.shape == (64, 64, 3) and window == (8, 0, 56, 64) and scale == 1 and padding[0] == (8, 8)
    assert np.array_equal(out[8:56], img) and not out[:8].any() and not out[56:].any()
    out2, window2, scale2, _, _ = MH.resize_image(img, min_dim=128, max_dim=128, mode="square")      # scaled up by 2
    assert out2.shape == (128, 128, 3) and scale2 == 2.0 and window2 == (16, 0, 112, 128)
    same, w, s, _, _ = MH.resize_image(img, mode="none")
    assert same is img and w == (0, 0, 48, 64) and s == 1
    wide = np.ones((64, 100, 3), np.uint8)
    p64, w64, s64, _, _ = MH.resize_image(wide, min_dim=64, mode="pad64")
    assert p64.shape == (64, 128, 3) and w64 == (0, 14, 64, 114) and s64 == 1 and p64[:, 14:114].all() and not p64[:, :14].any()
    cfg = graph_cfg()
    assert np.allclose(MH.mold_image(img, cfg)[0, 0], img[0, 0].astype(np.float32) - np.array([123.7, 116.8, 103.9]))
    m = MH.unmold_mask(np.ones((4, 4), np.float32), (2, 3, 10, 9), (16, 16, 3))
    assert m.dtype == bool and m.sum() == 8 * 6 and m[2:10, 3:9].all()


def test_oracle_dense_graphs_match_reference_graph_builders():
    """oracle/model.py vs the reference's build_resnet_fpn / rpn_graph / fpn_classifier_graph / build_fpn_mask_graph (both
    float64 evaluations rounded to fp32: 1e-5 relative to the tensor maximum)."""
    cfg = graph_cfg()
    p = graph_params(cfg)
    P = oracle.model.resnet_fpn(G["images"], p, cfg)
    for i, o in enumerate(P):
        assert o.shape == G["P%d" % (i + 2)].shape
        assert rel(o, G["P%d" % (i + 2)]) <= 1e-5, i
    logits, probs, bbox = oracle.model.rpn(G["P3"][:, 0], p)
    assert rel(logits, G["rpn_logits"]) <= 1e-5 and rel(probs, G["rpn_probs"]) <= 1e-5 and rel(bbox, G["rpn_bbox"]) <= 1e-5
    maps = [G["map%d" % i] for i in range(4)]
    cl, cp, cb = oracle.model.classifier(G["rois"], maps, G["meta"], p, cfg)
    assert rel(cl, G["cls_logits"]) <= 1e-5 and rel(cp, G["cls_probs"]) <= 1e-5 and rel(cb, G["cls_bbox"]) <= 1e-5
    mk = oracle.model.mask_head(G["rois"][:, :6], maps, G["meta"], p, cfg)
    assert mk.shape == G["mask"].shape and rel(mk, G["mask"]) <= 1e-5


def test_product_dense_graphs_on_cpu_match_reference_graph_builders():
    """model.py's backbone + FPN + RPN (torch fp32, here on the CPU device) vs the same fixture: fp32 accumulation through
    ~50 layers, 2e-4 relative to the tensor maximum."""
    import mulit_view_object_detection_b200 as m
    import torch
    cfg = graph_cfg()
    net = m.MaskRCNN("inference", cfg, params=graph_params(cfg), device="cpu")
    P = net.build_resnet_fpn(torch.from_numpy(G["images"]))
    for i, o in enumerate(P):
        assert tuple(o.shape) == G["P%d" % (i + 2)].shape
        assert rel(o.numpy(), G["P%d" % (i + 2)]) <= 2e-4, i
    logits, probs, bbox = net.rpn_graph(torch.from_numpy(G["P3"][:, 0]))
    assert rel(logits.numpy(), G["rpn_logits"]) <= 2e-4 and rel(probs.numpy(), G["rpn_probs"]) <= 2e-4
    assert rel(bbox.numpy(), G["rpn_bbox"]) <= 2e-4


def test_named_weights_round_trip(tmp_path):
    import mulit_view_object_detection_b200 as m
    from mulit_view_object_detection_b200 import weights_io
    cfg = graph_cfg()
    p = graph_params(cfg)
    path = str(tmp_path / "w.npz")
    weights_io.write_npz(path, MH.named_weights(p))
    net = m.MaskRCNN("inference", cfg, device="cpu", seed=99)
    assert abs(MH.checksum(net.params) - MH.checksum(p)) > 1.0
    net.load_weights(path)
    for name, q in MH.named_weights(p).items():
        for a, b in zip(q, MH.named_weights(net.params)[name]):
            assert np.array_equal(np.asarray(a, np.float32).reshape(-1), np.asarray(b, np.float32).reshape(-1)), name
    with pytest.raises(ValueError):
        m.MaskRCNN("training", cfg, device="cpu")


def test_mold_and_unmold_round_trip_on_host():
    """mold_inputs / unmold_detections (model_multi.py:2915-3017) are pure host code: a detection given in normalised coordinates of
    the padded network input comes back in pixels of the ORIGINAL image, zero-area boxes are dropped, masks land inside their boxes."""
    import mulit_view_object_detection_b200 as m
    cfg = graph_cfg()
    cfg.IMAGE_MIN_DIM = cfg.IMAGE_MAX_DIM = 64
    net = m.MaskRCNN("inference", cfg, device="cpu")
    views = [np.full((48, 64, 3), 100 + i, np.uint8) for i in range(2)]
    molded, metas, windows = net.mold_inputs(views)
    assert molded.shape == (2, 64, 64, 3) and molded.dtype == np.float64      # float32 image - float64 MEAN_PIXEL, as in the reference
    assert np.allclose(molded[0, 8, 0], 100 - np.array([123.7, 116.8, 103.9])) and np.allclose(molded[0, 0, 0], -np.array([123.7, 116.8, 103.9]))
    assert tuple(windows[0]) == (8, 0, 56, 64) and metas.shape == (2, 12 + cfg.NUM_CLASSES)
    assert tuple(metas[0, 1:4]) == (48, 64, 3) and tuple(metas[0, 4:7]) == (64, 64, 3) and tuple(metas[0, 7:11]) == (8, 0, 56, 64)
    # a box covering rows 20..40, columns 16..48 of the molded image = rows 12..32 of the original
    box_px = np.array([[20, 16, 40, 48]], np.float32)
    det = np.zeros((4, 6), np.float32)
    det[0, :4] = MH.norm_boxes(box_px, (64, 64))[0]
    det[0, 4:] = (3, 0.9)
    det[1, :4] = MH.norm_boxes(np.array([[30, 30, 30, 40]], np.float32), (64, 64))[0]      # zero height: dropped
    det[1, 4:] = (2, 0.8)
    masks = np.zeros((4, 28, 28, cfg.NUM_CLASSES), np.float32)
    masks[0, :, :, 3] = 1.0
    masks[0, :, :14, 3] = 0.2                                   # left half below the 0.5 threshold
    masks[1, :, :, 2] = 1.0
    boxes, class_ids, scores, full = net.unmold_detections(det, masks, (48, 64, 3), (64, 64, 3), windows[0])
    assert boxes.tolist() == [[12, 16, 32, 48]] and class_ids.tolist() == [3] and np.allclose(scores, [0.9])
    assert full.shape == (48, 64, 1) and full.dtype == bool
    assert full[12:32, 32:48, 0].all() and not full[12:32, 16:32, 0].any() and full.sum() == 20 * 16
    none = net.unmold_detections(np.zeros((4, 6), np.float32), masks, (48, 64, 3), (64, 64, 3), windows[0])
    assert none[0].shape == (0, 4) and none[3].shape == (48, 64, 0)
