"""GPU parity: PyramidROIAlign, NMS, refine_detections, ProposalLayer, ConvLSTM and 'ident'
through the C-ABI vs the oracle.  Keep-indices and levels bit-exact; crops bit-exact (the kernel
uses the oracle's individually rounded lerp); ConvLSTM / ident within the stated tolerance."""
import numpy as np
import pytest

import oracle
from helpers import small_cfg, to_dev, close
from mulit_view_object_detection_b200 import synthetic as syn

pytestmark = pytest.mark.gpu


def _m():
    import mulit_view_object_detection_b200 as m
    return m


def _maps(rng, B, C, sizes):
    return [np.maximum(rng.standard_normal((B, s, s, C)), 0).astype(np.float32) for s in sizes]


@pytest.mark.parametrize("pool,C,R", [((7, 7), 256, 200), ((14, 14), 32, 50), ((1, 1), 4, 16), ((3, 5), 64, 33)])
def test_pyramid_roi_align(pool, C, R):
    m = _m()
    rng = np.random.default_rng(0)
    B = 2
    boxes = syn.make_rois(rng, B, R)
    boxes[0, 0] = [0.0, 0.0, 1.0, 1.0]                 # whole image
    boxes[0, 1] = [0.9, 0.9, 1.3, 1.4]                 # partly outside -> extrapolation zeros
    boxes[1, 0] = [0.2, 0.3, 0.2, 0.3]                 # zero area -> level 2
    maps = _maps(rng, B, C, (40, 20, 10, 5))
    meta = syn.make_image_meta(B, (1024, 1024, 3), 5)
    layer = m.PyramidROIAlign(pool)
    out, lv = layer([to_dev(boxes)[0], meta] + to_dev(*maps), return_levels=True)
    o_lv = oracle.roi_levels(boxes, (1024, 1024, 3))
    assert np.array_equal(lv.cpu().numpy(), o_lv)
    assert set(np.unique(o_lv)) == {2, 3, 4, 5}
    o = oracle.pyramid_roi_align(boxes, (1024, 1024, 3), maps, pool)
    assert np.array_equal(out.cpu().numpy(), o)        # bit-exact crops


def test_roi_align_linear_ramp_and_whole_image():
    """ROIAlign of a linear ramp is the ramp; the whole-image box at ph=H returns the image."""
    m = _m()
    H = 16
    yy, xx = np.meshgrid(np.arange(H, dtype=np.float32), np.arange(H, dtype=np.float32), indexing="ij")
    ramp = np.stack([yy, xx, yy + xx, 2 * yy - xx], axis=-1)[None]
    boxes = np.array([[[0.0, 0.0, 1.0, 1.0]]], np.float32)
    meta = syn.make_image_meta(1, (7, 7, 3), 1)        # tiny image -> 4 + round(log2(7/224)) = -1 -> clamps to level 2
    maps = [ramp.astype(np.float32)] + [np.zeros((1, 4, 4, 4), np.float32)] * 3
    out = m.PyramidROIAlign((H, H))([to_dev(boxes)[0], meta] + to_dev(*maps))
    assert np.array_equal(out.cpu().numpy()[0, 0], ramp[0])


def _boxes_scores(rng, n, spread=0.3):
    c = rng.uniform(0.2, 0.8, (n, 2))
    s = rng.uniform(0.05, spread, (n, 2))
    boxes = np.concatenate([c - s / 2, c + s / 2], axis=1).astype(np.float32)
    scores = rng.permutation(n).astype(np.float32) / n          # distinct
    return boxes, scores


@pytest.mark.parametrize("n,thr,max_out", [(1, 0.5, 10), (37, 0.3, 100), (1000, 0.3, 100), (3000, 0.7, 1000), (6000, 0.7, 1000)])
def test_nms_keep_indices(n, thr, max_out):
    m = _m()
    rng = np.random.default_rng(n)
    boxes, scores = _boxes_scores(rng, n)
    keep, count = m.non_max_suppression(*to_dev(boxes, scores), max_out, thr)
    o = oracle.non_max_suppression(boxes, scores, max_out, thr)
    k = keep.cpu().numpy()
    assert int(count) == o.shape[0]
    assert np.array_equal(k[:o.shape[0]], o)                    # bit-exact keep list, selection order
    assert np.all(k[o.shape[0]:] == -1)


def test_nms_threshold_boundary_and_ties():
    """IoU exactly at the threshold is kept (strict >); equal scores resolve to the lower index."""
    m = _m()
    # two unit squares shifted by 0.5: inter 0.5, union 1.5, IoU = 1/3 (fp32 0.33333334)
    boxes = np.array([[0, 0, 1, 1], [0, 0.5, 1, 1.5], [0, 0, 1, 1], [5, 5, 6, 6]], np.float32)
    scores = np.array([0.9, 0.8, 0.9, 0.1], np.float32)
    thr = float(oracle.iou_tf(boxes[0], boxes[1:2])[0])
    keep, count = m.non_max_suppression(*to_dev(boxes, scores), 10, thr)
    assert keep.cpu().numpy()[:int(count)].tolist() == oracle.non_max_suppression(boxes, scores, 10, thr).tolist() == [0, 1, 3]
    keep, count = m.non_max_suppression(*to_dev(boxes, scores), 10, np.nextafter(np.float32(thr), np.float32(0)))
    assert keep.cpu().numpy()[:int(count)].tolist() == [0, 3]
    # flipped corners are normalised, degenerate boxes never suppress
    boxes = np.array([[1, 1, 0, 0], [0, 0, 1, 1], [0.5, 0.5, 0.5, 0.9]], np.float32)
    scores = np.array([0.5, 0.6, 0.7], np.float32)
    keep, count = m.non_max_suppression(*to_dev(boxes, scores), 10, 0.5)
    assert keep.cpu().numpy()[:int(count)].tolist() == oracle.non_max_suppression(boxes, scores, 10, 0.5).tolist() == [2, 1]


@pytest.mark.parametrize("N,K,min_conf", [(1000, 23, 0.7), (1000, 23, 0.0), (300, 41, 0.3), (64, 2, 0.5)])
def test_refine_detections(N, K, min_conf):
    m = _m()
    rng = np.random.default_rng(N + K)
    cfg = small_cfg(DETECTION_MIN_CONFIDENCE=min_conf, NUM_CLASSES=K)
    rois = syn.make_rois(rng, 1, N)[0]
    probs, deltas = syn.make_detection_inputs(rng, N, K)
    if min_conf:
        probs[:N // 3] = 0; probs[np.arange(N // 3), rng.integers(1, K, N // 3)] = rng.uniform(0.5, 1.0, N // 3).astype(np.float32)
        deltas *= 0.2
    window = np.array([0.05, 0.0, 0.95, 1.0], np.float32)
    det, keep, count = m.refine_detections_graph(*to_dev(rois, probs, deltas, window), cfg, return_keep=True)
    o_det, o_keep = oracle.refine_detections(rois, probs, deltas, window, cfg)
    n = o_keep.shape[0]
    assert n > 0 and int(count) == n
    assert np.array_equal(keep.cpu().numpy()[:n], o_keep)       # bit-exact keep indices
    assert np.array_equal(det.cpu().numpy(), o_det)             # boxes/class/score rows bit-exact


def test_detection_layer_batched():
    m = _m()
    rng = np.random.default_rng(5)
    B, N, K = 3, 200, 10
    cfg = small_cfg(DETECTION_MIN_CONFIDENCE=0.2, NUM_CLASSES=K, IMAGES_PER_GPU=B, IMAGE_SHAPE=np.array([128, 160, 3]))
    rois = syn.make_rois(rng, B, N)
    pd = [syn.make_detection_inputs(rng, N, K) for _ in range(B)]
    probs = np.stack([p for p, _ in pd]); deltas = np.stack([d for _, d in pd]) * 0.3
    meta = syn.make_image_meta(B, (128, 160, 3), K, window=(10, 0, 118, 160))
    out = m.DetectionLayer(cfg)(to_dev(rois, probs, deltas) + [meta])
    o = oracle.detection_layer(rois, probs, deltas, meta, cfg)
    assert out.shape == (B, 100, 6)
    assert np.array_equal(out.cpu().numpy(), o)


@pytest.mark.parametrize("hw,limit,count", [((128, 128), 600, 100), ((256, 320), 6000, 1000)])
def test_proposal_layer(hw, limit, count):
    m = _m()
    rng = np.random.default_rng(hw[0])
    B = 2
    anchors = syn.make_anchors(hw)
    A = anchors.shape[0]
    cfg = small_cfg(PRE_NMS_LIMIT=limit, IMAGES_PER_GPU=B)
    fg = rng.permutation(B * A).reshape(B, A).astype(np.float32) / (B * A)      # distinct scores
    probs = np.stack([1 - fg, fg], axis=-1).astype(np.float32)
    bbox = rng.normal(0, 0.5, (B, A, 4)).astype(np.float32)
    anc = np.broadcast_to(anchors, (B, A, 4)).copy()
    out = m.ProposalLayer(count, 0.7, cfg)(to_dev(probs, bbox, anc))
    o = oracle.proposal_layer(probs, bbox, anc, count, 0.7, cfg)
    assert np.array_equal(out.cpu().numpy(), o)


@pytest.mark.parametrize("B,X,Y,Z,C", [(1, 6, 5, 7, 8), (2, 8, 8, 8, 16), (1, 4, 4, 4, 36)])
def test_convlstm_step_and_sequence(B, X, Y, Z, C):
    m = _m()
    rng = np.random.default_rng(C)
    F = C
    V = 3
    grids = rng.standard_normal((B, V, X, Y, Z, C)).astype(np.float32)
    W = (rng.standard_normal((3, 3, 3, C + F, 4 * F)) * np.sqrt(2.0 / (27 * (C + F) + 4 * F))).astype(np.float32)
    b = rng.normal(0, 0.1, 4 * F).astype(np.float32)
    dW, db = to_dev(W, b)
    x0 = to_dev(grids[:, 0])[0]
    h, c = m.convlstm_step(x0, None, None, dW, db)
    oh, oc = oracle.convlstm_cell_step(grids[:, 0], np.zeros((B, X, Y, Z, F), np.float32), np.zeros((B, X, Y, Z, F), np.float32), W, b)
    # fp32 accumulation over K = 27*(C+F) terms vs the float64 oracle: 1e-5 relative + 2e-6 absolute
    close(h.cpu().numpy(), oh, rtol=1e-5, atol=2e-6)
    close(c.cpu().numpy(), oc, rtol=1e-5, atol=2e-6)
    # full recurrence through grid_reas('lstm3d'): ReLU -> ConvLSTM over views -> BN -> ReLU
    cfg = small_cfg(GRID_REAS="lstm3d", TOP_DOWN_PYRAMID_SIZE=F, nvox=X, nvox_z=Z)
    bn = (np.full(F, 1.1, np.float32), np.full(F, 0.02, np.float32), np.full(F, -0.01, np.float32), np.full(F, 0.9, np.float32))
    out = m.grid_reas(to_dev(grids)[0], "grid_reas_P4", cfg, params={"W": dW, "b": db, "bn": bn})
    o = oracle.grid_reas(grids, "grid_reas_P4", cfg, {"W": W, "b": b, "bn": bn})
    close(out.cpu().numpy(), o, rtol=1e-5, atol=5e-6)


def test_ident_fuse():
    m = _m()
    rng = np.random.default_rng(9)
    B, V, X, C, Cout = 1, 3, 6, 16, 24
    cfg = small_cfg(GRID_REAS="ident", NUM_VIEWS=V, nvox=X, nvox_z=X)
    grids = rng.standard_normal((B, V, X, X, X, C)).astype(np.float32)
    W = (rng.standard_normal((V * C, Cout)) * 0.2).astype(np.float32)
    b = rng.normal(0, 0.1, Cout).astype(np.float32)
    bn = (np.full(Cout, 0.9, np.float32), np.full(Cout, 0.1, np.float32), np.zeros(Cout, np.float32), np.ones(Cout, np.float32))
    dg, dW, db = to_dev(grids, W, b)
    out = m.grid_reas(dg, "grid_reas_P4", cfg, params={"weight": dW, "bias": db, "bn": bn})
    o = oracle.grid_reas(grids, "grid_reas_P4", cfg, {"weight": W, "bias": b, "bn": bn})
    close(out.cpu().numpy(), o, rtol=1e-5, atol=5e-6)
