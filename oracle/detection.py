"""Oracle for box refinement, greedy NMS, ``refine_detections_graph`` / ``DetectionLayer`` and
``ProposalLayer`` (test infrastructure, see oracle/__init__.py).

Restates mrcnn/model_multi.py:648-687 (box helpers), :1119-1258 (detection) and :690-767
(proposals).  ``tf.image.non_max_suppression`` and ``tf.nn.top_k`` are third-party TensorFlow
kernels, restated from their published algorithms (SURVEY.md spec F): greedy selection in
descending score order (stable: ties -> lower index first), IoU on min/max-normalised
corners, ``area <= 0 -> IoU = 0``, suppress iff ``IoU > threshold`` (strict).
"""
import numpy as np

from .geometry import F32


def exp_f32(x):
    """fp32 exp pinned as the correctly rounded value (float64 exp rounded once)."""
    with np.errstate(all="ignore"):
        return np.exp(np.asarray(x, dtype=np.float64)).astype(F32)


def apply_box_deltas(boxes, deltas):
    """``apply_box_deltas_graph`` (model_multi.py:648-669), fp32, same statement order."""
    boxes = np.asarray(boxes, dtype=F32)
    deltas = np.asarray(deltas, dtype=F32)
    half = F32(0.5)
    with np.errstate(all="ignore"):
        height = boxes[:, 2] - boxes[:, 0]
        width = boxes[:, 3] - boxes[:, 1]
        center_y = boxes[:, 0] + half * height
        center_x = boxes[:, 1] + half * width
        center_y = center_y + deltas[:, 0] * height
        center_x = center_x + deltas[:, 1] * width
        height = height * exp_f32(deltas[:, 2])
        width = width * exp_f32(deltas[:, 3])
        y1 = center_y - half * height
        x1 = center_x - half * width
        y2 = y1 + height
        x2 = x1 + width
    return np.stack([y1, x1, y2, x2], axis=1).astype(F32)


def clip_boxes(boxes, window):
    """``clip_boxes_graph`` (model_multi.py:672-687): max(min(v, hi), lo) per coordinate."""
    boxes = np.asarray(boxes, dtype=F32)
    wy1, wx1, wy2, wx2 = (F32(w) for w in window)
    y1 = np.maximum(np.minimum(boxes[:, 0], wy2), wy1)
    x1 = np.maximum(np.minimum(boxes[:, 1], wx2), wx1)
    y2 = np.maximum(np.minimum(boxes[:, 2], wy2), wy1)
    x2 = np.maximum(np.minimum(boxes[:, 3], wx2), wx1)
    return np.stack([y1, x1, y2, x2], axis=1).astype(F32)


def norm_boxes(boxes, shape):
    """``norm_boxes_graph`` (model_multi.py:3390-3405): (boxes - [0,0,1,1]) / ([h,w,h,w]-1)."""
    h, w = F32(shape[0]), F32(shape[1])
    scale = np.array([h, w, h, w], dtype=F32) - F32(1.0)
    shift = np.array([0.0, 0.0, 1.0, 1.0], dtype=F32)
    return ((np.asarray(boxes, dtype=F32) - shift) / scale).astype(F32)


def iou_tf(box_i, boxes_j):
    """IoU of one box against many with TensorFlow's NMS arithmetic (fp32)."""
    bi = np.asarray(box_i, dtype=F32)
    bj = np.asarray(boxes_j, dtype=F32).reshape(-1, 4)
    ymin_i, ymax_i = min(bi[0], bi[2]), max(bi[0], bi[2])
    xmin_i, xmax_i = min(bi[1], bi[3]), max(bi[1], bi[3])
    ymin_j = np.minimum(bj[:, 0], bj[:, 2])
    ymax_j = np.maximum(bj[:, 0], bj[:, 2])
    xmin_j = np.minimum(bj[:, 1], bj[:, 3])
    xmax_j = np.maximum(bj[:, 1], bj[:, 3])
    area_i = F32(F32(ymax_i - ymin_i) * F32(xmax_i - xmin_i))
    area_j = (ymax_j - ymin_j) * (xmax_j - xmin_j)
    iy0 = np.maximum(ymin_i, ymin_j)
    ix0 = np.maximum(xmin_i, xmin_j)
    iy1 = np.minimum(ymax_i, ymax_j)
    ix1 = np.minimum(xmax_i, xmax_j)
    inter = np.maximum(iy1 - iy0, F32(0)) * np.maximum(ix1 - ix0, F32(0))
    with np.errstate(all="ignore"):
        iou = inter / ((area_i + area_j) - inter)
    bad = (area_j <= 0) | (area_i <= 0)
    return np.where(bad, F32(0), iou).astype(F32)


def score_order(scores):
    """Descending score, ties -> lower index first (stable).  NaN scores sort last."""
    scores = np.asarray(scores, dtype=F32)
    key = np.where(np.isnan(scores), -np.inf, scores)
    return np.argsort(-key, kind="stable")


def non_max_suppression(boxes, scores, max_output_size, iou_threshold):
    """``tf.image.non_max_suppression`` -> indices (int32) into ``boxes`` in selection order."""
    boxes = np.asarray(boxes, dtype=F32).reshape(-1, 4)
    order = score_order(scores)
    thr = F32(iou_threshold)
    keep = []
    for idx in order:
        if len(keep) >= max_output_size:
            break
        if keep:
            iou = iou_tf(boxes[idx], boxes[np.array(keep)])
            if np.any(iou > thr):
                continue
        keep.append(int(idx))
    return np.array(keep, dtype=np.int32)


def refine_detections(rois, probs, deltas, window, cfg):
    """``refine_detections_graph(rois, probs, deltas, window, config)`` (model_multi.py:1119-1214).

    rois [N,4], probs [N,K], deltas [N,K,4], window [4] -> [DETECTION_MAX_INSTANCES, 6]
    rows (y1,x1,y2,x2,class_id,score), zero padded.  Also returns the kept roi indices
    (int32, in output order) as the second value."""
    rois = np.asarray(rois, dtype=F32)
    probs = np.asarray(probs, dtype=F32)
    deltas = np.asarray(deltas, dtype=F32)
    N = rois.shape[0]
    max_inst = cfg.DETECTION_MAX_INSTANCES
    class_ids = np.argmax(probs, axis=1).astype(np.int32)                   # :1135 (first max)
    ar = np.arange(N)
    class_scores = probs[ar, class_ids]                                     # :1138
    deltas_specific = deltas[ar, class_ids]                                 # :1140
    std = np.asarray(cfg.BBOX_STD_DEV, dtype=F32)
    refined = apply_box_deltas(rois, deltas_specific * std)                 # :1143-1144
    refined = clip_boxes(refined, window)                                   # :1146
    keep = np.nonzero(class_ids > 0)[0]                                     # :1151
    if cfg.DETECTION_MIN_CONFIDENCE:
        conf_keep = np.nonzero(class_scores >= F32(cfg.DETECTION_MIN_CONFIDENCE))[0]
        keep = np.intersect1d(keep, conf_keep)                              # :1153-1157 (sorted)
    pre_ids = class_ids[keep]
    pre_scores = class_scores[keep]
    pre_rois = refined[keep]
    nms_keep = []
    seen = []
    for cid in pre_ids:                                                     # tf.unique order
        if cid not in seen:
            seen.append(cid)
    for cid in seen:                                                        # :1166-1187
        ixs = np.nonzero(pre_ids == cid)[0]
        ck = non_max_suppression(pre_rois[ixs], pre_scores[ixs], max_inst,
                                 cfg.DETECTION_NMS_THRESHOLD)
        nms_keep.extend(keep[ixs[ck]].tolist())
    keep = np.intersect1d(keep, np.array(nms_keep, dtype=np.int64)).astype(np.int64)   # :1193
    scores_keep = class_scores[keep]
    num_keep = min(scores_keep.shape[0], max_inst)
    top = np.argsort(-scores_keep, kind="stable")[:num_keep]                # :1200 top_k
    keep = keep[top]
    det = np.zeros((max_inst, 6), dtype=F32)
    det[:num_keep, :4] = refined[keep]
    det[:num_keep, 4] = class_ids[keep].astype(F32)                         # :1207
    det[:num_keep, 5] = class_scores[keep]
    return det, keep.astype(np.int32)


def detection_layer(rois, mrcnn_class, mrcnn_bbox, image_meta, cfg):
    """``DetectionLayer.call`` (model_multi.py:1233-1255): window from ``image_meta`` columns
    7:11 normalised by the first image's shape (columns 4:6), then per-scene refinement."""
    image_meta = np.asarray(image_meta)
    image_shape = image_meta[0, 4:7]
    windows = norm_boxes(image_meta[:, 7:11], image_shape[:2])
    B = rois.shape[0]
    out = np.stack([refine_detections(rois[b], mrcnn_class[b], mrcnn_bbox[b], windows[b], cfg)[0]
                    for b in range(B)])
    return out.reshape(B, cfg.DETECTION_MAX_INSTANCES, 6)


def proposal_layer(rpn_probs, rpn_bbox, anchors, proposal_count, nms_threshold, cfg):
    """``ProposalLayer(proposal_count, nms_threshold, config)([probs, bbox, anchors])``
    (model_multi.py:690-767): fg score -> top-k PRE_NMS_LIMIT -> deltas*RPN_BBOX_STD_DEV ->
    apply -> clip to [0,1] -> NMS -> zero pad.  -> [B, proposal_count, 4]."""
    rpn_probs = np.asarray(rpn_probs, dtype=F32)
    rpn_bbox = np.asarray(rpn_bbox, dtype=F32)
    anchors = np.asarray(anchors, dtype=F32)
    B, A = rpn_probs.shape[:2]
    std = np.asarray(cfg.RPN_BBOX_STD_DEV, dtype=F32).reshape(1, 1, 4)
    scores = rpn_probs[:, :, 1]
    deltas = rpn_bbox * std
    limit = min(cfg.PRE_NMS_LIMIT, A)
    out = np.zeros((B, proposal_count, 4), dtype=F32)
    window = np.array([0, 0, 1, 1], dtype=F32)
    for b in range(B):
        ix = score_order(scores[b])[:limit]                                 # :723 top_k sorted
        boxes = clip_boxes(apply_box_deltas(anchors[b][ix], deltas[b][ix]), window)
        keep = non_max_suppression(boxes, scores[b][ix], proposal_count, nms_threshold)
        out[b, :keep.shape[0]] = boxes[keep]
    return out
