"""CPU oracle of the full ``model_multi`` inference graph.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Restates mrcnn/model_multi.py:2382-2553 (the 'inference' branch of ``MaskRCNN.build``): the TimeDistributed ResNet + FPN
(:497-641), the fusion neck (oracle.fusion_neck), the RPN (:1265-1306), ProposalLayer, the classifier head (:1335-1388),
DetectionLayer and the mask head (:1391-1444).  The dense 2-D layers are evaluated in float64 on the CPU (torch) from
the Keras definitions -- third-party kernels (Conv2D 'same'/'valid', MaxPool2D 'same', UpSampling2D, Conv2DTranspose,
BatchNormalization eps 1e-3), restated from their published semantics: parity with real Keras/TF bits unpinned; the WIRING is
pinned by tests/golden/model_graphs.npz, which make_golden.py produces by executing the reference's own graph builders.
Parameters: the dictionary of ``model_host.init_params`` (Keras layer names).  Every stage can be fed either the oracle's own
upstream values or the product's (stage-wise comparison: the discrete stages -- top-k, NMS -- are then compared on identical
inputs)."""
import math

import numpy as np
import torch
import torch.nn.functional as F

from .fusion import batch_norm_affine, fusion_neck
from .roi_align import pyramid_roi_align
from .detection import proposal_layer, detection_layer

F32 = np.float32


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(np.asarray(a), dtype=np.float64))


def _same(n, k, s):
    out = -(-n // s)                                   # TF SAME: ceil(n / s) outputs, the odd padding element at the END
    tot = max((out - 1) * s + k - n, 0)
    return tot // 2, tot - tot // 2


def conv2d(x, p, stride=1, padding="valid"):
    """keras Conv2D on channel-last x [N,H,W,C]; kernel [kh,kw,cin,cout]."""
    k, b = _t(p["kernel"]), _t(p["bias"])
    t = x.permute(0, 3, 1, 2)
    if padding == "same":
        (pt, pb), (pl, pr) = _same(t.shape[2], k.shape[0], stride), _same(t.shape[3], k.shape[1], stride)
        t = F.pad(t, (pl, pr, pt, pb))
    return F.conv2d(t, k.permute(3, 2, 0, 1), b, stride=stride).permute(0, 2, 3, 1)


def bn(x, p):
    scale, shift = batch_norm_affine(*p["bn"])         # fp32 affine as tf.nn.batch_normalization folds it
    return x * _t(scale) + _t(shift)


def resnet_blocks(architecture):
    """(stage, block, stride of the block's first conv, conv_block?) in ``resnet_graph`` order (:584-605): conv_block first in
    every stage (stride 1 in stage 2, else 2), then identity blocks b, c, ... (5 / 22 of them in stage 4)."""
    n4 = {"resnet50": 5, "resnet101": 22}[architecture]
    out = []
    for stage, n in ((2, 3), (3, 4), (4, n4 + 1), (5, 3)):
        for i in range(n):
            out.append((stage, chr(97 + i), (1 if stage == 2 else 2) if i == 0 else 1, i == 0))
    return out


def resnet_fpn(images, params, cfg):
    """``build_resnet_fpn`` (:609-641) -> [P2..P6], each [B,V,h,w,D] float32."""
    images = np.asarray(images)
    B, V = images.shape[:2]
    x = _t(images.reshape((B * V,) + images.shape[2:]))
    x = F.pad(x.permute(0, 3, 1, 2), (3, 3, 3, 3)).permute(0, 2, 3, 1)                                  # ZeroPadding2D((3,3)) :579
    x = torch.relu(bn(conv2d(x, params["conv1"], 2), params["bn_conv1"]))                               # :580-582
    t = x.permute(0, 3, 1, 2)
    (pt, pb), (pl, pr) = _same(t.shape[2], 3, 2), _same(t.shape[3], 3, 2)
    x = F.max_pool2d(F.pad(t, (pl, pr, pt, pb), value=float("-inf")), 3, 2).permute(0, 2, 3, 1)         # MaxPool 'same' :583
    stages, last = [], 2
    for stage, blk, stride, shortcut in resnet_blocks(getattr(cfg, "BACKBONE", "resnet101")):
        if stage != last:
            stages.append(x)
            last = stage
        cb, bb = "res%d%s_branch" % (stage, blk), "bn%d%s_branch" % (stage, blk)
        y = torch.relu(bn(conv2d(x, params[cb + "2a"], stride), params[bb + "2a"]))                     # :521-523 / :555-557
        y = torch.relu(bn(conv2d(y, params[cb + "2b"], 1, "same"), params[bb + "2b"]))
        y = bn(conv2d(y, params[cb + "2c"]), params[bb + "2c"])
        sc = bn(conv2d(x, params[cb + "1"], stride), params[bb + "1"]) if shortcut else x               # :564-565
        x = torch.relu(y + sc)
    stages.append(x)
    C2, C3, C4, C5 = stages
    up = lambda a: a.repeat_interleave(2, dim=1).repeat_interleave(2, dim=2)                            # UpSampling2D
    P5 = conv2d(C5, params["fpn_c5p5"])
    P4 = up(P5) + conv2d(C4, params["fpn_c4p4"])
    P3 = up(P4) + conv2d(C3, params["fpn_c3p3"])
    P2 = up(P3) + conv2d(C2, params["fpn_c2p2"])
    P2, P3, P4, P5 = (torch.relu(conv2d(p, params[n], 1, "same")) for p, n in ((P2, "fpn_p2"), (P3, "fpn_p3"), (P4, "fpn_p4"), (P5, "fpn_p5")))
    P6 = torch.relu(P5[:, ::2, ::2])                                                                    # MaxPool2D(1, strides 2) :638-639
    return [p.numpy().astype(F32).reshape((B, V) + tuple(p.shape[1:])) for p in (P2, P3, P4, P5, P6)]


def rpn(feature_map, params, anchor_stride=1):
    """``rpn_graph`` (:1265-1306) -> (logits [B,N,2], probs, bbox [B,N,4])."""
    x = _t(feature_map)
    shared = torch.relu(conv2d(x, params["rpn_conv_shared"], anchor_stride, "same"))
    logits = conv2d(shared, params["rpn_class_raw"]).reshape(x.shape[0], -1, 2)
    bbox = conv2d(shared, params["rpn_bbox_pred"]).reshape(x.shape[0], -1, 4)
    return logits.numpy().astype(F32), torch.softmax(logits, -1).numpy().astype(F32), bbox.numpy().astype(F32)


def classifier(rois, feature_maps, image_meta, params, cfg):
    """``fpn_classifier_graph`` (:1335-1388) -> (logits, probs [B,R,K], bbox [B,R,K,4])."""
    ps, K = int(cfg.POOL_SIZE), int(cfg.NUM_CLASSES)
    x = pyramid_roi_align(rois, np.asarray(image_meta)[0, 4:7], feature_maps, (ps, ps))
    B, R = x.shape[:2]
    x = _t(x.reshape((B * R,) + x.shape[2:]))
    x = torch.relu(bn(conv2d(x, params["mrcnn_class_conv1"]), params["mrcnn_class_bn1"]))
    x = torch.relu(bn(conv2d(x, params["mrcnn_class_conv2"]), params["mrcnn_class_bn2"]))
    shared = x[:, 0, 0, :]                                                                              # pool_squeeze :1374
    logits = shared @ _t(params["mrcnn_class_logits"]["kernel"]) + _t(params["mrcnn_class_logits"]["bias"])
    bbox = shared @ _t(params["mrcnn_bbox_fc"]["kernel"]) + _t(params["mrcnn_bbox_fc"]["bias"])
    return (logits.reshape(B, R, K).numpy().astype(F32), torch.softmax(logits, -1).reshape(B, R, K).numpy().astype(F32),
            bbox.reshape(B, R, K, 4).numpy().astype(F32))


def mask_head(rois, feature_maps, image_meta, params, cfg):
    """``build_fpn_mask_graph`` (:1391-1444) -> [B,N,2*ps,2*ps,K]."""
    ps = int(cfg.MASK_POOL_SIZE)
    x = pyramid_roi_align(rois, np.asarray(image_meta)[0, 4:7], feature_maps, (ps, ps))
    B, N = x.shape[:2]
    x = _t(x.reshape((B * N,) + x.shape[2:]))
    for i in range(1, 5):
        x = torch.relu(bn(conv2d(x, params["mrcnn_mask_conv%d" % i], 1, "same"), params["mrcnn_mask_bn%d" % i]))
    kd = _t(params["mrcnn_mask_deconv"]["kernel"]).permute(3, 2, 0, 1)                                  # keras [kh,kw,out,in]
    x = torch.relu(F.conv_transpose2d(x.permute(0, 3, 1, 2), kd, _t(params["mrcnn_mask_deconv"]["bias"]), stride=2)).permute(0, 2, 3, 1)
    x = torch.sigmoid(conv2d(x, params["mrcnn_mask"]))
    return x.reshape((B, N) + tuple(x.shape[1:])).numpy().astype(F32)


def predict(images, image_meta, anchors, Rcam, Kmat, params, cfg, given=None):
    """The inference graph, stage by stage.  ``given``: optional dict of upstream values to use INSTEAD of the oracle's own
    (keys 'P', 'maps', 'rpn_class', 'rpn_bbox', 'rpn_rois', 'mrcnn_class', 'mrcnn_bbox', 'detections')."""
    g = given or {}
    out = {}
    out["P"] = resnet_fpn(images, params, cfg)
    P = g.get("P", out["P"])
    if getattr(cfg, "VANILLA", False):
        B, D, z = P[0].shape[0], int(cfg.TOP_DOWN_PYRAMID_SIZE), int(cfg.IMAGE_SHAPE[0]) // 4
        out["maps"] = [np.zeros((B, z, z, D), F32)] * 2 + [np.ascontiguousarray(p[:, 0]) for p in P[2:]]
    else:
        out["maps"] = fusion_neck(P, np.asarray(Rcam, F32), np.asarray(Kmat, F32), cfg, params)
    maps = g.get("maps", out["maps"])
    r = [rpn(m, params) for m in maps]
    out["rpn_class"], out["rpn_bbox"] = np.concatenate([a[1] for a in r], 1), np.concatenate([a[2] for a in r], 1)
    out["rpn_rois"] = proposal_layer(g.get("rpn_class", out["rpn_class"]), g.get("rpn_bbox", out["rpn_bbox"]), anchors,
                                     int(cfg.POST_NMS_ROIS_INFERENCE), float(cfg.RPN_NMS_THRESHOLD), cfg)
    rois = g.get("rpn_rois", out["rpn_rois"])
    _, out["mrcnn_class"], out["mrcnn_bbox"] = classifier(rois, maps[:4], image_meta, params, cfg)
    out["detections"] = detection_layer(rois, g.get("mrcnn_class", out["mrcnn_class"]), g.get("mrcnn_bbox", out["mrcnn_bbox"]), image_meta, cfg)
    det = g.get("detections", out["detections"])
    out["mrcnn_mask"] = mask_head(np.ascontiguousarray(det[..., :4]), maps[:4], image_meta, params, cfg)
    return out
