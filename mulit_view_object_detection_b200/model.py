"""Full ``model_multi`` inference (BASELINE config c4): the reference's ``MaskRCNN`` in
'inference' mode (mrcnn/model_multi.py:2300-2560 ``build``, :3019-3082 ``detect``) around the B200 hot path.

Everything on the fusion path runs on this repo's CUDA kernels through the C-ABI (``layers``): ``unproj_feat`` +
``grid_reas`` + ``proj_grid`` + ``depth_sampling`` (``fusion_neck``), ``ProposalLayer``, ``PyramidROIAlign`` (classifier and
mask heads) and ``DetectionLayer``.  The dense 2-D convolutions around it -- the TimeDistributed ResNet-50/101 + FPN
(:497-641), the RPN (:1265-1332) and the two heads (:1335-1444) -- are plain library calls (cuDNN / cuBLAS through
``torch.nn.functional``, fp32 with TF32 off unless asked): SURVEY.md ranks them as having no custom-kernel upside.  They are
written against the Keras layer NAMES of the reference, so a parameter dict keyed by those names (``weights_io.read_npz``)
loads as is.

Tensors are channel-last at every interface the reference has ([B,V,H,W,C] feature maps, [B,R,ph,pw,C] crops); the dense
blocks run on channels_last NCHW views of the same memory.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

from . import layers as L
from .weights_io import compose_image_meta

from .model_host import (BN_EPS, MODEL_DEFAULTS, _cfg, _bn_p, compute_backbone_shapes, generate_pyramid_anchors, norm_boxes,   # noqa: F401
                         denorm_boxes, resize_image, mold_image, unmold_mask, resnet_blocks, init_params, named_weights)


# ---------------------------------------------------------------------------------------------------------------------
class _Dense:
    """The dense 2-D blocks on torch tensors (device-agnostic): parameters converted once to the layouts cuDNN wants."""

    def __init__(self, params, device):
        self.dev = device
        self.w, self.bn = {}, {}
        for name, p in params.items():
            if "kernel" in p:
                k = torch.as_tensor(np.asarray(p["kernel"]), dtype=torch.float32, device=device)
                b = torch.as_tensor(np.asarray(p["bias"]), dtype=torch.float32, device=device)
                if k.dim() == 4 and name != "mrcnn_mask_deconv":
                    k = k.permute(3, 2, 0, 1).contiguous(memory_format=torch.channels_last)      # HWIO -> OIHW
                elif k.dim() == 4:
                    k = k.permute(3, 2, 0, 1).contiguous()      # Conv2DTranspose kernel [kh,kw,out,in] -> torch [in,out,kh,kw]
                self.w[name] = (k, b)
                if name == "mrcnn_class_conv1":                 # 'valid' ps x ps conv over a ps x ps crop == one GEMM over (h, w, c)
                    self.gemm1 = k.permute(2, 3, 1, 0).reshape(-1, k.shape[0]).contiguous()
                if name == "mrcnn_class_conv2":
                    self.gemm2 = k.reshape(k.shape[0], -1).t().contiguous()
            elif "bn" in p and not name.startswith("grid_reas"):
                g, be, m, v = (torch.as_tensor(np.asarray(a), dtype=torch.float32, device=device).reshape(-1) for a in p["bn"])
                inv = torch.rsqrt(v + BN_EPS) * g
                self.bn[name] = (inv.contiguous(), (be - m * inv).contiguous())

    def conv(self, x, name, stride=1, pad=0):
        k, b = self.w[name]
        return F.conv2d(x, k, b, stride=stride, padding=pad)

    def bnorm(self, x, name, relu=True, add=None):
        s, t = self.bn[name]
        shape = (1, -1, 1, 1) if x.dim() == 4 else (1,) * (x.dim() - 1) + (-1,)
        y = x * s.view(shape) + t.view(shape)
        if add is not None:
            y = y + add
        return F.relu_(y) if relu else y


def _dense_precision(fn):
    """Run a graph piece with the model's dense-layer precision: cuDNN / cuBLAS fp32 (TF32 off, the reference computes in fp32)
    unless the model was built with ``allow_tf32=True``.  torch's global switches are restored afterwards."""
    import functools

    @functools.wraps(fn)
    def wrapped(self, *a, **k):
        old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
        torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = self.allow_tf32
        try:
            with torch.no_grad():
                return fn(self, *a, **k)
        finally:
            torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
    return wrapped


def _maxpool_same_3x3_s2(x):
    """KL.MaxPool2D((3,3), strides=(2,2), padding='same') (:581): TF SAME puts the odd padding element at the END."""
    H, W = x.shape[-2:]
    ph = max((math.ceil(H / 2) - 1) * 2 + 3 - H, 0)
    pw = max((math.ceil(W / 2) - 1) * 2 + 3 - W, 0)
    x = F.pad(x, (pw // 2, pw - pw // 2, ph // 2, ph - ph // 2), value=float("-inf"))
    return F.max_pool2d(x, 3, 2)


class MaskRCNN:
    """``MaskRCNN(mode='inference', config, model_dir)`` (model_multi.py:2300-2560): ``detect(images, Rcam, Kmat)`` and
    ``predict([molded_images, image_metas, anchors, Rcam, Kmat])`` with the reference's outputs
    [detections, mrcnn_class, mrcnn_bbox, mrcnn_mask, rpn_rois, rpn_class, rpn_bbox]."""

    def __init__(self, mode, config, model_dir=None, params=None, device=None, seed=0, allow_tf32=False):
        if mode != "inference":
            raise ValueError("only mode='inference' is built (SURVEY.md section 8(f) rank 4); training is out of scope")
        h, w = (int(v) for v in config.IMAGE_SHAPE[:2])
        if h / 2 ** 6 != int(h / 2 ** 6) or w / 2 ** 6 != int(w / 2 ** 6):                      # :2328-2332
            raise Exception("Image size must be dividable by 2 at least 6 times to avoid fractions when downscaling and upscaling."
                            "For example, use 256, 320, 384, 448, 512, ... etc. ")
        self.mode, self.config, self.model_dir = mode, config, model_dir
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.allow_tf32 = bool(allow_tf32)
        self.set_params(params if params is not None else init_params(config, seed))
        self._anchor_cache = {}

    # ---- parameters -------------------------------------------------------------------------------------------------
    def set_params(self, params):
        self.params = params
        self.dense = _Dense(params, self.device)
        # the neck's learnables once as device tensors / folded BatchNorm affines (layers.prepare_params)
        neck = {k: v for k, v in params.items() if k.startswith("grid_reas")}
        self.neck_params = L.prepare_params(neck, self.device) if self.device.type == "cuda" else neck

    def load_weights(self, filepath, by_name=True, exclude=None):
        """``load_weights`` (model_multi.py:2592-2642): Keras layers restored BY NAME from ``{layer: get_weights() list}`` --
        an ``.npz`` export (``weights_io.read_npz``) or, when h5py is importable, the Keras HDF5 file itself.  Conv2D / Dense:
        [kernel, bias]; BatchNorm: [gamma, beta, moving_mean, moving_variance]; the fusion-neck layers through
        ``weights_io.fusion_params_from_keras``.  Layers absent from the file keep their current values (by_name semantics)."""
        from . import weights_io
        named = weights_io.read_keras_hdf5(filepath) if str(filepath).endswith((".h5", ".hdf5")) else weights_io.read_npz(filepath)
        self.set_named_weights(named, exclude)

    def set_named_weights(self, named, exclude=None):
        from . import weights_io
        skip = set(exclude or ())
        params = {k: dict(v) for k, v in self.params.items()}
        for layer, ws in named.items():
            if layer in skip or layer not in params or layer.startswith("grid_reas"):
                continue
            if "kernel" in params[layer]:
                k, b = (np.asarray(a, np.float32) for a in ws[:2])
                if k.shape != params[layer]["kernel"].shape:
                    raise ValueError("layer %r: kernel shape %s != %s" % (layer, k.shape, params[layer]["kernel"].shape))
                params[layer] = {"kernel": k, "bias": b.reshape(-1)}
            elif "bn" in params[layer]:
                params[layer] = {"bn": tuple(np.asarray(a, np.float32).reshape(-1) for a in ws[:4])}
        if any(k.startswith("grid_reas") for k in named if k not in skip):
            params.update(weights_io.fusion_params_from_keras(named, self.config))
        self.set_params(params)

    def named_weights(self):
        return named_weights(self.params)

    # ---- graph pieces ----------------------------------------------------------------------------------------------
    def _block(self, x, stage, blk, stride, shortcut):
        d = self.dense
        cb, bb = "res%d%s_branch" % (stage, blk), "bn%d%s_branch" % (stage, blk)
        y = d.bnorm(d.conv(x, cb + "2a", stride=stride), bb + "2a")
        y = d.bnorm(d.conv(y, cb + "2b", pad=1), bb + "2b")
        y = d.conv(y, cb + "2c")
        sc = d.bnorm(d.conv(x, cb + "1", stride=stride), bb + "1", relu=False) if shortcut else x
        return d.bnorm(y, bb + "2c", relu=True, add=sc)                                          # Add -> relu (:531-533, :567-569)

    def resnet_graph(self, x):
        """``resnet_graph(input_image, architecture, stage5=True)`` (:572-607) on [N,3,H,W]; returns [C2, C3, C4, C5]."""
        d = self.dense
        x = d.bnorm(d.conv(x, "conv1", stride=2, pad=3), "bn_conv1")                            # ZeroPadding2D(3) + 7x7/2 valid
        x = _maxpool_same_3x3_s2(x)
        outs, last = [], 2
        for stage, blk, _, stride, shortcut in resnet_blocks(_cfg(self.config, "BACKBONE")):
            if stage != last:
                outs.append(x)
                last = stage
            x = self._block(x, stage, blk, stride, shortcut)
        outs.append(x)
        return outs

    @_dense_precision
    def build_resnet_fpn(self, input_image):
        """``build_resnet_fpn(input_image, config)`` (:609-641): images [B,V,H,W,3] -> [P2..P6], each [B,V,h,w,D] channel-last."""
        d = self.dense
        B, V = input_image.shape[:2]
        x = input_image.reshape((B * V,) + tuple(input_image.shape[2:])).permute(0, 3, 1, 2).contiguous(memory_format=torch.channels_last)
        C2, C3, C4, C5 = self.resnet_graph(x)
        up = lambda t: F.interpolate(t, scale_factor=2, mode="nearest")                         # UpSampling2D((2,2))
        P5 = d.conv(C5, "fpn_c5p5")
        P4 = up(P5) + d.conv(C4, "fpn_c4p4")
        P3 = up(P4) + d.conv(C3, "fpn_c3p3")
        P2 = up(P3) + d.conv(C2, "fpn_c2p2")
        P2, P3, P4, P5 = (F.relu_(d.conv(p, n, pad=1)) for p, n in ((P2, "fpn_p2"), (P3, "fpn_p3"), (P4, "fpn_p4"), (P5, "fpn_p5")))
        P6 = F.relu(P5[:, :, ::2, ::2])                                                         # MaxPool2D(1, strides=2) + relu
        return [p.permute(0, 2, 3, 1).contiguous().reshape((B, V) + (p.shape[2], p.shape[3], p.shape[1])) for p in (P2, P3, P4, P5, P6)]

    @_dense_precision
    def rpn_graph(self, feature_map):
        """``rpn_graph`` (:1265-1306) on one channel-last map [B,h,w,D] -> (logits [B,h*w*A,2], probs, bbox [B,h*w*A,4])."""
        d = self.dense
        x = feature_map.permute(0, 3, 1, 2)
        shared = F.relu_(d.conv(x, "rpn_conv_shared", stride=int(_cfg(self.config, "RPN_ANCHOR_STRIDE")), pad=1))
        Bn = x.shape[0]
        logits = d.conv(shared, "rpn_class_raw").permute(0, 2, 3, 1).reshape(Bn, -1, 2)
        bbox = d.conv(shared, "rpn_bbox_pred").permute(0, 2, 3, 1).reshape(Bn, -1, 4)
        return logits, torch.softmax(logits, dim=-1), bbox

    @_dense_precision
    def fpn_classifier_graph(self, rois, feature_maps, image_meta):
        """``fpn_classifier_graph`` (:1335-1388): PyramidROIAlign (K4) -> two FC layers as convolutions -> class / box heads."""
        d, cfg = self.dense, self.config
        ps, K = int(cfg.POOL_SIZE), int(cfg.NUM_CLASSES)
        x = L.PyramidROIAlign([ps, ps])([rois, image_meta] + feature_maps)                       # [B,R,ps,ps,D]
        B, R = x.shape[:2]
        x = d.bnorm(x.reshape(B * R, -1) @ d.gemm1 + d.w["mrcnn_class_conv1"][1], "mrcnn_class_bn1")
        x = d.bnorm(x @ d.gemm2 + d.w["mrcnn_class_conv2"][1], "mrcnn_class_bn2")
        wl, bl = d.w["mrcnn_class_logits"]
        logits = (x @ wl + bl).reshape(B, R, K)
        wb, bb = d.w["mrcnn_bbox_fc"]
        bbox = (x @ wb + bb).reshape(B, R, K, 4)
        return logits, torch.softmax(logits, dim=-1), bbox

    @_dense_precision
    def build_fpn_mask_graph(self, rois, feature_maps, image_meta):
        """``build_fpn_mask_graph`` (:1391-1444) -> [B, N, 2*MASK_POOL_SIZE, 2*MASK_POOL_SIZE, NUM_CLASSES]."""
        d, cfg = self.dense, self.config
        ps = int(cfg.MASK_POOL_SIZE)
        x = L.PyramidROIAlign([ps, ps])([rois, image_meta] + feature_maps)
        B, N = x.shape[:2]
        x = x.reshape((B * N,) + tuple(x.shape[2:])).permute(0, 3, 1, 2)                        # channels_last view, no copy
        for i in range(1, 5):
            x = d.bnorm(d.conv(x, "mrcnn_mask_conv%d" % i, pad=1), "mrcnn_mask_bn%d" % i)
        kd, bd = d.w["mrcnn_mask_deconv"]
        x = F.relu_(F.conv_transpose2d(x, kd, bd, stride=2))
        x = torch.sigmoid(d.conv(x, "mrcnn_mask"))
        return x.permute(0, 2, 3, 1).reshape(B, N, x.shape[2], x.shape[3], x.shape[1])

    # ---- the inference graph (:2382-2553) ---------------------------------------------------------------------------
    @_dense_precision
    def predict(self, inputs, return_features=False):
        molded_images, image_metas, anchors, Rcam, Kmat = inputs
        cfg, dev = self.config, self.device
        img = torch.as_tensor(molded_images, dtype=torch.float32, device=dev)
        Rcam = torch.as_tensor(Rcam, dtype=torch.float32, device=dev).contiguous()
        Kmat = torch.as_tensor(Kmat, dtype=torch.float32, device=dev).contiguous()
        anchors = torch.as_tensor(anchors, dtype=torch.float32, device=dev).contiguous()
        metas = np.asarray(image_metas, dtype=np.float32)
        P = self.build_resnet_fpn(img)                                                      # [P2..P6], [B,V,h,w,D]
        if getattr(cfg, "VANILLA", False):                                                  # :2411-2422
            B, D = img.shape[0], int(cfg.TOP_DOWN_PYRAMID_SIZE)
            z = int(cfg.IMAGE_SHAPE[0]) // 4
            zeros = torch.zeros((B, z, z, D), dtype=torch.float32, device=dev)
            maps = [zeros, zeros] + [p[:, 0].contiguous() for p in P[2:]]
        else:
            maps = L.fusion_neck(P, Rcam, Kmat, cfg, params=self.neck_params)                # PG2..PG6 (PG2 / PG3 zeros, :2406-2410)
        rpn_maps, mrcnn_maps = maps, maps[:4]
        outs = [self.rpn_graph(p) for p in rpn_maps]                                        # shared weights over the levels (:2437-2450)
        rpn_class_logits, rpn_class, rpn_bbox = (torch.cat([o[i] for o in outs], dim=1) for i in range(3))
        rpn_rois = L.ProposalLayer(proposal_count=int(cfg.POST_NMS_ROIS_INFERENCE), nms_threshold=float(cfg.RPN_NMS_THRESHOLD),
                                   name="ROI", config=cfg)([rpn_class.contiguous(), rpn_bbox.contiguous(), anchors])
        _, mrcnn_class, mrcnn_bbox = self.fpn_classifier_graph(rpn_rois, mrcnn_maps, metas)
        detections = L.DetectionLayer(cfg, name="mrcnn_detection")([rpn_rois, mrcnn_class.contiguous(), mrcnn_bbox.contiguous(), metas])
        mrcnn_mask = self.build_fpn_mask_graph(detections[..., :4].contiguous(), mrcnn_maps, metas)
        res = [detections, mrcnn_class, mrcnn_bbox, mrcnn_mask, rpn_rois, rpn_class, rpn_bbox]
        return (res, {"P": P, "maps": maps}) if return_features else res

    # ---- host side (:2915-3082, :3142-3162) -------------------------------------------------------------------------
    def get_anchors(self, image_shape):
        key = tuple(int(v) for v in image_shape)
        if key not in self._anchor_cache:
            cfg = self.config
            a = generate_pyramid_anchors(_cfg(cfg, "RPN_ANCHOR_SCALES"), _cfg(cfg, "RPN_ANCHOR_RATIOS"),
                                         compute_backbone_shapes(cfg, image_shape), _cfg(cfg, "BACKBONE_STRIDES"),
                                         _cfg(cfg, "RPN_ANCHOR_STRIDE"))
            self.anchors = a
            self._anchor_cache[key] = norm_boxes(a, image_shape[:2])
        return self._anchor_cache[key]

    def mold_inputs(self, images):
        cfg = self.config
        molded, metas, windows = [], [], []
        for image in images:
            m, window, scale, _, _ = resize_image(image, min_dim=_cfg(cfg, "IMAGE_MIN_DIM"), min_scale=_cfg(cfg, "IMAGE_MIN_SCALE"),
                                                  max_dim=_cfg(cfg, "IMAGE_MAX_DIM"), mode=_cfg(cfg, "IMAGE_RESIZE_MODE"))
            m = mold_image(m, cfg)
            metas.append(compose_image_meta(0, image.shape, m.shape, window, scale, np.zeros([cfg.NUM_CLASSES], dtype=np.int32)))
            molded.append(m)
            windows.append(window)
        return np.stack(molded), np.stack(metas), np.stack(windows)

    def unmold_detections(self, detections, mrcnn_mask, original_image_shape, image_shape, window):
        """:2954-3017."""
        zero_ix = np.where(detections[:, 4] == 0)[0]
        N = zero_ix[0] if zero_ix.shape[0] > 0 else detections.shape[0]
        boxes, class_ids, scores = detections[:N, :4], detections[:N, 4].astype(np.int32), detections[:N, 5]
        masks = mrcnn_mask[np.arange(N), :, :, class_ids]
        window = norm_boxes(window, image_shape[:2])
        wy1, wx1, wy2, wx2 = window
        shift = np.array([wy1, wx1, wy1, wx1])
        wh, ww = wy2 - wy1, wx2 - wx1
        boxes = np.divide(boxes - shift, np.array([wh, ww, wh, ww]))
        boxes = denorm_boxes(boxes, original_image_shape[:2])
        exclude_ix = np.where((boxes[:, 2] - boxes[:, 0]) * (boxes[:, 3] - boxes[:, 1]) <= 0)[0]
        if exclude_ix.shape[0] > 0:
            boxes, class_ids = np.delete(boxes, exclude_ix, axis=0), np.delete(class_ids, exclude_ix, axis=0)
            scores, masks = np.delete(scores, exclude_ix, axis=0), np.delete(masks, exclude_ix, axis=0)
            N = class_ids.shape[0]
        full = [unmold_mask(masks[i], boxes[i], original_image_shape) for i in range(N)]
        full = np.stack(full, axis=-1) if full else np.empty(tuple(original_image_shape[:2]) + (0,))
        return boxes, class_ids, scores, full

    def detect(self, images, Rcam, Kmat, verbose=0):
        """``detect(images, Rcam, Kmat)`` (:3019-3082): images = BATCH_SIZE scenes, each a list of NUM_VIEWS images."""
        cfg = self.config
        assert len(images) == cfg.BATCH_SIZE, "len(images) must be equal to BATCH_SIZE"
        molded, metas, windows = [], [], []
        for scene in images:
            m, meta, win = self.mold_inputs(scene)
            molded.append(m)
            metas.append(meta[0])               # the main view's meta / window describe the scene (the graph reads view 0, :245)
            windows.append(win[0])
        molded = np.stack(molded)
        image_shape = molded.shape[2:]
        anchors = np.broadcast_to(self.get_anchors(image_shape), (cfg.BATCH_SIZE,) + self.get_anchors(image_shape).shape).copy()
        detections, _, _, mrcnn_mask, _, _, _ = self.predict([molded, np.stack(metas), anchors, Rcam, Kmat])
        detections, mrcnn_mask = detections.cpu().numpy(), mrcnn_mask.cpu().numpy()
        results = []
        for i, scene in enumerate(images):
            rois, class_ids, scores, masks = self.unmold_detections(detections[i], mrcnn_mask[i], scene[0].shape, image_shape, windows[i])
            results.append({"rois": rois, "class_ids": class_ids, "scores": scores, "masks": masks})
        return results
