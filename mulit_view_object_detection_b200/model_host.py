"""Host side of the full ``model_multi`` inference path (no CUDA library needed): the reference's pure-NumPy helpers
(mrcnn/utils.py anchors / box normalisation / image resizing, mrcnn/model_multi.py:89-103, :3351-3356) and the parameter
dictionary of the inference graph keyed by the reference's Keras layer names, with Keras' default initialisers."""
import math

import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-3            # keras.layers.BatchNormalization default, which the reference's BatchNorm subclass keeps (:71-86)

# config attributes the full model reads on top of FusionConfig's (mrcnn/config.py:17-236)
MODEL_DEFAULTS = dict(
    BACKBONE="resnet101", BACKBONE_STRIDES=[4, 8, 16, 32, 64], FPN_CLASSIF_FC_LAYERS_SIZE=1024,
    RPN_ANCHOR_SCALES=(32, 64, 128, 256, 512), RPN_ANCHOR_RATIOS=[0.5, 1, 2], RPN_ANCHOR_STRIDE=1,
    IMAGE_RESIZE_MODE="square", IMAGE_MIN_DIM=640, IMAGE_MAX_DIM=640, IMAGE_MIN_SCALE=0,
    MEAN_PIXEL=np.array([123.7, 116.8, 103.9]),
)


def _cfg(config, name):
    return getattr(config, name) if hasattr(config, name) else MODEL_DEFAULTS[name]


# ---------------------------------------------------------------------------------------------------------------------
# host-side helpers (mrcnn/utils.py, pure NumPy in the reference)
def compute_backbone_shapes(config, image_shape):
    """model_multi.py:89-103."""
    return np.array([[int(math.ceil(image_shape[0] / s)), int(math.ceil(image_shape[1] / s))] for s in _cfg(config, "BACKBONE_STRIDES")])


def generate_anchors(scales, ratios, shape, feature_stride, anchor_stride):
    """utils.py:842-878: anchors of one pyramid level, (y1, x1, y2, x2) in pixels, row-major over (y, x, ratio)."""
    scales, ratios = np.meshgrid(np.array(scales), np.array(ratios))
    scales, ratios = scales.flatten(), ratios.flatten()
    heights, widths = scales / np.sqrt(ratios), scales * np.sqrt(ratios)
    ys = np.arange(0, shape[0], anchor_stride) * feature_stride
    xs = np.arange(0, shape[1], anchor_stride) * feature_stride
    xs, ys = np.meshgrid(xs, ys)
    bw, cx = np.meshgrid(widths, xs)
    bh, cy = np.meshgrid(heights, ys)
    centres = np.stack([cy, cx], axis=2).reshape([-1, 2])
    sizes = np.stack([bh, bw], axis=2).reshape([-1, 2])
    return np.concatenate([centres - 0.5 * sizes, centres + 0.5 * sizes], axis=1)


def generate_pyramid_anchors(scales, ratios, feature_shapes, feature_strides, anchor_stride):
    """utils.py:881-900: scale i belongs to level i, every ratio to every level."""
    return np.concatenate([generate_anchors(scales[i], ratios, feature_shapes[i], feature_strides[i], anchor_stride)
                           for i in range(len(scales))], axis=0)


def norm_boxes(boxes, shape):
    """utils.py:1112-1126."""
    h, w = shape
    scale = np.array([h - 1, w - 1, h - 1, w - 1])
    shift = np.array([0, 0, 1, 1])
    return np.divide((boxes - shift), scale).astype(np.float32)


def denorm_boxes(boxes, shape):
    """utils.py:1129-1143."""
    h, w = shape
    scale = np.array([h - 1, w - 1, h - 1, w - 1])
    shift = np.array([0, 0, 1, 1])
    return np.around(np.multiply(boxes, scale) + shift).astype(np.int32)


def _resize_bilinear(image, out_hw):
    """skimage.transform.resize(order=1, mode='constant', preserve_range=True) of utils.py:1146-1172 (third-party, absent here):
    restated as bilinear interpolation with half-pixel centres (torch).  Only reached when an input is not already at the
    network size; parity with skimage bits unpinned."""
    t = torch.from_numpy(np.ascontiguousarray(image, dtype=np.float32))
    t = t.permute(2, 0, 1)[None] if t.dim() == 3 else t[None, None]
    r = F.interpolate(t, size=tuple(int(v) for v in out_hw), mode="bilinear", align_corners=False)
    return (r[0].permute(1, 2, 0) if image.ndim == 3 else r[0, 0]).numpy()


def resize_image(image, min_dim=None, max_dim=None, min_scale=None, mode="square"):
    """utils.py:647-751 ('none', 'square', 'pad64'; 'crop' is training-only)."""
    dtype = image.dtype
    h, w = image.shape[:2]
    window, scale, padding = (0, 0, h, w), 1, [(0, 0), (0, 0), (0, 0)]
    if mode == "none":
        return image, window, scale, padding, None
    if min_dim:
        scale = max(1, min_dim / min(h, w))
    if min_scale and scale < min_scale:
        scale = min_scale
    if max_dim and mode == "square":
        image_max = max(h, w)
        if round(image_max * scale) > max_dim:
            scale = max_dim / image_max
    if scale != 1:
        image = _resize_bilinear(image, (round(h * scale), round(w * scale)))
    h, w = image.shape[:2]
    if mode == "square":
        top, left = (max_dim - h) // 2, (max_dim - w) // 2
        padding = [(top, max_dim - h - top), (left, max_dim - w - left), (0, 0)]
    elif mode == "pad64":
        if min_dim % 64:
            raise ValueError("Minimum dimension must be a multiple of 64")
        mh, mw = (h - h % 64 + 64) if h % 64 else h, (w - w % 64 + 64) if w % 64 else w
        top, left = (mh - h) // 2, (mw - w) // 2
        padding = [(top, mh - h - top), (left, mw - w - left), (0, 0)]
    else:
        raise ValueError("Mode {} not supported".format(mode))
    image = np.pad(image, padding, mode="constant", constant_values=0)
    window = (padding[0][0], padding[1][0], h + padding[0][0], w + padding[1][0])
    return image.astype(dtype), window, scale, padding, None


def mold_image(images, config):
    """model_multi.py:3351-3356."""
    return images.astype(np.float32) - np.asarray(_cfg(config, "MEAN_PIXEL"))


def unmold_mask(mask, bbox, image_shape):
    """utils.py:819-835 (the bilinear resize is the restatement above)."""
    y1, x1, y2, x2 = (int(v) for v in bbox)
    m = _resize_bilinear(mask, (y2 - y1, x2 - x1)) >= 0.5
    full = np.zeros(tuple(image_shape[:2]), dtype=bool)
    full[y1:y2, x1:x2] = m
    return full


# ---------------------------------------------------------------------------------------------------------------------
# parameters, keyed by the reference's Keras layer names
def _glorot(rng, shape, fan_in, fan_out):
    lim = math.sqrt(6.0 / (fan_in + fan_out))                       # keras 'glorot_uniform', the default kernel initializer
    return rng.uniform(-lim, lim, shape).astype(np.float32)


def _conv_p(rng, kh, kw, cin, cout):
    return {"kernel": _glorot(rng, (kh, kw, cin, cout), kh * kw * cin, kh * kw * cout), "bias": np.zeros(cout, np.float32)}


def _bn_p(c):
    return {"bn": (np.ones(c, np.float32), np.zeros(c, np.float32), np.zeros(c, np.float32), np.ones(c, np.float32))}


def resnet_blocks(architecture):
    """(stage, block letter, filters, first-conv stride, has shortcut conv) in graph order (resnet_graph, :572-607)."""
    n4 = {"resnet50": 5, "resnet101": 22}[architecture]
    out = []
    for stage, f, letters in ((2, [64, 64, 256], "abc"), (3, [128, 128, 512], "abcd"),
                              (4, [256, 256, 1024], "a" + "".join(chr(98 + i) for i in range(n4))), (5, [512, 512, 2048], "abc")):
        for i, blk in enumerate(letters):
            out.append((stage, blk, f, (1 if stage == 2 else 2) if i == 0 else 1, i == 0))
    return out


def init_params(config, seed=0):
    """Random parameters of the inference graph with Keras' default initialisers (glorot-uniform kernels, zero biases,
    identity BatchNorm statistics), keyed by the reference's layer names."""
    rng = np.random.default_rng(seed)
    P = {}
    P["conv1"] = _conv_p(rng, 7, 7, 3, 64)
    P["bn_conv1"] = _bn_p(64)
    cin = 64
    for stage, blk, (f1, f2, f3), stride, shortcut in resnet_blocks(_cfg(config, "BACKBONE")):
        cb, bb = "res%d%s_branch" % (stage, blk), "bn%d%s_branch" % (stage, blk)
        P[cb + "2a"], P[bb + "2a"] = _conv_p(rng, 1, 1, cin, f1), _bn_p(f1)
        P[cb + "2b"], P[bb + "2b"] = _conv_p(rng, 3, 3, f1, f2), _bn_p(f2)
        P[cb + "2c"], P[bb + "2c"] = _conv_p(rng, 1, 1, f2, f3), _bn_p(f3)
        if shortcut:
            P[cb + "1"], P[bb + "1"] = _conv_p(rng, 1, 1, cin, f3), _bn_p(f3)
        cin = f3
    D = int(config.TOP_DOWN_PYRAMID_SIZE)
    for name, c in (("fpn_c5p5", 2048), ("fpn_c4p4", 1024), ("fpn_c3p3", 512), ("fpn_c2p2", 256)):
        P[name] = _conv_p(rng, 1, 1, c, D)
    for name in ("fpn_p2", "fpn_p3", "fpn_p4", "fpn_p5"):
        P[name] = _conv_p(rng, 3, 3, D, D)
    A = len(_cfg(config, "RPN_ANCHOR_RATIOS"))
    P["rpn_conv_shared"] = _conv_p(rng, 3, 3, D, 512)
    P["rpn_class_raw"] = _conv_p(rng, 1, 1, 512, 2 * A)
    P["rpn_bbox_pred"] = _conv_p(rng, 1, 1, 512, 4 * A)
    K, fc, ps = int(config.NUM_CLASSES), int(_cfg(config, "FPN_CLASSIF_FC_LAYERS_SIZE")), int(config.POOL_SIZE)
    P["mrcnn_class_conv1"], P["mrcnn_class_bn1"] = _conv_p(rng, ps, ps, D, fc), _bn_p(fc)
    P["mrcnn_class_conv2"], P["mrcnn_class_bn2"] = _conv_p(rng, 1, 1, fc, fc), _bn_p(fc)
    P["mrcnn_class_logits"] = {"kernel": _glorot(rng, (fc, K), fc, K), "bias": np.zeros(K, np.float32)}
    P["mrcnn_bbox_fc"] = {"kernel": _glorot(rng, (fc, 4 * K), fc, 4 * K), "bias": np.zeros(4 * K, np.float32)}
    for i in range(1, 5):
        P["mrcnn_mask_conv%d" % i], P["mrcnn_mask_bn%d" % i] = _conv_p(rng, 3, 3, D if i == 1 else 128, 128), _bn_p(128)
    P["mrcnn_mask_deconv"] = {"kernel": _glorot(rng, (2, 2, 128, 128), 4 * 128, 4 * 128), "bias": np.zeros(128, np.float32)}
    P["mrcnn_mask"] = _conv_p(rng, 1, 1, 128, K)
    P.update(init_neck_params(config, rng))
    return P


def init_neck_params(config, rng):
    """Learnables of the fusion neck for ``config.GRID_REAS`` in the format ``layers.fusion_neck`` / ``weights_io.fusion_params_from_keras``
    use (model_multi.py:394-488): add / mean / max -> BatchNorm; ident -> 1x1x1 conv over the V*C concatenated views; lstm3d ->
    ConvLSTM kernel [3,3,3,C+F,4F]; conv3d -> the four U-Net convolutions + the two depthwise / 1x1 pairs of depth_sampling."""
    D, S, V = int(config.TOP_DOWN_PYRAMID_SIZE), int(config.samples), int(config.NUM_VIEWS)
    mode = getattr(config, "GRID_REAS", "add")
    ident_bn = lambda c: (np.ones(c, np.float32), np.zeros(c, np.float32), np.zeros(c, np.float32), np.ones(c, np.float32))
    out = {}
    for lvl in (2, 3, 4, 5, 6):
        if mode in ("add", "mean", "max"):
            g = {"bn": ident_bn(D)}
        elif mode == "ident":
            g = {"weight": _glorot(rng, (V * D, D), V * D, D), "bias": np.zeros(D, np.float32), "bn": ident_bn(D)}
        elif mode == "lstm3d":
            g = {"W": _glorot(rng, (3, 3, 3, 2 * D, 4 * D), 27 * 2 * D, 27 * 4 * D), "b": np.zeros(4 * D, np.float32), "bn": ident_bn(D)}
        elif mode == "conv3d":
            def c3(cin, cout, transposed=False):                    # Conv3DTranspose kernels are [3,3,3,out,in]
                shape = (3, 3, 3, cout, cin) if transposed else (3, 3, 3, cin, cout)
                return {"W": _glorot(rng, shape, 27 * cin, 27 * cout), "b": np.zeros(cout, np.float32), "bn": ident_bn(cout)}
            g = {"conv1": c3(V * D, 2 * D), "conv2": c3(2 * D, 4 * D), "deconv1": c3(4 * D, 2 * D, True), "deconv2": c3(4 * D, D, True)}
        else:
            raise ValueError("GRID_REAS=%r" % (mode,))
        out["grid_reas_P%d" % lvl] = g
        if mode == "conv3d":
            out["grid_reas_depth_PG%d" % lvl] = {
                "dw1": {"w": np.ones(D * S, np.float32), "b": np.zeros(D * S, np.float32)},
                "conv1": {"W": _glorot(rng, (D * S, 512), D * S, 512), "b": np.zeros(512, np.float32), "bn": ident_bn(512)},
                "dw2": {"w": np.ones(512, np.float32), "b": np.zeros(512, np.float32)},
                "conv2": {"W": _glorot(rng, (512, D), 512, D), "b": np.zeros(D, np.float32), "bn": ident_bn(D)}}
        else:
            out["grid_reas_depth_PG%d" % lvl] = {"weight": _glorot(rng, (S,), S, 1), "bias": 0.0, "bn": (1.0, 0.0, 0.0, 1.0)}
    return out


def randomize(params, seed=1):
    """Non-trivial BatchNorm statistics and biases on top of ``init_params`` (tests / fixtures: with identity statistics and
    zero biases a wiring mistake around a BatchNorm or a bias would go unnoticed)."""
    rng = np.random.default_rng(seed)
    out = {}
    for name, p in params.items():
        p = dict(p)
        if name.startswith("grid_reas_depth") and "weight" in p:
            p["bias"] = float(rng.normal(0, 0.05))
            p["bn"] = (float(rng.uniform(0.8, 1.2)), float(rng.normal(0, 0.05)), float(rng.normal(0, 0.05)), float(rng.uniform(0.7, 1.3)))
        elif name.startswith("grid_reas") and any(isinstance(v, dict) for v in p.values()):
            for k, q in p.items():                                  # conv3d neck: {'conv1': {'W','b','bn'}, ..., 'dw1': {'w','b'}}
                q = dict(q)
                if "bn" in q:
                    c = np.asarray(q["bn"][0]).size
                    q["bn"] = (rng.uniform(0.8, 1.2, c).astype(np.float32), rng.normal(0, 0.05, c).astype(np.float32),
                               rng.normal(0, 0.05, c).astype(np.float32), rng.uniform(0.7, 1.3, c).astype(np.float32))
                if "b" in q:
                    q["b"] = rng.normal(0, 0.05, np.asarray(q["b"]).shape).astype(np.float32)
                if "w" in q:
                    q["w"] = rng.uniform(0.8, 1.2, np.asarray(q["w"]).shape).astype(np.float32)
                p[k] = q
        elif "bn" in p:
            c = np.asarray(p["bn"][0]).size
            p["bn"] = (rng.uniform(0.8, 1.2, c).astype(np.float32), rng.normal(0, 0.05, c).astype(np.float32),
                       rng.normal(0, 0.05, c).astype(np.float32), rng.uniform(0.7, 1.3, c).astype(np.float32))
        elif "kernel" in p:
            p["bias"] = rng.normal(0, 0.05, np.asarray(p["bias"]).shape).astype(np.float32)
        out[name] = p
    return out


def named_weights(params):
    """The dense layers as ``{layer: get_weights() list}`` -- Conv2D / Dense [kernel, bias], BatchNorm [gamma, beta, mean, var]
    (what ``weights_io.write_npz`` stores and ``MaskRCNN.set_named_weights`` takes)."""
    out = {}
    for name, p in params.items():
        if name.startswith("grid_reas"):
            continue
        out[name] = [p["kernel"], p["bias"]] if "kernel" in p else list(p["bn"])
    return out


def checksum(params):
    """Order-independent float64 checksum of a parameter dictionary (fixtures store it to detect a drifting initialiser)."""
    tot = 0.0
    def leaves(v):
        if isinstance(v, dict):
            for k in sorted(v):
                yield from leaves(v[k])
        elif isinstance(v, (tuple, list)):
            for a in v:
                yield from leaves(a)
        else:
            yield v
    for name in sorted(params):
        for key in sorted(params[name]):
            for a in leaves(params[name][key]):
                a = np.asarray(a, dtype=np.float64)
                tot += float(np.sum(a * np.cos(np.arange(a.size, dtype=np.float64).reshape(a.shape))))
    return tot
