#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ by EXECUTING THE REFERENCE'S OWN PYTHON.

Run in the build container only (needs /root/reference; the GPU box never runs this):

    python tests/golden/make_golden.py

mrcnn/utils.py, mrcnn/recurrent.py and mrcnn/model_multi.py are imported unmodified from
/root/reference over the eager NumPy ``tf``/``keras`` stand-in in tf1_shim.py, then the hot-path
functions are called on small seeded inputs and inputs + outputs are stored as .npz files.
tests/test_golden.py checks the oracle (CPU) and tests/test_gpu_golden.py the CUDA kernels against
these files.  Fixtures are small (a few hundred KB in total).
"""
import io
import os
import sys
import contextlib
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

import tf1_shim  # noqa: E402

tf = tf1_shim.install()
sys.path.insert(0, REF)
with contextlib.redirect_stdout(io.StringIO()):
    from mrcnn import utils as ref_utils          # noqa: E402
    from mrcnn import model_multi as mm           # noqa: E402

from mulit_view_object_detection_b200.config import FusionConfig   # noqa: E402
from mulit_view_object_detection_b200 import synthetic as syn      # noqa: E402

T = tf1_shim.T


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def arr(x):
    return np.array(np.asarray(x).view(np.ndarray))


def save(name, **kw):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **kw)
    print("wrote %-28s %6.1f KB" % (name + ".npz", os.path.getsize(path) / 1024.0))


def cfg_dict(cfg, keys):
    return {"cfg_" + k: np.asarray(getattr(cfg, k)) for k in keys}


GRID_KEYS = ["nvox", "nvox_z", "vmin", "vmax", "vsize", "vmin_z", "vmax_z", "vsize_z", "samples", "IMAGE_SHAPE",
             "IMAGES_PER_GPU", "NUM_VIEWS"]


def gen_unproject_project():
    cases = [
        ("fusion_a", dict(nvox=6, nvox_z=5, samples=4, NUM_VIEWS=3, IMAGES_PER_GPU=2), 12, 12, 4, 8),
        # voxel sizes that are NOT exactly representable: exercises the sequential tf.range fill
        ("fusion_b", dict(nvox=7, nvox_z=9, vmin=-2.5, vmax=2.5, vmin_z=1.0, vmax_z=10.0, samples=5, NUM_VIEWS=2,
                          IMAGES_PER_GPU=1), 10, 14, 8, 10),
    ]
    for name, kw, fh, fw, C, P in cases:
        cfg = FusionConfig(**kw)
        B, V = cfg.BATCH_SIZE, cfg.NUM_VIEWS
        feats, Rcam, Kmat = syn.make_scene(cfg, B, V, fh, fw, C, seed=zlib.crc32(name.encode()) % 1000)
        per_view = quiet(mm.unproj_feat, [T(feats), T(Rcam), T(Kmat)], cfg)
        summed = tf.keras.backend.sum(per_view, axis=1)                       # grid_reas 'add', model_multi.py:402
        rays = quiet(mm.proj_grid, [summed, T(Rcam), T(Kmat)], cfg, P)
        save(name, feats=feats, Rcam=Rcam, Kmat=Kmat, proj_size=np.int32(P), per_view=arr(per_view), summed=arr(summed),
             rays=arr(rays), **cfg_dict(cfg, GRID_KEYS))


def gen_boxes():
    rng = np.random.default_rng(11)
    boxes = syn.make_rois(rng, 1, 64, pad_frac=0.1)[0]
    deltas = rng.normal(0, 0.5, (64, 4)).astype(np.float32)
    window = np.array([0.1, 0.05, 0.9, 1.0], np.float32)
    applied = mm.apply_box_deltas_graph(T(boxes), T(deltas))
    clipped = mm.clip_boxes_graph(applied, T(window))
    # the reference's NumPy twins (mrcnn/utils.py) on the same inputs, for cross-reading
    np_applied = ref_utils.apply_box_deltas(boxes.astype(np.float64), deltas.astype(np.float64))
    save("boxes", boxes=boxes, deltas=deltas, window=window, applied=arr(applied), clipped=arr(clipped),
         utils_applied_f64=np_applied)


def gen_nms_utils():
    """mrcnn/utils.py:381-415 non_max_suppression / :319-337 compute_iou -- pure NumPy, run unmodified."""
    rng = np.random.default_rng(12)
    n = 300
    c = rng.uniform(0.2, 0.8, (n, 2))
    s = rng.uniform(0.05, 0.3, (n, 2))
    boxes = np.concatenate([c - s / 2, c + s / 2], axis=1).astype(np.float32)
    scores = (rng.permutation(n).astype(np.float32) + 1) / n
    out = {}
    for thr in (0.3, 0.5, 0.7):
        out["keep_%02d" % int(thr * 10)] = ref_utils.non_max_suppression(boxes, scores, thr)
    area = (boxes[:, 2] - boxes[:, 0]) * (boxes[:, 3] - boxes[:, 1])
    iou0 = ref_utils.compute_iou(boxes[0], boxes, area[0], area)
    save("nms_utils", boxes=boxes, scores=scores, iou_row0=iou0, **out)


def gen_refine():
    rng = np.random.default_rng(13)
    N, K = 120, 7
    cfg = FusionConfig(NUM_CLASSES=K, DETECTION_MIN_CONFIDENCE=0.3, DETECTION_MAX_INSTANCES=20)
    rois = syn.make_rois(rng, 1, N)[0]
    probs, deltas = syn.make_detection_inputs(rng, N, K)
    deltas *= 0.3
    window = np.array([0.0, 0.02, 1.0, 0.97], np.float32)
    det = quiet(mm.refine_detections_graph, T(rois), T(probs), T(deltas), T(window), cfg)
    cfg0 = FusionConfig(NUM_CLASSES=K, DETECTION_MIN_CONFIDENCE=0, DETECTION_MAX_INSTANCES=20)
    det0 = quiet(mm.refine_detections_graph, T(rois), T(probs), T(deltas), T(window), cfg0)
    save("refine", rois=rois, probs=probs, deltas=deltas, window=window, det=arr(det), det_noconf=arr(det0))
    # batched through DetectionLayer
    B = 2
    cfgb = FusionConfig(NUM_CLASSES=K, DETECTION_MIN_CONFIDENCE=0.3, DETECTION_MAX_INSTANCES=20, IMAGES_PER_GPU=B,
                        IMAGE_SHAPE=np.array([96, 128, 3]))
    rois_b = syn.make_rois(rng, B, N)
    pd = [syn.make_detection_inputs(rng, N, K) for _ in range(B)]
    probs_b = np.stack([p for p, _ in pd])
    deltas_b = np.stack([d for _, d in pd]) * 0.3
    meta = syn.make_image_meta(B, (96, 128, 3), K, window=(8, 0, 88, 128))
    layer = mm.DetectionLayer(cfgb)
    out = quiet(layer.call, [T(rois_b), T(probs_b), T(deltas_b), T(meta)])
    save("detection_layer", rois=rois_b, probs=probs_b, deltas=deltas_b, image_meta=meta, out=arr(out))


def gen_roi_align():
    rng = np.random.default_rng(14)
    B, R, C = 2, 40, 8
    boxes = syn.make_rois(rng, B, R)
    boxes[0, 0] = [0.0, 0.0, 1.0, 1.0]
    boxes[0, 1] = [0.85, 0.9, 1.2, 1.3]
    maps = [np.maximum(rng.standard_normal((B, s, s, C)), 0).astype(np.float32) for s in (32, 16, 8, 4)]
    meta = syn.make_image_meta(B, (512, 512, 3), 3)
    for pool in ((7, 7), (3, 5)):
        layer = mm.PyramidROIAlign(pool)
        out = quiet(layer.call, [T(boxes), T(meta)] + [T(m) for m in maps])
        save("roi_align_%dx%d" % pool, boxes=boxes, image_meta=meta, P2=maps[0], P3=maps[1], P4=maps[2], P5=maps[3],
             pool=np.asarray(pool), out=arr(out))


def gen_proposals():
    rng = np.random.default_rng(15)
    B = 2
    anchors = syn.make_anchors((64, 64), scales=(8, 16, 32), strides=(4, 8, 16))
    A = anchors.shape[0]
    cfg = FusionConfig(PRE_NMS_LIMIT=150, IMAGES_PER_GPU=B)
    fg = (rng.permutation(B * A).reshape(B, A).astype(np.float32) + 1) / (B * A + 1)
    probs = np.stack([1 - fg, fg], axis=-1).astype(np.float32)
    bbox = rng.normal(0, 0.5, (B, A, 4)).astype(np.float32)
    anc = np.broadcast_to(anchors, (B, A, 4)).copy()
    layer = mm.ProposalLayer(proposal_count=40, nms_threshold=0.7, config=cfg)
    out = quiet(layer.call, [T(probs), T(bbox), T(anc)])
    save("proposals", probs=probs, bbox=bbox, anchors=anc, out=arr(out), pre_nms_limit=np.int32(150),
         proposal_count=np.int32(40), nms_threshold=np.float32(0.7))


def gen_convlstm():
    rng = np.random.default_rng(16)
    B, X, Y, Z, C = 1, 4, 3, 5, 4
    F = C
    cell = mm.ConvLSTMCell(shape=[X, Y, Z], kernel=[3, 3, 3], filters=F)
    W = (rng.standard_normal((3, 3, 3, C + F, 4 * F)) * 0.15).astype(np.float32)
    b = rng.normal(0, 0.1, 4 * F).astype(np.float32)
    cell.W, cell.bias = T(W), T(b)
    xs = rng.standard_normal((B, 3, X, Y, Z, C)).astype(np.float32)
    c = np.zeros((B, X, Y, Z, F), np.float32)
    h = np.zeros((B, X, Y, Z, F), np.float32)
    hs, cs = [], []
    for t in range(3):
        out, (c, h) = quiet(cell.call, T(xs[:, t]), (T(c), T(h)))
        hs.append(arr(h))
        cs.append(arr(c))
    save("convlstm", x=xs, W=W, b=b, h=np.stack(hs, 1), c=np.stack(cs, 1))


def gen_convlstm_sequence():
    """ConvRNN3D.call + get_initial_state (mrcnn/recurrent.py:143-173, 230-300) over the view axis, as convlstm() drives it
    (model_multi.py:109-123): zero initial states shaped like the INPUT, cell applied per view, last output returned.  The
    Keras RNN base class is replaced by a plain object carrying the attributes `call` reads; K.rnn is a Python loop."""
    from mrcnn import recurrent as rec
    rng = np.random.default_rng(20)
    B, V, X, Y, Z, C = 1, 4, 3, 4, 5, 4
    F = C
    cell = mm.ConvLSTMCell(shape=[X, Y, Z], kernel=[3, 3, 3], filters=F)
    W = (rng.standard_normal((3, 3, 3, C + F, 4 * F)) * 0.15).astype(np.float32)
    b = rng.normal(0, 0.1, 4 * F).astype(np.float32)
    cell.W, cell.bias = T(W), T(b)
    cell.kernel_shape = [3, 3, 3, C, 4 * F]            # what ConvLSTMCell.build sets (recurrent.py:437-441)

    def k_rnn(step, inputs, initial_states, constants=None, go_backwards=False, mask=None, input_length=None):
        states, outs = list(initial_states), []
        for t in range(arr(inputs).shape[1]):
            out, states = step(T(arr(inputs)[:, t]), tuple(states))
            states = list(states)
            outs.append(arr(out))
        return T(outs[-1]), T(np.stack(outs, 1)), states

    rec.K.zeros_like = lambda x: T(np.zeros_like(arr(x)))
    rec.K.sum = lambda x, axis=None: T(np.sum(arr(x), axis=axis, dtype=np.float32))
    rec.K.int_shape = lambda x: tuple(int(d) for d in arr(x).shape)
    rec.K.image_data_format = lambda: "channels_last"
    rec.K.rnn = k_rnn
    rec.has_arg = lambda fn, name: False
    rec.to_list = lambda x, allow_tuple=False: list(x)

    class Stub:
        pass
    layer = Stub()
    layer.cell, layer.stateful, layer.states = cell, False, [None, None]
    layer.go_backwards, layer._num_constants, layer.return_sequences, layer.return_state = False, None, False, False
    layer.get_initial_state = lambda inputs: rec.ConvRNN3D.get_initial_state(layer, inputs)
    xs = rng.standard_normal((B, V, X, Y, Z, C)).astype(np.float32)
    out = quiet(rec.ConvRNN3D.call, layer, T(xs))
    save("convlstm_sequence", x=xs, W=W, b=b, out=arr(out))


def gen_poses():
    """mrcnn/utils.py:1175-1218 quat2rot / vec2rot -- pure NumPy, run unmodified."""
    rng = np.random.default_rng(17)
    q = rng.normal(0, 1, (5, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)      # quat2rot is a rotation only for unit quaternions
    vec = rng.normal(0, 2, (5, 9))
    save("poses", quat=q, quat_R=np.stack([ref_utils.quat2rot(list(x)) for x in q]),
         vec=vec, vec_R=np.stack([ref_utils.vec2rot(x) for x in vec]))


# ---- the reference's grid_reas / depth_sampling wiring (mrcnn/model_multi.py:394-488) -----------------------------------
# These functions are Keras graph builders.  They are EXECUTED here with small functional stand-ins for the Keras layers
# they instantiate (weights looked up by the layer NAME the reference gives them), so the fixtures pin everything the
# reference authored: transposes / reshapes (view-major channel order, the c*S+s depth order), ReLU placement, layer
# sequence, the skip-concat order and the layer names.  The arithmetic inside a stand-in layer is the oracle's
# restatement of the Keras / TF kernel (third-party: unpinned, see oracle/__init__.py).
import oracle  # noqa: E402

LAYER_WEIGHTS = {}          # layer name -> get_weights() list, filled by the generators below


class _Lambda:
    def __init__(self, fn, name=None, **k):
        self.fn = fn

    def __call__(self, x):
        return self.fn(x)


class _Activation:
    def __init__(self, kind, **k):
        assert kind == "relu"

    def __call__(self, x):
        return T(np.maximum(arr(x), np.float32(0)))


class _Conv3D:
    def __init__(self, filters, kernel_size, strides=(1, 1, 1), padding="valid", name=None, **k):
        assert padding == "same"
        self.name, self.stride, self.filters = name, int(strides[0]), filters

    def __call__(self, x):
        W, b = LAYER_WEIGHTS[self.name]
        assert W.shape[-1] == self.filters
        return T((oracle.conv3d_strided_same(arr(x), W, self.stride) + b.astype(np.float64)).astype(np.float32))


class _Conv3DTranspose(_Conv3D):
    def __call__(self, x):
        W, b = LAYER_WEIGHTS[self.name]
        assert W.shape[-2] == self.filters
        return T((oracle.conv3d_transpose_same(arr(x), W, self.stride) + b.astype(np.float64)).astype(np.float32))


class _TimeDistributed:
    """KL.TimeDistributed(BatchNorm(), name) (add_bn_layer, :501-502) and KL.TimeDistributed(KL.Conv2D(1,(1,1)), name) (:483)."""

    def __init__(self, layer, name=None, **k):
        self.layer, self.name = layer, name

    def __call__(self, x, training=None):
        if isinstance(self.layer, _Conv2D):
            return self.layer(x, name=self.name)
        scale, shift = oracle.batch_norm_affine(*LAYER_WEIGHTS[self.name])
        return T(arr(x) * scale + shift)


class _Conv2D:
    def __init__(self, filters, kernel_size, padding="valid", name=None, **k):
        assert tuple(kernel_size) == (1, 1)
        self.name, self.filters = name, filters

    def __call__(self, x, name=None):
        W, b = LAYER_WEIGHTS[name or self.name]
        W = np.asarray(W).reshape(W.shape[-2], W.shape[-1])
        return T((arr(x).astype(np.float64) @ W.astype(np.float64)).astype(np.float32) + b)


class _DepthwiseConv2D:
    def __init__(self, kernel_size, depth_multiplier=1, name=None, **k):
        assert tuple(kernel_size) == (1, 1) and depth_multiplier == 1
        self.name = name

    def __call__(self, x):
        w, b = LAYER_WEIGHTS[self.name]
        return T(arr(x) * np.asarray(w).reshape(-1) + b)


def _install_layers():
    KL = mm.KL
    KL.Lambda, KL.Activation, KL.Conv3D, KL.Conv3DTranspose = _Lambda, _Activation, _Conv3D, _Conv3DTranspose
    KL.TimeDistributed, KL.Conv2D, KL.DepthwiseConv2D = _TimeDistributed, _Conv2D, _DepthwiseConv2D
    tf.keras.backend.sum = lambda x, axis=None: T(np.sum(arr(x), axis=axis, dtype=np.float32))


def _rand_bn(rng, n):
    return [rng.uniform(0.8, 1.2, n).astype(np.float32), rng.normal(0, 0.05, n).astype(np.float32),
            rng.normal(0, 0.05, n).astype(np.float32), rng.uniform(0.7, 1.3, n).astype(np.float32)]


def _flat(named):
    return {"w__%s__%d" % (k, i): np.asarray(a) for k, ws in named.items() for i, a in enumerate(ws)}


def gen_grid_reas():
    _install_layers()
    rng = np.random.default_rng(18)
    B, V, X, Z, C, F = 1, 3, 4, 8, 4, 4
    grids = rng.standard_normal((B, V, X, X, Z, C)).astype(np.float32)
    scope = "grid_reas_P4"
    for mode in ("add", "ident", "conv3d", "conv3d_tc"):
        if mode == "conv3d_tc":        # channel counts the tensor-core path accepts (32-channel K chunks), still a small file
            V, C, F, mode = 1, 32, 16, "conv3d"
            grids = rng.standard_normal((B, V, X, X, Z, C)).astype(np.float32)
            tag = "conv3d_tc"
        else:
            tag = mode
        cfg = FusionConfig(GRID_REAS=mode, NUM_VIEWS=V, nvox=X, nvox_z=Z, TOP_DOWN_PYRAMID_SIZE=F)
        cfg.TRAIN_BN = False
        LAYER_WEIGHTS.clear()
        mm.reused_lay.clear()
        if mode == "add":
            LAYER_WEIGHTS[scope + "_batch_norm"] = _rand_bn(rng, C)
        elif mode == "ident":
            LAYER_WEIGHTS[scope + "ident_conv"] = [(rng.standard_normal((1, 1, 1, V * C, F)) * 0.3).astype(np.float32),
                                                   rng.normal(0, 0.1, F).astype(np.float32)]
            LAYER_WEIGHTS[scope + "_batch_norm"] = _rand_bn(rng, F)
        else:
            nc, nb = scope + "_3D_conv", scope + "_batch_norm"
            for suf, bsuf, shp, nout in (("_1", "_1", (3, 3, 3, V * C, 2 * F), 2 * F), ("_2", "_2", (3, 3, 3, 2 * F, 4 * F), 4 * F),
                                         ("_deconv_1", "deconv_1", (3, 3, 3, 2 * F, 4 * F), 2 * F),
                                         ("_deconv_2", "deconv_2", (3, 3, 3, F, 4 * F), F)):
                LAYER_WEIGHTS[nc + suf] = [(rng.standard_normal(shp) * 0.1).astype(np.float32), rng.normal(0, 0.1, nout).astype(np.float32)]
                LAYER_WEIGHTS[nb + bsuf] = _rand_bn(rng, nout)
        out = quiet(mm.grid_reas, T(grids), scope, cfg)
        save("grid_reas_" + tag, grids=grids, out=arr(out), V=np.int32(V), F=np.int32(F), **_flat(LAYER_WEIGHTS))


def gen_depth_sampling():
    _install_layers()
    rng = np.random.default_rng(19)
    B, S, P, C, F = 2, 5, 3, 4, 4
    x = np.maximum(rng.standard_normal((B, S, P, P, C)), 0).astype(np.float32)
    name = "grid_reas_depth_PG4"
    for mode in ("add", "conv3d"):
        cfg = FusionConfig(GRID_REAS=mode, samples=S, TOP_DOWN_PYRAMID_SIZE=F)
        cfg.TRAIN_BN = False
        LAYER_WEIGHTS.clear()
        if mode == "conv3d":
            for i, (cin, cout) in enumerate(((C * S, 512), (512, F)), 1):
                LAYER_WEIGHTS[name + "_DepthwiseConv_%d" % i] = [rng.uniform(0.5, 1.5, (1, 1, cin, 1)).astype(np.float32),
                                                                 rng.normal(0, 0.1, cin).astype(np.float32)]
                LAYER_WEIGHTS[name + "2DConv_%d" % i] = [(rng.standard_normal((1, 1, cin, cout)) * cin ** -0.5).astype(np.float32),
                                                          rng.normal(0, 0.1, cout).astype(np.float32)]
                LAYER_WEIGHTS[name + "bn_%d" % i] = _rand_bn(rng, cout)
        else:
            LAYER_WEIGHTS[name + "2DConv"] = [rng.normal(0.1, 0.3, (1, 1, S, 1)).astype(np.float32), np.array([0.05], np.float32)]
            LAYER_WEIGHTS[name + "bn_deconv"] = _rand_bn(rng, 1)
        out = quiet(mm.depth_sampling, T(x), cfg, name)
        save("depth_sampling_" + mode, x=x, out=arr(out), S=np.int32(S), F=np.int32(F), **_flat(LAYER_WEIGHTS))


# ---- the dense graphs of the full model (mrcnn/model_multi.py:497-641 backbone + FPN, :1265-1306 RPN, :1335-1444 heads) --
# SURVEY.md section 8(f) rank 4.  Again the reference's OWN graph builders are executed; the Keras layers they instantiate
# are functional stand-ins evaluated in float64 on the CPU (torch), weights looked up by the reference's layer names.  That
# pins the wiring the product's model.py has to reproduce: strides and padding placement, block order, shortcut placement,
# the top-down additions, P6, the (h, w, anchor) order of the RPN reshapes, the FC-as-convolution flattening order, the
# transposed-convolution kernel layout.  The weights are model_host.init_params + randomize (seeded); the fixture stores
# their checksum, the inputs and the outputs.
def _t64(x):
    import torch
    return torch.from_numpy(np.ascontiguousarray(arr(x), dtype=np.float64))


def _same_pads(n, k, s):
    out = -(-n // s)
    total = max((out - 1) * s + k - n, 0)
    return total // 2, total - total // 2


def _act(y, kind):
    if kind in (None, "linear"):
        return y
    if kind == "relu":
        return np.maximum(y, 0)
    if kind == "sigmoid":
        return 1.0 / (1.0 + np.exp(-y))
    if kind == "softmax":
        e = np.exp(y - y.max(axis=-1, keepdims=True))
        return e / e.sum(axis=-1, keepdims=True)
    raise AssertionError(kind)


class _MConv2D:
    def __init__(self, filters, kernel_size, strides=(1, 1), padding="valid", activation=None, use_bias=True, name=None, **k):
        self.filters, self.name, self.padding, self.activation = filters, name, padding.lower(), activation
        self.strides = (strides, strides) if isinstance(strides, int) else tuple(strides)
        self.kernel_size = tuple(kernel_size)

    def __call__(self, x, name=None):
        import torch.nn.functional as F_
        W, b = LAYER_WEIGHTS[name or self.name]
        assert W.shape[:2] == self.kernel_size and W.shape[-1] == self.filters, (name or self.name, W.shape)
        t = _t64(x).permute(0, 3, 1, 2)
        if self.padding == "same":
            (pt, pb), (pl, pr) = _same_pads(t.shape[2], W.shape[0], self.strides[0]), _same_pads(t.shape[3], W.shape[1], self.strides[1])
            t = F_.pad(t, (pl, pr, pt, pb))
        y = F_.conv2d(t, _t64(W).permute(3, 2, 0, 1), _t64(b), stride=self.strides)
        return T(_act(y.permute(0, 2, 3, 1).numpy(), self.activation).astype(np.float32))


class _MConv2DTranspose(_MConv2D):
    def __call__(self, x, name=None):
        import torch.nn.functional as F_
        W, b = LAYER_WEIGHTS[name or self.name]                      # keras: [kh, kw, out, in]
        assert self.padding == "valid" and W.shape[2] == self.filters
        y = F_.conv_transpose2d(_t64(x).permute(0, 3, 1, 2), _t64(W).permute(3, 2, 0, 1), _t64(b), stride=self.strides)
        return T(_act(y.permute(0, 2, 3, 1).numpy(), self.activation).astype(np.float32))


class _MDense:
    def __init__(self, units, activation=None, name=None, **k):
        self.units, self.activation, self.name = units, activation, name

    def __call__(self, x, name=None):
        W, b = LAYER_WEIGHTS[name or self.name]
        assert W.shape[-1] == self.units
        return T(_act(arr(x).astype(np.float64) @ W.astype(np.float64) + b.astype(np.float64), self.activation).astype(np.float32))


class _MActivation:
    def __init__(self, kind, name=None, **k):
        self.kind = kind

    def __call__(self, x, name=None):
        return T(_act(arr(x).astype(np.float64), self.kind).astype(np.float32))


class _MZeroPadding2D:
    def __init__(self, padding, **k):
        self.p = padding

    def __call__(self, x, name=None):
        (a, b) = self.p
        return T(np.pad(arr(x), ((0, 0), (a, a), (b, b), (0, 0))))


class _MMaxPool2D:
    def __init__(self, pool_size=(2, 2), strides=None, padding="valid", **k):
        self.k = tuple(pool_size)
        self.s = (strides, strides) if isinstance(strides, int) else tuple(strides)
        self.padding = padding.lower()

    def __call__(self, x, name=None):
        import torch.nn.functional as F_
        t = _t64(x).permute(0, 3, 1, 2)
        if self.padding == "same":
            (pt, pb), (pl, pr) = _same_pads(t.shape[2], self.k[0], self.s[0]), _same_pads(t.shape[3], self.k[1], self.s[1])
            t = F_.pad(t, (pl, pr, pt, pb), value=float("-inf"))
        return T(F_.max_pool2d(t, self.k, self.s).permute(0, 2, 3, 1).numpy().astype(np.float32))


class _MUpSampling2D:
    def __init__(self, size=(2, 2), **k):
        self.size = tuple(size)

    def __call__(self, x, name=None):
        return T(np.repeat(np.repeat(arr(x), self.size[0], axis=1), self.size[1], axis=2))


class _MAdd:
    def __init__(self, **k):
        pass

    def __call__(self, xs):
        return T(arr(xs[0]) + arr(xs[1]))


class _MReshape:
    def __init__(self, shape, name=None, **k):
        self.shape = tuple(int(v) for v in shape)

    def __call__(self, x):
        return T(arr(x).reshape((arr(x).shape[0],) + self.shape))


class _MTimeDistributed:
    def __init__(self, layer, name=None, **k):
        self.layer, self.name = layer, name

    def __call__(self, x, training=None):
        a = arr(x)
        if isinstance(self.layer, mm.BatchNorm):
            assert training is False or training is None
            scale, shift = oracle.batch_norm_affine(*LAYER_WEIGHTS[self.name])
            return T(a * scale + shift)
        flat = T(a.reshape((a.shape[0] * a.shape[1],) + a.shape[2:]))
        y = arr(self.layer(flat, name=self.name))
        return T(y.reshape(a.shape[:2] + y.shape[1:]))


def _install_model_layers():
    KL = mm.KL
    KL.Lambda, KL.Activation, KL.TimeDistributed, KL.Conv2D, KL.Conv2DTranspose = _Lambda, _MActivation, _MTimeDistributed, _MConv2D, _MConv2DTranspose
    KL.Dense, KL.ZeroPadding2D, KL.MaxPool2D, KL.UpSampling2D, KL.Add, KL.Reshape = _MDense, _MZeroPadding2D, _MMaxPool2D, _MUpSampling2D, _MAdd, _MReshape
    mm.PyramidROIAlign.__call__ = lambda self, inputs: self.call(inputs)       # keras.engine.Layer.__call__ -> call (in memory only)
    mm.K.squeeze = lambda x, axis: T(np.squeeze(arr(x), axis))
    mm.K.int_shape = lambda x: tuple(int(v) for v in arr(x).shape)


def gen_model_graphs():
    from mulit_view_object_detection_b200 import model_host as MH
    _install_model_layers()
    cfg = FusionConfig(IMAGE_SHAPE=np.array([64, 64, 3]), NUM_VIEWS=2, IMAGES_PER_GPU=1, TOP_DOWN_PYRAMID_SIZE=16, NUM_CLASSES=5,
                       BACKBONE="resnet50", POOL_SIZE=3, MASK_POOL_SIZE=4, FPN_CLASSIF_FC_LAYERS_SIZE=32)
    cfg.TRAIN_BN = False
    params = MH.randomize(MH.init_params(cfg, seed=5), seed=6)
    LAYER_WEIGHTS.clear()
    LAYER_WEIGHTS.update(MH.named_weights(params))
    rng = np.random.default_rng(7)
    images = rng.normal(0, 40, (1, 2, 64, 64, 3)).astype(np.float32)
    P = quiet(mm.build_resnet_fpn, T(images), cfg)
    P = [arr(p) for p in P]
    rpn = [arr(a) for a in quiet(mm.rpn_graph, T(P[1][:, 0]), 3, 1)]
    maps = [rng.standard_normal((1, 64 // s, 64 // s, 16)).astype(np.float32) for s in (4, 8, 16, 32)]
    rois = syn.make_rois(rng, 1, 12, pad_frac=0.1)
    meta = MH_meta(cfg)
    cls = [arr(a) for a in quiet(mm.fpn_classifier_graph, T(rois), [T(m) for m in maps], T(meta), cfg.POOL_SIZE, cfg.NUM_CLASSES,
                                 train_bn=False, fc_layers_size=cfg.FPN_CLASSIF_FC_LAYERS_SIZE)]
    mask = arr(quiet(mm.build_fpn_mask_graph, T(rois[:, :6]), [T(m) for m in maps], T(meta), cfg.MASK_POOL_SIZE, cfg.NUM_CLASSES,
                     train_bn=False))
    # the reference's pure-NumPy host helpers, run unmodified (mrcnn/utils.py:842-900, :1112-1143; model_multi.py:89-103)
    cfg.BACKBONE_STRIDES, cfg.COMPUTE_BACKBONE_SHAPE = [4, 8, 16, 32, 64], None
    shapes = mm.compute_backbone_shapes(cfg, (128, 192, 3))
    anchors = ref_utils.generate_pyramid_anchors((32, 64, 128, 256, 512), [0.5, 1, 2], shapes, [4, 8, 16, 32, 64], 1)
    save("model_graphs", images=images, **{"P%d" % (i + 2): p for i, p in enumerate(P)},
         rpn_logits=rpn[0], rpn_probs=rpn[1], rpn_bbox=rpn[2],
         **{"map%d" % i: m for i, m in enumerate(maps)}, rois=rois, meta=meta,
         cls_logits=cls[0], cls_probs=cls[1], cls_bbox=cls[2], mask=mask,
         weights_checksum=np.float64(MH.checksum(params)),
         backbone_shapes=shapes, anchors=anchors.astype(np.float64), anchors_norm=ref_utils.norm_boxes(anchors, (128, 192)),
         denorm=ref_utils.denorm_boxes(ref_utils.norm_boxes(anchors[:64], (128, 192)), (128, 192)))


def MH_meta(cfg):
    h, w = int(cfg.IMAGE_SHAPE[0]), int(cfg.IMAGE_SHAPE[1])
    return mm.compose_image_meta(0, (h, w, 3), (h, w, 3), (0, 0, h, w), 1.0, np.zeros([cfg.NUM_CLASSES], dtype=np.int32))[None].astype(np.float32)


if __name__ == "__main__":
    if len(sys.argv) > 1:                      # python make_golden.py gen_model_graphs ...: only the named generators
        for g in sys.argv[1:]:
            globals()[g]()
        sys.exit(0)
    gen_grid_reas()
    gen_depth_sampling()
    gen_unproject_project()
    gen_boxes()
    gen_nms_utils()
    gen_refine()
    gen_roi_align()
    gen_proposals()
    gen_convlstm()
    gen_convlstm_sequence()
    gen_poses()
    gen_model_graphs()          # last: it swaps in its own stand-ins for the Keras layers
