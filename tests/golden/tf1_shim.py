"""A small EAGER NumPy stand-in for the ``tensorflow`` / ``keras`` namespaces, just large enough to
import /root/reference/mrcnn/{utils,recurrent,model_multi}.py and to EXECUTE the reference's own
Python for the fusion hot path (unproj_feat, proj_grid, nearest3, apply_box_deltas_graph,
clip_boxes_graph, refine_detections_graph, PyramidROIAlign.call, ProposalLayer.call,
DetectionLayer.call, ConvLSTMCell.call).

TensorFlow 1.x and Keras 2.x are not installable in this image.  What this shim pins is every
decision the REFERENCE made -- operand order, reshapes/transposes, meshgrid indexing, index
layout, gate order, set/top-k plumbing -- because that code runs unmodified.  The TensorFlow
kernels themselves are restated here from their published algorithms, deliberately written
independently of ``oracle/`` (plain loops / different vectorisation) so that agreement between
the two is a real cross-check:
  * fp32 everywhere; matmul contraction ascending k, no FMA;
  * ``gather_nd``: out-of-range index -> zeros (TF GPU kernel);
  * ``range`` / ``linspace``: TF1 RangeOp / LinSpaceOp fill order;
  * ``round``: half to even; ``exp`` / ``log``: correctly rounded fp32;
  * ``image.crop_and_resize`` / ``image.non_max_suppression`` / ``nn.top_k``: TF kernels.
Used only by tests/golden/make_golden.py (build container; never on the GPU box).
"""
import contextlib
import sys
import types

import numpy as np

F32 = np.float32


# ---------------------------------------------------------------------------------------------
class Dim(int):
    @property
    def value(self):
        return int(self)


class TShape(tuple):
    def as_list(self):
        return [int(d) for d in self]

    @property
    def ndims(self):
        return len(self)


class TT(np.ndarray):
    """ndarray that quacks like a tf.Tensor where the reference needs it to."""
    __array_priority__ = 100

    def __array_finalize__(self, obj):
        pass

    @property
    def shape(self):
        return TShape(Dim(d) for d in np.ndarray.shape.__get__(self))

    def get_shape(self):
        return self.shape

    def set_shape(self, shape):
        return None

    def __array_ufunc__(self, ufunc, method, *inputs, **kwargs):
        conv = []
        for x in inputs:
            if isinstance(x, TT):
                x = x.view(np.ndarray)
            elif isinstance(x, np.ndarray) and x.dtype == np.float64:
                x = x.astype(F32)            # convert_to_tensor(dtype hint = float32)
            elif isinstance(x, (np.float64,)):
                x = F32(x)
            conv.append(x)
        if "out" in kwargs:
            kwargs["out"] = tuple(o.view(np.ndarray) if isinstance(o, TT) else o for o in kwargs["out"])
        with np.errstate(all="ignore"):
            res = getattr(ufunc, method)(*conv, **kwargs)
        if isinstance(res, tuple):
            return tuple(_wrap(r) for r in res)
        return _wrap(res)

    def __bool__(self):
        return bool(self.view(np.ndarray))

    def __index__(self):
        return int(self.view(np.ndarray))

    def __hash__(self):
        return id(self)


def _wrap(a):
    if isinstance(a, np.ndarray):
        if a.dtype == np.float64:
            a = a.astype(F32)
        return a.view(TT)
    if isinstance(a, np.float64):
        return np.asarray(F32(a)).view(TT)
    if isinstance(a, np.generic):
        return np.asarray(a).view(TT)
    return a


def T(x, dtype=None):
    """convert_to_tensor: python floats / float64 arrays -> float32, python ints -> int32."""
    if isinstance(x, TT) and dtype is None:
        return x
    if isinstance(x, Variable):
        x = x.value
    a = np.asarray(x)
    if dtype is not None:
        a = a.astype(dtype)
    elif a.dtype == np.float64:
        a = a.astype(F32)
    elif a.dtype == np.int64 and not isinstance(x, np.ndarray):
        a = a.astype(np.int32)
    return np.asarray(a, order="C").view(TT)


def _raw(x):
    return np.asarray(T(x)).view(np.ndarray)


class Variable:
    """tf.Variable for ``repeat_tensor`` (model_multi.py:331-336): ``v[i].assign(n)``."""

    def __init__(self, init, **kw):
        self.value = np.array(_raw(init))

    def __getitem__(self, i):
        var = self

        class _Slot:
            def assign(self_inner, v):
                var.value[i] = v
                return T(var.value)
        return _Slot()


# ---------------------------------------------------------------------------------------------
def _dtype(d):
    if d is None:
        return None
    if isinstance(d, str):
        return np.dtype(d).type
    return d


def _matmul(a, b, name=None, **kw):
    a, b = _raw(a).astype(F32), _raw(b).astype(F32)
    k = a.shape[-1]
    acc = None
    for kk in range(k):                               # ascending k, every product and sum rounded
        term = (a[..., :, kk, None] * b[..., None, kk, :]).astype(F32)
        acc = term if acc is None else (acc + term).astype(F32)
    return T(acc)


def _range(start, limit=None, delta=1, dtype=None, name=None):
    if limit is None:
        start, limit = 0, start
    vals = [np.asarray(_raw(v)) for v in (start, limit, delta)]
    if all(np.issubdtype(v.dtype, np.integer) for v in vals) and dtype is None:
        return T(np.arange(int(vals[0]), int(vals[1]), int(vals[2]), dtype=np.int32))
    s, l, d = (F32(v) for v in vals)
    size = int(np.ceil(np.abs(F32(F32(l - s) / d))))
    out = np.empty(size, F32)
    val = s
    for i in range(size):                             # RangeOp: sequential accumulation
        out[i] = val
        val = F32(val + d)
    return T(out)


def _linspace(start, stop, num, name=None):
    s, e, n = F32(start), F32(stop), int(num)
    out = np.empty(n, F32)
    if n == 1:
        out[0] = s
        return T(out)
    step = F32(F32(e - s) / F32(n - 1))
    for i in range(n):
        out[i] = F32(s + F32(step * F32(i)))
    return T(out)


def _meshgrid(*args, indexing="xy"):
    return [T(m) for m in np.meshgrid(*[_raw(a) for a in args], indexing=indexing)]


def _shape_list(shape):
    if isinstance(shape, (TT, np.ndarray)):
        return [int(s) for s in np.asarray(shape).reshape(-1)]
    if isinstance(shape, (int, np.integer)):
        return [int(shape)]
    return [int(np.asarray(s)) for s in shape]


def _gather_nd(params, indices, name=None):
    p = _raw(params)
    idx = _raw(indices).astype(np.int64)
    k = idx.shape[-1]
    flat = idx.reshape(-1, k)
    out_tail = p.shape[k:]
    out = np.zeros((flat.shape[0],) + out_tail, dtype=p.dtype)
    ok = np.ones(flat.shape[0], bool)
    for a in range(k):
        ok &= (flat[:, a] >= 0) & (flat[:, a] < p.shape[a])
    good = flat[ok]
    out[ok] = p[tuple(good[:, a] for a in range(k))]                    # out-of-range rows stay zero (GPU kernel)
    return T(out.reshape(idx.shape[:-1] + out_tail))


def _gather(params, indices, axis=0, name=None):
    return T(np.take(_raw(params), _raw(indices).astype(np.int64), axis=int(axis)))


def _triangular_solve(matrix, rhs, lower=True, name=None):
    m, r = _raw(matrix).astype(F32), _raw(rhs).astype(F32)
    assert not lower
    n = m.shape[-1]
    m, r = np.broadcast_arrays(m[..., None], r[..., None, :, :]) if False else (m, r)
    x = [None] * n
    for i in range(n - 1, -1, -1):                    # back substitution, subtracting in ascending j
        acc = r[..., i, :]
        for j in range(i + 1, n):
            acc = (acc - (m[..., i, j, None] * x[j]).astype(F32)).astype(F32)
        x[i] = (acc / m[..., i, i, None]).astype(F32)
    return T(np.stack(x, axis=-2))


def _where(cond, x=None, y=None, name=None):
    c = _raw(cond)
    if x is None:
        return T(np.argwhere(c).astype(np.int64))
    return T(np.where(c, _raw(x), _raw(y)))


def _split(value, num_or_size_splits, axis=0, name=None):
    return [T(p) for p in np.split(_raw(value), num_or_size_splits, axis=int(axis))]


def _pad(tensor, paddings, mode="CONSTANT", constant_values=0, name=None):
    pads = [(int(np.asarray(a)), int(np.asarray(b))) for a, b in paddings]
    return T(np.pad(_raw(tensor), pads, mode="constant", constant_values=constant_values))


def _cast(x, dtype, name=None):
    a = _raw(x)
    d = _dtype(dtype)
    with np.errstate(all="ignore"):
        if np.issubdtype(d, np.integer) and np.issubdtype(a.dtype, np.floating):
            bad = ~np.isfinite(a) | (np.abs(a) >= 2.0 ** 31)
            out = np.where(bad, 0, a).astype(d)
            out = np.where(bad, np.iinfo(np.int32).min, out).astype(d)       # x86 cvttss2si result
            return T(out)
        return T(a.astype(d))


class _TopK:
    def __init__(self, values, indices):
        self.values, self.indices = values, indices

    def __getitem__(self, i):
        return (self.values, self.indices)[i]

    def __iter__(self):
        return iter((self.values, self.indices))


def _top_k(input, k=1, sorted=True, name=None):
    a = _raw(input)
    k = int(np.asarray(k))
    order = np.argsort(-a, axis=-1, kind="stable")[..., :k]              # ties: lower index first
    return _TopK(T(np.take_along_axis(a, order, axis=-1)), T(order.astype(np.int32)))


def _nms(boxes, scores, max_output_size, iou_threshold=0.5, name=None, **kw):
    """tf.image.non_max_suppression: greedy, IoU on normalised corners, suppress iff IoU > thr."""
    b = _raw(boxes).astype(F32)
    s = _raw(scores).astype(F32)
    thr = F32(iou_threshold)
    max_out = int(np.asarray(max_output_size))
    order = sorted(range(len(s)), key=lambda i: (-float(s[i]), i))
    sel = []

    def iou(i, j):
        yi0, yi1 = min(b[i, 0], b[i, 2]), max(b[i, 0], b[i, 2])
        xi0, xi1 = min(b[i, 1], b[i, 3]), max(b[i, 1], b[i, 3])
        yj0, yj1 = min(b[j, 0], b[j, 2]), max(b[j, 0], b[j, 2])
        xj0, xj1 = min(b[j, 1], b[j, 3]), max(b[j, 1], b[j, 3])
        ai = F32(F32(yi1 - yi0) * F32(xi1 - xi0))
        aj = F32(F32(yj1 - yj0) * F32(xj1 - xj0))
        if ai <= 0 or aj <= 0:
            return F32(0)
        ih = max(F32(min(yi1, yj1) - max(yi0, yj0)), F32(0))
        iw = max(F32(min(xi1, xj1) - max(xi0, xj0)), F32(0))
        inter = F32(ih * iw)
        return F32(inter / F32(F32(ai + aj) - inter))

    for i in order:
        if len(sel) >= max_out:
            break
        if all(not (iou(i, j) > thr) for j in sel):
            sel.append(i)
    return T(np.array(sel, dtype=np.int32))


def _crop_and_resize(image, boxes, box_ind, crop_size, method="bilinear", extrapolation_value=0, name=None):
    img = _raw(image).astype(F32)
    bx = _raw(boxes).astype(F32)
    bi = _raw(box_ind).astype(np.int64)
    ch, cw = int(crop_size[0]), int(crop_size[1])
    _, H, W, C = img.shape
    out = np.zeros((bx.shape[0], ch, cw, C), F32)
    for n in range(bx.shape[0]):
        y1, x1, y2, x2 = bx[n]
        hs = F32(F32(F32(y2 - y1) * F32(H - 1)) / F32(ch - 1)) if ch > 1 else F32(0)
        ws = F32(F32(F32(x2 - x1) * F32(W - 1)) / F32(cw - 1)) if cw > 1 else F32(0)
        for y in range(ch):
            in_y = F32(F32(y1 * F32(H - 1)) + F32(F32(y) * hs)) if ch > 1 else F32(F32(F32(0.5) * F32(y1 + y2)) * F32(H - 1))
            if not (in_y >= 0 and in_y <= H - 1):
                continue
            top, bot = int(np.floor(in_y)), int(np.ceil(in_y))
            ly = F32(in_y - F32(top))
            for x in range(cw):
                in_x = F32(F32(x1 * F32(W - 1)) + F32(F32(x) * ws)) if cw > 1 else F32(F32(F32(0.5) * F32(x1 + x2)) * F32(W - 1))
                if not (in_x >= 0 and in_x <= W - 1):
                    continue
                left, right = int(np.floor(in_x)), int(np.ceil(in_x))
                lx = F32(in_x - F32(left))
                tl, tr = img[bi[n], top, left], img[bi[n], top, right]
                bl, br = img[bi[n], bot, left], img[bi[n], bot, right]
                t = (tl + ((tr - tl) * lx).astype(F32)).astype(F32)
                bm = (bl + ((br - bl) * lx).astype(F32)).astype(F32)
                out[n, y, x] = (t + ((bm - t) * ly).astype(F32)).astype(F32)
    return T(out)


class _Sparse:
    def __init__(self, dense):
        self.dense = dense


def _set_intersection(a, b, name=None):
    ra, rb = _raw(a), _raw(b)
    rows = [np.intersect1d(ra[i], rb[i]) for i in range(ra.shape[0])]
    return _Sparse(T(np.stack(rows).astype(ra.dtype)))


def _unique(x, name=None):
    a = _raw(x)
    seen, out = set(), []
    for v in a.tolist():
        if v not in seen:
            seen.add(v)
            out.append(v)
    return (T(np.array(out, dtype=a.dtype)), None)


def _convolution(x, w, padding, **kw):
    """tf.nn.convolution, stride 1, SAME, channels-last, any rank; float64 accumulation."""
    x, w = _raw(x).astype(np.float64), _raw(w).astype(np.float64)
    assert padding == "SAME"
    nd = w.ndim - 2
    ks = w.shape[:nd]
    pad = [(0, 0)] + [((k - 1) // 2, k // 2) for k in ks] + [(0, 0)]
    xp = np.pad(x, pad)
    out = np.zeros(x.shape[:-1] + (w.shape[-1],))
    for off in np.ndindex(*ks):
        sl = (slice(None),) + tuple(slice(o, o + n) for o, n in zip(off, x.shape[1:-1])) + (slice(None),)
        out += xp[sl] @ w[off]
    return T(out.astype(F32))


def _sigmoid(x):
    a = _raw(x).astype(np.float64)
    return T((1.0 / (1.0 + np.exp(-a))).astype(F32))


def _unary64(fn):
    def f(x, name=None):
        with np.errstate(all="ignore"):
            return T(fn(_raw(x).astype(np.float64)).astype(F32))
    return f


@contextlib.contextmanager
def _scope(*a, **k):
    yield


class _Auto(types.ModuleType):
    """Module whose unknown attributes resolve to permissive dummy classes (so that
    ``class X(KL.Something)`` and ``from keras.x import y`` succeed at import time)."""

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        full = self.__name__ + "." + name
        if full in sys.modules:
            return sys.modules[full]
        cls = type(name, (object,), {"__init__": lambda self, *a, **k: None,
                                     "__call__": lambda self, *a, **k: None})
        setattr(self, name, cls)
        return cls


def _module(name, auto=True, **attrs):
    m = (_Auto if auto else types.ModuleType)(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def install():
    """Put the fake ``tensorflow`` / ``keras`` / ``skimage`` / ``distutils`` modules in sys.modules."""
    tf = _module("tensorflow", __version__="1.13.1")
    tf.float32, tf.int32, tf.int64, tf.bool = np.float32, np.int32, np.int64, np.bool_
    tf.dtypes = _module("tensorflow.dtypes", float32=np.float32, int32=np.int32)
    tf.newaxis = None
    tf.Variable = Variable
    tf.Tensor = TT

    tf.constant = lambda v, dtype=None, shape=None, name=None: (
        T(np.reshape(np.asarray(v, dtype=_dtype(dtype) or None), shape)) if shape is not None else T(v, _dtype(dtype)))
    tf.convert_to_tensor = lambda v, dtype=None, **k: T(v, _dtype(dtype))
    tf.zeros = lambda shape, dtype=np.float32, name=None: T(np.zeros(_shape_list(shape), _dtype(dtype)))
    tf.ones = lambda shape, dtype=np.float32, name=None: T(np.ones(_shape_list(shape), _dtype(dtype)))
    tf.zeros_like = lambda x, **k: T(np.zeros_like(_raw(x)))
    tf.eye = lambda n, batch_shape=None, dtype=np.float32, **k: T(
        np.broadcast_to(np.eye(n, dtype=_dtype(dtype)), tuple(batch_shape or ()) + (n, n)).copy())
    tf.shape = lambda x, name=None: T(np.array(np.ndarray.shape.__get__(np.asarray(_raw(x))), dtype=np.int32))
    tf.rank = lambda x: _raw(x).ndim
    tf.reshape = lambda x, shape, name=None: T(np.reshape(_raw(x), _shape_list(shape)))
    tf.transpose = lambda x, perm=None, name=None: T(np.transpose(_raw(x), perm))
    tf.matrix_transpose = lambda x, name=None: T(np.swapaxes(_raw(x), -1, -2))
    tf.expand_dims = lambda x, axis, name=None: T(np.expand_dims(_raw(x), int(axis)))
    tf.squeeze = lambda x, axis=None, name=None: T(np.squeeze(_raw(x), axis))
    tf.concat = lambda values, axis, name=None: T(np.concatenate([np.atleast_1d(_raw(v)) for v in values], axis=int(axis)))
    tf.stack = lambda values, axis=0, name=None: T(np.stack([_raw(v) for v in values], axis=int(axis)))
    tf.tile = lambda x, multiples, name=None: T(np.tile(_raw(x), _shape_list(multiples)))
    tf.split = _split
    tf.pad = _pad
    tf.cast = _cast
    tf.to_float = lambda x, name=None: _cast(x, np.float32)
    tf.to_int32 = lambda x, name=None: _cast(x, np.int32)
    tf.floor = lambda x, name=None: T(np.floor(_raw(x)))
    tf.round = lambda x, name=None: T(np.rint(_raw(x)))                  # half to even
    tf.sqrt = lambda x, name=None: T(np.sqrt(_raw(x).astype(F32)))
    tf.exp = _unary64(np.exp)
    tf.log = _unary64(np.log)
    tf.tanh = _unary64(np.tanh)
    tf.sigmoid = _sigmoid
    tf.maximum = lambda a, b, name=None: T(np.maximum(_raw(a), _raw(b)))
    tf.minimum = lambda a, b, name=None: T(np.minimum(_raw(a), _raw(b)))
    tf.equal = lambda a, b, name=None: T(np.equal(_raw(a), _raw(b)))
    tf.logical_or = lambda a, b, name=None: T(np.logical_or(_raw(a), _raw(b)))
    tf.logical_and = lambda a, b, name=None: T(np.logical_and(_raw(a), _raw(b)))
    tf.divide = lambda a, b, name=None: T(a) / T(b)
    tf.argmax = lambda x, axis=None, output_type=np.int64, name=None: T(np.argmax(_raw(x), axis=axis).astype(_dtype(output_type)))
    tf.reduce_sum = lambda x, axis=None, **k: T(np.sum(_raw(x), axis=axis, dtype=F32))
    tf.add_n = lambda xs, name=None: _add_n(xs)
    tf.matmul = _matmul
    tf.linalg = _module("tensorflow.linalg", matmul=_matmul)
    tf.matrix_triangular_solve = _triangular_solve
    tf.range = _range
    tf.linspace = _linspace
    tf.meshgrid = _meshgrid
    tf.gather = _gather
    tf.gather_nd = _gather_nd
    tf.where = _where
    tf.unique = _unique
    tf.boolean_mask = lambda x, m, name=None: T(_raw(x)[_raw(m).astype(bool)])
    tf.stop_gradient = lambda x, name=None: x
    tf.identity = lambda x, name=None: x
    tf.map_fn = lambda fn, elems, dtype=None, **k: T(np.stack([_raw(fn(e)) for e in T(elems)]).astype(_dtype(dtype) or None)) \
        if len(_raw(elems)) else T(np.zeros((0,), _dtype(dtype) or np.float32))
    tf.sparse_tensor_to_dense = lambda sp, **k: sp.dense
    tf.variable_scope = _scope
    tf.name_scope = _scope
    tf.control_dependencies = _scope
    tf.TensorShape = lambda dims: TShape(Dim(d) for d in dims)
    tf.constant_initializer = lambda v: None
    tf.ConfigProto = lambda **k: types.SimpleNamespace(gpu_options=types.SimpleNamespace(), log_device_placement=False)
    tf.Session = lambda **k: None

    tf.sets = _module("tensorflow.sets", set_intersection=_set_intersection)
    tf.image = _module("tensorflow.image", non_max_suppression=_nms, crop_and_resize=_crop_and_resize)
    tf.nn = _module("tensorflow.nn", top_k=_top_k, convolution=_convolution)
    tf.math = _module("tensorflow.math")
    contrib = _module("tensorflow.contrib")
    contrib.slim = _module("tensorflow.contrib.slim")
    contrib.slim.initializers = _module("tensorflow.contrib.slim.initializers", xavier_initializer=lambda **k: None)
    contrib.rnn = _module("tensorflow.contrib.rnn", LSTMStateTuple=lambda c, h: (c, h))
    tf.contrib = contrib
    kb = _module("tensorflow.keras.backend", sum=lambda x, axis=None: tf.reduce_sum(x, axis=axis))
    tf.keras = _module("tensorflow.keras", backend=kb)

    keras = _module("keras", __version__="2.2.4")
    for sub in ("backend", "layers", "engine", "models", "activations", "initializers", "regularizers", "constraints",
                "utils", "legacy"):
        setattr(keras, sub, _module("keras." + sub))
    for sub in ("keras.backend.tensorflow_backend", "keras.layers.recurrent", "keras.engine.base_layer",
                "keras.utils.conv_utils", "keras.utils.generic_utils", "keras.legacy.interfaces", "keras.legacy.layers"):
        _module(sub)
    sys.modules["keras.backend.tensorflow_backend"].set_session = lambda s: None
    sys.modules["keras.legacy.interfaces"].legacy_convlstm2d_support = lambda f: f
    sys.modules["keras.utils"].conv_utils = sys.modules["keras.utils.conv_utils"]
    sys.modules["keras.legacy"].interfaces = sys.modules["keras.legacy.interfaces"]

    for name in ("skimage", "skimage.color", "skimage.io", "skimage.transform", "imgaug", "h5py"):
        _module(name)
    try:
        import distutils.version  # noqa: F401
    except Exception:
        dv = _module("distutils.version", auto=False)

        class LooseVersion:
            def __init__(self, v):
                self.v = tuple(int(p) for p in str(v).split(".") if p.isdigit())

            def __ge__(self, o):
                return self.v >= o.v
        dv.LooseVersion = LooseVersion
        d = _module("distutils", auto=False)
        d.version = dv
    return tf


def _add_n(xs):
    acc = T(xs[0])
    for x in xs[1:]:
        acc = acc + T(x)                              # left to right
    return acc
