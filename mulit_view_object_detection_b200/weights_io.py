"""Host-side data formats either side of the fusion path: Keras layer names -> the parameter dictionaries the layers of
this package take, and the ``image_meta`` vector.

The reference stores its learnables in a Keras HDF5 checkpoint and restores them by layer name
(``MaskRCNN.load_weights``, mrcnn/model_multi.py:2592-2637).  The fusion path owns these layers
(names as built in mrcnn/model_multi.py:394-488, scopes from :2388-2404):

    grid_reas_P<l>_batch_norm                      add / ident / lstm3d   TimeDistributed(BatchNorm)  [gamma, beta, mean, var]
    grid_reas_P<l>ident_conv                       ident                  Conv3D 1x1x1                [kernel [1,1,1,V*C,F], bias]
    grid_reas_P<l>_convlstm3d                      lstm3d                 ConvRNN3D(ConvLSTMCell)     [kernel [3,3,3,C+F,4F], bias]
    grid_reas_P<l>_3D_conv_{1,2}                   conv3d                 Conv3D s2                   [kernel, bias]
    grid_reas_P<l>_3D_conv_deconv_{1,2}            conv3d                 Conv3DTranspose s2          [kernel [3,3,3,out,in], bias]
    grid_reas_P<l>_batch_norm_{1,2}, ..._batch_normdeconv_{1,2}   conv3d  BatchNorm (sic: no underscore before 'deconv', :434,:440)
    grid_reas_depth_PG<l>2DConv, ...bn_deconv      depth_sampling         TimeDistributed(Conv2D 1x1) [kernel [1,1,S,1], bias [1]], scalar BN
    grid_reas_depth_PG<l>_DepthwiseConv_{1,2}, ...2DConv_{1,2}, ...bn_{1,2}   depth_sampling, conv3d branch

``named`` below is ``{layer_name: [arrays in layer.get_weights() order]}``.  h5py is not part of this image, so the HDF5
reader is optional (used when h5py imports); :func:`read_npz` reads the same mapping from an ``.npz`` export
(keys ``"<layer_name>/<i>"``), which is what a one-line export script on the reference side produces."""
import numpy as np


def read_npz(path):
    """``{layer: [w0, w1, ...]}`` from an .npz whose keys are ``"<layer>/<index>"``."""
    named = {}
    with np.load(path) as z:
        for key in z.files:
            layer, _, i = key.rpartition("/")
            named.setdefault(layer, {})[int(i)] = z[key]
    return {k: [v[i] for i in sorted(v)] for k, v in named.items()}


def write_npz(path, named):
    np.savez(path, **{"%s/%d" % (k, i): np.asarray(w) for k, ws in named.items() for i, w in enumerate(ws)})


def read_keras_hdf5(path):
    """``{layer: [arrays]}`` from a Keras 2.x weight file (the layout ``load_weights`` walks, model_multi.py:2612-2614).
    Needs h5py, which this image does not ship: raises ImportError with that message otherwise."""
    try:
        import h5py
    except ImportError as e:                                   # pragma: no cover - environment dependent
        raise ImportError("reading Keras HDF5 checkpoints needs h5py (not installed here); export the layers with "
                          "write_npz on the reference side and use read_npz") from e
    named = {}
    with h5py.File(path, "r") as f:                            # pragma: no cover
        g = f["model_weights"] if "layer_names" not in f.attrs and "model_weights" in f else f
        for layer in (n.decode() if isinstance(n, bytes) else n for n in g.attrs["layer_names"]):
            names = [n.decode() if isinstance(n, bytes) else n for n in g[layer].attrs["weight_names"]]
            if names:
                named[layer] = [np.asarray(g[layer][n]) for n in names]
    return named


def _bn(named, name):
    w = named.get(name)
    if w is None:
        return None
    if len(w) != 4:
        raise ValueError("BatchNorm layer %r must hold [gamma, beta, moving_mean, moving_variance]" % name)
    return tuple(np.asarray(a, np.float32).reshape(-1) for a in w)


def _conv(named, name, bn_name):
    if name not in named:
        raise KeyError("checkpoint has no layer %r" % name)
    k, b = named[name][:2]
    out = {"W": np.asarray(k, np.float32), "b": np.asarray(b, np.float32).reshape(-1)}
    bn = _bn(named, bn_name)
    if bn is not None:
        out["bn"] = bn
    return out


def fusion_params_from_keras(named, config, levels=(2, 3, 4, 5, 6)):
    """The ``params`` dictionary of :func:`layers.fusion_neck` (also accepted by ``set_weights``) for ``config.GRID_REAS``."""
    mode, out = config.GRID_REAS, {}
    for lvl in levels:
        scope, depth = "grid_reas_P%d" % lvl, "grid_reas_depth_PG%d" % lvl
        bn_name = scope + "_batch_norm"
        g = {}
        if mode in ("add", "mean", "max"):
            if _bn(named, bn_name) is not None:
                g["bn"] = _bn(named, bn_name)
        elif mode == "ident":
            c = _conv(named, scope + "ident_conv", bn_name)
            k = c["W"]
            g = {"weight": k.reshape(k.shape[-2], k.shape[-1]), "bias": c["b"]}
            if "bn" in c:
                g["bn"] = c["bn"]
        elif mode == "lstm3d":
            c = _conv(named, scope + "_convlstm3d", bn_name)
            g = {"W": c["W"], "b": c["b"]}
            if "bn" in c:
                g["bn"] = c["bn"]
        elif mode == "conv3d":
            nc = scope + "_3D_conv"
            g = {"conv1": _conv(named, nc + "_1", bn_name + "_1"), "conv2": _conv(named, nc + "_2", bn_name + "_2"),
                 "deconv1": _conv(named, nc + "_deconv_1", bn_name + "deconv_1"),
                 "deconv2": _conv(named, nc + "_deconv_2", bn_name + "deconv_2")}
        else:
            raise ValueError("GRID_REAS=%r" % (mode,))
        out[scope] = g
        if mode == "conv3d":
            d = {}
            for i in (1, 2):
                dw = named[depth + "_DepthwiseConv_%d" % i]
                d["dw%d" % i] = {"w": np.asarray(dw[0], np.float32).reshape(-1), "b": np.asarray(dw[1], np.float32).reshape(-1)}
                c = _conv(named, depth + "2DConv_%d" % i, depth + "bn_%d" % i)
                c["W"] = c["W"].reshape(c["W"].shape[-2], c["W"].shape[-1])
                d["conv%d" % i] = c
        else:
            c = _conv(named, depth + "2DConv", depth + "bn_deconv")
            d = {"weight": c["W"].reshape(-1), "bias": float(c["b"][0])}
            if "bn" in c:
                d["bn"] = tuple(float(a[0]) for a in c["bn"])
        out[depth] = d
    return out


# ---- image_meta (mrcnn/model_multi.py:3278-3348): [id | original shape (3) | image shape (3) | window (4) | scale | active classes]
META_FIELDS = (("image_id", 0, 1), ("original_image_shape", 1, 4), ("image_shape", 4, 7), ("window", 7, 11), ("scale", 11, 12))


def compose_image_meta(image_id, original_image_shape, image_shape, window, scale, active_class_ids):
    parts = ([image_id], original_image_shape, image_shape, window, [scale], active_class_ids)
    return np.concatenate([np.asarray(p, dtype=np.float64).reshape(-1) for p in parts])


def parse_image_meta(meta):
    """[batch, 12 + NUM_CLASSES] -> dict of typed columns (ints for ids / shapes / window, float32 scale)."""
    meta = np.asarray(meta)
    out = {name: meta[:, lo:hi] for name, lo, hi in META_FIELDS}
    out["active_class_ids"] = meta[:, 12:]
    res = {k: v.astype(np.int32) for k, v in out.items()}
    res["image_id"] = res["image_id"][:, 0]
    res["scale"] = out["scale"][:, 0].astype(np.float32)
    return res
