"""Oracle for view fusion ``grid_reas`` and the recurrent ConvLSTM voxel update
(test infrastructure, see oracle/__init__.py).

Restates mrcnn/model_multi.py:394-463 (``add``, ``ident``, ``lstm3d``), the notebook
``mean`` (Notebook/projection.py:526-529) and mrcnn/recurrent.py:143-173,414-479.
"""
import numpy as np

from .geometry import F32

BN_EPS = 1e-3          # Keras BatchNormalization default epsilon (model_multi.py:501-502)


def batch_norm_affine(gamma, beta, mean, var, eps=BN_EPS):
    """Frozen-statistics BatchNorm as the affine ``x*scale + shift`` that
    ``tf.nn.batch_normalization`` evaluates: ``inv = rsqrt(var+eps)*gamma``,
    ``shift = beta - mean*inv`` (TRAIN_BN=False, mrcnn/config.py:208)."""
    gamma, beta, mean, var = (np.asarray(a, dtype=F32) for a in (gamma, beta, mean, var))
    inv = (F32(1.0) / np.sqrt(var + F32(eps))).astype(F32) * gamma
    shift = beta - mean * inv
    return inv.astype(F32), shift.astype(F32)


def fuse_views(grids, mode):
    """Reduce [B,V,X,Y,Z,C] over the view axis.

    ``sum``  : ``K.sum(x, axis=1)`` (model_multi.py:402), evaluated in ascending v.
    ``mean`` : ``sum * float32(1/V)``         -- defined by this oracle (parity unpinned:
    ``max``  : ``max_v`` of the zero-filled      the reference has neither, SURVEY.md spec B).
               per-view samples, ascending v"""
    grids = np.asarray(grids, dtype=F32)
    V = grids.shape[1]
    acc = grids[:, 0].copy()
    if mode in ("sum", "add", "mean"):
        for v in range(1, V):
            acc = acc + grids[:, v]
        if mode == "mean":
            acc = acc * F32(F32(1.0) / F32(V))
    elif mode == "max":
        for v in range(1, V):
            acc = np.maximum(acc, grids[:, v])
    else:
        raise ValueError(mode)
    return acc.astype(F32)


def channel_mean(grids):
    """The notebook's ``GRID_REAS='mean'`` (Notebook/projection.py:526-529,549): the mean is
    taken over the CHANNEL axis and the V per-view scalars become the channels, then ReLU.
    [B,V,X,Y,Z,C] -> [B,X,Y,Z,V]."""
    grids = np.asarray(grids, dtype=F32)
    C = grids.shape[-1]
    acc = grids[..., 0].copy()
    for c in range(1, C):
        acc = acc + grids[..., c]
    m = acc / F32(C)
    return np.maximum(np.transpose(m, (0, 2, 3, 4, 1)), F32(0)).astype(F32)


def ident_fuse(grids, weight, bias, bn):
    """``GRID_REAS='ident'`` (model_multi.py:443-455): ReLU -> concat views on the channel
    axis (view-major: v*C + c) -> Conv3D 1x1x1 (+bias) -> BN -> ReLU.
    weight [V*C, Cout] (the [1,1,1,V*C,Cout] Keras kernel squeezed), bias [Cout],
    bn = (scale, shift) from :func:`batch_norm_affine`.  Contraction accumulated in float64
    and rounded once (conv summation order is not observable; compare with a tolerance)."""
    grids = np.asarray(grids, dtype=F32)
    B, V, X, Y, Z, C = grids.shape
    x = np.maximum(grids, F32(0))
    x = np.transpose(x, (0, 2, 3, 4, 1, 5)).reshape(B, X, Y, Z, V * C)
    y = (x.astype(np.float64) @ np.asarray(weight, dtype=np.float64)).astype(F32)
    y = y + np.asarray(bias, dtype=F32)
    scale, shift = bn
    return np.maximum(y * scale + shift, F32(0)).astype(F32)


def _sigmoid(x):
    x = x.astype(np.float64)
    return (1.0 / (1.0 + np.exp(-x)))


def conv3d_same(x, W):
    """``tf.nn.convolution(x, W, 'SAME')`` for a [B,X,Y,Z,Cin] input and a [kx,ky,kz,Cin,Cout]
    filter, stride 1, zero padding (recurrent.py:457).  float64 accumulation."""
    x = np.asarray(x, dtype=np.float64)
    W = np.asarray(W, dtype=np.float64)
    B, X, Y, Z, Cin = x.shape
    kx, ky, kz, _, Cout = W.shape
    px, py, pz = kx // 2, ky // 2, kz // 2
    xp = np.zeros((B, X + kx - 1, Y + ky - 1, Z + kz - 1, Cin))
    xp[:, px:px + X, py:py + Y, pz:pz + Z] = x
    y = np.zeros((B, X, Y, Z, Cout))
    for a in range(kx):
        for b_ in range(ky):
            for c in range(kz):
                y += xp[:, a:a + X, b_:b_ + Y, c:c + Z] @ W[a, b_, c]
    return y


def _same_pad(n, k, s):
    """TensorFlow 'SAME': out = ceil(n/s); total = max((out-1)*s + k - n, 0); the odd element goes to the END."""
    out = -(-n // s)
    total = max((out - 1) * s + k - n, 0)
    return out, total // 2, total - total // 2


def conv3d_strided_same(x, W, stride):
    """``KL.Conv3D(strides=stride, padding='same')`` without bias (model_multi.py:415-428): TF 'SAME' padding (for an
    even size, k=3, stride 2: nothing before, one zero after, i.e. ``in = 2*o + k``).  x [B,X,Y,Z,Cin],
    W [kx,ky,kz,Cin,Cout] (cross-correlation, as TF).  float64 accumulation."""
    x = np.asarray(x, dtype=np.float64)
    W = np.asarray(W, dtype=np.float64)
    B, X, Y, Z, Cin = x.shape
    kx, ky, kz, _, Cout = W.shape
    (ox, bx, ax), (oy, by, ay), (oz, bz, az) = _same_pad(X, kx, stride), _same_pad(Y, ky, stride), _same_pad(Z, kz, stride)
    xp = np.zeros((B, X + bx + ax, Y + by + ay, Z + bz + az, Cin))
    xp[:, bx:bx + X, by:by + Y, bz:bz + Z] = x
    y = np.zeros((B, ox, oy, oz, Cout))
    for a in range(kx):
        for b_ in range(ky):
            for c in range(kz):
                y += xp[:, a:a + (ox - 1) * stride + 1:stride, b_:b_ + (oy - 1) * stride + 1:stride,
                        c:c + (oz - 1) * stride + 1:stride] @ W[a, b_, c]
    return y


def conv3d_transpose_same(x, Wt, stride):
    """``KL.Conv3DTranspose(strides=stride, padding='same')`` without bias (model_multi.py:430-441) =
    ``tf.nn.conv3d_transpose``: the gradient of the forward SAME convolution (of the 2x-sized tensor) with respect to
    its input, so ``out[s*o + k - pad_before] += x[o] . Wt[k]`` with the FORWARD conv's pad_before (0 for k=3, stride 2),
    output size ``stride * in``.  Wt is the Keras kernel [kx,ky,kz,Cout,Cin]."""
    x = np.asarray(x, dtype=np.float64)
    Wt = np.asarray(Wt, dtype=np.float64)
    B, X, Y, Z, Cin = x.shape
    kx, ky, kz, Cout, _ = Wt.shape
    OX, OY, OZ = X * stride, Y * stride, Z * stride
    pbx, pby, pbz = _same_pad(OX, kx, stride)[1], _same_pad(OY, ky, stride)[1], _same_pad(OZ, kz, stride)[1]
    big = np.zeros((B, OX + kx, OY + ky, OZ + kz, Cout))
    for a in range(kx):
        for b_ in range(ky):
            for c in range(kz):
                big[:, a:a + X * stride:stride, b_:b_ + Y * stride:stride, c:c + Z * stride:stride] += x @ Wt[a, b_, c].T
    return big[:, pbx:pbx + OX, pby:pby + OY, pbz:pbz + OZ]


def _conv_bn_relu(y, bias, bn):
    y = (y + np.asarray(bias, np.float64)).astype(F32)
    if bn is not None:
        scale, shift = batch_norm_affine(*bn)
        y = y * scale + shift
    return np.maximum(y, F32(0)).astype(F32)


def unet_fuse(grids, params):
    """``GRID_REAS='conv3d'`` (model_multi.py:406-441), the MLF-style U-Net: views concatenated on channels
    (view-major, :411-413) -> ReLU -> Conv3D(2F,s2) -> Conv3D(4F,s2) -> Conv3DTranspose(2F,s2) -> concat with the first
    encoder output (deconv first, :438) -> Conv3DTranspose(F,s2); BN + ReLU after every conv.
    params: {'conv1','conv2','deconv1','deconv2'} each {'W','b','bn' optional}."""
    grids = np.asarray(grids, dtype=F32)
    B, V, X, Y, Z, C = grids.shape
    x = np.transpose(grids, (0, 2, 3, 4, 1, 5)).reshape(B, X, Y, Z, V * C)
    x = np.maximum(x, F32(0))
    p = params
    conv1 = _conv_bn_relu(conv3d_strided_same(x, p["conv1"]["W"], 2), p["conv1"]["b"], p["conv1"].get("bn"))
    conv2 = _conv_bn_relu(conv3d_strided_same(conv1, p["conv2"]["W"], 2), p["conv2"]["b"], p["conv2"].get("bn"))
    deconv1 = _conv_bn_relu(conv3d_transpose_same(conv2, p["deconv1"]["W"], 2), p["deconv1"]["b"], p["deconv1"].get("bn"))
    cat = np.concatenate([deconv1, conv1], axis=4)
    return _conv_bn_relu(conv3d_transpose_same(cat, p["deconv2"]["W"], 2), p["deconv2"]["b"], p["deconv2"].get("bn"))


def depth_sampling_conv3d(x, params):
    """``depth_sampling`` 'conv3d' branch (model_multi.py:467-480): [B,S,P,P,C] -> [B,P,P,C*S] (channel c*S + s) ->
    DepthwiseConv2D 1x1 -> Conv2D 1x1 (512) -> BN -> ReLU -> DepthwiseConv2D 1x1 -> Conv2D 1x1 (F) -> BN -> ReLU.
    params: 'dw1' {'w' [C*S], 'b' [C*S]}, 'conv1' {'W' [C*S,512], 'b', 'bn'}, 'dw2' {'w' [512], 'b'}, 'conv2' {'W' [512,F], 'b', 'bn'}."""
    x = np.asarray(x, dtype=F32)
    B, S, P1, P2, C = x.shape
    y = np.transpose(x, (0, 2, 3, 4, 1)).reshape(B, P1, P2, C * S)
    y = y * np.asarray(params["dw1"]["w"], F32) + np.asarray(params["dw1"]["b"], F32)
    y = _conv_bn_relu(y.astype(np.float64) @ np.asarray(params["conv1"]["W"], np.float64), params["conv1"]["b"], params["conv1"].get("bn"))
    y = y * np.asarray(params["dw2"]["w"], F32) + np.asarray(params["dw2"]["b"], F32)
    return _conv_bn_relu(y.astype(np.float64) @ np.asarray(params["conv2"]["W"], np.float64), params["conv2"]["b"], params["conv2"].get("bn"))


def convlstm_cell_step(x, c_prev, h_prev, W, bias, forget_bias=1.0):
    """One ``ConvLSTMCell.call`` (mrcnn/recurrent.py:442-479, normalize=False):
    ``y = conv3d_SAME([x ; h_prev], W) + b``; gates split in the order
    ``j, i, f, o`` (:460-461); ``c = c_prev*sigmoid(f + forget_bias) + sigmoid(i)*tanh(j)``;
    ``h = tanh(c)*sigmoid(o)`` (:470-477).  Returns (h, c) as fp32."""
    xin = np.concatenate([np.asarray(x, F32), np.asarray(h_prev, F32)], axis=-1)
    y = conv3d_same(xin, W) + np.asarray(bias, dtype=np.float64)
    j, i, f, o = np.split(y, 4, axis=-1)
    c = np.asarray(c_prev, np.float64) * _sigmoid(f + forget_bias) + _sigmoid(i) * np.tanh(j)
    h = np.tanh(c) * _sigmoid(o)
    return h.astype(F32), c.astype(F32)


def convlstm(grids, W, bias, forget_bias=1.0):
    """``convlstm(grid, name, kernel, filters)`` (model_multi.py:109-123) = ``ConvRNN3D`` over
    the view axis with zero initial states shaped like the input (recurrent.py:143-173, so
    C == F), returning the last output only (return_sequences=False)."""
    grids = np.asarray(grids, dtype=F32)
    B, V, X, Y, Z, C = grids.shape
    F = W.shape[-1] // 4
    assert C == F, "initial state takes the input's channel count (recurrent.py:145-147)"
    c = np.zeros((B, X, Y, Z, F), F32)
    h = np.zeros((B, X, Y, Z, F), F32)
    for t in range(V):
        h, c = convlstm_cell_step(grids[:, t], c, h, W, bias, forget_bias)
    return h


def grid_reas(grids, scope, cfg, params=None):
    """``grid_reas(inputs, scope, config)`` (model_multi.py:394-463) for the modes on the hot
    path.  ``params`` carries the frozen learnables of ``scope``:
      add    : {'bn': (gamma,beta,mean,var)}
      ident  : {'weight' [V*C,Cout], 'bias' [Cout], 'bn': ...}
      lstm3d : {'W' [3,3,3,C+F,4F], 'b' [4F], 'bn': ...}
      conv3d : {'conv1','conv2','deconv1','deconv2'} each {'W','b','bn'} (see :func:`unet_fuse`)
      mean / max (oracle-defined, no BN in the reference): optional 'bn'."""
    params = params or {}
    mode = cfg.GRID_REAS
    bn = batch_norm_affine(*params["bn"]) if "bn" in params else None
    if mode == "add":
        x = fuse_views(grids, "sum")                            # :402
        if bn is not None:
            x = x * bn[0] + bn[1]                               # :403
        return np.maximum(x, F32(0)).astype(F32)                # :404
    if mode in ("mean", "max"):
        x = fuse_views(grids, mode)
        if bn is not None:
            x = np.maximum(x * bn[0] + bn[1], F32(0))
        return x.astype(F32)
    if mode == "ident":
        return ident_fuse(grids, params["weight"], params["bias"], bn)
    if mode == "conv3d":
        return unet_fuse(grids, params)
    if mode == "lstm3d":
        x = np.maximum(np.asarray(grids, F32), F32(0))          # :459
        h = convlstm(x, params["W"], params["b"])               # :460
        if bn is not None:
            h = h * bn[0] + bn[1]                               # :461
        return np.maximum(h, F32(0)).astype(F32)                # :462
    raise ValueError("GRID_REAS=%r is not on the hot path" % (mode,))


def fusion_neck(feature_maps, Rcam, Kmat, cfg, params, levels=(2, 3, 4, 5, 6)):
    """The fusion neck of ``MaskRCNN.build`` (model_multi.py:2382-2410): per level ``unproj_feat -> grid_reas -> proj_grid ->
    depth_sampling``; PG2 / PG3 replaced by zeros when ``VANILLA`` is false (:2406-2410)."""
    from .unproject import unproj_feat
    from .projection import proj_grid, depth_sampling
    ih = int(cfg.IMAGE_SHAPE[0])
    outs = []
    for lvl, fm in zip(levels, feature_maps):
        B = fm.shape[0]
        if not getattr(cfg, "VANILLA", False) and lvl in (2, 3):
            z = ih // (4 if lvl == 2 else 8)
            outs.append(np.zeros((B, z, z, int(cfg.TOP_DOWN_PYRAMID_SIZE)), F32))
            continue
        gp, dp = params.get("grid_reas_P%d" % lvl, {}), params["grid_reas_depth_PG%d" % lvl]
        fused = grid_reas(unproj_feat(fm, Rcam, Kmat, cfg), "grid_reas_P%d" % lvl, cfg, gp)
        rays = proj_grid(fused, Rcam, Kmat, cfg, ih // 2 ** lvl)
        if cfg.GRID_REAS == "conv3d":
            outs.append(depth_sampling_conv3d(rays, dp))
        else:
            outs.append(depth_sampling(rays, dp["weight"], dp.get("bias", 0.0), dp.get("bn")))
    return outs
