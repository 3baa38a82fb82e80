"""Pinned fp32 geometry helpers of the oracle (test infrastructure, see oracle/__init__.py).

Every function cites the reference lines it restates.  All values are np.float32 and every
arithmetic op is an individually rounded fp32 op (NumPy elementwise ufuncs never contract
a*b+c into an FMA).
"""
import numpy as np

F32 = np.float32


def tf1_range(start, limit, delta):
    """``tf.range(start, limit, delta)`` for float arguments, TF1 ``RangeOp<float>``.

    Call sites: mrcnn/model_multi.py:157-160 (voxel centres), :252 (pixel centres).
    TF casts the Python doubles to float32 scalars, sizes the output as
    ``ceil(|(limit - start) / delta|)`` in float arithmetic and fills it by SEQUENTIAL
    accumulation ``val += delta`` (third-party kernel, restated; SURVEY.md section 8 spec A.2).
    """
    start, limit, delta = F32(start), F32(limit), F32(delta)
    size = int(np.ceil(np.abs(F32(F32(limit - start) / delta))))
    out = np.empty(size, dtype=F32)
    val = start
    for i in range(size):
        out[i] = val
        val = F32(val + delta)
    return out


def tf1_linspace(start, stop, num):
    """``tf.linspace`` (TF1 ``LinSpaceOp<float>``): ``start + step * i`` with
    ``step = (stop - start) / (num - 1)`` in float.  Call site: model_multi.py:267."""
    start, stop = F32(start), F32(stop)
    out = np.empty(num, dtype=F32)
    if num == 1:
        out[0] = start
        return out
    step = F32(F32(stop - start) / F32(num - 1))
    for i in range(num):
        out[i] = F32(start + F32(step * F32(i)))
    return out


def matmul_seq(a, b):
    """Batched fp32 matmul with the contraction evaluated left-to-right in ascending k:
    ``((a0*b0 + a1*b1) + a2*b2) + a3*b3``.  Stands in for every small
    ``tf.matmul`` on the path (model_multi.py:143,147,180,183,281,283,290)."""
    a = np.asarray(a, dtype=F32)
    b = np.asarray(b, dtype=F32)
    k = a.shape[-1]
    assert b.shape[-2] == k
    acc = a[..., :, 0:1] * b[..., 0:1, :]
    for kk in range(1, k):
        acc = acc + a[..., :, kk:kk + 1] * b[..., kk:kk + 1, :]
    return acc.astype(F32)


def unproj_matrices(Rcam, Kmat):
    """KR_v = (K . [R_v^T | -R_v^T t_v]) . [[R_0|t_0],[0 0 0 1]]   -> [B,V,3,4]

    model_multi.py:135-147 (inverse pose, K.Rinv) and :175-180 (right-multiply by the main
    view's camera->world transform so the grid lives in the main camera frame)."""
    Rcam = np.asarray(Rcam, dtype=F32)
    Kmat = np.asarray(Kmat, dtype=F32)
    B, V = Rcam.shape[:2]
    Rt = np.swapaxes(Rcam[..., :3], -1, -2)                     # :137
    tr = Rcam[..., 3:4]                                         # :138
    tinv = -matmul_seq(Rt, tr)                                  # :143
    Rinv = np.concatenate([Rt, tinv], axis=-1)                  # [B,V,3,4]
    M = matmul_seq(Kmat[:, None], Rinv)                         # :147
    last = np.zeros((B, 1, 4), dtype=F32)
    last[:, 0, 3] = 1.0
    T0 = np.concatenate([Rcam[:, 0], last], axis=1)             # :175-177  [B,4,4]
    KR = matmul_seq(M, T0[:, None])                             # :180
    return KR


def grid_centres(cfg):
    """Voxel centre coordinates gx[X], gy[Y], gz[Z] (model_multi.py:157-160)."""
    g = tf1_range(cfg.vmin + cfg.vsize / 2.0, cfg.vmax, cfg.vsize)
    gz = tf1_range(cfg.vmin_z + cfg.vsize_z / 2.0, cfg.vmax_z, cfg.vsize_z)
    assert g.shape[0] == cfg.nvox, (g.shape, cfg.nvox)
    assert gz.shape[0] == cfg.nvox_z, (gz.shape, cfg.nvox_z)
    return g, g.copy(), gz


def proj_constants(cfg, proj_size):
    """Scalars of ``proj_grid`` (model_multi.py:238,267,294-296): rsz factor, depth samples,
    lo / hi / n of the (asymmetric-z) normalisation."""
    r = F32(float(proj_size) / cfg.IMAGE_SHAPE[0])
    z_s = tf1_linspace(cfg.vmin_z + cfg.vsize_z / 2.0, cfg.vmax_z - cfg.vsize_z / 2.0,
                       cfg.samples)
    lo = np.array([cfg.vmin, cfg.vmin, cfg.vmin_z + cfg.vsize_z / 2.0], dtype=F32)
    hi = np.array([cfg.vmax, cfg.vmax, cfg.vmax_z], dtype=F32)
    n = np.array([cfg.nvox * 1.0, cfg.nvox * 1.0, cfg.nvox_z * 1.0], dtype=F32)
    return r, z_s, lo, hi, n
