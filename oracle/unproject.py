"""Oracle for ``unproj_feat`` (test infrastructure, see oracle/__init__.py).

Restates mrcnn/model_multi.py:130-228 (grid in the main view's camera frame) and the
world-frame notebook variant Notebook/projection.py:47-151, op for op, in NumPy fp32.
"""
import numpy as np

from .geometry import F32, matmul_seq, unproj_matrices, grid_centres, tf1_range

INT_MIN = np.int32(-2 ** 31)
_COORD_LIMIT = F32(2.0 ** 30)


def unproject_coords(KR, xs, ys, zs, sx, sy):
    """Project points into every view's feature map.

    KR [B,V,3,4]; xs/ys/zs [B,N] point coordinates -> (u, w) [B,V,N] feature-map pixel
    coordinates.  model_multi.py:183-188: ``im_p = KR . (x,y,z,1)`` as 4-term dot products in
    ascending k, divide by depth first, then scale by fw/IMAGE_W resp. fh/IMAGE_H."""
    x = xs[:, None, :]
    y = ys[:, None, :]
    z = zs[:, None, :]
    one = F32(1.0)

    def row(i):
        k = KR[:, :, i, :]
        return ((k[..., 0:1] * x + k[..., 1:2] * y) + k[..., 2:3] * z) + k[..., 3:4] * one

    with np.errstate(all="ignore"):
        px, py, pz = row(0), row(1), row(2)
        u = (px / pz) * F32(sx)
        w = (py / pz) * F32(sy)
    return u.astype(F32), w.astype(F32)


def bilinear_taps(u, w, fh, fw):
    """floor / +1 taps, the four weights and per-tap in-bounds bits (model_multi.py:192-217).

    Returns x0,y0 (int32, INT_MIN where the coordinate is not usable), weights wa..wd (fp32)
    and ``valid`` (uint8; bit0=(y0,x0) bit1=(y1,x0) bit2=(y0,x1) bit3=(y1,x1)).
    Pinned: a non-finite or |coord| >= 2**30 coordinate makes all four taps invalid
    (undefined in the reference: inf/NaN weights times a zero-filled tap)."""
    with np.errstate(all="ignore"):
        ok = np.isfinite(u) & np.isfinite(w) & (np.abs(u) < _COORD_LIMIT) & (np.abs(w) < _COORD_LIMIT)
        us = np.where(ok, u, F32(0))
        ws = np.where(ok, w, F32(0))
        x0 = np.floor(us).astype(np.int32)
        y0 = np.floor(ws).astype(np.int32)
        x1 = x0 + 1
        y1 = y0 + 1
        x0f, x1f = x0.astype(F32), x1.astype(F32)
        y0f, y1f = y0.astype(F32), y1.astype(F32)
        wa = (x1f - us) * (y1f - ws)
        wb = (x1f - us) * (ws - y0f)
        wc = (us - x0f) * (y1f - ws)
        wd = (us - x0f) * (ws - y0f)
    inx0 = (x0 >= 0) & (x0 < fw)
    inx1 = (x1 >= 0) & (x1 < fw)
    iny0 = (y0 >= 0) & (y0 < fh)
    iny1 = (y1 >= 0) & (y1 < fh)
    valid = ((iny0 & inx0).astype(np.uint8) | ((iny1 & inx0).astype(np.uint8) << 1)
             | ((iny0 & inx1).astype(np.uint8) << 2) | ((iny1 & inx1).astype(np.uint8) << 3))
    valid = np.where(ok, valid, np.uint8(0)).astype(np.uint8)
    x0 = np.where(ok, x0, INT_MIN).astype(np.int32)
    y0 = np.where(ok, y0, INT_MIN).astype(np.int32)
    return x0, y0, (wa.astype(F32), wb.astype(F32), wc.astype(F32), wd.astype(F32)), valid


def gather_zero_fill(feats_bv, y, x, ok):
    """TF-GPU ``gather_nd`` semantics: an out-of-range (y, x) reads zeros
    (call sites model_multi.py:209-212).  feats_bv [fh,fw,C]; y,x,ok [N] -> [N,C]."""
    fh, fw, _ = feats_bv.shape
    yc = np.clip(y, 0, fh - 1)
    xc = np.clip(x, 0, fw - 1)
    out = feats_bv[yc, xc]
    return np.where(ok[:, None], out, F32(0)).astype(F32)


def _sample_views(feats, u, w):
    """Bilinear sampling with per-tap zero fill; ``((Ia+Ib)+Ic)+Id`` (tf.add_n, :220).
    feats [B,V,fh,fw,C], u/w [B,V,N] -> vals [B,V,N,C], x0,y0 [B,V,N], valid [B,V,N]."""
    B, V, fh, fw, C = feats.shape
    N = u.shape[-1]
    x0, y0, (wa, wb, wc, wd), valid = bilinear_taps(u, w, fh, fw)
    vals = np.zeros((B, V, N, C), dtype=F32)
    for b in range(B):
        for v in range(V):
            f = feats[b, v]
            vb = valid[b, v]
            xx0 = np.where(x0[b, v] == INT_MIN, 0, x0[b, v])
            yy0 = np.where(y0[b, v] == INT_MIN, 0, y0[b, v])
            Ia = gather_zero_fill(f, yy0, xx0, (vb & 1) != 0)
            Ib = gather_zero_fill(f, yy0 + 1, xx0, (vb & 2) != 0)
            Ic = gather_zero_fill(f, yy0, xx0 + 1, (vb & 4) != 0)
            Id = gather_zero_fill(f, yy0 + 1, xx0 + 1, (vb & 8) != 0)
            ok = (x0[b, v] != INT_MIN)[:, None]
            with np.errstate(all="ignore"):
                val = ((wa[b, v][:, None] * Ia + wb[b, v][:, None] * Ib)
                       + wc[b, v][:, None] * Ic) + wd[b, v][:, None] * Id
            vals[b, v] = np.where(ok, val, F32(0))
    return vals, x0, y0, valid


def unproj_feat(feats, Rcam, Kmat, cfg, return_aux=False):
    """``unproj_feat([feats, Rcam, Kmat], config)``  (mrcnn/model_multi.py:130-228).

    feats [B,V,fh,fw,C] f32, Rcam [B,V,3,4] camera->world, Kmat [B,3,3]
    -> [B,V,X,Y,Z,C] (index order ix,iy,iz after the transpose at :227).
    With ``return_aux`` also returns idx [B,V,X,Y,Z,2] = (y0,x0) int32 and
    valid [B,V,X,Y,Z] uint8 (4 tap bits)."""
    feats = np.ascontiguousarray(feats, dtype=F32)
    B, V, fh, fw, C = feats.shape
    KR = unproj_matrices(Rcam, Kmat)
    sy = F32(float(fh) / cfg.IMAGE_SHAPE[0])                    # :153
    sx = F32(float(fw) / cfg.IMAGE_SHAPE[1])                    # :154
    gx, gy, gz = grid_centres(cfg)
    X, Y, Z = gx.shape[0], gy.shape[0], gz.shape[0]
    # tf.meshgrid default 'xy' indexing: arrays [Y,X,Z] (:163-167)
    mx, my, mz = np.meshgrid(gx, gy, gz, indexing="xy")
    xs = np.broadcast_to(mx.reshape(1, -1), (B, X * Y * Z))
    ys = np.broadcast_to(my.reshape(1, -1), (B, X * Y * Z))
    zs = np.broadcast_to(mz.reshape(1, -1), (B, X * Y * Z))
    u, w = unproject_coords(KR, xs, ys, zs, sx, sy)
    vals, x0, y0, valid = _sample_views(feats, u, w)
    out = vals.reshape(B, V, Y, X, Z, C).transpose(0, 1, 3, 2, 4, 5)      # :223-227
    out = np.ascontiguousarray(out)
    if not return_aux:
        return out
    idx = np.stack([y0, x0], axis=-1).reshape(B, V, Y, X, Z, 2).transpose(0, 1, 3, 2, 4, 5)
    vld = valid.reshape(B, V, Y, X, Z).transpose(0, 1, 3, 2, 4)
    return out, np.ascontiguousarray(idx), np.ascontiguousarray(vld)


def notebook_grid(Rcam, cfg):
    """World-frame grid of the notebook variant (Notebook/projection.py:78-97): centres and
    ``grid_position = [R_0|t_0] . (0,0,grid_dist,1)`` per scene."""
    Rcam = np.asarray(Rcam, dtype=F32)
    B = Rcam.shape[0]
    g = tf1_range(cfg.vmin + cfg.vsize / 2.0, cfg.vmax, cfg.vsize)
    gz = tf1_range(-(cfg.nvox_z - 1) * 0.5 * cfg.vsize,
                   (cfg.nvox_z - 1) * 0.5 * cfg.vsize + cfg.vsize / 2, cfg.vsize)
    grid_dist = cfg.GRID_DIST if hasattr(cfg, "GRID_DIST") else 600 / 320 * cfg.vmax
    p = np.array([[0.0], [0.0], [grid_dist], [1.0]], dtype=F32)
    gp = matmul_seq(Rcam[:, 0], p[None])[..., 0]                # [B,3]
    return g, gz, gp


def unproj_feat_notebook(feats, Rcam, Kmat, cfg, return_aux=False):
    """World-axis-aligned variant (Notebook/projection.py:47-151): no T0 multiplication, grid
    centred at ``grid_position``; returns ``[grid, grid_position]``.  The reference only
    works for B=1 (its meshgrid flattens the batch offsets); here the B=1 behaviour is applied
    per scene."""
    feats = np.ascontiguousarray(feats, dtype=F32)
    Rcam = np.asarray(Rcam, dtype=F32)
    Kmat = np.asarray(Kmat, dtype=F32)
    B, V, fh, fw, C = feats.shape
    Rt = np.swapaxes(Rcam[..., :3], -1, -2)
    tinv = -matmul_seq(Rt, Rcam[..., 3:4])
    KR = matmul_seq(Kmat[:, None], np.concatenate([Rt, tinv], axis=-1))
    sy = F32(float(fh) / cfg.IMAGE_SHAPE[0])
    sx = F32(float(fw) / cfg.IMAGE_SHAPE[1])
    g, gz, gp = notebook_grid(Rcam, cfg)
    X = Y = g.shape[0]
    Z = gz.shape[0]
    xs = np.empty((B, X * Y * Z), F32)
    ys = np.empty_like(xs)
    zs = np.empty_like(xs)
    for b in range(B):
        mx, my, mz = np.meshgrid(g + gp[b, 0], g + gp[b, 1], gz + gp[b, 2], indexing="xy")
        xs[b], ys[b], zs[b] = mx.reshape(-1), my.reshape(-1), mz.reshape(-1)
    u, w = unproject_coords(KR, xs, ys, zs, sx, sy)
    vals, x0, y0, valid = _sample_views(feats, u, w)
    out = np.ascontiguousarray(vals.reshape(B, V, Y, X, Z, C).transpose(0, 1, 3, 2, 4, 5))
    if not return_aux:
        return out, gp
    idx = np.stack([y0, x0], axis=-1).reshape(B, V, Y, X, Z, 2).transpose(0, 1, 3, 2, 4, 5)
    vld = valid.reshape(B, V, Y, X, Z).transpose(0, 1, 3, 2, 4)
    return out, gp, np.ascontiguousarray(idx), np.ascontiguousarray(vld)
