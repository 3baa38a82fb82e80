"""Multi-threaded torch-CPU port of the oracle's unproject -> fuse -> project pipeline.
TEST / BASELINE INFRASTRUCTURE ONLY (see oracle/__init__.py).

This is the stand-in for "the reference's TF CPU path" that bench.py times on the GPU box's host
cores (TensorFlow is not installable here, and TF-CPU ``gather_nd`` raises on the out-of-range
taps this path produces).  It follows the reference's graph op for op -- coordinate matmuls,
floor, four full-size ``gather_nd`` results, four weighted products, ``add_n``, transpose,
``reduce_sum``, then the ray-sample ``gather_nd`` (mrcnn/model_multi.py:183-228, :402, :252-322)
-- with vectorised torch ops using every host thread.  Views are processed one at a time only to
bound memory; the arithmetic is identical to ``oracle.unproj_feat`` / ``oracle.proj_grid`` and
tests/test_oracle.py checks that bit for bit on small cases.
"""
import numpy as np
import torch

from . import geometry as G
from .projection import project_indices

INT_MIN = -2 ** 31


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def unproject_fuse(feats, Rcam, Kmat, cfg, mode="sum", x_slab=None):
    """feats [B,V,fh,fw,C] (numpy or torch, fp32) -> fused grid [B,X,Y,Z,C] (torch).
    ``x_slab=(x_begin, x_count)`` restricts the work to one x-slab of the grid (bounded sample for
    the CPU baseline); the result is then [B,x_count,Y,Z,C]."""
    feats = feats if isinstance(feats, torch.Tensor) else _t(feats)
    B, V, fh, fw, C = feats.shape
    KR = _t(G.unproj_matrices(np.asarray(Rcam), np.asarray(Kmat)))          # tiny: NumPy, pinned order
    sy = float(np.float32(float(fh) / cfg.IMAGE_SHAPE[0]))
    sx = float(np.float32(float(fw) / cfg.IMAGE_SHAPE[1]))
    gx, gy, gz = (_t(a) for a in G.grid_centres(cfg))
    if x_slab is not None:
        gx = gx[x_slab[0]:x_slab[0] + x_slab[1]].contiguous()
    X, Y, Z = gx.numel(), gy.numel(), gz.numel()
    # tf.meshgrid 'xy' -> [Y,X,Z], flattened row-major (model_multi.py:163-169)
    my, mx, mz = torch.meshgrid(gy, gx, gz, indexing="ij")
    x, y, z = mx.reshape(-1), my.reshape(-1), mz.reshape(-1)
    N = x.numel()
    out = torch.empty((B, X, Y, Z, C), dtype=torch.float32)
    for b in range(B):
        acc = None
        for v in range(V):
            k = KR[b, v]
            px = ((k[0, 0] * x + k[0, 1] * y) + k[0, 2] * z) + k[0, 3]
            py = ((k[1, 0] * x + k[1, 1] * y) + k[1, 2] * z) + k[1, 3]
            pz = ((k[2, 0] * x + k[2, 1] * y) + k[2, 2] * z) + k[2, 3]
            u = (px / pz) * sx
            w = (py / pz) * sy
            ok = torch.isfinite(u) & torch.isfinite(w) & (u.abs() < 2.0 ** 30) & (w.abs() < 2.0 ** 30)
            u = torch.where(ok, u, torch.zeros_like(u))
            w = torch.where(ok, w, torch.zeros_like(w))
            x0f, y0f = torch.floor(u), torch.floor(w)
            x0, y0 = x0f.to(torch.int64), y0f.to(torch.int64)
            x1f, y1f = x0f + 1.0, y0f + 1.0
            wa = (x1f - u) * (y1f - w)
            wb = (x1f - u) * (w - y0f)
            wc = (u - x0f) * (y1f - w)
            wd = (u - x0f) * (w - y0f)
            fmap = feats[b, v].reshape(fh * fw, C)

            def tap(yy, xx):            # gather_nd with TF-GPU zero fill
                inb = ok & (yy >= 0) & (yy < fh) & (xx >= 0) & (xx < fw)
                lin = (yy.clamp(0, fh - 1) * fw + xx.clamp(0, fw - 1))
                return fmap.index_select(0, lin) * inb.to(torch.float32).unsqueeze(1)

            Ia, Ib, Ic, Id = tap(y0, x0), tap(y0 + 1, x0), tap(y0, x0 + 1), tap(y0 + 1, x0 + 1)
            val = ((wa.unsqueeze(1) * Ia + wb.unsqueeze(1) * Ib) + wc.unsqueeze(1) * Ic) + wd.unsqueeze(1) * Id
            val = torch.where(ok.unsqueeze(1), val, torch.zeros_like(val))
            if acc is None:
                acc = val
            elif mode == "max":
                acc = torch.maximum(acc, val)
            else:
                acc = acc + val
        if mode == "mean":
            acc = acc * float(np.float32(1.0) / np.float32(V))
        out[b] = acc.reshape(Y, X, Z, C).permute(1, 0, 2, 3)                 # :223-227
    return out


def project(grid, Rcam, Kmat, cfg, proj_size):
    """fused grid [B,X,Y,Z,C] (torch) -> ray slices [B,S,P,P,C] (torch)."""
    idx, valid = project_indices(np.asarray(Rcam), np.asarray(Kmat), cfg, proj_size)   # small: NumPy, pinned order
    B, X, Y, Z, C = grid.shape
    S, ph, pw = idx.shape[1:4]
    out = torch.zeros((B, S, ph, pw, C), dtype=torch.float32)
    idx_t, valid_t = _t(idx.astype(np.int64)), _t(valid)
    for b in range(B):
        ii = idx_t[b].reshape(-1, 3)
        vv = valid_t[b].reshape(-1)
        lin = (ii[:, 0].clamp(0, X - 1) * Y + ii[:, 1].clamp(0, Y - 1)) * Z + ii[:, 2].clamp(0, Z - 1)
        g = grid[b].reshape(-1, C).index_select(0, lin) * vv.to(torch.float32).unsqueeze(1)
        out[b] = g.reshape(S, ph, pw, C)
    return out


def unproject_fuse_project(feats, Rcam, Kmat, cfg, proj_size, mode="sum"):
    fused = unproject_fuse(feats, Rcam, Kmat, cfg, mode)
    return project(fused, Rcam, Kmat, cfg, proj_size), fused
