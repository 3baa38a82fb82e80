"""Oracle for ``proj_grid`` / ``nearest3`` / ``depth_sampling`` (test infrastructure, see
oracle/__init__.py).

Restates mrcnn/model_multi.py:231-322, :357-369, :466-488 (non-``conv3d`` branch) and the
notebook variant Notebook/projection.py:253-339, in NumPy fp32 with the evaluation order of
SURVEY.md Appendix A.
"""
import numpy as np

from .geometry import F32, matmul_seq, proj_constants, tf1_linspace, tf1_range
from .fusion import batch_norm_affine

INT_MIN = np.int32(-2 ** 31)
_COORD_LIMIT = F32(2.0 ** 30)


def _as_hw(proj_size):
    if isinstance(proj_size, (tuple, list)):
        return int(proj_size[0]), int(proj_size[1])
    return int(proj_size), int(proj_size)


def _affine34(M, p):
    """[B,3,4] . (p,1) with p = (x,y,z) arrays of shape [B,...]: 4-term dots, ascending k."""
    x, y, z = p
    one = F32(1.0)
    sh = (slice(None),) + (None,) * (x.ndim - 1)

    def row(i):
        m = M[:, i, :]
        return ((m[:, 0][sh] * x + m[:, 1][sh] * y) + m[:, 2][sh] * z) + m[:, 3][sh] * one
    return row(0), row(1), row(2)


def project_indices(Rcam, Kmat, cfg, proj_size, view=0, notebook_grid_pos=None):
    """Voxel index of every ray sample: int32 [B,S,Ph,Pw,3] (ix,iy,iz) and valid [B,S,Ph,Pw].

    model_multi.py:238-298 + the ``tf.round`` of ``nearest3`` (:361, round-half-even).
    ``view`` selects which camera the rays belong to (the reference only does view 0, :245);
    the world->grid transform always uses view 0's pose (:279-290).  A non-square
    ``proj_size=(ph,pw)`` uses ``r = ph / IMAGE_SHAPE[0]`` (the reference is square-only).
    ``notebook_grid_pos`` [B,3] switches to the notebook variant
    (Notebook/projection.py:287-313)."""
    Rcam = np.asarray(Rcam, dtype=F32)
    Kmat = np.asarray(Kmat, dtype=F32)
    B = Rcam.shape[0]
    ph, pw = _as_hw(proj_size)
    r, z_s, lo, hi, n = proj_constants(cfg, ph)
    if notebook_grid_pos is not None:
        grid_dist = cfg.GRID_DIST if hasattr(cfg, "GRID_DIST") else 600 / 320 * cfg.vmax
        z_s = tf1_linspace(grid_dist - cfg.vmax * 0.8, grid_dist + cfg.vmax * 0.8, cfg.samples)
        lo = np.array([cfg.vmin, cfg.vmin, -cfg.nvox_z * 0.5 * cfg.vsize], dtype=F32)
        hi = np.array([cfg.vmax, cfg.vmax, cfg.nvox_z * 0.5 * cfg.vsize], dtype=F32)
    Kp = (Kmat * r).astype(F32)                                  # :239 (all nine entries)
    vs = tf1_range(0.5, ph, 1)                                   # rows   (:252)
    us = tf1_range(0.5, pw, 1)                                   # cols
    u = np.broadcast_to(us[None, None, :], (B, ph, pw))
    v = np.broadcast_to(vs[None, :, None], (B, ph, pw))
    k = lambda i, j: Kp[:, i, j][:, None, None]
    with np.errstate(all="ignore"):
        # back substitution of the upper-triangular solve (:263-264), rhs = (u, v, r)
        zc = np.broadcast_to(r / k(2, 2), (B, ph, pw)).astype(F32)
        yc = ((v - k(1, 2) * zc) / k(1, 1)).astype(F32)
        xc = (((u - k(0, 1) * yc) - k(0, 2) * zc) / k(0, 0)).astype(F32)
        zs = z_s[None, :, None, None]
        X = (xc[:, None] * zs, yc[:, None] * zs, zc[:, None] * zs)          # :270-272
        Xw = _affine34(Rcam[:, view], X)                                     # :283
        if notebook_grid_pos is None:
            R0t = np.swapaxes(Rcam[:, 0, :, :3], -1, -2)
            t0inv = -matmul_seq(R0t, Rcam[:, 0, :, 3:4])
            RT = np.concatenate([R0t, t0inv], axis=-1)                       # :279-281
            Xg = _affine34(RT, Xw)                                           # :290
        else:
            gp = np.asarray(notebook_grid_pos, dtype=F32)
            Xg = tuple(Xw[a] - gp[:, a][:, None, None, None] for a in range(3))
        q = [((Xg[a] - lo[a]) / (hi[a] - lo[a])) * n[a] for a in range(3)]  # :297-298
        q = [qa.astype(F32) for qa in q]
        ok = np.ones(q[0].shape, dtype=bool)
        for qa in q:
            ok &= np.isfinite(qa) & (np.abs(qa) < _COORD_LIMIT)
        idx = [np.rint(np.where(ok, qa, F32(0))).astype(np.int32) for qa in q]   # :361
    dims = (cfg.nvox, cfg.nvox, cfg.nvox_z)
    valid = ok.copy()
    for a in range(3):
        valid &= (idx[a] >= 0) & (idx[a] < dims[a])
    idx = np.stack([np.where(ok, ia, INT_MIN) for ia in idx], axis=-1).astype(np.int32)
    return idx, valid


def proj_grid(grid, Rcam, Kmat, cfg, proj_size, view=0, notebook_grid_pos=None,
              return_aux=False):
    """``proj_grid([grid, Rcam, Kmat], config, proj_size)`` (model_multi.py:231-322).

    grid [B,X,Y,Z,C] -> ray slices [B,S,Ph,Pw,C]; nearest-neighbour gather
    (``nearest3``, :357-369) with TF-GPU zero fill for out-of-range indices."""
    grid = np.asarray(grid, dtype=F32)
    B, X, Y, Z, C = grid.shape
    assert (X, Y, Z) == (cfg.nvox, cfg.nvox, cfg.nvox_z)
    idx, valid = project_indices(Rcam, Kmat, cfg, proj_size, view, notebook_grid_pos)
    S, ph, pw = idx.shape[1:4]
    out = np.zeros((B, S, ph, pw, C), dtype=F32)
    for b in range(B):
        ii = idx[b][valid[b]]
        out[b][valid[b]] = grid[b, ii[:, 0], ii[:, 1], ii[:, 2]]
    if return_aux:
        return out, idx, valid.astype(np.uint8)
    return out


def depth_sampling(x, weight, bias, bn):
    """``depth_sampling`` non-``conv3d`` branch (model_multi.py:481-487): one 1x1 conv over
    the S axis shared by all channels (weight [S], bias scalar), a scalar BatchNorm
    (bn = (gamma,beta,mean,var) scalars or None) and ReLU.  [B,S,P,P,C] -> [B,P,P,C].
    The sum over s is evaluated in ascending s."""
    x = np.asarray(x, dtype=F32)
    weight = np.asarray(weight, dtype=F32)
    S = x.shape[1]
    acc = x[:, 0] * weight[0]
    for s in range(1, S):
        acc = acc + x[:, s] * weight[s]
    acc = acc + F32(bias)
    if bn is not None:
        scale, shift = batch_norm_affine(*bn)
        acc = acc * scale + shift
    return np.maximum(acc, F32(0)).astype(F32)
