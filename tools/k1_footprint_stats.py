"""Footprint statistics of workload T for the K1 design decision.

For every (voxel tile, view) it measures the bounding box of the in-map bilinear taps of the tile's
voxels: that box is the K extent a tensor-core formulation of the interpolation
(out[128 voxels, C] += W[128, K] . F[K, C]) would have to contract over.  CPU only (NumPy); uses the
oracle's coordinate code, so it runs in the build container:  python tools/k1_footprint_stats.py
"""
import itertools
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from mulit_view_object_detection_b200 import synthetic as syn          # noqa: E402
from mulit_view_object_detection_b200.config import FusionConfig       # noqa: E402
from oracle.geometry import unproj_matrices, grid_centres, F32         # noqa: E402
from oracle.unproject import unproject_coords, bilinear_taps          # noqa: E402


def taps_of_scene(cfg, Rcam, Kmat, fh, fw):
    KR = unproj_matrices(Rcam, Kmat)
    sy = F32(float(fh) / cfg.IMAGE_SHAPE[0])
    sx = F32(float(fw) / cfg.IMAGE_SHAPE[1])
    gx, gy, gz = grid_centres(cfg)
    X, Y, Z = len(gx), len(gy), len(gz)
    mx, my, mz = np.meshgrid(gx, gy, gz, indexing="ij")        # [X,Y,Z]
    u, w = unproject_coords(KR, mx.reshape(1, -1), my.reshape(1, -1), mz.reshape(1, -1), sx, sy)
    x0, y0, _, valid = bilinear_taps(u, w, fh, fw)
    V = Rcam.shape[1]
    return x0.reshape(V, X, Y, Z), y0.reshape(V, X, Y, Z), valid.reshape(V, X, Y, Z)


def tile_boxes(x0, y0, valid, tile, fh, fw):
    """Per (view, tile): clipped tap bounding box (w, h) or (0, 0) when nothing is sampled."""
    V, X, Y, Z = x0.shape
    tx, ty, tz = tile
    out = []
    big = 1 << 20
    # in-map extent of the taps of each voxel
    inx0 = (valid & 0b0011) != 0          # column x0 has an in-map tap
    inx1 = (valid & 0b1100) != 0
    iny0 = (valid & 0b0101) != 0
    iny1 = (valid & 0b1010) != 0
    xmin = np.where(inx0, x0, np.where(inx1, x0 + 1, big))
    xmax = np.where(inx1, x0 + 1, np.where(inx0, x0, -big))
    ymin = np.where(iny0, y0, np.where(iny1, y0 + 1, big))
    ymax = np.where(iny1, y0 + 1, np.where(iny0, y0, -big))

    def red(a, f):
        a = a.reshape(V, X // tx, tx, Y // ty, ty, Z // tz, tz)
        return f(f(f(a, axis=6), axis=4), axis=2)
    bx0, bx1 = red(xmin, np.min), red(xmax, np.max)
    by0, by1 = red(ymin, np.min), red(ymax, np.max)
    any_valid = red((valid != 0).astype(np.int32), np.max) > 0
    w = np.where(any_valid, bx1 - bx0 + 1, 0)
    h = np.where(any_valid, by1 - by0 + 1, 0)
    nvalid = red((valid != 0).astype(np.int32), np.sum)
    return w, h, any_valid, nvalid


def main():
    T = dict(V=8, nvox=64, fh=40, fw=40, C=256)
    cfg = FusionConfig(nvox=T["nvox"], nvox_z=T["nvox"], samples=20, NUM_VIEWS=T["V"])
    scenes = 4
    _, Rcam, Kmat = syn.make_scene(cfg, scenes, T["V"], T["fh"], T["fw"], 4, seed=1000)
    shapes = [(4, 4, 8), (2, 2, 32), (2, 4, 16), (8, 8, 2), (4, 8, 4), (1, 2, 64), (2, 8, 8), (4, 2, 16)]
    print("tile      live%%  mean_area  K16(box)  p50  p90  p99   K(4x2 patches)  K(8x2 patches)  MMA us/scene@K16box (3 f16 MMAs, 1.9 GHz)")
    for tile in shapes:
        areas, k16, kp42, kp82, live, total = [], [], [], [], 0, 0
        for b in range(scenes):
            x0, y0, valid = taps_of_scene(cfg, Rcam[b:b + 1], Kmat[b:b + 1], T["fh"], T["fw"])
            w, h, anyv, _ = tile_boxes(x0, y0, valid, tile, T["fh"], T["fw"])
            total += anyv.size
            live += int(anyv.sum())
            a = (w * h)[anyv]
            areas.append(a)
            k16.append(((a + 15) // 16) * 16)
            kp42.append((((w + 3) // 4) * ((h + 1) // 2))[anyv] * 8)
            kp82.append((((w + 7) // 8) * ((h + 1) // 2))[anyv] * 16)
        a = np.concatenate(areas)
        k = np.concatenate(k16)
        k42 = np.concatenate(kp42)
        k42 = ((k42 + 15) // 16) * 16
        k82 = np.concatenate(kp82)
        # MMA time: per live (tile, view): K/16 k-steps x 3 MMAs x 128 cycles (M=128, N=256), 148 SMs, 1.9 GHz
        per_scene = live / scenes
        us = per_scene * (k.mean() / 16) * 3 * 128 / 148 / 1.9e3
        us42 = per_scene * (k42.mean() / 16) * 3 * 128 / 148 / 1.9e3
        print(f"{str(tile):9s} {100 * live / total:5.1f}  {a.mean():8.1f}  {k.mean():8.1f}  {np.percentile(k, 50):4.0f} {np.percentile(k, 90):4.0f} "
              f"{np.percentile(k, 99):4.0f}   {k42.mean():8.1f}        {k82.mean():8.1f}        {us:6.1f} (box) {us42:6.1f} (4x2)"
              f"   frac K<=16: {np.mean(k <= 16):.2f}  K<=32: {np.mean(k <= 32):.2f}")


if __name__ == "__main__":
    main()
