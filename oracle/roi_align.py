"""Oracle for ``PyramidROIAlign`` (test infrastructure, see oracle/__init__.py).

Restates mrcnn/model_multi.py:774-885: FPN level assignment (:816-828), one
``tf.image.crop_and_resize`` per level (:856-858) and the re-ordering back to box order
(:861-882).  ``crop_and_resize`` itself is a third-party TensorFlow kernel; its arithmetic
(``crop_and_resize_op``: one bilinear sample per bin, extrapolation value 0) is restated
from the published algorithm (SURVEY.md spec E).
"""
import numpy as np

from .geometry import F32

_LOG2 = F32(np.log(np.float64(2.0)))      # tf.log(2.0) as a float32 constant (:776)


def log_f32(x):
    """fp32 natural log, pinned as the correctly rounded value (float64 log rounded once)."""
    with np.errstate(all="ignore"):
        return np.log(np.asarray(x, dtype=np.float64)).astype(F32)


def roi_levels(boxes, image_shape):
    """FPN level per box, int32 in [2,5] (model_multi.py:810-828).

    ``lvl = clamp(4 + int32(round(log2(sqrt(h*w) / (224/sqrt(H*W))))), 2, 5)`` with
    round-half-even; non-finite (zero-area / negative-area boxes) pins to level 2, which is
    what ``4 + INT_MIN`` clamps to on the reference's hardware."""
    boxes = np.asarray(boxes, dtype=F32)
    y1, x1, y2, x2 = (boxes[..., i] for i in range(4))
    h = y2 - y1
    w = x2 - x1
    image_area = F32(float(image_shape[0]) * float(image_shape[1]))
    with np.errstate(all="ignore"):
        denom = F32(224.0) / np.sqrt(image_area).astype(F32)
        ratio = (np.sqrt(h * w).astype(F32) / denom).astype(F32)
        lvl_f = (log_f32(ratio) / _LOG2).astype(F32)
        finite = np.isfinite(lvl_f)
        r = np.rint(np.where(finite, lvl_f, F32(0)))
        r = np.clip(r, -64, 64).astype(np.int32)
    lvl = np.where(finite, np.minimum(5, np.maximum(2, 4 + r)), 2)
    return lvl.astype(np.int32)


def crop_and_resize(image, boxes, box_ind, crop_size):
    """``tf.image.crop_and_resize(image, boxes, box_ind, crop_size, 'bilinear', 0)``.

    image [B,H,W,C]; boxes [n,4] normalised (y1,x1,y2,x2); -> [n,ph,pw,C].  fp32 arithmetic
    in the kernel's own order: ``in_y = y1*(H-1) + iy*((y2-y1)*(H-1)/(ph-1))``;
    out of ``[0,H-1]`` -> 0; ``top = tl + (tr-tl)*lx``; ``val = top + (bottom-top)*ly``."""
    image = np.asarray(image, dtype=F32)
    boxes = np.asarray(boxes, dtype=F32)
    _, H, W, C = image.shape
    ph, pw = crop_size
    n = boxes.shape[0]
    out = np.zeros((n, ph, pw, C), dtype=F32)
    Hm1, Wm1 = F32(H - 1), F32(W - 1)
    for i in range(n):
        y1, x1, y2, x2 = boxes[i]
        img = image[box_ind[i]]
        with np.errstate(all="ignore"):
            hs = F32(F32(F32(y2 - y1) * Hm1) / F32(ph - 1)) if ph > 1 else F32(0)
            ws = F32(F32(F32(x2 - x1) * Wm1) / F32(pw - 1)) if pw > 1 else F32(0)
            iy = np.arange(ph, dtype=F32)
            ix = np.arange(pw, dtype=F32)
            if ph > 1:
                in_y = (F32(y1 * Hm1) + iy * hs).astype(F32)
            else:
                in_y = np.full(1, F32(F32(F32(0.5) * F32(y1 + y2)) * Hm1), dtype=F32)
            if pw > 1:
                in_x = (F32(x1 * Wm1) + ix * ws).astype(F32)
            else:
                in_x = np.full(1, F32(F32(F32(0.5) * F32(x1 + x2)) * Wm1), dtype=F32)
            oky = (in_y >= 0) & (in_y <= Hm1)          # NaN -> False -> extrapolation value
            okx = (in_x >= 0) & (in_x <= Wm1)
            sy = np.where(oky, in_y, F32(0))
            sx = np.where(okx, in_x, F32(0))
            top = np.floor(sy).astype(np.int64)
            bot = np.ceil(sy).astype(np.int64)
            left = np.floor(sx).astype(np.int64)
            right = np.ceil(sx).astype(np.int64)
            ly = (sy - np.floor(sy)).astype(F32)[:, None, None]
            lx = (sx - np.floor(sx)).astype(F32)[None, :, None]
            tl = img[top][:, left]
            tr = img[top][:, right]
            bl = img[bot][:, left]
            br = img[bot][:, right]
            t = tl + (tr - tl) * lx
            bm = bl + (br - bl) * lx
            val = t + (bm - t) * ly
        ok = oky[:, None, None] & okx[None, :, None]
        out[i] = np.where(ok, val, F32(0))
    return out


def pyramid_roi_align(boxes, image_shape, feature_maps, pool_shape, levels=None):
    """``PyramidROIAlign(pool_shape)([boxes, image_meta] + feature_maps)``.

    boxes [B,R,4]; ``image_shape`` = image_meta[0, 4:7] (:812); feature_maps = P2..P5, each
    [B,H_l,W_l,C] -> [B,R,ph,pw,C] in original (batch, box) order.  ``levels`` lets a test
    inject a precomputed level assignment."""
    boxes = np.asarray(boxes, dtype=F32)
    B, R, _ = boxes.shape
    C = feature_maps[0].shape[-1]
    ph, pw = pool_shape
    lv = roi_levels(boxes, image_shape) if levels is None else np.asarray(levels)
    out = np.zeros((B, R, ph, pw, C), dtype=F32)
    for i, level in enumerate(range(2, 6)):
        bi, ri = np.nonzero(lv == level)                         # tf.where, row-major
        if bi.size == 0:
            continue
        crops = crop_and_resize(feature_maps[i], boxes[bi, ri], bi, (ph, pw))
        out[bi, ri] = crops                                      # == sort by (batch, box)
    return out
