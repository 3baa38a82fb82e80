"""Seeded synthetic InteriorNet-shaped inputs for the fusion hot path.

Host-side NumPy only; used by ``bench.py``, ``__graft_entry__.smoke()`` and the tests so that
the kernels and the oracle always see the same bits.  There is no dataset access: shapes and
camera intrinsics follow samples/interior/interior_multi.py:150-156 (f=600, c=(320,320) for the
640x640 padded input) and the voxel box of :379-386.
"""
import numpy as np

F32 = np.float32


def _normalize(v):
    return v / np.linalg.norm(v)


def look_at_rotation(camera, look_at, up_point):
    """Camera->world rotation whose columns are the camera axes in world coordinates, built
    from an (eye, look-at, up-point) triple the way InteriorNet poses are interpreted
    (cf. mrcnn/utils.py:1210-1218)."""
    z = _normalize(look_at - camera)
    x = _normalize(np.cross(z, up_point - camera))
    y = -_normalize(np.cross(x, z))
    return np.stack([x, y, z], axis=1)


def small_rotation(rng, sigma):
    """Rotation matrix from a near-identity unit quaternion."""
    q = np.array([1.0, *(rng.normal(0.0, sigma, 3))])
    w, x, y, z = q / np.linalg.norm(q)
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


def intrinsics(image_hw=(640, 640)):
    """K of the InteriorNet loader: f=600, principal point (320, H/2)."""
    h, w = image_hw
    return np.array([[600.0, 0.0, w / 2.0], [0.0, 600.0, h / 2.0], [0.0, 0.0, 1.0]], dtype=F32)


def make_poses(rng, V, cfg, radius=1.5, jitter=0.05):
    """[V,3,4] camera->world poses: view 0 near the world origin, views 1..V-1 on an arc of
    ``radius`` metres around it looking at the grid centre; |depth| >= 1e-3 for every voxel
    corner region is enforced by re-drawing the jitter."""
    zc = 0.5 * (cfg.vmin_z + cfg.vmax_z)
    for _ in range(64):
        R0 = small_rotation(rng, 0.05)
        t0 = rng.normal(0.0, 0.5, 3)
        centre = R0 @ np.array([0.0, 0.0, zc]) + t0
        up_dir = R0 @ np.array([0.0, -1.0, 0.0])
        poses = [np.concatenate([R0, t0[:, None]], axis=1)]
        for v in range(1, V):
            th = np.pi * (v - 0.5 * V) / max(V, 2)
            off = np.array([radius * np.sin(th), 0.3 * radius * np.cos(2.0 * th), 0.2 * radius * (np.cos(th) - 1.0)])
            cam = t0 + R0 @ off + rng.normal(0.0, jitter, 3)
            Rv = look_at_rotation(cam, centre + rng.normal(0.0, jitter, 3), cam + up_dir)
            poses.append(np.concatenate([Rv, cam[:, None]], axis=1))
        poses = np.stack(poses)
        if _min_abs_depth(poses, cfg) >= 1e-3:
            return poses.astype(F32)
    raise RuntimeError("could not draw poses with |depth| >= 1e-3")


def _min_abs_depth(poses, cfg):
    gx = cfg.vmin + cfg.vsize * (np.arange(cfg.nvox) + 0.5)
    gz = cfg.vmin_z + cfg.vsize_z * (np.arange(cfg.nvox_z) + 0.5)
    X, Y, Z = np.meshgrid(gx, gx, gz, indexing="ij")
    pts = np.stack([X.ravel(), Y.ravel(), Z.ravel()])
    world = poses[0, :, :3] @ pts + poses[0, :, 3:4]
    m = np.inf
    for P in poses:
        depth = (P[:, :3].T @ (world - P[:, 3:4]))[2]
        m = min(m, np.abs(depth).min())
    return m


def make_scene(cfg, B, V, fh, fw, C, seed, image_hw=None):
    """feats [B,V,fh,fw,C] (N(0,1) then ReLU: FPN maps are post-ReLU,
    mrcnn/model_multi.py:630-640), Rcam [B,V,3,4], Kmat [B,3,3]."""
    rng = np.random.default_rng(seed)
    feats = np.maximum(rng.standard_normal((B, V, fh, fw, C), dtype=F32), F32(0))
    Rcam = np.stack([make_poses(rng, V, cfg) for _ in range(B)])
    hw = image_hw if image_hw is not None else tuple(int(s) for s in cfg.IMAGE_SHAPE[:2])
    Kmat = np.broadcast_to(intrinsics(hw), (B, 3, 3)).copy()
    return feats, Rcam, Kmat


def make_rois(rng, B, R, pad_frac=0.05):
    """[B,R,4] normalised (y1,x1,y2,x2): log-uniform side 0.02-0.6, ``pad_frac`` zero rows at
    the end of each scene (the reference zero-pads its proposal list)."""
    side_h = np.exp(rng.uniform(np.log(0.02), np.log(0.6), (B, R)))
    side_w = np.exp(rng.uniform(np.log(0.02), np.log(0.6), (B, R)))
    cy = rng.uniform(0.0, 1.0, (B, R))
    cx = rng.uniform(0.0, 1.0, (B, R))
    y1 = np.clip(cy - 0.5 * side_h, 0.0, 1.0)
    x1 = np.clip(cx - 0.5 * side_w, 0.0, 1.0)
    y2 = np.clip(cy + 0.5 * side_h, 0.0, 1.0)
    x2 = np.clip(cx + 0.5 * side_w, 0.0, 1.0)
    boxes = np.stack([y1, x1, y2, x2], axis=-1).astype(F32)
    npad = int(round(pad_frac * R))
    if npad:
        boxes[:, R - npad:] = 0
    return boxes


def make_detection_inputs(rng, R, K):
    """probs [R,K] = softmax(N(0,2)) with all top scores distinct, deltas [R,K,4] ~ N(0,0.5)."""
    logits = rng.normal(0.0, 2.0, (R, K))
    e = np.exp(logits - logits.max(axis=1, keepdims=True))
    probs = (e / e.sum(axis=1, keepdims=True)).astype(F32)
    top = probs.max(axis=1)
    while np.unique(top).size != top.size:          # enforce distinct scores
        probs = (probs * (1 + rng.uniform(-1e-3, 1e-3, probs.shape))).astype(F32)
        top = probs.max(axis=1)
    deltas = rng.normal(0.0, 0.5, (R, K, 4)).astype(F32)
    return probs, deltas


def make_image_meta(B, image_shape, num_classes, window=None):
    """``compose_image_meta`` layout (mrcnn/model_multi.py:3278-3300): id, original shape(3),
    image shape(3), window(4, pixels), scale, active class ids."""
    h, w = int(image_shape[0]), int(image_shape[1])
    win = window if window is not None else (0, 0, h, w)
    row = [0, h, w, 3, h, w, 3, *win, 1.0] + [1] * num_classes
    return np.tile(np.asarray(row, dtype=F32)[None], (B, 1))


def make_anchors(image_hw, scales=(32, 64, 128, 256, 512), ratios=(0.5, 1, 2),
                 strides=(4, 8, 16, 32, 64)):
    """Normalised pyramid anchors [A,4] (y1,x1,y2,x2): 3 ratios per location of each level."""
    H, W = image_hw
    out = []
    for scale, stride in zip(scales, strides):
        fh, fw = int(np.ceil(H / stride)), int(np.ceil(W / stride))
        r = np.asarray(ratios, dtype=np.float64)
        hs = scale / np.sqrt(r)
        ws = scale * np.sqrt(r)
        ys = np.arange(fh) * stride
        xs = np.arange(fw) * stride
        cy, cx, k = np.meshgrid(ys, xs, np.arange(len(ratios)), indexing="ij")
        h = hs[k]
        w = ws[k]
        boxes = np.stack([cy - 0.5 * h, cx - 0.5 * w, cy + 0.5 * h, cx + 0.5 * w], axis=-1).reshape(-1, 4)
        out.append(boxes)
    a = np.concatenate(out)
    scale_v = np.array([H - 1, W - 1, H - 1, W - 1], dtype=np.float64)
    shift = np.array([0, 0, 1, 1], dtype=np.float64)
    return ((a - shift) / scale_v).astype(F32)
