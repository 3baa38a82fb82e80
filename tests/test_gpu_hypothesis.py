"""GPU parity on hypothesis-drawn scenes: random shapes, ragged grids, arbitrary camera poses
(including cameras inside / behind the grid, where the reference samples behind-camera points, Appendix B quirk 1) and
random intrinsics.  Indices / masks / ray voxels / keep lists bit-exact, features and crops within 1e-5 relative."""
import numpy as np
import pytest

hypothesis = pytest.importorskip("hypothesis")
from hypothesis import given, settings, strategies as st, HealthCheck

import oracle
from helpers import small_cfg, to_dev, close

pytestmark = pytest.mark.gpu
SETTINGS = dict(max_examples=25, deadline=None, suppress_health_check=list(HealthCheck), derandomize=True)


def _m():
    import mulit_view_object_detection_b200 as m
    return m


def _random_pose(rng, wild):
    """camera->world [3,4]: a random rotation (small or arbitrary) and a translation near (or far from) the grid."""
    a = rng.normal(0, 1.5 if wild else 0.2, 3)
    th = np.linalg.norm(a) + 1e-12
    k = a / th
    K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    R = np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * (K @ K)
    t = rng.normal(0, 3.0 if wild else 0.5, 3)
    return np.concatenate([R, t[:, None]], axis=1).astype(np.float32)


@st.composite
def scenes(draw):
    seed = draw(st.integers(0, 2 ** 31 - 1))
    rng = np.random.default_rng(seed)
    B, V = draw(st.integers(1, 2)), draw(st.integers(1, 5))
    fh, fw = draw(st.integers(3, 24)), draw(st.integers(3, 24))
    C = 4 * draw(st.integers(1, 20))
    nvox, nvox_z = draw(st.integers(2, 14)), draw(st.integers(2, 18))
    wild = draw(st.booleans())
    cfg = small_cfg(nvox=nvox, nvox_z=nvox_z, NUM_VIEWS=V, samples=draw(st.integers(2, 9)), IMAGE_SHAPE=np.array([96, 96, 3]))
    feats = rng.standard_normal((B, V, fh, fw, C)).astype(np.float32)
    Rcam = np.stack([np.stack([_random_pose(rng, wild) for _ in range(V)]) for _ in range(B)])
    f = rng.uniform(40, 160)
    Kmat = np.broadcast_to(np.array([[f, 0, rng.uniform(30, 66)], [0, f * rng.uniform(0.8, 1.2), rng.uniform(30, 66)], [0, 0, 1]],
                                    np.float32), (B, 3, 3)).copy()
    return cfg, feats, Rcam, Kmat, draw(st.sampled_from(["sum", "mean", "max"])), draw(st.integers(2, 12))


@settings(**SETTINGS)
@given(scenes())
def test_random_scenes_unproject_fuse_project(case):
    m = _m()
    cfg, feats, Rcam, Kmat, mode, P = case
    d = to_dev(feats, Rcam, Kmat)
    per_view, idx, valid = m.unproj_feat(d, cfg, return_aux=True)
    o_views, o_idx, o_valid = oracle.unproj_feat(feats, Rcam, Kmat, cfg, return_aux=True)
    assert np.array_equal(idx.cpu().numpy(), o_idx)
    assert np.array_equal(valid.cpu().numpy(), o_valid)
    scale = max(1.0, float(np.abs(o_views).max()))
    close(per_view.cpu().numpy(), o_views, rtol=1e-5, atol=1e-6 * scale)
    rays, fused = m.unproject_fuse_project(*d, cfg, P, mode=mode)
    o_fused = oracle.fuse_views(o_views, mode)
    close(fused.cpu().numpy(), o_fused, rtol=1e-5, atol=2e-6 * scale)
    _, vox, pvalid = m.proj_grid([fused, d[1], d[2]], cfg, P, return_aux=True)
    o_vox, o_pvalid = oracle.project_indices(Rcam, Kmat, cfg, P)
    assert np.array_equal(vox.cpu().numpy(), o_vox)
    assert np.array_equal(pvalid.cpu().numpy().astype(bool), o_pvalid.astype(bool))
    close(rays.cpu().numpy(), oracle.proj_grid(fused.cpu().numpy(), Rcam, Kmat, cfg, P), rtol=0, atol=0)   # a pure gather


@settings(**SETTINGS)
@given(st.integers(0, 2 ** 31 - 1), st.integers(1, 300), st.sampled_from([0.3, 0.5, 0.7]), st.integers(1, 50))
def test_random_nms_keep_lists(seed, n, thr, max_out):
    m = _m()
    rng = np.random.default_rng(seed)
    y1, x1 = rng.uniform(0, 0.8, n), rng.uniform(0, 0.8, n)
    boxes = np.stack([y1, x1, y1 + rng.uniform(0.0, 0.3, n), x1 + rng.uniform(0.0, 0.3, n)], 1).astype(np.float32)
    if n > 3:
        boxes[1] = boxes[0]                      # exact duplicates and a degenerate (zero-area) box
        boxes[2, 2:] = boxes[2, :2]
    scores = rng.permutation(n).astype(np.float32) / n          # distinct
    keep, count = m.non_max_suppression(*to_dev(boxes, scores), max_out, thr)
    ref = oracle.non_max_suppression(boxes, scores, max_out, thr)
    k = keep.cpu().numpy()
    assert int(count) == ref.shape[0] and np.array_equal(k[:ref.shape[0]], ref) and np.all(k[ref.shape[0]:] == -1)


@settings(**SETTINGS)
@given(st.integers(0, 2 ** 31 - 1), st.integers(1, 40), st.sampled_from([(7, 7), (14, 14), (2, 5), (1, 1)]), st.integers(1, 6))
def test_random_roi_align(seed, R, pool, c4):
    m = _m()
    rng = np.random.default_rng(seed)
    C, B = 4 * c4, 2
    maps = [rng.standard_normal((B, s, s + (i % 2), C)).astype(np.float32) for i, s in enumerate((24, 12, 6, 3))]
    y1, x1 = rng.uniform(-0.1, 0.9, (B, R)), rng.uniform(-0.1, 0.9, (B, R))
    boxes = np.stack([y1, x1, y1 + rng.uniform(0, 0.9, (B, R)), x1 + rng.uniform(0, 0.9, (B, R))], -1).astype(np.float32)
    boxes[:, -1] = 0                                              # zero-padded proposal (level 2, quirk 13)
    meta = np.tile(np.array([[0, 96, 96, 3, 96, 96, 3, 0, 0, 96, 96, 1.0] + [1] * 3], np.float32), (B, 1))
    got, lv = m.PyramidROIAlign(pool)([to_dev(boxes)[0], meta] + to_dev(*maps), return_levels=True)
    assert np.array_equal(lv.cpu().numpy(), oracle.roi_levels(boxes, (96, 96, 3)))
    assert np.array_equal(got.cpu().numpy(), oracle.pyramid_roi_align(boxes, (96, 96, 3), maps, pool))     # bit-exact crops


@st.composite
def conv_cases(draw):
    kind = draw(st.sampled_from(["conv1", "conv3", "conv_s2", "deconv_s2"]))
    B, V = draw(st.integers(1, 2)), draw(st.integers(1, 3))
    C = 32 * draw(st.integers(1, 4))                       # 32 / 96: tf32 split, 64 / 128: fp16 split
    C2 = draw(st.sampled_from([0, 0, C]))
    Cout = 16 * draw(st.integers(1, 20))                   # not a multiple of the 256-column tile: ragged N
    dims = [draw(st.integers(1, 5)) for _ in range(3)]
    if kind == "conv_s2":
        dims = [2 * d for d in dims]
    return kind, B, V, C, C2, Cout, dims, draw(st.integers(0, 2 ** 31 - 1)), draw(st.booleans())


@settings(max_examples=30, deadline=None, suppress_health_check=list(HealthCheck), derandomize=True)
@given(conv_cases())
def test_random_tensor_core_convolutions(case):
    """mvf_conv3d_tc over random kinds / shapes / channel counts (both operand-split formats, ragged tiles, several view
    sources, the skip-concat second source) against the oracle's Conv3D / Conv3DTranspose."""
    m = _m()
    kind, B, V, C, C2, Cout, (X, Y, Z), seed, relu_in = case
    rng = np.random.default_rng(seed)
    Cin = V * C + C2
    k = 1 if kind == "conv1" else 3
    fan = (8 if kind == "deconv_s2" else k ** 3) * Cin
    wshape = (k, k, k, Cout, Cin) if kind == "deconv_s2" else (k, k, k, Cin, Cout)
    W = (rng.standard_normal(wshape) / np.sqrt(fan)).astype(np.float32)
    b = rng.normal(0, 0.1, Cout).astype(np.float32)
    x = rng.standard_normal((B, V, X, Y, Z, C)).astype(np.float32)
    x2 = rng.standard_normal((B, X, Y, Z, C2)).astype(np.float32) if C2 else None
    lkind = {"conv1": "conv", "conv3": "conv"}.get(kind, kind)
    conv = m.Conv3dTensorCore(*to_dev(W, b), lkind, V=V, C2=C2)
    got = conv(to_dev(x)[0], x2=to_dev(x2)[0] if C2 else None, relu_in=relu_in, relu_out=False).cpu().numpy()
    xin = np.transpose(x, (0, 2, 3, 4, 1, 5)).reshape(B, X, Y, Z, V * C)
    if C2:
        xin = np.concatenate([xin, x2], axis=-1)
    if relu_in:
        xin = np.maximum(xin, 0)
    if kind == "conv_s2":
        ref = oracle.conv3d_strided_same(xin, W, 2)
    elif kind == "deconv_s2":
        ref = oracle.conv3d_transpose_same(xin, W, 2)
    else:
        ref = oracle.fusion.conv3d_same(xin, W)
    ref = (ref + b).astype(np.float32)
    assert got.shape == ref.shape
    close(got, ref, rtol=1e-5, atol=5e-6)


@settings(max_examples=12, deadline=None, suppress_health_check=list(HealthCheck), derandomize=True)
@given(st.integers(0, 2 ** 31 - 1), st.sampled_from([32, 64, 96, 128]), st.integers(1, 5), st.integers(1, 5), st.integers(1, 9),
       st.sampled_from([(0, 0), (1, 0), (0, 1), (1, 1)]))
def test_random_convlstm_slab_steps(seed, C, X, Y, Z, halo):
    """mvf_convlstm_step_tc_slab over random slab shapes, halo combinations and both operand-split formats (C % 64): the slab
    step on a padded input equals the oracle's full SAME convolution on the same padded extent, restricted to the interior."""
    m = _m()
    rng = np.random.default_rng(seed)
    F = 64 * ((C + 63) // 64) if C % 64 else C              # F must be a multiple of 64; C may be 32 / 96 (tf32 split)
    lo, hi = halo
    Xin = X + lo + hi
    W = (rng.standard_normal((3, 3, 3, C + F, 4 * F)) * np.sqrt(2.0 / (27 * (C + F) + 4 * F))).astype(np.float32)
    b = rng.normal(0, 0.1, 4 * F).astype(np.float32)
    x = rng.standard_normal((1, Xin, Y, Z, C)).astype(np.float32)
    hp = np.tanh(rng.standard_normal((1, Xin, Y, Z, F))).astype(np.float32) * 0.5
    cp = rng.standard_normal((1, X, Y, Z, F)).astype(np.float32)
    cell = m.ConvLSTMTensorCore(*to_dev(W, b), 1.0)
    h, c = cell.step_slab(*to_dev(x, hp, cp), (lo, hi), relu_in=True)
    cpad = np.zeros((1, Xin, Y, Z, F), np.float32)
    cpad[:, lo:lo + X] = cp
    # the oracle pads with zeros on every side; planes next to a halo see the halo data, planes at a missing halo see zeros:
    # exactly what the slab kernel reads (TMA zero fill beyond the tensor)
    oh, oc = oracle.convlstm_cell_step(np.maximum(x, 0), cpad, hp, W, b)
    close(h.cpu().numpy()[:, lo:lo + X], oh[:, lo:lo + X], rtol=1e-5, atol=3e-6)
    close(c.cpu().numpy(), oc[:, lo:lo + X], rtol=1e-5, atol=3e-6)
