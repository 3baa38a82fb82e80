"""GPU parity of the tensor-core conv3d family (mvf_conv3d_tc: Conv3D stride 1 / stride 2, Conv3DTranspose stride 2) against
the oracle restatement of the Keras/TF layers the reference uses in grid_reas 'conv3d' (mrcnn/model_multi.py:406-441) and in
the depth_sampling 'conv3d' branch (:467-480).

Tolerance: the reference convolves in fp32; the 3xTF32 split carries ~2^-21 per product and the oracle accumulates in float64,
so outputs (O(1) magnitude after BN + ReLU) are compared at rtol=1e-5, atol=5e-6."""
import numpy as np
import pytest

import oracle
from helpers import to_dev, close, small_cfg

pytestmark = pytest.mark.gpu


def _m():
    import mulit_view_object_detection_b200 as m
    return m


def _bn(rng, n):
    return (rng.uniform(0.8, 1.2, n).astype(np.float32), rng.normal(0, 0.05, n).astype(np.float32),
            rng.normal(0, 0.05, n).astype(np.float32), rng.uniform(0.7, 1.3, n).astype(np.float32))


def _layer(rng, shape, fan_in, nout):
    return {"W": (rng.standard_normal(shape) * np.sqrt(1.0 / fan_in)).astype(np.float32),
            "b": rng.normal(0, 0.1, nout).astype(np.float32), "bn": _bn(rng, nout)}


def _dev_params(p):
    out = {}
    for k, v in p.items():
        out[k] = {kk: (to_dev(vv)[0] if kk in ("W", "b", "w") else vv) for kk, vv in v.items()}
    return out


@pytest.mark.parametrize("X,Y,Z,Cin,Cout", [(4, 4, 8, 32, 32), (6, 4, 10, 64, 48), (2, 8, 4, 32, 272), (4, 4, 4, 128, 256)])
def test_conv3d_stride2_matches_oracle(X, Y, Z, Cin, Cout):
    m = _m()
    rng = np.random.default_rng(X * Cout)
    p = _layer(rng, (3, 3, 3, Cin, Cout), 27 * Cin, Cout)
    x = rng.standard_normal((2, X, Y, Z, Cin)).astype(np.float32)
    conv = m.Conv3dTensorCore(*to_dev(p["W"], p["b"]), "conv_s2", bn=p["bn"])
    got = conv(to_dev(x)[0], relu_in=True)
    ref = oracle.fusion._conv_bn_relu(oracle.conv3d_strided_same(np.maximum(x, 0), p["W"], 2), p["b"], p["bn"])
    assert tuple(got.shape) == ref.shape
    close(got.cpu().numpy(), ref, rtol=1e-5, atol=5e-6)


@pytest.mark.parametrize("X,Y,Z,Cin,Cout", [(2, 2, 4, 32, 32), (3, 2, 5, 64, 48), (1, 4, 2, 32, 272), (2, 2, 2, 192, 64)])
def test_conv3d_transpose_stride2_matches_oracle(X, Y, Z, Cin, Cout):
    m = _m()
    rng = np.random.default_rng(Z * Cout)
    p = _layer(rng, (3, 3, 3, Cout, Cin), 8 * Cin, Cout)          # Keras Conv3DTranspose kernel: [k,k,k,filters,in]
    x = rng.standard_normal((2, X, Y, Z, Cin)).astype(np.float32)
    conv = m.Conv3dTensorCore(*to_dev(p["W"], p["b"]), "deconv_s2", bn=p["bn"])
    got = conv(to_dev(x)[0])
    ref = oracle.fusion._conv_bn_relu(oracle.conv3d_transpose_same(x, p["W"], 2), p["b"], p["bn"])
    assert tuple(got.shape) == ref.shape
    close(got.cpu().numpy(), ref, rtol=1e-5, atol=5e-6)


def test_conv3d_stride1_matches_oracle():
    m = _m()
    rng = np.random.default_rng(2)
    p = _layer(rng, (3, 3, 3, 32, 64), 27 * 32, 64)
    x = rng.standard_normal((1, 3, 5, 6, 32)).astype(np.float32)
    conv = m.Conv3dTensorCore(*to_dev(p["W"], p["b"]), "conv", bn=p["bn"])
    got = conv(to_dev(x)[0], relu_out=False)
    y = (oracle.fusion.conv3d_same(x, p["W"]) + p["b"]).astype(np.float32)
    scale, shift = oracle.batch_norm_affine(*p["bn"])
    close(got.cpu().numpy(), y * scale + shift, rtol=1e-5, atol=5e-6)


def test_stride2_rejects_odd_dims():
    m = _m()
    rng = np.random.default_rng(0)
    p = _layer(rng, (3, 3, 3, 32, 32), 27 * 32, 32)
    conv = m.Conv3dTensorCore(*to_dev(p["W"], p["b"]), "conv_s2")
    with pytest.raises(ValueError):
        conv(to_dev(rng.standard_normal((1, 3, 4, 4, 32)).astype(np.float32))[0])


@pytest.mark.parametrize("B,V,X,Z,C,F", [(1, 3, 8, 8, 32, 32), (2, 2, 4, 12, 32, 16), (1, 2, 8, 8, 64, 64)])
def test_grid_reas_conv3d_unet(B, V, X, Z, C, F):
    """grid_reas('conv3d') (model_multi.py:406-441): U-Net over the view-concatenated grids."""
    m = _m()
    rng = np.random.default_rng(V * 10 + Z)
    cfg = small_cfg(GRID_REAS="conv3d", NUM_VIEWS=V, nvox=X, nvox_z=Z, TOP_DOWN_PYRAMID_SIZE=F)
    grids = rng.standard_normal((B, V, X, X, Z, C)).astype(np.float32)
    params = {"conv1": _layer(rng, (3, 3, 3, V * C, 2 * F), 27 * V * C, 2 * F),
              "conv2": _layer(rng, (3, 3, 3, 2 * F, 4 * F), 27 * 2 * F, 4 * F),
              "deconv1": _layer(rng, (3, 3, 3, 2 * F, 4 * F), 8 * 4 * F, 2 * F),
              "deconv2": _layer(rng, (3, 3, 3, F, 4 * F), 8 * 4 * F, F)}
    n0 = m.launch_count()
    out = m.grid_reas(to_dev(grids)[0], "unet_%d_%d" % (V, Z), cfg, params=_dev_params(params))
    ref = oracle.grid_reas(grids, "grid_reas_P4", cfg, params)
    assert tuple(out.shape) == (B, X, X, Z, F)
    close(out.cpu().numpy(), ref, rtol=1e-5, atol=5e-6)
    assert m.launch_count() - n0 >= 4 + 4 + 2          # 4 weight preparations, 4 GEMM launches, 2 parity re-layout passes (stride 2)


def test_depth_sampling_conv3d_branch():
    """depth_sampling 'conv3d' branch (model_multi.py:467-480): depthwise 1x1 -> 1x1 conv (512) -> BN -> ReLU, twice."""
    m = _m()
    rng = np.random.default_rng(6)
    B, S, P, C, H, F = 2, 4, 6, 32, 64, 32
    cfg = small_cfg(GRID_REAS="conv3d", samples=S, TOP_DOWN_PYRAMID_SIZE=F)
    x = np.maximum(rng.standard_normal((B, S, P, P, C)), 0).astype(np.float32)
    params = {"dw1": {"w": rng.uniform(0.5, 1.5, C * S).astype(np.float32), "b": rng.normal(0, 0.1, C * S).astype(np.float32)},
              "conv1": _layer(rng, (C * S, H), C * S, H),
              "dw2": {"w": rng.uniform(0.5, 1.5, H).astype(np.float32), "b": rng.normal(0, 0.1, H).astype(np.float32)},
              "conv2": _layer(rng, (H, F), H, F)}
    out = m.depth_sampling(to_dev(x)[0], cfg, "PG4_depth", params=_dev_params(params))
    ref = oracle.depth_sampling_conv3d(x, params)
    assert tuple(out.shape) == (B, P, P, F)
    close(out.cpu().numpy(), ref, rtol=1e-5, atol=5e-6)


def test_fp16_split_keeps_fp32_accuracy_over_a_wide_dynamic_range():
    """The fp16 operand split scales each tensor by a power of two taken from its max; elements down to ~4e-7 of the max
    are still represented to 1e-5 relative (a2 goes subnormal with 2^-24 spacing in the scaled domain).  Inputs spanning
    six decades: the output must stay within 1e-5 of its own scale."""
    m = _m()
    rng = np.random.default_rng(12)
    Cin, Cout = 64, 64                                   # multiples of 64: the f16-rate path
    p = _layer(rng, (3, 3, 3, Cin, Cout), 27 * Cin, Cout)
    mag = 10.0 ** rng.uniform(-3, 3, (1, 4, 4, 8, Cin))
    x = (mag * rng.choice([-1.0, 1.0], mag.shape)).astype(np.float32)
    conv = m.Conv3dTensorCore(*to_dev(p["W"], p["b"]), "conv")
    got = conv(to_dev(x)[0], relu_out=False).cpu().numpy()
    ref = (oracle.fusion.conv3d_same(x, p["W"]) + p["b"]).astype(np.float32)
    scale = float(np.abs(ref).max())
    assert np.abs(got - ref).max() <= 1e-5 * scale
    # and a tensor of small values only (max 1e-3) is as accurate as one of O(1) values: the scale follows the data
    xs = (x / np.abs(x).max() * 1e-3).astype(np.float32)
    conv0 = m.Conv3dTensorCore(to_dev(p["W"])[0], to_dev(np.zeros(Cout, np.float32))[0], "conv")       # no bias: outputs ~1e-4
    got_s = conv0(to_dev(xs)[0], relu_out=False).cpu().numpy()
    ref_s = oracle.fusion.conv3d_same(xs, p["W"]).astype(np.float32)
    assert np.abs(got_s - ref_s).max() <= 1e-5 * float(np.abs(ref_s).max())


def test_fusion_neck_conv3d_mode():
    """The shipped configuration (GRID_REAS='conv3d', interior_multi.py:391,419) through fusion_neck: unproj_feat -> U-Net ->
    proj_grid -> depth_sampling conv3d branch, fp16-rate convolutions with the operand scale bounded by max|features|."""
    m = _m()
    from helpers import scene
    rng = np.random.default_rng(21)
    B, V, C, F, S = 1, 2, 64, 64, 4
    cfg = small_cfg(GRID_REAS="conv3d", NUM_VIEWS=V, nvox=8, nvox_z=8, samples=S, TOP_DOWN_PYRAMID_SIZE=F,
                    IMAGE_SHAPE=np.array([128, 128, 3]), VANILLA=True)
    levels = (4, 5)
    fmaps = []
    for lvl in levels:
        f, Rcam, Kmat = scene(cfg, B, V, 128 >> lvl, 128 >> lvl, C, seed=6, image_hw=(128, 128))
        fmaps.append(f)
    params = {}
    for lvl in levels:
        params["grid_reas_P%d" % lvl] = {"conv1": _layer(rng, (3, 3, 3, V * C, 2 * F), 27 * V * C, 2 * F),
                                         "conv2": _layer(rng, (3, 3, 3, 2 * F, 4 * F), 27 * 2 * F, 4 * F),
                                         "deconv1": _layer(rng, (3, 3, 3, 2 * F, 4 * F), 8 * 4 * F, 2 * F),
                                         "deconv2": _layer(rng, (3, 3, 3, F, 4 * F), 8 * 4 * F, F)}
        params["grid_reas_depth_PG%d" % lvl] = {
            "dw1": {"w": rng.uniform(0.5, 1.5, F * S).astype(np.float32), "b": rng.normal(0, 0.1, F * S).astype(np.float32)},
            "conv1": _layer(rng, (F * S, 64), F * S, 64),
            "dw2": {"w": rng.uniform(0.5, 1.5, 64).astype(np.float32), "b": rng.normal(0, 0.1, 64).astype(np.float32)},
            "conv2": _layer(rng, (64, F), 64, F)}
    prepared = m.prepare_params(params)
    dfm, (dR, dK) = to_dev(*fmaps), to_dev(Rcam, Kmat)
    outs = m.fusion_neck(dfm, dR, dK, cfg, params=prepared, levels=levels)
    refs = oracle.fusion_neck(fmaps, Rcam, Kmat, cfg, params, levels=levels)
    for o, r in zip(outs, refs):
        assert tuple(o.shape) == r.shape
        close(o.cpu().numpy(), r, rtol=1e-5, atol=5e-6)
    # the neck took the direct path (K1 writes the operand halves of the first conv); it must equal the materialised path
    # (fp32 per-view grids + split pass with the same scale bound) bit for bit
    import torch
    direct = m.unproject_unet_fuse(dfm[0], dR, dK, "grid_reas_P4", cfg, prepared["grid_reas_P4"])
    assert direct is not None
    bound = dfm[0].abs().amax().reshape(1)
    mat = m.grid_reas(m.unproj_feat([dfm[0], dR, dK], cfg), "grid_reas_P4", cfg, params=prepared["grid_reas_P4"], act_amax=bound)
    assert torch.equal(direct, mat)
