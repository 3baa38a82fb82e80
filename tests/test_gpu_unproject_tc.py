"""GPU parity of K1T -- unproj_feat + grid_reas 'add' evaluated on the tensor cores (mvf_unproject_fuse_tc, csrc/unproject_tc.cu) --
against the oracle restatement of mrcnn/model_multi.py:130-228 / :401-404 and against the CUDA-core slot kernel.

Tolerance: north_star's 1e-5 relative for fused features, atol 1e-6 (values that cancel to ~0).  K1T splits features and bilinear
weights into two fp16 halves each (22 mantissa bits) and accumulates the three cross products in fp32; measured error on workload T:
max |err| / (1e-5 |ref| + 1e-6) = 0.2.  Voxel->pixel indices and validity masks are not produced by K1T (side outputs stay on the slot
kernel, where they are bit-exact); the coordinate arithmetic is the same code."""
import numpy as np
import pytest

import oracle
from helpers import small_cfg, scene, to_dev, close, random_bn

pytestmark = pytest.mark.gpu


def _m():
    import mulit_view_object_detection_b200 as m
    return m


@pytest.mark.parametrize("nvox,nvox_z,V,C,B,mode", [
    (16, 16, 3, 64, 1, "sum"),          # odd view count: the second compute half idles for the last view
    (12, 10, 2, 128, 2, "mean"),        # grid not a multiple of the 4x4x8 tile: clipped TMA stores, two scenes
    (8, 24, 1, 256, 1, "sum"),          # single view
    (20, 8, 5, 192, 1, "mean"),         # C = 192: three 64-channel blocks
    (24, 24, 4, 256, 1, "sum"),
])
def test_k1t_matches_oracle(nvox, nvox_z, V, C, B, mode):
    m = _m()
    cfg = small_cfg(nvox=nvox, nvox_z=nvox_z, NUM_VIEWS=V, IMAGE_SHAPE=np.array([640, 640, 3]))
    feats, Rcam, Kmat = scene(cfg, B, V, 40, 40, C, seed=nvox + V + C)
    got = m.unproject_fuse(*to_dev(feats, Rcam, Kmat), cfg, mode=mode, tensor_cores=True)
    want = oracle.fuse_views(oracle.unproj_feat(feats, Rcam, Kmat, cfg), mode)
    assert got.shape == want.shape
    close(got.cpu().numpy(), want)
    assert (want != 0).mean() > 0.2


def test_k1t_bn_relu_epilogue_and_non_square_map():
    """grid_reas 'add' = ReLU(BN(sum_v)) (model_multi.py:401-404) fused into the epilogue; 30x40 map of a 480x640 image."""
    m = _m()
    rng = np.random.default_rng(3)
    C = 128
    cfg = small_cfg(nvox=16, nvox_z=16, NUM_VIEWS=3, IMAGE_SHAPE=np.array([480, 640, 3]))
    feats, Rcam, Kmat = scene(cfg, 1, 3, 30, 40, C, seed=11, image_hw=(480, 640))
    bn = random_bn(rng, C)
    got = m.unproject_fuse(*to_dev(feats, Rcam, Kmat), cfg, mode="sum", bn=bn, relu_out=True, tensor_cores=True)
    scale, shift = oracle.batch_norm_affine(*bn)
    want = np.maximum(oracle.fuse_views(oracle.unproj_feat(feats, Rcam, Kmat, cfg), "sum") * scale + shift, 0).astype(np.float32)
    close(got.cpu().numpy(), want, rtol=2e-5, atol=2e-6)          # one more rounded multiply-add than the bare sum
    ref = m.unproject_fuse(*to_dev(feats, Rcam, Kmat), cfg, mode="sum", bn=bn, relu_out=True, tensor_cores=False)
    close(got.cpu().numpy(), ref.cpu().numpy(), rtol=2e-5, atol=2e-6)


@pytest.mark.parametrize("scale", [1e-3, 1.0, 3e3])
def test_k1t_feature_magnitudes(scale):
    """The fp16 operand split is scaled by a power of two taken from max|features| per scene: parity must hold across magnitudes,
    and inside one tensor down to small elements (the synthetic features span ~4 decades below their maximum)."""
    m = _m()
    cfg = small_cfg(nvox=16, nvox_z=16, NUM_VIEWS=2, IMAGE_SHAPE=np.array([640, 640, 3]))
    feats, Rcam, Kmat = scene(cfg, 1, 2, 40, 40, 64, seed=5)
    feats = (feats * np.float32(scale)).astype(np.float32)
    got = m.unproject_fuse(*to_dev(feats, Rcam, Kmat), cfg, mode="sum", tensor_cores=True)
    want = oracle.fuse_views(oracle.unproj_feat(feats, Rcam, Kmat, cfg), "sum")
    close(got.cpu().numpy(), want, rtol=1e-5, atol=1e-6 * scale)


def test_k1t_zero_features_and_invisible_tiles():
    import torch
    m = _m()
    cfg = small_cfg(nvox=16, nvox_z=16, NUM_VIEWS=2, IMAGE_SHAPE=np.array([640, 640, 3]))
    feats, Rcam, Kmat = scene(cfg, 1, 2, 40, 40, 64, seed=6)
    d = to_dev(np.zeros_like(feats), Rcam, Kmat)
    assert float(m.unproject_fuse(*d, cfg, mode="sum", tensor_cores=True).abs().max()) == 0.0
    # cameras looking away from the grid: every tile is empty -> exact zeros, written by the epilogue's empty-tile path
    Rback = Rcam.copy()
    Rback[:, :, :, 2] *= -1.0
    Rback[:, :, :, 0] *= -1.0
    got = m.unproject_fuse(*to_dev(feats, Rback, Kmat), cfg, mode="sum", tensor_cores=True)
    ref = m.unproject_fuse(*to_dev(feats, Rback, Kmat), cfg, mode="sum", tensor_cores=False)
    close(got.cpu().numpy(), ref.cpu().numpy())
    assert torch.isfinite(got).all()


def test_k1t_workload_T_against_slot_kernel_and_dispatch():
    """Workload T (8 views, 64^3, 256 channels): K1T vs the slot kernel within the feature tolerance; the automatic dispatch takes
    K1T (3 launches: amax, split, MMA kernel) for this configuration and the slot kernel (1 launch) for max fusion."""
    import torch
    m = _m()
    cfg = small_cfg(nvox=64, nvox_z=64, samples=20, NUM_VIEWS=8, IMAGE_SHAPE=np.array([640, 640, 3]))
    feats, Rcam, Kmat = scene(cfg, 1, 8, 40, 40, 256, seed=1000)
    d = to_dev(feats, Rcam, Kmat)
    n0 = m.launch_count()
    auto = m.unproject_fuse(*d, cfg, mode="sum")
    assert m.launch_count() - n0 == 2          # split kernel + tensor-core kernel
    tc = m.unproject_fuse(*d, cfg, mode="sum", tensor_cores=True)
    assert torch.equal(auto, tc)                                        # deterministic: same bits run to run
    slot = m.unproject_fuse(*d, cfg, mode="sum", tensor_cores=False)
    err = ((tc - slot).abs() / (1e-5 * slot.abs() + 1e-6)).max().item()
    assert err <= 1.0, err
    n0 = m.launch_count()
    m.unproject_fuse(*d, cfg, mode="max")
    assert m.launch_count() - n0 == 1


def test_k1t_batch_and_slab_invariance():
    """A scene's result does not depend on its batch (per-scene operand scale), and x-slabs aligned with the 4-voxel tiles
    reproduce the full grid bit for bit; a slab that cuts tiles is routed to the slot kernel by the automatic dispatch."""
    import torch
    m = _m()
    cfg = small_cfg(nvox=32, nvox_z=32, NUM_VIEWS=4, IMAGE_SHAPE=np.array([640, 640, 3]))
    feats, Rcam, Kmat = scene(cfg, 3, 4, 40, 40, 128, seed=77)
    feats[1] *= 37.0                                                     # scenes with different magnitudes
    d = to_dev(feats, Rcam, Kmat)
    full = m.unproject_fuse(*d, cfg, mode="sum", tensor_cores=True)
    for b in range(3):
        one = m.unproject_fuse(d[0][b:b + 1], d[1][b:b + 1], d[2][b:b + 1], cfg, mode="sum", tensor_cores=True)
        assert torch.equal(one[0], full[b])
    for xb, xc in ((0, 8), (8, 16), (24, 8)):
        slab = m.unproject_fuse(*d, cfg, mode="sum", x_slab=(xb, xc))
        assert torch.equal(slab, full[:, xb:xb + xc])
    n0 = m.launch_count()
    halo = m.unproject_fuse(*d, cfg, mode="sum", x_slab=(7, 10))
    assert m.launch_count() - n0 == 1                                    # slot kernel
    close(halo.cpu().numpy(), full[:, 7:17].cpu().numpy())


def test_k1t_rejects_unsupported_configurations():
    m = _m()
    cfg = small_cfg(nvox=8, nvox_z=8, NUM_VIEWS=2)
    feats, Rcam, Kmat = scene(cfg, 1, 2, 40, 40, 64, seed=1)
    d = to_dev(feats, Rcam, Kmat)
    with pytest.raises(ValueError):
        m.unproject_fuse(*d, cfg, mode="max", tensor_cores=True)
    with pytest.raises(ValueError):
        m.unproject_fuse(*d, cfg, mode="sum", relu_in=True, tensor_cores=True)
    f32, R, K = scene(cfg, 1, 2, 40, 40, 32, seed=1)
    with pytest.raises(ValueError):
        m.unproject_fuse(*to_dev(f32, R, K), cfg, mode="sum", tensor_cores=True)
