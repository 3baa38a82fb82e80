"""GPU parity at BASELINE.json's full sizes (workload T: 8 views, 64^3 voxels, 256 channels, 40x40 features; config c5's 96^3
grid; the compiled maximum of 192 voxels per axis), where the NumPy oracle is too slow for the feature tensors:
  * voxel->pixel indices, tap validity masks and ray voxel indices: still compared bit-exactly with the oracle (they do not
    depend on the channel count, so the oracle runs with 4 channels);
  * features through size-independent properties: the fused sum equals the sum of the unfused per-view grids, linearity in the
    features, a constant field unprojects to the sum of the bilinear weights (exactly 1 wherever all four taps are inside
    the map), proj_grid is a pure gather of the grid at the reported voxel indices (bit-exact), and the x-slab / scene
    decompositions tile the full result bit-exactly."""
import numpy as np
import pytest

import oracle
from helpers import small_cfg, scene, to_dev

pytestmark = pytest.mark.gpu

T = dict(V=8, C=256, nvox=64, fh=40, fw=40, P=40, S=20)


def _m():
    import mulit_view_object_detection_b200 as m
    return m


def _cfg(nvox=64, V=8):
    return small_cfg(nvox=nvox, nvox_z=nvox, samples=T["S"], NUM_VIEWS=V, IMAGE_SHAPE=np.array([640, 640, 3]))


@pytest.mark.parametrize("nvox,V", [(64, 8), (96, 8), (192, 2)])
def test_indices_and_masks_bit_exact_at_full_size(nvox, V):
    m = _m()
    cfg = _cfg(nvox, V)
    feats, Rcam, Kmat = scene(cfg, 1, V, T["fh"], T["fw"], 4, seed=nvox + V)
    _, idx, valid = m.unproj_feat(to_dev(feats, Rcam, Kmat), cfg, return_aux=True)
    _, o_idx, o_valid = oracle.unproj_feat(feats, Rcam, Kmat, cfg, return_aux=True)
    assert np.array_equal(idx.cpu().numpy(), o_idx)
    assert np.array_equal(valid.cpu().numpy(), o_valid)
    dR, dK = to_dev(Rcam, Kmat)
    import torch
    grid = torch.zeros((1, nvox, nvox, nvox, 4), device="cuda")
    _, vox, pvalid = m.proj_grid([grid, dR, dK], cfg, T["P"], return_aux=True)
    o_vox, o_pvalid = oracle.project_indices(Rcam, Kmat, cfg, T["P"])
    assert np.array_equal(vox.cpu().numpy(), o_vox)
    assert np.array_equal(pvalid.cpu().numpy().astype(bool), o_pvalid.astype(bool))


def test_workload_T_feature_properties():
    import torch
    m = _m()
    cfg = _cfg()
    feats, Rcam, Kmat = scene(cfg, 1, T["V"], T["fh"], T["fw"], T["C"], seed=1000)
    d_f, d_R, d_K = to_dev(feats, Rcam, Kmat)
    per_view, idx, valid = m.unproj_feat([d_f, d_R, d_K], cfg, return_aux=True)           # [1,8,64,64,64,256], 2.1 GB
    fused = m.unproject_fuse(d_f, d_R, d_K, cfg, mode="sum")
    # (1) fused sum == ascending-view sum of the unfused grids (the kernel accumulates FMA chains: 1e-5 relative)
    ref = per_view[:, 0].clone()
    for v in range(1, T["V"]):
        ref += per_view[:, v]
    err = (fused - ref).abs().max().item()
    assert err <= 1e-5 * ref.abs().max().item() + 1e-6, err
    # max fusion is a pure selection: bit-exact against the unfused grids
    assert torch.equal(m.unproject_fuse(d_f, d_R, d_K, cfg, mode="max"), per_view.max(dim=1).values)
    # (2) a voxel whose taps are all outside every map is exactly zero; voxels seen by no view stay zero after fusion
    unseen = (valid == 0).all(dim=1)
    assert float(fused[unseen].abs().max()) == 0.0
    # (3) linearity in the features
    g = torch.Generator(device="cuda"); g.manual_seed(1)
    f2 = torch.randn(d_f.shape, device="cuda", generator=g)
    lhs = m.unproject_fuse(2.5 * d_f + f2, d_R, d_K, cfg, mode="sum")
    rhs = 2.5 * fused + m.unproject_fuse(f2, d_R, d_K, cfg, mode="sum")
    scale = rhs.abs().max().item()
    assert (lhs - rhs).abs().max().item() <= 2e-5 * scale
    del lhs, rhs, f2, per_view, ref
    # (4) constant field: the per-view sample is the sum of the in-map bilinear weights -- 1 where all four taps are inside
    ones = torch.ones_like(d_f)
    pv1 = m.unproj_feat([ones, d_R, d_K], cfg)
    inside = valid == 15
    assert (pv1[inside] - 1.0).abs().max().item() <= 4e-7
    assert float(pv1[valid == 0].abs().max()) == 0.0
    assert pv1.max().item() <= 1.0 + 4e-7 and pv1.min().item() >= 0.0
    del pv1, ones
    # (5) proj_grid is a gather of the fused grid at the reported voxel indices: bit-exact, zero outside the grid
    rays, vox, pvalid = m.proj_grid([fused, d_R, d_K], cfg, T["P"], return_aux=True)
    vx = vox[0].long().clamp_(0, T["nvox"] - 1)
    gathered = fused[0][vx[..., 0], vx[..., 1], vx[..., 2]] * pvalid[0].unsqueeze(-1).float()
    assert torch.equal(rays[0], gathered)
    # (6) x-slabs tile the grid bit-exactly (slab ownership / reduce-scatter layouts)
    for xb, xc in ((0, 8), (24, 16), (56, 8)):
        slab = m.unproject_fuse(d_f, d_R, d_K, cfg, mode="sum", x_slab=(xb, xc))
        assert torch.equal(slab, fused[:, xb:xb + xc])
    # (7) the fused host entry (pinned host buffers) returns the same ray slices
    pipe = m.HostPipeline(cfg, 1, T["V"], T["fh"], T["fw"], T["C"], T["P"], mode="sum")
    h_in = [torch.from_numpy(a).pin_memory() for a in (feats, Rcam, Kmat)]
    h_out = pipe.empty_output()
    pipe(h_in[0], h_in[1], h_in[2], h_out)
    assert torch.equal(h_out, rays.cpu())


def test_batched_scenes_equal_single_scene_calls():
    """16 scenes in one launch (the bench step) == 16 one-scene launches, bit for bit."""
    import torch
    m = _m()
    cfg = _cfg(32)
    feats, Rcam, Kmat = scene(cfg, 5, T["V"], T["fh"], T["fw"], T["C"], seed=77)
    d = to_dev(feats, Rcam, Kmat)
    rays, fused = m.unproject_fuse_project(*d, cfg, T["P"], mode="sum")
    for b in (0, 3, 4):
        r1, f1 = m.unproject_fuse_project(d[0][b:b + 1], d[1][b:b + 1], d[2][b:b + 1], cfg, T["P"], mode="sum")
        assert torch.equal(f1[0], fused[b]) and torch.equal(r1[0], rays[b])


def test_config_c1_against_the_oracle():
    """BASELINE.json configs[0]: 2-view scene, 256-ch 60x80 P4 features, 32^3 grid, fp32 unproject -> mean-fuse -> project, at
    its full size against the NumPy oracle (the reference's CPU-runnable case): indices / masks / ray voxels bit-exact,
    features and ray slices within 1e-5.  The non-square map takes the (60, 80) generalisation of proj_grid."""
    m = _m()
    cfg = small_cfg(nvox=32, nvox_z=32, samples=20, NUM_VIEWS=2, IMAGE_SHAPE=np.array([480, 640, 3]))
    feats, Rcam, Kmat = scene(cfg, 1, 2, 60, 80, 256, seed=1001, image_hw=(480, 640))
    d = to_dev(feats, Rcam, Kmat)
    per_view, idx, valid = m.unproj_feat(d, cfg, return_aux=True)
    o_views, o_idx, o_valid = oracle.unproj_feat(feats, Rcam, Kmat, cfg, return_aux=True)
    assert np.array_equal(idx.cpu().numpy(), o_idx) and np.array_equal(valid.cpu().numpy(), o_valid)
    np.testing.assert_allclose(per_view.cpu().numpy(), o_views, rtol=1e-5, atol=1e-6)
    rays, fused = m.unproject_fuse_project(*d, cfg, (60, 80), mode="mean")
    o_fused = oracle.fuse_views(o_views, "mean")
    np.testing.assert_allclose(fused.cpu().numpy(), o_fused, rtol=1e-5, atol=1e-6)
    _, vox, pvalid = m.proj_grid([fused, d[1], d[2]], cfg, (60, 80), return_aux=True)
    o_vox, o_pvalid = oracle.project_indices(Rcam, Kmat, cfg, (60, 80))
    assert np.array_equal(vox.cpu().numpy(), o_vox) and np.array_equal(pvalid.cpu().numpy().astype(bool), o_pvalid.astype(bool))
    o_rays = oracle.proj_grid(o_fused, Rcam, Kmat, cfg, (60, 80))
    assert rays.shape == o_rays.shape == (1, 20, 60, 80, 256)
    np.testing.assert_allclose(rays.cpu().numpy(), o_rays, rtol=1e-5, atol=1e-6)
    assert float(np.abs(o_rays).max()) > 0 and (o_valid != 0).mean() > 0.3


def test_config_c2_fusion_against_the_oracle():
    """BASELINE.json configs[1], fusion part at the P4 level: 4 views, 256-ch 40x40 features (640x640 padded input), 48^3 grid,
    max-fuse + projection, full size against the oracle (max is a selection: fused grid within 1e-5 of the oracle's, rays a
    bit-exact gather of our own grid); the ROIAlign / NMS halves of c2 run at 1000 / 6000 boxes in test_gpu_heads.py."""
    import torch
    m = _m()
    cfg = small_cfg(nvox=48, nvox_z=48, samples=20, NUM_VIEWS=4, IMAGE_SHAPE=np.array([640, 640, 3]))
    feats, Rcam, Kmat = scene(cfg, 1, 4, 40, 40, 256, seed=2001)
    d = to_dev(feats, Rcam, Kmat)
    rays, fused = m.unproject_fuse_project(*d, cfg, 40, mode="max")
    o_views, o_idx, o_valid = oracle.unproj_feat(feats, Rcam, Kmat, cfg, return_aux=True)
    o_fused = oracle.fuse_views(o_views, "max")
    np.testing.assert_allclose(fused.cpu().numpy(), o_fused, rtol=1e-5, atol=1e-6)
    _, idx, valid = m.unproj_feat(d, cfg, return_aux=True)
    assert np.array_equal(idx.cpu().numpy(), o_idx) and np.array_equal(valid.cpu().numpy(), o_valid)
    o_rays = oracle.proj_grid(fused.cpu().numpy(), Rcam, Kmat, cfg, 40)
    assert np.array_equal(rays.cpu().numpy(), o_rays)


# ---- feature VALUES at full size against the CPU oracle -------------------------------------------------------------------------
# oracle/torch_cpu.py is pinned bit for bit to the NumPy oracle and to the reference-generated golden fixtures by
# tests/test_oracle_torch_cpu.py; it does a quarter slab of workload T in well under a second per mode.
# Tolerance: north_star's 1e-5 relative for fused features (K1 sums FMA chains, the reference sums four rounded products, then the
# views); atol 1e-6 covers elements that cancel to ~0.  Ray slices are a pure gather of the fused grid, so the same bar applies.
@pytest.mark.parametrize("mode", ["sum", "mean", "max"])
def test_workload_T_features_against_the_oracle(mode):
    import torch
    from oracle import torch_cpu
    m = _m()
    cfg = _cfg()
    feats, Rcam, Kmat = scene(cfg, 1, T["V"], T["fh"], T["fw"], T["C"], seed=1000)
    d = to_dev(feats, Rcam, Kmat)
    rays, fused = m.unproject_fuse_project(*d, cfg, T["P"], mode=mode)
    fused_h = fused.cpu()
    o_fused = torch.empty_like(fused_h)
    for xb in range(0, T["nvox"], 16):                                        # four quarter slabs = the full scene
        o_fused[:, xb:xb + 16] = torch_cpu.unproject_fuse(feats, Rcam, Kmat, cfg, mode, x_slab=(xb, 16))
    np.testing.assert_allclose(fused_h.numpy(), o_fused.numpy(), rtol=1e-5, atol=1e-6)
    assert (o_fused != 0).float().mean().item() > 0.3
    o_rays = torch_cpu.project(o_fused, Rcam, Kmat, cfg, T["P"])
    np.testing.assert_allclose(rays.cpu().numpy(), o_rays.numpy(), rtol=1e-5, atol=1e-6)
    assert (o_rays != 0).any()


def test_c5_grid_slab_against_the_oracle():
    """Config c5's grid (96^3, 8 views, 256 channels): one 12-plane x-slab of the fused grid -- the unit the reduce-scatter /
    slab-owner layouts hand to each rank -- and the ray slices that slab contributes, against the CPU oracle."""
    import torch
    from oracle import torch_cpu
    m = _m()
    cfg = _cfg(96)
    feats, Rcam, Kmat = scene(cfg, 1, T["V"], T["fh"], T["fw"], T["C"], seed=5000)
    d = to_dev(feats, Rcam, Kmat)
    xb, xc = 36, 12
    slab = m.unproject_fuse(*d, cfg, mode="sum", x_slab=(xb, xc))
    o_slab = torch_cpu.unproject_fuse(feats, Rcam, Kmat, cfg, "sum", x_slab=(xb, xc))
    np.testing.assert_allclose(slab.cpu().numpy(), o_slab.numpy(), rtol=1e-5, atol=1e-6)
    assert (o_slab != 0).float().mean().item() > 0.3
    # the slab's share of the ray slices: project a grid that is zero outside the slab
    rays = m.proj_grid([slab, d[1], d[2]], cfg, T["P"], x_slab=(xb, xc))
    o_full = torch.zeros((1, 96, 96, 96, T["C"]))
    o_full[:, xb:xb + xc] = o_slab
    o_rays = torch_cpu.project(o_full, Rcam, Kmat, cfg, T["P"])
    np.testing.assert_allclose(rays.cpu().numpy(), o_rays.numpy(), rtol=1e-5, atol=1e-6)
    assert (o_rays != 0).any()
