"""GPU parity: unproject / fuse / project kernels through the C-ABI vs the oracle.
Bit-exact for voxel->pixel indices, validity masks and ray voxel indices; 1e-5 relative for
features (tolerances in tests/helpers.py)."""
import numpy as np
import pytest
import torch

import oracle
from helpers import small_cfg, scene, to_dev, close, random_bn, RTOL, ATOL

pytestmark = pytest.mark.gpu


def _m():
    import mulit_view_object_detection_b200 as m
    return m


@pytest.mark.parametrize("B,V,fh,fw,C,nvox,nvox_z", [
    (1, 2, 40, 40, 32, 16, 16),
    (2, 3, 30, 40, 64, 12, 20),       # ragged: dims not multiples of the 4x4x16 brick
    (1, 5, 20, 20, 256, 8, 8),        # C=256 (2 float4 per lane), V not dividing 32
    (1, 1, 10, 10, 4, 5, 3),          # single view, single float4
    (1, 8, 40, 40, 128, 16, 16),
])
def test_unproj_feat_matches_oracle(B, V, fh, fw, C, nvox, nvox_z):
    m = _m()
    cfg = small_cfg(nvox=nvox, nvox_z=nvox_z, NUM_VIEWS=V)
    feats, Rcam, Kmat = scene(cfg, B, V, fh, fw, C, seed=B * 100 + V)
    out, idx, valid = m.unproj_feat(to_dev(feats, Rcam, Kmat), cfg, return_aux=True)
    o_out, o_idx, o_valid = oracle.unproj_feat(feats, Rcam, Kmat, cfg, return_aux=True)
    assert np.array_equal(idx.cpu().numpy(), o_idx)            # bit-exact indices
    assert np.array_equal(valid.cpu().numpy(), o_valid)        # bit-exact validity masks
    close(out.cpu().numpy(), o_out)
    assert 0.3 < (o_valid == 15).mean() < 1.0                  # the case exercises both in- and out-of-bounds taps


@pytest.mark.parametrize("mode", ["sum", "mean", "max"])
@pytest.mark.parametrize("V,C", [(3, 32), (8, 256)])
def test_fused_reduction_matches_oracle(mode, V, C):
    m = _m()
    cfg = small_cfg(NUM_VIEWS=V)
    feats, Rcam, Kmat = scene(cfg, 2, V, 40, 40, C, seed=3)
    fused = m.unproject_fuse(*to_dev(feats, Rcam, Kmat), cfg, mode=mode)
    o = oracle.fuse_views(oracle.unproj_feat(feats, Rcam, Kmat, cfg), mode)
    close(fused.cpu().numpy(), o)


def test_fused_add_bn_relu_equals_grid_reas():
    m = _m()
    cfg = small_cfg(GRID_REAS="add")
    rng = np.random.default_rng(0)
    feats, Rcam, Kmat = scene(cfg, 1, 3, 40, 40, 32, seed=4)
    bn = random_bn(rng, 32)
    d = to_dev(feats, Rcam, Kmat)
    fused = m.unproject_fuse(*d, cfg, mode="sum", bn=bn, relu_out=True)
    # drop-in path: materialise unproj_feat, then grid_reas
    per_view = m.unproj_feat(d, cfg)
    reas = m.grid_reas(per_view, "grid_reas_P4", cfg, params={"bn": bn})
    o = oracle.grid_reas(oracle.unproj_feat(feats, Rcam, Kmat, cfg), "grid_reas_P4", cfg, {"bn": bn})
    close(fused.cpu().numpy(), o, atol=1e-5)
    close(reas.cpu().numpy(), o, atol=1e-5)


@pytest.mark.parametrize("mode", ["mean", "max"])
def test_grid_reas_mean_max(mode):
    m = _m()
    cfg = small_cfg(GRID_REAS=mode)
    feats, Rcam, Kmat = scene(cfg, 1, 3, 40, 40, 16, seed=5)
    per_view = m.unproj_feat(to_dev(feats, Rcam, Kmat), cfg)
    out = m.grid_reas(per_view, "s", cfg)
    o = oracle.grid_reas(oracle.unproj_feat(feats, Rcam, Kmat, cfg), "s", cfg)
    close(out.cpu().numpy(), o)


def test_view_shard_and_x_slab_compose():
    """Rmain + view shards summed == all views; x-slabs concatenated == whole grid."""
    m = _m()
    cfg = small_cfg(NUM_VIEWS=4)
    feats, Rcam, Kmat = scene(cfg, 1, 4, 40, 40, 32, seed=6)
    df, dR, dK = to_dev(feats, Rcam, Kmat)
    full = m.unproject_fuse(df, dR, dK, cfg, mode="sum")
    a = m.unproject_fuse(df[:, :2].contiguous(), dR[:, :2].contiguous(), dK, cfg, mode="sum", Rmain=dR[:, 0].contiguous())
    b = m.unproject_fuse(df[:, 2:].contiguous(), dR[:, 2:].contiguous(), dK, cfg, mode="sum", Rmain=dR[:, 0].contiguous())
    close((a + b).cpu().numpy(), full.cpu().numpy())
    s0 = m.unproject_fuse(df, dR, dK, cfg, mode="sum", x_slab=(0, 6))
    s1 = m.unproject_fuse(df, dR, dK, cfg, mode="sum", x_slab=(6, 10))
    import torch
    assert torch.equal(torch.cat([s0, s1], dim=1), full)


@pytest.mark.parametrize("proj_size,C,nvox,samples", [(40, 32, 16, 8), ((12, 20), 64, 12, 5), (10, 256, 8, 20), (7, 4, 5, 33)])
def test_proj_grid_matches_oracle(proj_size, C, nvox, samples):
    m = _m()
    cfg = small_cfg(nvox=nvox, nvox_z=nvox, samples=samples)
    rng = np.random.default_rng(1)
    _, Rcam, Kmat = scene(cfg, 2, 2, 8, 8, 4, seed=7)
    grid = rng.standard_normal((2, nvox, nvox, nvox, C)).astype(np.float32)
    out, vox, valid = m.proj_grid(to_dev(grid, Rcam, Kmat), cfg, proj_size, return_aux=True)
    o_out, o_vox, o_valid = oracle.proj_grid(grid, Rcam, Kmat, cfg, proj_size, return_aux=True)
    assert np.array_equal(vox.cpu().numpy(), o_vox)            # bit-exact voxel indices
    assert np.array_equal(valid.cpu().numpy(), o_valid)
    assert np.array_equal(out.cpu().numpy(), o_out)            # nearest-neighbour copy: exact
    assert 0.05 < o_valid.mean() < 1.0


def test_proj_grid_other_view_and_slabs():
    m = _m()
    cfg = small_cfg()
    rng = np.random.default_rng(2)
    _, Rcam, Kmat = scene(cfg, 1, 3, 8, 8, 4, seed=8)
    grid = rng.standard_normal((1, 16, 16, 16, 8)).astype(np.float32)
    dg, dR, dK = to_dev(grid, Rcam, Kmat)
    out = m.proj_grid([dg, dR, dK], cfg, 24, view=2)
    o = oracle.proj_grid(grid, Rcam, Kmat, cfg, 24, view=2)
    assert np.array_equal(out.cpu().numpy(), o)
    # slab-local projection: each ray sample comes from exactly one slab, the others give 0
    full = m.proj_grid([dg, dR, dK], cfg, 24)
    parts = [m.proj_grid([dg[:, a:a + n].contiguous(), dR, dK], cfg, 24, x_slab=(a, n)) for a, n in ((0, 5), (5, 11))]
    import torch
    assert torch.equal(parts[0] + parts[1], full)


def test_depth_sampling_and_fused_collapse():
    m = _m()
    cfg = small_cfg(samples=6)
    rng = np.random.default_rng(3)
    _, Rcam, Kmat = scene(cfg, 2, 2, 8, 8, 4, seed=9)
    grid = np.maximum(rng.standard_normal((2, 16, 16, 16, 32)), 0).astype(np.float32)
    p = {"weight": rng.normal(0, 0.5, 6).astype(np.float32), "bias": 0.1, "bn": (1.2, -0.05, 0.3, 0.8)}
    d = to_dev(grid, Rcam, Kmat)
    rays = m.proj_grid(d, cfg, 20)
    ds = m.depth_sampling(rays, cfg, "grid_reas_depth_PG4", params=p)
    fused = m.proj_grid_depth_sampling(d, cfg, 20, "grid_reas_depth_PG4", params=p)
    o = oracle.depth_sampling(oracle.proj_grid(grid, Rcam, Kmat, cfg, 20), p["weight"], p["bias"], p["bn"])
    close(ds.cpu().numpy(), o, atol=1e-5)
    close(fused.cpu().numpy(), o, atol=1e-5)


def test_notebook_world_grid_variant():
    m = _m()
    cfg = small_cfg(nvox=12, nvox_z=10, GRID_DIST=5.0)
    feats, Rcam, Kmat = scene(cfg, 1, 2, 40, 40, 8, seed=10)
    d = to_dev(feats, Rcam, Kmat)
    grid, gpos, idx, valid = m.unproj_feat_notebook(d, cfg, return_aux=True)
    o_grid, o_gp, o_idx, o_valid = oracle.unproj_feat_notebook(feats, Rcam, Kmat, cfg, return_aux=True)
    assert np.array_equal(gpos.cpu().numpy(), o_gp)
    assert np.array_equal(idx.cpu().numpy(), o_idx)
    assert np.array_equal(valid.cpu().numpy(), o_valid)
    close(grid.cpu().numpy(), o_grid)
    fused = m.view_reduce(grid, "sum")
    rays, vox, pv = m.proj_grid([fused, gpos, d[1], d[2]], cfg, 16, return_aux=True)
    o_rays, o_vox, o_pv = oracle.proj_grid(oracle.fuse_views(o_grid, "sum"), Rcam, Kmat, cfg, 16,
                                           notebook_grid_pos=o_gp, return_aux=True)
    assert np.array_equal(vox.cpu().numpy(), o_vox)
    assert np.array_equal(pv.cpu().numpy(), o_pv)
    close(rays.cpu().numpy(), o_rays)


def test_pipeline_matches_oracle_and_host_entry():
    m = _m()
    import torch
    cfg = small_cfg(NUM_VIEWS=4)
    feats, Rcam, Kmat = scene(cfg, 2, 4, 40, 40, 64, seed=11)
    rays, fused = m.unproject_fuse_project(*to_dev(feats, Rcam, Kmat), cfg, proj_size=40, mode="max")
    o_fused = oracle.fuse_views(oracle.unproj_feat(feats, Rcam, Kmat, cfg), "max")
    o_rays = oracle.proj_grid(o_fused, Rcam, Kmat, cfg, 40)
    close(fused.cpu().numpy(), o_fused)
    close(rays.cpu().numpy(), o_rays)
    # same thing through the HOST-buffer entry point
    pipe = m.HostPipeline(cfg, 2, 4, 40, 40, 64, 40, mode="max")
    h = [torch.from_numpy(a).pin_memory() for a in (feats, Rcam, Kmat)]
    h_out = pipe.empty_output()
    pipe(h[0], h[1], h[2], h_out)
    assert torch.equal(h_out, rays.cpu())


def test_identity_pose_known_answer():
    """K = diag(8,8,1) + principal point 32, identity pose, voxel centres chosen so that
    z = 2 planes hit integer pixels: the unprojected value IS the feature value."""
    m = _m()
    from mulit_view_object_detection_b200.config import FusionConfig
    cfg = FusionConfig(nvox=8, nvox_z=4, vmin=-2.0, vmax=2.0, vmin_z=0.0, vmax_z=16.0, samples=4,
                       IMAGE_SHAPE=np.array([64, 64, 3]))
    R = np.zeros((1, 1, 3, 4), np.float32); R[0, 0, :, :3] = np.eye(3)
    K = np.array([[[8, 0, 32], [0, 8, 32], [0, 0, 1]]], np.float32)
    f = np.random.default_rng(0).random((1, 1, 64, 64, 4)).astype(np.float32)
    out, idx, valid = m.unproj_feat(to_dev(f, R, K), cfg, return_aux=True)
    out, idx, valid = out.cpu().numpy(), idx.cpu().numpy(), valid.cpu().numpy()
    gx = -1.75 + 0.5 * np.arange(8)          # centres; gz = 2, 6, 10, 14
    iz = 0                                   # z = 2: u = 4x + 32 is an integer
    for ix in range(8):
        for iy in range(8):
            u, w = int(4 * gx[ix] + 32), int(4 * gx[iy] + 32)
            assert tuple(idx[0, 0, ix, iy, iz]) == (w, u)
            assert valid[0, 0, ix, iy, iz] == 15
            assert np.array_equal(out[0, 0, ix, iy, iz], f[0, 0, w, u])


def test_errors():
    m = _m()
    import torch
    cfg = small_cfg()
    feats, Rcam, Kmat = scene(cfg, 1, 3, 40, 40, 32, seed=0)
    with pytest.raises(ValueError):          # CPU tensors: no fallback
        m.unproj_feat([torch.from_numpy(feats), torch.from_numpy(Rcam), torch.from_numpy(Kmat)], cfg)
    d = to_dev(feats, Rcam, Kmat)
    with pytest.raises(ValueError):          # C % 4 != 0
        m.unproj_feat([d[0][..., :30].contiguous(), d[1], d[2]], cfg)
    with pytest.raises(ValueError):          # pose/feature batch mismatch
        m.unproj_feat([d[0], d[1][:, :2].contiguous(), d[2]], cfg)
    bad = small_cfg(nvox=16, vsize=0.1)      # tf.range would not give nvox centres
    with pytest.raises(ValueError):
        m.unproj_feat(d, bad)


@pytest.mark.parametrize("mode,vanilla", [("add", True), ("add", False), ("ident", True), ("mean", True)])
def test_fusion_neck_matches_oracle(mode, vanilla):
    """MaskRCNN.build's fusion neck (model_multi.py:2382-2410): every level through unproj_feat -> grid_reas -> proj_grid ->
    depth_sampling; 'add' runs the fused K1 + K3b pair."""
    import mulit_view_object_detection_b200 as m
    rng = np.random.default_rng(17)
    B, V, C, S = 1, 3, 32, 5
    cfg = small_cfg(GRID_REAS=mode, NUM_VIEWS=V, nvox=12, nvox_z=12, samples=S, TOP_DOWN_PYRAMID_SIZE=C,
                    IMAGE_SHAPE=np.array([128, 128, 3]), VANILLA=vanilla)
    levels = (2, 3, 4, 5, 6)
    fmaps, Rcam, Kmat = [], None, None
    for lvl in levels:
        f, Rcam, Kmat = scene(cfg, B, V, 128 >> lvl, 128 >> lvl, C, seed=3, image_hw=(128, 128))
        fmaps.append(f)
    params = {}
    for lvl in levels:
        bn = (rng.uniform(0.8, 1.2, C).astype(np.float32), rng.normal(0, 0.05, C).astype(np.float32),
              rng.normal(0, 0.05, C).astype(np.float32), rng.uniform(0.7, 1.3, C).astype(np.float32))
        g = {"bn": bn}
        if mode == "ident":
            g.update(weight=(rng.standard_normal((V * C, C)) * 0.1).astype(np.float32), bias=rng.normal(0, 0.1, C).astype(np.float32))
        if mode == "mean":
            g = {}
        params["grid_reas_P%d" % lvl] = g
        params["grid_reas_depth_PG%d" % lvl] = {"weight": rng.normal(0.1, 0.3, S).astype(np.float32), "bias": 0.05,
                                                "bn": (1.1, 0.02, -0.01, 0.9)}
    dparams = {}
    for k, v in params.items():                       # grid_reas learnables live on the device; depth / bn entries are host scalars
        on_dev = k.startswith("grid_reas_P")
        dparams[k] = {kk: (to_dev(vv)[0] if on_dev and kk in ("weight", "bias") else vv) for kk, vv in v.items()}
    d = [to_dev(f)[0] for f in fmaps]
    dR, dK = to_dev(Rcam, Kmat)
    outs = m.fusion_neck(d, dR, dK, cfg, params=dparams)
    refs = oracle.fusion_neck(fmaps, Rcam, Kmat, cfg, params)
    prepared = m.fusion_neck(d, dR, dK, cfg, params=m.prepare_params(params))       # same bits through prepared parameters
    assert all(torch.equal(a, b) for a, b in zip(outs, prepared))
    assert len(outs) == 5
    for lvl, o, r in zip(levels, outs, refs):
        assert tuple(o.shape) == r.shape, lvl
        close(o.cpu().numpy(), r, rtol=1e-5, atol=5e-6)
    if not vanilla:
        assert float(outs[0].abs().max()) == 0.0 and float(outs[1].abs().max()) == 0.0


def test_notebook_channel_mean():
    """Notebook GRID_REAS='mean' (Notebook/projection.py:526-529,549): mean over channels, views become channels, ReLU."""
    import mulit_view_object_detection_b200 as m
    rng = np.random.default_rng(2)
    for shape in ((2, 3, 4, 5, 6, 64), (1, 5, 3, 3, 3, 4), (1, 2, 2, 2, 2, 260)):
        g = rng.standard_normal(shape).astype(np.float32)
        out = m.channel_mean(to_dev(g)[0])
        ref = oracle.channel_mean(g)
        assert tuple(out.shape) == ref.shape
        close(out.cpu().numpy(), ref, rtol=1e-5, atol=1e-6)


def test_neck_level_host_entry_matches_device_path():
    """mvf_fusion_neck_level_host: pinned features in, PG [B,P,P,C] out == unproject_fuse(+BN+ReLU) -> proj_grid_depth_sampling."""
    import mulit_view_object_detection_b200 as m
    rng = np.random.default_rng(4)
    B, V, C, S, P = 3, 4, 64, 6, 20
    cfg = small_cfg(nvox=16, nvox_z=16, samples=S, NUM_VIEWS=V)
    feats, Rcam, Kmat = scene(cfg, B, V, 24, 24, C, seed=9)
    bn = random_bn(rng, C)
    depth = {"weight": rng.normal(0.1, 0.3, S).astype(np.float32), "bias": 0.03, "bn": (1.1, 0.02, -0.01, 0.9)}
    pipe = m.HostPipeline(cfg, B, V, 24, 24, C, P, mode="sum", depth=depth, bn=bn, relu_out=True)
    h_in = [torch.from_numpy(a).pin_memory() for a in (feats, Rcam, Kmat)]
    h_out = pipe.empty_output()
    assert tuple(h_out.shape) == (B, P, P, C)
    pipe(h_in[0], h_in[1], h_in[2], h_out)
    d = to_dev(feats, Rcam, Kmat)
    fused = m.unproject_fuse(*d, cfg, mode="sum", bn=bn, relu_out=True)
    ref = m.proj_grid_depth_sampling([fused, d[1], d[2]], cfg, P, "depth", params=depth)
    assert torch.equal(h_out, ref.cpu())
    o_fused = oracle.grid_reas(oracle.unproj_feat(feats, Rcam, Kmat, cfg), "g", small_cfg(GRID_REAS="add"), {"bn": bn})
    o = oracle.depth_sampling(oracle.proj_grid(o_fused, Rcam, Kmat, cfg, P), depth["weight"], depth["bias"], depth["bn"])
    close(h_out.numpy(), o, rtol=1e-5, atol=2e-6)


def test_ident_neck_direct_operand_path():
    """GRID_REAS='ident' with 64-channel features: K1 writes the fp16 operand halves of the 1x1x1 conv directly
    (mvf_unproject_split_f16 + MVF_FLAG_PRESPLIT); checked against the oracle and against the materialised path."""
    import mulit_view_object_detection_b200 as m
    rng = np.random.default_rng(31)
    B, V, C, F = 2, 3, 64, 48
    cfg = small_cfg(GRID_REAS="ident", NUM_VIEWS=V, nvox=10, nvox_z=6, TOP_DOWN_PYRAMID_SIZE=F)
    feats, Rcam, Kmat = scene(cfg, B, V, 20, 20, C, seed=13)
    params = {"weight": (rng.standard_normal((V * C, F)) * 0.1).astype(np.float32), "bias": rng.normal(0, 0.1, F).astype(np.float32),
              "bn": random_bn(rng, F)}
    d = to_dev(feats, Rcam, Kmat)
    dparams = {"weight": to_dev(params["weight"])[0], "bias": to_dev(params["bias"])[0], "bn": params["bn"]}
    n0 = m.launch_count()
    direct = m.unproject_ident_fuse(*d, "grid_reas_P4", cfg, dparams)
    assert direct is not None and m.launch_count() - n0 <= 4          # weight prep (2), K1, GEMM: no split / amax pass
    o = oracle.grid_reas(oracle.unproj_feat(feats, Rcam, Kmat, cfg), "grid_reas_P4", cfg, params)
    close(direct.cpu().numpy(), o, rtol=1e-5, atol=5e-6)
    mat = m.grid_reas(m.unproj_feat(d, cfg), "grid_reas_P4", cfg, params=dparams)
    close(direct.cpu().numpy(), mat.cpu().numpy(), rtol=1e-5, atol=5e-6)


@pytest.mark.parametrize("nvox,nvox_z,C,V", [(7, 9, 36, 3), (16, 16, 256, 8), (5, 3, 4, 1)])
def test_kernels_write_only_their_output(nvox, nvox_z, C, V):
    """Guard bands around the outputs of K1 and K3 (ragged grids, channel counts that are not a multiple of the warp tile):
    nothing outside [out, out + numel) is written."""
    import mulit_view_object_detection_b200 as m
    cfg = small_cfg(nvox=nvox, nvox_z=nvox_z, samples=5, NUM_VIEWS=V)
    feats, Rcam, Kmat = scene(cfg, 2, V, 11, 13, C, seed=3)
    d = to_dev(feats, Rcam, Kmat)
    G, SENT = 4096, 12345.0

    def guarded(shape):
        n = int(np.prod(shape))
        buf = torch.full((n + 2 * G,), SENT, device="cuda")
        return buf, buf[G:G + n].view(shape)

    for mode, shape in (("sum", (2, nvox, nvox, nvox_z, C)), ("max", (2, nvox, nvox, nvox_z, C)), ("none", (2, V, nvox, nvox, nvox_z, C))):
        buf, out = guarded(shape)
        m.unproject_fuse(*d, cfg, mode=mode, out=out)
        torch.cuda.synchronize()
        assert bool((buf[:G] == SENT).all()) and bool((buf[-G:] == SENT).all()), mode
        assert bool((out != SENT).all())                       # and every output element was written
    fused = m.unproject_fuse(*d, cfg, mode="sum")
    buf, rays = guarded((2, 5, 9, 6, C))
    m.proj_grid([fused, d[1], d[2]], cfg, (9, 6), out=rays)
    torch.cuda.synchronize()
    assert bool((buf[:G] == SENT).all()) and bool((buf[-G:] == SENT).all()) and bool((rays != SENT).all())


def test_hot_cell_known_answer():
    """Known answer: features that are 1 on the four pixels of ONE bilinear cell and 0 elsewhere.  A voxel
    that projects into that cell samples exactly its four pixels, so its value is the weight sum 1; a voxel whose cell shares
    no pixel with it gets exactly 0; nothing exceeds 1.  Independent of the oracle's bilinear arithmetic."""
    import mulit_view_object_detection_b200 as m
    cfg = small_cfg(nvox=24, nvox_z=24, NUM_VIEWS=2)
    feats, Rcam, Kmat = scene(cfg, 1, 2, 40, 40, 8, seed=5)
    d = to_dev(feats, Rcam, Kmat)
    _, idx, valid = m.unproj_feat(d, cfg, return_aux=True)
    idx, valid = idx.cpu().numpy()[0], valid.cpu().numpy()[0]                   # [V,X,Y,Z,2] (y0,x0), [V,X,Y,Z]
    full = valid[0] == 15
    cells, counts = np.unique(idx[0][full].reshape(-1, 2), axis=0, return_counts=True)
    y0, x0 = cells[np.argmax(counts)]                                          # the most populated interior cell of view 0
    hot = np.zeros_like(feats)
    hot[0, 0, y0:y0 + 2, x0:x0 + 2, :] = 1.0
    out = m.unproj_feat([to_dev(hot)[0], d[1], d[2]], cfg).cpu().numpy()[0, 0]    # view 0 grid [X,Y,Z,C]
    in_cell = full & (idx[0][..., 0] == y0) & (idx[0][..., 1] == x0)
    assert in_cell.sum() > 0
    assert np.abs(out[in_cell] - 1.0).max() <= 4e-7
    far = (np.abs(idx[0][..., 0] - y0) > 1) | (np.abs(idx[0][..., 1] - x0) > 1)
    assert np.all(out[far] == 0.0)
    assert out.max() <= 1.0 + 4e-7 and out.min() >= 0.0


def test_constant_field_round_trip_known_answer():
    """Known answer: unproject a constant field from the main view only, then project back: a ray sample
    reads 1 where its voxel lies inside the grid AND all four taps of that voxel are inside the view-0 map, 0 outside the grid."""
    import mulit_view_object_detection_b200 as m
    cfg = small_cfg(nvox=20, nvox_z=20, samples=9, NUM_VIEWS=1)
    feats, Rcam, Kmat = scene(cfg, 1, 1, 40, 40, 4, seed=2)
    d = to_dev(np.ones_like(feats), Rcam, Kmat)
    per_view, _, valid = m.unproj_feat(d, cfg, return_aux=True)
    rays, vox, pvalid = m.proj_grid([per_view[:, 0].contiguous(), d[1], d[2]], cfg, 20, return_aux=True)
    rays, vox, pvalid, valid = rays.cpu().numpy()[0], vox.cpu().numpy()[0], pvalid.cpu().numpy()[0].astype(bool), valid.cpu().numpy()[0, 0]
    assert np.all(rays[~pvalid] == 0.0)                                        # outside the grid: zero fill
    v = vox[pvalid]
    tap_bits = valid[v[:, 0], v[:, 1], v[:, 2]]
    inside = rays[pvalid][tap_bits == 15]
    assert inside.size > 0 and np.abs(inside - 1.0).max() <= 4e-7              # inside the frustum: the constant comes back
    assert np.all(rays[pvalid][tap_bits == 0] == 0.0)                          # voxels the view does not see
