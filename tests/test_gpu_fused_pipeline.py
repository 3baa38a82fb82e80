"""GPU parity of the device-resident fused pipeline (``mvf_unproject_fuse_project``) and of the cross-kernel overlap inside K1T:
the feature split runs as its own persistent kernel UNDER the tensor-core unprojection (programmatic dependent launch, per-scene
device counters, a per-call generation token).  Results must be the SAME BITS as the plain calls (unproject_fuse, proj_grid:
model_multi.py:130-228, :401-404, :231-322) whatever else is in flight on the stream -- the failures this guards against only
showed when older kernels were still running at enqueue time and the previous call had used another batch size -- and,
through them, the oracle's values.  Also: repeated calls on the shared workspace, one-scene batches, slot-kernel modes, BN + ReLU,
workload T at 16 scenes."""
import numpy as np
import pytest

import oracle
from helpers import small_cfg, scene, to_dev, close

pytestmark = pytest.mark.gpu


def _m():
    import mulit_view_object_detection_b200 as m
    return m


def _cfg(nvox, V, samples=6, image=320):
    return small_cfg(nvox=nvox, nvox_z=nvox, samples=samples, NUM_VIEWS=V, IMAGE_SHAPE=np.array([image, image, 3]))


@pytest.mark.parametrize("B,mode,Cc,bn", [(5, "sum", 64, False), (4, "mean", 128, True), (1, "sum", 64, False),
                                          (3, "max", 64, False), (6, "sum", 24, True), (7, "sum", 256, False), (40, "sum", 64, False)])
def test_fused_equals_plain_calls_bit_exact(B, mode, Cc, bn):
    import torch
    m = _m()
    V, nvox, fh, fw, P = 3, 16, 20, 20, 10
    cfg = _cfg(nvox, V)
    feats, Rcam, Kmat = scene(cfg, B, V, fh, fw, Cc, seed=100 * B + Cc)
    d = to_dev(feats, Rcam, Kmat)
    rng = np.random.default_rng(B)
    bnp = (rng.uniform(0.5, 1.5, Cc).astype(np.float32), rng.normal(0, 0.1, Cc).astype(np.float32),
           rng.normal(0, 0.1, Cc).astype(np.float32), rng.uniform(0.5, 1.5, Cc).astype(np.float32)) if bn else None
    # reference: the CUDA-core slot kernel + plain projection (no cross-kernel overlap anywhere)
    ref_grid = m.unproject_fuse(*d, cfg, mode=mode, bn=bnp, relu_out=bn, tensor_cores=False)
    for _ in range(3):                                    # back-to-back calls share the workspace and its counters
        rays, grid = m.unproject_fuse_project(*d, cfg, P, mode=mode, bn=bnp, relu_out=bn)
    alone = m.unproject_fuse(*d, cfg, mode=mode, bn=bnp, relu_out=bn)          # K1T (or slot kernel) without the projection
    torch.cuda.synchronize()
    assert torch.equal(grid, alone)                       # the overlapped pipeline and the standalone call agree bit for bit
    assert torch.equal(rays, m.proj_grid([grid, d[1], d[2]], cfg, P))          # the projection read exactly the final grid
    close(grid.cpu().numpy(), ref_grid.cpu().numpy(), rtol=2e-5 if bn else 1e-5, atol=2e-6 if bn else 1e-6)
    if not bn and B <= 7:                                 # and the values are the oracle's
        o_grid = oracle.fuse_views(oracle.unproj_feat(feats, Rcam, Kmat, cfg), mode)
        close(grid.cpu().numpy(), o_grid)                  # rtol 1e-5, atol 1e-6 (helpers.py)
        close(rays.cpu().numpy(), oracle.proj_grid(o_grid, Rcam, Kmat, cfg, P))


def test_fused_into_caller_buffers_and_argument_checks():
    import torch
    m = _m()
    B, V, nvox, fh, fw, Cc, P = 4, 2, 16, 12, 16, 64, 8
    cfg = _cfg(nvox, V)
    feats, Rcam, Kmat = scene(cfg, B, V, fh, fw, Cc, seed=7)
    d = to_dev(feats, Rcam, Kmat)
    grid = torch.full((B, nvox, nvox, nvox, Cc), -1.0, device="cuda")
    rays = torch.full((B, 6, P, P, Cc), -1.0, device="cuda")
    r, g = m.unproject_fuse_project(*d, cfg, P, grid_out=grid, out=rays)
    assert r.data_ptr() == rays.data_ptr() and g.data_ptr() == grid.data_ptr()
    ref_r, ref_g = m.unproject_fuse_project(*d, cfg, P, tensor_cores=False)
    torch.cuda.synchronize()
    close(grid.cpu().numpy(), ref_g.cpu().numpy())
    assert torch.equal(rays, m.proj_grid([grid, d[1], d[2]], cfg, P))
    with pytest.raises(ValueError):
        m.unproject_fuse_project(*d, cfg, P, out=rays[:2])
    with pytest.raises(ValueError):
        m.unproject_fuse_project(d[0], d[1][:2], d[2], cfg, P)


def test_fused_workload_T_16_scenes():
    """The bench configuration: 16 scenes x 8 views, 64^3, 256 channels -- every scene's grid and ray slices from the overlapped
    pipeline equal the one-scene-at-a-time plain calls bit for bit (a projection that started early would read stale voxels)."""
    import torch
    m = _m()
    B, V, Cc, P = 16, 8, 256, 40
    cfg = _cfg(64, V, samples=20, image=640)
    feats, Rcam, Kmat = scene(cfg, B, V, 40, 40, Cc, seed=1000)
    d = to_dev(feats, Rcam, Kmat)
    grid = torch.zeros((B, 64, 64, 64, Cc), device="cuda")
    rays = torch.zeros((B, 20, P, P, Cc), device="cuda")
    for _ in range(2):
        grid.fill_(-3.0)                                  # a stale-read would pick this value up
        m.unproject_fuse_project(*d, cfg, P, grid_out=grid, out=rays)
    torch.cuda.synchronize()
    for b in (0, 1, 7, 15):
        one = [t[b:b + 1].contiguous() for t in d]
        g1 = m.unproject_fuse(*one, cfg, mode="sum")
        r1 = m.proj_grid([g1, one[1], one[2]], cfg, P)
        assert torch.equal(grid[b:b + 1], g1), b
        assert torch.equal(rays[b:b + 1], r1), b


def test_overlap_with_busy_stream_and_changing_batch_sizes():
    """The sequences that exposed the round-2 races: a kernel still running when the call is enqueued (no host sync), after a call
    with another batch size (other workspace layout, other tensor maps).  Every result must match the slot kernel."""
    import torch
    m = _m()
    B, V, Cc, P = 16, 8, 256, 40
    cfg = _cfg(64, V, samples=20, image=640)
    feats, Rcam, Kmat = scene(cfg, B, V, 40, 40, Cc, seed=1000)
    d = to_dev(feats, Rcam, Kmat)
    slot = m.unproject_fuse(*d, cfg, mode="sum", tensor_cores=False)
    one = [t[0:1].contiguous() for t in d]
    grid = torch.zeros((B, 64, 64, 64, Cc), device="cuda")
    rays = torch.zeros((B, 20, P, P, Cc), device="cuda")

    def bad(g):
        torch.cuda.synchronize()
        return int(((g - slot).abs() > (1e-5 * slot.abs() + 1e-6)).sum())

    for rep in range(2):
        m.unproject_fuse(*one, cfg, mode="sum"); torch.cuda.synchronize()
        grid.fill_(-3.0)                                   # still running when the next call is enqueued
        m.unproject_fuse_project(*d, cfg, P, grid_out=grid, out=rays)
        assert bad(grid) == 0
        assert torch.equal(rays, m.proj_grid([grid, d[1], d[2]], cfg, P))
        m.unproject_fuse(*one, cfg, mode="sum"); torch.cuda.synchronize()
        rays.fill_(0.0)
        m.unproject_fuse(*d, cfg, mode="sum", out=grid)
        assert bad(grid) == 0
        m.unproject_fuse_project(*d, cfg, P, grid_out=grid, out=rays)
        grid.fill_(-3.0)
        m.unproject_fuse_project(*d, cfg, P, grid_out=grid, out=rays)
        c = grid.clone()                                   # consumer enqueued right behind the call
        assert bad(c) == 0
