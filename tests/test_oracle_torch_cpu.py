"""CPU: pins oracle/torch_cpu.py -- the thing bench.py's `cpu_baseline` leg and `--impl reference` arm time, and the checker of
the full-size GPU feature tests -- to the NumPy oracle and to the golden fixtures produced by the reference's own Python
(tests/golden/fusion_a/b.npz: mrcnn/model_multi.py:130-228 unproj_feat, :402 K.sum, :231-322 proj_grid).

Bar: bit-exact (np.array_equal).  Both restatements evaluate the same individually rounded fp32 ops in the same order; torch's
CPU elementwise kernels do not contract a*b+c into an FMA, so the bits agree."""
import os

import numpy as np
import pytest
import torch

import oracle
from oracle import torch_cpu
from helpers import small_cfg, scene
from test_golden import load, cfg_from


@pytest.mark.parametrize("name", ["fusion_a", "fusion_b"])
def test_torch_cpu_matches_reference_golden(name):
    d = load(name)
    cfg = cfg_from(d)
    fused = torch_cpu.unproject_fuse(d["feats"], d["Rcam"], d["Kmat"], cfg, "sum")
    assert np.array_equal(fused.numpy(), d["summed"])                       # reference unproj_feat + K.sum(axis=1)
    rays = torch_cpu.project(fused, d["Rcam"], d["Kmat"], cfg, int(d["proj_size"]))
    assert np.array_equal(rays.numpy(), d["rays"])                          # reference proj_grid + nearest3
    rays2, fused2 = torch_cpu.unproject_fuse_project(d["feats"], d["Rcam"], d["Kmat"], cfg, int(d["proj_size"]))
    assert np.array_equal(rays2.numpy(), d["rays"]) and np.array_equal(fused2.numpy(), d["summed"])


@pytest.mark.parametrize("mode", ["sum", "mean", "max"])
@pytest.mark.parametrize("threads", [1, 4])
def test_torch_cpu_matches_numpy_oracle(mode, threads):
    old = torch.get_num_threads()
    torch.set_num_threads(threads)
    try:
        cfg = small_cfg(nvox=12, nvox_z=10, samples=6, NUM_VIEWS=3, IMAGE_SHAPE=np.array([640, 640, 3]))
        feats, Rcam, Kmat = scene(cfg, 2, 3, 40, 40, 16, seed=31)
        per_view = oracle.unproj_feat(feats, Rcam, Kmat, cfg)
        want = oracle.fuse_views(per_view, mode)
        got = torch_cpu.unproject_fuse(feats, Rcam, Kmat, cfg, mode)
        assert np.array_equal(got.numpy(), want)
        assert (want != 0).mean() > 0.2
        want_rays = oracle.proj_grid(want, Rcam, Kmat, cfg, 20)
        got_rays = torch_cpu.project(got, Rcam, Kmat, cfg, 20)
        assert np.array_equal(got_rays.numpy(), want_rays)
        assert (want_rays != 0).any()
    finally:
        torch.set_num_threads(old)


def test_torch_cpu_x_slab_is_a_slice_of_the_full_grid():
    """The bounded sample bench.py times (an x-slab of the grid) is exactly that slab of the full result."""
    cfg = small_cfg(nvox=12, nvox_z=8, samples=4, NUM_VIEWS=2, IMAGE_SHAPE=np.array([640, 640, 3]))
    feats, Rcam, Kmat = scene(cfg, 1, 2, 40, 40, 8, seed=5)
    full = torch_cpu.unproject_fuse(feats, Rcam, Kmat, cfg, "sum")
    for xb, xc in ((0, 3), (4, 5), (9, 3)):
        slab = torch_cpu.unproject_fuse(feats, Rcam, Kmat, cfg, "sum", x_slab=(xb, xc))
        assert torch.equal(slab, full[:, xb:xb + xc])


def test_torch_cpu_non_square_map_and_image():
    """configs[0] geometry (60x80 map of a 480x640 image, non-square proj_size) in miniature."""
    cfg = small_cfg(nvox=8, nvox_z=8, samples=5, NUM_VIEWS=2, IMAGE_SHAPE=np.array([480, 640, 3]))
    feats, Rcam, Kmat = scene(cfg, 1, 2, 30, 40, 8, seed=9, image_hw=(480, 640))
    want = oracle.fuse_views(oracle.unproj_feat(feats, Rcam, Kmat, cfg), "mean")
    got = torch_cpu.unproject_fuse(feats, Rcam, Kmat, cfg, "mean")
    assert np.array_equal(got.numpy(), want)
    assert np.array_equal(torch_cpu.project(got, Rcam, Kmat, cfg, (6, 8)).numpy(), oracle.proj_grid(want, Rcam, Kmat, cfg, (6, 8)))
