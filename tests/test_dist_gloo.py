"""CPU: the multi-GPU sharding logic (mulit_view_object_detection_b200/dist.py) at world_size 2 over
gloo.  The compute steps are injected from the oracle (no GPU here); what is under test is the
view / slab / scene partitioning, the Rmain plumbing and the collectives."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from helpers import small_cfg, scene


class OracleOps:
    """dist.py's ``ops`` interface on torch CPU tensors, backed by the NumPy oracle."""

    def unproject_fuse(self, feats, Rcam, Kmat, config, mode, Rmain=None, x_slab=None, tensor_cores=None):
        f, R, K = feats.numpy(), Rcam.numpy(), Kmat.numpy()
        if Rmain is not None:                       # oracle takes the main pose as view 0: prepend it with zero features
            R = np.concatenate([Rmain.numpy()[:, None], R], axis=1)
            f = np.concatenate([np.zeros_like(f[:, :1]), f], axis=1)
            per_view = oracle.unproj_feat(f, R, K, config)[:, 1:]
        else:
            per_view = oracle.unproj_feat(f, R, K, config)
        fused = oracle.fuse_views(per_view, mode)
        if x_slab is not None:
            fused = fused[:, x_slab[0]:x_slab[0] + x_slab[1]]
        return torch.from_numpy(np.ascontiguousarray(fused))

    def proj_grid(self, grid, Rcam, Kmat, config, proj_size, x_slab=None):
        g = grid.numpy()
        if x_slab is not None:                      # a slab only answers for its own x range
            full = np.zeros((g.shape[0], config.nvox) + g.shape[2:], np.float32)
            full[:, x_slab[0]:x_slab[0] + x_slab[1]] = g
            g = full
        return torch.from_numpy(oracle.proj_grid(g, Rcam.numpy(), Kmat.numpy(), config, proj_size))

    def proj_collapse_linear(self, grid, Rcam, Kmat, config, proj_size, weight, x_slab=None):
        rays = self.proj_grid(grid, Rcam, Kmat, config, proj_size, x_slab=x_slab).numpy()          # [B,S,P,P,C]
        w = np.asarray(weight, np.float32).reshape(1, -1, 1, 1, 1)
        return torch.from_numpy((rays * w).sum(axis=1, dtype=np.float32))

    def depth_affine_relu(self, x, depth):
        gamma, beta, mean, var = (np.float32(np.asarray(a).reshape(-1)[0]) for a in depth["bn"])
        inv = np.float32(1.0) / np.sqrt(var + np.float32(1e-3)) * gamma
        y = (x.numpy() + np.float32(depth["bias"])) * inv + (beta - mean * inv)
        return torch.from_numpy(np.maximum(y, 0).astype(np.float32))

    def scale(self, grid, factor):
        return grid * np.float32(factor)

    def convlstm_cell(self, W, bias):
        return OracleCell(W.numpy(), bias.numpy())

    def affine_relu(self, h, bn):
        scale, shift = oracle.batch_norm_affine(*bn)
        return torch.from_numpy(np.maximum(h.numpy() * scale + shift, 0).astype(np.float32))


class OracleCell:
    """ConvLSTMCell.call on a slab with halo planes: SAME convolution over the padded extent, interior kept."""

    def __init__(self, W, b):
        self.W, self.b = W, b

    def step_slab(self, x, h_prev, c_prev, halo, relu_in=False, act_amax=None):
        lo, hi = halo
        xn = x.numpy()
        B, Xin = xn.shape[:2]
        F = self.b.shape[0] // 4
        if relu_in:
            xn = np.maximum(xn, 0)
        hp = h_prev.numpy() if h_prev is not None else np.zeros(xn.shape[:4] + (F,), np.float32)
        cp = np.zeros_like(hp)
        if c_prev is not None:
            cp[:, lo:Xin - hi] = c_prev.numpy()
        h, c = oracle.convlstm_cell_step(xn, cp, hp, self.W, self.b)
        h_out = np.zeros_like(h)
        h_out[:, lo:Xin - hi] = h[:, lo:Xin - hi]            # halo planes are the neighbours' to fill
        return torch.from_numpy(h_out), torch.from_numpy(np.ascontiguousarray(c[:, lo:Xin - hi]))


DEPTH = {"weight": np.array([0.4, 0.3, 0.2, 0.1], np.float32), "bias": 0.05, "bn": (1.2, 0.1, -0.05, 0.8)}


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world_size, port, strategy, mode, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world_size))
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    try:
        from mulit_view_object_detection_b200 import dist as mvd
        cfg = small_cfg(nvox=8, nvox_z=6, samples=4, NUM_VIEWS=3)
        feats, Rcam, Kmat = scene(cfg, 2, 3, 12, 12, 8, seed=5)
        t = [torch.from_numpy(a) for a in (feats, Rcam, Kmat)]
        ops = OracleOps()
        if strategy == "scene":
            rays, mine = mvd.scene_shard(*t, cfg, 10, mode=mode, ops=ops, gather=True)
        elif strategy == "auto":
            rays = mvd.fuse_project_auto(*t, cfg, 10, mode=mode, ops=ops)                       # 2 scenes, 2 ranks: scene sharding
            one = mvd.fuse_project_auto(*[x[:1] for x in t], cfg, 10, mode=mode, ops=ops)         # 1 scene, 2 ranks: slab owner
            assert torch.allclose(one, rays[:1], rtol=1e-5, atol=1e-6)
        elif strategy == "slab_owner_depth":
            rays, _ = mvd.slab_owner(*t, cfg, 10, mode=mode, ops=ops, depth=DEPTH)
        elif strategy == "slab_owner_depth_empty":                       # more ranks than x-planes: rank 2 owns an empty slab
            cfg = small_cfg(nvox=2, nvox_z=6, samples=4, NUM_VIEWS=3)
            rays, slab = mvd.slab_owner(*t, cfg, 10, mode=mode, ops=ops, depth=DEPTH)
            assert slab.shape[1] == (1 if rank < 2 else 0)
        elif strategy == "slab_owner_scatter":
            part, _ = mvd.slab_owner(*t, cfg, 10, mode=mode, ops=ops, scatter_scenes=True)   # rank r keeps scene r
            parts = [torch.empty_like(part) for _ in range(world_size)]
            dist.all_gather(parts, part)
            rays = torch.cat(parts, dim=0)
        else:
            fn = {"allreduce": mvd.view_shard_allreduce, "reduce_scatter": mvd.view_shard_reduce_scatter,
                  "slab_owner": mvd.slab_owner}[strategy]
            rays, _ = fn(*t, cfg, 10, mode=mode, ops=ops)
        np.save(os.path.join(out_dir, "rays_%d.npy" % rank), rays.numpy())
    finally:
        dist.destroy_process_group()


def _lstm_case():
    cfg = small_cfg(nvox=6, nvox_z=4, samples=4, NUM_VIEWS=3, GRID_REAS="lstm3d", TOP_DOWN_PYRAMID_SIZE=4)
    feats, Rcam, Kmat = scene(cfg, 1, 3, 12, 12, 4, seed=8)
    rng = np.random.default_rng(3)
    W = (rng.standard_normal((3, 3, 3, 8, 16)) * 0.15).astype(np.float32)
    b = rng.normal(0, 0.1, 16).astype(np.float32)
    bn = (np.full(4, 1.1, np.float32), np.full(4, 0.02, np.float32), np.full(4, -0.01, np.float32), np.full(4, 0.9, np.float32))
    return cfg, feats, Rcam, Kmat, W, b, bn


def _lstm_worker(rank, world_size, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world_size))
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    try:
        from mulit_view_object_detection_b200 import dist as mvd
        cfg, feats, Rcam, Kmat, W, b, bn = _lstm_case()
        t = [torch.from_numpy(a) for a in (feats, Rcam, Kmat)]
        rays, slab = mvd.lstm_slab(*t, cfg, {"W": torch.from_numpy(W), "b": torch.from_numpy(b), "bn": bn}, proj_size=10,
                                   ops=OracleOps())
        np.save(os.path.join(out_dir, "rays_%d.npy" % rank), rays.numpy())
        np.save(os.path.join(out_dir, "slab_%d.npy" % rank), slab.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world_size", [2, 3])
def test_lstm_slab_halo_exchange_matches_single_process(world_size, tmp_path):
    """Recurrent fusion over x-slabs with a 1-voxel halo of h exchanged per step == the unsharded recurrence."""
    mp.spawn(_lstm_worker, args=(world_size, _free_port(), str(tmp_path)), nprocs=world_size, join=True)
    cfg, feats, Rcam, Kmat, W, b, bn = _lstm_case()
    fused = oracle.grid_reas(oracle.unproj_feat(feats, Rcam, Kmat, cfg), "grid_reas_P4", cfg, {"W": W, "b": b, "bn": bn})
    ref = oracle.proj_grid(fused, Rcam, Kmat, cfg, 10)
    slabs = [np.load(os.path.join(str(tmp_path), "slab_%d.npy" % r)) for r in range(world_size)]
    np.testing.assert_allclose(np.concatenate(slabs, axis=1), fused, rtol=1e-6, atol=1e-7)
    for r in range(world_size):
        np.testing.assert_allclose(np.load(os.path.join(str(tmp_path), "rays_%d.npy" % r)), ref, rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("strategy,mode", [("allreduce", "sum"), ("allreduce", "max"), ("allreduce", "mean"),
                                           ("reduce_scatter", "sum"), ("reduce_scatter", "max"),
                                           ("slab_owner", "sum"), ("slab_owner", "max"), ("scene", "sum"),
                                           ("slab_owner_scatter", "sum"), ("auto", "sum")])
def test_two_rank_strategies_match_single_process(strategy, mode, tmp_path):
    world_size = 2
    mp.spawn(_worker, args=(world_size, _free_port(), strategy, mode, str(tmp_path)), nprocs=world_size, join=True)
    cfg = small_cfg(nvox=8, nvox_z=6, samples=4, NUM_VIEWS=3)
    feats, Rcam, Kmat = scene(cfg, 2, 3, 12, 12, 8, seed=5)
    ref = oracle.proj_grid(oracle.fuse_views(oracle.unproj_feat(feats, Rcam, Kmat, cfg), mode), Rcam, Kmat, cfg, 10)
    for r in range(world_size):
        got = np.load(os.path.join(str(tmp_path), "rays_%d.npy" % r))
        # sum order differs between shardings (per-rank partial sums): 1e-5 relative; max is exact
        np.testing.assert_allclose(got, ref, rtol=1e-5, atol=1e-6)
        if mode == "max":
            assert np.array_equal(got, ref)


def test_partition_helpers():
    from mulit_view_object_detection_b200 import dist as mvd
    for V, W in ((8, 2), (8, 8), (5, 4), (3, 8)):
        cover = []
        for r in range(W):
            lo, hi = mvd.view_slice(V, r, W)
            cover += list(range(lo, hi))
        assert cover == list(range(V))
    for X, W in ((64, 8), (48, 4), (10, 4)):
        spans = [mvd.slab_bounds(X, r, W) for r in range(W)]
        assert spans[0][0] == 0 and sum(c for _, c in spans) == X
        assert all(spans[i][0] + spans[i][1] == spans[i + 1][0] for i in range(W - 1))
    assert mvd.scene_indices(7, 1, 3) == [1, 4]


@pytest.mark.parametrize("strategy,world_size,nvox", [("slab_owner_depth", 2, 8), ("slab_owner_depth", 3, 8), ("slab_owner_depth_empty", 3, 2)])
def test_slab_owner_with_depth_collapse(strategy, world_size, nvox, tmp_path):
    """dist.slab_owner(depth=...): each rank collapses its own slab's ray samples linearly, the all-reduce moves [B,P,P,C], bias +
    BatchNorm + ReLU follow the sum -- equal to depth_sampling(proj_grid(full grid)) (model_multi.py:481-487); also with an EMPTY
    slab (more ranks than x-planes)."""
    mp.spawn(_worker, args=(world_size, _free_port(), strategy, "sum", str(tmp_path)), nprocs=world_size, join=True)
    cfg = small_cfg(nvox=nvox, nvox_z=6, samples=4, NUM_VIEWS=3)
    feats, Rcam, Kmat = scene(small_cfg(nvox=8, nvox_z=6, samples=4, NUM_VIEWS=3), 2, 3, 12, 12, 8, seed=5)
    fused = oracle.fuse_views(oracle.unproj_feat(feats, Rcam, Kmat, cfg), "sum")
    ref = oracle.depth_sampling(oracle.proj_grid(fused, Rcam, Kmat, cfg, 10), DEPTH["weight"], DEPTH["bias"], DEPTH["bn"])
    for r in range(world_size):
        got = np.load(os.path.join(str(tmp_path), "rays_%d.npy" % r))
        np.testing.assert_allclose(got, ref, rtol=1e-5, atol=1e-6)
    assert (ref > 0).any()
