"""GPU (needs >= 2 devices): the cooperative strategies of dist.py over NCCL vs the single-GPU
pipeline.  Skipped on a 1-GPU box; the gloo twin in test_dist_gloo.py always runs on CPU."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world_size, port, out_dir):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world_size))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world_size, device_id=torch.device("cuda", rank))
    try:
        import mulit_view_object_detection_b200 as m
        from mulit_view_object_detection_b200 import dist as mvd
        from helpers import small_cfg, scene
        cfg = small_cfg(nvox=16, nvox_z=16, samples=6, NUM_VIEWS=4)
        feats, Rcam, Kmat = scene(cfg, 2, 4, 40, 40, 64, seed=21)
        d = [torch.from_numpy(a).cuda() for a in (feats, Rcam, Kmat)]
        ref, _ = m.unproject_fuse_project(*d, cfg, 24, mode="sum")
        res = {}
        for name, fn in (("allreduce", mvd.view_shard_allreduce), ("reduce_scatter", mvd.view_shard_reduce_scatter),
                         ("slab_owner", mvd.slab_owner)):
            rays, _ = fn(*d, cfg, 24, mode="sum")
            res[name] = float((rays - ref).abs().max() / ref.abs().max())
        # slab owner with the depth collapse: linear collapse per slab, all-reduce of PG, bias / BN / ReLU after the sum
        depth = {"weight": torch.linspace(0.3, 0.05, 6).cuda(), "bias": 0.02, "bn": (1.1, 0.05, -0.02, 0.9)}
        fused_ref = m.unproject_fuse(*d, cfg, mode="sum")
        pg_ref = m.proj_grid_depth_sampling([fused_ref, d[1], d[2]], cfg, 24, "depth", params=depth)
        pg, _ = mvd.slab_owner(*d, cfg, 24, mode="sum", depth=depth)
        res["slab_owner_depth"] = float((pg - pg_ref).abs().max() / pg_ref.abs().max())
        rays, _ = mvd.scene_shard(*d, cfg, 24, mode="sum", gather=True)
        res["scene"] = float((rays - ref).abs().max())
        refm, _ = m.unproject_fuse_project(*d, cfg, 24, mode="max")
        raysm, _ = mvd.view_shard_reduce_scatter(*d, cfg, 24, mode="max")
        res["max_exact"] = float((raysm - refm).abs().max())
        # recurrent fusion over x-slabs with halo exchange vs the single-GPU grid_reas('lstm3d') + proj_grid
        rng = np.random.default_rng(4)
        Cc = 64
        W = torch.from_numpy((rng.standard_normal((3, 3, 3, 2 * Cc, 4 * Cc)) * 0.02).astype(np.float32)).cuda()
        bb = torch.from_numpy(rng.normal(0, 0.1, 4 * Cc).astype(np.float32)).cuda()
        lcfg = small_cfg(nvox=16, nvox_z=16, samples=6, NUM_VIEWS=4, GRID_REAS="lstm3d", TOP_DOWN_PYRAMID_SIZE=Cc)
        d1 = [t[:1].contiguous() for t in d]
        full = m.grid_reas(m.unproj_feat(d1, lcfg), "lstm_ref", lcfg, params={"W": W, "b": bb})
        ref_rays = m.proj_grid([full, d1[1], d1[2]], lcfg, 24)
        lrays, slab = mvd.lstm_slab(*d1, lcfg, {"W": W, "b": bb}, proj_size=24)
        xb, xc = mvd.slab_bounds(16, rank, world_size)
        res["lstm_slab"] = float((slab - full[:, xb:xb + xc]).abs().max())
        res["lstm_rays"] = float((lrays - ref_rays).abs().max())
        torch.cuda.synchronize()
        np.save(os.path.join(out_dir, "res_%d.npy" % rank), np.array([res[k] for k in sorted(res)]))
    finally:
        dist.destroy_process_group()


def test_nccl_strategies_match_single_gpu(tmp_path):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        allreduce, lstm_rays, lstm_slab, max_exact, reduce_scatter, scene_err, slab_owner, slab_owner_depth = \
            np.load(os.path.join(str(tmp_path), "res_%d.npy" % r))
        assert slab_owner_depth < 1e-5
        assert lstm_slab == 0.0 and lstm_rays == 0.0        # same tiles, same K order: bit-identical to the unsharded run
        assert allreduce < 1e-5 and reduce_scatter < 1e-5 and slab_owner < 1e-5      # partial-sum order differs
        assert scene_err == 0.0 and max_exact == 0.0
