"""Multi-GPU sharding of the fusion path: one process per GPU, ``torch.distributed`` (NCCL over
NVLink 5 / NVSwitch on the B200 box, gloo in the CPU tests) for the plumbing.

The reference has no multi-GPU path at all (its ``GPU_COUNT > 1`` branch imports a module that
does not exist, mrcnn/model_multi.py:2557-2559), so this layer is new; SURVEY.md section 8(e) is its spec.
Five strategies, all producing the same ray slices ``proj_grid(grid_reas(unproj_feat(...)))``:

* ``scene_shard``          scenes are independent: rank r takes scenes r::world, no collective.
* ``view_shard_allreduce`` rank r unprojects its views into a full-size partial grid, one
                           ``all_reduce(sum|max)`` over N*C*4 bytes, every rank projects.
* ``view_shard_reduce_scatter``  partial grids are reduce-scattered into x-slabs, each rank
                           projects from its slab only (a ray sample comes from exactly one
                           slab, the others contribute 0) and the small ray tensor is all-reduced.
* ``lstm_slab``            recurrent fusion (GRID_REAS='lstm3d') cannot shard views (the ConvLSTM is sequential in
                           the view axis): ranks own x-slabs, unproject view t for their slab + 1-voxel halo, run the
                           ConvLSTM step slab-locally and exchange the two boundary planes of ``h`` with their
                           neighbours after every step (isend/irecv), then project slab-locally.
* ``slab_owner``           features are tiny (13 MB for 8 views): every rank holds all views,
                           unprojects ALL of them for its own x-slab (no grid exchange at all),
                           projects locally, all-reduces the ray tensor.

The compute steps go through an ``ops`` object (default: the CUDA layers of this package); the
CPU tests inject an oracle-backed ``ops`` so that the sharding / collective logic is covered with
``gloo`` at world_size 2 without a GPU.
"""
import torch
import torch.distributed as dist


class CudaOps:
    """The product path: kernels of libmvfusion.so on torch CUDA tensors."""

    def unproject_fuse(self, feats, Rcam, Kmat, config, mode, Rmain=None, x_slab=None, tensor_cores=None):
        from . import layers
        return layers.unproject_fuse(feats, Rcam, Kmat, config, mode=mode, Rmain=Rmain, x_slab=x_slab, tensor_cores=tensor_cores)

    def proj_grid(self, grid, Rcam, Kmat, config, proj_size, x_slab=None):
        from . import layers
        return layers.proj_grid([grid, Rcam, Kmat], config, proj_size, x_slab=x_slab)

    def proj_collapse_linear(self, grid, Rcam, Kmat, config, proj_size, weight, x_slab=None):
        """The LINEAR part of proj_grid + depth_sampling for one grid slab: sum_s w_s * sample_s, [B,P,P,C] (K3b without bias /
        BatchNorm / ReLU).  Slabs add up exactly: a ray sample lies in exactly one slab, the others contribute 0."""
        from . import layers
        params = {"weight": weight, "bias": 0.0, "bn": None}
        return layers.proj_grid_depth_sampling([grid, Rcam, Kmat], config, proj_size, "depth", params=params, x_slab=x_slab,
                                               linear=True)

    def depth_affine_relu(self, x, depth):
        """ReLU(BN_scalar(x + bias)) of depth_sampling (model_multi.py:483-487), applied AFTER the cross-rank sum."""
        from . import layers
        return layers.depth_affine_relu(x, depth)

    def scale(self, grid, factor):
        """grid * factor through the view-reduce kernel (V=1, per-channel scale)."""
        from . import layers, _lib
        C = grid.shape[-1]
        scale = torch.full((C,), float(factor), dtype=torch.float32, device=grid.device)
        shift = torch.zeros((C,), dtype=torch.float32, device=grid.device)
        out = torch.empty_like(grid)
        B = grid.shape[0]
        rc = _lib.lib.mvf_view_reduce(layers._ptr(grid), B, 1, grid.numel() // (B * C), C, _lib.FUSE_SUM, 0,
                                      layers._ptr(scale), layers._ptr(shift), layers._ptr(out), layers._stream())
        _lib.check(rc, "mvf_view_reduce")
        return out


    def convlstm_cell(self, W, bias):
        from . import layers
        return layers.ConvLSTMTensorCore(W, bias, 1.0)

    def affine_relu(self, h, bn):
        from . import layers
        scale, shift = layers._bn_affine(bn if bn is not None else layers._default_bn(h.shape[-1]), h.shape[-1], h.device)
        return layers._affine_relu(h.contiguous(), scale, shift)


def world(group=None):
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def view_slice(V, rank, world_size):
    """Contiguous, balanced split of the view axis: views [lo, hi) of rank ``rank``."""
    base, rem = divmod(V, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def slab_bounds(X, rank, world_size):
    """x-slab (x_begin, x_count) of rank ``rank``; X must divide evenly for reduce-scatter."""
    base, rem = divmod(X, world_size)
    lo = rank * base + min(rank, rem)
    return lo, base + (1 if rank < rem else 0)


def scene_indices(B, rank, world_size):
    return list(range(rank, B, world_size))


def _reduce_op(mode):
    return dist.ReduceOp.MAX if mode == "max" else dist.ReduceOp.SUM


def _all_reduce(t, op, group):
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=op, group=group)
    return t


def _reduce_scatter_x(partial, op, group):
    """[B,X,Y,Z,C] partial grids -> this rank's reduced x-slab [B,X/W,Y,Z,C]."""
    rank, ws = world(group)
    if ws == 1:
        return partial
    B, X = partial.shape[:2]
    if X % ws:
        raise ValueError("reduce-scatter by slab needs nvox (%d) divisible by the world size (%d)" % (X, ws))
    xs = X // ws
    out = torch.empty((B, xs) + tuple(partial.shape[2:]), dtype=partial.dtype, device=partial.device)
    if dist.get_backend(group) == "gloo":                # gloo has no reduce_scatter: reduce, then slice
        full = partial.clone()
        dist.all_reduce(full, op=op, group=group)
        out.copy_(full[:, rank * xs:(rank + 1) * xs])
        return out
    for b in range(B):                                   # per scene the x-slabs are contiguous chunks
        dist.reduce_scatter_tensor(out[b], partial[b].contiguous(), op=op, group=group)
    return out


def _pipelined(B, ws, compute, exchange, finish):
    """Software pipeline over scenes: the collective of scene b runs (async, on the backend's own stream) while scene b+1 is
    being unprojected on the compute stream.  ``compute(b)`` -> tensor, ``exchange(t)`` -> (tensor, work | None),
    ``finish(b, t)`` consumes the exchanged tensor."""
    pending = None
    for b in range(B):
        t = compute(b)
        if pending is not None:
            pb, pt, pw = pending
            if pw is not None:
                pw.wait()
            finish(pb, pt)
        if ws > 1:
            t, w = exchange(t)
        else:
            w = None
        pending = (b, t, w)
    pb, pt, pw = pending
    if pw is not None:
        pw.wait()
    finish(pb, pt)


def view_shard_allreduce(feats, Rcam, Kmat, config, proj_size, mode="sum", group=None, ops=None):
    """``feats`` / ``Rcam`` hold ALL views on every rank (or at least this rank's slice is valid);
    each rank unprojects views ``view_slice(V, rank, world)`` only, scene by scene, and the all-reduce of scene b overlaps the
    unprojection of scene b+1.  Returns (rays, fused grid)."""
    ops = ops or CudaOps()
    rank, ws = world(group)
    B, V = feats.shape[0], feats.shape[1]
    lo, hi = view_slice(V, rank, ws)
    Rmain = Rcam[:, 0].contiguous()
    local_mode = "max" if mode == "max" else "sum"
    g = config
    fused = torch.empty((B, g.nvox, g.nvox, g.nvox_z, feats.shape[-1]), dtype=feats.dtype, device=feats.device)

    def compute(b):
        if hi > lo:
            return ops.unproject_fuse(feats[b:b + 1, lo:hi].contiguous(), Rcam[b:b + 1, lo:hi].contiguous(), Kmat[b:b + 1], config,
                                      local_mode, Rmain=Rmain[b:b + 1])
        # more ranks than views: identity element
        return torch.full((1,) + tuple(fused.shape[1:]), float("-inf") if mode == "max" else 0.0, dtype=feats.dtype, device=feats.device)

    def exchange(t):
        return t, dist.all_reduce(t, op=_reduce_op(mode), group=group, async_op=True)

    def finish(b, t):
        fused[b:b + 1].copy_(ops.scale(t, 1.0 / V) if mode == "mean" else t)

    _pipelined(B, ws, compute, exchange, finish)
    rays = ops.proj_grid(fused, Rcam, Kmat, config, proj_size)
    return rays, fused


def view_shard_reduce_scatter(feats, Rcam, Kmat, config, proj_size, mode="sum", group=None, ops=None):
    """View-sharded unprojection, reduce-scatter of the grid by x-slab, slab-local projection,
    all-reduce(sum) of the ray slices.  Returns (rays, this rank's grid slab)."""
    ops = ops or CudaOps()
    rank, ws = world(group)
    B, V = feats.shape[0], feats.shape[1]
    lo, hi = view_slice(V, rank, ws)
    if hi <= lo:
        raise ValueError("view sharding needs at least one view per rank (V=%d, world=%d)" % (V, ws))
    if config.nvox % ws:
        raise ValueError("reduce-scatter by slab needs nvox (%d) divisible by the world size (%d)" % (config.nvox, ws))
    Rmain = Rcam[:, 0].contiguous()
    xb, xc = slab_bounds(config.nvox, rank, ws)
    g = config
    slab = torch.empty((B, xc, g.nvox, g.nvox_z, feats.shape[-1]), dtype=feats.dtype, device=feats.device)
    gloo = ws > 1 and dist.get_backend(group) == "gloo"                    # gloo has no reduce_scatter: reduce, then slice

    def compute(b):
        return ops.unproject_fuse(feats[b:b + 1, lo:hi].contiguous(), Rcam[b:b + 1, lo:hi].contiguous(), Kmat[b:b + 1], config,
                                  "max" if mode == "max" else "sum", Rmain=Rmain[b:b + 1])

    def exchange(t):                                                        # per scene the x-slabs are contiguous chunks
        if gloo:
            return t, dist.all_reduce(t, op=_reduce_op(mode), group=group, async_op=True)
        out = torch.empty((1, xc) + tuple(t.shape[2:]), dtype=t.dtype, device=t.device)
        return out, dist.reduce_scatter_tensor(out[0], t[0], op=_reduce_op(mode), group=group, async_op=True)

    def finish(b, t):
        part = t[:, xb:xb + xc] if (gloo or ws == 1) else t
        slab[b:b + 1].copy_(ops.scale(part.contiguous(), 1.0 / V) if mode == "mean" else part)

    _pipelined(B, ws, compute, exchange, finish)
    rays = ops.proj_grid(slab, Rcam, Kmat, config, proj_size, x_slab=(xb, xc))
    rays = _all_reduce(rays, dist.ReduceOp.SUM, group)
    return rays, slab


def _reduce_scatter_scenes(rays, group):
    """Sum over ranks, each rank keeping the contiguous block of B/world scenes it owns (half the bytes of an all-reduce)."""
    rank, ws = world(group)
    if ws == 1:
        return rays
    B = rays.shape[0]
    if B % ws:
        raise ValueError("scatter_scenes needs the scene count (%d) divisible by the world size (%d)" % (B, ws))
    nb = B // ws
    if dist.get_backend(group) == "gloo":                # gloo has no reduce_scatter: reduce, then slice
        dist.all_reduce(rays, op=dist.ReduceOp.SUM, group=group)
        return rays[rank * nb:(rank + 1) * nb].contiguous()
    out = torch.empty((nb,) + tuple(rays.shape[1:]), dtype=rays.dtype, device=rays.device)
    dist.reduce_scatter_tensor(out, rays.contiguous(), op=dist.ReduceOp.SUM, group=group)
    return out


def slab_owner(feats, Rcam, Kmat, config, proj_size, mode="sum", group=None, ops=None, scatter_scenes=False, depth=None):
    """Owner-computes: every rank holds all views and fuses ALL of them for its own x-slab;
    the only exchange is the all-reduce(sum) of the ray slices -- or, with ``scatter_scenes``, a reduce-scatter that
    leaves rank r with the ray slices of scenes [r*B/W, (r+1)*B/W) only (the downstream heads are data parallel over
    scenes).  Returns (rays, grid slab).

    ``depth`` = the depth_sampling learnables ({'weight' [S], 'bias', 'bn'}, model_multi.py:481-487): the neck's real boundary.
    The depth collapse is linear up to its bias, so each rank collapses its own slab's samples (K3b, no bias / BN / ReLU),
    the all-reduce moves [B,P,P,C] (1.6 MB per T-scene instead of 32.8 MB of mostly-zero ray slices), and bias + BatchNorm +
    ReLU are applied after the sum.  Returns (PG [B,P,P,C], grid slab).  A rank whose slab is empty (more ranks than x-planes)
    contributes zeros."""
    ops = ops or CudaOps()
    rank, ws = world(group)
    xb, xc = slab_bounds(config.nvox, rank, ws)
    slab = ops.unproject_fuse(feats, Rcam, Kmat, config, mode, x_slab=(xb, xc))
    if depth is not None:
        pg = ops.proj_collapse_linear(slab, Rcam, Kmat, config, proj_size, depth["weight"], x_slab=(xb, xc))
        pg = _reduce_scatter_scenes(pg, group) if scatter_scenes else _all_reduce(pg, dist.ReduceOp.SUM, group)
        return ops.depth_affine_relu(pg, depth), slab
    rays = ops.proj_grid(slab, Rcam, Kmat, config, proj_size, x_slab=(xb, xc))
    rays = _reduce_scatter_scenes(rays, group) if scatter_scenes else _all_reduce(rays, dist.ReduceOp.SUM, group)
    return rays, slab


def scene_shard(feats, Rcam, Kmat, config, proj_size, mode="sum", group=None, ops=None, gather=False):
    """Data parallel over scenes: rank r processes scenes r::world with no collective.
    With ``gather`` the per-rank ray slices are all-gathered back into scene order."""
    ops = ops or CudaOps()
    rank, ws = world(group)
    B = feats.shape[0]
    mine = scene_indices(B, rank, ws)
    idx = torch.as_tensor(mine, dtype=torch.long, device=feats.device)
    f, R, K = feats.index_select(0, idx), Rcam.index_select(0, idx), Kmat.index_select(0, idx)
    fused = ops.unproject_fuse(f, R, K, config, mode)
    rays = ops.proj_grid(fused, R, K, config, proj_size)
    if not gather or ws == 1:
        return rays, mine
    if B % ws:
        raise ValueError("gather needs the scene count (%d) divisible by the world size (%d)" % (B, ws))
    parts = [torch.empty_like(rays) for _ in range(ws)]
    dist.all_gather(parts, rays, group=group)
    out = torch.empty((B,) + tuple(rays.shape[1:]), dtype=rays.dtype, device=rays.device)
    for r in range(ws):
        out[r::ws] = parts[r]
    return out, mine


def _peer(group, r):
    return dist.get_global_rank(group, r) if group is not None else r


def exchange_halo(h, lo, hi, group=None):
    """``h`` [B, lo+Xs+hi, Y, Z, F] with valid interior planes: the first interior plane goes to rank-1 (it becomes
    that rank's high halo plane), the last interior plane to rank+1 (its low halo plane); the received planes are
    written into this rank's halo planes.  2 x Y*Z*F*4 bytes per neighbour (4 MB at 64^2 x 256)."""
    rank, ws = world(group)
    if ws == 1 or not (lo or hi):
        return h
    Xin = h.shape[1]
    ops, recv_lo, recv_hi = [], None, None
    if lo:
        send = h[:, lo].contiguous()
        recv_lo = torch.empty_like(send)
        ops += [dist.P2POp(dist.isend, send, _peer(group, rank - 1), group),
                dist.P2POp(dist.irecv, recv_lo, _peer(group, rank - 1), group)]
    if hi:
        send = h[:, Xin - 1 - hi].contiguous()
        recv_hi = torch.empty_like(send)
        ops += [dist.P2POp(dist.isend, send, _peer(group, rank + 1), group),
                dist.P2POp(dist.irecv, recv_hi, _peer(group, rank + 1), group)]
    for req in dist.batch_isend_irecv(ops):
        req.wait()
    if lo:
        h[:, 0].copy_(recv_lo)
    if hi:
        h[:, Xin - 1].copy_(recv_hi)
    return h


def lstm_slab(feats, Rcam, Kmat, config, params, proj_size=None, group=None, ops=None):
    """GRID_REAS='lstm3d' (model_multi.py:457-462: ReLU -> ConvLSTM over the views -> BN -> ReLU) with the grid split
    into x-slabs.  ``params``: {'W' [3,3,3,C+F,4F], 'b' [4F], 'bn' optional}.  Every rank holds all views' features
    (13 MB).  Returns (ray slices or None, this rank's fused slab [B,Xs,Y,Z,F])."""
    ops = ops or CudaOps()
    rank, ws = world(group)
    xb, xc = slab_bounds(config.nvox, rank, ws)
    if xc < 1:
        raise ValueError("lstm_slab needs at least one x-plane per rank (nvox=%d, world=%d)" % (config.nvox, ws))
    lo, hi = (1 if rank > 0 else 0), (1 if rank < ws - 1 else 0)
    V = feats.shape[1]
    Rmain = Rcam[:, 0].contiguous()
    cell = ops.convlstm_cell(params["W"], params["b"])
    h = c = None
    for t in range(V):
        # view t unprojected for the slab and its halo planes; a 1-view 'sum' is the per-view grid itself
        # (slot kernel: its per-voxel arithmetic does not depend on the slab, which keeps the sharded recurrence bit-identical
        # to the unsharded one at every world size)
        x_t = ops.unproject_fuse(feats[:, t:t + 1].contiguous(), Rcam[:, t:t + 1].contiguous(), Kmat, config, "sum",
                                 Rmain=Rmain, x_slab=(xb - lo, xc + lo + hi), tensor_cores=False)
        # the fp16 operand split scales by a power of two taken from max|operand|: agree on it across the slabs
        # (4-byte all-reduce) so that the sharded recurrence is bit-identical to the unsharded one
        amax = x_t.amax().clamp_min(0).reshape(1)
        if h is not None:
            amax = torch.maximum(amax, h.abs().amax().reshape(1))
        _all_reduce(amax, dist.ReduceOp.MAX, group)
        h, c = cell.step_slab(x_t, h, c, (lo, hi), relu_in=True, act_amax=amax)
        if t + 1 < V:
            exchange_halo(h, lo, hi, group)
    slab = ops.affine_relu(h[:, lo:lo + xc], params.get("bn"))
    if proj_size is None:
        return None, slab
    rays = ops.proj_grid(slab, Rcam, Kmat, config, proj_size, x_slab=(xb, xc))
    rays = _all_reduce(rays, dist.ReduceOp.SUM, group)
    return rays, slab


def fuse_project_auto(feats, Rcam, Kmat, config, proj_size, mode="sum", group=None, ops=None):
    """Pick the sharding that moves the fewest bytes (measured in profiles/r1_scaling.json): enough scenes
    for every rank -> ``scene_shard`` (no exchange at all; rays gathered back into scene order); fewer scenes than ranks ->
    ``slab_owner`` (features replicated, one all-reduce of the ray slices).  View sharding is never chosen: its partial grids
    cross NVLink (25 GB per rank at config c5) and it is slower than one GPU.  Returns the full ray slices [B,S,P,P,C]."""
    rank, ws = world(group)
    B = feats.shape[0]
    if ws == 1 or (B >= ws and B % ws == 0):
        rays, _ = scene_shard(feats, Rcam, Kmat, config, proj_size, mode=mode, group=group, ops=ops, gather=True)
        return rays
    rays, _ = slab_owner(feats, Rcam, Kmat, config, proj_size, mode=mode, group=group, ops=ops)
    return rays
