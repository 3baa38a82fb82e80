"""GPU parity of the tensor-core (tcgen05, 3xTF32) ConvLSTM step against the oracle restatement of
ConvLSTMCell.call (mrcnn/recurrent.py:442-479) and against the exact-fp32 CUDA-core step.

Tolerance: the reference convolves in fp32; the 3xTF32 split (a_hi*b_hi + a_hi*b_lo + a_lo*b_hi, fp32
accumulation in TMEM) carries ~2^-21 per product, so gates agree to ~1e-6 and the bounded outputs
h, c are compared at rtol=1e-5, atol=3e-6."""
import numpy as np
import pytest

import oracle
from helpers import to_dev, close

pytestmark = pytest.mark.gpu


def _m():
    import mulit_view_object_detection_b200 as m
    return m


def _weights(rng, C, F):
    W = (rng.standard_normal((3, 3, 3, C + F, 4 * F)) * np.sqrt(2.0 / (27 * (C + F) + 4 * F))).astype(np.float32)
    b = rng.normal(0, 0.1, 4 * F).astype(np.float32)
    return W, b


@pytest.mark.parametrize("B,X,Y,Z,C", [(1, 4, 4, 8, 64), (2, 2, 8, 8, 64), (1, 3, 5, 6, 64), (1, 2, 2, 32, 128)])
def test_tc_step_matches_oracle(B, X, Y, Z, C):
    m = _m()
    rng = np.random.default_rng(C + X)
    F = C
    W, b = _weights(rng, C, F)
    x0 = rng.standard_normal((B, X, Y, Z, C)).astype(np.float32)
    x1 = rng.standard_normal((B, X, Y, Z, C)).astype(np.float32)
    dW, db, dx0, dx1 = to_dev(W, b, x0, x1)
    cell = m.ConvLSTMTensorCore(dW, db, 1.0)
    zeros = np.zeros((B, X, Y, Z, F), np.float32)
    # step 1: zero initial state (h_prev = c_prev = NULL), ReLU on load (model_multi.py:459)
    h1, c1 = cell.step(dx0, None, None, relu_in=True)
    oh1, oc1 = oracle.convlstm_cell_step(np.maximum(x0, 0), zeros, zeros, W, b)
    close(h1.cpu().numpy(), oh1, rtol=1e-5, atol=3e-6)
    close(c1.cpu().numpy(), oc1, rtol=1e-5, atol=3e-6)
    # step 2: recurrent state present
    h2, c2 = cell.step(dx1, h1, c1, relu_in=False)
    oh2, oc2 = oracle.convlstm_cell_step(x1, oc1, oh1, W, b)
    close(h2.cpu().numpy(), oh2, rtol=1e-5, atol=3e-6)
    close(c2.cpu().numpy(), oc2, rtol=1e-5, atol=3e-6)
    # and the exact-fp32 CUDA-core kernel agrees with the tensor-core one
    h2f, c2f = m.convlstm_step(dx1, h1, c1, dW, db)
    close(h2.cpu().numpy(), h2f.cpu().numpy(), rtol=1e-5, atol=3e-6)


def test_tc_rejects_unsupported_shapes():
    m = _m()
    rng = np.random.default_rng(0)
    W, b = _weights(rng, 16, 16)
    with pytest.raises(ValueError):
        m.ConvLSTMTensorCore(*to_dev(W, b))


def test_grid_reas_lstm3d_uses_tensor_cores():
    """grid_reas('lstm3d') (model_multi.py:457-462): ReLU -> ConvLSTM over the views -> BN -> ReLU, C = F = 64."""
    m = _m()
    from helpers import small_cfg
    rng = np.random.default_rng(5)
    B, V, X, Z, C = 1, 3, 4, 8, 64
    grids = rng.standard_normal((B, V, X, X, Z, C)).astype(np.float32)
    W, b = _weights(rng, C, C)
    bn = (np.full(C, 1.1, np.float32), np.full(C, 0.02, np.float32), np.full(C, -0.01, np.float32), np.full(C, 0.9, np.float32))
    cfg = small_cfg(GRID_REAS="lstm3d", TOP_DOWN_PYRAMID_SIZE=C, nvox=X, nvox_z=Z)
    dW, db = to_dev(W, b)
    n0 = m.launch_count()
    out = m.grid_reas(to_dev(grids)[0], "grid_reas_P4", cfg, params={"W": dW, "b": db, "bn": bn})
    o = oracle.grid_reas(grids, "grid_reas_P4", cfg, {"W": W, "b": b, "bn": bn})
    close(out.cpu().numpy(), o, rtol=1e-5, atol=5e-6)
    assert m.launch_count() - n0 >= 1 + V            # prepare + one tensor-core step per view (plus the hi/lo split passes)


@pytest.mark.parametrize("B,V,X,Y,Z,C,Cout", [(1, 3, 4, 4, 8, 32, 64), (2, 2, 3, 5, 6, 64, 48), (1, 4, 2, 2, 32, 64, 256),
                                              (1, 2, 4, 4, 8, 32, 320)])
def test_ident_tc_matches_oracle(B, V, X, Y, Z, C, Cout):
    """grid_reas('ident') (model_multi.py:443-455) on the tensor cores vs the oracle and vs the CUDA-core kernel."""
    m = _m()
    from helpers import small_cfg
    rng = np.random.default_rng(V * 100 + Cout)
    cfg = small_cfg(GRID_REAS="ident", NUM_VIEWS=V, nvox=X, nvox_z=Z)
    grids = rng.standard_normal((B, V, X, Y, Z, C)).astype(np.float32)
    W = (rng.standard_normal((V * C, Cout)) * 0.1).astype(np.float32)
    b = rng.normal(0, 0.1, Cout).astype(np.float32)
    bn = (np.full(Cout, 0.9, np.float32), np.full(Cout, 0.1, np.float32), np.zeros(Cout, np.float32), np.ones(Cout, np.float32))
    dg, dW, db = to_dev(grids, W, b)
    params = {"weight": dW, "bias": db, "bn": bn}
    out = m.grid_reas(dg, "ident_tc_%d_%d" % (V, Cout), cfg, params=params, tensor_cores=True)
    o = oracle.grid_reas(grids, "grid_reas_P4", cfg, {"weight": W, "bias": b, "bn": bn})
    close(out.cpu().numpy(), o, rtol=1e-5, atol=5e-6)
    ref = m.grid_reas(dg, "ident_fp32", cfg, params=params, tensor_cores=False)
    close(out.cpu().numpy(), ref.cpu().numpy(), rtol=1e-5, atol=5e-6)


def test_tc_slab_steps_with_halo_match_full_grid():
    """mvf_convlstm_step_tc_slab: two x-slabs with a 1-voxel halo of x and h (planes copied by hand here, exchanged by
    dist.exchange_halo in the multi-GPU path) reproduce the full-grid recurrence bit for bit (same tiles, same K order)."""
    import torch
    m = _m()
    rng = np.random.default_rng(11)
    B, X, Y, Z, C = 1, 6, 4, 8, 64
    W, b = _weights(rng, C, C)
    xs = [rng.standard_normal((B, X, Y, Z, C)).astype(np.float32) for _ in range(3)]
    dW, db = to_dev(W, b)
    dx = to_dev(*xs)
    cell = m.ConvLSTMTensorCore(dW, db, 1.0)
    # slabs: rank 0 owns x in [0,4) (halo on the high side), rank 1 owns [4,6) (halo on the low side)
    spans = [(0, 4, 0, 1), (4, 2, 1, 0)]
    hs, cs = [None, None], [None, None]
    h = c = None
    for t in range(3):
        # the operand scale of the fp16 split comes from max|operand| over the WHOLE grid (dist.lstm_slab all-reduces it)
        amax = dx[t].amax().clamp_min(0).reshape(1)
        if h is not None:
            amax = torch.maximum(amax, h.abs().amax().reshape(1))
        h, c = cell.step(dx[t], h, c, relu_in=True)                      # the unsharded recurrence
        new = []
        for r, (xb, xc, lo, hi) in enumerate(spans):
            x_pad = dx[t][:, xb - lo:xb + xc + hi].contiguous()
            new.append(cell.step_slab(x_pad, hs[r], cs[r], (lo, hi), relu_in=True, act_amax=amax))
        (h0, c0), (h1, c1) = new
        h0[:, 4].copy_(h1[:, 1])          # rank 1's first interior plane -> rank 0's high halo
        h1[:, 0].copy_(h0[:, 3])          # rank 0's last interior plane  -> rank 1's low halo
        hs, cs = [h0, h1], [c0, c1]
    got_h = torch.cat([hs[0][:, :4], hs[1][:, 1:]], dim=1)
    got_c = torch.cat(cs, dim=1)
    assert torch.equal(got_h, h) and torch.equal(got_c, c)


# ---- config c3's channel count: C = F = 256, K = 27 * 512 = 13 824 (the accumulation-error worst case of DESIGN 3.3) -------------
# Tolerance: h and c are bounded (|h| < 1, |c| < t after t steps); the float64-accumulating oracle vs the 3xFP16 split with
# promoted fp32 accumulation differs by <= ~3e-6 absolute on the O(1) gate pre-activations, so the comparison is carried by
# atol (3e-6) for most elements -- rtol=1e-5 only matters for |value| > 0.3.
C3_RTOL, C3_ATOL = 1e-5, 3e-6


def test_tc_step_matches_oracle_at_c3_channels():
    """ConvLSTMTensorCore.step at C = F = 256 vs oracle.convlstm_cell_step (recurrent.py:442-479) on a 4x4x8 grid:
    first step (no h: x chunks only) and a recurrent step (x and h chunks, K = 13 824)."""
    m = _m()
    rng = np.random.default_rng(256)
    B, X, Y, Z, C = 1, 4, 4, 8, 256
    W, b = _weights(rng, C, C)
    x0 = rng.standard_normal((B, X, Y, Z, C)).astype(np.float32)
    x1 = rng.standard_normal((B, X, Y, Z, C)).astype(np.float32)
    dW, db, dx0, dx1 = to_dev(W, b, x0, x1)
    cell = m.ConvLSTMTensorCore(dW, db, 1.0)
    zeros = np.zeros((B, X, Y, Z, C), np.float32)
    h1, c1 = cell.step(dx0, None, None, relu_in=True)
    oh1, oc1 = oracle.convlstm_cell_step(np.maximum(x0, 0), zeros, zeros, W, b)
    close(h1.cpu().numpy(), oh1, rtol=C3_RTOL, atol=C3_ATOL)
    close(c1.cpu().numpy(), oc1, rtol=C3_RTOL, atol=C3_ATOL)
    h2, c2 = cell.step(dx1, h1, c1, relu_in=False)
    oh2, oc2 = oracle.convlstm_cell_step(x1, oc1, oh1, W, b)
    close(h2.cpu().numpy(), oh2, rtol=C3_RTOL, atol=C3_ATOL)
    close(c2.cpu().numpy(), oc2, rtol=C3_RTOL, atol=C3_ATOL)
    assert float(np.abs(oh2).max()) > 0.05            # the gates are not saturated at zero


def test_tc_slab_step_matches_oracle_at_c3_channels():
    """mvf_convlstm_step_tc_slab at C = F = 256: two x-slabs with hand-copied halo planes vs the oracle's full-grid step
    (with and without h), and bit-identical to the unsharded tensor-core step."""
    import torch
    m = _m()
    rng = np.random.default_rng(257)
    B, X, Y, Z, C = 1, 4, 4, 8, 256
    W, b = _weights(rng, C, C)
    xs = [rng.standard_normal((B, X, Y, Z, C)).astype(np.float32) for _ in range(2)]
    dW, db = to_dev(W, b)
    dx = to_dev(*xs)
    cell = m.ConvLSTMTensorCore(dW, db, 1.0)
    spans = [(0, 2, 0, 1), (2, 2, 1, 0)]
    hs, cs = [None, None], [None, None]
    h = c = None
    oh = oc = np.zeros((B, X, Y, Z, C), np.float32)
    for t in range(2):
        amax = dx[t].amax().clamp_min(0).reshape(1)
        if h is not None:
            amax = torch.maximum(amax, h.abs().amax().reshape(1))
        h, c = cell.step(dx[t], h, c, relu_in=True)
        oh, oc = oracle.convlstm_cell_step(np.maximum(xs[t], 0), oc, oh, W, b)
        new = []
        for r, (xb, xc, lo, hi) in enumerate(spans):
            x_pad = dx[t][:, xb - lo:xb + xc + hi].contiguous()
            new.append(cell.step_slab(x_pad, hs[r], cs[r], (lo, hi), relu_in=True, act_amax=amax))
        (h0, c0), (h1, c1) = new
        h0[:, 2].copy_(h1[:, 1])
        h1[:, 0].copy_(h0[:, 1])
        hs, cs = [h0, h1], [c0, c1]
        got_h = torch.cat([hs[0][:, :2], hs[1][:, 1:]], dim=1)
        got_c = torch.cat(cs, dim=1)
        assert torch.equal(got_h, h) and torch.equal(got_c, c)
        close(got_h.cpu().numpy(), oh, rtol=C3_RTOL, atol=C3_ATOL)
        close(got_c.cpu().numpy(), oc, rtol=C3_RTOL, atol=C3_ATOL)


def test_tc_step_vs_fp32_cuda_kernel_at_c3_size():
    """One recurrent step of config c3 at full size (64^3 voxels, C = F = 256, K = 13 824) against the exact-fp32 CUDA-core kernel
    (mvf_convlstm_step: fp32 FMA accumulation, itself ~1e-6 away from exact arithmetic at this K).  Measured on B200:
    max |dh| = 4.6e-6, max |dc| = 7.0e-6 with |c| up to ~2.  Both sides carry rounding error here (the oracle tests above compare
    with float64 accumulation at atol=3e-6), so the bar is rtol=1e-5 / atol=8e-6 (the 7.0e-6 maximum sits on an element with
    |c| < 0.1), and 1e-5 absolute overall."""
    import torch
    m = _m()
    g = torch.Generator(device="cuda")
    g.manual_seed(0)
    X, C = 64, 256
    W = torch.randn((3, 3, 3, 2 * C, 4 * C), device="cuda", generator=g) * (2.0 / (27 * 2 * C + 4 * C)) ** 0.5
    b = torch.randn(4 * C, device="cuda", generator=g) * 0.1
    x = torch.randn((1, X, X, X, C), device="cuda", generator=g).relu_()
    cell = m.ConvLSTMTensorCore(W, b, 1.0)
    h1, c1 = cell.step(x, None, None)
    h2, c2 = cell.step(x, h1, c1)
    h2f, c2f = m.convlstm_step(x, h1, c1, W, b)
    eh = (h2 - h2f).abs().max().item()
    ec = (c2 - c2f).abs().max().item()
    assert eh <= 1e-5 and ec <= 1e-5, (eh, ec)
    assert bool(((h2 - h2f).abs() <= 8e-6 + C3_RTOL * h2f.abs()).all()), eh
    assert bool(((c2 - c2f).abs() <= 8e-6 + C3_RTOL * c2f.abs()).all()), ec
    assert h2f.abs().max().item() > 0.1
