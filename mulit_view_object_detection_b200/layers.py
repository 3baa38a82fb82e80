"""Host-side mirror of the reference's fusion-path layers over libmvfusion.so.

Same names, argument order, ``config`` attribute names and output layouts as the Keras/TF1
layers in mrcnn/model_multi.py and mrcnn/recurrent.py, operating on contiguous fp32
channel-last **torch CUDA tensors** in place of TF tensors:

    unproj_feat([feats, Rcam, Kmat], config)                      model_multi.py:130
    grid_reas(x, scope, config)                                   model_multi.py:394
    convlstm(grid, name, kernel, filters)                         model_multi.py:109
    proj_grid([grid, Rcam, Kmat], config, proj_size)              model_multi.py:231
    depth_sampling(x, config, name)                               model_multi.py:466
    PyramidROIAlign(pool_shape)([boxes, image_meta] + maps)       model_multi.py:779
    refine_detections_graph(rois, probs, deltas, window, config)  model_multi.py:1119
    DetectionLayer(config)([rois, cls, bbox, image_meta])         model_multi.py:1217
    ProposalLayer(count, nms_thr, config)([probs, bbox, anchors]) model_multi.py:690

plus the fused entries ``unproject_fuse`` / ``unproject_fuse_project`` that never materialise
the per-view grids.  torch is plumbing only (device memory, streams); every op is a kernel of
libmvfusion.so launched on ``torch.cuda.current_stream()``.  There is no CPU path: a non-CUDA
tensor raises ``ValueError`` (the reference raises on CPU too -- TF-CPU ``gather_nd`` rejects
the out-of-range taps this path produces).

Learnable state (frozen inference weights) is looked up by layer scope/name in ``weights``,
mirroring the reference's ``reused_lay`` registry (model_multi.py:112-117); missing entries
use the Keras initial values (BatchNorm gamma=1, beta=0, mean=0, var=1).
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import lib, check, grid_from_config

BN_EPS = 1e-3                      # Keras BatchNormalization default (model_multi.py:501-502)

# scope/name -> dict of tensors; the analogue of the reference's `reused_lay`
weights = {}
reused_lay = weights


def set_weights(scope, **tensors):
    weights[scope] = tensors


# ------------------------------------------------------------------------------------------------
def _cuda(t, name, dtype=torch.float32):
    if not isinstance(t, torch.Tensor):
        raise ValueError("%s must be a torch tensor (got %r)" % (name, type(t)))
    if not t.is_cuda:
        raise ValueError("%s must live on a CUDA device: this path has no CPU implementation" % name)
    if t.dtype != dtype:
        raise ValueError("%s must be %s (got %s)" % (name, dtype, t.dtype))
    return t.contiguous()


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class PreparedBN:
    """A frozen BatchNorm already folded to device-side (scale, shift) -- see :func:`prepare_params`."""

    def __init__(self, scale, shift):
        self.scale, self.shift = scale, shift


class PreparedDepth:
    """depth_sampling learnables already on the device / folded to host scalars -- see :func:`prepare_params`."""

    def __init__(self, w, bias, inv, shift):
        self.w, self.bias, self.inv, self.shift = w, bias, inv, shift


def prepare_params(params, device="cuda"):
    """Convert a parameter dictionary (NumPy arrays, BN 4-tuples: what ``weights_io.fusion_params_from_keras`` returns) ONCE
    into device tensors and folded BatchNorm affines, so that the per-call host work of the layers is pointer passing only
    (and a :func:`fusion_neck` call can be captured in a CUDA graph).  The structure of the dictionary is preserved."""
    device = torch.device(device)

    def conv(v, key):
        if isinstance(v, dict):
            if "weight" in v and np.ndim(v["weight"]) == 1 and "dw1" not in v and key.startswith("grid_reas_depth"):
                w, bias, inv, shift = _depth_params(key, v, int(np.size(v["weight"])), device)
                return PreparedDepth(w, bias, inv, shift)
            return {k: conv(x, k) for k, x in v.items()}
        if key == "bn" and not isinstance(v, PreparedBN):
            n = int(np.size(v[0])) if not isinstance(v[0], torch.Tensor) else v[0].numel()
            return PreparedBN(*_bn_affine(v, n, device))
        if isinstance(v, np.ndarray):
            return torch.from_numpy(np.ascontiguousarray(v, dtype=np.float32)).to(device)
        return v

    return {k: conv(v, k) for k, v in params.items()}


def _bn_affine(bn, C_, device):
    """(scale, shift) of a frozen BatchNorm as tf.nn.batch_normalization evaluates it:
    inv = rsqrt(var + eps) * gamma ; shift = beta - mean * inv."""
    if bn is None:
        return None, None
    if isinstance(bn, PreparedBN):
        return bn.scale, bn.shift
    gamma, beta, mean, var = (torch.as_tensor(a, dtype=torch.float32, device=device).reshape(-1) for a in bn)
    inv = torch.rsqrt(var + BN_EPS) * gamma
    shift = beta - mean * inv
    if inv.numel() == 1 and C_ > 1:
        inv, shift = inv.expand(C_), shift.expand(C_)
    return inv.contiguous(), shift.contiguous()


def _default_bn(C_):
    return (np.ones(C_, np.float32), np.zeros(C_, np.float32), np.zeros(C_, np.float32), np.ones(C_, np.float32))


def _image_hw(config):
    return int(config.IMAGE_SHAPE[0]), int(config.IMAGE_SHAPE[1])


_FUSE = {"none": _lib.FUSE_NONE, "sum": _lib.FUSE_SUM, "add": _lib.FUSE_SUM, "mean": _lib.FUSE_MEAN,
         "max": _lib.FUSE_MAX}


# ------------------------------------------------------------------------------------------------
_k1t_ws = {}          # device -> scratch of the tensor-core path (fp16 operand halves of the features), grown on demand


def _k1t_workspace(device, nbytes):
    ws = _k1t_ws.get(device)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty((nbytes,), dtype=torch.uint8, device=device)
        _k1t_ws[device] = ws
    return ws


def unproject_fuse(feats, Rcam, Kmat, config, mode="sum", bn=None, relu_in=False, relu_out=False,
                   Rmain=None, x_slab=None, world_grid=False, return_aux=False, out=None, tensor_cores=None):
    """K1: unproject every view and reduce over views on chip (the per-view grids never exist).

    ``tensor_cores``: None = automatic -- the linear reductions (sum / mean, no ReLU in front, C % 64 == 0, C <= 256, no side
    outputs) run as K1T on tcgen05 (``mvf_unproject_fuse_tc``), everything else on the CUDA-core slot kernel
    (``mvf_unproject_fuse``); True / False force either (True raises when the configuration does not qualify).

    mode 'none' -> [B,V,Xs,Y,Z,C] (= unproj_feat), else [B,Xs,Y,Z,C].
    ``Rmain`` [B,3,4]: main-view pose when ``Rcam`` is a shard of the views.
    ``x_slab`` (x_begin, x_count): compute one x-slab of the grid only.
    ``return_aux``: also return idx int32 [B,V,Xs,Y,Z,2] and valid uint8 [B,V,Xs,Y,Z]."""
    feats = _cuda(feats, "feats")
    Rcam = _cuda(Rcam, "Rcam")
    Kmat = _cuda(Kmat, "Kmat")
    if feats.dim() != 5 or Rcam.dim() != 4 or Kmat.dim() != 3:
        raise ValueError("expected feats [B,V,fh,fw,C], Rcam [B,V,3,4], Kmat [B,3,3]")
    B, V, fh, fw, Cc = feats.shape
    if tuple(Rcam.shape) != (B, V, 3, 4) or tuple(Kmat.shape) != (B, 3, 3):
        raise ValueError("Rcam %s / Kmat %s do not match feats %s" % (tuple(Rcam.shape), tuple(Kmat.shape), tuple(feats.shape)))
    if Rmain is not None:
        Rmain = _cuda(Rmain, "Rmain")
        if tuple(Rmain.shape) != (B, 3, 4):
            raise ValueError("Rmain must be [B,3,4]")
    g = grid_from_config(config)
    xb, xc = (0, g.nvox) if x_slab is None else (int(x_slab[0]), int(x_slab[1]))
    Y, Z = g.nvox, g.nvox_z
    m = _FUSE[mode]
    shape = (B, V, xc, Y, Z, Cc) if m == _lib.FUSE_NONE else (B, xc, Y, Z, Cc)
    if out is None:
        out = torch.empty(shape, dtype=torch.float32, device=feats.device)
    elif tuple(out.shape) != shape or not out.is_contiguous():
        raise ValueError("out must be a contiguous %s tensor" % (shape,))
    idx = valid = gpos = None
    if return_aux:
        idx = torch.empty((B, V, xc, Y, Z, 2), dtype=torch.int32, device=feats.device)
        valid = torch.empty((B, V, xc, Y, Z), dtype=torch.uint8, device=feats.device)
    if world_grid:
        gpos = torch.empty((B, 3), dtype=torch.float32, device=feats.device)
    scale, shift = _bn_affine(bn, Cc, feats.device)
    flags = (_lib.FLAG_RELU_IN if relu_in else 0) | (_lib.FLAG_RELU_OUT if relu_out else 0) | \
            (_lib.FLAG_WORLD_GRID if world_grid else 0)
    grid_dist = float(getattr(config, "GRID_DIST", 600 / 320 * config.vmax))
    ih, iw = _image_hw(config)
    if xc == 0:                       # empty slab of a sharded caller: nothing to compute
        tensor_cores = False
    eligible = bool(lib.mvf_unproject_fuse_tc_supported(V, Cc, m, flags)) and not return_aux and not world_grid
    if tensor_cores is None and x_slab is not None and (xb % 4 != 0 or (xc % 4 != 0 and xb + xc != g.nvox)):
        # K1T works on 4x4x8 voxel tiles anchored at the slab origin; a slab that cuts the full grid's tiles would see other
        # per-tile footprints and K orders, i.e. last-bit differences from the unsharded result.  Such slabs (the 1-voxel halo
        # slabs of dist.lstm_slab) stay on the slot kernel, whose per-voxel arithmetic is independent of the slab.
        eligible = False
    if tensor_cores and not eligible:
        raise ValueError("the tensor-core unprojection needs mode sum/mean, no relu_in, C % 64 == 0, C <= 256 and no side outputs")
    if eligible and tensor_cores is not False:
        nbytes = lib.mvf_unproject_fuse_tc_workspace_bytes(B, V, fh, fw, Cc)
        ws = _k1t_workspace(feats.device, nbytes)
        rc = lib.mvf_unproject_fuse_tc(_ptr(feats), _ptr(Rcam), _ptr(Rmain), _ptr(Kmat), C.byref(g), B, V, fh, fw, Cc,
                                       ih, iw, m, flags, grid_dist, xb, xc, _ptr(scale), _ptr(shift), _ptr(out),
                                       _ptr(ws), ws.numel(), _stream())
        check(rc, "mvf_unproject_fuse_tc")
        return out
    rc = lib.mvf_unproject_fuse(_ptr(feats), _ptr(Rcam), _ptr(Rmain), _ptr(Kmat), C.byref(g), B, V, fh, fw, Cc,
                                ih, iw, m, flags, grid_dist, xb, xc, _ptr(scale), _ptr(shift),
                                _ptr(out), _ptr(idx), _ptr(valid), _ptr(gpos), _stream())
    check(rc, "mvf_unproject_fuse")
    res = (out,)
    if world_grid:
        res += (gpos,)
    if return_aux:
        res += (idx, valid)
    return res[0] if len(res) == 1 else res


def unproj_feat(inputs, config, return_aux=False):
    """``unproj_feat([feats, Rcam, Kmat], config)`` -> [B,V,X,Y,Z,C]  (model_multi.py:130-228)."""
    feats, Rcam, Kmat = inputs
    return unproject_fuse(feats, Rcam, Kmat, config, mode="none", return_aux=return_aux)


def unproj_feat_notebook(inputs, config, return_aux=False):
    """World-frame variant, returns ``[grid, grid_position]`` (Notebook/projection.py:47-151)."""
    feats, Rcam, Kmat = inputs
    return unproject_fuse(feats, Rcam, Kmat, config, mode="none", world_grid=True, return_aux=return_aux)


def view_reduce(x, mode, bn=None, relu_in=False, relu_out=False):
    x = _cuda(x, "x")
    if x.dim() != 6:
        raise ValueError("expected [B,V,X,Y,Z,C]")
    B, V, X, Y, Z, Cc = x.shape
    out = torch.empty((B, X, Y, Z, Cc), dtype=torch.float32, device=x.device)
    scale, shift = _bn_affine(bn, Cc, x.device)
    flags = (_lib.FLAG_RELU_IN if relu_in else 0) | (_lib.FLAG_RELU_OUT if relu_out else 0)
    rc = lib.mvf_view_reduce(_ptr(x), B, V, X * Y * Z, Cc, _FUSE[mode], flags, _ptr(scale), _ptr(shift), _ptr(out), _stream())
    check(rc, "mvf_view_reduce")
    return out


def channel_mean(x):
    """The notebook's ``GRID_REAS='mean'`` (Notebook/projection.py:526-529,549): mean over the CHANNEL axis, the V per-view
    scalars become the channels, ReLU.  [B,V,X,Y,Z,C] -> [B,X,Y,Z,V]."""
    x = _cuda(x, "inputs")
    B, V, X, Y, Z, Cc = x.shape
    out = torch.empty((B, X, Y, Z, V), dtype=torch.float32, device=x.device)
    check(lib.mvf_channel_mean(_ptr(x), B, V, X * Y * Z, Cc, _ptr(out), _stream()), "mvf_channel_mean")
    return out


def convlstm_step(x, h_prev, c_prev, W, bias, forget_bias=1.0, relu_in=False):
    """One ``ConvLSTMCell.call`` (mrcnn/recurrent.py:442-479): returns (h, c)."""
    x = _cuda(x, "x")
    W = _cuda(W, "W")
    bias = _cuda(bias, "bias")
    B, X, Y, Z, Cc = x.shape
    F = W.shape[-1] // 4
    if tuple(W.shape) != (3, 3, 3, Cc + F, 4 * F):
        raise ValueError("W must be [3,3,3,C+F,4F] (recurrent.py:423-426), got %s" % (tuple(W.shape),))
    h = torch.empty((B, X, Y, Z, F), dtype=torch.float32, device=x.device)
    c = torch.empty_like(h)
    hp = _cuda(h_prev, "h_prev") if h_prev is not None else None
    cp = _cuda(c_prev, "c_prev") if c_prev is not None else None
    rc = lib.mvf_convlstm_step(_ptr(x), _ptr(hp), _ptr(cp), _ptr(W), _ptr(bias), float(forget_bias), B, X, Y, Z, Cc, F,
                               _lib.FLAG_RELU_IN if relu_in else 0, _ptr(h), _ptr(c), _stream())
    check(rc, "mvf_convlstm_step")
    return h, c


def tensor_core_eligible(C_, F):
    """The tcgen05 path of the ConvLSTM step needs 32-channel K chunks and 64-filter tiles."""
    return C_ % 32 == 0 and F % 64 == 0


class ConvLSTMTensorCore:
    """ConvLSTMCell.call (mrcnn/recurrent.py:442-479) on the tensor cores: weights are split / transposed once
    (``mvf_convlstm_prepare``), the scratch for the hi/lo halves of x and h is allocated once."""

    def __init__(self, W, bias, forget_bias=1.0):
        W = _cuda(W, "W")
        self.bias = _cuda(bias, "bias")
        self.F = W.shape[-1] // 4
        self.C = W.shape[-2] - self.F
        if tuple(W.shape) != (3, 3, 3, self.C + self.F, 4 * self.F):
            raise ValueError("W must be [3,3,3,C+F,4F] (recurrent.py:423-426), got %s" % (tuple(W.shape),))
        if not tensor_core_eligible(self.C, self.F):
            raise ValueError("tensor-core ConvLSTM needs C % 32 == 0 and F % 64 == 0")
        self.forget_bias = float(forget_bias)
        nbytes = lib.mvf_convlstm_wsplit_bytes(self.C, self.F)
        self.wsplit = torch.empty(nbytes // 4, dtype=torch.float32, device=W.device)
        check(lib.mvf_convlstm_prepare(_ptr(W), self.C, self.F, _ptr(self.wsplit), _stream()), "mvf_convlstm_prepare")
        self._ws = None

    def step(self, x, h_prev, c_prev, relu_in=False):
        x = _cuda(x, "x")
        B, X, Y, Z, Cc = x.shape
        if Cc != self.C:
            raise ValueError("x has %d channels, the weights expect %d" % (Cc, self.C))
        need = lib.mvf_convlstm_tc_workspace_bytes(B, X, Y, Z, self.C, self.F)
        if need and (self._ws is None or self._ws.numel() * 4 < need or self._ws.device != x.device):
            self._ws = torch.empty(need // 4, dtype=torch.float32, device=x.device)
        h = torch.empty((B, X, Y, Z, self.F), dtype=torch.float32, device=x.device)
        c = torch.empty_like(h)
        hp = _cuda(h_prev, "h_prev") if h_prev is not None else None
        cp = _cuda(c_prev, "c_prev") if c_prev is not None else None
        rc = lib.mvf_convlstm_step_tc(_ptr(x), _ptr(hp), _ptr(cp), _ptr(self.wsplit), _ptr(self.bias), self.forget_bias,
                                      B, X, Y, Z, self.C, self.F, _lib.FLAG_RELU_IN if relu_in else 0,
                                      _ptr(h), _ptr(c), _ptr(self._ws) if need else None, self._ws.numel() * 4 if need else 0, _stream())
        check(rc, "mvf_convlstm_step_tc")
        return h, c


    def step_slab(self, x, h_prev, c_prev, halo, relu_in=False, h_out=None, act_amax=None):
        """Slab form (mvf_convlstm_step_tc_slab): ``x`` [B,lo+Xs+hi,Y,Z,C] and ``h_prev`` / returned ``h`` [B,lo+Xs+hi,Y,Z,F]
        carry the halo planes ``halo = (lo, hi)``; ``c_prev`` / returned ``c`` are [B,Xs,Y,Z,F].  Only the interior planes
        of ``h`` are written; the caller fills the halo planes (dist.exchange_halo).  ``act_amax``: 1-element device tensor
        holding max(|relu?(x)|, |h_prev|) over the WHOLE grid (all slabs), so that every slab splits its operands with
        the same power-of-two scale."""
        x = _cuda(x, "x")
        lo, hi = int(halo[0]), int(halo[1])
        B, Xin, Y, Z, Cc = x.shape
        Xs = Xin - lo - hi
        if Cc != self.C or Xs <= 0:
            raise ValueError("bad slab: x %s, halo %s, C %d" % (tuple(x.shape), (lo, hi), self.C))
        need = lib.mvf_convlstm_tc_workspace_bytes(B, Xin, Y, Z, self.C, self.F)
        if need and (self._ws is None or self._ws.numel() * 4 < need or self._ws.device != x.device):
            self._ws = torch.empty(need // 4, dtype=torch.float32, device=x.device)
        h = h_out if h_out is not None else torch.zeros((B, Xin, Y, Z, self.F), dtype=torch.float32, device=x.device)
        c = torch.empty((B, Xs, Y, Z, self.F), dtype=torch.float32, device=x.device)
        hp = _cuda(h_prev, "h_prev") if h_prev is not None else None
        cp = _cuda(c_prev, "c_prev") if c_prev is not None else None
        rc = lib.mvf_convlstm_step_tc_slab(_ptr(x), _ptr(hp), _ptr(cp), _ptr(self.wsplit), _ptr(self.bias), self.forget_bias,
                                           B, Xs, Y, Z, self.C, self.F, lo, hi, _lib.FLAG_RELU_IN if relu_in else 0,
                                           _ptr(h), _ptr(c), _ptr(self._ws) if need else None, self._ws.numel() * 4 if need else 0,
                                           _ptr(act_amax), _stream())
        check(rc, "mvf_convlstm_step_tc_slab")
        return h, c


_cells = {}      # one prepared cell per layer name, like the reference's `reused_lay` (model_multi.py:112-117)


def _cached_cell(name, p):
    key = (name, p["W"].data_ptr(), p["W"]._version, p["b"].data_ptr(), p["b"]._version)
    hit = _cells.get(name)
    if hit is None or hit[0] != key:
        hit = (key, ConvLSTMTensorCore(p["W"], p["b"], 1.0))
        _cells[name] = hit
    return hit[1]


class IdentTensorCore:
    """grid_reas 'ident' (model_multi.py:443-455) on the tensor cores: the 1x1x1 conv over the view-concatenated
    channels is a [N, V*C] x [V*C, Cout] GEMM; weights are split / transposed once (``mvf_ident_prepare``)."""

    def __init__(self, weight, V, C_, Cout):
        weight = _cuda(weight, "weight").reshape(V * C_, Cout)
        self.V, self.C, self.Cout = V, C_, Cout
        self.wsplit = torch.empty(lib.mvf_ident_wsplit_bytes(V, C_, Cout) // 4, dtype=torch.float32, device=weight.device)
        check(lib.mvf_ident_prepare(_ptr(weight), V, C_, Cout, _ptr(self.wsplit), _stream()), "mvf_ident_prepare")
        self._ws = None

    def __call__(self, x, bias, scale, shift):
        B, V, X, Y, Z, Cc = x.shape
        need = lib.mvf_ident_tc_workspace_bytes(B, V, X, Y, Z, Cc, self.Cout)
        if need and (self._ws is None or self._ws.numel() * 4 < need or self._ws.device != x.device):
            self._ws = torch.empty(need // 4, dtype=torch.float32, device=x.device)
        out = torch.empty((B, X, Y, Z, self.Cout), dtype=torch.float32, device=x.device)
        rc = lib.mvf_ident_fuse_tc(_ptr(x), _ptr(self.wsplit), _ptr(bias), _ptr(scale), _ptr(shift), B, V, X, Y, Z, Cc,
                                   self.Cout, _ptr(out), _ptr(self._ws) if need else None, self._ws.numel() * 4 if need else 0, _stream())
        check(rc, "mvf_ident_fuse_tc")
        return out


_idents = {}


class Conv3dTensorCore:
    """One plain 3-D convolution of the fusion path on the tensor cores (``mvf_conv3d_tc``): Conv3D k=1|3 stride 1,
    Conv3D k=3 stride 2, or Conv3DTranspose k=3 stride 2, all 'same' padded, with bias -> BN -> ReLU in the epilogue
    (model_multi.py:415-441, :443-455, :472-480).  ``W`` is the Keras kernel; it is split / transposed once."""

    KINDS = {"conv": _lib.CONV_S1, "conv_s2": _lib.CONV_S2, "deconv_s2": _lib.DECONV_S2}

    def __init__(self, W, bias, kind, V=1, C2=0, bn=None, chan_interleave=0):
        W = _cuda(W, "W")
        self.kind = self.KINDS[kind]
        if W.dim() == 2:
            W = W.reshape((1, 1, 1) + tuple(W.shape))
        self.ksize = int(W.shape[0])
        if self.kind == _lib.DECONV_S2:
            self.Cout, Cin = int(W.shape[3]), int(W.shape[4])
        else:
            Cin, self.Cout = int(W.shape[3]), int(W.shape[4])
        if (Cin - C2) % V:
            raise ValueError("input channels %d - %d do not split into %d views" % (Cin, C2, V))
        self.V, self.C, self.C2, self.Cin = V, (Cin - C2) // V, C2, Cin
        self.bias = _cuda(bias, "bias")
        self.scale, self.shift = _bn_affine(bn, self.Cout, W.device)
        nbytes = lib.mvf_conv3d_wsplit_bytes(self.kind, self.ksize, Cin, self.Cout)
        self.wsplit = torch.empty(max(nbytes, 4) // 4, dtype=torch.float32, device=W.device)
        check(lib.mvf_conv3d_prepare(_ptr(W), self.kind, self.ksize, self.V, self.C, self.C2, self.Cout, int(chan_interleave),
                                     _ptr(self.wsplit), _stream()), "mvf_conv3d_prepare")
        self._ws = None

    def out_dims(self, X, Y, Z):
        if self.kind == _lib.CONV_S2:
            return X // 2, Y // 2, Z // 2
        if self.kind == _lib.DECONV_S2:
            return 2 * X, 2 * Y, 2 * Z
        return X, Y, Z

    def workspace(self, B, X, Y, Z, device):
        """The device scratch of a call on a [B,V,X,Y,Z,C] input (None when the library splits inside the GEMM)."""
        need = lib.mvf_conv3d_tc_workspace_bytes(self.kind, self.ksize, B, self.V, X, Y, Z, self.C, self.C2, self.Cout)
        if need and (self._ws is None or self._ws.numel() * 4 < need or self._ws.device != device):
            self._ws = torch.empty(need // 4, dtype=torch.float32, device=device)
        return self._ws if need else None

    def presplit_workspace(self, B, X, Y, Z, device):
        """Scratch for operand halves written by ``mvf_unproject_split_f16``: hi + lo fp16 of [B,V,X,Y,Z,C] + the scale cell."""
        need = 4 * B * self.V * X * Y * Z * self.C + 256
        if self._ws is None or self._ws.numel() * 4 < need or self._ws.device != device:
            self._ws = torch.empty((need + 3) // 4, dtype=torch.float32, device=device)
        return self._ws

    def call_presplit(self, B, X, Y, Z, relu_out=True):
        """Run on operand halves already written into ``self.presplit_workspace(...)`` by ``mvf_unproject_split_f16``."""
        ws = self.presplit_workspace(B, X, Y, Z, self.bias.device)
        OX, OY, OZ = self.out_dims(X, Y, Z)
        out = torch.empty((B, OX, OY, OZ, self.Cout), dtype=torch.float32, device=self.bias.device)
        flags = _lib.FLAG_PRESPLIT | (_lib.FLAG_RELU_OUT if relu_out else 0)
        rc = lib.mvf_conv3d_tc(None, None, _ptr(self.wsplit), _ptr(self.bias), _ptr(self.scale), _ptr(self.shift), None, None,
                               self.kind, self.ksize, B, self.V, X, Y, Z, self.C, self.C2, self.Cout, flags,
                               _ptr(out), _ptr(ws), ws.numel() * 4, None, _stream())
        check(rc, "mvf_conv3d_tc")
        return out

    def __call__(self, x, x2=None, relu_in=False, relu_out=True, pre=None, act_amax=None):
        """x [B,V,X,Y,Z,C] (or [B,X,Y,Z,C] when V == 1); x2 [B,X,Y,Z,C2] appended on channels; ``pre`` = (scale, shift)
        per input channel [V*C] applied before the conv (a depthwise 1x1); ``act_amax``: 1-element device tensor bounding
        max|x|, |x2| (skips the max-reduction pass of the fp16 operand split)."""
        x = _cuda(x, "x")
        if x.dim() == 5:
            x = x.unsqueeze(1)
        B, V, X, Y, Z, Cc = x.shape
        if V != self.V or Cc != self.C or (x2 is None) != (self.C2 == 0):
            raise ValueError("input %s does not match the prepared weights (V=%d, C=%d, C2=%d)" % (tuple(x.shape), self.V, self.C, self.C2))
        if x2 is not None:
            x2 = _cuda(x2, "x2")
            if tuple(x2.shape) != (B, X, Y, Z, self.C2):
                raise ValueError("x2 must be %s" % ((B, X, Y, Z, self.C2),))
        # 0 when the library fuses the hi/lo split into the GEMM (large 1x1x1 convolutions)
        need = lib.mvf_conv3d_tc_workspace_bytes(self.kind, self.ksize, B, V, X, Y, Z, Cc, self.C2, self.Cout)
        if need and (self._ws is None or self._ws.numel() * 4 < need or self._ws.device != x.device):
            self._ws = torch.empty(need // 4, dtype=torch.float32, device=x.device)
        OX, OY, OZ = self.out_dims(X, Y, Z)
        out = torch.empty((B, OX, OY, OZ, self.Cout), dtype=torch.float32, device=x.device)
        ps = psh = None
        if pre is not None:
            ps, psh = _cuda(pre[0], "pre_scale"), _cuda(pre[1], "pre_shift")
        flags = (_lib.FLAG_RELU_IN if relu_in else 0) | (_lib.FLAG_RELU_OUT if relu_out else 0)
        rc = lib.mvf_conv3d_tc(_ptr(x), _ptr(x2), _ptr(self.wsplit), _ptr(self.bias), _ptr(self.scale), _ptr(self.shift),
                               _ptr(ps), _ptr(psh), self.kind, self.ksize, B, V, X, Y, Z, Cc, self.C2, self.Cout, flags,
                               _ptr(out), _ptr(self._ws) if need else None, self._ws.numel() * 4 if need else 0,
                               _ptr(act_amax), _stream())
        check(rc, "mvf_conv3d_tc")
        return out


_convs = {}


def _cached_conv(name, p, kind, **kw):
    W = p["W"]
    key = (W.data_ptr() if isinstance(W, torch.Tensor) else id(W), getattr(W, "_version", 0), kind, tuple(sorted(kw.items())))
    hit = _convs.get(name)
    if hit is None or hit[0] != key:
        hit = (key, Conv3dTensorCore(p["W"], p["b"], kind, bn=p.get("bn", _default_bn(int(np.prod(p["b"].shape)))), **kw))
        _convs[name] = hit
    return hit[1]


def unet_fuse(x, scope, config, params, act_amax=None, conv1_out=None):
    """``GRID_REAS='conv3d'`` (model_multi.py:406-441): the MLF U-Net over the view-concatenated grids, four tensor-core
    convolutions; the per-view grids [B,V,X,Y,Z,C] are consumed in place (no transpose / reshape / concat copies).
    ``conv1_out``: the first convolution's output when the caller already produced it (``unproject_unet_fuse``)."""
    name = scope + "_3D_conv"
    if conv1_out is None:
        x = _cuda(x, "inputs")
        B, V, X, Y, Z, Cc = x.shape
        if X % 4 or Y % 4 or Z % 4:
            raise ValueError("the conv3d U-Net halves the grid twice: nvox and nvox_z must be multiples of 4")
        conv1 = _cached_conv(name + "_1", params["conv1"], "conv_s2", V=V)(x, relu_in=True, act_amax=act_amax)   # :415-421
    else:
        conv1 = conv1_out
    conv2 = _cached_conv(name + "_2", params["conv2"], "conv_s2")(conv1)                                # :423-428
    deconv1 = _cached_conv(name + "_deconv_1", params["deconv1"], "deconv_s2")(conv2)                   # :430-436
    C2 = conv1.shape[-1]
    return _cached_conv(name + "_deconv_2", params["deconv2"], "deconv_s2", C2=C2)(deconv1, x2=conv1)   # :437-441


def depth_sampling_conv3d(x, name, params):
    """``depth_sampling`` 'conv3d' branch (model_multi.py:467-480): the [B,S,P,P,C] ray slices are S sources of C channels
    for a 1x1 conv over C*S inputs (the reference's channel order c*S+s is folded into the prepared weights), each conv
    preceded by its depthwise 1x1 (a per-channel affine applied while the operand is split)."""
    x = _cuda(x, "x")
    B, S, P1, P2, Cc = x.shape
    dev = x.device
    dw1w = _cuda(params["dw1"]["w"], "dw1.w").reshape(Cc, S).t().contiguous().reshape(-1)      # (c*S+s) -> (s*C+c)
    dw1b = _cuda(params["dw1"]["b"], "dw1.b").reshape(Cc, S).t().contiguous().reshape(-1)
    c1 = _cached_conv(name + "2DConv_1", params["conv1"], "conv", V=S, chan_interleave=S)
    y = c1(x.reshape(B, S, 1, P1, P2, Cc), pre=(dw1w, dw1b))                                     # :472-475
    c2 = _cached_conv(name + "2DConv_2", params["conv2"], "conv")
    y = c2(y, pre=(_cuda(params["dw2"]["w"], "dw2.w"), _cuda(params["dw2"]["b"], "dw2.b")))       # :477-480
    return y.reshape(B, P1, P2, -1)


def convlstm(grid, name, kernel=(3, 3, 3), filters=32, params=None, relu_in=False, tensor_cores=None):
    """``convlstm(grid, name, kernel, filters)`` (model_multi.py:109-123): ConvRNN3D over the view
    axis with zero initial state, last output only.  Weights: ``weights[name] = {'W','b'}``."""
    grid = _cuda(grid, "grid")
    if tuple(kernel) != (3, 3, 3):
        raise ValueError("only the reference's 3x3x3 kernel is built")
    p = params if params is not None else weights.get(name)
    if p is None:
        raise ValueError("no weights registered for ConvLSTM %r (set_weights(name, W=..., b=...))" % name)
    if grid.shape[-1] != filters:
        raise ValueError("initial state takes the input's channel count, so C must equal filters (recurrent.py:145-147)")
    V = grid.shape[1]
    h = c = None
    F = p["W"].shape[-1] // 4
    use_tc = tensor_core_eligible(grid.shape[-1], F) if tensor_cores is None else bool(tensor_cores)
    if use_tc:                                  # tcgen05 implicit GEMM (3xTF32); same maths, ~1e-6 relative
        cell = _cached_cell(name, p)
        for t in range(V):
            h, c = cell.step(grid[:, t].contiguous(), h, c, relu_in=relu_in)
        return h
    for t in range(V):
        h, c = convlstm_step(grid[:, t], h, c, p["W"], p["b"], 1.0, relu_in=relu_in)
    return h


def grid_reas(inputs, scope, config, kernel=(3, 3, 3), params=None, tensor_cores=None, act_amax=None):
    """``grid_reas(inputs, scope, config)`` (model_multi.py:394-463) on a materialised
    [B,V,X,Y,Z,C] tensor; modes add | mean | max | ident | lstm3d."""
    x = _cuda(inputs, "inputs")
    mode = config.GRID_REAS
    p = params if params is not None else weights.get(scope, {})
    Cc = x.shape[-1]
    if mode == "add":
        return view_reduce(x, "sum", bn=p.get("bn", _default_bn(Cc)), relu_out=True)
    if mode in ("mean", "max"):
        has_bn = "bn" in p
        return view_reduce(x, mode, bn=p.get("bn"), relu_out=has_bn)
    if mode == "ident":
        B, V, X, Y, Z, _ = x.shape
        Wt = _cuda(p["weight"], "weight")
        Cout = Wt.shape[-1]
        bias = _cuda(p["bias"], "bias")
        scale, shift = _bn_affine(p.get("bn", _default_bn(Cout)), Cout, x.device)
        use_tc = (Cc % 32 == 0 and Cout % 16 == 0) if tensor_cores is None else bool(tensor_cores)
        if use_tc:                              # tcgen05 GEMM (3xTF32)
            key = (Wt.data_ptr(), Wt._version, V, Cc, Cout)
            hit = _idents.get(scope)
            if hit is None or hit[0] != key:
                hit = (key, IdentTensorCore(Wt, V, Cc, Cout))
                _idents[scope] = hit
            return hit[1](x, bias, scale, shift)
        out = torch.empty((B, X, Y, Z, Cout), dtype=torch.float32, device=x.device)
        rc = lib.mvf_ident_fuse(_ptr(x), _ptr(Wt.reshape(V * Cc, Cout)), _ptr(bias), _ptr(scale), _ptr(shift),
                                B, V, X * Y * Z, Cc, Cout, _ptr(out), _stream())
        check(rc, "mvf_ident_fuse")
        return out
    if mode == "conv3d":
        return unet_fuse(x, scope, config, p, act_amax=act_amax)
    if mode == "lstm3d":
        h = convlstm(x, scope + "_convlstm3d", kernel=kernel, filters=config.TOP_DOWN_PYRAMID_SIZE,
                     params={"W": p["W"], "b": p["b"]} if "W" in p else None, relu_in=True)
        scale, shift = _bn_affine(p.get("bn", _default_bn(h.shape[-1])), h.shape[-1], h.device)
        return _affine_relu(h, scale, shift)
    raise ValueError("GRID_REAS=%r is not built (hot path: add, mean, max, ident, conv3d, lstm3d)" % (mode,))


def _affine_relu(h, scale, shift):
    """BN affine + ReLU on [B,X,Y,Z,F] via the view-reduce kernel with V=1 (one launch, no torch math)."""
    B = h.shape[0]
    Cc = h.shape[-1]
    out = torch.empty_like(h)
    N = h.numel() // (B * Cc)
    rc = lib.mvf_view_reduce(_ptr(h), B, 1, N, Cc, _lib.FUSE_SUM, _lib.FLAG_RELU_OUT, _ptr(scale), _ptr(shift), _ptr(out), _stream())
    check(rc, "mvf_view_reduce")
    return out


# ------------------------------------------------------------------------------------------------
def _as_hw(proj_size):
    if isinstance(proj_size, (tuple, list)):
        return int(proj_size[0]), int(proj_size[1])
    return int(proj_size), int(proj_size)


def _proj_common(grid, Rcam, Kmat, config, proj_size, view, x_slab, grid_pos):
    grid = _cuda(grid, "grid")
    Rcam = _cuda(Rcam, "Rcam")
    Kmat = _cuda(Kmat, "Kmat")
    if grid.dim() != 5 or Rcam.dim() != 4:
        raise ValueError("expected grid [B,X,Y,Z,C], Rcam [B,V,3,4]")
    B, Xs, Y, Z, Cc = grid.shape
    g = grid_from_config(config)
    xb, xc = (0, g.nvox) if x_slab is None else (int(x_slab[0]), int(x_slab[1]))
    if (Xs, Y, Z) != (xc, g.nvox, g.nvox_z):
        raise ValueError("grid %s does not match config/slab (%d,%d,%d)" % (tuple(grid.shape), xc, g.nvox, g.nvox_z))
    Rview = Rcam[:, view].contiguous()
    Rmain = Rcam[:, 0].contiguous() if view != 0 else None
    ph, pw = _as_hw(proj_size)
    flags = _lib.FLAG_WORLD_GRID if grid_pos is not None else 0
    gp = _cuda(grid_pos, "grid_pos") if grid_pos is not None else None
    grid_dist = float(getattr(config, "GRID_DIST", 600 / 320 * config.vmax))
    return grid, Rview, Rmain, Kmat, gp, g, B, Cc, ph, pw, flags, grid_dist, xb, xc


def proj_grid(inputs, config, proj_size, view=0, x_slab=None, return_aux=False, out=None):
    """``proj_grid([grid, Rcam, Kmat], config, proj_size)`` -> [B,S,P,P,C]  (model_multi.py:231-322).
    The notebook variant takes ``[grid, grid_pos, Rcam, Kmat]`` (Notebook/projection.py:253).
    ``proj_size`` may be (ph, pw); ``view`` picks the camera the rays belong to (reference: 0)."""
    if len(inputs) == 4:
        grid, grid_pos, Rcam, Kmat = inputs
    else:
        (grid, Rcam, Kmat), grid_pos = inputs, None
    grid, Rview, Rmain, Kmat, gp, g, B, Cc, ph, pw, flags, gd, xb, xc = _proj_common(
        grid, Rcam, Kmat, config, proj_size, view, x_slab, grid_pos)
    S = int(config.samples)
    shape = (B, S, ph, pw, Cc)
    if out is None:
        out = torch.empty(shape, dtype=torch.float32, device=grid.device)
    elif tuple(out.shape) != shape or not out.is_contiguous():
        raise ValueError("out must be a contiguous %s tensor" % (shape,))
    vox = valid = None
    if return_aux:
        vox = torch.empty((B, S, ph, pw, 3), dtype=torch.int32, device=grid.device)
        valid = torch.empty((B, S, ph, pw), dtype=torch.uint8, device=grid.device)
    rc = lib.mvf_project_rays(_ptr(grid), _ptr(Rview), _ptr(Rmain), _ptr(Kmat), _ptr(gp), C.byref(g), B, Cc,
                              _image_hw(config)[0], ph, pw, S, flags, gd, xb, xc, _ptr(out), _ptr(vox), _ptr(valid), _stream())
    check(rc, "mvf_project_rays")
    return (out, vox, valid) if return_aux else out


def _depth_params(name, params, S, device):
    p = params if params is not None else weights.get(name)
    if p is None:
        raise ValueError("no weights registered for depth_sampling %r (set_weights(name, weight=[S], bias=..., bn=...))" % name)
    if isinstance(p, PreparedDepth):
        return p.w, p.bias, p.inv, p.shift
    w = torch.as_tensor(p["weight"], dtype=torch.float32, device=device).reshape(-1).contiguous()
    if w.numel() != S:
        raise ValueError("depth conv weight must have S=%d entries" % S)
    bias = float(p.get("bias", 0.0))
    gamma, beta, mean, var = (float(np.asarray(a).reshape(-1)[0]) for a in (p.get("bn") or (1.0, 0.0, 0.0, 1.0)))
    inv = np.float32(np.float32(1.0) / np.sqrt(np.float32(var) + np.float32(BN_EPS))) * np.float32(gamma)
    shift = np.float32(beta) - np.float32(mean) * inv
    return w, bias, float(inv), float(shift)


def depth_sampling(x, config, name, params=None):
    """``depth_sampling(x, config, name)`` non-conv3d branch: [B,S,P,P,C] -> [B,P,P,C]
    (model_multi.py:481-487).  Weights: ``{'weight' [S], 'bias', 'bn' (4 scalars)}``."""
    if config.GRID_REAS == "conv3d":
        p = params if params is not None else weights.get(name)
        if p is None:
            raise ValueError("no weights registered for depth_sampling %r" % name)
        return depth_sampling_conv3d(x, name, p)
    x = _cuda(x, "x")
    B, S, P1, P2, Cc = x.shape
    w, bias, inv, shift = _depth_params(name, params, S, x.device)
    out = torch.empty((B, P1, P2, Cc), dtype=torch.float32, device=x.device)
    rc = lib.mvf_depth_collapse(_ptr(x), B, S, P1 * P2, Cc, _ptr(w), bias, inv, shift, _lib.FLAG_RELU_OUT, _ptr(out), _stream())
    check(rc, "mvf_depth_collapse")
    return out


def depth_affine_relu(x, depth):
    """ReLU(BN_scalar(x + bias)): the non-linear tail of depth_sampling (model_multi.py:483-487) on an already collapsed
    [B,P,P,C] tensor (mvf_depth_collapse with S = 1, weight 1)."""
    x = _cuda(x, "x")
    B, P1, P2, Cc = x.shape
    one = torch.ones(1, dtype=torch.float32, device=x.device)
    p = dict(depth)
    p["weight"] = one
    w, bias, inv, shift = _depth_params("depth", p, 1, x.device)
    out = torch.empty_like(x)
    rc = lib.mvf_depth_collapse(_ptr(x), B, 1, P1 * P2, Cc, _ptr(w), bias, inv, shift, _lib.FLAG_RELU_OUT, _ptr(out), _stream())
    check(rc, "mvf_depth_collapse")
    return out


def proj_grid_depth_sampling(inputs, config, proj_size, name, params=None, view=0, x_slab=None, linear=False):
    """Fused ``depth_sampling(proj_grid(...))``: the ray slices [B,S,P,P,C] are never written.
    ``linear``: only the linear part ``sum_s w_s * sample_s`` (no bias, BatchNorm or ReLU) -- what one grid slab contributes;
    slabs add up and ``depth_affine_relu`` finishes the layer after the cross-rank sum (dist.slab_owner)."""
    grid, Rcam, Kmat = inputs
    grid, Rview, Rmain, Kmat, gp, g, B, Cc, ph, pw, flags, gd, xb, xc = _proj_common(
        grid, Rcam, Kmat, config, proj_size, view, x_slab, None)
    S = int(config.samples)
    w, bias, inv, shift = _depth_params(name, params, S, grid.device)
    if linear:
        bias, inv, shift = 0.0, 1.0, 0.0
    out = torch.empty((B, ph, pw, Cc), dtype=torch.float32, device=grid.device)
    rc = lib.mvf_project_depth_collapse(_ptr(grid), _ptr(Rview), _ptr(Rmain), _ptr(Kmat), None, C.byref(g), B, Cc,
                                        _image_hw(config)[0], ph, pw, S, flags | (0 if linear else _lib.FLAG_RELU_OUT), gd, xb, xc,
                                        _ptr(w), bias, inv, shift, _ptr(out), _stream())
    check(rc, "mvf_project_depth_collapse")
    return out


def unproject_fuse_project(feats, Rcam, Kmat, config, proj_size, mode="sum", bn=None, relu_out=False,
                           grid_out=None, out=None, tensor_cores=None):
    """The fused pipeline: unproj_feat -> grid_reas(sum|mean|max [+BN+ReLU]) -> proj_grid; returns (ray slices [B,S,P,P,C],
    fused grid [B,X,Y,Z,C]).  One C call (``mvf_unproject_fuse_project``): K1T with its feature split running under it (or the slot
    kernel for the modes K1T does not take), then the projection reading the main-view poses in place -- same kernels, same bits
    as ``unproject_fuse`` followed by ``proj_grid``.
    ``tensor_cores=False`` forces the two plain calls on the slot kernel."""
    if tensor_cores is False:
        fused = unproject_fuse(feats, Rcam, Kmat, config, mode=mode, bn=bn, relu_out=relu_out, out=grid_out, tensor_cores=False)
        return proj_grid([fused, Rcam, Kmat], config, proj_size, out=out), fused
    feats, Rcam, Kmat = _cuda(feats, "feats"), _cuda(Rcam, "Rcam"), _cuda(Kmat, "Kmat")
    if feats.dim() != 5 or Rcam.dim() != 4 or Kmat.dim() != 3:
        raise ValueError("expected feats [B,V,fh,fw,C], Rcam [B,V,3,4], Kmat [B,3,3]")
    B, V, fh, fw, Cc = feats.shape
    if tuple(Rcam.shape) != (B, V, 3, 4) or tuple(Kmat.shape) != (B, 3, 3):
        raise ValueError("Rcam %s / Kmat %s do not match feats %s" % (tuple(Rcam.shape), tuple(Kmat.shape), tuple(feats.shape)))
    g = grid_from_config(config)
    ph, pw = _as_hw(proj_size)
    S = int(config.samples)
    m = _FUSE[mode]
    if m == _lib.FUSE_NONE:
        raise ValueError("unproject_fuse_project needs a view reduction (sum | mean | max)")
    gshape, rshape = (B, g.nvox, g.nvox, g.nvox_z, Cc), (B, S, ph, pw, Cc)
    if grid_out is None:
        grid_out = torch.empty(gshape, dtype=torch.float32, device=feats.device)
    elif tuple(grid_out.shape) != gshape or not grid_out.is_contiguous():
        raise ValueError("grid_out must be a contiguous %s tensor" % (gshape,))
    if out is None:
        out = torch.empty(rshape, dtype=torch.float32, device=feats.device)
    elif tuple(out.shape) != rshape or not out.is_contiguous():
        raise ValueError("out must be a contiguous %s tensor" % (rshape,))
    scale, shift = _bn_affine(bn, Cc, feats.device)
    flags = _lib.FLAG_RELU_OUT if relu_out else 0
    if tensor_cores and not lib.mvf_unproject_fuse_tc_supported(V, Cc, m, flags):
        raise ValueError("the tensor-core unprojection needs mode sum/mean, C % 64 == 0 and C <= 256")
    ws = _k1t_workspace(feats.device, lib.mvf_unproject_fuse_project_workspace_bytes(B, V, fh, fw, Cc))
    ih, iw = _image_hw(config)
    rc = lib.mvf_unproject_fuse_project(_ptr(feats), _ptr(Rcam), _ptr(Kmat), C.byref(g), B, V, fh, fw, Cc, ih, iw, m, flags,
                                        _ptr(scale), _ptr(shift), ph, pw, S, _ptr(grid_out), _ptr(out), _ptr(ws), ws.numel(), _stream())
    check(rc, "mvf_unproject_fuse_project")
    return out, grid_out


def unproject_unet_fuse(feats, Rcam, Kmat, scope, config, params):
    """``grid_reas(unproj_feat(...))`` for GRID_REAS='conv3d' (model_multi.py:2382-2392 with :406-441) without materialising the
    per-view grids: K1 writes them once, as the fp16 operand halves of the U-Net's first convolution
    (``mvf_unproject_split_f16``; the fp32 grids and the 4.3 GB split pass over them never exist at 64^3 x 8 views), scaled by
    a power of two taken from max|features| (a bound: bilinear weights are in [0,1] and sum to at most 1).
    Needs C % 64 == 0 and grid dims divisible by 4; returns None when the shapes do not qualify (caller falls back)."""
    feats, Rcam, Kmat = _cuda(feats, "feats"), _cuda(Rcam, "Rcam"), _cuda(Kmat, "Kmat")
    B, V, fh, fw, Cc = feats.shape
    g = grid_from_config(config)
    X, Z = g.nvox, g.nvox_z
    if Cc % 64 or X % 4 or Z % 4:
        return None
    conv1 = _cached_conv(scope + "_3D_conv_1", params["conv1"], "conv_s2", V=V)
    ws = conv1.presplit_workspace(B, X, X, Z, feats.device)
    bound = feats.abs().amax().reshape(1)
    ih, iw = _image_hw(config)
    rc = lib.mvf_unproject_split_f16(_ptr(feats), _ptr(Rcam), None, _ptr(Kmat), C.byref(g), B, V, fh, fw, Cc, ih, iw,
                                     _lib.FLAG_RELU_IN, 1, _ptr(bound), _ptr(ws), ws.numel() * 4, _stream())
    if rc == _lib.MVF_EUNSUPPORTED:
        return None
    check(rc, "mvf_unproject_split_f16")
    try:
        c1 = conv1.call_presplit(B, X, X, Z)
    except ValueError:                                  # the library chose the tf32 format for these channel counts
        return None
    return unet_fuse(None, scope, config, params, conv1_out=c1)


def unproject_ident_fuse(feats, Rcam, Kmat, scope, config, params):
    """``grid_reas(unproj_feat(...))`` for GRID_REAS='ident' (model_multi.py:2382-2392 with :443-455) without materialising the
    per-view grids: K1 writes ReLU(unprojected views) as the fp16 operand halves of the 1x1x1 convolution, which then runs at
    the f16 tensor rate (three MMAs per product).  Needs C % 64 == 0; returns None otherwise (caller falls back)."""
    feats, Rcam, Kmat = _cuda(feats, "feats"), _cuda(Rcam, "Rcam"), _cuda(Kmat, "Kmat")
    B, V, fh, fw, Cc = feats.shape
    if Cc % 64:
        return None
    g = grid_from_config(config)
    X, Z = g.nvox, g.nvox_z
    Wt = _cuda(params["weight"], "weight")
    Cout = Wt.shape[-1]
    p = {"W": Wt.reshape(V * Cc, Cout), "b": _cuda(params["bias"], "bias"), "bn": params.get("bn", _default_bn(Cout))}
    try:
        conv = _cached_conv(scope + "ident_conv/presplit", p, "conv", V=V, chan_interleave=-1)
    except ValueError:
        return None
    ws = conv.presplit_workspace(B, X, X, Z, feats.device)
    bound = feats.abs().amax().reshape(1)
    ih, iw = _image_hw(config)
    rc = lib.mvf_unproject_split_f16(_ptr(feats), _ptr(Rcam), None, _ptr(Kmat), C.byref(g), B, V, fh, fw, Cc, ih, iw,
                                     _lib.FLAG_RELU_IN, 0, _ptr(bound), _ptr(ws), ws.numel() * 4, _stream())
    if rc == _lib.MVF_EUNSUPPORTED:
        return None
    check(rc, "mvf_unproject_split_f16")
    return conv.call_presplit(B, X, X, Z)


def fusion_neck(feature_maps, Rcam, Kmat, config, params=None, levels=(2, 3, 4, 5, 6), proj_sizes=None):
    """The fusion neck of ``MaskRCNN.build`` (model_multi.py:2382-2410): per pyramid level
    ``unproj_feat -> grid_reas -> proj_grid -> depth_sampling``, the caller of every kernel of the path.

    feature_maps: [P2..P6], each [B,V,fh,fw,C];  returns [PG2..PG6], each [B,P,P,C'] with P = IMAGE_SHAPE[0] / stride
    (160/80/40/20/10 at 640).  ``params[name]`` holds the learnables of ``grid_reas_P<l>`` and ``grid_reas_depth_PG<l>``
    (default: the ``weights`` registry).  With ``config.VANILLA`` false PG2 and PG3 are all-zero tensors (the reference
    computes and then discards them, :2406-2410; they are not computed here).
    GRID_REAS='add' takes the fused path: K1 (unproject + sum + BN + ReLU in registers) then K3b (projection + depth
    collapse), two launches per level and neither the per-view grids nor the ray slices ever exist."""
    params = params if params is not None else weights
    ih = int(config.IMAGE_SHAPE[0])
    outs = []
    for i, (lvl, fm) in enumerate(zip(levels, feature_maps)):
        fm = _cuda(fm, "P%d" % lvl)
        P_ = int(proj_sizes[i]) if proj_sizes is not None else ih // (2 ** lvl)
        B, Cc = fm.shape[0], fm.shape[-1]
        if not getattr(config, "VANILLA", False) and lvl in (2, 3):
            z = ih // (4 if lvl == 2 else 8)                        # :2407-2410 (both extents from IMAGE_SHAPE[0])
            outs.append(torch.zeros((B, z, z, int(config.TOP_DOWN_PYRAMID_SIZE)), dtype=torch.float32, device=fm.device))
            continue
        gname, dname = "grid_reas_P%d" % lvl, "grid_reas_depth_PG%d" % lvl
        gp, dp = params.get(gname, {}), params.get(dname)
        if config.GRID_REAS == "add":
            fused = unproject_fuse(fm, Rcam, Kmat, config, mode="sum", bn=gp.get("bn", _default_bn(Cc)), relu_out=True)
            outs.append(proj_grid_depth_sampling([fused, Rcam, Kmat], config, P_, dname, params=dp))
            continue
        fused = None
        if config.GRID_REAS == "conv3d":
            fused = unproject_unet_fuse(fm, Rcam, Kmat, gname, config, gp)
        elif config.GRID_REAS == "ident" and "weight" in gp:
            fused = unproject_ident_fuse(fm, Rcam, Kmat, gname, config, gp)
        if fused is None:
            per_view = unproj_feat([fm, Rcam, Kmat], config)
            # bilinear weights are in [0,1] and sum to at most 1, so the unprojected grids are bounded by max|features|: the
            # U-Net's first conv takes its fp16 operand scale from this (tiny) reduction instead of a pass over the grids
            bound = fm.abs().amax().reshape(1) if config.GRID_REAS == "conv3d" else None
            fused = grid_reas(per_view, gname, config, params=gp, act_amax=bound)
            del per_view
        rays = proj_grid([fused, Rcam, Kmat], config, P_)
        outs.append(depth_sampling(rays, config, dname, params=dp))
    return outs


# ------------------------------------------------------------------------------------------------
class PyramidROIAlign:
    """``PyramidROIAlign(pool_shape)([boxes, image_meta] + feature_maps)`` (model_multi.py:779-885)."""

    def __init__(self, pool_shape, **kwargs):
        self.pool_shape = tuple(pool_shape)

    def __call__(self, inputs, return_levels=False):
        return self.call(inputs, return_levels)

    def call(self, inputs, return_levels=False):
        boxes = _cuda(inputs[0], "boxes")
        image_meta = inputs[1]
        maps = [_cuda(m, "feature_map") for m in inputs[2:]]
        if len(maps) != 4:
            raise ValueError("PyramidROIAlign takes the four maps P2..P5")
        # image_shape = parse_image_meta_graph(image_meta)['image_shape'][0]   (:812, meta columns 4:7)
        meta0 = image_meta[0, 4:6]
        ih, iw = (int(v) for v in (meta0.tolist() if isinstance(meta0, torch.Tensor) else meta0))
        B, R, _ = boxes.shape
        Cc = maps[0].shape[-1]
        for m in maps:
            if m.shape[0] != B or m.shape[-1] != Cc:
                raise ValueError("feature maps must share batch and channel sizes")
        ph, pw = self.pool_shape
        out = torch.empty((B, R, ph, pw, Cc), dtype=torch.float32, device=boxes.device)
        lv = torch.empty((B, R), dtype=torch.int32, device=boxes.device) if return_levels else None
        ptrs = (C.c_void_p * 4)(*[m.data_ptr() for m in maps])
        H = (C.c_int * 4)(*[m.shape[1] for m in maps])
        W = (C.c_int * 4)(*[m.shape[2] for m in maps])
        rc = lib.mvf_pyramid_roi_align(_ptr(boxes), ptrs, H, W, B, R, Cc, ih, iw, ph, pw, _ptr(out), _ptr(lv), _stream())
        check(rc, "mvf_pyramid_roi_align")
        return (out, lv) if return_levels else out

    def compute_output_shape(self, input_shape):
        return input_shape[0][:2] + self.pool_shape + (input_shape[2][-1],)


def non_max_suppression(boxes, scores, max_output_size, iou_threshold, class_ids=None):
    """``tf.image.non_max_suppression(boxes, scores, max_output_size, iou_threshold)`` ->
    (keep int32 [max_output_size] padded with -1, count).  boxes [n,4] or batched [p,n,4]."""
    boxes = _cuda(boxes, "boxes")
    scores = _cuda(scores, "scores")
    batched = boxes.dim() == 3
    b3 = boxes if batched else boxes[None]
    nprob, n, _ = b3.shape
    cls = _cuda(class_ids, "class_ids", torch.int32) if class_ids is not None else None
    keep = torch.empty((nprob, max_output_size), dtype=torch.int32, device=boxes.device)
    count = torch.empty((nprob,), dtype=torch.int32, device=boxes.device)
    nbytes = lib.mvf_nms_workspace_bytes(nprob, n)
    ws = torch.empty((nbytes,), dtype=torch.uint8, device=boxes.device)
    rc = lib.mvf_nms(_ptr(b3), _ptr(scores), _ptr(cls), nprob, n, float(iou_threshold), int(max_output_size),
                     int(max_output_size), _ptr(keep), _ptr(count), _ptr(ws), nbytes, _stream())
    check(rc, "mvf_nms")
    return (keep, count) if batched else (keep[0], count[0])


def _refine_batched(rois, probs, deltas, windows, config, return_keep=False):
    rois = _cuda(rois, "rois")
    probs = _cuda(probs, "probs")
    deltas = _cuda(deltas, "deltas")
    windows = _cuda(windows, "windows")
    B, N, K = probs.shape
    max_inst = int(config.DETECTION_MAX_INSTANCES)
    det = torch.empty((B, max_inst, 6), dtype=torch.float32, device=rois.device)
    keep = torch.empty((B, max_inst), dtype=torch.int32, device=rois.device)
    count = torch.empty((B,), dtype=torch.int32, device=rois.device)
    nbytes = lib.mvf_refine_detections_workspace_bytes(B, N)
    ws = torch.empty((nbytes,), dtype=torch.uint8, device=rois.device)
    std = (C.c_float * 4)(*[float(np.float32(v)) for v in np.asarray(config.BBOX_STD_DEV).reshape(-1)])
    min_conf = float(np.float32(config.DETECTION_MIN_CONFIDENCE)) if config.DETECTION_MIN_CONFIDENCE else 0.0
    rc = lib.mvf_refine_detections(_ptr(rois), _ptr(probs), _ptr(deltas), _ptr(windows), std, B, N, K, min_conf,
                                   float(np.float32(config.DETECTION_NMS_THRESHOLD)), max_inst, _ptr(det), _ptr(keep),
                                   _ptr(count), _ptr(ws), nbytes, _stream())
    check(rc, "mvf_refine_detections")
    return (det, keep, count) if return_keep else det


def refine_detections_graph(rois, probs, deltas, window, config, return_keep=False):
    """``refine_detections_graph(rois, probs, deltas, window, config)`` -> [max_inst, 6]
    (model_multi.py:1119-1214).  One scene: rois [N,4], probs [N,K], deltas [N,K,4], window [4]."""
    res = _refine_batched(rois[None], probs[None], deltas[None], _cuda(window, "window")[None], config, return_keep)
    if return_keep:
        return res[0][0], res[1][0], res[2][0]
    return res[0]


def norm_boxes_graph(boxes, shape):
    """``norm_boxes_graph`` (model_multi.py:3390-3405) on host values (window normalisation)."""
    h, w = np.float32(shape[0]), np.float32(shape[1])
    scale = np.array([h, w, h, w], dtype=np.float32) - np.float32(1.0)
    shift = np.array([0.0, 0.0, 1.0, 1.0], dtype=np.float32)
    return ((np.asarray(boxes, dtype=np.float32) - shift) / scale).astype(np.float32)


class DetectionLayer:
    """``DetectionLayer(config)([rois, mrcnn_class, mrcnn_bbox, image_meta])`` -> [B, max_inst, 6]
    (model_multi.py:1217-1258)."""

    def __init__(self, config=None, **kwargs):
        self.config = config

    def __call__(self, inputs):
        return self.call(inputs)

    def call(self, inputs):
        rois, mrcnn_class, mrcnn_bbox, image_meta = inputs
        meta = image_meta.detach().cpu().numpy() if isinstance(image_meta, torch.Tensor) else np.asarray(image_meta)
        image_shape = meta[0, 4:7]                                         # :1241
        windows = norm_boxes_graph(meta[:, 7:11], image_shape[:2])         # :1242
        wd = torch.as_tensor(windows, device=rois.device)
        det = _refine_batched(rois, mrcnn_class, mrcnn_bbox, wd, self.config)
        return det.reshape(self.config.BATCH_SIZE, self.config.DETECTION_MAX_INSTANCES, 6)

    def compute_output_shape(self, input_shape):
        return (None, self.config.DETECTION_MAX_INSTANCES, 6)


class ProposalLayer:
    """``ProposalLayer(proposal_count, nms_threshold, config)([rpn_probs, rpn_bbox, anchors])`` ->
    [B, proposal_count, 4]  (model_multi.py:690-767)."""

    def __init__(self, proposal_count, nms_threshold, config=None, **kwargs):
        self.config = config
        self.proposal_count = int(proposal_count)
        self.nms_threshold = float(nms_threshold)

    def __call__(self, inputs):
        return self.call(inputs)

    def call(self, inputs):
        probs = _cuda(inputs[0], "rpn_probs")
        bbox = _cuda(inputs[1], "rpn_bbox")
        anchors = _cuda(inputs[2], "anchors")
        B, A, _ = probs.shape
        out = torch.empty((B, self.proposal_count, 4), dtype=torch.float32, device=probs.device)
        count = torch.empty((B,), dtype=torch.int32, device=probs.device)
        limit = int(self.config.PRE_NMS_LIMIT)
        nbytes = lib.mvf_proposals_workspace_bytes(B, A, limit)
        ws = torch.empty((nbytes,), dtype=torch.uint8, device=probs.device)
        std = (C.c_float * 4)(*[float(np.float32(v)) for v in np.asarray(self.config.RPN_BBOX_STD_DEV).reshape(-1)])
        rc = lib.mvf_proposals(_ptr(probs), _ptr(bbox), _ptr(anchors), std, B, A, limit, self.proposal_count,
                               float(np.float32(self.nms_threshold)), _ptr(out), _ptr(count), _ptr(ws), nbytes, _stream())
        check(rc, "mvf_proposals")
        return out

    def compute_output_shape(self, input_shape):
        return (None, self.proposal_count, 4)


# ------------------------------------------------------------------------------------------------
class HostPipeline:
    """End-to-end entry over HOST buffers (the reference's only host->device crossing is
    ``keras_model.predict``, model_multi.py:3067-3068): pinned feats/Rcam/Kmat in, pinned ray slices
    out, device scratch owned by this object.  One call = H2D + K1 + K3 + D2H + stream sync."""

    def __init__(self, config, B, V, fh, fw, Cc, proj_size, mode="sum", device=None, depth=None, bn=None, relu_out=False):
        """``depth`` = the depth_sampling learnables ({'weight' [S], 'bias', 'bn'}): the call then returns one level of the
        fusion neck, PG [B,P,P,C] (``mvf_fusion_neck_level_host``), instead of the ray slices; ``bn`` / ``relu_out``: the
        grid_reas BatchNorm + ReLU fused into K1."""
        self.config = config
        self.g = grid_from_config(config)
        self.shape = (B, V, fh, fw, Cc)
        self.ph, self.pw = _as_hw(proj_size)
        self.S = int(config.samples)
        self.mode = _FUSE[mode]
        device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.nbytes = lib.mvf_pipeline_host_workspace_bytes(C.byref(self.g), B, V, fh, fw, Cc, self.ph, self.pw, self.S)
        self.ws = torch.empty((self.nbytes,), dtype=torch.uint8, device=device)
        self.h2d_bytes = 4 * (B * V * fh * fw * Cc + B * V * 12 + B * 9)
        self.depth = _depth_params("depth", depth, self.S, device) if depth is not None else None
        self.bn = _bn_affine(bn, Cc, device)
        self.flags = _lib.FLAG_RELU_OUT if relu_out else 0
        self.d2h_bytes = 4 * B * (1 if self.depth else self.S) * self.ph * self.pw * Cc
        # the three copy / compute streams and their events: owned by this object, created on `device` (mvf_host_aux_create)
        self.aux = C.c_void_p()
        with torch.cuda.device(device):
            check(lib.mvf_host_aux_create(C.byref(self.aux)), "mvf_host_aux_create")

    def __del__(self):
        aux = getattr(self, "aux", None)
        if aux:
            lib.mvf_host_aux_destroy(aux)
            self.aux = None

    def empty_output(self):
        B, _, _, _, Cc = self.shape
        shape = (B, self.ph, self.pw, Cc) if self.depth else (B, self.S, self.ph, self.pw, Cc)
        return torch.empty(shape, dtype=torch.float32).pin_memory()

    def __call__(self, h_feats, h_Rcam, h_Kmat, h_out):
        for t, n in ((h_feats, "feats"), (h_Rcam, "Rcam"), (h_Kmat, "Kmat"), (h_out, "out")):
            if t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous():
                raise ValueError("%s must be a contiguous fp32 HOST tensor" % n)
        B, V, fh, fw, Cc = self.shape
        if tuple(h_feats.shape) != self.shape:
            raise ValueError("feats shape %s != %s" % (tuple(h_feats.shape), self.shape))
        ih, iw = _image_hw(self.config)
        if self.depth:
            w, bias, inv, shift = self.depth
            rc = lib.mvf_fusion_neck_level_host(_ptr(h_feats), _ptr(h_Rcam), _ptr(h_Kmat), C.byref(self.g), B, V, fh, fw, Cc,
                                                ih, iw, self.mode, self.flags, _ptr(self.bn[0]), _ptr(self.bn[1]),
                                                self.ph, self.pw, self.S, _ptr(w), bias, inv, shift,
                                                _ptr(h_out), _ptr(self.ws), self.nbytes, self.aux, _stream())
            check(rc, "mvf_fusion_neck_level_host")
            return h_out
        rc = lib.mvf_unproject_fuse_project_host(_ptr(h_feats), _ptr(h_Rcam), _ptr(h_Kmat), C.byref(self.g), B, V, fh, fw, Cc,
                                                 ih, iw, self.mode, self.flags, _ptr(self.bn[0]), _ptr(self.bn[1]),
                                                 self.ph, self.pw, self.S, _ptr(h_out), _ptr(self.ws), self.nbytes, self.aux, _stream())
        check(rc, "mvf_unproject_fuse_project_host")
        return h_out
