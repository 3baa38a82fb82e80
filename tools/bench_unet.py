"""Time the conv3d family at grid_reas('conv3d') size (model_multi.py:406-441): V views x 256 channels, X^3 grid, F = 256,
one scene; per-layer CUDA-event timings (split passes included) and useful TFLOP/s; plus grid_reas('ident') and the
depth_sampling conv3d branch.  usage: python tools/bench_unet.py [X=48] [V=4] [out.json]"""
import json, sys
import torch
sys.path.insert(0, '.')
import mulit_view_object_detection_b200 as m

X = int(sys.argv[1]) if len(sys.argv) > 1 else 48
V = int(sys.argv[2]) if len(sys.argv) > 2 else 4
C = F = 256
dev = torch.device('cuda:0')
g = torch.Generator(device=dev); g.manual_seed(0)


def rnd(*shape, scale=1.0):
    return torch.randn(shape, device=dev, generator=g) * scale


def timed(fn, n=3):
    for _ in range(2):
        out = fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        out = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, out


res = {"grid": X, "views": V, "C": C, "F": F, "layers": []}
grids = rnd(1, V, X, X, X, C)
layers = [("conv1  Conv3D s2 %d->%d" % (V * C, 2 * F), "conv_s2", (3, 3, 3, V * C, 2 * F), dict(V=V), (X // 2) ** 3 * 27 * V * C * 2 * F),
          ("conv2  Conv3D s2 %d->%d" % (2 * F, 4 * F), "conv_s2", (3, 3, 3, 2 * F, 4 * F), {}, (X // 4) ** 3 * 27 * 2 * F * 4 * F),
          ("deconv1 Conv3DT s2 %d->%d" % (4 * F, 2 * F), "deconv_s2", (3, 3, 3, 2 * F, 4 * F), {}, (X // 4) ** 3 * 27 * 4 * F * 2 * F),
          ("deconv2 Conv3DT s2 %d->%d" % (4 * F, F), "deconv_s2", (3, 3, 3, F, 4 * F), dict(C2=2 * F), (X // 2) ** 3 * 27 * 4 * F * F)]
x, skip, total_ms, total_flop = grids, None, 0.0, 0.0
for i, (name, kind, wshape, kw, macs) in enumerate(layers):
    fan = wshape[3] * 27 if kind != "deconv_s2" else wshape[4] * 8
    conv = m.Conv3dTensorCore(rnd(*wshape, scale=fan ** -0.5), rnd(wshape[4] if kind != "deconv_s2" else wshape[3], scale=0.1), kind, **kw)
    if i == 3:
        ms, out = timed(lambda: conv(x, x2=skip))
    else:
        ms, out = timed(lambda: conv(x, relu_in=(i == 0)))
    if i == 0:
        skip = out
    x = out
    flop = 2.0 * macs
    total_ms += ms; total_flop += flop
    res["layers"].append({"layer": name, "ms": ms, "useful_tflops": flop / ms / 1e9, "out": list(out.shape)})
    print("%-34s %8.3f ms  %7.1f TFLOP/s useful   out %s" % (name, ms, flop / ms / 1e9, tuple(out.shape)))
res["unet_ms"] = total_ms; res["unet_useful_tflops"] = total_flop / total_ms / 1e9
print("U-Net total %.3f ms, %.1f TFLOP/s useful" % (total_ms, total_flop / total_ms / 1e9))

ident = m.Conv3dTensorCore(rnd(V * C, F, scale=(V * C) ** -0.5), rnd(F, scale=0.1), "conv", V=V)
ms, _ = timed(lambda: ident(grids, relu_in=True))
fl = 2.0 * X ** 3 * V * C * F
res["ident"] = {"ms": ms, "useful_tflops": fl / ms / 1e9, "hbm_gbs_in_plus_out": (grids.numel() + X ** 3 * F) * 4 / ms / 1e6}
print("ident 1x1x1 %d->%d: %.3f ms  %.1f TFLOP/s useful  (reads %.2f GB of per-view grids: %.0f GB/s)"
      % (V * C, F, ms, fl / ms / 1e9, grids.numel() * 4 / 1e9, res["ident"]["hbm_gbs_in_plus_out"]))

S, P = 20, 40
rays = rnd(1, S, P, P, C).relu_()
params = {"dw1": {"w": rnd(C * S) * 0.1 + 1, "b": rnd(C * S, scale=0.1)}, "conv1": {"W": rnd(C * S, 512, scale=(C * S) ** -0.5), "b": rnd(512, scale=0.1)},
          "dw2": {"w": rnd(512) * 0.1 + 1, "b": rnd(512, scale=0.1)}, "conv2": {"W": rnd(512, F, scale=512 ** -0.5), "b": rnd(F, scale=0.1)}}
ms, out = timed(lambda: m.depth_sampling_conv3d(rays, "bench_depth", params))
fl = 2.0 * P * P * (C * S * 512 + 512 * F)
res["depth_sampling_conv3d"] = {"ms": ms, "useful_tflops": fl / ms / 1e9}
print("depth_sampling conv3d branch (S=%d, P=%d): %.3f ms  %.1f TFLOP/s useful" % (S, P, ms, fl / ms / 1e9))
if len(sys.argv) > 3:
    json.dump(res, open(sys.argv[3], "w"), indent=1)

# ---- one level of the conv3d neck end to end (model_multi.py:2382-2404 with GRID_REAS='conv3d'): unproj_feat (per-view grids) ->
# U-Net -> proj_grid -> depth_sampling conv3d branch, device resident
import numpy as np
from mulit_view_object_detection_b200 import synthetic as syn
cfg = m.FusionConfig(nvox=X, nvox_z=X, samples=S, NUM_VIEWS=V, GRID_REAS="conv3d", IMAGE_SHAPE=np.array([640, 640, 3]),
                     TOP_DOWN_PYRAMID_SIZE=F, VANILLA=True)
feats, Rcam, Kmat = syn.make_scene(cfg, 1, V, P, P, C, seed=5)
d = [torch.from_numpy(a).to(dev) for a in (feats, Rcam, Kmat)]
wl = lambda *shape, fan: rnd(*shape, scale=fan ** -0.5)
nparams = {"grid_reas_P4": {"conv1": {"W": wl(3, 3, 3, V * C, 2 * F, fan=27 * V * C), "b": rnd(2 * F, scale=0.1)},
                            "conv2": {"W": wl(3, 3, 3, 2 * F, 4 * F, fan=27 * 2 * F), "b": rnd(4 * F, scale=0.1)},
                            "deconv1": {"W": wl(3, 3, 3, 2 * F, 4 * F, fan=8 * 4 * F), "b": rnd(2 * F, scale=0.1)},
                            "deconv2": {"W": wl(3, 3, 3, F, 4 * F, fan=8 * 4 * F), "b": rnd(F, scale=0.1)}},
           "grid_reas_depth_PG4": params}
ms_k1, per_view = timed(lambda: m.unproj_feat(d, cfg))
ms_neck, pg = timed(lambda: m.fusion_neck([d[0]], d[1], d[2], cfg, params=nparams, levels=(4,)))
res["k1_none_ms"] = ms_k1
res["neck_conv3d_P4_ms"] = ms_neck
print("unproj_feat (per-view grids, %.2f GB): %.3f ms = %.0f GB/s written;  conv3d neck level P4 end to end: %.3f ms  -> PG %s"
      % (per_view.numel() * 4 / 1e9, ms_k1, per_view.numel() * 4 / ms_k1 / 1e6, ms_neck, tuple(pg[0].shape)))
if len(sys.argv) > 3:
    json.dump(res, open(sys.argv[3], "w"), indent=1)

# ---- one level of the 'ident' neck: materialised (unproj_feat + 1x1x1 conv with the in-kernel tf32 converter) vs direct (K1 writes
# the conv's fp16 operand halves)
icfg = m.FusionConfig(nvox=X, nvox_z=X, samples=S, NUM_VIEWS=V, GRID_REAS="ident", IMAGE_SHAPE=np.array([640, 640, 3]),
                      TOP_DOWN_PYRAMID_SIZE=F)
ip = {"weight": rnd(V * C, F, scale=(V * C) ** -0.5), "bias": rnd(F, scale=0.1)}
ms_mat, _ = timed(lambda: m.grid_reas(m.unproj_feat(d, icfg), "bench_ident", icfg, params=ip))
ms_dir, _ = timed(lambda: m.unproject_ident_fuse(d[0], d[1], d[2], "bench_ident", icfg, ip))
res["ident_level_materialised_ms"] = ms_mat; res["ident_level_direct_ms"] = ms_dir
print("ident level (unproj_feat -> 1x1x1 conv): materialised %.3f ms, direct operand path %.3f ms" % (ms_mat, ms_dir))
if len(sys.argv) > 3:
    json.dump(res, open(sys.argv[3], "w"), indent=1)
