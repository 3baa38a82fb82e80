"""Config c2 head kernels on one B200: PyramidROIAlign (K4), refine_detections / ProposalLayer / NMS (K5), grid_reas 'ident'
(K2b) and the max-fuse pipeline at 48^3 -- CUDA-event timings over device-resident synthetic inputs."""
import json, sys
import numpy as np, torch
sys.path.insert(0, '.')
import mulit_view_object_detection_b200 as m
from mulit_view_object_detection_b200 import synthetic as syn

dev = torch.device('cuda:0')
rng = np.random.default_rng(2000)
HBM = 6560.0
try:
    HBM = float(json.load(open('MEASURED_PEAKS.json'))['hbm_gbs'])
except Exception:
    pass


def timed(fn, n=20, warm=3, reps=10):
    """Average GPU time of one call: `reps` calls are captured in a CUDA graph (so Python/ctypes launch overhead is not
    on the clock), the graph is replayed `n` times with an L2 flush (write of a 160 MB buffer) before each replay."""
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    graph = torch.cuda.CUDAGraph()
    try:
        with torch.cuda.stream(side):
            fn()
            torch.cuda.synchronize()
            with torch.cuda.graph(graph, stream=side):
                for _ in range(reps):
                    fn()
    except Exception as e:                     # a wrapper that synchronises cannot be captured: time eagerly
        torch.cuda.synchronize()
        print("not capturable (%s): eager timing" % str(e)[:60], file=sys.stderr)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n * reps):
            fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / (n * reps)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    flush = torch.empty(160 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2
    tot = 0.0
    for _ in range(n):
        flush.zero_()
        e0.record(); graph.replay(); e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / n / reps


out = {}
B, C, img = 1, 256, 640
cfg = m.FusionConfig(nvox=48, nvox_z=48, samples=20, NUM_VIEWS=4, GRID_REAS="max", IMAGES_PER_GPU=B,
                     IMAGE_SHAPE=np.array([img, img, 3]), TOP_DOWN_PYRAMID_SIZE=C)
# ---- K4 PyramidROIAlign: 1000 x 7x7 (classifier head) and 100 x 14x14 (mask head) over P2..P5
maps = [torch.from_numpy(np.maximum(rng.standard_normal((B, img // s, img // s, C), dtype=np.float32), 0)).to(dev) for s in (4, 8, 16, 32)]
meta = syn.make_image_meta(B, (img, img, 3), 25)       # host array: the layer only reads the image shape from it
for R, pool in ((1000, 7), (100, 14)):
    boxes = torch.from_numpy(syn.make_rois(rng, B, R)).to(dev)
    layer = m.PyramidROIAlign([pool, pool])
    ms = timed(lambda: layer([boxes, meta] + maps))
    lower = 4 * C * R * pool * pool * 2          # read one vector (L2-resident maps) + write one vector per bin
    out["roi_align_%dx%dx%d" % (R, pool, pool)] = {"ms": ms, "algorithmic_bytes_lower_bound": lower,
                                                  "achieved_gbs": lower / ms / 1e6, "frac_of_hbm_peak": lower / ms / 1e6 / HBM}
# ---- K5 refine_detections_graph: 1000 rois, 25 classes
K = 25
probs, deltas = syn.make_detection_inputs(rng, 1000, K)
rois = syn.make_rois(rng, 1, 1000)[0]
d = [torch.from_numpy(a).to(dev) for a in (rois, probs, deltas)]
window = torch.tensor([0.0, 0.0, 1.0, 1.0], device=dev)
out["refine_detections_1000x25"] = {"ms": timed(lambda: m.refine_detections_graph(d[0], d[1], d[2], window, cfg))}
# ---- K5 ProposalLayer: 102300 anchors -> top 6000 -> NMS(0.7) -> 1000
anchors = syn.make_anchors((img, img))
A = anchors.shape[0]
fg = rng.permutation(A).astype(np.float32) / A
pr = np.stack([1 - fg, fg], -1)[None].astype(np.float32)
bb = rng.normal(0, 0.5, (1, A, 4)).astype(np.float32)
dp = [torch.from_numpy(a).to(dev) for a in (pr, bb, anchors[None].copy())]
prop = m.ProposalLayer(1000, 0.7, cfg)
out["proposal_layer_%d_anchors" % A] = {"ms": timed(lambda: prop(dp))}
# ---- K5 plain NMS on 6000 boxes
b6 = torch.from_numpy(syn.make_rois(rng, 1, 6000, pad_frac=0)[0]).to(dev)
s6 = torch.from_numpy(rng.permutation(6000).astype(np.float32) / 6000).to(dev)
out["nms_6000_boxes"] = {"ms": timed(lambda: m.non_max_suppression(b6, s6, 1000, 0.7))}
# ---- c2 fusion: 4 views, P4 40x40x256, 48^3, max-fuse + projection (one scene, device resident)
feats, Rcam, Kmat = syn.make_scene(cfg, 1, 4, 40, 40, C, seed=2001)
df, dR, dK = (torch.from_numpy(a).to(dev) for a in (feats, Rcam, Kmat))
ms = timed(lambda: m.unproject_fuse_project(df, dR, dK, cfg, proj_size=40, mode="max"))
alg = 4 * C * (4 * 40 * 40 + 48 ** 3 + 2 * 20 * 40 * 40)
out["c2_fusion_max_48cubed_P4"] = {"ms": ms, "voxel_samples_per_s": 4 * 48 ** 3 / ms * 1e3, "algorithmic_bytes": alg,
                                   "frac_of_hbm_peak": alg / ms / 1e6 / HBM}
# ---- c2 fusion at every pyramid level: P2..P5 maps 160/80/40/20 squared, projected at the same sizes (model_multi.py:2393-2397)
for lvl, fs in ((2, 160), (3, 80), (4, 40), (5, 20)):
    f_l, R_l, K_l = syn.make_scene(cfg, 1, 4, fs, fs, C, seed=2001)
    dl = [torch.from_numpy(a).to(dev) for a in (f_l, R_l, K_l)]
    grid_l = torch.empty((1, 48, 48, 48, C), device=dev)
    rays_l = torch.empty((1, 20, fs, fs, C), device=dev)
    ms1 = timed(lambda: m.unproject_fuse(dl[0], dl[1], dl[2], cfg, mode="max", out=grid_l), n=10)
    ms3 = timed(lambda: m.proj_grid([grid_l, dl[1], dl[2]], cfg, fs, out=rays_l), n=10)
    a1, a3 = 4 * C * (4 * fs * fs + 48 ** 3), 4 * C * 2 * 20 * fs * fs
    out["c2_fusion_max_48cubed_P%d" % lvl] = {"k1_ms": ms1, "k3_ms": ms3, "k1_frac_of_hbm_peak": a1 / ms1 / 1e6 / HBM,
                                              "k3_frac_of_hbm_peak": a3 / ms3 / 1e6 / HBM,
                                              "voxel_samples_per_s": 4 * 48 ** 3 / (ms1 + ms3) * 1e3}
    del dl, grid_l, rays_l
# ---- the whole 'add' neck (model_multi.py:2382-2410) at c2 shapes, all five levels computed (VANILLA=True)
ncfg = m.FusionConfig(nvox=48, nvox_z=48, samples=20, NUM_VIEWS=4, GRID_REAS="add", IMAGES_PER_GPU=1, VANILLA=True,
                      IMAGE_SHAPE=np.array([img, img, 3]), TOP_DOWN_PYRAMID_SIZE=C)
fm, Rn, Kn = [], None, None
for fs in (160, 80, 40, 20, 10):
    f_l, Rn, Kn = syn.make_scene(ncfg, 1, 4, fs, fs, C, seed=2001)
    fm.append(torch.from_numpy(f_l).to(dev))
Rn, Kn = torch.from_numpy(Rn).to(dev), torch.from_numpy(Kn).to(dev)
nparams = {"grid_reas_depth_PG%d" % l: {"weight": np.full(20, 0.05, np.float32), "bias": 0.0} for l in (2, 3, 4, 5, 6)}
nparams.update({"grid_reas_P%d" % l: {"bn": (np.ones(C, np.float32), np.zeros(C, np.float32), np.zeros(C, np.float32), np.ones(C, np.float32))}
                for l in (2, 3, 4, 5, 6)})
nparams = m.prepare_params(nparams, dev)
out["fusion_neck_add_P2_P6_48cubed"] = {"ms": timed(lambda: m.fusion_neck(fm, Rn, Kn, ncfg, params=nparams), n=10, reps=3),
                                        "launches": "2 per level (K1 sum+BN+ReLU, K3b projection + depth collapse)"}
# ---- K2b grid_reas 'ident': [V*C -> C] 1x1x1 conv at 4 views, 48^3
V = 4
icfg = m.FusionConfig(nvox=48, nvox_z=48, samples=20, NUM_VIEWS=V, GRID_REAS="ident", IMAGES_PER_GPU=1,
                      IMAGE_SHAPE=np.array([img, img, 3]), TOP_DOWN_PYRAMID_SIZE=C)
per_view = torch.randn((1, V, 48, 48, 48, C), device=dev).relu_()
Wi = torch.randn((V * C, C), device=dev) * 0.03
bi = torch.zeros(C, device=dev)
ms = timed(lambda: m.grid_reas(per_view, "grid_reas_P4", icfg, params={"weight": Wi, "bias": bi}), n=5)
out["ident_fuse_4x256_to_256_48cubed"] = {"ms": ms, "tflops": 2.0 * 48 ** 3 * V * C * C / ms / 1e9}
print(json.dumps(out, indent=1))
