"""K3 / K3b at workload T (16 scenes): proj_grid (ray slices) and the fused projection + depth collapse, ms per scene and
fraction of the HBM roofline against their algorithmic bytes."""
import json, sys, numpy as np, torch
sys.path.insert(0, '.')
import mulit_view_object_detection_b200 as m
from mulit_view_object_detection_b200 import synthetic as syn
HBM = 6560.0
try:
    HBM = float(json.load(open('MEASURED_PEAKS.json'))['hbm_gbs'])
except Exception:
    pass
cfg = m.FusionConfig(nvox=64, nvox_z=64, samples=20, NUM_VIEWS=8, IMAGE_SHAPE=np.array([640, 640, 3]))
B, C, P, S = 16, 256, 40, 20
feats, Rcam, Kmat = syn.make_scene(cfg, B, 8, 40, 40, C, seed=1000)
d = [torch.from_numpy(a).cuda() for a in (feats, Rcam, Kmat)]
grid = m.unproject_fuse(*d, cfg, mode="sum")
rays = torch.empty((B, S, P, P, C), device='cuda')
depth = m.prepare_params({"grid_reas_depth_PG4": {"weight": np.full(S, 0.05, np.float32), "bias": 0.0}})["grid_reas_depth_PG4"]
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')


def timed(fn, n=10):
    for _ in range(3):
        fn()
    tot = 0.0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(n):
        flush.zero_()                    # the grid must come from HBM, not from a warm L2
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / n / B


ms3 = timed(lambda: m.proj_grid([grid, d[1], d[2]], cfg, P, out=rays))
ms3b = timed(lambda: m.proj_grid_depth_sampling([grid, d[1], d[2]], cfg, P, "grid_reas_depth_PG4", params=depth))
a3, a3b = 4 * C * 2 * S * P * P, 4 * C * (S * P * P + P * P)
print("K3  proj_grid                     %.4f ms/scene  %.0f GB/s algorithmic = %.2f of HBM peak" % (ms3, a3 / ms3 / 1e6, a3 / ms3 / 1e6 / HBM))
print("K3b proj_grid + depth collapse    %.4f ms/scene  %.0f GB/s algorithmic = %.2f of HBM peak" % (ms3b, a3b / ms3b / 1e6, a3b / ms3b / 1e6 / HBM))
