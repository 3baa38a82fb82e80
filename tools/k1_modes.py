"""K1 at workload T (16 scenes) in every fusion mode: ms per scene."""
import sys, numpy as np, torch
sys.path.insert(0, '.')
import mulit_view_object_detection_b200 as m
from mulit_view_object_detection_b200 import synthetic as syn
cfg = m.FusionConfig(nvox=64, nvox_z=64, samples=20, NUM_VIEWS=8, IMAGE_SHAPE=np.array([640, 640, 3]))
B = 16
feats, Rcam, Kmat = syn.make_scene(cfg, B, 8, 40, 40, 256, seed=1000)
d = [torch.from_numpy(a).cuda() for a in (feats, Rcam, Kmat)]
grid = torch.empty((B, 64, 64, 64, 256), device='cuda')
bn = (np.ones(256, np.float32), np.zeros(256, np.float32), np.zeros(256, np.float32), np.ones(256, np.float32))
for name, kw in (("sum", dict(mode="sum")), ("sum+bn+relu", dict(mode="sum", bn=bn, relu_out=True)), ("mean", dict(mode="mean")),
                 ("max", dict(mode="max")), ("sum relu_in", dict(mode="sum", relu_in=True))):
    for _ in range(3):
        m.unproject_fuse(*d, cfg, out=grid, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        m.unproject_fuse(*d, cfg, out=grid, **kw)
    e1.record(); torch.cuda.synchronize()
    print("%-14s %.4f ms/scene" % (name, e0.elapsed_time(e1) / 5 / B))
