"""One launch each of the head kernels at config-c2 sizes (for ncu): PyramidROIAlign 1000x7x7x256, NMS of 6000 boxes,
refine_detections 1000x25.  usage: ncu --set full -k regex:'roi_align|nms_|refine' python tools/profile_heads.py"""
import sys
import numpy as np, torch
sys.path.insert(0, '.')
import mulit_view_object_detection_b200 as m
from mulit_view_object_detection_b200 import synthetic as syn

dev = torch.device('cuda:0')
rng = np.random.default_rng(2000)
B, C, img = 1, 256, 640
cfg = m.FusionConfig(nvox=48, nvox_z=48, samples=20, NUM_VIEWS=4, IMAGE_SHAPE=np.array([img, img, 3]))
maps = [torch.from_numpy(np.maximum(rng.standard_normal((B, img // s, img // s, C), dtype=np.float32), 0)).to(dev) for s in (4, 8, 16, 32)]
meta = syn.make_image_meta(B, (img, img, 3), 25)
boxes = torch.from_numpy(syn.make_rois(rng, B, 1000)).to(dev)
layer = m.PyramidROIAlign([7, 7])
b6 = torch.from_numpy(syn.make_rois(rng, 1, 6000, pad_frac=0)[0]).to(dev)
s6 = torch.from_numpy(rng.permutation(6000).astype(np.float32) / 6000).to(dev)
probs, deltas = syn.make_detection_inputs(rng, 1000, 25)
d = [torch.from_numpy(a).to(dev) for a in (syn.make_rois(rng, 1, 1000)[0], probs, deltas)]
window = torch.tensor([0.0, 0.0, 1.0, 1.0], device=dev)
for _ in range(2):
    layer([boxes, meta] + maps)
    keep, cnt = m.non_max_suppression(b6, s6, 1000, 0.7)
    m.refine_detections_graph(d[0], d[1], d[2], window, cfg)
torch.cuda.synchronize()
print("kept", int(cnt))
