"""NMS / ProposalLayer / refine_detections timings only (the K5 rows of tools/bench_heads.py), CUDA-graph replays with an L2 flush."""
import json, sys
import numpy as np, torch
sys.path.insert(0, '.')
import mulit_view_object_detection_b200 as m
from mulit_view_object_detection_b200 import synthetic as syn

dev = torch.device('cuda:0')
rng = np.random.default_rng(2000)


def timed(fn, n=20, warm=3, reps=10):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        fn()
        torch.cuda.synchronize()
        with torch.cuda.graph(graph, stream=side):
            for _ in range(reps):
                fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    flush = torch.empty(160 << 20, dtype=torch.uint8, device=dev)
    tot = 0.0
    for _ in range(n):
        flush.zero_()
        e0.record(); graph.replay(); e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / n / reps


out = {}
img = 640
cfg = m.FusionConfig(nvox=48, nvox_z=48, samples=20, NUM_VIEWS=4, GRID_REAS="max", IMAGES_PER_GPU=1, IMAGE_SHAPE=np.array([img, img, 3]))
K = 25
probs, deltas = syn.make_detection_inputs(rng, 1000, K)
rois = syn.make_rois(rng, 1, 1000)[0]
d = [torch.from_numpy(a).to(dev) for a in (rois, probs, deltas)]
window = torch.tensor([0.0, 0.0, 1.0, 1.0], device=dev)
out["refine_detections_1000x25_ms"] = timed(lambda: m.refine_detections_graph(d[0], d[1], d[2], window, cfg))
anchors = syn.make_anchors((img, img))
A = anchors.shape[0]
fg = rng.permutation(A).astype(np.float32) / A
pr = np.stack([1 - fg, fg], -1)[None].astype(np.float32)
bb = rng.normal(0, 0.5, (1, A, 4)).astype(np.float32)
dp = [torch.from_numpy(a).to(dev) for a in (pr, bb, anchors[None].copy())]
prop = m.ProposalLayer(1000, 0.7, cfg)
out["proposal_layer_%d_anchors_ms" % A] = timed(lambda: prop(dp))
b6 = torch.from_numpy(syn.make_rois(rng, 1, 6000, pad_frac=0)[0]).to(dev)
s6 = torch.from_numpy(rng.permutation(6000).astype(np.float32) / 6000).to(dev)
out["nms_6000_boxes_ms"] = timed(lambda: m.non_max_suppression(b6, s6, 1000, 0.7))
# heavy overlap: the sweep has to visit every candidate
c = rng.random((6000, 2)).astype(np.float32) * 0.2 + 0.4
hw = rng.random((6000, 2)).astype(np.float32) * 0.1 + 0.3
bo = torch.from_numpy(np.concatenate([c - hw / 2, c + hw / 2], 1)).to(dev)
out["nms_6000_overlapping_boxes_ms"] = timed(lambda: m.non_max_suppression(bo, s6, 1000, 0.7))
print(json.dumps(out, indent=1))
