import sys
sys.path.insert(0, '/root/repo')
import torch, bench
import mulit_view_object_detection_b200 as m
from mulit_view_object_detection_b200 import synthetic as syn
T = bench.T; B = 16
cfg = bench.make_config(B)
feats, Rcam, Kmat = syn.make_scene(cfg, B, T["V"], T["fh"], T["fw"], T["C"], seed=1000)
d = [torch.from_numpy(a).cuda() for a in (feats, Rcam, Kmat)]
X = T["nvox"]
grid = torch.empty((B, X, X, X, T["C"]), dtype=torch.float32, device='cuda')
rays = torch.empty((B, T["S"], T["P"], T["P"], T["C"]), dtype=torch.float32, device='cuda')
def timed(fn, n=20):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
mode = sys.argv[1] if len(sys.argv) > 1 else "time"
def run_all(tag):
    print(tag, "k1t alone ms", timed(lambda: m.unproject_fuse(*d, cfg, mode="sum", out=grid)))
    print(tag, "k3 alone ms", timed(lambda: m.proj_grid([grid, d[1], d[2]], cfg, T["P"], out=rays)))
    print(tag, "fused ms", timed(lambda: m.unproject_fuse_project(*d, cfg, T["P"], mode="sum", grid_out=grid, out=rays)))
if mode == "time":
    run_all("legacy-stream")
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        def timed(fn, n=20):
            for _ in range(5): fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            for _ in range(n): fn()
            e1.record(st); torch.cuda.synchronize()
            return e0.elapsed_time(e1) / n
        run_all("side-stream")
if mode == "time0":
    print("k1t alone ms", timed(lambda: m.unproject_fuse(*d, cfg, mode="sum", out=grid)))
    print("k3 alone ms", timed(lambda: m.proj_grid([grid, d[1], d[2]], cfg, T["P"], out=rays)))
    print("fused ms", timed(lambda: m.unproject_fuse_project(*d, cfg, T["P"], mode="sum", grid_out=grid, out=rays)))
    print("two calls ms", timed(lambda: (m.unproject_fuse(*d, cfg, mode="sum", out=grid), m.proj_grid([grid, d[1], d[2]], cfg, T["P"], out=rays))))
else:
    for _ in range(2):
        m.unproject_fuse(*d, cfg, mode="sum", out=grid)
    for _ in range(2):
        m.unproject_fuse_project(*d, cfg, T["P"], mode="sum", grid_out=grid, out=rays)
    torch.cuda.synchronize()
