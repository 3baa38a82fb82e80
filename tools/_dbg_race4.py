import sys
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np, torch
import mulit_view_object_detection_b200 as m
from helpers import small_cfg, scene, to_dev
B, V, Cc, P = 16, 8, 256, 40
cfg = small_cfg(nvox=64, nvox_z=64, samples=20, NUM_VIEWS=V, IMAGE_SHAPE=np.array([640, 640, 3]))
feats, Rcam, Kmat = scene(cfg, B, V, 40, 40, Cc, seed=1000)
d = to_dev(feats, Rcam, Kmat)
slot = m.unproject_fuse(*d, cfg, mode="sum", tensor_cores=False)
def nbad(g, s):
    return int(((g - s).abs() > (1e-5 * s.abs() + 1e-6)).sum())
grid = torch.zeros((B, 64, 64, 64, Cc), device="cuda")
rays = torch.zeros((B, 20, P, P, Cc), device="cuda")
seq = sys.argv[1] if len(sys.argv) > 1 else "F1F1F1"
one = [t[0:1].contiguous() for t in d]
for i, c in enumerate(seq):
    if c == "F":
        grid.fill_(-3.0)
        m.unproject_fuse_project(*d, cfg, P, grid_out=grid, out=rays)
        torch.cuda.synchronize()
        print(i, "fused16: bad per scene", [nbad(grid[b], slot[b]) for b in (0, 1, 2, 15)])
    elif c == "S":
        g = m.unproject_fuse(*d, cfg, mode="sum")
        torch.cuda.synchronize()
        print(i, "standalone16: bad per scene", [nbad(g[b], slot[b]) for b in (0, 1, 2, 15)])
    elif c == "1":
        g1 = m.unproject_fuse(*one, cfg, mode="sum")
        torch.cuda.synchronize()
        print(i, "standalone B=1 scene 0: bad", nbad(g1[0], slot[0]))
    elif c == "2":
        two = [t[0:2].contiguous() for t in d]
        g2 = m.unproject_fuse(*two, cfg, mode="sum")
        torch.cuda.synchronize()
        print(i, "standalone B=2: bad", [nbad(g2[b], slot[b]) for b in (0, 1)])
print("---- controlled sequences")
def b1():
    m.unproject_fuse(*one, cfg, mode="sum"); torch.cuda.synchronize()
def check(tag, g):
    torch.cuda.synchronize()
    print(tag, [nbad(g[b], slot[b]) for b in (0, 1, 15)])
for rep in range(2):
    b1(); grid.fill_(-3.0); torch.cuda.synchronize(); m.unproject_fuse_project(*d, cfg, P, grid_out=grid, out=rays); check("a) 1, fill, sync, F ", grid)
    b1(); grid.fill_(-3.0); m.unproject_fuse_project(*d, cfg, P, grid_out=grid, out=rays); check("b) 1, fill, F       ", grid)
    b1(); grid.fill_(-3.0); m.unproject_fuse(*d, cfg, mode="sum", out=grid); check("c) 1, fill, S       ", grid)
    b1(); m.unproject_fuse_project(*d, cfg, P, grid_out=grid, out=rays); check("d) 1, F             ", grid)
    b1(); rays.fill_(0.0); m.unproject_fuse_project(*d, cfg, P, grid_out=grid, out=rays); check("e) 1, small fill, F ", grid)
    m.unproject_fuse_project(*d, cfg, P, grid_out=grid, out=rays); grid.fill_(-3.0); m.unproject_fuse_project(*d, cfg, P, grid_out=grid, out=rays); check("f) F, fill, F       ", grid)
