"""K1T bring-up diagnostics: structured feature fields isolate the A (weights / K range) and B (pixel / channel layout) operands."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import mulit_view_object_detection_b200 as m
from mulit_view_object_detection_b200 import synthetic as syn


def run(name, nv, V, C, B, field, seed=7, fh=40, fw=40):
    cfg = m.FusionConfig(nvox=nv, nvox_z=nv, samples=8, NUM_VIEWS=V)
    feats, Rcam, Kmat = syn.make_scene(cfg, B, V, fh, fw, C, seed=seed)
    if field == "const":
        feats[:] = 1.0
    elif field == "chan":
        feats[:] = (np.arange(C, dtype=np.float32) + 1)[None, None, None, None, :]
    elif field == "pix":
        feats[:] = (np.arange(fh * fw, dtype=np.float32).reshape(fh, fw) + 1)[None, None, :, :, None]
    elif field == "px":
        feats[:] = (np.arange(fw, dtype=np.float32) + 1)[None, None, None, :, None]
    elif field == "py":
        feats[:] = (np.arange(fh, dtype=np.float32) + 1)[None, None, :, None, None]
    d = [torch.from_numpy(a).cuda() for a in (feats, Rcam, Kmat)]
    ref = m.unproject_fuse(*d, cfg, mode="sum", tensor_cores=False)
    torch.cuda.synchronize()
    got = m.unproject_fuse(*d, cfg, mode="sum", tensor_cores=True)
    torch.cuda.synchronize()
    err = (got - ref).abs()
    bad = err > (1e-5 * ref.abs() + 1e-5)
    badvox = bad.any(dim=-1)
    nb = int(badvox.sum())
    print("%-28s max|err| %.3e  bad voxels %d / %d  bad elems %d" % (name, err.max().item(), nb, badvox.numel(), int(bad.sum())), flush=True)
    if nb:
        idx = badvox.nonzero()[:6].cpu().numpy()
        for i in idx:
            b, x, y, z = i
            g = got[b, x, y, z].cpu().numpy(); r = ref[b, x, y, z].cpu().numpy()
            print("    vox b%d x%d y%d z%d tile(%d,%d,%d) m=%d  got[:4]=%s ref[:4]=%s  nbadch=%d" %
                  (b, x, y, z, x // 4, y // 4, z // 8, ((x % 4) * 4 + y % 4) * 8 + z % 8, np.round(g[:4], 4), np.round(r[:4], 4), int(bad[b, x, y, z].sum())), flush=True)
        tiles = {}
        for i in badvox.nonzero().cpu().numpy():
            tiles[(i[0], i[1] // 4, i[2] // 4, i[3] // 8)] = tiles.get((i[0], i[1] // 4, i[2] // 4, i[3] // 8), 0) + 1
        print("    bad tiles: %d of %d; sample %s" % (len(tiles), badvox.numel() // 128, list(tiles.items())[:8]), flush=True)


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which in ("all", "a"):
        for field in ("const", "chan", "px", "py", "pix", "rand"):
            run("nv16 V1 C64 " + field, 16, 1, 64, 1, field)
        run("nv64 V1 C64 rand", 64, 1, 64, 1, "rand")
        run("nv64 V1 C256 const", 64, 1, 256, 1, "const")
        run("nv64 V1 C256 chan", 64, 1, 256, 1, "chan")
        run("nv64 V1 C256 rand", 64, 1, 256, 1, "rand")
        run("nv64 V8 C256 rand", 64, 8, 256, 1, "rand")
    if which in ("all", "b"):
        run("nv24 V4 C128 B2 rand", 24, 4, 128, 2, "rand")
