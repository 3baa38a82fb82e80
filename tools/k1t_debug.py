"""Bring-up / A-B script for the tensor-core unprojection (K1T) on a GPU box: parity against the CUDA-core slot kernel and the
oracle, descriptor variants (DEBUG_ENV builds: MVF_K1T_DBG bit 0 / 1 swap the A / B descriptor offsets), and timing on T."""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child(dbg):
    os.environ["MVF_K1T_DBG"] = str(dbg)
    import numpy as np
    import torch
    import mulit_view_object_detection_b200 as m
    from mulit_view_object_detection_b200 import synthetic as syn
    out = {}
    for name, (nv, V, C, B) in {"small": (16, 3, 64, 1), "mid": (24, 4, 128, 2), "T1": (64, 8, 256, 1)}.items():
        cfg = m.FusionConfig(nvox=nv, nvox_z=nv, samples=8, NUM_VIEWS=V)
        feats, Rcam, Kmat = syn.make_scene(cfg, B, V, 40, 40, C, seed=7)
        d = [torch.from_numpy(a).cuda() for a in (feats, Rcam, Kmat)]
        ref = m.unproject_fuse(*d, cfg, mode="sum", tensor_cores=False)
        torch.cuda.synchronize()
        got = m.unproject_fuse(*d, cfg, mode="sum", tensor_cores=True)
        torch.cuda.synchronize()
        err = (got - ref).abs().max().item()
        rel = ((got - ref).abs() / (ref.abs() * 1e-5 + 1e-6)).max().item()
        out[name] = (err, rel, ref.abs().max().item(), float(got.double().sum()), float(ref.double().sum()))
        print("dbg=%d %-6s max|err|=%.3e  max err/(1e-5|ref|+1e-6)=%.3f  max|ref|=%.3f  sum got %.6e ref %.6e" % ((dbg, name) + out[name]), flush=True)
    return out


def timing():
    import torch
    import mulit_view_object_detection_b200 as m
    from mulit_view_object_detection_b200 import synthetic as syn
    cfg = m.FusionConfig(nvox=64, nvox_z=64, samples=20, NUM_VIEWS=8)
    B = 16
    feats, Rcam, Kmat = syn.make_scene(cfg, B, 8, 40, 40, 256, seed=1000)
    d = [torch.from_numpy(a).cuda() for a in (feats, Rcam, Kmat)]
    grid = torch.empty((B, 64, 64, 64, 256), device="cuda")
    for tc in (False, True):
        for _ in range(3):
            m.unproject_fuse(*d, cfg, mode="sum", out=grid, tensor_cores=tc)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        n = 30
        for _ in range(n):
            m.unproject_fuse(*d, cfg, mode="sum", out=grid, tensor_cores=tc)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        try:
            clk = torch.cuda.clock_rate()
        except Exception:
            clk = -1
        print("timing tensor_cores=%s: %.3f ms per %d scenes = %.1f us/scene -> %.3f of 6560 GB/s  (SM clock after the loop %d MHz)" %
              (tc, ms, B, ms * 1e3 / B, (1024.0 * (12800 + 262144) * B / (ms * 1e-3) / 1e9) / 6560.0, clk), flush=True)


def prof():
    """Per-role cycle counters of CTA 0 (DEBUG_ENV build)."""
    import ctypes
    import torch
    import mulit_view_object_detection_b200 as m
    from mulit_view_object_detection_b200 import synthetic as syn, _lib
    cfg = m.FusionConfig(nvox=64, nvox_z=64, samples=20, NUM_VIEWS=8)
    B = 16
    feats, Rcam, Kmat = syn.make_scene(cfg, B, 8, 40, 40, 256, seed=1000)
    d = [torch.from_numpy(a).cuda() for a in (feats, Rcam, Kmat)]
    grid = torch.empty((B, 64, 64, 64, 256), device="cuda")
    for _ in range(3):
        m.unproject_fuse(*d, cfg, mode="sum", out=grid, tensor_cores=True)
    torch.cuda.synchronize()
    buf = (ctypes.c_ulonglong * 32)()
    _lib.lib.mvf_debug_k1t_prof.restype = ctypes.c_int
    assert _lib.lib.mvf_debug_k1t_prof(buf) == 0
    v = list(buf)
    names = {0: "producer g0 [total, record wait, ring-slot wait, produce, -, K-steps, -, fence+arrive]", 8: "producer g1",
             16: "geometry [total, record-slot wait] / MMA warp [total, full wait, acc_empty wait, record wait]",
             24: "epilogue w0 [total, acc_full wait, -, work, wait_read, tmem wait::ld, scale+st.shared, fence+syncwarp+tma store]"}
    for base, nm in names.items():
        print("%-70s %s" % (nm, v[base:base + 8]))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        child(int(sys.argv[2]))
    elif len(sys.argv) > 1 and sys.argv[1] == "prof":
        prof()
    elif len(sys.argv) > 1 and sys.argv[1] == "timing":
        timing()
    else:
        for dbg in (0, 1, 2, 3):
            r = subprocess.run([sys.executable, __file__, "child", str(dbg)], timeout=60)
            print("dbg=%d rc=%d" % (dbg, r.returncode), flush=True)
        subprocess.run([sys.executable, __file__, "timing"], timeout=300)
