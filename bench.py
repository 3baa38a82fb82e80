#!/usr/bin/env python
"""bench.py -- fused unproject -> fuse -> project throughput (voxel-samples/s) on B200.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # CPU arm (oracle port on host cores)

Workload "T" (BASELINE.md section 4 / SURVEY.md section 8(d)): 8-view InteriorNet-shaped scenes, 256-ch P4
features 40x40 (640x640 padded input), 64^3 voxel grid, sum fusion, proj_grid P=40, S=20.
One *step* = one pass of the pipeline over a batch of ``--scenes`` synthetic scenes per GPU;
scenes are independent, so ranks shard scenes with no data-path collective (weak scaling).
One JSON line is printed by rank 0.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np

METRIC = "voxel_samples_per_s"
UNIT = "voxel-samples/s"
T = dict(V=8, C=256, nvox=64, fh=40, fw=40, P=40, S=20, image=640)
FALLBACK_HBM_GBS = 6650.0          # /opt/skills/guides/B200_PROFILING.md fallback


def make_config(B):
    from mulit_view_object_detection_b200.config import FusionConfig
    return FusionConfig(nvox=T["nvox"], nvox_z=T["nvox"], samples=T["S"], NUM_VIEWS=T["V"], GRID_REAS="add",
                        IMAGES_PER_GPU=B, IMAGE_SHAPE=np.array([T["image"], T["image"], 3]),
                        TOP_DOWN_PYRAMID_SIZE=T["C"])


def workload_config(args, n_gpus):
    return {"workload": "T: %d-view scene, 256-ch P4 40x40 features, %d^3 grid, unproject->sum-fuse->project(P=40,S=20), fp32"
                        % (T["V"], T["nvox"]),
            "scenes_per_gpu": args.scenes, "views": T["V"], "grid": [T["nvox"]] * 3, "channels": T["C"],
            "feature_hw": [T["fh"], T["fw"]], "proj": T["P"], "samples": T["S"],
            "sharding": "scenes sharded across ranks, no data-path collective" if n_gpus > 1 else "single GPU",
            "l2_policy": "inputs larger than L2: per-step working set = %d scenes x %.0f MB (features %.0f MB, grids %.0f MB)"
                         % (args.scenes, sum(algorithmic_bytes(1)) / 1e6, args.scenes * 13.1, args.scenes * T["nvox"] ** 3 * 1024 / 1e6)}


def algorithmic_bytes(B):
    """B_alg = 4*C*(V*fh*fw + N + 2*S*P*P) per scene; K1 = read every feature vector once
    + write the fused grid once; K3 = read one grid vector + write one output vector per ray sample."""
    N = T["nvox"] ** 3
    k1 = 4 * T["C"] * (T["V"] * T["fh"] * T["fw"] + N) * B
    k3 = 4 * T["C"] * (2 * T["S"] * T["P"] * T["P"]) * B
    return k1, k3


class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons through NVML while the timed region runs."""

    def __init__(self, index, period=0.01):
        super().__init__(daemon=True)
        self.period, self.samples, self.reasons, self.stop_flag = period, [], set(), False
        self.power_w = []
        self.sm_max = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.sm_max = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                self.power_w.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(self.period)

    def result(self):
        self.stop_flag = True
        if self.is_alive():
            self.join(timeout=1.0)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.sm_max, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.sm_max,
                "reasons": sorted(self.reasons), "samples": len(self.samples),
                "sm_mhz_min": float(np.min(self.samples)),
                "power_w_median": float(np.median(self.power_w)) if self.power_w else None,
                "power_w_max": float(np.max(self.power_w)) if self.power_w else None}


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def k1_traffic_bytes(B):
    """DRAM bytes per K1 launch, STATIC: taken from the committed `ncu --set full` capture of the same kernel on the same workload
    (profiles/k1_traffic.json: dram__bytes_read.sum + dram__bytes_write.sum per scene) and scaled to B scenes -- not measured in
    this run (ncu cannot run inside the timed bench)."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "k1_traffic.json")))
        return float(d["dram_bytes_per_scene"]) * B, "static: %s" % d.get("source", "profiles/k1_traffic.json")
    except Exception:
        return None, None


# ---------------------------------------------------------------------------------------------
def cpu_reference_run(scenes, slab, steps, warmup, threads=None):
    """The oracle's torch-CPU port (the stand-in for the reference's TF CPU path) on a bounded
    sample: ``scenes`` T-scenes restricted to an x-slab of ``slab`` of the 64 x-planes."""
    import torch
    from oracle import torch_cpu
    from mulit_view_object_detection_b200 import synthetic as syn
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    cfg = make_config(scenes)
    feats, Rcam, Kmat = syn.make_scene(cfg, scenes, T["V"], T["fh"], T["fw"], T["C"], seed=1234)
    feats_t = torch.from_numpy(feats)
    X = T["nvox"]
    full = torch.zeros((scenes, X, X, X, T["C"]), dtype=torch.float32)

    def step():
        part = torch_cpu.unproject_fuse(feats_t, Rcam, Kmat, cfg, "sum", x_slab=(0, slab))
        full[:, :slab] = part
        return torch_cpu.project(full, Rcam, Kmat, cfg, T["P"])

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    samples = scenes * T["V"] * slab * X * X * steps
    return samples / dt, dt / steps, threads


def run_reference(args):
    """CPU arm: the reference's algorithm on the host cores.  Imports only the pure-host modules of the package (config,
    synthetic) and oracle/ -- libmvfusion.so is never mapped and no GPU is touched in this process."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    slab = 16
    warmup = max(args.warmup, 3)
    value, s_per_step, threads = cpu_reference_run(1, slab, args.steps, warmup)
    sample = "1 scene of workload T restricted to an x-slab of %d/64 planes (%d voxel-samples per step) + full proj_grid" \
             % (slab, T["V"] * slab * 64 * 64)
    cfgd = workload_config(args, args.gpus)
    cfgd["workload"] += " -- CPU arm: bounded sample per step = " + sample
    cfgd["scenes_per_gpu"] = 1
    cfgd["sample_x_planes"] = slab
    cfgd["sharding"] = "rank 0 only, host cores"
    cfgd["l2_policy"] = "n/a (CPU arm)"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": warmup, "ms_per_step": s_per_step * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": cfgd,
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
            "native_library_loaded": any("libmvfusion" in l for l in open("/proc/self/maps")) if os.path.exists("/proc/self/maps") else None,
            "note": "TensorFlow/Keras are not installable here (no wheel, no network) and TF-CPU gather_nd raises on this "
                    "path's out-of-range taps; this arm times oracle/torch_cpu.py, the op-for-op torch-CPU port of the "
                    "reference graph (pinned bit for bit to the NumPy oracle and the reference-generated fixtures by "
                    "tests/test_oracle_torch_cpu.py), on all host threads"}
    print(json.dumps(line))
    return 0


def measured_bf16_peak():
    try:
        d = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), "measured (MEASURED_PEAKS.json bf16_tflops_sustained)"
    except Exception:
        return 1400.0, "fallback (B200_PROFILING.md)"


def convlstm_line(dev, local=0, steps=12, warmup=4):
    """K2: one ConvLSTM step of workload c3 (C = F = 256, 64^3 voxels, recurrent state present) on the tensor cores.
    Roofline: tensor.  `achieved` counts the f16 MMA work actually issued (3 MMAs per product: the fp32-parity split
    a*2^s = a1 + a2 in fp16 halves, fp32 accumulation); `useful` is the conv's own 2*M*K*N.  Peak = the measured
    sustained bf16/f16 rate.

    Reproducibility: the one-time costs (weight split / transpose, workspace allocation + first touch, first
    launch) are timed separately; then `warmup` untimed steps, then `steps` steps each bracketed by its own CUDA events, with
    SM clock / power / throttle reasons sampled DURING the loop.  The headline is the MEDIAN step; min and max are reported."""
    import torch
    import mulit_view_object_detection_b200 as m
    X, C = 64, 256
    g = torch.Generator(device=dev)
    g.manual_seed(0)
    W = torch.randn((3, 3, 3, 2 * C, 4 * C), device=dev, generator=g) * (2.0 / (27 * 2 * C + 4 * C)) ** 0.5
    b = torch.randn(4 * C, device=dev, generator=g) * 0.1
    x = torch.randn((1, X, X, X, C), device=dev, generator=g).relu_()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    cell = m.ConvLSTMTensorCore(W, b, 1.0)
    torch.cuda.synchronize()
    t_prepare = time.perf_counter() - t0
    t0 = time.perf_counter()
    h, c = cell.step(x, None, None)                 # allocates + first-touches the operand workspace, first launch
    torch.cuda.synchronize()
    t_first = time.perf_counter() - t0
    for _ in range(warmup):
        cell.step(x, h, c)
    torch.cuda.synchronize()
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(steps)]
    sampler = ClockSampler(local, period=0.005)
    sampler.start()
    for k in range(steps):
        ev[k][0].record()
        h2, c2 = cell.step(x, h, c)
        ev[k][1].record()
    torch.cuda.synchronize()
    clocks = sampler.result()
    per = sorted(e[0].elapsed_time(e[1]) for e in ev)
    ms = float(np.median(per))
    flop = 2.0 * X ** 3 * 27 * 2 * C * 4 * C
    peak, src = measured_bf16_peak()
    return {"workload": "c3 step: ConvLSTM 3x3x3, C=F=256, 64^3 voxels, K=13824, N=1024, fp32-parity 3xFP16 split on tcgen05 "
                        "(operand split passes included)",
            "ms_per_step": ms, "ms_min": per[0], "ms_max": per[-1], "steps": steps, "warmup": warmup,
            "one_time_ms": {"prepare_weights": t_prepare * 1e3, "first_step_incl_workspace_alloc": t_first * 1e3},
            "clocks": clocks,
            "useful_tflops": flop / ms / 1e9,
            "roofline": {"bound": "tensor", "achieved": 3.0 * flop / ms / 1e9, "peak": peak, "unit": "TFLOP/s",
                         "frac": 3.0 * flop / ms / 1e9 / peak, "traffic": None,
                         "peak_source": src, "kernel": "convlstm_tc_kernel<false,true> (K2)"},
            "checksum": float(h2.double().sum())}


def c2_line(dev, peak):
    """Config c2 extras (4-view scene, 48^3 grid, max-fuse + PyramidROIAlign + NMS on one B200): device-resident timings of the
    head kernels and of the max-fuse pipeline at the P4 level, CUDA events over 20 back-to-back calls each."""
    import torch
    import mulit_view_object_detection_b200 as m
    from mulit_view_object_detection_b200 import synthetic as syn
    rng = np.random.default_rng(2000)
    C, img = 256, 640
    cfg = m.FusionConfig(nvox=48, nvox_z=48, samples=20, NUM_VIEWS=4, GRID_REAS="max", IMAGE_SHAPE=np.array([img, img, 3]))

    def timed(fn, n=20):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    maps = [torch.from_numpy(np.maximum(rng.standard_normal((1, img // s, img // s, C), dtype=np.float32), 0)).to(dev) for s in (4, 8, 16, 32)]
    meta = syn.make_image_meta(1, (img, img, 3), 25)
    boxes = torch.from_numpy(syn.make_rois(rng, 1, 1000)).to(dev)
    roi = m.PyramidROIAlign([7, 7])
    b6 = torch.from_numpy(syn.make_rois(rng, 1, 6000, pad_frac=0)[0]).to(dev)
    s6 = torch.from_numpy(rng.permutation(6000).astype(np.float32) / 6000).to(dev)
    probs, deltas = syn.make_detection_inputs(rng, 1000, 25)
    det = [torch.from_numpy(a).to(dev) for a in (syn.make_rois(rng, 1, 1000)[0], probs, deltas)]
    window = torch.tensor([0.0, 0.0, 1.0, 1.0], device=dev)
    feats, Rcam, Kmat = syn.make_scene(cfg, 1, 4, 40, 40, C, seed=2001)
    d = [torch.from_numpy(a).to(dev) for a in (feats, Rcam, Kmat)]
    grid = torch.empty((1, 48, 48, 48, C), device=dev)
    rays = torch.empty((1, 20, 40, 40, C), device=dev)
    ms_roi = timed(lambda: roi([boxes, meta] + maps))
    ms_fuse = timed(lambda: m.unproject_fuse_project(*d, cfg, 40, mode="max", grid_out=grid, out=rays))
    lower = 4 * C * 1000 * 49 * 2
    return {"workload": "c2: 4 views, 48^3, max-fuse + projection at P4 (one scene); PyramidROIAlign 1000x7x7x256; NMS of 6000 boxes; "
                        "refine_detections 1000x25 (timings include the Python/ctypes launch path)",
            "fusion_max_P4_ms": ms_fuse, "fusion_voxel_samples_per_s": 4 * 48 ** 3 / ms_fuse * 1e3,
            "roi_align_1000x7x7_ms": ms_roi, "roi_align_frac_of_hbm_peak_vs_2_vectors_per_bin": lower / ms_roi / 1e6 / peak,
            "nms_6000_boxes_ms": timed(lambda: m.non_max_suppression(b6, s6, 1000, 0.7)),
            "refine_detections_1000x25_ms": timed(lambda: m.refine_detections_graph(det[0], det[1], det[2], window, cfg))}


def c4_line(dev, scenes=8, views=5, steps=3, grid_reas="add", channels=256):
    """Config c4: full model_multi Mask R-CNN inference (model.py: MaskRCNN.predict), ResNet-101 + FPN backbone, 5 views per scene,
    batch 8 scenes on one GPU, 640x640 inputs, 256-channel pyramid, GRID_REAS='add', the reference InferenceConfig's 40^3 grid
    (samples/interior/interior_multi.py:397-421), random-init weights.  The fusion neck, ProposalLayer, PyramidROIAlign and
    DetectionLayer run on this repo's kernels; backbone / RPN / head convolutions are cuDNN / cuBLAS fp32 (TF32 off), timed a
    second time with TF32 allowed.  Inputs resident in HBM; the per-stage times come from CUDA events around each stage."""
    import torch
    import mulit_view_object_detection_b200 as m
    from mulit_view_object_detection_b200 import synthetic as syn
    cfg = m.FusionConfig(IMAGE_SHAPE=np.array([640, 640, 3]), NUM_VIEWS=views, IMAGES_PER_GPU=scenes, TOP_DOWN_PYRAMID_SIZE=channels,
                         NUM_CLASSES=23, nvox=40, nvox_z=40, samples=20, GRID_REAS=grid_reas, BACKBONE="resnet101",
                         DETECTION_MIN_CONFIDENCE=0.0)
    out = {"workload": "c4: MaskRCNN.predict, ResNet-101 + FPN, %d scenes x %d views, 640x640, %d-ch pyramid, 40^3 grid, GRID_REAS=%s, "
                       "1000 proposals, 100 detections + masks, random-init weights, inputs in HBM" % (scenes, views, channels, grid_reas)}
    with torch.no_grad():
        net = m.MaskRCNN("inference", cfg, device=dev, seed=4)
        rng = np.random.default_rng(4000)
        images = torch.from_numpy(rng.normal(0, 50, (scenes, views, 640, 640, 3)).astype(np.float32)).to(dev)
        _, Rcam, Kmat = syn.make_scene(cfg, scenes, views, 8, 8, 4, seed=4001)
        meta = np.stack([m.weights_io.compose_image_meta(0, (640, 640, 3), (640, 640, 3), (0, 0, 640, 640), 1.0,
                                                         np.zeros(cfg.NUM_CLASSES, np.int32)) for _ in range(scenes)]).astype(np.float32)
        a = net.get_anchors((640, 640, 3))
        anchors = torch.from_numpy(np.broadcast_to(a, (scenes,) + a.shape).copy()).to(dev)
        inputs = [images, meta, anchors, torch.from_numpy(Rcam).to(dev), torch.from_numpy(Kmat).to(dev)]
        for tf32 in (False, True):
            net.allow_tf32 = tf32
            for _ in range(2):
                res = net.predict(inputs)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n0 = m.launch_count()
            e0.record()
            for _ in range(steps):
                res = net.predict(inputs)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
            key = "tf32_dense" if tf32 else "fp32_dense"
            out[key] = {"ms_per_batch": ms, "scenes_per_s": scenes / ms * 1e3, "views_per_s": scenes * views / ms * 1e3}
            if not tf32:
                out["library_kernel_launches_per_batch"] = (m.launch_count() - n0) // steps
                out["detections_checksum"] = float(res[0].double().sum())
                out["n_detections"] = int((res[0][..., 4] > 0).sum())
        # stage split (fp32): backbone+FPN | fusion neck | RPN + proposals | heads
        net.allow_tf32 = False
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        ev[0].record()
        P = net.build_resnet_fpn(images)
        ev[1].record()
        maps = m.fusion_neck(P, inputs[3], inputs[4], cfg, params=net.neck_params)
        ev[2].record()
        torch.cuda.synchronize()
        out["stage_ms_fp32"] = {"backbone_fpn": ev[0].elapsed_time(ev[1]), "fusion_neck_levels_4_5_6": ev[1].elapsed_time(ev[2]),
                                "rpn_proposals_heads": out["fp32_dense"]["ms_per_batch"] - ev[0].elapsed_time(ev[2])}
        del net, images, P, maps, res
    torch.cuda.empty_cache()
    return out


def cooperative_lines(dev, world, rank):
    """The cooperative (strong-scaling) splits BASELINE.json names, measured at EVERY --gpus N so that the driver's SCALE record
    carries them: all ranks work on the SAME scenes and the timed step includes the NCCL exchange.
      c3_lstm_slab            config c3: one 8-view 64^3 scene, recurrent fusion C = F = 256, x-slabs + 1-voxel halo of h per step
      c5_slab_owner           config c5: 32 scenes x 8 views, 96^3: every rank fuses all views for its x-slab, collapses the depth
                              axis of its own ray samples linearly, all-reduce of PG [B,P,P,C], bias/BN/ReLU after the sum
      c5_view_reduce_scatter  config c5 as BASELINE.json words it: views sharded, per-scene reduce-scatter of the partial grids by
                              x-slab (pipelined against the next scene's unprojection), slab-local projection, all-reduce of rays
    Each entry: ms_per_step (CUDA events, max over ranks), checksum, and the max error of scene 0 against the plain single-GPU
    path computed in the same run by rank 0."""
    import torch
    import torch.distributed as dist
    import mulit_view_object_detection_b200 as m
    from mulit_view_object_detection_b200 import synthetic as syn, dist as mvd
    out = {}
    solo = None
    if world > 1:
        groups = [dist.new_group(ranks=[r]) for r in range(world)]        # collective: every rank creates every group
        solo = groups[rank]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            r = fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            r = fn()
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), r

    def rel_err(a, b):
        return float(((a - b).abs() / (1e-5 * b.abs() + 1e-6)).max().item())

    # ---- c3: recurrent fusion of one scene
    V, C, X = 8, 256, 64
    cfg = make_config(1)
    feats, Rcam, Kmat = syn.make_scene(cfg, 1, V, T["fh"], T["fw"], C, seed=3000)
    d = [torch.from_numpy(a).to(dev) for a in (feats, Rcam, Kmat)]
    g = torch.Generator(device=dev)
    g.manual_seed(0)
    params = {"W": torch.randn((3, 3, 3, 2 * C, 4 * C), device=dev, generator=g) * (2.0 / (27 * 2 * C + 4 * C)) ** 0.5,
              "b": torch.randn(4 * C, device=dev, generator=g) * 0.1}
    ms, (rays, _) = timed(lambda: mvd.lstm_slab(*d, cfg, params, proj_size=T["P"]), steps=3, warmup=2)
    err = None
    if rank == 0 and world > 1:
        ref, _ = mvd.lstm_slab(*d, cfg, params, proj_size=T["P"], group=solo)
        err = float((rays - ref).abs().max().item())                       # the sharded recurrence is bit-identical by design
    barrier()
    flop = 2.0 * X ** 3 * 27 * 2 * C * 4 * C * V
    out["c3_lstm_slab"] = {"workload": "c3: one 8-view 64^3 scene, ConvLSTM 3x3x3 C=F=256 (3xFP16 split on tcgen05), x-slabs + halo of h per step, proj_grid",
                           "ms_per_step": ms, "useful_tflops": flop / ms / 1e9, "checksum": float(rays.double().sum()),
                           "max_abs_err_vs_single_gpu": err, "exchange": "2 planes of h (4 MB each) per neighbour per view step (isend/irecv) + 4-byte all-reduce of the operand scale + all-reduce of the ray slices (32.8 MB)"}
    del rays, params
    torch.cuda.empty_cache()

    # ---- c5: 32 scenes x 8 views, 96^3
    Bc, Xc = 32, 96
    from mulit_view_object_detection_b200.config import FusionConfig
    cfg5 = FusionConfig(nvox=Xc, nvox_z=Xc, samples=T["S"], NUM_VIEWS=V, GRID_REAS="add", IMAGES_PER_GPU=Bc,
                        IMAGE_SHAPE=np.array([T["image"], T["image"], 3]), TOP_DOWN_PYRAMID_SIZE=C)
    feats, Rcam, Kmat = syn.make_scene(cfg5, Bc, V, T["fh"], T["fw"], C, seed=5000)
    d = [torch.from_numpy(a).to(dev) for a in (feats, Rcam, Kmat)]
    depth = {"weight": torch.full((T["S"],), 1.0 / T["S"], device=dev), "bias": 0.01, "bn": (1.1, 0.02, -0.01, 0.9)}
    ms, (pg, _) = timed(lambda: mvd.slab_owner(*d, cfg5, T["P"], mode="sum", depth=depth), steps=3, warmup=2)
    err = None
    if rank == 0:
        one = [t[:1].contiguous() for t in d]
        fused = m.unproject_fuse(*one, cfg5, mode="sum")
        ref = m.proj_grid_depth_sampling([fused, one[1], one[2]], cfg5, T["P"], "depth", params=depth)
        err = rel_err(pg[:1], ref)
        del fused, ref
    barrier()
    out["c5_slab_owner"] = {"workload": "c5: 32 scenes x 8 views, 96^3, slab owner + linear depth collapse per slab, PG [B,P,P,C] out",
                            "ms_per_step": ms, "voxel_samples_per_s": Bc * V * Xc ** 3 / ms * 1e3, "checksum": float(pg.double().sum()),
                            "max_err_vs_single_gpu_in_tolerance_units": err,
                            "exchange": "one all-reduce of PG: %d scenes x 1.6 MB = %.0f MB" % (Bc, Bc * T["P"] * T["P"] * C * 4 / 1e6)}
    del pg
    torch.cuda.empty_cache()
    if world <= V and Xc % world == 0:
        ms, (rays, _) = timed(lambda: mvd.view_shard_reduce_scatter(*d, cfg5, T["P"], mode="sum"), steps=2, warmup=1)
        err = None
        if rank == 0:
            one = [t[:1].contiguous() for t in d]
            ref, _ = m.unproject_fuse_project(*one, cfg5, T["P"], mode="sum")
            err = rel_err(rays[:1], ref)
            del ref
        barrier()
        out["c5_view_reduce_scatter"] = {"workload": "c5: 32 scenes x 8 views, 96^3, views sharded, per-scene reduce-scatter of the partial grids "
                                                     "by x-slab (pipelined), slab-local projection, all-reduce of ray slices",
                                         "ms_per_step": ms, "voxel_samples_per_s": Bc * V * Xc ** 3 / ms * 1e3, "checksum": float(rays.double().sum()),
                                         "max_err_vs_single_gpu_in_tolerance_units": err,
                                         "exchange": "reduce-scatter of %.1f GB of partial grids per rank + all-reduce of %.2f GB of ray slices"
                                                     % (Bc * Xc ** 3 * C * 4 / 1e9, Bc * T["S"] * T["P"] * T["P"] * C * 4 / 1e9)}
        del rays
    torch.cuda.empty_cache()
    return out


# ---------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist
    import mulit_view_object_detection_b200 as m
    from mulit_view_object_detection_b200 import synthetic as syn

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")
    # bind the rank to the CPUs (hence, by first touch, the pinned pages) of its GPU's NUMA node before anything is allocated
    numa_node = None
    if not args.no_numa_bind:
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        try:
            from pcie_probe import bind_to_gpu_numa
            numa_node = bind_to_gpu_numa(local)
        except Exception:
            numa_node = None
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B = args.scenes
    cfg = make_config(B)
    stream = torch.cuda.current_stream()

    # synthetic scenes: one seeded scene block per rank, resident in HBM before the timed region
    feats, Rcam, Kmat = syn.make_scene(cfg, B, T["V"], T["fh"], T["fw"], T["C"], seed=1000 + rank)
    d_feats, d_R, d_K = (torch.from_numpy(a).to(dev) for a in (feats, Rcam, Kmat))
    X = T["nvox"]
    grid = torch.empty((B, X, X, X, T["C"]), dtype=torch.float32, device=dev)
    rays = torch.empty((B, T["S"], T["P"], T["P"], T["C"]), dtype=torch.float32, device=dev)

    def step():
        # one C call (mvf_unproject_fuse_project): feature split, tensor-core unprojection and projection queued with programmatic
        # stream serialization, overlapping scene by scene
        m.unproject_fuse_project(d_feats, d_R, d_K, cfg, T["P"], mode="sum", grid_out=grid, out=rays)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()

    # ---- timed region: device-resident inputs; CUDA events on the launching stream around every step
    K = args.steps
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(K)]
    sampler = ClockSampler(local)
    sampler.start()
    n0 = m.launch_count()
    barrier()
    for k in range(K):
        ev[k][0].record(stream)
        step()
        ev[k][1].record(stream)
    barrier()
    launches = m.launch_count() - n0
    clocks = sampler.result()
    # the two kernels of the step on their own (same inputs, same launch path, back to back): K1T with its split (the split runs
    # under it from the second scene on) for the roofline object, K3 for the pipeline object
    for _ in range(2):          # (first use of a kernel -- here torch's gather of the main-view poses inside proj_grid -- loads it: tens of ms)
        m.unproject_fuse(d_feats, d_R, d_K, cfg, mode="sum", out=grid)
        m.proj_grid([grid, d_R, d_K], cfg, T["P"], out=rays)
    torch.cuda.synchronize()
    ek = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    ek[0].record(stream)
    for _ in range(K):
        m.unproject_fuse(d_feats, d_R, d_K, cfg, mode="sum", out=grid)
    ek[1].record(stream)
    for _ in range(K):
        m.proj_grid([grid, d_R, d_K], cfg, T["P"], out=rays)
    ek[2].record(stream)
    torch.cuda.synchronize()
    # the CUDA-core slot kernel on the same inputs, for the record (not part of the timed step)
    es0, es1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    m.unproject_fuse(d_feats, d_R, d_K, cfg, mode="sum", out=grid, tensor_cores=False)
    es0.record(stream)
    for _ in range(3):
        m.unproject_fuse(d_feats, d_R, d_K, cfg, mode="sum", out=grid, tensor_cores=False)
    es1.record(stream)
    torch.cuda.synchronize()
    k1_slot_ms = es0.elapsed_time(es1) / 3
    m.unproject_fuse(d_feats, d_R, d_K, cfg, mode="sum", out=grid)            # leave the K1T result in `grid`
    total_ms = ev[0][0].elapsed_time(ev[K - 1][1])
    k1_ms = ek[0].elapsed_time(ek[1]) / K
    k3_ms = ek[1].elapsed_time(ek[2]) / K
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    voxel_samples_step = world * B * T["V"] * X ** 3
    value = voxel_samples_step * K / (total_ms * 1e-3)

    # ---- end to end through the C-ABI host entry: pinned host buffers in, pinned ray slices out
    pipe = m.HostPipeline(cfg, B, T["V"], T["fh"], T["fw"], T["C"], T["P"], mode="sum")
    h_in = [torch.from_numpy(a).pin_memory() for a in (feats, Rcam, Kmat)]
    h_out = pipe.empty_output()
    for _ in range(3):
        pipe(h_in[0], h_in[1], h_in[2], h_out)
    barrier()
    Ke = max(3, min(K, 10))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(Ke):
        pipe(h_in[0], h_in[1], h_in[2], h_out)
    e1.record(stream)
    barrier()
    te = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = voxel_samples_step * Ke / (float(te.item()) * 1e-3)
    checksum = float(h_out.double().sum())          # the device->host result is really read

    # ---- the same crossing at the boundary the reference model has (model_multi.py:2382-2404): features in, PG out --
    # grid_reas BatchNorm + ReLU fused into K1, depth_sampling fused into the projection, so only [B,P,P,C] comes back
    depth = {"weight": np.full(T["S"], 1.0 / T["S"], np.float32), "bias": 0.0, "bn": (1.0, 0.0, 0.0, 1.0)}
    ones, zeros = np.ones(T["C"], np.float32), np.zeros(T["C"], np.float32)
    neck = m.HostPipeline(cfg, B, T["V"], T["fh"], T["fw"], T["C"], T["P"], mode="sum", depth=depth,
                          bn=(ones, zeros, zeros, ones), relu_out=True)
    n_out = neck.empty_output()
    for _ in range(3):
        neck(h_in[0], h_in[1], h_in[2], n_out)
    barrier()
    e0.record(stream)
    for _ in range(Ke):
        neck(h_in[0], h_in[1], h_in[2], n_out)
    e1.record(stream)
    barrier()
    tn = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tn, op=dist.ReduceOp.MAX)
    neck_value = voxel_samples_step * Ke / (float(tn.item()) * 1e-3)
    neck_checksum = float(n_out.double().sum())

    coop = None
    pipe_bytes, neck_bytes = (pipe.h2d_bytes, pipe.d2h_bytes), (neck.h2d_bytes, neck.d2h_bytes)
    if not args.no_cooperative:
        del grid, rays, pipe, neck, h_out, n_out
        torch.cuda.empty_cache()
        coop = cooperative_lines(dev, world, rank)
    c4_dp = None
    if world > 1 and not args.no_model:
        # config c4 as BASELINE.json words it: full model inference, batch of 8 scenes x 5 views PER GPU, data parallel (no
        # collective: scenes are independent); aggregate scenes/s over the slowest rank
        try:
            mine = c4_line(dev)
            t = torch.tensor([mine["fp32_dense"]["ms_per_batch"], mine["tf32_dense"]["ms_per_batch"]], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            c4_dp = {"workload": mine["workload"] + " -- per GPU, %d GPUs data parallel" % world,
                     "fp32_dense": {"ms_per_batch_max_over_ranks": float(t[0]), "scenes_per_s": world * 8 / float(t[0]) * 1e3},
                     "tf32_dense": {"ms_per_batch_max_over_ranks": float(t[1]), "scenes_per_s": world * 8 / float(t[1]) * 1e3},
                     "rank0_stage_ms_fp32": mine["stage_ms_fp32"]}
        except Exception as e:
            c4_dp = {"error": "%s: %s" % (type(e).__name__, e)}
    if rank == 0:
        peak, peak_src = measured_peak()
        k1_bytes, k3_bytes = algorithmic_bytes(B)
        achieved = k1_bytes / (k1_ms * 1e-3) / 1e9
        traffic, traffic_src = k1_traffic_bytes(B)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": max(args.warmup, 3),
            "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(args, world),
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": pipe_bytes[0], "d2h_bytes_per_step": pipe_bytes[1],
                    "steps": Ke, "api": "mvf_unproject_fuse_project_host (pinned host buffers, H2D + K1 + K3 + D2H + sync)",
                    "checksum": checksum,
                    "bound": "host link: %.0f MB cross PCIe per rank per step (%.1f GB/s per rank achieved here, both directions "
                             "together); see profiles/r2_pcie_probe.json for the measured copy ceiling at 1/2/4/8 ranks"
                             % ((pipe_bytes[0] + pipe_bytes[1]) / 1e6,
                                (pipe_bytes[0] + pipe_bytes[1]) * world / 1e9 / (voxel_samples_step / e2e_value) / world),
                    "numa_node": numa_node},
            "e2e_neck": {"value": neck_value, "unit": UNIT, "h2d_bytes_per_step": neck_bytes[0], "d2h_bytes_per_step": neck_bytes[1],
                         "steps": Ke, "api": "mvf_fusion_neck_level_host (features in, depth-sampled PG [B,P,P,C] out: H2D + K1(+BN+ReLU) "
                                             "+ K3b + D2H + sync) -- the host/device boundary of the reference model; extra to `e2e`",
                         "checksum": neck_checksum},
            "gpu_launches": int(launches),
            "roofline": {"kernel": "unproject_tc_kernel (K1T: tcgen05 unprojection) timed on its own, back-to-back calls of mvf_unproject_fuse_tc; "
                                   "the time includes its feature-split kernel, which runs under it from the second scene on",
                         "bound": "hbm", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                         "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": k1_bytes, "ms_per_launch": k1_ms,
                         "share_of_step": k1_ms / (total_ms / K),
                         "share_note": "K1T alone / overlapped step: in the step the projection of scene b-1 runs under the unprojection of scene b",
                         "cuda_core_slot_kernel_ms": k1_slot_ms,
                         "cuda_core_slot_kernel_frac": k1_bytes / (k1_slot_ms * 1e-3) / 1e9 / peak},
            "pipeline": {"algorithmic_bytes_per_step": k1_bytes + k3_bytes, "achieved_gbs": (k1_bytes + k3_bytes) / (total_ms / K * 1e-3) / 1e9,
                         "frac_of_hbm_peak": (k1_bytes + k3_bytes) / (total_ms / K * 1e-3) / 1e9 / peak,
                         "step": "mvf_unproject_fuse_project: split | K1T | K3 overlapped scene by scene (programmatic dependent launch + device counters)",
                         "k1_alone_ms": k1_ms, "k3_alone_ms": k3_ms, "sum_of_kernels_alone_ms": k1_ms + k3_ms,
                         "k3_achieved_gbs": k3_bytes / (k3_ms * 1e-3) / 1e9},
        }
        if coop is not None:
            line["cooperative"] = coop
        if world == 1 and not args.no_convlstm:
            line["k2_convlstm"] = convlstm_line(dev, local)
            line["c2_heads"] = c2_line(dev, peak)
            if not args.no_model:
                try:
                    line["c4_model"] = c4_line(dev)
                except Exception as e:                       # the headline line must not depend on the extra
                    line["c4_model"] = {"error": "%s: %s" % (type(e).__name__, e)}
                try:                                         # the neck the reference's shipped configs use (interior_multi.py:391-419)
                    line["c4_model_conv3d"] = c4_line(dev, grid_reas="conv3d", channels=64)
                except Exception as e:
                    line["c4_model_conv3d"] = {"error": "%s: %s" % (type(e).__name__, e)}
        if c4_dp is not None:
            line["c4_model"] = c4_dp
        if world == 1 and not args.no_cpu_baseline:
            n_cpu = 24
            v, s_per, threads = cpu_reference_run(1, 16, n_cpu, 1)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": "%d passes over 1 scene of workload T restricted to an x-slab of 16/64 planes (524288 "
                                              "voxel-samples per pass) + full proj_grid, oracle/torch_cpu.py on all host threads, "
                                              "%.1f s of CPU work" % (n_cpu, s_per * n_cpu)}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def run_cooperative(args):
    """Strong-scaling measurement of the cooperative strategies: all ranks work on the
    SAME --scenes scenes; the timed step includes the NCCL collective.  Extra to the headline line."""
    import torch
    import torch.distributed as dist
    import mulit_view_object_detection_b200 as m
    from mulit_view_object_detection_b200 import synthetic as syn, dist as mvd

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B = args.scenes
    cfg = make_config(B)
    feats, Rcam, Kmat = syn.make_scene(cfg, B, T["V"], T["fh"], T["fw"], T["C"], seed=1000)   # same scenes on every rank
    d = [torch.from_numpy(a).to(dev) for a in (feats, Rcam, Kmat)]
    if args.strategy == "lstm_slab":
        # config c3: recurrent fusion (ConvLSTM over the 8 views, C = F = 256) with the grid split into x-slabs,
        # one halo exchange of h per step; the tensor cores do 59 TFLOP of useful work per scene
        g = torch.Generator(device=dev)
        g.manual_seed(0)
        Cc = T["C"]
        params = {"W": torch.randn((3, 3, 3, 2 * Cc, 4 * Cc), device=dev, generator=g) * (2.0 / (27 * 2 * Cc + 4 * Cc)) ** 0.5,
                  "b": torch.randn(4 * Cc, device=dev, generator=g) * 0.1}
        fn = lambda f, R, K, cfg_, P, mode: mvd.lstm_slab(f, R, K, cfg_, params, proj_size=P)
    else:
        fn = {"view_allreduce": mvd.view_shard_allreduce, "view_reduce_scatter": mvd.view_shard_reduce_scatter,
              "slab_owner": mvd.slab_owner,
              "slab_owner_scatter": lambda *a, **k: mvd.slab_owner(*a, scatter_scenes=True, **k)}[args.strategy]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3) if args.strategy != "lstm_slab" else 2):     # 2: the first calls create the NCCL channels
        fn(*d, cfg, T["P"], mode="sum")
    barrier()
    stream = torch.cuda.current_stream()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n0 = m.launch_count()
    e0.record(stream)
    for _ in range(args.steps):
        rays, _ = fn(*d, cfg, T["P"], mode="sum")
    e1.record(stream)
    barrier()
    launches = m.launch_count() - n0
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    if rank == 0:
        value = B * T["V"] * T["nvox"] ** 3 * args.steps / (total_ms * 1e-3)
        cfgd = workload_config(args, world)
        cfgd["sharding"] = args.strategy
        if args.strategy == "lstm_slab":
            cfgd["workload"] = "c3: %d-view scene, %d^3 grid, recurrent voxel fusion (ConvLSTM 3x3x3, C=F=256, 3xFP16 split on tcgen05), " \
                               "x-slabs + 1-voxel halo of h exchanged per step, then proj_grid" % (T["V"], T["nvox"])
            flop = 2.0 * T["nvox"] ** 3 * 27 * 2 * T["C"] * 4 * T["C"] * T["V"] * B
            extra = {"useful_tflops": flop * args.steps / (total_ms * 1e-3) / 1e12}
        else:
            extra = {}
        print(json.dumps({"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                          "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps, "higher_is_better": True,
                          "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfgd,
                          "gpu_launches": int(launches), "checksum": float(rays.double().sum()), **extra}))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--scenes", type=int, default=16, help="scenes per GPU per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-convlstm", action="store_true", help="skip the extra K2 (ConvLSTM on tensor cores) measurement")
    ap.add_argument("--no-model", action="store_true", help="skip the extra c4 (full model_multi inference) measurement")
    ap.add_argument("--no-numa-bind", action="store_true", help="do not bind the rank to its GPU's NUMA node")
    ap.add_argument("--no-cooperative", action="store_true", help="skip the cooperative multi-GPU splits (c3 / c5 strong scaling)")
    ap.add_argument("--strategy", default="scene", choices=["scene", "view_allreduce", "view_reduce_scatter", "slab_owner", "slab_owner_scatter", "lstm_slab"],
                    help="multi-GPU sharding: scene (default, weak scaling, no collective) or one of the cooperative "
                         "strategies of dist.py on a FIXED batch of --scenes scenes (strong scaling)")
    ap.add_argument("--nvox", type=int, default=64, help="voxels per grid axis (64 = workload T; 96 = config c5)")
    args = ap.parse_args()
    T["nvox"] = args.nvox
    if args.impl == "reference":
        return run_reference(args)
    if args.strategy != "scene":
        return run_cooperative(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
