"""Host <-> device copy ceiling per rank and in aggregate (why e2e scales 1.66x at 8 GPUs).

    python tools/pcie_probe.py                                   # 1 GPU
    python -m torch.distributed.run --nproc-per-node N ... tools/pcie_probe.py [--numa]

Each rank copies the e2e step's volumes (210 MB H2D, 524 MB D2H) between PINNED host buffers and its GPU: H2D alone, D2H alone,
both at once on two streams -- all ranks simultaneously (barrier before each phase), CUDA events, max over ranks.  --numa binds
the process (CPU affinity, hence first-touch placement of the pinned pages) to the NUMA node of its GPU before allocating."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def gpu_numa_node(index):
    """NUMA node and CPU list of GPU `index` from sysfs (None when the platform does not expose it)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        bus = bus.lower()
        if len(bus.split(":")[0]) == 8:                 # nvml pads the domain to 8 hex digits, sysfs uses 4
            bus = bus[4:]
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bus).read())
        if node < 0:
            return None, None
        cpus = open("/sys/devices/system/node/node%d/cpulist" % node).read().strip()
        out = set()
        for part in cpus.split(","):
            a, _, b = part.partition("-")
            out.update(range(int(a), int(b or a) + 1))
        return node, sorted(out)
    except Exception:
        return None, None


def bind_to_gpu_numa(index):
    node, cpus = gpu_numa_node(index)
    if node is None or not cpus:
        return None
    try:
        os.sched_setaffinity(0, cpus)
        return node
    except Exception:
        return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--numa", action="store_true")
    ap.add_argument("--h2d-mb", type=float, default=209.7)
    ap.add_argument("--d2h-mb", type=float, default=524.3)
    ap.add_argument("--reps", type=int, default=10)
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    node = bind_to_gpu_numa(local) if args.numa else None
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n_in, n_out = int(args.h2d_mb * 1e6) // 4, int(args.d2h_mb * 1e6) // 4
    h_in = torch.empty(n_in, dtype=torch.float32).pin_memory()
    h_out = torch.empty(n_out, dtype=torch.float32).pin_memory()
    h_in.fill_(1.0); h_out.fill_(0.0)                   # touch the pages
    d_in = torch.empty(n_in, dtype=torch.float32, device=dev)
    d_out = torch.ones(n_out, dtype=torch.float32, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def phase(do_in, do_out):
        for _ in range(2):
            if do_in:
                d_in.copy_(h_in, non_blocking=True)
            if do_out:
                h_out.copy_(d_out, non_blocking=True)
        barrier()
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record()
        s1.wait_event(e0); s2.wait_event(e0)
        for _ in range(args.reps):
            if do_in:
                with torch.cuda.stream(s1):
                    d_in.copy_(h_in, non_blocking=True)
            if do_out:
                with torch.cuda.stream(s2):
                    h_out.copy_(d_out, non_blocking=True)
        e1.record(s1); e2.record(s2)
        torch.cuda.current_stream().wait_event(e1); torch.cuda.current_stream().wait_event(e2)
        e3 = torch.cuda.Event(enable_timing=True)
        e3.record()
        barrier()
        ms = e0.elapsed_time(e3) / args.reps
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    res = {"world": world, "numa_bound": args.numa, "numa_node_rank0": node, "cpus": len(os.sched_getaffinity(0)),
           "h2d_mb": args.h2d_mb, "d2h_mb": args.d2h_mb}
    ms = phase(True, False); res["h2d_alone_gbs_per_rank"] = args.h2d_mb / ms
    ms = phase(False, True); res["d2h_alone_gbs_per_rank"] = args.d2h_mb / ms
    ms = phase(True, True); res["both_ms"] = ms
    res["both_gbs_per_rank"] = (args.h2d_mb + args.d2h_mb) / ms
    res["both_gbs_aggregate"] = world * res["both_gbs_per_rank"]
    # the e2e step of bench.py moves exactly these volumes per rank: its ceiling in voxel-samples/s
    res["e2e_ceiling_voxel_samples_per_s"] = world * 16 * 8 * 64 ** 3 / (ms * 1e-3)
    if rank == 0:
        print(json.dumps(res))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
