"""Host<->device copy ceiling of the box, all GPUs at once (the bound of the end-to-end figure).

    python tools/pcie_probe.py                                                        # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
        tools/pcie_probe.py                                                           # 8 GPUs concurrently

Every rank binds to the CPUs of its GPU's NUMA node when the box exposes more than one (bench.bind_near_gpu), allocates
pinned buffers of the bench step's sizes (853 MB in, 1473 MB out), and times -- all ranks between the same barriers,
max over ranks -- D2H alone, H2D alone and both directions at once.  Rank 0 prints one JSON line with the per-GPU and
aggregate rates, the topology (`nvidia-smi topo -m`), the NUMA nodes the process can see and the clocks."""
import json
import os
import subprocess
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import ClockSampler, bind_near_gpu  # noqa: E402


def main():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = bind_near_gpu(local) if "--no-bind" not in sys.argv else "binding disabled (--no-bind)"
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
    n_out, n_in = 1473 * 2 ** 20 // 4, 853 * 2 ** 20 // 4
    d_out = torch.empty(n_out, dtype=torch.float32, device=dev)
    h_out = torch.empty(n_out, dtype=torch.float32, pin_memory=True)
    d_in = torch.empty(n_in, dtype=torch.float32, device=dev)
    h_in = torch.empty(n_in, dtype=torch.float32, pin_memory=True)
    h_in.zero_()
    h_out.zero_()
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def run(d2h, h2d):
        barrier()
        t = time.perf_counter()
        if d2h:
            with torch.cuda.stream(s1):
                h_out.copy_(d_out, non_blocking=True)
        if h2d:
            with torch.cuda.stream(s2):
                d_in.copy_(h_in, non_blocking=True)
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        return float(dt.item())

    sampler = ClockSampler(local)
    sampler.start()
    res = {}
    for name, d2h, h2d in (("d2h_only", True, False), ("h2d_only", False, True), ("both", True, True)):
        run(d2h, h2d)
        t = min(run(d2h, h2d) for _ in range(5))
        res[name] = {"ms": 1e3 * t,
                     "d2h_GBs_per_gpu": n_out * 4 / t / 1e9 if d2h else None,
                     "h2d_GBs_per_gpu": n_in * 4 / t / 1e9 if h2d else None,
                     "aggregate_GBs": world * ((n_out * 4 if d2h else 0) + (n_in * 4 if h2d else 0)) / t / 1e9}
    clocks = sampler.stop()
    if rank == 0:
        nodes = sorted(d for d in os.listdir("/sys/devices/system/node") if d.startswith("node") and d[4:].isdigit()) \
            if os.path.isdir("/sys/devices/system/node") else []
        try:
            topo = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=20).stdout
        except Exception as e:
            topo = "unavailable: %s" % e
        print(json.dumps({"n_gpus": world, "host_cpus": len(os.sched_getaffinity(0)), "numa_nodes_visible": nodes,
                          "rank0_binding": numa, "copies": res, "clocks": clocks, "topology": topo.splitlines()}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
