"""D2H bandwidth per rank, alone and concurrently, with and without NUMA-local pinning."""
import os, sys, time, subprocess
import torch, torch.distributed as dist
rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1: dist.init_process_group("nccl", device_id=torch.device("cuda", local))
if rank == 0:
    print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout)
    print(subprocess.run(["bash", "-c", "lscpu | grep -i -E 'numa|socket|model name|^CPU\\(s\\)'"], capture_output=True, text=True).stdout)
def aff_for(idx):
    out = subprocess.run(["nvidia-smi", "topo", "-C", "-i", str(idx)], capture_output=True, text=True).stdout
    return out.strip()
print("rank", rank, "affinity now", len(os.sched_getaffinity(0)), "topo -C:", aff_for(local), flush=True)
n = 1 << 28   # 2 GiB of doubles
src = torch.empty(n, dtype=torch.float64, device="cuda").normal_()
def run(tag):
    dst = torch.empty(n, dtype=torch.float64, pin_memory=True)
    dst.copy_(src); torch.cuda.synchronize()
    for mode in ("alone", "concurrent"):
        if world > 1: dist.barrier()
        for r in range(world if mode == "alone" else 1):
            if world > 1: dist.barrier()
            if mode == "concurrent" or r == rank:
                t0 = time.perf_counter(); dst.copy_(src, non_blocking=True); torch.cuda.synchronize(); dt = time.perf_counter() - t0
                print("%s rank %d %s: %.1f GB/s" % (tag, rank, mode, n * 8 / dt / 1e9), flush=True)
            if world > 1: dist.barrier()
    del dst
run("default")
try:
    import pynvml
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(local)
    pynvml.nvmlDeviceSetCpuAffinity(h)
    print("rank", rank, "affinity after nvml", sorted(os.sched_getaffinity(0))[:4], len(os.sched_getaffinity(0)), flush=True)
    run("nvml-affinity")
except Exception as e:
    print("nvml affinity failed", e)
if world > 1: dist.destroy_process_group()
