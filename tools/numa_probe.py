"""Development aid: PCIe H2D bandwidth of every GPU from pinned memory, one process per GPU at the same time,
with and without binding the process to the GPU's NUMA-local cores (from nvidia-smi topo / sysfs)."""
import os, sys, subprocess, time
import torch
import torch.distributed as dist

rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); lr = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))

def cpus_of_gpu(i):
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(i)
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        bus = bus.lower()
        if len(bus.split(":")[0]) == 8: bus = bus[4:]
        node = open(f"/sys/bus/pci/devices/{bus}/numa_node").read().strip()
        cl = open(f"/sys/bus/pci/devices/{bus}/local_cpulist").read().strip()
        return node, cl
    except Exception as e:
        return "?", str(e)

def parse(cl):
    out = set()
    for part in cl.split(","):
        if "-" in part:
            a, b = part.split("-"); out.update(range(int(a), int(b) + 1))
        elif part: out.add(int(part))
    return out

def bw(nbytes=64 << 20, reps=20):
    h = torch.empty(nbytes, dtype=torch.uint8).pin_memory(); h.fill_(1)
    d = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    for _ in range(3): d.copy_(h, non_blocking=True)
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): d.copy_(h, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    return nbytes * reps / (e0.elapsed_time(e1) * 1e-3) / 1e9

node, cl = cpus_of_gpu(lr)
aff0 = sorted(os.sched_getaffinity(0))
b0 = bw()
try:
    os.sched_setaffinity(0, parse(cl) & set(aff0) or set(aff0))
    b1 = bw()
except Exception as e:
    b1 = float("nan")
res = torch.tensor([b0, b1], device="cuda"); allr = [torch.zeros_like(res) for _ in range(world)]
dist.all_gather(allr, res)
print(f"rank {rank}: gpu numa node {node}, local cpus {cl[:40]}, initial affinity {len(aff0)} cpus [{aff0[0]}..{aff0[-1]}], H2D {b0:.1f} -> bound {b1:.1f} GB/s", flush=True)
if rank == 0:
    print("sum unbound %.1f GB/s, bound %.1f GB/s" % (sum(float(a[0]) for a in allr), sum(float(a[1]) for a in allr)))
    print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout[:3000])
dist.destroy_process_group()
