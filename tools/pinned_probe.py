"""Does the host-link bandwidth depend on WHICH pinned allocation a transfer uses (hypervisor / NUMA placement),
or on when it runs?  Allocates several pinned buffers, times H2D and D2H of each in rounds.  Development aid,
run under gpurun:  python tools/pinned_probe.py [n_buffers] [MiB]"""
import sys, time
import torch

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
mib = int(sys.argv[2]) if len(sys.argv) > 2 else 64
dev = torch.device("cuda:0")
d = torch.empty(mib << 20, dtype=torch.uint8, device=dev)
bufs = [torch.empty(mib << 20, dtype=torch.uint8).pin_memory() for _ in range(n)]
for b in bufs:
    b.fill_(1)


def bw(fn, reps=5):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record(); torch.cuda.synchronize()
    return (mib << 20) * reps / (a.elapsed_time(b) * 1e-3) / 1e9


for rnd in range(4):
    h2d = [bw(lambda b=b: d.copy_(b, non_blocking=True)) for b in bufs]
    d2h = [bw(lambda b=b: b.copy_(d, non_blocking=True)) for b in bufs]
    print(f"round {rnd}: H2D GB/s " + " ".join(f"{x:5.1f}" for x in h2d) + "   D2H GB/s " + " ".join(f"{x:5.1f}" for x in d2h), flush=True)
    time.sleep(1.0)
# both directions at once on two streams (what the pipeline does)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
d2 = torch.empty_like(d)
for i in range(0, n - 1, 2):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    s1.wait_event(a); s2.wait_event(a)
    for _ in range(5):
        with torch.cuda.stream(s1):
            d.copy_(bufs[i], non_blocking=True)
        with torch.cuda.stream(s2):
            bufs[i + 1].copy_(d2, non_blocking=True)
    torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)
    b.record(); torch.cuda.synchronize()
    t = a.elapsed_time(b) * 1e-3
    print(f"duplex buffers {i},{i+1}: {(mib << 20) * 5 / t / 1e9:5.1f} GB/s each way")
