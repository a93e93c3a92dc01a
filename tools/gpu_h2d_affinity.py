"""Developer tool: does the H2D bandwidth of a pinned buffer depend on the vCPU that allocated and first
touched it?  (The boxes are VMs that show one NUMA node; the e2e figure moves by 30 % between processes.)"""
import os, time, torch
T, N, V = 1000, 256, 48
dev = torch.empty(T, N, V, device="cuda")
all_cpus = sorted(os.sched_getaffinity(0))
res = []
for cpu in all_cpus:
    os.sched_setaffinity(0, {cpu})
    host = torch.empty(T, N, V).fill_(1.0).pin_memory()      # allocated and touched on this vCPU
    for _ in range(2):
        dev.copy_(host, non_blocking=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        dev.copy_(host, non_blocking=True)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 10
    res.append((cpu, host.numel() * 4 / dt / 1e9))
    del host
os.sched_setaffinity(0, set(all_cpus))
print("GB/s by allocating vCPU:", " ".join(f"{c}:{g:.1f}" for c, g in res))
# and: same buffer, copy ISSUED from different vCPUs
host = torch.empty(T, N, V).fill_(1.0).pin_memory()
out = []
for cpu in all_cpus[::3]:
    os.sched_setaffinity(0, {cpu})
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(10):
        dev.copy_(host, non_blocking=True)
    torch.cuda.synchronize()
    out.append((cpu, host.numel() * 4 / ((time.perf_counter() - t0) / 10) / 1e9))
print("GB/s by issuing vCPU (one buffer):", " ".join(f"{c}:{g:.1f}" for c, g in out))
