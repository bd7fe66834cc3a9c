import os, time, torch
print("cpu_count", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
for f in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us", "/sys/fs/cgroup/cpu/cpu.cfs_period_us"):
    try:
        print(f, open(f).read().strip())
    except Exception as e:
        print(f, "n/a")
print("torch threads default", torch.get_num_threads(), torch.get_num_interop_threads())
a = torch.randn(4096, 4096); b = torch.randn(4096, 4096)
x = torch.randn(2, 320, 128, 96); w = torch.randn(320, 320, 3, 3)
for nt in (8, 16, 32, 64, 128):
    torch.set_num_threads(nt)
    a @ b
    t0 = time.time(); a @ b; a @ b; dt = (time.time() - t0) / 2
    torch.nn.functional.conv2d(x, w, padding=1)
    t1 = time.time(); torch.nn.functional.conv2d(x, w, padding=1); dc = time.time() - t1
    print(f"threads {nt}: matmul {2*4096**3/dt/1e9:.0f} GFLOP/s, conv {2*2*320*320*9*128*96/dc/1e9:.0f} GFLOP/s", flush=True)
