python - <<'PY'
import os, time
print("cpu_count", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
for f in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us", "/sys/fs/cgroup/cpu/cpu.cfs_period_us"):
    try: print(f, open(f).read().strip())
    except Exception as e: print(f, "n/a")
print(open("/proc/loadavg").read())
import torch
print("torch threads default", torch.get_num_threads())
a = torch.randn(2048, 2560).bfloat16(); w = torch.randn(9728, 2560).bfloat16()
for nt in (128, 32, 16, 8, 4):
    torch.set_num_threads(nt)
    torch.nn.functional.linear(a[:64], w)
    t=time.time()
    for _ in range(3): torch.nn.functional.linear(a[:64], w)
    t1=(time.time()-t)/3
    t=time.time()
    for _ in range(3): torch.nn.functional.linear(a[:1], w)
    t2=(time.time()-t)/3
    print(f"threads {nt}: linear 64x2560x9728 bf16 {t1*1e3:.1f} ms ; 1x {t2*1e3:.1f} ms")
PY
