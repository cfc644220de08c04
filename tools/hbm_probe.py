"""HBM probe: pure write (memset), pure read (sum) and copy bandwidth on this GPU, for the roofline of the streaming kernels."""
import torch, time
n = 3 * 2 ** 30            # doubles: 24 GiB
a = torch.empty(n, dtype=torch.float64, device="cuda")
b = torch.empty(n // 2, dtype=torch.float64, device="cuda")
def timeit(f, reps=5):
    f(); torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(reps):
        ev0.record(); f(); ev1.record(); torch.cuda.synchronize()
        best = min(best, ev0.elapsed_time(ev1))
    return best
t = timeit(lambda: a.zero_()); print(f"memset  {a.numel() * 8 / t / 1e6:8.1f} GB/s")
t = timeit(lambda: a.fill_(1.5)); print(f"fill    {a.numel() * 8 / t / 1e6:8.1f} GB/s")
t = timeit(lambda: a.sum()); print(f"read    {a.numel() * 8 / t / 1e6:8.1f} GB/s")
t = timeit(lambda: b.copy_(a[: n // 2])); print(f"copy    {2 * b.numel() * 8 / t / 1e6:8.1f} GB/s (read + write)")
