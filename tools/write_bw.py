"""Write-only / read-only / copy HBM bandwidth on this GPU (context for the Gram roofline)."""
import torch

n = 1 << 30  # doubles: 8 GiB
a = torch.empty(n, dtype=torch.float64, device="cuda")
b = torch.empty(n, dtype=torch.float64, device="cuda")


def timeit(fn, reps=5):
    fn()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


gb = n * 8e-9
t = timeit(lambda: a.fill_(1.5))
print(f"fill_ (write only)      : {t:.3f} ms  {gb / t * 1e3:.0f} GB/s")
t = timeit(lambda: a.zero_())
print(f"zero_ (memset)          : {t:.3f} ms  {gb / t * 1e3:.0f} GB/s")
t = timeit(lambda: b.copy_(a))
print(f"copy_ (read + write)    : {t:.3f} ms  {2 * gb / t * 1e3:.0f} GB/s (sum of both directions)")
t = timeit(lambda: a.sum())
print(f"sum (read only)         : {t:.3f} ms  {gb / t * 1e3:.0f} GB/s")
t = timeit(lambda: torch.add(a, 1.0, out=a))
print(f"add in place (r + w)    : {t:.3f} ms  {2 * gb / t * 1e3:.0f} GB/s (sum of both directions)")
