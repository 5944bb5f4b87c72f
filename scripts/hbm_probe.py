"""Write-only / read-only / copy bandwidth of this GPU at the size of the src pass's output (context for the src-pass roofline)."""
import json
import torch

n = 542080 * 1024 * 2 * 2  # bytes of dk + dv on the headline graph
x = torch.empty(n, dtype=torch.uint8, device="cuda")
y = torch.empty(n, dtype=torch.uint8, device="cuda")
big = torch.empty(6 * n // 4, dtype=torch.float32, device="cuda")


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


res = {}
ms = timeit(lambda: x.zero_())
res["memset_GBps"] = n / ms / 1e6
ms = timeit(lambda: y.copy_(x))
res["copy_GBps_read_plus_write"] = 2 * n / ms / 1e6
xs = big
ms = timeit(lambda: xs.sum())
res["read_sum_GBps"] = xs.numel() * 4 / ms / 1e6
print(json.dumps(res))
