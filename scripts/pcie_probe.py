"""Host<->device copy bandwidth of the box (context for bench.py's `e2e`): pinned 1 GiB buffers, H2D alone, D2H alone, both at once
on two streams.  Prints one JSON line."""
import json
import time

import torch

n = 1 << 30
h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_a = torch.empty(n, dtype=torch.uint8, device="cuda")
d_b = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(h2d, d2h, reps=5):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1):
                d_a.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                h_out.copy_(d_b, non_blocking=True)
    torch.cuda.synchronize()
    return reps * n / (time.perf_counter() - t0) / 1e9


run(True, True, 1)
print(json.dumps({"h2d_alone_GBps": round(run(True, False), 1), "d2h_alone_GBps": round(run(False, True), 1),
                  "both_at_once_GBps_per_direction": round(run(True, True), 1)}))
