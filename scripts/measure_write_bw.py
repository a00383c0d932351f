"""Measure HBM write-only vs copy bandwidth on the box (context for the slot kernel's roofline:
its traffic is ~100 % writes, while MEASURED_PEAKS.json's hbm_gbs is a read+write copy)."""
import json
import torch

def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e-3

nbytes = 8 << 30
x = torch.empty(nbytes // 4, dtype=torch.float32, device="cuda")
y = torch.empty(nbytes // 4, dtype=torch.float32, device="cuda")
res = {}
res["fill_gbs"] = nbytes / timeit(lambda: x.fill_(1.0)) / 1e9
res["zero_gbs"] = nbytes / timeit(lambda: x.zero_()) / 1e9
res["copy_gbs"] = 2 * nbytes / timeit(lambda: y.copy_(x)) / 1e9
res["read_sum_gbs"] = nbytes / timeit(lambda: x.sum()) / 1e9
xc = x.view(torch.complex64)
res["fill_complex_gbs"] = nbytes / timeit(lambda: xc.fill_(1 + 1j)) / 1e9
print(json.dumps(res))
