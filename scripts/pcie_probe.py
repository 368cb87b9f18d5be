#!/usr/bin/env python
"""Development helper (GPU box): host<->device copy rates from pinned memory, alone and both directions at once
(the e2e path of the batched entry is PCIe-bound once packing runs on the device)."""
import torch, time
dev = torch.device("cuda", 0)
for mb in (64, 1024):
    n = mb << 20
    h = torch.empty(n, dtype=torch.uint8).pin_memory(); h.random_(0, 5)
    h2 = torch.empty(n // 4, dtype=torch.uint8).pin_memory()
    d = torch.empty(n, dtype=torch.uint8, device=dev); d2 = torch.empty(n // 4, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    def t(fn, reps=5):
        fn(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps): fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / reps
    a = t(lambda: d.copy_(h, non_blocking=True))
    b = t(lambda: h2.copy_(d2, non_blocking=True))
    def both():
        with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
        with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
    c = t(both)
    print(f"{mb} MiB: H2D {n / a / 1e9:.1f} GB/s, D2H {n / 4 / b / 1e9:.1f} GB/s, both at once: {c * 1e3:.2f} ms "
          f"(H2D alone {a * 1e3:.2f} ms, D2H alone {b * 1e3:.2f} ms)")
import os
print("cpus", os.cpu_count())
