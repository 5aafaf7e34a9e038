"""Pinned host -> device copy rate of one layer call's token matrix (T x D fp32 = 59 MB), as one copy or split over
several streams: python tools/h2d_rate.py"""
import torch

dev = torch.device("cuda:0")
n = 38432 * 384
h = torch.empty(n, dtype=torch.float32).pin_memory()
d = torch.empty(n, dtype=torch.float32, device=dev)
for parts in (1, 2, 4):
    streams = [torch.cuda.Stream() for _ in range(parts)]
    step = (n + parts - 1) // parts
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for it in range(6):
        torch.cuda.synchronize()
        a.record()
        for r in range(12):
            for i, s in enumerate(streams):
                s.wait_event(a) if r == 0 else None
                with torch.cuda.stream(s):
                    d[i * step:(i + 1) * step].copy_(h[i * step:(i + 1) * step], non_blocking=True)
        for s in streams:
            torch.cuda.current_stream().wait_stream(s)
        b.record()
        b.synchronize()
        best = min(best, a.elapsed_time(b))
    print(f"{parts} stream(s): 12 x {n * 4 / 1e6:.0f} MB in {best:.2f} ms = {12 * n * 4 / best / 1e6:.1f} GB/s")
