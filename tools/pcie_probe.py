"""Scratch: pinned H2D / D2H bandwidth on the box (ceiling for the e2e number)."""
import time, torch
dev = torch.device("cuda:0")
for mb in (48, 96, 1024):
    n = mb << 20
    h = torch.empty(n, dtype=torch.uint8).pin_memory(); d = torch.empty(n, dtype=torch.uint8, device=dev)
    h2 = torch.empty(n, dtype=torch.uint8).pin_memory(); d2 = torch.empty(n, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    for name, fn in (("D2H", lambda: h.copy_(d, non_blocking=True)), ("H2D", lambda: d.copy_(h, non_blocking=True))):
        for _ in range(2): fn()
        torch.cuda.synchronize(); t = time.perf_counter()
        for _ in range(8): fn()
        torch.cuda.synchronize(); dt = time.perf_counter() - t
        print(f"{name} {mb} MiB: {8 * n / dt / 1e9:.1f} GB/s")
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(8):
        with torch.cuda.stream(s1): h.copy_(d, non_blocking=True)
        with torch.cuda.stream(s2): d2.copy_(h2, non_blocking=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t
    print(f"both {mb} MiB: {8 * n / dt / 1e9:.1f} GB/s each direction")
