"""D2H of one 8.3 MB frame: one copy vs halves / quarters on separate streams (tools only)."""
import time, torch
n = 8294400
d = torch.empty(n, dtype=torch.uint8, device="cuda")
h = torch.empty(n, dtype=torch.uint8).pin_memory()
for parts in (1, 2, 4, 8):
    streams = [torch.cuda.Stream() for _ in range(parts)]
    step = n // parts
    def frame():
        for k, s in enumerate(streams):
            with torch.cuda.stream(s):
                h[k * step:(k + 1) * step].copy_(d[k * step:(k + 1) * step], non_blocking=True)
    for _ in range(20): frame()
    torch.cuda.synchronize()
    t0 = time.perf_counter(); reps = 400
    for _ in range(reps): frame()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"{parts} stream(s): {n * reps / dt / 1e9:.2f} GB/s, {dt / reps * 1e6:.1f} us per frame")
# one big copy for reference
big = torch.empty(1 << 30, dtype=torch.uint8, device="cuda"); hb = torch.empty(1 << 30, dtype=torch.uint8).pin_memory()
hb.copy_(big); torch.cuda.synchronize()
t0 = time.perf_counter(); hb.copy_(big, non_blocking=True); torch.cuda.synchronize()
print(f"1 GiB copy: {(1 << 30) / (time.perf_counter() - t0) / 1e9:.2f} GB/s")
