import torch, time
d = torch.empty(1473 * 2**20 // 4, dtype=torch.float32, device="cuda")
h = torch.empty_like(d, device="cpu", pin_memory=True)
d2 = torch.empty(853 * 2**20 // 4, dtype=torch.float32, device="cuda")
h2 = torch.empty_like(d2, device="cpu", pin_memory=True)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(both):
    torch.cuda.synchronize(); t = time.perf_counter()
    with torch.cuda.stream(s1): h.copy_(d, non_blocking=True)
    if both:
        with torch.cuda.stream(s2): d2.copy_(h2, non_blocking=True)
    torch.cuda.synchronize(); return time.perf_counter() - t
for both in (False, True):
    run(both); ts = [run(both) for _ in range(5)]
    t = min(ts)
    print("both" if both else "d2h only", "%.2f ms  d2h %.1f GB/s" % (t * 1e3, d.numel() * 4 / t / 1e9), ("h2d %.1f GB/s" % (d2.numel() * 4 / t / 1e9)) if both else "")
