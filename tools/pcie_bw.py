"""PCIe copy bandwidth of the box: pinned host <-> device, one stream vs two (does a second copy engine help?)."""
import time, torch
dev = torch.device('cuda', 0)
n = 400 * 1024 * 1024 // 8
d = torch.empty(n, dtype=torch.float64, device=dev).normal_()
h = torch.empty(n, dtype=torch.float64, pin_memory=True)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, reps=5):
    fn(); torch.cuda.synchronize(); best = 1e9
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
    return n * 8 / best / 1e9
def d2h_1():
    with torch.cuda.stream(s1): h.copy_(d, non_blocking=True)
def d2h_2():
    m = n // 2
    with torch.cuda.stream(s1): h[:m].copy_(d[:m], non_blocking=True)
    with torch.cuda.stream(s2): h[m:].copy_(d[m:], non_blocking=True)
def h2d_1():
    with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
def h2d_2():
    m = n // 2
    with torch.cuda.stream(s1): d[:m].copy_(h[:m], non_blocking=True)
    with torch.cuda.stream(s2): d[m:].copy_(h[m:], non_blocking=True)
def both():
    with torch.cuda.stream(s1): h.copy_(d, non_blocking=True)
    with torch.cuda.stream(s2): d2.copy_(h2, non_blocking=True)
d2 = torch.empty_like(d); h2 = torch.empty(n, dtype=torch.float64, pin_memory=True)
print('D2H one stream  %.1f GB/s' % t(d2h_1)); print('D2H two streams %.1f GB/s' % t(d2h_2))
print('H2D one stream  %.1f GB/s' % t(h2d_1)); print('H2D two streams %.1f GB/s' % t(h2d_2))
print('D2H + H2D concurrently: %.1f GB/s each direction' % t(both))
