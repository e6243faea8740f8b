"""Diagnostic (GPU box): one-direction and simultaneous two-direction PCIe copy rates with page-locked memory."""
import time
import torch
n = 1 << 28
h1 = torch.empty(n, dtype=torch.uint8).pin_memory(); h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
g1 = torch.empty(n, dtype=torch.uint8, device="cuda"); g2 = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(h2d, d2h, reps=6):
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1): g1.copy_(h1, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2): h2.copy_(g2, non_blocking=True)
    torch.cuda.synchronize(); return reps * n / (time.perf_counter() - t) / 1e9
run(True, True, 1)
print("h2d alone GB/s", run(True, False)); print("d2h alone GB/s", run(False, True))
print("both at once, GB/s per direction", run(True, True))
