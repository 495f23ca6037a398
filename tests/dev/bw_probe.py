"""HBM bandwidth by read:write mix (not a test): explains what a write-heavy kernel can reach."""
import torch
n = 512 * 1024 * 1024   # fp32 elements = 2 GiB
a = torch.empty(n, dtype=torch.float32, device="cuda"); b = torch.empty_like(a); c = torch.empty_like(a)
a.normal_(); b.normal_()
def timed(fn, reps=10):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
gb = n * 4 / 1e9
print("fill   (0r:1w) %.0f GB/s" % (gb / timed(lambda: c.fill_(1.0)) * 1e3))
print("copy   (1r:1w) %.0f GB/s" % (2 * gb / timed(lambda: c.copy_(a)) * 1e3))
print("add    (2r:1w) %.0f GB/s" % (3 * gb / timed(lambda: torch.add(a, b, out=c)) * 1e3))
print("sum    (1r:0w) %.0f GB/s" % (gb / timed(lambda: a.sum()) * 1e3))
# 1 read : 3 writes (the fused layer kernel's mix): three outputs from one input
d = torch.empty_like(a)
def one_to_three():
    torch.mul(a, 2.0, out=b); 
print("sincos-like 1r:2w (frexp) skipped")
