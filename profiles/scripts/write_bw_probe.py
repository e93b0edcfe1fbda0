"""Write-only, read-only and copy bandwidth of the box's B200 (torch kernels as the yardstick): is a pure write stream
(what the splice is when its parents sit in L2) bounded by the same figure as a copy?"""
import torch

def timed(fn, reps=10):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best

n = 1 << 30                                  # 4 GiB of int32
x = torch.empty(n, dtype=torch.int32, device="cuda")
y = torch.empty(n, dtype=torch.int32, device="cuda")
gb = n * 4 / 1e9
t = timed(lambda: x.zero_());            print("write only (zero_ 4 GiB):   %.3f ms  %.0f GB/s" % (t, gb / t * 1e3))
t = timed(lambda: x.fill_(7));           print("write only (fill_ 4 GiB):   %.3f ms  %.0f GB/s" % (t, gb / t * 1e3))
t = timed(lambda: torch.cuda.memset if False else x.zero_());
t = timed(lambda: y.copy_(x));           print("copy (read + write 8 GiB):  %.3f ms  %.0f GB/s" % (t, 2 * gb / t * 1e3))
t = timed(lambda: x.sum());              print("read only (sum 4 GiB):      %.3f ms  %.0f GB/s" % (t, gb / t * 1e3))
xs = x[: n // 4]
t = timed(lambda: xs.zero_());           print("write only (zero_ 1 GiB):   %.3f ms  %.0f GB/s" % (t, gb / 4 / t * 1e3))
