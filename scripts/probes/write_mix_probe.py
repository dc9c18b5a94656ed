"""hardware probe (not product code): HBM bandwidth of a pure write stream, a pure read stream and the copy that
MEASURED_PEAKS.json is defined by, with plain torch ops -- context for the C2 roofline (the headline kernel writes 4 bytes
for every byte it reads)."""
import torch
dev = torch.device("cuda:0")
n = 1 << 30  # bytes
a = torch.empty(2 * n, dtype=torch.uint8, device=dev)
b = torch.empty(2 * n, dtype=torch.uint8, device=dev)
def timed(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
t = timed(lambda: a.zero_());                 print("write only  (zero_ 2 GiB)        %7.1f GB/s" % (2 * n / t / 1e6))
t = timed(lambda: a.view(torch.int64).sum()); print("read only   (sum of 2 GiB)       %7.1f GB/s" % (2 * n / t / 1e6))
t = timed(lambda: b.copy_(a));                print("copy        (2 GiB -> 2 GiB)     %7.1f GB/s (read + write bytes)" % (4 * n / t / 1e6))
a16 = a.view(torch.float16)[: n // 4]   # 0.5 GiB read
t = timed(lambda: torch.add(a16.float(), 1.0, out=b.view(torch.float32)[: n // 4][: a16.numel()]) if False else b.view(torch.float32)[: a16.numel()].copy_(a16));
print("1 : 2 mix   (0.5 GiB f16 -> 1 GiB f32 convert) %7.1f GB/s (read + write bytes)" % ((a16.numel() * 6) / t / 1e6))
a8 = a[: n // 2]
t = timed(lambda: b.view(torch.float32)[: a8.numel()].copy_(a8));
print("1 : 4 mix   (0.5 GiB u8 -> 2 GiB f32 convert)  %7.1f GB/s (read + write bytes)" % ((a8.numel() * 5) / t / 1e6))
