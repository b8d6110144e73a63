"""HBM write-bandwidth probes (pure write, strided 1 KB chunks) for the roofline discussion."""
import torch

F, L, T = 1024, 2562, 938
out = torch.empty((F, L, T), dtype=torch.complex64, device="cuda")
nbytes = out.numel() * 8


def timeit(fn, n=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


ms = timeit(lambda: out.view(torch.float32).zero_())
print("memset 19.7 GB: %.3f ms -> %.2f TB/s" % (ms, nbytes / ms / 1e9))
ms = timeit(lambda: out.fill_(1.0))
print("fill_  19.7 GB: %.3f ms -> %.2f TB/s" % (ms, nbytes / ms / 1e9))
src = torch.empty_like(out)
ms = timeit(lambda: out.copy_(src))
print("copy   2x19.7 GB: %.3f ms -> %.2f TB/s (read+write)" % (ms, 2 * nbytes / ms / 1e9))
for w in (128, 256, 512):
    sub = out[:, :, :w]
    ms = timeit(lambda: sub.fill_(1.0))
    print("strided fill, %4d B chunks at 7504 B stride: %.3f ms -> %.2f TB/s" % (w * 8, ms, sub.numel() * 8 / ms / 1e9))
