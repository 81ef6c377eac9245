"""HBM bandwidth by access mix on this GPU (torch kernels, CUDA events, best of 10): write-only (fill), read-only (sum),
copy (read + write).  MEASURED_PEAKS.json's hbm_gbs is the copy figure (read + write bytes)."""
import torch

n = 1 << 30   # 1 Gi floats = 4 GiB
a = torch.empty(n, dtype=torch.float32, device="cuda")
b = torch.empty(n, dtype=torch.float32, device="cuda")


def best(fn, bytes_moved, reps=10):
    fn(); torch.cuda.synchronize()
    t = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        t.append(e0.elapsed_time(e1))
    return bytes_moved / (min(t) * 1e-3) / 1e9


print(f"write-only  (fill_)  : {best(lambda: a.fill_(1.0), n * 4):8.0f} GB/s")
print(f"write-only  (zero_)  : {best(lambda: a.zero_(), n * 4):8.0f} GB/s")
print(f"read-only   (sum)    : {best(lambda: a.sum(), n * 4):8.0f} GB/s")
print(f"copy        (copy_)  : {best(lambda: b.copy_(a), n * 8):8.0f} GB/s (read + write bytes)")
print(f"read 2 : write 1 (add): {best(lambda: torch.add(a, b, out=b), n * 12):8.0f} GB/s (all bytes)")
