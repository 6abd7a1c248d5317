"""Diagnostic: pinned D2H bandwidth alone and while kernels run on another stream (why is e2e PCIe-bound?)."""
import time
import torch

dev = torch.device("cuda:0")
n = 176947200  # sigma of C3: 19.66M nodes * 9 doubles
src = torch.randn(n, dtype=torch.float64, device=dev)
dst = torch.empty(n, dtype=torch.float64, pin_memory=True)
side = torch.cuda.Stream()
a = torch.randn(64 * 1024 * 1024, dtype=torch.float64, device=dev)
b = torch.empty_like(a)


def d2h(label, busy):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(side):
        e0.record(side)
        dst.copy_(src, non_blocking=True)
        e1.record(side)
    if busy:
        for _ in range(200):
            b.copy_(a)          # HBM-bound kernels on the default stream
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"{label}: {n * 8 / ms / 1e6:.1f} GB/s ({ms:.1f} ms)")


for _ in range(2):
    d2h("D2H alone", False)
d2h("D2H with concurrent HBM-bound kernels", True)
h = torch.empty(19660800, dtype=torch.float64, pin_memory=True)
t = torch.empty(19660800, dtype=torch.float64, device=dev)
torch.cuda.synchronize()
t0 = time.time()
t.copy_(h, non_blocking=True)
torch.cuda.synchronize()
print(f"H2D 157 MB: {157.3 / (time.time() - t0) / 1e3:.1f} GB/s")
