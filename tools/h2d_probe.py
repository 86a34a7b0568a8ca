import torch, time
n = 1 << 30
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
for chunk in (n, 8 << 20, 1 << 20):
    torch.cuda.synchronize(); t = time.perf_counter()
    for o in range(0, n, chunk):
        d[o:o + chunk].copy_(h[o:o + chunk], non_blocking=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t
    print(f"H2D pinned chunk={chunk>>20}MiB: {n/dt/1e9:.1f} GB/s")
torch.cuda.synchronize(); t = time.perf_counter(); h.copy_(d); torch.cuda.synchronize(); print(f"D2H pinned: {n/(time.perf_counter()-t)/1e9:.1f} GB/s")
