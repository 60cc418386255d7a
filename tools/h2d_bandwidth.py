import torch, time
for mb in (2, 16, 128):
    n = mb << 20
    h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    d = torch.empty(n, dtype=torch.uint8, device="cuda")
    for _ in range(3): d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): d.copy_(h, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    e0.record()
    for _ in range(10): h.copy_(d, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    ms2 = e0.elapsed_time(e1) / 10
    print(f"{mb} MiB: H2D {ms*1e3:.1f} us ({n/ms/1e6:.1f} GB/s)  D2H {ms2*1e3:.1f} us ({n/ms2/1e6:.1f} GB/s)")
