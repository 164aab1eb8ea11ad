import torch, time
def t(fn, n=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n
for gb in (1, 4, 10):
    n = gb * (1 << 30)
    d = torch.empty(n, dtype=torch.uint8, device='cuda')
    t0 = time.perf_counter(); h = torch.empty(n, dtype=torch.uint8, pin_memory=True); ta = time.perf_counter() - t0
    dt = t(lambda: h.copy_(d, non_blocking=True))
    dt2 = t(lambda: d.copy_(h, non_blocking=True))
    print('%d GiB: pin alloc %.2fs  D2H %.1f GB/s  H2D %.1f GB/s' % (gb, ta, n / dt / 1e9, n / dt2 / 1e9), flush=True)
    # 2D strided D2H: rows of 512 KiB out of a 4 MiB pitch
    if gb == 4:
        dv = d.view(-1, 4 << 20); hv = h.view(-1, 4 << 20)
        dt3 = t(lambda: hv[:, :512 << 10].copy_(dv[:, :512 << 10], non_blocking=True))
        print('   2D 512KiB/4MiB pitch D2H %.1f GB/s' % (dv.shape[0] * (512 << 10) / dt3 / 1e9), flush=True)
    del d, h
