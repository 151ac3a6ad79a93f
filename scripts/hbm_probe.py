"""Read-only HBM bandwidth references on this GPU: torch reductions and cuBLAS GEMV over the sample matrix."""
import torch
for N in (21504, 125000):
    S = torch.randn((N, 3000), device='cuda')
    q = torch.randn(3000, device='cuda')
    out = torch.empty(N, device='cuda')
    def timeit(fn, reps=50):
        for _ in range(5): fn()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps): fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps * 1e3
    gb = N * 3000 * 4 / 1e9
    for name, fn in (("torch.mv (cuBLAS gemv)", lambda: torch.mv(S, q, out=out)), ("S.sum()", lambda: S.sum()),
                     ("S.abs().max() fused? no: amax", lambda: torch.amax(S))):
        us = timeit(fn)
        print("N=%d %-32s %.1f us  %.0f GB/s" % (N, name, us, gb / us * 1e6))
    del S
