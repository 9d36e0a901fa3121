"""Per-shape timing of the tcgen05 linear kernels on the layer shapes of the Berlin model."""
import sys, torch
sys.path.insert(0, '.')
from tomosar2height_b200.linear import linear, _launch_fwd, _launch_wgrad, _cache, colsum

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 524288
shapes = [(64, 32), (32, 32), (32, 64), (64, 128), (128, 64), (128, 256), (256, 128), (256, 512), (512, 256), (512, 1024), (1024, 512)]
def timeit(fn, n=5):
    fn(); torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n
print(f"rows={rows}")
print(f"{'K':>5} {'N':>5} | {'fwd ms':>8} {'TF/s':>7} {'GB/s':>7} | {'wgrad ms':>8} {'TF/s':>7} {'GB/s':>7} | {'cublas ms':>9} {'TF/s':>6} | colsum ms")
for K, N in shapes:
    x = torch.randn(rows, K, device='cuda'); w = torch.randn(N, K, device='cuda') / K ** .5; b = torch.randn(N, device='cuda')
    gy = torch.randn(rows, N, device='cuda')
    hi, lo = _cache.get(w)
    out = torch.empty(rows, N, device='cuda')
    dw = torch.empty(N, K, device='cuda')
    t_f = timeit(lambda: _launch_fwd(x, None, hi, lo, N, b, True, None, None, out))
    t_w = timeit(lambda: _launch_wgrad(gy, x, True, dw))
    t_c = timeit(lambda: torch.addmm(b, x, w.t()))
    t_s = timeit(lambda: colsum(gy))
    fl = 2 * rows * K * N
    by = 4 * rows * (K + N)
    print(f"{K:5d} {N:5d} | {t_f:8.3f} {fl/t_f/1e9:7.1f} {by/t_f/1e6:7.0f} | {t_w:8.3f} {fl/t_w/1e9:7.1f} {by/t_w/1e6:7.0f} | {t_c:9.3f} {fl/t_c/1e9:6.1f} | {t_s:.3f}")
