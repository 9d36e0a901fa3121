"""Per-shape timing of the tcgen05 linear kernels on the layer shapes of the Berlin model."""
import sys, torch
sys.path.insert(0, '.')
from tomosar2height_b200.linear import linear, _launch_fwd, _launch_wgrad, _cache, colsum, use_f16
from tomosar2height_b200 import _lib
from tomosar2height_b200._lib import ptr

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 524288
if len(sys.argv) > 3:
    only = [(int(sys.argv[2]), int(sys.argv[3]))]
else:
    only = None
shapes = only or [(64, 32), (32, 32), (32, 64), (64, 128), (128, 64), (128, 256), (256, 128), (256, 512), (512, 256), (512, 1024), (1024, 512)]
def timeit(fn, n=5):
    fn(); torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n
print(f"rows={rows}")
print(f"{'K':>5} {'N':>5} | {'fwd ms':>8} {'TF/s':>7} {'GB/s':>7} | {'wgrad ms':>8} {'TF/s':>7} {'GB/s':>7} | {'cublas ms':>9} {'TF/s':>6} | colsum ms | tf32x3 fwd ms TF/s | absmax ms | f16 max rel err")
for K, N in shapes:
    x = torch.randn(rows, K, device='cuda'); w = torch.randn(N, K, device='cuda') / K ** .5; b = torch.randn(N, device='cuda')
    gy = torch.randn(rows, N, device='cuda')
    split = _cache.get(w, f16=use_f16(N, K))
    split32 = _cache.get(w)
    out = torch.empty(rows, N, device='cuda')
    dw = torch.empty(N, K, device='cuda')
    t_f = timeit(lambda: _launch_fwd(x, None, split, N, b, True, None, None, out))
    t_32 = timeit(lambda: _launch_fwd(x, None, split32, N, b, True, None, None, out))
    slot = torch.empty(1, dtype=torch.int32, device='cuda')
    t_a = timeit(lambda: _lib.call("t2h_absmax", ptr(x), K, K, None, 0, 0, rows, ptr(slot)))
    _launch_fwd(x, None, split, N, b, True, None, None, out)
    ref = torch.addmm(b.double(), x[:4096].relu().double(), w.double().t())
    err = ((out[:4096].double() - ref).abs().max() / ref.abs().max()).item()
    t_w = timeit(lambda: _launch_wgrad(gy, x, True, dw))
    refw = gy.double().t() @ x.relu().double()
    err_w = ((dw.double() - refw).abs().max() / refw.abs().max()).item()
    t_c = timeit(lambda: torch.addmm(b, x, w.t()))
    t_s = timeit(lambda: colsum(gy))
    fl = 2 * rows * K * N
    by = 4 * rows * (K + N)
    print(f"{K:5d} {N:5d} | {t_f:8.3f} {fl/t_f/1e9:7.1f} {by/t_f/1e6:7.0f} | {t_w:8.3f} {fl/t_w/1e9:7.1f} {by/t_w/1e6:7.0f} | {t_c:9.3f} {fl/t_c/1e9:6.1f} | {t_s:.3f} | {t_32:8.3f} {fl/t_32/1e9:7.1f} | {t_a:.3f} | {err:.2e} | wgrad err {err_w:.2e}")
