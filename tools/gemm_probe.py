"""Per-shape, per-epilogue timing of t2h_linear_fwd_f16 / t2h_linear_wgrad_f16 (CUDA events, inputs > L2).

    python tools/gemm_probe.py [rows]
"""
import sys

import torch

sys.path.insert(0, ".")
from tomosar2height_b200 import _lib  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1048576
dev = torch.device("cuda")
ptr = _lib.ptr


def slot_of(*ts):
    s = torch.zeros(1, dtype=torch.int32, device=dev)
    a, b = ts[0], (ts[1] if len(ts) > 1 else None)
    _lib.call("t2h_absmax", ptr(a), a.stride(0), a.shape[1], ptr(b), 0 if b is None else b.stride(0),
              0 if b is None else b.shape[1], a.shape[0], ptr(s))
    return s


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


print(f"rows = {rows}")
for K, N in [(256, 512), (512, 256), (128, 256), (256, 128), (512, 1024), (1024, 512)]:
    x = torch.randn(rows, K, device=dev)
    w = torch.randn(N, K, device=dev) / K ** 0.5
    bias = torch.randn(N, device=dev)
    aux = torch.randn(rows, N, device=dev)
    out = torch.empty(rows, N, device=dev)
    ws_, xs_ = slot_of(w), slot_of(x)
    hi, lo = torch.empty(N, K, dtype=torch.float16, device=dev), torch.empty(N, K, dtype=torch.float16, device=dev)
    _lib.call("t2h_split_f16", ptr(w), w.numel(), ptr(ws_), ptr(hi), ptr(lo))
    oslot = torch.zeros(1, dtype=torch.int32, device=dev)
    flops = 2.0 * rows * K * N

    def fwd(bias_=None, relu=False, mask=None, res=None, omax=None):
        _lib.call("t2h_linear_fwd_f16", ptr(x), x.stride(0), K, None, 0, 0, rows, ptr(xs_), ptr(hi), ptr(lo), ptr(ws_), N,
                  ptr(bias_), int(relu), ptr(mask), 0 if mask is None else mask.stride(0), ptr(res),
                  0 if res is None else res.stride(0), ptr(out), out.stride(0), ptr(omax))

    line = f"fwd  K={K:5d} N={N:5d}:"
    for name, kw in [("plain", {}), ("bias+relu_in+outmax", dict(bias_=bias, relu=True, omax=oslot)),
                     ("mask", dict(mask=aux)), ("residual", dict(bias_=bias, res=aux)),
                     ("mask+res", dict(mask=aux, res=aux))]:
        ms = timeit(lambda: fwd(**kw))
        line += f"  {name} {ms:.3f} ms {flops / ms * 1e-9:6.1f} TF/s |"
    print(line)
    g = torch.randn(rows, N, device=dev)
    gs = slot_of(g)
    n_ws = int(_lib.load().t2h_linear_wgrad_workspace_bytes(rows, N, K))
    wsb = torch.empty(n_ws, dtype=torch.uint8, device=dev)
    dw, db = torch.empty(N, K, device=dev), torch.empty(N, device=dev)
    ms = timeit(lambda: _lib.call("t2h_linear_wgrad_f16", ptr(g), g.stride(0), ptr(gs), ptr(x), x.stride(0), ptr(xs_), rows, N, K,
                                  1, ptr(wsb), n_ws, ptr(dw), dw.stride(0), ptr(db)))
    print(f"wgrad K={K:5d} N={N:5d}:  {ms:.3f} ms {flops / ms * 1e-9:6.1f} TF/s")
    del x, w, aux, out, g
