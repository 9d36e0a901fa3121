"""Config-5 microbenchmark: segmented max/mean and bilinear sample fwd+bwd vs the HBM roofline.

python tools/bench_ops.py [--full]   -> one JSON line per (op, N, R, C)
Points: B tiles of 262144 clustered points (N = B * 262144 total), plane R x R, C channels.
Bytes are the ALGORITHMIC bytes of SURVEY.md §8(d) (tomosar2height_b200/profiling.py).
"""
import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden"))
from cases import synthetic_cloud
import tomosar2height_b200.functional as T
from tomosar2height_b200.topology import Topology

PEAK = 6547.8
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass


def timeit(fn, n=10, flush=None):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        if flush is not None:
            flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    full = "--full" in sys.argv
    tiles_list = [4, 16, 64, 244] if full else [4]          # 1M, 4M, 16M, 64M points
    grid = [(R, C) for R in ([64, 128, 256, 512] if full else [32, 64, 128, 256]) for C in ([32, 64, 128] if full else [32, 128, 512])]
    if "--quick" in sys.argv:
        grid = [(256, 32), (128, 128), (32, 512)]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2
    for tiles in tiles_list:
        n_per = 262144
        cloud = synthetic_cloud(tiles, n_per, seed=1, clustered="--uniform" not in sys.argv).cuda()
        n = tiles * n_per
        for R, C in grid:
            if n * C * 4 * 3 > 60e9:
                continue
            topo = Topology(cloud, max(R, 256))
            lvl = topo.level(R)
            M = tiles * R * R
            rows = torch.randn(n, C, device="cuda")
            plane = torch.randn(tiles, R, R, C, device="cuda")
            gplane = torch.randn(M, C, device="cuda")
            res = {}
            if C <= 128:
                pooled, arg = T.seg_max_pool(rows, lvl, return_arg=True)
                res["S1_fwd seg_max+gather"] = (timeit(lambda: T.seg_max_pool(rows, lvl), flush=flush), 8 * n * C + 4 * n + 4 * M * C)
                r2 = rows.clone().requires_grad_(True)
                out = T.seg_max_pool(r2, lvl)
                res["S1_bwd"] = (timeit(lambda: torch.autograd.grad(out, r2, rows, retain_graph=True), flush=flush), 8 * n * C + 4 * M * C)
            res["S2_fwd seg_mean"] = (timeit(lambda: T.seg_mean(rows, lvl), flush=flush), 4 * n * C + 4 * n + 4 * M * C)
            r3 = rows.clone().requires_grad_(True)
            pm = T.seg_mean(r3, lvl)
            res["S2_bwd"] = (timeit(lambda: torch.autograd.grad(pm, r3, gplane, retain_graph=True), flush=flush), 4 * n * C + 4 * n + 4 * M * C)
            res["G1_fwd sample"] = (timeit(lambda: T.bilinear_sample(plane, lvl), flush=flush), 4 * M * C + 8 * n + 4 * n * C)
            p2 = plane.clone().requires_grad_(True)
            sm = T.bilinear_sample(p2, lvl)
            res["G2_bwd sample"] = (timeit(lambda: torch.autograd.grad(sm, p2, rows, retain_graph=True), flush=flush), 4 * n * C + 8 * n + 4 * M * C)
            for op, (ms, nbytes) in res.items():
                gbs = nbytes / ms / 1e6
                print(json.dumps({"op": op, "points": n, "R": R, "C": C, "ms": round(ms, 4), "GB/s": round(gbs, 1), "frac_of_hbm_peak": round(gbs / PEAK, 3)}))
            del topo, rows, plane, gplane, r3, pm, p2, sm
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
