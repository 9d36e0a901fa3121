import os, sys, torch
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests", "golden"))
from cases import synthetic_cloud
from tomosar2height_b200 import _lib
from tomosar2height_b200._lib import ptr
from tomosar2height_b200.topology import Topology
tiles, n_per = 4, 262144
cloud = synthetic_cloud(tiles, n_per, seed=1).cuda()
topo = Topology(cloud, 256)
n = tiles * n_per
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def timeit(fn, reps=10):
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
    return sorted(ts)[len(ts)//2]
for C, R in [(32, 256), (64, 256), (128, 128)]:
    lvl = topo.level(R)
    g = torch.randn(n, C, device="cuda")
    out = torch.empty(tiles, R, R, C, device="cuda")
    xyz = lvl.xyz_sorted
    nb = int(_lib.load().t2h_bilinear_sample_bwd_workspace_bytes(R, C, n, lvl.n_seg, lvl.morton))
    ws = torch.empty(nb, dtype=torch.uint8, device="cuda")
    def tiled():
        _lib.call("t2h_bilinear_sample_bwd", ptr(g), n, R, C, ptr(xyz), 4, None, ptr(lvl.keys), ptr(lvl.cell_start), lvl.n_seg, lvl.shift, lvl.morton, ptr(ws), nb, ptr(out))
    def gather():
        _lib.call("t2h_bilinear_sample_bwd", ptr(g), n, R, C, ptr(xyz), 4, None, None, ptr(lvl.cell_start), lvl.n_seg, lvl.shift, lvl.morton, None, 0, ptr(out))
    a = timeit(tiled); ref = out.clone(); b = timeit(gather)
    print(C, R, "wtile ms", round(a, 4), "gather ms", round(b, 4), "max diff", float((out - ref).abs().max()), "bytes", (4*n*C + 8*n + 4*lvl.n_seg*C)/1e6)
