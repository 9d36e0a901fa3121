"""One launch of every segment / sampling operator per plane level, bracketed by cudaProfilerStart/Stop:

ncu --profile-from-start off --set full --clock-control none --import-source on -o gpurun_out/prof_ops python tools/profile_ops.py
"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from cases import synthetic_cloud
import tomosar2height_b200.functional as T
from tomosar2height_b200.topology import Topology

levels = [(32, 256), (128, 128), (512, 32)] if "--all" not in sys.argv else [(32, 256), (64, 256), (128, 128), (256, 64), (512, 32)]
tiles, n_per = 4, 262144
cloud = synthetic_cloud(tiles, n_per, seed=1).cuda()
topo = Topology(cloud, 256)
n = tiles * n_per
for rep in range(2):
    if rep == 1:
        torch.cuda.synchronize(); torch.cuda.profiler.start()
    for C, R in levels:
        lvl = topo.level(R)
        rows = torch.randn(n, C, device="cuda")
        plane = torch.randn(tiles, R, R, C, device="cuda").requires_grad_(True)
        if C <= 128:
            r2 = rows.clone().requires_grad_(True)
            T.seg_max_pool(r2, lvl).backward(rows)
        r3 = rows.clone().requires_grad_(True)
        T.seg_mean(r3, lvl).sum().backward()
        T.bilinear_sample(plane, lvl).backward(rows)
        del rows, plane
torch.cuda.synchronize(); torch.cuda.profiler.stop()
