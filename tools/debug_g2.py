import sys, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests/golden')
from cases import synthetic_cloud
import tomosar2height_b200.functional as T
from tomosar2height_b200.topology import Topology
from torch.profiler import profile, ProfilerActivity
for clustered in (True, False):
    cloud = synthetic_cloud(4, 262144, seed=1, clustered=clustered).cuda()
    topo = Topology(cloud, 256)
    for R, C in ((256, 32), (128, 128), (32, 512)):
        lvl = topo.level(R)
        plane = torch.randn(4, R, R, C, device='cuda', requires_grad=True)
        rows = torch.randn(4 * 262144, C, device='cuda')
        sm = T.bilinear_sample(plane, lvl)
        torch.autograd.grad(sm, plane, rows, retain_graph=True); torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            torch.autograd.grad(sm, plane, rows, retain_graph=True); torch.cuda.synchronize()
        for e in prof.key_averages():
            if e.device_time_total > 20: print(clustered, R, C, e.key[:60], round(e.device_time_total / 1e3, 3), 'ms')
