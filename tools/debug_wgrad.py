import torch, sys
sys.path.insert(0, '.')
from tomosar2height_b200 import _lib
from tomosar2height_b200.linear import _launch_wgrad
torch.manual_seed(0)
for rows, n, k in [(32, 32, 32), (64, 32, 32), (32, 128, 64)]:
    g = torch.zeros(rows, n, device='cuda'); x = torch.zeros(rows, k, device='cuda')
    g[3, 5] = 1.0; g[9, 1] = 2.0
    x[3] = torch.arange(1, k + 1, device='cuda').float(); x[9] = 100 + torch.arange(k, device='cuda').float()
    dw = torch.full((n, k), -7.0, device='cuda')
    _launch_wgrad(g, x, False, dw)
    torch.cuda.synchronize()
    ref = g.t() @ x
    print(rows, n, k, 'max err', (dw - ref).abs().max().item(), 'nonzero got', int((dw != 0).sum()), 'ref', int((ref != 0).sum()))
    nz = (dw != 0).nonzero()
    print(' got rows', sorted(set(nz[:, 0].tolist()))[:10], 'row5', dw[5, :8].tolist(), 'row1', dw[1, :8].tolist())
