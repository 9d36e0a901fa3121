"""Per-shape time table of the GEMM / conv kernels over one eager training micro-step."""
import sys, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests/golden')
import oracle
import tomosar2height_b200 as t2h
from tomosar2height_b200.profiling import KernelTimer, shape_table
from cases import synthetic_cloud, synthetic_targets
mb, N = int(sys.argv[1]) if len(sys.argv) > 1 else 4, 262144
cfg = t2h.berlin_config()
params = oracle.synth_state_dict(oracle.reference_param_shapes(cfg), seed=0)
model = t2h.TomoSAR2Height(cfg); model.load_state_dict(params); model = model.cuda().train()
cloud = synthetic_cloud(mb, N, 1).cuda(); dsm, _ = synthetic_targets(mb, 512, 1); dsm = dsm.cuda()
def step():
    pa, _ = model(input_cloud=cloud)
    ((pa.squeeze(-1) - dsm).abs().mean(dim=(1, 2)).sum()).backward()
for _ in range(2): step()
torch.cuda.synchronize()
with KernelTimer(n_rows=mb * N) as kt:
    step()
tab = shape_table(kt)
tot = sum(r[3] for r in tab)
print(f"GEMM/conv kernel time {tot:.1f} ms per micro-batch of {mb} tiles")
for name, shp, n, ms, tf in tab[:40]:
    print(f"{ms:8.3f} ms {ms/tot*100:5.1f}%  n={n:3d} {tf:7.1f} TF/s  {name[4:]:16s} rows={shp[0]:8d} K={shp[1]:5d} N={shp[2]:5d} {shp[3] if len(shp) > 3 else ''}")
