import sys, os, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests/golden')
import oracle
from cases import CASES, make_cfg, synthetic_cloud, synthetic_targets
import tomosar2height_b200 as t2h
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
name = sys.argv[1] if len(sys.argv) > 1 else "munich_small"
spec = CASES[name]; cfg = make_cfg(**spec["cfg"])
params = oracle.synth_state_dict(oracle.reference_param_shapes(cfg), seed=spec["seed"])
model = t2h.TomoSAR2Height(cfg); model.load_state_dict(params); model = model.cuda()
B, N = spec["B"], 2 * spec["N"]
size = cfg.model.decoder_pixel_kwargs.output_size
cloud = synthetic_cloud(B, N, seed=spec["seed"] + 50)
dsm, image = synthetic_targets(B, size, spec["seed"] + 50, with_image=cfg.use_image)
P64 = {k: v.double().requires_grad_(True) for k, v in params.items()}
pa64, pb64 = oracle.oracle_forward(P64, cfg, cloud.double(), None if image is None else image.double(), aten=False)
g = torch.Generator().manual_seed(7)
wa = torch.randn(B, size, size, 1, generator=g, dtype=torch.float64)
wb = torch.randn(B, size, size, 1, generator=g, dtype=torch.float64)
((pa64 * wa).mean() + (0.0 if pb64 is None else (pb64 * wb).mean())).backward()
pa, pb = model(input_cloud=cloud.cuda(), input_image=None if image is None else image.cuda())
((pa * wa.float().cuda()).mean() + (0.0 if pb is None else (pb * wb.float().cuda()).mean())).backward()
# the same network in fp32 on the CPU (oracle), to separate kernel error from fp32 conditioning
P32 = {k: v.clone().requires_grad_(True) for k, v in params.items()}
pa32, pb32 = oracle.oracle_forward(P32, cfg, cloud, image, aten=False)
((pa32 * wa.float()).mean() + (0.0 if pb32 is None else (pb32 * wb.float()).mean())).backward()
rows32 = []
for pname in P32:
    if P64[pname].grad is None or P32[pname].grad is None: continue
    denom = max(P64[pname].grad.abs().max().item(), 1e-12)
    rows32.append(((P32[pname].grad.double() - P64[pname].grad).abs().max().item() / denom, pname))
rows32.sort(reverse=True)
print("CPU fp32 oracle vs fp64: worst", rows32[:3])
print("heights rel err", ((pa.detach().cpu().double() - pa64.detach()).abs().max() / pa64.abs().max()).item())
rows = []
for pname, p in model.named_parameters():
    g64 = P64[pname].grad
    if g64 is None or p.grad is None: continue
    denom = max(g64.abs().max().item(), 1e-12)
    rows.append(((p.grad.cpu().double() - g64).abs().max().item() / denom, pname, denom))
rows.sort(reverse=True)
for r in rows[:12]: print(f"{r[0]:.3e}  {r[1]}  max|g|={r[2]:.3e}")
import statistics
e_gpu = sorted(r[0] for r in rows); e_cpu = sorted(r[0] for r in rows32)
print("GPU  median %.2e p90 %.2e max %.2e" % (statistics.median(e_gpu), e_gpu[int(.9*len(e_gpu))], e_gpu[-1]))
print("CPU32 median %.2e p90 %.2e max %.2e" % (statistics.median(e_cpu), e_cpu[int(.9*len(e_cpu))], e_cpu[-1]))
