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
oracle.oracle_loss(pa64, pb64, dsm, cfg.use_footprint).backward()
pa, pb = model(input_cloud=cloud.cuda(), input_image=None if image is None else image.cuda())
loss = torch.nn.functional.l1_loss(pa.squeeze(), dsm.cuda().squeeze())
if cfg.use_footprint:
    loss = loss + 10.0 * torch.nn.functional.binary_cross_entropy_with_logits(pb.squeeze(), (dsm.cuda().squeeze() > 0.0001).float())
loss.backward()
print("heights rel err", ((pa.detach().cpu().double() - pa64.detach()).abs().max() / pa64.abs().max()).item())
rows = []
for pname, p in model.named_parameters():
    g64 = P64[pname].grad
    if g64 is None or p.grad is None: continue
    denom = max(g64.abs().max().item(), 1e-12)
    rows.append(((p.grad.cpu().double() - g64).abs().max().item() / denom, pname, denom))
rows.sort(reverse=True)
for r in rows[:12]: print(f"{r[0]:.3e}  {r[1]}  max|g|={r[2]:.3e}")
