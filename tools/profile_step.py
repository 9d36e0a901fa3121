"""torch.profiler kernel table for one training step (device time per kernel), for triage only."""
import sys, os, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests/golden')
import oracle
import tomosar2height_b200 as t2h
from cases import synthetic_cloud, synthetic_targets
from torch.profiler import profile, ProfilerActivity
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.benchmark = True
tiles, mb, N = int(sys.argv[1]) if len(sys.argv) > 1 else 4, 2, 262144
cfg = t2h.berlin_config()
params = oracle.synth_state_dict(oracle.reference_param_shapes(cfg), seed=0)
model = t2h.TomoSAR2Height(cfg); model.load_state_dict(params); model = model.cuda().train()
opt = torch.optim.AdamW(model.parameters(), lr=1e-4)
cloud = synthetic_cloud(tiles, N, 1).cuda(); dsm, _ = synthetic_targets(tiles, 512, 1); dsm = dsm.cuda()
def step():
    opt.zero_grad(set_to_none=False)
    for i in range(0, tiles, mb):
        pa, _ = model(input_cloud=cloud[i:i + mb])
        ((pa.squeeze(-1) - dsm[i:i + mb]).abs().mean(dim=(1, 2)).sum()).backward()
    opt.step()
for _ in range(2): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step(); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=70))
