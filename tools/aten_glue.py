"""Which ATen kernels (adds, copies, ReLU backward, clamps, pooling ...) remain in one eager training micro-step, by
operator and input shapes (torch.profiler, CUDA time).   python tools/aten_glue.py [micro_batch_tiles]"""
import sys, collections, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests/golden')
import oracle
import tomosar2height_b200 as t2h
from cases import synthetic_cloud, synthetic_targets
from torch.profiler import profile, ProfilerActivity
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
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True) as prof:
    step()
    torch.cuda.synchronize()
rows = collections.defaultdict(lambda: [0, 0.0])
for e in prof.key_averages(group_by_input_shape=True):
    if e.key.startswith("aten::") and e.self_device_time_total > 0:
        k = (e.key, str(e.input_shapes)[:110])
        rows[k][0] += e.count; rows[k][1] += e.self_device_time_total / 1e3
tot = sum(v[1] for v in rows.values())
print(f"ATen kernels in one micro-step of {mb} tiles: {tot:.2f} ms device time")
for (name, shp), (n, ms) in sorted(rows.items(), key=lambda kv: -kv[1][1])[:45]:
    print(f"{ms:8.3f} ms  n={n:4d}  {name:34s} {shp}")
