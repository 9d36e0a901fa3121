"""Which operands still need a standalone t2h_absmax pass in one training micro-step (shape, count, caller)."""
import sys, collections, traceback, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests/golden')
import oracle
import tomosar2height_b200 as t2h
from tomosar2height_b200 import _lib, linear as L
from cases import synthetic_cloud, synthetic_targets
mb, N = 4, 262144
cfg = t2h.berlin_config()
params = oracle.synth_state_dict(oracle.reference_param_shapes(cfg), seed=0)
model = t2h.TomoSAR2Height(cfg); model.load_state_dict(params); model = model.cuda().train()
cloud = synthetic_cloud(mb, N, 1).cuda(); dsm, _ = synthetic_targets(mb, 512, 1); dsm = dsm.cuda()
def step():
    pa, _ = model(input_cloud=cloud)
    ((pa.squeeze(-1) - dsm).abs().mean(dim=(1, 2)).sum()).backward()
step(); torch.cuda.synchronize()
seen = collections.Counter()
orig = _lib.call
def traced(name, *a):
    if name == "t2h_absmax":
        frames = [f.name for f in traceback.extract_stack()[:-1] if 'tomosar2height_b200' in f.filename]
        seen[(a[6], a[2], a[5], '>'.join(frames[-4:]))] += 1
    return orig(name, *a)
_lib.call = traced; L._lib.call = traced
step(); torch.cuda.synchronize()
tot = 0
for (rows, k1, k2, who), n in sorted(seen.items(), key=lambda kv: -kv[0][0] * (kv[0][1] + kv[0][2]) * kv[1]):
    gb = 4 * rows * (k1 + k2) * n / 1e9
    tot += gb
    print(f"{n:3d} x rows={rows:8d} k={k1}+{k2}  {gb:6.2f} GB  {who}")
print(f"total {tot:.1f} GB per micro-batch (~{tot / 6.2:.2f} ms)")
