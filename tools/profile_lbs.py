"""Short profiling target for ncu: SMPL-X full LBS from the motion representation, 32 clips x 143 frames (row L2 / N3)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rohm_b200 import glue, synthetic
from rohm_b200.body_model import BodyModel, kernels_for
dev = torch.device('cuda:0')
bm = BodyModel.create('', device=dev, seed=0)
ds = synthetic.make_dataset('pose', seed=3, realistic_std=True)
B, T = 32, 143
x = synthetic.plausible_motion(B, T, 7, ds).to(dev)
mean, std = glue.stats_on(ds, dev)
k = kernels_for(bm, dev, B * T, with_vertices=True)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
for _ in range(reps):
    j, v = k.from_repr(x, mean, std, want_vertices=True)
torch.cuda.synchronize()
print("done", float(v.abs().max()))
