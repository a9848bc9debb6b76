"""Short profiling target for ncu: a few forwards of BASELINE configs[2] (TrajNet + TrajControl, 64 clips x 144 frames)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rohm_b200 import synthetic
from rohm_b200.trajnet import TrajNet

dev = torch.device('cuda:0')
ds = synthetic.make_dataset('traj')
m = TrajNet(time_dim=32, mid_dim=512, cond_dim=13, traj_feat_dim=13, trajcontrol=True, device=dev, dataset=ds, repr_abs_only=True)
m.load_state_dict(synthetic.synth_state_dict(m, 2)); m.to(dev).eval()
B, T = 64, 144
batch = {k: v.to(dev) for k, v in synthetic.trajnet_batch(B, T, 5, control=True).items()}
batch['x_t'] = torch.randn(B, T, 13, device=dev)
ts = torch.full((B,), 500, device=dev, dtype=torch.long)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    out = m(batch, ts)
torch.cuda.synchronize()
print("done", bool(torch.isfinite(out).all()))
