"""Small invocations of the second-session kernels for compute-sanitizer (memcheck / racecheck): a TrajNet + TrajControl
forward at 2 clips (split-K convolutions with up to 8 K ranges, gn_mish_split_kernel, sum_split_kernel) and a 130-frame SMPL-X
LBS call with vertices (fused launch, TMA vertex stores over a ragged row and column tile).
compute-sanitizer --tool racecheck python tools/sanitizer_target.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rohm_b200 import synthetic
from rohm_b200.body_model import BodyModel
from rohm_b200.trajnet import TrajNet

dev = torch.device('cuda:0')
ds = synthetic.make_dataset('traj')
m = TrajNet(time_dim=32, mid_dim=512, cond_dim=13, traj_feat_dim=13, trajcontrol=True, device=dev, dataset=ds, repr_abs_only=True)
m.load_state_dict(synthetic.synth_state_dict(m, 2)); m.to(dev).eval()
B, T = 2, 144
batch = {k: v.to(dev) for k, v in synthetic.trajnet_batch(B, T, 5, control=True).items()}
batch['x_t'] = torch.randn(B, T, 13, device=dev)
out = m(batch, torch.full((B,), 500, device=dev, dtype=torch.long))
torch.cuda.synchronize()
print("trajnet forward finite:", bool(torch.isfinite(out).all()))

bm = BodyModel.create('', device=dev, seed=0)
N = 130
g = torch.Generator().manual_seed(1)
o = bm(transl=torch.randn(N, 3, generator=g).to(dev), global_orient=(0.3 * torch.randn(N, 3, generator=g)).to(dev),
       body_pose=(0.3 * torch.randn(N, 63, generator=g)).to(dev), betas=torch.randn(N, 10, generator=g).to(dev))
torch.cuda.synchronize()
print("lbs vertices finite:", bool(torch.isfinite(o.vertices).all()), tuple(o.vertices.shape), o.vertices.stride())
