"""Short profiling target for ncu: full LBS (N3-sized, 32 clips x 143 frames) and the skating-guidance kernels."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rohm_b200 import synthetic
from rohm_b200.body_model import BodyModel, kernels_for
dev = torch.device('cuda:0')
bm = BodyModel.create('', device=dev, seed=0)
N = 32 * 143
g = torch.Generator().manual_seed(0)
go, bp = 0.3 * torch.randn(N, 3, generator=g), 0.3 * torch.randn(N, 63, generator=g)
be, tr = torch.randn(N, 10, generator=g), torch.randn(N, 3, generator=g)
for _ in range(3):
    out = bm(transl=tr.to(dev), global_orient=go.to(dev), body_pose=bp.to(dev), betas=be.to(dev))
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    out = bm(transl=tr.to(dev), global_orient=go.to(dev), body_pose=bp.to(dev), betas=be.to(dev))
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"full LBS {N} frames: {ms:.3f} ms -> {N/ms*1e3:.0f} frames/s, {N*126280/ms/1e6:.1f} GB/s algorithmic (126280 B/frame)")
ds = synthetic.make_dataset('pose', seed=3, realistic_std=True)
x = synthetic.plausible_motion(32, 143, 1, ds).to(dev)
k = kernels_for(bm, dev, 32 * 143, with_vertices=False)
mean, std = torch.from_numpy(ds.Mean).to(dev), torch.from_numpy(ds.Std).to(dev)
for _ in range(3):
    gr = k.skating_guidance(x, mean, std)
torch.cuda.synchronize()
e0.record()
for _ in range(20):
    gr = k.skating_guidance(x, mean, std)
e1.record(); torch.cuda.synchronize()
print(f"skating guidance 32x143: {e0.elapsed_time(e1)/20*1e3:.1f} us per call, grad absmax {float(gr.abs().max()):.3e}")
