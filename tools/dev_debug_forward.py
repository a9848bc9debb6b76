"""Developer aid: one PoseNet forward with synchronous launches so that a faulting kernel is reported at its launch site.
ROHM_B200_GRAPH=0 CUDA_LAUNCH_BLOCKING=1 python tools/dev_debug_forward.py [B T]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rohm_b200 import synthetic
from rohm_b200.posenet import PoseNet
from oracle import posenet_oracle

B, T = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (2, 144)
dev = torch.device('cuda:0')
ds = synthetic.make_dataset('pose')
m = PoseNet(dataset=ds, body_feat_dim=294, latent_dim=512, ff_size=1024, num_layers=8, num_heads=4, device=dev, traj_feat_dim=22)
sd = synthetic.synth_state_dict(m, 1)
m.load_state_dict(sd); m.to(dev).eval()
g = torch.Generator().manual_seed(5)
x = torch.randn(B, 294, 1, T, generator=g)
cond = synthetic.posenet_batch(B, T, 7)['cond']
ts = torch.randint(0, 1000, (B,), generator=g)
try:
    out = m({'x_t': x.to(dev), 'cond': cond.to(dev)}, ts.to(dev))
    torch.cuda.synchronize()
    ref = posenet_oracle.posenet_forward(sd, x, cond, ts)
    print("max err", float((out.cpu() - ref).abs().max()), "finite", bool(torch.isfinite(out).all()))
except Exception as e:
    print("FAILED:", str(e)[:600])
