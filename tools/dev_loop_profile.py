"""Developer check: where does the host time of one sampling step go?"""
import argparse, cProfile, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rohm_b200 import diffusion, synthetic
from rohm_b200.posenet import PoseNet
dev = torch.device('cuda:0')
ds = synthetic.make_dataset('pose')
m = PoseNet(dataset=ds, body_feat_dim=294, latent_dim=512, ff_size=1024, num_layers=8, num_heads=4, device=dev, traj_feat_dim=22)
m.load_state_dict(synthetic.synth_state_dict(m, 1)); m.to(dev).eval()
args = argparse.Namespace(noise_schedule='cosine', sigma_small=True)
d = diffusion.create_gaussian_diffusion(args, diffusion, diffusion.SpacedDiffusionPoseNet, 1000, 'ddim200', dev)
B, T = 32, 144
batch = {'cond': synthetic.posenet_batch(B, T, 7, device=dev)['cond']}
shape = [B, 294, 1, T]
d.p_sample_loop(m, batch, shape, clip_denoised=False); torch.cuda.synchronize()
# host-only time of engine.forward
e = m._engine
x = torch.randn(shape, device=dev); ts = torch.zeros(B, dtype=torch.long, device=dev); out = torch.empty_like(x)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(200): e.forward(x, ts, out)
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"engine.forward: host {1e6*(t1-t0)/200:.1f} us/call, device-complete {1e6*(t2-t0)/200:.1f} us/call")
torch.cuda.synchronize(); t0 = time.perf_counter()
d.p_sample_loop(m, batch, shape, clip_denoised=False)
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"p_sample_loop(200 steps): host {1e3*(t1-t0)/200:.3f} ms/step, total {1e3*(t2-t0)/200:.3f} ms/step")
pr = cProfile.Profile(); pr.enable()
d.p_sample_loop(m, batch, shape, clip_denoised=False)
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats('cumulative').print_stats(22)
