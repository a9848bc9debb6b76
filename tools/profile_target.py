"""Short profiling target for ncu: a few steps of the BASELINE configs[1] PoseNet chain (32 clips x 144 frames)."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rohm_b200 import diffusion, synthetic
from rohm_b200.posenet import PoseNet

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 4
dev = torch.device('cuda:0')
ds = synthetic.make_dataset('pose')
m = PoseNet(dataset=ds, body_feat_dim=294, latent_dim=512, ff_size=1024, num_layers=8, num_heads=4, device=dev, traj_feat_dim=22)
m.load_state_dict(synthetic.synth_state_dict(m, 1)); m.to(dev).eval()
args = argparse.Namespace(noise_schedule='cosine', sigma_small=True)
d = diffusion.create_gaussian_diffusion(args, diffusion, diffusion.SpacedDiffusionPoseNet, 1000, f'ddim{steps}' if 1000 % steps == 0 else '', dev)
B, T = 32, 144
batch = {'cond': synthetic.posenet_batch(B, T, 7, device=dev)['cond']}
torch.manual_seed(0)
out = d.p_sample_loop(m, batch, [B, 294, 1, T], clip_denoised=False)
torch.cuda.synchronize()
print("done", float(out.abs().max()))
