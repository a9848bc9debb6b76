"""Timing of BASELINE configs[2] (TrajNet + TrajControl, 64 clips x 144 frames): one forward and a 100-step ancestral loop."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rohm_b200 import diffusion, synthetic
from rohm_b200.trajnet import TrajNet

dev = torch.device('cuda:0')
for control in (False, True):
    ds = synthetic.make_dataset('traj')
    m = TrajNet(time_dim=32, mid_dim=512, cond_dim=13, traj_feat_dim=13, trajcontrol=control, device=dev, dataset=ds,
                repr_abs_only=True)
    m.load_state_dict(synthetic.synth_state_dict(m, 2)); m.to(dev).eval()
    B, T = 64, 144
    batch = {k: v.to(dev) for k, v in synthetic.trajnet_batch(B, T, 5, control=control).items()}
    batch['x_t'] = torch.randn(B, T, 13, device=dev)
    ts = torch.full((B,), 500, device=dev, dtype=torch.long)
    for _ in range(3): m(batch, ts)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): m(batch, ts)
    e1.record(); torch.cuda.synchronize()
    fwd = e0.elapsed_time(e1) / 20
    args = argparse.Namespace(noise_schedule='cosine', sigma_small=True)
    d = diffusion.create_gaussian_diffusion(args, diffusion, diffusion.SpacedDiffusionTrajNet, 100, '', dev)
    d.p_sample_loop(m, batch, [B, T, 13], clip_denoised=False); torch.cuda.synchronize()
    e0.record()
    out = d.p_sample_loop(m, batch, [B, T, 13], clip_denoised=False)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    gflop = 1.144 if control else 0.723
    print(f"TrajNet control={control}: forward B{B} T{T} {fwd:.3f} ms ({B * gflop / fwd:.1f} TFLOP/s algorithmic), "
          f"100-step loop {ms:.1f} ms -> {B / ms * 1e3:.1f} clips/s at 100 steps, {B / ms * 1e2:.2f} clips/s at 1000 steps; "
          f"finite={bool(torch.isfinite(out).all())}", flush=True)
