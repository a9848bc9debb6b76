"""Developer check (run on a B200): PoseNet engine vs the CPU oracle + step timing."""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rohm_b200 import synthetic, diffusion
from rohm_b200.posenet import PoseNet
from oracle import posenet_oracle

dev = torch.device('cuda:0')
ds = synthetic.make_dataset('pose')
m = PoseNet(dataset=ds, body_feat_dim=294, latent_dim=512, ff_size=1024, num_layers=8, num_heads=4, device=dev, traj_feat_dim=22)
sd = synthetic.synth_state_dict(m, 1)
m.load_state_dict(sd); m.to(dev).eval()

for (B, T) in [(2, 16), (3, 143), (2, 144)]:
    g = torch.Generator().manual_seed(5)
    x = torch.randn(B, 294, 1, T, generator=g)
    cond = synthetic.posenet_batch(B, T, 7)['cond']
    ts = torch.randint(0, 1000, (B,), generator=g)
    ref = posenet_oracle.posenet_forward(sd, x, cond, ts)
    ref64 = posenet_oracle.posenet_forward(sd, x.double(), cond.double(), ts)
    for prec, name in ((3, 'tf32x3'), (2, 'f16x2'), (1, 'tf32')):
        m.precision = prec
        out = m({'x_t': x.to(dev), 'cond': cond.to(dev)}, ts.to(dev)).cpu()
        print(f"B{B} T{T} {name}: max|gpu-oracle32| {float((out-ref).abs().max()):.3e}  max|gpu-oracle64| {float((out-ref64).abs().max()):.3e}  "
              f"max|oracle32-oracle64| {float((ref-ref64).abs().max()):.3e}  max|ref| {float(ref.abs().max()):.2f}", flush=True)

# timing of the forward at the bench shape
for prec, name in ((3, 'tf32x3'), (2, 'f16x2'), (1, 'tf32')):
    m.precision = prec
    B, T = 32, 144
    x = torch.randn(B, 294, 1, T, device=dev)
    cond = synthetic.posenet_batch(B, T, 7, device=dev)['cond']
    ts = torch.full((B,), 500, device=dev, dtype=torch.long)
    batch = {'x_t': x, 'cond': cond}
    for _ in range(3): m(batch, ts)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): m(batch, ts)
    e1.record(); torch.cuda.synchronize()
    print(f"{name}: forward B{B} T{T}: {e0.elapsed_time(e1)/20:.3f} ms, launches {m._engine.launches_per_forward}", flush=True)

# a short sampling loop through the public API
args = argparse.Namespace(noise_schedule='cosine', sigma_small=True)
m.precision = None
d = diffusion.create_gaussian_diffusion(args, diffusion, diffusion.SpacedDiffusionPoseNet, 1000, '', dev)
B, T = 32, 144
batch = {'cond': synthetic.posenet_batch(B, T, 7, device=dev)['cond']}
torch.manual_seed(0)
torch.cuda.synchronize(); t0 = time.time()
_, out = d.eval_losses(model=m, batch=batch, shape=[B, 294, 1, T], progress=False, clip_denoised=False, cond_fn_with_grad=False, compute_loss=False)
torch.cuda.synchronize(); dt = time.time() - t0
print(f"1000-step loop B{B}: {dt:.2f} s -> {B/dt:.2f} clips/s; finite={bool(torch.isfinite(out).all())} absmax={float(out.abs().max()):.2f}")
