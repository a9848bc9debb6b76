"""2-GPU check under torchrun (NCCL over NVLink), world size 2: clip-sharded PoseNet sampling with guidance.
 (1) unguided: sample_sharded (parity noise) is bit-identical to the single-GPU run of the whole batch;
 (2) guided gradient: with parallel.global_guidance the gathered skating gradient equals the unsharded one (4-float
     all-reduce per guided step), the default per-shard contract does not;
 (3) a guided sharded 12-step tail runs through sample_sharded in both modes.
torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/n2_sharded_check.py"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from rohm_b200 import diffusion, parallel, synthetic
from rohm_b200.body_model import kernels_for
from rohm_b200.posenet import PoseNet

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", rank)))
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
ds = synthetic.make_dataset('pose', seed=3, realistic_std=True)
m = PoseNet(dataset=ds, body_feat_dim=294, latent_dim=512, ff_size=1024, num_layers=8, num_heads=4, device=dev, traj_feat_dim=22)
m.load_state_dict({k: v.cpu() for k, v in synthetic.synth_state_dict(m, 1).items()})
m.to(dev).eval()
a = argparse.Namespace(noise_schedule='cosine', sigma_small=True)
B, T = 6, 143
cond = synthetic.plausible_motion(B, T, 21, ds).to(dev)
shape = [B, 294, 1, T]

# (1) unguided, 20 respaced steps
d = diffusion.create_gaussian_diffusion(a, diffusion, diffusion.SpacedDiffusionPoseNet, 1000, 'ddim20', dev)
torch.manual_seed(5)
out_sh = parallel.sample_sharded(d, m, {'cond': cond}, shape, progress=False, clip_denoised=False, cond_fn_with_grad=False)
torch.manual_seed(5)
parallel.ShardedNoise(B, 0, 1).install(d)
_, out_full = d.eval_losses(model=m, batch={'cond': cond}, shape=shape, compute_loss=False, progress=False, clip_denoised=False,
                            cond_fn_with_grad=False)
d._randn, d._randn_like = torch.randn, torch.randn_like
same = torch.equal(out_sh, out_full)

# (2) guidance gradient, global vs per-shard normalisers
mean, std = torch.from_numpy(ds.Mean).to(dev), torch.from_numpy(ds.Std).to(dev)
k = kernels_for(m.smplx_model, dev, B * T, with_vertices=False)
full = k.skating_guidance(cond, mean, std)
lo, hi = parallel.shard_bounds(B, rank, world)
mine = cond[lo:hi].contiguous()
parallel.global_guidance(m)
g_glob = m.guide_skating_with_smpl({}, {'pred_xstart': mine}, None, compute_grad='x_0')
parallel.global_guidance(m, enable=False)
g_shard = m.guide_skating_with_smpl({}, {'pred_xstart': mine}, None, compute_grad='x_0')
got = parallel.gather_clips(g_glob, B)
got_shard = parallel.gather_clips(g_shard, B)
scale = float(full.abs().max())
e_glob, e_shard = float((got - full).abs().max()) / scale, float((got_shard - full).abs().max()) / scale

# (3) a guided tail through sample_sharded, both modes (runs; the chain is chaotic)
dg = diffusion.create_gaussian_diffusion(a, diffusion, diffusion.SpacedDiffusionPoseNet, 1000, "12" + ",0" * 19, dev)
finite = []
for glob in (False, True):
    parallel.global_guidance(m, enable=glob)
    torch.manual_seed(7)
    o = parallel.sample_sharded(dg, m, {'cond': cond}, shape, progress=False, clip_denoised=False, cond_fn_with_grad=True,
                                grad_type='amass')
    finite.append(bool(torch.isfinite(o).all()))
parallel.global_guidance(m, enable=False)
if rank == 0:
    print(f"n2 check: unguided sharded == single-GPU bit-exact: {same}; guidance gradient vs unsharded, relative: "
          f"exact-global mode {e_glob:.2e}, per-shard mode {e_shard:.2e}; guided sharded tails finite: {finite}")
    assert same and e_glob < 1e-5 and e_shard > 1e-3 and all(finite)
dist.barrier()
dist.destroy_process_group()
