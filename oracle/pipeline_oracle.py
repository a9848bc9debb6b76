"""Oracle (test infrastructure only): the multi-round TrajNet -> PoseNet inference of test_amass_full.py:231-384 on the CPU,
assembled from the other oracle modules (denoisers, samplers, kinematics, glue).  Flag set of the shipped AMASS configs:
input_noise, iter2_cond_noisy_traj, iter2_cond_noisy_pose, repr_abs_only, mask_scheme in {'lower', 'upper', None},
cond_fn_with_grad, grad_type='amass'.  Pinned by tests/golden/pipeline.npz (the unmodified reference run through the same
call sequence, tools/gen_golden.py:gen_pipeline).
"""
import torch

from . import diffusion_oracle as do
from . import glue_oracle as go
from . import kinematics_oracle as ko
from . import posenet_oracle, trajnet_oracle


def posenet_guided_step(tables, tmap, i, x_t, cond, sd_pose, mean_p, std_p, body_model, noise, guided=True, guide_last=50):
    """One p_sample_with_grad step of the PoseNet sampler (gaussian_diffusion_posenet.py:436-480) -> (x_{t-1}, x0)."""
    B = x_t.shape[0]
    with torch.no_grad():
        x0 = posenet_oracle.posenet_forward(sd_pose, x_t, cond, torch.full((B,), tmap[i], dtype=torch.long))
    g = None
    if guided and i <= guide_last:
        gr = ko.guide_skating(x0, mean_p, std_p, body_model)
        g = [(3e6, gr)] if gr.dim() != 0 else None
    return do.p_sample_step(tables, i, x_t, x0, noise, g), x0


def run_rounds(sd_pose, sd_traj, sd_ctrl, ds_pose, ds_traj, body_model, pose, traj, pose_steps, traj_steps, rounds,
               noise_pose, noise_traj, mask_scheme='lower', guided=True, guide_last=50, pose_respacing='', teacher=None,
               teacher_steps=()):
    """pose / traj: the dataloader batches (CPU tensors, see rohm_b200.synthetic.pipeline_batches); noise_*: objects with
    randn(*shape) / randn_like(x) replaying the reference's global-generator draws (one stream per diffusion module).
    Returns a list with one dict per round: val_traj, traj_full, cond, val_pose (free-running).

    teacher: optional mapping with the reference's per-round results (tests/golden/pipeline.npz: r{k}_val_pose, r{k}_xt{i}).
    When given, round k+1 is conditioned on the teacher's round-k PoseNet output (the guided PoseNet chain is chaotic, so
    stages are compared one at a time) and each dict also carries 'tf': {i: x_{i-1} computed from the teacher's x_i}."""
    tp, mp_ = do.create_diffusion('cosine', pose_steps, pose_respacing)
    tt, mt_ = do.create_diffusion('cosine', traj_steps, '')
    mean_p, std_p = torch.from_numpy(ds_pose.Mean), torch.from_numpy(ds_pose.Std)
    B, T = traj['cond'].shape[0], traj['cond'].shape[1]
    out = []
    val_pose = None
    noisy_p = pose['motion_repr_noisy'][:, 0:-1]
    n_pose = len(tp['betas'])
    for it in range(rounds):
        x_T = noise_traj.randn(B, T, 13)
        if it == 0:
            fn = lambda x, t: trajnet_oracle.trajnet_forward(sd_traj, x, traj['cond'], torch.full((B,), t, dtype=torch.long))
        else:
            cc = go.pose_to_control_cond(val_pose, T, 272)
            fn = lambda x, t: trajnet_oracle.trajnet_forward(sd_ctrl, x, traj['cond'], torch.full((B,), t, dtype=torch.long),
                                                             control_cond=cc)
        with torch.no_grad():
            val_traj, _ = do.p_sample_loop(tt, mt_, fn, x_T, lambda i: noise_traj.randn_like(x_T))
        _, traj_full = go.traj_to_full_repr(val_traj, traj['motion_repr_clean'], ds_traj.Mean, ds_traj.Std, ds_pose.Mean,
                                            ds_pose.Std, body_model)
        cond = go.build_pose_cond(noisy_p, traj_full, mask_scheme, apply_mask=mask_scheme is not None)
        x = noise_pose.randn(B, 294, 1, T - 1)
        noises = {i: noise_pose.randn_like(x) for i in range(n_pose - 1, -1, -1)}  # the reference draws one per step, in order
        for i in range(n_pose - 1, -1, -1):
            x, _ = posenet_guided_step(tp, mp_, i, x, cond, sd_pose, mean_p, std_p, body_model, noises[i], guided, guide_last)
        res = {'val_traj': val_traj, 'traj_full': traj_full, 'cond': cond, 'val_pose': x.detach()}
        val_pose = res['val_pose']
        if teacher is not None:
            res['tf'] = {}
            for i in teacher_steps:
                xt = torch.from_numpy(teacher[f"r{it}_xt{i}"])
                res['tf'][i] = posenet_guided_step(tp, mp_, i, xt, cond, sd_pose, mean_p, std_p, body_model, noises[i], guided,
                                                   guide_last)[0].detach()
            val_pose = torch.from_numpy(teacher[f"r{it}_val_pose"])
        out.append(res)
    return out
