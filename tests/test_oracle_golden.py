"""CPU: pins the oracle (oracle/*.py) to golden vectors produced by the unmodified reference (tools/gen_golden.py)
and to the known-answer values of SURVEY.md Appendix F."""
import numpy as np
import pytest
import torch

from helpers import NoiseTape, TOL, golden, posenet_state_dict
from oracle import diffusion_oracle as do
from oracle import kinematics_oracle as ko
from oracle import posenet_oracle, trajnet_oracle
from rohm_b200 import synthetic

TABLES = ("betas", "alphas_cumprod", "alphas_cumprod_prev", "alphas_cumprod_next", "sqrt_alphas_cumprod",
          "sqrt_one_minus_alphas_cumprod", "log_one_minus_alphas_cumprod", "sqrt_recip_alphas_cumprod",
          "sqrt_recipm1_alphas_cumprod", "posterior_variance", "posterior_log_variance_clipped",
          "posterior_mean_coef1", "posterior_mean_coef2")


def _resp(meta):
    sched, steps, resp = [str(x) for x in meta]
    return sched, int(steps), resp


def test_schedule_tables_bit_exact():
    g = golden("schedules.npz")
    for c in range(int(g["n_cases"])):
        sched, steps, resp = _resp(g[f"c{c}_meta"])
        tables, tmap = do.create_diffusion(sched, steps, resp)
        assert tmap == g[f"c{c}_timestep_map"].tolist()
        for t in TABLES:
            assert np.array_equal(tables[t], g[f"c{c}_{t}"]), (sched, steps, resp, t)


def test_space_timesteps_known_answers():
    g = golden("schedules.npz")
    assert sorted(do.space_timesteps(300, [10, 15, 20])) == g["space_300_10_15_20"].tolist()
    assert sorted(do.space_timesteps(1000, "ddim100")) == g["space_1000_ddim100"].tolist()
    assert sorted(do.space_timesteps(1000, "100")) == g["space_1000_100"].tolist()
    assert sorted(do.space_timesteps(1000, "7,13,29")) == g["space_1000_7_13_29"].tolist()
    assert int(g["ddim300_raises"]) == 1
    with pytest.raises(ValueError):
        do.space_timesteps(1000, "ddim300")
    # reference docstring example (respace.py:16-18)
    s = sorted(do.space_timesteps(300, [10, 15, 20]))
    assert len(s) == 45 and s[:11] == [0, 11, 22, 33, 44, 55, 66, 77, 88, 99, 100]


def test_appendix_f_known_answers():
    t, tmap = do.create_diffusion("cosine", 1000, "")
    assert tmap == list(range(1000))
    assert t["betas"][0] == 4.128422482196914e-05 and t["betas"][1] == 4.614175273665033e-05 and t["betas"][-1] == 0.999
    assert t["alphas_cumprod"][500] == 0.49228517244880304
    assert t["posterior_mean_coef1"][0] == 1.0 and t["posterior_mean_coef2"][0] == 0.0
    assert t["posterior_mean_coef1"][-1] == 0.0015568917154901703
    assert t["posterior_log_variance_clipped"][0] == t["posterior_log_variance_clipped"][1] == -10.734082532465003
    assert t["posterior_variance"][50] == 0.00027514289686454517
    t, tmap = do.create_diffusion("cosine", 1000, "ddim100")
    assert tmap == list(range(0, 1000, 10))
    assert t["betas"][-1] == 0.7755724061584093 and t["posterior_mean_coef2"][-1] == 0.47341577964733306
    t, tmap = do.create_diffusion("cosine", 1000, "100")
    assert tmap[-3:] == [979, 989, 999] and t["betas"][-1] == 0.9999899991985922


def test_posenet_forward_matches_reference():
    g = golden("posenet_forward.npz")
    sd, _ = posenet_state_dict(int(g["weight_seed"]))
    for c in range(int(g["n_cases"])):
        B, T, s = [int(v) for v in g[f"c{c}_meta"]]
        gen = torch.Generator().manual_seed(s)
        x = torch.randn(B, 294, 1, T, generator=gen)
        cond = synthetic.posenet_batch(B, T, s + 100)['cond']
        ts = torch.randint(0, 1000, (B,), generator=gen)
        assert np.array_equal(ts.numpy(), g[f"c{c}_timesteps"])
        y = posenet_oracle.posenet_forward(sd, x, cond, ts)
        err = float((y - torch.from_numpy(g[f"c{c}_out"])).abs().max())
        assert err < 2e-5, (c, err)


def _trajnet_sd(seed, control):
    from rohm_b200.trajnet import TrajNet
    ds = synthetic.make_dataset('traj')
    m = TrajNet(time_dim=32, mid_dim=512, cond_dim=13, traj_feat_dim=13, trajcontrol=control, device=None, dataset=ds,
                repr_abs_only=True)
    return synthetic.synth_state_dict(m, seed)


def test_trajnet_forward_matches_reference():
    g = golden("trajnet_forward.npz")
    for c in range(int(g["n_cases"])):
        B, T, s, control = [int(v) for v in g[f"c{c}_meta"]]
        sd = _trajnet_sd(int(g["weight_seed"]), bool(control))
        gen = torch.Generator().manual_seed(s)
        x = torch.randn(B, T, 13, generator=gen)
        batch = synthetic.trajnet_batch(B, T, s + 100, control=bool(control))
        ts = torch.randint(0, 100, (B,), generator=gen)
        y = trajnet_oracle.trajnet_forward(sd, x, batch['cond'], ts, batch.get('control_cond'))
        err = float((y - torch.from_numpy(g[f"c{c}_out"])).abs().max())
        assert err < 2e-5, (c, err)


def test_sampling_config1_trajnet_50_steps():
    """BASELINE config 1: TrajNet vanilla, 1 clip, 50 DDPM steps -- oracle loop vs the reference's eval_losses."""
    g = golden("sampling.npz")
    B, T, bseed, nseed, steps = [int(v) for v in g["traj50_meta"]]
    sd = _trajnet_sd(2, False)
    batch = synthetic.trajnet_batch(B, T, bseed)
    tables, tmap = do.create_diffusion("cosine", steps, "")
    tape = NoiseTape(nseed)
    x_T = tape.randn(B, T, 13)
    model = lambda x, t: trajnet_oracle.trajnet_forward(sd, x, batch['cond'], torch.full((B,), t, dtype=torch.long))
    y, _ = do.p_sample_loop(tables, tmap, model, x_T, lambda i: tape.randn_like(x_T))
    err = float((y - torch.from_numpy(g["traj50_out"])).abs().max())
    assert err < TOL, err


def test_sampling_posenet_respaced():
    g = golden("sampling.npz")
    B, T, bseed, nseed, steps = [int(v) for v in g["pose_ddim20_meta"]]
    sd, _ = posenet_state_dict(1)
    cond = synthetic.posenet_batch(B, T, bseed)['cond']
    tables, tmap = do.create_diffusion("cosine", 1000, "ddim20")
    assert len(tmap) == steps
    tape = NoiseTape(nseed)
    x_T = tape.randn(B, 294, 1, T)
    model = lambda x, t: posenet_oracle.posenet_forward(sd, x, cond, torch.full((B,), t, dtype=torch.long))
    y, _ = do.p_sample_loop(tables, tmap, model, x_T, lambda i: tape.randn_like(x_T))
    err = float((y - torch.from_numpy(g["pose_ddim20_out"])).abs().max())
    assert err < TOL, err


def test_rotation_chain_matches_reference():
    g = golden("kinematics.npz")
    r6 = torch.from_numpy(g["rot6d_in"])
    R = ko.rot6d_to_rotmat(r6)
    assert float((R - torch.from_numpy(g["rotmat_out"])).abs().max()) < 1e-6
    aa = ko.rotmat_to_aa(R)
    assert float((aa - torch.from_numpy(g["aa_out"])).abs().max()) < 1e-5


def _kin_inputs(g):
    B, T, seed, dseed = [int(v) for v in g["kin_meta"]]
    ds = synthetic.make_dataset('pose', seed=dseed, realistic_std=True)
    x = synthetic.plausible_motion(B, T, seed, ds)
    return ds, x


def test_joint_recovery_and_skating_gradient_match_reference():
    g = golden("kinematics.npz")
    ds, x = _kin_inputs(g)
    mean, std = torch.from_numpy(ds.Mean), torch.from_numpy(ds.Std)
    rep = ko.split_repr(x[:, :, 0].permute(0, 2, 1) * std + mean)
    body = synthetic.smplx_like_model(0)
    assert float((ko.joints_from_abs_traj(rep) - torch.from_numpy(g["abs_traj_joints"])).abs().max()) < 1e-6
    assert float((ko.joints_from_smplx(rep, body) - torch.from_numpy(g["smplx_joints"])).abs().max()) < 1e-5
    grad = ko.guide_skating(x, mean, std, body)
    ref = torch.from_numpy(g["skating_grad"])
    assert grad.shape == ref.shape and float(ref.abs().max()) > 1e-3
    assert float((grad - ref).abs().max()) < 1e-5 * max(1.0, float(ref.abs().max()))


def test_sampling_posenet_guided():
    """p_sample_with_grad with grad_type='amass' on the last 6 steps of the 1000-step chain (every step guided,
    started from q_sample(init_image)).  The guided chain is ill-conditioned (3e6-weighted gradient of a loss with
    hard masks), so each step is checked teacher-forced from the reference's own x_t."""
    g = golden("sampling.npz")
    B, T, bseed, nseed, skip = [int(v) for v in g["pose_guided_meta"]]
    sd, _ = posenet_state_dict(1)
    ds = synthetic.make_dataset('pose', seed=3, realistic_std=True)
    mean, std = torch.from_numpy(ds.Mean), torch.from_numpy(ds.Std)
    init = synthetic.plausible_motion(B, T, bseed, ds)
    cond = init.clone()
    body = synthetic.smplx_like_model(0)
    tables, tmap = do.create_diffusion("cosine", 1000, "")
    tape = NoiseTape(nseed)
    first = 1000 - skip - 1
    x_start = do.q_sample(tables, first, init, tape.randn(B, 294, 1, T))
    ref_xt, ref_x0, ref_s = (torch.from_numpy(g[k]) for k in ("pose_guided_xt", "pose_guided_x0", "pose_guided_sample"))
    assert float((x_start - ref_xt[0]).abs().max()) < 1e-5
    n_guided = 0
    for k, i in enumerate(range(first, -1, -1)):
        x = ref_xt[k]
        x0 = posenet_oracle.posenet_forward(sd, x, cond, torch.full((B,), tmap[i], dtype=torch.long))
        assert float((x0 - ref_x0[k]).abs().max()) < 2e-5 * max(1.0, float(ref_xt[k].abs().max()))  # |x_t| reaches ~1e2 here
        noise = tape.randn_like(x)
        gr = ko.guide_skating(ref_x0[k], mean, std, body)
        n_guided += int(gr.dim() > 0)
        y = do.p_sample_step(tables, i, x, ref_x0[k], noise, [(3e6, gr)] if gr.dim() > 0 else None)
        err = float((y - ref_s[k]).abs().max())
        assert err < TOL * max(1.0, float(ref_s[k].abs().max())), (k, err)
    assert n_guided == first + 1


def test_smplx_restatement_self_consistency():
    """The LBS restatement has no reference fixtures (parity unpinned): check identities instead."""
    body = synthetic.smplx_like_model(0, num_verts=2000)
    N = 3
    g = torch.Generator().manual_seed(9)
    betas = torch.randn(N, 10, generator=g)
    transl = torch.randn(N, 3, generator=g)
    zeros = torch.zeros(N, 3)
    # zero pose: joints are the regressed rest joints + translation, vertices are the shaped template + translation
    j, v = ko.smplx_forward(body, zeros, torch.zeros(N, 63), betas, transl, return_verts=True, dtype=torch.float64)
    m64 = {k: (t.double() if torch.is_tensor(t) else t) for k, t in body.items()}
    v_shaped = m64["v_template"] + torch.einsum('bl,mkl->bmk', torch.cat([betas.double(), torch.zeros(N, 10).double()], 1), m64["shapedirs"])
    J = torch.einsum('bik,ji->bjk', v_shaped, m64["J_regressor"])
    assert float((j - (J + transl.double()[:, None])).abs().max()) < 1e-6
    assert float((v - (v_shaped + transl.double()[:, None])).abs().max()) < 1e-6
    # a pure global rotation rotates every joint about the root joint
    go = torch.tensor([[0.3, -0.2, 0.5]]).repeat(N, 1)
    j2, _ = ko.smplx_forward(body, go, torch.zeros(N, 63), betas, transl, return_verts=False, dtype=torch.float64)
    R = ko.batch_rodrigues(go.double())
    expect = torch.einsum('nij,nkj->nki', R, J - J[:, :1]) + J[:, :1] + transl.double()[:, None]
    assert float((j2 - expect).abs().max()) < 1e-6
    # Rodrigues(rotmat_to_aa(R)) == R for generic rotations
    R6 = ko.rot6d_to_rotmat(torch.randn(32, 6, generator=g).double())
    assert float((ko.batch_rodrigues(ko.rotmat_to_aa(R6)) - R6).abs().max()) < 1e-5
