"""GPU parity tests: SMPL-X joints FK / full LBS, skating-guidance gradient (analytic VJP vs the reference's autograd)
and guided PoseNet sampling steps."""
import argparse

import numpy as np
import pytest
import torch

from helpers import NoiseTape, TOL, golden
from oracle import diffusion_oracle as do
from oracle import kinematics_oracle as ko
from oracle import posenet_oracle
from rohm_b200 import diffusion, synthetic
from rohm_b200.body_model import BodyModel, kernels_for
from rohm_b200.posenet import PoseNet

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def body(cuda_device):
    return BodyModel.create('', device=cuda_device, seed=0), synthetic.smplx_like_model(0)


def _params(N, seed):
    g = torch.Generator().manual_seed(seed)
    return (0.3 * torch.randn(N, 3, generator=g), 0.3 * torch.randn(N, 63, generator=g), torch.randn(N, 10, generator=g),
            torch.randn(N, 3, generator=g))


def test_joints_and_vertices_match_oracle(body, cuda_device):
    bm, model = body
    for N in (1, 5, 37):
        go, bp, be, tr = _params(N, 10 + N)
        out = bm(transl=tr.to(cuda_device), global_orient=go.to(cuda_device), body_pose=bp.to(cuda_device),
                 betas=be.to(cuda_device), jaw_pose=torch.zeros(N, 3), left_hand_pose=torch.zeros(N, 45),
                 expression=torch.zeros(N, 10))
        j64, v64 = ko.smplx_forward(model, go, bp, be, tr, return_verts=True, dtype=torch.float64)
        assert out.joints.shape == (N, 55, 3) and out.vertices.shape == (N, 10475, 3)
        assert float((out.joints.cpu().double() - j64).abs().max()) < 2e-5
        assert float((out.vertices.cpu().double() - v64).abs().max()) < 5e-5
        j32, v32 = ko.smplx_forward(model, go, bp, be, tr, return_verts=True)
        assert float((out.vertices.cpu() - v32).abs().max()) < TOL


def test_zero_pose_and_edge_rotations(body, cuda_device):
    bm, model = body
    N = 4
    go = torch.tensor([[0., 0, 0], [3.1415, 0, 0], [0, 1e-7, 0], [1.2, -2.0, 0.7]])
    bp = torch.zeros(N, 63)
    bp[3] = 0.5
    be, tr = torch.zeros(N, 10), torch.zeros(N, 3)
    out = bm(transl=tr.to(cuda_device), global_orient=go.to(cuda_device), body_pose=bp.to(cuda_device),
             betas=be.to(cuda_device))
    j, v = ko.smplx_forward(model, go, bp, be, tr, return_verts=True, dtype=torch.float64)
    assert float((out.joints.cpu().double() - j).abs().max()) < 2e-5
    assert float((out.vertices.cpu().double() - v).abs().max()) < 5e-5


def test_pitched_tma_vertex_store_equals_the_dense_store(body, cuda_device, monkeypatch):
    """The fused LBS launch writes pitched rows (16-byte-aligned frames, TMA bulk stores; the default) and dense rows (4-byte
    stores, ROHM_B200_LBS_TMA_STORE=0) from the same accumulators: bit-identical vertices, ragged last row / column tiles
    included (N = 1, 129, 300 frames; 3 V = 327 x 96 + 33 columns)."""
    from rohm_b200.body_model import BodyKernels
    bm, _ = body
    monkeypatch.setenv("ROHM_B200_LBS_TMA_STORE", "1")
    k_tma = BodyKernels(bm, cuda_device, 300, True)
    monkeypatch.setenv("ROHM_B200_LBS_TMA_STORE", "0")
    k_dense = BodyKernels(bm, cuda_device, 300, True)
    assert k_dense.vertex_pitch == 0
    if not k_tma.vertex_pitch:
        pytest.skip("fused LBS launch not in use on this handle")
    assert k_tma.vertex_pitch % 4 == 0 and k_tma.vertex_pitch >= 3 * k_tma.V
    for N in (1, 129, 300):
        go, bp, be, tr = (t.to(cuda_device) for t in _params(N, 40 + N))
        j1, v1 = k_tma.forward(go, bp, be, tr, True)
        j2, v2 = k_dense.forward(go, bp, be, tr, True)
        assert v1.shape == v2.shape == (N, k_tma.V, 3) and v2.is_contiguous()
        assert v1.data_ptr() % 16 == 0 and (N == 1 or v1.stride() == (k_tma.vertex_pitch, 3, 1))  # torch normalises size-1 strides
        assert torch.equal(v1, v2) and torch.equal(j1, j2)
        assert torch.equal(v1.contiguous().view(N, -1), v2.view(N, -1))


def _motion(B, T, seed, dseed=3):
    ds = synthetic.make_dataset('pose', seed=dseed, realistic_std=True)
    return ds, synthetic.plausible_motion(B, T, seed, ds)


def test_recover_joints_from_repr_matches_reference_golden(body, cuda_device):
    """recover_from_repr_smpl('smplx_params') incl. the rot6d -> rotmat -> axis-angle chain, vs the reference."""
    bm, model = body
    g = golden("kinematics.npz")
    B, T, seed, dseed = [int(v) for v in g["kin_meta"]]
    ds, x = _motion(B, T, seed, dseed)
    k = kernels_for(bm, cuda_device, B * T, with_vertices=False)
    mean, std = torch.from_numpy(ds.Mean).to(cuda_device), torch.from_numpy(ds.Std).to(cuda_device)
    joints = k.from_repr(x.to(cuda_device), mean, std, want_vertices=False)
    assert float((joints.cpu() - torch.from_numpy(g["smplx_joints"])).abs().max()) < 2e-5


def _posenet(cuda_device, ds):
    m = PoseNet(dataset=ds, body_feat_dim=294, latent_dim=512, ff_size=1024, num_layers=8, num_heads=4,
                device=cuda_device, traj_feat_dim=22)
    sd = {k: v.cpu() for k, v in synthetic.synth_state_dict(m, 1).items()}
    m.load_state_dict(sd)
    m.to(cuda_device).eval()
    return m, sd


def test_skating_gradient_matches_reference_golden(cuda_device):
    """guide_skating_with_smpl: analytic CUDA VJP vs the reference's autograd through the body model."""
    g = golden("kinematics.npz")
    B, T, seed, dseed = [int(v) for v in g["kin_meta"]]
    ds, x = _motion(B, T, seed, dseed)
    m, _ = _posenet(cuda_device, ds)
    xg = x.to(cuda_device)
    grad = m.guide_skating_with_smpl({'x_t': xg}, {'pred_xstart': xg}, None, compute_grad='x_0').cpu()
    ref = torch.from_numpy(g["skating_grad"])
    scale = float(ref.abs().max())
    assert scale > 1e-3
    assert float((grad - ref).abs().max()) < 2e-4 * scale
    assert float(grad[:, :22].abs().max()) == 0.0 and float(grad[:, -4:].abs().max()) == 0.0


@pytest.mark.parametrize("B,T,seed", [(1, 2, 1), (3, 50, 2), (8, 143, 3)])
def test_skating_gradient_matches_oracle_autograd(cuda_device, B, T, seed):
    ds, x = _motion(B, T, seed)
    m, _ = _posenet(cuda_device, ds)
    body = synthetic.smplx_like_model(0)
    xg = x.to(cuda_device)
    grad = m.guide_skating_with_smpl({'x_t': xg}, {'pred_xstart': xg}, None, compute_grad='x_0').cpu()
    ref = ko.guide_skating(x.double(), torch.from_numpy(ds.Mean).double(), torch.from_numpy(ds.Std).double(), body)
    if ref.dim() == 0:
        assert float(grad.abs().max()) == 0.0
    else:
        scale = float(ref.abs().max())
        assert float((grad.double() - ref).abs().max()) < 2e-4 * scale


def test_nothing_skates_gives_zero_gradient(cuda_device):
    ds, x = _motion(2, 10, 5)
    x[:, -4:] = 0.0  # no predicted contact anywhere
    m, _ = _posenet(cuda_device, ds)
    xg = x.to(cuda_device)
    grad = m.guide_skating_with_smpl({'x_t': xg}, {'pred_xstart': xg}, None, compute_grad='x_0')
    assert float(grad.abs().max()) == 0.0


def test_guided_steps_match_reference_golden(cuda_device):
    """p_sample_with_grad(grad_type='amass'), every step guided, teacher-forced from the reference's own x_t (the
    guided chain is ill-conditioned, see tests/test_oracle_golden.py)."""
    g = golden("sampling.npz")
    B, T, bseed, nseed, skip = [int(v) for v in g["pose_guided_meta"]]
    ds = synthetic.make_dataset('pose', seed=3, realistic_std=True)
    m, sd = _posenet(cuda_device, ds)
    init = synthetic.plausible_motion(B, T, bseed, ds)
    batch = {'cond': init.clone().to(cuda_device)}
    args = argparse.Namespace(noise_schedule='cosine', sigma_small=True)
    d = diffusion.create_gaussian_diffusion(args, diffusion, diffusion.SpacedDiffusionPoseNet, 1000, '', cuda_device)
    tape = NoiseTape(nseed, cuda_device)
    d._randn, d._randn_like = tape.randn, tape.randn_like
    first = 1000 - skip - 1
    t_rows = d._t_rows(B, cuda_device)
    x_start = d.q_sample(init.to(cuda_device), t_rows[first], tape.randn(B, 294, 1, T))
    ref_xt, ref_s = torch.from_numpy(g["pose_guided_xt"]), torch.from_numpy(g["pose_guided_sample"])
    assert float((x_start.cpu() - ref_xt[0]).abs().max()) < 1e-5
    for k, i in enumerate(range(first, -1, -1)):
        out = d.p_sample_with_grad(m, batch, ref_xt[k].to(cuda_device), t_rows[i], clip_denoised=False,
                                   grad_type='amass', _step_index=i)
        err = float((out['sample'].cpu() - ref_s[k]).abs().max())
        # the 3e6-weighted gradient amplifies the 2e-5 denoiser difference; samples reach |x| ~ 700 on this synthetic
        # input, so the bound is relative to the sample magnitude
        assert err < 1e-3 * max(1.0, float(ref_s[k].abs().max())), (k, err)


def test_full_lbs_throughput_shape(body, cuda_device):
    """N3-sized call: 32 clips x 143 frames of vertices (575 MB) runs and agrees with the oracle on a sample."""
    bm, model = body
    N = 32 * 143
    go, bp, be, tr = _params(N, 99)
    out = bm(transl=tr.to(cuda_device), global_orient=go.to(cuda_device), body_pose=bp.to(cuda_device),
             betas=be.to(cuda_device))
    idx = torch.tensor([0, 127, 128, 1000, 4479, 4480, N - 1])  # first / last rows of row tiles, the ragged last tile
    j, v = ko.smplx_forward(model, go[idx], bp[idx], be[idx], tr[idx], return_verts=True, dtype=torch.float64)
    assert float((out.vertices[idx.to(cuda_device)].cpu().double() - v).abs().max()) < 5e-5


def test_global_guidance_mode_reproduces_the_unsharded_gradient(cuda_device):
    """Clip-sharded guidance with batch-GLOBAL normalisers (the 4-float all-reduce of rohm_b200.parallel.global_guidance,
    emulated here by adding the other shard's sums): the concatenated shard gradients equal the unsharded gradient, while the
    default per-shard normalisation does not."""
    B, T = 6, 40
    ds, x = _motion(B, T, 11)
    m, _ = _posenet(cuda_device, ds)
    bm = m.smplx_model
    mean, std = torch.from_numpy(ds.Mean).to(cuda_device), torch.from_numpy(ds.Std).to(cuda_device)
    xg = x.to(cuda_device)
    full = kernels_for(bm, cuda_device, B * T, with_vertices=False).skating_guidance(xg, mean, std)
    halves = [xg[:2].contiguous(), xg[2:].contiguous()]  # ragged shards: 2 + 4 clips
    k = kernels_for(bm, cuda_device, B * T, with_vertices=False)
    local = []
    for h in halves:  # pass 1: every shard's own sums
        rec = {}
        k.skating_guidance_global(h, mean, std, lambda s, rec=rec: rec.setdefault("s", s.clone()))
        local.append(rec["s"])
    total = local[0] + local[1]
    grads = [k.skating_guidance_global(h, mean, std, lambda s: s.copy_(total)) for h in halves]
    got = torch.cat(grads, dim=0)
    scale = float(full.abs().max())
    assert scale > 0 and float((got - full).abs().max()) < 1e-5 * scale
    per_shard = torch.cat([k.skating_guidance(h, mean, std) for h in halves], dim=0)
    assert float((per_shard - full).abs().max()) > 1e-2 * scale  # the default contract is per-shard semantics
    # and through the model hook
    m.guidance_sum_reducer = lambda s: s.copy_(total)
    g2 = m.guide_skating_with_smpl({}, {'pred_xstart': halves[1]}, None, compute_grad='x_0')
    del m.guidance_sum_reducer
    assert torch.equal(g2, grads[1])
