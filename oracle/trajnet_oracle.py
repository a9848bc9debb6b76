"""Oracle (test infrastructure only): TrajNet conv U-Net (+ TrajControl branch) forward, torch-CPU functional ops.

Works on a reference-format state dict.  Reference: model/trajnet.py:10-75 (ControlNet), :177-275 (TrajNet.forward),
model/heads.py:20-106 (ResidualTemporalBlock, SinusoidalPosEmb, Downsample1d, Upsample1d, Conv1dBlock).
"""
import math

import torch
import torch.nn.functional as F


def sinusoidal_pos_emb(t, dim):
    """heads.py:57-69."""
    half = dim // 2
    e = math.log(10000) / (half - 1)
    e = torch.exp(torch.arange(half, dtype=torch.float32) * -e)
    e = t[:, None].to(torch.float32) * e[None, :]
    return torch.cat((e.sin(), e.cos()), dim=-1)


def conv1d_block(x, sd, p, groups=8):
    """heads.py:90-106: Conv1d(k, pad=k//2) -> GroupNorm(8) -> Mish."""
    w = sd[p + "block.0.weight"]
    x = F.conv1d(x, w, sd[p + "block.0.bias"], padding=w.shape[-1] // 2)
    x = F.group_norm(x, groups, sd[p + "block.2.weight"], sd[p + "block.2.bias"], 1e-5)
    return F.mish(x)


def rtb(x, t, sd, p):
    """heads.py:43-54 ResidualTemporalBlock; t=None for blocks built with input_t=False."""
    out = conv1d_block(x, sd, p + "blocks.0.")
    if (p + "time_mlp.1.weight") in sd and t is not None:
        out = out + F.linear(F.mish(t), sd[p + "time_mlp.1.weight"], sd[p + "time_mlp.1.bias"])[:, :, None]
    out = conv1d_block(out, sd, p + "blocks.1.")
    if (p + "residual_conv.weight") in sd:
        res = F.conv1d(x, sd[p + "residual_conv.weight"], sd[p + "residual_conv.bias"])
    else:
        res = x
    return out + res


def down(x, sd, p):
    return F.conv1d(x, sd[p + "conv.weight"], sd[p + "conv.bias"], stride=2, padding=1)


def up(x, sd, p):
    return F.conv_transpose1d(x, sd[p + "conv.weight"], sd[p + "conv.bias"], stride=2, padding=1)


def c1x1(x, sd, p):
    return F.conv1d(x, sd[p + "weight"], sd[p + "bias"])


def trajnet_forward(sd, x_t, cond, time, control_cond=None, time_dim=32):
    """x_t, cond: [B, T, traj_dim]; time: int64 [B]; control_cond: [B, T, 272] or None (vanilla TrajNet).
    Returns [B, T, traj_dim]."""
    dtype = x_t.dtype
    sd = {k: v.to(dtype) for k, v in sd.items()}
    trajcontrol = control_cond is not None
    # time_mlp (trajnet.py:120-125)
    t = sinusoidal_pos_emb(time, time_dim).to(dtype)
    t = F.linear(t, sd["time_mlp.1.weight"], sd["time_mlp.1.bias"])
    t = F.linear(F.mish(t), sd["time_mlp.3.weight"], sd["time_mlp.3.bias"])

    # condition pyramid (trajnet.py:192-208)
    c = cond.permute(0, 2, 1)
    h_cond = []
    c = rtb(c, None, sd, "cond_enc1.")
    h_cond.append(c)
    c = down(c, sd, "cond_downsample1.")
    c = rtb(c, None, sd, "cond_enc2.")
    h_cond.append(c)
    c = down(c, sd, "cond_downsample2.")
    c = rtb(c, None, sd, "cond_enc3.")
    h_cond.append(c)
    c = down(c, sd, "cond_downsample3.")
    c = rtb(c, None, sd, "cond_enc4.")
    h_cond.append(c)

    z = None
    if trajcontrol:  # trajnet.py:43-75
        k = control_cond.permute(0, 2, 1)
        k = c1x1(k, sd, "controlnet.control_zero_conv_0.")
        k = rtb(k, t, sd, "controlnet.control_enc1.")
        z1 = c1x1(k, sd, "controlnet.control_zero_conv_1.")
        k = down(torch.cat([k, h_cond[0]], dim=1), sd, "controlnet.control_downsample1.")
        k = rtb(k, t, sd, "controlnet.control_enc2.")
        z2 = c1x1(k, sd, "controlnet.control_zero_conv_2.")
        k = down(torch.cat([k, h_cond[1]], dim=1), sd, "controlnet.control_downsample2.")
        k = rtb(k, t, sd, "controlnet.control_enc3.")
        z3 = c1x1(k, sd, "controlnet.control_zero_conv_3.")
        k = down(torch.cat([k, h_cond[2]], dim=1), sd, "controlnet.control_downsample3.")
        k = rtb(k, t, sd, "controlnet.control_enc4.")
        z4 = c1x1(k, sd, "controlnet.control_zero_conv_4.")
        k = down(torch.cat([k, h_cond[3]], dim=1), sd, "controlnet.control_downsample4.")
        k = rtb(k, t, sd, "controlnet.control_mid_block1.")
        k = rtb(k, t, sd, "controlnet.control_mid_block2.")
        zm = c1x1(k, sd, "controlnet.control_zero_conv_mid.")
        z = (z1, z2, z3, z4, zm)

    # U-Net (trajnet.py:216-275)
    x = x_t.permute(0, 2, 1)
    h = []
    x = rtb(x, t, sd, "diff_enc1.")
    h.append(x)
    x = down(torch.cat([x, h_cond[0]], dim=1), sd, "diff_downsample1.")
    x = rtb(x, t, sd, "diff_enc2.")
    h.append(x)
    x = down(torch.cat([x, h_cond[1]], dim=1), sd, "diff_downsample2.")
    x = rtb(x, t, sd, "diff_enc3.")
    h.append(x)
    x = down(torch.cat([x, h_cond[2]], dim=1), sd, "diff_downsample3.")
    x = rtb(x, t, sd, "diff_enc4.")
    h.append(x)
    x = down(torch.cat([x, h_cond[3]], dim=1), sd, "diff_downsample4.")
    x = rtb(x, t, sd, "diff_mid_block1.")
    x = rtb(x, t, sd, "diff_mid_block2.")
    if trajcontrol:
        x = x + z[4]
    x = rtb(torch.cat([up(x, sd, "diff_upsample4."), h[3]], dim=1), t, sd, "diff_dec4.")
    if trajcontrol:
        x = x + z[3]
    x = rtb(torch.cat([up(x, sd, "diff_upsample3."), h[2]], dim=1), t, sd, "diff_dec3.")
    if trajcontrol:
        x = x + z[2]
    x = rtb(torch.cat([up(x, sd, "diff_upsample2."), h[1]], dim=1), t, sd, "diff_dec2.")
    if trajcontrol:
        x = x + z[1]
    x = rtb(torch.cat([up(x, sd, "diff_upsample1."), h[0]], dim=1), t, sd, "diff_dec1.")
    if trajcontrol:
        x = x + z[0]
    x = conv1d_block(x, sd, "diff_final_conv.0.")
    x = c1x1(x, sd, "diff_final_conv.1.")
    return x.permute(0, 2, 1)
