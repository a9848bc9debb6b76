"""Oracle (test infrastructure only): PoseNet denoiser forward, restated with torch-CPU functional ops.

Works directly on a reference-format state dict (keys as in reference model/posenet.py:59-72).
Reference: model/posenet.py:75-96 (forward), model/heads.py:112-176 (PositionalEncoding, TimestepEmbedder,
InputProcess, OutputProcess), torch.nn.TransformerEncoderLayer with norm_first=False, activation=gelu (exact erf),
layer_norm_eps=1e-5, dropout inactive (eval).
"""
import math

import torch
import torch.nn.functional as F


def positional_encoding(max_len, d_model, dtype=torch.float32):
    """heads.py:117-122 (the `pe` buffer, squeezed to [max_len, d_model])."""
    pe = torch.zeros(max_len, d_model)
    position = torch.arange(0, max_len, dtype=torch.float).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, d_model, 2).float() * (-math.log(10000.0) / d_model))
    pe[:, 0::2] = torch.sin(position * div_term)
    pe[:, 1::2] = torch.cos(position * div_term)
    return pe.to(dtype)


def encoder_layer(x, sd, prefix, num_heads):
    """One post-norm nn.TransformerEncoderLayer on x: [S, B, D]."""
    S, B, D = x.shape
    dh = D // num_heads
    qkv = F.linear(x, sd[prefix + "self_attn.in_proj_weight"], sd[prefix + "self_attn.in_proj_bias"])
    q, k, v = qkv.split(D, dim=-1)
    # [S, B, H, dh] -> [B, H, S, dh]
    q = q.reshape(S, B, num_heads, dh).permute(1, 2, 0, 3)
    k = k.reshape(S, B, num_heads, dh).permute(1, 2, 0, 3)
    v = v.reshape(S, B, num_heads, dh).permute(1, 2, 0, 3)
    att = torch.softmax(torch.matmul(q, k.transpose(-1, -2)) / math.sqrt(dh), dim=-1)
    ctx = torch.matmul(att, v).permute(2, 0, 1, 3).reshape(S, B, D)
    a = F.linear(ctx, sd[prefix + "self_attn.out_proj.weight"], sd[prefix + "self_attn.out_proj.bias"])
    x = F.layer_norm(x + a, (D,), sd[prefix + "norm1.weight"], sd[prefix + "norm1.bias"], 1e-5)
    h = F.gelu(F.linear(x, sd[prefix + "linear1.weight"], sd[prefix + "linear1.bias"]))
    h = F.linear(h, sd[prefix + "linear2.weight"], sd[prefix + "linear2.bias"])
    x = F.layer_norm(x + h, (D,), sd[prefix + "norm2.weight"], sd[prefix + "norm2.bias"], 1e-5)
    return x


def posenet_forward(sd, x_t, cond, timesteps, num_layers=8, num_heads=4, traj_feat_dim=22):
    """x_t, cond: [B, C, 1, T]; timesteps: int64 [B] (ORIGINAL timesteps, i.e. after _WrappedModel).
    Returns [B, C, 1, T] with channels [0:traj_feat_dim] copied from cond (posenet.py:94-95)."""
    dtype = x_t.dtype
    sd = {k: v.to(dtype) if v.is_floating_point() else v for k, v in sd.items()}
    B, C, _, T = x_t.shape
    pe = sd["sequence_pos_encoder.pe"][:, 0, :]  # [5000, D]
    D = pe.shape[1]
    # TimestepEmbedder (heads.py:145-146)
    emb = F.linear(pe[timesteps], sd["embed_timestep.time_embed.0.weight"], sd["embed_timestep.time_embed.0.bias"])
    emb = F.linear(F.silu(emb), sd["embed_timestep.time_embed.2.weight"], sd["embed_timestep.time_embed.2.bias"])
    emb = emb.unsqueeze(0)  # [1, B, D]
    # InputProcess x2 (heads.py:156-160)
    xs = x_t.permute(3, 0, 1, 2).reshape(T, B, C)
    cs = cond.permute(3, 0, 1, 2).reshape(T, B, C)
    x = F.linear(xs, sd["input_process.poseEmbedding.weight"], sd["input_process.poseEmbedding.bias"]) + F.linear(
        cs, sd["input_process_cond.poseEmbedding.weight"], sd["input_process_cond.poseEmbedding.bias"])
    xseq = torch.cat((emb, x), dim=0)  # [T+1, B, D]
    xseq = xseq + pe[: T + 1].unsqueeze(1)
    for l in range(num_layers):
        xseq = encoder_layer(xseq, sd, f"seqTransEncoder.layers.{l}.", num_heads)
    out = xseq[1:]
    out = F.linear(out, sd["output_process.poseFinal.weight"], sd["output_process.poseFinal.bias"])  # [T, B, 272]
    out = out.reshape(T, B, -1, 1).permute(1, 2, 3, 0)  # [B, 272, 1, T]
    return torch.cat([cond[:, 0:traj_feat_dim], out], dim=1)
