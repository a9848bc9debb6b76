"""PoseNet: drop-in for reference model/posenet.py:11-96 (constructor, attributes, state-dict keys, call signature),
with the forward pass executed by the CUDA engine behind ``rohm_posenet_*`` (include/rohm_b200.h).

The module owns its parameters in torch containers named exactly like the reference so ``load_state_dict(strict=True)``
works on released checkpoints; on the first forward (and whenever parameters change) they are repacked into the
engine's TF32 hi/lo layout.  There is no eager / CPU forward: calling the model without a B200 raises.
"""
import ctypes as C
import os

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from ._lib import RohmB200Error
from .heads import InputProcess, OutputProcess, PositionalEncoding, TimestepEmbedder


DEFAULT_PRECISION = "f16x2"


def _precision_from_env(supports_f16=True):
    """ROHM_B200_PRECISION: 'f16x2' (default: fp16 hi/lo pairs, fp32-grade), 'tf32x3' (TF32 hi/lo pairs, fp32-grade),
    'tf32' (single pass, fast, ~1e-3).  Engines without an fp16 path (the LBS blend GEMM) run 'f16x2' as 'tf32x3'."""
    v = os.environ.get("ROHM_B200_PRECISION", DEFAULT_PRECISION).lower()
    if v in ("f16x2", "fp16x2", "parity"):
        return _lib.PRECISION_F16X2 if supports_f16 else _lib.PRECISION_TF32X3
    if v in ("tf32x3", "3xtf32", "fp32"):
        return _lib.PRECISION_TF32X3
    if v in ("tf32", "fast"):
        return _lib.PRECISION_TF32
    raise RohmB200Error(f"ROHM_B200_PRECISION={v!r}: expected 'f16x2' (default), 'tf32x3' or 'tf32' (fast)")


class PoseNetEngine:
    """Owns one rohm_posenet handle (device weights + workspace) sized for (max_batch, max_frames)."""

    def __init__(self, module, device, max_batch, max_frames, precision):
        self.lib = _lib.load()
        self.ctx = _lib.ctx(device.index)
        self.device = device
        self.max_batch, self.max_frames, self.precision = max_batch, max_frames, precision
        sd = {k: v.detach().to(device=device, dtype=torch.float32).contiguous() for k, v in module.state_dict().items()
              if not k.startswith("smplx_model.")}
        self._keep = sd  # keeps the source tensors alive during create (the library copies them)
        L = module.num_layers
        layers = (_lib.PoseNetLayerW * L)()
        for l in range(L):
            p = f"seqTransEncoder.layers.{l}."
            for field, key in (("in_proj_w", "self_attn.in_proj_weight"), ("in_proj_b", "self_attn.in_proj_bias"),
                               ("out_proj_w", "self_attn.out_proj.weight"), ("out_proj_b", "self_attn.out_proj.bias"),
                               ("lin1_w", "linear1.weight"), ("lin1_b", "linear1.bias"),
                               ("lin2_w", "linear2.weight"), ("lin2_b", "linear2.bias"),
                               ("norm1_w", "norm1.weight"), ("norm1_b", "norm1.bias"),
                               ("norm2_w", "norm2.weight"), ("norm2_b", "norm2.bias")):
                setattr(layers[l], field, sd[p + key].data_ptr())
        pe = sd["sequence_pos_encoder.pe"].reshape(-1, module.latent_dim).contiguous()
        self._keep["__pe2d"] = pe
        w = _lib.PoseNetW()
        w.d_model, w.ff_size, w.num_layers, w.num_heads = module.latent_dim, module.ff_size, L, module.num_heads
        w.in_feats, w.out_feats, w.traj_feats = module.input_feats, module.output_process.output_feats, module.traj_feat_dim
        w.pe_len = pe.shape[0]
        for field, key in (("in_w", "input_process.poseEmbedding.weight"), ("in_b", "input_process.poseEmbedding.bias"),
                           ("cond_w", "input_process_cond.poseEmbedding.weight"),
                           ("cond_b", "input_process_cond.poseEmbedding.bias"),
                           ("t0_w", "embed_timestep.time_embed.0.weight"), ("t0_b", "embed_timestep.time_embed.0.bias"),
                           ("t2_w", "embed_timestep.time_embed.2.weight"), ("t2_b", "embed_timestep.time_embed.2.bias"),
                           ("out_w", "output_process.poseFinal.weight"), ("out_b", "output_process.poseFinal.bias")):
            setattr(w, field, sd[key].data_ptr())
        w.pe = pe.data_ptr()
        w.layers = layers
        if w.in_feats != w.traj_feats + w.out_feats:
            raise RohmB200Error(f"PoseNet: body_feat_dim ({w.in_feats}) must equal traj_feat_dim ({w.traj_feats}) + "
                                f"pose_feat_dim ({w.out_feats})")
        handle = C.c_void_p()
        with torch.cuda.device(device):
            rc = self.lib.rohm_posenet_create(self.ctx, C.byref(w), max_batch, max_frames, precision, C.byref(handle))
        _lib.check(rc, self.ctx)
        self.handle = handle
        self._keep = None  # the library owns its copies now
        if os.environ.get("ROHM_B200_PDL", "1") == "0":
            self.lib.rohm_posenet_set_option(handle, 1, 0)
        if os.environ.get("ROHM_B200_GRAPH", "1") == "0":
            self.lib.rohm_posenet_set_option(handle, 0, 0)
        # the condition whose step-invariant embedding the engine currently holds: a STRONG reference (so the caching
        # allocator cannot hand its address to a different tensor while it is cached) plus its version counter
        self.cond_ref = None
        self.cond_version = -1
        from . import ops
        self.op_key = ops.register_engine(self)

    def __del__(self):
        h = getattr(self, "handle", None)
        try:
            from . import ops
            ops.unregister_engine(getattr(self, "op_key", 0))
        except Exception:
            pass
        if h:
            try:
                self.lib.rohm_posenet_destroy(h)
            except Exception:
                pass
            self.handle = None

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def set_cond(self, cond):
        B, _, _, T = cond.shape
        rc = self.lib.rohm_posenet_set_cond(self.handle, C.c_void_p(cond.data_ptr()), B, T, self._stream())
        _lib.check(rc, self.ctx)

    def forward(self, x_t, timesteps, out=None):
        """The denoiser call, through the custom op torch.ops.rohm.posenet_forward (an explicit `out` skips the op layer)."""
        if out is None:
            return torch.ops.rohm.posenet_forward(self.op_key, x_t, timesteps)
        return self._forward_impl(x_t, timesteps, out)

    def _forward_impl(self, x_t, timesteps, out=None):
        B, _, _, T = x_t.shape
        if out is None:
            out = torch.empty_like(x_t)
        rc = self.lib.rohm_posenet_forward(self.handle, C.c_void_p(x_t.data_ptr()), C.c_void_p(timesteps.data_ptr()),
                                           C.c_void_p(out.data_ptr()), B, T, self._stream())
        _lib.check(rc, self.ctx)
        return out

    def sample_step(self, x_t, timesteps, coef_row):
        """One whole ancestral step as one graph launch (rohm_posenet_sample_step): -> (pred_xstart, x_{t-1}); the noise is
        what torch.randn_like(x_t) would have drawn (torch's CUDA generator is advanced accordingly)."""
        from .ops import cuda_generator_state
        B, _, _, T = x_t.shape
        x0, nxt = torch.empty_like(x_t), torch.empty_like(x_t)
        gen, seed, offset = cuda_generator_state(x_t.device)
        inc = C.c_uint64(0)
        rc = self.lib.rohm_posenet_sample_step(self.handle, C.c_void_p(x_t.data_ptr()), C.c_void_p(timesteps.data_ptr()),
                                               C.c_void_p(x0.data_ptr()), C.c_void_p(nxt.data_ptr()),
                                               C.c_void_p(coef_row.data_ptr()), seed, offset, C.byref(inc), B, T, self._stream())
        _lib.check(rc, self.ctx)
        gen.set_offset(offset + int(inc.value))
        return x0, nxt

    def profile(self, x_t, timesteps):
        """One forward with per-kernel CUDA-event timing -> ({category: ms}, {category: launches})."""
        B, _, _, T = x_t.shape
        out = torch.empty_like(x_t)
        ms = (C.c_float * 4)()
        n = (C.c_int * 4)()
        rc = self.lib.rohm_posenet_profile(self.handle, C.c_void_p(x_t.data_ptr()), C.c_void_p(timesteps.data_ptr()),
                                           C.c_void_p(out.data_ptr()), B, T, self._stream(), ms, n)
        _lib.check(rc, self.ctx)
        names = ("gemm", "attention", "layernorm", "other")
        return {k: float(ms[i]) for i, k in enumerate(names)}, {k: int(n[i]) for i, k in enumerate(names)}

    @property
    def launches_per_forward(self):
        return int(self.lib.rohm_posenet_launches_per_forward(self.handle))


def _create_body_model(body_model_path, device):
    """The SMPL-X body model submodule (reference posenet.py:57-58).  Uses the real ``smplx`` package when it is
    installed (so checkpoint keys under ``smplx_model.*`` match); otherwise the package's own BodyModel."""
    try:
        import smplx  # noqa: F401  third-party, optional
        m = smplx.create(model_path=body_model_path, model_type="smplx", gender='neutral', flat_hand_mean=True,
                         use_pca=False)
        return m.to(device) if device is not None else m
    except ImportError:
        from .body_model import BodyModel
        return BodyModel.create(body_model_path, device=device)


class PoseNet(nn.Module):
    def __init__(self, dataset, body_feat_dim, nfeats=1,
                 latent_dim=256, ff_size=1024, num_layers=8, num_heads=4, dropout=0.1, activation="gelu",
                 body_model_path='',
                 device=None,
                 traj_feat_dim=4,
                 weight_loss_rec_repr_full_body=0.0,
                 weight_loss_repr_foot_contact_mse=0.0,
                 weight_loss_joint_pos_global=0.0,
                 weight_loss_joint_vel_global=0.0, weight_loss_joint_smooth=0.0,
                 weight_loss_foot_skating=0.0,
                 start_skating_loss_epoch=0,
                 ):
        super().__init__()
        if activation != "gelu":
            raise RohmB200Error("PoseNet: only activation='gelu' (the configuration RoHM ships) is implemented")
        self.dataset = dataset
        self.body_feat_dim = body_feat_dim
        self.nfeats = nfeats
        self.traj_feat_dim = traj_feat_dim
        self.foot_joint_index_list = [7, 10, 8, 11]  # left ankle, left toe, right ankle, right toe
        self.foot_skating_vel_thres = 0.1
        self.fps = 30
        self.latent_dim = latent_dim
        self.ff_size = ff_size
        self.num_layers = num_layers
        self.num_heads = num_heads
        self.dropout = dropout
        self.activation = activation
        self.input_feats = self.body_feat_dim * self.nfeats
        self.normalize_output = False
        self.device = device
        self.weight_loss_rec_repr_full_body = weight_loss_rec_repr_full_body
        self.weight_loss_repr_foot_contact_mse = weight_loss_repr_foot_contact_mse
        self.weight_loss_joint_pos_global = weight_loss_joint_pos_global
        self.weight_loss_joint_vel_global = weight_loss_joint_vel_global
        self.weight_loss_joint_smooth = weight_loss_joint_smooth
        self.weight_loss_foot_skating = weight_loss_foot_skating
        self.start_skating_loss_epoch = start_skating_loss_epoch

        self.smplx_model = _create_body_model(body_model_path, device)
        self.input_process = InputProcess(self.input_feats, self.latent_dim)
        self.input_process_cond = InputProcess(self.input_feats, self.latent_dim)
        self.sequence_pos_encoder = PositionalEncoding(self.latent_dim, self.dropout)
        enc_layer = nn.TransformerEncoderLayer(d_model=self.latent_dim, nhead=self.num_heads,
                                               dim_feedforward=self.ff_size, dropout=self.dropout,
                                               activation=self.activation)
        self.seqTransEncoder = nn.TransformerEncoder(enc_layer, num_layers=self.num_layers,
                                                     enable_nested_tensor=False)
        self.embed_timestep = TimestepEmbedder(self.latent_dim, self.sequence_pos_encoder)
        self.output_process = OutputProcess(self.dataset.pose_feat_dim, self.latent_dim, self.nfeats)

        self.precision = None  # None -> ROHM_B200_PRECISION env (default f16x2)
        self._engine = None
        self._engine_fingerprint = None

    # ---------------------------------------------------------------- engine management
    def _fingerprint(self):
        return tuple((p.data_ptr(), p._version) for p in self.parameters())

    def invalidate_engine(self):
        """Forces the weights to be repacked on the next forward (call after mutating parameters in place)."""
        self._engine = None

    def _apply(self, fn, *a, **k):
        self._engine = None
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, state_dict, strict=True, **k):
        """nn.Module.load_state_dict, plus: when the body model is the package's own BodyModel (no ``smplx`` package
        installed), the ``smplx_model.*`` entries of a reference checkpoint are adopted by buffer name
        (BodyModel.load_smplx_state) instead of being matched key by key -- real smplx registers more buffers than
        RoHM's calls need, so a strict key match could never succeed."""
        self._engine = None
        from .body_model import BodyModel
        if isinstance(self.smplx_model, BodyModel):
            prefix = "smplx_model."
            body = {key[len(prefix):]: v for key, v in state_dict.items() if key.startswith(prefix)}
            rest = {key: v for key, v in state_dict.items() if not key.startswith(prefix)}
            own = {prefix + n: b for n, b in self.smplx_model.state_dict().items()}
            foreign = [n for n in body if (prefix + n) not in own or own[prefix + n].shape != body[n].shape]
            if foreign:
                self.smplx_model.load_smplx_state(body)
                rest.update({prefix + n: b for n, b in self.smplx_model.state_dict().items()})
            else:
                rest.update({prefix + n: v for n, v in body.items()})
                for key, v in own.items():
                    rest.setdefault(key, v)
            state_dict = rest
        return super().load_state_dict(state_dict, strict=strict, **k)

    def engine(self, B, T, device):
        if self.training:
            raise RohmB200Error("PoseNet: the CUDA engine implements the inference path (model.eval()); training is "
                                "out of scope")
        prec = self.precision if self.precision is not None else _precision_from_env()
        e = self._engine
        if (e is None or e.device != device or B > e.max_batch or T > e.max_frames or e.precision != prec):
            mb = max(B, e.max_batch if e is not None and e.device == device else 0)
            mf = max(T, e.max_frames if e is not None and e.device == device else 0)
            self._engine = None
            e = PoseNetEngine(self, device, mb, mf, prec)
            self._engine = e
            self._engine_fingerprint = self._fingerprint()
        return e

    def prepare_cond(self, cond):
        """Runs the step-invariant part of the forward for this condition tensor if it is new or was modified."""
        if cond.device.type != "cuda":
            raise RohmB200Error("PoseNet: batch tensors must live on a CUDA device (no CPU path)")
        B, Cc, _, T = cond.shape
        e = self.engine(B, T, cond.device)
        # Same tensor OBJECT, unmodified since it was embedded -> reuse.  Identity (not data_ptr): a freed condition's
        # address and version count can be handed to the next batch's tensor by the caching allocator.
        if e.cond_ref is not cond or e.cond_version != cond._version:
            fp = self._fingerprint()  # parameters are re-checked once per new condition, not per step
            if fp != self._engine_fingerprint:
                self._engine = None
                e = self.engine(B, T, cond.device)
            c = cond if (cond.is_contiguous() and cond.dtype == torch.float32) else cond.contiguous().float()
            e.set_cond(c)
            e.cond_ref, e.cond_version = cond, cond._version
        return e

    def invalidate_cond(self):
        """Forget the cached step-invariant condition embedding (the samplers call this at the start of every loop,
        so a condition can never outlive the loop it was embedded for)."""
        if self._engine is not None:
            self._engine.cond_ref, self._engine.cond_version = None, -1

    # ---------------------------------------------------------------- test-time guidance
    def _norm_stats(self, device):
        key = (str(device), id(self.dataset))
        if getattr(self, "_norm_cache_key", None) != key:
            self._norm_cache = (torch.from_numpy(np.ascontiguousarray(self.dataset.Mean, dtype=np.float32)).to(device),
                                torch.from_numpy(np.ascontiguousarray(self.dataset.Std, dtype=np.float32)).to(device))
            self._norm_cache_key = key
        return self._norm_cache

    def guide_skating_with_smpl(self, batch, out, denoise_t, compute_grad='x_t'):
        """Gradient of -(foot-skating loss from SMPL-X joints + from the joint-based representation) w.r.t. x_t or the
        predicted x_0 (reference posenet.py:196-257), [bs, body_feat_dim, 1, T], trajectory and contact channels zero.
        One fused analytic forward+VJP (rohm_skating_guidance) instead of autograd through the body model; when nothing
        skates the result is an all-zero tensor (the reference returns a 0-dim zero), so no host sync is needed."""
        from .body_model import kernels_for
        x = batch['x_t'] if compute_grad == 'x_t' else out['pred_xstart']
        x = x.detach()
        x = x if (x.is_contiguous() and x.dtype == torch.float32) else x.contiguous().float()
        if self.dataset.traj_feat_dim != 22 or x.shape[1] != 294:
            raise RohmB200Error("guide_skating_with_smpl: implemented for the 294-channel representation with the "
                                "22-channel trajectory block (the configuration RoHM ships)")
        B, _, _, T = x.shape
        mean, std = self._norm_stats(x.device)
        k = kernels_for(self.smplx_model, x.device, B * T, with_vertices=False)
        reducer = getattr(self, "guidance_sum_reducer", None)
        if reducer is not None:
            # clip-sharded run reproducing the unsharded batch: the loss normalisers are batch-wide counts (reference
            # posenet.py:230-233), so the four sums are all-reduced over the ranks (rohm_b200.parallel.global_guidance)
            return k.skating_guidance_global(x, mean, std, reducer)
        return k.skating_guidance(x, mean, std)

    def _camera_affine(self, batch, device):
        """[B, 3, 4] canonical -> camera map of guide_2d_projection_with_smpl (reference posenet.py:285-297):
        p_cam = inv(cam_R) (inv(transf_matrix) p_cano - cam_t).  Per-clip 4x4 inverses: host-sized work, cached per
        transf_matrix tensor so the guided steps of one loop compute it once."""
        tm = batch['transf_matrix']
        cache = getattr(self, "_cam_cache", None)
        if cache is not None and cache[0] is tm and cache[1] == tm._version:
            return cache[2]
        cano2scene = torch.linalg.inv(tm.to(device=device, dtype=torch.float32))  # [B, 4, 4]
        cam_R = torch.as_tensor(self.dataset.cam_R, dtype=torch.float32, device=device).reshape(3, 3)
        cam_t = torch.as_tensor(self.dataset.cam_t, dtype=torch.float32, device=device).reshape(3)
        Rinv = torch.linalg.inv(cam_R)
        M = Rinv @ cano2scene[:, 0:3, 0:3]                                   # [B, 3, 3]
        m = (Rinv @ (cano2scene[:, 0:3, 3] - cam_t).unsqueeze(-1))           # [B, 3, 1]
        aff = torch.cat([M, m], dim=-1).contiguous()
        self._cam_cache = (tm, tm._version, aff)
        return aff

    def guide_2d_projection_with_smpl(self, batch, out, denoise_t, compute_grad='x_t'):
        """Gradient of -(2-D reprojection loss of 10 SMPL-X body joints against batch['keypoints_2d']) w.r.t. x_t or
        the predicted x_0 (reference posenet.py:260-317), [bs, body_feat_dim, 1, T], trajectory and contact channels
        zero.  One analytic forward + VJP kernel (rohm_projection_guidance) instead of autograd through the body model."""
        from .body_model import kernels_for
        x = batch['x_t'] if compute_grad == 'x_t' else out['pred_xstart']
        x = x.detach()
        x = x if (x.is_contiguous() and x.dtype == torch.float32) else x.contiguous().float()
        if self.dataset.traj_feat_dim != 22 or x.shape[1] != 294:
            raise RohmB200Error("guide_2d_projection_with_smpl: implemented for the 294-channel representation with the "
                                "22-channel trajectory block (the configuration RoHM ships)")
        B, _, _, T = x.shape
        dev = x.device
        mean, std = self._norm_stats(dev)
        f32 = lambda t: t.to(device=dev, dtype=torch.float32).contiguous()
        kp = f32(batch['keypoints_2d'])
        if kp.dim() != 4 or kp.shape[0] != B or kp.shape[1] < T or kp.shape[2] != 22 or kp.shape[3] != 3:
            raise RohmB200Error(f"guide_2d_projection_with_smpl: batch['keypoints_2d'] must be [{B}, >={T}, 22, 3], got "
                                f"{tuple(kp.shape)}")
        k = kernels_for(self.smplx_model, dev, B * T, with_vertices=False)
        return k.projection_guidance(x, mean, std, self._camera_affine(batch, dev), f32(batch['focal_length']),
                                     f32(batch['camera_center']), kp)

    def compute_losses_with_smpl(self, batch, model_output, smplx_model=None, epoch=0):
        """The evaluation loss dictionary of reference posenet.py:99-193 (what eval_losses returns with its default
        compute_loss=True, test_posenet.py:178); off the hot path, see rohm_b200/eval_losses.py."""
        from .eval_losses import posenet_losses
        return posenet_losses(self, batch, model_output, smplx_model, epoch)

    # ---------------------------------------------------------------- forward
    def forward(self, batch, timesteps):
        """batch['x_t'], batch['cond']: [bs, body_feat_dim, 1, T]; timesteps: [bs] int -> [bs, body_feat_dim, 1, T]
        (channels [0, traj_feat_dim) are batch['cond'][:, :traj_feat_dim], the rest is the denoised pose)."""
        x_t, cond = batch['x_t'], batch['cond']
        if x_t.dim() != 4 or x_t.shape[2] != 1 or x_t.shape != cond.shape or x_t.shape[1] != self.input_feats:
            raise RohmB200Error(f"PoseNet: expected x_t/cond of shape [B, {self.input_feats}, 1, T], got "
                                f"{tuple(x_t.shape)} / {tuple(cond.shape)}")
        if timesteps.is_floating_point():
            # the reference indexes pe[timesteps]: a float index raises there too (rescale_timesteps is never enabled)
            raise RohmB200Error("PoseNet: timesteps must be an integer tensor (they index the positional table)")
        e = self.prepare_cond(cond)
        x = x_t if (x_t.is_contiguous() and x_t.dtype == torch.float32) else x_t.contiguous().float()
        ts = timesteps.to(device=x.device, dtype=torch.int64).contiguous()
        return e.forward(x, ts)
