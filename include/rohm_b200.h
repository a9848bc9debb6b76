/* rohm_b200 -- C ABI of the B200-native RoHM hot path.
 *
 * The reference (sanweiliti/RoHM) is pure Python/PyTorch and has no FFI; its boundary is a set of Python symbols
 * (layer A).  This header is layer B: the plain-C entry points the Python drop-in
 * (rohm_b200/dropin/{model,diffusion,utils}) binds with ctypes.  Each entry point names the reference code it
 * replaces (file:line @ 57ba22c).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer to fp32 (or int64 where stated) unless marked "host";
 *   - `stream` is a cudaStream_t passed as void*; all calls are asynchronous on it;
 *   - no entry point allocates device memory except the *_create functions;
 *   - return value: 0 on success, negative rohm_status otherwise; rohm_last_error(ctx) gives the message;
 *   - one ctx per device/stream user; handles are not thread-safe.
 */
#ifndef ROHM_B200_H_
#define ROHM_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define ROHM_API __attribute__((visibility("default")))
#else
#define ROHM_API
#endif

typedef enum {
  ROHM_OK = 0,
  ROHM_ERR_INVALID = -1,   /* bad argument / unsupported shape */
  ROHM_ERR_CUDA = -2,      /* CUDA runtime or driver error */
  ROHM_ERR_NO_DEVICE = -3, /* no sm_100 device */
  ROHM_ERR_STATE = -4      /* call order violated (e.g. forward before set_cond) */
} rohm_status;

/* Arithmetic mode of the tensor-core GEMMs. */
typedef enum {
  ROHM_PRECISION_TF32X3 = 3, /* error-compensated TF32 hi/lo pairs: fp32-grade results (parity mode) */
  ROHM_PRECISION_F16X2 = 2,  /* error-compensated fp16 hi/lo pairs: the same 2 x 11 significant bits per value in half
                                the bytes (fp32-grade results, the default of every engine; activations must stay below 1.3e5) */
  ROHM_PRECISION_TF32 = 1    /* single-pass TF32: ~1e-3 relative (fast mode) */
} rohm_precision;

typedef struct rohm_ctx rohm_ctx;
typedef struct rohm_posenet rohm_posenet;
typedef struct rohm_trajnet rohm_trajnet;
typedef struct rohm_body rohm_body;

ROHM_API int rohm_version(void);
ROHM_API int rohm_ctx_create(int device, rohm_ctx** out);
ROHM_API void rohm_ctx_destroy(rohm_ctx* ctx);
ROHM_API const char* rohm_last_error(const rohm_ctx* ctx);

/* ------------------------------------------------------------------------------------------------------------
 * Sampler arithmetic (diffusion/gaussian_diffusion_posenet.py and _trajnet.py)
 * ---------------------------------------------------------------------------------------------------------- */

/* One ancestral update over n_clips clips of clip_elems contiguous elements each (all operands share the layout):
 *   mean = c1*x0 + c2*x_t                           q_posterior_mean_variance  :212-234, p_mean_variance :236-280
 *   mean += gs_k*grad_k        for k < n_grads      p_sample_with_grad         :461-477 (gs = weight*variance[t])
 *   out  = mean + sigma*noise                       p_sample :426-434          (sigma = (t!=0)*exp(0.5*logvar[t]))
 * coef points to device rows of ROHM_DDPM_COEFS floats {c1, c2, sigma, gs0, gs1, unused, unused, unused}; clip b reads
 * row coef + b*coef_clip_stride (stride 0 = one row for the whole batch, e.g. a row of a per-step table uploaded once).
 * Products and sums are rounded individually (no FMA contraction) so the result matches the reference's chain of
 * separate elementwise kernels bit for bit.  grads may be NULL when n_grads == 0.  `out` may alias x_t. */
#define ROHM_DDPM_COEFS 8
ROHM_API int rohm_ddpm_step(rohm_ctx* ctx, const float* x0, const float* x_t, const float* noise, const float* grad0,
                            const float* grad1, int n_grads, float* out, int64_t n_clips, int64_t clip_elems,
                            const float* coef, int64_t coef_clip_stride, void* stream);

/* rohm_ddpm_step with the noise = th.randn_like(x) (:426, :458) drawn INSIDE the kernel: Philox4_32_10(seed, offset) consumed
 * exactly as torch's CUDA normal_ kernel consumes it for a tensor of n_clips*clip_elems fp32 elements on this device, so
 * with (seed, offset) = the state of torch's CUDA generator the result is bit-identical to
 * rohm_ddpm_step(..., noise = torch.randn_like(x), ...).  *offset_increment (host, optional) receives the amount the caller
 * must advance the generator's offset by afterwards (what torch itself would have added). */
ROHM_API int rohm_ddpm_step_philox(rohm_ctx* ctx, const float* x0, const float* x_t, const float* grad0, const float* grad1,
                                   int n_grads, float* out, int64_t n_clips, int64_t clip_elems, const float* coef,
                                   int64_t coef_clip_stride, uint64_t seed, uint64_t offset, uint64_t* offset_increment,
                                   void* stream);

/* q_sample :192-210:  out = sqrt_ac*x_start + sqrt_1m_ac*noise. */
ROHM_API int rohm_q_sample(rohm_ctx* ctx, const float* x_start, const float* noise, float* out, int64_t n, float sqrt_ac,
                  float sqrt_one_minus_ac, void* stream);

/* DDIM update as INTENDED by ddim_sample :665-715 (unreachable in the reference):
 *   eps  = (sqrt_recip_ac*x_t - x0) / sqrt_recipm1_ac
 *   out  = x0*sqrt(ac_prev) + sqrt(1 - ac_prev - sigma^2)*eps + nonzero*sigma*noise
 * The four scalars are precomputed on the host from the float64 tables. */
ROHM_API int rohm_ddim_step(rohm_ctx* ctx, const float* x0, const float* x_t, const float* noise, float* out, int64_t n,
                   float sqrt_recip_ac, float sqrt_recipm1_ac, float sqrt_ac_prev, float dir_coef, float sigma,
                   void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * PoseNet denoiser (model/posenet.py:75-96, model/heads.py:112-176, nn.TransformerEncoder post-norm/gelu)
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct {
  const float* in_proj_w;  /* self_attn.in_proj_weight [3D, D] */
  const float* in_proj_b;  /* [3D] */
  const float* out_proj_w; /* self_attn.out_proj.weight [D, D] */
  const float* out_proj_b;
  const float* lin1_w; /* linear1.weight [F, D] */
  const float* lin1_b;
  const float* lin2_w; /* linear2.weight [D, F] */
  const float* lin2_b;
  const float* norm1_w;
  const float* norm1_b;
  const float* norm2_w;
  const float* norm2_b;
} rohm_posenet_layer;

typedef struct {
  int d_model;     /* latent_dim (512) */
  int ff_size;     /* 1024 */
  int num_layers;  /* 8 */
  int num_heads;   /* 4 */
  int in_feats;    /* body_feat_dim (294) */
  int out_feats;   /* pose_feat_dim (272) */
  int traj_feats;  /* channels copied from cond to the output (22) */
  int pe_len;      /* rows of the positional table (5000) */
  const float* in_w;   /* input_process.poseEmbedding.weight [D, in_feats] */
  const float* in_b;
  const float* cond_w; /* input_process_cond.poseEmbedding.weight [D, in_feats] */
  const float* cond_b;
  const float* pe;     /* sequence_pos_encoder.pe as [pe_len, D] */
  const float* t0_w;   /* embed_timestep.time_embed.0 [D, D] */
  const float* t0_b;
  const float* t2_w;   /* embed_timestep.time_embed.2 [D, D] */
  const float* t2_b;
  const float* out_w;  /* output_process.poseFinal.weight [out_feats, D] */
  const float* out_b;
  const rohm_posenet_layer* layers; /* host array of num_layers entries */
} rohm_posenet_weights;

/* Copies and repacks the weights (TF32 hi/lo split, K-padded) into library-owned device memory and allocates the
 * activation workspace for up to max_batch clips of max_frames frames. */
ROHM_API int rohm_posenet_create(rohm_ctx* ctx, const rohm_posenet_weights* w, int max_batch, int max_frames, int precision,
                        rohm_posenet** out);
ROHM_API void rohm_posenet_destroy(rohm_posenet* pn);

/* Step-invariant part of PoseNet.forward (posenet.py:86, 90-91): input_process_cond(cond) + positional rows.
 * cond: [B, in_feats, 1, T] contiguous.  Must be called whenever batch['cond'], B or T change. */
ROHM_API int rohm_posenet_set_cond(rohm_posenet* pn, const float* cond, int B, int T, void* stream);

/* PoseNet.forward (posenet.py:75-96).  x_t: [B, in_feats, 1, T]; timesteps: int64 [B] (original, un-respaced);
 * out: [B, in_feats, 1, T] with channels [0, traj_feats) copied from the cond given to set_cond. */
ROHM_API int rohm_posenet_forward(rohm_posenet* pn, const float* x_t, const int64_t* timesteps, float* out, int B, int T,
                         void* stream);

/* One whole ancestral step of the PoseNet sampler (p_mean_variance :236-280 + p_sample :388-434) as ONE graph launch:
 * x0_out = PoseNet(x_t, timesteps) followed by x_next = c1 x0 + c2 x_t + sigma N(0, I) with the noise drawn in the update
 * kernel exactly as torch.randn_like(x_t) would draw it from (seed, offset) (see rohm_ddpm_step_philox).  coef_row: device
 * row {c1, c2, sigma, ...} of this step (shared by all clips).  *offset_increment: what to advance the generator by. */
ROHM_API int rohm_posenet_sample_step(rohm_posenet* pn, const float* x_t, const int64_t* timesteps, float* x0_out,
                                      float* x_next, const float* coef_row, uint64_t seed, uint64_t offset,
                                      uint64_t* offset_increment, int B, int T, void* stream);

/* Same as rohm_posenet_forward but with CUDA events recorded on `stream` around every kernel launch; synchronises the
 * stream and returns, per category {0: tensor-core GEMM, 1: attention, 2: LayerNorm, 3: pack/unpack/time-token}, the
 * summed device milliseconds (host float[4]) and the number of launches (host int[4]).  For bench.py's roofline. */
ROHM_API int rohm_posenet_profile(rohm_posenet* pn, const float* x_t, const int64_t* timesteps, float* out, int B, int T,
                         void* stream, float* ms_by_category, int* launches_by_category);

/* Options: 0 = replay the forward as a CUDA graph (default 1; the graph is captured on first use per (B, T) and its
 * three caller-memory pointers are patched per call); 1 = programmatic dependent launch on the GEMMs (default 1). */
ROHM_API int rohm_posenet_set_option(rohm_posenet* pn, int option, int value);

/* Kernel launches issued by the last forward (for bench.py's gpu_launches accounting). */
ROHM_API int rohm_posenet_launches_per_forward(const rohm_posenet* pn);

/* ------------------------------------------------------------------------------------------------------------
 * TrajNet denoiser + TrajControl branch (model/trajnet.py:10-75, 177-275; blocks model/heads.py:20-106)
 * ---------------------------------------------------------------------------------------------------------- */

/* Parameters are handed over by their reference state-dict key ("diff_enc1.blocks.0.block.0.weight",
 * "controlnet.control_zero_conv_0.bias", "time_mlp.1.weight", ...): `names[i]` (host strings), `ptrs[i]` (device fp32),
 * `numels[i]`.  Every key the architecture needs must be present with the reference's shape; extra keys (e.g. the
 * never-evaluated cond_downsample4.*) are ignored.  `frames` must be a multiple of 16.  The library repacks all
 * convolution weights into GEMM layout (TF32 hi/lo, tap-major K) and owns its copies. */
ROHM_API int rohm_trajnet_create(rohm_ctx* ctx, int n_params, const char* const* names, const float* const* ptrs,
                                 const int64_t* numels, int time_dim, int cond_dim, int traj_feat_dim, int mid_dim,
                                 int trajcontrol, int control_cond_dim, int max_batch, int frames, int precision,
                                 rohm_trajnet** out);
ROHM_API void rohm_trajnet_destroy(rohm_trajnet* tn);

/* Step-invariant part of TrajNet.forward: the condition pyramid cond_enc1..4 (trajnet.py:192-208) and, with
 * TrajControl, control_zero_conv_0(control_cond) (:51-52).  cond: [B, frames, cond_dim]; control_cond:
 * [B, frames, control_cond_dim] or NULL for the vanilla network.  Call whenever batch['cond'] / ['control_cond'] change. */
ROHM_API int rohm_trajnet_set_cond(rohm_trajnet* tn, const float* cond, const float* control_cond, int B, void* stream);

/* TrajNet.forward (trajnet.py:177-275).  x_t: [B, frames, traj_feat_dim]; time: int64 [B]; out: same shape as x_t. */
ROHM_API int rohm_trajnet_forward(rohm_trajnet* tn, const float* x_t, const int64_t* time, float* out, int B,
                                  void* stream);
/* One whole ancestral step of the TrajNet sampler (gaussian_diffusion_trajnet.py:388-434, p_sample without cond_fn):
 * x0_out = TrajNet.forward(x_t, time), x_next = coef1 x0 + coef2 x_t + sigma z with z drawn inside the update kernel exactly as
 * torch.randn_like(x_t) would draw it from (seed, offset) -- forward and update are ONE graph launch.  Arguments as
 * rohm_posenet_sample_step (coef_row: device float[8] of the step). */
ROHM_API int rohm_trajnet_sample_step(rohm_trajnet* tn, const float* x_t, const int64_t* time, float* x0_out, float* x_next,
                                      const float* coef_row, uint64_t seed, uint64_t offset, uint64_t* offset_increment, int B,
                                      void* stream);
ROHM_API int rohm_trajnet_set_option(rohm_trajnet* tn, int option, int value); /* 0: CUDA-graph replay, 1: programmatic dependent launch (both default 1) */
ROHM_API int rohm_trajnet_launches_per_forward(const rohm_trajnet* tn);

/* ------------------------------------------------------------------------------------------------------------
 * SMPL-X body model: joints FK, skating guidance, full LBS
 * (third-party smplx==0.1.28 lbs.py / body_models.py as called from data_loaders/motion_representation.py:373-398,
 *  model/posenet.py:196-257; see DESIGN.md for provenance -- the body model is not part of the reference tree)
 * ---------------------------------------------------------------------------------------------------------- */

/* v_template [V,3], shapedirs [V,3,shape_comps] (first 10 components = betas), posedirs [486, V*3],
 * J_regressor [55,V], lbs_weights [V,55] (device fp32); parents_host: host int[55] (-1 for the root).
 * max_frames: capacity in frames (B*T) of the per-call workspace.  with_vertices = 0 skips the LBS data (joints and
 * guidance only; posedirs / lbs_weights may then be NULL). */
ROHM_API int rohm_body_create(rohm_ctx* ctx, const float* v_template, const float* shapedirs, int shape_comps,
                              const float* posedirs, const float* J_regressor, const float* lbs_weights,
                              const int* parents_host, int num_verts, int64_t max_frames, int with_vertices,
                              int precision, rohm_body** out);
ROHM_API void rohm_body_destroy(rohm_body* bd);
/* 1 if rohm_body_forward computes the vertices with the fused blend-GEMM + skinning launch, 0 if the handle uses the
 * two-kernel path (no vertex support, TF32 precision, ROHM_B200_FUSED_LBS=0, or a body model whose 32-vertex tiles touch
 * more than 16 bones).  Introspection only. */
ROHM_API int rohm_body_uses_fused_lbs(const rohm_body* bd);

/* Row pitch, in floats, of the `vertices` buffers handed to rohm_body_forward / rohm_body_from_repr[_layout] from now on:
 * frame n's vertices start at vertices + n * pitch.  0 (the default) = dense [N, V, 3] as smplx returns them
 * (reference: body_model output `.vertices`, test_amass_full.py:392-428).  A pitch that is a multiple of 4 floats (>= 3 V;
 * 16-byte-aligned buffer) lets the fused launch write the vertices with TMA bulk stores instead of 4-byte stores (3 V =
 * 31425 floats is not 16-byte divisible); only valid when rohm_body_uses_fused_lbs() is 1. */
ROHM_API int rohm_body_set_vertex_pitch(rohm_body* bd, int64_t pitch_floats);

/* SMPLX.forward with jaw / eyes / hands / expression = 0 (exactly how RoHM calls it): global_orient [N,3], body_pose
 * [N,63] axis-angle, betas [N,10], transl [N,3] -> joints [N, num_joints, 3] (first num_joints <= 55 posed joints +
 * transl; NULL to skip) and vertices [N, V, 3] (NULL to skip). */
ROHM_API int rohm_body_forward(rohm_body* bd, const float* global_orient, const float* body_pose, const float* betas,
                               const float* transl, int64_t N, float* joints, int num_joints, float* vertices,
                               void* stream);

/* recover_from_repr_smpl(recover_mode='smplx_params') on a normalised representation x [B,294,1,T] with the dataset's
 * mean/std [294]: denormalise, rot6d -> rotmat -> axis-angle (kornia route), then rohm_body_forward. */
ROHM_API int rohm_body_from_repr(rohm_body* bd, const float* x, const float* mean, const float* stdv, int B, int T,
                                 float* joints, int num_joints, float* vertices, void* stream);

/* The same with an explicit layout: channels_last = 0 -> x is [B,294,1,T] (PoseNet tensors); 1 -> x is [B,T,294] (the
 * drivers' tensors, test_amass_full.py:279-293, 386-428). */
ROHM_API int rohm_body_from_repr_layout(rohm_body* bd, const float* x, int channels_last, const float* mean,
                                        const float* stdv, int B, int T, float* joints, int num_joints, float* vertices,
                                        void* stream);

/* PoseNet.guide_skating_with_smpl(compute_grad='x_0'): grad [B,294,1,T] = d(-(loss_smpl + loss_abs))/d x0 with the
 * channels [0,22) and the 4 contact channels zero.  Analytic VJP (no autograd); all-zero if nothing skates.
 * loss_out: optional device float[4] = {sum_abs, count_abs, sum_smpl, count_smpl}. */
ROHM_API int rohm_skating_guidance(rohm_body* bd, const float* x0, const float* mean, const float* stdv, int B, int T,
                                   float* grad, float* loss_out, void* stream);

/* rohm_skating_guidance in two halves, for clip-sharded runs that reproduce the reference's BATCH-GLOBAL normalisers
 * (posenet.py:230-233, 242-248): _sums computes this shard's {sum_abs, count_abs, sum_smpl, count_smpl} into sums_out (device
 * float[4]; per-frame state stays in the handle), the caller all-reduces the four floats over the ranks, _backward produces
 * this shard's gradient with the global sums.  _sums followed by _backward with the same sums == rohm_skating_guidance. */
ROHM_API int rohm_skating_guidance_sums(rohm_body* bd, const float* x0, const float* mean, const float* stdv, int B, int T,
                                        float* sums_out, void* stream);
ROHM_API int rohm_skating_guidance_backward(rohm_body* bd, const float* x0, const float* mean, const float* stdv, int B, int T,
                                            const float* sums, float* grad, void* stream);

/* PoseNet.guide_2d_projection_with_smpl(compute_grad='x_0') (model/posenet.py:260-317, utils/other_utils.py:150-185):
 * grad [B,294,1,T] = d(-loss_2d)/d x0, loss_2d = mean over (clip, frame, 10 selected joints, 2) of
 * |perspective_projection(camera <- scene <- canonical joints) - keypoints| * confidence; channels [0,22) and the 4 contact
 * channels zero.  cam_affine [B,12]: rows of the 3x4 map canonical -> camera coordinates per clip
 * (inv(cam_R) (inv(transf_matrix) p - cam_t)); focal, center [B,2]; keypoints_2d [B, kp_frames >= T, 22, 3] = (u, v, conf).
 * Analytic VJP through the 22-joint kinematic tree (no autograd).  loss_out: optional device float = the un-normalised sum. */
ROHM_API int rohm_projection_guidance(rohm_body* bd, const float* x0, const float* mean, const float* stdv, int B, int T,
                                      const float* cam_affine, const float* focal, const float* center,
                                      const float* keypoints_2d, int kp_frames, float* grad, float* loss_out,
                                      void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Either side of the sampling loops: the drivers' inter-round glue and representation recovery, on the device
 * ---------------------------------------------------------------------------------------------------------- */

/* test_amass_full.py:268-311 (per-clip host loop in the reference): TrajNet output traj_out [B,T,traj_dim] (z-scored,
 * traj_dim = 13 for repr_abs_only else <= 22) is scattered into repr_clean [B,T,294] -> composite_out [B,T,294]
 * (z-scored with the trajectory dataset's mean/std) -> SMPL-X joints (rot6d -> axis-angle -> FK) -> get_repr_smplx
 * (data_loaders/motion_representation.py:187-282: forward direction, root quaternion incl. the first-NaN repair, velocities,
 * global-orient 6-D / angular velocity, translation) -> traj_full_out [B,T-1,22], z-scored with the pose dataset's stats. */
ROHM_API int rohm_traj_glue(rohm_body* bd, const float* traj_out, int traj_dim, const float* repr_clean,
                            const float* traj_mean, const float* traj_std, const float* pose_mean, const float* pose_std,
                            int B, int T, float* composite_out, float* traj_full_out, void* stream);

/* The last stage of rohm_traj_glue on its own: get_repr_smplx's trajectory block (motion_representation.py:187-282) from
 * joints [B,T,22,3], global-orient axis-angles [B*T,3] and translations [B*T,3] -> traj_full_out [B,T-1,22], z-scored with
 * mean/stdv (the first 22 entries are read). */
ROHM_API int rohm_traj_repr_from_joints(rohm_ctx* ctx, const float* joints, const float* global_orient_aa,
                                        const float* transl, const float* mean, const float* stdv, int B, int T,
                                        float* traj_full_out, void* stream);

/* test_amass_full.py:256-258: control_cond [B,T,cond_feats] from the PoseNet output pose_out [B,traj_feats+cond_feats,1,Tp]
 * (frames [0,Tp) copied, frames [Tp,T) repeat frame Tp-1). */
ROHM_API int rohm_pose_to_control_cond(rohm_ctx* ctx, const float* pose_out, int B, int Tp, int T, int traj_feats,
                                       int cond_feats, float* control_cond, void* stream);

/* test_amass_full.py:320-370: PoseNet condition cond_out [B,294,1,Tp] = src (channel-major [B,294,1,src_T] or channels-last
 * [B,src_T,294]) with channels [0,22) replaced by traj_full [B,Tp,22] (NULL keeps src) and channels >= 22 zeroed where
 * chan_keep[c] == 0 (294 bytes, NULL = keep all: the 'lower' / 'upper' joint masks), where frame_lo[b] <= t < frame_hi[b]
 * (int [B], NULL = none: the 'full' scheme) and, with zero_contact, in the 4 contact channels. */
ROHM_API int rohm_build_pose_cond(rohm_ctx* ctx, const float* src, int src_channel_major, int src_T, const float* traj_full,
                                  const unsigned char* chan_keep, const int* frame_lo, const int* frame_hi,
                                  int zero_contact, int B, int Tp, float* cond_out, void* stream);

/* rot6d_to_rotmat (quaternion.py:482-501) and rotation_matrix_to_angle_axis (konia_transform.py:317-340 -> :350-444 ->
 * :561-631) on n 6-D rotations: aa [n,3] and/or rotmat [n,9] (row-major), either may be NULL. */
ROHM_API int rohm_rot6d_to_aa(rohm_ctx* ctx, const float* rot6d, int64_t n, float* aa, float* rotmat, void* stream);

/* recover_from_repr_smpl(recover_mode='joint_abs_traj' | 'joint_rel_traj') (motion_representation.py:285-371) on a z-scored
 * representation in either layout -> joints [B,T,22,3]. */
ROHM_API int rohm_joints_from_traj(rohm_ctx* ctx, const float* x, int channels_last, const float* mean, const float* stdv,
                                   int B, int T, int relative, float* joints, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ROHM_B200_H_ */
