// Internal layout of a rohm_body handle, shared by the translation units that run kernels against the body model
// (body.cu: FK / LBS / skating guidance; glue.cu: inter-round glue, 2-D reprojection guidance, representation recovery).
#pragma once
#include "common.h"
#include "gemm.cuh"

struct rohm_body {
  rohm_ctx* ctx = nullptr;
  rohm::DevicePool pool;
  int V = 0, sd_comps = 0, passes = 3;
  int kind = rohm::kKindTf32;  // operand element type of the blend GEMM (kKindF16 in ROHM_PRECISION_F16X2)
  int64_t max_frames = 0;
  float *Jt = nullptr, *Jd = nullptr;
  const float* lbs_w = nullptr;  // dense weights copy
  float* lbs_w_copy = nullptr;
  int* bone_idx = nullptr;
  float* bone_w = nullptr;
  bool sparse_ok = true;
  rohm::PackedWeight blend;  // [V*3 (padded), kBlendK]
  // per-frame workspace
  float *go = nullptr, *bp = nullptr, *betas = nullptr, *transl = nullptr, *A = nullptr, *feat_h = nullptr,
        *feat_l = nullptr, *vposed = nullptr;
  float *foot = nullptr, *gdir = nullptr, *sums = nullptr;
  int* parents_dev = nullptr;      // [55] kinematic tree on the device (glue.cu kernels; body.cu keeps a __constant__ copy)
  float* jwork = nullptr;          // [max_frames, 22, 3] scratch joints (glue)
  float* gwork = nullptr;          // [max_frames, 22, 3] scratch joint gradients (2-D guidance)
  rohm::GemmParams g_blend{};
};
