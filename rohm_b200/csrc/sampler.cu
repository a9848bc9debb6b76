// Sampler arithmetic: fused DDPM posterior step, q_sample, DDIM step.  HBM-bound elementwise kernels:
// 128-bit vectorised, grid sized as a multiple of the SM count, every product/sum individually rounded so the
// result is bit-identical to the reference's chain of separate elementwise ops
// (diffusion/gaussian_diffusion_posenet.py:212-234, 426-434, 461-479, 192-210, 696-715).
#include <curand_kernel.h>

#include "common.h"

namespace rohm {
namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ float ddpm_one(float x0, float xt, float nz, float g0, float g1, int n_grads, float c1,
                                          float c2, float sigma, float gs0, float gs1) {
  float mean = __fadd_rn(__fmul_rn(c1, x0), __fmul_rn(c2, xt));
  if (n_grads > 0) mean = __fadd_rn(mean, __fmul_rn(gs0, g0));
  if (n_grads > 1) mean = __fadd_rn(mean, __fmul_rn(gs1, g1));
  return __fadd_rn(mean, __fmul_rn(sigma, nz));
}

// grid = (chunks per clip, clips); each clip reads its own coefficient row.
template <bool kVec>
__global__ void __launch_bounds__(kThreads) ddpm_step_kernel(const float* __restrict__ x0, const float* x_t,
                                                             const float* __restrict__ noise,
                                                             const float* __restrict__ g0,
                                                             const float* __restrict__ g1, int n_grads, float* out,
                                                             int64_t clip_elems, const float* __restrict__ coef,
                                                             int64_t coef_stride) {
  const int64_t clip = blockIdx.y;
  const float* cf = coef + clip * coef_stride;
  const float c1 = cf[0], c2 = cf[1], sigma = cf[2], gs0 = cf[3], gs1 = cf[4];
  const int64_t base = clip * clip_elems;
  const int64_t tid = static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x;
  const int64_t nthreads = static_cast<int64_t>(gridDim.x) * kThreads;
  if (kVec) {
    const int64_t n4 = clip_elems >> 2;  // host guarantees clip_elems % 4 == 0 and 16-byte aligned bases
    const float4* a4 = reinterpret_cast<const float4*>(x0 + base);
    const float4* b4 = reinterpret_cast<const float4*>(x_t + base);
    const float4* z4 = reinterpret_cast<const float4*>(noise + base);
    const float4* p4 = reinterpret_cast<const float4*>(n_grads > 0 ? g0 + base : x0 + base);
    const float4* q4 = reinterpret_cast<const float4*>(n_grads > 1 ? g1 + base : x0 + base);
    float4* o4 = reinterpret_cast<float4*>(out + base);
    for (int64_t i = tid; i < n4; i += nthreads) {
      const float4 a = a4[i], b = b4[i], z = z4[i];
      float4 p = make_float4(0.f, 0.f, 0.f, 0.f), q = p;
      if (n_grads > 0) p = p4[i];
      if (n_grads > 1) q = q4[i];
      float4 r;
      r.x = ddpm_one(a.x, b.x, z.x, p.x, q.x, n_grads, c1, c2, sigma, gs0, gs1);
      r.y = ddpm_one(a.y, b.y, z.y, p.y, q.y, n_grads, c1, c2, sigma, gs0, gs1);
      r.z = ddpm_one(a.z, b.z, z.z, p.z, q.z, n_grads, c1, c2, sigma, gs0, gs1);
      r.w = ddpm_one(a.w, b.w, z.w, p.w, q.w, n_grads, c1, c2, sigma, gs0, gs1);
      o4[i] = r;
    }
  } else {
    for (int64_t i = tid; i < clip_elems; i += nthreads) {
      const int64_t k = base + i;
      out[k] = ddpm_one(x0[k], x_t[k], noise[k], n_grads > 0 ? g0[k] : 0.f, n_grads > 1 ? g1[k] : 0.f, n_grads, c1, c2,
                        sigma, gs0, gs1);
    }
  }
}

__global__ void __launch_bounds__(kThreads) q_sample_kernel(const float* __restrict__ xs,
                                                            const float* __restrict__ noise, float* out, int64_t n,
                                                            float a, float b) {
  const int64_t tid = static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x;
  const int64_t nthreads = static_cast<int64_t>(gridDim.x) * kThreads;
  for (int64_t i = tid; i < n; i += nthreads) out[i] = __fadd_rn(__fmul_rn(a, xs[i]), __fmul_rn(b, noise[i]));
}

__global__ void __launch_bounds__(kThreads) ddim_step_kernel(const float* __restrict__ x0, const float* x_t,
                                                             const float* __restrict__ noise, float* out, int64_t n,
                                                             float sr, float srm1, float sap, float dir, float sigma) {
  const int64_t tid = static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x;
  const int64_t nthreads = static_cast<int64_t>(gridDim.x) * kThreads;
  for (int64_t i = tid; i < n; i += nthreads) {
    const float eps = __fdiv_rn(__fsub_rn(__fmul_rn(sr, x_t[i]), x0[i]), srm1);
    const float mean = __fadd_rn(__fmul_rn(x0[i], sap), __fmul_rn(dir, eps));
    out[i] = __fadd_rn(mean, __fmul_rn(sigma, noise[i]));
  }
}

// The same update with the Gaussian noise drawn inside the kernel, bit-identical to what `torch.randn_like(x)` would have
// produced from the same generator state (ATen distribution_elementwise_grid_stride_kernel with curand_normal4, unroll 4):
// virtual thread `vidx` of torch's grid (G = grid * 256 threads) owns elements vidx + G * j; its j-th normal is component
// j % 4 of its (j / 4)-th curand_normal4 call on Philox4_32_10(seed, subsequence = vidx, offset).  One real thread per
// virtual thread, so the launch reads x0 / x_t and writes x_{t-1} once and the noise never touches memory.
__global__ void __launch_bounds__(kThreads) ddpm_step_philox_kernel(const float* __restrict__ x0, const float* x_t,
                                                                    const float* __restrict__ g0,
                                                                    const float* __restrict__ g1, int n_grads, float* out,
                                                                    int64_t numel, int64_t clip_elems,
                                                                    const float* __restrict__ coef, int64_t coef_stride,
                                                                    unsigned long long seed, unsigned long long offset,
                                                                    int64_t G, int iters) {
  // as the programmatic dependent of the denoiser's last kernel (posenet.cu): the Philox state is set up while that kernel
  // drains; x0 / x_t are read after it has completed.  A no-op for a plain launch.
  const int64_t vidx = static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x;
  curandStatePhilox4_32_10_t state;
  if (vidx < G) curand_init(seed, static_cast<unsigned long long>(vidx), offset, &state);
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if (vidx >= G) return;
  for (int k = 0; k < iters; ++k) {
    const float4 nz = curand_normal4(&state);
    const float z[4] = {nz.x, nz.y, nz.z, nz.w};
#pragma unroll
    for (int ii = 0; ii < 4; ++ii) {
      const int64_t li = vidx + G * (4 * k + ii);
      if (li < numel) {
        const int64_t clip = li / clip_elems;
        const float* cf = coef + clip * coef_stride;
        out[li] = ddpm_one(x0[li], x_t[li], z[ii], n_grads > 0 ? g0[li] : 0.f, n_grads > 1 ? g1[li] : 0.f, n_grads, cf[0],
                           cf[1], cf[2], cf[3], cf[4]);
      }
    }
  }
}

int grid_for(const rohm_ctx* ctx, int64_t work_items) {
  const int sms = ctx->sm_count > 0 ? ctx->sm_count : 148;
  int64_t blocks = (work_items + kThreads - 1) / kThreads;
  const int64_t cap = static_cast<int64_t>(sms) * 8;  // 8 resident CTAs of 256 threads per SM
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return static_cast<int>(blocks);
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace
}  // namespace rohm

using namespace rohm;

extern "C" int rohm_ddpm_step(rohm_ctx* ctx, const float* x0, const float* x_t, const float* noise, const float* grad0,
                              const float* grad1, int n_grads, float* out, int64_t n_clips, int64_t clip_elems,
                              const float* coef, int64_t coef_clip_stride, void* stream) {
  if (ctx == nullptr) return ROHM_ERR_INVALID;
  rohm::DeviceGuard device_guard__(ctx);
  if (n_clips < 0 || clip_elems < 0 || x0 == nullptr || x_t == nullptr || noise == nullptr || out == nullptr ||
      coef == nullptr || n_grads < 0 || n_grads > 2 || (n_grads > 0 && grad0 == nullptr) ||
      (n_grads > 1 && grad1 == nullptr) || n_clips > 65535)
    return fail(ctx, ROHM_ERR_INVALID, "rohm_ddpm_step: bad arguments");
  if (n_clips == 0 || clip_elems == 0) return ROHM_OK;
  const bool vec = (clip_elems % 4 == 0) && aligned16(x0) && aligned16(x_t) && aligned16(noise) && aligned16(out) &&
                   (n_grads < 1 || aligned16(grad0)) && (n_grads < 2 || aligned16(grad1));
  const int sms = ctx->sm_count > 0 ? ctx->sm_count : 148;
  const int64_t items = vec ? clip_elems / 4 : clip_elems;
  int64_t bx = (items + kThreads - 1) / kThreads;
  const int64_t cap = (static_cast<int64_t>(sms) * 8 + n_clips - 1) / n_clips;  // ~8 resident CTAs per SM in total
  if (bx > cap) bx = cap;
  if (bx < 1) bx = 1;
  dim3 grid(static_cast<unsigned>(bx), static_cast<unsigned>(n_clips));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (vec)
    ddpm_step_kernel<true><<<grid, kThreads, 0, st>>>(x0, x_t, noise, grad0, grad1, n_grads, out, clip_elems, coef,
                                                      coef_clip_stride);
  else
    ddpm_step_kernel<false><<<grid, kThreads, 0, st>>>(x0, x_t, noise, grad0, grad1, n_grads, out, clip_elems, coef,
                                                       coef_clip_stride);
  ROHM_CUDA(ctx, cudaGetLastError());
  return ROHM_OK;
}

// Launch geometry of torch's normal_ kernel for `numel` fp32 elements on this device (ATen calc_execution_policy):
// G = virtual threads, iters = curand_normal4 calls per thread, *increment = what the generator's offset advances by.
static int torch_normal_policy(rohm_ctx* ctx, int64_t numel, int64_t* G, int* iters, unsigned long long* increment) {
  int threads_per_sm = 0;
  ROHM_CUDA(ctx, cudaDeviceGetAttribute(&threads_per_sm, cudaDevAttrMaxThreadsPerMultiProcessor, ctx->device));
  const int64_t blocks_per_sm = threads_per_sm / 256;
  int64_t grid = (numel + 255) / 256;
  const int64_t cap = static_cast<int64_t>(ctx->sm_count) * blocks_per_sm;
  if (grid > cap) grid = cap;
  *G = grid * 256;
  *iters = static_cast<int>((numel - 1) / (*G * 4) + 1);
  *increment = static_cast<unsigned long long>(*iters) * 4ull;
  return ROHM_OK;
}

extern "C" int rohm_ddpm_step_philox(rohm_ctx* ctx, const float* x0, const float* x_t, const float* grad0, const float* grad1,
                                     int n_grads, float* out, int64_t n_clips, int64_t clip_elems, const float* coef,
                                     int64_t coef_clip_stride, uint64_t seed, uint64_t offset, uint64_t* offset_increment,
                                     void* stream) {
  if (ctx == nullptr) return ROHM_ERR_INVALID;
  rohm::DeviceGuard device_guard__(ctx);
  if (n_clips < 0 || clip_elems < 0 || x0 == nullptr || x_t == nullptr || out == nullptr || coef == nullptr || n_grads < 0 ||
      n_grads > 2 || (n_grads > 0 && grad0 == nullptr) || (n_grads > 1 && grad1 == nullptr))
    return fail(ctx, ROHM_ERR_INVALID, "rohm_ddpm_step_philox: bad arguments");
  const int64_t numel = n_clips * clip_elems;
  if (offset_increment != nullptr) *offset_increment = 0;
  if (numel == 0) return ROHM_OK;
  int64_t G = 0;
  int iters = 0;
  unsigned long long inc = 0;
  int rc = torch_normal_policy(ctx, numel, &G, &iters, &inc);
  if (rc != ROHM_OK) return rc;
  if (offset_increment != nullptr) *offset_increment = inc;
  ddpm_step_philox_kernel<<<static_cast<unsigned>(G / kThreads), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      x0, x_t, grad0, grad1, n_grads, out, numel, clip_elems, coef, coef_clip_stride, seed, offset, G, iters);
  ROHM_CUDA(ctx, cudaGetLastError());
  return ROHM_OK;
}

// For the denoiser engines that append the update to their forward graph (posenet.cu): the kernel's address (to find its
// graph node) and a launch with explicit geometry.
namespace rohm {
const void* ddpm_step_philox_kernel_address() { return reinterpret_cast<const void*>(ddpm_step_philox_kernel); }
int ddpm_step_philox_policy(rohm_ctx* ctx, int64_t numel, int64_t* G, int* iters, unsigned long long* increment) {
  return torch_normal_policy(ctx, numel, G, iters, increment);
}
cudaError_t launch_ddpm_step_philox(const float* x0, const float* x_t, float* out, int64_t numel, int64_t clip_elems,
                                    const float* coef, unsigned long long seed, unsigned long long offset, int64_t G, int iters,
                                    cudaStream_t st, bool pdl) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(static_cast<unsigned>(G / kThreads)), cfg.blockDim = dim3(kThreads), cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr, cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, ddpm_step_philox_kernel, x0, x_t, static_cast<const float*>(nullptr),
                            static_cast<const float*>(nullptr), 0, out, numel, clip_elems, coef, static_cast<int64_t>(0), seed,
                            offset, G, iters);
}
}  // namespace rohm

extern "C" int rohm_q_sample(rohm_ctx* ctx, const float* x_start, const float* noise, float* out, int64_t n,
                             float sqrt_ac, float sqrt_one_minus_ac, void* stream) {
  if (ctx == nullptr) return ROHM_ERR_INVALID;
  rohm::DeviceGuard device_guard__(ctx);
  if (n < 0 || x_start == nullptr || noise == nullptr || out == nullptr)
    return fail(ctx, ROHM_ERR_INVALID, "rohm_q_sample: bad arguments");
  if (n == 0) return ROHM_OK;
  q_sample_kernel<<<grid_for(ctx, n), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(x_start, noise, out, n, sqrt_ac,
                                                                                       sqrt_one_minus_ac);
  ROHM_CUDA(ctx, cudaGetLastError());
  return ROHM_OK;
}

extern "C" int rohm_ddim_step(rohm_ctx* ctx, const float* x0, const float* x_t, const float* noise, float* out,
                              int64_t n, float sqrt_recip_ac, float sqrt_recipm1_ac, float sqrt_ac_prev, float dir_coef,
                              float sigma, void* stream) {
  if (ctx == nullptr) return ROHM_ERR_INVALID;
  rohm::DeviceGuard device_guard__(ctx);
  if (n < 0 || x0 == nullptr || x_t == nullptr || noise == nullptr || out == nullptr)
    return fail(ctx, ROHM_ERR_INVALID, "rohm_ddim_step: bad arguments");
  if (n == 0) return ROHM_OK;
  ddim_step_kernel<<<grid_for(ctx, n), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      x0, x_t, noise, out, n, sqrt_recip_ac, sqrt_recipm1_ac, sqrt_ac_prev, dir_coef, sigma);
  ROHM_CUDA(ctx, cudaGetLastError());
  return ROHM_OK;
}
