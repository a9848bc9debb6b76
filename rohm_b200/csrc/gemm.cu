// tcgen05 / TMA implementation of the segmented-A 3xTF32 GEMM declared in gemm.cuh.
#include "gemm.cuh"
#include "ptx.cuh"

#include <cudaTypedefs.h>
#include <cstdio>
#include <mutex>

namespace rohm {

namespace {

constexpr int kEpiWarps = 8;
constexpr int kThreads = 32 * (2 + kEpiWarps);
constexpr int kSmemBudget = 200 * 1024;  // ring buffer budget; + 1 KB alignment slack stays below 227 KB

template <int BLOCK_N, int PASSES>
struct TileCfg {
  static constexpr int kABytes = kGemmBlockM * kGemmBlockK * 4;  // 16 KB
  static constexpr int kBBytes = BLOCK_N * kGemmBlockK * 4;
  static constexpr int kSplit = (PASSES == 3) ? 2 : 1;
  static constexpr int kStageBytes = kSplit * (kABytes + kBBytes);
  static constexpr int kStagesRaw = kSmemBudget / kStageBytes;
  static constexpr int kStages = kStagesRaw > 8 ? 8 : kStagesRaw;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024;
  // PASSES == 3 keeps two accumulators per stage: columns [0, BLOCK_N) take the leading hi*hi products, columns
  // [BLOCK_N, 2*BLOCK_N) the two small cross terms.  The tensor core truncates when it adds into the
  // accumulator, so keeping the ~2^-11-sized terms out of the big sum cuts the rounding count of the main
  // accumulator by 3x (measured: error grows linearly with the number of accumulating MMAs).
  static constexpr int kAccCols = (PASSES == 3 ? 2 : 1) * BLOCK_N;
  // two accumulator stages: the epilogue drains tile i while the MMA warp already accumulates tile i+1
  static constexpr int kAccStages = 2;
  static constexpr int kTmemNeed = kAccStages * kAccCols;
  static constexpr uint32_t kTmemCols = kTmemNeed <= 32 ? 32 : kTmemNeed <= 64 ? 64 : kTmemNeed <= 128 ? 128 : kTmemNeed <= 256 ? 256 : 512;
  static_assert(kTmemNeed <= 512, "accumulators exceed TMEM");
  static_assert(kStages >= 2, "need at least a double buffer");
  static_assert(BLOCK_N % 16 == 0 && BLOCK_N >= 16 && BLOCK_N <= 256, "UMMA N constraint for M=128");
  static_assert(BLOCK_N % 32 == 0, "epilogue walks TMEM in 32-column chunks");
};

__device__ __forceinline__ float apply_act(float x, int act) {
  if (act == kActGelu) {
    return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));  // exact (erf) GELU, nn.GELU default
  } else if (act == kActSilu) {
    return x / (1.0f + expf(-x));
  } else if (act == kActMish) {
    // x * tanh(softplus(x)); softplus with torch's threshold=20
    float sp = x > 20.0f ? x : log1pf(expf(x));
    return x * tanhf(sp);
  }
  return x;
}

// Persistent, warp-specialised: grid = min(#tiles, #SMs); every CTA walks tiles blockIdx.x, +gridDim.x, ...
// (N-tile index fastest, so CTAs running concurrently share the same A rows in L2).
template <int BLOCK_N, int PASSES>
__global__ void __launch_bounds__(kThreads, 1) gemm_tile_kernel(const __grid_constant__ GemmParams p) {
  using Cfg = TileCfg<BLOCK_N, PASSES>;
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t full_bar[Cfg::kStages];
  __shared__ uint64_t empty_bar[Cfg::kStages];
  __shared__ uint64_t tmem_full_bar[Cfg::kAccStages];
  __shared__ uint64_t tmem_empty_bar[Cfg::kAccStages];
  __shared__ uint32_t tmem_base_smem;

  // SWIZZLE_128B tiles need 1024-byte alignment.
  const uint32_t raw_addr = ptx::smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int tiles_n = (p.grid_n_cols + BLOCK_N - 1) / BLOCK_N;
  const int tiles_m = (p.grid_m_rows + kGemmBlockM - 1) / kGemmBlockM;
  const int num_tiles = tiles_m * tiles_n;

  int total_iters = 0;
  for (int s = 0; s < p.num_segs; ++s) total_iters += p.seg_kblocks[s];

  if (warp_idx == 0 && lane == 0) {
    for (int s = 0; s < p.num_segs; ++s) {
      ptx::prefetch_tmap(&p.a_hi[s]);
      if (PASSES == 3) ptx::prefetch_tmap(&p.a_lo[s]);
    }
    ptx::prefetch_tmap(&p.b_hi);
    if (PASSES == 3) ptx::prefetch_tmap(&p.b_lo);
    for (int i = 0; i < Cfg::kStages; ++i) {
      ptx::mbar_init(&full_bar[i], 1);
      ptx::mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < Cfg::kAccStages; ++i) {
      ptx::mbar_init(&tmem_full_bar[i], 1);
      ptx::mbar_init(&tmem_empty_bar[i], kEpiWarps);
    }
    ptx::fence_barrier_init();
  }
  if (warp_idx == 1) {
    ptx::tmem_alloc<Cfg::kTmemCols>(&tmem_base_smem);
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = tmem_base_smem;

  // Everything above overlaps the tail of the previous kernel under programmatic dependent launch.
  ptx::pdl_wait_prior_grid();

  if (warp_idx == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = (tile / tiles_n) * kGemmBlockM;
        const int n0 = (tile % tiles_n) * BLOCK_N;
        int kcol = 0;
        for (int s = 0; s < p.num_segs; ++s) {
          const int row = m0 * p.seg_row_mul[s] + p.seg_row_shift[s];
          const int nkb = p.seg_kblocks[s];
          for (int kb = 0; kb < nkb; ++kb, ++it, kcol += kGemmBlockK) {
            const int stage = it % Cfg::kStages;
            const uint32_t phase = (it / Cfg::kStages) & 1;
            ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* st = smem + stage * Cfg::kStageBytes;
            ptx::mbar_expect_tx(&full_bar[stage], Cfg::kStageBytes);
            ptx::tma_load_2d(st, &p.a_hi[s], &full_bar[stage], kb * kGemmBlockK, row);
            ptx::tma_load_2d(st + Cfg::kSplit * Cfg::kABytes, &p.b_hi, &full_bar[stage], kcol, n0);
            if (PASSES == 3) {
              ptx::tma_load_2d(st + Cfg::kABytes, &p.a_lo[s], &full_bar[stage], kb * kGemmBlockK, row);
              ptx::tma_load_2d(st + 2 * Cfg::kABytes + Cfg::kBBytes, &p.b_lo, &full_bar[stage], kcol, n0);
            }
          }
        }
      }
    }
  } else if (warp_idx == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = ptx::make_idesc(/*TF32*/ 2, kGemmBlockM, BLOCK_N);
      int it = 0, tcount = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tcount) {
        const int acc_stage = tcount % Cfg::kAccStages;
        const uint32_t acc_phase = (tcount / Cfg::kAccStages) & 1;
        ptx::mbar_wait(&tmem_empty_bar[acc_stage], acc_phase ^ 1);  // epilogue has drained this accumulator
        ptx::tc_fence_after_sync();
        const uint32_t acc = tmem_base + static_cast<uint32_t>(acc_stage * Cfg::kAccCols);
        for (int ki = 0; ki < total_iters; ++ki, ++it) {
          const int stage = it % Cfg::kStages;
          const uint32_t phase = (it / Cfg::kStages) & 1;
          ptx::mbar_wait(&full_bar[stage], phase);
          ptx::tc_fence_after_sync();
          const uint32_t st = ptx::smem_u32(smem + stage * Cfg::kStageBytes);
          const uint64_t a_hi = ptx::make_desc_sw128_kmajor(st);
          const uint64_t b_hi = ptx::make_desc_sw128_kmajor(st + Cfg::kSplit * Cfg::kABytes);
          const uint64_t a_lo = ptx::make_desc_sw128_kmajor(st + Cfg::kABytes);
          const uint64_t b_lo = ptx::make_desc_sw128_kmajor(st + 2 * Cfg::kABytes + Cfg::kBBytes);
#pragma unroll
          for (int k = 0; k < kGemmBlockK / 8; ++k) {
            // advancing K by 8 fp32 = 32 bytes inside the 128-byte swizzle span: +2 in the (>>4) address field
            const uint64_t koff = static_cast<uint64_t>(k * 2);
            const uint32_t first = (ki > 0 || k > 0) ? 1u : 0u;
            if (PASSES == 3) {
              ptx::mma_tf32_ss(acc + BLOCK_N, a_lo + koff, b_hi + koff, idesc, first);
              ptx::mma_tf32_ss(acc + BLOCK_N, a_hi + koff, b_lo + koff, idesc, 1u);
              ptx::mma_tf32_ss(acc, a_hi + koff, b_hi + koff, idesc, first);
            } else {
              ptx::mma_tf32_ss(acc, a_hi + koff, b_hi + koff, idesc, first);
            }
          }
          ptx::mma_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs have read it
        }
        ptx::mma_commit(&tmem_full_bar[acc_stage]);  // accumulator complete
      }
    }
  } else {
    // ===================== epilogue (warps 2..9) =====================
    // TMEM lane quarter = warp_idx % 4 (hardware rule); the two warps sharing a quarter alternate 32-column chunks.
    const int q = warp_idx & 3;
    const int half = (warp_idx - 2) >> 2;
    const bool vec_ok = ((p.N & 3) == 0);
    int tcount = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tcount) {
      const int m0 = (tile / tiles_n) * kGemmBlockM;
      const int n0 = (tile % tiles_n) * BLOCK_N;
      const int acc_stage = tcount % Cfg::kAccStages;
      const uint32_t acc_phase = (tcount / Cfg::kAccStages) & 1;
      const int m = m0 + q * 32 + lane;
      const bool row_ok = m < p.M;
      bool row_real = true;
      int clip = 0;
      if (p.clip_rows > 0) {
        clip = m / p.clip_rows;
        row_real = (m - clip * p.clip_rows) < p.clip_valid;
      }
      const int64_t orow = static_cast<int64_t>(m) * p.out_row_mul + p.out_row_add;

      // The residual tile comes from L2/HBM: fetch the first chunk while the MMA warp is still accumulating (these
      // warps are otherwise idle), and each following chunk while the current one is being processed.
      float4 rcur[8], rnext[8];
      auto residual_vec_ok = [&](int c0) {
        const int nb = n0 + c0;
        return p.residual != nullptr && row_ok && vec_ok && c0 < BLOCK_N && nb + 32 <= p.N;
      };
      auto fetch_residual = [&](int c0, float4(&dst)[8]) {
        const float4* r = reinterpret_cast<const float4*>(p.residual + orow * p.ldr + n0 + c0);
#pragma unroll
        for (int j = 0; j < 8; ++j) dst[j] = r[j];
      };
      if (residual_vec_ok(half * 32)) fetch_residual(half * 32, rcur);

      ptx::mbar_wait(&tmem_full_bar[acc_stage], acc_phase);
      ptx::tc_fence_after_sync();

#pragma unroll 1
      for (int c0 = half * 32; c0 < BLOCK_N; c0 += 64) {
        uint32_t raw[32], raw2[32];
        const uint32_t taddr = tmem_base + static_cast<uint32_t>(acc_stage * Cfg::kAccCols) +
                               (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(c0);
        ptx::tmem_ld_32x32(taddr, raw);
        if (PASSES == 3) ptx::tmem_ld_32x32(taddr + BLOCK_N, raw2);
        const bool have_res = residual_vec_ok(c0);
        const bool next_res = residual_vec_ok(c0 + 64);
        if (next_res) fetch_residual(c0 + 64, rnext);
        ptx::tmem_ld_wait();
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(raw[j]);
        if (PASSES == 3) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] += __uint_as_float(raw2[j]);
        }
        const int nb = n0 + c0;
        if (nb >= p.N) continue;  // warp-uniform
        const bool full = vec_ok && (nb + 32 <= p.N);

        if (row_ok) {
          if (p.bias != nullptr) {
            if (full) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                const float4 t4 = __ldg(reinterpret_cast<const float4*>(p.bias + nb + j));
                v[j] += t4.x, v[j + 1] += t4.y, v[j + 2] += t4.z, v[j + 3] += t4.w;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (nb + j < p.N) v[j] += __ldg(p.bias + nb + j);
            }
          }
          if (p.act != kActNone) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = apply_act(v[j], p.act);
          }
          if (p.residual != nullptr) {
            if (have_res) {
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                v[4 * j] += rcur[j].x, v[4 * j + 1] += rcur[j].y, v[4 * j + 2] += rcur[j].z, v[4 * j + 3] += rcur[j].w;
              }
            } else {
              const float* r = p.residual + orow * p.ldr + nb;
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (nb + j < p.N) v[j] += r[j];
            }
          }
          if (!row_real) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = 0.0f;
          }
        }
        if (next_res) {
#pragma unroll
          for (int j = 0; j < 8; ++j) rcur[j] = rnext[j];
        }

        // ---- GroupNorm partial statistics over real rows ----
        if (p.gn_stats != nullptr) {
          const int gs = p.gn_group_size;
          const bool contrib = row_ok && row_real;
          const int clip0 = __shfl_sync(0xffffffffu, clip, 0);
          const bool uniform = __all_sync(0xffffffffu, clip == clip0);
          for (int jg = 0; jg < 32 && nb + jg < p.N; jg += (gs < 32 ? gs : 32)) {
            const int span = gs < 32 ? gs : 32;
            float s1 = 0.0f, s2 = 0.0f;
            if (contrib) {
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                if (j >= jg && j < jg + span && nb + j < p.N) {
                  s1 += v[j];
                  s2 += v[j] * v[j];
                }
              }
            }
            const int g = (nb + jg) / gs;
            if (uniform) {
#pragma unroll
              for (int off = 16; off > 0; off >>= 1) {
                s1 += __shfl_xor_sync(0xffffffffu, s1, off);
                s2 += __shfl_xor_sync(0xffffffffu, s2, off);
              }
              if (lane == 0) {
                double* dst = p.gn_stats + (static_cast<int64_t>(clip0) * p.gn_groups + g) * 2;
                atomicAdd(dst, static_cast<double>(s1));
                atomicAdd(dst + 1, static_cast<double>(s2));
              }
            } else if (contrib) {
              double* dst = p.gn_stats + (static_cast<int64_t>(clip) * p.gn_groups + g) * 2;
              atomicAdd(dst, static_cast<double>(s1));
              atomicAdd(dst + 1, static_cast<double>(s2));
            }
          }
        }

        if (row_ok) {
          if (p.out != nullptr) {
            float* o = p.out + orow * p.ldo + nb;
            if (full) {
#pragma unroll
              for (int j = 0; j < 32; j += 4)
                *reinterpret_cast<float4*>(o + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (nb + j < p.N) o[j] = v[j];
            }
          }
          if (p.out_hi != nullptr) {
            float* oh = p.out_hi + orow * p.lds + nb;
            float* ol = p.out_lo + orow * p.lds + nb;
            if (full) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                float4 h, l;
                h.x = ptx::to_tf32(v[j]), h.y = ptx::to_tf32(v[j + 1]);
                h.z = ptx::to_tf32(v[j + 2]), h.w = ptx::to_tf32(v[j + 3]);
                l.x = v[j] - h.x, l.y = v[j + 1] - h.y, l.z = v[j + 2] - h.z, l.w = v[j + 3] - h.w;
                *reinterpret_cast<float4*>(oh + j) = h;
                *reinterpret_cast<float4*>(ol + j) = l;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (nb + j < p.N) {
                  const float h = ptx::to_tf32(v[j]);
                  oh[j] = h;
                  ol[j] = v[j] - h;
                }
            }
          }
        }
      }
      // all TMEM reads of this accumulator stage are complete (tcgen05.wait::ld above): hand it back
      ptx::tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&tmem_empty_bar[acc_stage]);
    }
  }

  __syncthreads();
  ptx::pdl_launch_dependents();
  if (warp_idx == 1) {
    ptx::tc_fence_after_sync();
    ptx::tmem_dealloc<Cfg::kTmemCols>(tmem_base);
  }
}

__global__ void split_tf32_kernel(const float* __restrict__ x, float* __restrict__ hi, float* __restrict__ lo,
                                  int64_t n) {
  int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (; i < n; i += stride) {
    const float v = x[i];
    const float h = ptx::to_tf32(v);
    hi[i] = h;
    lo[i] = v - h;
  }
}

PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres);
    if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(ptr);
  });
  return fn;
}

template <int BLOCK_N, int PASSES>
cudaError_t launch_cfg(const GemmParams& p, int m_rows, int n_cols, cudaStream_t stream, bool pdl) {
  using Cfg = TileCfg<BLOCK_N, PASSES>;
  auto kern = gemm_tile_kernel<BLOCK_N, PASSES>;
  static bool attr_set = false;
  if (!attr_set) {  // normally done up front by gemm_init_attributes(); kept for stand-alone users of launch_gemm
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  static int num_sms = 0;
  if (num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || num_sms <= 0) num_sms = 148;
  }
  GemmParams q = p;
  q.grid_m_rows = m_rows;
  q.grid_n_cols = n_cols;
  const int tiles = ((n_cols + BLOCK_N - 1) / BLOCK_N) * ((m_rows + kGemmBlockM - 1) / kGemmBlockM);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(tiles < num_sms ? tiles : num_sms, 1, 1);
  cfg.blockDim = dim3(kThreads, 1, 1);
  cfg.dynamicSmemBytes = Cfg::kSmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, q);
}

}  // namespace

int make_tmap_2d(CUtensorMap* map, const float* base, int64_t rows, int64_t cols, int64_t ld, int box_rows,
                 int row_elem_stride) {
  auto fn = get_encode_fn();
  if (fn == nullptr) return -1;
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstride[1] = {static_cast<cuuint64_t>(ld) * sizeof(float)};
  // With a traversal stride s the box spans box_rows * s tensor rows and TMA delivers every s-th of them
  // (ceil(boxDim / elementStride) elements), so smem still receives exactly box_rows rows.
  cuuint32_t box[2] = {static_cast<cuuint32_t>(kGemmBlockK), static_cast<cuuint32_t>(box_rows * row_elem_stride)};
  cuuint32_t estride[2] = {1u, static_cast<cuuint32_t>(row_elem_stride)};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstride, box, estride,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return static_cast<int>(r);
}

cudaError_t launch_gemm(const GemmParams& p, int m_rows, int n_cols, int block_n, int passes, cudaStream_t stream,
                        bool pdl) {
#define ROHM_GEMM_CASE(BN)                                                             \
  case BN:                                                                             \
    return passes == 3 ? launch_cfg<BN, 3>(p, m_rows, n_cols, stream, pdl)             \
                       : launch_cfg<BN, 1>(p, m_rows, n_cols, stream, pdl);
  switch (block_n) {
    ROHM_GEMM_CASE(32)
    ROHM_GEMM_CASE(64)
    ROHM_GEMM_CASE(96)
    ROHM_GEMM_CASE(128)
    default:
      return cudaErrorInvalidValue;
  }
#undef ROHM_GEMM_CASE
}

template <int BLOCK_N, int PASSES>
static cudaError_t set_attr() {
  return cudaFuncSetAttribute(gemm_tile_kernel<BLOCK_N, PASSES>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              TileCfg<BLOCK_N, PASSES>::kSmemBytes);
}

cudaError_t gemm_init_attributes() {
  cudaError_t e;
  if ((e = set_attr<32, 1>()) != cudaSuccess) return e;
  if ((e = set_attr<32, 3>()) != cudaSuccess) return e;
  if ((e = set_attr<64, 1>()) != cudaSuccess) return e;
  if ((e = set_attr<64, 3>()) != cudaSuccess) return e;
  if ((e = set_attr<96, 1>()) != cudaSuccess) return e;
  if ((e = set_attr<96, 3>()) != cudaSuccess) return e;
  if ((e = set_attr<128, 1>()) != cudaSuccess) return e;
  if ((e = set_attr<128, 3>()) != cudaSuccess) return e;
  return cudaSuccess;
}

cudaError_t launch_split_tf32(const float* x, float* hi, float* lo, int64_t n, cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  const int threads = 256;
  int64_t blocks = (n + threads - 1) / threads;
  if (blocks > 148 * 16) blocks = 148 * 16;
  split_tf32_kernel<<<static_cast<unsigned>(blocks), threads, 0, stream>>>(x, hi, lo, n);
  return cudaGetLastError();
}

}  // namespace rohm
