// K5 / K6 — row normalisations (HBM-bound, one pass, 16-byte vector loads, warp-shuffle + smem block reduction).
//   rmsnorm   : HF-4.31 LlamaRMSNorm semantics: fp32 variance, normalised value rounded to
//               bf16, then multiplied by the bf16 weight (second rounding).
//   layernorm : nn.LayerNorm / SAM LayerNorm2d-on-NHWC rows (model/segment_anything_med2d/modeling/common.py:31-45):
//               fp32 statistics, single rounding on output. Optional fused GELU(erf) for the mask-decoder upscaler.
//   pool3_ln  : TokenCompressor front half (model/medplib/model/medplib_arch.py:67-77): AdaptiveAvgPool1d(576->256)
//               = mean of 3 consecutive tokens starting at floor(i*576/256), fused with LayerNorm(4096).
#include "internal.h"
#include "ptx.cuh"

namespace mpl {

template <int THREADS>
__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();  // protect red[] reuse across calls
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = (lane < THREADS / 32) ? red[lane] : 0.0f;
  t = warp_sum(t);
  return t;
}

constexpr int NORM_THREADS = 128;
constexpr int NORM_MAX_VEC = 4;  // supports D <= 128*8*4 = 4096

// x,y: [rows, D] bf16 with leading dims; D % 8 == 0, D <= 4096.
__global__ void __launch_bounds__(NORM_THREADS) rmsnorm_kernel(const __nv_bfloat16* __restrict__ x, long long ldx,
                                                               const __nv_bfloat16* __restrict__ w,
                                                               __nv_bfloat16* __restrict__ y, long long ldy, int D,
                                                               float eps) {
  griddep_wait();  // programmatic dependent launch: the previous kernel's outputs are read from here on
  griddep_launch_dependents();
  __shared__ float red[NORM_THREADS / 32];
  const long long row = blockIdx.x;
  const __nv_bfloat16* xr = x + row * ldx;
  __nv_bfloat16* yr = y + row * ldy;
  float v[NORM_MAX_VEC][8];
  float ss = 0.0f;
#pragma unroll
  for (int i = 0; i < NORM_MAX_VEC; ++i) {
    const int c = (i * NORM_THREADS + threadIdx.x) * 8;
    if (c < D) {
      const uint4 raw = *reinterpret_cast<const uint4*>(xr + c);
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = __bfloat1622float2(h[e]);
        v[i][2 * e] = f.x;
        v[i][2 * e + 1] = f.y;
        ss += f.x * f.x + f.y * f.y;
      }
    }
  }
  ss = block_sum<NORM_THREADS>(ss, red);
  const float rstd = rsqrtf(ss / static_cast<float>(D) + eps);
#pragma unroll
  for (int i = 0; i < NORM_MAX_VEC; ++i) {
    const int c = (i * NORM_THREADS + threadIdx.x) * 8;
    if (c < D) {
      const uint4 wraw = *reinterpret_cast<const uint4*>(w + c);
      const __nv_bfloat162* wh = reinterpret_cast<const __nv_bfloat162*>(&wraw);
      uint4 o;
      uint32_t* op = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 wf = __bfloat1622float2(wh[e]);
        op[e] = pack_bf16(wf.x * bf16_round(v[i][2 * e] * rstd), wf.y * bf16_round(v[i][2 * e + 1] * rstd));
      }
      *reinterpret_cast<uint4*>(yr + c) = o;
    }
  }
}

// Generic LayerNorm over the last dim; any D % 8 == 0 (loops when D > 4096 is not needed here).
// act: 0 none, MPL_ACT_GELU fused on the output.
__global__ void __launch_bounds__(NORM_THREADS) layernorm_kernel(const __nv_bfloat16* __restrict__ x, long long ldx,
                                                                 const __nv_bfloat16* __restrict__ w,
                                                                 const __nv_bfloat16* __restrict__ b,
                                                                 __nv_bfloat16* __restrict__ y, long long ldy, int D,
                                                                 float eps, int act) {
  griddep_wait();  // programmatic dependent launch: the previous kernel's outputs are read from here on
  griddep_launch_dependents();
  __shared__ float red[NORM_THREADS / 32];
  const long long row = blockIdx.x;
  const __nv_bfloat16* xr = x + row * ldx;
  __nv_bfloat16* yr = y + row * ldy;
  float v[NORM_MAX_VEC][8];
  float s = 0.0f;
#pragma unroll
  for (int i = 0; i < NORM_MAX_VEC; ++i) {
    const int c = (i * NORM_THREADS + threadIdx.x) * 8;
    if (c < D) {
      const uint4 raw = *reinterpret_cast<const uint4*>(xr + c);
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = __bfloat1622float2(h[e]);
        v[i][2 * e] = f.x;
        v[i][2 * e + 1] = f.y;
        s += f.x + f.y;
      }
    }
  }
  const float mean = block_sum<NORM_THREADS>(s, red) / static_cast<float>(D);
  float sq = 0.0f;
#pragma unroll
  for (int i = 0; i < NORM_MAX_VEC; ++i) {
    const int c = (i * NORM_THREADS + threadIdx.x) * 8;
    if (c < D) {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float d = v[i][e] - mean;
        sq += d * d;
      }
    }
  }
  const float rstd = rsqrtf(block_sum<NORM_THREADS>(sq, red) / static_cast<float>(D) + eps);
#pragma unroll
  for (int i = 0; i < NORM_MAX_VEC; ++i) {
    const int c = (i * NORM_THREADS + threadIdx.x) * 8;
    if (c < D) {
      const uint4 wraw = *reinterpret_cast<const uint4*>(w + c);
      const uint4 braw = *reinterpret_cast<const uint4*>(b + c);
      const __nv_bfloat162* wh = reinterpret_cast<const __nv_bfloat162*>(&wraw);
      const __nv_bfloat162* bh = reinterpret_cast<const __nv_bfloat162*>(&braw);
      uint4 o;
      uint32_t* op = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 wf = __bfloat1622float2(wh[e]);
        const float2 bf = __bfloat1622float2(bh[e]);
        float o0 = (v[i][2 * e] - mean) * rstd * wf.x + bf.x;
        float o1 = (v[i][2 * e + 1] - mean) * rstd * wf.y + bf.y;
        if (act == MPL_ACT_GELU) {
          o0 = bf16_round(o0);
          o1 = bf16_round(o1);
          o0 = 0.5f * o0 * (1.0f + erff(o0 * 0.70710678118654752f));
          o1 = 0.5f * o1 * (1.0f + erff(o1 * 0.70710678118654752f));
        }
        op[e] = pack_bf16(o0, o1);
      }
      *reinterpret_cast<uint4*>(yr + c) = o;
    }
  }
}

// x: [N, Tin, D] -> y: [N, Tout, D]; window of output i = tokens [floor(i*Tin/Tout), ceil((i+1)*Tin/Tout)).
// The pooled value is rounded to bf16 (the reference's pool output dtype) before LayerNorm.
__global__ void __launch_bounds__(NORM_THREADS) pool_ln_kernel(const __nv_bfloat16* __restrict__ x,
                                                               const __nv_bfloat16* __restrict__ w,
                                                               const __nv_bfloat16* __restrict__ b,
                                                               __nv_bfloat16* __restrict__ y, int Tin, int Tout, int D,
                                                               float eps) {
  griddep_wait();  // programmatic dependent launch: the previous kernel's outputs are read from here on
  griddep_launch_dependents();
  __shared__ float red[NORM_THREADS / 32];
  const int n = blockIdx.x / Tout;
  const int i = blockIdx.x % Tout;
  const int t0 = static_cast<int>((static_cast<long long>(i) * Tin) / Tout);
  const int t1 = static_cast<int>((static_cast<long long>(i + 1) * Tin + Tout - 1) / Tout);
  const float inv = 1.0f / static_cast<float>(t1 - t0);
  const __nv_bfloat16* xb = x + (static_cast<long long>(n) * Tin + t0) * D;
  __nv_bfloat16* yr = y + (static_cast<long long>(n) * Tout + i) * D;
  float v[NORM_MAX_VEC][8];
  float s = 0.0f;
  const int nt = t1 - t0;
  constexpr int WMAX = 4;  // windows of up to 4 tokens (576 -> 256: always 3) have every load in flight before the first use
  if (nt <= WMAX) {
    uint4 raw[NORM_MAX_VEC][WMAX];
#pragma unroll
    for (int k = 0; k < NORM_MAX_VEC; ++k) {
      const int c = (k * NORM_THREADS + threadIdx.x) * 8;
#pragma unroll
      for (int t = 0; t < WMAX; ++t)
        raw[k][t] = (c < D && t < nt) ? *reinterpret_cast<const uint4*>(xb + static_cast<long long>(t) * D + c)
                                      : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int k = 0; k < NORM_MAX_VEC; ++k) {
      const int c = (k * NORM_THREADS + threadIdx.x) * 8;
      if (c < D) {
#pragma unroll
        for (int e = 0; e < 8; ++e) v[k][e] = 0.0f;
#pragma unroll
        for (int t = 0; t < WMAX; ++t) {  // (absent tokens are +0: the sum is unchanged, same order as the loop below)
          const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw[k][t]);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 f = __bfloat1622float2(h[e]);
            v[k][2 * e] += f.x;
            v[k][2 * e + 1] += f.y;
          }
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          v[k][e] = bf16_round(v[k][e] * inv);
          s += v[k][e];
        }
      }
    }
  } else {
#pragma unroll
    for (int k = 0; k < NORM_MAX_VEC; ++k) {
      const int c = (k * NORM_THREADS + threadIdx.x) * 8;
      if (c < D) {
#pragma unroll
        for (int e = 0; e < 8; ++e) v[k][e] = 0.0f;
        for (int t = 0; t < nt; ++t) {
          const uint4 raw = *reinterpret_cast<const uint4*>(xb + static_cast<long long>(t) * D + c);
          const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 f = __bfloat1622float2(h[e]);
            v[k][2 * e] += f.x;
            v[k][2 * e + 1] += f.y;
          }
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          v[k][e] = bf16_round(v[k][e] * inv);
          s += v[k][e];
        }
      }
    }
  }
  const float mean = block_sum<NORM_THREADS>(s, red) / static_cast<float>(D);
  float sq = 0.0f;
#pragma unroll
  for (int k = 0; k < NORM_MAX_VEC; ++k) {
    const int c = (k * NORM_THREADS + threadIdx.x) * 8;
    if (c < D) {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float d = v[k][e] - mean;
        sq += d * d;
      }
    }
  }
  const float rstd = rsqrtf(block_sum<NORM_THREADS>(sq, red) / static_cast<float>(D) + eps);
#pragma unroll
  for (int k = 0; k < NORM_MAX_VEC; ++k) {
    const int c = (k * NORM_THREADS + threadIdx.x) * 8;
    if (c < D) {
      const uint4 wraw = *reinterpret_cast<const uint4*>(w + c);
      const uint4 braw = *reinterpret_cast<const uint4*>(b + c);
      const __nv_bfloat162* wh = reinterpret_cast<const __nv_bfloat162*>(&wraw);
      const __nv_bfloat162* bh = reinterpret_cast<const __nv_bfloat162*>(&braw);
      uint4 o;
      uint32_t* op = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 wf = __bfloat1622float2(wh[e]);
        const float2 bf = __bfloat1622float2(bh[e]);
        op[e] = pack_bf16((v[k][2 * e] - mean) * rstd * wf.x + bf.x, (v[k][2 * e + 1] - mean) * rstd * wf.y + bf.y);
      }
      *reinterpret_cast<uint4*>(yr + c) = o;
    }
  }
}

static bool norm_args_ok(const void* x, const void* y, int D, long long ldx, long long ldy) {
  return x != nullptr && y != nullptr && D > 0 && D <= NORM_THREADS * 8 * NORM_MAX_VEC && (D % 8) == 0 &&
         (ldx % 8) == 0 && (ldy % 8) == 0;
}

}  // namespace mpl

extern "C" int mpl_rmsnorm(const void* x, long long ldx, const void* weight, void* y, long long ldy, int rows, int D,
                           float eps, void* stream) {
  if (rows <= 0) return MPL_OK;
  if (!mpl::norm_args_ok(x, y, D, ldx, ldy) || weight == nullptr) return MPL_ERR_ARG;
  mpl::launch_pdl(mpl::rmsnorm_kernel, dim3(rows), dim3(mpl::NORM_THREADS), 0, static_cast<cudaStream_t>(stream), 
      static_cast<const __nv_bfloat16*>(x), ldx, static_cast<const __nv_bfloat16*>(weight),
      static_cast<__nv_bfloat16*>(y), ldy, D, eps);
  return mpl::launch_status();
}

extern "C" int mpl_layernorm(const void* x, long long ldx, const void* weight, const void* bias, void* y,
                             long long ldy, int rows, int D, float eps, int act, void* stream) {
  if (rows <= 0) return MPL_OK;
  if (!mpl::norm_args_ok(x, y, D, ldx, ldy) || weight == nullptr || bias == nullptr) return MPL_ERR_ARG;
  mpl::launch_pdl(mpl::layernorm_kernel, dim3(rows), dim3(mpl::NORM_THREADS), 0, static_cast<cudaStream_t>(stream), 
      static_cast<const __nv_bfloat16*>(x), ldx, static_cast<const __nv_bfloat16*>(weight),
      static_cast<const __nv_bfloat16*>(bias), static_cast<__nv_bfloat16*>(y), ldy, D, eps, act);
  return mpl::launch_status();
}

extern "C" int mpl_pool_layernorm(const void* x, const void* weight, const void* bias, void* y, int n, int t_in,
                                  int t_out, int D, float eps, void* stream) {
  if (n <= 0 || t_out <= 0) return MPL_OK;
  if (!mpl::norm_args_ok(x, y, D, D, D) || weight == nullptr || bias == nullptr || t_in < t_out) return MPL_ERR_ARG;
  mpl::launch_pdl(mpl::pool_ln_kernel, dim3(n * t_out), dim3(mpl::NORM_THREADS), 0, static_cast<cudaStream_t>(stream), 
      static_cast<const __nv_bfloat16*>(x), static_cast<const __nv_bfloat16*>(weight),
      static_cast<const __nv_bfloat16*>(bias), static_cast<__nv_bfloat16*>(y), t_in, t_out, D, eps);
  return mpl::launch_status();
}
