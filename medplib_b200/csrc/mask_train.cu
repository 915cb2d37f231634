// Backward kernels of the grounding head for the train step (SURVEY.md §8 a-10, a-13, a-14, a-15 with
// inference=False): text_hidden_fcs, the SAM-Med2D two-way mask decoder, postprocess_masks and the four mask losses.
//   reference autograd being replaced: model/MedPLIB.py:456-559 over
//   model/segment_anything_med2d/modeling/{mask_decoder.py:71-153, transformer.py:16-244}.
// The whole head moves < 2 MB and < 1 GFLOP per mask, so these are latency-bound small kernels: plain fp32 FMA tiles
// through shared memory, coalesced along the contiguous dimension, no tensor cores (nothing here is GEMM-bound).
//   gemm_small       C[M,N] (+)= op(A) op(B), general strides (covers dX = dY W, dW = dY^T X and outer products)
//   col_sum          out[n] += sum_m X[m,n]               (bias gradients, bf16 -> f32 accumulation)
//   layernorm_bwd    nn.LayerNorm / LayerNorm2d-on-NHWC-rows backward (dx, dweight, dbias)
//   act_fwd/act_bwd  GELU(erf) / ReLU
//   attn_small_bwd   softmax(q k^T * scale) v backward for head_dim <= 32, Tq*Tk tiny (one CTA per head)
//   bilinear_bwd     adjoint of mpl_bilinear_resize (gather form: deterministic, no atomics)
//   mask_losses_bwd  d(BCE, Dice, IoU-MSE, Focal)/d(mask logits, predicted IoU) in one pass
#include "internal.h"
#include "ptx.cuh"

namespace mpl {

using bf16_t = __nv_bfloat16;
#define ST(s) static_cast<cudaStream_t>(s)

__device__ __forceinline__ float ld_any(const void* p, int is_f32, long long i) {
  return is_f32 ? static_cast<const float*>(p)[i] : __bfloat162float(static_cast<const bf16_t*>(p)[i]);
}

// ------------------------------------------------------------------------------------------------- small GEMM
// C[m,n] (+)= sum_k A[m*sam + k*sak] * B[k*sbk + n*sbn]; A, B bf16 or f32; C bf16 (overwrite) or f32 (accumulate flag).
constexpr int GS_T = 32;
__global__ void __launch_bounds__(256) gemm_small_kernel(const void* __restrict__ A, int a_f32, long long sam,
                                                         long long sak, const void* __restrict__ B, int b_f32,
                                                         long long sbk, long long sbn, void* __restrict__ C, int c_f32,
                                                         long long ldc, int accumulate, int M, int N, int K,
                                                         int k_per_split) {
  __shared__ float sA[GS_T][GS_T + 1], sB[GS_T][GS_T + 1];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  const int m0 = blockIdx.y * GS_T, n0 = blockIdx.x * GS_T;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  // gridDim.z > 1: split-K (fp32 accumulate mode only; partial sums meet in fp32 atomics)
  const int k_beg = blockIdx.z * k_per_split;
  const int k_end = min(K, k_beg + k_per_split);
  for (int k0 = k_beg; k0 < k_end; k0 += GS_T) {
    // choose the fastest-varying index per operand so the global reads coalesce along its contiguous dimension
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = ty + 8 * i;
      {
        int m, k;
        if (sak == 1) { m = m0 + r; k = k0 + tx; } else { m = m0 + tx; k = k0 + r; }
        const float v = (m < M && k < k_end) ? ld_any(A, a_f32, m * sam + k * sak) : 0.f;
        sA[m - m0][k - k0] = v;
      }
      {
        int k, n;
        if (sbn == 1) { k = k0 + r; n = n0 + tx; } else { k = k0 + tx; n = n0 + r; }
        const float v = (k < k_end && n < N) ? ld_any(B, b_f32, k * sbk + n * sbn) : 0.f;
        sB[k - k0][n - n0] = v;
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < GS_T; ++kk) {
      const float b = sB[kk][tx];
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[i] = fmaf(sA[ty + 8 * i][kk], b, acc[i]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty + 8 * i, n = n0 + tx;
    if (m < M && n < N) {
      const long long o = m * ldc + n;
      if (c_f32) {
        float* c = static_cast<float*>(C);
        if (gridDim.z > 1)
          atomicAdd(c + o, acc[i]);
        else
          c[o] = accumulate ? c[o] + acc[i] : acc[i];
      } else {
        static_cast<bf16_t*>(C)[o] = __float2bfloat16_rn(acc[i]);
      }
    }
  }
}

// out[n] += sum_m X[m*ld + n]   (one thread per column, coalesced across the warp; blockIdx.y splits the rows)
constexpr int CS_ROWS = 128;
__global__ void __launch_bounds__(256) col_sum_kernel(const void* __restrict__ X, int x_f32, long long ld,
                                                      float* __restrict__ out, int M, int N) {
  const int n = blockIdx.x * 256 + threadIdx.x;
  if (n >= N) return;
  const int m_beg = blockIdx.y * CS_ROWS, m_end = min(M, m_beg + CS_ROWS);
  float acc = 0.f;
  for (int m = m_beg; m < m_end; ++m) acc += ld_any(X, x_f32, m * ld + n);
  if (gridDim.y > 1)
    atomicAdd(out + n, acc);
  else
    out[n] += acc;
}

// ------------------------------------------------------------------------------------------------- LayerNorm bwd
// One warp per row (D <= 4096): xhat = (x - mean) * rstd, g = dy * w;
// dx = rstd * (g - mean(g) - xhat * mean(g * xhat)); dweight += dy * xhat, dbias += dy (fp32 atomics over rows).
__global__ void __launch_bounds__(256) layernorm_bwd_kernel(const bf16_t* __restrict__ x, long long ldx,
                                                            const bf16_t* __restrict__ w,
                                                            const bf16_t* __restrict__ dy, long long lddy,
                                                            bf16_t* __restrict__ dx, long long lddx,
                                                            float* __restrict__ dw, float* __restrict__ db, int rows,
                                                            int D, float eps) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= rows) return;
  const bf16_t* xr = x + row * ldx;
  const bf16_t* dr = dy + row * lddy;
  float s = 0.f, ss = 0.f;
  for (int c = lane; c < D; c += 32) {
    const float v = __bfloat162float(xr[c]);
    s += v;
    ss += v * v;
  }
  s = warp_sum(s);
  ss = warp_sum(ss);
  const float mean = s / D;
  const float rstd = rsqrtf(fmaxf(ss / D - mean * mean, 0.f) + eps);
  float sg = 0.f, sgx = 0.f;
  for (int c = lane; c < D; c += 32) {
    const float xh = (__bfloat162float(xr[c]) - mean) * rstd;
    const float g = __bfloat162float(dr[c]) * __bfloat162float(w[c]);
    sg += g;
    sgx += g * xh;
  }
  sg = warp_sum(sg) / D;
  sgx = warp_sum(sgx) / D;
  for (int c = lane; c < D; c += 32) {
    const float xh = (__bfloat162float(xr[c]) - mean) * rstd;
    const float d = __bfloat162float(dr[c]);
    const float g = d * __bfloat162float(w[c]);
    dx[row * lddx + c] = __float2bfloat16_rn(rstd * (g - sg - xh * sgx));
    if (dw != nullptr) atomicAdd(dw + c, d * xh);
    if (db != nullptr) atomicAdd(db + c, d);
  }
}

// ------------------------------------------------------------------------------------------------- activations
__global__ void __launch_bounds__(256) act_fwd_kernel(const bf16_t* __restrict__ x, bf16_t* __restrict__ y,
                                                      long long n, int act) {
  const long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
  if (i >= n) return;
  const float v = __bfloat162float(x[i]);
  float r = v;
  if (act == MPL_ACT_GELU) r = 0.5f * v * (1.f + erff(v * 0.70710678118654752f));
  if (act == MPL_ACT_RELU) r = v > 0.f ? v : 0.f;
  y[i] = __float2bfloat16_rn(r);
}
// GELU: x = the activation's INPUT; ReLU: x may be input or output (sign test only). dx may alias dy.
__global__ void __launch_bounds__(256) act_bwd_kernel(const bf16_t* __restrict__ x, const bf16_t* __restrict__ dy,
                                                      bf16_t* __restrict__ dx, long long n, int act) {
  const long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
  if (i >= n) return;
  const float v = __bfloat162float(x[i]), d = __bfloat162float(dy[i]);
  float r = d;
  if (act == MPL_ACT_GELU) {
    const float cdf = 0.5f * (1.f + erff(v * 0.70710678118654752f));
    const float pdf = 0.3989422804014327f * __expf(-0.5f * v * v);
    r = d * (cdf + v * pdf);
  }
  if (act == MPL_ACT_RELU) r = v > 0.f ? d : 0.f;
  dx[i] = __float2bfloat16_rn(r);
}

// ------------------------------------------------------------------------------------------------- small attention bwd
// One CTA per (head, batch). q [B*Tq, ldq], k/v [B*Tk, ld] (batches stacked along rows), head h owns columns
// [h*d, (h+1)*d), d <= 32.
// K, V and the dK, dV accumulators live in shared memory as fp32 (row pitch d+1); every warp walks query rows.
__global__ void __launch_bounds__(256) attn_small_bwd_kernel(
    const bf16_t* __restrict__ q, long long ldq, const bf16_t* __restrict__ k, long long ldk,
    const bf16_t* __restrict__ v, long long ldv, const bf16_t* __restrict__ dO, long long ldo, bf16_t* __restrict__ dq,
    long long lddq, bf16_t* __restrict__ dk, long long lddk, bf16_t* __restrict__ dv, long long lddv, int Tq, int Tk,
    int d, float scale) {
  extern __shared__ float sm[];
  const int P = d + 1;
  float* sK = sm;
  float* sV = sK + Tk * P;
  float* sdK = sV + Tk * P;
  float* sdV = sdK + Tk * P;
  float* wbuf = sdV + Tk * P;  // per warp: p[Tk], ds[Tk], q[32], do[32]
  const int h = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int col0 = h * d;
  {
    const long long bq = static_cast<long long>(blockIdx.y) * Tq, bk = static_cast<long long>(blockIdx.y) * Tk;
    q += bq * ldq;
    dO += bq * ldo;
    dq += bq * lddq;
    k += bk * ldk;
    v += bk * ldv;
    dk += bk * lddk;
    dv += bk * lddv;
  }
  for (int i = threadIdx.x; i < Tk * d; i += 256) {
    const int j = i / d, c = i % d;
    sK[j * P + c] = __bfloat162float(k[j * ldk + col0 + c]);
    sV[j * P + c] = __bfloat162float(v[j * ldv + col0 + c]);
    sdK[j * P + c] = 0.f;
    sdV[j * P + c] = 0.f;
  }
  __syncthreads();
  float* wp = wbuf + warp * (2 * Tk + 64);
  float* wds = wp + Tk;
  float* wq = wds + Tk;
  float* wdo = wq + 32;
  for (int i = warp; i < Tq; i += 8) {
    if (lane < d) {
      wq[lane] = __bfloat162float(q[i * ldq + col0 + lane]);
      wdo[lane] = __bfloat162float(dO[i * ldo + col0 + lane]);
    }
    __syncwarp();
    float mx = -INFINITY;
    for (int j = lane; j < Tk; j += 32) {
      float s = 0.f, dp = 0.f;
      for (int c = 0; c < d; ++c) {
        s = fmaf(wq[c], sK[j * P + c], s);
        dp = fmaf(wdo[c], sV[j * P + c], dp);
      }
      s *= scale;
      wp[j] = s;
      wds[j] = dp;
      mx = fmaxf(mx, s);
    }
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < Tk; j += 32) {
      const float e = __expf(wp[j] - mx);
      wp[j] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    const float inv = 1.f / sum;
    float delta = 0.f;
    for (int j = lane; j < Tk; j += 32) {
      const float p = wp[j] * inv;
      wp[j] = p;
      delta += p * wds[j];
    }
    delta = warp_sum(delta);
    for (int j = lane; j < Tk; j += 32) wds[j] = wp[j] * (wds[j] - delta) * scale;
    __syncwarp();
    if (lane < d) {
      float acc = 0.f;
      const float qc = wq[lane], doc = wdo[lane];
      for (int j = 0; j < Tk; ++j) {
        const float ds = wds[j], p = wp[j];
        acc = fmaf(ds, sK[j * P + lane], acc);
        atomicAdd(&sdK[j * P + lane], ds * qc);
        atomicAdd(&sdV[j * P + lane], p * doc);
      }
      dq[i * lddq + col0 + lane] = __float2bfloat16_rn(acc);
    }
    __syncwarp();
  }
  __syncthreads();
  for (int i = threadIdx.x; i < Tk * d; i += 256) {
    const int j = i / d, c = i % d;
    dk[j * lddk + col0 + c] = __float2bfloat16_rn(sdK[j * P + c]);
    dv[j * lddv + col0 + c] = __float2bfloat16_rn(sdV[j * P + c]);
  }
}

// ------------------------------------------------------------------------------------------------- bilinear bwd
// dx[n, iy, ix] = sum over output pixels whose 2x2 footprint (exactly as the forward computes it) touches (iy, ix).
__device__ __forceinline__ void bil_src(int o, float s, int In, int* i0, int* i1, float* l) {
  float f = s * (o + 0.5f) - 0.5f;
  f = f < 0.f ? 0.f : f;
  const int a = min(static_cast<int>(f), In - 1);
  *i0 = a;
  *i1 = min(a + 1, In - 1);
  *l = f - a;
}
__global__ void __launch_bounds__(256) bilinear_bwd_kernel(const void* __restrict__ dy, int dy_f32, int Hout, int Wout,
                                                           bf16_t* __restrict__ dx, long long dx_sn, long long dx_sy,
                                                           int Hin, int Win, int N) {
  const long long gid = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
  if (gid >= static_cast<long long>(N) * Hin * Win) return;
  const int ix = gid % Win, iy = (gid / Win) % Hin, n = gid / (static_cast<long long>(Win) * Hin);
  const float sy = static_cast<float>(Hin) / Hout, sx = static_cast<float>(Win) / Wout;
  // candidate output range: source coordinate within (i-1, i+1), widened by one and clamped
  const int oy_lo = max(0, static_cast<int>(floorf((iy - 1 + 0.5f) / sy - 0.5f)) - 1);
  const int oy_hi = min(Hout - 1, static_cast<int>(ceilf((iy + 1 + 0.5f) / sy - 0.5f)) + 1);
  const int ox_lo = max(0, static_cast<int>(floorf((ix - 1 + 0.5f) / sx - 0.5f)) - 1);
  const int ox_hi = min(Wout - 1, static_cast<int>(ceilf((ix + 1 + 0.5f) / sx - 0.5f)) + 1);
  float acc = 0.f;
  for (int oy = oy_lo; oy <= oy_hi; ++oy) {
    int y0, y1;
    float ly;
    bil_src(oy, sy, Hin, &y0, &y1, &ly);
    float wy = 0.f;
    if (y0 == iy) wy += 1.f - ly;
    if (y1 == iy) wy += ly;
    if (wy == 0.f) continue;
    for (int ox = ox_lo; ox <= ox_hi; ++ox) {
      int x0, x1;
      float lx;
      bil_src(ox, sx, Win, &x0, &x1, &lx);
      float wx = 0.f;
      if (x0 == ix) wx += 1.f - lx;
      if (x1 == ix) wx += lx;
      if (wx == 0.f) continue;
      acc += wy * wx * ld_any(dy, dy_f32, (static_cast<long long>(n) * Hout + oy) * Wout + ox);
    }
  }
  dx[n * dx_sn + iy * dx_sy + ix] = __float2bfloat16_rn(acc);
}

// ------------------------------------------------------------------------------------------------- mask losses bwd
// Given sums6 = {sum bce, sum p, sum t, sum p t, focal_pos, focal_neg} of the forward (mpl_mask_losses) and
// dl[4] = d total / d {bce, dice, iou, focal}:  dpred[i] (bf16) and dpred_iou (f32[1]).
__global__ void __launch_bounds__(256) mask_losses_bwd_kernel(const bf16_t* __restrict__ pred,
                                                              const float* __restrict__ gt,
                                                              const bf16_t* __restrict__ pred_iou,
                                                              const float* __restrict__ sums,
                                                              const float* __restrict__ dl, long long n,
                                                              bf16_t* __restrict__ dpred,
                                                              float* __restrict__ dpred_iou) {
  const long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
  const float sp = sums[1], st = sums[2], I = sums[3];
  const float d_bce = dl[0], d_dice = dl[1], d_iou = dl[2], d_focal = dl[3];
  const float U = sp + st;                 // dice denominator
  const float Un = sp + st - I;            // IoU union
  const float iou = (I + 1e-7f) / (Un + 1e-7f);
  const float piou = __bfloat162float(*pred_iou);
  if (i == 0 && dpred_iou != nullptr) *dpred_iou = d_iou * (-2.f) * (iou - piou);
  if (i >= n) return;
  const float x = __bfloat162float(pred[i]), t = gt[i];
  const float p = 1.f / (1.f + __expf(-x));
  const float dpdx = p * (1.f - p);
  float g = d_bce * (p - t) / static_cast<float>(n);
  // dice = 1 - (2 I + e) / (U + e)
  const float e6 = 1e-6f;
  g += d_dice * (-(2.f * t * (U + e6) - (2.f * I + e6)) / ((U + e6) * (U + e6))) * dpdx;
  // (iou - piou)^2, iou = (I + e) / (Un + e): d iou / d p_i = (t (Un + e) - (I + e)(1 - t)) / (Un + e)^2
  const float e7 = 1e-7f;
  g += d_iou * 2.f * (iou - piou) * ((t * (Un + e7) - (I + e7) * (1.f - t)) / ((Un + e7) * (Un + e7))) * dpdx;
  // focal (gamma 2, alpha .25), normalised by n + 1e-12
  const float e12 = 1e-12f, al = 0.25f;
  const float dpos = -al * t * (-2.f * (1.f - p) * __logf(p + e12) + (1.f - p) * (1.f - p) / (p + e12));
  const float dneg = -(1.f - al) * (1.f - t) * (2.f * p * __logf(1.f - p + e12) - p * p / (1.f - p + e12));
  g += d_focal * (dpos + dneg) / (static_cast<float>(n) + e12) * dpdx;
  dpred[i] = __float2bfloat16_rn(g);
}

// ------------------------------------------------------------------------------------------------- dropout
// out = (accumulate ? out : 0) + x * mask * scale  (peft's lora_dropout: forward on the adapter input, backward on the
// adapter's input gradient); 8 elements per thread, n % 8 == 0.
__global__ void __launch_bounds__(256) mask_scale_kernel(const bf16_t* __restrict__ x,
                                                         const unsigned char* __restrict__ mask, float scale,
                                                         bf16_t* __restrict__ out, int accumulate, long long n) {
  const long long i = (static_cast<long long>(blockIdx.x) * 256 + threadIdx.x) * 8;
  if (i >= n) return;
  const uint4 xv = *reinterpret_cast<const uint4*>(x + i);
  const uint2 mv = *reinterpret_cast<const uint2*>(mask + i);
  uint4 ov = accumulate ? *reinterpret_cast<const uint4*>(out + i) : make_uint4(0, 0, 0, 0);
  const bf16_t* xe = reinterpret_cast<const bf16_t*>(&xv);
  const unsigned char* me = reinterpret_cast<const unsigned char*>(&mv);
  bf16_t* oe = reinterpret_cast<bf16_t*>(&ov);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float v = me[j] ? __bfloat162float(xe[j]) * scale : 0.f;
    oe[j] = __float2bfloat16_rn(accumulate ? __bfloat162float(oe[j]) + bf16_round(v) : v);
  }
  *reinterpret_cast<uint4*>(out + i) = ov;
}

// ------------------------------------------------------------------------------------------------- token pool backward
// Adjoint of AdaptiveAvgPool1d over the token axis (TokenCompressor / MaskTokenEncoder, medplib_arch.py:67-108):
// window j = [floor(j*Tin/Tout), ceil((j+1)*Tin/Tout)); dx[n,t,:] = sum over the windows that contain t of dy[n,j,:]/len_j.
// Gather form (one CTA per input token): deterministic, every dy row read ~once.
__global__ void __launch_bounds__(128) token_pool_bwd_kernel(const bf16_t* __restrict__ dy, bf16_t* __restrict__ dx,
                                                             int t_in, int t_out, int D) {
  const int n = blockIdx.x / t_in, t = blockIdx.x % t_in;
  // candidate windows: j with start_j <= t < end_j; start_j is non-decreasing, so scan around floor(t*Tout/Tin)
  int j0 = static_cast<int>((static_cast<long long>(t) * t_out) / t_in);
  while (j0 > 0 && (static_cast<long long>(j0) * t_in + t_out - 1) / t_out > t) --j0;  // end_{j0-1} = ceil(j0*Tin/Tout) > t
  for (int c = threadIdx.x * 2; c < D; c += 256) {
    float a0 = 0.f, a1 = 0.f;
    for (int j = j0; j < t_out; ++j) {
      const int st = static_cast<int>((static_cast<long long>(j) * t_in) / t_out);
      if (st > t) break;
      const int en = static_cast<int>((static_cast<long long>(j + 1) * t_in + t_out - 1) / t_out);
      if (t >= en) continue;
      const float inv = 1.0f / static_cast<float>(en - st);
      const __nv_bfloat162 v = *reinterpret_cast<const __nv_bfloat162*>(dy + (static_cast<long long>(n) * t_out + j) * D + c);
      a0 += __bfloat162float(v.x) * inv;
      a1 += __bfloat162float(v.y) * inv;
    }
    *reinterpret_cast<__nv_bfloat162*>(dx + (static_cast<long long>(n) * t_in + t) * D + c) = __floats2bfloat162_rn(a0, a1);
  }
}

// AdaptiveAvgPool1d over the token axis itself (fp32 mean, one rounding): y[n,j,:] = mean_{t in window j} x[n,t,:]
__global__ void __launch_bounds__(128) token_pool_kernel(const bf16_t* __restrict__ x, bf16_t* __restrict__ y, int t_in,
                                                         int t_out, int D) {
  const int n = blockIdx.x / t_out, j = blockIdx.x % t_out;
  const int st = static_cast<int>((static_cast<long long>(j) * t_in) / t_out);
  const int en = static_cast<int>((static_cast<long long>(j + 1) * t_in + t_out - 1) / t_out);
  const float inv = 1.0f / static_cast<float>(en - st);
  for (int c = threadIdx.x * 2; c < D; c += 256) {
    float a0 = 0.f, a1 = 0.f;
    for (int t = st; t < en; ++t) {
      const __nv_bfloat162 v = *reinterpret_cast<const __nv_bfloat162*>(x + (static_cast<long long>(n) * t_in + t) * D + c);
      a0 += __bfloat162float(v.x);
      a1 += __bfloat162float(v.y);
    }
    *reinterpret_cast<__nv_bfloat162*>(y + (static_cast<long long>(n) * t_out + j) * D + c) =
        __floats2bfloat162_rn(a0 * inv, a1 * inv);
  }
}

// ------------------------------------------------------------------------------------------------- col2im (conv dgrad)
// Adjoint of mpl_im2col_nhwc: dcols bf16 [B*Ho*Wo, kh*kw*C] (column (ky*kw+kx)*C+c) -> dx bf16 [B,H,W,C];
// dx[b,y,x,c] = sum over taps (ky,kx) and outputs (oy,ox) with oy*stride - pad + ky == y, ox*stride - pad + kx == x.
// Gather form, one CTA per input pixel.
__global__ void __launch_bounds__(128) col2im_nhwc_kernel(const bf16_t* __restrict__ dcols, bf16_t* __restrict__ dx, int H,
                                                          int W, int C, int kh, int kw, int stride, int pad, int Ho,
                                                          int Wo) {
  const int r = blockIdx.x;
  const int b = r / (H * W), p = r % (H * W);
  const int y = p / W, x = p % W;
  for (int c = threadIdx.x * 2; c < C; c += 256) {
    float a0 = 0.f, a1 = 0.f;
    for (int ky = 0; ky < kh; ++ky) {
      const int ny = y + pad - ky;
      if (ny < 0 || ny % stride != 0 || ny / stride >= Ho) continue;
      for (int kx = 0; kx < kw; ++kx) {
        const int nx = x + pad - kx;
        if (nx < 0 || nx % stride != 0 || nx / stride >= Wo) continue;
        const long long row = (static_cast<long long>(b) * Ho + ny / stride) * Wo + nx / stride;
        const __nv_bfloat162 v =
            *reinterpret_cast<const __nv_bfloat162*>(dcols + row * kh * kw * C + (ky * kw + kx) * C + c);
        a0 += __bfloat162float(v.x);
        a1 += __bfloat162float(v.y);
      }
    }
    *reinterpret_cast<__nv_bfloat162*>(dx + static_cast<long long>(r) * C + c) = __floats2bfloat162_rn(a0, a1);
  }
}

}  // namespace mpl

using mpl::bf16_t;

extern "C" int mpl_mask_scale_bf16(const void* x, const unsigned char* mask, float scale, void* out, int accumulate,
                                   long long n, void* stream) {
  if (n <= 0) return MPL_OK;
  if (x == nullptr || mask == nullptr || out == nullptr) return MPL_ERR_ARG;
  if ((n % 8) != 0 || (reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(out) & 15) ||
      (reinterpret_cast<uintptr_t>(mask) & 7))
    return MPL_ERR_ALIGN;
  mpl::mask_scale_kernel<<<static_cast<unsigned>((n / 8 + 255) / 256), 256, 0, ST(stream)>>>(
      static_cast<const bf16_t*>(x), mask, scale, static_cast<bf16_t*>(out), accumulate, n);
  return mpl::launch_status();
}

extern "C" int mpl_gemm_small(const void* A, int a_is_f32, long long a_stride_m, long long a_stride_k, const void* B,
                              int b_is_f32, long long b_stride_k, long long b_stride_n, void* C, int c_is_f32,
                              long long ldc, int accumulate, int M, int N, int K, void* stream) {
  if (M <= 0 || N <= 0) return MPL_OK;
  if (A == nullptr || B == nullptr || C == nullptr || K <= 0) return MPL_ERR_ARG;
  if (accumulate && !c_is_f32) return MPL_ERR_ARG;
  dim3 grid((N + 31) / 32, (M + 31) / 32, 1);
  int k_per_split = (K + 31) / 32 * 32;
  if (accumulate && c_is_f32 && K >= 512) {
    // few output tiles and a long reduction (weight gradients over thousands of rows): split K so the launch fills
    // the machine
    const long long tiles = static_cast<long long>(grid.x) * grid.y;
    long long want = (2LL * mpl::num_sms() + tiles - 1) / tiles;
    const long long max_split = K / 256;
    if (want > max_split) want = max_split;
    if (want > 1) {
      k_per_split = static_cast<int>(((K + want - 1) / want + 31) / 32 * 32);
      grid.z = static_cast<unsigned>((K + k_per_split - 1) / k_per_split);
    }
  }
  mpl::gemm_small_kernel<<<grid, 256, 0, ST(stream)>>>(A, a_is_f32, a_stride_m, a_stride_k, B, b_is_f32, b_stride_k,
                                                       b_stride_n, C, c_is_f32, ldc, accumulate, M, N, K, k_per_split);
  return mpl::launch_status();
}

extern "C" int mpl_col_sum(const void* X, int x_is_f32, long long ld, float* out, int M, int N, void* stream) {
  if (M <= 0 || N <= 0) return MPL_OK;
  if (X == nullptr || out == nullptr) return MPL_ERR_ARG;
  mpl::col_sum_kernel<<<dim3((N + 255) / 256, (M + mpl::CS_ROWS - 1) / mpl::CS_ROWS), 256, 0, ST(stream)>>>(
      X, x_is_f32, ld, out, M, N);
  return mpl::launch_status();
}

extern "C" int mpl_layernorm_bwd(const void* x, long long ldx, const void* weight, const void* dy, long long lddy,
                                 void* dx, long long lddx, float* dweight, float* dbias, int rows, int D, float eps,
                                 void* stream) {
  if (rows <= 0) return MPL_OK;
  if (x == nullptr || weight == nullptr || dy == nullptr || dx == nullptr || D <= 0) return MPL_ERR_ARG;
  mpl::layernorm_bwd_kernel<<<(rows + 7) / 8, 256, 0, ST(stream)>>>(
      static_cast<const bf16_t*>(x), ldx, static_cast<const bf16_t*>(weight), static_cast<const bf16_t*>(dy), lddy,
      static_cast<bf16_t*>(dx), lddx, dweight, dbias, rows, D, eps);
  return mpl::launch_status();
}

extern "C" int mpl_act_fwd(const void* x, void* y, long long n, int act, void* stream) {
  if (n <= 0) return MPL_OK;
  if (x == nullptr || y == nullptr) return MPL_ERR_ARG;
  mpl::act_fwd_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, ST(stream)>>>(
      static_cast<const bf16_t*>(x), static_cast<bf16_t*>(y), n, act);
  return mpl::launch_status();
}
extern "C" int mpl_act_bwd(const void* x, const void* dy, void* dx, long long n, int act, void* stream) {
  if (n <= 0) return MPL_OK;
  if (x == nullptr || dy == nullptr || dx == nullptr) return MPL_ERR_ARG;
  mpl::act_bwd_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, ST(stream)>>>(
      static_cast<const bf16_t*>(x), static_cast<const bf16_t*>(dy), static_cast<bf16_t*>(dx), n, act);
  return mpl::launch_status();
}

extern "C" int mpl_attn_small_bwd(const void* q, long long ldq, const void* k, long long ldk, const void* v,
                                  long long ldv, const void* d_o, long long ldo, void* dq, long long lddq, void* dk,
                                  long long lddk, void* dv, long long lddv, int batch, int Tq, int Tk, int H, int head_dim,
                                  float scale, void* stream) {
  if (Tq <= 0 || Tk <= 0 || H <= 0 || batch <= 0) return MPL_OK;
  if (q == nullptr || k == nullptr || v == nullptr || d_o == nullptr || dq == nullptr || dk == nullptr ||
      dv == nullptr)
    return MPL_ERR_ARG;
  if (head_dim <= 0 || head_dim > 32) return MPL_ERR_UNSUPPORTED;
  const size_t smem = (static_cast<size_t>(4) * Tk * (head_dim + 1) + 8 * (2 * static_cast<size_t>(Tk) + 64)) * 4;
  if (smem > 200 * 1024) return MPL_ERR_UNSUPPORTED;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(mpl::attn_small_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) !=
        cudaSuccess)
      return MPL_ERR_CUDA;
    attr_set = true;
  }
  mpl::attn_small_bwd_kernel<<<dim3(H, batch), 256, smem, ST(stream)>>>(
      static_cast<const bf16_t*>(q), ldq, static_cast<const bf16_t*>(k), ldk, static_cast<const bf16_t*>(v), ldv,
      static_cast<const bf16_t*>(d_o), ldo, static_cast<bf16_t*>(dq), lddq, static_cast<bf16_t*>(dk), lddk,
      static_cast<bf16_t*>(dv), lddv, Tq, Tk, head_dim, scale);
  return mpl::launch_status();
}

extern "C" int mpl_bilinear_resize_bwd(const void* dy, int dy_is_f32, int Hout, int Wout, void* dx, long long dx_stride_n,
                                       long long dx_stride_y, int Hin, int Win, int N, void* stream) {
  if (N <= 0) return MPL_OK;
  if (dy == nullptr || dx == nullptr || Hin <= 0 || Win <= 0 || Hout <= 0 || Wout <= 0) return MPL_ERR_ARG;
  const long long total = static_cast<long long>(N) * Hin * Win;
  mpl::bilinear_bwd_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, ST(stream)>>>(
      dy, dy_is_f32, Hout, Wout, static_cast<bf16_t*>(dx), dx_stride_n, dx_stride_y, Hin, Win, N);
  return mpl::launch_status();
}

extern "C" int mpl_mask_losses_bwd(const void* pred, const float* gt, const void* pred_iou, const float* sums6,
                                   const float* dloss4, long long n, void* dpred, float* dpred_iou, void* stream) {
  if (n <= 0) return MPL_OK;
  if (pred == nullptr || gt == nullptr || pred_iou == nullptr || sums6 == nullptr || dloss4 == nullptr ||
      dpred == nullptr)
    return MPL_ERR_ARG;
  mpl::mask_losses_bwd_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, ST(stream)>>>(
      static_cast<const bf16_t*>(pred), gt, static_cast<const bf16_t*>(pred_iou), sums6, dloss4, n,
      static_cast<bf16_t*>(dpred), dpred_iou);
  return mpl::launch_status();
}

extern "C" int mpl_token_pool_bwd(const void* dy, void* dx, int n, int t_in, int t_out, int D, void* stream) {
  if (n <= 0) return MPL_OK;
  if (dy == nullptr || dx == nullptr || t_in <= 0 || t_out <= 0) return MPL_ERR_ARG;
  if (D % 2 != 0) return MPL_ERR_ALIGN;
  mpl::token_pool_bwd_kernel<<<n * t_in, 128, 0, ST(stream)>>>(static_cast<const mpl::bf16_t*>(dy),
                                                               static_cast<mpl::bf16_t*>(dx), t_in, t_out, D);
  return mpl::launch_status();
}

extern "C" int mpl_col2im_nhwc(const void* dcols, void* dx, int B, int H, int W, int C, int kh, int kw, int stride, int pad,
                               void* stream) {
  if (B <= 0) return MPL_OK;
  if (dcols == nullptr || dx == nullptr || kh <= 0 || kw <= 0 || stride <= 0) return MPL_ERR_ARG;
  if (C % 2 != 0) return MPL_ERR_ALIGN;
  const int Ho = (H + 2 * pad - kh) / stride + 1, Wo = (W + 2 * pad - kw) / stride + 1;
  mpl::col2im_nhwc_kernel<<<B * H * W, 128, 0, ST(stream)>>>(static_cast<const mpl::bf16_t*>(dcols),
                                                             static_cast<mpl::bf16_t*>(dx), H, W, C, kh, kw, stride, pad, Ho,
                                                             Wo);
  return mpl::launch_status();
}

extern "C" int mpl_token_pool(const void* x, void* y, int n, int t_in, int t_out, int D, void* stream) {
  if (n <= 0) return MPL_OK;
  if (x == nullptr || y == nullptr || t_in <= 0 || t_out <= 0) return MPL_ERR_ARG;
  if (D % 2 != 0) return MPL_ERR_ALIGN;
  mpl::token_pool_kernel<<<n * t_out, 128, 0, ST(stream)>>>(static_cast<const mpl::bf16_t*>(x), static_cast<mpl::bf16_t*>(y),
                                                            t_in, t_out, D);
  return mpl::launch_status();
}
