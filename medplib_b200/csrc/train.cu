// Train-step kernels (SURVEY.md §8 a-15 losses, a-17 train step): everything the backward pass of the LLaMA-MoE stack
// with LoRA adapters needs besides the tcgen05 GEMM (which also runs every dX = dY·W contraction, against transposed
// weight copies kept resident in HBM), plus the fused cross-entropy, the mask losses and the AdamW update.
//   reference: train_ds_medplib.py:599-625 (DeepSpeed engine backward/step), peft 0.10 LoRA Linear
//   (y = W x + (alpha/r) B A x), HF-4.31 LlamaDecoderLayer autograd, deepspeed.moe top1gating autograd,
//   model/MedPLIB.py:26-124 (mask losses), medplib_moe_llama.py:399-421 (shifted CE).
// All HBM-bound except attention backward (mma.sync via wmma fragments; tcgen05 version is future work).
#include <mma.h>

#include <cstdlib>

#include "internal.h"
#include "ptx.cuh"

namespace mpl {

using bf16_t = __nv_bfloat16;

template <int THREADS>
__device__ __forceinline__ float block_sum_t(float v, float* red) {
  v = warp_sum(v);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = (lane < THREADS / 32) ? red[lane] : 0.0f;
  return warp_sum(t);
}
template <int THREADS>
__device__ __forceinline__ float block_max_t(float v, float* red) {
  v = warp_max(v);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = (lane < THREADS / 32) ? red[lane] : -INFINITY;
  return warp_max(t);
}

// ------------------------------------------------------------------------------------------------- transpose
// out[c, r] = in[r, c]; 64x64 tiles through shared memory (scalar 2-byte accesses: any alignment / leading dim).
__global__ void __launch_bounds__(256) transpose_kernel(const bf16_t* __restrict__ in, long long ldi,
                                                        bf16_t* __restrict__ out, long long ldo, int rows, int cols) {
  __shared__ bf16_t tile[64][66];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const long long r0 = static_cast<long long>(blockIdx.y) * 64, c0 = static_cast<long long>(blockIdx.x) * 64;
  for (int i = ty; i < 64; i += 8) {
    const long long r = r0 + i;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const long long c = c0 + tx + 32 * h;
      tile[i][tx + 32 * h] = (r < rows && c < cols) ? in[r * ldi + c] : __float2bfloat16(0.0f);
    }
  }
  __syncthreads();
  for (int i = ty; i < 64; i += 8) {
    const long long c = c0 + i;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const long long r = r0 + tx + 32 * h;
      if (c < cols && r < rows) out[c * ldo + r] = tile[tx + 32 * h][i];
    }
  }
}

// ------------------------------------------------------------------------------------------------- LoRA rank-r ops
constexpr int LORA_MAX_R = 16;

// u[m, j] = scale * sum_k x[m,k] * A[j,k]   (one warp per row; A is served by L1/L2)
__global__ void __launch_bounds__(256) lora_down_kernel(const bf16_t* __restrict__ x, long long ldx,
                                                        const bf16_t* __restrict__ A, long long lda, void* __restrict__ u,
                                                        int u_f32, int M, int K, int r, float scale) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long m = static_cast<long long>(blockIdx.x) * 8 + warp;
  if (m >= M) return;
  float acc[LORA_MAX_R];
#pragma unroll
  for (int j = 0; j < LORA_MAX_R; ++j) acc[j] = 0.0f;
  const bf16_t* xr = x + m * ldx;
  for (int c = lane * 8; c < K; c += 256) {
    const uint4 raw = *reinterpret_cast<const uint4*>(xr + c);
    const __nv_bfloat162* xp = reinterpret_cast<const __nv_bfloat162*>(&raw);
    float xv[8];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 f = __bfloat1622float2(xp[i]);
      xv[2 * i] = f.x;
      xv[2 * i + 1] = f.y;
    }
#pragma unroll
    for (int j = 0; j < LORA_MAX_R; ++j) {
      if (j < r) {
        const uint4 ar = *reinterpret_cast<const uint4*>(A + static_cast<long long>(j) * lda + c);
        const __nv_bfloat162* ap = reinterpret_cast<const __nv_bfloat162*>(&ar);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 f = __bfloat1622float2(ap[i]);
          acc[j] += xv[2 * i] * f.x + xv[2 * i + 1] * f.y;
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < LORA_MAX_R; ++j) acc[j] = warp_sum(acc[j]);
  if (lane == 0) {
    for (int j = 0; j < r; ++j) {
      if (u_f32)
        static_cast<float*>(u)[m * r + j] = acc[j] * scale;
      else
        static_cast<bf16_t*>(u)[m * r + j] = __float2bfloat16_rn(acc[j] * scale);
    }
  }
}

// y[m,n] = bf16(y[m,n] + bf16(scale * bf16(sum_j u[m,j] * Bm[n*sn + j*sr])))   (in place; 2 columns per thread)
constexpr int UP_ROWS = 32;
__global__ void __launch_bounds__(256) lora_up_add_kernel(bf16_t* __restrict__ y, long long ldy,
                                                          const void* __restrict__ u, int u_f32,
                                                          const bf16_t* __restrict__ Bm, long long sn, long long sr,
                                                          float scale, int M, int N, int r) {
  __shared__ float us[UP_ROWS][LORA_MAX_R];
  const int m0 = blockIdx.y * UP_ROWS;
  for (int i = threadIdx.x; i < UP_ROWS * LORA_MAX_R; i += 256) {
    const int rr = i / LORA_MAX_R, j = i % LORA_MAX_R;
    const long long m = m0 + rr;
    float v = 0.0f;
    if (m < M && j < r)
      v = u_f32 ? static_cast<const float*>(u)[m * r + j] : __bfloat162float(static_cast<const bf16_t*>(u)[m * r + j]);
    us[rr][j] = v;
  }
  __syncthreads();
  const long long n = (static_cast<long long>(blockIdx.x) * 256 + threadIdx.x) * 2;
  if (n >= N) return;
  const bool two = n + 1 < N;
  float b0[LORA_MAX_R], b1[LORA_MAX_R];
#pragma unroll
  for (int j = 0; j < LORA_MAX_R; ++j) {
    b0[j] = j < r ? __bfloat162float(Bm[n * sn + j * sr]) : 0.0f;
    b1[j] = (j < r && two) ? __bfloat162float(Bm[(n + 1) * sn + j * sr]) : 0.0f;
  }
  const int rows = min(UP_ROWS, M - m0);
  for (int rr = 0; rr < rows; ++rr) {
    float a0 = 0.0f, a1 = 0.0f;
#pragma unroll
    for (int j = 0; j < LORA_MAX_R; ++j) {
      a0 += us[rr][j] * b0[j];
      a1 += us[rr][j] * b1[j];
    }
    bf16_t* yp = y + static_cast<long long>(m0 + rr) * ldy + n;
    const float y0 = __bfloat162float(yp[0]) + bf16_round(scale * bf16_round(a0));
    yp[0] = __float2bfloat16_rn(y0);
    if (two) {
      const float y1 = __bfloat162float(yp[1]) + bf16_round(scale * bf16_round(a1));
      yp[1] = __float2bfloat16_rn(y1);
    }
  }
}

// out[n*sn + j*sr] += scale * sum_m X[m,n] * U[m,j]   (fp32 atomics; grid.y splits M)
constexpr int WG_ROWS = 32;
__global__ void __launch_bounds__(256) rank_wgrad_kernel(const bf16_t* __restrict__ X, long long ldx,
                                                         const void* __restrict__ U, int u_f32, float* __restrict__ out,
                                                         long long sn, long long sr, float scale, int M, int N, int r,
                                                         int rows_per_split) {
  __shared__ float us[WG_ROWS][LORA_MAX_R];
  const long long n = (static_cast<long long>(blockIdx.x) * 256 + threadIdx.x) * 2;
  const bool ok0 = n < N, ok1 = n + 1 < N;
  float a0[LORA_MAX_R], a1[LORA_MAX_R];
#pragma unroll
  for (int j = 0; j < LORA_MAX_R; ++j) a0[j] = a1[j] = 0.0f;
  const int mbeg = blockIdx.y * rows_per_split;
  const int mend = min(M, mbeg + rows_per_split);
  for (int m0 = mbeg; m0 < mend; m0 += WG_ROWS) {
    __syncthreads();
    for (int i = threadIdx.x; i < WG_ROWS * LORA_MAX_R; i += 256) {
      const int rr = i / LORA_MAX_R, j = i % LORA_MAX_R;
      const long long m = m0 + rr;
      float v = 0.0f;
      if (m < mend && j < r)
        v = u_f32 ? static_cast<const float*>(U)[m * r + j] : __bfloat162float(static_cast<const bf16_t*>(U)[m * r + j]);
      us[rr][j] = v;
    }
    __syncthreads();
    const int rows = min(WG_ROWS, mend - m0);
    for (int rr = 0; rr < rows; ++rr) {
      const bf16_t* xp = X + static_cast<long long>(m0 + rr) * ldx + n;
      const float x0 = ok0 ? __bfloat162float(xp[0]) : 0.0f;
      const float x1 = ok1 ? __bfloat162float(xp[1]) : 0.0f;
#pragma unroll
      for (int j = 0; j < LORA_MAX_R; ++j) {
        a0[j] += x0 * us[rr][j];
        a1[j] += x1 * us[rr][j];
      }
    }
  }
#pragma unroll
  for (int j = 0; j < LORA_MAX_R; ++j) {
    if (j < r) {
      if (ok0) atomicAdd(out + n * sn + j * sr, scale * a0[j]);
      if (ok1) atomicAdd(out + (n + 1) * sn + j * sr, scale * a1[j]);
    }
  }
}


// ---- bandwidth-shaped variants for r <= 8 with 16-byte aligned rows (the train step's shapes). All three are pure
// HBM streams (x / y / dy read once): every lane moves 16 bytes per access, >= 4 accesses in flight per lane, and the
// rank-r factors live in shared memory / registers.
constexpr int LR8 = 8;
constexpr int LD8_THREADS = 512;
// u[m, j] = scale * sum_k x[m,k] A[j,k]: A (r x K) staged ONCE per CTA in shared memory (r*K*2 <= 200 KB), one CTA
// per SM, every warp sweeps FOUR rows at a time so each staged A chunk is read once per four rows.
__global__ void __launch_bounds__(LD8_THREADS) lora_down8_kernel(const bf16_t* __restrict__ x, long long ldx,
                                                                 const bf16_t* __restrict__ A, long long lda,
                                                                 void* __restrict__ u, int u_f32, int M, int K, int r,
                                                                 float scale, int rows_per_cta) {
  extern __shared__ __align__(16) unsigned char ld8_smem[];
  bf16_t* sA = reinterpret_cast<bf16_t*>(ld8_smem);  // [r][K] bf16
  for (int i = threadIdx.x * 8; i < r * K; i += LD8_THREADS * 8) {
    const int j = i / K, c = i % K;
    *reinterpret_cast<uint4*>(sA + i) = *reinterpret_cast<const uint4*>(A + static_cast<long long>(j) * lda + c);
  }
  __syncthreads();
  constexpr int NW = LD8_THREADS / 32;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long m_beg = static_cast<long long>(blockIdx.x) * rows_per_cta;
  const long long m_end = m_beg + rows_per_cta < M ? m_beg + rows_per_cta : M;
  for (long long m = m_beg + warp; m < m_end; m += 4 * NW) {
    float acc[4][LR8];
    const bf16_t* xr[4];
    bool ok[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      ok[q] = m + q * NW < m_end;
      xr[q] = x + (ok[q] ? m + q * NW : m) * ldx;
#pragma unroll
      for (int j = 0; j < LR8; ++j) acc[q][j] = 0.0f;
    }
    for (int c = lane * 8; c < K; c += 256) {
      uint4 raw[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) raw[q] = ld_nc_v4(xr[q] + c);
#pragma unroll
      for (int j = 0; j < LR8; ++j) {
        if (j < r) {
          const uint4 ar = *reinterpret_cast<const uint4*>(sA + j * K + c);
          const __nv_bfloat162* ap = reinterpret_cast<const __nv_bfloat162*>(&ar);
          float2 f[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) f[i] = __bfloat1622float2(ap[i]);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const __nv_bfloat162* xp = reinterpret_cast<const __nv_bfloat162*>(&raw[q]);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float2 xv = __bfloat1622float2(xp[i]);
              acc[q][j] = fmaf(xv.x, f[i].x, acc[q][j]);
              acc[q][j] = fmaf(xv.y, f[i].y, acc[q][j]);
            }
          }
        }
      }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
      for (int j = 0; j < LR8; ++j) acc[q][j] = warp_sum(acc[q][j]);
    if (lane == 0) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (!ok[q]) continue;
        const long long mm = m + q * NW;
        for (int j = 0; j < r; ++j) {
          if (u_f32)
            static_cast<float*>(u)[mm * r + j] = acc[q][j] * scale;
          else
            static_cast<bf16_t*>(u)[mm * r + j] = __float2bfloat16_rn(acc[q][j] * scale);
        }
      }
    }
  }
}

// y[m, n..n+8) += rank-r update; CTA = 2048 columns x 16 rows, 8 rows of loads in flight per lane.
constexpr int UP8_ROWS = 16;
__global__ void __launch_bounds__(256) lora_up_add8_kernel(bf16_t* __restrict__ y, long long ldy,
                                                           const void* __restrict__ u, int u_f32,
                                                           const bf16_t* __restrict__ Bm, long long sn, long long sr,
                                                           float scale, int M, int N, int r) {
  __shared__ float us[UP8_ROWS][LR8];
  const int m0 = blockIdx.y * UP8_ROWS;
  if (threadIdx.x < UP8_ROWS * LR8) {
    const int rr = threadIdx.x / LR8, j = threadIdx.x % LR8;
    const long long m = m0 + rr;
    float v = 0.0f;
    if (m < M && j < r)
      v = u_f32 ? static_cast<const float*>(u)[m * r + j] : __bfloat162float(static_cast<const bf16_t*>(u)[m * r + j]);
    us[rr][j] = v;
  }
  __syncthreads();
  const long long n = (static_cast<long long>(blockIdx.x) * 256 + threadIdx.x) * 8;
  if (n >= N) return;
  float b[8][LR8];  // b[i][j] = Bm[(n+i)*sn + j*sr]
  if (sr == 1 && sn == 8 && r == 8) {  // [N, 8] row-major: one 16-byte load per column
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const uint4 raw = *reinterpret_cast<const uint4*>(Bm + (n + i) * 8);
      const bf16_t* bv = reinterpret_cast<const bf16_t*>(&raw);
#pragma unroll
      for (int j = 0; j < LR8; ++j) b[i][j] = __bfloat162float(bv[j]);
    }
  } else if (sn == 1 && (sr % 8) == 0) {  // [r, N] row-major: one 16-byte load per rank row
#pragma unroll
    for (int j = 0; j < LR8; ++j) {
      uint4 raw = make_uint4(0, 0, 0, 0);
      if (j < r) raw = *reinterpret_cast<const uint4*>(Bm + j * sr + n);
      const bf16_t* bv = reinterpret_cast<const bf16_t*>(&raw);
#pragma unroll
      for (int i = 0; i < 8; ++i) b[i][j] = __bfloat162float(bv[i]);
    }
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < LR8; ++j) b[i][j] = j < r ? __bfloat162float(Bm[(n + i) * sn + j * sr]) : 0.0f;
  }
  const int rows = min(UP8_ROWS, M - m0);
  bf16_t* y0 = y + static_cast<long long>(m0) * ldy + n;
#pragma unroll
  for (int half = 0; half < UP8_ROWS / 8; ++half) {
    uint4 raw[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int rr = half * 8 + q;
      raw[q] = rr < rows ? *reinterpret_cast<const uint4*>(y0 + static_cast<long long>(rr) * ldy) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int rr = half * 8 + q;
      if (rr >= rows) continue;
      bf16_t* yv = reinterpret_cast<bf16_t*>(&raw[q]);
      float uu[LR8];
#pragma unroll
      for (int j = 0; j < LR8; ++j) uu[j] = us[rr][j];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float a = 0.0f;
#pragma unroll
        for (int j = 0; j < LR8; ++j) a = fmaf(uu[j], b[i][j], a);
        yv[i] = __float2bfloat16_rn(__bfloat162float(yv[i]) + bf16_round(scale * bf16_round(a)));
      }
      *reinterpret_cast<uint4*>(y0 + static_cast<long long>(rr) * ldy) = raw[q];
    }
  }
}

// out[n, j] += scale * sum_m X[m, n] U[m, j]: CTA = 256 columns (lane x 8) x WG8_ROWS rows (warp w takes rows w, w+8,
// ...), warps combine through shared-memory atomics, then ONE global atomic per (column, j) per CTA.
constexpr int WG8_ROWS = 256;
__global__ void __launch_bounds__(256) rank_wgrad8_kernel(const bf16_t* __restrict__ X, long long ldx,
                                                          const void* __restrict__ U, int u_f32, float* __restrict__ out,
                                                          long long sn, long long sr, float scale, int M, int N, int r) {
  __shared__ float us[WG8_ROWS][LR8];
  __shared__ float sacc[256][LR8 + 1];
  const int mbeg = blockIdx.y * WG8_ROWS;
  const int rows = min(WG8_ROWS, M - mbeg);
  for (int i = threadIdx.x; i < WG8_ROWS * LR8; i += 256) {
    const int rr = i / LR8, j = i % LR8;
    const long long m = mbeg + rr;
    float v = 0.0f;
    if (rr < rows && j < r)
      v = u_f32 ? static_cast<const float*>(U)[m * r + j] : __bfloat162float(static_cast<const bf16_t*>(U)[m * r + j]);
    us[rr][j] = v;
  }
  for (int i = threadIdx.x; i < 256 * (LR8 + 1); i += 256) (&sacc[0][0])[i] = 0.0f;
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long n = static_cast<long long>(blockIdx.x) * 256 + lane * 8;
  if (n < N) {
    float acc[8][LR8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < LR8; ++j) acc[i][j] = 0.0f;
    const bf16_t* xp = X + static_cast<long long>(mbeg) * ldx + n;
    for (int rr = warp; rr < rows; rr += 32) {
      uint4 raw[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int row = rr + 8 * q;
        raw[q] = row < rows ? ld_nc_v4(xp + static_cast<long long>(row) * ldx) : make_uint4(0, 0, 0, 0);
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int row = min(rr + 8 * q, WG8_ROWS - 1);
        const bf16_t* xv = reinterpret_cast<const bf16_t*>(&raw[q]);
        float uu[LR8];
#pragma unroll
        for (int j = 0; j < LR8; ++j) uu[j] = us[row][j];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float xf = __bfloat162float(xv[i]);
#pragma unroll
          for (int j = 0; j < LR8; ++j) acc[i][j] = fmaf(xf, uu[j], acc[i][j]);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < LR8; ++j) atomicAdd(&sacc[lane * 8 + i][j], acc[i][j]);
  }
  __syncthreads();
  const long long nc = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
  if (nc < N) {
#pragma unroll
    for (int j = 0; j < LR8; ++j)
      if (j < r) atomicAdd(out + nc * sn + j * sr, scale * sacc[threadIdx.x][j]);
  }
}

// ---- mma.sync (m16n8k16, bf16 -> fp32) variants of the three rank-r (r <= 8) operators. The rank dimension maps onto
// the 8-wide N (or the zero-padded 16-wide M) of the warp-level MMA, which removes ~95 % of the issued instructions of
// the FMA versions above and leaves these kernels as pure 16-byte-per-lane HBM streams. The contraction index may be
// permuted freely as long as both operands use the same permutation, so every lane feeds fragments straight from the
// 16-byte chunk it loaded (no shared-memory transposes).
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                               uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t u32_of(const uint4& v, int i) {
  return i == 0 ? v.x : (i == 1 ? v.y : (i == 2 ? v.z : v.w));
}

// u[m, j] = scale * sum_k x[m,k] A[j,k].  CTA = 16 rows x 8 K-splits (one warp each); K % 32 == 0.
// Lane (g = lane/4, t = lane%4) loads cols [32 blk + 8t, +8) of rows g and g+8 of x, and of row g of A (zero if g >= r).
__global__ void __launch_bounds__(256) lora_down_mma_kernel(const bf16_t* __restrict__ x, long long ldx,
                                                            const bf16_t* __restrict__ A, long long lda,
                                                            void* __restrict__ u, int u_f32, int M, int K, int r,
                                                            float scale, bf16_t* __restrict__ u_pad, int pad_col) {
  __shared__ float red[8][16][8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const long long m0 = static_cast<long long>(blockIdx.x) * 16;
  const bool ok0 = m0 + g < M, ok1 = m0 + g + 8 < M, okA = g < r;
  const bf16_t* x0 = x + (ok0 ? m0 + g : 0) * ldx + 8 * t;
  const bf16_t* x1 = x + (ok1 ? m0 + g + 8 : 0) * ldx + 8 * t;
  const bf16_t* al = A + static_cast<long long>(okA ? g : 0) * lda + 8 * t;
  const uint4 zero = make_uint4(0, 0, 0, 0);
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  const int nblk = K / 32;
  int blk = warp;
  for (; blk + 24 < nblk; blk += 32) {  // 4 blocks (stride 8) per pass: 8 x-loads in flight per lane
    uint4 xa[4], xb[4], av[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int c = (blk + 8 * q) * 32;
      xa[q] = ok0 ? ld_nc_v4(x0 + c) : zero;
      xb[q] = ok1 ? ld_nc_v4(x1 + c) : zero;
      av[q] = okA ? *reinterpret_cast<const uint4*>(al + c) : zero;
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      mma_bf16_16816(acc, xa[q].x, xb[q].x, xa[q].y, xb[q].y, av[q].x, av[q].y);
      mma_bf16_16816(acc, xa[q].z, xb[q].z, xa[q].w, xb[q].w, av[q].z, av[q].w);
    }
  }
  for (; blk < nblk; blk += 8) {
    const int c = blk * 32;
    const uint4 xa = ok0 ? ld_nc_v4(x0 + c) : zero;
    const uint4 xb = ok1 ? ld_nc_v4(x1 + c) : zero;
    const uint4 av = okA ? *reinterpret_cast<const uint4*>(al + c) : zero;
    mma_bf16_16816(acc, xa.x, xb.x, xa.y, xb.y, av.x, av.y);
    mma_bf16_16816(acc, xa.z, xb.z, xa.w, xb.w, av.z, av.w);
  }
  red[warp][g][2 * t] = acc[0];
  red[warp][g][2 * t + 1] = acc[1];
  red[warp][g + 8][2 * t] = acc[2];
  red[warp][g + 8][2 * t + 1] = acc[3];
  __syncthreads();
  if (threadIdx.x < 128) {
    const int row = threadIdx.x >> 3, j = threadIdx.x & 7;
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) v += red[w][row][j];
    const long long m = m0 + row;
    if (m < M && j < r) {
      if (u_f32)
        static_cast<float*>(u)[m * r + j] = v * scale;
      else
        static_cast<bf16_t*>(u)[m * r + j] = __float2bfloat16_rn(v * scale);
      // second copy for the GEMM's extension k-block: columns [pad_col, pad_col + r) of a [M, 64] bf16 operand
      if (u_pad != nullptr) u_pad[m * 64 + pad_col + j] = __float2bfloat16_rn(v * scale);
    }
  }
}

// Extension-operand rows of one adapter: dst[n, col + j] = bf16(scale * src[n * sn + j * sr]) for j < r (dst bf16 [N, 64],
// zero elsewhere -- written once at allocation). One launch packs EVERY adapter of the model (device-side item table).
struct LoraPackItem {
  const bf16_t* src;
  bf16_t* dst;
  long long sn, sr;
  int N, r, col;
  float scale;
  long long dn, dj;  // destination strides (64, 1 for an extension operand; 1, N for a plain transposed copy [r, N])
};
__global__ void __launch_bounds__(256) lora_pack_kernel(const LoraPackItem* __restrict__ items) {
  const LoraPackItem it = items[blockIdx.y];
  for (long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x; i < static_cast<long long>(it.N) * it.r;
       i += static_cast<long long>(gridDim.x) * 256) {
    const long long n = i / it.r;
    const int j = static_cast<int>(i % it.r);
    it.dst[n * it.dn + (it.col + j) * it.dj] = __float2bfloat16_rn(it.scale * __bfloat162float(it.src[n * it.sn + j * it.sr]));
  }
}

// y[m,n] = bf16(y + bf16(scale * bf16(sum_j u[m,j] Bm[n*sn + j*sr]))).  CTA = 256 columns (warp: 32) x UPM_ROWS rows.
constexpr int UPM_ROWS = 64;
__global__ void __launch_bounds__(256) lora_up_add_mma_kernel(bf16_t* __restrict__ y, long long ldy,
                                                              const void* __restrict__ u, int u_f32,
                                                              const bf16_t* __restrict__ Bm, long long sn, long long sr,
                                                              float scale, int M, int N, int r) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const long long col0 = static_cast<long long>(blockIdx.x) * 256 + warp * 32;
  if (col0 >= N) return;
  // B fragments of the 4 MMAs of this warp's 32 columns: n' = g <-> column col0 + 8 (g / 2) + 2 q + (g % 2)
  uint32_t bq[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const long long n = col0 + 8 * (g >> 1) + 2 * q + (g & 1);
    float lo = 0.f, hi = 0.f;
    if (n < N) {
      if (2 * t < r) lo = __bfloat162float(Bm[n * sn + (2 * t) * sr]);
      if (2 * t + 1 < r) hi = __bfloat162float(Bm[n * sn + (2 * t + 1) * sr]);
    }
    bq[q] = pack_bf16(lo, hi);
  }
  const bool col_ok = col0 + 8 * t < N;
  const long long mbeg = static_cast<long long>(blockIdx.y) * UPM_ROWS;
#pragma unroll 1
  for (int rg = 0; rg < UPM_ROWS / 16; rg += 2) {  // two 16-row groups per pass: 4 y-loads in flight per lane
    long long rows[4];
    uint4 yv[4];
    uint32_t ua[4];
#pragma unroll
    for (int h = 0; h < 4; ++h) {
      rows[h] = mbeg + (rg + (h >> 1)) * 16 + g + 8 * (h & 1);
      const bool ok = rows[h] < M;
      yv[h] = (ok && col_ok) ? *reinterpret_cast<const uint4*>(y + rows[h] * ldy + col0 + 8 * t) : make_uint4(0, 0, 0, 0);
      float lo = 0.f, hi = 0.f;
      if (ok) {
        if (u_f32) {
          const float* up = static_cast<const float*>(u) + rows[h] * r;
          if (2 * t < r) lo = up[2 * t];
          if (2 * t + 1 < r) hi = up[2 * t + 1];
        } else {
          const bf16_t* up = static_cast<const bf16_t*>(u) + rows[h] * r;
          if (2 * t < r) lo = __bfloat162float(up[2 * t]);
          if (2 * t + 1 < r) hi = __bfloat162float(up[2 * t + 1]);
        }
      }
      ua[h] = pack_bf16(lo, hi);
    }
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      float c[4][4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        c[q][0] = c[q][1] = c[q][2] = c[q][3] = 0.f;
        mma_bf16_16816(c[q], ua[2 * half], ua[2 * half + 1], 0u, 0u, bq[q], 0u);
      }
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {  // hh = 0: row g, hh = 1: row g + 8
        const int h = 2 * half + hh;
        if (rows[h] >= M || !col_ok) continue;
        bf16_t* e = reinterpret_cast<bf16_t*>(&yv[h]);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            const float a = c[q][2 * hh + i];
            e[2 * q + i] = __float2bfloat16_rn(__bfloat162float(e[2 * q + i]) + bf16_round(scale * bf16_round(a)));
          }
        }
        *reinterpret_cast<uint4*>(y + rows[h] * ldy + col0 + 8 * t) = yv[h];
      }
    }
  }
}

// out[n*sn + j*sr] += scale * sum_m X[m,n] U[m,j].  CTA = 64 columns x WGM_ROWS rows, 8 warps = 8 row ranges.
// Lane (g, t): rank row j = g; loads rows base + 4t + {0,1,2,3}, cols [col0 + 8g, +8) of X (the MMA's B operand,
// n' = g for the i-th of its 8 columns) and U[those rows][g] (the A operand, rows j >= 8 are zero padding).
constexpr int WGM_ROWS = 512;
__global__ void __launch_bounds__(256) rank_wgrad_mma_kernel(const bf16_t* __restrict__ X, long long ldx,
                                                             const void* __restrict__ U, int u_f32,
                                                             float* __restrict__ out, long long sn, long long sr,
                                                             float scale, int M, int N, int r) {
  __shared__ float red[8][8][64 + 1];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const long long col0 = static_cast<long long>(blockIdx.x) * 64;
  const bool col_ok = col0 + 8 * g < N;
  const long long rbeg = static_cast<long long>(blockIdx.y) * WGM_ROWS + warp * (WGM_ROWS / 8);
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
  const uint4 zero = make_uint4(0, 0, 0, 0);
#pragma unroll 1
  for (int step = 0; step < WGM_ROWS / 8 / 16; step += 2) {  // 2 x 16 rows per pass: 8 X-loads in flight per lane
    uint4 raw[2][4];
    uint32_t a0[2], a2[2];
#pragma unroll
    for (int s2 = 0; s2 < 2; ++s2) {
      const long long base = rbeg + (step + s2) * 16 + 4 * t;
      float uv[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const long long m = base + q;
        const bool ok = m < M;
        raw[s2][q] = (ok && col_ok) ? ld_nc_v4(X + m * ldx + col0 + 8 * g) : zero;
        uv[q] = 0.f;
        if (ok && g < r)
          uv[q] = u_f32 ? static_cast<const float*>(U)[m * r + g] : __bfloat162float(static_cast<const bf16_t*>(U)[m * r + g]);
      }
      a0[s2] = pack_bf16(uv[0], uv[1]);
      a2[s2] = pack_bf16(uv[2], uv[3]);
    }
#pragma unroll
    for (int s2 = 0; s2 < 2; ++s2) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const uint32_t sel = (i & 1) ? 0x7632u : 0x5410u;
        const uint32_t b0 = __byte_perm(u32_of(raw[s2][0], i >> 1), u32_of(raw[s2][1], i >> 1), sel);
        const uint32_t b1 = __byte_perm(u32_of(raw[s2][2], i >> 1), u32_of(raw[s2][3], i >> 1), sel);
        mma_bf16_16816(acc[i], a0[s2], 0u, a2[s2], 0u, b0, b1);
      }
    }
  }
  // acc[i][0..1] = out^T[j = g][col0 + 8 (2t) + i], [col0 + 8 (2t + 1) + i]
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    red[warp][g][16 * t + i] = acc[i][0];
    red[warp][g][16 * t + 8 + i] = acc[i][1];
  }
  __syncthreads();
  for (int o = threadIdx.x; o < 8 * 64; o += 256) {
    const int j = o >> 6, c = o & 63;
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) v += red[w][j][c];
    const long long n = col0 + c;
    if (j < r && n < N) atomicAdd(out + n * sn + j * sr, scale * v);
  }
}

// ------------------------------------------------------------------------------------------------- RMSNorm backward
// y = w * (x * rstd):  dx = rstd * (g - xhat * mean(g * xhat)), g = dy * w;  out = bf16(dx + add);  dw += dy * xhat
constexpr int RB_THREADS = 128;
__global__ void __launch_bounds__(RB_THREADS) rmsnorm_bwd_kernel(const bf16_t* __restrict__ x, long long ldx,
                                                                 const bf16_t* __restrict__ w,
                                                                 const bf16_t* __restrict__ dy, long long lddy,
                                                                 const bf16_t* __restrict__ add, long long ldadd,
                                                                 bf16_t* __restrict__ dx, long long lddx,
                                                                 float* __restrict__ dw, int D, float eps) {
  __shared__ float red[RB_THREADS / 32];
  const long long row = blockIdx.x;
  const bf16_t* xr = x + row * ldx;
  const bf16_t* gr = dy + row * lddy;
  float xv[4][8], gv[4][8];
  float ss = 0.0f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = (i * RB_THREADS + threadIdx.x) * 8;
    if (c < D) {
      const uint4 xr4 = *reinterpret_cast<const uint4*>(xr + c);
      const uint4 gr4 = *reinterpret_cast<const uint4*>(gr + c);
      const bf16_t* xp = reinterpret_cast<const bf16_t*>(&xr4);
      const bf16_t* gp = reinterpret_cast<const bf16_t*>(&gr4);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        xv[i][e] = __bfloat162float(xp[e]);
        gv[i][e] = __bfloat162float(gp[e]);  // dy; multiplied by w below
        ss += xv[i][e] * xv[i][e];
      }
    }
  }
  ss = block_sum_t<RB_THREADS>(ss, red);
  const float rstd = rsqrtf(ss / static_cast<float>(D) + eps);
  float dot = 0.0f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = (i * RB_THREADS + threadIdx.x) * 8;
    if (c < D) {
      const uint4 wr4 = *reinterpret_cast<const uint4*>(w + c);
      const bf16_t* wp = reinterpret_cast<const bf16_t*>(&wr4);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float xhat = xv[i][e] * rstd;
        if (dw != nullptr) atomicAdd(dw + c + e, gv[i][e] * xhat);
        gv[i][e] *= __bfloat162float(wp[e]);  // g = dy * w
        xv[i][e] = xhat;
        dot += gv[i][e] * xhat;
      }
    }
  }
  dot = block_sum_t<RB_THREADS>(dot, red) / static_cast<float>(D);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = (i * RB_THREADS + threadIdx.x) * 8;
    if (c < D) {
      float o[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] = rstd * (gv[i][e] - xv[i][e] * dot);
      if (add != nullptr) {
        const uint4 ar4 = *reinterpret_cast<const uint4*>(add + row * ldadd + c);
        const bf16_t* ap = reinterpret_cast<const bf16_t*>(&ar4);
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] += __bfloat162float(ap[e]);
      }
      uint4 ov;
      ov.x = pack_bf16(o[0], o[1]);
      ov.y = pack_bf16(o[2], o[3]);
      ov.z = pack_bf16(o[4], o[5]);
      ov.w = pack_bf16(o[6], o[7]);
      *reinterpret_cast<uint4*>(dx + row * lddx + c) = ov;
    }
  }
}

// ------------------------------------------------------------------------------------------------- SiLU(g) * u
__global__ void __launch_bounds__(256) silu_mul_kernel(const bf16_t* __restrict__ g, const bf16_t* __restrict__ u,
                                                       bf16_t* __restrict__ h, long long n8) {
  const long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
  if (i >= n8) return;
  const uint4 gr = reinterpret_cast<const uint4*>(g)[i], ur = reinterpret_cast<const uint4*>(u)[i];
  const bf16_t* gp = reinterpret_cast<const bf16_t*>(&gr);
  const bf16_t* up = reinterpret_cast<const bf16_t*>(&ur);
  float o[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const float gv = __bfloat162float(gp[e]);
    o[e] = bf16_round(gv / (1.0f + __expf(-gv))) * __bfloat162float(up[e]);
  }
  uint4 ov;
  ov.x = pack_bf16(o[0], o[1]);
  ov.y = pack_bf16(o[2], o[3]);
  ov.z = pack_bf16(o[4], o[5]);
  ov.w = pack_bf16(o[6], o[7]);
  reinterpret_cast<uint4*>(h)[i] = ov;
}
// dg = dh * u * sigma(g) * (1 + g * (1 - sigma(g)));  du = dh * silu(g)   (dg may alias g, du may alias u)
__global__ void __launch_bounds__(256) silu_mul_bwd_kernel(const bf16_t* g, const bf16_t* u,
                                                           const bf16_t* __restrict__ dh, bf16_t* dg, bf16_t* du,
                                                           long long n8) {
  const long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
  if (i >= n8) return;
  const uint4 gr = reinterpret_cast<const uint4*>(g)[i], ur = reinterpret_cast<const uint4*>(u)[i];
  const uint4 dr = reinterpret_cast<const uint4*>(dh)[i];
  const bf16_t* gp = reinterpret_cast<const bf16_t*>(&gr);
  const bf16_t* up = reinterpret_cast<const bf16_t*>(&ur);
  const bf16_t* dp = reinterpret_cast<const bf16_t*>(&dr);
  float og[8], ou[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const float gv = __bfloat162float(gp[e]), uv = __bfloat162float(up[e]), dv = __bfloat162float(dp[e]);
    const float sg = __fdividef(1.0f, 1.0f + __expf(-gv));  // (no IEEE slow-path call between the 8 elements)
    og[e] = dv * uv * sg * (1.0f + gv * (1.0f - sg));
    ou[e] = dv * gv * sg;
  }
  uint4 a, b;
  a.x = pack_bf16(og[0], og[1]); a.y = pack_bf16(og[2], og[3]); a.z = pack_bf16(og[4], og[5]); a.w = pack_bf16(og[6], og[7]);
  b.x = pack_bf16(ou[0], ou[1]); b.y = pack_bf16(ou[2], ou[3]); b.z = pack_bf16(ou[4], ou[5]); b.w = pack_bf16(ou[6], ou[7]);
  reinterpret_cast<uint4*>(dg)[i] = a;
  reinterpret_cast<uint4*>(du)[i] = b;
}

// ------------------------------------------------------------------------------------------------- attention backward
// delta[bh, t] = sum_d dO[b,t,h,d] * O[b,t,h,d]   (one warp per row)
__global__ void __launch_bounds__(256) attn_delta_kernel(const bf16_t* __restrict__ o, const bf16_t* __restrict__ dO,
                                                         long long sb, long long st, long long sh, float* __restrict__ delta,
                                                         int B, int H, int T, int hd) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long idx = static_cast<long long>(blockIdx.x) * 8 + warp;  // (b*H + h)*T + t
  if (idx >= static_cast<long long>(B) * H * T) return;
  const int t = idx % T;
  const long long bh = idx / T;
  const int h = bh % H;
  const long long b = bh / H;
  const long long off = b * sb + t * st + h * sh;
  float acc = 0.0f;
  for (int c = lane; c < hd; c += 32) acc += __bfloat162float(o[off + c]) * __bfloat162float(dO[off + c]);
  acc = warp_sum(acc);
  if (lane == 0) delta[idx] = acc;
}

struct AttnBwdParams {
  const bf16_t *q, *k, *v, *dO;  // q,k,v: rotated q / k and v with the forward's strides; dO with the o strides
  long long q_sb, q_st, q_sh, k_sb, k_st, k_sh, v_sb, v_st, v_sh, o_sb, o_st, o_sh;
  const float *lse, *delta;  // [B*H, T]
  float* dq;                 // f32 [B, T, H, D] contiguous, zero-initialised (atomic accumulation over key blocks)
  bf16_t *dk, *dv;           // written with the k / v strides of dk_s*, dv_s*
  long long dk_sb, dk_st, dk_sh, dv_sb, dv_st, dv_sh;
  int B, H, T;
  float scale;
  int causal;
  const unsigned char* kv_mask;
  long long kv_mask_stride;
};

constexpr int AB_BM = 64, AB_BN = 64, AB_THREADS = 256;
// One CTA per (key block, head, batch): keeps K_j, V_j and the dK_j, dV_j accumulators resident and walks the query
// blocks that can see these keys. S = Q K^T and dP = dO V^T (recomputed), P = 2^(S*scale*log2e - lse),
// dS = P * (dP - delta) * scale;  dV += P^T dO, dK += dS^T Q, dQ += dS K (fp32 atomics).
template <int D>
__global__ void __launch_bounds__(AB_THREADS) attn_bwd_kernel(const AttnBwdParams p) {
  using namespace nvcuda;
  constexpr int LDT = D + 8;    // bf16 row pitch of the [64][D] tiles
  constexpr int LDS = AB_BN + 4;  // fp32 row pitch of S / dP
  constexpr int LDP = AB_BN + 8;  // bf16 row pitch of P / dS
  extern __shared__ __align__(128) uint8_t ab_smem[];
  bf16_t* sK = reinterpret_cast<bf16_t*>(ab_smem);
  bf16_t* sV = sK + AB_BN * LDT;
  bf16_t* sQ0 = sV + AB_BN * LDT;             // Q / dO tiles are double-buffered: the next query block streams in
  bf16_t* sdO0 = sQ0 + 2 * AB_BM * LDT;       // (cp.async) while the current one is being multiplied
  float* sS = reinterpret_cast<float*>(sdO0 + 2 * AB_BM * LDT);
  float* sdP = sS + AB_BM * LDS;
  bf16_t* sP = reinterpret_cast<bf16_t*>(sdP + AB_BM * LDS);
  bf16_t* sdS = sP + AB_BM * LDP;
  float* sLse = reinterpret_cast<float*>(sdS + AB_BM * LDP);
  float* sDelta = sLse + AB_BM;
  float* sdQ = sS;  // [64][D + 4] fp32 staging of the dQ tile; reuses S + dP (+ P, dS) once they are consumed
  constexpr int LDQ = D + 4;
  static_assert(AB_BM * LDQ * 4 <= 2 * AB_BM * LDS * 4 + 2 * AB_BM * LDP * 2, "dQ staging must fit");

  const int warp = threadIdx.x >> 5;
  const int b = blockIdx.z, h = blockIdx.y;
  const int k0 = blockIdx.x * AB_BN;
  const int T = p.T;
  const bf16_t* qg = p.q + b * p.q_sb + h * p.q_sh;
  const bf16_t* kg = p.k + b * p.k_sb + h * p.k_sh;
  const bf16_t* vg = p.v + b * p.v_sb + h * p.v_sh;
  const bf16_t* og = p.dO + b * p.o_sb + h * p.o_sh;
  const long long bh = static_cast<long long>(b) * p.H + h;
  const unsigned char* mrow = p.kv_mask ? p.kv_mask + static_cast<long long>(b) * p.kv_mask_stride : nullptr;
  const float sl2 = p.scale * 1.4426950408889634f;

  auto load_tile = [&](bf16_t* dst, const bf16_t* src, long long stride_t, int t0) {
    constexpr int CPR = D / 8;
    for (int idx = threadIdx.x; idx < 64 * CPR; idx += AB_THREADS) {
      const int r = idx / CPR, c = idx % CPR;
      const int t = t0 + r;
      uint4 val = make_uint4(0, 0, 0, 0);
      if (t < T) val = *reinterpret_cast<const uint4*>(src + static_cast<long long>(t) * stride_t + c * 8);
      *reinterpret_cast<uint4*>(dst + r * LDT + c * 8) = val;
    }
  };
  auto load_tile_async = [&](bf16_t* dst, const bf16_t* src, long long stride_t, int t0) {
    constexpr int CPR = D / 8;
    for (int idx = threadIdx.x; idx < 64 * CPR; idx += AB_THREADS) {
      const int r = idx / CPR, c = idx % CPR;
      const int t = t0 + r;
      const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(dst + r * LDT + c * 8));
      const int sz = t < T ? 16 : 0;  // src-size 0: the 16 bytes are zero-filled
      const bf16_t* g = src + static_cast<long long>(t < T ? t : 0) * stride_t + c * 8;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(g), "r"(sz) : "memory");
    }
  };
  load_tile(sK, kg, p.k_st, k0);
  load_tile(sV, vg, p.v_st, k0);

  // persistent accumulators: this warp owns key rows [16*(warp/2), +16) x columns [64*(warp%2), +64) of dK and dV
  const int acc_r = (warp >> 1) * 16, acc_c = (warp & 1) * (D / 2);
  constexpr int NF = D / 32;  // 16-wide fragments per warp per matrix
  wmma::fragment<wmma::accumulator, 16, 16, 16, float> dKf[NF], dVf[NF];
#pragma unroll
  for (int i = 0; i < NF; ++i) {
    wmma::fill_fragment(dKf[i], 0.0f);
    wmma::fill_fragment(dVf[i], 0.0f);
  }

  const int q_begin = p.causal ? (k0 / AB_BM) * AB_BM : 0;
  load_tile_async(sQ0, qg, p.q_st, q_begin);
  load_tile_async(sdO0, og, p.o_st, q_begin);
  asm volatile("cp.async.commit_group;" ::: "memory");
  int buf = 0;
  for (int q0 = q_begin; q0 < T; q0 += AB_BM, buf ^= 1) {
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();  // this block's tiles have landed for every thread; previous iteration done with the other buffer
    bf16_t* sQ = sQ0 + buf * AB_BM * LDT;
    bf16_t* sdO = sdO0 + buf * AB_BM * LDT;
    if (q0 + AB_BM < T) {  // prefetch the next query block into the buffer the previous iteration used
      load_tile_async(sQ0 + (buf ^ 1) * AB_BM * LDT, qg, p.q_st, q0 + AB_BM);
      load_tile_async(sdO0 + (buf ^ 1) * AB_BM * LDT, og, p.o_st, q0 + AB_BM);
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
    if (threadIdx.x < AB_BM) {
      const int t = q0 + threadIdx.x;
      sLse[threadIdx.x] = t < T ? p.lse[bh * T + t] : INFINITY;
      sDelta[threadIdx.x] = t < T ? p.delta[bh * T + t] : 0.0f;
    }
    __syncthreads();
    // ---- S = Q K^T, dP = dO V^T : 4x4 fragments each; warp w computes fragments (w/2, 2*(w%2)) and (w/2, 2*(w%2)+1)
    {
      const int fr = (warp >> 1) * 16, fc = (warp & 1) * 32;
      wmma::fragment<wmma::accumulator, 16, 16, 16, float> s0, s1, d0, d1;
      wmma::fill_fragment(s0, 0.0f); wmma::fill_fragment(s1, 0.0f);
      wmma::fill_fragment(d0, 0.0f); wmma::fill_fragment(d1, 0.0f);
#pragma unroll
      for (int kk = 0; kk < D; kk += 16) {
        wmma::fragment<wmma::matrix_a, 16, 16, 16, bf16_t, wmma::row_major> aq, ao;
        wmma::fragment<wmma::matrix_b, 16, 16, 16, bf16_t, wmma::col_major> bk0, bk1, bv0, bv1;
        wmma::load_matrix_sync(aq, sQ + fr * LDT + kk, LDT);
        wmma::load_matrix_sync(ao, sdO + fr * LDT + kk, LDT);
        wmma::load_matrix_sync(bk0, sK + fc * LDT + kk, LDT);
        wmma::load_matrix_sync(bk1, sK + (fc + 16) * LDT + kk, LDT);
        wmma::load_matrix_sync(bv0, sV + fc * LDT + kk, LDT);
        wmma::load_matrix_sync(bv1, sV + (fc + 16) * LDT + kk, LDT);
        wmma::mma_sync(s0, aq, bk0, s0);
        wmma::mma_sync(s1, aq, bk1, s1);
        wmma::mma_sync(d0, ao, bv0, d0);
        wmma::mma_sync(d1, ao, bv1, d1);
      }
      wmma::store_matrix_sync(sS + fr * LDS + fc, s0, LDS, wmma::mem_row_major);
      wmma::store_matrix_sync(sS + fr * LDS + fc + 16, s1, LDS, wmma::mem_row_major);
      wmma::store_matrix_sync(sdP + fr * LDS + fc, d0, LDS, wmma::mem_row_major);
      wmma::store_matrix_sync(sdP + fr * LDS + fc + 16, d1, LDS, wmma::mem_row_major);
    }
    __syncthreads();
    // ---- P and dS
    for (int idx = threadIdx.x; idx < AB_BM * AB_BN; idx += AB_THREADS) {
      const int r = idx / AB_BN, c = idx % AB_BN;
      const int qi = q0 + r, kj = k0 + c;
      bool ok = qi < T && kj < T;
      if (p.causal) ok = ok && kj <= qi;
      if (ok && mrow != nullptr) ok = mrow[kj] != 0;
      float pv = 0.0f, ds = 0.0f;
      if (ok) {
        pv = exp2f(sS[r * LDS + c] * sl2 - sLse[r]);
        ds = pv * (sdP[r * LDS + c] - sDelta[r]) * p.scale;
      }
      sP[r * LDP + c] = __float2bfloat16_rn(pv);
      sdS[r * LDP + c] = __float2bfloat16_rn(ds);
    }
    __syncthreads();
    // ---- dV += P^T dO ; dK += dS^T Q   (A = P^T read column-major from sP)
#pragma unroll
    for (int kk = 0; kk < AB_BM; kk += 16) {
      wmma::fragment<wmma::matrix_a, 16, 16, 16, bf16_t, wmma::col_major> ap, ads;
      wmma::load_matrix_sync(ap, sP + kk * LDP + acc_r, LDP);
      wmma::load_matrix_sync(ads, sdS + kk * LDP + acc_r, LDP);
#pragma unroll
      for (int i = 0; i < NF; ++i) {
        wmma::fragment<wmma::matrix_b, 16, 16, 16, bf16_t, wmma::row_major> bo, bq;
        wmma::load_matrix_sync(bo, sdO + kk * LDT + acc_c + i * 16, LDT);
        wmma::load_matrix_sync(bq, sQ + kk * LDT + acc_c + i * 16, LDT);
        wmma::mma_sync(dVf[i], ap, bo, dVf[i]);
        wmma::mma_sync(dKf[i], ads, bq, dKf[i]);
      }
    }
    // ---- dQ tile = dS K : rows [16*(warp/2), +16) x columns [64*(warp%2), +64)
    wmma::fragment<wmma::accumulator, 16, 16, 16, float> dQf[NF];
#pragma unroll
    for (int i = 0; i < NF; ++i) wmma::fill_fragment(dQf[i], 0.0f);
#pragma unroll
    for (int kk = 0; kk < AB_BN; kk += 16) {
      wmma::fragment<wmma::matrix_a, 16, 16, 16, bf16_t, wmma::row_major> ads;
      wmma::load_matrix_sync(ads, sdS + acc_r * LDP + kk, LDP);
#pragma unroll
      for (int i = 0; i < NF; ++i) {
        wmma::fragment<wmma::matrix_b, 16, 16, 16, bf16_t, wmma::row_major> bk;
        wmma::load_matrix_sync(bk, sK + kk * LDT + acc_c + i * 16, LDT);
        wmma::mma_sync(dQf[i], ads, bk, dQf[i]);
      }
    }
    __syncthreads();  // every warp finished reading sP / sdS / sS / sdP before the staging overwrites them
#pragma unroll
    for (int i = 0; i < NF; ++i)
      wmma::store_matrix_sync(sdQ + acc_r * LDQ + acc_c + i * 16, dQf[i], LDQ, wmma::mem_row_major);
    __syncthreads();
    // 16-byte vector reductions (red.global.add.v4.f32, sm_90+): a quarter of the atomic instructions
    for (int idx = threadIdx.x; idx < AB_BM * (D / 4); idx += AB_THREADS) {
      const int r = idx / (D / 4), c = (idx % (D / 4)) * 4;
      const int qi = q0 + r;
      if (qi < T) {
        float4* dst = reinterpret_cast<float4*>(p.dq + ((static_cast<long long>(b) * T + qi) * p.H + h) * D + c);
        const float* sp = sdQ + r * LDQ + c;
        atomicAdd(dst, make_float4(sp[0], sp[1], sp[2], sp[3]));
      }
    }
  }
  // ---- write dK_j, dV_j (stage through shared memory as fp32)
  __syncthreads();
  float* stage = sS;  // [64][LDQ]
  bf16_t* dkg = p.dk + b * p.dk_sb + h * p.dk_sh;
  bf16_t* dvg = p.dv + b * p.dv_sb + h * p.dv_sh;
#pragma unroll
  for (int pass = 0; pass < 2; ++pass) {
#pragma unroll
    for (int i = 0; i < NF; ++i)
      wmma::store_matrix_sync(stage + acc_r * LDQ + acc_c + i * 16, pass == 0 ? dKf[i] : dVf[i], LDQ,
                              wmma::mem_row_major);
    __syncthreads();
    for (int idx = threadIdx.x; idx < AB_BN * D; idx += AB_THREADS) {
      const int r = idx / D, c = idx % D;
      const int kj = k0 + r;
      if (kj < T) {
        bf16_t* dst = pass == 0 ? dkg + static_cast<long long>(kj) * p.dk_st : dvg + static_cast<long long>(kj) * p.dv_st;
        dst[c] = __float2bfloat16_rn(stage[r * LDQ + c]);
      }
    }
    __syncthreads();
  }
}

// RoPE backward (rotation by -angle): dq from the fp32 accumulator -> bf16 rows; dk in place.
//   dx1 = dy1*cos + dy2*sin ; dx2 = dy2*cos - dy1*sin
__global__ void __launch_bounds__(256) rope_bwd_kernel(const float* __restrict__ dq32, bf16_t* __restrict__ dq,
                                                       bf16_t* __restrict__ dk, long long ld,
                                                       const bf16_t* __restrict__ cos_t, const bf16_t* __restrict__ sin_t,
                                                       int T, int H, int hd, int pos0) {
  const long long row = blockIdx.x;
  const int t = row % T;
  const int half = hd / 2;
  const bf16_t* cr = cos_t + static_cast<long long>(pos0 + t) * hd;
  const bf16_t* sr = sin_t + static_cast<long long>(pos0 + t) * hd;
  for (int w = threadIdx.x; w < H * half; w += 256) {
    const int h = w / half, c = w % half;
    const float cs = __bfloat162float(cr[c]), sn = __bfloat162float(sr[c]);
    const long long o1 = row * ld + h * hd + c, o2 = o1 + half;
    {
      const long long s1 = (row * H + h) * hd + c;
      const float y1 = dq32[s1], y2 = dq32[s1 + half];
      dq[o1] = __float2bfloat16_rn(y1 * cs + y2 * sn);
      dq[o2] = __float2bfloat16_rn(y2 * cs - y1 * sn);
    }
    {
      const float y1 = __bfloat162float(dk[o1]), y2 = __bfloat162float(dk[o2]);
      dk[o1] = __float2bfloat16_rn(y1 * cs + y2 * sn);
      dk[o2] = __float2bfloat16_rn(y2 * cs - y1 * sn);
    }
  }
}

// ------------------------------------------------------------------------------------------------- MoE backward
// combine: out[s] = res[s] + sum_j gate[s,j] * y[slot[s,j]]
//   dy[slot[s,j]] = gate[s,j] * dout[s];  dgate[s,j] = <dout[s], y[slot[s,j]]>   (0 for dropped routes)
__global__ void __launch_bounds__(128) moe_combine_bwd_kernel(const bf16_t* __restrict__ dout, long long ldd,
                                                              const bf16_t* __restrict__ y, const int* __restrict__ slot,
                                                              const float* __restrict__ gate, bf16_t* __restrict__ dy,
                                                              float* __restrict__ dgate, int k, int D) {
  __shared__ float red[4];
  const int s = blockIdx.x;
  const bf16_t* dr = dout + static_cast<long long>(s) * ldd;
  for (int j = 0; j < k; ++j) {
    const int sl = slot[s * k + j];
    if (sl < 0) {
      if (threadIdx.x == 0) dgate[s * k + j] = 0.0f;
      continue;
    }
    const float g = bf16_round(gate[s * k + j]);
    float acc = 0.0f;
    for (int c = threadIdx.x * 8; c < D; c += 128 * 8) {
      const uint4 d4 = *reinterpret_cast<const uint4*>(dr + c);
      const uint4 y4 = *reinterpret_cast<const uint4*>(y + static_cast<long long>(sl) * D + c);
      const bf16_t* dp = reinterpret_cast<const bf16_t*>(&d4);
      const bf16_t* yp = reinterpret_cast<const bf16_t*>(&y4);
      float o[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float dv = __bfloat162float(dp[e]);
        acc += dv * __bfloat162float(yp[e]);
        o[e] = g * dv;
      }
      uint4 ov;
      ov.x = pack_bf16(o[0], o[1]); ov.y = pack_bf16(o[2], o[3]); ov.z = pack_bf16(o[4], o[5]); ov.w = pack_bf16(o[6], o[7]);
      *reinterpret_cast<uint4*>(dy + static_cast<long long>(sl) * D + c) = ov;
    }
    acc = block_sum_t<128>(acc, red);
    if (threadIdx.x == 0) dgate[s * k + j] = acc;
  }
}

// top-1 router backward: gates = softmax(logits); gate value = gates[s, expert[s]] when kept.
//   dg[e] = aux_scale * E * ce[e] / S + (e == expert[s] && kept ? dgate[s] : 0);  dlogits = g * (dg - sum_f g_f dg_f)
//   dh[s,:] += sum_e dlogits[s,e] * wg[e,:]    (one warp per token)
__global__ void __launch_bounds__(128) moe_router_bwd_kernel(const float* __restrict__ gates, const int* __restrict__ expert,
                                                             const int* __restrict__ slot, const float* __restrict__ dgate,
                                                             const int* __restrict__ exp_counts, float aux_scale,
                                                             const float* __restrict__ wg, float* __restrict__ dlogits,
                                                             bf16_t* __restrict__ dh, long long ldh, int S, int D, int E,
                                                             int k) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int s = blockIdx.x * 4 + warp;
  if (s >= S) return;
  float g[MPL_MAX_EXPERTS], dl[MPL_MAX_EXPERTS];
#pragma unroll
  for (int e = 0; e < MPL_MAX_EXPERTS; ++e) g[e] = e < E ? gates[static_cast<long long>(s) * E + e] : 0.0f;
  // gradient of the loss w.r.t. the softmax output of each chosen expert
  const int ex1 = expert[static_cast<long long>(s) * k];
  const bool kept1 = slot[static_cast<long long>(s) * k] >= 0;
  int ex2 = -1;
  float dsel1 = kept1 ? dgate[static_cast<long long>(s) * k] : 0.0f, dsel2 = 0.0f;
  if (k == 2) {
    // top2gating: the two kept gate values are renormalised by max(g1 + g2, eps) (dropped ones count as 0);
    // n_i = g_i / den  ->  dg_i = (dn_i - (dn_1 n_1 + dn_2 n_2)) / den  (the clamp passes no gradient to the sum)
    ex2 = expert[static_cast<long long>(s) * 2 + 1];
    const bool kept2 = slot[static_cast<long long>(s) * 2 + 1] >= 0;
    float g1 = 0.0f, g2 = 0.0f;
#pragma unroll
    for (int e = 0; e < MPL_MAX_EXPERTS; ++e) {
      if (e == ex1 && kept1) g1 = g[e];
      if (e == ex2 && kept2) g2 = g[e];
    }
    const float eps = 1.1920928955078125e-07f;  // torch.finfo(float32).eps
    const float sum = g1 + g2, den = fmaxf(sum, eps);
    const float dn1 = dsel1, dn2 = kept2 ? dgate[static_cast<long long>(s) * 2 + 1] : 0.0f;
    const float t = sum > eps ? (dn1 * g1 + dn2 * g2) / den : 0.0f;
    dsel1 = kept1 ? (dn1 - t) / den : 0.0f;
    dsel2 = kept2 ? (dn2 - t) / den : 0.0f;
  }
  float inner = 0.0f;
#pragma unroll
  for (int e = 0; e < MPL_MAX_EXPERTS; ++e) {
    float dg = 0.0f;
    if (e < E) {
      // l_aux = E * sum_e mean_s(gates)[e] * mean_s(mask1)[e] for both gatings (top-2: mean(me * ce) * E * E)
      dg = aux_scale * E * (static_cast<float>(exp_counts[e]) / S) / S;
      if (e == ex1) dg += dsel1;
      if (e == ex2) dg += dsel2;
    }
    dl[e] = dg;
    inner += g[e] * dg;
  }
#pragma unroll
  for (int e = 0; e < MPL_MAX_EXPERTS; ++e) {
    dl[e] = g[e] * (dl[e] - inner);
    if (e < E && lane == 0) dlogits[static_cast<long long>(s) * E + e] = dl[e];
  }
  bf16_t* hr = dh + static_cast<long long>(s) * ldh;
  for (int c = lane * 2; c < D; c += 64) {
    float a0 = __bfloat162float(hr[c]), a1 = __bfloat162float(hr[c + 1]);
#pragma unroll
    for (int e = 0; e < MPL_MAX_EXPERTS; ++e)
      if (e < E) {
        a0 += dl[e] * wg[static_cast<long long>(e) * D + c];
        a1 += dl[e] * wg[static_cast<long long>(e) * D + c + 1];
      }
    hr[c] = __float2bfloat16_rn(a0);
    hr[c + 1] = __float2bfloat16_rn(a1);
  }
}

// ------------------------------------------------------------------------------------------------- MoE buffer tails
// Expert buffers are [E * C, width] with expert e owning rows [e*C, e*C + kept[e]): the GEMMs / dispatch / combine-backward
// write exactly those rows, the later full-buffer passes (SiLU*up, rank-r weight gradients) need the REST to be zero.
// Zeroing only rows [e*C + kept[e], (e+1)*C) -- a third of the buffer at capacity_factor 1.5 -- replaces torch.zeros.
__global__ void __launch_bounds__(256) zero_tail_rows_kernel(bf16_t* __restrict__ buf, long long ld, int C, int width,
                                                             const int* __restrict__ kept) {
  const int e = blockIdx.y;
  const int k = min(max(kept[e], 0), C);
  const long long n_chunks = static_cast<long long>(C - k) * (width / 8);
  bf16_t* base = buf + (static_cast<long long>(e) * C + k) * ld;
  for (long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x; i < n_chunks;
       i += static_cast<long long>(gridDim.x) * 256) {
    const long long r = i / (width / 8), c = (i % (width / 8)) * 8;
    *reinterpret_cast<uint4*>(base + r * ld + c) = make_uint4(0, 0, 0, 0);
  }
}

// ------------------------------------------------------------------------------------------------- cross entropy
// Row r: label < 0 -> ignored. lse[r] = logsumexp(logits[r,:]); acc[0] += lse - logits[label]; acc[1] += 1 (valid rows)
constexpr int CE_THREADS = 512;
__global__ void __launch_bounds__(CE_THREADS) ce_fwd_kernel(const float* __restrict__ logits, long long ld,
                                                            const long long* __restrict__ labels, int V,
                                                            float* __restrict__ lse, float* __restrict__ acc) {
  __shared__ float red[CE_THREADS / 32];
  const long long r = blockIdx.x;
  const long long lab = labels[r];
  if (lab < 0) {
    if (threadIdx.x == 0) lse[r] = 0.0f;
    return;
  }
  const float* row = logits + r * ld;
  float m = -INFINITY;
  for (int c = threadIdx.x; c < V; c += CE_THREADS) m = fmaxf(m, row[c]);
  m = block_max_t<CE_THREADS>(m, red);
  float sum = 0.0f;
  for (int c = threadIdx.x; c < V; c += CE_THREADS) sum += __expf(row[c] - m);
  sum = block_sum_t<CE_THREADS>(sum, red);
  if (threadIdx.x == 0) {
    const float l = m + logf(sum);
    lse[r] = l;
    atomicAdd(acc, l - row[lab]);
    atomicAdd(acc + 1, 1.0f);
  }
}
// dlogits[r, c] = (softmax - onehot) * gout / n_valid  (bf16, row pitch ldd, columns [V, ldd) zero-filled)
__global__ void __launch_bounds__(CE_THREADS) ce_bwd_kernel(const float* __restrict__ logits, long long ld,
                                                            const long long* __restrict__ labels, int V,
                                                            const float* __restrict__ lse, const float* __restrict__ acc,
                                                            const float* __restrict__ gout, bf16_t* __restrict__ dl,
                                                            long long ldd) {
  const long long r = blockIdx.x;
  const long long lab = labels[r];
  bf16_t* out = dl + r * ldd;
  if (lab < 0) {
    for (int c = threadIdx.x; c < ldd; c += CE_THREADS) out[c] = __float2bfloat16(0.0f);
    return;
  }
  const float sc = (gout ? *gout : 1.0f) / fmaxf(acc[1], 1.0f);
  const float* row = logits + r * ld;
  const float l = lse[r];
  for (int c = threadIdx.x; c < ldd; c += CE_THREADS) {
    float v = 0.0f;
    if (c < V) v = (__expf(row[c] - l) - (c == lab ? 1.0f : 0.0f)) * sc;
    out[c] = __float2bfloat16_rn(v);
  }
}

// ------------------------------------------------------------------------------------------------- embedding backward
// adjoint of mpl_gather_rows: idx >= 0 -> dtable[idx] += dx[r]; idx <= -2 -> dfeats[-idx-2] += dx[r]
__global__ void __launch_bounds__(128) scatter_add_rows_kernel(const bf16_t* __restrict__ dx, long long ldx,
                                                               const int* __restrict__ idx, float* __restrict__ dtable,
                                                               long long ldt, float* __restrict__ dfeats, long long ldf,
                                                               int D) {
  const long long r = blockIdx.x;
  const int i = idx[r];
  float* dst = nullptr;
  if (i >= 0 && dtable != nullptr) dst = dtable + static_cast<long long>(i) * ldt;
  if (i <= -2 && dfeats != nullptr) dst = dfeats + static_cast<long long>(-i - 2) * ldf;
  if (dst == nullptr) return;
  const bf16_t* src = dx + r * ldx;
  for (int c = threadIdx.x; c < D; c += 128) atomicAdd(dst + c, __bfloat162float(src[c]));
}

// ------------------------------------------------------------------------------------------------- optimizer
// DETERMINISTIC: every data-parallel rank must derive the same clip factor from the same (all-reduced) gradients, or the
// replicas' parameters drift apart one ulp at a time. Block partials go to a scratch array; the last block to finish
// (atomic ticket) adds them in index order. *out is ADDED to (callers zero it), like the atomic version it replaces.
constexpr int SUMSQ_MAX_BLOCKS = 1024;
__device__ float g_sumsq_partial[SUMSQ_MAX_BLOCKS];
__device__ unsigned int g_sumsq_ticket = 0;
__global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ g, long long n, float* __restrict__ out) {
  __shared__ float red[8];
  __shared__ bool last;
  float acc = 0.0f;
  for (long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * 256) {
    const float v = g[i];
    acc += v * v;
  }
  acc = block_sum_t<256>(acc, red);
  if (threadIdx.x == 0) {
    g_sumsq_partial[blockIdx.x] = acc;
    __threadfence();
    last = atomicAdd(&g_sumsq_ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  float t = 0.0f;
  for (unsigned int i = threadIdx.x; i < gridDim.x; i += 256) t += __ldcg(&g_sumsq_partial[i]);  // fixed assignment
  t = block_sum_t<256>(t, red);                                                                   // fixed tree
  if (threadIdx.x == 0) {
    *out += t;
    g_sumsq_ticket = 0;  // ready for the next launch (launches on one stream are serialised)
  }
}
// AdamW (torch.optim.AdamW / DeepSpeed FusedAdam adam_w_mode): fp32 master + moments; clip = min(1, max_norm/(norm+1e-6))
__global__ void __launch_bounds__(256) adamw_kernel(float* __restrict__ master, float* __restrict__ m, float* __restrict__ v,
                                                    const float* __restrict__ g, void* __restrict__ param, int param_bf16,
                                                    long long n, float lr, float b1, float b2, float eps, float wd,
                                                    float bc1, float bc2, const float* __restrict__ sumsq, float max_norm,
                                                    float grad_scale) {
  const long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
  if (i >= n) return;
  float clip = grad_scale;
  if (sumsq != nullptr && max_norm > 0.0f) {
    const float norm = sqrtf(*sumsq) * grad_scale;
    clip *= fminf(1.0f, max_norm / (norm + 1e-6f));
  }
  const float gi = g[i] * clip;
  float w = master[i];
  w *= 1.0f - lr * wd;
  const float mi = b1 * m[i] + (1.0f - b1) * gi;
  const float vi = b2 * v[i] + (1.0f - b2) * gi * gi;
  m[i] = mi;
  v[i] = vi;
  const float denom = sqrtf(vi) / sqrtf(bc2) + eps;
  w -= (lr / bc1) * mi / denom;
  master[i] = w;
  if (param_bf16)
    static_cast<bf16_t*>(param)[i] = __float2bfloat16_rn(w);
  else
    static_cast<float*>(param)[i] = w;
}

// The same update over the WHOLE gradient arena in one launch: `chunks` (device, 4 x int64 per chunk: arena offset, length,
// parameter base address, is_bf16 | element offset within the parameter << 1) maps 4096-element chunks to parameters.
constexpr int ADAMW_CHUNK = 4096;
__global__ void __launch_bounds__(256) adamw_multi_kernel(float* __restrict__ master, float* __restrict__ m,
                                                          float* __restrict__ v, const float* __restrict__ g,
                                                          const long long* __restrict__ chunks, float lr, float b1,
                                                          float b2, float eps, float wd, float bc1, float bc2,
                                                          const float* __restrict__ sumsq, float max_norm,
                                                          float grad_scale) {
  const long long* c = chunks + static_cast<long long>(blockIdx.x) * 4;
  const long long off = c[0], len = c[1], pbase = c[2], meta = c[3];
  const bool is_bf16 = (meta & 1) != 0;
  const long long poff = meta >> 1;
  float clip = grad_scale;
  if (sumsq != nullptr && max_norm > 0.0f) {
    const float norm = sqrtf(*sumsq) * grad_scale;
    clip *= fminf(1.0f, max_norm / (norm + 1e-6f));
  }
  const float rbc2 = rsqrtf(bc2), step = lr / bc1;
  for (long long i = threadIdx.x; i < len; i += 256) {
    const long long a = off + i;
    const float gi = g[a] * clip;
    float w = master[a];
    w *= 1.0f - lr * wd;
    const float mi = b1 * m[a] + (1.0f - b1) * gi;
    const float vi = b2 * v[a] + (1.0f - b2) * gi * gi;
    m[a] = mi;
    v[a] = vi;
    const float denom = sqrtf(vi) * rbc2 + eps;
    w -= step * mi / denom;
    master[a] = w;
    if (is_bf16)
      reinterpret_cast<bf16_t*>(pbase)[poff + i] = __float2bfloat16_rn(w);
    else
      reinterpret_cast<float*>(pbase)[poff + i] = w;
  }
}

// ------------------------------------------------------------------------------------------------- mask losses
// One CTA per mask. out[0..3] = BCE-with-logits mean, Dice loss, IoU-MSE loss, Focal loss (model/MedPLIB.py:26-124);
// sums[0..5] = sum bce, sum p, sum t, sum p*t, focal_pos, focal_neg (kept for the backward).
constexpr int ML_THREADS = 1024;
__global__ void __launch_bounds__(ML_THREADS) mask_loss_kernel(const bf16_t* __restrict__ pred, const float* __restrict__ gt,
                                                               const bf16_t* __restrict__ pred_iou, long long n,
                                                               float* __restrict__ out, float* __restrict__ sums) {
  __shared__ float red[ML_THREADS / 32];
  float a[6] = {0, 0, 0, 0, 0, 0};
  for (long long i = threadIdx.x; i < n; i += ML_THREADS) {
    const float x = __bfloat162float(pred[i]), t = gt[i];
    const float pr = 1.0f / (1.0f + __expf(-x));
    a[0] += fmaxf(x, 0.0f) - x * t + log1pf(__expf(-fabsf(x)));
    a[1] += pr;
    a[2] += t;
    a[3] += pr * t;
    a[4] += -0.25f * t * (1.0f - pr) * (1.0f - pr) * logf(pr + 1e-12f);
    a[5] += -0.75f * (1.0f - t) * pr * pr * logf(1.0f - pr + 1e-12f);
  }
  float s[6];
#pragma unroll
  for (int j = 0; j < 6; ++j) s[j] = block_sum_t<ML_THREADS>(a[j], red);
  if (threadIdx.x == 0) {
    const float nf = static_cast<float>(n);
    out[0] = s[0] / nf / (1.0f + 1e-8f);
    out[1] = 1.0f - (2.0f * s[3] + 1e-6f) / (s[1] + s[2] + 1e-6f);
    const float iou = (s[3] + 1e-7f) / (s[1] + s[2] - s[3] + 1e-7f);
    const float pi = pred_iou ? __bfloat162float(*pred_iou) : 0.0f;
    out[2] = (iou - pi) * (iou - pi);
    out[3] = (s[4] + s[5]) / (nf + 1e-12f);
    if (sums != nullptr)
      for (int j = 0; j < 6; ++j) sums[j] = s[j];
  }
}

}  // namespace mpl

// ================================================================================================= C ABI
using namespace mpl;
#define ST(s) static_cast<cudaStream_t>(s)

extern "C" int mpl_transpose_bf16(const void* in, long long ld_in, void* out, long long ld_out, int rows, int cols,
                                  void* stream) {
  if (rows <= 0 || cols <= 0) return MPL_OK;
  if (in == nullptr || out == nullptr) return MPL_ERR_ARG;
  dim3 grid((cols + 63) / 64, (rows + 63) / 64);
  transpose_kernel<<<grid, 256, 0, ST(stream)>>>(static_cast<const bf16_t*>(in), ld_in, static_cast<bf16_t*>(out), ld_out,
                                                 rows, cols);
  return launch_status();
}

// mpl_lora_down + a second, padded copy of u for the GEMM's extension k-block (mpl_gemm_args.ext_a): u_pad bf16 [M, 64],
// columns [pad_col, pad_col + r). MPL_ERR_UNSUPPORTED when the shape does not take the tensor-core kernel.
extern "C" int mpl_lora_down_ext(const void* x, long long ldx, const void* A, long long lda, void* u, int u_is_f32, int M,
                                 int K, int r, float scale, void* u_pad, int pad_col, void* stream) {
  if (M <= 0) return MPL_OK;
  if (x == nullptr || A == nullptr || u == nullptr || u_pad == nullptr || r < 1 || pad_col < 0 || pad_col + r > 64)
    return MPL_ERR_ARG;
  if (!(r <= 8 && K % 32 == 0 && ldx % 8 == 0 && lda % 8 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 &&
        (reinterpret_cast<uintptr_t>(A) & 15) == 0))
    return MPL_ERR_UNSUPPORTED;
  lora_down_mma_kernel<<<(M + 15) / 16, 256, 0, ST(stream)>>>(static_cast<const bf16_t*>(x), ldx,
                                                              static_cast<const bf16_t*>(A), lda, u, u_is_f32, M, K, r, scale,
                                                              static_cast<bf16_t*>(u_pad), pad_col);
  return launch_status();
}

// items: device array of n_items {src, dst, sn, sr, N, r, col, scale, dn, dj} (see LoraPackItem; 64 bytes each)
extern "C" int mpl_lora_pack(const void* items, int n_items, void* stream) {
  if (n_items <= 0) return MPL_OK;
  if (items == nullptr) return MPL_ERR_ARG;
  dim3 grid(32, static_cast<unsigned>(n_items));
  lora_pack_kernel<<<grid, 256, 0, ST(stream)>>>(static_cast<const LoraPackItem*>(items));
  return launch_status();
}

extern "C" int mpl_lora_down(const void* x, long long ldx, const void* A, long long lda, void* u, int u_is_f32, int M,
                             int K, int r, float scale, void* stream) {
  if (M <= 0) return MPL_OK;
  if (x == nullptr || A == nullptr || u == nullptr || r < 1 || r > LORA_MAX_R) return MPL_ERR_ARG;
  if (K % 8 != 0 || ldx % 8 != 0 || lda % 8 != 0) return MPL_ERR_ALIGN;
  if (r <= 8 && K % 32 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(A) & 15) == 0) {
    lora_down_mma_kernel<<<(M + 15) / 16, 256, 0, ST(stream)>>>(static_cast<const bf16_t*>(x), ldx,
                                                                static_cast<const bf16_t*>(A), lda, u, u_is_f32, M, K, r,
                                                                scale, nullptr, 0);
    return launch_status();
  }
  if (r <= LR8 && static_cast<size_t>(r) * K * 2 <= 200 * 1024 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 &&
      (reinterpret_cast<uintptr_t>(A) & 15) == 0) {
    static bool attr = false;
    if (!attr) {
      if (cudaFuncSetAttribute(lora_down8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess)
        return MPL_ERR_CUDA;
      attr = true;
    }
    // one CTA per SM-sized slab of rows (A is staged once per CTA)
    int rows = (M + num_sms() - 1) / num_sms();
    if (rows < 16) rows = 16;
    const int grid = (M + rows - 1) / rows;
    lora_down8_kernel<<<grid, LD8_THREADS, static_cast<size_t>(r) * K * 2, ST(stream)>>>(
        static_cast<const bf16_t*>(x), ldx, static_cast<const bf16_t*>(A), lda, u, u_is_f32, M, K, r, scale, rows);
    return launch_status();
  }
  lora_down_kernel<<<(M + 7) / 8, 256, 0, ST(stream)>>>(static_cast<const bf16_t*>(x), ldx, static_cast<const bf16_t*>(A),
                                                        lda, u, u_is_f32, M, K, r, scale);
  return launch_status();
}

extern "C" int mpl_lora_up_add(void* y, long long ldy, const void* u, int u_is_f32, const void* Bm, long long bm_stride_n,
                               long long bm_stride_r, float scale, int M, int N, int r, void* stream) {
  if (M <= 0 || N <= 0) return MPL_OK;
  if (y == nullptr || u == nullptr || Bm == nullptr || r < 1 || r > LORA_MAX_R) return MPL_ERR_ARG;
  if (r <= 8 && N % 8 == 0 && ldy % 8 == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0) {
    dim3 gm((N + 255) / 256, (M + UPM_ROWS - 1) / UPM_ROWS);
    lora_up_add_mma_kernel<<<gm, 256, 0, ST(stream)>>>(static_cast<bf16_t*>(y), ldy, u, u_is_f32,
                                                       static_cast<const bf16_t*>(Bm), bm_stride_n, bm_stride_r, scale, M,
                                                       N, r);
    return launch_status();
  }
  if (r <= LR8 && N % 8 == 0 && ldy % 8 == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0) {
    dim3 g8((N + 2047) / 2048, (M + UP8_ROWS - 1) / UP8_ROWS);
    lora_up_add8_kernel<<<g8, 256, 0, ST(stream)>>>(static_cast<bf16_t*>(y), ldy, u, u_is_f32,
                                                    static_cast<const bf16_t*>(Bm), bm_stride_n, bm_stride_r, scale, M, N, r);
    return launch_status();
  }
  dim3 grid((N + 511) / 512, (M + UP_ROWS - 1) / UP_ROWS);
  lora_up_add_kernel<<<grid, 256, 0, ST(stream)>>>(static_cast<bf16_t*>(y), ldy, u, u_is_f32,
                                                   static_cast<const bf16_t*>(Bm), bm_stride_n, bm_stride_r, scale, M, N, r);
  return launch_status();
}

extern "C" int mpl_rank_wgrad(const void* X, long long ldx, const void* U, int u_is_f32, float* out, long long out_stride_n,
                              long long out_stride_r, float scale, int M, int N, int r, void* stream) {
  if (M <= 0 || N <= 0) return MPL_OK;
  if (X == nullptr || U == nullptr || out == nullptr || r < 1 || r > LORA_MAX_R) return MPL_ERR_ARG;
  if (r <= 8 && N % 8 == 0 && ldx % 8 == 0 && (reinterpret_cast<uintptr_t>(X) & 15) == 0) {
    dim3 gm((N + 63) / 64, (M + WGM_ROWS - 1) / WGM_ROWS);
    rank_wgrad_mma_kernel<<<gm, 256, 0, ST(stream)>>>(static_cast<const bf16_t*>(X), ldx, U, u_is_f32, out, out_stride_n,
                                                      out_stride_r, scale, M, N, r);
    return launch_status();
  }
  if (r <= LR8 && N % 8 == 0 && ldx % 8 == 0 && (reinterpret_cast<uintptr_t>(X) & 15) == 0) {
    dim3 g8((N + 255) / 256, (M + WG8_ROWS - 1) / WG8_ROWS);
    rank_wgrad8_kernel<<<g8, 256, 0, ST(stream)>>>(static_cast<const bf16_t*>(X), ldx, U, u_is_f32, out, out_stride_n,
                                                   out_stride_r, scale, M, N, r);
    return launch_status();
  }
  const int gx = (N + 511) / 512;
  int split = (2 * num_sms() + gx - 1) / gx;
  const int max_split = (M + WG_ROWS - 1) / WG_ROWS;
  if (split > max_split) split = max_split;
  if (split < 1) split = 1;
  int rows = (M + split - 1) / split;
  rows = (rows + WG_ROWS - 1) / WG_ROWS * WG_ROWS;
  split = (M + rows - 1) / rows;
  dim3 grid(gx, split);
  rank_wgrad_kernel<<<grid, 256, 0, ST(stream)>>>(static_cast<const bf16_t*>(X), ldx, U, u_is_f32, out, out_stride_n,
                                                  out_stride_r, scale, M, N, r, rows);
  return launch_status();
}

extern "C" int mpl_rmsnorm_bwd(const void* x, long long ldx, const void* weight, const void* dy, long long lddy,
                               const void* add, long long ldadd, void* dx, long long lddx, float* dweight, int rows, int D,
                               float eps, void* stream) {
  if (rows <= 0) return MPL_OK;
  if (x == nullptr || weight == nullptr || dy == nullptr || dx == nullptr) return MPL_ERR_ARG;
  if (D % 8 != 0 || D > 4096 || ldx % 8 != 0 || lddy % 8 != 0 || lddx % 8 != 0 || (add != nullptr && ldadd % 8 != 0))
    return MPL_ERR_ALIGN;
  rmsnorm_bwd_kernel<<<rows, RB_THREADS, 0, ST(stream)>>>(
      static_cast<const bf16_t*>(x), ldx, static_cast<const bf16_t*>(weight), static_cast<const bf16_t*>(dy), lddy,
      static_cast<const bf16_t*>(add), ldadd, static_cast<bf16_t*>(dx), lddx, dweight, D, eps);
  return launch_status();
}

extern "C" int mpl_silu_mul(const void* g, const void* u, void* h, long long n, void* stream) {
  if (n <= 0) return MPL_OK;
  if (g == nullptr || u == nullptr || h == nullptr) return MPL_ERR_ARG;
  if (n % 8 != 0) return MPL_ERR_ALIGN;
  const long long n8 = n / 8;
  silu_mul_kernel<<<static_cast<unsigned>((n8 + 255) / 256), 256, 0, ST(stream)>>>(
      static_cast<const bf16_t*>(g), static_cast<const bf16_t*>(u), static_cast<bf16_t*>(h), n8);
  return launch_status();
}
extern "C" int mpl_silu_mul_bwd(const void* g, const void* u, const void* dh, void* dg, void* du, long long n,
                                void* stream) {
  if (n <= 0) return MPL_OK;
  if (g == nullptr || u == nullptr || dh == nullptr || dg == nullptr || du == nullptr) return MPL_ERR_ARG;
  if (n % 8 != 0) return MPL_ERR_ALIGN;
  const long long n8 = n / 8;
  silu_mul_bwd_kernel<<<static_cast<unsigned>((n8 + 255) / 256), 256, 0, ST(stream)>>>(
      static_cast<const bf16_t*>(g), static_cast<const bf16_t*>(u), static_cast<const bf16_t*>(dh),
      static_cast<bf16_t*>(dg), static_cast<bf16_t*>(du), n8);
  return launch_status();
}

extern "C" int mpl_attention_bwd(const mpl_attn_bwd_args* a, void* stream) {
  if (a == nullptr || a->q == nullptr || a->k == nullptr || a->v == nullptr || a->o == nullptr || a->d_o == nullptr ||
      a->lse == nullptr || a->delta == nullptr || a->dq_f32 == nullptr || a->dk == nullptr || a->dv == nullptr)
    return MPL_ERR_ARG;
  if (a->B <= 0 || a->H <= 0 || a->T <= 0) return MPL_OK;
  if (a->head_dim != 128 && a->head_dim != 64) return MPL_ERR_UNSUPPORTED;
  const long long strides[] = {a->q_stride[0], a->q_stride[1], a->q_stride[2], a->k_stride[0], a->k_stride[1],
                               a->k_stride[2], a->v_stride[0], a->v_stride[1], a->v_stride[2], a->o_stride[0],
                               a->o_stride[1], a->o_stride[2]};
  for (long long s : strides)
    if (s % 8 != 0) return MPL_ERR_ALIGN;
  const long long rows = static_cast<long long>(a->B) * a->H * a->T;
  attn_delta_kernel<<<static_cast<unsigned>((rows + 7) / 8), 256, 0, ST(stream)>>>(
      static_cast<const bf16_t*>(a->o), static_cast<const bf16_t*>(a->d_o), a->o_stride[0], a->o_stride[1],
      a->o_stride[2], a->delta, a->B, a->H, a->T, a->head_dim);
  if (launch_status() != MPL_OK) return MPL_ERR_CUDA;
  if (attention_bwd_tc_supported(*a)) return attention_bwd_tc(*a, ST(stream));  // tcgen05 kernel (attention_tc.cu)
  if (a->head_dim == 128 && a->dk_stride[1] % 2 == 0 && a->dv_stride[1] % 2 == 0 && getenv("MPL_ATTN_BWD_WMMA") == nullptr)
    return attn_bwd_fa2(*a, ST(stream));  // mma.sync FA2-style kernel (attention.cu); the wmma kernel below: d = 64
  AttnBwdParams p;
  p.q = static_cast<const bf16_t*>(a->q);
  p.k = static_cast<const bf16_t*>(a->k);
  p.v = static_cast<const bf16_t*>(a->v);
  p.dO = static_cast<const bf16_t*>(a->d_o);
  p.q_sb = a->q_stride[0]; p.q_st = a->q_stride[1]; p.q_sh = a->q_stride[2];
  p.k_sb = a->k_stride[0]; p.k_st = a->k_stride[1]; p.k_sh = a->k_stride[2];
  p.v_sb = a->v_stride[0]; p.v_st = a->v_stride[1]; p.v_sh = a->v_stride[2];
  p.o_sb = a->o_stride[0]; p.o_st = a->o_stride[1]; p.o_sh = a->o_stride[2];
  p.lse = a->lse;
  p.delta = a->delta;
  p.dq = a->dq_f32;
  p.dk = static_cast<bf16_t*>(a->dk);
  p.dv = static_cast<bf16_t*>(a->dv);
  p.dk_sb = a->dk_stride[0]; p.dk_st = a->dk_stride[1]; p.dk_sh = a->dk_stride[2];
  p.dv_sb = a->dv_stride[0]; p.dv_st = a->dv_stride[1]; p.dv_sh = a->dv_stride[2];
  p.B = a->B; p.H = a->H; p.T = a->T;
  p.scale = a->scale;
  p.causal = a->causal;
  p.kv_mask = a->kv_mask;
  p.kv_mask_stride = a->kv_mask_stride > 0 ? a->kv_mask_stride : a->T;
  dim3 grid((a->T + AB_BN - 1) / AB_BN, a->H, a->B);
  auto smem_bytes = [](int D) {
    return 6 * 64 * (D + 8) * 2 + 2 * 64 * (AB_BN + 4) * 4 + 2 * 64 * (AB_BN + 8) * 2 + 2 * 64 * 4;
  };
  if (a->head_dim == 128) {
    static bool set = false;
    const int sm = smem_bytes(128);
    if (!set) {
      if (cudaFuncSetAttribute(attn_bwd_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm) != cudaSuccess)
        return MPL_ERR_CUDA;
      set = true;
    }
    attn_bwd_kernel<128><<<grid, AB_THREADS, sm, ST(stream)>>>(p);
  } else {
    static bool set = false;
    const int sm = smem_bytes(64);
    if (!set) {
      if (cudaFuncSetAttribute(attn_bwd_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm) != cudaSuccess)
        return MPL_ERR_CUDA;
      set = true;
    }
    attn_bwd_kernel<64><<<grid, AB_THREADS, sm, ST(stream)>>>(p);
  }
  return launch_status();
}

extern "C" int mpl_rope_bwd(const float* dq_f32, void* dq, void* dk, long long ld, const void* cos_t, const void* sin_t,
                            int B, int T, int H, int head_dim, int pos0, void* stream) {
  if (B <= 0 || T <= 0) return MPL_OK;
  if (dq_f32 == nullptr || dq == nullptr || dk == nullptr || cos_t == nullptr || sin_t == nullptr) return MPL_ERR_ARG;
  rope_bwd_kernel<<<B * T, 256, 0, ST(stream)>>>(dq_f32, static_cast<bf16_t*>(dq), static_cast<bf16_t*>(dk), ld,
                                                 static_cast<const bf16_t*>(cos_t), static_cast<const bf16_t*>(sin_t), T, H,
                                                 head_dim, pos0);
  return launch_status();
}

extern "C" int mpl_moe_combine_bwd(const void* dout, long long ldd, const void* y, const int* slot, const float* gate,
                                   void* dy, float* dgate, int S, int k, int D, void* stream) {
  if (S <= 0) return MPL_OK;
  if (dout == nullptr || y == nullptr || slot == nullptr || gate == nullptr || dy == nullptr || dgate == nullptr)
    return MPL_ERR_ARG;
  if (D % 8 != 0 || ldd % 8 != 0) return MPL_ERR_ALIGN;
  moe_combine_bwd_kernel<<<S, 128, 0, ST(stream)>>>(static_cast<const bf16_t*>(dout), ldd, static_cast<const bf16_t*>(y),
                                                    slot, gate, static_cast<bf16_t*>(dy), dgate, k, D);
  return launch_status();
}

extern "C" int mpl_moe_router_bwd(const float* gates, const int* expert, const int* slot, const float* dgate,
                                  const int* exp_counts, float aux_scale, const float* wg, float* dlogits, void* dh,
                                  long long ldh, int S, int D, int E, int k, void* stream) {
  if (S <= 0) return MPL_OK;
  if (gates == nullptr || expert == nullptr || slot == nullptr || dgate == nullptr || exp_counts == nullptr ||
      wg == nullptr || dlogits == nullptr || dh == nullptr || E < 1 || E > MPL_MAX_EXPERTS || (k != 1 && k != 2))
    return MPL_ERR_ARG;
  if (D % 2 != 0) return MPL_ERR_ALIGN;
  moe_router_bwd_kernel<<<(S + 3) / 4, 128, 0, ST(stream)>>>(gates, expert, slot, dgate, exp_counts, aux_scale, wg, dlogits,
                                                             static_cast<bf16_t*>(dh), ldh, S, D, E, k);
  return launch_status();
}

extern "C" int mpl_ce_fwd(const float* logits, long long ld, const long long* labels, int rows, int V, float* lse,
                          float* acc, void* stream) {
  if (rows <= 0) return MPL_OK;
  if (logits == nullptr || labels == nullptr || lse == nullptr || acc == nullptr) return MPL_ERR_ARG;
  ce_fwd_kernel<<<rows, CE_THREADS, 0, ST(stream)>>>(logits, ld, labels, V, lse, acc);
  return launch_status();
}
extern "C" int mpl_ce_bwd(const float* logits, long long ld, const long long* labels, int rows, int V, const float* lse,
                          const float* acc, const float* grad_out, void* dlogits, long long ldd, void* stream) {
  if (rows <= 0) return MPL_OK;
  if (logits == nullptr || labels == nullptr || lse == nullptr || acc == nullptr || dlogits == nullptr || ldd < V)
    return MPL_ERR_ARG;
  ce_bwd_kernel<<<rows, CE_THREADS, 0, ST(stream)>>>(logits, ld, labels, V, lse, acc, grad_out,
                                                     static_cast<bf16_t*>(dlogits), ldd);
  return launch_status();
}

extern "C" int mpl_scatter_add_rows(const void* dx, long long ldx, const int* idx, float* dtable, long long ld_table,
                                    float* dfeats, long long ld_feats, int rows, int D, void* stream) {
  if (rows <= 0) return MPL_OK;
  if (dx == nullptr || idx == nullptr) return MPL_ERR_ARG;
  scatter_add_rows_kernel<<<rows, 128, 0, ST(stream)>>>(static_cast<const bf16_t*>(dx), ldx, idx, dtable, ld_table, dfeats,
                                                        ld_feats, D);
  return launch_status();
}

extern "C" int mpl_sumsq_f32(const float* g, long long n, float* out, void* stream) {
  if (n <= 0) return MPL_OK;
  if (g == nullptr || out == nullptr) return MPL_ERR_ARG;
  long long blocks = (n + 255) / 256;
  if (blocks > 4 * num_sms()) blocks = 4 * num_sms();
  if (blocks > SUMSQ_MAX_BLOCKS) blocks = SUMSQ_MAX_BLOCKS;
  sumsq_kernel<<<static_cast<unsigned>(blocks), 256, 0, ST(stream)>>>(g, n, out);
  return launch_status();
}

extern "C" int mpl_adamw(float* master, float* m, float* v, const float* grad, void* param, int param_is_bf16, long long n,
                         float lr, float beta1, float beta2, float eps, float weight_decay, int step, const float* sumsq,
                         float max_norm, float grad_scale, void* stream) {
  if (n <= 0) return MPL_OK;
  if (master == nullptr || m == nullptr || v == nullptr || grad == nullptr || param == nullptr || step < 1)
    return MPL_ERR_ARG;
  const float bc1 = 1.0f - powf(beta1, static_cast<float>(step)), bc2 = 1.0f - powf(beta2, static_cast<float>(step));
  adamw_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, ST(stream)>>>(master, m, v, grad, param, param_is_bf16, n, lr,
                                                                              beta1, beta2, eps, weight_decay, bc1, bc2,
                                                                              sumsq, max_norm, grad_scale);
  return launch_status();
}

extern "C" int mpl_adamw_multi(float* master, float* m, float* v, const float* grad, const long long* chunks,
                               int n_chunks, float lr, float beta1, float beta2, float eps, float weight_decay, int step,
                               const float* sumsq, float max_norm, float grad_scale, void* stream) {
  if (n_chunks <= 0) return MPL_OK;
  if (master == nullptr || m == nullptr || v == nullptr || grad == nullptr || chunks == nullptr || step < 1)
    return MPL_ERR_ARG;
  const float bc1 = 1.0f - powf(beta1, static_cast<float>(step)), bc2 = 1.0f - powf(beta2, static_cast<float>(step));
  adamw_multi_kernel<<<n_chunks, 256, 0, ST(stream)>>>(master, m, v, grad, chunks, lr, beta1, beta2, eps, weight_decay, bc1,
                                                       bc2, sumsq, max_norm, grad_scale);
  return launch_status();
}

extern "C" int mpl_mask_losses(const void* pred, const float* gt, const void* pred_iou, long long n, float* out4,
                               float* sums6, void* stream) {
  if (pred == nullptr || gt == nullptr || out4 == nullptr || n <= 0) return MPL_ERR_ARG;
  mask_loss_kernel<<<1, ML_THREADS, 0, ST(stream)>>>(static_cast<const bf16_t*>(pred), gt,
                                                     static_cast<const bf16_t*>(pred_iou), n, out4, sums6);
  return launch_status();
}

extern "C" int mpl_zero_tail_rows(void* buf, long long ld, int groups, int C, int width, const int* kept, void* stream) {
  if (groups <= 0 || C <= 0) return MPL_OK;
  if (buf == nullptr || kept == nullptr) return MPL_ERR_ARG;
  if (width % 8 != 0 || ld % 8 != 0 || (reinterpret_cast<uintptr_t>(buf) & 15) != 0) return MPL_ERR_ALIGN;
  dim3 grid(static_cast<unsigned>(mpl::num_sms() * 2), static_cast<unsigned>(groups));
  zero_tail_rows_kernel<<<grid, 256, 0, ST(stream)>>>(static_cast<bf16_t*>(buf), ld, C, width, kept);
  return launch_status();
}
