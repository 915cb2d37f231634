// K3 — skinny GEMM for decode-time and tiny-M linears:  C[M,N] = epilogue(A[M,K] · W[N,K]^T),  M <= 16.
//
// HBM-bound: the weight matrix is streamed exactly once with 16-byte loads straight from global memory into
// mma.sync (m16n8k16, bf16 -> fp32) A-fragments — the 16 weight rows of a CTA take the MMA "M" role and the
// (<= 8 per tile) activation rows the "N" role, so no shared-memory staging and no FFMA/LDS bottleneck at M = 8.
// A thread's 16-byte load covers k = k0 + 8t .. 8t+7; the same k-permutation is applied to the activation
// fragment, which leaves the dot products unchanged. The 8 warps of a CTA interleave over K in 64-byte
// segments (so a CTA reads 512 contiguous bytes of each weight row per step) and reduce through shared memory.
// Replaces the same nn.Linear call sites as K1 when M is the decode batch.
#include "internal.h"
#include "ptx.cuh"

namespace mpl {

constexpr int SK_WARPS = 8;
constexpr int SK_THREADS = SK_WARPS * 32;
constexpr int SK_ROWS = 16;  // weight rows per CTA
constexpr int SK_UNROLL = 4;

struct SkinnyParams {
  const __nv_bfloat16* A;
  long long lda;
  const __nv_bfloat16* Wn[3];  // blockIdx.y selects the weight matrix / output / bias (fused q,k,v projections)
  const __nv_bfloat16* W2;
  long long ldw;
  void* Cn[3];
  long long ldc;
  const __nv_bfloat16* biasn[3];
  const __nv_bfloat16* residual;
  long long ldr;
  const float* row_scale;
  const int* m_dev;
  int M, N, K;
  int act;
  int out_f32;
};

__device__ __forceinline__ void hmma_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                           uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

__device__ __forceinline__ uint4 ldg_stream(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}

__device__ __forceinline__ float skinny_act(float v, int act) {
  switch (act) {
    case MPL_ACT_GELU:
      return 0.5f * v * (1.0f + erff(v * 0.70710678118654752f));
    case MPL_ACT_QUICK_GELU:
      return v / (1.0f + __expf(-1.702f * v));
    case MPL_ACT_RELU:
      return fmaxf(v, 0.0f);
    case MPL_ACT_SILU:
      return v / (1.0f + __expf(-v));
    case MPL_ACT_SIGMOID:
      return 1.0f / (1.0f + __expf(-v));
    default:
      return v;
  }
}

template <int MT, bool DUAL>
__global__ void __launch_bounds__(SK_THREADS) skinny_gemm_kernel(const SkinnyParams p) {
  constexpr int NW = DUAL ? 2 : 1;
  __shared__ float red[SK_WARPS][NW][SK_ROWS][MT * 8 + 1];
  int M = p.M;
  if (p.m_dev != nullptr) M = min(M, *p.m_dev);
  if (M <= 0) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int which = blockIdx.y;
  const __nv_bfloat16* W = which == 0 ? p.Wn[0] : (which == 1 ? p.Wn[1] : p.Wn[2]);
  void* Cout = which == 0 ? p.Cn[0] : (which == 1 ? p.Cn[1] : p.Cn[2]);
  const __nv_bfloat16* bias = which == 0 ? p.biasn[0] : (which == 1 ? p.biasn[1] : p.biasn[2]);
  const int n0 = blockIdx.x * SK_ROWS;
  const int r0 = min(n0 + g, p.N - 1);
  const int r1 = min(n0 + g + 8, p.N - 1);
  const __nv_bfloat16* w0p[NW];
  const __nv_bfloat16* w1p[NW];
  w0p[0] = W + static_cast<long long>(r0) * p.ldw;
  w1p[0] = W + static_cast<long long>(r1) * p.ldw;
  if (DUAL) {
    w0p[NW - 1] = p.W2 + static_cast<long long>(r0) * p.ldw;
    w1p[NW - 1] = p.W2 + static_cast<long long>(r1) * p.ldw;
  }
  const __nv_bfloat16* xp[MT];
  bool xok[MT];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
    const int m = mt * 8 + g;
    xok[mt] = m < M;
    xp[mt] = p.A + static_cast<long long>(xok[mt] ? m : 0) * p.lda;
  }
  float acc[NW][MT][4];
#pragma unroll
  for (int w = 0; w < NW; ++w)
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[w][mt][i] = 0.0f;

  const int chunks = (p.K + 31) / 32;
  for (int j0 = warp; j0 < chunks; j0 += SK_WARPS * SK_UNROLL) {
    uint4 wa[SK_UNROLL][NW], wb[SK_UNROLL][NW], xb[SK_UNROLL][MT];
#pragma unroll
    for (int u = 0; u < SK_UNROLL; ++u) {
      const int k = (j0 + u * SK_WARPS) * 32 + t * 8;
      const bool ok = k < p.K;
#pragma unroll
      for (int w = 0; w < NW; ++w) {
        wa[u][w] = ok ? ldg_stream(w0p[w] + k) : make_uint4(0, 0, 0, 0);
        wb[u][w] = ok ? ldg_stream(w1p[w] + k) : make_uint4(0, 0, 0, 0);
      }
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
        xb[u][mt] = (ok && xok[mt]) ? *reinterpret_cast<const uint4*>(xp[mt] + k) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int u = 0; u < SK_UNROLL; ++u)
#pragma unroll
      for (int w = 0; w < NW; ++w)
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          hmma_16816(acc[w][mt], wa[u][w].x, wb[u][w].x, wa[u][w].y, wb[u][w].y, xb[u][mt].x, xb[u][mt].y);
          hmma_16816(acc[w][mt], wa[u][w].z, wb[u][w].z, wa[u][w].w, wb[u][w].w, xb[u][mt].z, xb[u][mt].w);
        }
  }
  // C fragment: c0,c1 -> (weight row g, m = 2t, 2t+1); c2,c3 -> (weight row g+8, same m)
#pragma unroll
  for (int w = 0; w < NW; ++w)
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      red[warp][w][g][mt * 8 + t * 2] = acc[w][mt][0];
      red[warp][w][g][mt * 8 + t * 2 + 1] = acc[w][mt][1];
      red[warp][w][g + 8][mt * 8 + t * 2] = acc[w][mt][2];
      red[warp][w][g + 8][mt * 8 + t * 2 + 1] = acc[w][mt][3];
    }
  __syncthreads();
  for (int idx = threadIdx.x; idx < SK_ROWS * MT * 8; idx += SK_THREADS) {
    const int r = idx % SK_ROWS;
    const int m = idx / SK_ROWS;
    const int n = n0 + r;
    if (m >= M || n >= p.N) continue;
    float v = 0.0f, v2 = 0.0f;
#pragma unroll
    for (int w = 0; w < SK_WARPS; ++w) {
      v += red[w][0][r][m];
      if (DUAL) v2 += red[w][NW - 1][r][m];
    }
    const bool f32 = p.out_f32 != 0;
    if (DUAL) {
      const float gte = bf16_round(v), up = bf16_round(v2);
      v = bf16_round(gte / (1.0f + __expf(-gte))) * up;
    }
    if (bias != nullptr) v += __bfloat162float(bias[n]);
    if (p.act != MPL_ACT_NONE) v = skinny_act(f32 ? v : bf16_round(v), p.act);
    if (p.row_scale != nullptr) v = (f32 ? v : bf16_round(v)) * p.row_scale[m];
    if (p.residual != nullptr)
      v = (f32 ? v : bf16_round(v)) + __bfloat162float(p.residual[static_cast<long long>(m) * p.ldr + n]);
    if (f32)
      reinterpret_cast<float*>(Cout)[static_cast<long long>(m) * p.ldc + n] = v;
    else
      reinterpret_cast<__nv_bfloat16*>(Cout)[static_cast<long long>(m) * p.ldc + n] = __float2bfloat16_rn(v);
  }
}

int skinny_gemm_bf16(const mpl_gemm_args& a, cudaStream_t stream) {
  if (a.M <= 0 || a.N <= 0) return MPL_OK;
  const int nb = a.nb < 1 ? 1 : a.nb;
  if (a.M > 16 || nb > 3 || (a.B2 != nullptr && nb != 1)) return MPL_ERR_UNSUPPORTED;
  if (a.K <= 0 || a.A == nullptr || a.B[0] == nullptr || a.C[0] == nullptr) return MPL_ERR_ARG;
  if ((a.K % 8) != 0 || (a.lda % 8) != 0 || (a.ldb % 8) != 0 || (reinterpret_cast<uintptr_t>(a.A) & 15) != 0 ||
      (reinterpret_cast<uintptr_t>(a.B[0]) & 15) != 0)
    return MPL_ERR_ALIGN;
  SkinnyParams p;
  p.A = static_cast<const __nv_bfloat16*>(a.A);
  p.lda = a.lda;
  for (int i = 0; i < 3; ++i) {
    p.Wn[i] = static_cast<const __nv_bfloat16*>(a.B[i < nb ? i : 0]);
    p.Cn[i] = a.C[i < nb ? i : 0];
    p.biasn[i] = static_cast<const __nv_bfloat16*>(a.bias[i < nb ? i : 0]);
  }
  p.W2 = static_cast<const __nv_bfloat16*>(a.B2);
  p.ldw = a.ldb;
  p.ldc = a.ldc;
  p.residual = static_cast<const __nv_bfloat16*>(a.residual);
  p.ldr = a.ldr;
  p.row_scale = a.row_scale;
  p.m_dev = a.m_dev;
  p.M = a.M;
  p.N = a.N;
  p.K = a.K;
  p.act = a.act;
  p.out_f32 = a.out_dtype == MPL_DT_F32;
  const dim3 grid((a.N + SK_ROWS - 1) / SK_ROWS, nb);
  const bool dual = a.B2 != nullptr;
  if (a.M <= 8) {
    if (dual)
      skinny_gemm_kernel<1, true><<<grid, SK_THREADS, 0, stream>>>(p);
    else
      skinny_gemm_kernel<1, false><<<grid, SK_THREADS, 0, stream>>>(p);
  } else {
    if (dual)
      skinny_gemm_kernel<2, true><<<grid, SK_THREADS, 0, stream>>>(p);
    else
      skinny_gemm_kernel<2, false><<<grid, SK_THREADS, 0, stream>>>(p);
  }
  return mpl::launch_status();
}

}  // namespace mpl

extern "C" int mpl_skinny_gemm_bf16(const mpl_gemm_args* args, void* stream) {
  if (args == nullptr) return MPL_ERR_ARG;
  return mpl::skinny_gemm_bf16(*args, static_cast<cudaStream_t>(stream));
}

// Dispatcher used by the host side: tensor-core tiles for M > 16, streaming kernel otherwise.
namespace mpl {
int linear_bf16(const mpl_gemm_args& a, cudaStream_t stream) {
  if (a.M <= 16 && (a.K % 8) == 0) return skinny_gemm_bf16(a, stream);
  return gemm_bf16(a, stream);
}
}  // namespace mpl
extern "C" int mpl_linear_bf16(const mpl_gemm_args* args, void* stream) {
  if (args == nullptr) return MPL_ERR_ARG;
  return mpl::linear_bf16(*args, static_cast<cudaStream_t>(stream));
}
