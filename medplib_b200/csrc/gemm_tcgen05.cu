// K1 — bf16 GEMM on the 5th-gen tensor cores:  C[M,N] = epilogue(A[M,K] · W[N,K]^T)
//
// Replaces every nn.Linear on the dense path of the reference (SURVEY.md §8a rows a-1,a-2,a-3,a-7,a-9,a-10,a-11;
// reference call sites e.g. model/medplib/model/multimodal_projector/builder.py:39-46 and the HF LlamaAttention /
// LlamaMLP linears driven from model/medplib/model/language_model/medplib_moe_llama.py:123-147).
//
// Design (one CTA per SM, persistent over output tiles, warp-specialised):
//   warp 0      TMA producer: cp.async.bulk.tensor 2-D boxes {64 x 128} of A and {64 x BN} of W, 128-B swizzle,
//               into a STAGES-deep shared-memory ring guarded by full/empty mbarriers
//   warp 1      allocates TMEM, then one lane issues tcgen05.mma (128 x BN x 16, bf16 -> fp32 in TMEM) and
//               tcgen05.commit's the ring slot back to the producer / the accumulator to the epilogue
//   warps 2..5  epilogue: tcgen05.ld the fp32 accumulator (each warp owns its 32-lane TMEM quadrant), apply
//               bias / activation / SiLU(gate)*up / residual, round to bf16 at the same points the reference's
//               eager bf16 path rounds, and store. Two accumulator stages let the epilogue of tile i overlap
//               the main loop of tile i+1.
// Tiles are ordered m-fastest so CTAs running concurrently share one W panel (W streams from HBM once, A from L2).
#include <cuda.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <type_traits>
#include <unordered_map>
#include <vector>

#include "internal.h"
#include "ptx.cuh"

namespace mpl {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int GEMM_THREADS = 192;
constexpr int LORA_STAGE_BYTES = 2 * 256 * 16;  // staged b rows of up to two rank-8 LoRA terms (launches that carry them)

struct GemmDevParams {
  int M, N, K;
  int nb;
  void* C[3];
  long long ldc;
  const __nv_bfloat16* bias[3];
  const __nv_bfloat16* residual;
  long long ldr;
  const float* row_scale;
  const int* m_dev;
  int act;
  int out_f32;
  int dual;
  // grouped mode (groups > 0): group g multiplies A rows [g * a_group_rows, +m_dev[g]) by its own weight matrix (tensor
  // maps in GroupMaps) into C + g * c_group_stride; ONE launch walks the tiles of all groups (m fastest inside a group)
  int groups;
  long long a_group_rows;
  long long c_group_stride;  // elements
  __nv_bfloat16* sb_g;  // SiLU*up backward in the epilogue: v = dh -> g <- dh u silu'(g), u <- dh silu(g) (in place), no C
  __nv_bfloat16* sb_u;
  __nv_bfloat16* dual_g;  // dual mode, optional: the rounded gate / up projections themselves (train forward keeps them)
  __nv_bfloat16* dual_u;
  int ext;  // 1: one extra k-block whose operands come from the extension maps (rank-r adapters inside the accumulator)
  // fused LoRA up-projections (see mpl_gemm_args): term t adds to output matrix lora_mat[t]
  const void* lora_u[2];
  const __nv_bfloat16* lora_b[2];
  float lora_scale[2];
  int lora_u_f32[2];
  int lora_mat[2];
  int lora_r;
};

struct GroupMaps {
  CUtensorMap b[MPL_MAX_EXPERTS];
  CUtensorMap b2[MPL_MAX_EXPERTS];
};

// tile index -> (group / weight matrix, first row, first column, rows of the group)
struct TileCoord {
  int g, m0, n0, M;
};
struct TileSpace {
  int groups, tiles_n1, nb, out_bn, num_tiles, bm;  // bm = rows of a tile: 128 (one CTA) or 256 (a CTA pair)
  int tm[MPL_MAX_EXPERTS], Mg[MPL_MAX_EXPERTS];
  int tiles_m, M;
  __device__ __forceinline__ TileCoord at(int tile) const {
    TileCoord c;
    if (groups > 0) {
      int g = 0, t = tile;
#pragma unroll 1
      while (g + 1 < groups && t >= tm[g] * tiles_n1) t -= tm[g++] * tiles_n1;
      c.g = g;
      c.m0 = (t % tm[g]) * bm;
      c.n0 = (t / tm[g]) * out_bn;
      c.M = Mg[g];
    } else {
      const int nt = tile / tiles_m;
      c.g = nt / tiles_n1;  // which weight matrix (nb > 1)
      c.m0 = (tile % tiles_m) * bm;
      c.n0 = (nt % tiles_n1) * out_bn;
      c.M = M;
    }
    return c;
  }
};
__device__ __forceinline__ TileSpace make_tile_space(const GemmDevParams& p, int out_bn, int bm) {
  TileSpace ts;
  ts.bm = bm;
  ts.groups = p.groups;
  ts.out_bn = out_bn;
  ts.nb = p.nb;
  ts.tiles_n1 = (p.N + out_bn - 1) / out_bn;
  if (p.groups > 0) {
    int total = 0;
#pragma unroll 1
    for (int g = 0; g < p.groups; ++g) {
      const int Mg = min(p.M, p.m_dev != nullptr ? p.m_dev[g] : p.M);
      ts.Mg[g] = Mg;
      ts.tm[g] = (Mg + bm - 1) / bm;
      total += ts.tm[g] * ts.tiles_n1;
    }
    ts.num_tiles = total;
    ts.tiles_m = 0;
    ts.M = 0;
  } else {
    int M = p.M;
    if (p.m_dev != nullptr) M = min(M, *p.m_dev);
    ts.M = M;
    ts.tiles_m = (M + bm - 1) / bm;
    ts.num_tiles = ts.tiles_m * ts.tiles_n1 * p.nb;
  }
  return ts;
}

// SiLU(gate) of the dual epilogue: fast divide by default (see apply_act); -DMPL_SILU_IEEE_DIV restores div.rn
#ifdef MPL_SILU_IEEE_DIV
#define MPL_SILU_DIV(a, b) ((a) / (b))
#else
#define MPL_SILU_DIV(a, b) __fdividef((a), (b))
#endif
__device__ __forceinline__ float apply_act(float v, int act) {
  switch (act) {
    case MPL_ACT_GELU:
      return 0.5f * v * (1.0f + erff(v * 0.70710678118654752f));
    // __fdividef (MUFU.RCP + FMUL, <= 2 ulp of fp32, no slow-path call): an IEEE divide carries a predicated CALL per
    // element that keeps the scheduler from interleaving the 32 elements of a chunk; the result is rounded to bf16
    case MPL_ACT_QUICK_GELU:
      return __fdividef(v, 1.0f + __expf(-1.702f * v));
    case MPL_ACT_RELU:
      return fmaxf(v, 0.0f);
    case MPL_ACT_SILU:
      return __fdividef(v, 1.0f + __expf(-v));
    case MPL_ACT_SIGMOID:
      return __fdividef(1.0f, 1.0f + __expf(-v));
    default:
      return v;
  }
}

// Epilogue of one accumulator row (one thread = one output row of the tile): tcgen05.ld 32 columns at a time, apply
// bias / activation / SiLU(gate)*up / row scale / residual with the reference's bf16 rounding points, store.
// t_row = TMEM address of this thread's lane at the first column of the accumulator.
template <int BN>
__device__ __forceinline__ void epilogue_rows(const GemmDevParams& p, const __nv_bfloat16* bias, void* Cout, int row, int M,
                                              int n0, int out_bn, uint32_t t_row, bool vec_ok, bool res_vec_ok,
                                              int which, uint8_t* s_lora) {
  const bool row_ok = row < M;
  // fused LoRA up-projections (rank 8): this row's coefficients (rounded to bf16 like the MMA operand of the unfused
  // kernel) once per tile; the tile's rows of b ([out_bn][8] bf16 = 16 B per column, per term) are staged in shared
  // memory by the four epilogue warps together -- read straight from global memory they cost one L2 round trip per
  // 8 columns on the epilogue's critical path
  float lu[2][8];
  bool lon[2] = {false, false};
  if (p.lora_r > 0) {
    const int et = static_cast<int>(threadIdx.x) - 64;  // 0..127 over the epilogue warps
    asm volatile("bar.sync 1, 128;" ::: "memory");       // the previous tile's reads of the staging area are done
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      const bool term = p.lora_u[t] != nullptr && p.lora_mat[t] == which;
      lon[t] = term && row_ok;
      if (term) {
        for (int cidx = et; cidx < out_bn; cidx += 128) {
          const int col = min(n0 + cidx, p.N - 1);
          *reinterpret_cast<uint4*>(s_lora + (t * 256 + cidx) * 16) =
              *reinterpret_cast<const uint4*>(p.lora_b[t] + static_cast<long long>(col) * 8);
        }
      }
      if (lon[t]) {
        if (p.lora_u_f32[t]) {
          const float4* up = reinterpret_cast<const float4*>(static_cast<const float*>(p.lora_u[t]) + static_cast<long long>(row) * 8);
          const float4 a0 = up[0], a1 = up[1];
          lu[t][0] = bf16_round(a0.x), lu[t][1] = bf16_round(a0.y), lu[t][2] = bf16_round(a0.z), lu[t][3] = bf16_round(a0.w);
          lu[t][4] = bf16_round(a1.x), lu[t][5] = bf16_round(a1.y), lu[t][6] = bf16_round(a1.z), lu[t][7] = bf16_round(a1.w);
        } else {
          const uint4 raw = *reinterpret_cast<const uint4*>(static_cast<const __nv_bfloat16*>(p.lora_u[t]) + static_cast<long long>(row) * 8);
          const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 f = __bfloat1622float2(h2[e]);
            lu[t][2 * e] = f.x;
            lu[t][2 * e + 1] = f.y;
          }
        }
      }
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");  // staging complete
  }
  const float rscale = (p.row_scale != nullptr && row_ok) ? p.row_scale[row] : 1.0f;
  const bool f32 = p.out_f32 != 0;
  const int nlim = min(p.N, n0 + out_bn);  // columns of this tile (the last 32-column chunk may be partial)
#pragma unroll 1
  for (int c = 0; c < (out_bn + 31) / 32; ++c) {
    uint32_t r[32];
    float v[32];
    tmem_ld_32x32(t_row + c * 32, r);
    if (p.dual) {
      uint32_t r2[32];
      tmem_ld_32x32(t_row + BN / 2 + c * 32, r2);
      tmem_ld_wait();
      if (p.dual_g != nullptr && row_ok && n0 + c * 32 < nlim) {
        // the train forward keeps gate(x) and up(x) for the backward: same rounding as the two separate GEMMs gave
        const int nc0 = n0 + c * 32;
        __nv_bfloat16* gp = p.dual_g + static_cast<long long>(row) * p.ldc + nc0;
        __nv_bfloat16* up = p.dual_u + static_cast<long long>(row) * p.ldc + nc0;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (nc0 + q * 8 + 8 <= nlim && vec_ok) {
            uint4 og, ou;
            og.x = pack_bf16(__uint_as_float(r[q * 8 + 0]), __uint_as_float(r[q * 8 + 1]));
            og.y = pack_bf16(__uint_as_float(r[q * 8 + 2]), __uint_as_float(r[q * 8 + 3]));
            og.z = pack_bf16(__uint_as_float(r[q * 8 + 4]), __uint_as_float(r[q * 8 + 5]));
            og.w = pack_bf16(__uint_as_float(r[q * 8 + 6]), __uint_as_float(r[q * 8 + 7]));
            ou.x = pack_bf16(__uint_as_float(r2[q * 8 + 0]), __uint_as_float(r2[q * 8 + 1]));
            ou.y = pack_bf16(__uint_as_float(r2[q * 8 + 2]), __uint_as_float(r2[q * 8 + 3]));
            ou.z = pack_bf16(__uint_as_float(r2[q * 8 + 4]), __uint_as_float(r2[q * 8 + 5]));
            ou.w = pack_bf16(__uint_as_float(r2[q * 8 + 6]), __uint_as_float(r2[q * 8 + 7]));
            *reinterpret_cast<uint4*>(gp + q * 8) = og;
            *reinterpret_cast<uint4*>(up + q * 8) = ou;
          } else {
#pragma unroll
            for (int j = q * 8; j < q * 8 + 8; ++j)
              if (nc0 + j < nlim) {
                gp[j] = __float2bfloat16_rn(__uint_as_float(r[j]));
                up[j] = __float2bfloat16_rn(__uint_as_float(r2[j]));
              }
          }
        }
      }
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        // reference: down(silu(gate(x)) * up(x)) with every intermediate rounded to bf16
        const float g = bf16_round(__uint_as_float(r[j]));
        const float u = bf16_round(__uint_as_float(r2[j]));
        const float s = bf16_round(MPL_SILU_DIV(g, 1.0f + __expf(-g)));
        v[j] = s * u;
      }
    } else {
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
    }
    const int nc = n0 + c * 32;
    if (nc >= nlim || !row_ok) continue;
    const bool full = (nc + 32 <= nlim);
    if (bias != nullptr) {
      if (full && (reinterpret_cast<uintptr_t>(bias + nc) & 15) == 0) {  // four 16-byte broadcast loads instead of 32 scalar ones
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const uint4 bw = *reinterpret_cast<const uint4*>(bias + nc + q * 8);
          const __nv_bfloat162* bh = reinterpret_cast<const __nv_bfloat162*>(&bw);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 f = __bfloat1622float2(bh[e]);
            v[q * 8 + 2 * e] += f.x;
            v[q * 8 + 2 * e + 1] += f.y;
          }
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (full || nc + j < nlim) v[j] += __bfloat162float(bias[nc + j]);
      }
    }
    if (p.act != MPL_ACT_NONE) {
      // The activation is chosen ONCE per chunk and each case is a straight-line unrolled loop: with the switch inside
      // the element loop every element was its own basic block, i.e. 32 serial exp / divide dependency chains on the
      // one epilogue warp a scheduler has (CLIP fc1 at M = 577: 42 us against 16 us without the activation).
      if (!f32) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = bf16_round(v[j]);
      }
      switch (p.act) {
#define MPL_ACT_LOOP(A)                                       \
  case A:                                                     \
    _Pragma("unroll") for (int j = 0; j < 32; ++j) v[j] = apply_act(v[j], A); \
    break;
        MPL_ACT_LOOP(MPL_ACT_GELU)
        MPL_ACT_LOOP(MPL_ACT_QUICK_GELU)
        MPL_ACT_LOOP(MPL_ACT_RELU)
        MPL_ACT_LOOP(MPL_ACT_SILU)
        MPL_ACT_LOOP(MPL_ACT_SIGMOID)
#undef MPL_ACT_LOOP
        default:
          break;
      }
    }
    if (p.row_scale != nullptr) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = (f32 ? v[j] : bf16_round(v[j])) * rscale;
    }
    if (p.lora_r > 0 && (lon[0] || lon[1])) {
      // v = bf16(bf16(v) + bf16(scale * bf16(sum_j u[j] b[n][j]))), b from the staged rows (broadcast reads)
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        if (!lon[t]) continue;
        const float sc = p.lora_scale[t];
        const uint8_t* sb = s_lora + (t * 256 + c * 32) * 16;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const uint4 bw = *reinterpret_cast<const uint4*>(sb + j * 16);
          const __nv_bfloat162* bh = reinterpret_cast<const __nv_bfloat162*>(&bw);
          float acc = 0.0f;
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 f = __bfloat1622float2(bh[e]);
            acc += lu[t][2 * e] * f.x + lu[t][2 * e + 1] * f.y;
          }
          v[j] = bf16_round(bf16_round(v[j]) + bf16_round(sc * bf16_round(acc)));
        }
      }
    }
    if (p.sb_g != nullptr) {
      // this GEMM is the dgrad of down_proj: v = dh. The SiLU(gate) * up backward consumes it right here (dh rounded to bf16
      // as the separate pass read it): dg = dh u s (1 + g (1 - s)), du = dh g s with s = sigmoid(g), written over g and u
      __nv_bfloat16* gp = p.sb_g + static_cast<long long>(row) * p.ldc + nc;
      __nv_bfloat16* up = p.sb_u + static_cast<long long>(row) * p.ldc + nc;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (nc + q * 8 + 8 <= nlim && vec_ok) {
          uint4 gr = *reinterpret_cast<const uint4*>(gp + q * 8), ur = *reinterpret_cast<const uint4*>(up + q * 8);
          __nv_bfloat162* g2 = reinterpret_cast<__nv_bfloat162*>(&gr);
          __nv_bfloat162* u2 = reinterpret_cast<__nv_bfloat162*>(&ur);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 gf = __bfloat1622float2(g2[e]), uf = __bfloat1622float2(u2[e]);
            const float d0 = bf16_round(v[q * 8 + 2 * e]), d1 = bf16_round(v[q * 8 + 2 * e + 1]);
            const float s0 = 1.0f / (1.0f + __expf(-gf.x)), s1 = 1.0f / (1.0f + __expf(-gf.y));
            g2[e] = __floats2bfloat162_rn(d0 * uf.x * s0 * (1.0f + gf.x * (1.0f - s0)), d1 * uf.y * s1 * (1.0f + gf.y * (1.0f - s1)));
            u2[e] = __floats2bfloat162_rn(d0 * gf.x * s0, d1 * gf.y * s1);
          }
          *reinterpret_cast<uint4*>(gp + q * 8) = gr;
          *reinterpret_cast<uint4*>(up + q * 8) = ur;
        } else {
#pragma unroll
          for (int j = q * 8; j < q * 8 + 8; ++j)
            if (nc + j < nlim) {
              const float gv = __bfloat162float(gp[j]), uv = __bfloat162float(up[j]), dv = bf16_round(v[j]);
              const float sg = 1.0f / (1.0f + __expf(-gv));
              gp[j] = __float2bfloat16_rn(dv * uv * sg * (1.0f + gv * (1.0f - sg)));
              up[j] = __float2bfloat16_rn(dv * gv * sg);
            }
        }
      }
      continue;  // nothing else leaves this tile
    }
    if (p.residual != nullptr) {
      if (!f32) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = bf16_round(v[j]);
      }
      const __nv_bfloat16* rp = p.residual + static_cast<long long>(row) * p.ldr + nc;
      if (full && res_vec_ok) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const uint4 rv = *reinterpret_cast<const uint4*>(rp + q * 8);
          const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&rv);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 f = __bfloat1622float2(h[e]);
            v[q * 8 + e * 2] += f.x;
            v[q * 8 + e * 2 + 1] += f.y;
          }
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (nc + j < nlim) v[j] += __bfloat162float(rp[j]);
      }
    }
    if (f32) {
      float* cp = reinterpret_cast<float*>(Cout) + static_cast<long long>(row) * p.ldc + nc;
      if (full && vec_ok) {
#pragma unroll
        for (int q = 0; q < 8; ++q)
          *reinterpret_cast<float4*>(cp + q * 4) = make_float4(v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (nc + j < nlim) cp[j] = v[j];
      }
    } else {
      __nv_bfloat16* cp = reinterpret_cast<__nv_bfloat16*>(Cout) + static_cast<long long>(row) * p.ldc + nc;
      if (full && vec_ok) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint4 o;
          o.x = pack_bf16(v[q * 8 + 0], v[q * 8 + 1]);
          o.y = pack_bf16(v[q * 8 + 2], v[q * 8 + 3]);
          o.z = pack_bf16(v[q * 8 + 4], v[q * 8 + 5]);
          o.w = pack_bf16(v[q * 8 + 6], v[q * 8 + 7]);
          *reinterpret_cast<uint4*>(cp + q * 8) = o;
        }
      } else if (vec_ok) {
        // partial chunk (tile widths that are not a multiple of 32, or the matrix edge): whole 8-column groups
        // still go out as 16-byte stores
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (nc + q * 8 + 8 <= nlim) {
            uint4 o;
            o.x = pack_bf16(v[q * 8 + 0], v[q * 8 + 1]);
            o.y = pack_bf16(v[q * 8 + 2], v[q * 8 + 3]);
            o.z = pack_bf16(v[q * 8 + 4], v[q * 8 + 5]);
            o.w = pack_bf16(v[q * 8 + 6], v[q * 8 + 7]);
            *reinterpret_cast<uint4*>(cp + q * 8) = o;
          } else {
#pragma unroll
            for (int j = q * 8; j < q * 8 + 8; ++j)
              if (nc + j < nlim) cp[j] = __float2bfloat16_rn(v[j]);
          }
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (nc + j < nlim) cp[j] = __float2bfloat16_rn(v[j]);
      }
    }
  }
}

template <int BN>
struct GemmCfg {
  // BN: any multiple of 16 in [128, 256] (the UMMA N of a 128-row tile); the launcher picks the width whose tile count
  // fills the 148 SMs best (M = 615 prefill: 145 tiles of 144 columns for o_proj instead of 110 of 192)
  static_assert(BN % 16 == 0 && BN >= 64 && BN <= 256, "tile width");
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGES_FIT = (232448 - 1024 - 256 - 512 /*static smem*/) / (A_BYTES + B_BYTES);
  static constexpr int STAGES = STAGES_FIT > 6 ? 6 : STAGES_FIT;
  static constexpr int ACC_STRIDE = BN <= 128 ? 128 : 256;  // column offset between the two accumulator stages
  static constexpr int TMEM_COLS = 2 * ACC_STRIDE;          // power of two >= 32
  static constexpr int SMEM_BYTES = STAGES * (A_BYTES + B_BYTES) + 1024 /*align slack*/ + 256 /*barriers*/;
};

template <int BN>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                         const __grid_constant__ CUtensorMap tmB1, const __grid_constant__ CUtensorMap tmB2,
                         const GemmDevParams p, const __grid_constant__ GroupMaps gmaps,
                         const __grid_constant__ CUtensorMap tmAx, const __grid_constant__ CUtensorMap tmBx) {
  using Cfg = GemmCfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * Cfg::A_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sB + STAGES * Cfg::B_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int out_bn = p.dual ? BN / 2 : BN;
  __shared__ TileSpace ts;  // (per-group tile counts are indexed dynamically: shared memory, not a local array)
  const int kblocks = (p.K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (p.dual || p.nb > 2) tma_prefetch_desc(&tmB2);
    if (p.nb > 1) tma_prefetch_desc(&tmB1);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], 4);  // one arrive per epilogue warp
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
  // programmatic dependent launch: everything above overlapped the previous kernel's tail; its outputs (A, the residual,
  // the device-side row counts) are touched only from here on
  griddep_wait();
  griddep_launch_dependents();
  if (threadIdx.x == 0) ts = make_tile_space(p, out_bn, BM);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int num_tiles = ts.num_tiles;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const TileCoord tc = ts.at(tile);
        const int which = tc.g, n0 = tc.n0;
        const int m0 = p.groups > 0 ? static_cast<int>(tc.g * p.a_group_rows) + tc.m0 : tc.m0;  // row in the A buffer
        for (int kb = 0; kb < kblocks + p.ext; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_expect_tx(&full_bar[stage], Cfg::A_BYTES + Cfg::B_BYTES);
          if (kb == kblocks) {
            // extension k-block: [u_0 | u_1 | ... | 0] x [s B_0 | s B_1 | ... | 0]^T -- the rank-r adapters of this output
            // enter the fp32 accumulator as 64 more columns of K (rows of the weight-side tensor: matrix / group major)
            tma_load_2d(sA + stage * Cfg::A_BYTES, &tmAx, &full_bar[stage], 0, m0);
            if (p.dual) {
              tma_load_2d(sB + stage * Cfg::B_BYTES, &tmBx, &full_bar[stage], 0, n0);
              tma_load_2d(sB + stage * Cfg::B_BYTES + Cfg::B_BYTES / 2, &tmBx, &full_bar[stage], 0, p.N + n0);
            } else {
              tma_load_2d(sB + stage * Cfg::B_BYTES, &tmBx, &full_bar[stage], 0, which * p.N + n0);
            }
            if (++stage == STAGES) {
              stage = 0;
              phase ^= 1;
            }
            continue;
          }
          tma_load_2d(sA + stage * Cfg::A_BYTES, &tmA, &full_bar[stage], kb * BK, m0);
          if (p.groups > 0) {
            // per-group weight maps: a __grid_constant__ array, indexed in param space
            if (p.dual) {
              tma_load_2d(sB + stage * Cfg::B_BYTES, &gmaps.b[which], &full_bar[stage], kb * BK, n0);
              tma_load_2d(sB + stage * Cfg::B_BYTES + Cfg::B_BYTES / 2, &gmaps.b2[which], &full_bar[stage], kb * BK, n0);
            } else {
              tma_load_2d(sB + stage * Cfg::B_BYTES, &gmaps.b[which], &full_bar[stage], kb * BK, n0);
            }
          } else if (p.dual) {
            tma_load_2d(sB + stage * Cfg::B_BYTES, &tmB, &full_bar[stage], kb * BK, n0);
            tma_load_2d(sB + stage * Cfg::B_BYTES + Cfg::B_BYTES / 2, &tmB2, &full_bar[stage], kb * BK, n0);
          } else {
            // explicit branches: the tensor maps must be addressed in param space (no local copies)
            if (which == 0)
              tma_load_2d(sB + stage * Cfg::B_BYTES, &tmB, &full_bar[stage], kb * BK, n0);
            else if (which == 1)
              tma_load_2d(sB + stage * Cfg::B_BYTES, &tmB1, &full_bar[stage], kb * BK, n0);
            else
              tma_load_2d(sB + stage * Cfg::B_BYTES, &tmB2, &full_bar[stage], kb * BK, n0);
          }
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (single thread)
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(BM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * Cfg::ACC_STRIDE;
        for (int kb = 0; kb < kblocks + p.ext; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(sA + stage * Cfg::A_BYTES);
          const uint32_t b_addr = smem_u32(sB + stage * Cfg::B_BYTES);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t da = umma_desc_k_sw128(a_addr + k * 32);
            const uint64_t db = umma_desc_k_sw128(b_addr + k * 32);
            umma_bf16_ss(d_tmem, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);  // slot reusable once these MMAs retire
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(&tfull_bar[acc]);  // accumulator complete
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue (warps 2..5)
    const int quad = warp & 3;  // TMEM lane quadrant this warp may read
    int acc = 0;
    uint32_t acc_phase = 0;
    const bool vec_ok = p.out_f32 ? ((p.ldc & 3) == 0) : ((p.ldc & 7) == 0);
    const bool res_vec_ok = (p.ldr & 7) == 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const TileCoord tc = ts.at(tile);
      const int which = tc.g, m0 = tc.m0, n0 = tc.n0, M = tc.M;
      const __nv_bfloat16* bias = p.groups > 0 ? nullptr : (which == 0 ? p.bias[0] : (which == 1 ? p.bias[1] : p.bias[2]));
      void* Cout = which == 0 ? p.C[0] : (which == 1 ? p.C[1] : p.C[2]);
      if (p.groups > 0)
        Cout = static_cast<char*>(p.C[0]) + static_cast<long long>(which) * p.c_group_stride * (p.out_f32 ? 4 : 2);
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      epilogue_rows<BN>(p, bias, Cout, m0 + quad * 32 + lane, M, n0, out_bn,
                        tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * Cfg::ACC_STRIDE, vec_ok, res_vec_ok,
                        which, reinterpret_cast<uint8_t*>(tmem_slot) + 64);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
}

// ------------------------------------------------------------------------------------------------ CTA-pair variant
// Same GEMM on `cta_group::2`: a cluster of two CTAs (one TPC) owns a 256 x BN tile. Each CTA stages ITS 128 rows of A
// and HALF of the weight rows (BN/2) per k-block -- (128 + BN/2) x 128 B instead of (128 + BN) x 128 B for the same
// 128 x BN of output per CTA -- and one thread of the leader (cluster rank 0) issues tcgen05.mma.cta_group::2
// (M = 256, N = BN), which reads both CTAs' shared memory and writes 128 accumulator lanes into each CTA's TMEM. The
// kernel exists because the one-CTA kernel is bound by the L2 -> shared-memory fill, not by the tensor pipe
// (profiles/r02_gemm_tile_widths.md): a third fewer staged bytes per MAC at BN = 256.
//   full[s]    lives in the leader: ONE arrive.expect_tx by the leader's producer for the bytes of BOTH CTAs; the peer's TMA
//              loads complete on the leader's barrier (cp.async.bulk.tensor.cta_group::2)
//   empty[s]   in each CTA, signalled by the leader's tcgen05.commit multicast to both CTAs
//   tfull[a]   in each CTA (multicast commit); tempty[a] in the leader, 8 arrivals = 4 epilogue warps x 2 CTAs
// SiLU(gate)*up: the leader stages the gate rows, the peer the up rows of the same features, so the accumulator columns
// are [gate | up] exactly as in the one-CTA kernel and the epilogue is shared.
template <int BN>
struct PairCfg {
  static_assert(BN % 16 == 0 && BN >= 128 && BN <= 256, "pair tile width");
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = (BN / 2) * BK * 2;  // per CTA
  static constexpr int STAGES_FIT = (232448 - 1024 - 256 - 512 - LORA_STAGE_BYTES) / (A_BYTES + B_BYTES);
  static constexpr int STAGES = STAGES_FIT > 7 ? 7 : STAGES_FIT;
  static constexpr int ACC_STRIDE = BN <= 128 ? 128 : 256;
  static constexpr int TMEM_COLS = 2 * ACC_STRIDE;
  static constexpr int SMEM_BYTES = STAGES * (A_BYTES + B_BYTES) + 1024 + 256;
};

template <int BN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_tcgen05_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                              const __grid_constant__ CUtensorMap tmB1, const __grid_constant__ CUtensorMap tmB2,
                              const GemmDevParams p, const __grid_constant__ GroupMaps gmaps,
                              const __grid_constant__ CUtensorMap tmAx, const __grid_constant__ CUtensorMap tmBx) {
  using Cfg = PairCfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * Cfg::A_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sB + STAGES * Cfg::B_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();  // 0 = leader (issues the MMAs), 1 = peer
  const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;

  const int out_bn = p.dual ? BN / 2 : BN;
  __shared__ TileSpace ts;
  const int kblocks = (p.K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (p.dual || p.nb > 2) tma_prefetch_desc(&tmB2);
    if (p.nb > 1) tma_prefetch_desc(&tmB1);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], 8);  // 4 epilogue warps of each CTA (only the leader's copy is used)
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc_pair<Cfg::TMEM_COLS>(tmem_slot);
  griddep_wait();
  griddep_launch_dependents();
  if (threadIdx.x == 0) ts = make_tile_space(p, out_bn, 2 * BM);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // both CTAs' barriers are initialised and their TMEM allocated before anything crosses the pair
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int num_tiles = ts.num_tiles;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (one thread in EACH CTA)
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs) {
        const TileCoord tc = ts.at(tile);
        const int which = tc.g;
        const int m0 = (p.groups > 0 ? static_cast<int>(tc.g * p.a_group_rows) + tc.m0 : tc.m0) + static_cast<int>(rank) * BM;
        // weight rows of this CTA: the two halves of the tile's columns, or (dual) gate rows in the leader / up rows in the peer
        const int nrow = p.dual ? tc.n0 : tc.n0 + static_cast<int>(rank) * (BN / 2);
        const bool second = p.dual && rank == 1;  // the peer reads W2 (up)
        for (int kb = 0; kb < kblocks + p.ext; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          const uint32_t fb = mapa_shared(smem_u32(&full_bar[stage]), 0);
          if (rank == 0) mbar_expect_tx(&full_bar[stage], 2 * (Cfg::A_BYTES + Cfg::B_BYTES));
          uint8_t* dstB = sB + stage * Cfg::B_BYTES;
          if (kb == kblocks) {  // extension k-block (see the one-CTA kernel)
            tma_load_2d_pair(sA + stage * Cfg::A_BYTES, &tmAx, fb, 0, m0);
            tma_load_2d_pair(dstB, &tmBx, fb, 0, p.dual ? static_cast<int>(rank) * p.N + tc.n0 : which * p.N + nrow);
            if (++stage == STAGES) {
              stage = 0;
              phase ^= 1;
            }
            continue;
          }
          tma_load_2d_pair(sA + stage * Cfg::A_BYTES, &tmA, fb, kb * BK, m0);
          if (p.groups > 0) {
            if (second)
              tma_load_2d_pair(dstB, &gmaps.b2[which], fb, kb * BK, nrow);
            else
              tma_load_2d_pair(dstB, &gmaps.b[which], fb, kb * BK, nrow);
          } else if (second) {
            tma_load_2d_pair(dstB, &tmB2, fb, kb * BK, nrow);
          } else if (which == 0) {
            tma_load_2d_pair(dstB, &tmB, fb, kb * BK, nrow);
          } else if (which == 1) {
            tma_load_2d_pair(dstB, &tmB1, fb, kb * BK, nrow);
          } else {
            tma_load_2d_pair(dstB, &tmB2, fb, kb * BK, nrow);
          }
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer: one thread of the LEADER
    if (lane == 0 && rank == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(2 * BM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs) {
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * Cfg::ACC_STRIDE;
        for (int kb = 0; kb < kblocks + p.ext; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(sA + stage * Cfg::A_BYTES);
          const uint32_t b_addr = smem_u32(sB + stage * Cfg::B_BYTES);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t da = umma_desc_k_sw128(a_addr + k * 32);
            const uint64_t db = umma_desc_k_sw128(b_addr + k * 32);
            umma_bf16_ss_pair(d_tmem, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit_pair(&empty_bar[stage], 3);  // the slot is reusable in BOTH CTAs once these MMAs retire
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit_pair(&tfull_bar[acc], 3);  // accumulator complete (each CTA's epilogue waits on its own copy)
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue (warps 2..5 of each CTA: its own 128 rows)
    const int quad = warp & 3;
    int acc = 0;
    uint32_t acc_phase = 0;
    const bool vec_ok = p.out_f32 ? ((p.ldc & 3) == 0) : ((p.ldc & 7) == 0);
    const bool res_vec_ok = (p.ldr & 7) == 0;
    for (int tile = pair; tile < num_tiles; tile += num_pairs) {
      const TileCoord tc = ts.at(tile);
      const int which = tc.g, m0 = tc.m0 + static_cast<int>(rank) * BM, n0 = tc.n0, M = tc.M;
      const __nv_bfloat16* bias = p.groups > 0 ? nullptr : (which == 0 ? p.bias[0] : (which == 1 ? p.bias[1] : p.bias[2]));
      void* Cout = which == 0 ? p.C[0] : (which == 1 ? p.C[1] : p.C[2]);
      if (p.groups > 0)
        Cout = static_cast<char*>(p.C[0]) + static_cast<long long>(which) * p.c_group_stride * (p.out_f32 ? 4 : 2);
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      epilogue_rows<BN>(p, bias, Cout, m0 + quad * 32 + lane, M, n0, out_bn,
                        tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * Cfg::ACC_STRIDE, vec_ok, res_vec_ok,
                        which, reinterpret_cast<uint8_t*>(tmem_slot) + 64);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_shared(smem_u32(&tempty_bar[acc]), 0));
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // nobody leaves (or frees TMEM) while the other CTA may still signal its barriers / read its smem
  if (warp == 1) tmem_dealloc_pair<Cfg::TMEM_COLS>(tmem_base);
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(sym);
  });
  return fn;
}

// 2-D bf16 row-major tensor [rows, cols] with leading dimension ld (elements); box = {64 cols, box_rows}.
static int make_tmap(CUtensorMap* out, const void* ptr, long long rows, long long cols, long long ld, int box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) return MPL_ERR_DRIVER;
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0 || ((ld * 2) & 15) != 0) return MPL_ERR_ALIGN;
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * 2};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(BK), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? MPL_OK : MPL_ERR_DRIVER;
}

// Generic bf16 tensor-map encoder shared with the streaming GEMM (128-B swizzle, 256-B L2 promotion, zero OOB fill).
// strides_bytes has rank-1 entries (dimension 0 is contiguous).
int encode_tmap_bf16(void* out, const void* ptr, int rank, const unsigned long long* dims,
                     const unsigned long long* strides_bytes, const unsigned* box) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) return MPL_ERR_DRIVER;
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0) return MPL_ERR_ALIGN;
  cuuint64_t d[5], st[4];
  cuuint32_t b[5], es[5];
  for (int i = 0; i < rank; ++i) {
    d[i] = dims[i];
    b[i] = box[i];
    es[i] = 1;
    if (i + 1 < rank) {
      st[i] = strides_bytes[i];
      if (st[i] & 15) return MPL_ERR_ALIGN;
    }
  }
  CUresult r = fn(static_cast<CUtensorMap*>(out), CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(ptr), d, st,
                  b, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? MPL_OK : MPL_ERR_DRIVER;
}

// Optional in-situ timing of every tcgen05 GEMM launch (bench.py roofline): CUDA events recorded on the launching
// stream around each launch, summed by mpl_profile_gemm_read after a synchronise.
static bool g_prof = false;
static bool g_prof_suppress = false;        // set while a fork/join region is timed as ONE interval by its caller
static std::vector<cudaEvent_t> g_prof_ev;  // pairs (start, stop)
static size_t g_prof_used = 0;
static cudaEvent_t prof_event() {
  if (g_prof_used == g_prof_ev.size()) {
    cudaEvent_t e;
    cudaEventCreate(&e);
    g_prof_ev.push_back(e);
  }
  return g_prof_ev[g_prof_used++];
}

static int g_num_sms = 0;
int num_sms() {
  if (g_num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  return g_num_sms;
}

// Tile widths compiled in (UMMA N: any multiple of 16; these cover the wave counts that matter). One-CTA tiles are
// 128 x W, CTA-pair tiles 256 x W.
#define MPL_TILE_WIDTHS(X) X(64) X(96) X(128) X(144) X(160) X(176) X(192) X(208) X(224) X(240) X(256)
#define MPL_PAIR_WIDTHS(X) X(128) X(160) X(176) X(192) X(224) X(240) X(256)
constexpr int PAIR_FLAG = 1000;  // tile_n = PAIR_FLAG + W selects the CTA-pair kernel with width W

// Pick the tile shape with the lowest (waves x per-tile time) estimate for an [rows, N] output (x `mats` weight matrices /
// `groups` row groups of `rows` each) on a persistent grid of one CTA (or one CTA pair) per SM (per TPC). Measured on
// B200 (tools/gemm_tiles.py, profiles/r02_gemm_tile_widths.md): at prefill sizes the kernels are bound by the L2 ->
// shared-memory fill (~100 GB/s per SM, 13-15 TB/s over the chip), so a tile costs ~ the rows it stages per k-block --
// 128 + W for one CTA, 128 + W/2 per CTA of a pair -- plus a fixed part; one-CTA widths that leave only 4 ring stages
// (240, 256) pay ~10 % more. Ties go to the wider tile. MPL_GEMM_PAIR=0 keeps the one-CTA kernel.
static bool pair_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("MPL_GEMM_PAIR");
    on = (e == nullptr || atoi(e) != 0) ? 1 : 0;
  }
  return on != 0;
}
static int pair_min_rows() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("MPL_GEMM_PAIR_MIN_ROWS");
    v = e != nullptr ? atoi(e) : 640;  // (measured: 650 rows per expert at T = 1299 already favour pairs, 615 do not)
  }
  return v;
}
static int pick_tile_n(long long rows, int row_sets, int N, int dual, bool lora = false) {
  const int sms = num_sms();
  static const int widths[] = {
#define MPL_BN_LIST(W) W,
      MPL_TILE_WIDTHS(MPL_BN_LIST)
  };
  static const int pair_widths[] = {MPL_PAIR_WIDTHS(MPL_BN_LIST)
#undef MPL_BN_LIST
  };
  int best = 256;
  double best_cost = 1e300;
  for (int i = static_cast<int>(sizeof(widths) / sizeof(widths[0])) - 1; i >= 0; --i) {
    const int bn = widths[i];
    if (lora && bn == 224) continue;  // 5 stages of 44 KB leave no room for the staged LoRA rows
    const int out_bn = dual ? bn / 2 : bn;
    const long long tiles = row_sets * ((rows + BM - 1) / BM) * ((N + out_bn - 1) / out_bn);
    const long long waves = (tiles + sms - 1) / sms;
    const double c = static_cast<double>(waves) * (128 + bn + 22) * (bn > 224 ? 1.1 : 1.0);
    if (c < best_cost) {
      best = bn;
      best_cost = c;
    }
  }
  // CTA pairs pay off once a weight panel is re-used by enough row tiles to be served from L2 (M = 5112: 1.45-1.6
  // PFLOP/s against 1.2-1.3 for one CTA); at M = 615 the weights stream from DRAM in 128-byte pieces per row and both
  // kernels sit at the same ~2.5 us stage round trip (profiles/r02_gemm_tile_widths.md)
  if (pair_enabled() && rows >= pair_min_rows()) {
    const int pairs = sms / 2;
    for (int i = static_cast<int>(sizeof(pair_widths) / sizeof(pair_widths[0])) - 1; i >= 0; --i) {
      const int bn = pair_widths[i];
      const int out_bn = dual ? bn / 2 : bn;
      const long long tiles = row_sets * ((rows + 2 * BM - 1) / (2 * BM)) * ((N + out_bn - 1) / out_bn);
      const long long waves = (tiles + pairs - 1) / pairs;
      const double c = static_cast<double>(waves) * (128 + bn / 2 + 22);
      if (c < best_cost) {
        best = PAIR_FLAG + bn;
        best_cost = c;
      }
    }
  }
  return best;
}

template <int BN, bool PAIR>
static int launch_any(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmB1, const CUtensorMap& tmB2,
                      const GemmDevParams& p, const GroupMaps& gm, long long tiles, cudaStream_t stream,
                      const CUtensorMap* tmAx = nullptr, const CUtensorMap* tmBx = nullptr) {
  using Cfg = std::conditional_t<PAIR, PairCfg<BN>, GemmCfg<BN>>;  // (only the selected kernel is instantiated)
  constexpr int smem = Cfg::SMEM_BYTES;
  auto kernel = [] {
    if constexpr (PAIR)
      return &gemm_bf16_tcgen05_pair_kernel<BN>;
    else
      return &gemm_bf16_tcgen05_kernel<BN>;
  }();
  constexpr int smem_max = smem + LORA_STAGE_BYTES <= 232448 ? smem + LORA_STAGE_BYTES : smem;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max) != cudaSuccess) return MPL_ERR_CUDA;
    attr_set = true;
  }
  const int smem_launch = smem + (p.lora_r > 0 ? LORA_STAGE_BYTES : 0);
  if (smem_launch > smem_max) return MPL_ERR_UNSUPPORTED;  // (one-CTA width 224: no room for the staged LoRA rows)
  const int units = PAIR ? num_sms() / 2 : num_sms();  // CTAs, or CTA pairs (one per TPC)
  int grid = static_cast<int>(tiles < units ? tiles : units);
  if (grid < 1) grid = 1;
  if (PAIR) grid *= 2;
  const bool prof = g_prof && !g_prof_suppress;
  if (prof) cudaEventRecord(prof_event(), stream);
  launch_pdl(kernel, dim3(grid), dim3(GEMM_THREADS), smem_launch, stream, tmA, tmB, tmB1, tmB2, p, gm,
             tmAx != nullptr ? *tmAx : tmA, tmBx != nullptr ? *tmBx : tmB);
  if (prof) cudaEventRecord(prof_event(), stream);
  return mpl::launch_status();
}

template <int BN, bool PAIR>
static int launch_gemm(const mpl_gemm_args& a, cudaStream_t stream) {
  const int dual = a.B2 != nullptr;
  const int nb = a.nb < 1 ? 1 : a.nb;
  if (nb > 3 || (dual && nb != 1)) return MPL_ERR_ARG;
  const int box_b = (PAIR || dual) ? BN / 2 : BN;  // weight rows per TMA box
  CUtensorMap tmA, tmB, tmB1, tmB2;
  int rc = make_tmap(&tmA, a.A, a.M, a.K, a.lda, BM);
  if (rc) return rc;
  rc = make_tmap(&tmB, a.B[0], a.N, a.K, a.ldb, box_b);
  if (rc) return rc;
  tmB1 = tmB;
  tmB2 = tmB;
  if (dual) {
    rc = make_tmap(&tmB2, a.B2, a.N, a.K, a.ldb, box_b);
    if (rc) return rc;
  }
  if (nb > 1) {
    rc = make_tmap(&tmB1, a.B[1], a.N, a.K, a.ldb, box_b);
    if (rc) return rc;
  }
  if (nb > 2) {
    rc = make_tmap(&tmB2, a.B[2], a.N, a.K, a.ldb, box_b);
    if (rc) return rc;
  }
  GemmDevParams p;
  memset(&p, 0, sizeof(p));
  p.M = a.M;
  p.N = a.N;
  p.K = a.K;
  p.nb = nb;
  for (int i = 0; i < 3; ++i) {
    p.C[i] = a.C[i];
    p.bias[i] = static_cast<const __nv_bfloat16*>(a.bias[i]);
  }
  p.ldc = a.ldc;
  p.residual = static_cast<const __nv_bfloat16*>(a.residual);
  p.ldr = a.ldr;
  p.row_scale = a.row_scale;
  p.m_dev = a.m_dev;
  p.act = a.act;
  p.out_f32 = a.out_dtype == MPL_DT_F32;
  p.dual = dual;
  if (a.silu_bwd_g != nullptr || a.silu_bwd_u != nullptr) {
    if (dual || nb != 1 || a.silu_bwd_g == nullptr || a.silu_bwd_u == nullptr || a.residual != nullptr) return MPL_ERR_ARG;
    p.sb_g = static_cast<__nv_bfloat16*>(a.silu_bwd_g);
    p.sb_u = static_cast<__nv_bfloat16*>(a.silu_bwd_u);
  }
  if (a.dual_g != nullptr || a.dual_u != nullptr) {
    if (!dual || a.dual_g == nullptr || a.dual_u == nullptr || a.out_dtype == MPL_DT_F32) return MPL_ERR_ARG;
    p.dual_g = static_cast<__nv_bfloat16*>(a.dual_g);
    p.dual_u = static_cast<__nv_bfloat16*>(a.dual_u);
  }
  if (a.lora_r != 0) {
    if (a.lora_r != 8 || dual || a.out_dtype == MPL_DT_F32) return MPL_ERR_UNSUPPORTED;
    p.lora_r = a.lora_r;
    for (int t = 0; t < 2; ++t) {
      if (a.lora_u[t] != nullptr && (a.lora_b[t] == nullptr || (reinterpret_cast<uintptr_t>(a.lora_b[t]) & 15) != 0))
        return MPL_ERR_ALIGN;
      p.lora_u[t] = a.lora_u[t];
      p.lora_b[t] = static_cast<const __nv_bfloat16*>(a.lora_b[t]);
      p.lora_scale[t] = a.lora_scale[t];
      p.lora_u_f32[t] = a.lora_u_f32[t];
      p.lora_mat[t] = a.lora_mat[t];
    }
  }
  const int out_bn = dual ? BN / 2 : BN;
  const int bm = PAIR ? 2 * BM : BM;
  const long long tiles = static_cast<long long>((a.M + bm - 1) / bm) * ((a.N + out_bn - 1) / out_bn) * nb;
  static const GroupMaps no_groups = {};
  CUtensorMap tmAx, tmBx;
  if (a.ext_a != nullptr) {
    // rank-r adapters as 64 more columns of K: ext_a bf16 [M, 64], ext_b bf16 [(nb, or 2 with B2) * N, 64], zero where unused
    if (a.ext_b == nullptr) return MPL_ERR_ARG;
    rc = make_tmap(&tmAx, a.ext_a, a.M, BK, BK, BM);
    if (rc == MPL_OK) rc = make_tmap(&tmBx, a.ext_b, static_cast<long long>(dual ? 2 : nb) * a.N, BK, BK, box_b);
    if (rc) return rc;
    p.ext = 1;
    return launch_any<BN, PAIR>(tmA, tmB, tmB1, tmB2, p, no_groups, tiles, stream, &tmAx, &tmBx);
  }
  return launch_any<BN, PAIR>(tmA, tmB, tmB1, tmB2, p, no_groups, tiles, stream);
}

int gemm_bf16(const mpl_gemm_args& a, cudaStream_t stream) {
  if (a.M <= 0 || a.N <= 0) return MPL_OK;
  if (a.K <= 0 || a.A == nullptr || a.B[0] == nullptr || (a.C[0] == nullptr && a.silu_bwd_g == nullptr)) return MPL_ERR_ARG;
  const int dual = a.B2 != nullptr;
  const int nb = a.nb < 1 ? 1 : a.nb;
  int bn = a.tile_n;
  if (bn == 0) bn = pick_tile_n(a.M, nb, a.N, dual, a.lora_r != 0);
  switch (bn) {
#define MPL_BN_CASE(W) \
  case W:              \
    return launch_gemm<W, false>(a, stream);
    MPL_TILE_WIDTHS(MPL_BN_CASE)
#undef MPL_BN_CASE
#define MPL_BN_CASE(W)   \
  case PAIR_FLAG + W:    \
    return launch_gemm<W, true>(a, stream);
    MPL_PAIR_WIDTHS(MPL_BN_CASE)
#undef MPL_BN_CASE
    default:
      return MPL_ERR_ARG;
  }
}

template <int BN, bool PAIR>
static int launch_grouped(const mpl_grouped_gemm_args& a, cudaStream_t stream) {
  const int dual = a.B2[0] != nullptr;
  const int box_b = (PAIR || dual) ? BN / 2 : BN;
  const long long group_rows = a.a_group_stride / a.lda;
  CUtensorMap tmA;
  GroupMaps gm;
  memset(&gm, 0, sizeof(gm));
  // ONE map over the whole expert buffer: group g starts at row g * group_rows; rows past a group's count are read
  // (they belong to the next group or are zero-filled past the end) but never stored
  int rc = make_tmap(&tmA, a.A, group_rows * (a.groups - 1) + a.M, a.K, a.lda, BM);
  for (int g = 0; g < a.groups && rc == MPL_OK; ++g) {
    rc = make_tmap(&gm.b[g], a.B[g], a.N, a.K, a.ldb, box_b);
    if (rc == MPL_OK && dual) rc = make_tmap(&gm.b2[g], a.B2[g], a.N, a.K, a.ldb, box_b);
  }
  if (rc != MPL_OK) return rc;
  GemmDevParams p;
  memset(&p, 0, sizeof(p));
  p.M = a.M;
  p.N = a.N;
  p.K = a.K;
  p.nb = 1;
  p.C[0] = a.C;
  p.ldc = a.ldc;
  p.residual = static_cast<const __nv_bfloat16*>(a.residual);
  p.ldr = a.ldr;
  p.m_dev = a.m_dev;
  p.act = a.act;
  p.out_f32 = a.out_dtype == MPL_DT_F32;
  p.dual = dual;
  p.groups = a.groups;
  p.a_group_rows = group_rows;
  p.c_group_stride = a.c_group_stride;
  const int out_bn = dual ? BN / 2 : BN;
  const int bm = PAIR ? 2 * BM : BM;
  const long long tiles = static_cast<long long>((a.M + bm - 1) / bm) * ((a.N + out_bn - 1) / out_bn) * a.groups;
  return launch_any<BN, PAIR>(tmA, gm.b[0], gm.b[0], gm.b[0], p, gm, tiles, stream);
}

// Grouped (per-expert) GEMM for M > 16: ONE tcgen05 launch walks the tiles of every group (device-side row counts via
// m_dev; m fastest inside a group so concurrently running CTAs share a weight panel). Round 1 launched one kernel per
// expert on forked streams: at prefill sizes every such launch was a partial wave (profiles/r01_ncu_gemm_prefill.md).
int grouped_gemm_bf16(const mpl_grouped_gemm_args& a, cudaStream_t stream) {
  if (a.row_map != nullptr || a.a_row_map != nullptr) return MPL_ERR_UNSUPPORTED;
  if (a.groups < 1 || a.groups > MPL_MAX_EXPERTS || a.M <= 0 || a.N <= 0) return a.groups < 1 || a.M <= 0 ? MPL_OK : MPL_ERR_ARG;
  if (a.K <= 0 || a.A == nullptr || a.C == nullptr || a.lda <= 0 || a.a_group_stride % a.lda != 0) return MPL_ERR_ARG;
  for (int g = 0; g < a.groups; ++g)
    if (a.B[g] == nullptr) return MPL_ERR_ARG;
  // tile shape for the EXPECTED rows per group (rows spread evenly over the groups; m_total_hint = rows of all groups
  // together, default = the capacity of every group); the true per-group counts are read on the device
  const int dual = a.B2[0] != nullptr;
  const long long rows_total = a.m_total_hint > 0 ? a.m_total_hint : static_cast<long long>(a.M) * a.groups;
  long long per_group = (rows_total + a.groups - 1) / a.groups;
  if (per_group > a.M) per_group = a.M;
  int bn = a.tile_n;
  if (bn == 0) bn = pick_tile_n(per_group, a.groups, a.N, dual);
  switch (bn) {
#define MPL_BN_CASE(W) \
  case W:              \
    return launch_grouped<W, false>(a, stream);
    MPL_TILE_WIDTHS(MPL_BN_CASE)
#undef MPL_BN_CASE
#define MPL_BN_CASE(W)   \
  case PAIR_FLAG + W:    \
    return launch_grouped<W, true>(a, stream);
    MPL_PAIR_WIDTHS(MPL_BN_CASE)
#undef MPL_BN_CASE
    default:
      return MPL_ERR_ARG;
  }
}

}  // namespace mpl

extern "C" int mpl_profile_gemm(int enable) {
  mpl::g_prof = enable != 0;
  mpl::g_prof_used = 0;
  return MPL_OK;
}
extern "C" int mpl_profile_gemm_read(float* total_ms, int* launches) {
  if (cudaDeviceSynchronize() != cudaSuccess) return MPL_ERR_CUDA;
  float sum = 0.0f;
  for (size_t i = 0; i + 1 < mpl::g_prof_used; i += 2) {
    float ms = 0.0f;
    if (cudaEventElapsedTime(&ms, mpl::g_prof_ev[i], mpl::g_prof_ev[i + 1]) != cudaSuccess) return MPL_ERR_CUDA;
    sum += ms;
  }
  if (total_ms) *total_ms = sum;
  if (launches) *launches = static_cast<int>(mpl::g_prof_used / 2);
  mpl::g_prof_used = 0;
  return MPL_OK;
}

extern "C" int mpl_gemm_bf16(const mpl_gemm_args* args, void* stream) {
  if (args == nullptr) return MPL_ERR_ARG;
  return mpl::gemm_bf16(*args, static_cast<cudaStream_t>(stream));
}
