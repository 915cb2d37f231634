// K2-tc — flash attention forward on the 5th-gen tensor cores:  O = softmax(scale * Q K^T + masks) V
//
// Replaces, for head_dim 64 / 128 and Tq >= 32, the eager matmul-softmax-matmul of HF LlamaAttention (causal + key
// padding, prefill and training forward; medplib_moe_llama.py:126-135) and HF CLIPAttention
// (clip_encoder.py:53-57, App. A.2). Same contract as mpl_attention (fp32 softmax, P rounded to bf16 before P.V like
// `softmax(...).to(bf16) @ v`, optional log-sum-exp for the backward).
//
// One CTA per (128-query tile, batch x head), six warps:
//   warp 0      TMA producer: Q once, then K_j / V_j tiles of 128 keys into a two-stage ring (4-D tensor maps over the
//               caller's strided [B, T, H, d] / KV-cache [B, H, Tmax, d] tensors, 128-byte swizzle, zero fill past T)
//   warp 1      tcgen05.mma issuer (one lane): S_j = Q K_j^T (128 x 128 x d, K-major A and B) into one of two TMEM
//               accumulators; O += P_j V_j (128 x d x 128) with V as an MN-MAJOR B operand (V stays [keys, d] as it
//               lies in memory: no transpose pass) and P_j read from shared memory
//   warps 2..5  softmax: thread r owns query row r = TMEM lane r. tcgen05.ld the scores (two passes: row max, then
//               exp2 / row sum from ONE TMEM read of the row), write P_j as bf16 into the swizzled K-major tile the MMA
//               reads. The fp32 output accumulates in TMEM (O += P_j V_j) and is rescaled there, lazily, only when a
//               row maximum jumps by more than 2^8; the P.V MMA of tile j overlaps the score pass of tile j + 1.
// S_{j+1} is issued as soon as K_{j+1} has landed, i.e. the tensor core computes the next scores while the softmax
// warps are busy with the current ones.
#include <cuda.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "internal.h"
#include "ptx.cuh"

namespace mpl {

constexpr int AT_BM = 128;  // queries per CTA
constexpr int AT_BN = 128;  // keys per tile
constexpr int AT_THREADS = 192;

struct AttnTcParams {
  __nv_bfloat16* o;
  long long o_sb, o_st, o_sh;
  float* lse;
  const unsigned char* kv_mask;
  long long kv_mask_stride;
  int B, H, Tq, Tk, causal;
  float scale_log2;
};

__device__ __forceinline__ void tma_load_4d(void* smem_dst, const void* tmap, uint64_t* bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::
          "r"(smem_u32(smem_dst)),
      "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

// MN-major operand tile in shared memory, 128-byte swizzle: the tile is [K rows][64 MN elements = 128 B] per 64-wide
// MN block (what a TMA box {64, rows} writes), 8-row swizzle atoms of 1024 B along K (stride byte offset), MN blocks
// `lbo_bytes` apart (leading byte offset). Canonical form ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units.
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
__device__ __forceinline__ float ex2_approx(float x) {  // one MUFU: 2^x, flush-to-zero, 2^-inf = 0
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__host__ __device__ constexpr uint32_t umma_idesc_bf16_bmn(uint32_t m, uint32_t n) {
  return umma_idesc_bf16(m, n) | (1u << 16);  // b_major = MN
}

template <int D>
struct AttnTcCfg {
  static constexpr int KB = D / 64;                    // 64-wide blocks of the head dimension
  static constexpr int Q_BYTES = AT_BM * D * 2;        // [KB][128 rows][128 B]
  static constexpr int KV_BYTES = AT_BN * D * 2;       // one of K / V
  static constexpr int P_BYTES = AT_BM * AT_BN * 2;    // [2 key blocks][128 rows][128 B]
  static constexpr int STAGES = 2;
  static constexpr int SMEM_BYTES = Q_BYTES + STAGES * 2 * KV_BYTES + 2 * P_BYTES + 1024 + 256;  // P double-buffered
  static constexpr int TMEM_COLS = 512;  // S0 | S1 | PV (128 + 128 + D columns), power of two
};

template <int D>
__global__ void __launch_bounds__(AT_THREADS, 1)
attn_fwd_tcgen05_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                        const __grid_constant__ CUtensorMap tmV, const AttnTcParams p) {
  using Cfg = AttnTcCfg<D>;
  constexpr int KB = Cfg::KB;
  extern __shared__ uint8_t at_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(at_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sKV = sQ + Cfg::Q_BYTES;  // stage s: K at s * 2 * KV_BYTES, V right after it
  uint8_t* sP = sKV + Cfg::STAGES * 2 * Cfg::KV_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * Cfg::P_BYTES);
  uint64_t* q_full = bars;
  uint64_t* k_full = bars + 1;    // [2]  K and V have separate barriers: a K stage is free as soon as S_j is done
  uint64_t* k_empty = bars + 3;   // [2]  (long before P_j V_j), so K_{j+2} is requested two tiles ahead of its use
  uint64_t* v_full = bars + 5;    // [2]
  uint64_t* v_empty = bars + 7;   // [2]
  uint64_t* s_full = bars + 9;    // [2]
  uint64_t* s_empty = bars + 11;  // [2]
  uint64_t* p_full = bars + 13;
  uint64_t* pv_full = bars + 14;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 15);
  uint32_t* mask_words = tmem_slot + 2;  // [2][4]: key-padding bits of the current tile, double-buffered by tile parity

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_qt = (p.Tq + AT_BM - 1) / AT_BM;
  const int qt = n_qt - 1 - static_cast<int>(blockIdx.x);  // heavy (late) causal tiles first
  const int bh = blockIdx.y, b = bh / p.H, h = bh % p.H;
  const int q0 = qt * AT_BM;
  const int shift = p.Tk - p.Tq;  // key j visible to query i iff j <= i + shift
  int k_end = p.Tk;
  if (p.causal) k_end = min(p.Tk, q0 + AT_BM + shift);
  const int n_tiles = k_end > 0 ? (k_end + AT_BN - 1) / AT_BN : 0;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&k_empty[s], 1);
      mbar_init(&v_full[s], 1);
      mbar_init(&v_empty[s], 1);
      mbar_init(&s_full[s], 1);
      mbar_init(&s_empty[s], 4);
    }
    mbar_init(p_full, 4);
    mbar_init(pv_full, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
  griddep_wait();  // programmatic dependent launch: q, k, v of the previous kernels are read from here on
  griddep_launch_dependents();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tm_S0 = tmem_base, tm_PV = tmem_base + 2 * AT_BN;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0 && n_tiles > 0) {
      mbar_expect_tx(q_full, Cfg::Q_BYTES);
      for (int kb = 0; kb < KB; ++kb) tma_load_4d(sQ + kb * (AT_BM * 128), &tmQ, q_full, kb * 64, q0, h, b);
      for (int j = 0; j < n_tiles; ++j) {
        const int s = j & 1;
        const uint32_t ph = ((j >> 1) & 1) ^ 1;
        uint8_t* sK = sKV + s * 2 * Cfg::KV_BYTES;
        uint8_t* sV = sK + Cfg::KV_BYTES;
        mbar_wait(&k_empty[s], ph);
        mbar_expect_tx(&k_full[s], Cfg::KV_BYTES);
        for (int kb = 0; kb < KB; ++kb) tma_load_4d(sK + kb * (AT_BN * 128), &tmK, &k_full[s], kb * 64, j * AT_BN, h, b);
        mbar_wait(&v_empty[s], ph);
        mbar_expect_tx(&v_full[s], Cfg::KV_BYTES);
        for (int kb = 0; kb < KB; ++kb) tma_load_4d(sV + kb * (AT_BN * 128), &tmV, &v_full[s], kb * 64, j * AT_BN, h, b);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0 && n_tiles > 0) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(AT_BM, AT_BN);
      constexpr uint32_t idesc_pv = umma_idesc_bf16_bmn(AT_BM, D);
      const uint32_t q_addr = smem_u32(sQ), p_addr = smem_u32(sP);
      auto issue_s = [&](int j) {
        const int s = j & 1;
        mbar_wait(&k_full[s], (j >> 1) & 1);
        mbar_wait(&s_empty[s], ((j >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t k_addr = smem_u32(sKV + s * 2 * Cfg::KV_BYTES);
#pragma unroll
        for (int kb = 0; kb < KB; ++kb)
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16_ss(tm_S0 + s * AT_BN, umma_desc_k_sw128(q_addr + kb * (AT_BM * 128) + k * 32),
                         umma_desc_k_sw128(k_addr + kb * (AT_BN * 128) + k * 32), idesc_s, (kb | k) != 0 ? 1u : 0u);
        umma_commit(&s_full[s]);
        umma_commit(&k_empty[s]);  // K_j is not needed again
      };
      mbar_wait(q_full, 0);
      issue_s(0);
      for (int j = 0; j < n_tiles; ++j) {
        if (j + 1 < n_tiles) issue_s(j + 1);
        const int s = j & 1;
        mbar_wait(p_full, j & 1);  // P_j written (and O rescaled if the row maximum jumped)
        mbar_wait(&v_full[s], (j >> 1) & 1);
        tc_fence_after();
        const uint32_t v_addr = smem_u32(sKV + s * 2 * Cfg::KV_BYTES + Cfg::KV_BYTES);
#pragma unroll
        for (int kk = 0; kk < AT_BN / 16; ++kk)
          umma_bf16_ss(tm_PV, umma_desc_k_sw128(p_addr + s * Cfg::P_BYTES + (kk >> 2) * (AT_BM * 128) + (kk & 3) * 32),
                       umma_desc_mn_sw128(v_addr + kk * 2048, AT_BN * 128), idesc_pv, (j | kk) != 0 ? 1u : 0u);
        umma_commit(pv_full);       // O += P_j V_j complete: O may be rescaled, the P tile overwritten
        umma_commit(&v_empty[s]);   // V_j stage reusable
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax warps (2..5)
    const int quad = warp & 3;
    const int row = quad * 32 + lane;  // query row inside the tile = TMEM lane
    const int qi = q0 + row;
    const bool row_ok = qi < p.Tq;
    const uint32_t t_lane = static_cast<uint32_t>(quad * 32) << 16;
    const int wg_tid = (warp - 2) * 32 + lane;  // 0..127, linear id inside the softmax group
    // m: the reference maximum the exponentials are taken against. It follows the running row maximum LAZILY: only
    // when a tile's maximum exceeds it by more than 2^8 is the output (in TMEM) rescaled -- with P in bf16 (relative
    // precision) and fp32 sums, values up to 2^8 above "1" lose nothing, and after the first tile a rescale is rare.
    float m = -INFINITY, l = 0.0f;
    const float sl2 = p.scale_log2;
    const int kmax = p.causal ? min(p.Tk - 1, qi + shift) : p.Tk - 1;  // last key this row may see
    const unsigned char* mrow = p.kv_mask != nullptr ? p.kv_mask + static_cast<long long>(b) * p.kv_mask_stride : nullptr;
    for (int j = 0; j < n_tiles; ++j) {
      const int s = j & 1;
      const int c0 = j * AT_BN;
      // every row of the tile sees every key of the tile: no per-element tests (all but the diagonal / last tiles)
      const bool full = mrow == nullptr && c0 + AT_BN <= p.Tk && (!p.causal || c0 + AT_BN - 1 <= q0 + shift);
      uint32_t mw[4] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu};
      if (mrow != nullptr) {
        // key-padding bits of this tile: one ballot per warp, shared through shared memory (named barrier of the group)
        const int key = c0 + wg_tid;
        const bool on = key < p.Tk && mrow[key] != 0;
        const unsigned int bal = __ballot_sync(0xffffffffu, on);
        if (lane == 0) mask_words[s * 4 + (warp - 2)] = bal;
        asm volatile("bar.sync 2, 128;" ::: "memory");
#pragma unroll
        for (int w = 0; w < 4; ++w) mw[w] = mask_words[s * 4 + w];
      }
      mbar_wait(&s_full[s], (j >> 1) & 1);
      tc_fence_after();
      const uint32_t t_s = tm_S0 + s * AT_BN + t_lane;
      // the whole score row in registers (one TMEM read), raw (unscaled) values
      uint32_t r[AT_BN];
#pragma unroll
      for (int c = 0; c < AT_BN / 32; ++c) tmem_ld_32x32(t_s + c * 32, *reinterpret_cast<uint32_t(*)[32]>(&r[c * 32]));
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_empty[s]);  // S_j is in registers: the tensor core may start S_{j+2}
      if (!full) {
#pragma unroll
        for (int c = 0; c < AT_BN / 32; ++c) {
          const int lim = kmax - (c0 + c * 32);  // columns 0..lim of this chunk are visible
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (!(i <= lim && ((mw[c] >> i) & 1u))) r[c * 32 + i] = 0xff800000u;  // -inf
        }
      }
      float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;  // four independent chains
#pragma unroll
      for (int i = 0; i < AT_BN; i += 4) {
        mx0 = fmaxf(mx0, __uint_as_float(r[i]));
        mx1 = fmaxf(mx1, __uint_as_float(r[i + 1]));
        mx2 = fmaxf(mx2, __uint_as_float(r[i + 2]));
        mx3 = fmaxf(mx3, __uint_as_float(r[i + 3]));
      }
      const float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * sl2;  // scale > 0: max commutes with it
      const bool bump = mx > m + 8.0f;  // (m = -inf on the first tile: any finite maximum bumps)
      const float m_new = bump ? mx : m;
      const float corr = bump ? ex2_approx(m - m_new) : 1.0f;  // 0 when m was -inf
      const float msafe = (m_new == -INFINITY) ? 0.0f : m_new;
      l *= corr;
      // P_j = 2^(x * scale - m_new) as bf16 into the swizzled K-major tile (buffer j & 1: P_{j-2} V_{j-2} is complete,
      // this thread waited for it a tile ago), row sum in fp32 (4 partial sums)
      float sum0 = 0.0f, sum1 = 0.0f, sum2 = 0.0f, sum3 = 0.0f;
      uint8_t* sPj = sP + s * Cfg::P_BYTES;
#pragma unroll
      for (int c = 0; c < AT_BN / 32; ++c) {
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          const float p0 = ex2_approx(fmaf(__uint_as_float(r[c * 32 + i]), sl2, -msafe));
          const float p1 = ex2_approx(fmaf(__uint_as_float(r[c * 32 + i + 1]), sl2, -msafe));
          const float p2 = ex2_approx(fmaf(__uint_as_float(r[c * 32 + i + 2]), sl2, -msafe));
          const float p3 = ex2_approx(fmaf(__uint_as_float(r[c * 32 + i + 3]), sl2, -msafe));
          sum0 += p0;
          sum1 += p1;
          sum2 += p2;
          sum3 += p3;
          pk[i >> 1] = pack_bf16(p0, p1);
          pk[(i >> 1) + 1] = pack_bf16(p2, p3);
        }
        // 32 keys = 64 bytes = four 16-byte chunks of this row's 128-byte line in key block c / 2
        uint8_t* line = sPj + (c >> 1) * (AT_BM * 128) + row * 128;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int chunk = ((c & 1) * 4 + q) ^ (row & 7);
          *reinterpret_cast<uint4*>(line + chunk * 16) = make_uint4(pk[q * 4], pk[q * 4 + 1], pk[q * 4 + 2], pk[q * 4 + 3]);
        }
      }
      if (j > 0) {
        // O += P_{j-1} V_{j-1} has been running behind this tile's softmax. It must be complete before O is rescaled
        // (rare) and before P_j V_j is issued; waiting here also keeps this thread's view of the barrier one phase behind
        mbar_wait(pv_full, (j - 1) & 1);
        tc_fence_after();
        if (__any_sync(0xffffffffu, bump)) {
#pragma unroll
          for (int c = 0; c < D / 32; ++c) {
            uint32_t t[32];
            tmem_ld_32x32(tm_PV + t_lane + c * 32, t);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) t[i] = __float_as_uint(__uint_as_float(t[i]) * corr);
            tmem_st_32x32(tm_PV + t_lane + c * 32, t);
          }
          tmem_st_wait();
        }
      }
      l += (sum0 + sum1) + (sum2 + sum3);
      m = m_new;
      tc_fence_before();
      fence_proxy_async();  // generic-proxy stores of P -> visible to the tensor core's async proxy
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
    }
    if (n_tiles > 0) {
      mbar_wait(pv_full, (n_tiles - 1) & 1);
      tc_fence_after();
    }
    const float inv = l > 0.0f ? 1.0f / l : 0.0f;
    __nv_bfloat16* op = p.o + static_cast<long long>(b) * p.o_sb + static_cast<long long>(qi) * p.o_st +
                        static_cast<long long>(h) * p.o_sh;
#pragma unroll
    for (int c = 0; c < D / 32; ++c) {
      uint32_t t[32];
      if (n_tiles > 0) {
        tmem_ld_32x32(tm_PV + t_lane + c * 32, t);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) t[i] = 0u;
      }
      if (row_ok) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint4 v;
          v.x = pack_bf16(__uint_as_float(t[q * 8 + 0]) * inv, __uint_as_float(t[q * 8 + 1]) * inv);
          v.y = pack_bf16(__uint_as_float(t[q * 8 + 2]) * inv, __uint_as_float(t[q * 8 + 3]) * inv);
          v.z = pack_bf16(__uint_as_float(t[q * 8 + 4]) * inv, __uint_as_float(t[q * 8 + 5]) * inv);
          v.w = pack_bf16(__uint_as_float(t[q * 8 + 6]) * inv, __uint_as_float(t[q * 8 + 7]) * inv);
          *reinterpret_cast<uint4*>(op + c * 32 + q * 8) = v;
        }
      }
    }
    if (row_ok && p.lse != nullptr)
      p.lse[static_cast<long long>(bh) * p.Tq + qi] = l > 0.0f ? m + log2f(l) : INFINITY;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
}

// [B, T, H, d] view of a strided tensor (element strides sb, st, sh; d contiguous) as a 4-D tensor map {d, T, H, B}
static int make_tmap4(CUtensorMap* out, const void* ptr, int d, long long T, int H, int B, long long sb, long long st,
                      long long sh, int box_rows) {
  const unsigned long long dims[4] = {static_cast<unsigned long long>(d), static_cast<unsigned long long>(T),
                                      static_cast<unsigned long long>(H), static_cast<unsigned long long>(B)};
  const unsigned long long strides[3] = {static_cast<unsigned long long>(st) * 2, static_cast<unsigned long long>(sh) * 2,
                                         static_cast<unsigned long long>(sb) * 2};
  const unsigned box[4] = {64u, static_cast<unsigned>(box_rows), 1u, 1u};
  return encode_tmap_bf16(out, ptr, 4, dims, strides, box);
}

template <int D>
static int launch_attn_tc(const mpl_attn_args& a, cudaStream_t stream) {
  using Cfg = AttnTcCfg<D>;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(attn_fwd_tcgen05_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES) !=
        cudaSuccess)
      return MPL_ERR_CUDA;
    attr_set = true;
  }
  CUtensorMap tmQ, tmK, tmV;
  int rc = make_tmap4(&tmQ, a.q, D, a.Tq, a.H, a.B, a.q_stride[0], a.q_stride[1], a.q_stride[2], AT_BM);
  if (rc == MPL_OK) rc = make_tmap4(&tmK, a.k, D, a.Tk, a.H, a.B, a.k_stride[0], a.k_stride[1], a.k_stride[2], AT_BN);
  if (rc == MPL_OK) rc = make_tmap4(&tmV, a.v, D, a.Tk, a.H, a.B, a.v_stride[0], a.v_stride[1], a.v_stride[2], AT_BN);
  if (rc != MPL_OK) return rc;
  AttnTcParams p;
  p.o = static_cast<__nv_bfloat16*>(a.o);
  p.o_sb = a.o_stride[0];
  p.o_st = a.o_stride[1];
  p.o_sh = a.o_stride[2];
  p.lse = a.lse;
  p.kv_mask = a.kv_mask;
  p.kv_mask_stride = a.kv_mask_stride > 0 ? a.kv_mask_stride : a.Tk;
  p.B = a.B;
  p.H = a.H;
  p.Tq = a.Tq;
  p.Tk = a.Tk;
  p.causal = a.causal;
  p.scale_log2 = a.scale * 1.4426950408889634f;
  dim3 grid((a.Tq + AT_BM - 1) / AT_BM, a.B * a.H);
  launch_pdl(attn_fwd_tcgen05_kernel<D>, grid, dim3(AT_THREADS), Cfg::SMEM_BYTES, stream, tmQ, tmK, tmV, p);
  return launch_status();
}

// ------------------------------------------------------------------------------------------------ backward (d = 128)
// FlashAttention backward on the tensor cores (replaces torch autograd through HF-4.31 LlamaAttention in the train step,
// SURVEY.md §8 a-17; round 1 ran it on mma.sync at 140 TFLOP/s). One CTA owns a tile of 128 keys of one (batch, head) and
// walks the query tiles that can see it; five GEMMs per (query tile, key tile), all 128 x 128 x 128 on tcgen05.mma with
// fp32 accumulators in TMEM (512 columns: S then dQ | dP | dV | dK):
//   S  = Q K^T        A = Q  [q][d]   K-major    B = K  [key][d] K-major
//   dP = dO V^T       A = dO [q][d]   K-major    B = V  [key][d] K-major
//   dV += P^T dO      A = P  [q][key] MN-major (M = key)         B = dO [q][d]  MN-major (N = d)
//   dK += dS^T Q      A = dS [q][key] MN-major                   B = Q  [q][d]  MN-major
//   dQ  = dS K        A = dS [q][key] K-major                    B = K  [key][d] MN-major
// i.e. every operand is read from the tile exactly as TMA (Q, K, V, dO) or the softmax warps (P, dS) wrote it: no
// transposes. Warp 0 = TMA producer, warp 1 = MMA issuer (one lane), warps 2..5 = one thread per query row: P = 2^(S c
// - lse), dS = P (dP - delta) scale as bf16 into the swizzled tiles, then the dQ tile is reduced into the fp32 dQ with
// 16-byte vector REDs (query tiles are shared by several key-tile CTAs). dK / dV leave TMEM once, at the end.
struct AttnBwdTcParams {
  const float *lse, *delta;  // [B*H, T]
  float* dq;                 // f32 [B, T, H, 128] contiguous, zero-initialised
  __nv_bfloat16 *dk, *dv;
  long long dk_sb, dk_st, dk_sh, dv_sb, dv_st, dv_sh;
  const unsigned char* kv_mask;
  long long kv_mask_stride;
  int B, H, T, causal;
  float scale, scale_log2;
};
constexpr int ABT_TILE = 128 * 128 * 2;  // one 128 x 128 bf16 tile: [2 blocks of 64 columns][128 rows][128 B]
constexpr int ABT_SMEM = 6 * ABT_TILE + 1024 + 256;
__host__ __device__ constexpr uint32_t umma_idesc_bf16_amn_bmn(uint32_t m, uint32_t n) {
  return umma_idesc_bf16(m, n) | (1u << 15) | (1u << 16);  // a_major = MN, b_major = MN
}

__global__ void __launch_bounds__(AT_THREADS, 1)
attn_bwd_tcgen05_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                        const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmdO,
                        const AttnBwdTcParams p) {
  constexpr int D = 128;
  extern __shared__ uint8_t at_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(at_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sK = smem;
  uint8_t* sV = sK + ABT_TILE;
  uint8_t* sQ = sV + ABT_TILE;
  uint8_t* sdO = sQ + ABT_TILE;
  uint8_t* sP = sdO + ABT_TILE;
  uint8_t* sdS = sP + ABT_TILE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sdS + ABT_TILE);
  uint64_t* kv_full = bars;
  uint64_t* q_full = bars + 1;
  uint64_t* q_empty = bars + 2;
  uint64_t* s_full = bars + 3;
  uint64_t* dp_full = bars + 4;
  uint64_t* pds_full = bars + 5;
  uint64_t* dq_full = bars + 6;
  uint64_t* dq_empty = bars + 7;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  uint32_t* mask_words = tmem_slot + 2;  // [4]: key-padding bits of this CTA's 128 keys

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j = blockIdx.x;  // key tile (heavy causal tiles have the small j: launched first)
  const int bh = blockIdx.y, b = bh / p.H, h = bh % p.H;
  const int k0 = j * AT_BN;
  const int n_qt = (p.T + AT_BM - 1) / AT_BM;
  const int i_begin = p.causal ? j : 0;  // first query tile that sees these keys
  const int n_it = n_qt - i_begin;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    tma_prefetch_desc(&tmdO);
    mbar_init(kv_full, 1);
    mbar_init(q_full, 1);
    mbar_init(q_empty, 1);
    mbar_init(s_full, 1);
    mbar_init(dp_full, 1);
    mbar_init(pds_full, 4);
    mbar_init(dq_full, 1);
    mbar_init(dq_empty, 4);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  griddep_wait();
  griddep_launch_dependents();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tm_S = tmem_base, tm_dP = tmem_base + 128, tm_dV = tmem_base + 256, tm_dK = tmem_base + 384;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0 && n_it > 0) {
      mbar_expect_tx(kv_full, 2 * ABT_TILE);
      for (int kb = 0; kb < 2; ++kb) {
        tma_load_4d(sK + kb * (128 * 128), &tmK, kv_full, kb * 64, k0, h, b);
        tma_load_4d(sV + kb * (128 * 128), &tmV, kv_full, kb * 64, k0, h, b);
      }
      for (int it = 0; it < n_it; ++it) {
        const int q0 = (i_begin + it) * AT_BM;
        if (it > 0) mbar_wait(q_empty, (it - 1) & 1);  // the MMAs that read Q / dO / P / dS of the previous tile retired
        mbar_expect_tx(q_full, 2 * ABT_TILE);
        for (int kb = 0; kb < 2; ++kb) {
          tma_load_4d(sQ + kb * (128 * 128), &tmQ, q_full, kb * 64, q0, h, b);
          tma_load_4d(sdO + kb * (128 * 128), &tmdO, q_full, kb * 64, q0, h, b);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0 && n_it > 0) {
      constexpr uint32_t id_kk = umma_idesc_bf16(128, 128);           // A K-major, B K-major
      constexpr uint32_t id_mm = umma_idesc_bf16_amn_bmn(128, 128);   // A MN-major, B MN-major
      constexpr uint32_t id_km = umma_idesc_bf16_bmn(128, 128);       // A K-major, B MN-major
      const uint32_t aK = smem_u32(sK), aV = smem_u32(sV), aQ = smem_u32(sQ), aO = smem_u32(sdO), aP = smem_u32(sP),
                     aS = smem_u32(sdS);
      mbar_wait(kv_full, 0);
      for (int it = 0; it < n_it; ++it) {
        mbar_wait(q_full, it & 1);
        if (it > 0) mbar_wait(dq_empty, (it - 1) & 1);  // dQ of the previous tile has left the S columns
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const uint32_t off = (kk >> 2) * (128 * 128) + (kk & 3) * 32;
          umma_bf16_ss(tm_S, umma_desc_k_sw128(aQ + off), umma_desc_k_sw128(aK + off), id_kk, kk != 0 ? 1u : 0u);
        }
        umma_commit(s_full);
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const uint32_t off = (kk >> 2) * (128 * 128) + (kk & 3) * 32;
          umma_bf16_ss(tm_dP, umma_desc_k_sw128(aO + off), umma_desc_k_sw128(aV + off), id_kk, kk != 0 ? 1u : 0u);
        }
        umma_commit(dp_full);
        mbar_wait(pds_full, it & 1);  // P and dS of this tile are in shared memory
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)  // contraction over the 128 queries, 16 per instruction
          umma_bf16_ss(tm_dV, umma_desc_mn_sw128(aP + kk * 2048, 128 * 128), umma_desc_mn_sw128(aO + kk * 2048, 128 * 128),
                       id_mm, (it | kk) != 0 ? 1u : 0u);
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)
          umma_bf16_ss(tm_dK, umma_desc_mn_sw128(aS + kk * 2048, 128 * 128), umma_desc_mn_sw128(aQ + kk * 2048, 128 * 128),
                       id_mm, (it | kk) != 0 ? 1u : 0u);
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)  // contraction over the 128 keys
          umma_bf16_ss(tm_S, umma_desc_k_sw128(aS + (kk >> 2) * (128 * 128) + (kk & 3) * 32),
                       umma_desc_mn_sw128(aK + kk * 2048, 128 * 128), id_km, kk != 0 ? 1u : 0u);
        umma_commit(q_empty);  // Q, dO, P, dS may be overwritten
        umma_commit(dq_full);  // dQ tile (and, after the last tile, dK / dV) complete
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax / reduction warps (2..5)
    const int quad = warp & 3;
    const int row = quad * 32 + lane;  // TMEM lane: query row of the tile (S, dP, dQ) / key row (dV, dK)
    const uint32_t t_lane = static_cast<uint32_t>(quad * 32) << 16;
    const int wg_tid = (warp - 2) * 32 + lane;
    uint32_t mw[4] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu};
    {
      const int key = k0 + wg_tid;
      bool on = key < p.T;
      if (on && p.kv_mask != nullptr) on = p.kv_mask[static_cast<long long>(b) * p.kv_mask_stride + key] != 0;
      const unsigned int bal = __ballot_sync(0xffffffffu, on);
      if (lane == 0) mask_words[warp - 2] = bal;
      asm volatile("bar.sync 2, 128;" ::: "memory");
#pragma unroll
      for (int w = 0; w < 4; ++w) mw[w] = mask_words[w];
    }
    const bool keys_full = (mw[0] & mw[1] & mw[2] & mw[3]) == 0xffffffffu;
    const float sl2 = p.scale_log2, sc = p.scale;
    for (int it = 0; it < n_it; ++it) {
      const int q0 = (i_begin + it) * AT_BM;
      const int q = q0 + row;
      const bool row_ok = q < p.T;
      const float lse_q = row_ok ? p.lse[static_cast<long long>(bh) * p.T + q] : INFINITY;
      const float delta_q = row_ok ? p.delta[static_cast<long long>(bh) * p.T + q] : 0.0f;
      const bool full = keys_full && (!p.causal || q0 >= k0 + AT_BN - 1);  // no per-element visibility tests needed
      const int kmax = p.causal ? q - k0 : AT_BN;  // columns 0..kmax of this row are visible (causal)
      mbar_wait(s_full, it & 1);
      mbar_wait(dp_full, it & 1);
      tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t rs[32], rd[32];
        tmem_ld_32x32(tm_S + t_lane + c * 32, rs);
        tmem_ld_32x32(tm_dP + t_lane + c * 32, rd);
        tmem_ld_wait();
        uint32_t pk[16], dk_[16];
        const uint32_t mwc = c == 0 ? mw[0] : (c == 1 ? mw[1] : (c == 2 ? mw[2] : mw[3]));
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          float p0 = ex2_approx(fmaf(__uint_as_float(rs[i]), sl2, -lse_q));
          float p1 = ex2_approx(fmaf(__uint_as_float(rs[i + 1]), sl2, -lse_q));
          if (!full) {
            if (!(c * 32 + i <= kmax && ((mwc >> i) & 1u))) p0 = 0.0f;
            if (!(c * 32 + i + 1 <= kmax && ((mwc >> (i + 1)) & 1u))) p1 = 0.0f;
          }
          const float d0 = p0 * (__uint_as_float(rd[i]) - delta_q) * sc;
          const float d1 = p1 * (__uint_as_float(rd[i + 1]) - delta_q) * sc;
          pk[i >> 1] = pack_bf16(p0, p1);
          dk_[i >> 1] = pack_bf16(d0, d1);
        }
        // 32 keys = 64 bytes = four 16-byte chunks of this row's 128-byte line in key block c / 2 (128-B swizzle)
        uint8_t* lineP = sP + (c >> 1) * (128 * 128) + row * 128;
        uint8_t* lineS = sdS + (c >> 1) * (128 * 128) + row * 128;
#pragma unroll
        for (int qq = 0; qq < 4; ++qq) {
          const int chunk = ((c & 1) * 4 + qq) ^ (row & 7);
          *reinterpret_cast<uint4*>(lineP + chunk * 16) = make_uint4(pk[qq * 4], pk[qq * 4 + 1], pk[qq * 4 + 2], pk[qq * 4 + 3]);
          *reinterpret_cast<uint4*>(lineS + chunk * 16) =
              make_uint4(dk_[qq * 4], dk_[qq * 4 + 1], dk_[qq * 4 + 2], dk_[qq * 4 + 3]);
        }
      }
      tc_fence_before();
      fence_proxy_async();  // generic-proxy stores of P / dS -> visible to the tensor core's async proxy
      __syncwarp();
      if (lane == 0) mbar_arrive(pds_full);
      // dQ tile of this (query tile, key tile) pair -> fp32 dQ in global memory
      mbar_wait(dq_full, it & 1);
      tc_fence_after();
      float* dq_row = p.dq + ((static_cast<long long>(b) * p.T + q) * p.H + h) * D;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t t[32];
        tmem_ld_32x32(tm_S + t_lane + c * 32, t);
        tmem_ld_wait();
        if (row_ok) {
#pragma unroll
          for (int qq = 0; qq < 8; ++qq)
            atomicAdd(reinterpret_cast<float4*>(dq_row + c * 32 + qq * 4),
                      make_float4(__uint_as_float(t[qq * 4]), __uint_as_float(t[qq * 4 + 1]), __uint_as_float(t[qq * 4 + 2]),
                                  __uint_as_float(t[qq * 4 + 3])));
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(dq_empty);
    }
    // dK / dV of this key tile (complete with the last dq_full): thread = key row. (tcgen05.ld is .sync.aligned: every
    // lane of the warp takes part, only the stores are guarded)
    const int key = k0 + row;
    if (n_it > 0) {
      __nv_bfloat16* dvp = p.dv + static_cast<long long>(b) * p.dv_sb + static_cast<long long>(key) * p.dv_st +
                           static_cast<long long>(h) * p.dv_sh;
      __nv_bfloat16* dkp = p.dk + static_cast<long long>(b) * p.dk_sb + static_cast<long long>(key) * p.dk_st +
                           static_cast<long long>(h) * p.dk_sh;
#pragma unroll 1
      for (int c = 0; c < 8; ++c) {
        uint32_t t[32];
        tmem_ld_32x32((c < 4 ? tm_dV : tm_dK) + t_lane + (c & 3) * 32, t);
        tmem_ld_wait();
        if (key < p.T) {
          __nv_bfloat16* dst = (c < 4 ? dvp : dkp) + (c & 3) * 32;
#pragma unroll
          for (int qq = 0; qq < 4; ++qq) {
            uint4 v;
            v.x = pack_bf16(__uint_as_float(t[qq * 8 + 0]), __uint_as_float(t[qq * 8 + 1]));
            v.y = pack_bf16(__uint_as_float(t[qq * 8 + 2]), __uint_as_float(t[qq * 8 + 3]));
            v.z = pack_bf16(__uint_as_float(t[qq * 8 + 4]), __uint_as_float(t[qq * 8 + 5]));
            v.w = pack_bf16(__uint_as_float(t[qq * 8 + 6]), __uint_as_float(t[qq * 8 + 7]));
            *reinterpret_cast<uint4*>(dst + qq * 8) = v;
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem_base);
}

bool attention_bwd_tc_supported(const mpl_attn_bwd_args& a) {
  static int enabled = -1;
  if (enabled < 0) {
    const char* e = getenv("MPL_ATTN_BWD_TC");
    enabled = (e == nullptr || atoi(e) != 0) ? 1 : 0;
  }
  if (!enabled || a.head_dim != 128 || a.T < 1) return false;
  const long long strides[] = {a.q_stride[0], a.q_stride[1], a.q_stride[2], a.k_stride[0], a.k_stride[1], a.k_stride[2],
                               a.v_stride[0], a.v_stride[1], a.v_stride[2], a.o_stride[0], a.o_stride[1], a.o_stride[2],
                               a.dk_stride[0], a.dk_stride[1], a.dk_stride[2], a.dv_stride[0], a.dv_stride[1],
                               a.dv_stride[2]};
  for (long long s : strides)
    if (s % 8 != 0 || s <= 0) return false;
  const uintptr_t ptrs[] = {reinterpret_cast<uintptr_t>(a.q),  reinterpret_cast<uintptr_t>(a.k),
                            reinterpret_cast<uintptr_t>(a.v),  reinterpret_cast<uintptr_t>(a.d_o),
                            reinterpret_cast<uintptr_t>(a.dk), reinterpret_cast<uintptr_t>(a.dv),
                            reinterpret_cast<uintptr_t>(a.dq_f32)};
  for (uintptr_t q : ptrs)
    if (q % 16 != 0) return false;
  return true;
}

// a.delta must already hold rowsum(dO * O) (attn_delta_kernel, train.cu); dq_f32 zero-initialised by the caller
int attention_bwd_tc(const mpl_attn_bwd_args& a, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(attn_bwd_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ABT_SMEM) != cudaSuccess)
      return MPL_ERR_CUDA;
    attr_set = true;
  }
  CUtensorMap tmQ, tmK, tmV, tmdO;
  int rc = make_tmap4(&tmQ, a.q, 128, a.T, a.H, a.B, a.q_stride[0], a.q_stride[1], a.q_stride[2], AT_BM);
  if (rc == MPL_OK) rc = make_tmap4(&tmK, a.k, 128, a.T, a.H, a.B, a.k_stride[0], a.k_stride[1], a.k_stride[2], AT_BN);
  if (rc == MPL_OK) rc = make_tmap4(&tmV, a.v, 128, a.T, a.H, a.B, a.v_stride[0], a.v_stride[1], a.v_stride[2], AT_BN);
  if (rc == MPL_OK) rc = make_tmap4(&tmdO, a.d_o, 128, a.T, a.H, a.B, a.o_stride[0], a.o_stride[1], a.o_stride[2], AT_BM);
  if (rc != MPL_OK) return rc;
  AttnBwdTcParams p;
  p.lse = a.lse;
  p.delta = a.delta;
  p.dq = a.dq_f32;
  p.dk = static_cast<__nv_bfloat16*>(a.dk);
  p.dv = static_cast<__nv_bfloat16*>(a.dv);
  p.dk_sb = a.dk_stride[0];
  p.dk_st = a.dk_stride[1];
  p.dk_sh = a.dk_stride[2];
  p.dv_sb = a.dv_stride[0];
  p.dv_st = a.dv_stride[1];
  p.dv_sh = a.dv_stride[2];
  p.kv_mask = a.kv_mask;
  p.kv_mask_stride = a.kv_mask_stride > 0 ? a.kv_mask_stride : a.T;
  p.B = a.B;
  p.H = a.H;
  p.T = a.T;
  p.causal = a.causal;
  p.scale = a.scale;
  p.scale_log2 = a.scale * 1.4426950408889634f;
  dim3 grid((a.T + AT_BN - 1) / AT_BN, a.B * a.H);
  launch_pdl(attn_bwd_tcgen05_kernel, grid, dim3(AT_THREADS), ABT_SMEM, stream, tmQ, tmK, tmV, tmdO, p);
  return launch_status();
}

// true when the tcgen05 kernel can take this problem (the caller falls back to the mma.sync kernels otherwise)
bool attention_tc_supported(const mpl_attn_args& a) {
  static int enabled = -1;
  if (enabled < 0) {
    const char* e = getenv("MPL_ATTN_TC");
    enabled = (e == nullptr || atoi(e) != 0) ? 1 : 0;
  }
  if (!enabled) return false;
  if (a.head_dim != 64 && a.head_dim != 128) return false;
  if (a.Tq < 32 || a.Tk < 1 || a.rel_h != nullptr || a.rel_w != nullptr || a.tk_dev != nullptr) return false;
  const long long strides[] = {a.q_stride[0], a.q_stride[1], a.q_stride[2], a.k_stride[0], a.k_stride[1], a.k_stride[2],
                               a.v_stride[0], a.v_stride[1], a.v_stride[2], a.o_stride[0], a.o_stride[1], a.o_stride[2]};
  for (long long s : strides)
    if (s % 8 != 0 || s <= 0) return false;
  const uintptr_t ptrs[] = {reinterpret_cast<uintptr_t>(a.q), reinterpret_cast<uintptr_t>(a.k),
                            reinterpret_cast<uintptr_t>(a.v), reinterpret_cast<uintptr_t>(a.o)};
  for (uintptr_t q : ptrs)
    if (q % 16 != 0) return false;
  return true;
}

int attention_tc(const mpl_attn_args& a, cudaStream_t stream) {
  return a.head_dim == 128 ? launch_attn_tc<128>(a, stream) : launch_attn_tc<64>(a, stream);
}

}  // namespace mpl
