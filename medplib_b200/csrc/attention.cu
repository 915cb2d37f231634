// K2 / K2b / K2c / K3 — attention.
//   flash_fwd   : tiled online-softmax attention (fp32 softmax, bf16 P·V) on mma.sync m16n8k16 fragments,
//                 64 query rows x 64 keys per step, XOR-swizzled shared memory + ldmatrix. Used for
//                   * LLaMA prefill: causal, head_dim 128, optional key-padding mask (HF-4.31 LlamaAttention,
//                     SURVEY.md App. A.1; call site model/medplib/model/language_model/medplib_moe_llama.py:127-135)
//                   * CLIP ViT-L: non-causal, 577 tokens, head_dim 64 (App. A.2; clip_encoder.py:53-57)
//                   * SAM-Med2D: windows of 196 / global 256 tokens, head_dim 64, decomposed relative-position
//                     bias rel_h[q, k/kw] + rel_w[q, k%kw] added in-kernel
//                     (model/segment_anything_med2d/modeling/image_encoder.py:280-296,381-421)
//                   * SAM mask decoder token<->image attention, head_dim 16/32
//                     (model/segment_anything_med2d/modeling/transformer.py:185-244)
//   decode      : one query token per sequence against the KV cache (HBM-bound): 8 lanes per key, 16-byte loads,
//                 warp-shuffle reductions, 8 warps per (batch, head) combined through shared memory.
// Layout: q/k/v/o are addressed as base + b*stride_b + t*stride_t + h*stride_h + d (elements), so fused-QKV
// buffers, KV caches and per-head views need no copies.
#include "internal.h"
#include "ptx.cuh"

namespace mpl {

struct AttnParams {
  const __nv_bfloat16* q;
  const __nv_bfloat16* k;
  const __nv_bfloat16* v;
  __nv_bfloat16* o;
  long long q_sb, q_st, q_sh;
  long long k_sb, k_st, k_sh;
  long long v_sb, v_st, v_sh;
  long long o_sb, o_st, o_sh;
  int B, H, Tq, Tk;
  float scale;
  int causal;                    // key j visible to query i iff j <= i + (Tk - Tq)
  const unsigned char* kv_mask;  // [B, Tk] 1 = attend, or NULL
  long long kv_mask_stride;
  const float* rel_h;            // [B*H, Tq, rel_kh] or NULL
  const float* rel_w;            // [B*H, Tq, rel_kw]
  int rel_kh, rel_kw;
  const int* tk_dev;             // optional device-side Tk (decode under CUDA graphs)
  float* scratch;                // decode split-K partials: [B*H] int counters (zeroed) then [B*H, nsplit, D+2] floats
  int nsplit;
  float* lse;                    // prefill only, optional: [B*H, Tq] log2-domain log-sum-exp of the scaled scores (training)
};

__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void mma_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

constexpr int FA_BM = 64, FA_BN = 64, FA_THREADS = 128;

// Byte offset of 16-byte chunk `c` of row `r` inside a [rows][D] bf16 tile with XOR swizzle.
template <int D>
__device__ __forceinline__ uint32_t swz(int r, int c) {
  constexpr int CPR = D / 8;
  constexpr int MASK = CPR < 8 ? CPR - 1 : 7;
  return static_cast<uint32_t>(r * (D * 2) + ((c ^ (r & MASK)) << 4));
}

template <int D>
__device__ __forceinline__ void load_tile(uint32_t smem_base, const __nv_bfloat16* g, long long stride_t, int t0,
                                          int t_end, int rows) {
  constexpr int CPR = D / 8;  // 16-byte chunks per row
  for (int idx = threadIdx.x; idx < rows * CPR; idx += FA_THREADS) {
    const int r = idx / CPR, c = idx % CPR;
    const int t = t0 + r;
    const bool ok = t < t_end;
    const __nv_bfloat16* src = g + static_cast<long long>(ok ? t : t0) * stride_t + c * 8;
    cp_async16(smem_base + swz<D>(r, c), src, ok);
  }
}

template <int D>
__global__ void __launch_bounds__(FA_THREADS) flash_fwd_kernel(const AttnParams p) {
  extern __shared__ __align__(128) uint8_t fa_smem[];
  const uint32_t sQ = smem_u32(fa_smem);
  const uint32_t sK0 = sQ + FA_BM * D * 2;       // K / V tiles are double-buffered: the next key block streams in
  const uint32_t sV0 = sK0 + 2 * FA_BN * D * 2;  // (cp.async) while the current one is being multiplied
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t4 = lane & 3;
  const int b = blockIdx.z, h = blockIdx.y;
  const int q0 = blockIdx.x * FA_BM;
  const int Tk = p.tk_dev ? *p.tk_dev : p.Tk;
  const int off = Tk - p.Tq;
  const __nv_bfloat16* qg = p.q + b * p.q_sb + h * p.q_sh;
  const __nv_bfloat16* kg = p.k + b * p.k_sb + h * p.k_sh;
  const __nv_bfloat16* vg = p.v + b * p.v_sb + h * p.v_sh;

  load_tile<D>(sQ, qg, p.q_st, q0, p.Tq, FA_BM);
  load_tile<D>(sK0, kg, p.k_st, 0, Tk, FA_BN);  // first key block (k_end >= 1 whenever there is a query)
  load_tile<D>(sV0, vg, p.v_st, 0, Tk, FA_BN);
  asm volatile("cp.async.commit_group;" ::: "memory");
  cp_async_wait_all();
  __syncthreads();
  // Q fragments for this warp's 16 rows, all of D
  uint32_t qf[D / 16][4];
#pragma unroll
  for (int kk = 0; kk < D / 16; ++kk) {
    const int r = warp * 16 + (lane & 15);
    const int c = kk * 2 + (lane >> 4);
    ldmatrix_x4(qf[kk], sQ + swz<D>(r, c));
  }
  float o[D / 8][4];
#pragma unroll
  for (int i = 0; i < D / 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) o[i][j] = 0.0f;
  float row_max[2] = {-INFINITY, -INFINITY};
  float row_sum[2] = {0.0f, 0.0f};
  const float sl2 = p.scale * 1.4426950408889634f;
  const int qrow[2] = {q0 + warp * 16 + g, q0 + warp * 16 + g + 8};

  int k_end = Tk;
  if (p.causal) k_end = min(Tk, q0 + FA_BM + off);
  const unsigned char* mrow = p.kv_mask ? p.kv_mask + static_cast<long long>(b) * p.kv_mask_stride : nullptr;
  const float* relh[2] = {nullptr, nullptr};
  const float* relw[2] = {nullptr, nullptr};
  if (p.rel_h) {
    const long long bh = static_cast<long long>(b) * p.H + h;
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      const int qr = min(qrow[rr], p.Tq - 1);
      relh[rr] = p.rel_h + (bh * p.Tq + qr) * p.rel_kh;
      relw[rr] = p.rel_w + (bh * p.Tq + qr) * p.rel_kw;
    }
  }

  int kbuf = 0;
  for (int k0 = 0; k0 < k_end; k0 += FA_BN, kbuf ^= 1) {
    cp_async_wait_all();
    __syncthreads();  // this block's K / V landed for every thread; the previous iteration's reads of the other buffer done
    const uint32_t sK = sK0 + kbuf * FA_BN * D * 2, sV = sV0 + kbuf * FA_BN * D * 2;
    if (k0 + FA_BN < k_end) {
      load_tile<D>(sK0 + (kbuf ^ 1) * FA_BN * D * 2, kg, p.k_st, k0 + FA_BN, Tk, FA_BN);
      load_tile<D>(sV0 + (kbuf ^ 1) * FA_BN * D * 2, vg, p.v_st, k0 + FA_BN, Tk, FA_BN);
      asm volatile("cp.async.commit_group;" ::: "memory");
    }

    float s[FA_BN / 8][4];
#pragma unroll
    for (int i = 0; i < FA_BN / 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) s[i][j] = 0.0f;
#pragma unroll
    for (int kk = 0; kk < D / 16; ++kk) {
#pragma unroll
      for (int np = 0; np < FA_BN / 16; ++np) {
        uint32_t kf[4];
        const int r = np * 16 + (lane & 7) + ((lane >> 4) << 3);
        const int c = kk * 2 + ((lane >> 3) & 1);
        ldmatrix_x4(kf, sK + swz<D>(r, c));
        mma_16816(s[np * 2], qf[kk], kf[0], kf[1]);
        mma_16816(s[np * 2 + 1], qf[kk], kf[2], kf[3]);
      }
    }
    // scale (log2 domain), bias, masks
    float tile_max[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int nt = 0; nt < FA_BN / 8; ++nt) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int rr = j >> 1;
        const int key = k0 + nt * 8 + t4 * 2 + (j & 1);
        float val = s[nt][j] * sl2;
        bool ok = key < Tk;
        if (p.causal) ok = ok && (key <= qrow[rr] + off);
        if (mrow != nullptr && ok) ok = mrow[key] != 0;
        if (relh[0] != nullptr && ok)
          val += (relh[rr][key / p.rel_kw] + relw[rr][key % p.rel_kw]) * 1.4426950408889634f;
        val = ok ? val : -INFINITY;
        s[nt][j] = val;
        tile_max[rr] = fmaxf(tile_max[rr], val);
      }
    }
    float corr[2], mnew[2];
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      float m = tile_max[rr];
      m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
      m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
      mnew[rr] = fmaxf(row_max[rr], m);
      const float msafe = (mnew[rr] == -INFINITY) ? 0.0f : mnew[rr];
      corr[rr] = exp2f(row_max[rr] - msafe);  // row_max = -inf -> 0
      row_max[rr] = mnew[rr];
      mnew[rr] = msafe;
      row_sum[rr] *= corr[rr];
    }
#pragma unroll
    for (int i = 0; i < D / 8; ++i) {
      o[i][0] *= corr[0];
      o[i][1] *= corr[0];
      o[i][2] *= corr[1];
      o[i][3] *= corr[1];
    }
    uint32_t pf[FA_BN / 16][4];
#pragma unroll
    for (int nt = 0; nt < FA_BN / 8; ++nt) {
      const float p0 = exp2f(s[nt][0] - mnew[0]);
      const float p1 = exp2f(s[nt][1] - mnew[0]);
      const float p2 = exp2f(s[nt][2] - mnew[1]);
      const float p3 = exp2f(s[nt][3] - mnew[1]);
      row_sum[0] += p0 + p1;
      row_sum[1] += p2 + p3;
      pf[nt >> 1][(nt & 1) * 2] = pack_bf16(p0, p1);
      pf[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16(p2, p3);
    }
#pragma unroll
    for (int kk = 0; kk < FA_BN / 16; ++kk) {
#pragma unroll
      for (int dp = 0; dp < D / 16; ++dp) {
        uint32_t vf[4];
        const int r = kk * 16 + (lane & 7) + (((lane >> 3) & 1) << 3);
        const int c = dp * 2 + (lane >> 4);
        ldmatrix_x4_trans(vf, sV + swz<D>(r, c));
        mma_16816(o[dp * 2], pf[kk], vf[0], vf[1]);
        mma_16816(o[dp * 2 + 1], pf[kk], vf[2], vf[3]);
      }
    }
  }
  // finalize
#pragma unroll
  for (int rr = 0; rr < 2; ++rr) {
    float sres = row_sum[rr];
    sres += __shfl_xor_sync(0xffffffffu, sres, 1);
    sres += __shfl_xor_sync(0xffffffffu, sres, 2);
    if (p.lse != nullptr && t4 == 0 && qrow[rr] < p.Tq)
      p.lse[(static_cast<long long>(b) * p.H + h) * p.Tq + qrow[rr]] = sres > 0.0f ? row_max[rr] + log2f(sres) : INFINITY;
    row_sum[rr] = sres > 0.0f ? 1.0f / sres : 0.0f;
  }
  __nv_bfloat16* og = p.o + b * p.o_sb + h * p.o_sh;
#pragma unroll
  for (int rr = 0; rr < 2; ++rr) {
    if (qrow[rr] < p.Tq) {
      __nv_bfloat16* orow = og + static_cast<long long>(qrow[rr]) * p.o_st;
#pragma unroll
      for (int i = 0; i < D / 8; ++i) {
        const uint32_t packed = pack_bf16(o[i][rr * 2] * row_sum[rr], o[i][rr * 2 + 1] * row_sum[rr]);
        *reinterpret_cast<uint32_t*>(orow + i * 8 + t4 * 2) = packed;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------ decode (Tq == 1)
constexpr int DEC_WARPS = 8;
template <int D>
__global__ void __launch_bounds__(DEC_WARPS * 32) decode_attn_kernel(const AttnParams p) {
  static_assert(D == 128, "decode kernel is written for head_dim 128");
  __shared__ float s_o[DEC_WARPS][D];
  __shared__ float s_m[DEC_WARPS], s_l[DEC_WARPS];
  const int b = blockIdx.y, h = blockIdx.x;
  griddep_launch_dependents();  // the o_proj streaming GEMM may start prefetching its weights
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int grp = lane >> 3, gl = lane & 7;  // 4 keys per warp step, 8 lanes x 16 dims per key
  const int Tk = p.tk_dev ? *p.tk_dev : p.Tk;
  // split-K over the keys (flash-decoding): blockIdx.z owns keys [k_lo, k_hi), a multiple of 32 keys per split
  const int nsplit = gridDim.z;
  const int per = ((Tk + nsplit - 1) / nsplit + 31) & ~31;
  const int k_lo = blockIdx.z * per;
  const int k_hi = min(Tk, k_lo + per);
  const __nv_bfloat16* qg = p.q + b * p.q_sb + h * p.q_sh + gl * 16;
  const __nv_bfloat16* kg = p.k + b * p.k_sb + h * p.k_sh + gl * 16;
  const __nv_bfloat16* vg = p.v + b * p.v_sb + h * p.v_sh + gl * 16;
  const unsigned char* mrow = p.kv_mask ? p.kv_mask + static_cast<long long>(b) * p.kv_mask_stride : nullptr;
  float qf[16];
  {
    const uint4 a = *reinterpret_cast<const uint4*>(qg);
    const uint4 c = *reinterpret_cast<const uint4*>(qg + 8);
    const __nv_bfloat162* ha = reinterpret_cast<const __nv_bfloat162*>(&a);
    const __nv_bfloat162* hc = reinterpret_cast<const __nv_bfloat162*>(&c);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 fa = __bfloat1622float2(ha[e]), fc = __bfloat1622float2(hc[e]);
      qf[2 * e] = fa.x;
      qf[2 * e + 1] = fa.y;
      qf[8 + 2 * e] = fc.x;
      qf[8 + 2 * e + 1] = fc.y;
    }
  }
  const float sl2 = p.scale * 1.4426950408889634f;
  float m = -INFINITY, l = 0.0f;
  float acc[16];
#pragma unroll
  for (int e = 0; e < 16; ++e) acc[e] = 0.0f;
  // The trip count is warp-uniform (full-mask shuffles inside): groups past the end compute a masked dummy key.
  for (int k0 = k_lo + warp * 4; k0 < k_hi; k0 += DEC_WARPS * 4) {
    const int key = k0 + grp;
    const bool valid = key < k_hi;
    const int kk = valid ? key : Tk - 1;
    const __nv_bfloat16* kr = kg + static_cast<long long>(kk) * p.k_st;
    const __nv_bfloat16* vr = vg + static_cast<long long>(kk) * p.v_st;
    const uint4 ka = *reinterpret_cast<const uint4*>(kr), kc = *reinterpret_cast<const uint4*>(kr + 8);
    const uint4 va = *reinterpret_cast<const uint4*>(vr), vc = *reinterpret_cast<const uint4*>(vr + 8);
    const __nv_bfloat162* hka = reinterpret_cast<const __nv_bfloat162*>(&ka);
    const __nv_bfloat162* hkc = reinterpret_cast<const __nv_bfloat162*>(&kc);
    float dot = 0.0f;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 fa = __bfloat1622float2(hka[e]), fc = __bfloat1622float2(hkc[e]);
      dot += qf[2 * e] * fa.x + qf[2 * e + 1] * fa.y + qf[8 + 2 * e] * fc.x + qf[8 + 2 * e + 1] * fc.y;
    }
    dot += __shfl_xor_sync(0xffffffffu, dot, 1);
    dot += __shfl_xor_sync(0xffffffffu, dot, 2);
    dot += __shfl_xor_sync(0xffffffffu, dot, 4);
    float sc = dot * sl2;
    if (!valid || (mrow != nullptr && mrow[kk] == 0)) sc = -INFINITY;
    const float mn = fmaxf(m, sc);
    const float msafe = (mn == -INFINITY) ? 0.0f : mn;
    const float corr = exp2f(m - msafe);
    // P is rounded to bf16 before multiplying V, as the reference's softmax(...).to(bf16) @ v does
    const float pexp = exp2f(sc - msafe);
    const float pv = bf16_round(pexp);
    l = l * corr + pexp;
    m = mn;
    const __nv_bfloat162* hva = reinterpret_cast<const __nv_bfloat162*>(&va);
    const __nv_bfloat162* hvc = reinterpret_cast<const __nv_bfloat162*>(&vc);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 fa = __bfloat1622float2(hva[e]), fc = __bfloat1622float2(hvc[e]);
      acc[2 * e] = acc[2 * e] * corr + pv * fa.x;
      acc[2 * e + 1] = acc[2 * e + 1] * corr + pv * fa.y;
      acc[8 + 2 * e] = acc[8 + 2 * e] * corr + pv * fc.x;
      acc[8 + 2 * e + 1] = acc[8 + 2 * e + 1] * corr + pv * fc.y;
    }
  }
  // merge the 4 key groups of the warp (lanes differing in bits 3,4)
#pragma unroll
  for (int sh = 8; sh <= 16; sh <<= 1) {
    const float m2 = __shfl_xor_sync(0xffffffffu, m, sh);
    const float l2 = __shfl_xor_sync(0xffffffffu, l, sh);
    const float mn = fmaxf(m, m2);
    const float msafe = (mn == -INFINITY) ? 0.0f : mn;
    const float c1 = exp2f(m - msafe), c2 = exp2f(m2 - msafe);
    l = l * c1 + l2 * c2;
#pragma unroll
    for (int e = 0; e < 16; ++e) {
      const float a2 = __shfl_xor_sync(0xffffffffu, acc[e], sh);
      acc[e] = acc[e] * c1 + a2 * c2;
    }
    m = mn;
  }
  if (grp == 0) {
#pragma unroll
    for (int e = 0; e < 16; ++e) s_o[warp][gl * 16 + e] = acc[e];
    if (gl == 0) {
      s_m[warp] = m;
      s_l[warp] = l;
    }
  }
  __syncthreads();
  __shared__ int s_last;
  float mm = -INFINITY, lt = 0.0f, ot = 0.0f;
  if (threadIdx.x < D) {
#pragma unroll
    for (int w = 0; w < DEC_WARPS; ++w) mm = fmaxf(mm, s_m[w]);
    const float msafe = (mm == -INFINITY) ? 0.0f : mm;
#pragma unroll
    for (int w = 0; w < DEC_WARPS; ++w) {
      const float c = exp2f(s_m[w] - msafe);
      lt += s_l[w] * c;
      ot += s_o[w][threadIdx.x] * c;
    }
  }
  __nv_bfloat16* optr = p.o + b * p.o_sb + h * p.o_sh;
  if (nsplit == 1) {
    if (threadIdx.x < D) optr[threadIdx.x] = __float2bfloat16_rn(lt > 0.0f ? ot / lt : 0.0f);
    return;
  }
  // publish this split's (max, sum, acc); the last CTA to arrive for (b,h) merges all splits
  const int bh = b * gridDim.x + h;
  int* counters = reinterpret_cast<int*>(p.scratch);
  float* part = p.scratch + gridDim.x * gridDim.y + static_cast<long long>(bh) * nsplit * (D + 2);
  if (threadIdx.x < D) {
    float* mine = part + blockIdx.z * (D + 2);
    mine[threadIdx.x] = ot;
    if (threadIdx.x == 0) {
      mine[D] = mm;
      mine[D + 1] = lt;
    }
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(&counters[bh], 1) == nsplit - 1);
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  if (threadIdx.x < D) {
    float gm = -INFINITY;
    for (int z = 0; z < nsplit; ++z) gm = fmaxf(gm, __ldcg(part + z * (D + 2) + D));
    const float gsafe = (gm == -INFINITY) ? 0.0f : gm;
    float gl_ = 0.0f, go = 0.0f;
    for (int z = 0; z < nsplit; ++z) {
      const float c = exp2f(__ldcg(part + z * (D + 2) + D) - gsafe);
      gl_ += __ldcg(part + z * (D + 2) + D + 1) * c;
      go += __ldcg(part + z * (D + 2) + threadIdx.x) * c;
    }
    optr[threadIdx.x] = __float2bfloat16_rn(gl_ > 0.0f ? go / gl_ : 0.0f);
    if (threadIdx.x == 0) counters[bh] = 0;  // self-cleaning for the next launch
  }
}

template <int D>
static int launch_flash(const AttnParams& p, cudaStream_t stream) {
  constexpr int smem = (FA_BM + 4 * FA_BN) * D * 2;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(flash_fwd_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess)
      return MPL_ERR_CUDA;
    attr_set = true;
  }
  dim3 grid((p.Tq + FA_BM - 1) / FA_BM, p.H, p.B);
  flash_fwd_kernel<D><<<grid, FA_THREADS, smem, stream>>>(p);
  return mpl::launch_status();
}

// ------------------------------------------------------------------------------------------ training backward (d = 128)
// FlashAttention-2 style backward of causal self-attention on mma.sync m16n8k16 with register-resident accumulators
// (replaces torch autograd through HF-4.31 LlamaAttention's eager softmax(QK^T)V in the train step, a-17).
// One CTA (4 warps) owns 64 keys; warp w owns keys [16w, 16w+16) and keeps dK_w, dV_w (16 x 128 fp32 each) in registers
// while it walks the query blocks that can see these keys. Everything is computed TRANSPOSED (keys are the M rows):
//   S^T = K_w Q^T, dP^T = V_w dO^T  ->  P^T = 2^(S^T c - lse_q), dS^T = P^T (dP^T - delta_q) scale   (C fragments)
//   dV_w += P^T dO, dK_w += dS^T Q   (the C fragments are re-packed as A fragments: no shared-memory round trip)
//   dQ += dS K needs dS with queries as rows: dS^T goes through an 8 KB shared tile once, warp w then produces the
//   16 x 128 slice of its 16 queries over this CTA's 64 keys and reduces it into the fp32 dQ with 8-byte vector REDs.
// Q / dO tiles are double-buffered with cp.async; tiles use the forward kernel's XOR swizzle and ldmatrix patterns.
struct AttnBwd2Params {
  const __nv_bfloat16 *q, *k, *v, *dO;
  long long q_sb, q_st, q_sh, k_sb, k_st, k_sh, v_sb, v_st, v_sh, o_sb, o_st, o_sh;
  const float *lse, *delta;  // [B*H, T]
  float* dq;                 // f32 [B, T, H, 128] contiguous, zero-initialised
  __nv_bfloat16 *dk, *dv;
  long long dk_sb, dk_st, dk_sh, dv_sb, dv_st, dv_sh;
  int B, H, T;
  float scale;
  int causal;
  const unsigned char* kv_mask;
  long long kv_mask_stride;
};

constexpr int AB2_SMEM = 6 * 64 * 128 * 2 + 64 * 64 * 2 + 4 * 64 * 4;  // K, V, 2x(Q, dO), dS^T, 2x(lse, delta)

__global__ void __launch_bounds__(FA_THREADS, 2) attn_bwd_fa2_kernel(const AttnBwd2Params p) {
  constexpr int D = 128;
  extern __shared__ __align__(128) uint8_t ab2_smem[];
  const uint32_t sK = smem_u32(ab2_smem);
  const uint32_t sV = sK + 64 * D * 2;
  const uint32_t sQ0 = sV + 64 * D * 2;
  const uint32_t sdO0 = sQ0 + 2 * 64 * D * 2;
  const uint32_t sdS = sdO0 + 2 * 64 * D * 2;
  uint8_t* sdS_ptr = ab2_smem + 6 * 64 * D * 2;
  float* sLse = reinterpret_cast<float*>(ab2_smem + 6 * 64 * D * 2 + 64 * 64 * 2);  // [2][64]
  float* sDelta = sLse + 128;                                                        // [2][64]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t4 = lane & 3;
  const int b = blockIdx.z, h = blockIdx.y;
  const int k0 = blockIdx.x * 64;
  const int T = p.T;
  const __nv_bfloat16* qg = p.q + b * p.q_sb + h * p.q_sh;
  const __nv_bfloat16* kg = p.k + b * p.k_sb + h * p.k_sh;
  const __nv_bfloat16* vg = p.v + b * p.v_sb + h * p.v_sh;
  const __nv_bfloat16* og = p.dO + b * p.o_sb + h * p.o_sh;
  const long long bh = static_cast<long long>(b) * p.H + h;
  const unsigned char* mrow = p.kv_mask ? p.kv_mask + static_cast<long long>(b) * p.kv_mask_stride : nullptr;
  const float sl2 = p.scale * 1.4426950408889634f;

  const int q_begin = p.causal ? k0 : 0;
  load_tile<D>(sK, kg, p.k_st, k0, T, 64);
  load_tile<D>(sV, vg, p.v_st, k0, T, 64);
  load_tile<D>(sQ0, qg, p.q_st, q_begin, T, 64);
  load_tile<D>(sdO0, og, p.o_st, q_begin, T, 64);
  asm volatile("cp.async.commit_group;" ::: "memory");

  float dv[D / 8][4], dk[D / 8][4];
#pragma unroll
  for (int i = 0; i < D / 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) dv[i][j] = dk[i][j] = 0.0f;
  // this lane's two key rows and their validity (bounds + key-padding mask)
  const int key0 = k0 + warp * 16 + g, key1 = key0 + 8;
  const bool kok0 = key0 < T && (mrow == nullptr || mrow[key0] != 0);
  const bool kok1 = key1 < T && (mrow == nullptr || mrow[key1] != 0);

  int buf = 0;
  for (int q0 = q_begin; q0 < T; q0 += 64, buf ^= 1) {
    cp_async_wait_all();
    __syncthreads();  // tiles of this iteration landed; every warp is done with the previous iteration
    const uint32_t sQ = sQ0 + buf * 64 * D * 2, sdO = sdO0 + buf * 64 * D * 2;
    if (q0 + 64 < T) {
      load_tile<D>(sQ0 + (buf ^ 1) * 64 * D * 2, qg, p.q_st, q0 + 64, T, 64);
      load_tile<D>(sdO0 + (buf ^ 1) * 64 * D * 2, og, p.o_st, q0 + 64, T, 64);
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
    if (threadIdx.x < 64) {
      const int t = q0 + threadIdx.x;
      sLse[buf * 64 + threadIdx.x] = t < T ? p.lse[bh * T + t] : INFINITY;
      sDelta[buf * 64 + threadIdx.x] = t < T ? p.delta[bh * T + t] : 0.0f;
    }
    __syncthreads();

    // ---- S^T = K_w Q^T and dP^T = V_w dO^T : [16 keys x 64 queries] each
    float st[8][4], dpt[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) st[i][j] = dpt[i][j] = 0.0f;
#pragma unroll
    for (int kk = 0; kk < D / 16; ++kk) {
      uint32_t ka[4], va[4];
      {
        const int r = warp * 16 + (lane & 15);
        const int c = kk * 2 + (lane >> 4);
        ldmatrix_x4(ka, sK + swz<D>(r, c));
        ldmatrix_x4(va, sV + swz<D>(r, c));
      }
#pragma unroll
      for (int np = 0; np < 4; ++np) {
        uint32_t qb[4], ob[4];
        const int r = np * 16 + (lane & 7) + ((lane >> 4) << 3);
        const int c = kk * 2 + ((lane >> 3) & 1);
        ldmatrix_x4(qb, sQ + swz<D>(r, c));
        ldmatrix_x4(ob, sdO + swz<D>(r, c));
        mma_16816(st[np * 2], ka, qb[0], qb[1]);
        mma_16816(st[np * 2 + 1], ka, qb[2], qb[3]);
        mma_16816(dpt[np * 2], va, ob[0], ob[1]);
        mma_16816(dpt[np * 2 + 1], va, ob[2], ob[3]);
      }
    }
    // ---- P^T, dS^T (C fragments: rows = keys g / g+8, columns = queries nt*8 + 2*t4 + {0,1}) -> A fragments
    uint32_t pa[4][4], dsa[4][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      float pv[4], dsv[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int ql = nt * 8 + t4 * 2 + (j & 1);
        const int qi = q0 + ql;
        const int key = (j >> 1) ? key1 : key0;
        bool ok = ((j >> 1) ? kok1 : kok0) && qi < T;
        if (p.causal) ok = ok && key <= qi;
        float pr = 0.0f, ds = 0.0f;
        if (ok) {
          pr = exp2f(st[nt][j] * sl2 - sLse[buf * 64 + ql]);
          ds = pr * (dpt[nt][j] - sDelta[buf * 64 + ql]) * p.scale;
        }
        pv[j] = pr;
        dsv[j] = ds;
      }
      pa[nt >> 1][(nt & 1) * 2] = pack_bf16(pv[0], pv[1]);
      pa[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16(pv[2], pv[3]);
      const uint32_t d01 = pack_bf16(dsv[0], dsv[1]), d23 = pack_bf16(dsv[2], dsv[3]);
      dsa[nt >> 1][(nt & 1) * 2] = d01;
      dsa[nt >> 1][(nt & 1) * 2 + 1] = d23;
      // dS^T tile for the dQ pass: row = key (within the CTA's 64), 16-byte chunk nt, element pair t4
      const int r0 = warp * 16 + g, r1 = r0 + 8;
      *reinterpret_cast<uint32_t*>(sdS_ptr + swz<64>(r0, nt) + t4 * 4) = d01;
      *reinterpret_cast<uint32_t*>(sdS_ptr + swz<64>(r1, nt) + t4 * 4) = d23;
    }
    // ---- dV_w += P^T dO ; dK_w += dS^T Q   (contraction over the 64 queries)
#pragma unroll
    for (int kq = 0; kq < 4; ++kq) {
#pragma unroll
      for (int dp = 0; dp < D / 16; ++dp) {
        uint32_t ob[4], qb[4];
        const int r = kq * 16 + (lane & 7) + (((lane >> 3) & 1) << 3);
        const int c = dp * 2 + (lane >> 4);
        ldmatrix_x4_trans(ob, sdO + swz<D>(r, c));
        ldmatrix_x4_trans(qb, sQ + swz<D>(r, c));
        mma_16816(dv[dp * 2], pa[kq], ob[0], ob[1]);
        mma_16816(dv[dp * 2 + 1], pa[kq], ob[2], ob[3]);
        mma_16816(dk[dp * 2], dsa[kq], qb[0], qb[1]);
        mma_16816(dk[dp * 2 + 1], dsa[kq], qb[2], qb[3]);
      }
    }
    __syncthreads();  // dS^T of all four warps is in shared memory
    // ---- dQ[16 queries of this warp, 128] = dS[16 x 64 keys] K[64 keys x 128], reduced into the fp32 dQ
    {
      float dq[D / 8][4];
#pragma unroll
      for (int i = 0; i < D / 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dq[i][j] = 0.0f;
#pragma unroll
      for (int kkk = 0; kkk < 4; ++kkk) {
        uint32_t da[4];
        {
          // A = dS (rows = queries) read from the [key][query] tile: transposed 8x8 blocks
          const int r = kkk * 16 + (lane & 7) + ((lane >> 4) << 3);
          const int c = warp * 2 + ((lane >> 3) & 1);
          ldmatrix_x4_trans(da, sdS + swz<64>(r, c));
        }
#pragma unroll
        for (int dp = 0; dp < D / 16; ++dp) {
          uint32_t kb[4];
          const int r = kkk * 16 + (lane & 7) + (((lane >> 3) & 1) << 3);
          const int c = dp * 2 + (lane >> 4);
          ldmatrix_x4_trans(kb, sK + swz<D>(r, c));
          mma_16816(dq[dp * 2], da, kb[0], kb[1]);
          mma_16816(dq[dp * 2 + 1], da, kb[2], kb[3]);
        }
      }
#pragma unroll
      for (int rr = 0; rr < 2; ++rr) {
        const int qi = q0 + warp * 16 + g + rr * 8;
        if (qi < T) {
          float* drow = p.dq + ((static_cast<long long>(b) * T + qi) * p.H + h) * D + t4 * 2;
#pragma unroll
          for (int i = 0; i < D / 8; ++i)
            atomicAdd(reinterpret_cast<float2*>(drow + i * 8), make_float2(dq[i][rr * 2], dq[i][rr * 2 + 1]));
        }
      }
    }
  }
  // ---- write dK_w, dV_w
  __nv_bfloat16* dkg = p.dk + b * p.dk_sb + h * p.dk_sh;
  __nv_bfloat16* dvg = p.dv + b * p.dv_sb + h * p.dv_sh;
#pragma unroll
  for (int rr = 0; rr < 2; ++rr) {
    const int key = rr ? key1 : key0;
    if (key < T) {
      __nv_bfloat16* kr = dkg + static_cast<long long>(key) * p.dk_st + t4 * 2;
      __nv_bfloat16* vr = dvg + static_cast<long long>(key) * p.dv_st + t4 * 2;
#pragma unroll
      for (int i = 0; i < D / 8; ++i) {
        *reinterpret_cast<uint32_t*>(kr + i * 8) = pack_bf16(dk[i][rr * 2], dk[i][rr * 2 + 1]);
        *reinterpret_cast<uint32_t*>(vr + i * 8) = pack_bf16(dv[i][rr * 2], dv[i][rr * 2 + 1]);
      }
    }
  }
}

// Host entry used by mpl_attention_bwd (train.cu) for head_dim 128.
int attn_bwd_fa2(const mpl_attn_bwd_args& a, cudaStream_t stream) {
  AttnBwd2Params p;
  p.q = static_cast<const __nv_bfloat16*>(a.q);
  p.k = static_cast<const __nv_bfloat16*>(a.k);
  p.v = static_cast<const __nv_bfloat16*>(a.v);
  p.dO = static_cast<const __nv_bfloat16*>(a.d_o);
  p.q_sb = a.q_stride[0]; p.q_st = a.q_stride[1]; p.q_sh = a.q_stride[2];
  p.k_sb = a.k_stride[0]; p.k_st = a.k_stride[1]; p.k_sh = a.k_stride[2];
  p.v_sb = a.v_stride[0]; p.v_st = a.v_stride[1]; p.v_sh = a.v_stride[2];
  p.o_sb = a.o_stride[0]; p.o_st = a.o_stride[1]; p.o_sh = a.o_stride[2];
  p.lse = a.lse;
  p.delta = a.delta;
  p.dq = a.dq_f32;
  p.dk = static_cast<__nv_bfloat16*>(a.dk);
  p.dv = static_cast<__nv_bfloat16*>(a.dv);
  p.dk_sb = a.dk_stride[0]; p.dk_st = a.dk_stride[1]; p.dk_sh = a.dk_stride[2];
  p.dv_sb = a.dv_stride[0]; p.dv_st = a.dv_stride[1]; p.dv_sh = a.dv_stride[2];
  p.B = a.B; p.H = a.H; p.T = a.T;
  p.scale = a.scale;
  p.causal = a.causal;
  p.kv_mask = a.kv_mask;
  p.kv_mask_stride = a.kv_mask_stride > 0 ? a.kv_mask_stride : a.T;
  static bool set = false;
  if (!set) {
    if (cudaFuncSetAttribute(attn_bwd_fa2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AB2_SMEM) != cudaSuccess)
      return MPL_ERR_CUDA;
    set = true;
  }
  dim3 grid((a.T + 63) / 64, a.H, a.B);
  attn_bwd_fa2_kernel<<<grid, FA_THREADS, AB2_SMEM, stream>>>(p);
  return launch_status();
}

}  // namespace mpl


namespace mpl {
// ------------------------------------------------------------------------------------------------- small attention
// The SAM mask decoder's two-way attention (transformer.py:185-244): 8 heads of 16 (cross) or 32 (self) channels, 6 token
// rows against 256 image rows and back -- 50-180 KB per call, pure latency on the mma.sync tile kernel (one 64-row tile
// per CTA mostly empty: 8-16 us per launch, profiles/r02_ncu_maskdec_attention.md). Here ONE WARP owns one query row:
// lane l scores keys l, l + 32, ... (Tk <= 256: at most 8 scores per lane, kept in registers), the row maximum and sum are
// warp-shuffle reductions, the normalised probabilities are rounded to bf16 exactly where the eager reference rounds them
// (softmax output -> bf16 -> @ V), every lane accumulates its keys' share of P V and a last shuffle reduction adds the lanes.
template <int HD>
__global__ void __launch_bounds__(128) attn_small_kernel(const AttnParams p) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long item = static_cast<long long>(blockIdx.x) * 4 + warp;
  if (item >= static_cast<long long>(p.B) * p.H * p.Tq) return;
  const int i = static_cast<int>(item % p.Tq);
  const long long bh = item / p.Tq;
  const int h = static_cast<int>(bh % p.H);
  const long long b = bh / p.H;
  const __nv_bfloat16* qr = p.q + b * p.q_sb + i * p.q_st + h * p.q_sh;
  const __nv_bfloat16* kb = p.k + b * p.k_sb + h * p.k_sh;
  const __nv_bfloat16* vb = p.v + b * p.v_sb + h * p.v_sh;
  float qf[HD];
#pragma unroll
  for (int c = 0; c < HD / 8; ++c) {
    const uint4 raw = *reinterpret_cast<const uint4*>(qr + c * 8);
    const __nv_bfloat16* hv = reinterpret_cast<const __nv_bfloat16*>(&raw);
#pragma unroll
    for (int e = 0; e < 8; ++e) qf[c * 8 + e] = __bfloat162float(hv[e]);
  }
  const float sl2 = p.scale * 1.4426950408889634f;
  float sc[8];
  float mx = -INFINITY;
#pragma unroll
  for (int t = 0; t < 8; ++t) {
    const int j = lane + 32 * t;
    sc[t] = -INFINITY;
    if (j < p.Tk) {
      const __nv_bfloat16* kr = kb + static_cast<long long>(j) * p.k_st;
      float dot = 0.0f;
#pragma unroll
      for (int c = 0; c < HD / 8; ++c) {
        const uint4 raw = *reinterpret_cast<const uint4*>(kr + c * 8);
        const __nv_bfloat16* hv = reinterpret_cast<const __nv_bfloat16*>(&raw);
#pragma unroll
        for (int e = 0; e < 8; ++e) dot += qf[c * 8 + e] * __bfloat162float(hv[e]);
      }
      sc[t] = dot * sl2;
      mx = fmaxf(mx, sc[t]);
    }
  }
  mx = warp_max(mx);
  float l = 0.0f;
#pragma unroll
  for (int t = 0; t < 8; ++t) {
    sc[t] = lane + 32 * t < p.Tk ? exp2f(sc[t] - mx) : 0.0f;
    l += sc[t];
  }
  l = warp_sum(l);
  const float inv = 1.0f / l;
  float acc[HD];
#pragma unroll
  for (int e = 0; e < HD; ++e) acc[e] = 0.0f;
#pragma unroll
  for (int t = 0; t < 8; ++t) {
    const int j = lane + 32 * t;
    if (j < p.Tk) {
      const float pj = bf16_round(sc[t] * inv);
      const __nv_bfloat16* vr = vb + static_cast<long long>(j) * p.v_st;
#pragma unroll
      for (int c = 0; c < HD / 8; ++c) {
        const uint4 raw = *reinterpret_cast<const uint4*>(vr + c * 8);
        const __nv_bfloat16* hv = reinterpret_cast<const __nv_bfloat16*>(&raw);
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[c * 8 + e] += pj * __bfloat162float(hv[e]);
      }
    }
  }
#pragma unroll
  for (int e = 0; e < HD; ++e) acc[e] = warp_sum(acc[e]);
  if (lane == 0) {
    __nv_bfloat16* orow = p.o + b * p.o_sb + i * p.o_st + h * p.o_sh;
#pragma unroll
    for (int e = 0; e < HD; e += 2)
      *reinterpret_cast<uint32_t*>(orow + e) = pack_bf16(acc[e], acc[e + 1]);
  }
}
static bool attn_small_enabled() {
  static const bool on = [] {
    const char* c = getenv("MPL_ATTN_SMALL");
    return c == nullptr || atoi(c) != 0;
  }();
  return on;
}
}  // namespace mpl

extern "C" int mpl_attention(const mpl_attn_args* a, void* stream_) {
  using namespace mpl;
  if (a == nullptr || a->q == nullptr || a->k == nullptr || a->v == nullptr || a->o == nullptr) return MPL_ERR_ARG;
  if (a->B <= 0 || a->H <= 0 || a->Tq <= 0) return MPL_OK;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  AttnParams p;
  p.q = static_cast<const __nv_bfloat16*>(a->q);
  p.k = static_cast<const __nv_bfloat16*>(a->k);
  p.v = static_cast<const __nv_bfloat16*>(a->v);
  p.o = static_cast<__nv_bfloat16*>(a->o);
  p.q_sb = a->q_stride[0]; p.q_st = a->q_stride[1]; p.q_sh = a->q_stride[2];
  p.k_sb = a->k_stride[0]; p.k_st = a->k_stride[1]; p.k_sh = a->k_stride[2];
  p.v_sb = a->v_stride[0]; p.v_st = a->v_stride[1]; p.v_sh = a->v_stride[2];
  p.o_sb = a->o_stride[0]; p.o_st = a->o_stride[1]; p.o_sh = a->o_stride[2];
  p.B = a->B; p.H = a->H; p.Tq = a->Tq; p.Tk = a->Tk;
  p.scale = a->scale;
  p.causal = a->causal;
  p.kv_mask = a->kv_mask;
  p.kv_mask_stride = a->kv_mask_stride > 0 ? a->kv_mask_stride : a->Tk;
  p.rel_h = a->rel_h; p.rel_w = a->rel_w; p.rel_kh = a->rel_kh; p.rel_kw = a->rel_kw;
  p.tk_dev = a->tk_dev;
  p.lse = a->lse;
  // 16-byte vector access on every row
  const long long strides[] = {p.q_sb, p.q_st, p.q_sh, p.k_sb, p.k_st, p.k_sh, p.v_sb, p.v_st, p.v_sh};
  for (long long s : strides)
    if (s % 8 != 0) return MPL_ERR_ALIGN;
  if (a->Tq == 1 && a->head_dim == 128 && a->rel_h == nullptr) {
    int nsplit = 1;
    if (a->scratch != nullptr) {
      const int bh = p.B * p.H;
      nsplit = (2 * num_sms() + bh - 1) / bh;
      const int by_keys = (p.Tk + 63) / 64;  // at least 64 keys per split
      if (nsplit > by_keys) nsplit = by_keys;
      if (nsplit > 32) nsplit = 32;
      while (nsplit > 1 && static_cast<long long>(bh) * 4 + static_cast<long long>(bh) * nsplit * 130 * 4 > a->scratch_bytes)
        --nsplit;
      if (nsplit < 1) nsplit = 1;
    }
    p.scratch = static_cast<float*>(a->scratch);
    p.nsplit = nsplit;
    dim3 grid(p.H, p.B, nsplit);
    decode_attn_kernel<128><<<grid, DEC_WARPS * 32, 0, stream>>>(p);
    return mpl::launch_status();
  }
  if (p.tk_dev != nullptr) return MPL_ERR_UNSUPPORTED;  // device-side Tk only on the decode path
  if (attention_tc_supported(*a)) return attention_tc(*a, stream);  // tcgen05 + TMA tiles (attention_tc.cu)
  if (p.o_st % 2 != 0 || p.o_sh % 2 != 0 || p.o_sb % 2 != 0) return MPL_ERR_ALIGN;
  if ((a->head_dim == 16 || a->head_dim == 32) && a->Tq <= 64 && a->Tk >= 1 && a->Tk <= 256 && !a->causal && a->kv_mask == nullptr &&
      a->rel_h == nullptr && a->lse == nullptr && attn_small_enabled()) {
    const long long items = static_cast<long long>(p.B) * p.H * p.Tq;
    const unsigned grid = static_cast<unsigned>((items + 3) / 4);
    if (a->head_dim == 16)
      attn_small_kernel<16><<<grid, 128, 0, stream>>>(p);
    else
      attn_small_kernel<32><<<grid, 128, 0, stream>>>(p);
    return mpl::launch_status();
  }
  switch (a->head_dim) {
    case 16: return launch_flash<16>(p, stream);
    case 32: return launch_flash<32>(p, stream);
    case 64: return launch_flash<64>(p, stream);
    case 128: return launch_flash<128>(p, stream);
    default: return MPL_ERR_UNSUPPORTED;
  }
}
