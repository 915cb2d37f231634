// K9 — one persistent kernel per decode step of the LLaMA-MoE stack (T == 1, B <= 8 sequences, top-1 routing).
//
// Replaces, for the autoregressive steps of MedPLIBForCausalLM.generate / evaluate (model/MedPLIB.py:592-606,
// model/medplib/model/language_model/medplib_moe_llama.py:110-305,451-485), the ~7 launches per layer of the general
// runner (llama_stack.cu). A decode step is HBM-bound: 13-22 GB of weights are read once per token, and every kernel
// boundary costs the weight stream ~8 us of ramp-down / launch / ramp-up (measured: 152 us per layer against a 62 us
// roofline). Here ONE cooperative grid (one CTA per SM) runs all layers:
//   * a producer thread per CTA walks the whole step's weight-tile schedule and keeps a 6 x 16 KB TMA ring full; it does
//     not take part in grid barriers, so the HBM stream continues across phase boundaries (the next phase's weights are
//     already in shared memory when the consumers arrive) — the only stall is the wait for the router's expert choice;
//   * 8 consumer warps run the phases of a layer, separated by grid barriers (atomic counter in L2):
//       P1  q,k,v = RMSNorm(x) Wqkv^T           activations normalised once per CTA into shared memory
//       P2  RoPE(q, k_new) + KV append + split-K attention over the cache, last-arriver merge per (b, h)
//       P3  x += attn Wo^T
//       P4  h = RMSNorm(x); router logits / softmax / top-1 / capacity slots — recomputed by every CTA (no barrier)
//       P5  h1 = SiLU(h Wgate_e^T) * (h Wup_e^T)  only experts that received tokens are streamed
//       P6  x += gate * (h1 Wdown_e^T)          MoE combine fused (dense layers: plain residual)
//     then the final RMSNorm. The GEMM core is the streaming kernel's (skinny_gemm.cu): mma.sync m16n8k16 with the
//     weight rows as the M operand, k split over the warps, cross-warp reduction through shared memory.
// Rounding points are those of the general path (bf16 after every linear, RMSNorm's two roundings, RoPE's three, P
// rounded before P·V, bf16(gate) * bf16(y)), so both paths produce the same bits.
#include <cooperative_groups.h>
#include <cuda.h>

#include <cmath>
#include <cstring>
#include <vector>

#include "internal.h"
#include "ptx.cuh"

namespace mpl {

constexpr int DK_CONSUMERS = 8;
constexpr int DK_THREADS = (DK_CONSUMERS + 1) * 32;
constexpr int DK_STAGES = 6;
constexpr int DK_STAGE_BYTES = 16384;  // [8 k-blocks][16 rows][128 B swizzled]  (dual: 2 x [8][8 rows][128 B])
constexpr int DK_KC = 512;
constexpr int DK_MAXB = 8;
constexpr int DK_MAXE = MPL_MAX_EXPERTS;
constexpr int DK_RP = 9;  // reduction row pitch (floats)
constexpr int DK_RED_FLOATS = DK_CONSUMERS * 16 * DK_RP;

// Per-layer device-resident description (the "decode plan"): tensor maps + the small raw pointers.
struct alignas(64) DecLayerDev {
  CUtensorMap wq, wk, wv, wo;  // 3-D {64, N, K/64}, box {64, 16, 8}
  CUtensorMap wgate[DK_MAXE];  // box {64, 8, 8}
  CUtensorMap wup[DK_MAXE];
  CUtensorMap wdown[DK_MAXE];  // box {64, 16, 8}
  const __nv_bfloat16* input_ln;
  const __nv_bfloat16* post_ln;
  const float* wg;  // NULL: dense layer
  int n_experts;
  int pad_[9];
};
constexpr long long DK_SYNC_BYTES = 256;  // [0] grid-barrier counter, [1] exit counter

struct DecParams {
  const DecLayerDev* layers;
  unsigned int* sync;
  __nv_bfloat16* x;         // [B, D] in/out
  __nv_bfloat16* out_norm;  // [B, D] or NULL
  const __nv_bfloat16* final_norm;
  __nv_bfloat16* qkv;   // [B, 3D]
  __nv_bfloat16* attn;  // [B, D]
  __nv_bfloat16* h1;    // [E*B, F]
  float* attn_part;     // split-K partials [B*H][nsplit][130]
  int* attn_cnt;        // [B*H] zero-initialised, self-cleaning
  __nv_bfloat16* kc;
  __nv_bfloat16* vc;
  long long cache_layer;  // elements per layer of the KV cache
  const __nv_bfloat16* cos_t;
  const __nv_bfloat16* sin_t;
  const unsigned char* kv_mask;
  long long kv_mask_stride;
  const int* pos_dev;
  float* gate_logits;  // [L, B, Emax] or NULL
  float* l_aux;        // [L] or NULL
  int* exp_counts;     // [L, Emax] or NULL
  int B, D, H, F, L, Tmax, pos, nsplit, Emax;
  int cap[DK_MAXE + 1];  // capacity for a layer with E experts (index E)
  float eps, scale;
};

struct Ring {
  uint8_t* base;
  uint64_t* full;
  uint64_t* empty;
  int stage;
  uint32_t phase;
  __device__ __forceinline__ void advance() {
    if (++stage == DK_STAGES) {
      stage = 0;
      phase ^= 1;
    }
  }
};

// routing decision of the current layer, written by consumer thread 0, read by everyone (and by the producer)
struct RouteSmem {
  int kept[DK_MAXE];
  int tok_of_slot[DK_MAXE][DK_MAXB];
  float gate_of_slot[DK_MAXE][DK_MAXB];
  unsigned int amask;
  int moe;
  float logits[DK_MAXB][DK_MAXE];
  float gates[DK_MAXB][DK_MAXE];
};

__device__ __forceinline__ void dk_hmma(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                        uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint4 dk_lds128(uint32_t addr) {
  uint4 r;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(addr));
  return r;
}
__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// Grid barrier for the consumer warps (the producer thread never waits here). Same structure as cooperative
// groups' grid.sync(): CTA barrier, one thread fences + arrives + spins + fences (the gpu-scope fence also invalidates
// this SM's L1, so plain loads after the barrier see the other CTAs' writes), CTA barrier.
__device__ __forceinline__ void grid_sync(unsigned int* ctr, unsigned int& target) {
  consumer_sync();
  if (threadIdx.x == 0) {
    target += gridDim.x;
    __threadfence();
    atomicAdd(ctr, 1u);
    while (ld_acquire_u32(ctr) < target) {
    }
    __threadfence();
  }
  consumer_sync();
}

// ------------------------------------------------------------------------------------------------ producer side
__device__ __forceinline__ void produce_tile(Ring& r, const CUtensorMap* m0, const CUtensorMap* m1, int n0, int chunks) {
  for (int c = 0; c < chunks; ++c) {
    mbar_wait(&r.empty[r.stage], r.phase ^ 1);
    uint8_t* dst = r.base + r.stage * DK_STAGE_BYTES;
    mbar_expect_tx(&r.full[r.stage], DK_STAGE_BYTES);
    tma_load_3d(dst, m0, &r.full[r.stage], 0, n0, c * (DK_KC / 64));
    if (m1 != nullptr) tma_load_3d(dst + DK_STAGE_BYTES / 2, m1, &r.full[r.stage], 0, n0, c * (DK_KC / 64));
    r.advance();
  }
}

// ------------------------------------------------------------------------------------------------ consumer GEMM core
// One tile: 16 weight rows (DUAL: 8 gate + 8 up rows) x K, activation row of this lane's column group `arow` (generic
// pointer into shared or global memory, or NULL for an empty column). Leaves the CTA-reduced sums in rbuf[row][m].
template <bool DUAL>
__device__ __forceinline__ void consume_tile(Ring& r, int chunks, int K, const __nv_bfloat16* arow, float* rbuf) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  constexpr int ROWS = DUAL ? 8 : 16;
  float acc0[4] = {0.f, 0.f, 0.f, 0.f};  // one accumulator, same HMMA order as the streaming kernel: identical bits
  const uint4 zero = make_uint4(0, 0, 0, 0);
  uint4 xn0 = zero, xn1 = zero;
  {
    const int k = warp * 64 + t * 8;
    if (arow != nullptr && k < K) xn0 = *reinterpret_cast<const uint4*>(arow + k);
    if (arow != nullptr && k + 32 < K) xn1 = *reinterpret_cast<const uint4*>(arow + k + 32);
  }
  for (int c = 0; c < chunks; ++c) {
    const uint4 xb0 = xn0, xb1 = xn1;
    if (c + 1 < chunks) {  // next chunk's activation fragment in flight while this chunk's weights are consumed
      const int k = (c + 1) * DK_KC + warp * 64 + t * 8;
      xn0 = (arow != nullptr && k < K) ? *reinterpret_cast<const uint4*>(arow + k) : zero;
      xn1 = (arow != nullptr && k + 32 < K) ? *reinterpret_cast<const uint4*>(arow + k + 32) : zero;
    }
    mbar_wait(&r.full[r.stage], r.phase);
    const uint32_t sbase = smem_u32(r.base + r.stage * DK_STAGE_BYTES) + warp * (ROWS * 128) + g * 128;
    const uint32_t hi = DUAL ? DK_STAGE_BYTES / 2 : 1024;
    {
      const uint32_t sw = static_cast<uint32_t>((t ^ g) * 16);
      const uint4 wa = dk_lds128(sbase + sw), wb = dk_lds128(sbase + hi + sw);
      dk_hmma(acc0, wa.x, wb.x, wa.y, wb.y, xb0.x, xb0.y);
      dk_hmma(acc0, wa.z, wb.z, wa.w, wb.w, xb0.z, xb0.w);
    }
    {
      const uint32_t sw = static_cast<uint32_t>(((4 + t) ^ g) * 16);
      const uint4 wa = dk_lds128(sbase + sw), wb = dk_lds128(sbase + hi + sw);
      dk_hmma(acc0, wa.x, wb.x, wa.y, wb.y, xb1.x, xb1.y);
      dk_hmma(acc0, wa.z, wb.z, wa.w, wb.w, xb1.z, xb1.w);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&r.empty[r.stage]);
    r.advance();
  }
  // C fragment: c0,c1 -> (weight row g, m = 2t, 2t+1); c2,c3 -> (weight row g+8, same m)
  float* rw = rbuf + warp * 16 * DK_RP;
  rw[g * DK_RP + t * 2] = acc0[0];
  rw[g * DK_RP + t * 2 + 1] = acc0[1];
  rw[(g + 8) * DK_RP + t * 2] = acc0[2];
  rw[(g + 8) * DK_RP + t * 2 + 1] = acc0[3];
  consumer_sync();
}
__device__ __forceinline__ float reduce_rows(const float* rbuf, int r, int m) {
  float v = 0.0f;
#pragma unroll
  for (int w = 0; w < DK_CONSUMERS; ++w) v += rbuf[(w * 16 + r) * DK_RP + m];
  return v;
}

// RMSNorm of the B activation rows into shared memory (one warp per row), HF LlamaRMSNorm roundings:
// w * bf16(x * rstd). ln == NULL: plain copy.
__device__ __forceinline__ void stage_rows(const DecParams& p, const __nv_bfloat16* src, const __nv_bfloat16* ln,
                                           uint8_t* s_a, int pitch) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp < p.B) {
    const __nv_bfloat16* xr = src + static_cast<long long>(warp) * p.D;
    uint8_t* dst = s_a + warp * pitch;
    float rstd = 1.0f;
    if (ln != nullptr) {
      float ss = 0.0f;
      for (int k = lane * 8; k < p.D; k += 256) {
        const uint4 raw = *reinterpret_cast<const uint4*>(xr + k);
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 f = __bfloat1622float2(h[i]);
          ss += f.x * f.x + f.y * f.y;
        }
      }
      ss = warp_sum(ss);
      rstd = rsqrtf(ss / static_cast<float>(p.D) + p.eps);
    }
    for (int k = lane * 8; k < p.D; k += 256) {
      uint4 v = *reinterpret_cast<const uint4*>(xr + k);
      if (ln != nullptr) {
        const uint4 w = *reinterpret_cast<const uint4*>(ln + k);
        const __nv_bfloat162* xp = reinterpret_cast<const __nv_bfloat162*>(&v);
        const __nv_bfloat162* wp = reinterpret_cast<const __nv_bfloat162*>(&w);
        uint4 o;
        uint32_t* op = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 xf = __bfloat1622float2(xp[i]), wf = __bfloat1622float2(wp[i]);
          op[i] = pack_bf16(wf.x * bf16_round(xf.x * rstd), wf.y * bf16_round(xf.y * rstd));
        }
        v = o;
      }
      *reinterpret_cast<uint4*>(dst + k * 2) = v;
    }
  }
}

// ------------------------------------------------------------------------------------------------ attention item
// One (b, h, split) of the decode attention: 8 warps, 4 keys per warp step, 8 lanes x 16 dims per key. q and the new
// key are rotated on the fly from the q,k,v buffer; the split that owns position `pos` appends k,v to the cache.
__device__ __forceinline__ void rope16(const __nv_bfloat16* src, const __nv_bfloat16* cr, const __nv_bfloat16* sr,
                                       int gl, float (&out)[16]) {
  // dims d = gl*16 .. +15 of a 128-wide head; partner = d + 64 (first half, rotated with a minus sign) or d - 64
  const int d0 = gl * 16;
  const bool first = gl < 4;
  const __nv_bfloat16* own = src + d0;
  const __nv_bfloat16* par = src + (first ? d0 + 64 : d0 - 64);
#pragma unroll
  for (int e = 0; e < 16; ++e) {
    const float f = __bfloat162float(own[e]), fp = __bfloat162float(par[e]);
    const float c = __bfloat162float(cr[d0 + e]), s = __bfloat162float(sr[d0 + e]);
    const float a = bf16_round(f * c);
    const float b = bf16_round((first ? -fp : fp) * s);
    out[e] = bf16_round(a + b);
  }
}

__device__ void attention_item(const DecParams& p, int layer, int b, int h, int z, int Tk, float* s_f) {
  constexpr int D = 128;
  float* s_o = s_f;                     // [8][128]
  float* s_m = s_f + DK_CONSUMERS * D;  // [8]
  float* s_l = s_m + DK_CONSUMERS;      // [8]
  int* s_last = reinterpret_cast<int*>(s_l + DK_CONSUMERS);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int grp = lane >> 3, gl = lane & 7;
  const int nsplit = p.nsplit;
  const int per = ((Tk + nsplit - 1) / nsplit + 31) & ~31;
  const int k_lo = z * per;
  const int k_hi = min(Tk, k_lo + per);
  const int pos = Tk - 1;
  const __nv_bfloat16* qrow = p.qkv + static_cast<long long>(b) * 3 * p.D + h * D;
  const __nv_bfloat16* krow = qrow + p.D;
  const __nv_bfloat16* vrow = qrow + 2 * p.D;
  const __nv_bfloat16* cr = p.cos_t + static_cast<long long>(pos) * D;
  const __nv_bfloat16* sr = p.sin_t + static_cast<long long>(pos) * D;
  const long long head_off = (static_cast<long long>(b) * p.H + h) * p.Tmax * D;
  __nv_bfloat16* kc = p.kc + layer * p.cache_layer + head_off;
  __nv_bfloat16* vc = p.vc + layer * p.cache_layer + head_off;
  const unsigned char* mrow = p.kv_mask ? p.kv_mask + static_cast<long long>(b) * p.kv_mask_stride : nullptr;
  float qf[16];
  rope16(qrow, cr, sr, gl, qf);
  const float sl2 = p.scale * 1.4426950408889634f;
  float m = -INFINITY, l = 0.0f;
  float acc[16];
#pragma unroll
  for (int e = 0; e < 16; ++e) acc[e] = 0.0f;
  for (int k0 = k_lo + warp * 4; k0 < k_hi; k0 += DK_CONSUMERS * 4) {
    const int key = k0 + grp;
    const bool valid = key < k_hi;
    const int kk = valid ? key : k_lo;
    float kf[16], vf[16];
    if (kk == pos) {
      rope16(krow, cr, sr, gl, kf);
#pragma unroll
      for (int e = 0; e < 16; ++e) vf[e] = __bfloat162float(vrow[gl * 16 + e]);
      if (valid) {  // KV-cache append (exactly one lane group of one split owns the new position)
        uint4 o0, o1;
        o0.x = pack_bf16(kf[0], kf[1]); o0.y = pack_bf16(kf[2], kf[3]); o0.z = pack_bf16(kf[4], kf[5]); o0.w = pack_bf16(kf[6], kf[7]);
        o1.x = pack_bf16(kf[8], kf[9]); o1.y = pack_bf16(kf[10], kf[11]); o1.z = pack_bf16(kf[12], kf[13]); o1.w = pack_bf16(kf[14], kf[15]);
        __nv_bfloat16* kd = kc + static_cast<long long>(pos) * D + gl * 16;
        *reinterpret_cast<uint4*>(kd) = o0;
        *reinterpret_cast<uint4*>(kd + 8) = o1;
        __nv_bfloat16* vd = vc + static_cast<long long>(pos) * D + gl * 16;
        *reinterpret_cast<uint4*>(vd) = *reinterpret_cast<const uint4*>(vrow + gl * 16);
        *reinterpret_cast<uint4*>(vd + 8) = *reinterpret_cast<const uint4*>(vrow + gl * 16 + 8);
      }
    } else {
      const __nv_bfloat16* kr = kc + static_cast<long long>(kk) * D + gl * 16;
      const __nv_bfloat16* vr = vc + static_cast<long long>(kk) * D + gl * 16;
      const uint4 ka = *reinterpret_cast<const uint4*>(kr), kb = *reinterpret_cast<const uint4*>(kr + 8);
      const uint4 va = *reinterpret_cast<const uint4*>(vr), vb = *reinterpret_cast<const uint4*>(vr + 8);
      const __nv_bfloat162* hka = reinterpret_cast<const __nv_bfloat162*>(&ka);
      const __nv_bfloat162* hkb = reinterpret_cast<const __nv_bfloat162*>(&kb);
      const __nv_bfloat162* hva = reinterpret_cast<const __nv_bfloat162*>(&va);
      const __nv_bfloat162* hvb = reinterpret_cast<const __nv_bfloat162*>(&vb);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 a = __bfloat1622float2(hka[e]), c = __bfloat1622float2(hkb[e]);
        kf[2 * e] = a.x; kf[2 * e + 1] = a.y; kf[8 + 2 * e] = c.x; kf[8 + 2 * e + 1] = c.y;
        const float2 a2 = __bfloat1622float2(hva[e]), c2 = __bfloat1622float2(hvb[e]);
        vf[2 * e] = a2.x; vf[2 * e + 1] = a2.y; vf[8 + 2 * e] = c2.x; vf[8 + 2 * e + 1] = c2.y;
      }
    }
    // same association as the general decode kernel: pairs (e, e+1) of the low and the high 8 dims per step
    float dot = 0.0f;
#pragma unroll
    for (int e = 0; e < 4; ++e)
      dot += qf[2 * e] * kf[2 * e] + qf[2 * e + 1] * kf[2 * e + 1] + qf[8 + 2 * e] * kf[8 + 2 * e] +
             qf[8 + 2 * e + 1] * kf[8 + 2 * e + 1];
    dot += __shfl_xor_sync(0xffffffffu, dot, 1);
    dot += __shfl_xor_sync(0xffffffffu, dot, 2);
    dot += __shfl_xor_sync(0xffffffffu, dot, 4);
    float sc = dot * sl2;
    if (!valid || (mrow != nullptr && mrow[kk] == 0)) sc = -INFINITY;
    const float mn = fmaxf(m, sc);
    const float msafe = (mn == -INFINITY) ? 0.0f : mn;
    const float corr = exp2f(m - msafe);
    const float pexp = exp2f(sc - msafe);
    const float pv = bf16_round(pexp);  // P rounded to bf16 before P·V (softmax(...).to(bf16) @ v)
    l = l * corr + pexp;
    m = mn;
#pragma unroll
    for (int e = 0; e < 16; ++e) acc[e] = acc[e] * corr + pv * vf[e];
  }
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
    for (int e = 0; e < 16; ++e) s_o[warp * D + gl * 16 + e] = acc[e];
    if (gl == 0) {
      s_m[warp] = m;
      s_l[warp] = l;
    }
  }
  consumer_sync();
  float mm = -INFINITY, lt = 0.0f, ot = 0.0f;
  if (threadIdx.x < D) {
#pragma unroll
    for (int w = 0; w < DK_CONSUMERS; ++w) mm = fmaxf(mm, s_m[w]);
    const float msafe = (mm == -INFINITY) ? 0.0f : mm;
#pragma unroll
    for (int w = 0; w < DK_CONSUMERS; ++w) {
      const float c = exp2f(s_m[w] - msafe);
      lt += s_l[w] * c;
      ot += s_o[w * D + threadIdx.x] * c;
    }
  }
  __nv_bfloat16* optr = p.attn + static_cast<long long>(b) * p.D + h * D;
  if (nsplit == 1) {
    if (threadIdx.x < D) optr[threadIdx.x] = __float2bfloat16_rn(lt > 0.0f ? ot / lt : 0.0f);
    consumer_sync();  // s_o / s_m are reused by the next item
    return;
  }
  const int bh = b * p.H + h;
  float* part = p.attn_part + static_cast<long long>(bh) * nsplit * (D + 2);
  if (threadIdx.x < D) {
    float* mine = part + z * (D + 2);
    mine[threadIdx.x] = ot;
    if (threadIdx.x == 0) {
      mine[D] = mm;
      mine[D + 1] = lt;
    }
  }
  __threadfence();
  consumer_sync();
  if (threadIdx.x == 0) *s_last = (atomicAdd(&p.attn_cnt[bh], 1) == nsplit - 1);
  consumer_sync();
  if (*s_last) {
    __threadfence();
    if (threadIdx.x < D) {
      float gm = -INFINITY;
      for (int zz = 0; zz < nsplit; ++zz) gm = fmaxf(gm, __ldcg(part + zz * (D + 2) + D));
      const float gsafe = (gm == -INFINITY) ? 0.0f : gm;
      float gl_ = 0.0f, go = 0.0f;
      for (int zz = 0; zz < nsplit; ++zz) {
        const float c = exp2f(__ldcg(part + zz * (D + 2) + D) - gsafe);
        gl_ += __ldcg(part + zz * (D + 2) + D + 1) * c;
        go += __ldcg(part + zz * (D + 2) + threadIdx.x) * c;
      }
      optr[threadIdx.x] = __float2bfloat16_rn(gl_ > 0.0f ? go / gl_ : 0.0f);
      if (threadIdx.x == 0) p.attn_cnt[bh] = 0;
    }
  }
  consumer_sync();
}

// ------------------------------------------------------------------------------------------------ the kernel
__global__ void __launch_bounds__(DK_THREADS, 1) llama_decode_kernel(const DecParams p) {
  extern __shared__ uint8_t dk_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(dk_raw) + 1023) & ~uintptr_t(1023));
  const int pitch = p.D * 2 + 64;  // activation row pitch in shared memory (conflict-free 16-byte reads)
  uint8_t* s_a = smem + DK_STAGES * DK_STAGE_BYTES;
  float* red = reinterpret_cast<float*>(s_a + DK_MAXB * pitch);  // [2][8][16][9], also the attention scratch
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(red + 2 * DK_RED_FLOATS);
  uint64_t* empty_bar = full_bar + DK_STAGES;
  uint64_t* route_bar = empty_bar + DK_STAGES;
  RouteSmem* rt = reinterpret_cast<RouteSmem*>(route_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int G = gridDim.x;
  if (threadIdx.x == 0) {
    for (int s = 0; s < DK_STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], DK_CONSUMERS);
    }
    mbar_init(route_bar, 1);
    fence_mbar_init();
  }
  __syncthreads();
  Ring ring{smem, full_bar, empty_bar, 0, 0};
  const int D = p.D, F = p.F, B = p.B;
  const int tiles_d = (D + 15) / 16;         // 16-row tiles of a [D, *] matrix
  const int tiles_f = (F + 7) / 8;           // 8+8-row tiles of the gate/up pair
  const int chunks_d = (D + DK_KC - 1) / DK_KC;
  const int chunks_f = (F + DK_KC - 1) / DK_KC;

  if (warp == DK_CONSUMERS) {
    // ============================================================ producer: the whole step's weight schedule
    if (lane != 0) return;
    uint32_t route_phase = 0;
    for (int l = 0; l < p.L; ++l) {
      const DecLayerDev* L = p.layers + l;
      for (int tile = blockIdx.x; tile < 3 * tiles_d; tile += G) {
        const int which = tile / tiles_d;
        const CUtensorMap* tm = which == 0 ? &L->wq : (which == 1 ? &L->wk : &L->wv);
        produce_tile(ring, tm, nullptr, (tile % tiles_d) * 16, chunks_d);
      }
      for (int tile = blockIdx.x; tile < tiles_d; tile += G) produce_tile(ring, &L->wo, nullptr, tile * 16, chunks_d);
      mbar_wait(route_bar, route_phase);  // expert choice of this layer
      route_phase ^= 1;
      const unsigned int amask = rt->amask;
      const int nact = __popc(amask);
      for (int tile = blockIdx.x; tile < nact * tiles_f; tile += G) {
        const int e = __fns(amask, 0, tile / tiles_f + 1);
        produce_tile(ring, &L->wgate[e], &L->wup[e], (tile % tiles_f) * 8, chunks_d);
      }
      for (int tile = blockIdx.x; tile < nact * tiles_d; tile += G) {
        const int e = __fns(amask, 0, tile / tiles_d + 1);
        produce_tile(ring, &L->wdown[e], nullptr, (tile % tiles_d) * 16, chunks_f);
      }
    }
    return;
  }

  // ============================================================== consumers
  const int g = lane >> 2;
  unsigned int bar_target = 0;
  int buf = 0;
  const int pos = p.pos_dev ? *p.pos_dev : p.pos;
  const int Tk = pos + 1;
  for (int l = 0; l < p.L; ++l) {
    const DecLayerDev* L = p.layers + l;
    // ---------------------------------------------------------- P1: q,k,v = RMSNorm(x) Wqkv^T
    stage_rows(p, p.x, L->input_ln, s_a, pitch);
    consumer_sync();
    {
      const __nv_bfloat16* arow = g < B ? reinterpret_cast<const __nv_bfloat16*>(s_a + g * pitch) : nullptr;
      for (int tile = blockIdx.x; tile < 3 * tiles_d; tile += G) {
        float* rbuf = red + buf * DK_RED_FLOATS;
        buf ^= 1;
        consume_tile<false>(ring, chunks_d, D, arow, rbuf);
        const int which = tile / tiles_d, n0 = (tile % tiles_d) * 16;
        if (threadIdx.x < 16 * DK_MAXB) {
          const int r = threadIdx.x & 15, m = threadIdx.x >> 4;
          if (m < B && n0 + r < D)
            p.qkv[static_cast<long long>(m) * 3 * D + which * D + n0 + r] = __float2bfloat16_rn(reduce_rows(rbuf, r, m));
        }
      }
    }
    grid_sync(p.sync, bar_target);
    // ---------------------------------------------------------- P2: RoPE + KV append + attention over the cache
    for (int item = blockIdx.x; item < B * p.H * p.nsplit; item += G) {
      const int z = item % p.nsplit, bh = item / p.nsplit;
      attention_item(p, l, bh / p.H, bh % p.H, z, Tk, red);
    }
    grid_sync(p.sync, bar_target);
    // ---------------------------------------------------------- P3: x += attn Wo^T
    stage_rows(p, p.attn, nullptr, s_a, pitch);
    consumer_sync();
    {
      const __nv_bfloat16* arow = g < B ? reinterpret_cast<const __nv_bfloat16*>(s_a + g * pitch) : nullptr;
      for (int tile = blockIdx.x; tile < tiles_d; tile += G) {
        float* rbuf = red + buf * DK_RED_FLOATS;
        buf ^= 1;
        consume_tile<false>(ring, chunks_d, D, arow, rbuf);
        const int n0 = tile * 16;
        if (threadIdx.x < 16 * DK_MAXB) {
          const int r = threadIdx.x & 15, m = threadIdx.x >> 4;
          if (m < B && n0 + r < D) {
            __nv_bfloat16* xp = p.x + static_cast<long long>(m) * D + n0 + r;
            *xp = __float2bfloat16_rn(bf16_round(reduce_rows(rbuf, r, m)) + __bfloat162float(*xp));
          }
        }
      }
    }
    grid_sync(p.sync, bar_target);
    // ---------------------------------------------------------- P4: h = RMSNorm(x), router (every CTA, no barrier)
    stage_rows(p, p.x, L->post_ln, s_a, pitch);
    const int E = L->wg != nullptr ? L->n_experts : 1;
    if (L->wg != nullptr && warp < B) {
      // logits from the stored (bf16-rounded) h, same per-lane order as moe_router_kernel
      __syncwarp();
      const uint8_t* hr = s_a + warp * pitch;
      float acc[DK_MAXE];
#pragma unroll
      for (int e = 0; e < DK_MAXE; ++e) acc[e] = 0.0f;
      for (int c = lane * 8; c < D; c += 256) {
        const uint4 raw = *reinterpret_cast<const uint4*>(hr + c * 2);
        const __nv_bfloat162* hp = reinterpret_cast<const __nv_bfloat162*>(&raw);
        float xv[8];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 f = __bfloat1622float2(hp[i]);
          xv[2 * i] = f.x;
          xv[2 * i + 1] = f.y;
        }
#pragma unroll
        for (int e = 0; e < DK_MAXE; ++e) {
          if (e < E) {
            const float4 w0 = *reinterpret_cast<const float4*>(L->wg + static_cast<long long>(e) * D + c);
            const float4 w1 = *reinterpret_cast<const float4*>(L->wg + static_cast<long long>(e) * D + c + 4);
            acc[e] += xv[0] * w0.x + xv[1] * w0.y + xv[2] * w0.z + xv[3] * w0.w + xv[4] * w1.x + xv[5] * w1.y +
                      xv[6] * w1.z + xv[7] * w1.w;
          }
        }
      }
#pragma unroll
      for (int e = 0; e < DK_MAXE; ++e) acc[e] = warp_sum(acc[e]);
      if (lane == 0) {
        float mx = -INFINITY;
#pragma unroll
        for (int e = 0; e < DK_MAXE; ++e)
          if (e < E) mx = fmaxf(mx, acc[e]);
        float ex[DK_MAXE];
        float sum = 0.0f;
#pragma unroll
        for (int e = 0; e < DK_MAXE; ++e)
          if (e < E) {
            ex[e] = expf(acc[e] - mx);
            sum += ex[e];
          }
#pragma unroll
        for (int e = 0; e < DK_MAXE; ++e)
          if (e < E) {
            rt->logits[warp][e] = acc[e];
            rt->gates[warp][e] = ex[e] / sum;
          }
      }
    }
    consumer_sync();
    if (threadIdx.x == 0) {
      // top-1 + capacity slots in token order (torch.cumsum), as moe_scan_kernel / moe_route_small_kernel
      const bool moe = L->wg != nullptr;
      const int C = p.cap[E];
      int cnt[DK_MAXE];
      float me[DK_MAXE];
      for (int e = 0; e < E; ++e) cnt[e] = 0, me[e] = 0.0f, rt->kept[e] = 0;
      for (int s = 0; s < B; ++s) {
        int i1 = 0;
        float gsel = 1.0f;
        if (moe) {
          float best = rt->gates[s][0];
          for (int e = 0; e < E; ++e) {
            me[e] += rt->gates[s][e];
            if (rt->gates[s][e] > best) best = rt->gates[s][e], i1 = e;
          }
          gsel = best;
        }
        const int loc = cnt[i1]++;
        if (loc < C || !moe) {
          rt->tok_of_slot[i1][loc] = s;
          rt->gate_of_slot[i1][loc] = gsel;
          rt->kept[i1] = loc + 1;
        }
      }
      unsigned int am = 0;
      for (int e = 0; e < E; ++e)
        if (rt->kept[e] > 0) am |= 1u << e;
      rt->amask = am;
      rt->moe = moe ? 1 : 0;
      if (moe && blockIdx.x == 0) {
        float aux = 0.0f;
        for (int e = 0; e < E; ++e) {
          aux += (me[e] / B) * (static_cast<float>(cnt[e]) / B);
          if (p.exp_counts != nullptr) p.exp_counts[l * E + e] = cnt[e];
        }
        if (p.l_aux != nullptr) p.l_aux[l] = aux * E;
        if (p.gate_logits != nullptr)
          for (int s = 0; s < B; ++s)
            for (int e = 0; e < E; ++e) p.gate_logits[(static_cast<long long>(l) * B + s) * E + e] = rt->logits[s][e];
      }
      mbar_arrive(route_bar);  // release: the producer may read the expert choice
    }
    consumer_sync();
    const unsigned int amask = rt->amask;
    const int nact = __popc(amask);
    // ---------------------------------------------------------- P5: h1 = SiLU(h Wgate^T) * (h Wup^T) per active expert
    for (int tile = blockIdx.x; tile < nact * tiles_f; tile += G) {
      const int e = __fns(amask, 0, tile / tiles_f + 1);
      const int n0 = (tile % tiles_f) * 8;
      const int M = rt->kept[e];
      const __nv_bfloat16* arow =
          g < M ? reinterpret_cast<const __nv_bfloat16*>(s_a + rt->tok_of_slot[e][g] * pitch) : nullptr;
      float* rbuf = red + buf * DK_RED_FLOATS;
      buf ^= 1;
      consume_tile<true>(ring, chunks_d, D, arow, rbuf);
      if (threadIdx.x < 8 * DK_MAXB) {
        const int r = threadIdx.x & 7, m = threadIdx.x >> 3;
        if (m < M && n0 + r < F) {
          const float gte = bf16_round(reduce_rows(rbuf, r, m)), up = bf16_round(reduce_rows(rbuf, r + 8, m));
          const float v = bf16_round(gte / (1.0f + __expf(-gte))) * up;
          p.h1[(static_cast<long long>(e) * B + m) * F + n0 + r] = __float2bfloat16_rn(v);
        }
      }
    }
    grid_sync(p.sync, bar_target);
    // ---------------------------------------------------------- P6: x += gate * (h1 Wdown^T)
    for (int tile = blockIdx.x; tile < nact * tiles_d; tile += G) {
      const int e = __fns(amask, 0, tile / tiles_d + 1);
      const int n0 = (tile % tiles_d) * 16;
      const int M = rt->kept[e];
      const __nv_bfloat16* arow = g < M ? p.h1 + (static_cast<long long>(e) * B + g) * F : nullptr;
      float* rbuf = red + buf * DK_RED_FLOATS;
      buf ^= 1;
      consume_tile<false>(ring, chunks_f, F, arow, rbuf);
      if (threadIdx.x < 16 * DK_MAXB) {
        const int r = threadIdx.x & 15, m = threadIdx.x >> 4;
        if (m < M && n0 + r < D) {
          float v = reduce_rows(rbuf, r, m);
          if (rt->moe) v = bf16_round(v) * bf16_round(rt->gate_of_slot[e][m]);  // combine_weights.type_as(x)
          __nv_bfloat16* xp = p.x + static_cast<long long>(rt->tok_of_slot[e][m]) * D + n0 + r;
          *xp = __float2bfloat16_rn(bf16_round(v) + __bfloat162float(*xp));
        }
      }
    }
    grid_sync(p.sync, bar_target);
  }
  // ------------------------------------------------------------ final RMSNorm (hidden_states[-1])
  if (p.out_norm != nullptr && blockIdx.x == 0) {
    stage_rows(p, p.x, p.final_norm, s_a, pitch);
    consumer_sync();
    for (int i = threadIdx.x; i < B * (D / 8); i += DK_CONSUMERS * 32) {
      const int m = i / (D / 8), c = (i % (D / 8)) * 8;
      *reinterpret_cast<uint4*>(p.out_norm + static_cast<long long>(m) * D + c) =
          *reinterpret_cast<const uint4*>(s_a + m * pitch + c * 2);
    }
  }
  // reset the barrier counter for the next launch once every CTA is past its last wait
  if (threadIdx.x == 0) {
    if (atomicAdd(p.sync + 1, 1u) == static_cast<unsigned int>(G) - 1) {
      p.sync[0] = 0;
      p.sync[1] = 0;
      __threadfence();
    }
  }
}

// ------------------------------------------------------------------------------------------------ host side
static int tmap3(CUtensorMap* out, const void* W, int N, int K, int rows) {
  const unsigned long long dims[3] = {64ull, static_cast<unsigned long long>(N), static_cast<unsigned long long>(K / 64)};
  const unsigned long long strides[2] = {static_cast<unsigned long long>(K) * 2, 128ull};
  const unsigned box[3] = {64u, static_cast<unsigned>(rows), static_cast<unsigned>(DK_KC / 64)};
  return encode_tmap_bf16(out, W, 3, dims, strides, box);
}

static long long decode_smem_bytes(int D) {
  return 1024 + static_cast<long long>(DK_STAGES) * DK_STAGE_BYTES + static_cast<long long>(DK_MAXB) * (D * 2 + 64) +
         2LL * DK_RED_FLOATS * 4 + (2 * DK_STAGES + 2) * 8 + static_cast<long long>(sizeof(RouteSmem)) + 64;
}

bool llama_decode_supported(const mpl_llama_model& m, const mpl_llama_io& io) {
  if (io.decode_plan == nullptr || io.T != 1 || io.B < 1 || io.B > DK_MAXB) return false;
  if (io.hidden_states != nullptr || io.moe_noise != nullptr) return false;
  if (m.top_k != 1 || (m.hidden % 64) != 0 || (m.ffn % 64) != 0 || m.hidden != m.n_heads * 128) return false;
  if (io.attn_scratch == nullptr) return false;
  if (decode_smem_bytes(m.hidden) > 227 * 1024) return false;
  return true;
}

long long llama_decode_plan_bytes(const mpl_llama_model& m) {
  return DK_SYNC_BYTES + static_cast<long long>(m.n_layers) * static_cast<long long>(sizeof(DecLayerDev));
}

int llama_decode_plan_build(const mpl_llama_model& m, void* plan_dev, cudaStream_t st) {
  if (m.layers == nullptr || plan_dev == nullptr) return MPL_ERR_ARG;
  if ((m.hidden % 64) != 0 || (m.ffn % 64) != 0) return MPL_ERR_UNSUPPORTED;
  std::vector<DecLayerDev> host(m.n_layers);
  const int D = m.hidden, F = m.ffn;
  for (int l = 0; l < m.n_layers; ++l) {
    const mpl_llama_layer& L = m.layers[l];
    DecLayerDev& d = host[l];
    memset(&d, 0, sizeof(d));
    int rc = tmap3(&d.wq, L.wq, D, D, 16);
    if (rc == MPL_OK) rc = tmap3(&d.wk, L.wk, D, D, 16);
    if (rc == MPL_OK) rc = tmap3(&d.wv, L.wv, D, D, 16);
    if (rc == MPL_OK) rc = tmap3(&d.wo, L.wo, D, D, 16);
    const int E = L.wg != nullptr ? L.n_experts : 1;
    if (E > DK_MAXE) return MPL_ERR_UNSUPPORTED;
    for (int e = 0; e < E && rc == MPL_OK; ++e) {
      rc = tmap3(&d.wgate[e], L.w_gate[e], F, D, 8);
      if (rc == MPL_OK) rc = tmap3(&d.wup[e], L.w_up[e], F, D, 8);
      if (rc == MPL_OK) rc = tmap3(&d.wdown[e], L.w_down[e], D, F, 16);
    }
    if (rc != MPL_OK) return rc;
    d.input_ln = static_cast<const __nv_bfloat16*>(L.input_ln);
    d.post_ln = static_cast<const __nv_bfloat16*>(L.post_ln);
    d.wg = L.wg;
    d.n_experts = E;
  }
  char* base = static_cast<char*>(plan_dev);
  if (cudaMemsetAsync(base, 0, DK_SYNC_BYTES, st) != cudaSuccess) return MPL_ERR_CUDA;
  if (cudaMemcpyAsync(base + DK_SYNC_BYTES, host.data(), host.size() * sizeof(DecLayerDev), cudaMemcpyHostToDevice,
                      st) != cudaSuccess)
    return MPL_ERR_CUDA;
  if (cudaStreamSynchronize(st) != cudaSuccess) return MPL_ERR_CUDA;  // `host` goes out of scope
  return MPL_OK;
}

// ws: qkv [B,3D] | attn [B,D] | h1 [Emax*B, F]  (carved by the caller from the stack workspace)
int llama_decode_step(const mpl_llama_model& m, const mpl_llama_io& io, void* qkv, void* attn, void* h1,
                      const int* cap_by_e, int emax, cudaStream_t st) {
  const int D = m.hidden, H = m.n_heads;
  DecParams p;
  memset(&p, 0, sizeof(p));
  char* plan = static_cast<char*>(const_cast<void*>(io.decode_plan));
  p.sync = reinterpret_cast<unsigned int*>(plan);
  p.layers = reinterpret_cast<const DecLayerDev*>(plan + DK_SYNC_BYTES);
  p.x = static_cast<__nv_bfloat16*>(io.x);
  p.out_norm = static_cast<__nv_bfloat16*>(io.out_norm);
  p.final_norm = static_cast<const __nv_bfloat16*>(m.final_norm);
  p.qkv = static_cast<__nv_bfloat16*>(qkv);
  p.attn = static_cast<__nv_bfloat16*>(attn);
  p.h1 = static_cast<__nv_bfloat16*>(h1);
  p.kc = static_cast<__nv_bfloat16*>(io.k_cache);
  p.vc = static_cast<__nv_bfloat16*>(io.v_cache);
  p.cache_layer = static_cast<long long>(io.B) * H * io.Tmax * 128;
  p.cos_t = static_cast<const __nv_bfloat16*>(m.rope_cos);
  p.sin_t = static_cast<const __nv_bfloat16*>(m.rope_sin);
  p.kv_mask = io.kv_mask;
  p.kv_mask_stride = io.kv_mask_stride;
  p.pos_dev = io.pos_dev;
  p.gate_logits = io.gate_logits;
  p.l_aux = io.l_aux;
  p.exp_counts = io.exp_counts;
  p.B = io.B;
  p.D = D;
  p.H = H;
  p.F = m.ffn;
  p.L = m.n_layers;
  p.Tmax = io.Tmax;
  p.pos = io.past_len;
  p.Emax = emax;
  for (int e = 0; e <= DK_MAXE; ++e) p.cap[e] = cap_by_e[e];
  p.eps = m.rms_eps;
  p.scale = 1.0f / sqrtf(128.0f);
  const int G = num_sms();
  // split-K over the keys: aim at ~6 work items per CTA, at least 64 keys per split, within the scratch buffer
  const int bh = io.B * H;
  const int Tk = io.past_len + 1;
  int nsplit = (6 * G + bh - 1) / bh;
  const int by_keys = (Tk + 63) / 64;
  if (nsplit > by_keys) nsplit = by_keys;
  if (nsplit > 32) nsplit = 32;
  if (nsplit < 1) nsplit = 1;
  while (nsplit > 1 && static_cast<long long>(bh) * 4 + 256 + static_cast<long long>(bh) * nsplit * 130 * 4 > io.attn_scratch_bytes)
    --nsplit;
  p.nsplit = nsplit;
  p.attn_cnt = static_cast<int*>(io.attn_scratch);
  p.attn_part = reinterpret_cast<float*>(static_cast<char*>(io.attn_scratch) + ((static_cast<long long>(bh) * 4 + 255) & ~255LL));
  const int smem = static_cast<int>(decode_smem_bytes(D));
  static int attr_smem = 0;
  if (attr_smem < smem) {
    if (cudaFuncSetAttribute(llama_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess)
      return MPL_ERR_CUDA;
    attr_smem = smem;
  }
  void* args[] = {&p};
  if (cudaLaunchCooperativeKernel(reinterpret_cast<void*>(llama_decode_kernel), dim3(G), dim3(DK_THREADS), args, smem,
                                  st) != cudaSuccess) {
    cudaGetLastError();
    return MPL_ERR_CUDA;
  }
  return launch_status();
}

}  // namespace mpl
