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
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "internal.h"
#include "ptx.cuh"

namespace mpl {

constexpr int DK_CONSUMERS = 8;
constexpr int DK_THREADS = (DK_CONSUMERS + 2) * 32;  // 8 consumer warps + TMA producer warp + L2 prefetch warp
constexpr int DK_STAGES = 6;
constexpr int DK_STAGE_BYTES = 16384;  // [8 k-blocks][16 rows][128 B swizzled]  (dual: 2 x [8][8 rows][128 B])
constexpr int DK_KC = 512;
constexpr int DK_MAXB = 8;
constexpr int DK_MAXE = MPL_MAX_EXPERTS;
constexpr int DK_RP = 9;  // reduction row pitch (floats)
constexpr int DK_RED_FLOATS = DK_CONSUMERS * 16 * DK_RP;
constexpr int DK_WG_SMEM = 32768;  // router weights staged in shared memory when E*D*4 fits

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
  float* attn_part;     // split-K partials [B*H][nsplit][132]
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
  int B, D, H, F, L, Tmax, pos, nsplit, Emax, timing_layer;
  int kla, ela, spec;  // L2 prefetch look-ahead (16 KB chunks per CTA): known stream, expert stream; speculate all experts
  int cap[DK_MAXE + 1];  // capacity for a layer with E experts (index E)
  float eps, scale;
};

// Optional per-phase timestamps (dev tool): consumer thread 0 of every CTA records %globaltimer at the phase boundaries
// of layer `timing_layer`.
constexpr int DK_TSLOTS = 32;
__device__ unsigned long long g_dk_times[160 * DK_TSLOTS];
static int g_timing_layer = -1;
__device__ __forceinline__ unsigned long long globaltimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define DK_STAMP(i)                                                                              \
  do {                                                                                           \
    if (l == p.timing_layer && threadIdx.x == 0 && blockIdx.x < 160) g_dk_times[blockIdx.x * DK_TSLOTS + (i)] = globaltimer(); \
  } while (0)

#define DK_STAMP_L(layer_, i)                                                                           \
  do {                                                                                                   \
    if ((layer_) == p.timing_layer && threadIdx.x == 0 && blockIdx.x < 160) g_dk_times[blockIdx.x * DK_TSLOTS + (i)] = globaltimer(); \
  } while (0)

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
  int cnt[DK_MAXE];   // tokens whose first choice is e (before the capacity cut)
  float me[DK_MAXE];  // sum of the gate probabilities of e
  int kept[DK_MAXE];
  int tok_of_slot[DK_MAXE][DK_MAXB];
  float gate_of_slot[DK_MAXE][DK_MAXB];
  unsigned int amask;
  int moe;
  float logits[DK_MAXB][DK_MAXE];
  float gates[DK_MAXB][DK_MAXE];
  // L2 prefetcher coordination (monotonic counters, written by the producer / consumer thread 0, polled by the prefetcher)
  unsigned int pf_kn;        // chunks of the known stream (q,k,v,o of all layers) the producer has issued so far
  unsigned int pf_ex;        // (layer << 16) | chunks of that layer's expert stream issued so far
  unsigned int pf_route;     // 1 + last layer whose expert choice is published
  unsigned int pf_amask[4];  // expert masks of the last layers, slot l & 3
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
__device__ __forceinline__ unsigned int ld_relaxed_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// Grid barrier for the consumer warps (the producer thread never waits here). Same structure as cooperative
// groups' grid.sync(): CTA barrier, one thread fences + arrives + spins + fences (the gpu-scope fence also invalidates
// this SM's L1, so plain loads after the barrier see the other CTAs' writes), CTA barrier.
__device__ __forceinline__ void grid_sync(unsigned int* ctr, unsigned int& target) {
  consumer_sync();
  if (threadIdx.x == 0) {
    target += gridDim.x;
    // release (cumulative over the CTA's writes ordered by the bar.sync above) ... relaxed polling ... one acquire
    // fence, which also invalidates this SM's L1 (SASS: MEMBAR.ALL.GPU + RED / LDG.STRONG / MEMBAR + CCTL.IVALL)
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
    while (ld_relaxed_u32(ctr) < target) {
    }
    asm volatile("fence.acq_rel.gpu;" ::: "memory");
  }
  consumer_sync();
}

// ------------------------------------------------------------------------------------------------ producer side
__device__ __forceinline__ void tma_prefetch_3d(const void* tmap, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];" ::"l"(tmap), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void produce_tile(Ring& r, const CUtensorMap* m0, const CUtensorMap* m1, int n0, int chunks,
                                             volatile unsigned int* progress, unsigned int& count) {
  for (int c = 0; c < chunks; ++c) {
    mbar_wait(&r.empty[r.stage], r.phase ^ 1);
    uint8_t* dst = r.base + r.stage * DK_STAGE_BYTES;
    mbar_expect_tx(&r.full[r.stage], DK_STAGE_BYTES);
    tma_load_3d(dst, m0, &r.full[r.stage], 0, n0, c * (DK_KC / 64));
    if (m1 != nullptr) tma_load_3d(dst + DK_STAGE_BYTES / 2, m1, &r.full[r.stage], 0, n0, c * (DK_KC / 64));
    r.advance();
    *progress = ++count;
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
  // activation fragments run two chunks ahead of the weights (global-memory A of the down projection: L2 latency)
  auto load_a = [&](int c, uint4& x0, uint4& x1) {
    const int k = c * DK_KC + warp * 64 + t * 8;
    x0 = (arow != nullptr && c < chunks && k < K) ? *reinterpret_cast<const uint4*>(arow + k) : zero;
    x1 = (arow != nullptr && c < chunks && k + 32 < K) ? *reinterpret_cast<const uint4*>(arow + k + 32) : zero;
  };
  uint4 xa0, xa1, xn0, xn1;
  load_a(0, xa0, xa1);
  load_a(1, xn0, xn1);
  for (int c = 0; c < chunks; ++c) {
    const uint4 xb0 = xa0, xb1 = xa1;
    xa0 = xn0;
    xa1 = xn1;
    load_a(c + 2, xn0, xn1);
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

// Activation staging: the B rows go global (L2) -> registers -> shared memory with EVERY load of a thread issued before
// its first store (thread t owns the 16-byte vectors t and t + 256 of every row: one L2 round trip for the whole block,
// and not queued behind the weight ring in the TMA unit as a bulk copy would be). With a norm weight (already in shared
// memory, completion on ln_bar / ln_phase) the rows are RMS-normalised on the way with HF LlamaRMSNorm's roundings
// w * bf16(x * rstd); the sum of squares is reduced lane -> warp (shuffles) -> CTA (fixed order 0..7).
struct ActStage {
  uint8_t* s_a;  // [B][pitch]
  int pitch;
};
__device__ __forceinline__ void dk_bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __noinline__ void stage_rows(ActStage st, const __nv_bfloat16* __restrict__ src, const uint8_t* s_ln,
                                        uint64_t* ln_bar, uint32_t ln_phase, int B, int D, float eps, float* s_part,
                                        unsigned long long* ts = nullptr) {
  constexpr int NV = 2;  // vectors per thread and row: D <= 4096
  constexpr int RG = 4;  // rows per register group (spills are ruinous here: shared memory leaves almost no L1)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* const s_a = st.s_a;
  const int pitch = st.pitch;
  const uint4 zero = make_uint4(0, 0, 0, 0);
  if (s_ln != nullptr) mbar_wait(ln_bar, ln_phase);
  if (ts != nullptr) ts[0] = globaltimer();
#pragma unroll 1
  for (int m0 = 0; m0 < B; m0 += RG) {
    uint4 v[RG][NV];
#pragma unroll
    for (int r = 0; r < RG; ++r)
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const int k = (threadIdx.x + j * DK_CONSUMERS * 32) * 8;
        v[r][j] = (m0 + r < B && k < D) ? *reinterpret_cast<const uint4*>(src + static_cast<long long>(m0 + r) * D + k)
                                        : zero;
      }
    if (s_ln != nullptr) {
      float ss[RG];
#pragma unroll
      for (int r = 0; r < RG; ++r) {
        ss[r] = 0.0f;
#pragma unroll
        for (int j = 0; j < NV; ++j) {
          const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v[r][j]);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float2 f = __bfloat1622float2(h[i]);
            ss[r] += f.x * f.x + f.y * f.y;
          }
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1)
#pragma unroll
        for (int r = 0; r < RG; ++r) ss[r] += __shfl_xor_sync(0xffffffffu, ss[r], o);  // independent trees
      if (lane == 0) {
#pragma unroll
        for (int r = 0; r < RG; ++r) s_part[r * DK_CONSUMERS + warp] = ss[r];
      }
      if (ts != nullptr && m0 == 0) ts[1] = globaltimer();
      consumer_sync();
      if (ts != nullptr && m0 == 0) ts[2] = globaltimer();
#pragma unroll
      for (int r = 0; r < RG; ++r) {
        float tot = 0.0f;
#pragma unroll
        for (int wi = 0; wi < DK_CONSUMERS; ++wi) tot += s_part[r * DK_CONSUMERS + wi];
        const float rstd = rsqrtf(tot / static_cast<float>(D) + eps);
#pragma unroll
        for (int j = 0; j < NV; ++j) {
          const int k = (threadIdx.x + j * DK_CONSUMERS * 32) * 8;
          const uint4 w = k < D ? *reinterpret_cast<const uint4*>(s_ln + k * 2) : zero;
          const __nv_bfloat162* xp = reinterpret_cast<const __nv_bfloat162*>(&v[r][j]);
          const __nv_bfloat162* wp = reinterpret_cast<const __nv_bfloat162*>(&w);
          uint4 o;
          uint32_t* op = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float2 xf = __bfloat1622float2(xp[i]), wf = __bfloat1622float2(wp[i]);
            op[i] = pack_bf16(wf.x * bf16_round(xf.x * rstd), wf.y * bf16_round(xf.y * rstd));
          }
          v[r][j] = o;
        }
      }
    }
#pragma unroll
    for (int r = 0; r < RG; ++r)
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const int k = (threadIdx.x + j * DK_CONSUMERS * 32) * 8;
        if (m0 + r < B && k < D) *reinterpret_cast<uint4*>(s_a + (m0 + r) * pitch + k * 2) = v[r][j];
      }
    if (ts != nullptr && m0 == 0) ts[3] = globaltimer();
    consumer_sync();  // rows complete for every reader; s_part reusable
  }
  if (ts != nullptr) ts[4] = globaltimer();
}

// L2 prefetch of small per-layer tensors (router weights, norm weights) well before they are needed.
__device__ __forceinline__ void prefetch_l2(const void* ptr, long long bytes) {
  const char* c = static_cast<const char*>(ptr);
  for (long long off = static_cast<long long>(threadIdx.x) * 128; off < bytes; off += DK_CONSUMERS * 32 * 128)
    asm volatile("prefetch.global.L2 [%0];" ::"l"(c + off));
}

// ------------------------------------------------------------------------------------------------ attention item
// One (b, h, split) of the decode attention: 8 warps, 4 keys per warp step, 8 lanes x 16 dims per key. q and the new
// key are rotated on the fly from the q,k,v buffer; the split that owns position `pos` appends k,v to the cache.
__device__ __forceinline__ void unpack16(const uint4& a, const uint4& b, float (&f)[16]) {
  const __nv_bfloat162* ha = reinterpret_cast<const __nv_bfloat162*>(&a);
  const __nv_bfloat162* hb = reinterpret_cast<const __nv_bfloat162*>(&b);
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float2 x = __bfloat1622float2(ha[e]), y = __bfloat1622float2(hb[e]);
    f[2 * e] = x.x;
    f[2 * e + 1] = x.y;
    f[8 + 2 * e] = y.x;
    f[8 + 2 * e + 1] = y.y;
  }
}
// RoPE of dims d = gl*16 .. +15 of a 128-wide head (partner = d + 64 for the first half, rotated with a minus sign, or
// d - 64), HF apply_rotary_pos_emb in bf16: every product and the sum rounded. Result packed as bf16.
__device__ __forceinline__ void rope16(const __nv_bfloat16* src, const __nv_bfloat16* cr, const __nv_bfloat16* sr,
                                       int gl, uint4& o0, uint4& o1) {
  const int d0 = gl * 16;
  const bool first = gl < 4;
  const uint4* own = reinterpret_cast<const uint4*>(src + d0);
  const uint4* par = reinterpret_cast<const uint4*>(src + (first ? d0 + 64 : d0 - 64));
  const uint4* cp = reinterpret_cast<const uint4*>(cr + d0);
  const uint4* sp = reinterpret_cast<const uint4*>(sr + d0);
  const uint4 a0 = own[0], a1 = own[1], p0 = par[0], p1 = par[1], c0 = cp[0], c1 = cp[1], s0 = sp[0], s1 = sp[1];
  float f[16], fp[16], c[16], sn[16], out[16];
  unpack16(a0, a1, f);
  unpack16(p0, p1, fp);
  unpack16(c0, c1, c);
  unpack16(s0, s1, sn);
#pragma unroll
  for (int e = 0; e < 16; ++e) {
    const float a = bf16_round(f[e] * c[e]);
    const float b = bf16_round((first ? -fp[e] : fp[e]) * sn[e]);
    out[e] = a + b;
  }
  o0.x = pack_bf16(out[0], out[1]); o0.y = pack_bf16(out[2], out[3]); o0.z = pack_bf16(out[4], out[5]); o0.w = pack_bf16(out[6], out[7]);
  o1.x = pack_bf16(out[8], out[9]); o1.y = pack_bf16(out[10], out[11]); o1.z = pack_bf16(out[12], out[13]); o1.w = pack_bf16(out[14], out[15]);
}

// P2. One (b, h, split) per WARP: 4 keys per step (8 lanes x 16 dims per key), 4 steps (16 keys, 8 KB) in flight per
// warp, no CTA-level synchronisation: the 4 key groups merge by shuffles, splits merge through global scratch by the
// last warp to arrive (atomic counter per (b, h), self-cleaning). The cached keys do not depend on this layer's q,k,v,
// so the first 16 keys of the item are requested BEFORE the grid barrier that ends P1 and arrive while it is waiting.
__device__ __noinline__ void attention_phase(const DecParams& p, int layer, int Tk, unsigned int& bar_target) {
  constexpr int D = 128;
  constexpr int UNR = 2;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int grp = lane >> 3, gl = lane & 7;
  const int nsplit = p.nsplit;
  const int per = ((Tk + nsplit - 1) / nsplit + 31) & ~31;
  const int pos = Tk - 1;
  const int n_items = p.B * p.H * nsplit;
  const int stride = gridDim.x * DK_CONSUMERS;
  const __nv_bfloat16* cr = p.cos_t + static_cast<long long>(pos) * D;
  const __nv_bfloat16* sr = p.sin_t + static_cast<long long>(pos) * D;
  const float sl2 = p.scale * 1.4426950408889634f;
  bool synced = false;
  // items are dealt round-robin over the CTAs first, so every SM pulls on the KV cache
  for (int item = warp * gridDim.x + blockIdx.x; item < n_items || !synced; item += stride) {
    const bool has = item < n_items;
    const int z = has ? item % nsplit : 0, bh = has ? item / nsplit : 0;
    const int b = bh / p.H, h = bh % p.H;
    const int k_lo = z * per;
    const int k_hi = has ? min(Tk, k_lo + per) : k_lo;
    const __nv_bfloat16* qrow = p.qkv + static_cast<long long>(b) * 3 * p.D + h * D;
    const __nv_bfloat16* krow = qrow + p.D;
    const __nv_bfloat16* vrow = qrow + 2 * p.D;
    const long long head_off = (static_cast<long long>(b) * p.H + h) * p.Tmax * D;
    __nv_bfloat16* kc = p.kc + layer * p.cache_layer + head_off;
    __nv_bfloat16* vc = p.vc + layer * p.cache_layer + head_off;
    const unsigned char* mrow = p.kv_mask ? p.kv_mask + static_cast<long long>(b) * p.kv_mask_stride : nullptr;

    // cached K/V of one key for this lane's 16 dims (packed bf16); the new position is filled in later
    auto load_cached = [&](int k0, uint4& ka, uint4& kb, uint4& va, uint4& vb) {
      const int key = k0 + grp;
      const int kk = (key < k_hi && key != pos) ? key : k_lo;
      if (kk == pos) {  // (k_lo == pos: nothing cached to read)
        ka = kb = va = vb = make_uint4(0, 0, 0, 0);
        return;
      }
      const __nv_bfloat16* kr = kc + static_cast<long long>(kk) * D + gl * 16;
      const __nv_bfloat16* vr = vc + static_cast<long long>(kk) * D + gl * 16;
      ka = *reinterpret_cast<const uint4*>(kr);
      kb = *reinterpret_cast<const uint4*>(kr + 8);
      va = *reinterpret_cast<const uint4*>(vr);
      vb = *reinterpret_cast<const uint4*>(vr + 8);
    };
    // the new position: rotate k from the q,k,v buffer, append k,v to the cache (one lane group of one split owns it)
    auto fix_new = [&](int k0, uint4& ka, uint4& kb, uint4& va, uint4& vb) {
      if (k0 + grp != pos || pos >= k_hi) return;
      rope16(krow, cr, sr, gl, ka, kb);
      va = *reinterpret_cast<const uint4*>(vrow + gl * 16);
      vb = *reinterpret_cast<const uint4*>(vrow + gl * 16 + 8);
      __nv_bfloat16* kd = kc + static_cast<long long>(pos) * D + gl * 16;
      *reinterpret_cast<uint4*>(kd) = ka;
      *reinterpret_cast<uint4*>(kd + 8) = kb;
      __nv_bfloat16* vd = vc + static_cast<long long>(pos) * D + gl * 16;
      *reinterpret_cast<uint4*>(vd) = va;
      *reinterpret_cast<uint4*>(vd + 8) = vb;
    };

    // two register buffers of UNR steps (8 keys) each, ping-pong: one is consumed while the other is in flight
    uint4 ka[2][UNR], kb[2][UNR], va[2][UNR], vb[2][UNR];
    auto load_round = [&](int kb0, int w) {
#pragma unroll
      for (int u = 0; u < UNR; ++u)
        if (kb0 + 4 * u < k_hi) load_cached(kb0 + 4 * u, ka[w][u], kb[w][u], va[w][u], vb[w][u]);
    };
    load_round(k_lo, 0);
    load_round(k_lo + 4 * UNR, 1);
    if (!synced) {
      DK_STAMP_L(layer, 2);
      grid_sync(p.sync, bar_target);  // q,k,v of this layer are complete; the first keys are already in flight
      DK_STAMP_L(layer, 3);
      synced = true;
    }
    if (!has) break;
    float qf[16];
    {
      uint4 q0, q1;
      rope16(qrow, cr, sr, gl, q0, q1);
      unpack16(q0, q1, qf);
    }
    float m = -INFINITY, l = 0.0f;
    float acc[16];
#pragma unroll
    for (int e = 0; e < 16; ++e) acc[e] = 0.0f;
    auto compute_round = [&](int kb0, int w) {
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        if (kb0 + 4 * u >= k_hi) break;  // warp-uniform
        fix_new(kb0 + 4 * u, ka[w][u], kb[w][u], va[w][u], vb[w][u]);
        const int key = kb0 + 4 * u + grp;
        const bool valid = key < k_hi;
        const int kk = valid ? key : k_lo;
        float kf[16], vf[16];
        unpack16(ka[w][u], kb[w][u], kf);
        unpack16(va[w][u], vb[w][u], vf);
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
    };
    for (int kb0 = k_lo; kb0 < k_hi; kb0 += 8 * UNR) {
      compute_round(kb0, 0);
      load_round(kb0 + 8 * UNR, 0);
      if (kb0 + 4 * UNR < k_hi) {
        compute_round(kb0 + 4 * UNR, 1);
        load_round(kb0 + 12 * UNR, 1);
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
  __nv_bfloat16* optr = p.attn + static_cast<long long>(b) * p.D + h * D;
  if (nsplit == 1) {
    if (grp == 0) {
      const float inv = l > 0.0f ? 1.0f / l : 0.0f;
      uint4 o0, o1;
      o0.x = pack_bf16(acc[0] * inv, acc[1] * inv); o0.y = pack_bf16(acc[2] * inv, acc[3] * inv);
      o0.z = pack_bf16(acc[4] * inv, acc[5] * inv); o0.w = pack_bf16(acc[6] * inv, acc[7] * inv);
      o1.x = pack_bf16(acc[8] * inv, acc[9] * inv); o1.y = pack_bf16(acc[10] * inv, acc[11] * inv);
      o1.z = pack_bf16(acc[12] * inv, acc[13] * inv); o1.w = pack_bf16(acc[14] * inv, acc[15] * inv);
      *reinterpret_cast<uint4*>(optr + gl * 16) = o0;
      *reinterpret_cast<uint4*>(optr + gl * 16 + 8) = o1;
    }
    continue;
  }

  float* part = p.attn_part + static_cast<long long>(bh) * nsplit * (D + 4);
  if (grp == 0) {
    float* mine = part + z * (D + 4);
#pragma unroll
    for (int q4 = 0; q4 < 4; ++q4)
      *reinterpret_cast<float4*>(mine + gl * 16 + q4 * 4) =
          make_float4(acc[q4 * 4], acc[q4 * 4 + 1], acc[q4 * 4 + 2], acc[q4 * 4 + 3]);
    if (gl == 0) {
      mine[D] = m;
      mine[D + 1] = l;
    }
  }
  __threadfence();
  __syncwarp();
  int last = 0;
  if (lane == 0) last = (atomicAdd(&p.attn_cnt[bh], 1) == nsplit - 1);
  last = __shfl_sync(0xffffffffu, last, 0);
  if (!last) continue;
  __threadfence();
  // this warp merges all splits. Lane zz first fetches (max, sum) of split zz -- one round trip for all splits --
  // then every lane accumulates its 4 dims over the splits with the loads of 8 splits in flight at a time.
  const float mz = lane < nsplit ? __ldcg(part + lane * (D + 4) + D) : -INFINITY;
  const float lz = lane < nsplit ? __ldcg(part + lane * (D + 4) + D + 1) : 0.0f;
  const float gm = warp_max(mz);
  const float gsafe = (gm == -INFINITY) ? 0.0f : gm;
  const float cz = exp2f(mz - gsafe);  // 0 for absent / empty splits
  const float gl_ = warp_sum(lz * cz);
  float go[4] = {0.f, 0.f, 0.f, 0.f};
  for (int z0 = 0; z0 < nsplit; z0 += 8) {
    float4 o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j)
      o[j] = z0 + j < nsplit ? __ldcg(reinterpret_cast<const float4*>(part + (z0 + j) * (D + 4) + lane * 4))
                             : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float c = __shfl_sync(0xffffffffu, cz, (z0 + j) & 31);
      if (z0 + j < nsplit) {
        go[0] += o[j].x * c;
        go[1] += o[j].y * c;
        go[2] += o[j].z * c;
        go[3] += o[j].w * c;
      }
    }
  }
  const float inv = gl_ > 0.0f ? 1.0f / gl_ : 0.0f;
  uint2 o;
  o.x = pack_bf16(go[0] * inv, go[1] * inv);
  o.y = pack_bf16(go[2] * inv, go[3] * inv);
  *reinterpret_cast<uint2*>(optr + lane * 4) = o;
  if (lane == 0) p.attn_cnt[bh] = 0;  // self-cleaning for the next layer / launch
  }
}

// ------------------------------------------------------------------------------------------------ the kernel
__global__ void __launch_bounds__(DK_THREADS, 1) llama_decode_kernel(const __grid_constant__ DecParams p) {
  extern __shared__ uint8_t dk_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(dk_raw) + 1023) & ~uintptr_t(1023));
  const int pitch = p.D * 2 + 64;  // activation row pitch in shared memory (conflict-free 16-byte reads)
  uint8_t* s_a = smem + DK_STAGES * DK_STAGE_BYTES;
  float* red = reinterpret_cast<float*>(s_a + p.B * pitch);  // [2][8][16][9]; staging / router scratch between phases
  // (shared memory is sized by the actual B: what it does not take stays L1, which the few spilled values need)
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(red + 2 * DK_RED_FLOATS);
  uint64_t* empty_bar = full_bar + DK_STAGES;
  uint64_t* route_bar = empty_bar + DK_STAGES;
  uint64_t* act_bar = route_bar + 1;
  uint64_t* wg_bar = route_bar + 2;  // post-attention norm weight + router weights of the layer
  uint64_t* lnin_bar = route_bar + 3;  // input norm weight of the next layer (or the final norm)
  RouteSmem* rt = reinterpret_cast<RouteSmem*>(route_bar + 4);
  uint8_t* s_ln_in = reinterpret_cast<uint8_t*>(rt) + ((sizeof(RouteSmem) + 127) & ~size_t(127));  // [D] bf16
  uint8_t* s_ln_post = s_ln_in + p.D * 2;                                                           // [D] bf16
  float* s_wg = reinterpret_cast<float*>(s_ln_post + p.D * 2);  // [E][D] f32 router weights (when they fit)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int G = gridDim.x;
  if (threadIdx.x == 0) {
    for (int s = 0; s < DK_STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], DK_CONSUMERS);
    }
    mbar_init(route_bar, 1);
    mbar_init(act_bar, 1);
    mbar_init(wg_bar, 1);
    mbar_init(lnin_bar, 1);
    fence_mbar_init();
    rt->pf_kn = rt->pf_ex = rt->pf_route = 0;
    rt->pf_amask[0] = rt->pf_amask[1] = rt->pf_amask[2] = rt->pf_amask[3] = 1u;
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
    unsigned int kn = 0;
    volatile unsigned int* pf_kn = &rt->pf_kn;
    volatile unsigned int* pf_ex = &rt->pf_ex;
    for (int l = 0; l < p.L; ++l) {
      const DecLayerDev* L = p.layers + l;
      for (int tile = blockIdx.x; tile < 3 * tiles_d; tile += G) {
        const int which = tile / tiles_d;
        const CUtensorMap* tm = which == 0 ? &L->wq : (which == 1 ? &L->wk : &L->wv);
        produce_tile(ring, tm, nullptr, (tile % tiles_d) * 16, chunks_d, pf_kn, kn);
      }
      for (int tile = blockIdx.x; tile < tiles_d; tile += G)
        produce_tile(ring, &L->wo, nullptr, tile * 16, chunks_d, pf_kn, kn);
      mbar_wait(route_bar, route_phase);  // expert choice of this layer
      route_phase ^= 1;
      const unsigned int amask = rt->amask;
      const int nact = __popc(amask);
      unsigned int ex = static_cast<unsigned int>(l) << 16;
      *pf_ex = ex;
      for (int tile = blockIdx.x; tile < nact * tiles_f; tile += G) {
        const int e = __fns(amask, 0, tile / tiles_f + 1);
        produce_tile(ring, &L->wgate[e], &L->wup[e], (tile % tiles_f) * 8, chunks_d, pf_ex, ex);
      }
      for (int tile = blockIdx.x; tile < nact * tiles_d; tile += G) {
        const int e = __fns(amask, 0, tile / tiles_d + 1);
        produce_tile(ring, &L->wdown[e], nullptr, (tile % tiles_d) * 16, chunks_f, pf_ex, ex);
      }
    }
    return;
  }
  if (warp == DK_CONSUMERS + 1) {
    // ============================================================ L2 prefetcher
    // Walks the same per-CTA tile schedule as the producer, a bounded number of 16 KB chunks AHEAD of it, with
    // cp.async.bulk.prefetch.tensor (HBM -> L2, fire and forget): while the consumers sit in a grid barrier, the
    // attention phase, activation staging or the router -- when the 6-stage ring is full and the producer is blocked --
    // HBM keeps streaming into L2, and the ring refills from L2 afterwards. Two independent cursors:
    //   known stream   q,k,v,o tiles of ALL layers (never depends on data): up to p.kla chunks ahead, across layers;
    //   expert stream  gate/up/down tiles of one layer: needs that layer's expert choice (or, with p.spec, assumes every
    //                  expert is hit -- B >= 4 -- and starts early): up to p.ela chunks ahead.
    // Every wait is a poll of a monotonic shared-memory counter that the producer / consumers advance no matter what
    // this thread does, a cursor that has fallen behind jumps forward, so the thread cannot block anyone and always exits.
    if (lane != 0 || (p.kla <= 0 && p.ela <= 0)) return;
    const int bid = blockIdx.x;
    auto cnt = [&](int total) { return total > bid ? (total - bid + G - 1) / G : 0; };
    const int n1 = cnt(3 * tiles_d), n3 = cnt(tiles_d);
    const unsigned int per_layer = static_cast<unsigned int>((n1 + n3) * chunks_d);
    const unsigned int total_known = p.kla > 0 ? per_layer * static_cast<unsigned int>(p.L) : 0u;
    const unsigned int kla = static_cast<unsigned int>(p.kla), ela = static_cast<unsigned int>(p.ela);
    volatile RouteSmem* vr = rt;
    unsigned int kn_pf = 0, ex_pf = 0;
    int le = p.ela > 0 ? 0 : p.L;
    while (kn_pf < total_known || le < p.L) {
      bool progress = false;
      if (le < p.L) {
        const bool known = static_cast<int>(vr->pf_route) > le;
        if (known || p.spec) {
          const DecLayerDev* Lp = p.layers + le;
          const int E = Lp->wg != nullptr ? Lp->n_experts : 1;
          unsigned int amask = known ? vr->pf_amask[le & 3] : ((1u << E) - 1u);
          amask &= (1u << E) - 1u;
          if (amask == 0) amask = 1u;
          const int nact = __popc(amask);
          const int n5 = cnt(nact * tiles_f), n6 = cnt(nact * tiles_d);
          const unsigned int c5 = static_cast<unsigned int>(n5 * chunks_d);
          const unsigned int total_e = c5 + static_cast<unsigned int>(n6 * chunks_f);
          const unsigned int pos = vr->pf_ex;
          const int ll = static_cast<int>(pos >> 16);
          if (ll > le) {  // the producer is already past this layer
            ++le;
            ex_pf = 0;
            progress = true;
          } else {
            const unsigned int base = ll == le ? (pos & 0xffffu) : 0u;
            if (ex_pf < base) ex_pf = base;
            if (ex_pf < total_e && ex_pf < base + ela) {
              if (ex_pf < c5) {
                const int tile = bid + static_cast<int>(ex_pf / chunks_d) * G, c = static_cast<int>(ex_pf % chunks_d);
                const int e = __fns(amask, 0, tile / tiles_f + 1) & (DK_MAXE - 1);
                tma_prefetch_3d(&Lp->wgate[e], 0, (tile % tiles_f) * 8, c * (DK_KC / 64));
                tma_prefetch_3d(&Lp->wup[e], 0, (tile % tiles_f) * 8, c * (DK_KC / 64));
              } else {
                const unsigned int r = ex_pf - c5;
                const int tile = bid + static_cast<int>(r / chunks_f) * G, c = static_cast<int>(r % chunks_f);
                const int e = __fns(amask, 0, tile / tiles_d + 1) & (DK_MAXE - 1);
                tma_prefetch_3d(&Lp->wdown[e], 0, (tile % tiles_d) * 16, c * (DK_KC / 64));
              }
              ++ex_pf;
              progress = true;
            }
            if (ex_pf >= total_e) {
              ++le;
              ex_pf = 0;
              progress = true;
            }
          }
        }
      }
      if (kn_pf < total_known) {
        const unsigned int kl = vr->pf_kn;
        if (kn_pf < kl) kn_pf = kl;
        if (kn_pf < total_known && kn_pf < kl + kla) {
          const DecLayerDev* Lp = p.layers + kn_pf / per_layer;
          const unsigned int r = kn_pf % per_layer;
          const int ti = static_cast<int>(r / chunks_d), c = static_cast<int>(r % chunks_d);
          if (ti < n1) {
            const int tile = bid + ti * G, which = tile / tiles_d;
            const CUtensorMap* tm = which == 0 ? &Lp->wq : (which == 1 ? &Lp->wk : &Lp->wv);
            tma_prefetch_3d(tm, 0, (tile % tiles_d) * 16, c * (DK_KC / 64));
          } else {
            tma_prefetch_3d(&Lp->wo, 0, (bid + (ti - n1) * G) * 16, c * (DK_KC / 64));
          }
          ++kn_pf;
          progress = true;
        }
      }
      if (!progress) __nanosleep(100);
    }
    return;
  }

  // ============================================================== consumers
  const int g = lane >> 2;
  unsigned int bar_target = 0;
  int buf = 0;
  const ActStage act{s_a, pitch};
  uint32_t wg_phase = 0, lnin_phase = 0;
  if (threadIdx.x == 0) {  // input norm weight of layer 0
    mbar_expect_tx(lnin_bar, static_cast<uint32_t>(D * 2));
    dk_bulk_g2s(s_ln_in, p.layers[0].input_ln, static_cast<uint32_t>(D * 2), lnin_bar);
  }
  const int pos = p.pos_dev ? *p.pos_dev : p.pos;
  const int Tk = pos + 1;
  for (int l = 0; l < p.L; ++l) {
    const DecLayerDev* L = p.layers + l;
    DK_STAMP(0);
    const int E = L->wg != nullptr ? L->n_experts : 1;
    const bool wg_smem = L->wg != nullptr && static_cast<long long>(E) * D * 4 <= DK_WG_SMEM;
    if (threadIdx.x == 0) {  // small weights of this layer: in shared memory long before P4 needs them
      fence_proxy_async();
      mbar_expect_tx(wg_bar, static_cast<uint32_t>(D * 2 + (wg_smem ? E * D * 4 : 0)));
      dk_bulk_g2s(s_ln_post, L->post_ln, static_cast<uint32_t>(D * 2), wg_bar);
      if (wg_smem)
        for (int e = 0; e < E; ++e) dk_bulk_g2s(s_wg + e * D, L->wg + static_cast<long long>(e) * D, D * 4, wg_bar);
    }
    // ---------------------------------------------------------- P1: q,k,v = RMSNorm(x) Wqkv^T
    unsigned long long* const tsp =
        (l == p.timing_layer && threadIdx.x == 0 && blockIdx.x < 160) ? g_dk_times + blockIdx.x * DK_TSLOTS : nullptr;
    DK_STAMP(14);
    stage_rows(act, p.x, s_ln_in, lnin_bar, lnin_phase, B, D, p.eps, red, tsp != nullptr ? tsp + 15 : nullptr);
    lnin_phase ^= 1;
    consumer_sync();
    DK_STAMP(1);
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
    // ---------------------------------------------------------- P2 (includes the grid barrier that ends P1)
    attention_phase(p, l, Tk, bar_target);
    DK_STAMP(4);
    grid_sync(p.sync, bar_target);
    DK_STAMP(5);
    // ---------------------------------------------------------- P3: x += attn Wo^T
    stage_rows(act, p.attn, nullptr, nullptr, 0u, B, D, p.eps, red);
    if (threadIdx.x == 0) {  // s_ln_in is free (all CTA threads are past P1's staging): next layer's input norm weight
      fence_proxy_async();
      mbar_expect_tx(lnin_bar, static_cast<uint32_t>(D * 2));
      dk_bulk_g2s(s_ln_in, l + 1 < p.L ? p.layers[l + 1].input_ln : p.final_norm, static_cast<uint32_t>(D * 2), lnin_bar);
    }
    consumer_sync();
    DK_STAMP(6);
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
    DK_STAMP(7);
    grid_sync(p.sync, bar_target);
    DK_STAMP(8);
    // ---------------------------------------------------------- P4: h = RMSNorm(x), router (every CTA, no barrier)
    stage_rows(act, p.x, s_ln_post, wg_bar, wg_phase, B, D, p.eps, red,
               tsp != nullptr ? tsp + 20 : nullptr);  // (wg_bar also covers the router weights)
    wg_phase ^= 1;
    DK_STAMP(25);
    if (L->wg != nullptr) {
      // router logits from the stored (bf16-rounded) h, fp32 like DeepSpeed's TopKGate; all 8 warps share every row
      // (thread t owns k = 8t, 8t + 2048, ...), partial sums reduced lane -> warp -> CTA in a fixed order
      const float* wgp = wg_smem ? s_wg : L->wg;  // generic pointer: shared-memory copy when it fits
      float* s_rp = red + 64;                     // [B][DK_MAXE][8 warps]
#pragma unroll 1
      for (int e0 = 0; e0 < E; e0 += 2)
#pragma unroll 1
      for (int m0 = 0; m0 < B; m0 += 4) {
        float acc[4][2];
#pragma unroll
        for (int m = 0; m < 4; ++m) acc[m][0] = acc[m][1] = 0.0f;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int c = (threadIdx.x + j * DK_CONSUMERS * 32) * 8;
          if (c < D) {
            float4 wv[2][2];
#pragma unroll
            for (int ee = 0; ee < 2; ++ee) {
              const int e = min(e0 + ee, E - 1);
              wv[ee][0] = *reinterpret_cast<const float4*>(wgp + static_cast<long long>(e) * D + c);
              wv[ee][1] = *reinterpret_cast<const float4*>(wgp + static_cast<long long>(e) * D + c + 4);
            }
#pragma unroll
            for (int m = 0; m < 4; ++m) {
              if (m0 + m < B) {
                const uint4 raw = *reinterpret_cast<const uint4*>(s_a + (m0 + m) * pitch + c * 2);
                const __nv_bfloat162* hp = reinterpret_cast<const __nv_bfloat162*>(&raw);
                float xv[8];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const float2 f = __bfloat1622float2(hp[i]);
                  xv[2 * i] = f.x;
                  xv[2 * i + 1] = f.y;
                }
#pragma unroll
                for (int ee = 0; ee < 2; ++ee) {
                  const float4 w0 = wv[ee][0], w1 = wv[ee][1];
                  acc[m][ee] += xv[0] * w0.x + xv[1] * w0.y + xv[2] * w0.z + xv[3] * w0.w + xv[4] * w1.x +
                                xv[5] * w1.y + xv[6] * w1.z + xv[7] * w1.w;
                }
              }
            }
          }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
#pragma unroll
          for (int m = 0; m < 4; ++m) {
            acc[m][0] += __shfl_xor_sync(0xffffffffu, acc[m][0], o);
            acc[m][1] += __shfl_xor_sync(0xffffffffu, acc[m][1], o);
          }
        if (lane == 0) {
#pragma unroll
          for (int m = 0; m < 4; ++m) {
            if (m0 + m < B) {
              s_rp[((m0 + m) * DK_MAXE + e0) * DK_CONSUMERS + warp] = acc[m][0];
              if (e0 + 1 < E) s_rp[((m0 + m) * DK_MAXE + e0 + 1) * DK_CONSUMERS + warp] = acc[m][1];
            }
          }
        }
      }
      consumer_sync();
      if (threadIdx.x < B * E) {
        const int m = threadIdx.x / E, e = threadIdx.x % E;
        float t = 0.0f;
#pragma unroll
        for (int w = 0; w < DK_CONSUMERS; ++w) t += s_rp[(m * DK_MAXE + e) * DK_CONSUMERS + w];
        rt->logits[m][e] = t;
      }
      consumer_sync();
      if (threadIdx.x < B) {
        const int m = threadIdx.x;
        float mx = -INFINITY;
        for (int e = 0; e < E; ++e) mx = fmaxf(mx, rt->logits[m][e]);
        float sum = 0.0f;
        for (int e = 0; e < E; ++e) sum += expf(rt->logits[m][e] - mx);
        for (int e = 0; e < E; ++e) rt->gates[m][e] = expf(rt->logits[m][e] - mx) / sum;
      }
    }
    consumer_sync();
    DK_STAMP(26);
    if (threadIdx.x == 0) {
      // top-1 + capacity slots in token order (torch.cumsum), as moe_scan_kernel / moe_route_small_kernel
      const bool moe = L->wg != nullptr;
      const int C = p.cap[E];
      // (counters live in shared memory: a dynamically indexed local array would sit in local memory, i.e. L2)
      int* cnt = rt->cnt;
      float* me = rt->me;
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
      rt->pf_amask[l & 3] = am;
      __threadfence_block();
      *const_cast<volatile unsigned int*>(&rt->pf_route) = static_cast<unsigned int>(l) + 1u;
      if (moe && blockIdx.x == 0) {
        float aux = 0.0f;
        for (int e = 0; e < E; ++e) {
          aux += (me[e] / B) * (static_cast<float>(cnt[e]) / B);
          if (p.exp_counts != nullptr) p.exp_counts[l * p.Emax + e] = cnt[e];
        }
        if (p.l_aux != nullptr) p.l_aux[l] = aux * E;
        if (p.gate_logits != nullptr)
          for (int s = 0; s < B; ++s)
            for (int e = 0; e < E; ++e) p.gate_logits[static_cast<long long>(l) * B * p.Emax + s * E + e] = rt->logits[s][e];
      }
      mbar_arrive(route_bar);  // release: the producer may read the expert choice
    }
    consumer_sync();
    DK_STAMP(9);
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
    DK_STAMP(10);
    grid_sync(p.sync, bar_target);
    DK_STAMP(11);
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
    DK_STAMP(12);
    grid_sync(p.sync, bar_target);
    DK_STAMP(13);
  }
  // ------------------------------------------------------------ final RMSNorm (hidden_states[-1])
  if (p.out_norm != nullptr && blockIdx.x == 0) {
    stage_rows(act, p.x, s_ln_in, lnin_bar, lnin_phase, B, D, p.eps, red);
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

static long long decode_smem_bytes(int D, int B) {
  return 1024 + static_cast<long long>(DK_STAGES) * DK_STAGE_BYTES + static_cast<long long>(B) * (D * 2 + 64) +
         2LL * DK_RED_FLOATS * 4 + (2 * DK_STAGES + 4) * 8 + static_cast<long long>(sizeof(RouteSmem)) + 128 + 2 * D * 2 +
         DK_WG_SMEM + 64;
}

bool llama_decode_supported(const mpl_llama_model& m, const mpl_llama_io& io) {
  if (io.decode_plan == nullptr || io.T != 1 || io.B < 1 || io.B > DK_MAXB) return false;
  if (io.hidden_states != nullptr || io.moe_noise != nullptr) return false;
  if (m.top_k != 1 || (m.hidden % 64) != 0 || (m.ffn % 64) != 0 || m.hidden != m.n_heads * 128) return false;
  if (io.attn_scratch == nullptr) return false;
  if (decode_smem_bytes(m.hidden, io.B) > 227 * 1024 || m.hidden > 4096) return false;
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
  p.timing_layer = g_timing_layer;
  {
    // L2 prefetch look-ahead per CTA in 16 KB chunks (x 148 CTAs must stay well inside the 126 MB L2)
    static int kla = -1, ela = -1, spec = -1;
    if (kla < 0) {
      const char* a = getenv("MPL_DK_KLA");
      const char* b = getenv("MPL_DK_ELA");
      const char* c = getenv("MPL_DK_SPEC");
      kla = a != nullptr ? atoi(a) : 16;
      ela = b != nullptr ? atoi(b) : 8;
      spec = c != nullptr ? atoi(c) : 4;  // speculate "every expert is hit" from this many sequences up (0: never)
    }
    p.kla = kla;
    p.ela = ela > 0xffff ? 0xffff : ela;
    p.spec = (spec > 0 && io.B >= spec) ? 1 : 0;
  }
  for (int e = 0; e <= DK_MAXE; ++e) p.cap[e] = cap_by_e[e];
  p.eps = m.rms_eps;
  p.scale = 1.0f / sqrtf(128.0f);
  const int G = num_sms();
  // split-K over the keys: one work item per consumer warp of the grid, at least 32 keys per split, within the scratch
  const int bh = io.B * H;
  const int Tk = io.past_len + 1;
  int nsplit = (G * DK_CONSUMERS) / bh;  // one (b, h, split) per warp, at most one round
  const int by_keys = (Tk + 31) / 32;
  if (nsplit > by_keys) nsplit = by_keys;
  if (nsplit > 32) nsplit = 32;
  if (nsplit < 1) nsplit = 1;
  while (nsplit > 1 && static_cast<long long>(bh) * 4 + 256 + static_cast<long long>(bh) * nsplit * 132 * 4 > io.attn_scratch_bytes)
    --nsplit;
  p.nsplit = nsplit;
  p.attn_cnt = static_cast<int*>(io.attn_scratch);
  p.attn_part = reinterpret_cast<float*>(static_cast<char*>(io.attn_scratch) + ((static_cast<long long>(bh) * 4 + 255) & ~255LL));
  const int smem = static_cast<int>(decode_smem_bytes(D, io.B));
  static int attr_smem = 0;
  if (attr_smem < smem) {
    const cudaError_t e = cudaFuncSetAttribute(llama_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) {
      fprintf(stderr, "medplib_b200: decode kernel smem attribute (%d B): %s\n", smem, cudaGetErrorString(e));
      return MPL_ERR_CUDA;
    }
    attr_smem = smem;
  }
  void* args[] = {&p};
  const cudaError_t le = cudaLaunchCooperativeKernel(reinterpret_cast<void*>(llama_decode_kernel), dim3(G),
                                                     dim3(DK_THREADS), args, smem, st);
  if (le != cudaSuccess) {
    fprintf(stderr, "medplib_b200: decode kernel launch (grid %d, smem %d B): %s\n", G, smem, cudaGetErrorString(le));
    cudaGetLastError();
    return MPL_ERR_CUDA;
  }
  return launch_status();
}

}  // namespace mpl

// Dev tool: layer >= 0 enables per-phase timestamps of that layer in the decode kernel; out (host, [160*16] u64) != NULL
// copies the last recorded stamps back (after a device synchronise).
extern "C" int mpl_debug_decode_timing(int layer, unsigned long long* out) {
  mpl::g_timing_layer = layer;
  if (out != nullptr) {
    if (cudaDeviceSynchronize() != cudaSuccess) return MPL_ERR_CUDA;
    if (cudaMemcpyFromSymbol(out, mpl::g_dk_times, sizeof(unsigned long long) * 160 * mpl::DK_TSLOTS) != cudaSuccess)
      return MPL_ERR_CUDA;
  }
  return MPL_OK;
}
