// K9 — one persistent kernel per decode step of the LLaMA-MoE stack (T == 1, B <= 8 sequences, top-1 routing).
//
// Replaces, for the autoregressive steps of MedPLIBForCausalLM.generate / evaluate (model/MedPLIB.py:592-606,
// model/medplib/model/language_model/medplib_moe_llama.py:110-305,451-485), the ~7 launches per layer of the general
// runner (llama_stack.cu). A decode step is HBM-bound: 13-22 GB of weights are read once per token. ONE cooperative
// grid (one CTA per SM) runs all layers:
//   * a producer thread per CTA walks the whole step's weight-tile schedule and keeps a 6 x 16 KB TMA ring full; it does
//     not take part in grid barriers, so the next phase's first tiles are in shared memory when the consumers arrive;
//   * a prefetch thread walks the same schedule a bounded distance ahead with cp.async.bulk.prefetch.tensor (HBM -> L2)
//     so that HBM keeps streaming while the consumers sit in a barrier / attention / the router;
//   * 8 consumer warps run, per layer, four weight phases separated by grid barriers (atomic counter in L2):
//       step 0  q,k,v = RMSNorm(x) Wqkv^T, then RoPE + KV append + split-K attention over the cache
//       step 1  x += attn Wo^T
//       step 2  h = RMSNorm(x); router logits / softmax / top-1 / capacity slots (recomputed by every CTA, no
//               barrier); h1 = SiLU(h Wgate_e^T) * (h Wup_e^T) for the experts that received tokens
//       step 3  x += gate * (h1 Wdown_e^T)      MoE combine fused (dense layers: plain residual)
//     then the final RMSNorm. The GEMM core is mma.sync m16n8k16 with the weight rows as the M operand, k split over
//     the warps, cross-warp reduction through shared memory.
// CODE SIZE IS A FIRST-ORDER CONCERN HERE (profiles/r02_ncu_decode_kernel.md): a layer executes every phase's code
// exactly once, so anything outside the tile loop runs from a cold instruction cache — round 1's 300 KB of unrolled,
// four-times-instantiated code spent 87 % of its inter-phase samples in `stall_no_inst`. The consumer side is
// therefore ONE loop over (layer, step) with a single instance of the staging code, the tile loop and the grid barrier,
// row loops are real loops, and helpers used twice are not inlined.
// Rounding points are those of the general path (bf16 after every linear, RMSNorm's two roundings, RoPE's three, P
// rounded before P·V, bf16(gate) * bf16(y)); fp32 accumulation orders differ from it only inside the attention.
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
constexpr int DK_MAX_STAGES = 12;  // ring depth is chosen per launch: as many 16 KB stages as shared memory holds
constexpr int DK_STAGE_BYTES = 16384;  // [8 k-blocks][16 rows][128 B swizzled]  (dual: 2 x [8][8 rows][128 B])
constexpr int DK_KC = 512;
constexpr int DK_MAXB = 8;
constexpr int DK_MAXE = MPL_MAX_EXPERTS;
constexpr int DK_RP = 9;  // reduction row pitch (floats)
constexpr int DK_RED_FLOATS = DK_CONSUMERS * 16 * DK_RP;
constexpr int DK_WG_SMEM = 32768;  // router weights staged in shared memory when E*D*4 fits AND it costs no ring stage

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
  const int* rope_pos;  // [B] or NULL: per-sequence RoPE position of the new token (cache column stays `pos`)
  float* gate_logits;  // [L, B, Emax] or NULL
  float* l_aux;        // [L] or NULL
  int* exp_counts;     // [L, Emax] or NULL
  int B, D, H, F, L, Tmax, pos, nsplit, Emax, timing_layer;
  int stages, wg_smem;  // ring depth; bytes reserved for a shared-memory copy of the router weights (0: read from L2)
  int la, spec, evict_first;  // L2 prefetch look-ahead (16 KB chunks per CTA, in consumption order); walk past undecided routers
  int cap[DK_MAXE + 1];  // capacity for a layer with E experts (index E)
  float eps, scale;
};

// Optional per-phase timestamps (dev tool): consumer thread 0 of every CTA records %globaltimer at the phase boundaries
// of layer `timing_layer`.
constexpr int DK_TSLOTS = 32;
__device__ unsigned long long g_dk_times[160 * DK_TSLOTS];
static int g_timing_layer = -1;
static bool g_dprof = false;  // bench.py: CUDA events around every launch of the decode kernel
static std::vector<cudaEvent_t> g_dprof_ev;
static size_t g_dprof_used = 0;
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
  int n;  // stages
  __device__ __forceinline__ void advance() {
    if (++stage == n) {
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
  int nact;            // number of experts that received tokens
  int act[DK_MAXE];    // their indices, ascending
  float logits[DK_MAXB][DK_MAXE];
  float gates[DK_MAXB][DK_MAXE];
  // L2 prefetcher coordination (monotonic counters, written by the producer / consumer thread 0, polled by the prefetcher)
  unsigned int pf_pos;       // producer position: (layer << 16) | chunks of that layer issued so far
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

__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// Grid barrier for the consumer warps (the producer thread never waits here). Same structure as cooperative
// groups' grid.sync(): CTA barrier, one thread fences + arrives + spins + fences (the gpu-scope fence also invalidates
// this SM's L1, so plain loads after the barrier see the other CTAs' writes), CTA barrier.
__device__ __noinline__ void grid_sync(unsigned int* ctr, unsigned int& target) {
  consumer_sync();
  if (threadIdx.x == 0) {
    target += gridDim.x;
    // release (cumulative over the CTA's writes ordered by the bar.sync above) ... relaxed polling ... one acquire
    // fence, which also invalidates this SM's L1 (SASS: MEMBAR.ALL.GPU + RED / LDG.STRONG / MEMBAR + CCTL.IVALL)
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
#ifdef MPL_DK_BARRIER_FENCE  // (round-2 form: relaxed polling + one acquire fence)
    while (ld_relaxed_u32(ctr) < target) {
    }
    asm volatile("fence.acq_rel.gpu;" ::: "memory");
#else
    // acquire polling (the form of CUTLASS's GenericBarrier::wait_*): the load that observes the last arrival is the
    // acquire operation itself, no trailing MEMBAR on the critical path of the last arriver
    while (ld_acquire_u32(ctr) < target) {
    }
#endif
  }
  consumer_sync();
}

// ------------------------------------------------------------------------------------------------ producer side
__device__ __forceinline__ void tma_prefetch_3d(const void* tmap, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];" ::"l"(tmap), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
// weights are read exactly once per step: with an evict-first policy the stream does not push the lines the prefetcher
// has parked in L2 (normal priority) out before they are used
__device__ __forceinline__ void tma_load_3d_hint(void* smem_dst, const void* tmap, uint64_t* bar, int c0, int c1, int c2,
                                                 uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4, "
      "%5}], [%2], %6;" ::"r"(smem_u32(smem_dst)),
      "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "l"(policy)
      : "memory");
}
__device__ __forceinline__ void produce_tile(Ring& r, const CUtensorMap* m0, const CUtensorMap* m1, int n0, int chunks,
                                             volatile unsigned int* progress, unsigned int& count, uint64_t policy) {
  for (int c = 0; c < chunks; ++c) {
    mbar_wait(&r.empty[r.stage], r.phase ^ 1);
    uint8_t* dst = r.base + r.stage * DK_STAGE_BYTES;
    mbar_expect_tx(&r.full[r.stage], DK_STAGE_BYTES);
    if (policy != 0) {
      tma_load_3d_hint(dst, m0, &r.full[r.stage], 0, n0, c * (DK_KC / 64), policy);
      if (m1 != nullptr)
        tma_load_3d_hint(dst + DK_STAGE_BYTES / 2, m1, &r.full[r.stage], 0, n0, c * (DK_KC / 64), policy);
    } else {
      tma_load_3d(dst, m0, &r.full[r.stage], 0, n0, c * (DK_KC / 64));
      if (m1 != nullptr) tma_load_3d(dst + DK_STAGE_BYTES / 2, m1, &r.full[r.stage], 0, n0, c * (DK_KC / 64));
    }
    r.advance();
    *progress = ++count;
  }
}

// ------------------------------------------------------------------------------------------------ consumer helpers
__device__ __forceinline__ float reduce_rows(const float* rbuf, int r, int m) {
  float v = 0.0f;
#pragma unroll
  for (int w = 0; w < DK_CONSUMERS; ++w) v += rbuf[(w * 16 + r) * DK_RP + m];
  return v;
}
__device__ __forceinline__ void dk_bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void dk_prefetch_l2(const void* gsrc, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gsrc), "r"(bytes) : "memory");
}
__device__ __forceinline__ uint4 lds_v4(const void* p) { return dk_lds128(smem_u32(p)); }
__device__ __forceinline__ void sts_v4(void* p, const uint4& v) {
  asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(smem_u32(p)), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}
__device__ __forceinline__ float sumsq8(const uint4& v, float a) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __bfloat1622float2(h[i]);
    a += f.x * f.x + f.y * f.y;
  }
  return a;
}
// HF LlamaRMSNorm's roundings: w * bf16(x * rstd)
__device__ __forceinline__ uint4 norm8(const uint4& x, const uint4& w, float rstd) {
  const __nv_bfloat162* xp = reinterpret_cast<const __nv_bfloat162*>(&x);
  const __nv_bfloat162* wp = reinterpret_cast<const __nv_bfloat162*>(&w);
  uint4 o;
  uint32_t* op = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 xf = __bfloat1622float2(xp[i]), wf = __bfloat1622float2(wp[i]);
    op[i] = pack_bf16(wf.x * bf16_round(xf.x * rstd), wf.y * bf16_round(xf.y * rstd));
  }
  return o;
}

// ------------------------------------------------------------------------------------------------ attention
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
// d - 64), HF apply_rotary_pos_emb in bf16: every product and the sum rounded. Result packed as bf16. (Not inlined:
// used for q and for the new key; a second copy would only add cold code.)
__device__ __noinline__ void rope16(const __nv_bfloat16* src, const __nv_bfloat16* cr, const __nv_bfloat16* sr, int gl,
                                    uint4& o0, uint4& o1) {
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

// work split of the decode attention over the grid's consumer warps: item = (b * H + h) * nsplit + z
struct AttnItem {
  int b, h, z, bh, k_lo, k_hi;
};
__device__ __forceinline__ AttnItem attn_item(const DecParams& p, int item, int Tk) {
  AttnItem it;
  const int per = ((Tk + p.nsplit - 1) / p.nsplit + 31) & ~31;
  it.z = item % p.nsplit;
  it.bh = item / p.nsplit;
  it.b = it.bh / p.H;
  it.h = it.bh % p.H;
  it.k_lo = it.z * per;
  it.k_hi = min(Tk, it.k_lo + per);
  return it;
}

// One (b, h, split) per WARP: 4 keys per step (8 lanes x 16 dims per key), two steps per round with ONE online-softmax
// update per round, two rounds (16 keys, 8 KB) in flight per warp, no CTA-level synchronisation: the 4 key groups merge
// by shuffles, splits merge through global scratch by the last warp to arrive (atomic counter per (b, h), self-cleaning).
// The cached keys do not depend on this layer's q,k,v, so the first 16 keys of the item are requested BEFORE the grid
// barrier that ends the q,k,v phase and arrive while it is waiting. The new position (k rotated here, k,v appended to
// the cache) is folded in after the cached keys by the split that owns it.
__device__ __forceinline__ void attention_phase(const DecParams& p, int layer, int Tk, unsigned int& bar_target) {
  constexpr int D = 128;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int grp = lane >> 3, gl = lane & 7;
  const int nsplit = p.nsplit;
  const int pos = Tk - 1;
  const int n_items = p.B * p.H * nsplit;
  const int stride = gridDim.x * DK_CONSUMERS;
  const float sl2 = p.scale * 1.4426950408889634f;
  bool synced = false;
  // items are dealt round-robin over the CTAs first, so every SM pulls on the KV cache
#pragma unroll 1
  for (int item = warp * gridDim.x + blockIdx.x; item < n_items || !synced; item += stride) {
    const bool has = item < n_items;
    const AttnItem it = attn_item(p, has ? item : 0, Tk);
    const int k_lo = it.k_lo;
    const int k_end = has ? min(it.k_hi, pos) : k_lo;  // cached keys of this split: [k_lo, k_end)
    const int rp = p.rope_pos != nullptr ? p.rope_pos[it.b] : pos;  // rotation angle of this sequence's new token
    const __nv_bfloat16* cr = p.cos_t + static_cast<long long>(rp) * D;
    const __nv_bfloat16* sr = p.sin_t + static_cast<long long>(rp) * D;
    const __nv_bfloat16* qrow = p.qkv + static_cast<long long>(it.b) * 3 * p.D + it.h * D;
    const long long head_off = (static_cast<long long>(it.b) * p.H + it.h) * p.Tmax * D;
    __nv_bfloat16* kc = p.kc + layer * p.cache_layer + head_off;
    __nv_bfloat16* vc = p.vc + layer * p.cache_layer + head_off;
    const unsigned char* mrow = p.kv_mask ? p.kv_mask + static_cast<long long>(it.b) * p.kv_mask_stride : nullptr;

    // two register buffers of two steps (8 keys) each, ping-pong: one is consumed while the other is in flight
    uint4 ka[2][2], kb[2][2], va[2][2], vb[2][2];
    auto load_round = [&](int kb0, int w) {
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        if (kb0 + 4 * u < k_end) {  // warp-uniform
          const int key = kb0 + 4 * u + grp;
          const int kk = key < k_end ? key : k_lo;
          const __nv_bfloat16* kr = kc + static_cast<long long>(kk) * D + gl * 16;
          const __nv_bfloat16* vr = vc + static_cast<long long>(kk) * D + gl * 16;
          ka[w][u] = *reinterpret_cast<const uint4*>(kr);
          kb[w][u] = *reinterpret_cast<const uint4*>(kr + 8);
          va[w][u] = *reinterpret_cast<const uint4*>(vr);
          vb[w][u] = *reinterpret_cast<const uint4*>(vr + 8);
        } else {  // (a zero weight times a stale NaN pattern would still poison the accumulator)
          ka[w][u] = kb[w][u] = va[w][u] = vb[w][u] = make_uint4(0, 0, 0, 0);
        }
      }
    };
    load_round(k_lo, 0);
    load_round(k_lo + 8, 1);
    if (!synced) {
      DK_STAMP_L(layer, 16);
      grid_sync(p.sync, bar_target);  // q,k,v of this layer are complete; the first keys are already in flight
      DK_STAMP_L(layer, 17);
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
    // score of the key held in (ka, kb) against q (log2 domain), -inf when masked / out of range
    auto score = [&](const uint4& k0, const uint4& k1, bool valid, int key) {
      float kf[16];
      unpack16(k0, k1, kf);
      float dot = 0.0f;
#pragma unroll
      for (int e = 0; e < 4; ++e)
        dot += qf[2 * e] * kf[2 * e] + qf[2 * e + 1] * kf[2 * e + 1] + qf[8 + 2 * e] * kf[8 + 2 * e] +
               qf[8 + 2 * e + 1] * kf[8 + 2 * e + 1];
      dot += __shfl_xor_sync(0xffffffffu, dot, 1);
      dot += __shfl_xor_sync(0xffffffffu, dot, 2);
      dot += __shfl_xor_sync(0xffffffffu, dot, 4);
      float sc = dot * sl2;
      if (!valid || (mrow != nullptr && mrow[key] == 0)) sc = -INFINITY;
      return sc;
    };
    // online-softmax update with two keys at once (P rounded to bf16 before P·V: softmax(...).to(bf16) @ v)
    auto update2 = [&](float s0, float s1, const uint4& v00, const uint4& v01, const uint4& v10, const uint4& v11) {
      const float mn = fmaxf(m, fmaxf(s0, s1));
      const float msafe = (mn == -INFINITY) ? 0.0f : mn;
      const float corr = exp2f(m - msafe);
      const float p0 = exp2f(s0 - msafe), p1 = exp2f(s1 - msafe);
      const float r0 = bf16_round(p0), r1 = bf16_round(p1);
      l = l * corr + (p0 + p1);
      m = mn;
      float f0[16], f1[16];
      unpack16(v00, v01, f0);
      unpack16(v10, v11, f1);
#pragma unroll
      for (int e = 0; e < 16; ++e) acc[e] = (acc[e] * corr + r0 * f0[e]) + r1 * f1[e];
    };
    auto compute_round = [&](int kb0, int w) {
      const int key0 = kb0 + grp, key1 = kb0 + 4 + grp;
      const bool v0 = key0 < k_end, v1 = key1 < k_end;
      const float s0 = score(ka[w][0], kb[w][0], v0, v0 ? key0 : k_lo);
      const float s1 = (kb0 + 4 < k_end) ? score(ka[w][1], kb[w][1], v1, v1 ? key1 : k_lo) : -INFINITY;
      update2(s0, s1, va[w][0], vb[w][0], va[w][1], vb[w][1]);
    };
#pragma unroll 1
    for (int kb0 = k_lo; kb0 < k_end; kb0 += 16) {
      compute_round(kb0, 0);
      load_round(kb0 + 16, 0);
      if (kb0 + 8 < k_end) {
        compute_round(kb0 + 8, 1);
        load_round(kb0 + 24, 1);
      }
    }
    if (pos >= k_lo && pos < it.k_hi) {
      // the new position belongs to this split: rotate k, append k,v to the cache, add the key (all four lane groups
      // compute the same score -- the shuffles need the whole warp -- but only group 0 accounts for it)
      const __nv_bfloat16* krow = qrow + p.D;
      const __nv_bfloat16* vrow = qrow + 2 * p.D;
      uint4 kn0, kn1;
      rope16(krow, cr, sr, gl, kn0, kn1);
      const uint4 vn0 = *reinterpret_cast<const uint4*>(vrow + gl * 16);
      const uint4 vn1 = *reinterpret_cast<const uint4*>(vrow + gl * 16 + 8);
      if (grp == 0) {
        __nv_bfloat16* kd = kc + static_cast<long long>(pos) * D + gl * 16;
        *reinterpret_cast<uint4*>(kd) = kn0;
        *reinterpret_cast<uint4*>(kd + 8) = kn1;
        __nv_bfloat16* vd = vc + static_cast<long long>(pos) * D + gl * 16;
        *reinterpret_cast<uint4*>(vd) = vn0;
        *reinterpret_cast<uint4*>(vd + 8) = vn1;
      }
      const float s0 = score(kn0, kn1, grp == 0, pos);
      // (ka[0][1] .. are dead here: any finite-free operand works for the absent second key, its weight is exp2(-inf) = 0)
      update2(s0, -INFINITY, vn0, vn1, make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0, 0));
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
    __nv_bfloat16* optr = p.attn + static_cast<long long>(it.b) * p.D + it.h * D;
    if (nsplit == 1) {
      if (grp == 0) {
        const float inv = l > 0.0f ? __frcp_rn(l) : 0.0f;
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
    float* part = p.attn_part + static_cast<long long>(it.bh) * nsplit * (D + 4);
    if (grp == 0) {
      float* mine = part + it.z * (D + 4);
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4)
        *reinterpret_cast<float4*>(mine + gl * 16 + q4 * 4) =
            make_float4(acc[q4 * 4], acc[q4 * 4 + 1], acc[q4 * 4 + 2], acc[q4 * 4 + 3]);
      if (gl == 0) {
        mine[D] = m;
        mine[D + 1] = l;
      }
    }
    // ONE acq_rel atomic by lane 0 instead of a sequentially-consistent fence by all 32 lanes on either side of a relaxed
    // one (two MEMBAR.SC.GPU per item on the critical path of the phase): __syncwarp orders the lanes' partial stores
    // before the release, the last arriver's acquire before the merge's loads (which are ld.cg: L2, never a stale L1 line)
    __syncwarp();
    int last = 0;
    if (lane == 0) {
      unsigned int old;
      asm volatile("atom.add.acq_rel.gpu.global.u32 %0, [%1], 1;" : "=r"(old) : "l"(p.attn_cnt + it.bh) : "memory");
      last = old == static_cast<unsigned int>(nsplit - 1);
    }
    last = __shfl_sync(0xffffffffu, last, 0);
    if (!last) continue;
    __syncwarp();
    // this warp merges all splits. Lane zz first fetches (max, sum) of split zz -- one round trip for all splits --
    // then every lane accumulates its 4 dims over the splits, 8 splits' loads in flight at a time.
    const float mz = lane < nsplit ? __ldcg(part + lane * (D + 4) + D) : -INFINITY;
    const float lz = lane < nsplit ? __ldcg(part + lane * (D + 4) + D + 1) : 0.0f;
    const float gm = warp_max(mz);
    const float gsafe = (gm == -INFINITY) ? 0.0f : gm;
    const float cz = exp2f(mz - gsafe);  // 0 for absent / empty splits
    const float gl_ = warp_sum(lz * cz);
    float go[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
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
    const float inv = gl_ > 0.0f ? __frcp_rn(gl_) : 0.0f;
    uint2 o;
    o.x = pack_bf16(go[0] * inv, go[1] * inv);
    o.y = pack_bf16(go[2] * inv, go[3] * inv);
    *reinterpret_cast<uint2*>(optr + lane * 4) = o;
    if (lane == 0) p.attn_cnt[it.bh] = 0;  // self-cleaning for the next layer / launch
  }
}

// ------------------------------------------------------------------------------------------------ the kernel
__global__ void __launch_bounds__(DK_THREADS, 1) llama_decode_kernel(const __grid_constant__ DecParams p) {
  extern __shared__ uint8_t dk_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(dk_raw) + 1023) & ~uintptr_t(1023));
  const int pitch = p.D * 2 + 64;  // activation row pitch in shared memory (conflict-free 16-byte reads)
  uint8_t* s_a = smem + p.stages * DK_STAGE_BYTES;
  float* red = reinterpret_cast<float*>(s_a + p.B * pitch);  // [2][8][16][9]; staging / router scratch between phases
  // (shared memory is sized by the actual B: what it does not take stays L1, which the few spilled values need)
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(red + 2 * DK_RED_FLOATS);
  uint64_t* empty_bar = full_bar + DK_MAX_STAGES;
  uint64_t* route_bar = empty_bar + DK_MAX_STAGES;
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
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], DK_CONSUMERS);
    }
    mbar_init(route_bar, 1);
    mbar_init(act_bar, 1);
    mbar_init(wg_bar, 1);
    mbar_init(lnin_bar, 1);
    fence_mbar_init();
    rt->pf_pos = rt->pf_route = 0;
    rt->nact = 1;
    rt->act[0] = 0;
    rt->amask = 1u;
    rt->pf_amask[0] = rt->pf_amask[1] = rt->pf_amask[2] = rt->pf_amask[3] = 1u;
  }
  __syncthreads();
  Ring ring{smem, full_bar, empty_bar, 0, 0, p.stages};
  const int D = p.D, F = p.F, B = p.B;
  const int tiles_d = (D + 15) / 16;         // 16-row tiles of a [D, *] matrix
  const int tiles_f = (F + 7) / 8;           // 8+8-row tiles of the gate/up pair
  const int chunks_d = (D + DK_KC - 1) / DK_KC;
  const int chunks_f = (F + DK_KC - 1) / DK_KC;

  if (warp == DK_CONSUMERS) {
    // ============================================================ producer: the whole step's weight schedule
    // publishes its position (layer << 16 | chunks issued in this layer) for the L2 prefetcher
    if (lane != 0) return;
    uint32_t route_phase = 0;
    volatile unsigned int* pf_pos = &rt->pf_pos;
    uint64_t policy = 0;
    if (p.evict_first) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
    for (int l = 0; l < p.L; ++l) {
      const DecLayerDev* L = p.layers + l;
      unsigned int pos = static_cast<unsigned int>(l) << 16;
      *pf_pos = pos;
      for (int tile = blockIdx.x; tile < 3 * tiles_d; tile += G) {
        const int which = tile / tiles_d;
        const CUtensorMap* tm = which == 0 ? &L->wq : (which == 1 ? &L->wk : &L->wv);
        produce_tile(ring, tm, nullptr, (tile % tiles_d) * 16, chunks_d, pf_pos, pos, policy);
      }
      for (int tile = blockIdx.x; tile < tiles_d; tile += G)
        produce_tile(ring, &L->wo, nullptr, tile * 16, chunks_d, pf_pos, pos, policy);
      mbar_wait(route_bar, route_phase);  // expert choice of this layer
      route_phase ^= 1;
      const int nact = rt->nact;
      for (int tile = blockIdx.x; tile < nact * tiles_f; tile += G) {
        const int e = rt->act[tile / tiles_f];
        produce_tile(ring, &L->wgate[e], &L->wup[e], (tile % tiles_f) * 8, chunks_d, pf_pos, pos, policy);
      }
      for (int tile = blockIdx.x; tile < nact * tiles_d; tile += G) {
        const int e = rt->act[tile / tiles_d];
        produce_tile(ring, &L->wdown[e], nullptr, (tile % tiles_d) * 16, chunks_f, pf_pos, pos, policy);
      }
    }
    *pf_pos = static_cast<unsigned int>(p.L) << 16;
    return;
  }
  if (warp == DK_CONSUMERS + 1) {
    // ============================================================ L2 prefetcher
    // Walks the same per-CTA chunk schedule as the producer, IN CONSUMPTION ORDER, at most p.la 16 KB chunks ahead of
    // it, with cp.async.bulk.prefetch.tensor (HBM -> L2, fire and forget). While the consumers sit in a grid barrier,
    // the attention phase, activation staging or the tail of an unbalanced phase -- the 6-stage ring is full and the
    // producer is blocked -- HBM keeps streaming the next chunks into L2 and the ring refills from there. (Measured:
    // prefetching further ahead than ~40 MB across the grid, or out of consumption order, is evicted before its use
    // and doubles the traffic.) The expert tiles of a layer need its routing decision: with p.spec (B >= 4: every
    // expert is hit almost surely) the cursor walks on assuming all experts, otherwise it waits for the decision.
    // Every wait polls a monotonic shared-memory word that the producer / consumers advance regardless of this thread,
    // and a cursor that has fallen behind jumps forward: the thread can neither block anyone nor fail to exit.
    if (lane != 0 || p.la <= 0) return;
    const int bid = blockIdx.x;
    auto cnt = [&](int total) { return total > bid ? (total - bid + G - 1) / G : 0; };
    const int n1 = cnt(3 * tiles_d), n3 = cnt(tiles_d);
    const unsigned int nk = static_cast<unsigned int>((n1 + n3) * chunks_d);
    const unsigned int la = static_cast<unsigned int>(p.la);
    volatile RouteSmem* vr = rt;
    int pl = 0;            // cursor: layer, chunk within the layer
    unsigned int poff = 0;
    int cur_l = -1, cur_E = 1;  // cached per-layer facts
    auto experts_of = [&](int layer) {
      if (layer != cur_l) {
        const DecLayerDev* Lp = p.layers + layer;
        cur_E = Lp->wg != nullptr ? Lp->n_experts : 1;
        cur_l = layer;
      }
      return cur_E;
    };
    // chunks of layer `layer`'s expert part under mask `amask`
    auto expert_chunks = [&](unsigned int amask, unsigned int& c5) {
      const int nact = __popc(amask);
      c5 = static_cast<unsigned int>(cnt(nact * tiles_f) * chunks_d);
      return c5 + static_cast<unsigned int>(cnt(nact * tiles_d) * chunks_f);
    };
    auto mask_of = [&](int layer, bool& known) {
      const int E = experts_of(layer);
      const unsigned int full = (1u << E) - 1u;
      known = static_cast<int>(vr->pf_route) > layer;
      unsigned int m = known ? (vr->pf_amask[layer & 3] & full) : full;
      return m != 0 ? m : 1u;
    };
    // chunk `off` of a layer's q,k,v / o part
    auto issue_known = [&](const DecLayerDev* Lp, unsigned int off) {
      const int ti = static_cast<int>(off / chunks_d), c = static_cast<int>(off % chunks_d);
      if (ti < n1) {
        const int tile = bid + ti * G, which = tile / tiles_d;
        const CUtensorMap* tm = which == 0 ? &Lp->wq : (which == 1 ? &Lp->wk : &Lp->wv);
        tma_prefetch_3d(tm, 0, (tile % tiles_d) * 16, c * (DK_KC / 64));
      } else {
        tma_prefetch_3d(&Lp->wo, 0, (bid + (ti - n1) * G) * 16, c * (DK_KC / 64));
      }
    };
    while (pl < p.L) {
      const unsigned int lv = vr->pf_pos;
      const int ll = static_cast<int>(lv >> 16);
      const unsigned int loff = lv & 0xffffu;
      if (ll >= p.L) break;
      if (pl < ll || (pl == ll && poff < loff)) {  // fell behind the producer: jump to it
        pl = ll;
        poff = loff;
      }
      // distance to the producer in chunks (the cursor is never more than one layer ahead)
      unsigned int dist;
      if (pl == ll) {
        dist = poff - loff;
      } else {
        bool kn;
        unsigned int c5;
        const unsigned int end_ll = nk + expert_chunks(mask_of(ll, kn), c5);
        dist = poff + (end_ll > loff ? end_ll - loff : 0u);
        if (pl > ll + 1) dist = la;  // cannot happen (la << chunks per layer); be safe
      }
      if (dist >= la) {
        __nanosleep(64);
        continue;
      }
      const DecLayerDev* Lp = p.layers + pl;
      if (poff < nk) {
        issue_known(Lp, poff);
        ++poff;
        continue;
      }
      bool known;
      const unsigned int amask = mask_of(pl, known);
      if (!known && !p.spec) {
        // (measured, round 2: parking the NEXT layer's q,k,v tiles in L2 during this window does not pay -- the lines do
        // not survive the 270 MB expert stream even with evict-first on the ring loads, and get fetched twice)
        __nanosleep(64);
        continue;
      }
      unsigned int c5;
      const unsigned int total_e = expert_chunks(amask, c5);
      const unsigned int ex = poff - nk;
      if (ex >= total_e) {
        ++pl;
        poff = 0;
        continue;
      }
      if (ex < c5) {
        const int tile = bid + static_cast<int>(ex / chunks_d) * G, c = static_cast<int>(ex % chunks_d);
        const int e = __fns(amask, 0, tile / tiles_f + 1) & (DK_MAXE - 1);
        tma_prefetch_3d(&Lp->wgate[e], 0, (tile % tiles_f) * 8, c * (DK_KC / 64));
        tma_prefetch_3d(&Lp->wup[e], 0, (tile % tiles_f) * 8, c * (DK_KC / 64));
      } else {
        const unsigned int r = ex - c5;
        const int tile = bid + static_cast<int>(r / chunks_f) * G, c = static_cast<int>(r % chunks_f);
        const int e = __fns(amask, 0, tile / tiles_d + 1) & (DK_MAXE - 1);
        tma_prefetch_3d(&Lp->wdown[e], 0, (tile % tiles_d) * 16, c * (DK_KC / 64));
      }
      ++poff;
    }
    return;
  }

  // ============================================================== consumers
  // ONE loop over (layer, step) with a single instance of every piece of code (see the header: cold code is fetched at
  // L2 latency, so size and straight-line-ness decide the time between the weight phases).
  const int g = lane >> 2, t4 = lane & 3;
  unsigned int bar_target = 0;
  int buf = 0;
  uint32_t wg_phase = 0, lnin_phase = 0;
  const int k0 = threadIdx.x * 8, k1 = (threadIdx.x + DK_CONSUMERS * 32) * 8;  // this thread's two 8-element slices
  const bool in0 = k0 < D, in1 = k1 < D;
  const float inv_d = 1.0f / static_cast<float>(D);  // exact for the power-of-two widths
  // small per-layer weights (post-attention norm + router) go to shared memory one layer ahead
  auto issue_small = [&](int layer) {
    const DecLayerDev* Ln = p.layers + layer;
    const float* wgn = Ln->wg;
    const int En = wgn != nullptr ? Ln->n_experts : 1;
    const bool fits = wgn != nullptr && static_cast<long long>(En) * D * 4 <= p.wg_smem;
    fence_proxy_async();
    mbar_expect_tx(wg_bar, static_cast<uint32_t>(D * 2 + (fits ? En * D * 4 : 0)));
    dk_bulk_g2s(s_ln_post, Ln->post_ln, static_cast<uint32_t>(D * 2), wg_bar);
    if (fits)
      for (int e = 0; e < En; ++e) dk_bulk_g2s(s_wg + e * D, wgn + static_cast<long long>(e) * D, D * 4, wg_bar);
  };
  if (threadIdx.x == 0) {
    mbar_expect_tx(lnin_bar, static_cast<uint32_t>(D * 2));
    dk_bulk_g2s(s_ln_in, p.layers[0].input_ln, static_cast<uint32_t>(D * 2), lnin_bar);
    issue_small(0);
  }
  const int pos = p.pos_dev ? *p.pos_dev : p.pos;
  const int Tk = pos + 1;
  const float* wg_l = nullptr;
  int E = 1;
  bool wg_smem = false;
  const int n_steps = 4 * p.L + (p.out_norm != nullptr && blockIdx.x == 0 ? 1 : 0);  // + the final RMSNorm (CTA 0)
#pragma unroll 1
  for (int it = 0; it < n_steps; ++it) {
    const int l = it >> 2, step = it & 3;
    const bool fin = l >= p.L;  // final norm pseudo-step
    const DecLayerDev* L = p.layers + (fin ? p.L - 1 : l);
    DK_STAMP(step * 4);
    if (step == 0 && !fin) {  // (loads that step 2 needs: issued here, long before)
      wg_l = L->wg;
      E = wg_l != nullptr ? L->n_experts : 1;
      wg_smem = wg_l != nullptr && static_cast<long long>(E) * D * 4 <= p.wg_smem;
    }
    // ---------------------------------------------------------------------------------------------- staging
    // steps 0 / 2: RMSNorm(x) rows; step 1: attention output rows; step 3: nothing (its A operand is h1 in global memory)
    if (step != 3) {
      const __nv_bfloat16* src = step == 1 ? p.attn : p.x;
      // global (L2) -> shared memory with 16-byte cp.async copies, all rows in flight at once: one L2 round trip and no
      // registers held meanwhile (10 warps on 4 schedulers cap a thread at 168 registers)
#pragma unroll 1
      for (int r = 0; r < B; ++r) {
        const __nv_bfloat16* row = src + static_cast<long long>(r) * D;
        if (in0)
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(s_a + r * pitch + k0 * 2)),
                       "l"(row + k0)
                       : "memory");
        if (in1)
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(s_a + r * pitch + k1 * 2)),
                       "l"(row + k1)
                       : "memory");
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
      if (step == 0) DK_STAMP(25);
      if (step == 1) {
        if (threadIdx.x == 0) {  // s_ln_in is free (every thread is past step 0's staging): next layer's input norm
          fence_proxy_async();
          mbar_expect_tx(lnin_bar, static_cast<uint32_t>(D * 2));
          dk_bulk_g2s(s_ln_in, l + 1 < p.L ? p.layers[l + 1].input_ln : p.final_norm, static_cast<uint32_t>(D * 2),
                      lnin_bar);
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
      } else {
        const uint8_t* s_ln = step == 0 ? s_ln_in : s_ln_post;
        if (step == 0) {
          mbar_wait(lnin_bar, lnin_phase);
          lnin_phase ^= 1;
        } else {
          mbar_wait(wg_bar, wg_phase);  // (also covers the router weights)
          wg_phase ^= 1;
        }
        const uint4 w0 = in0 ? lds_v4(s_ln + k0 * 2) : make_uint4(0, 0, 0, 0);
        const uint4 w1 = in1 ? lds_v4(s_ln + k1 * 2) : make_uint4(0, 0, 0, 0);
        if (step == 0) DK_STAMP(26);
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        if (step == 0) DK_STAMP(27);
        // sum of squares of this thread's slices; lane -> warp (shuffles) -> CTA (fixed order 0..7). Rows go in groups of
        // FOUR with every shared-memory load of a group issued before its arithmetic: the loads / stores are volatile asm
        // (kept in program order), so a row-at-a-time loop is ONE serial load -> FMA -> shuffle chain per row on the two
        // warps a scheduler has -- 0.22 us per row measured (B = 8: 1.8 us here, 3.3 us in the normalising loop below).
        const uint4 zero4 = make_uint4(0, 0, 0, 0);
        if (B == 1) {  // (the single-sequence step keeps its own minimal code: nothing to interleave)
          float ss = 0.0f;
          if (in0) ss = sumsq8(lds_v4(s_a + k0 * 2), ss);
          if (in1) ss = sumsq8(lds_v4(s_a + k1 * 2), ss);
          ss = warp_sum(ss);
          if (lane == 0) red[warp] = ss;
        }
#pragma unroll 1
        for (int r0 = 0; r0 < B && B > 1; r0 += 4) {
          uint4 xa[4], xb[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            xa[j] = xb[j] = zero4;
            if (r0 + j < B) {
              const uint8_t* rowp = s_a + (r0 + j) * pitch;
              if (in0) xa[j] = lds_v4(rowp + k0 * 2);
              if (in1) xb[j] = lds_v4(rowp + k1 * 2);
            }
          }
          float ss[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) ss[j] = sumsq8(xb[j], sumsq8(xa[j], 0.0f));
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
            for (int j = 0; j < 4; ++j) ss[j] += __shfl_xor_sync(0xffffffffu, ss[j], o);
          }
          if (lane == 0) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
              if (r0 + j < B) red[(r0 + j) * DK_CONSUMERS + warp] = ss[j];
          }
        }
        consumer_sync();
        if (step == 0) DK_STAMP(28);
        if (B == 1) {
          const float4 p0 = *reinterpret_cast<const float4*>(red);
          const float4 p1 = *reinterpret_cast<const float4*>(red + 4);
          const float tot = ((((((p0.x + p0.y) + p0.z) + p0.w) + p1.x) + p1.y) + p1.z) + p1.w;
          const float rstd = rsqrtf(__fmul_rn(tot, inv_d) + p.eps);
          if (in0) sts_v4(s_a + k0 * 2, norm8(lds_v4(s_a + k0 * 2), w0, rstd));
          if (in1) sts_v4(s_a + k1 * 2, norm8(lds_v4(s_a + k1 * 2), w1, rstd));
        }
#pragma unroll 1
        for (int r0 = 0; r0 < B && B > 1; r0 += 4) {
          uint4 xa[4], xb[4];
          float rstd[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            xa[j] = xb[j] = zero4;
            rstd[j] = 0.0f;
            if (r0 + j < B) {
              const uint8_t* rowp = s_a + (r0 + j) * pitch;
              if (in0) xa[j] = lds_v4(rowp + k0 * 2);
              if (in1) xb[j] = lds_v4(rowp + k1 * 2);
              const float4 p0 = *reinterpret_cast<const float4*>(red + (r0 + j) * DK_CONSUMERS);
              const float4 p1 = *reinterpret_cast<const float4*>(red + (r0 + j) * DK_CONSUMERS + 4);
              const float tot = ((((((p0.x + p0.y) + p0.z) + p0.w) + p1.x) + p1.y) + p1.z) + p1.w;
              rstd[j] = rsqrtf(__fmul_rn(tot, inv_d) + p.eps);
            }
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (r0 + j < B) {
              xa[j] = norm8(xa[j], w0, rstd[j]);
              xb[j] = norm8(xb[j], w1, rstd[j]);
            }
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (r0 + j < B) {
              if (in0) sts_v4(s_a + (r0 + j) * pitch + k0 * 2, xa[j]);
              if (in1) sts_v4(s_a + (r0 + j) * pitch + k1 * 2, xb[j]);
            }
          }
        }
      }
      if (fin) {  // hidden_states[-1] = RMSNorm(x) with the final norm weight (CTA 0 only)
        consumer_sync();
        for (int i = threadIdx.x; i < B * (D / 8); i += DK_CONSUMERS * 32) {
          const int m = i / (D / 8), c = (i % (D / 8)) * 8;
          *reinterpret_cast<uint4*>(p.out_norm + static_cast<long long>(m) * D + c) = lds_v4(s_a + m * pitch + c * 2);
        }
        break;
      }
    }
    // ---------------------------------------------------------------------------------------------- router (step 2)
    if (step == 2) {
      DK_STAMP(23);
      if (wg_l != nullptr) {
        // logits from the stored (bf16-rounded) h in fp32 like DeepSpeed's TopKGate. Thread t owns the two slices it
        // has just written (its own shared-memory stores: readable without a barrier); partial sums lane -> warp
        // (shuffles) -> CTA (fixed order 0..7, by the thread that owns the row)
        const float* wgp = wg_smem ? s_wg : wg_l;  // generic pointer: shared-memory copy when it fits
        float* s_rp = red + 64;                    // [B][DK_MAXE][8 warps]
        // (weights of expert e + 1 are requested while expert e is computed: they may come from L2)
        auto load_w = [&](int e, float4& a, float4& b, float4& c, float4& d) {
          a = b = c = d = make_float4(0.f, 0.f, 0.f, 0.f);
          if (e < E) {
            const float* we = wgp + static_cast<long long>(e) * D;
            if (in0) a = *reinterpret_cast<const float4*>(we + k0), b = *reinterpret_cast<const float4*>(we + k0 + 4);
            if (in1) c = *reinterpret_cast<const float4*>(we + k1), d = *reinterpret_cast<const float4*>(we + k1 + 4);
          }
        };
        float4 wa, wb, wc, wd, na, nb, nc, nd;
        load_w(0, na, nb, nc, nd);
#pragma unroll 1
        for (int e = 0; e < E; ++e) {
          wa = na, wb = nb, wc = nc, wd = nd;
          load_w(e + 1, na, nb, nc, nd);
          // rows in groups of four, loads first (see the RMSNorm staging above: 0.42 us per row otherwise)
          auto dot8 = [](const uint4& raw, const float4& a0, const float4& a1, float acc) {
            const __nv_bfloat162* hp = reinterpret_cast<const __nv_bfloat162*>(&raw);
            const float2 f0 = __bfloat1622float2(hp[0]), f1 = __bfloat1622float2(hp[1]);
            const float2 f2 = __bfloat1622float2(hp[2]), f3 = __bfloat1622float2(hp[3]);
            acc += f0.x * a0.x + f0.y * a0.y + f1.x * a0.z + f1.y * a0.w + f2.x * a1.x + f2.y * a1.y + f3.x * a1.z +
                   f3.y * a1.w;
            return acc;
          };
          if (B == 1) {
            float acc = 0.0f;
            if (in0) acc = dot8(lds_v4(s_a + k0 * 2), wa, wb, acc);
            if (in1) acc = dot8(lds_v4(s_a + k1 * 2), wc, wd, acc);
            acc = warp_sum(acc);
            if (lane == 0) s_rp[e * DK_CONSUMERS + warp] = acc;
          }
#pragma unroll 1
          for (int m0 = 0; m0 < B && B > 1; m0 += 4) {
            uint4 xa[4], xb[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              xa[j] = xb[j] = make_uint4(0, 0, 0, 0);
              if (m0 + j < B) {
                const uint8_t* rowp = s_a + (m0 + j) * pitch;
                if (in0) xa[j] = lds_v4(rowp + k0 * 2);
                if (in1) xb[j] = lds_v4(rowp + k1 * 2);
              }
            }
            float acc[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              acc[j] = 0.0f;
              if (m0 + j < B) {
                if (in0) acc[j] = dot8(xa[j], wa, wb, acc[j]);
                if (in1) acc[j] = dot8(xb[j], wc, wd, acc[j]);
              }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
              for (int j = 0; j < 4; ++j) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], o);
            }
            if (lane == 0) {
#pragma unroll
              for (int j = 0; j < 4; ++j)
                if (m0 + j < B) s_rp[((m0 + j) * DK_MAXE + e) * DK_CONSUMERS + warp] = acc[j];
            }
          }
        }
        DK_STAMP(29);
        consumer_sync();
        DK_STAMP(30);
        if (threadIdx.x < B) {  // the thread that owns row m: logits, softmax (rows 0..7 all live in warp 0)
          const int m = threadIdx.x;
          float mx = -INFINITY;
          _Pragma("unroll 1") for (int e = 0; e < E; ++e) {
            float v = 0.0f;
#pragma unroll
            for (int w = 0; w < DK_CONSUMERS; ++w) v += s_rp[(m * DK_MAXE + e) * DK_CONSUMERS + w];
            rt->logits[m][e] = v;
            mx = fmaxf(mx, v);
          }
          float sum = 0.0f;
          _Pragma("unroll 1") for (int e = 0; e < E; ++e) sum += expf(rt->logits[m][e] - mx);
          _Pragma("unroll 1") for (int e = 0; e < E; ++e) rt->gates[m][e] = expf(rt->logits[m][e] - mx) / sum;
        }
        if (warp == 0) __syncwarp();
      }
      DK_STAMP(24);
      if (warp == 0) {
        // top-1 + capacity slots in token order (torch.cumsum), as moe_scan_kernel / moe_route_small_kernel: lane = token,
        // a token's slot is the number of earlier tokens with the same choice (ballot + popc) -- the serial loop over the
        // tokens this replaces cost 2.5 us per layer at B = 8 (one thread, dependent shared-memory round trips)
        const bool moe = wg_l != nullptr;
        const int C = p.cap[E];
        const bool tok = lane < B;
        int i1 = 0;
        float gsel = 1.0f;
        if (moe && tok) {
          float best = rt->gates[lane][0];
          _Pragma("unroll 1") for (int e = 1; e < E; ++e) {
            const float gv = rt->gates[lane][e];
            if (gv > best) best = gv, i1 = e;
          }
          gsel = best;
        }
        unsigned int am = 0;
        int na = 0;
        _Pragma("unroll 1") for (int e = 0; e < E; ++e) {
          const bool mine = tok && i1 == e;
          const unsigned int votes = __ballot_sync(0xffffffffu, mine);
          const int n_e = __popc(votes);
          const int kept_e = moe ? min(n_e, C) : n_e;
          if (mine) {
            const int loc = __popc(votes & ((1u << lane) - 1u));
            if (loc < C || !moe) {
              rt->tok_of_slot[e][loc] = lane;
              rt->gate_of_slot[e][loc] = gsel;
            }
          }
          if (lane == 0) {
            rt->cnt[e] = n_e;
            rt->kept[e] = kept_e;
            if (kept_e > 0) rt->act[na] = e;
          }
          if (kept_e > 0) am |= 1u << e, ++na;
        }
        __syncwarp();
        if (lane == 0) {
          rt->amask = am;
          rt->nact = na;
          rt->moe = moe ? 1 : 0;
          rt->pf_amask[l & 3] = am;
          __threadfence_block();
          *const_cast<volatile unsigned int*>(&rt->pf_route) = static_cast<unsigned int>(l) + 1u;
          mbar_arrive(route_bar);  // release: the producer may read the expert choice
        }
      }
      consumer_sync();  // staged rows + routing decision visible to every warp
      if (threadIdx.x == 0) {  // off the critical path: statistics, next layer's small weights
        if (wg_l != nullptr && blockIdx.x == 0) {
          float aux = 0.0f;
          _Pragma("unroll 1") for (int e = 0; e < E; ++e) {
            float me_e = 0.0f;  // sum of the gate probabilities of e, in token order
            _Pragma("unroll 1") for (int sq = 0; sq < B; ++sq) me_e += rt->gates[sq][e];
            rt->me[e] = me_e;
            aux += (me_e / B) * (static_cast<float>(rt->cnt[e]) / B);
            if (p.exp_counts != nullptr) p.exp_counts[l * p.Emax + e] = rt->cnt[e];
          }
          if (p.l_aux != nullptr) p.l_aux[l] = aux * E;
          if (p.gate_logits != nullptr)
            _Pragma("unroll 1") for (int sq = 0; sq < B; ++sq)
              _Pragma("unroll 1") for (int e = 0; e < E; ++e)
                p.gate_logits[static_cast<long long>(l) * B * p.Emax + sq * E + e] = rt->logits[sq][e];
        }
        if (l + 1 < p.L) issue_small(l + 1);  // every warp is past its reads of s_ln_post / s_wg
      }
    } else if (step != 3) {
      consumer_sync();  // staged rows visible to every warp
    }
    DK_STAMP(step * 4 + 1);
    // ---------------------------------------------------------------------------------------------- weight tiles
    // 16 weight rows (step 2: 8 gate + 8 up rows) x K per tile; a warp owns one 64-wide k block of every 512-wide chunk
    {
      const bool dual = step == 2;
      const int nact = rt->nact;
      const int per_e = step == 2 ? tiles_f : tiles_d;  // tiles per weight matrix
      const int n_tiles = step == 0 ? 3 * tiles_d : (step == 1 ? tiles_d : nact * per_e);
      const int chunks = step == 3 ? chunks_f : chunks_d;
      const int K = step == 3 ? F : D;
      const uint32_t warp_off = static_cast<uint32_t>(warp * ((dual ? 8 : 16) * 128) + g * 128);
      const uint32_t hi = dual ? DK_STAGE_BYTES / 2 : 1024;
      const uint32_t sw0 = static_cast<uint32_t>((t4 ^ g) * 16), sw1 = static_cast<uint32_t>(((4 + t4) ^ g) * 16);
      const uint4 zero = make_uint4(0, 0, 0, 0);
#pragma unroll 1
      for (int tile = blockIdx.x; tile < n_tiles; tile += G) {
        const int mat = tile / per_e, n0 = (tile - mat * per_e) * (dual ? 8 : 16);
        int e = 0, M = B;
        const __nv_bfloat16* arow = nullptr;  // activation row of this lane's column (generic: shared or global)
        if (step >= 2) {
          e = rt->act[mat];
          M = rt->kept[e];
          if (g < M)
            arow = step == 2 ? reinterpret_cast<const __nv_bfloat16*>(s_a + rt->tok_of_slot[e][g] * pitch)
                             : p.h1 + (static_cast<long long>(e) * B + g) * F;
        } else if (g < B) {
          arow = reinterpret_cast<const __nv_bfloat16*>(s_a + g * pitch);
        }
        float* rbuf = red + buf * DK_RED_FLOATS;
        buf ^= 1;
        float acc0[4] = {0.f, 0.f, 0.f, 0.f};
        // activation fragments run two chunks ahead of the weights (step 3 reads them from global memory: L2 latency)
        const int kw = warp * 64 + t4 * 8;
        auto load_a = [&](int c, uint4& x0, uint4& x1) {
          const int k = c * DK_KC + kw;
          x0 = (arow != nullptr && c < chunks && k < K) ? *reinterpret_cast<const uint4*>(arow + k) : zero;
          x1 = (arow != nullptr && c < chunks && k + 32 < K) ? *reinterpret_cast<const uint4*>(arow + k + 32) : zero;
        };
        uint4 xa0, xa1, xn0, xn1;
        load_a(0, xa0, xa1);
        load_a(1, xn0, xn1);
#pragma unroll 1
        for (int c = 0; c < chunks; ++c) {
          const uint4 xb0 = xa0, xb1 = xa1;
          xa0 = xn0;
          xa1 = xn1;
          load_a(c + 2, xn0, xn1);
          mbar_wait(&ring.full[ring.stage], ring.phase);
          const uint32_t sbase = smem_u32(ring.base + ring.stage * DK_STAGE_BYTES) + warp_off;
          {
            const uint4 wa = dk_lds128(sbase + sw0), wb = dk_lds128(sbase + hi + sw0);
            dk_hmma(acc0, wa.x, wb.x, wa.y, wb.y, xb0.x, xb0.y);
            dk_hmma(acc0, wa.z, wb.z, wa.w, wb.w, xb0.z, xb0.w);
          }
          {
            const uint4 wa = dk_lds128(sbase + sw1), wb = dk_lds128(sbase + hi + sw1);
            dk_hmma(acc0, wa.x, wb.x, wa.y, wb.y, xb1.x, xb1.y);
            dk_hmma(acc0, wa.z, wb.z, wa.w, wb.w, xb1.z, xb1.w);
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&ring.empty[ring.stage]);
          ring.advance();
        }
        // C fragment: c0,c1 -> (weight row g, m = 2t, 2t+1); c2,c3 -> (weight row g+8, same m)
        float* rw = rbuf + warp * 16 * DK_RP;
        rw[g * DK_RP + t4 * 2] = acc0[0];
        rw[g * DK_RP + t4 * 2 + 1] = acc0[1];
        rw[(g + 8) * DK_RP + t4 * 2] = acc0[2];
        rw[(g + 8) * DK_RP + t4 * 2 + 1] = acc0[3];
        consumer_sync();
        // epilogue: 16 rows (step 2: 8 gate/up pairs) x M sequences, one value per thread
        if (threadIdx.x < 16 * DK_MAXB) {
          const int r = dual ? (threadIdx.x & 7) : (threadIdx.x & 15);
          const int m = dual ? (threadIdx.x >> 3) : (threadIdx.x >> 4);
          const int n = n0 + r;
          if (step == 0) {
            if (m < B && n < D)
              p.qkv[static_cast<long long>(m) * 3 * D + mat * D + n] = __float2bfloat16_rn(reduce_rows(rbuf, r, m));
          } else if (step == 1) {
            if (m < B && n < D) {
              __nv_bfloat16* xp = p.x + static_cast<long long>(m) * D + n;
              *xp = __float2bfloat16_rn(bf16_round(reduce_rows(rbuf, r, m)) + __bfloat162float(*xp));
            }
          } else if (step == 2) {
            if (m < M && m < DK_MAXB && n < F) {
              const float gte = bf16_round(reduce_rows(rbuf, r, m)), up = bf16_round(reduce_rows(rbuf, r + 8, m));
              const float v = bf16_round(gte / (1.0f + __expf(-gte))) * up;
              p.h1[(static_cast<long long>(e) * B + m) * F + n] = __float2bfloat16_rn(v);
            }
          } else {
            if (m < M && n < D) {
              float v = reduce_rows(rbuf, r, m);
              if (rt->moe) v = bf16_round(v) * bf16_round(rt->gate_of_slot[e][m]);  // combine_weights.type_as(x)
              __nv_bfloat16* xp = p.x + static_cast<long long>(rt->tok_of_slot[e][m]) * D + n;
              *xp = __float2bfloat16_rn(bf16_round(v) + __bfloat162float(*xp));
            }
          }
        }
      }
    }
    DK_STAMP(step * 4 + 2);
    // ---------------------------------------------------------------------------------------------- phase end
    if (step == 0) attention_phase(p, l, Tk, bar_target);  // (includes the grid barrier that ends the q,k,v tiles)
    DK_STAMP(step == 0 ? 18 : step * 4 + 3);
    grid_sync(p.sync, bar_target);
    DK_STAMP(step == 0 ? 3 : 19 + step);
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

// shared memory of a launch: ring | activation rows | reduction buffers | barriers | routing state | norm weights |
// (optional) router weights. The ring takes every 16 KB stage that fits: the weight stream is latency-bound by the
// bytes in flight per SM (6 stages = 96 KB gave 46 GB/s per SM, i.e. the whole chip only just at the HBM rate and any
// imbalance below it), so depth is worth more than the shared-memory copy of the router weights.
static long long decode_smem_fixed(int D, int B) {
  return 1024 + static_cast<long long>(B) * (D * 2 + 64) + 2LL * DK_RED_FLOATS * 4 + (2 * DK_MAX_STAGES + 4) * 8 +
         static_cast<long long>(sizeof(RouteSmem)) + 128 + 2 * D * 2 + 64;
}
static void decode_smem_plan(int D, int B, int emax, int* stages, int* wg_smem, long long* total) {
  const long long cap = 227 * 1024;
  const long long fixed = decode_smem_fixed(D, B);
  int st = static_cast<int>((cap - fixed) / DK_STAGE_BYTES);
  if (st > DK_MAX_STAGES) st = DK_MAX_STAGES;
  const char* env = getenv("MPL_DK_STAGES");
  if (env != nullptr && atoi(env) >= 2 && atoi(env) < st) st = atoi(env);
  long long wg = static_cast<long long>(emax) * D * 4;
  if (wg > DK_WG_SMEM || fixed + static_cast<long long>(st) * DK_STAGE_BYTES + wg > cap) wg = 0;
  *stages = st;
  *wg_smem = static_cast<int>(wg);
  *total = fixed + static_cast<long long>(st) * DK_STAGE_BYTES + wg;
}

bool llama_decode_supported(const mpl_llama_model& m, const mpl_llama_io& io) {
  if (io.decode_plan == nullptr || io.T != 1 || io.B < 1 || io.B > DK_MAXB) return false;
  if (io.hidden_states != nullptr || io.moe_noise != nullptr) return false;
  if (m.top_k != 1 || (m.hidden % 64) != 0 || (m.ffn % 64) != 0 || m.hidden != m.n_heads * 128) return false;
  if (io.attn_scratch == nullptr) return false;
  if (decode_smem_fixed(m.hidden, io.B) + 4LL * DK_STAGE_BYTES > 227 * 1024 || m.hidden > 4096) return false;
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
  p.rope_pos = io.rope_pos;
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
    // L2 prefetch look-ahead per CTA in 16 KB chunks (x 148 CTAs: 16 chunks = 38 MB, well inside the 126 MB L2)
    static int la = -1, spec = -1, evict = 0;
    if (la < 0) {
      const char* a = getenv("MPL_DK_LA");
      const char* c = getenv("MPL_DK_SPEC");
      const char* d = getenv("MPL_DK_EVICT");
      evict = d != nullptr ? atoi(d) : 1;
      la = a != nullptr ? atoi(a) : 16;
      spec = c != nullptr ? atoi(c) : 4;  // walk past an undecided router assuming "every expert is hit" from this B up
    }
    p.la = la > 0x3fff ? 0x3fff : la;
    p.spec = (spec > 0 && io.B >= spec) ? 1 : 0;
    p.evict_first = evict;
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
  long long smem_ll = 0;
  decode_smem_plan(D, io.B, emax, &p.stages, &p.wg_smem, &smem_ll);
  const int smem = static_cast<int>(smem_ll);
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
  cudaEvent_t pe0 = nullptr, pe1 = nullptr;
  if (g_dprof) {
    if (g_dprof_used + 2 > g_dprof_ev.size()) {
      cudaEvent_t a, b;
      if (cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&b) != cudaSuccess) return MPL_ERR_CUDA;
      g_dprof_ev.push_back(a);
      g_dprof_ev.push_back(b);
    }
    pe0 = g_dprof_ev[g_dprof_used], pe1 = g_dprof_ev[g_dprof_used + 1];
    g_dprof_used += 2;
    cudaEventRecord(pe0, st);
  }
  const cudaError_t le = cudaLaunchCooperativeKernel(reinterpret_cast<void*>(llama_decode_kernel), dim3(G),
                                                     dim3(DK_THREADS), args, smem, st);
  if (le != cudaSuccess) {
    fprintf(stderr, "medplib_b200: decode kernel launch (grid %d, smem %d B): %s\n", G, smem, cudaGetErrorString(le));
    cudaGetLastError();
    return MPL_ERR_CUDA;
  }
  if (pe1 != nullptr) cudaEventRecord(pe1, st);
  return launch_status();
}

}  // namespace mpl

// In-situ timing of the decode kernel's launches (bench.py roofline): CUDA events on the launch stream around each one.
extern "C" int mpl_profile_decode(int enable) {
  mpl::g_dprof = enable != 0;
  mpl::g_dprof_used = 0;
  return MPL_OK;
}
extern "C" int mpl_profile_decode_read(float* total_ms, int* launches) {
  if (cudaDeviceSynchronize() != cudaSuccess) return MPL_ERR_CUDA;
  float tot = 0.0f;
  for (size_t i = 0; i + 1 < mpl::g_dprof_used; i += 2) {
    float ms = 0.0f;
    if (cudaEventElapsedTime(&ms, mpl::g_dprof_ev[i], mpl::g_dprof_ev[i + 1]) != cudaSuccess) return MPL_ERR_CUDA;
    tot += ms;
  }
  if (total_ms) *total_ms = tot;
  if (launches) *launches = static_cast<int>(mpl::g_dprof_used / 2);
  mpl::g_dprof_used = 0;
  return MPL_OK;
}

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
