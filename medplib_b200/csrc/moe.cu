// K4 — MoE router + top-k token scatter / gather (HBM-bound, warp-shuffle reductions, 16-byte row copies).
//
// Replaces deepspeed.moe.layer.MoE as the reference calls it (model/MedPLIB.py:253-263,
// model/medplib/model/language_model/medplib_moe_llama.py:141-147,604-614; DeepSpeed 0.13.1 semantics restated in
// SURVEY.md App. A.3 / oracle/moe.py): fp32 gate GEMV -> softmax -> top-1 (or top-2) -> capacity -> slot per token.
// DeepSpeed materialises dispatch/combine as dense one-hot einsums over [S,E,C]; here a token's row is copied
// straight to row e*C+slot of the expert buffer and gathered back scaled by its gate value. Expert loads stay on the
// device (kept[e] feeds the expert GEMMs' m_dev), so there is no exp_counts.to('cpu') sync per layer.
//   router  : one warp per token, E dot products of length D against the fp32 gate, shuffle-reduced
//   scan    : one CTA; per-thread token segments -> per-expert counts -> exclusive scan -> slots (deterministic,
//             position order = torch.cumsum order)
//   dispatch: one CTA per (token, route): 16-byte vector copy of the row into its slot
//   combine : one CTA per token: out = residual + bf16(sum_j bf16(gate_j) * y[slot_j])
#include "internal.h"
#include "ptx.cuh"

namespace mpl {

constexpr int MOE_MAX_E = 8;

__global__ void __launch_bounds__(128) moe_router_kernel(const __nv_bfloat16* __restrict__ h, long long ldh,
                                                         const float* __restrict__ wg, int S, int D, int E,
                                                         float* __restrict__ logits, float* __restrict__ gates) {
  griddep_wait();  // programmatic dependent launch: the previous kernel's outputs are read from here on
  griddep_launch_dependents();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int s = blockIdx.x * 4 + warp;
  if (s >= S) return;
  const __nv_bfloat16* hr = h + static_cast<long long>(s) * ldh;
  float acc[MOE_MAX_E];
#pragma unroll
  for (int e = 0; e < MOE_MAX_E; ++e) acc[e] = 0.0f;
  for (int c = lane * 8; c < D; c += 256) {
    const uint4 raw = *reinterpret_cast<const uint4*>(hr + c);
    const __nv_bfloat162* hp = reinterpret_cast<const __nv_bfloat162*>(&raw);
    float x[8];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 f = __bfloat1622float2(hp[i]);
      x[2 * i] = f.x;
      x[2 * i + 1] = f.y;
    }
#pragma unroll
    for (int e = 0; e < MOE_MAX_E; ++e) {
      if (e < E) {
        const float4 w0 = *reinterpret_cast<const float4*>(wg + static_cast<long long>(e) * D + c);
        const float4 w1 = *reinterpret_cast<const float4*>(wg + static_cast<long long>(e) * D + c + 4);
        acc[e] += x[0] * w0.x + x[1] * w0.y + x[2] * w0.z + x[3] * w0.w + x[4] * w1.x + x[5] * w1.y + x[6] * w1.z +
                  x[7] * w1.w;
      }
    }
  }
#pragma unroll
  for (int e = 0; e < MOE_MAX_E; ++e) acc[e] = warp_sum(acc[e]);
  if (lane == 0) {
    float m = -INFINITY;
#pragma unroll
    for (int e = 0; e < MOE_MAX_E; ++e)
      if (e < E) m = fmaxf(m, acc[e]);
    float ex[MOE_MAX_E];
    float sum = 0.0f;
#pragma unroll
    for (int e = 0; e < MOE_MAX_E; ++e)
      if (e < E) {
        ex[e] = expf(acc[e] - m);
        sum += ex[e];
      }
#pragma unroll
    for (int e = 0; e < MOE_MAX_E; ++e)
      if (e < E) {
        logits[static_cast<long long>(s) * E + e] = acc[e];
        gates[static_cast<long long>(s) * E + e] = ex[e] / sum;
      }
  }
}

// RMSNorm fused into the router: h = w * bf16(x * rstd) (HF LlamaRMSNorm roundings, stored for the expert GEMMs) and
// the router logits from the stored (rounded) h. One 128-thread CTA per token, the row in registers (D <= 4096: every
// global load is issued before the first use), lane -> warp (shuffles) -> CTA (shared memory, fixed order) reductions.
__global__ void __launch_bounds__(128) moe_norm_router_kernel(const __nv_bfloat16* __restrict__ x, long long ldx,
                                                              const __nv_bfloat16* __restrict__ ln_w, float eps,
                                                              __nv_bfloat16* __restrict__ h, long long ldh,
                                                              const float* __restrict__ wg, int D, int E,
                                                              float* __restrict__ logits, float* __restrict__ gates) {
  griddep_wait();  // programmatic dependent launch: the previous kernel's outputs are read from here on
  griddep_launch_dependents();
  constexpr int NV = 4;  // 16-byte vectors per thread: D <= 128 * 4 * 8
  __shared__ float red[4][MOE_MAX_E + 1];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int s = blockIdx.x;
  const __nv_bfloat16* xr = x + static_cast<long long>(s) * ldx;
  uint4 v[NV], wr[NV];
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int c = (j * 128 + threadIdx.x) * 8;
    v[j] = c < D ? *reinterpret_cast<const uint4*>(xr + c) : make_uint4(0, 0, 0, 0);
    wr[j] = c < D ? *reinterpret_cast<const uint4*>(ln_w + c) : make_uint4(0, 0, 0, 0);
  }
  float ss = 0.0f;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const __nv_bfloat162* hp = reinterpret_cast<const __nv_bfloat162*>(&v[j]);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 f = __bfloat1622float2(hp[i]);
      ss += f.x * f.x + f.y * f.y;
    }
  }
  ss = warp_sum(ss);
  if (lane == 0) red[warp][MOE_MAX_E] = ss;
  __syncthreads();
  const float rstd = rsqrtf((((red[0][MOE_MAX_E] + red[1][MOE_MAX_E]) + red[2][MOE_MAX_E]) + red[3][MOE_MAX_E]) /
                                static_cast<float>(D) + eps);
  float acc[MOE_MAX_E];
#pragma unroll
  for (int e = 0; e < MOE_MAX_E; ++e) acc[e] = 0.0f;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int c = (j * 128 + threadIdx.x) * 8;
    if (c < D) {
      const __nv_bfloat162* xp = reinterpret_cast<const __nv_bfloat162*>(&v[j]);
      const __nv_bfloat162* wp = reinterpret_cast<const __nv_bfloat162*>(&wr[j]);
      float hv[8];
      uint4 o;
      uint32_t* op = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 xf = __bfloat1622float2(xp[i]), wf = __bfloat1622float2(wp[i]);
        hv[2 * i] = bf16_round(wf.x * bf16_round(xf.x * rstd));
        hv[2 * i + 1] = bf16_round(wf.y * bf16_round(xf.y * rstd));
        op[i] = pack_bf16(hv[2 * i], hv[2 * i + 1]);
      }
      *reinterpret_cast<uint4*>(h + static_cast<long long>(s) * ldh + c) = o;
#pragma unroll
      for (int e = 0; e < MOE_MAX_E; ++e) {
        if (e < E) {
          const float4 w0 = __ldg(reinterpret_cast<const float4*>(wg + static_cast<long long>(e) * D + c));
          const float4 w1 = __ldg(reinterpret_cast<const float4*>(wg + static_cast<long long>(e) * D + c + 4));
          acc[e] += hv[0] * w0.x + hv[1] * w0.y + hv[2] * w0.z + hv[3] * w0.w + hv[4] * w1.x + hv[5] * w1.y +
                    hv[6] * w1.z + hv[7] * w1.w;
        }
      }
    }
  }
#pragma unroll
  for (int e = 0; e < MOE_MAX_E; ++e)
    if (e < E) {
      acc[e] = warp_sum(acc[e]);
      if (lane == 0) red[warp][e] = acc[e];
    }
  __syncthreads();
  if (threadIdx.x == 0) {
    float lg[MOE_MAX_E];
    float m = -INFINITY;
#pragma unroll
    for (int e = 0; e < MOE_MAX_E; ++e)
      if (e < E) {
        lg[e] = ((red[0][e] + red[1][e]) + red[2][e]) + red[3][e];
        m = fmaxf(m, lg[e]);
      }
    float ex[MOE_MAX_E];
    float sum = 0.0f;
#pragma unroll
    for (int e = 0; e < MOE_MAX_E; ++e)
      if (e < E) {
        ex[e] = expf(lg[e] - m);
        sum += ex[e];
      }
#pragma unroll
    for (int e = 0; e < MOE_MAX_E; ++e)
      if (e < E) {
        logits[static_cast<long long>(s) * E + e] = lg[e];
        gates[static_cast<long long>(s) * E + e] = ex[e] / sum;
      }
  }
}

constexpr int SCAN_THREADS = 1024;

__device__ __forceinline__ int block_sum_int(int v, int* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) red[warp] = v;
  __syncthreads();
  int t = red[lane];  // SCAN_THREADS / 32 == 32 partial sums
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  __syncthreads();
  return t;
}

// One CTA. Selection (top-1 / top-2), capacity, slots, kept counts, exp_counts, l_aux.
__global__ void __launch_bounds__(SCAN_THREADS) moe_scan_kernel(const float* __restrict__ logits,
                                                                const float* __restrict__ gates,
                                                                const float* __restrict__ noise, int S, int E, int k,
                                                                int C, int* __restrict__ expert,
                                                                float* __restrict__ gate, int* __restrict__ slot,
                                                                int* __restrict__ kept, int* __restrict__ exp_counts,
                                                                float* __restrict__ l_aux, int skey_ok) {
  griddep_wait();  // programmatic dependent launch: the previous kernel's outputs are read from here on
  griddep_launch_dependents();
  extern __shared__ float skey[];  // [S] when skey_ok: RTS keys of one expert
  __shared__ int cnt[SCAN_THREADS][MOE_MAX_E];  // per-thread segment counts, then exclusive offsets
  __shared__ int total1[MOE_MAX_E], total2[MOE_MAX_E];
  __shared__ float me_sum[MOE_MAX_E];
  __shared__ float red[SCAN_THREADS / 32];
  __shared__ int ired[SCAN_THREADS / 32];
  const int tid = threadIdx.x;
  const int seg = (S + SCAN_THREADS - 1) / SCAN_THREADS;
  const int s0 = tid * seg, s1 = min(S, s0 + seg);

  // ---- pass 0: choose experts; me = sum_s gates[s,e]
  float me_loc[MOE_MAX_E];
#pragma unroll
  for (int e = 0; e < MOE_MAX_E; ++e) me_loc[e] = 0.0f;
  for (int s = s0; s < s1; ++s) {
    const float* g = gates + static_cast<long long>(s) * E;
    int i1 = 0;
    float best = g[0];
    for (int e = 0; e < E; ++e) {
      if (e < MOE_MAX_E) me_loc[e] += g[e];
      if (g[e] > best) {
        best = g[e];
        i1 = e;
      }
    }
    expert[s * k] = i1;
    if (k == 2) {
      const float* lg = logits + static_cast<long long>(s) * E;
      int i2 = -1;
      float b2 = -INFINITY;
      for (int e = 0; e < E; ++e) {
        if (e == i1) continue;
        const float v = lg[e] + (noise ? noise[static_cast<long long>(s) * E + e] : 0.0f);
        if (i2 < 0 || v > b2) {
          b2 = v;
          i2 = e;
        }
      }
      expert[s * k + 1] = i2;
    }
  }
  for (int e = 0; e < E; ++e) {
    float v = warp_sum(me_loc[e]);
    __syncthreads();
    if ((tid & 31) == 0) red[tid >> 5] = v;
    __syncthreads();
    if (tid < 32) {
      float t = red[tid];
      t = warp_sum(t);
      if (tid == 0) me_sum[e] = t;
    }
  }
  __syncthreads();

  // ---- routes: j = 0 (first choice), j = 1 (second choice; its locations start after all first choices)
  for (int j = 0; j < k; ++j) {
    for (int e = 0; e < E; ++e) cnt[tid][e] = 0;
    for (int s = s0; s < s1; ++s) cnt[tid][expert[s * k + j]] += 1;
    __syncthreads();
    // exclusive scan over threads, one warp per expert
    const int w = tid >> 5, lane = tid & 31;
    if (w < E) {
      int run = 0;
      for (int base = 0; base < SCAN_THREADS; base += 32) {
        const int v = cnt[base + lane][w];
        int inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int n = __shfl_up_sync(0xffffffffu, inc, o);
          if (lane >= o) inc += n;
        }
        cnt[base + lane][w] = run + inc - v;
        run += __shfl_sync(0xffffffffu, inc, 31);
      }
      if (lane == 0) {
        if (j == 0)
          total1[w] = run;
        else
          total2[w] = run;
      }
    }
    __syncthreads();
    // top-1 with Random Token Selection: when an expert overflows, keep the C tokens with the largest uniforms.
    const bool rts = (k == 1 && noise != nullptr);
    if (rts && skey_ok) {
      // keep the C largest uniforms of an overflowing expert (ties: earlier position first): the expert's keys are
      // staged in shared memory and the C-th largest is found by bisection on the fp32 bit pattern (non-negative
      // floats order like unsigned ints): 31 block-wide counts instead of an O(S^2) ranking
      for (int e = 0; e < E; ++e) {
        if (total1[e] <= C) continue;  // uniform branch
        __syncthreads();
        for (int t = tid; t < S; t += SCAN_THREADS)
          skey[t] = expert[t] == e ? noise[static_cast<long long>(t) * E + e] : -1.0f;
        __syncthreads();
        auto count_ge = [&](unsigned v) {
          int c = 0;
          for (int t = tid; t < S; t += SCAN_THREADS) {
            const float kf = skey[t];
            c += (kf >= 0.0f && __float_as_uint(kf) >= v) ? 1 : 0;
          }
          return block_sum_int(c, ired);
        };
        unsigned lo = 0u, hi = 0x7f800000u;  // count_ge(lo) >= C > count_ge(hi)
        while (hi - lo > 1u) {
          const unsigned mid = lo + (hi - lo) / 2u;
          if (count_ge(mid) >= C)
            lo = mid;
          else
            hi = mid;
        }
        const unsigned thr = lo;
        const int n_greater = count_ge(thr + 1u);
        const int n_equal = count_ge(thr) - n_greater;
        int allow = C - n_greater;  // >= 1
        if (n_equal > allow) {      // exact ties at the threshold (rare): the earliest `allow` of them stay
          if (tid == 0) {
            for (int t = 0; t < S; ++t) {
              const float kf = skey[t];
              if (kf >= 0.0f && __float_as_uint(kf) == thr) {
                if (allow > 0)
                  --allow;
                else
                  skey[t] = -0.5f;
              }
            }
          }
          __syncthreads();
        }
        for (int s = s0; s < s1; ++s)
          if (expert[s] == e) {
            const float kf = skey[s];
            slot[s] = (kf >= 0.0f && __float_as_uint(kf) >= thr) ? -2 : -1;  // -2: kept, location resolved below
          }
      }
      __syncthreads();
    }
    int run[MOE_MAX_E];
#pragma unroll
    for (int e = 0; e < MOE_MAX_E; ++e) run[e] = (e < E) ? cnt[tid][e] : 0;
    for (int s = s0; s < s1; ++s) {
      const int e = expert[s * k + j];
      int loc = 0;
#pragma unroll
      for (int q = 0; q < MOE_MAX_E; ++q)
        if (q == e) loc = run[q]++;
      if (j == 1) loc += total1[e];
      bool keep = loc < C;
      if (rts && total1[e] > C) {
        if (skey_ok) continue;  // decided above
        const float u = noise[static_cast<long long>(s) * E + e];
        int rank = 0;
        for (int t = 0; t < S; ++t) {
          if (expert[t] != e || t == s) continue;
          const float ut = noise[static_cast<long long>(t) * E + e];
          if (ut > u || (ut == u && t < s)) ++rank;
        }
        keep = rank < C;
        loc = -2;  // resolved below (locations are the cumsum over KEPT tokens)
      }
      slot[s * k + j] = keep ? loc : -1;
    }
    __syncthreads();
    if (rts) {
      // overflowing experts: locations = cumsum over the KEPT tokens in position order (DeepSpeed recomputes
      // locations1 after the RTS mask) — a second block scan over the kept flags
      bool any = false;
      for (int e = 0; e < E; ++e) any = any || total1[e] > C;
      if (any) {
        for (int e = 0; e < E; ++e) cnt[tid][e] = 0;
        for (int s = s0; s < s1; ++s) {
          const int e = expert[s];
          if (total1[e] > C && slot[s] == -2) cnt[tid][e] += 1;
        }
        __syncthreads();
        if (w < E) {
          int run2 = 0;
          for (int base = 0; base < SCAN_THREADS; base += 32) {
            const int v = cnt[base + lane][w];
            int inc = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
              const int n = __shfl_up_sync(0xffffffffu, inc, o);
              if (lane >= o) inc += n;
            }
            cnt[base + lane][w] = run2 + inc - v;
            run2 += __shfl_sync(0xffffffffu, inc, 31);
          }
        }
        __syncthreads();
        int run2[MOE_MAX_E];
#pragma unroll
        for (int e = 0; e < MOE_MAX_E; ++e) run2[e] = (e < E) ? cnt[tid][e] : 0;
        for (int s = s0; s < s1; ++s) {
          const int e = expert[s];
          if (total1[e] > C && slot[s] == -2) {
#pragma unroll
            for (int q = 0; q < MOE_MAX_E; ++q)
              if (q == e) slot[s] = run2[q]++;
          }
        }
        __syncthreads();
      }
    }
  }
  // ---- finalize: global rows, gate values, counts, l_aux
  for (int s = s0; s < s1; ++s) {
    float gsel[2] = {0.0f, 0.0f};
    for (int j = 0; j < k; ++j) {
      const int e = expert[s * k + j];
      const int loc = slot[s * k + j];
      gsel[j] = loc >= 0 ? gates[static_cast<long long>(s) * E + e] : 0.0f;
      slot[s * k + j] = loc >= 0 ? e * C + loc : -1;
    }
    if (k == 2) {
      const float denom = fmaxf(gsel[0] + gsel[1], 1.1920928955078125e-07f);
      gsel[0] /= denom;
      gsel[1] /= denom;
    }
    for (int j = 0; j < k; ++j) gate[s * k + j] = gsel[j];
  }
  if (tid < E) {
    const int c1 = total1[tid];
    exp_counts[tid] = c1;
    int kc = min(c1, C);
    if (k == 2) kc = min(c1 + total2[tid], C);
    kept[tid] = kc;
  }
  if (tid == 0) {
    float acc = 0.0f;
    for (int e = 0; e < E; ++e) acc += (me_sum[e] / S) * (static_cast<float>(total1[e]) / S);
    *l_aux = acc * E;
  }
}

__global__ void __launch_bounds__(128) moe_dispatch_kernel(const __nv_bfloat16* __restrict__ h, long long ldh,
                                                           const int* __restrict__ slot,
                                                           __nv_bfloat16* __restrict__ xperm, int k, int D) {
  griddep_wait();  // programmatic dependent launch: the previous kernel's outputs are read from here on
  griddep_launch_dependents();
  const int r = blockIdx.x;  // token * k + route
  const int dst = slot[r];
  if (dst < 0) return;
  const uint4* src = reinterpret_cast<const uint4*>(h + static_cast<long long>(r / k) * ldh);
  uint4* out = reinterpret_cast<uint4*>(xperm + static_cast<long long>(dst) * D);
  for (int c = threadIdx.x; c < D / 8; c += 128) out[c] = src[c];
}

// ln_w != NULL: the row just written is also normalised for its next consumer -- h_out[s] = ln_w * bf16(out[s] * rstd),
// LlamaRMSNorm of the NEXT decoder layer's input (or the final norm) -- while it is still in registers (D <= 4096).
__global__ void __launch_bounds__(128) moe_combine_kernel(const __nv_bfloat16* __restrict__ y,
                                                          const int* __restrict__ slot,
                                                          const float* __restrict__ gate,
                                                          const __nv_bfloat16* __restrict__ residual, long long ldr,
                                                          __nv_bfloat16* __restrict__ out, long long ldo, int k,
                                                          int D, const __nv_bfloat16* __restrict__ ln_w, float eps,
                                                          __nv_bfloat16* __restrict__ h_out, long long ldh) {
  griddep_wait();  // programmatic dependent launch: the previous kernel's outputs are read from here on
  griddep_launch_dependents();
  __shared__ float red[4];
  const int s = blockIdx.x;
  int sl[2] = {-1, -1};
  float g[2] = {0.0f, 0.0f};
  for (int j = 0; j < k; ++j) {
    sl[j] = slot[s * k + j];
    g[j] = bf16_round(gate[s * k + j]);  // combine_weights.type_as(input)
  }
  const __nv_bfloat16* rr = residual ? residual + static_cast<long long>(s) * ldr : nullptr;
  __nv_bfloat16* orow = out + static_cast<long long>(s) * ldo;
  uint4 kept[4];  // the row's outputs (bf16) for the fused RMSNorm: 4 x 8 columns per thread = D <= 4096
  float ss = 0.0f;
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int c = (it * 128 + threadIdx.x) * 8;
    if (c >= D) continue;
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.0f;
    for (int j = 0; j < k; ++j) {
      if (sl[j] < 0) continue;
      const uint4 raw = *reinterpret_cast<const uint4*>(y + static_cast<long long>(sl[j]) * D + c);
      const __nv_bfloat162* hp = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 f = __bfloat1622float2(hp[i]);
        acc[2 * i] += g[j] * f.x;
        acc[2 * i + 1] += g[j] * f.y;
      }
    }
    if (rr != nullptr) {
      const uint4 raw = *reinterpret_cast<const uint4*>(rr + c);
      const __nv_bfloat162* hp = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 f = __bfloat1622float2(hp[i]);
        acc[2 * i] = bf16_round(acc[2 * i]) + f.x;
        acc[2 * i + 1] = bf16_round(acc[2 * i + 1]) + f.y;
      }
    }
    uint4 o;
    o.x = pack_bf16(acc[0], acc[1]);
    o.y = pack_bf16(acc[2], acc[3]);
    o.z = pack_bf16(acc[4], acc[5]);
    o.w = pack_bf16(acc[6], acc[7]);
    *reinterpret_cast<uint4*>(orow + c) = o;
    kept[it] = o;
    if (ln_w != nullptr) {
      const __nv_bfloat162* hp = reinterpret_cast<const __nv_bfloat162*>(&o);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 f = __bfloat1622float2(hp[i]);
        ss += f.x * f.x + f.y * f.y;
      }
    }
  }
  if (ln_w == nullptr) return;  // (uniform over the CTA)
  ss = warp_sum(ss);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
  __syncthreads();
  const float tot = ((red[0] + red[1]) + red[2]) + red[3];
  const float rstd = rsqrtf(tot / static_cast<float>(D) + eps);
  __nv_bfloat16* hrow = h_out + static_cast<long long>(s) * ldh;
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int c = (it * 128 + threadIdx.x) * 8;
    if (c >= D) continue;
    const uint4 wraw = *reinterpret_cast<const uint4*>(ln_w + c);
    const __nv_bfloat162* wh = reinterpret_cast<const __nv_bfloat162*>(&wraw);
    const __nv_bfloat162* xh = reinterpret_cast<const __nv_bfloat162*>(&kept[it]);
    uint4 o;
    uint32_t* op = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 wf = __bfloat1622float2(wh[e]), xf = __bfloat1622float2(xh[e]);
      op[e] = pack_bf16(wf.x * bf16_round(xf.x * rstd), wf.y * bf16_round(xf.y * rstd));
    }
    *reinterpret_cast<uint4*>(hrow + c) = o;
  }
}

// ------------------------------------------------------------------------------------------------ small-S fused path
// Decode-time MoE front end for S <= 64 tokens in ONE CTA: [RMSNorm] -> router -> top-k -> slots -> dispatch.
// Replaces five launches (rmsnorm, router, scan, dispatch + the host-visible bookkeeping) of the general path; also
// emits the slot -> token map and per-slot gate values so the expert down-projection can do the combine in its
// epilogue. Slot order = token order (torch.cumsum), like the general path. noise: only the k = 2 Gumbel term.
constexpr int SMALL_S = 64;
constexpr int SMALL_THREADS = 1024;
__global__ void __launch_bounds__(SMALL_THREADS) moe_route_small_kernel(
    const __nv_bfloat16* __restrict__ x, long long ldx, const __nv_bfloat16* __restrict__ ln_w, float eps,
    const float* __restrict__ wg, const float* __restrict__ noise, int S, int D, int E, int k, int C,
    __nv_bfloat16* __restrict__ h, long long ldh, float* __restrict__ logits, float* __restrict__ gates,
    int* __restrict__ expert, float* __restrict__ gate, int* __restrict__ slot, int* __restrict__ kept,
    int* __restrict__ exp_counts, float* __restrict__ l_aux, __nv_bfloat16* __restrict__ xperm,
    int* __restrict__ tok_of_slot, float* __restrict__ gate_of_slot) {
  __shared__ float s_logit[SMALL_S][MOE_MAX_E];
  __shared__ float s_gates[SMALL_S][MOE_MAX_E];
  __shared__ int s_slot[SMALL_S][2];
  __shared__ float s_part[32][MOE_MAX_E + 1];  // per-warp partials: [0] sum of squares, [1+e] router dot products
  griddep_launch_dependents();  // the next kernel (a PDL streaming GEMM) may become resident; it waits for our results
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // ---- phase 1: a group of `wpt` warps per token (the whole row in one or two 16-byte loads per lane when S is
  // small): normalise, store h, router logits from the stored (rounded) values
  int wpt = 16;
  while (wpt > 1 && wpt * S > 32) wpt >>= 1;
  const int per_round = 32 / wpt;
  const int grp = warp / wpt, wig = warp % wpt;
  const int step = wpt * 256;
  for (int s0 = 0; s0 < S; s0 += per_round) {
    const int s = s0 + grp;
    const bool live = s < S && grp < per_round;
    const __nv_bfloat16* xr = x + static_cast<long long>(live ? s : 0) * ldx;
    __nv_bfloat16* hr = h + static_cast<long long>(live ? s : 0) * ldh;
    float rstd = 1.0f;
    if (ln_w != nullptr) {
      float ss = 0.0f;
      if (live) {
#pragma unroll 4
        for (int c = (wig * 32 + lane) * 8; c < D; c += step) {
          const uint4 raw = *reinterpret_cast<const uint4*>(xr + c);
          const __nv_bfloat162* hp = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float2 f = __bfloat1622float2(hp[i]);
            ss += f.x * f.x + f.y * f.y;
          }
        }
      }
      ss = warp_sum(ss);
      if (lane == 0) s_part[warp][0] = ss;
      __syncthreads();
      float tot = 0.0f;
      for (int w = 0; w < wpt; ++w) tot += s_part[grp * wpt + w][0];
      rstd = rsqrtf(tot / static_cast<float>(D) + eps);
    }
    float acc[MOE_MAX_E];
#pragma unroll
    for (int e = 0; e < MOE_MAX_E; ++e) acc[e] = 0.0f;
    if (live) {
#pragma unroll 2
      for (int c = (wig * 32 + lane) * 8; c < D; c += step) {
        const uint4 raw = *reinterpret_cast<const uint4*>(xr + c);
        const __nv_bfloat162* hp = reinterpret_cast<const __nv_bfloat162*>(&raw);
        float v[8];
        if (ln_w != nullptr) {
          const uint4 wraw = *reinterpret_cast<const uint4*>(ln_w + c);
          const __nv_bfloat162* wp = reinterpret_cast<const __nv_bfloat162*>(&wraw);
          uint4 o;
          uint32_t* op = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float2 f = __bfloat1622float2(hp[i]), wf = __bfloat1622float2(wp[i]);
            const float a = bf16_round(wf.x * bf16_round(f.x * rstd)), b = bf16_round(wf.y * bf16_round(f.y * rstd));
            v[2 * i] = a;
            v[2 * i + 1] = b;
            op[i] = pack_bf16(a, b);
          }
          *reinterpret_cast<uint4*>(hr + c) = o;
        } else {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float2 f = __bfloat1622float2(hp[i]);
            v[2 * i] = f.x;
            v[2 * i + 1] = f.y;
          }
          if (hr != xr) *reinterpret_cast<uint4*>(hr + c) = raw;
        }
#pragma unroll
        for (int e = 0; e < MOE_MAX_E; ++e) {
          if (e < E) {
            const float4 w0 = *reinterpret_cast<const float4*>(wg + static_cast<long long>(e) * D + c);
            const float4 w1 = *reinterpret_cast<const float4*>(wg + static_cast<long long>(e) * D + c + 4);
            acc[e] += v[0] * w0.x + v[1] * w0.y + v[2] * w0.z + v[3] * w0.w + v[4] * w1.x + v[5] * w1.y +
                      v[6] * w1.z + v[7] * w1.w;
          }
        }
      }
    }
#pragma unroll
    for (int e = 0; e < MOE_MAX_E; ++e)
      if (e < E) acc[e] = warp_sum(acc[e]);
    __syncthreads();  // s_part[.][0] fully consumed above
    if (lane == 0) {
#pragma unroll
      for (int e = 0; e < MOE_MAX_E; ++e)
        if (e < E) s_part[warp][1 + e] = acc[e];
    }
    __syncthreads();
    if (live && wig == 0 && lane == 0) {
      float lg[MOE_MAX_E];
      float m = -INFINITY;
#pragma unroll
      for (int e = 0; e < MOE_MAX_E; ++e)
        if (e < E) {
          float t = 0.0f;
          for (int w = 0; w < wpt; ++w) t += s_part[grp * wpt + w][1 + e];
          lg[e] = t;
          m = fmaxf(m, t);
        }
      float ex[MOE_MAX_E];
      float sum = 0.0f;
#pragma unroll
      for (int e = 0; e < MOE_MAX_E; ++e)
        if (e < E) {
          ex[e] = expf(lg[e] - m);
          sum += ex[e];
        }
#pragma unroll
      for (int e = 0; e < MOE_MAX_E; ++e)
        if (e < E) {
          const float gv = ex[e] / sum;
          s_logit[s][e] = lg[e];
          s_gates[s][e] = gv;
          logits[static_cast<long long>(s) * E + e] = lg[e];
          gates[static_cast<long long>(s) * E + e] = gv;
        }
    }
    __syncthreads();
  }
  // ---- phase 2: sequential slot assignment (S <= 64)
  if (threadIdx.x == 0) {
    int c1[MOE_MAX_E], c2[MOE_MAX_E];
    float me[MOE_MAX_E];
    for (int e = 0; e < E; ++e) c1[e] = c2[e] = 0, me[e] = 0.0f;
    for (int s = 0; s < S; ++s) {
      int i1 = 0;
      float best = s_gates[s][0];
      for (int e = 0; e < E; ++e) {
        me[e] += s_gates[s][e];
        if (s_gates[s][e] > best) best = s_gates[s][e], i1 = e;
      }
      expert[s * k] = i1;
      s_slot[s][0] = c1[i1]++;
      if (k == 2) {
        int i2 = -1;
        float b2 = -INFINITY;
        for (int e = 0; e < E; ++e) {
          if (e == i1) continue;
          const float v = s_logit[s][e] + (noise ? noise[static_cast<long long>(s) * E + e] : 0.0f);
          if (i2 < 0 || v > b2) b2 = v, i2 = e;
        }
        expert[s * k + 1] = i2;
        s_slot[s][1] = c2[i2]++;
      }
    }
    float aux = 0.0f;
    for (int e = 0; e < E; ++e) {
      exp_counts[e] = c1[e];
      kept[e] = min(k == 2 ? c1[e] + c2[e] : c1[e], C);
      aux += (me[e] / S) * (static_cast<float>(c1[e]) / S);
    }
    *l_aux = aux * E;
    for (int s = 0; s < S; ++s) {
      float gsel[2] = {0.0f, 0.0f};
      int row[2] = {-1, -1};
      for (int j = 0; j < k; ++j) {
        const int e = expert[s * k + j];
        int loc = s_slot[s][j] + (j == 1 ? c1[e] : 0);
        if (loc < C) {
          row[j] = e * C + loc;
          gsel[j] = s_gates[s][e];
        }
      }
      if (k == 2) {
        const float denom = fmaxf(gsel[0] + gsel[1], 1.1920928955078125e-07f);
        gsel[0] /= denom;
        gsel[1] /= denom;
      }
      for (int j = 0; j < k; ++j) {
        s_slot[s][j] = row[j];
        slot[s * k + j] = row[j];
        gate[s * k + j] = gsel[j];
        if (row[j] >= 0 && tok_of_slot != nullptr) {
          tok_of_slot[row[j]] = s;
          gate_of_slot[row[j]] = gsel[j];
        }
      }
    }
  }
  __syncthreads();
  // ---- phase 3: dispatch (rows are L2-hot)
  if (xperm == nullptr) return;  // the expert GEMM gathers its rows through tok_of_slot instead
  for (int r = warp; r < S * k; r += SMALL_THREADS / 32) {
    const int dst = s_slot[r / k][r % k];
    if (dst < 0) continue;
    const uint4* src = reinterpret_cast<const uint4*>(h + static_cast<long long>(r / k) * ldh);
    uint4* out = reinterpret_cast<uint4*>(xperm + static_cast<long long>(dst) * D);
    for (int c = lane; c < D / 8; c += 32) out[c] = src[c];
  }
}

static int moe_scan_launch(const mpl_moe_route_args& a, cudaStream_t stream);

// RMSNorm(x) -> h, router logits / gates, then the scan: what mpl_rmsnorm + moe_route do, in two launches instead of
// three and without re-reading h. Returns MPL_ERR_UNSUPPORTED for widths the register-resident kernel does not cover.
int moe_norm_route(const mpl_moe_route_args& a, const void* x, long long ldx, const void* ln_w, float eps,
                   cudaStream_t stream) {
  if (a.S <= 0) return MPL_OK;
  if (a.D > 4096 || (a.D % 8) != 0 || (a.ldh % 8) != 0 || (ldx % 8) != 0) return MPL_ERR_UNSUPPORTED;
  if (a.h == nullptr || x == nullptr || ln_w == nullptr || a.wg == nullptr || a.logits == nullptr || a.gates == nullptr)
    return MPL_ERR_ARG;
  if (a.E < 1 || a.E > MOE_MAX_E || a.k < 1 || a.k > 2 || a.k > a.E || a.capacity < 1) return MPL_ERR_UNSUPPORTED;
  launch_pdl(moe_norm_router_kernel, dim3(a.S), dim3(128), 0, stream, static_cast<const __nv_bfloat16*>(x), ldx,
                                                 static_cast<const __nv_bfloat16*>(ln_w), eps,
                                                 static_cast<__nv_bfloat16*>(const_cast<void*>(a.h)), a.ldh, a.wg, a.D, a.E,
                                                 a.logits, a.gates);
  const int rc = launch_status();
  if (rc != MPL_OK) return rc;
  return moe_scan_launch(a, stream);
}

int moe_route(const mpl_moe_route_args& a, cudaStream_t stream) {
  if (a.S <= 0) return MPL_OK;
  if (a.h == nullptr || a.wg == nullptr || a.logits == nullptr || a.gates == nullptr || a.expert == nullptr ||
      a.gate == nullptr || a.slot == nullptr || a.kept == nullptr || a.exp_counts == nullptr || a.l_aux == nullptr)
    return MPL_ERR_ARG;
  if (a.E < 1 || a.E > MOE_MAX_E || a.k < 1 || a.k > 2 || a.k > a.E || a.capacity < 1) return MPL_ERR_UNSUPPORTED;
  if ((a.D % 8) != 0 || (a.ldh % 8) != 0) return MPL_ERR_ALIGN;
  launch_pdl(moe_router_kernel, dim3((a.S + 3) / 4), dim3(128), 0, stream, static_cast<const __nv_bfloat16*>(a.h), a.ldh, a.wg, a.S, a.D,
                                                      a.E, a.logits, a.gates);
  const int rc = launch_status();
  if (rc != MPL_OK) return rc;
  return moe_scan_launch(a, stream);
}

static int moe_scan_launch(const mpl_moe_route_args& a, cudaStream_t stream) {
  if (a.expert == nullptr || a.gate == nullptr || a.slot == nullptr || a.kept == nullptr || a.exp_counts == nullptr ||
      a.l_aux == nullptr)
    return MPL_ERR_ARG;
  size_t skey_bytes = 0;
  if (a.k == 1 && a.noise != nullptr && static_cast<size_t>(a.S) * 4 <= 160 * 1024) {
    skey_bytes = static_cast<size_t>(a.S) * 4;
    static bool attr = false;
    if (!attr) {
      if (cudaFuncSetAttribute(moe_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024) != cudaSuccess)
        return MPL_ERR_CUDA;
      attr = true;
    }
  }
  launch_pdl(moe_scan_kernel, dim3(1), dim3(SCAN_THREADS), skey_bytes, stream, a.logits, a.gates, a.noise, a.S, a.E, a.k, a.capacity, a.expert,
                                                           a.gate, a.slot, a.kept, a.exp_counts, a.l_aux,
                                                           skey_bytes > 0 ? 1 : 0);
  return launch_status();
}

int moe_dispatch(const void* h, long long ldh, const int* slot, void* xperm, int S, int k, int D,
                 cudaStream_t stream) {
  if (S <= 0) return MPL_OK;
  if (h == nullptr || slot == nullptr || xperm == nullptr) return MPL_ERR_ARG;
  if ((D % 8) != 0 || (ldh % 8) != 0) return MPL_ERR_ALIGN;
  launch_pdl(moe_dispatch_kernel, dim3(S * k), dim3(128), 0, stream, static_cast<const __nv_bfloat16*>(h), ldh, slot,
                                                 static_cast<__nv_bfloat16*>(xperm), k, D);
  return mpl::launch_status();
}


int moe_combine(const void* y, const int* slot, const float* gate, const void* residual, long long ldr, void* out,
                long long ldo, int S, int k, int D, cudaStream_t stream, const void* ln_w, float eps, void* h_out,
                long long ldh) {
  if (S <= 0) return MPL_OK;
  if (y == nullptr || slot == nullptr || gate == nullptr || out == nullptr) return MPL_ERR_ARG;
  if ((D % 8) != 0 || (ldr % 8) != 0 || (ldo % 8) != 0 || k < 1 || k > 2 || D > 4096) return MPL_ERR_ALIGN;
  if (ln_w != nullptr && (h_out == nullptr || (ldh % 8) != 0)) return MPL_ERR_ARG;
  launch_pdl(moe_combine_kernel, dim3(S), dim3(128), 0, stream, static_cast<const __nv_bfloat16*>(y), slot, gate,
             static_cast<const __nv_bfloat16*>(residual), ldr, static_cast<__nv_bfloat16*>(out), ldo, k, D,
             static_cast<const __nv_bfloat16*>(ln_w), eps, static_cast<__nv_bfloat16*>(h_out), ldh);
  return mpl::launch_status();
}

int moe_route_small(const mpl_moe_route_args& a, const void* x, long long ldx, const void* ln_w, float eps, void* h,
                    long long ldh, void* xperm, int* tok_of_slot, float* gate_of_slot, cudaStream_t stream) {
  if (a.S <= 0) return MPL_OK;
  if (a.S > SMALL_S || (a.k == 1 && a.noise != nullptr)) return MPL_ERR_UNSUPPORTED;
  if (x == nullptr || h == nullptr || a.wg == nullptr || a.logits == nullptr ||
      a.gates == nullptr || a.expert == nullptr || a.gate == nullptr || a.slot == nullptr || a.kept == nullptr ||
      a.exp_counts == nullptr || a.l_aux == nullptr)
    return MPL_ERR_ARG;
  if (a.E < 1 || a.E > MOE_MAX_E || a.k < 1 || a.k > 2 || a.k > a.E || a.capacity < 1) return MPL_ERR_UNSUPPORTED;
  if ((a.D % 8) != 0 || (ldx % 8) != 0 || (ldh % 8) != 0) return MPL_ERR_ALIGN;
  moe_route_small_kernel<<<1, SMALL_THREADS, 0, stream>>>(static_cast<const __nv_bfloat16*>(x), ldx,
                                               static_cast<const __nv_bfloat16*>(ln_w), eps, a.wg, a.noise, a.S, a.D,
                                               a.E, a.k, a.capacity, static_cast<__nv_bfloat16*>(h), ldh, a.logits,
                                               a.gates, a.expert, a.gate, a.slot, a.kept, a.exp_counts, a.l_aux,
                                               static_cast<__nv_bfloat16*>(xperm), tok_of_slot, gate_of_slot);
  return launch_status();
}

}  // namespace mpl

extern "C" int mpl_moe_route_small(const mpl_moe_route_args* a, const void* x, long long ldx, const void* ln_weight,
                                   float ln_eps, void* h, long long ldh, void* xperm, int* tok_of_slot,
                                   float* gate_of_slot, void* stream) {
  if (a == nullptr) return MPL_ERR_ARG;
  return mpl::moe_route_small(*a, x, ldx, ln_weight, ln_eps, h, ldh, xperm, tok_of_slot, gate_of_slot,
                              static_cast<cudaStream_t>(stream));
}

extern "C" int mpl_moe_route(const mpl_moe_route_args* a, void* stream) {
  if (a == nullptr) return MPL_ERR_ARG;
  return mpl::moe_route(*a, static_cast<cudaStream_t>(stream));
}
extern "C" int mpl_moe_dispatch(const void* h, long long ldh, const int* slot, void* xperm, int S, int k, int D,
                                void* stream) {
  return mpl::moe_dispatch(h, ldh, slot, xperm, S, k, D, static_cast<cudaStream_t>(stream));
}
extern "C" int mpl_moe_combine(const void* y, const int* slot, const float* gate, const void* residual, long long ldr,
                               void* out, long long ldo, int S, int k, int D, void* stream) {
  return mpl::moe_combine(y, slot, gate, residual, ldr, out, ldo, S, k, D, static_cast<cudaStream_t>(stream));
}
