// Geometric region sampler: the non-GEMM steps of GeoRegionSampler.forward
// (model/rp_sampler/GeoSampler.py:229-345) as coalesced HBM kernels with warp-shuffle reductions; the two grouped
// Linears per stage, flatten_projector and dim_projector run on the tcgen05 / streaming GEMMs (gemm_tcgen05.cu).
//
// Data layout: a stage's point set is ONE bf16 "point table" [R regions, N points, ld] whose row is
//   [ d feature columns | row/H | col/W | zero padding up to ld ]        (ld = round_up(d + 2, 64))
// i.e. exactly the rows torch.cat([fea, xy], -1) builds at :302-303, padded so that they are GEMM operands as they are.
//   geo_point_table : point_sample (:31-56, 263-276) -> stage-0 table         one CTA per (region, point)
//   geo_fps         : farthest_point_sample (:59-80)                          one 8-warp CTA per region, points in registers
//   geo_knn         : square_distance + topk (:101-136)                       one warp per anchor, k rounds of a
//                                                                             lexicographic (distance, index) warp-min
//   geo_group       : local - anchor | anchor (:302-308)  -> GEMM operands    one CTA per (region, anchor, neighbour)
//   geo_ln_pool     : ReLU -> LayerNorm -> Avg/Max pool over the k neighbours (:139-157, 317) -> next stage's table
// Rounding points are the eager bf16 reference's: coordinates and every elementwise result rounded to bf16, distances
// as oracle/geo.py::fps_dist_bf16 / knn_dist_bf16 write them out, LayerNorm in fp32 rounded once, pooling in fp32.
// Ties: FPS takes the first maximum (torch.max on CPU); kNN takes the k smallest by (distance, index) in that order.
#include "internal.h"
#include "ptx.cuh"

namespace mpl {

typedef __nv_bfloat16 bf;

// ---------------------------------------------------------------------------------------------------- point table
// fmap bf16 [n_img, h*w, C]; pts f32 [R, P, 2] = (row / H, col / W); img_of_region int [R].
__global__ void __launch_bounds__(128) geo_point_table_kernel(const bf* __restrict__ fmap, const int* __restrict__ img_of,
                                                              const float* __restrict__ pts, int P, int h, int w, int C,
                                                              bf* __restrict__ table, int ld) {
  const int rp = blockIdx.x, r = rp / P;
  const float py = bf16_round(pts[2 * rp]), px = bf16_round(pts[2 * rp + 1]);  // .type(original_dtype), flipped to (x, y)
  // (2.0 * point_coords - 1.0) is formed on the bf16 coordinates before point_sample's .float() (:51): a bf16 grid
  const float gx = bf16_round(2.0f * px - 1.0f), gy = bf16_round(2.0f * py - 1.0f);
  const float fx = (gx + 1.0f) * 0.5f * (w - 1), fy = (gy + 1.0f) * 0.5f * (h - 1);
  const int x0 = static_cast<int>(floorf(fx)), y0 = static_cast<int>(floorf(fy));
  const float lx = fx - x0, ly = fy - y0;
  const bf* base = fmap + static_cast<long long>(img_of[r]) * h * w * C;
  bf* out = table + static_cast<long long>(rp) * ld;
  for (int c = threadIdx.x * 8; c < ld; c += 128 * 8) {
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = 0.0f;
    if (c < C) {
      // same accumulation order as grid_sample's CPU kernel is not guaranteed; fp32 sums of four products, rounded once
#pragma unroll
      for (int dy = 0; dy < 2; ++dy)
#pragma unroll
        for (int dx = 0; dx < 2; ++dx) {
          const int xx = x0 + dx, yy = y0 + dy;
          if (xx < 0 || xx >= w || yy < 0 || yy >= h) continue;
          const float wgt = (dx ? lx : 1.0f - lx) * (dy ? ly : 1.0f - ly);
          const uint4 raw = *reinterpret_cast<const uint4*>(base + (static_cast<long long>(yy) * w + xx) * C + c);
          const bf* hv = reinterpret_cast<const bf*>(&raw);
#pragma unroll
          for (int e = 0; e < 8; ++e) v[e] += wgt * __bfloat162float(hv[e]);
        }
    }
    uint4 o;
    bf* ov = reinterpret_cast<bf*>(&o);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int col = c + e;
      ov[e] = __float2bfloat16_rn(col < C ? v[e] : (col == C ? py : (col == C + 1 ? px : 0.0f)));
    }
    *reinterpret_cast<uint4*>(out + c) = o;
  }
}

// ---------------------------------------------------------------------------------------------------- FPS
// xy = table + d (row pitch ld); one CTA of 8 warps per region, N <= 1024 points: thread t holds points t, t + 256, ...
// The S iterations are one dependent chain (the next centroid is the arg-max of the running minimum distance): per
// iteration a thread updates its <= 4 distances, the warp reduces (value, index) by shuffles -- first maximum: the lower
// index wins a tie --, the eight warp results meet in a double-buffered shared-memory slot (ONE CTA barrier per
// iteration) and every thread reduces them redundantly. (One warp per region with 32 points per lane took 1.7 us per
// iteration, 217 us for S = 128: profiles/r02_ncu_geo.md.)
constexpr int GEO_FPS_THREADS = 256;
constexpr int GEO_FPS_PPT = 4;  // points per thread
__global__ void __launch_bounds__(GEO_FPS_THREADS) geo_fps_kernel(const bf* __restrict__ xy, long long ld, int N, int S,
                                                                  const int* __restrict__ start, int* __restrict__ fps_idx) {
  __shared__ float2 s_xy[GEO_FPS_THREADS * GEO_FPS_PPT];
  __shared__ float s_best[2][GEO_FPS_THREADS / 32];
  __shared__ int s_bidx[2][GEO_FPS_THREADS / 32];
  const int r = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bf* base = xy + static_cast<long long>(r) * N * ld;
  float px[GEO_FPS_PPT], py[GEO_FPS_PPT], dist[GEO_FPS_PPT];
#pragma unroll
  for (int i = 0; i < GEO_FPS_PPT; ++i) {
    const int n = i * GEO_FPS_THREADS + tid;
    px[i] = n < N ? __bfloat162float(base[n * ld]) : 0.0f;
    py[i] = n < N ? __bfloat162float(base[n * ld + 1]) : 0.0f;
    dist[i] = 1e10f;
    s_xy[n] = make_float2(px[i], py[i]);
  }
  __syncthreads();
  int far = start[r];
  for (int s = 0; s < S; ++s) {
    if (tid == 0) fps_idx[r * S + s] = far;
    const float2 c = s_xy[far];
    float best = -2.0f;
    int bi = 0x7fffffff;
#pragma unroll
    for (int i = 0; i < GEO_FPS_PPT; ++i) {
      const int n = i * GEO_FPS_THREADS + tid;
      if (n < N) {
        const float dx = bf16_round(px[i] - c.x), dy = bf16_round(py[i] - c.y);
        const float d = bf16_round(bf16_round(dx * dx) + bf16_round(dy * dy));
        dist[i] = fminf(dist[i], d);
        if (dist[i] > best) {  // ascending n within the thread: strict > keeps the first maximum
          best = dist[i];
          bi = n;
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ob > best || (ob == best && oi < bi)) {
        best = ob;
        bi = oi;
      }
    }
    const int buf = s & 1;
    if (lane == 0) {
      s_best[buf][warp] = best;
      s_bidx[buf][warp] = bi;
    }
    __syncthreads();  // (the other buffer is rewritten only after the NEXT barrier: every thread has read this one by then)
    best = s_best[buf][0];
    bi = s_bidx[buf][0];
#pragma unroll
    for (int q = 1; q < GEO_FPS_THREADS / 32; ++q) {
      const float ob = s_best[buf][q];
      const int oi = s_bidx[buf][q];
      if (ob > best || (ob == best && oi < bi)) {
        best = ob;
        bi = oi;
      }
    }
    far = bi;
  }
}

// ---------------------------------------------------------------------------------------------------- kNN
// one warp per anchor (r, s); the k smallest of N by (distance, index), written in that order.
constexpr int GEO_PPL = 32;  // candidates per lane: N <= 32 * GEO_PPL
__global__ void __launch_bounds__(128) geo_knn_kernel(const bf* __restrict__ xy, long long ld, int N, int S, int k,
                                                      const int* __restrict__ fps_idx, int* __restrict__ knn_idx,
                                                      int anchors) {
  const int a = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (a >= anchors) return;
  const int r = a / S;
  const bf* base = xy + static_cast<long long>(r) * N * ld;
  const int ai = fps_idx[a];
  const float qx = __bfloat162float(base[ai * ld]), qy = __bfloat162float(base[ai * ld + 1]);
  const float sq = bf16_round(bf16_round(qx * qx) + bf16_round(qy * qy));
  unsigned long long key[GEO_PPL];
#pragma unroll
  for (int i = 0; i < GEO_PPL; ++i) {
    const int n = i * 32 + lane;
    key[i] = ~0ull;
    if (n < N) {
      const float x = __bfloat162float(base[n * ld]), y = __bfloat162float(base[n * ld + 1]);
      const float dot = bf16_round(qx * x + qy * y);  // products of bf16 pairs are exact in fp32: one rounding, as the matmul
      const float sx = bf16_round(bf16_round(x * x) + bf16_round(y * y));
      const float d = bf16_round(bf16_round(-2.0f * dot + sq) + sx);
      unsigned int u = __float_as_uint(d);
      u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);  // order-preserving map of the float (the distance may round below 0)
      key[i] = (static_cast<unsigned long long>(u) << 32) | static_cast<unsigned int>(n);
    }
  }
  for (int j = 0; j < k; ++j) {
    unsigned long long best = ~0ull;
#pragma unroll
    for (int i = 0; i < GEO_PPL; ++i) best = key[i] < best ? key[i] : best;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long ob = __shfl_xor_sync(0xffffffffu, best, o);
      best = ob < best ? ob : best;
    }
    const int n = static_cast<int>(best & 0xffffffffu);
    if (lane == 0) knn_idx[static_cast<long long>(a) * k + j] = n;
    if ((n & 31) == lane) {
#pragma unroll
      for (int i = 0; i < GEO_PPL; ++i)
        if (i == (n >> 5)) key[i] = ~0ull;
    }
  }
}

// ---------------------------------------------------------------------------------------------------- grouping
// row (r, s, j):  a1[row, :] = table[r, knn] - table[r, fps]  (bf16, the padding stays 0),  a2[row, ld:2ld] = table[r, fps].
__global__ void __launch_bounds__(128) geo_group_kernel(const bf* __restrict__ table, int ld, int N, int S, int k,
                                                        const int* __restrict__ fps_idx, const int* __restrict__ knn_idx,
                                                        bf* __restrict__ a1, bf* __restrict__ a2) {
  const long long row = blockIdx.x;
  const int a = static_cast<int>(row / k), r = a / S;
  const bf* loc = table + (static_cast<long long>(r) * N + knn_idx[row]) * ld;
  const bf* anc = table + (static_cast<long long>(r) * N + fps_idx[a]) * ld;
  bf* o1 = a1 + row * ld;
  bf* o2 = a2 + row * 2 * ld + ld;
  for (int c = threadIdx.x * 8; c < ld; c += 128 * 8) {
    const uint4 lr = *reinterpret_cast<const uint4*>(loc + c), ar = *reinterpret_cast<const uint4*>(anc + c);
    const bf* lv = reinterpret_cast<const bf*>(&lr);
    const bf* av = reinterpret_cast<const bf*>(&ar);
    uint4 o;
    bf* ov = reinterpret_cast<bf*>(&o);
#pragma unroll
    for (int e = 0; e < 8; ++e) ov[e] = __float2bfloat16_rn(__bfloat162float(lv[e]) - __bfloat162float(av[e]));
    *reinterpret_cast<uint4*>(o1 + c) = o;
    *reinterpret_cast<uint4*>(o2 + c) = ar;
  }
}

// ---------------------------------------------------------------------------------------------------- LN + pool
// y bf16 [anchors, k, D] (ReLU already applied by the GEMM epilogue) -> LayerNorm(D) per row -> pool over k ->
// out[anchor, 0:D] (row pitch ldo); with xy_src != NULL also out[anchor, D:D+2] = the anchor's coordinates and zeros up
// to ldo (the next stage's point table). mode 0: AvgPool1d(k), 1: AdaptiveMaxPool1d(1).
constexpr int GEO_LN_THREADS = 128;
constexpr int GEO_LN_VEC = 4;  // D <= 128 * 8 * 4
__global__ void __launch_bounds__(GEO_LN_THREADS) geo_ln_pool_kernel(const bf* __restrict__ y, int k, int D,
                                                                     const bf* __restrict__ w, const bf* __restrict__ b,
                                                                     float eps, int mode, const bf* __restrict__ xy_src,
                                                                     long long ld_src, int N, int S,
                                                                     const int* __restrict__ fps_idx,
                                                                     bf* __restrict__ out, long long ldo) {
  __shared__ float red[2][GEO_LN_THREADS / 32];
  const int a = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float acc[GEO_LN_VEC][8];
#pragma unroll
  for (int i = 0; i < GEO_LN_VEC; ++i)
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[i][e] = mode == 0 ? 0.0f : -INFINITY;
  for (int j = 0; j < k; ++j) {
    const bf* yr = y + (static_cast<long long>(a) * k + j) * D;
    float v[GEO_LN_VEC][8];
    float s = 0.0f;
#pragma unroll
    for (int i = 0; i < GEO_LN_VEC; ++i) {
      const int c = (i * GEO_LN_THREADS + threadIdx.x) * 8;
      if (c < D) {
        const uint4 raw = *reinterpret_cast<const uint4*>(yr + c);
        const bf* hv = reinterpret_cast<const bf*>(&raw);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          v[i][e] = __bfloat162float(hv[e]);
          s += v[i][e];
        }
      }
    }
    s = warp_sum(s);
    if (lane == 0) red[0][warp] = s;
    __syncthreads();
    float tot = 0.0f;
#pragma unroll
    for (int q = 0; q < GEO_LN_THREADS / 32; ++q) tot += red[0][q];
    const float mean = tot / static_cast<float>(D);
    float sq = 0.0f;
#pragma unroll
    for (int i = 0; i < GEO_LN_VEC; ++i) {
      const int c = (i * GEO_LN_THREADS + threadIdx.x) * 8;
      if (c < D) {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float dlt = v[i][e] - mean;
          sq += dlt * dlt;
        }
      }
    }
    sq = warp_sum(sq);
    if (lane == 0) red[1][warp] = sq;
    __syncthreads();
    float tq = 0.0f;
#pragma unroll
    for (int q = 0; q < GEO_LN_THREADS / 32; ++q) tq += red[1][q];
    const float rstd = rsqrtf(tq / static_cast<float>(D) + eps);
#pragma unroll
    for (int i = 0; i < GEO_LN_VEC; ++i) {
      const int c = (i * GEO_LN_THREADS + threadIdx.x) * 8;
      if (c < D) {
        const uint4 wr = *reinterpret_cast<const uint4*>(w + c), br = *reinterpret_cast<const uint4*>(b + c);
        const bf* wv = reinterpret_cast<const bf*>(&wr);
        const bf* bv = reinterpret_cast<const bf*>(&br);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float o = bf16_round((v[i][e] - mean) * rstd * __bfloat162float(wv[e]) + __bfloat162float(bv[e]));
          acc[i][e] = mode == 0 ? acc[i][e] + o : fmaxf(acc[i][e], o);
        }
      }
    }
    // red[0] is rewritten only after the next row's first __syncthreads-separated read of red[1]: no hazard across rows
  }
  bf* orow = out + static_cast<long long>(a) * ldo;
  const float inv_k = 1.0f / static_cast<float>(k);
#pragma unroll
  for (int i = 0; i < GEO_LN_VEC; ++i) {
    const int c = (i * GEO_LN_THREADS + threadIdx.x) * 8;
    if (c < D) {
      uint4 o;
      bf* ov = reinterpret_cast<bf*>(&o);
#pragma unroll
      for (int e = 0; e < 8; ++e) ov[e] = __float2bfloat16_rn(mode == 0 ? acc[i][e] * inv_k : acc[i][e]);
      *reinterpret_cast<uint4*>(orow + c) = o;
    }
  }
  if (xy_src != nullptr) {
    const int r = a / S;
    const bf* src = xy_src + (static_cast<long long>(r) * N + fps_idx[a]) * ld_src;
    for (int c = D + threadIdx.x; c < ldo; c += GEO_LN_THREADS)
      orow[c] = c < D + 2 ? src[c - D] : __float2bfloat16_rn(0.0f);
  }
}

// Warp-per-row form for D <= 1024 (the CLIP-L width): each of the CTA's four warps takes every fourth neighbour row, holds
// it in registers (4 x 16-byte loads per lane, all in flight), gets its LayerNorm statistics from warp shuffles alone (no
// CTA barrier per row: the row-at-a-time kernel above walks k rows through two barriers each and ran at 0.11 of the HBM
// peak, profiles/r02_ncu_geo.md) and pools into registers; the four partial pools meet once in shared memory.
__global__ void __launch_bounds__(GEO_LN_THREADS) geo_ln_pool_warp_kernel(const bf* __restrict__ y, int k, int D,
                                                                          const bf* __restrict__ w, const bf* __restrict__ b,
                                                                          float eps, int mode, const bf* __restrict__ xy_src,
                                                                          long long ld_src, int N, int S,
                                                                          const int* __restrict__ fps_idx,
                                                                          bf* __restrict__ out, long long ldo) {
  __shared__ float part[GEO_LN_THREADS / 32][1024];
  const int a = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float acc[4][8];
  float wv[4][8], bv[4][8];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = (i * 32 + lane) * 8;
    uint4 wr = make_uint4(0, 0, 0, 0), br = wr;
    if (c < D) {
      wr = *reinterpret_cast<const uint4*>(w + c);
      br = *reinterpret_cast<const uint4*>(b + c);
    }
    const bf* wp = reinterpret_cast<const bf*>(&wr);
    const bf* bp = reinterpret_cast<const bf*>(&br);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      wv[i][e] = __bfloat162float(wp[e]);
      bv[i][e] = __bfloat162float(bp[e]);
      acc[i][e] = mode == 0 ? 0.0f : -INFINITY;
    }
  }
  const float inv_d = 1.0f / static_cast<float>(D);
  for (int j = warp; j < k; j += GEO_LN_THREADS / 32) {
    const bf* yr = y + (static_cast<long long>(a) * k + j) * D;
    uint4 raw[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int c = (i * 32 + lane) * 8;
      raw[i] = c < D ? *reinterpret_cast<const uint4*>(yr + c) : make_uint4(0, 0, 0, 0);
    }
    float v[4][8];
    float s = 0.0f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const bf* hv = reinterpret_cast<const bf*>(&raw[i]);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        v[i][e] = __bfloat162float(hv[e]);
        s += v[i][e];
      }
    }
    const float mean = warp_sum(s) * inv_d;
    float sq = 0.0f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if ((i * 32 + lane) * 8 < D) {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float dlt = v[i][e] - mean;
          sq += dlt * dlt;
        }
      }
    }
    const float rstd = rsqrtf(warp_sum(sq) * inv_d + eps);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float o = bf16_round((v[i][e] - mean) * rstd * wv[i][e] + bv[i][e]);
        acc[i][e] = mode == 0 ? acc[i][e] + o : fmaxf(acc[i][e], o);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = (i * 32 + lane) * 8;
    if (c < D) {
#pragma unroll
      for (int e = 0; e < 8; ++e) part[warp][c + e] = acc[i][e];
    }
  }
  __syncthreads();
  bf* orow = out + static_cast<long long>(a) * ldo;
  const float inv_k = 1.0f / static_cast<float>(k);
  for (int c = threadIdx.x; c < D; c += GEO_LN_THREADS) {
    float r = part[0][c];
#pragma unroll
    for (int q = 1; q < GEO_LN_THREADS / 32; ++q) r = mode == 0 ? r + part[q][c] : fmaxf(r, part[q][c]);
    orow[c] = __float2bfloat16_rn(mode == 0 ? r * inv_k : r);
  }
  if (xy_src != nullptr) {
    const int r = a / S;
    const bf* src = xy_src + (static_cast<long long>(r) * N + fps_idx[a]) * ld_src;
    for (int c = D + threadIdx.x; c < ldo; c += GEO_LN_THREADS)
      orow[c] = c < D + 2 ? src[c - D] : __float2bfloat16_rn(0.0f);
  }
}

}  // namespace mpl

using namespace mpl;

extern "C" int mpl_geo_point_table(const void* fmap, const int* img_of_region, const float* pts, int R, int P, int h,
                                   int w, int C, void* table, int ld, void* stream) {
  if (R <= 0 || P <= 0) return MPL_OK;
  if (fmap == nullptr || img_of_region == nullptr || pts == nullptr || table == nullptr) return MPL_ERR_ARG;
  if ((C % 8) != 0 || (ld % 8) != 0 || ld < C + 2) return MPL_ERR_ALIGN;
  geo_point_table_kernel<<<R * P, 128, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const bf*>(fmap), img_of_region, pts, P, h, w, C, static_cast<bf*>(table), ld);
  return launch_status();
}

extern "C" int mpl_geo_fps(const void* xy, long long ld, int R, int N, int S, const int* start, int* fps_idx,
                           void* stream) {
  if (R <= 0 || S <= 0) return MPL_OK;
  if (xy == nullptr || start == nullptr || fps_idx == nullptr || N <= 0) return MPL_ERR_ARG;
  if (N > GEO_FPS_THREADS * GEO_FPS_PPT) return MPL_ERR_UNSUPPORTED;
  geo_fps_kernel<<<R, GEO_FPS_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const bf*>(xy), ld, N, S, start,
                                                                             fps_idx);
  return launch_status();
}

extern "C" int mpl_geo_knn(const void* xy, long long ld, int R, int N, int S, int k, const int* fps_idx, int* knn_idx,
                           void* stream) {
  if (R <= 0 || S <= 0 || k <= 0) return MPL_OK;
  if (xy == nullptr || fps_idx == nullptr || knn_idx == nullptr || k > N) return MPL_ERR_ARG;
  if (N > 32 * GEO_PPL) return MPL_ERR_UNSUPPORTED;
  const int anchors = R * S;
  geo_knn_kernel<<<(anchors + 3) / 4, 128, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const bf*>(xy), ld, N, S, k,
                                                                                 fps_idx, knn_idx, anchors);
  return launch_status();
}

extern "C" int mpl_geo_group(const void* table, int ld, int R, int N, int S, int k, const int* fps_idx,
                             const int* knn_idx, void* a1, void* a2, void* stream) {
  if (R <= 0 || S <= 0 || k <= 0) return MPL_OK;
  if (table == nullptr || fps_idx == nullptr || knn_idx == nullptr || a1 == nullptr || a2 == nullptr) return MPL_ERR_ARG;
  if ((ld % 8) != 0) return MPL_ERR_ALIGN;
  geo_group_kernel<<<R * S * k, 128, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const bf*>(table), ld, N, S, k, fps_idx, knn_idx, static_cast<bf*>(a1), static_cast<bf*>(a2));
  return launch_status();
}

extern "C" int mpl_geo_ln_pool(const void* y, int R, int S, int k, int D, const void* weight, const void* bias, float eps,
                               int mode, const void* xy_src, long long ld_src, int N, const int* fps_idx, void* out,
                               long long ldo, void* stream) {
  if (R <= 0 || S <= 0 || k <= 0) return MPL_OK;
  if (y == nullptr || weight == nullptr || bias == nullptr || out == nullptr || (mode != 0 && mode != 1)) return MPL_ERR_ARG;
  if (xy_src != nullptr && (fps_idx == nullptr || ldo < D + 2)) return MPL_ERR_ARG;
  if ((D % 8) != 0 || (ldo % 8) != 0) return MPL_ERR_ALIGN;
  if (D > GEO_LN_THREADS * 8 * GEO_LN_VEC) return MPL_ERR_UNSUPPORTED;
  if (D <= 1024 && k >= GEO_LN_THREADS / 32) {  // (max pooling over fewer rows than warps would pool a -inf partial: row form)
    geo_ln_pool_warp_kernel<<<R * S, GEO_LN_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const bf*>(y), k, D, static_cast<const bf*>(weight), static_cast<const bf*>(bias), eps, mode,
        static_cast<const bf*>(xy_src), ld_src, N, S, fps_idx, static_cast<bf*>(out), ldo);
    return launch_status();
  }
  geo_ln_pool_kernel<<<R * S, GEO_LN_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const bf*>(y), k, D, static_cast<const bf*>(weight), static_cast<const bf*>(bias), eps, mode,
      static_cast<const bf*>(xy_src), ld_src, N, S, fps_idx, static_cast<bf*>(out), ldo);
  return launch_status();
}
