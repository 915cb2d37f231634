// HBM-bound data-movement / elementwise kernels of the hot path (coalesced 16-byte accesses, warp-shuffle reductions).
//   rope_kv        : HF-4.31 apply_rotary_pos_emb on q,k in place + KV-cache append   
//   gather_rows    : multimodal splice / embedding lookup / window (un)partition / pixel shuffle as one row gather
//                    (model/medplib/model/medplib_arch.py:296-527; image_encoder.py:299-345)
//   argmax         : greedy next-token selection over fp32 logits (HF greedy_search)
//   im2col_patch   : non-overlapping patch extraction for the CLIP / SAM patch-embedding convs
//   im2col_nhwc    : general kxk/stride/pad im2col on token-major (NHWC) activations, optional per-channel gate
//                    (Adapter_Layer channel gate + conv3x3 s2, image_encoder.py:44-46; neck conv3x3 :145-149;
//                    MaskTokenEncoder convs medplib_arch.py:84-93)
//   clip_embed     : [cls ; patches] + position embeddings (CLIPVisionEmbeddings, App. A.2)
//   sam_relpos     : decomposed relative-position terms q.Rh, q.Rw (image_encoder.py:381-421)
//   col_mean       : AdaptiveAvgPool2d(1) over tokens (Adapter_Layer :43-45)
//   convt4s2_col2im: col2im of ConvTranspose2d(k4,s2,p1) computed as a GEMM + ReLU + skip (Adapter_Layer :33-36,48-51)
//   add            : bf16 + {bf16|f32} -> bf16 (token / key positional-encoding adds, transformer.py:160-175)
//   bilinear_resize: F.interpolate(mode="bilinear", align_corners=False) of postprocess_masks (MedPLIB.py:682-701)
//   region_sample  : point_sample + masked mean of extract_region_feature (medplib_arch.py:39-64,580-614)
#include "internal.h"
#include "ptx.cuh"

namespace mpl {

// ------------------------------------------------------------------------------------------------- rope + kv append
// grid = B*T rows, block = 256: thread -> (head, 8-element chunk of the first half).
__global__ void __launch_bounds__(256) rope_kv_kernel(__nv_bfloat16* __restrict__ q, __nv_bfloat16* __restrict__ k,
                                                      const __nv_bfloat16* __restrict__ v, long long ld,
                                                      const __nv_bfloat16* __restrict__ cos_t,
                                                      const __nv_bfloat16* __restrict__ sin_t,
                                                      __nv_bfloat16* __restrict__ kc, __nv_bfloat16* __restrict__ vc,
                                                      int T, int H, int hd, int Tmax, int pos0,
                                                      const int* __restrict__ pos_dev) {
  griddep_wait();  // programmatic dependent launch: the previous kernel's outputs are read from here on
  griddep_launch_dependents();
  const int row = blockIdx.x;
  const int b = row / T, t = row % T;
  const int pos = (pos_dev ? *pos_dev : pos0) + t;
  const int half = hd / 2;
  const int cph = half / 8;  // 16-byte chunks per half head
  const __nv_bfloat16* cr = cos_t + static_cast<long long>(pos) * hd;
  const __nv_bfloat16* sr = sin_t + static_cast<long long>(pos) * hd;
  for (int w = threadIdx.x; w < H * cph; w += 256) {
    const int h = w / cph, c = (w % cph) * 8;
    const long long off = static_cast<long long>(row) * ld + h * hd + c;
    const uint4 c1 = *reinterpret_cast<const uint4*>(cr + c), c2 = *reinterpret_cast<const uint4*>(cr + half + c);
    const uint4 s1 = *reinterpret_cast<const uint4*>(sr + c), s2 = *reinterpret_cast<const uint4*>(sr + half + c);
    const __nv_bfloat16* c1p = reinterpret_cast<const __nv_bfloat16*>(&c1);
    const __nv_bfloat16* c2p = reinterpret_cast<const __nv_bfloat16*>(&c2);
    const __nv_bfloat16* s1p = reinterpret_cast<const __nv_bfloat16*>(&s1);
    const __nv_bfloat16* s2p = reinterpret_cast<const __nv_bfloat16*>(&s2);
    const long long coff = kc ? ((static_cast<long long>(b) * H + h) * Tmax + pos) * hd + c : 0;
#pragma unroll
    for (int which = 0; which < 2; ++which) {
      __nv_bfloat16* base = which == 0 ? q : k;
      if (base == nullptr) continue;
      const uint4 a1 = *reinterpret_cast<const uint4*>(base + off);
      const uint4 a2 = *reinterpret_cast<const uint4*>(base + off + half);
      const __nv_bfloat16* x1 = reinterpret_cast<const __nv_bfloat16*>(&a1);
      const __nv_bfloat16* x2 = reinterpret_cast<const __nv_bfloat16*>(&a2);
      uint4 o1, o2;
      __nv_bfloat16* y1 = reinterpret_cast<__nv_bfloat16*>(&o1);
      __nv_bfloat16* y2 = reinterpret_cast<__nv_bfloat16*>(&o2);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float f1 = __bfloat162float(x1[i]), f2 = __bfloat162float(x2[i]);
        // (x * cos) + (rotate_half(x) * sin), every product and the sum rounded to bf16 like the eager reference
        const float p1 = bf16_round(f1 * __bfloat162float(c1p[i])), r1 = bf16_round(-f2 * __bfloat162float(s1p[i]));
        const float p2 = bf16_round(f2 * __bfloat162float(c2p[i])), r2 = bf16_round(f1 * __bfloat162float(s2p[i]));
        y1[i] = __float2bfloat16_rn(p1 + r1);
        y2[i] = __float2bfloat16_rn(p2 + r2);
      }
      *reinterpret_cast<uint4*>(base + off) = o1;
      *reinterpret_cast<uint4*>(base + off + half) = o2;
      if (which == 1 && kc != nullptr) {
        *reinterpret_cast<uint4*>(kc + coff) = o1;
        *reinterpret_cast<uint4*>(kc + coff + half) = o2;
      }
    }
    if (vc != nullptr && v != nullptr) {
      *reinterpret_cast<uint4*>(vc + coff) = *reinterpret_cast<const uint4*>(v + off);
      *reinterpret_cast<uint4*>(vc + coff + half) = *reinterpret_cast<const uint4*>(v + off + half);
    }
  }
}

// ------------------------------------------------------------------------------------------------- gather rows
// idx >= 0: table row; idx <= -2: feats row (-idx - 2); idx == -1: zeros. D % 8 == 0.
__global__ void __launch_bounds__(128) gather_rows_kernel(const __nv_bfloat16* __restrict__ table, long long ldt,
                                                          const __nv_bfloat16* __restrict__ feats, long long ldf,
                                                          const int* __restrict__ idx, __nv_bfloat16* __restrict__ out,
                                                          long long ldo, int D) {
  griddep_wait();  // programmatic dependent launch: the previous kernel's outputs are read from here on
  griddep_launch_dependents();
  const int r = blockIdx.x;
  const int i = idx[r];
  uint4* o = reinterpret_cast<uint4*>(out + static_cast<long long>(r) * ldo);
  if (i == -1) {
    for (int c = threadIdx.x; c < D / 8; c += 128) o[c] = make_uint4(0, 0, 0, 0);
    return;
  }
  const uint4* src = i >= 0 ? reinterpret_cast<const uint4*>(table + static_cast<long long>(i) * ldt)
                            : reinterpret_cast<const uint4*>(feats + static_cast<long long>(-i - 2) * ldf);
  for (int c = threadIdx.x; c < D / 8; c += 128) o[c] = src[c];
}

// ------------------------------------------------------------------------------------------------- argmax
__global__ void __launch_bounds__(1024) argmax_kernel(const float* __restrict__ x, long long ld, int V,
                                                      long long* __restrict__ out) {
  __shared__ float sv[32];
  __shared__ int si[32];
  const float* xr = x + static_cast<long long>(blockIdx.x) * ld;
  float best = -INFINITY;
  int bi = 0x7fffffff;
  for (int i = threadIdx.x; i < V; i += 1024) {
    const float v = xr[i];
    if (v > best || (v == best && i < bi)) {
      best = v;
      bi = i;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > best || (ov == best && oi < bi)) {
      best = ov;
      bi = oi;
    }
  }
  if ((threadIdx.x & 31) == 0) {
    sv[threadIdx.x >> 5] = best;
    si[threadIdx.x >> 5] = bi;
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    best = sv[threadIdx.x];
    bi = si[threadIdx.x];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > best || (ov == best && oi < bi)) {
        best = ov;
        bi = oi;
      }
    }
    if (threadIdx.x == 0) out[blockIdx.x] = bi == 0x7fffffff ? 0 : bi;
  }
}

// ------------------------------------------------------------------------------------------------- im2col (patches)
// img bf16 [B,C,H,W]; out [B*gh*gw, Kpad], column = (c*P + ky)*P + kx (= conv weight.view(out, -1) order).
__global__ void __launch_bounds__(256) im2col_patch_kernel(const __nv_bfloat16* __restrict__ img,
                                                           __nv_bfloat16* __restrict__ out, int C, int H, int W, int P,
                                                           int gh, int gw, int Kpad) {
  const int r = blockIdx.x;
  const int b = r / (gh * gw), g = r % (gh * gw);
  const int gy = g / gw, gx = g % gw;
  const int K = C * P * P;
  __nv_bfloat16* o = out + static_cast<long long>(r) * Kpad;
  for (int kk = threadIdx.x; kk < Kpad; kk += 256) {
    __nv_bfloat16 v = __float2bfloat16_rn(0.0f);
    if (kk < K) {
      const int c = kk / (P * P), rem = kk % (P * P);
      const int ky = rem / P, kx = rem % P;
      v = img[((static_cast<long long>(b) * C + c) * H + gy * P + ky) * W + gx * P + kx];
    }
    o[kk] = v;
  }
}

// x bf16 [B,H,W,C] (token-major) -> out [B*Ho*Wo, kh*kw*C], column = (ky*kw + kx)*C + c; zero padding.
// gate (bf16 [B,C] or NULL): x is multiplied by its channel gate (rounded to bf16) on the way.
__global__ void __launch_bounds__(128) im2col_nhwc_kernel(const __nv_bfloat16* __restrict__ x,
                                                          const __nv_bfloat16* __restrict__ gate,
                                                          __nv_bfloat16* __restrict__ out, int H, int W, int C, int kh,
                                                          int kw, int stride, int pad, int Ho, int Wo) {
  const int r = blockIdx.x;
  const int b = r / (Ho * Wo), p = r % (Ho * Wo);
  const int oy = p / Wo, ox = p % Wo;
  const int cpr = C / 8;
  uint4* o = reinterpret_cast<uint4*>(out + static_cast<long long>(r) * kh * kw * C);
  for (int w = threadIdx.x; w < kh * kw * cpr; w += 128) {
    const int tap = w / cpr, c = (w % cpr) * 8;
    const int iy = oy * stride - pad + tap / kw, ix = ox * stride - pad + tap % kw;
    uint4 val = make_uint4(0, 0, 0, 0);
    if (iy >= 0 && iy < H && ix >= 0 && ix < W) {
      val = *reinterpret_cast<const uint4*>(x + ((static_cast<long long>(b) * H + iy) * W + ix) * C + c);
      if (gate != nullptr) {
        const uint4 gv = *reinterpret_cast<const uint4*>(gate + static_cast<long long>(b) * C + c);
        const __nv_bfloat162* xp = reinterpret_cast<const __nv_bfloat162*>(&val);
        const __nv_bfloat162* gp = reinterpret_cast<const __nv_bfloat162*>(&gv);
        uint32_t* op = reinterpret_cast<uint32_t*>(&val);
        uint32_t tmp[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 a = __bfloat1622float2(xp[i]), g2 = __bfloat1622float2(gp[i]);
          tmp[i] = pack_bf16(a.x * g2.x, a.y * g2.y);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) op[i] = tmp[i];
      }
    }
    o[w] = val;
  }
}

// ------------------------------------------------------------------------------------------------- CLIP embeddings
__global__ void __launch_bounds__(128) clip_embed_kernel(const __nv_bfloat16* __restrict__ patch,
                                                         const __nv_bfloat16* __restrict__ cls,
                                                         const __nv_bfloat16* __restrict__ pos,
                                                         __nv_bfloat16* __restrict__ out, int np, int D) {
  const int r = blockIdx.x;  // b*(np+1) + t
  const int b = r / (np + 1), t = r % (np + 1);
  const __nv_bfloat16* src = t == 0 ? cls : patch + (static_cast<long long>(b) * np + t - 1) * D;
  const __nv_bfloat16* pr = pos + static_cast<long long>(t) * D;
  __nv_bfloat16* o = out + static_cast<long long>(r) * D;
  for (int c = threadIdx.x * 8; c < D; c += 128 * 8) {
    const uint4 a = *reinterpret_cast<const uint4*>(src + c), p = *reinterpret_cast<const uint4*>(pr + c);
    const __nv_bfloat162* ap = reinterpret_cast<const __nv_bfloat162*>(&a);
    const __nv_bfloat162* pp = reinterpret_cast<const __nv_bfloat162*>(&p);
    uint4 ov;
    uint32_t* op = reinterpret_cast<uint32_t*>(&ov);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 x = __bfloat1622float2(ap[i]), y = __bfloat1622float2(pp[i]);
      op[i] = pack_bf16(x.x + y.x, x.y + y.y);
    }
    *reinterpret_cast<uint4*>(o + c) = ov;
  }
}

// ------------------------------------------------------------------------------------------------- SAM rel-pos
// One thread per output: rel_h[bh, t, kk] = bf16(q[bh,t,:] . Rh[y - kk + hh - 1, :]), same for w.
__global__ void __launch_bounds__(256) sam_relpos_kernel(const __nv_bfloat16* __restrict__ q, long long q_sb,
                                                         long long q_st, long long q_sh,
                                                         const __nv_bfloat16* __restrict__ rph,
                                                         const __nv_bfloat16* __restrict__ rpw,
                                                         float* __restrict__ rel_h, float* __restrict__ rel_w, int B,
                                                         int H, int hh, int ww, int hd) {
  const long long total = static_cast<long long>(B) * H * hh * ww * (hh + ww);
  const long long gid = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
  if (gid >= total) return;
  const int j = gid % (hh + ww);
  long long rest = gid / (hh + ww);
  const int t = rest % (hh * ww);
  rest /= (hh * ww);
  const int h = rest % H, b = rest / H;
  const int y = t / ww, x = t % ww;
  const __nv_bfloat16* qr = q + b * q_sb + t * q_st + h * q_sh;
  const __nv_bfloat16* rr = j < hh ? rph + static_cast<long long>(y - j + hh - 1) * hd
                                   : rpw + static_cast<long long>(x - (j - hh) + ww - 1) * hd;
  float acc = 0.0f;
  for (int c = 0; c < hd; c += 8) {
    const uint4 a = *reinterpret_cast<const uint4*>(qr + c), r = *reinterpret_cast<const uint4*>(rr + c);
    const __nv_bfloat162* ap = reinterpret_cast<const __nv_bfloat162*>(&a);
    const __nv_bfloat162* rp = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 u = __bfloat1622float2(ap[i]), w = __bfloat1622float2(rp[i]);
      acc += u.x * w.x + u.y * w.y;
    }
  }
  acc = bf16_round(acc);
  const long long bh = static_cast<long long>(b) * H + h;
  if (j < hh)
    rel_h[(bh * hh * ww + t) * hh + j] = acc;
  else
    rel_w[(bh * hh * ww + t) * ww + (j - hh)] = acc;
}

// ------------------------------------------------------------------------------------------------- column mean
// block = 32 channel-chunks (8 channels each) x 8 row groups; rows strided over the groups, reduced through smem.
__global__ void __launch_bounds__(256) col_mean_kernel(const __nv_bfloat16* __restrict__ x,
                                                       __nv_bfloat16* __restrict__ out, int T, int C) {
  __shared__ float part[8][32][8];
  const int b = blockIdx.y;
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int c = (blockIdx.x * 32 + cx) * 8;
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.0f;
  if (c < C) {
    const __nv_bfloat16* xb = x + static_cast<long long>(b) * T * C + c;
    for (int t = ry; t < T; t += 8) {
      const uint4 a = *reinterpret_cast<const uint4*>(xb + static_cast<long long>(t) * C);
      const __nv_bfloat162* ap = reinterpret_cast<const __nv_bfloat162*>(&a);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 f = __bfloat1622float2(ap[i]);
        acc[2 * i] += f.x;
        acc[2 * i + 1] += f.y;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) part[ry][cx][i] = acc[i];
  __syncthreads();
  if (ry == 0 && c < C) {
    const float inv = 1.0f / static_cast<float>(T);
    float tot[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float v = 0.0f;
#pragma unroll
      for (int r = 0; r < 8; ++r) v += part[r][cx][i];
      tot[i] = v * inv;
    }
    uint4 o;
    o.x = pack_bf16(tot[0], tot[1]);
    o.y = pack_bf16(tot[2], tot[3]);
    o.z = pack_bf16(tot[4], tot[5]);
    o.w = pack_bf16(tot[6], tot[7]);
    *reinterpret_cast<uint4*>(out + static_cast<long long>(b) * C + c) = o;
  }
}

// ------------------------------------------------------------------------------------------------- convT col2im
// cols f32 [B*Hi*Wi, 16*C], column = (ky*4 + kx)*C + co. ConvTranspose2d(k=4, s=2, p=1): oy = 2*iy - 1 + ky.
// out[b,oy,ox,co] = bf16(skip + relu(bf16(sum of the (up to 4) contributing taps))).
__global__ void __launch_bounds__(128) convt4s2_col2im_kernel(const float* __restrict__ cols,
                                                              const __nv_bfloat16* __restrict__ skip,
                                                              __nv_bfloat16* __restrict__ out, int Hi, int Wi, int C) {
  const int Ho = 2 * Hi, Wo = 2 * Wi;
  const int r = blockIdx.x;
  const int b = r / (Ho * Wo), p = r % (Ho * Wo);
  const int oy = p / Wo, ox = p % Wo;
  for (int c = threadIdx.x * 4; c < C; c += 128 * 4) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int ky = 0; ky < 4; ++ky) {
      const int ny = oy + 1 - ky;
      if (ny < 0 || (ny & 1) || (ny >> 1) >= Hi) continue;
#pragma unroll
      for (int kx = 0; kx < 4; ++kx) {
        const int nx = ox + 1 - kx;
        if (nx < 0 || (nx & 1) || (nx >> 1) >= Wi) continue;
        const long long row = (static_cast<long long>(b) * Hi + (ny >> 1)) * Wi + (nx >> 1);
        const float4 v = *reinterpret_cast<const float4*>(cols + row * 16 * C + (ky * 4 + kx) * C + c);
        acc.x += v.x;
        acc.y += v.y;
        acc.z += v.z;
        acc.w += v.w;
      }
    }
    float o[4] = {fmaxf(bf16_round(acc.x), 0.f), fmaxf(bf16_round(acc.y), 0.f), fmaxf(bf16_round(acc.z), 0.f),
                  fmaxf(bf16_round(acc.w), 0.f)};
    const long long off = static_cast<long long>(r) * C + c;
    if (skip != nullptr) {
#pragma unroll
      for (int i = 0; i < 4; ++i) o[i] += __bfloat162float(skip[off + i]);
    }
    uint2 ov;
    ov.x = pack_bf16(o[0], o[1]);
    ov.y = pack_bf16(o[2], o[3]);
    *reinterpret_cast<uint2*>(out + off) = ov;
  }
}

// ------------------------------------------------------------------------------------------------- adds
// out = bf16(a + b) with b bf16 or f32; b is broadcast over rows when b_rows < rows (row = i / D).
__global__ void __launch_bounds__(256) add_kernel(const __nv_bfloat16* __restrict__ a, const void* __restrict__ b,
                                                  int b_f32, __nv_bfloat16* __restrict__ out, long long n,
                                                  long long b_period) {
  griddep_wait();  // programmatic dependent launch: the previous kernel's outputs are read from here on
  griddep_launch_dependents();
  const long long i = (static_cast<long long>(blockIdx.x) * 256 + threadIdx.x) * 8;
  if (i >= n) return;
  const long long j = i % b_period;
  const uint4 av = *reinterpret_cast<const uint4*>(a + i);
  const __nv_bfloat162* ap = reinterpret_cast<const __nv_bfloat162*>(&av);
  float bv[8];
  if (b_f32) {
    const float4 b0 = *reinterpret_cast<const float4*>(static_cast<const float*>(b) + j);
    const float4 b1 = *reinterpret_cast<const float4*>(static_cast<const float*>(b) + j + 4);
    bv[0] = b0.x; bv[1] = b0.y; bv[2] = b0.z; bv[3] = b0.w;
    bv[4] = b1.x; bv[5] = b1.y; bv[6] = b1.z; bv[7] = b1.w;
  } else {
    const uint4 braw = *reinterpret_cast<const uint4*>(static_cast<const __nv_bfloat16*>(b) + j);
    const __nv_bfloat162* bp = reinterpret_cast<const __nv_bfloat162*>(&braw);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float2 f = __bfloat1622float2(bp[q]);
      bv[2 * q] = f.x;
      bv[2 * q + 1] = f.y;
    }
  }
  uint4 o;
  uint32_t* op = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float2 f = __bfloat1622float2(ap[q]);
    op[q] = pack_bf16(f.x + bv[2 * q], f.y + bv[2 * q + 1]);
  }
  *reinterpret_cast<uint4*>(out + i) = o;
}

// ------------------------------------------------------------------------------------------------- bilinear resize
// in bf16 [N, Hin(rows of ld_in), Win] -> out [N, Hout, Wout] (bf16 or f32); align_corners=False, fp32 math.
__global__ void __launch_bounds__(256) bilinear_kernel(const __nv_bfloat16* __restrict__ in, long long in_sn,
                                                       long long in_sy, int Hin, int Win, void* __restrict__ out,
                                                       int out_f32, int Hout, int Wout, int N) {
  const long long gid = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
  const long long total = static_cast<long long>(N) * Hout * Wout;
  if (gid >= total) return;
  const int ox = gid % Wout;
  const int oy = (gid / Wout) % Hout;
  const int n = gid / (static_cast<long long>(Wout) * Hout);
  const float sy = static_cast<float>(Hin) / Hout, sx = static_cast<float>(Win) / Wout;
  float fy = sy * (oy + 0.5f) - 0.5f, fx = sx * (ox + 0.5f) - 0.5f;
  fy = fy < 0.f ? 0.f : fy;
  fx = fx < 0.f ? 0.f : fx;
  const int y0 = min(static_cast<int>(fy), Hin - 1), x0 = min(static_cast<int>(fx), Win - 1);
  const int y1 = min(y0 + 1, Hin - 1), x1 = min(x0 + 1, Win - 1);
  const float ly = fy - y0, lx = fx - x0;
  const __nv_bfloat16* base = in + n * in_sn;
  const float v00 = __bfloat162float(base[y0 * in_sy + x0]), v01 = __bfloat162float(base[y0 * in_sy + x1]);
  const float v10 = __bfloat162float(base[y1 * in_sy + x0]), v11 = __bfloat162float(base[y1 * in_sy + x1]);
  const float v = (1.f - ly) * ((1.f - lx) * v00 + lx * v01) + ly * ((1.f - lx) * v10 + lx * v11);
  if (out_f32)
    static_cast<float*>(out)[gid] = v;
  else
    static_cast<__nv_bfloat16*>(out)[gid] = __float2bfloat16_rn(v);
}

// ------------------------------------------------------------------------------------------------- region sample
// fmap bf16 [h*w, C]; pts f32 [P,2] = (x, y) in [0,1] (already rounded the way the reference rounds them);
// grid = bf16(2 * pts - 1); grid_sample(bilinear, align_corners=True, zeros padding) in fp32 -> bf16 per point -> mean
// over points -> bf16.
__global__ void __launch_bounds__(128) region_sample_kernel(const __nv_bfloat16* __restrict__ fmap,
                                                            const float* __restrict__ pts, int P, int h, int w, int C,
                                                            __nv_bfloat16* __restrict__ out) {
  const int c = blockIdx.x * 128 + threadIdx.x;
  if (c >= C) return;
  float acc = 0.0f;
  for (int p = 0; p < P; ++p) {
    // point_sample (medplib_arch.py:39-64) forms 2 * coords - 1 on the run-dtype (bf16) coordinates BEFORE its .float():
    // the grid itself is rounded to bf16 (2 * c is exact, the subtraction is not)
    const float gx = bf16_round(2.0f * pts[2 * p] - 1.0f), gy = bf16_round(2.0f * pts[2 * p + 1] - 1.0f);
    const float fx = (gx + 1.0f) * 0.5f * (w - 1), fy = (gy + 1.0f) * 0.5f * (h - 1);
    const int x0 = static_cast<int>(floorf(fx)), y0 = static_cast<int>(floorf(fy));
    const float lx = fx - x0, ly = fy - y0;
    float v = 0.0f;
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        const int xx = x0 + dx, yy = y0 + dy;
        if (xx < 0 || xx >= w || yy < 0 || yy >= h) continue;
        const float wgt = (dx ? lx : 1.0f - lx) * (dy ? ly : 1.0f - ly);
        v += wgt * __bfloat162float(fmap[(static_cast<long long>(yy) * w + xx) * C + c]);
      }
    acc += bf16_round(v);
  }
  // mean over zero points is NaN in the reference, then nan_to_num -> 0
  out[c] = __float2bfloat16_rn(P > 0 ? acc / static_cast<float>(P) : 0.0f);
}

static inline int launched() { return launch_status(); }

}  // namespace mpl

using namespace mpl;
typedef __nv_bfloat16 bf;

extern "C" int mpl_rope_kv(void* q, void* k, const void* v, long long ld, const void* cos_t, const void* sin_t,
                           void* k_cache, void* v_cache, int B, int T, int H, int head_dim, int Tmax, int pos0,
                           const int* pos_dev, void* stream) {
  if (B <= 0 || T <= 0) return MPL_OK;
  if (cos_t == nullptr || sin_t == nullptr || (q == nullptr && k == nullptr)) return MPL_ERR_ARG;
  if ((head_dim % 16) != 0 || (ld % 8) != 0) return MPL_ERR_ALIGN;
  if (k_cache != nullptr && pos_dev == nullptr && pos0 + T > Tmax) return MPL_ERR_ARG;
  launch_pdl(rope_kv_kernel, dim3(B * T), dim3(256), 0, static_cast<cudaStream_t>(stream), 
      static_cast<bf*>(q), static_cast<bf*>(k), static_cast<const bf*>(v), ld, static_cast<const bf*>(cos_t),
      static_cast<const bf*>(sin_t), static_cast<bf*>(k_cache), static_cast<bf*>(v_cache), T, H, head_dim, Tmax, pos0,
      pos_dev);
  return launched();
}

extern "C" int mpl_gather_rows(const void* table, long long ld_table, const void* feats, long long ld_feats,
                               const int* idx, void* out, long long ld_out, int rows, int D, void* stream) {
  if (rows <= 0) return MPL_OK;
  if (idx == nullptr || out == nullptr) return MPL_ERR_ARG;
  if ((D % 8) != 0 || (ld_table % 8) != 0 || (ld_feats % 8) != 0 || (ld_out % 8) != 0) return MPL_ERR_ALIGN;
  launch_pdl(gather_rows_kernel, dim3(rows), dim3(128), 0, static_cast<cudaStream_t>(stream), 
      static_cast<const bf*>(table), ld_table, static_cast<const bf*>(feats), ld_feats, idx, static_cast<bf*>(out),
      ld_out, D);
  return launched();
}

extern "C" int mpl_argmax_f32(const float* x, long long ld, int rows, int V, long long* out, void* stream) {
  if (rows <= 0) return MPL_OK;
  if (x == nullptr || out == nullptr || V <= 0) return MPL_ERR_ARG;
  argmax_kernel<<<rows, 1024, 0, static_cast<cudaStream_t>(stream)>>>(x, ld, V, out);
  return launched();
}

extern "C" int mpl_im2col_patch(const void* img, void* out, int B, int C, int H, int W, int P, int Kpad,
                                void* stream) {
  if (B <= 0) return MPL_OK;
  if (img == nullptr || out == nullptr || P <= 0 || H % P || W % P || Kpad < C * P * P) return MPL_ERR_ARG;
  const int gh = H / P, gw = W / P;
  im2col_patch_kernel<<<B * gh * gw, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const bf*>(img), static_cast<bf*>(out), C, H, W, P, gh, gw, Kpad);
  return launched();
}

extern "C" int mpl_im2col_nhwc(const void* x, const void* gate, void* out, int B, int H, int W, int C, int kh, int kw,
                               int stride, int pad, void* stream) {
  if (B <= 0) return MPL_OK;
  if (x == nullptr || out == nullptr || kh <= 0 || kw <= 0 || stride <= 0) return MPL_ERR_ARG;
  if ((C % 8) != 0) return MPL_ERR_ALIGN;
  const int Ho = (H + 2 * pad - kh) / stride + 1, Wo = (W + 2 * pad - kw) / stride + 1;
  im2col_nhwc_kernel<<<B * Ho * Wo, 128, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const bf*>(x), static_cast<const bf*>(gate), static_cast<bf*>(out), H, W, C, kh, kw, stride, pad, Ho,
      Wo);
  return launched();
}

extern "C" int mpl_clip_embed(const void* patch, const void* cls, const void* pos, void* out, int B, int n_patches,
                              int D, void* stream) {
  if (B <= 0) return MPL_OK;
  if (patch == nullptr || cls == nullptr || pos == nullptr || out == nullptr) return MPL_ERR_ARG;
  if ((D % 8) != 0) return MPL_ERR_ALIGN;
  clip_embed_kernel<<<B * (n_patches + 1), 128, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const bf*>(patch), static_cast<const bf*>(cls), static_cast<const bf*>(pos), static_cast<bf*>(out),
      n_patches, D);
  return launched();
}

extern "C" int mpl_sam_relpos(const void* q, long long q_sb, long long q_st, long long q_sh, const void* rel_pos_h,
                              const void* rel_pos_w, float* rel_h, float* rel_w, int B, int H, int hh, int ww,
                              int head_dim, void* stream) {
  if (B <= 0) return MPL_OK;
  if (q == nullptr || rel_pos_h == nullptr || rel_pos_w == nullptr || rel_h == nullptr || rel_w == nullptr)
    return MPL_ERR_ARG;
  if ((head_dim % 8) != 0 || (q_sb % 8) != 0 || (q_st % 8) != 0 || (q_sh % 8) != 0) return MPL_ERR_ALIGN;
  const long long total = static_cast<long long>(B) * H * hh * ww * (hh + ww);
  sam_relpos_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const bf*>(q), q_sb, q_st, q_sh, static_cast<const bf*>(rel_pos_h), static_cast<const bf*>(rel_pos_w),
      rel_h, rel_w, B, H, hh, ww, head_dim);
  return launched();
}

extern "C" int mpl_col_mean(const void* x, void* out, int B, int T, int C, void* stream) {
  if (B <= 0) return MPL_OK;
  if (x == nullptr || out == nullptr || T <= 0) return MPL_ERR_ARG;
  if ((C % 8) != 0) return MPL_ERR_ALIGN;
  dim3 grid((C / 8 + 31) / 32, B);
  col_mean_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const bf*>(x), static_cast<bf*>(out),
                                                                      T, C);
  return launched();
}

extern "C" int mpl_convt4s2_col2im(const float* cols, const void* skip, void* out, int B, int Hi, int Wi, int C,
                                   void* stream) {
  if (B <= 0) return MPL_OK;
  if (cols == nullptr || out == nullptr) return MPL_ERR_ARG;
  if ((C % 4) != 0) return MPL_ERR_ALIGN;
  convt4s2_col2im_kernel<<<B * 4 * Hi * Wi, 128, 0, static_cast<cudaStream_t>(stream)>>>(
      cols, static_cast<const bf*>(skip), static_cast<bf*>(out), Hi, Wi, C);
  return launched();
}

extern "C" int mpl_add(const void* a, const void* b, int b_is_f32, void* out, long long n, long long b_period,
                       void* stream) {
  if (n <= 0) return MPL_OK;
  if (a == nullptr || b == nullptr || out == nullptr) return MPL_ERR_ARG;
  if ((n % 8) != 0 || b_period <= 0 || (b_period % 8) != 0) return MPL_ERR_ALIGN;
  launch_pdl(add_kernel, dim3(static_cast<unsigned>((n / 8 + 255) / 256)), dim3(256), 0, static_cast<cudaStream_t>(stream), 
      static_cast<const bf*>(a), b, b_is_f32, static_cast<bf*>(out), n, b_period);
  return launched();
}

extern "C" int mpl_bilinear_resize(const void* in, long long in_stride_n, long long in_stride_y, int Hin, int Win,
                                   void* out, int out_dtype, int Hout, int Wout, int N, void* stream) {
  if (N <= 0) return MPL_OK;
  if (in == nullptr || out == nullptr || Hin <= 0 || Win <= 0 || Hout <= 0 || Wout <= 0) return MPL_ERR_ARG;
  const long long total = static_cast<long long>(N) * Hout * Wout;
  bilinear_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const bf*>(in), in_stride_n, in_stride_y, Hin, Win, out, out_dtype == MPL_DT_F32, Hout, Wout, N);
  return launched();
}

extern "C" int mpl_region_sample_mean(const void* fmap, const float* pts, int P, int h, int w, int C, void* out,
                                      void* stream) {
  if (fmap == nullptr || out == nullptr || (P > 0 && pts == nullptr)) return MPL_ERR_ARG;
  region_sample_kernel<<<(C + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const bf*>(fmap), pts, P, h, w, C, static_cast<bf*>(out));
  return launched();
}
