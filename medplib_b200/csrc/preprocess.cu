// Image input pipeline: u8 decoded image -> resized / padded / normalised tensor, one launch per batch.
//
// PIL's BILINEAR resize is two integer convolutions with a u8 rounding in between, so the kernel keeps that structure:
// a CTA owns a band of R output rows of one job, runs the horizontal pass over just the source rows that band's
// vertical taps touch into shared memory (u8), then the vertical pass + the per-level fp32 table + the centre pad
// straight to the [C, L, L] output.  HBM traffic is the source once per job (bands overlap by the filter support,
// absorbed by L2) and the output once; there is no intermediate image in HBM.  Integer arithmetic follows
// Pillow's libImaging/Resample.c (ImagingResampleHorizontal_8bpc / Vertical_8bpc) so results are bit-exact.
#include <cuda_bf16.h>
#include <stdint.h>

#include "internal.h"

namespace {

constexpr int kThreads = 256;
constexpr int kPrecisionBits = 32 - 8 - 2;
constexpr int kSmemBudget = 200 * 1024;

// Upper bound of the source rows touched by R consecutive output rows: the band's centres span (R-1)*scale and each
// end reaches support + 0.5 further (scale = in/out, support = max(scale, 1)); rounded up in integers.
__host__ __device__ inline int band_rows_bound(int in_size, int out_size, int R) {
  long long sup = (in_size + out_size - 1) / out_size;
  if (sup < 1) sup = 1;
  const long long span = (static_cast<long long>(R - 1) * in_size + out_size - 1) / out_size;
  const long long b = span + 2 * sup + 2;
  return b > in_size ? in_size : static_cast<int>(b);
}

__host__ __device__ inline int pick_band(int H, int new_h, int new_w, int C) {
  for (int R = 8; R >= 1; R >>= 1)
    if (static_cast<long long>(band_rows_bound(H, new_h, R)) * new_w * C <= kSmemBudget) return R;
  return 0;
}

__device__ __forceinline__ int clip8(int acc) {
  const int v = acc >> kPrecisionBits;
  return v < 0 ? 0 : (v > 255 ? 255 : v);
}

__device__ __forceinline__ unsigned ld_word(uintptr_t addr, uintptr_t last) {
  return __ldg(reinterpret_cast<const unsigned*>(addr < last ? addr : last));
}

__global__ void __launch_bounds__(kThreads) preprocess_kernel(const mpl_preprocess_job* __restrict__ jobs) {
  extern __shared__ unsigned char tmp[];  // [rows][new_w * C] u8: the horizontal pass of this band's source rows
  __shared__ mpl_preprocess_job sj;
  __shared__ float s_lut[3 * 256];
  const int tid = threadIdx.x;
  if (tid == 0) sj = jobs[blockIdx.y];
  __syncthreads();
  const int C = sj.C, L = sj.out_size, new_h = sj.new_h, new_w = sj.new_w;
  const int R = pick_band(sj.H, new_h, new_w, C);
  const int y0 = blockIdx.x * R;
  if (R == 0 || y0 >= L) return;
  const int y1 = min(y0 + R, L);
  const int ry0 = max(y0 - sj.pad_top, 0), ry1 = min(y1 - sj.pad_top, new_h);
  const int pitch = new_w * C;
  int in_lo = 0;
  if (ry0 < ry1) {
    in_lo = sj.bound_y[2 * ry0];
    const int nrows = sj.bound_y[2 * (ry1 - 1)] + sj.bound_y[2 * (ry1 - 1) + 1] - in_lo;
    const unsigned char* src = static_cast<const unsigned char*>(sj.src);
    // Source bytes are read as aligned 32-bit words and funnel-shifted to the pixel boundary: 4 taps of an RGB pixel
    // are 12 bytes = 3 words (one word for a single-channel mask), so a tap group costs 3 word loads instead of 12
    // byte loads.  Words past the image end are clamped to the last valid word: those positions only ever meet the
    // zero coefficients that pad every tap list to a multiple of 4 (coef_x is tap-major [ks_x][new_w]).
    const uintptr_t wlast =
        (reinterpret_cast<uintptr_t>(src) + static_cast<uintptr_t>(sj.H - 1) * sj.src_stride + sj.W * C - 1) & ~uintptr_t(3);
    for (int idx = tid; idx < nrows * new_w; idx += kThreads) {
      const int r = idx / new_w, xx = idx - r * new_w;
      const int xmin = __ldg(sj.bound_x + 2 * xx), n = __ldg(sj.bound_x + 2 * xx + 1);
      const int* k = sj.coef_x + xx;
      const uintptr_t a = reinterpret_cast<uintptr_t>(src) + static_cast<uintptr_t>(in_lo + r) * sj.src_stride +
                          static_cast<uintptr_t>(xmin) * C;
      const unsigned sh = static_cast<unsigned>(a & 3) * 8;
      uintptr_t wp = a & ~uintptr_t(3);
      unsigned char* d = tmp + r * pitch + xx * C;
      unsigned w0 = ld_word(wp, wlast);
      if (C == 3) {
        int a0 = 1 << (kPrecisionBits - 1), a1 = a0, a2 = a0;
        for (int g = 0; g < n; g += 4) {
          const unsigned w1 = ld_word(wp + 4, wlast), w2 = ld_word(wp + 8, wlast), w3 = ld_word(wp + 12, wlast);
          const unsigned u0 = __funnelshift_r(w0, w1, sh), u1 = __funnelshift_r(w1, w2, sh), u2 = __funnelshift_r(w2, w3, sh);
          const int k0 = __ldg(k + g * new_w), k1 = __ldg(k + (g + 1) * new_w), k2 = __ldg(k + (g + 2) * new_w),
                    k3 = __ldg(k + (g + 3) * new_w);
          a0 += static_cast<int>(u0 & 0xff) * k0;
          a1 += static_cast<int>((u0 >> 8) & 0xff) * k0;
          a2 += static_cast<int>((u0 >> 16) & 0xff) * k0;
          a0 += static_cast<int>(u0 >> 24) * k1;
          a1 += static_cast<int>(u1 & 0xff) * k1;
          a2 += static_cast<int>((u1 >> 8) & 0xff) * k1;
          a0 += static_cast<int>((u1 >> 16) & 0xff) * k2;
          a1 += static_cast<int>(u1 >> 24) * k2;
          a2 += static_cast<int>(u2 & 0xff) * k2;
          a0 += static_cast<int>((u2 >> 8) & 0xff) * k3;
          a1 += static_cast<int>((u2 >> 16) & 0xff) * k3;
          a2 += static_cast<int>(u2 >> 24) * k3;
          w0 = w3;
          wp += 12;
        }
        d[0] = static_cast<unsigned char>(clip8(a0));
        d[1] = static_cast<unsigned char>(clip8(a1));
        d[2] = static_cast<unsigned char>(clip8(a2));
      } else {
        int a0 = 1 << (kPrecisionBits - 1);
        for (int g = 0; g < n; g += 4) {
          const unsigned w1 = ld_word(wp + 4, wlast);
          const unsigned u0 = __funnelshift_r(w0, w1, sh);
          a0 += static_cast<int>(u0 & 0xff) * __ldg(k + g * new_w);
          a0 += static_cast<int>((u0 >> 8) & 0xff) * __ldg(k + (g + 1) * new_w);
          a0 += static_cast<int>((u0 >> 16) & 0xff) * __ldg(k + (g + 2) * new_w);
          a0 += static_cast<int>(u0 >> 24) * __ldg(k + (g + 3) * new_w);
          w0 = w1;
          wp += 4;
        }
        d[0] = static_cast<unsigned char>(clip8(a0));
      }
    }
  }
  if (sj.lut != nullptr)
    for (int i = tid; i < C * 256; i += kThreads) s_lut[i] = __ldg(sj.lut + i);
  __syncthreads();

  const int rows = y1 - y0;
  const int total = C * rows * L;
  for (int idx = tid; idx < total; idx += kThreads) {
    const int xo = idx % L, t = idx / L;
    const int yo = y0 + t % rows, c = t / rows;
    const int ry = yo - sj.pad_top, rx = xo - sj.pad_left;
    float val = sj.pad_value[c];
    if (ry >= 0 && ry < new_h && rx >= 0 && rx < new_w) {
      const int ymin = __ldg(sj.bound_y + 2 * ry) - in_lo, n = __ldg(sj.bound_y + 2 * ry + 1);
      const int* k = sj.coef_y + static_cast<long long>(ry) * sj.ks_y;
      const unsigned char* s = tmp + ymin * pitch + rx * C + c;
      int acc = 1 << (kPrecisionBits - 1);
      for (int y = 0; y < n; ++y) acc += s[y * pitch] * __ldg(k + y);
      const int level = clip8(acc);
      val = sj.lut != nullptr ? s_lut[c * 256 + level] : static_cast<float>(level);
    }
    const long long o = (static_cast<long long>(c) * L + yo) * L + xo;
    if (sj.out_dtype == MPL_DT_F32)
      static_cast<float*>(sj.dst)[o] = val;
    else if (sj.out_dtype == MPL_DT_BF16)
      static_cast<__nv_bfloat16*>(sj.dst)[o] = __float2bfloat16_rn(val);
    else
      static_cast<unsigned char*>(sj.dst)[o] = static_cast<unsigned char>(val);
  }
}

}  // namespace

extern "C" int mpl_preprocess_band_rows(int in_size, int out_size, int R) {
  if (in_size <= 0 || out_size <= 0 || R <= 0) return MPL_ERR_ARG;
  return band_rows_bound(in_size, out_size, R);
}

extern "C" int mpl_preprocess_images(const mpl_preprocess_job* jobs_host, const mpl_preprocess_job* jobs_dev,
                                     int n_jobs, void* stream) {
  if (n_jobs <= 0) return MPL_OK;
  if (jobs_host == nullptr || jobs_dev == nullptr || n_jobs > 65535) return MPL_ERR_ARG;
  int bands = 0;
  long long smem = 0;
  for (int i = 0; i < n_jobs; ++i) {
    const mpl_preprocess_job& j = jobs_host[i];
    if (j.src == nullptr || j.dst == nullptr || j.coef_x == nullptr || j.bound_x == nullptr || j.coef_y == nullptr ||
        j.bound_y == nullptr || (j.C != 1 && j.C != 3) || j.H <= 0 || j.W <= 0 || j.new_h <= 0 || j.new_w <= 0 ||
        j.pad_top < 0 || j.pad_left < 0 || j.pad_top + j.new_h > j.out_size || j.pad_left + j.new_w > j.out_size ||
        j.ks_x <= 0 || (j.ks_x & 3) != 0 || j.ks_y <= 0 || j.src_stride < static_cast<long long>(j.W) * j.C ||
        (j.out_dtype != MPL_DT_BF16 && j.out_dtype != MPL_DT_F32 && j.out_dtype != MPL_DT_U8))
      return MPL_ERR_ARG;
    if ((reinterpret_cast<uintptr_t>(j.src) & 3) != 0) return MPL_ERR_ALIGN;  // word loads never reach below src
    const int R = pick_band(j.H, j.new_h, j.new_w, j.C);
    if (R == 0) return MPL_ERR_UNSUPPORTED;
    const int b = (j.out_size + R - 1) / R;
    bands = b > bands ? b : bands;
    const long long need = static_cast<long long>(band_rows_bound(j.H, j.new_h, R)) * j.new_w * j.C;
    smem = need > smem ? need : smem;
  }
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(preprocess_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget) != cudaSuccess)
      return MPL_ERR_CUDA;
    attr_set = true;
  }
  preprocess_kernel<<<dim3(bands, n_jobs), kThreads, static_cast<size_t>(smem), static_cast<cudaStream_t>(stream)>>>(
      jobs_dev);
  return mpl::launch_status();
}
