// Image input pipeline: u8 decoded image -> resized / padded / normalised tensor, one launch per batch.
//
// PIL's BILINEAR resize is two integer convolutions with a u8 rounding in between, so the kernel keeps that structure:
// a CTA owns a band of R output rows of one job, streams just the source rows that band's vertical taps touch through a
// double-buffered cp.async stage (RB rows at a time, 16-byte chunks, so the HBM/L2 latency is paid once per stage, not
// once per tap), runs the horizontal pass from that stage into shared memory (u8), then the vertical pass + the
// per-level fp32 table + the centre pad straight to the [C, L, L] output.  HBM traffic is the source once per job (bands overlap by the filter support,
// absorbed by L2) and the output once; there is no intermediate image in HBM.  Integer arithmetic follows
// Pillow's libImaging/Resample.c (ImagingResampleHorizontal_8bpc / Vertical_8bpc) so results are bit-exact.
//
// MPL_CPU_EMULATION: tests/dev/preprocess_emu.cpp compiles THIS file with g++ (one std::thread per CUDA thread, a
// std::barrier for __syncthreads, memcpy for cp.async) so the index arithmetic, shared-memory layout, staging and
// clamping of both loop structures are checked against the oracle without a GPU. Test infrastructure only; the
// preprocessor switches below do not change the device code (the SASS of the nvcc build is unaffected).
#include <stdint.h>
#include <stdlib.h>
#ifdef MPL_CPU_EMULATION
#include "cuda_emu.h"  // tests/dev
#include "../../include/medplib_b200.h"
#else
#include <cuda_bf16.h>

#include "internal.h"
#endif

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

// Bytes of one staged source row: the row's 16-byte-aligned footprint (up to 15 bytes of misalignment in front) plus the
// 4-tap group over-read behind it (< 16 bytes), in 16-byte chunks.
__host__ __device__ inline int stage_pitch(int W, int C) { return ((W * C + 15) / 16 + 2) * 16; }

// (R output rows per band, RB source rows per cp.async stage) as R | RB << 8, or 0 when nothing fits.
__host__ __device__ inline int pick_band(int H, int W, int new_h, int new_w, int C) {
  const int Rs[6] = {8, 4, 4, 2, 2, 1}, RBs[6] = {4, 4, 2, 2, 1, 1};
  for (int t = 0; t < 6; ++t) {
    const long long need = 2LL * RBs[t] * stage_pitch(W, C) + static_cast<long long>(band_rows_bound(H, new_h, Rs[t])) * new_w * C;
    if (need <= kSmemBudget) return Rs[t] | (RBs[t] << 8);
  }
  return 0;
}

__host__ __device__ inline long long band_smem(int H, int W, int new_h, int new_w, int C, int rrb) {
  return 2LL * (rrb >> 8) * stage_pitch(W, C) + static_cast<long long>(band_rows_bound(H, new_h, rrb & 255)) * new_w * C;
}

__device__ __forceinline__ int clip8(int acc) {
  const int v = acc >> kPrecisionBits;
  return v < 0 ? 0 : (v > 255 ? 255 : v);
}

__device__ __forceinline__ void cp_async16(unsigned char* dst_smem, uintptr_t src_global) {
#ifdef MPL_CPU_EMULATION
  mpl_emu::copy16(dst_smem, src_global);
}
__device__ __forceinline__ void cp_async_commit() {}
template <int N>
__device__ __forceinline__ void cp_async_wait() {}
#else
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(static_cast<unsigned>(__cvta_generic_to_shared(dst_smem))),
               "l"(src_global)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
#endif

// V4 = candidate loop structure (MPL_PREPROCESS_V4=1), identical arithmetic: 128 pixel columns x 2 row groups per CTA,
// so neither pass divides by a runtime value, and the vertical pass computes the channels of a pixel in one thread with
// the row's coefficients loaded once per tap. Not the default until it has been validated and timed on a B200.
template <bool V4>
__global__ void __launch_bounds__(kThreads) preprocess_kernel(const mpl_preprocess_job* __restrict__ jobs) {
#ifdef MPL_CPU_EMULATION
  unsigned char* smem = mpl_emu::dynamic_smem();
#else
  extern __shared__ __align__(16) unsigned char smem[];  // [2][RB][stage_pitch] staged source rows | tmp
#endif
  __shared__ mpl_preprocess_job sj;
  __shared__ float s_lut[3 * 256];
  const int tid = threadIdx.x;
  if (tid == 0) sj = jobs[blockIdx.y];
  __syncthreads();
  const int C = sj.C, L = sj.out_size, new_h = sj.new_h, new_w = sj.new_w;
  const int rrb = pick_band(sj.H, sj.W, new_h, new_w, C);
  const int R = rrb & 255, RB = rrb >> 8;
  const int y0 = blockIdx.x * R;
  if (rrb == 0 || y0 >= L) return;
  const int y1 = min(y0 + R, L);
  const int ry0 = max(y0 - sj.pad_top, 0), ry1 = min(y1 - sj.pad_top, new_h);
  const int pitch = new_w * C;
  const int sp = stage_pitch(sj.W, C), cpr = sp / 16;
  unsigned char* tmp = smem + 2 * RB * sp;  // [rows][new_w * C] u8: the horizontal pass of this band's source rows
  int in_lo = 0;
  if (ry0 < ry1) {
    in_lo = sj.bound_y[2 * ry0];
    const int nrows = sj.bound_y[2 * (ry1 - 1)] + sj.bound_y[2 * (ry1 - 1) + 1] - in_lo;
    const uintptr_t src = reinterpret_cast<uintptr_t>(sj.src);  // 16-byte aligned (checked by the host entry)
    // chunks past the image end are clamped to its last chunk: those bytes only ever meet the zero coefficients that
    // pad every tap list to a multiple of 4 (coef_x is tap-major [ks_x][new_w])
    const uintptr_t last_chunk = (src + static_cast<uintptr_t>(sj.H - 1) * sj.src_stride + sj.W * C - 1) & ~uintptr_t(15);
    auto issue = [&](int stage, int r0) {
      const int nb = min(RB, nrows - r0);
      unsigned char* buf = smem + stage * RB * sp;
      for (int idx = tid; idx < nb * cpr; idx += kThreads) {
        const int i = idx / cpr, v = idx - i * cpr;
        uintptr_t g = ((src + static_cast<uintptr_t>(in_lo + r0 + i) * sj.src_stride) & ~uintptr_t(15)) + 16 * v;
        g = g < last_chunk ? g : last_chunk;
        cp_async16(buf + i * sp + 16 * v, g);
      }
      cp_async_commit();
    };
    issue(0, 0);
    int stage = 0;
    for (int r0 = 0; r0 < nrows; r0 += RB, stage ^= 1) {
      if (r0 + RB < nrows) {
        issue(stage ^ 1, r0 + RB);
        cp_async_wait<1>();
      } else {
        cp_async_wait<0>();
      }
      __syncthreads();
      const int nb = min(RB, nrows - r0);
      const unsigned char* buf = smem + stage * RB * sp;
      // 4 taps of an RGB pixel are 12 bytes = 3 aligned words funnel-shifted to the pixel boundary (1 word for a mask)
      const int n_items = V4 ? ((nb + 1) >> 1) * ((new_w + 127) & ~127) * 2 : nb * new_w;
      for (int idx = tid; idx < n_items; idx += kThreads) {
        int i, xx;
        if (V4) {  // item = (row pair, 128-column block, row group, column): shifts and masks only
          const int blk = idx >> 8;  // one 256-thread step = 128 columns x 2 rows
          const int cb = (new_w + 127) >> 7;
          int pair = 0, b = blk;
          while (b >= cb) {  // at most RB / 2 - 1 = 1 iteration
            b -= cb;
            ++pair;
          }
          i = pair * 2 + ((idx >> 7) & 1);
          xx = (b << 7) + (idx & 127);
          if (i >= nb || xx >= new_w) continue;
        } else {
          i = idx / new_w;
          xx = idx - i * new_w;
        }
        const int xmin = __ldg(sj.bound_x + 2 * xx), n = __ldg(sj.bound_x + 2 * xx + 1);
        const int* k = sj.coef_x + xx;
        const int mis = static_cast<int>((src + static_cast<uintptr_t>(in_lo + r0 + i) * sj.src_stride) & 15);
        const int a = mis + xmin * C;
        const unsigned sh = static_cast<unsigned>(a & 3) * 8;
        const unsigned* wp = reinterpret_cast<const unsigned*>(buf + i * sp + (a & ~3));
        unsigned char* d = tmp + (r0 + i) * pitch + xx * C;
        unsigned w0 = wp[0];
        if (C == 3) {
          int a0 = 1 << (kPrecisionBits - 1), a1 = a0, a2 = a0;
          for (int g = 0; g < n; g += 4) {
            const unsigned w1 = wp[1], w2 = wp[2], w3 = wp[3];
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
            wp += 3;
          }
          d[0] = static_cast<unsigned char>(clip8(a0));
          d[1] = static_cast<unsigned char>(clip8(a1));
          d[2] = static_cast<unsigned char>(clip8(a2));
        } else {
          int a0 = 1 << (kPrecisionBits - 1);
          for (int g = 0; g < n; g += 4) {
            const unsigned w1 = wp[1];
            const unsigned u0 = __funnelshift_r(w0, w1, sh);
            a0 += static_cast<int>(u0 & 0xff) * __ldg(k + g * new_w);
            a0 += static_cast<int>((u0 >> 8) & 0xff) * __ldg(k + (g + 1) * new_w);
            a0 += static_cast<int>((u0 >> 16) & 0xff) * __ldg(k + (g + 2) * new_w);
            a0 += static_cast<int>(u0 >> 24) * __ldg(k + (g + 3) * new_w);
            w0 = w1;
            wp += 1;
          }
          d[0] = static_cast<unsigned char>(clip8(a0));
        }
      }
      __syncthreads();  // the stage just read is the one the next iteration's cp.async overwrites
    }
  }
  if (sj.lut != nullptr)
    for (int i = tid; i < C * 256; i += kThreads) s_lut[i] = __ldg(sj.lut + i);
  __syncthreads();

  const int rows = y1 - y0;
  if (V4) {
    const int col = tid & 127, rg = tid >> 7;
    for (int r = rg; r < rows; r += 2) {
      const int yo = y0 + r, ry = yo - sj.pad_top;
      const bool in_y = ry >= 0 && ry < new_h;
      const int ymin = in_y ? __ldg(sj.bound_y + 2 * ry) - in_lo : 0, n = in_y ? __ldg(sj.bound_y + 2 * ry + 1) : 0;
      const int* k = sj.coef_y + static_cast<long long>(in_y ? ry : 0) * sj.ks_y;
      for (int xo = col; xo < L; xo += 128) {
        const int rx = xo - sj.pad_left;
        float v[3] = {sj.pad_value[0], sj.pad_value[1], sj.pad_value[2]};
        if (in_y && rx >= 0 && rx < new_w) {
          const unsigned char* s = tmp + ymin * pitch + rx * C;
          int a0 = 1 << (kPrecisionBits - 1), a1 = a0, a2 = a0;
          if (C == 3) {
            for (int y = 0; y < n; ++y) {
              const int kv = __ldg(k + y);
              a0 += s[0] * kv;
              a1 += s[1] * kv;
              a2 += s[2] * kv;
              s += pitch;
            }
          } else {
            for (int y = 0; y < n; ++y) {
              a0 += s[0] * __ldg(k + y);
              s += pitch;
            }
          }
          const int l0 = clip8(a0), l1 = clip8(a1), l2 = clip8(a2);
          const bool lut = sj.lut != nullptr;
          v[0] = lut ? s_lut[l0] : static_cast<float>(l0);
          v[1] = lut ? s_lut[256 + l1] : static_cast<float>(l1);
          v[2] = lut ? s_lut[512 + l2] : static_cast<float>(l2);
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          if (c >= C) break;
          const long long o = (static_cast<long long>(c) * L + yo) * L + xo;
          if (sj.out_dtype == MPL_DT_F32)
            static_cast<float*>(sj.dst)[o] = v[c];
          else if (sj.out_dtype == MPL_DT_BF16)
            static_cast<__nv_bfloat16*>(sj.dst)[o] = __float2bfloat16_rn(v[c]);
          else
            static_cast<unsigned char*>(sj.dst)[o] = static_cast<unsigned char>(v[c]);
        }
      }
    }
    return;
  }
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

// Validates the jobs and sizes the launch: grid.x = bands (the largest band count of any job), dynamic shared memory =
// the largest per-job need. Shared by the CUDA entry below and the CPU emulation harness.
static int plan_launch(const mpl_preprocess_job* jobs_host, int n_jobs, int* bands_out, long long* smem_out) {
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
    if ((reinterpret_cast<uintptr_t>(j.src) & 15) != 0) return MPL_ERR_ALIGN;  // aligned chunks never reach below src
    const int rrb = pick_band(j.H, j.W, j.new_h, j.new_w, j.C);
    if (rrb == 0) return MPL_ERR_UNSUPPORTED;
    const int R = rrb & 255;
    const int b = (j.out_size + R - 1) / R;
    bands = b > bands ? b : bands;
    const long long need = band_smem(j.H, j.W, j.new_h, j.new_w, j.C, rrb);
    smem = need > smem ? need : smem;
  }
  *bands_out = bands;
  *smem_out = smem;
  return MPL_OK;
}

#ifdef MPL_CPU_EMULATION
// variant 0 = default loops, 1 = the MPL_PREPROCESS_V4 candidate; blocks run one after the other, 256 std::threads each
extern "C" int mpl_emu_preprocess_images(const mpl_preprocess_job* jobs, int n_jobs, int variant) {
  if (n_jobs <= 0) return MPL_OK;
  int bands = 0;
  long long smem = 0;
  const int rc = plan_launch(jobs, n_jobs, &bands, &smem);
  if (rc != MPL_OK) return rc;
  if (smem > kSmemBudget) return MPL_ERR_UNSUPPORTED;
  mpl_emu::allowed_ranges().clear();
  for (int i = 0; i < n_jobs; ++i) {  // what the ABI lets the kernel read: the image, rounded up to 16 bytes
    const uintptr_t lo = reinterpret_cast<uintptr_t>(jobs[i].src);
    const uintptr_t hi = lo + static_cast<uintptr_t>(jobs[i].H - 1) * jobs[i].src_stride + jobs[i].W * jobs[i].C;
    mpl_emu::allowed_ranges().push_back({lo, (hi + 15) & ~uintptr_t(15)});
  }
  mpl_emu::run_grid(bands, n_jobs, kThreads, static_cast<size_t>(smem), [&]() {
    if (variant == 1)
      preprocess_kernel<true>(jobs);
    else
      preprocess_kernel<false>(jobs);
  });
  return mpl_emu::violations() == 0 ? MPL_OK : MPL_ERR_CUDA;
}
#else
extern "C" int mpl_preprocess_images(const mpl_preprocess_job* jobs_host, const mpl_preprocess_job* jobs_dev,
                                     int n_jobs, void* stream) {
  if (n_jobs <= 0) return MPL_OK;
  if (jobs_host == nullptr || jobs_dev == nullptr || n_jobs > 65535) return MPL_ERR_ARG;
  int bands = 0;
  long long smem = 0;
  const int rc = plan_launch(jobs_host, n_jobs, &bands, &smem);
  if (rc != MPL_OK) return rc;
  static int variant = -1;  // 0: default loops, 1: MPL_PREPROCESS_V4=1 (candidate, see preprocess_kernel)
  if (variant < 0) {
    if (cudaFuncSetAttribute(preprocess_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget) !=
            cudaSuccess ||
        cudaFuncSetAttribute(preprocess_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget) !=
            cudaSuccess)
      return MPL_ERR_CUDA;
    const char* e = getenv("MPL_PREPROCESS_V4");
    variant = (e != nullptr && e[0] == '1') ? 1 : 0;
  }
  const dim3 grid(bands, n_jobs);
  if (variant == 1)
    preprocess_kernel<true><<<grid, kThreads, static_cast<size_t>(smem), static_cast<cudaStream_t>(stream)>>>(jobs_dev);
  else
    preprocess_kernel<false><<<grid, kThreads, static_cast<size_t>(smem), static_cast<cudaStream_t>(stream)>>>(jobs_dev);
  return mpl::launch_status();
}
#endif
