// Minimal CPU stand-ins for the CUDA constructs csrc/preprocess.cu uses, so that g++ can compile the kernel unchanged
// and run it with one std::thread per CUDA thread (TEST INFRASTRUCTURE — see tests/dev/README.md).
//   threadIdx / blockIdx : thread_local          __syncthreads : std::barrier over the block's threads
//   __shared__ variables : function-level statics (blocks run one after the other, so they are per-block in effect)
//   dynamic shared memory: one 16-byte-aligned arena, poisoned before every block
//   cp.async 16-byte copy : memcpy, checked against the address ranges the C ABI allows the kernel to read
#pragma once
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <barrier>
#include <cmath>
#include <functional>
#include <thread>
#include <utility>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __shared__ static
#define __restrict__ __restrict

struct emu_uint3 {
  unsigned x, y, z;
};
inline thread_local emu_uint3 threadIdx{0, 0, 0}, blockIdx{0, 0, 0};

struct __nv_bfloat16 {
  uint16_t bits;
};
inline __nv_bfloat16 __float2bfloat16_rn(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return {static_cast<uint16_t>((u >> 16) | 0x40)};  // NaN
  u += 0x7fffu + ((u >> 16) & 1u);  // round to nearest even
  return {static_cast<uint16_t>(u >> 16)};
}
template <class T>
inline T __ldg(const T* p) {
  return *p;
}
inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned shift) {
  return static_cast<unsigned>(((static_cast<uint64_t>(hi) << 32) | lo) >> (shift & 31));
}
using std::max;
using std::min;

namespace mpl_emu {
inline std::barrier<>*& block_barrier() {
  static std::barrier<>* b = nullptr;
  return b;
}
inline unsigned char* dynamic_smem() {
  alignas(16) static unsigned char arena[232 * 1024];
  return arena;
}
inline std::vector<std::pair<uintptr_t, uintptr_t>>& allowed_ranges() {
  static std::vector<std::pair<uintptr_t, uintptr_t>> r;
  return r;
}
inline std::atomic<long>& violation_count() {
  static std::atomic<long> v{0};
  return v;
}
inline long violations() { return violation_count().load(); }
inline void copy16(unsigned char* dst, uintptr_t src) {
  bool ok = (src & 15) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0;
  bool inside = false;
  for (const auto& r : allowed_ranges()) inside = inside || (src >= r.first && src + 16 <= r.second);
  if (!ok || !inside) {
    violation_count()++;
    memset(dst, 0xEE, 16);
    return;
  }
  memcpy(dst, reinterpret_cast<const void*>(src), 16);
}
inline void run_grid(int gx, int gy, int threads, size_t smem_bytes, const std::function<void()>& kernel) {
  violation_count() = 0;
  for (int by = 0; by < gy; ++by)
    for (int bx = 0; bx < gx; ++bx) {
      memset(dynamic_smem(), 0xCD, smem_bytes);
      std::barrier<> bar(threads);
      block_barrier() = &bar;
      std::vector<std::thread> pool;
      pool.reserve(threads);
      for (int t = 0; t < threads; ++t)
        pool.emplace_back([&, t]() {
          threadIdx = {static_cast<unsigned>(t), 0, 0};
          blockIdx = {static_cast<unsigned>(bx), static_cast<unsigned>(by), 0};
          kernel();
          bar.arrive_and_drop();  // a thread that has returned no longer takes part in later barriers
        });
      for (auto& th : pool) th.join();
    }
}
}  // namespace mpl_emu
inline void __syncthreads() { mpl_emu::block_barrier()->arrive_and_wait(); }
