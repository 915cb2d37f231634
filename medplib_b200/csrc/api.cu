// Library probe entry points of the C ABI.
#include "internal.h"

#include <cstdlib>

namespace mpl {
unsigned long long g_launches = 0;
bool pdl_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("MPL_PDL");
    on = (e == nullptr || atoi(e) != 0) ? 1 : 0;
  }
  return on != 0;
}
}  // namespace mpl

extern "C" int mpl_version(void) { return 1; }
extern "C" long long mpl_launch_count(void) { return static_cast<long long>(mpl::g_launches); }

extern "C" int mpl_device_info(int* sm_count, int* cc_major, int* cc_minor) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return MPL_ERR_CUDA;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return MPL_ERR_CUDA;
  if (sm_count) *sm_count = prop.multiProcessorCount;
  if (cc_major) *cc_major = prop.major;
  if (cc_minor) *cc_minor = prop.minor;
  return MPL_OK;
}
