// Internal C++ declarations shared by the kernel translation units (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>

#include "../../include/medplib_b200.h"

namespace mpl {
int num_sms();
int gemm_bf16(const mpl_gemm_args& a, cudaStream_t stream);
}  // namespace mpl
