// Internal C++ declarations shared by the kernel translation units (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>

#include <cstring>

#include "../../include/medplib_b200.h"

namespace mpl {
// number of kernels this library has launched in this process (bench.py reports the delta over its timed region)
extern unsigned long long g_launches;
inline int launch_status(int n = 1) {
  g_launches += n;
  return cudaGetLastError() == cudaSuccess ? MPL_OK : MPL_ERR_CUDA;
}
int num_sms();

// Programmatic dependent launch: the kernel may become resident (and run its prologue up to its griddep_wait()) while the
// previous kernel of the stream is still finishing. ONLY for kernels that execute griddep_wait() before their first
// read or write of global memory another kernel may touch. MPL_PDL=0 in the environment turns the attribute off.
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
int encode_tmap_bf16(void* out, const void* ptr, int rank, const unsigned long long* dims,
                     const unsigned long long* strides_bytes, const unsigned* box);
int gemm_bf16(const mpl_gemm_args& a, cudaStream_t stream);
int skinny_gemm_bf16(const mpl_gemm_args& a, cudaStream_t stream);
int linear_bf16(const mpl_gemm_args& a, cudaStream_t stream);
int grouped_gemm_bf16(const mpl_grouped_gemm_args& a, cudaStream_t stream);
int skinny_grouped_gemm_bf16(const mpl_grouped_gemm_args& a, cudaStream_t stream);
int moe_route_small(const mpl_moe_route_args& a, const void* x, long long ldx, const void* ln_w, float eps, void* h,
                    long long ldh, void* xperm, int* tok_of_slot, float* gate_of_slot, cudaStream_t stream);
int moe_route(const mpl_moe_route_args& a, cudaStream_t stream);
int moe_norm_route(const mpl_moe_route_args& a, const void* x, long long ldx, const void* ln_w, float eps,
                   cudaStream_t stream);  // RMSNorm fused into the router (D <= 4096), else MPL_ERR_UNSUPPORTED
bool llama_decode_supported(const mpl_llama_model& m, const mpl_llama_io& io);
long long llama_decode_plan_bytes(const mpl_llama_model& m);
int llama_decode_plan_build(const mpl_llama_model& m, void* plan_dev, cudaStream_t st);
int llama_decode_step(const mpl_llama_model& m, const mpl_llama_io& io, void* qkv, void* attn, void* h1,
                      const int* cap_by_e, int emax, cudaStream_t st);
int attn_bwd_fa2(const mpl_attn_bwd_args& a, cudaStream_t stream);
// attention_tc.cu: tcgen05 + TMA flash-attention forward (head_dim 64 / 128, Tq >= 32, no relative-position bias)
bool attention_tc_supported(const mpl_attn_args& a);
int attention_tc(const mpl_attn_args& a, cudaStream_t stream);
// tcgen05 + TMA flash-attention backward (head_dim 128); delta precomputed, dq_f32 zeroed by the caller
bool attention_bwd_tc_supported(const mpl_attn_bwd_args& a);
int attention_bwd_tc(const mpl_attn_bwd_args& a, cudaStream_t stream);
int moe_dispatch(const void* h, long long ldh, const int* slot, void* xperm, int S, int k, int D, cudaStream_t stream);
int moe_combine(const void* y, const int* slot, const float* gate, const void* residual, long long ldr, void* out,
                long long ldo, int S, int k, int D, cudaStream_t stream, const void* ln_w = nullptr, float eps = 0.0f,
                void* h_out = nullptr, long long ldh = 0);  // ln_w: + RMSNorm of the combined row into h_out
}  // namespace mpl
