/* medplib_b200 — C ABI of the B200 (sm_100a) kernels behind the MedPLIB multimodal hot path.
 *
 * The reference (ShawnHuang497/MedPLIB) has no FFI: its hot path is Python nn.Modules
 * (model/MedPLIB.py:187 MedPLIBForCausalLM, model/LISA.py:180 LISAForCausalLM) calling torch / HF / DeepSpeed
 * ops. The drop-in boundary is therefore those Python classes (medplib_b200/model/*), and THIS header is the
 * C boundary their forward() calls through ctypes: plain device pointers, sizes and a cudaStream_t (as void*).
 *
 * Conventions
 *   - every function returns 0 (MPL_OK) or a negative MPL_ERR_* code; nothing throws across the ABI
 *   - all pointers are DEVICE pointers unless the parameter name ends in _host
 *   - buffers are owned by the caller (torch); the library never frees or retains them beyond the call
 *   - kernels are enqueued on `stream` (the caller's current stream) and never synchronise
 *   - bf16 = uint16_t storage (__nv_bfloat16), row-major, leading dimension in ELEMENTS
 * Each entry cites the reference code whose arithmetic it replaces (paths relative to the reference repo).
 */
#ifndef MEDPLIB_B200_H
#define MEDPLIB_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
  MPL_OK = 0,
  MPL_ERR_ARG = -1,    /* null pointer / bad size */
  MPL_ERR_ALIGN = -2,  /* pointer or leading dimension not 16-byte aligned where TMA needs it */
  MPL_ERR_DRIVER = -3, /* cuTensorMapEncodeTiled unavailable or failed */
  MPL_ERR_CUDA = -4,   /* launch failed: see cudaGetLastError */
  MPL_ERR_UNSUPPORTED = -5
};

enum { MPL_ACT_NONE = 0, MPL_ACT_GELU = 1, MPL_ACT_QUICK_GELU = 2, MPL_ACT_RELU = 3, MPL_ACT_SILU = 4 };
enum { MPL_DT_BF16 = 0, MPL_DT_F32 = 1 };

/* Library / device probe. Returns the ABI version; fills sm count and compute capability when non-null. */
int mpl_version(void);
int mpl_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ---------------------------------------------------------------------------------------------------------
 * K1  C[M,N] = epilogue(A[M,K] · W[N,K]^T)      tcgen05 + TMA, bf16 in, fp32 accumulate in TMEM
 * Replaces nn.Linear / F.linear on the dense path: HF LlamaAttention q/k/v/o_proj and LlamaMLP gate/up/down
 * (driven from model/medplib/model/language_model/medplib_moe_llama.py:123-147), CLIP linears
 * (model/medplib/model/multimodal_encoder/clip_encoder.py:53-57), mm_projector
 * (model/medplib/model/multimodal_projector/builder.py:39-46), region_fea_adapter / TokenCompressor.proj
 * (model/medplib/model/medplib_arch.py:73,131), text_hidden_fcs (model/MedPLIB.py:153-164), SAM-Med2D linears
 * and im2col'ed convs (model/segment_anything_med2d/modeling/image_encoder.py).
 * Epilogue order (each step rounds to bf16 first when the output is bf16, like the reference's eager ops):
 *   acc (+ bias) -> act -> (* row_scale[m]) -> (+ residual[m,n]) -> store
 * B2 != NULL selects the fused LlamaMLP front half: out[m,n] = silu(A·B[n]^T) * (A·B2[n]^T).
 * m_dev != NULL: the effective M is min(M, *m_dev) read on the device (MoE expert loads without a host sync).
 */
typedef struct {
  const void* A;   /* bf16 [M,K] */
  long long lda;
  const void* B;   /* bf16 [N,K]  (nn.Linear.weight layout) */
  const void* B2;  /* bf16 [N,K] or NULL */
  long long ldb;
  void* C;         /* bf16 or f32 [M,N] */
  long long ldc;
  const void* bias;     /* bf16 [N] or NULL */
  const void* residual; /* bf16 [M,N] or NULL */
  long long ldr;
  const float* row_scale; /* f32 [M] or NULL */
  const int* m_dev;       /* device int or NULL */
  int M, N, K;
  int act;       /* MPL_ACT_* */
  int out_dtype; /* MPL_DT_* */
  int tile_n;    /* 0 = auto, else 128 or 256 */
} mpl_gemm_args;
int mpl_gemm_bf16(const mpl_gemm_args* args, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MEDPLIB_B200_H */
