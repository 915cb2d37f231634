// Native runner of the CLIP ViT-L/14-336 vision tower as MedPLIB uses it: hidden_states[select_layer][:, 1:].
//
// Replaces CLIPVisionTower.forward (model/medplib/model/multimodal_encoder/clip_encoder.py:41-60) over HF-4.31
// CLIPVisionTransformer. Only the layers up to the selected hidden state are run (the reference
// runs all 24 and discards the last). Per layer: LayerNorm -> fused q,k,v GEMM (+bias) -> flash attention (577x577,
// head_dim 64, non-causal) -> out_proj GEMM (+bias, +residual) -> LayerNorm -> fc1 GEMM (+bias, quick-GELU) ->
// fc2 GEMM (+bias, +residual). The patch-embedding conv is an im2col + tcgen05 GEMM.
#include <cmath>
#include <cstring>

#include "internal.h"

namespace mpl {

static inline long long al(long long v) { return (v + 255) & ~255LL; }

struct ClipWs {
  char *cols, *patch, *x, *h, *qkv, *attn, *mlp;
  long long total;
};

static ClipWs clip_carve(const mpl_clip_model& m, int B, char* base) {
  const long long g = m.image_size / m.patch, np = g * g, T = np + 1, D = m.hidden;
  ClipWs w;
  long long off = 0;
  auto take = [&](long long bytes) {
    char* p = base ? base + off : nullptr;
    off += al(bytes);
    return p;
  };
  w.cols = take(B * np * m.k_pad * 2);
  w.patch = take(B * np * D * 2);
  w.x = take(B * T * D * 2);
  w.h = take(B * T * D * 2);
  w.qkv = take(B * T * 3 * D * 2);
  w.attn = take(B * T * D * 2);
  w.mlp = take(B * T * static_cast<long long>(m.mlp) * 2);
  w.total = off;
  return w;
}

static mpl_gemm_args gemm0(const void* A, long long lda, int M, int N, int K) {
  mpl_gemm_args g;
  memset(&g, 0, sizeof(g));
  g.A = A;
  g.lda = lda;
  g.ldb = K;
  g.ldc = N;
  g.M = M;
  g.N = N;
  g.K = K;
  g.nb = 1;
  return g;
}

#define MPL_TRY(expr)                \
  do {                               \
    const int rc__ = (expr);         \
    if (rc__ != MPL_OK) return rc__; \
  } while (0)

int clip_forward(const mpl_clip_model& m, const void* images, int B, void* feats, void* workspace, long long ws_bytes,
                 cudaStream_t st) {
  if (images == nullptr || feats == nullptr || workspace == nullptr || m.layers == nullptr) return MPL_ERR_ARG;
  if (B <= 0) return MPL_OK;
  const int g = m.image_size / m.patch, np = g * g, T = np + 1, D = m.hidden, H = m.n_heads, hd = D / H;
  const int S = B * T;
  const ClipWs w = clip_carve(m, B, static_cast<char*>(workspace));
  if (w.total > ws_bytes) return MPL_ERR_ARG;
  void* s_ = static_cast<void*>(st);
  MPL_TRY(mpl_im2col_patch(images, w.cols, B, 3, m.image_size, m.image_size, m.patch, m.k_pad, s_));
  {
    mpl_gemm_args a = gemm0(w.cols, m.k_pad, B * np, D, m.k_pad);
    a.B[0] = m.patch_w;
    a.C[0] = w.patch;
    MPL_TRY(linear_bf16(a, st));
  }
  MPL_TRY(mpl_clip_embed(w.patch, m.cls, m.pos, w.h, B, np, D, s_));
  MPL_TRY(mpl_layernorm(w.h, D, m.pre_ln_w, m.pre_ln_b, w.x, D, S, D, m.ln_eps, MPL_ACT_NONE, s_));
  for (int l = 0; l < m.n_layers; ++l) {
    const mpl_clip_layer& L = m.layers[l];
    MPL_TRY(mpl_layernorm(w.x, D, L.ln1_w, L.ln1_b, w.h, D, S, D, m.ln_eps, MPL_ACT_NONE, s_));
    {
      mpl_gemm_args a = gemm0(w.h, D, S, D, D);
      a.nb = 3;
      a.B[0] = L.wq;
      a.B[1] = L.wk;
      a.B[2] = L.wv;
      a.bias[0] = L.bq;
      a.bias[1] = L.bk;
      a.bias[2] = L.bv;
      a.C[0] = w.qkv;
      a.C[1] = w.qkv + static_cast<long long>(D) * 2;
      a.C[2] = w.qkv + static_cast<long long>(D) * 4;
      a.ldc = 3LL * D;
      MPL_TRY(linear_bf16(a, st));
    }
    {
      mpl_attn_args a;
      memset(&a, 0, sizeof(a));
      a.q = w.qkv;
      a.k = w.qkv + static_cast<long long>(D) * 2;
      a.v = w.qkv + static_cast<long long>(D) * 4;
      a.o = w.attn;
      for (int i = 0; i < 3; ++i) {
        long long* sp = i == 0 ? a.q_stride : (i == 1 ? a.k_stride : a.v_stride);
        sp[0] = static_cast<long long>(T) * 3 * D;
        sp[1] = 3LL * D;
        sp[2] = hd;
      }
      a.o_stride[0] = static_cast<long long>(T) * D;
      a.o_stride[1] = D;
      a.o_stride[2] = hd;
      a.B = B;
      a.H = H;
      a.Tq = T;
      a.Tk = T;
      a.head_dim = hd;
      a.scale = 1.0f / sqrtf(static_cast<float>(hd));  // 4.31 scales q before the bmm: identical for power-of-two hd
      MPL_TRY(mpl_attention(&a, s_));
    }
    {
      mpl_gemm_args a = gemm0(w.attn, D, S, D, D);
      a.B[0] = L.wo;
      a.bias[0] = L.bo;
      a.C[0] = w.x;
      a.residual = w.x;
      a.ldr = D;
      MPL_TRY(linear_bf16(a, st));
    }
    MPL_TRY(mpl_layernorm(w.x, D, L.ln2_w, L.ln2_b, w.h, D, S, D, m.ln_eps, MPL_ACT_NONE, s_));
    {
      mpl_gemm_args a = gemm0(w.h, D, S, m.mlp, D);
      a.B[0] = L.fc1_w;
      a.bias[0] = L.fc1_b;
      a.C[0] = w.mlp;
      a.act = MPL_ACT_QUICK_GELU;
      MPL_TRY(linear_bf16(a, st));
    }
    {
      mpl_gemm_args a = gemm0(w.mlp, m.mlp, S, D, m.mlp);
      a.B[0] = L.fc2_w;
      a.bias[0] = L.fc2_b;
      a.C[0] = w.x;
      a.residual = w.x;
      a.ldr = D;
      MPL_TRY(linear_bf16(a, st));
    }
  }
  // drop the CLS row of every image
  if (cudaMemcpy2DAsync(feats, static_cast<size_t>(np) * D * 2, w.x + static_cast<long long>(D) * 2,
                        static_cast<size_t>(T) * D * 2, static_cast<size_t>(np) * D * 2, B, cudaMemcpyDeviceToDevice,
                        st) != cudaSuccess)
    return MPL_ERR_CUDA;
  return MPL_OK;
}

}  // namespace mpl

extern "C" long long mpl_clip_workspace_bytes(const mpl_clip_model* m, int B) {
  if (m == nullptr || B <= 0) return 0;
  return mpl::clip_carve(*m, B, nullptr).total;
}

extern "C" int mpl_clip_forward(const mpl_clip_model* m, const void* images, int B, void* feats, void* workspace,
                                long long workspace_bytes, void* stream) {
  if (m == nullptr) return MPL_ERR_ARG;
  return mpl::clip_forward(*m, images, B, feats, workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
}
