// Native runners of the SAM-Med2D grounding head: ViT-B image encoder with per-block adapters, and the two-way
// mask decoder for one text prompt.
//
// Encoder replaces ImageEncoderViT.forward (model/segment_anything_med2d/modeling/image_encoder.py:151-162) called
// per image by get_visual_embs (model/MedPLIB.py:274-285); here the whole batch runs together, token-major.
//   patch conv = im2col + GEMM(+bias, +pos_embed as residual)
//   block: LN -> [window partition as a row gather] -> qkv GEMM -> q.Rh / q.Rw terms -> flash attention with the
//          decomposed rel-pos bias added in-kernel -> [unpartition gather] -> proj GEMM (+residual) -> LN ->
//          lin1 GEMM (+GELU) -> lin2 GEMM (+residual) ; adapter: token mean -> 2 tiny GEMMs (ReLU, sigmoid) ->
//          im2col(k3,s2,p1) with the channel gate fused -> GEMM (+ReLU) -> GEMM to f32 cols -> col2im(k4,s2,p1)+ReLU+
//          skip -> LN -> add
//   neck : 1x1 conv GEMM -> LN2d (row LN in token-major) -> im2col(k3) GEMM -> LN2d
// Decoder replaces MaskDecoder.predict_masks (mask_decoder.py:113-153) + TwoWayTransformer (transformer.py:62-244)
// + PromptEncoder.forward text path (prompt_encoder.py:140-187). Token-side linears (6 rows) use the streaming GEMM,
// image-side linears (grid*grid rows) the tcgen05 GEMM; softmax / LayerNorm reductions are warp-shuffle kernels.
#include <cmath>
#include <cstring>

#include "internal.h"

namespace mpl {

static inline long long al(long long v) { return (v + 255) & ~255LL; }

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

static int lin(const void* A, long long lda, int M, const void* W, const void* b, int N, int K, void* C, long long ldc,
               int act, const void* residual, long long ldr, int out_dtype, cudaStream_t st) {
  mpl_gemm_args a = gemm0(A, lda, M, N, K);
  a.B[0] = W;
  a.bias[0] = b;
  a.C[0] = C;
  a.ldc = ldc;
  a.act = act;
  a.residual = residual;
  a.ldr = ldr;
  a.out_dtype = out_dtype;
  return linear_bf16(a, st);
}

// ===================================================================================================== encoder
struct SamEncWs {
  char *cols, *x, *h, *win, *qkv, *attn, *attn_u, *mlp, *pooled, *g1, *gate, *acols, *s1, *u, *ad, *n0, *ncols;
  float *rel_h, *rel_w, *tcols;
  long long total;
};

static SamEncWs sam_enc_carve(const mpl_sam_encoder& m, int B, char* base) {
  const long long g = m.image_size / m.patch, T = g * g, D = m.hidden, O = m.out_chans;
  const long long ws = 14, gp = ((g + ws - 1) / ws) * ws, Tw = gp * gp;  // padded tokens per image in windowed blocks
  const long long Tm = Tw > T ? Tw : T;
  const long long smax = g > ws ? g : ws;
  SamEncWs w;
  long long off = 0;
  auto take = [&](long long bytes) {
    char* p = base ? base + off : nullptr;
    off += al(bytes);
    return p;
  };
  w.cols = take(B * T * 3 * m.patch * m.patch * 2);
  w.x = take(B * T * D * 2);
  w.h = take(B * T * D * 2);
  w.win = take(B * Tm * D * 2);
  w.qkv = take(B * Tm * 3 * D * 2);
  w.attn = take(B * Tm * D * 2);
  w.attn_u = take(B * T * D * 2);
  w.mlp = take(B * T * static_cast<long long>(m.mlp) * 2);
  w.pooled = take(B * D * 2);
  w.g1 = take(B * D * 2);
  w.gate = take(B * D * 2);
  w.acols = take(B * (T / 4) * 9 * D * 2);
  w.s1 = take(B * (T / 4) * D * 2);
  w.u = take(B * T * D * 2);
  w.ad = take(B * T * D * 2);
  w.n0 = take(B * T * O * 2);
  w.ncols = take(B * T * 9 * O * 2);
  w.rel_h = reinterpret_cast<float*>(take(B * m.n_heads * Tm * smax * 4));
  w.rel_w = reinterpret_cast<float*>(take(B * m.n_heads * Tm * smax * 4));
  w.tcols = reinterpret_cast<float*>(take(B * (T / 4) * 16 * D * 4));
  w.total = off;
  return w;
}

int sam_encoder_forward(const mpl_sam_encoder& m, const void* images, int B, const int* win_part,
                        const int* win_unpart, int n_windows, void* out, void* workspace, long long ws_bytes,
                        cudaStream_t st) {
  if (images == nullptr || out == nullptr || workspace == nullptr || m.blocks == nullptr) return MPL_ERR_ARG;
  if (B <= 0) return MPL_OK;
  const int g = m.image_size / m.patch, T = g * g, D = m.hidden, H = m.n_heads, hd = D / H, O = m.out_chans;
  const int S = B * T;
  const SamEncWs w = sam_enc_carve(m, B, static_cast<char*>(workspace));
  if (w.total > ws_bytes) return MPL_ERR_ARG;
  void* s_ = static_cast<void*>(st);
  const int Kp = 3 * m.patch * m.patch;
  MPL_TRY(mpl_im2col_patch(images, w.cols, B, 3, m.image_size, m.image_size, m.patch, Kp, s_));
  // x = conv(img) + bias + pos_embed  (pos_embed rows repeat per image: one GEMM per image keeps the residual simple)
  for (int b = 0; b < B; ++b)
    MPL_TRY(lin(w.cols + static_cast<long long>(b) * T * Kp * 2, Kp, T, m.patch_w, m.patch_b, D, Kp,
                w.x + static_cast<long long>(b) * T * D * 2, D, MPL_ACT_NONE, m.pos_embed, D, MPL_DT_BF16, st));
  for (int i = 0; i < m.depth; ++i) {
    const mpl_sam_block& L = m.blocks[i];
    MPL_TRY(mpl_layernorm(w.x, D, L.ln1_w, L.ln1_b, w.h, D, S, D, 1e-6f, MPL_ACT_NONE, s_));
    const char* attn_in = w.h;
    int nb_win = B, Tw = T, side = g;
    if (L.window > 0) {
      if (win_part == nullptr || win_unpart == nullptr) return MPL_ERR_ARG;
      side = L.window;
      Tw = side * side;
      nb_win = B * n_windows;
      MPL_TRY(mpl_gather_rows(w.h, D, nullptr, D, win_part, w.win, D, nb_win * Tw, D, s_));
      attn_in = w.win;
    }
    const int Sw = nb_win * Tw;
    MPL_TRY(lin(attn_in, D, Sw, L.qkv_w, L.qkv_b, 3 * D, D, w.qkv, 3LL * D, MPL_ACT_NONE, nullptr, 0, MPL_DT_BF16, st));
    MPL_TRY(mpl_sam_relpos(w.qkv, static_cast<long long>(Tw) * 3 * D, 3LL * D, hd, L.rel_pos_h, L.rel_pos_w, w.rel_h,
                           w.rel_w, nb_win, H, side, side, hd, s_));
    {
      mpl_attn_args a;
      memset(&a, 0, sizeof(a));
      a.q = w.qkv;
      a.k = w.qkv + static_cast<long long>(D) * 2;
      a.v = w.qkv + static_cast<long long>(D) * 4;
      a.o = w.attn;
      for (int j = 0; j < 3; ++j) {
        long long* sp = j == 0 ? a.q_stride : (j == 1 ? a.k_stride : a.v_stride);
        sp[0] = static_cast<long long>(Tw) * 3 * D;
        sp[1] = 3LL * D;
        sp[2] = hd;
      }
      a.o_stride[0] = static_cast<long long>(Tw) * D;
      a.o_stride[1] = D;
      a.o_stride[2] = hd;
      a.B = nb_win;
      a.H = H;
      a.Tq = Tw;
      a.Tk = Tw;
      a.head_dim = hd;
      a.scale = 1.0f / sqrtf(static_cast<float>(hd));
      a.rel_h = w.rel_h;
      a.rel_w = w.rel_w;
      a.rel_kh = side;
      a.rel_kw = side;
      MPL_TRY(mpl_attention(&a, s_));
    }
    const char* proj_in = w.attn;
    if (L.window > 0) {
      MPL_TRY(mpl_gather_rows(w.attn, D, nullptr, D, win_unpart, w.attn_u, D, S, D, s_));
      proj_in = w.attn_u;
    }
    MPL_TRY(lin(proj_in, D, S, L.proj_w, L.proj_b, D, D, w.x, D, MPL_ACT_NONE, w.x, D, MPL_DT_BF16, st));
    // x_norm = LN2(x); x = x + mlp(x_norm) (+ Adapter(x_norm))
    MPL_TRY(mpl_layernorm(w.x, D, L.ln2_w, L.ln2_b, w.h, D, S, D, 1e-6f, MPL_ACT_NONE, s_));
    MPL_TRY(lin(w.h, D, S, L.lin1_w, L.lin1_b, m.mlp, D, w.mlp, m.mlp, MPL_ACT_GELU, nullptr, 0, MPL_DT_BF16, st));
    MPL_TRY(lin(w.mlp, m.mlp, S, L.lin2_w, L.lin2_b, D, m.mlp, w.x, D, MPL_ACT_NONE, w.x, D, MPL_DT_BF16, st));
    if (L.ad_ch0 != nullptr) {
      const int Dq = D / 4, Th = (g / 2) * (g / 2);
      MPL_TRY(mpl_col_mean(w.h, w.pooled, B, T, D, s_));
      MPL_TRY(lin(w.pooled, D, B, L.ad_ch0, nullptr, Dq, D, w.g1, Dq, MPL_ACT_RELU, nullptr, 0, MPL_DT_BF16, st));
      MPL_TRY(lin(w.g1, Dq, B, L.ad_ch2, nullptr, D, Dq, w.gate, D, MPL_ACT_SIGMOID, nullptr, 0, MPL_DT_BF16, st));
      MPL_TRY(mpl_im2col_nhwc(w.h, w.gate, w.acols, B, g, g, D, 3, 3, 2, 1, s_));
      MPL_TRY(lin(w.acols, 9LL * D, B * Th, L.ad_conv, nullptr, D, 9 * D, w.s1, D, MPL_ACT_RELU, nullptr, 0,
                  MPL_DT_BF16, st));
      MPL_TRY(lin(w.s1, D, B * Th, L.ad_convt, nullptr, 16 * D, D, w.tcols, 16LL * D, MPL_ACT_NONE, nullptr, 0,
                  MPL_DT_F32, st));
      MPL_TRY(mpl_convt4s2_col2im(w.tcols, w.h, w.u, B, g / 2, g / 2, D, s_));
      MPL_TRY(mpl_layernorm(w.u, D, L.ad_norm_w, L.ad_norm_b, w.ad, D, S, D, 1e-5f, MPL_ACT_NONE, s_));
      MPL_TRY(mpl_add(w.x, w.ad, 0, w.x, static_cast<long long>(S) * D, static_cast<long long>(S) * D, s_));
    }
  }
  // neck
  MPL_TRY(lin(w.x, D, S, m.neck0_w, nullptr, O, D, w.n0, O, MPL_ACT_NONE, nullptr, 0, MPL_DT_BF16, st));
  MPL_TRY(mpl_layernorm(w.n0, O, m.neck1_w, m.neck1_b, w.n0, O, S, O, 1e-6f, MPL_ACT_NONE, s_));
  MPL_TRY(mpl_im2col_nhwc(w.n0, nullptr, w.ncols, B, g, g, O, 3, 3, 1, 1, s_));
  MPL_TRY(lin(w.ncols, 9LL * O, S, m.neck2_w, nullptr, O, 9 * O, out, O, MPL_ACT_NONE, nullptr, 0, MPL_DT_BF16, st));
  MPL_TRY(mpl_layernorm(out, O, m.neck3_w, m.neck3_b, out, O, S, O, 1e-6f, MPL_ACT_NONE, s_));
  return MPL_OK;
}

// ===================================================================================================== mask decoder
struct SamDecWs {
  char *keys, *kpe, *tok, *tpe, *qin, *q, *k, *v, *att, *mlp, *kq, *katt, *up0, *up0n, *up1, *ups, *hy1, *hy2, *hy3;
  long long total;
};

static SamDecWs sam_dec_carve(const mpl_sam_mask_decoder& m, char* base) {
  const long long T = static_cast<long long>(m.grid) * m.grid, D = m.dim, nt = m.n_mask_tokens + 2;
  SamDecWs w;
  long long off = 0;
  auto take = [&](long long bytes) {
    char* p = base ? base + off : nullptr;
    off += al(bytes);
    return p;
  };
  w.keys = take(T * D * 2);
  w.kpe = take(T * D * 2);
  w.tok = take(16 * D * 2);
  w.tpe = take(16 * D * 2);
  w.qin = take(16 * D * 2);
  w.q = take(16 * D * 2);
  w.k = take(T * D * 2);
  w.v = take(T * D * 2);
  w.att = take(16 * D * 2);
  w.mlp = take(16 * static_cast<long long>(m.mlp) * 2);
  w.kq = take(T * D * 2);
  w.katt = take(T * D * 2);
  w.up0 = take(T * D * 2);       // [T, 4 * D/4]
  w.up0n = take(T * D * 2);
  w.up1 = take(T * 4 * (D / 2) * 2);  // [4T, 4 * D/8]
  w.ups = take(T * 16 * (D / 8) * 2);
  w.hy1 = take(16 * D * 2);
  w.hy2 = take(16 * D * 2);
  w.hy3 = take(16 * D * 2);
  w.total = off;
  (void)nt;
  return w;
}

static int attn_call(const void* q, long long ldq, const void* k, long long ldk, const void* v, long long ldv, void* o,
                     long long ldo, int Tq, int Tk, int H, int hd, void* s_) {
  mpl_attn_args a;
  memset(&a, 0, sizeof(a));
  a.q = q;
  a.k = k;
  a.v = v;
  a.o = o;
  a.q_stride[0] = 0;
  a.q_stride[1] = ldq;
  a.q_stride[2] = hd;
  a.k_stride[0] = 0;
  a.k_stride[1] = ldk;
  a.k_stride[2] = hd;
  a.v_stride[0] = 0;
  a.v_stride[1] = ldv;
  a.v_stride[2] = hd;
  a.o_stride[0] = 0;
  a.o_stride[1] = ldo;
  a.o_stride[2] = hd;
  a.B = 1;
  a.H = H;
  a.Tq = Tq;
  a.Tk = Tk;
  a.head_dim = hd;
  a.scale = 1.0f / sqrtf(static_cast<float>(hd));
  return mpl_attention(&a, s_);
}

int sam_mask_decoder_forward(const mpl_sam_mask_decoder& m, const void* image_embedding, const void* text_embed,
                             void* low_res_mask, void* iou, void* workspace, long long ws_bytes, cudaStream_t st) {
  if (image_embedding == nullptr || text_embed == nullptr || low_res_mask == nullptr || iou == nullptr ||
      workspace == nullptr || m.layers == nullptr || m.dense_pe == nullptr || m.shuffle_idx == nullptr)
    return MPL_ERR_ARG;
  const int T = m.grid * m.grid, D = m.dim, H = m.n_heads, nt = m.n_mask_tokens + 2;
  const int Di = D / 2;  // internal dim of the cross attentions (attention_downsample_rate = 2)
  if (nt > 16) return MPL_ERR_UNSUPPORTED;
  const SamDecWs w = sam_dec_carve(m, static_cast<char*>(workspace));
  if (w.total > ws_bytes) return MPL_ERR_ARG;
  void* s_ = static_cast<void*>(st);
  const long long TD = static_cast<long long>(T) * D, ND = static_cast<long long>(nt) * D;
  auto d2d = [&](void* dst, const void* src, size_t bytes) {
    return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, st) == cudaSuccess ? MPL_OK : MPL_ERR_CUDA;
  };
  // tokens = [iou_token ; mask_tokens ; text]; query_pe = tokens (constant); keys = image + no_mask_embed
  MPL_TRY(d2d(w.tpe, m.iou_token, D * 2));
  MPL_TRY(d2d(w.tpe + D * 2, m.mask_tokens, static_cast<size_t>(m.n_mask_tokens) * D * 2));
  MPL_TRY(d2d(w.tpe + static_cast<long long>(m.n_mask_tokens + 1) * D * 2, text_embed, D * 2));
  MPL_TRY(d2d(w.tok, w.tpe, ND * 2));
  MPL_TRY(mpl_add(image_embedding, m.no_mask, 0, w.keys, TD, D, s_));

  auto token_to_image = [&](const mpl_sam_attn& A, const void* nw, const void* nb) -> int {
    // queries += Attn(q = queries + pe, k = keys + key_pe, v = keys); queries = LN(queries)
    MPL_TRY(mpl_add(w.tok, w.tpe, 0, w.qin, ND, ND, s_));
    MPL_TRY(mpl_add(w.keys, m.dense_pe, 1, w.kpe, TD, TD, s_));
    MPL_TRY(lin(w.qin, D, nt, A.q_w, A.q_b, Di, D, w.q, Di, 0, nullptr, 0, MPL_DT_BF16, st));
    MPL_TRY(lin(w.kpe, D, T, A.k_w, A.k_b, Di, D, w.k, Di, 0, nullptr, 0, MPL_DT_BF16, st));
    MPL_TRY(lin(w.keys, D, T, A.v_w, A.v_b, Di, D, w.v, Di, 0, nullptr, 0, MPL_DT_BF16, st));
    MPL_TRY(attn_call(w.q, Di, w.k, Di, w.v, Di, w.att, Di, nt, T, H, Di / H, s_));
    MPL_TRY(lin(w.att, Di, nt, A.o_w, A.o_b, D, Di, w.tok, D, 0, w.tok, D, MPL_DT_BF16, st));
    return mpl_layernorm(w.tok, D, nw, nb, w.tok, D, nt, D, 1e-5f, MPL_ACT_NONE, s_);
  };

  for (int i = 0; i < m.depth; ++i) {
    const mpl_sam_twoway_layer& L = m.layers[i];
    // (1) self attention of the tokens (first layer: no PE, output replaces the queries)
    const char* qk_in = w.tok;
    if (i > 0) {
      MPL_TRY(mpl_add(w.tok, w.tpe, 0, w.qin, ND, ND, s_));
      qk_in = w.qin;
    }
    MPL_TRY(lin(qk_in, D, nt, L.self_attn.q_w, L.self_attn.q_b, D, D, w.q, D, 0, nullptr, 0, MPL_DT_BF16, st));
    MPL_TRY(lin(qk_in, D, nt, L.self_attn.k_w, L.self_attn.k_b, D, D, w.k, D, 0, nullptr, 0, MPL_DT_BF16, st));
    MPL_TRY(lin(w.tok, D, nt, L.self_attn.v_w, L.self_attn.v_b, D, D, w.v, D, 0, nullptr, 0, MPL_DT_BF16, st));
    MPL_TRY(attn_call(w.q, D, w.k, D, w.v, D, w.att, D, nt, nt, H, D / H, s_));
    MPL_TRY(lin(w.att, D, nt, L.self_attn.o_w, L.self_attn.o_b, D, D, w.tok, D, 0, i > 0 ? w.tok : nullptr, D,
                MPL_DT_BF16, st));
    MPL_TRY(mpl_layernorm(w.tok, D, L.n1_w, L.n1_b, w.tok, D, nt, D, 1e-5f, MPL_ACT_NONE, s_));
    // (2) tokens attend to the image
    MPL_TRY(token_to_image(L.t2i, L.n2_w, L.n2_b));
    // (3) MLP on the tokens
    MPL_TRY(lin(w.tok, D, nt, L.lin1_w, L.lin1_b, m.mlp, D, w.mlp, m.mlp, MPL_ACT_RELU, nullptr, 0, MPL_DT_BF16, st));
    MPL_TRY(lin(w.mlp, m.mlp, nt, L.lin2_w, L.lin2_b, D, m.mlp, w.tok, D, 0, w.tok, D, MPL_DT_BF16, st));
    MPL_TRY(mpl_layernorm(w.tok, D, L.n3_w, L.n3_b, w.tok, D, nt, D, 1e-5f, MPL_ACT_NONE, s_));
    // (4) image attends to the tokens: q = keys + key_pe (unchanged since (2)), k = queries + pe, v = queries
    MPL_TRY(mpl_add(w.tok, w.tpe, 0, w.qin, ND, ND, s_));
    MPL_TRY(lin(w.kpe, D, T, L.i2t.q_w, L.i2t.q_b, Di, D, w.kq, Di, 0, nullptr, 0, MPL_DT_BF16, st));
    MPL_TRY(lin(w.qin, D, nt, L.i2t.k_w, L.i2t.k_b, Di, D, w.k, Di, 0, nullptr, 0, MPL_DT_BF16, st));
    MPL_TRY(lin(w.tok, D, nt, L.i2t.v_w, L.i2t.v_b, Di, D, w.v, Di, 0, nullptr, 0, MPL_DT_BF16, st));
    MPL_TRY(attn_call(w.kq, Di, w.k, Di, w.v, Di, w.katt, Di, T, nt, H, Di / H, s_));
    MPL_TRY(lin(w.katt, Di, T, L.i2t.o_w, L.i2t.o_b, D, Di, w.keys, D, 0, w.keys, D, MPL_DT_BF16, st));
    MPL_TRY(mpl_layernorm(w.keys, D, L.n4_w, L.n4_b, w.keys, D, T, D, 1e-5f, MPL_ACT_NONE, s_));
  }
  MPL_TRY(token_to_image(m.final_attn, m.nf_w, m.nf_b));

  // upscaling: convT(k2,s2) -> LN2d -> GELU -> convT(k2,s2) -> GELU, as two GEMMs in (pixel, tap) row order and one
  // final row gather into raster order
  const int C4 = D / 4, C8 = D / 8;
  MPL_TRY(lin(w.keys, D, T, m.up0_w, m.up0_b, 4 * C4, D, w.up0, 4LL * C4, 0, nullptr, 0, MPL_DT_BF16, st));
  MPL_TRY(mpl_layernorm(w.up0, C4, m.up_ln_w, m.up_ln_b, w.up0n, C4, 4 * T, C4, 1e-6f, MPL_ACT_GELU, s_));
  MPL_TRY(lin(w.up0n, C4, 4 * T, m.up1_w, m.up1_b, 4 * C8, C4, w.up1, 4LL * C8, MPL_ACT_GELU, nullptr, 0, MPL_DT_BF16,
              st));
  MPL_TRY(mpl_gather_rows(w.up1, C8, nullptr, C8, m.shuffle_idx, w.ups, C8, 16 * T, C8, s_));
  // hypernetwork MLP of mask token 0 and the mask product; IoU head on the iou token
  const char* mt0 = w.tok + static_cast<long long>(D) * 2;
  MPL_TRY(lin(mt0, D, 1, m.hyper_w[0], m.hyper_b[0], D, D, w.hy1, D, MPL_ACT_RELU, nullptr, 0, MPL_DT_BF16, st));
  MPL_TRY(lin(w.hy1, D, 1, m.hyper_w[1], m.hyper_b[1], D, D, w.hy2, D, MPL_ACT_RELU, nullptr, 0, MPL_DT_BF16, st));
  MPL_TRY(lin(w.hy2, D, 1, m.hyper_w[2], m.hyper_b[2], C8, D, w.hy3, C8, 0, nullptr, 0, MPL_DT_BF16, st));
  MPL_TRY(lin(w.hy3, C8, 1, w.ups, nullptr, 16 * T, C8, low_res_mask, 16LL * T, 0, nullptr, 0, MPL_DT_BF16, st));
  MPL_TRY(lin(w.tok, D, 1, m.iou_w[0], m.iou_b[0], D, D, w.hy1, D, MPL_ACT_RELU, nullptr, 0, MPL_DT_BF16, st));
  MPL_TRY(lin(w.hy1, D, 1, m.iou_w[1], m.iou_b[1], D, D, w.hy2, D, MPL_ACT_RELU, nullptr, 0, MPL_DT_BF16, st));
  MPL_TRY(lin(w.hy2, D, 1, m.iou_w[2], m.iou_b[2], m.n_mask_tokens, D, iou, m.n_mask_tokens, 0, nullptr, 0,
              MPL_DT_BF16, st));
  return MPL_OK;
}

}  // namespace mpl

extern "C" long long mpl_sam_encoder_workspace_bytes(const mpl_sam_encoder* m, int B) {
  if (m == nullptr || B <= 0) return 0;
  return mpl::sam_enc_carve(*m, B, nullptr).total;
}
extern "C" int mpl_sam_encoder_forward(const mpl_sam_encoder* m, const void* images, int B, const int* win_part,
                                       const int* win_unpart, int n_windows, void* out, void* workspace,
                                       long long workspace_bytes, void* stream) {
  if (m == nullptr) return MPL_ERR_ARG;
  return mpl::sam_encoder_forward(*m, images, B, win_part, win_unpart, n_windows, out, workspace, workspace_bytes,
                                  static_cast<cudaStream_t>(stream));
}
extern "C" long long mpl_sam_mask_decoder_workspace_bytes(const mpl_sam_mask_decoder* m) {
  if (m == nullptr) return 0;
  return mpl::sam_dec_carve(*m, nullptr).total;
}
extern "C" int mpl_sam_mask_decoder_forward(const mpl_sam_mask_decoder* m, const void* image_embedding,
                                            const void* text_embed, void* low_res_mask, void* iou, void* workspace,
                                            long long workspace_bytes, void* stream) {
  if (m == nullptr) return MPL_ERR_ARG;
  return mpl::sam_mask_decoder_forward(*m, image_embedding, text_embed, low_res_mask, iou, workspace, workspace_bytes,
                                       static_cast<cudaStream_t>(stream));
}
