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

enum { MPL_ACT_NONE = 0, MPL_ACT_GELU = 1, MPL_ACT_QUICK_GELU = 2, MPL_ACT_RELU = 3, MPL_ACT_SILU = 4, MPL_ACT_SIGMOID = 5 };
enum { MPL_DT_BF16 = 0, MPL_DT_F32 = 1, MPL_DT_U8 = 2 /* mpl_preprocess_images only */ };
#define MPL_MAX_EXPERTS 8

/* Library / device probe. Returns the ABI version; fills sm count and compute capability when non-null. */
int mpl_version(void);
int mpl_device_info(int* sm_count, int* cc_major, int* cc_minor);
/* Kernels launched by this library in this process so far (bench.py's gpu_launches is a delta of this). */
long long mpl_launch_count(void);

/* ---------------------------------------------------------------------------------------------------------
 * K1  C[M,N] = epilogue(A[M,K] · W[N,K]^T)      tcgen05 + TMA, bf16 in, fp32 accumulate in TMEM
 * Replaces nn.Linear / F.linear on the dense path: HF LlamaAttention q/k/v/o_proj and LlamaMLP gate/up/down
 * (driven from model/medplib/model/language_model/medplib_moe_llama.py:123-147), CLIP linears
 * (model/medplib/model/multimodal_encoder/clip_encoder.py:53-57), mm_projector
 * (model/medplib/model/multimodal_projector/builder.py:39-46), region_fea_adapter / TokenCompressor.proj
 * (model/medplib/model/medplib_arch.py:73,131), text_hidden_fcs (model/MedPLIB.py:153-164), SAM-Med2D linears
 * and im2col'ed convs (model/segment_anything_med2d/modeling/image_encoder.py).
 * Epilogue order (each step rounds to bf16 first when the output is bf16, like the reference's eager ops):
 *   acc (+ bias) -> act -> (* row_scale[m]) -> (+ LoRA terms) -> (+ residual[m,n]) -> store
 * nb in 1..3 weight matrices of identical shape share A in ONE launch (q/k/v projections): output i goes to C[i]
 *   with bias[i]; this fills the 148 SMs where a single N=4096 projection leaves 40% of them idle.
 * B2 != NULL (nb == 1) selects the fused LlamaMLP front half: out[m,n] = silu(A·B[0][n]^T) * (A·B2[n]^T).
 * m_dev != NULL: the effective M is min(M, *m_dev) read on the device (MoE expert loads without a host sync).
 */
typedef struct {
  const void* A;   /* bf16 [M,K] */
  long long lda;
  const void* B[3]; /* bf16 [N,K]  (nn.Linear.weight layout) */
  const void* B2;   /* bf16 [N,K] or NULL */
  long long ldb;
  void* C[3];       /* bf16 or f32 [M,N] */
  long long ldc;
  const void* bias[3];  /* bf16 [N] or NULL */
  const void* residual; /* bf16 [M,N] or NULL */
  long long ldr;
  const float* row_scale; /* f32 [M] or NULL */
  const int* m_dev;       /* device int or NULL */
  int M, N, K;
  int nb;        /* number of weight matrices, 1..3 */
  int act;       /* MPL_ACT_* */
  int out_dtype; /* MPL_DT_* */
  int tile_n;    /* 0 = auto, else a multiple of 16 in [128, 256] (144, 160, ... 256 are compiled in) */
  const void* ln_weight; /* streaming (M <= 16) path only: fused LlamaRMSNorm prologue on A, bf16 [K] or NULL */
  float ln_eps;
  /* Tensor-core path only, bf16 output: up to two fused rank-r LoRA up-projections (peft Linear.forward / its dgrad,
   * train_ds_medplib.py:294-302), applied after the activation and before the residual, in peft's bf16 rounding:
   *   C[lora_mat[t]][m,n] = bf16(C + bf16(lora_scale[t] * bf16(sum_j bf16(u_t[m,j]) * b_t[n,j])))   for t = 0, 1 in order.
   * u_t: [M, lora_r] row-major (bf16, or f32 when lora_u_f32[t]); b_t: bf16 [N, lora_r] row-major (lora_B.weight; for
   * the dgrad the transposed lora_A); lora_r in {0 (off), 8}; unused terms have lora_u[t] == NULL. */
  const void* lora_u[2];
  const void* lora_b[2];
  float lora_scale[2];
  int lora_u_f32[2];
  int lora_mat[2];
  int lora_r;
  /* Tensor-core path only: ONE extra k-block (64 columns of K) appended to the contraction, C[i] = epilogue(A B[i]^T +
   * ext_a ext_b[i]^T): the rank-r adapters of a LoRA-wrapped Linear (or of its dgrad) enter the fp32 accumulator on the
   * tensor cores -- ext_a = [u_0 | u_1 | ... | 0] bf16 [M, 64] (adapter t in columns [8t, 8t+8)), ext_b = bf16
   * [(nb, or 2 with B2) * N, 64], matrix-major, rows of matrix i = [s B_0 | ... ] with zeros for the adapters that do not
   * feed it. One rounding (the GEMM's) instead of peft's three; NULL = off. */
  const void* ext_a;
  const void* ext_b;
  /* B2 != NULL, tensor-core path: also store the two projections themselves, bf16 [M, N] with row pitch ldc each (the
   * train forward keeps gate(x) and up(x) for the backward); NULL = only silu(gate) * up goes out. */
  void* dual_g;
  void* dual_u;
  /* Tensor-core path, nb == 1: this GEMM is the dgrad of LlamaMLP.down_proj (its result is dh) and the SiLU(gate) * up
   * backward runs in its epilogue: g <- bf16(dh u s (1 + g (1 - s))), u <- bf16(dh g s), s = sigmoid(g), in place over
   * bf16 [M, N] buffers with row pitch ldc; C may be NULL (dh itself is not stored). */
  void* silu_bwd_g;
  void* silu_bwd_u;
} mpl_gemm_args;
int mpl_gemm_bf16(const mpl_gemm_args* args, void* stream);
/* In-situ timing of the tcgen05 GEMM launches (bench.py roofline): enable, run, then read the summed CUDA-event
 * durations (ms) and the launch count since enabling; reading synchronises the device and resets the counters. */
int mpl_profile_gemm(int enable);
int mpl_profile_gemm_read(float* total_ms, int* launches);
/* The same for the persistent decode-step kernel (one launch = one token step of all layers; replaces the per-token
 * HF forward of generate(), model/MedPLIB.py:592-606). */
int mpl_profile_decode(int enable);
int mpl_profile_decode_read(float* total_ms, int* launches);

/* K3  same contract as mpl_gemm_bf16 for M <= 16 (decode batch, [SEG] rows, mask-decoder tokens): HBM-bound
 * streaming kernel, weights read once with 16-byte loads into mma.sync fragments (nb <= 3 as grid.y).
 * mpl_linear_bf16 dispatches on M (<= 16 -> skinny, else tcgen05). */
int mpl_skinny_gemm_bf16(const mpl_gemm_args* args, void* stream);
int mpl_linear_bf16(const mpl_gemm_args* args, void* stream);

/* The experts of one MoE layer in ONE launch: group g multiplies rows [g*a_group_stride/lda ...) of A — the expert
 * buffer written by mpl_moe_dispatch / mpl_moe_route_small — by B[g] (and B2[g]: fused SiLU(gate)*up), with the row
 * count of every group read on the device (m_dev[g]; an expert without tokens costs nothing). row_map/row_gate (M <= 16
 * only): instead of writing group-major rows, scatter row m of group g to output row row_map[g*map_group_stride+m]
 * scaled by bf16(row_gate[..]) and add residual there — the DeepSpeed combine (einsum 'sec,ecm->sm') fused into the
 * down projection. */
typedef struct {
  const void* A;
  long long lda, a_group_stride;
  const void* B[MPL_MAX_EXPERTS];
  const void* B2[MPL_MAX_EXPERTS];
  long long ldb;
  void* C;
  long long ldc, c_group_stride;
  const void* residual;
  long long ldr;
  const int* m_dev;
  const int* a_row_map; /* M <= 16 only: A row m of group g = A[a_row_map[g*map_group_stride+m]] (fused MoE dispatch) */
  const int* row_map;
  const float* row_gate;
  long long map_group_stride;
  int groups, M, N, K, act, out_dtype;
  int m_dev_stable; /* 1: m_dev was already final before the PREVIOUS launch on this stream began, so the weight
                       stream may start before the dependency wait (programmatic dependent launch) */
  int m_total_hint; /* 0, or the number of rows all groups hold together (host-side estimate, e.g. S * top_k): picks the
                       tile width of the tensor-core path; the true per-group counts are always read from m_dev */
  int tile_n;       /* 0 = auto, else a multiple of 16 in [128, 256] (tensor-core path only) */
} mpl_grouped_gemm_args;
int mpl_grouped_gemm_bf16(const mpl_grouped_gemm_args* args, void* stream);

/* K5  LlamaRMSNorm (transformers 4.31 semantics; used at
 * model/medplib/model/language_model/medplib_moe_llama.py:123,139,286): y = w * bf16(x * rsqrt(mean(x^2)+eps)). */
int mpl_rmsnorm(const void* x, long long ldx, const void* weight, void* y, long long ldy, int rows, int D, float eps,
                void* stream);
/* K5  nn.LayerNorm over the last dim (CLIP, SAM blocks, mask decoder; LayerNorm2d of
 * model/segment_anything_med2d/modeling/common.py:31-45 on NHWC rows). act = MPL_ACT_GELU fuses the GELU that
 * follows LayerNorm2d in mask_decoder.py:53-59. D % 8 == 0, D <= 4096. */
int mpl_layernorm(const void* x, long long ldx, const void* weight, const void* bias, void* y, long long ldy,
                  int rows, int D, float eps, int act, void* stream);
/* K6  TokenCompressor front half (model/medplib/model/medplib_arch.py:67-77): AdaptiveAvgPool1d(t_in -> t_out)
 * over tokens fused with LayerNorm(D). x [n,t_in,D] -> y [n,t_out,D], contiguous bf16. */
int mpl_pool_layernorm(const void* x, const void* weight, const void* bias, void* y, int n, int t_in, int t_out,
                       int D, float eps, void* stream);

/* K2 / K2b / K2c / K3  softmax(scale * q k^T + bias + masks) v, fp32 softmax, bf16 in/out.
 * Replaces the eager matmul-softmax-matmul of HF LlamaAttention (causal + key padding),
 * HF CLIPAttention (App. A.2), SAM-Med2D Attention with decomposed relative positions
 * (model/segment_anything_med2d/modeling/image_encoder.py:280-296,381-421) and the mask decoder's Attention
 * (model/segment_anything_med2d/modeling/transformer.py:185-244).
 * Tensors are addressed base + b*stride[0] + t*stride[1] + h*stride[2] + d (elements; innermost contiguous).
 * head_dim in {16,32,64,128}. Tq == 1 with head_dim 128 takes the HBM-bound KV-cache decode kernel, where
 * tk_dev (device int) may override Tk so one CUDA graph serves every decode step. */
typedef struct {
  const void* q;
  const void* k;
  const void* v;
  void* o;
  long long q_stride[3], k_stride[3], v_stride[3], o_stride[3];
  int B, H, Tq, Tk, head_dim;
  float scale;
  int causal;                   /* key j visible to query i iff j <= i + (Tk - Tq) */
  const unsigned char* kv_mask; /* [B,Tk] 1 = attend, or NULL */
  long long kv_mask_stride;     /* row stride of kv_mask in bytes; 0 = Tk */
  const float* rel_h;           /* [B*H,Tq,rel_kh] f32 or NULL: bias = rel_h[q][k / rel_kw] + rel_w[q][k % rel_kw] */
  const float* rel_w;           /* [B*H,Tq,rel_kw] */
  int rel_kh, rel_kw;
  const int* tk_dev;
  void* scratch;             /* decode only, optional: zero-initialised once, >= B*H*4 + B*H*nsplit*520 bytes; enables
                                split-K over the keys (flash-decoding) so B*H < #SM still fills the GPU */
  long long scratch_bytes;
  float* lse; /* Tq > 1 only, optional (training): f32 [B*H, Tq] out, log2-domain log-sum-exp of the scaled + masked scores
                 (row i: max_j s_ij*log2e + log2 sum_j 2^(...)), consumed by mpl_attention_bwd; +inf for a fully masked row */
} mpl_attn_args;
int mpl_attention(const mpl_attn_args* args, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * K4  MoE router + top-k token scatter/gather. Replaces deepspeed.moe.layer.MoE (DeepSpeed 0.13.1 TopKGate /
 * top1gating / top2gating / MOELayer einsum dispatch+combine) as the reference calls it at
 * model/MedPLIB.py:253-263 and model/medplib/model/language_model/medplib_moe_llama.py:141-147,604-614.
 * mpl_moe_route: logits = h.float() @ wg^T (fp32), gates = softmax, top-k selection (k = 1 or 2), slot of every
 *   (token, route) in the [E*capacity, D] expert buffer in cumsum (position) order, -1 when dropped by capacity.
 *   noise (f32 [S,E] or NULL): k = 1 -> the Random-Token-Selection uniforms used when an expert overflows;
 *   k = 2 -> the Gumbel noise added to the logits for the second choice. gate[s,j] = combine weight (0 if dropped;
 *   for k = 2 the two kept values renormalised by their sum). kept[e] = rows used in expert e's buffer (feeds
 *   mpl_gemm_args.m_dev, no host sync); exp_counts[e] = first-choice tokens before dropping; l_aux as DeepSpeed.
 * mpl_moe_dispatch: xperm[slot[s,j]] = h[s].   mpl_moe_combine: out[s] = residual[s] + bf16(sum_j bf16(gate[s,j]) *
 *   y[slot[s,j]]) (residual may be NULL). */
typedef struct {
  const void* h; /* bf16 [S,D] */
  long long ldh;
  const float* wg;    /* f32 [E,D] */
  const float* noise; /* f32 [S,E] or NULL */
  int S, D, E, k, capacity;
  float* logits;   /* [S,E] out */
  float* gates;    /* [S,E] out */
  int* expert;     /* [S,k] out */
  float* gate;     /* [S,k] out */
  int* slot;       /* [S,k] out */
  int* kept;       /* [E] out */
  int* exp_counts; /* [E] out */
  float* l_aux;    /* [1] out */
} mpl_moe_route_args;
int mpl_moe_route(const mpl_moe_route_args* args, void* stream);
/* Decode-time variant for S <= 64 tokens, ONE launch: h = RMSNorm(x) (ln_weight may be NULL: x is already normalised),
 * route, assign slots, copy the rows into xperm; also writes slot -> token (tok_of_slot) and per-slot gate values for
 * mpl_grouped_gemm_bf16's fused combine. Not for k = 1 with RTS noise (training). */
int mpl_moe_route_small(const mpl_moe_route_args* args, const void* x, long long ldx, const void* ln_weight,
                        float ln_eps, void* h, long long ldh, void* xperm, int* tok_of_slot, float* gate_of_slot,
                        void* stream);
int mpl_moe_dispatch(const void* h, long long ldh, const int* slot, void* xperm, int S, int k, int D, void* stream);
int mpl_moe_combine(const void* y, const int* slot, const float* gate, const void* residual, long long ldr, void* out,
                    long long ldo, int S, int k, int D, void* stream);

/* RoPE (HF 4.31 apply_rotary_pos_emb with rotate_half; bf16 cos/sin tables [max_pos, head_dim])
 * applied in place to q and k rows ([B*T] rows of leading dimension ld, H heads of head_dim), position of row (b,t) =
 * (pos_dev ? *pos_dev : pos0) + t, and append of the rotated k and of v to the KV cache [B,H,Tmax,head_dim]
 * (k_cache/v_cache may be NULL). */
int mpl_rope_kv(void* q, void* k, const void* v, long long ld, const void* cos_t, const void* sin_t, void* k_cache,
                void* v_cache, int B, int T, int H, int head_dim, int Tmax, int pos0, const int* pos_dev, void* stream);

/* K9  row gather: out[r] = idx[r] >= 0 ? table[idx[r]] : idx[r] == -1 ? 0 : feats[-idx[r]-2]. One kernel for the
 * embedding lookup + multimodal splice (model/medplib/model/medplib_arch.py:296-527), SAM window (un)partition
 * (image_encoder.py:299-345) and the pixel shuffle of the k2/s2 transposed convs (mask_decoder.py:53-59). */
int mpl_gather_rows(const void* table, long long ld_table, const void* feats, long long ld_feats, const int* idx,
                    void* out, long long ld_out, int rows, int D, void* stream);

/* Greedy token selection (HF greedy_search argmax; lowest index wins ties): out[r] = argmax_v x[r, v]. */
int mpl_argmax_f32(const float* x, long long ld, int rows, int V, long long* out, void* stream);

/* Patch-embedding im2col for stride == kernel convs (CLIPVisionEmbeddings; SAM PatchEmbed image_encoder.py:424-455):
 * img bf16 [B,C,H,W] -> out [B*(H/P)*(W/P), Kpad], column (c*P+ky)*P+kx, columns >= C*P*P zero. */
int mpl_im2col_patch(const void* img, void* out, int B, int C, int H, int W, int P, int Kpad, void* stream);
/* General im2col on token-major activations x bf16 [B,H,W,C] -> [B*Ho*Wo, kh*kw*C], column (ky*kw+kx)*C+c, zero
 * padding; gate (bf16 [B,C] or NULL) multiplies x per channel first (Adapter_Layer channel gate, image_encoder.py:44-46). */
int mpl_im2col_nhwc(const void* x, const void* gate, void* out, int B, int H, int W, int C, int kh, int kw, int stride,
                    int pad, void* stream);
/* CLIPVisionEmbeddings tail: out[b,0] = cls + pos[0]; out[b,1+i] = patch[b,i] + pos[1+i]. */
int mpl_clip_embed(const void* patch, const void* cls, const void* pos, void* out, int B, int n_patches, int D,
                   void* stream);
/* add_decomposed_rel_pos terms (image_encoder.py:381-421): rel_h[bh,t,k] = bf16(q[b,t,h,:] . rel_pos_h[y-k+hh-1]),
 * rel_w likewise with x; outputs f32 [B*H, hh*ww, hh] / [.., ww] for mpl_attention. */
int mpl_sam_relpos(const void* q, long long q_sb, long long q_st, long long q_sh, const void* rel_pos_h,
                   const void* rel_pos_w, float* rel_h, float* rel_w, int B, int H, int hh, int ww, int head_dim,
                   void* stream);
/* Mean over tokens: x bf16 [B,T,C] -> out bf16 [B,C] (Adapter_Layer avg_pool, image_encoder.py:43-45). */
int mpl_col_mean(const void* x, void* out, int B, int T, int C, void* stream);
/* col2im of ConvTranspose2d(k=4,s=2,p=1) computed as cols = x @ W' (f32 [B*Hi*Wi, 16*C], column (ky*4+kx)*C+co):
 * out[b,oy,ox,:] = skip + relu(bf16(sum of taps)) (Adapter_Layer spatial branch + skip, image_encoder.py:33-36,48-51). */
int mpl_convt4s2_col2im(const float* cols, const void* skip, void* out, int B, int Hi, int Wi, int C, void* stream);
/* out = bf16(a + b); b is bf16 or f32 and repeats every b_period elements (positional-encoding adds,
 * transformer.py:160-175 — the dense PE is fp32, App. B-8). */
int mpl_add(const void* a, const void* b, int b_is_f32, void* out, long long n, long long b_period, void* stream);
/* F.interpolate(bilinear, align_corners=False) of postprocess_masks (model/MedPLIB.py:682-701): in bf16
 * [N, Hin, Win] with strides (elements) -> out [N,Hout,Wout] contiguous, out_dtype MPL_DT_*. */
int mpl_bilinear_resize(const void* in, long long in_stride_n, long long in_stride_y, int Hin, int Win, void* out,
                        int out_dtype, int Hout, int Wout, int N, void* stream);
/* extract_region_feature (model/medplib/model/medplib_arch.py:580-614): mean over P points of the bilinear
 * (align_corners=True, fp32) samples of fmap bf16 [h*w, C] at pts f32 [P,2] = (x,y) in [0,1] (bf16-representable); the
 * sampling grid is bf16(2 * pts - 1) as point_sample :51 forms it on the run-dtype coordinates; out bf16 [C]. */
int mpl_region_sample_mean(const void* fmap, const float* pts, int P, int h, int w, int C, void* out, void* stream);

/* GeoRegionSampler (model/rp_sampler/GeoSampler.py:162-345, behind --region_geo_sampler): the
 * non-GEMM steps. A stage's point set is a bf16 "point table" [R regions, N points, ld]: d feature columns, the two
 * normalised coordinates (row / H, col / W), zeros up to ld (ld % 8 == 0, ld >= d + 2) -- torch.cat([fea, xy], -1) of
 * :302-303 padded into a GEMM operand. Rounding points: the eager bf16 reference's. Ties: FPS = first maximum (torch.max
 * on CPU), kNN = the k smallest by (distance, index), listed in that order (the reference's topk(sorted=False) leaves
 * the choice among equal distances open).
 *   mpl_geo_point_table: point_sample :31-56,263-276. fmap bf16 [n_img, h*w, C], img_of_region int [R], pts f32
 *                        [R, P, 2] = (row / H, col / W) -> table [R, P, ld]
 *   mpl_geo_fps:         farthest_point_sample :59-80 from start[r]; xy = table + d (row pitch ld); N <= 1024
 *                        -> fps_idx int [R, S]
 *   mpl_geo_knn:         square_distance + topk :101-136 of the S anchors fps_idx among the N points -> knn_idx int [R, S, k]
 *   mpl_geo_group:       :302-308. row (r, s, j): a1[row, 0:ld] = table[r, knn] - table[r, fps] (bf16; the operand of
 *                        diff_projector), a2[row, ld:2ld] = table[r, fps] (a2 bf16 [R*S*k, 2*ld]; its first half is the
 *                        diff_projector GEMM's output: the operand of the agg_projector's 1x1 conv)
 *   mpl_geo_ln_pool:     ConvReLULN1D's LayerNorm :152-156 + AvgPool1d(k) (mode 0) / AdaptiveMaxPool1d(1) (mode 1) :317
 *                        over y bf16 [R*S, k, D] (ReLU applied by the GEMM epilogue) -> out[R*S, 0:D] (row pitch ldo);
 *                        with xy_src (= stage table + d, pitch ld_src) also the anchors' coordinates and zero padding:
 *                        out is then the next stage's point table. */
int mpl_geo_point_table(const void* fmap, const int* img_of_region, const float* pts, int R, int P, int h, int w, int C,
                        void* table, int ld, void* stream);
int mpl_geo_fps(const void* xy, long long ld, int R, int N, int S, const int* start, int* fps_idx, void* stream);
int mpl_geo_knn(const void* xy, long long ld, int R, int N, int S, int k, const int* fps_idx, int* knn_idx, void* stream);
int mpl_geo_group(const void* table, int ld, int R, int N, int S, int k, const int* fps_idx, const int* knn_idx, void* a1,
                  void* a2, void* stream);
int mpl_geo_ln_pool(const void* y, int R, int S, int k, int D, const void* weight, const void* bias, float eps, int mode,
                    const void* xy_src, long long ld_src, int N, const int* fps_idx, void* out, long long ldo, void* stream);

/* =========================================================================================================
 * Native stack runners: ONE call enqueues every kernel of a sub-model forward (no Python between kernels).
 * Weight structs hold device pointers into the caller's nn.Parameters (read in place; nothing is copied or
 * retained); arrays of per-layer structs are HOST arrays. workspace is caller-owned device scratch of at least
 * mpl_*_workspace_bytes(). All activations bf16 row-major (token-major).
 * ========================================================================================================= */
/* LLaMA decoder layer (HF 4.31 LlamaDecoderLayer as patched by MoELlamaDecoderLayer_forward,
 * model/medplib/model/language_model/medplib_moe_llama.py:110-162). wg == NULL: dense LlamaMLP from expert slot 0. */
typedef struct {
  const void* input_ln; /* bf16 [D] */
  const void* wq;       /* bf16 [D,D] */
  const void* wk;
  const void* wv;
  const void* wo;
  const void* post_ln;
  const float* wg; /* f32 [E,D]: mlp.deepspeed_moe.gate.wg.weight */
  int n_experts;
  const void* w_gate[MPL_MAX_EXPERTS]; /* bf16 [F,D]: ...experts.deepspeed_experts.{e}.gate_proj.weight */
  const void* w_up[MPL_MAX_EXPERTS];   /* bf16 [F,D] */
  const void* w_down[MPL_MAX_EXPERTS]; /* bf16 [D,F] */
} mpl_llama_layer;

typedef struct {
  int n_layers, hidden, n_heads, ffn;
  float rms_eps;
  int top_k;             /* 1 or 2 */
  float capacity_factor; /* train: moe.capacity_factor, eval: moe.eval_capacity_factor (caller picks) */
  int min_capacity;
  const mpl_llama_layer* layers; /* host array [n_layers] */
  const void* final_norm;        /* bf16 [D] */
  const void* rope_cos;          /* bf16 [rope_len, head_dim] (HF 4.31 cached tables) */
  const void* rope_sin;
  int rope_len;
} mpl_llama_model;

/* MoELlamaModel_forward (medplib_moe_llama.py:165-305) on inputs_embeds. Rows are (b,t) with t fastest. */
typedef struct {
  void* x;                    /* bf16 [B*T, D] in: inputs_embeds; out: output of the last decoder layer */
  void* out_norm;             /* bf16 [B*T, D] out: final RMSNorm (hidden_states[-1]); may be NULL */
  void* const* hidden_states; /* NULL or host array [n_layers] of bf16 [B*T,D] buffers receiving each layer's INPUT */
  int B, T, past_len;
  void* k_cache; /* bf16 [n_layers, B, H, Tmax, head_dim]; rows [past_len, past_len+T) are written */
  void* v_cache;
  int Tmax;
  const unsigned char* kv_mask; /* [B, >= past_len+T] 1 = attend (padding mask) or NULL */
  long long kv_mask_stride;
  const int* pos_dev; /* NULL, or device int overriding past_len (T == 1 decode under CUDA graphs) */
  const int* tk_dev;  /* NULL, or device int = *pos_dev + 1 */
  void* attn_scratch; /* optional zero-initialised scratch for split-K decode attention (see mpl_attn_args.scratch) */
  long long attn_scratch_bytes;
  const void* decode_plan; /* optional, from mpl_llama_decode_plan_build: T == 1 steps with B <= 8 and top-1 routing run
                              as ONE persistent cooperative kernel (llama_decode.cu) instead of ~7 launches per layer */
  const float* const* moe_noise; /* NULL or host array [n_layers] of f32 [B*T,E] (see mpl_moe_route) */
  float* gate_logits;            /* NULL or f32 [n_layers, B*T * Emax] out: layer l owns block l (indexed by TRANSFORMER layer,
                                    dense layers left untouched), rows inside it packed [B*T, E_l] (what a forward hook on wg
                                    observes); exp_counts likewise [n_layers, Emax] */
  float* l_aux;                  /* NULL or f32 [n_layers] out */
  int* exp_counts;               /* NULL or int [n_layers, E] out */
  void* workspace;
  long long workspace_bytes;
  const int* rope_pos; /* NULL, or device int [B] (decode step through decode_plan only): RoPE position of each sequence's
                          new token when it differs from the cache column past_len -- continuous batching, where the
                          sequences of a batch share the write column but have their own lengths (kv_mask hides the rest) */
} mpl_llama_io;
long long mpl_llama_workspace_bytes(const mpl_llama_model* model, int B, int T);
/* Device-resident decode plan (TMA descriptors of every weight matrix + the grid-barrier words). Build once per weight
 * set into a caller-owned device buffer of mpl_llama_decode_plan_bytes(); rebuild when a weight tensor moves. */
long long mpl_llama_decode_plan_bytes(const mpl_llama_model* model);
int mpl_llama_decode_plan_build(const mpl_llama_model* model, void* plan, void* stream);
/* Dev tool: layer >= 0 makes the decode kernel record per-phase %globaltimer stamps of that layer (-1: off); out != NULL
 * (host, 160*16 u64: [CTA][stamp]) receives the last recorded stamps after a device synchronise. */
int mpl_debug_decode_timing(int layer, unsigned long long* out);
int mpl_llama_forward(const mpl_llama_model* model, const mpl_llama_io* io, void* stream);

/* CLIP ViT encoder layer / tower (HF 4.31 CLIPEncoderLayer, CLIPVisionTransformer) as used by
 * CLIPVisionTower.forward (model/medplib/model/multimodal_encoder/clip_encoder.py:41-60). */
typedef struct {
  const void *ln1_w, *ln1_b;
  const void *wq, *bq, *wk, *bk, *wv, *bv, *wo, *bo;
  const void *ln2_w, *ln2_b;
  const void *fc1_w, *fc1_b, *fc2_w, *fc2_b;
} mpl_clip_layer;
typedef struct {
  int n_layers; /* layers to RUN = index of the selected hidden state (23 for select_layer=-2 of 24) */
  int hidden, n_heads, mlp, image_size, patch;
  int k_pad;    /* 3*patch*patch rounded up to a multiple of 8 */
  float ln_eps;
  const void* patch_w; /* bf16 [hidden, k_pad]: patch_embedding.weight.view(hidden,-1) zero-padded */
  const void* cls;     /* bf16 [hidden] */
  const void* pos;     /* bf16 [n_patches+1, hidden] */
  const void *pre_ln_w, *pre_ln_b;
  const mpl_clip_layer* layers; /* host array */
} mpl_clip_model;
long long mpl_clip_workspace_bytes(const mpl_clip_model* model, int B);
/* images bf16 [B,3,S,S] -> feats bf16 [B, n_patches, hidden] = hidden_states[n_layers][:, 1:]. */
int mpl_clip_forward(const mpl_clip_model* model, const void* images, int B, void* feats, void* workspace,
                     long long workspace_bytes, void* stream);

/* SAM-Med2D image encoder (model/segment_anything_med2d/modeling/image_encoder.py: ImageEncoderViT.forward :151-162,
 * Block.forward :214-238, Attention :280-296, Adapter_Layer :18-56). Conv weights are passed repacked for the
 * im2col GEMMs (done once by the host for these frozen weights):
 *   ad_conv  [C, 9*C]   = spatial.0.weight.permute(0,2,3,1).reshape(C, 9*C)            (co | ky,kx,ci)
 *   ad_convt [16*C, C]  = spatial.2.weight.permute(2,3,1,0).reshape(16*C, C)           (ky,kx,co | ci)
 *   neck2_w  [O, 9*O]   = neck.2.weight.permute(0,2,3,1).reshape(O, 9*O) */
typedef struct {
  const void *ln1_w, *ln1_b;
  const void *qkv_w, *qkv_b, *proj_w, *proj_b;
  const void *rel_pos_h, *rel_pos_w; /* bf16 [2*s-1, head_dim], s = window or grid size */
  int window;                        /* 0 = global attention */
  const void *ln2_w, *ln2_b;
  const void *lin1_w, *lin1_b, *lin2_w, *lin2_b;
  const void *ad_ch0, *ad_ch2; /* [C/4, C], [C, C/4]; ad_ch0 == NULL: block has no adapter */
  const void *ad_conv, *ad_convt;
  const void *ad_norm_w, *ad_norm_b;
} mpl_sam_block;
typedef struct {
  int depth, hidden, n_heads, mlp, image_size, patch, out_chans;
  const void *patch_w, *patch_b; /* bf16 [hidden, 3*patch*patch], [hidden] */
  const void* pos_embed;         /* bf16 [grid*grid, hidden] */
  const mpl_sam_block* blocks;   /* host array */
  const void *neck0_w, *neck1_w, *neck1_b, *neck2_w, *neck3_w, *neck3_b;
} mpl_sam_encoder;
long long mpl_sam_encoder_workspace_bytes(const mpl_sam_encoder* model, int B);
/* images bf16 [B,3,S,S] -> out bf16 [B, grid*grid, out_chans] (token-major = NCHW output permuted (0,2,3,1)).
 * win_part (device int [B*nw*ws*ws]) / win_unpart (device int [B*grid*grid]): mpl_gather_rows index maps of
 * window_partition / window_unpartition (image_encoder.py:299-345) for windowed blocks (-1 = zero padding). */
int mpl_sam_encoder_forward(const mpl_sam_encoder* model, const void* images, int B, const int* win_part,
                            const int* win_unpart, int n_windows, void* out, void* workspace,
                            long long workspace_bytes, void* stream);

/* SAM-Med2D mask decoder (model/segment_anything_med2d/modeling/mask_decoder.py:113-153 over
 * transformer.py:62-244) for ONE text prompt (batch 1, multimask_output=False), fed by PromptEncoder.forward
 * (prompt_encoder.py:140-187): tokens = [iou_token; mask_tokens; text_embed], src = image_embedding + no_mask_embed. */
typedef struct {
  const void *q_w, *q_b, *k_w, *k_b, *v_w, *v_b, *o_w, *o_b;
} mpl_sam_attn;
typedef struct {
  mpl_sam_attn self_attn, t2i, i2t;
  const void *n1_w, *n1_b, *n2_w, *n2_b, *n3_w, *n3_b, *n4_w, *n4_b;
  const void *lin1_w, *lin1_b, *lin2_w, *lin2_b;
} mpl_sam_twoway_layer;
typedef struct {
  int dim, n_heads, mlp, depth, n_mask_tokens, grid;
  const void* iou_token;   /* bf16 [dim] */
  const void* mask_tokens; /* bf16 [n_mask_tokens, dim] */
  const void* no_mask;     /* bf16 [dim]: prompt_encoder.no_mask_embed.weight */
  const float* dense_pe;   /* f32 [grid*grid, dim]: get_dense_pe() token-major (input independent, cached by host) */
  const mpl_sam_twoway_layer* layers; /* host array [depth] */
  mpl_sam_attn final_attn;
  const void *nf_w, *nf_b;
  const void *up0_w, *up0_b; /* repacked convT k2s2: bf16 [4*dim/4, dim] rows (ky,kx,co), bias tiled [4*dim/4] */
  const void *up_ln_w, *up_ln_b;
  const void *up1_w, *up1_b; /* bf16 [4*dim/8, dim/4], bias tiled */
  const int* shuffle_idx;    /* device int [16*grid*grid]: raster pixel -> row of the twice-upscaled buffer */
  const void *hyper_w[3], *hyper_b[3]; /* output_hypernetworks_mlps[0] (mask 0 is the one returned) */
  const void *iou_w[3], *iou_b[3];     /* iou_prediction_head */
} mpl_sam_mask_decoder;
long long mpl_sam_mask_decoder_workspace_bytes(const mpl_sam_mask_decoder* model);
/* image_embedding bf16 [grid*grid, dim] (token-major), text_embed bf16 [dim] -> low_res_mask bf16 [4*grid, 4*grid],
 * iou bf16 [n_mask_tokens] (caller takes [0]). */
int mpl_sam_mask_decoder_forward(const mpl_sam_mask_decoder* model, const void* image_embedding,
                                 const void* text_embed, void* low_res_mask, void* iou, void* workspace,
                                 long long workspace_bytes, void* stream);

/* =========================================================================================================
 * Train step: backward of the LLaMA-MoE stack with LoRA adapters, fused cross-entropy, mask
 * losses, AdamW. Replaces what torch autograd + peft 0.10 + DeepSpeed run for train_ds_medplib.py:599-625
 * (model_engine.backward / step). Every dX = dY·W contraction reuses mpl_gemm_bf16 against transposed weight copies
 * (made once with mpl_transpose_bf16; the frozen 7B weights fit twice in 180 GB).
 * ========================================================================================================= */
/* out[c, r] = in[r, c] (bf16; any alignment). */
int mpl_transpose_bf16(const void* in, long long ld_in, void* out, long long ld_out, int rows, int cols, void* stream);

/* peft LoRA Linear (y = x W^T + s * (x A^T) B^T, A [r,K], B [N,r], r <= 16):
 *   mpl_lora_down   u[m,j] = scale * sum_k x[m,k] A[j,k]            (u bf16 — the forward's rounding — or f32 [M,r])
 *   mpl_lora_up_add y[m,n] = bf16(y + bf16(scale * bf16(sum_j u[m,j] Bm[n*sn + j*sr])))   in place
 *   mpl_rank_wgrad  out[n*sn + j*sr] += scale * sum_m X[m,n] U[m,j]  (f32, atomic)  — dB = s dY^T u, dA = du^T x, dwg
 * backward: du = s * dY B (lora_down with B^T), dX += du A (lora_up_add with Bm = A, sn = 1, sr = K). */
int mpl_lora_down(const void* x, long long ldx, const void* A, long long lda, void* u, int u_is_f32, int M, int K, int r,
                  float scale, void* stream);
int mpl_lora_up_add(void* y, long long ldy, const void* u, int u_is_f32, const void* Bm, long long bm_stride_n,
                    long long bm_stride_r, float scale, int M, int N, int r, void* stream);
/* The adapters as an extension k-block of the base GEMM (mpl_gemm_args.ext_a / ext_b):
 *   mpl_lora_down_ext = mpl_lora_down + a second copy of u in columns [pad_col, pad_col + r) of u_pad bf16 [M, 64]
 *   mpl_lora_pack     fills the weight-side operands of EVERY adapter in one launch: items = device array of
 *                     {const bf16* src; bf16* dst; long long sn, sr; int N, r, col; float scale; long long dn, dj}
 *                     (64 bytes): dst[n * dn + (col + j) * dj] = bf16(scale * src[n * sn + j * sr]) -- dn, dj = 64, 1 for an
 *                     extension operand bf16 [N, 64] (zero elsewhere), 1, N for a plain transposed copy [r, N]. */
int mpl_lora_down_ext(const void* x, long long ldx, const void* A, long long lda, void* u, int u_is_f32, int M, int K,
                      int r, float scale, void* u_pad, int pad_col, void* stream);
int mpl_lora_pack(const void* items, int n_items, void* stream);
int mpl_rank_wgrad(const void* X, long long ldx, const void* U, int u_is_f32, float* out, long long out_stride_n,
                   long long out_stride_r, float scale, int M, int N, int r, void* stream);

/* LlamaRMSNorm backward: dx = bf16(rstd * (g - xhat * mean(g * xhat)) + add), g = dy * w, xhat = x * rstd;
 * add (bf16 or NULL) = the gradient arriving on the residual branch; dweight (f32 [D] or NULL) += sum_rows dy * xhat. */
int mpl_rmsnorm_bwd(const void* x, long long ldx, const void* weight, const void* dy, long long lddy, const void* add,
                    long long ldadd, void* dx, long long lddx, float* dweight, int rows, int D, float eps, void* stream);
/* LlamaMLP gate: h = bf16(silu(g)) * u and its backward (dg may alias g, du may alias u). n elements, n % 8 == 0. */
int mpl_silu_mul(const void* g, const void* u, void* h, long long n, void* stream);
int mpl_silu_mul_bwd(const void* g, const void* u, const void* dh, void* dg, void* du, long long n, void* stream);

/* Backward of mpl_attention (self-attention over T tokens, no cache; head_dim 64 or 128) given the forward's lse:
 * dq (f32 [B,T,H,d], zero-initialised by the caller, accumulated with atomics), dk / dv (bf16, own strides).
 * q and k are the ROTATED tensors the forward saw; mpl_rope_bwd then rotates dq / dk back. */
typedef struct {
  const void *q, *k, *v, *o, *d_o;
  long long q_stride[3], k_stride[3], v_stride[3], o_stride[3]; /* o strides address both o and d_o */
  const float* lse; /* f32 [B*H, T] from mpl_attn_args.lse */
  float* delta;     /* f32 [B*H, T] scratch: rowsum(dO * O) */
  float* dq_f32;
  void *dk, *dv;
  long long dk_stride[3], dv_stride[3];
  int B, H, T, head_dim;
  float scale;
  int causal;
  const unsigned char* kv_mask;
  long long kv_mask_stride;
} mpl_attn_bwd_args;
int mpl_attention_bwd(const mpl_attn_bwd_args* args, void* stream);
/* dq = bf16(R(-pos) dq_f32) into rows of leading dimension ld; dk rotated in place (see mpl_rope_kv). */
int mpl_rope_bwd(const float* dq_f32, void* dq, void* dk, long long ld, const void* cos_t, const void* sin_t, int B, int T,
                 int H, int head_dim, int pos0, void* stream);

/* MoE backward (DeepSpeed top-1 / top-2 gating autograd):
 *   mpl_moe_combine_bwd: dy[slot[s,j]] = gate[s,j] * dout[s]; dgate[s,j] = <dout[s], y[slot[s,j]]> (0 when dropped)
 *   (the dispatch's backward is mpl_moe_combine with unit gates)
 *   mpl_moe_router_bwd: dlogits = softmax-backward of (dgate on the chosen expert(s) + aux_scale * dl_aux/dgates);
 *   k = 2 first takes dgate back through top2gating's renormalisation g_i / max(g_1 + g_2, eps) over the kept choices
 *   (expert / slot / dgate are [S,k]); dh[s,:] += dlogits[s,:] wg. dwg = dlogits^T h via mpl_rank_wgrad. */
int mpl_moe_combine_bwd(const void* dout, long long ldd, const void* y, const int* slot, const float* gate, void* dy,
                        float* dgate, int S, int k, int D, void* stream);
int mpl_moe_router_bwd(const float* gates, const int* expert, const int* slot, const float* dgate, const int* exp_counts,
                       float aux_scale, const float* wg, float* dlogits, void* dh, long long ldh, int S, int D, int E,
                       int k, void* stream);

/* Zero rows [e*C + kept[e], (e+1)*C) of an expert buffer bf16 [groups*C, width] (row pitch ld): everything the per-expert
 * kernels leave unwritten, instead of zero-filling the whole buffer first. */
int mpl_zero_tail_rows(void* buf, long long ld, int groups, int C, int width, const int* kept, void* stream);

/* Shifted cross-entropy of medplib_moe_llama.py:399-421 on fp32 logits: labels i64 [rows] (already shifted; < 0 ignored).
 * fwd: lse[r], acc[0] += sum of row losses, acc[1] += valid rows (acc zeroed by the caller; loss = acc[0]/acc[1]).
 * bwd: dlogits bf16 [rows, ldd] = (softmax - onehot) * (*grad_out) / acc[1], columns [V, ldd) zero. */
int mpl_ce_fwd(const float* logits, long long ld, const long long* labels, int rows, int V, float* lse, float* acc,
               void* stream);
int mpl_ce_bwd(const float* logits, long long ld, const long long* labels, int rows, int V, const float* lse,
               const float* acc, const float* grad_out, void* dlogits, long long ldd, void* stream);
/* Adjoint of mpl_gather_rows (embedding lookup + multimodal splice): f32 atomic accumulation into the embedding-table
 * gradient (idx >= 0) and the feature-row gradient (idx <= -2); either target may be NULL. */
int mpl_scatter_add_rows(const void* dx, long long ldx, const int* idx, float* dtable, long long ld_table, float* dfeats,
                         long long ld_feats, int rows, int D, void* stream);

/* Optimizer (train_ds_medplib.py:398-411: AdamW betas (0.9, 0.95), weight decay 0, gradient clipping 1.0):
 * out[0] += sum g^2; AdamW on fp32 master weights + moments writing the bf16 / f32 parameter in place, with the clip
 * coefficient min(1, max_norm / (sqrt(*sumsq) * grad_scale + 1e-6)) applied on the fly (sumsq NULL: no clipping). */
int mpl_sumsq_f32(const float* g, long long n, float* out, void* stream);
int mpl_adamw(float* master, float* m, float* v, const float* grad, void* param, int param_is_bf16, long long n, float lr,
              float beta1, float beta2, float eps, float weight_decay, int step, const float* sumsq, float max_norm,
              float grad_scale, void* stream);
/* mpl_adamw over the whole flat gradient arena in ONE launch. chunks: device int64 [n_chunks, 4] = {arena offset, length
 * (<= 4096), parameter base address, is_bf16 | (element offset within the parameter << 1)}. */
int mpl_adamw_multi(float* master, float* m, float* v, const float* grad, const long long* chunks, int n_chunks, float lr,
                    float beta1, float beta2, float eps, float weight_decay, int step, const float* sumsq, float max_norm,
                    float grad_scale, void* stream);

/* The four mask losses of model/MedPLIB.py:26-124 for ONE mask in one pass: pred bf16 [n] logits, gt f32 [n] in {0,1},
 * pred_iou bf16 scalar -> out4 = {sigmoid_ce_loss, dice_loss, MaskIoULoss, FocalLoss}; sums6 (optional) = the six
 * reductions (sum bce, sum p, sum t, sum p t, focal_pos, focal_neg). */
int mpl_mask_losses(const void* pred, const float* gt, const void* pred_iou, long long n, float* out4, float* sums6,
                    void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Grounding-head backward (train step with seg_flag): what torch autograd runs for model/MedPLIB.py:456-559 over
 * text_hidden_fcs (:153-164), model/segment_anything_med2d/modeling/mask_decoder.py:71-153, transformer.py:16-244,
 * postprocess_masks (:682-701) and the four mask losses (:26-124). Latency-bound small kernels (< 1 GFLOP per mask).
 * --------------------------------------------------------------------------------------------------------- */
/* C[m*ldc + n] (+)= sum_k A[m*a_stride_m + k*a_stride_k] * B[k*b_stride_k + n*b_stride_n]; A, B bf16 or f32; C bf16
 * (overwritten) or f32 (accumulate = 1 adds). Strides express every transposition: dX = dY W, dW += dY^T X,
 * outer products. */
int mpl_gemm_small(const void* A, int a_is_f32, long long a_stride_m, long long a_stride_k, const void* B, int b_is_f32,
                   long long b_stride_k, long long b_stride_n, void* C, int c_is_f32, long long ldc, int accumulate,
                   int M, int N, int K, void* stream);
/* out[n] += sum_m X[m*ld + n] (bias gradients; M = 1: accumulate a bf16 / f32 vector into an fp32 gradient). */
int mpl_col_sum(const void* X, int x_is_f32, long long ld, float* out, int M, int N, void* stream);
/* nn.LayerNorm / LayerNorm2d (rows = NHWC pixels) backward; dweight / dbias f32 [D] accumulated (NULL: skipped). */
int mpl_layernorm_bwd(const void* x, long long ldx, const void* weight, const void* dy, long long lddy, void* dx,
                      long long lddx, float* dweight, float* dbias, int rows, int D, float eps, void* stream);
/* y = act(x) and dx = act'(x) dy for MPL_ACT_GELU (erf form; x = the input) / MPL_ACT_RELU (x = input or output). */
int mpl_act_fwd(const void* x, void* y, long long n, int act, void* stream);
int mpl_act_bwd(const void* x, const void* dy, void* dx, long long n, int act, void* stream);
/* Vision-side adapters of the ICL recipe (scripts/train_medplib_icl.sh: --sft_modules mm_token_compressor,mask_encoder):
 *   mpl_token_pool_bwd: adjoint of AdaptiveAvgPool1d over tokens (medplib_arch.py:73,104): dy bf16 [n,t_out,D] -> dx [n,t_in,D]
 *   mpl_col2im_nhwc:    adjoint of mpl_im2col_nhwc (dgrad of the MaskTokenEncoder's 3x3 stride-2 convs, :84-93):
 *                       dcols bf16 [B*Ho*Wo, kh*kw*C] -> dx bf16 [B,H,W,C] */
int mpl_token_pool(const void* x, void* y, int n, int t_in, int t_out, int D, void* stream); /* the pool itself */
int mpl_token_pool_bwd(const void* dy, void* dx, int n, int t_in, int t_out, int D, void* stream);
int mpl_col2im_nhwc(const void* dcols, void* dx, int B, int H, int W, int C, int kh, int kw, int stride, int pad,
                    void* stream);
/* Backward of softmax(scale q k^T) v for the mask decoder's attentions (transformer.py:185-244): q [batch*Tq, ld],
 * k / v [batch*Tk, ld] (batches stacked along rows), head h = columns [h*head_dim, (h+1)*head_dim), head_dim <= 32;
 * one CTA per (head, batch). */
int mpl_attn_small_bwd(const void* q, long long ldq, const void* k, long long ldk, const void* v, long long ldv,
                       const void* d_o, long long ldo, void* dq, long long lddq, void* dk, long long lddk, void* dv,
                       long long lddv, int batch, int Tq, int Tk, int H, int head_dim, float scale, void* stream);
/* peft lora_dropout (train_ds_medplib.py:296-300): out = (accumulate ? out : 0) + x * mask * scale over n bf16 elements
 * (mask: one byte per element, scale = 1 / (1 - p)); forward on the adapter's input, backward on its input gradient. */
int mpl_mask_scale_bf16(const void* x, const unsigned char* mask, float scale, void* out, int accumulate, long long n,
                        void* stream);
/* Adjoint of mpl_bilinear_resize: dy [N, Hout, Wout] (bf16 or f32) -> dx bf16 [N, Hin, Win] (strides given). */
int mpl_bilinear_resize_bwd(const void* dy, int dy_is_f32, int Hout, int Wout, void* dx, long long dx_stride_n,
                            long long dx_stride_y, int Hin, int Win, int N, void* stream);
/* Backward of mpl_mask_losses: dloss4 = d total / d {bce, dice, iou, focal} (device f32[4]), sums6 from the forward
 * -> dpred bf16 [n], dpred_iou f32[1] (optional). */
int mpl_mask_losses_bwd(const void* pred, const float* gt, const void* pred_iou, const float* sums6, const float* dloss4,
                        long long n, void* dpred, float* dpred_iou, void* stream);

/* =========================================================================================================
 * Image input pipeline: decoded u8 image -> model-ready tensor, one launch for a ragged batch.
 * Replaces, per image, datasets/LazySupervisedDataset.py:539-553 (and :516-517 for region masks):
 *   ResizeLongestSide.apply_image (model/segment_anything_med2d/utils/transforms.py:26-32) = PIL Image.resize(BILINEAR):
 *     antialiased two-pass convolution on 8-bit channels, 22-bit fixed-point coefficients, u8 rounding between the
 *     horizontal and the vertical pass (Pillow libImaging/Resample.c) — reproduced bit for bit;
 *   preprocess + pad_tensor_channelwise (:446-502) and CLIPImageProcessor rescale/normalize: every output value is a
 *     function of the resized u8 level, so the host passes a 256-entry fp32 table per channel (lut) and the pad value.
 * Each job turns src u8 [H,W,C] (interleaved, C = 1 or 3) into dst [C, out_size, out_size]: the new_h x new_w resized
 * image placed at (pad_top, pad_left), pad_value[c] elsewhere.  coef / bound are PIL's tables for one axis, built by
 * the host (medplib_b200/preprocess.py:pil_coeffs): bound int32 [n_out, 2] = (first tap, tap count); coef_y int32
 * [new_h, ks_y]; coef_x TAP-MAJOR int32 [ks_x, new_w] with ks_x a multiple of 4, zero-filled past each pixel's tap
 * count (the horizontal pass consumes taps four at a time from aligned 32-bit source words).  src must be 16-byte
 * aligned (MPL_ERR_ALIGN) and up to 15 bytes past the image's last byte, inside the same allocation, may be read.  jobs_host and jobs_dev hold the same n_jobs structs (the pointers inside
 * are device pointers); the host copy sizes the grid and the shared memory.  MPL_ERR_UNSUPPORTED when two staged source
 * rows plus one output row's taps do not fit 200 KB of shared memory (rows beyond ~30k pixels, downscales beyond ~100x). */
typedef struct {
  const void* src;
  long long src_stride; /* bytes between source rows */
  int H, W, C;
  int new_h, new_w;
  int out_size;
  int pad_top, pad_left;
  const int* coef_x;
  const int* bound_x;
  const int* coef_y;
  const int* bound_y;
  int ks_x, ks_y;
  const float* lut; /* f32 [C,256], or NULL: the u8 level itself */
  float pad_value[3];
  int out_dtype; /* MPL_DT_BF16 / MPL_DT_F32 / MPL_DT_U8 */
  void* dst;
} mpl_preprocess_job;
int mpl_preprocess_images(const mpl_preprocess_job* jobs_host, const mpl_preprocess_job* jobs_dev, int n_jobs,
                          void* stream);
/* Rows of the source a band of R output rows can touch (upper bound used to size shared memory; integer-only so the
 * host and the kernel agree). Exposed for the host-side tests. */
int mpl_preprocess_band_rows(int in_size, int out_size, int R);

#ifdef __cplusplus
}
#endif
#endif /* MEDPLIB_B200_H */
