// Native runner of the LLaMA-MoE decoder stack: one C-ABI call enqueues every kernel of a prefill / decode / train
// forward over all layers (no Python between kernels, no host synchronisation, no exp_counts.to('cpu')).
//
// Replaces MoELlamaModel_forward + MoELlamaDecoderLayer_forward
// (model/medplib/model/language_model/medplib_moe_llama.py:110-305) over HF-4.31 LlamaRMSNorm / LlamaAttention /
// LlamaMLP and deepspeed.moe.layer.MoE. Per layer:
//   rmsnorm -> fused q,k,v GEMM (tcgen05, or the streaming kernel when B*T <= 16) -> RoPE + KV-cache append ->
//   flash / decode attention over the cache -> o_proj GEMM (+residual epilogue) -> rmsnorm ->
//   router + scan -> dispatch -> per expert [gate|up GEMM with SiLU*mul epilogue -> down GEMM], sized on the device by
//   kept[e] -> combine (+residual)         (dense layers: gate|up -> down with residual epilogue)
// and a final rmsnorm. x is updated in place.
#include <cmath>
#include <cstring>

#include "internal.h"

namespace mpl {

static inline long long align_up(long long v) { return (v + 255) & ~255LL; }

struct LlamaWs {
  char* h;
  char* qkv;
  char* attn;
  char* xperm;
  char* h1;
  char* y;
  float* logits;
  float* gates;
  int* expert;
  float* gate;
  int* slot;
  int* kept;
  int* exp_counts;
  float* l_aux;
  int* tok_of_slot;
  float* gate_of_slot;
  long long total;
};

static int moe_capacity(const mpl_llama_model& m, int S, int E) {
  const double cf = static_cast<double>(m.capacity_factor) * (m.top_k == 2 ? 2.0 : 1.0);
  int c = static_cast<int>(std::ceil((static_cast<double>(S) / E) * cf));
  if (c < m.min_capacity) c = m.min_capacity;
  if (c < 1) c = 1;
  return c;
}

static int max_experts(const mpl_llama_model& m) {
  int e = 1;
  for (int i = 0; i < m.n_layers; ++i)
    if (m.layers[i].wg != nullptr && m.layers[i].n_experts > e) e = m.layers[i].n_experts;
  return e;
}

static LlamaWs carve(const mpl_llama_model& m, int B, int T, char* base) {
  const long long S = static_cast<long long>(B) * T, D = m.hidden, F = m.ffn;
  const int E = max_experts(m);
  long long erows = S;
  for (int i = 0; i < m.n_layers; ++i)
    if (m.layers[i].wg != nullptr) {
      const long long r = static_cast<long long>(m.layers[i].n_experts) * moe_capacity(m, S, m.layers[i].n_experts);
      if (r > erows) erows = r;
    }
  LlamaWs w;
  long long off = 0;
  auto take = [&](long long bytes) {
    char* p = base ? base + off : nullptr;
    off += align_up(bytes);
    return p;
  };
  w.h = take(S * D * 2);
  w.qkv = take(S * 3 * D * 2);
  w.attn = take(S * D * 2);
  w.xperm = take(erows * D * 2);
  w.h1 = take(erows * F * 2);
  w.y = take(erows * D * 2);
  w.logits = reinterpret_cast<float*>(take(S * E * 4));
  w.gates = reinterpret_cast<float*>(take(S * E * 4));
  w.expert = reinterpret_cast<int*>(take(S * 2 * 4));
  w.gate = reinterpret_cast<float*>(take(S * 2 * 4));
  w.slot = reinterpret_cast<int*>(take(S * 2 * 4));
  w.kept = reinterpret_cast<int*>(take(MPL_MAX_EXPERTS * 4));
  w.exp_counts = reinterpret_cast<int*>(take(MPL_MAX_EXPERTS * 4));
  w.l_aux = reinterpret_cast<float*>(take(256));
  w.tok_of_slot = reinterpret_cast<int*>(take(erows * 4));
  w.gate_of_slot = reinterpret_cast<float*>(take(erows * 4));
  w.total = off;
  return w;
}

static mpl_gemm_args gemm_base(const void* A, long long lda, int M, int N, int K) {
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
  g.act = MPL_ACT_NONE;
  g.out_dtype = MPL_DT_BF16;
  return g;
}

#define MPL_TRY(expr)            \
  do {                           \
    const int rc__ = (expr);     \
    if (rc__ != MPL_OK) return rc__; \
  } while (0)

int llama_forward(const mpl_llama_model& m, const mpl_llama_io& io, cudaStream_t st) {
  if (m.layers == nullptr || io.x == nullptr || io.k_cache == nullptr || io.v_cache == nullptr ||
      io.workspace == nullptr || m.rope_cos == nullptr || m.rope_sin == nullptr)
    return MPL_ERR_ARG;
  const int B = io.B, T = io.T, D = m.hidden, H = m.n_heads, F = m.ffn;
  if (B <= 0 || T <= 0) return MPL_OK;
  const int hd = D / H;
  const int S = B * T;
  const int past = io.past_len;
  if (io.pos_dev == nullptr && (past + T > io.Tmax || past + T > m.rope_len)) return MPL_ERR_ARG;
  const LlamaWs w = carve(m, B, T, static_cast<char*>(io.workspace));
  if (w.total > io.workspace_bytes) return MPL_ERR_ARG;
  void* st_ = static_cast<void*>(st);
  if (llama_decode_supported(m, io)) {
    // decode step: one persistent kernel for all layers (h1 holds E*B rows: erows >= E*C >= the rows it needs)
    int cap[MPL_MAX_EXPERTS + 1];
    for (int e = 0; e <= MPL_MAX_EXPERTS; ++e) cap[e] = e > 0 ? moe_capacity(m, S, e) : 0;
    const int emax = max_experts(m);
    if (static_cast<long long>(emax) * B * F * 2 <= (w.y - w.h1))
      return llama_decode_step(m, io, w.qkv, w.attn, w.h1, cap, emax, st);
  }
  if (io.rope_pos != nullptr) return MPL_ERR_UNSUPPORTED;  // per-sequence positions exist in the one-kernel decode step only
  const long long cache_layer = static_cast<long long>(B) * H * io.Tmax * hd;  // elements per layer
  const int emax_all = max_experts(m);  // router outputs: one [S * Emax] block per transformer layer, rows packed by E_l

  bool h_ready = false;  // w.h already holds RMSNorm(x) with this layer's input_ln (fused into the previous MoE combine)
  bool out_norm_done = false;
  for (int l = 0; l < m.n_layers; ++l) {
    const mpl_llama_layer& L = m.layers[l];
    if (io.hidden_states != nullptr && io.hidden_states[l] != nullptr)
      if (cudaMemcpyAsync(io.hidden_states[l], io.x, static_cast<size_t>(S) * D * 2, cudaMemcpyDeviceToDevice, st) !=
          cudaSuccess)
        return MPL_ERR_CUDA;
    // ---- attention block (decode: the RMSNorm runs as the prologue of the streaming q,k,v GEMM)
    const bool small = S <= 16;
    if (!small && !h_ready) MPL_TRY(mpl_rmsnorm(io.x, D, L.input_ln, w.h, D, S, D, m.rms_eps, st_));
    h_ready = false;
    {
      mpl_gemm_args g = gemm_base(small ? io.x : w.h, D, S, D, D);
      if (small) {
        g.ln_weight = L.input_ln;
        g.ln_eps = m.rms_eps;
      }
      g.nb = 3;
      g.B[0] = L.wq;
      g.B[1] = L.wk;
      g.B[2] = L.wv;
      g.C[0] = w.qkv;
      g.C[1] = w.qkv + static_cast<long long>(D) * 2;
      g.C[2] = w.qkv + static_cast<long long>(D) * 4;
      g.ldc = 3LL * D;
      MPL_TRY(linear_bf16(g, st));
    }
    char* kc = static_cast<char*>(io.k_cache) + l * cache_layer * 2;
    char* vc = static_cast<char*>(io.v_cache) + l * cache_layer * 2;
    MPL_TRY(mpl_rope_kv(w.qkv, w.qkv + static_cast<long long>(D) * 2, w.qkv + static_cast<long long>(D) * 4, 3LL * D,
                        m.rope_cos, m.rope_sin, kc, vc, B, T, H, hd, io.Tmax, past, io.pos_dev, st_));
    {
      mpl_attn_args a;
      memset(&a, 0, sizeof(a));
      a.q = w.qkv;
      a.k = kc;
      a.v = vc;
      a.o = w.attn;
      a.q_stride[0] = static_cast<long long>(T) * 3 * D;
      a.q_stride[1] = 3LL * D;
      a.q_stride[2] = hd;
      a.k_stride[0] = a.v_stride[0] = static_cast<long long>(H) * io.Tmax * hd;
      a.k_stride[1] = a.v_stride[1] = hd;
      a.k_stride[2] = a.v_stride[2] = static_cast<long long>(io.Tmax) * hd;
      a.o_stride[0] = static_cast<long long>(T) * D;
      a.o_stride[1] = D;
      a.o_stride[2] = hd;
      a.B = B;
      a.H = H;
      a.Tq = T;
      a.Tk = past + T;
      a.head_dim = hd;
      a.scale = 1.0f / sqrtf(static_cast<float>(hd));
      a.causal = T > 1;
      a.kv_mask = io.kv_mask;
      a.kv_mask_stride = io.kv_mask_stride;
      a.tk_dev = io.tk_dev;
      a.scratch = io.attn_scratch;
      a.scratch_bytes = io.attn_scratch_bytes;
      MPL_TRY(mpl_attention(&a, st_));
    }
    {
      mpl_gemm_args g = gemm_base(w.attn, D, S, D, D);
      g.B[0] = L.wo;
      g.C[0] = io.x;
      g.residual = io.x;
      g.ldr = D;
      MPL_TRY(linear_bf16(g, st));
    }
    // ---- FFN block
    if (L.wg == nullptr) {
      if (!small) MPL_TRY(mpl_rmsnorm(io.x, D, L.post_ln, w.h, D, S, D, m.rms_eps, st_));
      mpl_gemm_args g = gemm_base(small ? io.x : w.h, D, S, F, D);
      if (small) {
        g.ln_weight = L.post_ln;
        g.ln_eps = m.rms_eps;
      }
      g.B[0] = L.w_gate[0];
      g.B2 = L.w_up[0];
      g.C[0] = w.h1;
      MPL_TRY(linear_bf16(g, st));
      mpl_gemm_args d = gemm_base(w.h1, F, S, D, F);
      d.B[0] = L.w_down[0];
      d.C[0] = io.x;
      d.residual = io.x;
      d.ldr = D;
      MPL_TRY(linear_bf16(d, st));
      continue;
    }
    const int E = L.n_experts;
    const int C = moe_capacity(m, S, E);
    mpl_moe_route_args r;
    memset(&r, 0, sizeof(r));
    r.h = w.h;
    r.ldh = D;
    r.wg = L.wg;
    r.noise = io.moe_noise ? io.moe_noise[l] : nullptr;
    r.S = S;
    r.D = D;
    r.E = E;
    r.k = m.top_k;
    r.capacity = C;
    r.logits = io.gate_logits ? io.gate_logits + static_cast<long long>(l) * S * emax_all : w.logits;  // layer block = S * Emax
    r.gates = w.gates;
    r.expert = w.expert;
    r.gate = w.gate;
    r.slot = w.slot;
    r.kept = w.kept;
    r.exp_counts = io.exp_counts ? io.exp_counts + static_cast<long long>(l) * emax_all : w.exp_counts;
    r.l_aux = io.l_aux ? io.l_aux + l : w.l_aux;
    const bool fused_front = small && !(m.top_k == 1 && r.noise != nullptr);
    const bool gather = fused_front && C <= 16;  // streaming expert GEMMs read their rows through tok_of_slot
    if (fused_front) {
      // one launch: RMSNorm + router + slots (+ slot->token map for the fused dispatch / combine)
      MPL_TRY(moe_route_small(r, io.x, D, L.post_ln, m.rms_eps, w.h, D, gather ? nullptr : w.xperm, w.tok_of_slot,
                              w.gate_of_slot, st));
    } else {
      const int frc = moe_norm_route(r, io.x, D, L.post_ln, m.rms_eps, st);  // RMSNorm fused into the router
      if (frc == MPL_ERR_UNSUPPORTED) {
        MPL_TRY(mpl_rmsnorm(io.x, D, L.post_ln, w.h, D, S, D, m.rms_eps, st_));
        MPL_TRY(moe_route(r, st));
      } else if (frc != MPL_OK) {
        return frc;
      }
      MPL_TRY(moe_dispatch(w.h, D, w.slot, w.xperm, S, m.top_k, D, st));
    }
    const bool fused_combine = fused_front && m.top_k == 1 && C <= 16;
    {
      mpl_grouped_gemm_args g;
      memset(&g, 0, sizeof(g));
      g.A = gather ? w.h : w.xperm;
      g.lda = D;
      g.a_group_stride = static_cast<long long>(C) * D;
      if (gather) {
        g.a_row_map = w.tok_of_slot;
        g.map_group_stride = C;
      }
      for (int e = 0; e < E; ++e) {
        g.B[e] = L.w_gate[e];
        g.B2[e] = L.w_up[e];
      }
      g.ldb = D;
      g.C = w.h1;
      g.ldc = F;
      g.c_group_stride = static_cast<long long>(C) * F;
      g.m_dev = w.kept;
      g.groups = E;
      g.M = C;
      g.N = F;
      g.K = D;
      g.out_dtype = MPL_DT_BF16;
      g.m_total_hint = S * m.top_k;
      MPL_TRY(mpl_grouped_gemm_bf16(&g, st_));
      mpl_grouped_gemm_args d;
      memset(&d, 0, sizeof(d));
      d.A = w.h1;
      d.lda = F;
      d.a_group_stride = static_cast<long long>(C) * F;
      for (int e = 0; e < E; ++e) d.B[e] = L.w_down[e];
      d.ldb = F;
      d.m_dev = w.kept;
      d.m_dev_stable = 1;  // kept[] was written two launches back (router), before the gate/up GEMM started
      d.groups = E;
      d.M = C;
      d.N = D;
      d.K = F;
      d.out_dtype = MPL_DT_BF16;
      d.m_total_hint = S * m.top_k;
      if (fused_combine) {
        d.C = io.x;
        d.ldc = D;
        d.residual = io.x;
        d.ldr = D;
        d.row_map = w.tok_of_slot;
        d.row_gate = w.gate_of_slot;
        d.map_group_stride = C;
      } else {
        d.C = w.y;
        d.ldc = D;
        d.c_group_stride = static_cast<long long>(C) * D;
      }
      MPL_TRY(mpl_grouped_gemm_bf16(&d, st_));
    }
    if (fused_combine) continue;
    // combine (+ residual) and, while the row is in registers, the RMSNorm its next consumer needs: the next layer's
    // input_layernorm (into w.h) or the final norm (into out_norm)
    const void* next_ln = nullptr;
    void* next_h = nullptr;
    if (!small && D <= 4096) {
      if (l + 1 < m.n_layers) {
        next_ln = m.layers[l + 1].input_ln;
        next_h = w.h;
      } else if (io.out_norm != nullptr) {
        next_ln = m.final_norm;
        next_h = io.out_norm;
      }
    }
    MPL_TRY(moe_combine(w.y, w.slot, w.gate, io.x, D, io.x, D, S, m.top_k, D, st, next_ln, m.rms_eps, next_h, D));
    if (next_ln != nullptr) {
      if (l + 1 < m.n_layers)
        h_ready = true;
      else
        out_norm_done = true;
    }
  }
  if (io.out_norm != nullptr && !out_norm_done)
    MPL_TRY(mpl_rmsnorm(io.x, D, m.final_norm, io.out_norm, D, S, D, m.rms_eps, st_));
  return MPL_OK;
}

}  // namespace mpl

extern "C" long long mpl_llama_workspace_bytes(const mpl_llama_model* m, int B, int T) {
  if (m == nullptr || m->layers == nullptr || B <= 0 || T <= 0) return 0;
  return mpl::carve(*m, B, T, nullptr).total;
}

extern "C" long long mpl_llama_decode_plan_bytes(const mpl_llama_model* m) {
  if (m == nullptr || m->layers == nullptr) return 0;
  return mpl::llama_decode_plan_bytes(*m);
}
extern "C" int mpl_llama_decode_plan_build(const mpl_llama_model* m, void* plan, void* stream) {
  if (m == nullptr) return MPL_ERR_ARG;
  return mpl::llama_decode_plan_build(*m, plan, static_cast<cudaStream_t>(stream));
}

extern "C" int mpl_llama_forward(const mpl_llama_model* m, const mpl_llama_io* io, void* stream) {
  if (m == nullptr || io == nullptr) return MPL_ERR_ARG;
  return mpl::llama_forward(*m, *io, static_cast<cudaStream_t>(stream));
}
