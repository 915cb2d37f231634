// K3 — streaming GEMM for decode-time and tiny-M linears:  C[M,N] = epilogue(A[M,K] · W[N,K]^T),  M <= 16.
//
// HBM-bound: every weight byte is read exactly once, and the point of the design is that the weight stream never stops.
//  * PERSISTENT: one CTA per SM walks output tiles (16 weight rows x all of K; for SiLU(gate)*up: 8 gate rows + 8 up
//    rows of the same features, so the pairing is thread-local in the accumulator fragment). Tiles of all the weight
//    matrices of a launch (q,k,v; or the experts of an MoE layer, skipping experts without tokens) form one list.
//  * ONE producer thread fills a 6-stage shared-memory ring with TMA: the weight matrix is described as a 3-D tensor
//    {64 k, N rows, K/64 k-blocks}, so ONE cp.async.bulk.tensor instruction moves a 16-row x 512-k box (16 KB, 128-B
//    swizzle, completion on an mbarrier): 96 KB in flight per SM, independent of registers. (1-KB per-row bulk copies
//    capped the SM at ~19 GB/s; one big box per stage lets the TMA unit generate the lines itself.)
//  * PROGRAMMATIC DEPENDENT LAUNCH: weights never depend on the previous kernel, so the producer starts streaming as
//    soon as the CTA is resident, while the previous kernel in the stream is still finishing; only the consumers
//    execute griddepcontrol.wait before they touch activations. Every launch carries the programmatic-serialization
//    attribute and triggers its own dependents right after its wait, so two streaming kernels overlap head to tail.
//  * Eight consumer warps split each 512-element K chunk (warp w owns k-block w of the box), read weight fragments
//    with conflict-free 16-byte LDS (the 128-B swizzle spreads the 8 rows of a quarter-warp over all banks) and feed
//    mma.sync m16n8k16 (weights = the "M" operand, the <= 8 activation rows per tile = "N");
//    the k-permutation a 16-byte load implies is applied to both operands, which leaves the dot products unchanged.
//    Partial sums are reduced across warps through (double-buffered) shared memory once per tile.
// Fusions: bias / activation / SiLU(gate)*up / row scale / residual, an RMSNorm prologue (decode q,k,v straight from
// the un-normalised hidden state), the MoE dispatch (A rows gathered through the slot -> token map) and combine
// (scatter to the token row scaled by its gate value + residual).
// Replaces the same nn.Linear call sites as K1 when M is the decode batch.
#include <cuda.h>

#include <cstring>
#include <mutex>
#include <unordered_map>

#include "internal.h"
#include "ptx.cuh"

namespace mpl {

constexpr int SK_CONSUMERS = 8;
constexpr int SK_THREADS = (SK_CONSUMERS + 1) * 32;
constexpr int SK_ROWS = 16;                   // weight rows per tile (one mma.sync M-tile)
constexpr int SK_KC = 512;                    // K elements per stage
constexpr int SK_STAGE_BYTES = SK_ROWS * SK_KC * 2;  // [8 k-blocks][16 rows][128 B, swizzled]
constexpr int SK_STAGES = 6;
constexpr int SK_MAXY = 8;

struct SkinnyMaps {
  CUtensorMap w[SK_MAXY];   // slice -> weight matrix (3-D {64, N, K/64} when K % 64 == 0, else 2-D {K, N})
  CUtensorMap w2[SK_MAXY];  // second matrix of the SiLU(gate)*up pair
};

struct SkinnyParams {
  const __nv_bfloat16* A;
  long long lda;
  long long a_ystride;              // elements between the A blocks of consecutive slices (0 = shared A)
  int map3d;  // tensor maps are 3-D (one TMA instruction per stage) / 2-D (one per 64-wide k-block)
  void* C[SK_MAXY];
  long long ldc;
  const __nv_bfloat16* bias[SK_MAXY];
  const __nv_bfloat16* residual;
  long long ldr;
  const float* row_scale;
  const int* m_dev;  // device row count (slice y reads m_dev[y * m_dev_ystride])
  int m_dev_ystride;
  int m_dev_stable;      // m_dev was final before the previous launch started: the producer may read it before the wait
  const int* a_row_map;  // MoE dispatch: A row m of slice y = A[a_row_map[y * map_ystride + m]] (a_ystride ignored)
  const int* row_map;    // MoE combine: output/residual row of A-row m in slice y = row_map[y * map_ystride + m]
  const float* row_gate;
  int map_ystride;
  const __nv_bfloat16* ln_w;  // fused RMSNorm prologue on A (HF LlamaRMSNorm rounding) or NULL
  float ln_eps;
  int M, N, K;
  int ny;
  int act;
  int out_f32;
};

__device__ __forceinline__ void hmma_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                           uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 r;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(addr));
  return r;
}

__device__ __forceinline__ float skinny_act(float v, int act) {
  switch (act) {
    case MPL_ACT_GELU:
      return 0.5f * v * (1.0f + erff(v * 0.70710678118654752f));
    case MPL_ACT_QUICK_GELU:
      return v / (1.0f + __expf(-1.702f * v));
    case MPL_ACT_RELU:
      return fmaxf(v, 0.0f);
    case MPL_ACT_SILU:
      return v / (1.0f + __expf(-v));
    case MPL_ACT_SIGMOID:
      return 1.0f / (1.0f + __expf(-v));
    default:
      return v;
  }
}

// x (8 bf16) -> RMSNorm'ed bf16 with the reference's two roundings: w * bf16(x * rstd)
__device__ __forceinline__ uint4 rms_apply(uint4 x, uint4 w, float rstd) {
  const __nv_bfloat162* xp = reinterpret_cast<const __nv_bfloat162*>(&x);
  const __nv_bfloat162* wp = reinterpret_cast<const __nv_bfloat162*>(&w);
  uint4 o;
  uint32_t* op = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 xf = __bfloat1622float2(xp[i]), wf = __bfloat1622float2(wp[i]);
    op[i] = pack_bf16(wf.x * bf16_round(xf.x * rstd), wf.y * bf16_round(xf.y * rstd));
  }
  return o;
}

// bit mask of the slices that have rows (an expert without tokens contributes no tile: none of its bytes are read)
__device__ __forceinline__ uint32_t active_slices(const SkinnyParams& p) {
  if (p.m_dev == nullptr) return (1u << p.ny) - 1u;
  uint32_t m = 0;
  for (int y = 0; y < p.ny; ++y)
    if (ld_cg_s32(p.m_dev + y * p.m_dev_ystride) > 0) m |= 1u << y;
  return m;
}

template <int MT, bool DUAL>
__global__ void __launch_bounds__(SK_THREADS, 1)
skinny_gemm_kernel(const __grid_constant__ SkinnyMaps maps, const SkinnyParams p) {
  constexpr int OUT_ROWS = DUAL ? SK_ROWS / 2 : SK_ROWS;  // output features per tile
  constexpr int RP = MT * 8 + 1;
  constexpr int RED_FLOATS = SK_CONSUMERS * SK_ROWS * RP;
  extern __shared__ uint8_t sk_smem_raw[];
  uint8_t* sk_smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(sk_smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sk_smem + SK_STAGES * SK_STAGE_BYTES);
  uint64_t* empty_bar = full_bar + SK_STAGES;
  float* s_rstd = reinterpret_cast<float*>(empty_bar + SK_STAGES);  // [16]
  float* red = s_rstd + 16;                                         // [2][SK_CONSUMERS][SK_ROWS][RP]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_per_slice = (p.N + OUT_ROWS - 1) / OUT_ROWS;
  const int chunks = (p.K + SK_KC - 1) / SK_KC;

  if (threadIdx.x == 0) {
    for (int s = 0; s < SK_STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], SK_CONSUMERS);
    }
    fence_mbar_init();
  }
  __syncthreads();

  if (warp == SK_CONSUMERS) {
    // ------------------------------------------------------------ producer: one elected thread, TMA boxes
    if (lane != 0) return;
    if (p.m_dev != nullptr && !p.m_dev_stable) griddep_wait();
    const uint32_t amask = active_slices(p);
    const int total = __popc(amask) * tiles_per_slice;
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
      const int y = __fns(amask, 0, tile / tiles_per_slice + 1);
      const int n0 = (tile % tiles_per_slice) * OUT_ROWS;
      const CUtensorMap* m0 = &maps.w[y];
      const CUtensorMap* m1 = &maps.w2[y];
      for (int c = 0; c < chunks; ++c) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* dst = sk_smem + stage * SK_STAGE_BYTES;
        if (p.map3d) {
          mbar_expect_tx(&full_bar[stage], SK_STAGE_BYTES);
          tma_load_3d(dst, m0, &full_bar[stage], 0, n0, c * (SK_KC / 64));
          if (DUAL) tma_load_3d(dst + SK_STAGE_BYTES / 2, m1, &full_bar[stage], 0, n0, c * (SK_KC / 64));
        } else {
          const int kbs = min(SK_KC / 64, (p.K - c * SK_KC + 63) / 64);
          mbar_expect_tx(&full_bar[stage], static_cast<uint32_t>(kbs * SK_ROWS * 128));
          for (int kb = 0; kb < kbs; ++kb) {
            tma_load_2d(dst + kb * OUT_ROWS * 128, m0, &full_bar[stage], c * SK_KC + kb * 64, n0);
            if (DUAL)
              tma_load_2d(dst + SK_STAGE_BYTES / 2 + kb * OUT_ROWS * 128, m1, &full_bar[stage], c * SK_KC + kb * 64, n0);
          }
        }
        if (++stage == SK_STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
    return;
  }

  // -------------------------------------------------------------- consumers
  griddep_wait();  // activations / row counts / maps come from the preceding kernels
  if (threadIdx.x == 0) griddep_launch_dependents();
  const int g = lane >> 2, t = lane & 3;
  const int ctid = threadIdx.x;  // 0..255
  const uint32_t amask = active_slices(p);
  const int total = __popc(amask) * tiles_per_slice;
  if (p.ln_w != nullptr) {
    // RMSNorm statistics of the (<= MT*8) activation rows: every CTA recomputes them (M*K*2 bytes from L2)
    for (int m = warp; m < MT * 8; m += SK_CONSUMERS) {
      float ss = 0.0f;
      if (m < p.M) {
        const __nv_bfloat16* xr = p.A + static_cast<long long>(m) * p.lda;
        for (int k = lane * 8; k < p.K; k += 256) {
          const uint4 raw = *reinterpret_cast<const uint4*>(xr + k);
          const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float2 f = __bfloat1622float2(h[i]);
            ss += f.x * f.x + f.y * f.y;
          }
        }
      }
      ss = warp_sum(ss);
      if (lane == 0) s_rstd[m] = rsqrtf(ss / static_cast<float>(p.K) + p.ln_eps);
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");  // consumers only (the producer warp is busy copying)
  }
  float rstd[MT];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) rstd[mt] = (p.ln_w != nullptr) ? s_rstd[mt * 8 + g] : 1.0f;

  int stage = 0;
  uint32_t phase = 0;
  int buf = 0;
  for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
    const int y = __fns(amask, 0, tile / tiles_per_slice + 1);
    const int n0 = (tile % tiles_per_slice) * OUT_ROWS;
    int M = p.M;
    if (p.m_dev != nullptr) M = min(M, ld_cg_s32(p.m_dev + y * p.m_dev_ystride));
    const __nv_bfloat16* xp[MT];
    bool xok[MT];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      const int m = mt * 8 + g;
      xok[mt] = m < M;
      long long arow = m;
      const __nv_bfloat16* base = p.A + static_cast<long long>(y) * p.a_ystride;
      if (p.a_row_map != nullptr) {
        base = p.A;
        arow = xok[mt] ? p.a_row_map[y * p.map_ystride + m] : 0;
        if (arow < 0) {
          arow = 0;
          xok[mt] = false;
        }
      }
      xp[mt] = base + (xok[mt] ? arow : 0) * p.lda;
    }
    float acc[MT][4];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[mt][i] = 0.0f;

    for (int c = 0; c < chunks; ++c) {
      const int kbase = c * SK_KC + warp * 64;  // this warp's 64-element slice of the chunk
      // activation fragments first (global / L1), so their latency overlaps the wait on the weights
      uint4 xb[2][MT];
#pragma unroll
      for (int h2 = 0; h2 < 2; ++h2) {
        const int k = kbase + h2 * 32 + t * 8;
        const bool ok = k < p.K;
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          uint4 v = (ok && xok[mt]) ? *reinterpret_cast<const uint4*>(xp[mt] + k) : make_uint4(0, 0, 0, 0);
          if (p.ln_w != nullptr && ok && xok[mt])
            v = rms_apply(v, *reinterpret_cast<const uint4*>(p.ln_w + k), rstd[mt]);
          xb[h2][mt] = v;
        }
      }
      mbar_wait(&full_bar[stage], phase);
      // stage layout: [k-block = warp][row][128 B]; 16-byte chunk j of row r sits at chunk j ^ (r & 7) (128-B swizzle).
      // DUAL: gate rows in the first half of the stage, up rows in the second (8-row boxes).
      const uint32_t sbase = smem_u32(sk_smem + stage * SK_STAGE_BYTES) + warp * (OUT_ROWS * 128) + g * 128;
#pragma unroll
      for (int h2 = 0; h2 < 2; ++h2) {
        const int kin = warp * 64 + h2 * 32 + t * 8;  // element offset inside the chunk
        const bool ok = c * SK_KC + kin < p.K;
        const uint32_t sw = static_cast<uint32_t>(((h2 * 4 + t) ^ g) * 16);
        uint4 wa = lds128(sbase + sw);
        uint4 wb = lds128(sbase + (DUAL ? SK_STAGE_BYTES / 2 : 1024) + sw);
        if (!ok) {  // beyond K the stage holds stale bytes
          wa = make_uint4(0, 0, 0, 0);
          wb = make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          hmma_16816(acc[mt], wa.x, wb.x, wa.y, wb.y, xb[h2][mt].x, xb[h2][mt].y);
          hmma_16816(acc[mt], wa.z, wb.z, wa.w, wb.w, xb[h2][mt].z, xb[h2][mt].w);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_bar[stage]);
      if (++stage == SK_STAGES) {
        stage = 0;
        phase ^= 1;
      }
    }
    // C fragment: c0,c1 -> (weight row g, m = 2t, 2t+1); c2,c3 -> (weight row g+8, same m)
    float* rbuf = red + buf * RED_FLOATS;
    buf ^= 1;  // double-buffered: the next tile's partials never race with this tile's epilogue reads
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      float* rw = rbuf + warp * SK_ROWS * RP;
      rw[g * RP + mt * 8 + t * 2] = acc[mt][0];
      rw[g * RP + mt * 8 + t * 2 + 1] = acc[mt][1];
      rw[(g + 8) * RP + mt * 8 + t * 2] = acc[mt][2];
      rw[(g + 8) * RP + mt * 8 + t * 2 + 1] = acc[mt][3];
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");
    void* Cout = p.C[y];
    const __nv_bfloat16* bias = p.bias[y];
    for (int idx = ctid; idx < OUT_ROWS * MT * 8; idx += SK_CONSUMERS * 32) {
      const int r = idx % OUT_ROWS;
      const int m = idx / OUT_ROWS;
      const int n = n0 + r;
      if (m >= M || n >= p.N) continue;
      float v = 0.0f, v2 = 0.0f;
#pragma unroll
      for (int w = 0; w < SK_CONSUMERS; ++w) {
        v += rbuf[(w * SK_ROWS + r) * RP + m];
        if (DUAL) v2 += rbuf[(w * SK_ROWS + r + OUT_ROWS) * RP + m];
      }
      const bool f32 = p.out_f32 != 0;
      if (DUAL) {
        const float gte = bf16_round(v), up = bf16_round(v2);
        v = bf16_round(gte / (1.0f + __expf(-gte))) * up;
      }
      if (bias != nullptr) v += __bfloat162float(bias[n]);
      if (p.act != MPL_ACT_NONE) v = skinny_act(f32 ? v : bf16_round(v), p.act);
      if (p.row_scale != nullptr) v = (f32 ? v : bf16_round(v)) * p.row_scale[m];
      long long orow = m;
      if (p.row_map != nullptr) {
        // MoE combine: out[token] = residual[token] + bf16(bf16(gate) * y)   (combine_weights.type_as(x))
        const int slot = y * p.map_ystride + m;
        orow = p.row_map[slot];
        if (orow < 0) continue;
        v = bf16_round(v) * bf16_round(p.row_gate[slot]);
      }
      if (p.residual != nullptr) v = (f32 ? v : bf16_round(v)) + __bfloat162float(p.residual[orow * p.ldr + n]);
      if (f32)
        reinterpret_cast<float*>(Cout)[orow * p.ldc + n] = v;
      else
        reinterpret_cast<__nv_bfloat16*>(Cout)[orow * p.ldc + n] = __float2bfloat16_rn(v);
    }
  }
}

template <int MT, bool DUAL>
static int launch_skinny(const SkinnyMaps& maps, const SkinnyParams& p, cudaStream_t stream) {
  constexpr int smem = 1024 + SK_STAGES * SK_STAGE_BYTES + 2 * SK_STAGES * 8 + 16 * 4 +
                       2 * SK_CONSUMERS * SK_ROWS * (MT * 8 + 1) * 4;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(skinny_gemm_kernel<MT, DUAL>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) !=
        cudaSuccess)
      return MPL_ERR_CUDA;
    attr_set = true;
  }
  const long long tiles = static_cast<long long>((p.N + (DUAL ? 8 : 16) - 1) / (DUAL ? 8 : 16)) * p.ny;
  const int grid = static_cast<int>(tiles < num_sms() ? tiles : num_sms());
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(SK_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (cudaLaunchKernelEx(&cfg, skinny_gemm_kernel<MT, DUAL>, maps, p) != cudaSuccess) {
    ++g_launches;
    return MPL_ERR_CUDA;
  }
  return launch_status();
}

// Host-side weight description of a launch (the device side only sees tensor maps).
struct SkinnyWeights {
  const void* W[SK_MAXY];
  const void* W2[SK_MAXY];
  long long ldw;
};

// Tensor maps are pure functions of (pointer, shape, box); encoding costs a driver call, so they are cached.
struct TmKey {
  const void* ptr;
  long long ld;
  int N, K, rows;
  bool operator==(const TmKey& o) const { return ptr == o.ptr && ld == o.ld && N == o.N && K == o.K && rows == o.rows; }
};
struct TmKeyHash {
  size_t operator()(const TmKey& k) const {
    size_t h = reinterpret_cast<size_t>(k.ptr) * 0x9E3779B97F4A7C15ull;
    h ^= static_cast<size_t>(k.ld) * 0xC2B2AE3D27D4EB4Full + static_cast<size_t>(k.N) * 1315423911u +
         static_cast<size_t>(k.K) * 2654435761u + static_cast<size_t>(k.rows);
    return h;
  }
};
static std::unordered_map<TmKey, CUtensorMap, TmKeyHash> g_tm_cache;
static std::mutex g_tm_mutex;

static int weight_tmap(CUtensorMap* out, const void* W, int N, int K, long long ldw, int rows) {
  const TmKey key{W, ldw, N, K, rows};
  std::lock_guard<std::mutex> lock(g_tm_mutex);
  auto it = g_tm_cache.find(key);
  if (it != g_tm_cache.end()) {
    *out = it->second;
    return MPL_OK;
  }
  int rc;
  if ((K % 64) == 0) {
    const unsigned long long dims[3] = {64ull, static_cast<unsigned long long>(N), static_cast<unsigned long long>(K / 64)};
    const unsigned long long strides[2] = {static_cast<unsigned long long>(ldw) * 2, 128ull};
    const unsigned box[3] = {64u, static_cast<unsigned>(rows), static_cast<unsigned>(SK_KC / 64)};
    rc = encode_tmap_bf16(out, W, 3, dims, strides, box);
  } else {
    const unsigned long long dims[2] = {static_cast<unsigned long long>(K), static_cast<unsigned long long>(N)};
    const unsigned long long strides[1] = {static_cast<unsigned long long>(ldw) * 2};
    const unsigned box[2] = {64u, static_cast<unsigned>(rows)};
    rc = encode_tmap_bf16(out, W, 2, dims, strides, box);
  }
  if (rc != MPL_OK) return rc;
  if (g_tm_cache.size() > 16384) g_tm_cache.clear();
  g_tm_cache.emplace(key, *out);
  return MPL_OK;
}

static int skinny_dispatch(SkinnyParams& p, const SkinnyWeights& w, int ny, bool dual, cudaStream_t stream) {
  p.ny = ny;
  if (p.ln_w != nullptr && (p.a_ystride != 0 || p.a_row_map != nullptr)) return MPL_ERR_UNSUPPORTED;
  if ((p.K % 8) != 0 || (p.lda % 8) != 0 || (w.ldw % 8) != 0 || (p.a_ystride % 8) != 0 ||
      (reinterpret_cast<uintptr_t>(p.A) & 15) != 0)
    return MPL_ERR_ALIGN;
  p.map3d = (p.K % 64) == 0;
  SkinnyMaps maps;
  const int rows = dual ? SK_ROWS / 2 : SK_ROWS;
  for (int i = 0; i < ny; ++i) {
    int rc = weight_tmap(&maps.w[i], w.W[i], p.N, p.K, w.ldw, rows);
    if (rc == MPL_OK && dual) rc = weight_tmap(&maps.w2[i], w.W2[i], p.N, p.K, w.ldw, rows);
    if (rc != MPL_OK) return rc;
  }
  if (p.M <= 8) return dual ? launch_skinny<1, true>(maps, p, stream) : launch_skinny<1, false>(maps, p, stream);
  return dual ? launch_skinny<2, true>(maps, p, stream) : launch_skinny<2, false>(maps, p, stream);
}

int skinny_gemm_bf16(const mpl_gemm_args& a, cudaStream_t stream) {
  if (a.M <= 0 || a.N <= 0) return MPL_OK;
  const int nb = a.nb < 1 ? 1 : a.nb;
  if (a.M > 16 || nb > 3 || (a.B2 != nullptr && nb != 1)) return MPL_ERR_UNSUPPORTED;
  if (a.K <= 0 || a.A == nullptr || a.B[0] == nullptr || a.C[0] == nullptr) return MPL_ERR_ARG;
  SkinnyParams p;
  SkinnyWeights w;
  memset(&p, 0, sizeof(p));
  memset(&w, 0, sizeof(w));
  p.A = static_cast<const __nv_bfloat16*>(a.A);
  p.lda = a.lda;
  for (int i = 0; i < nb; ++i) {
    w.W[i] = a.B[i];
    w.W2[i] = a.B2;
    p.C[i] = a.C[i];
    p.bias[i] = static_cast<const __nv_bfloat16*>(a.bias[i]);
  }
  w.ldw = a.ldb;
  p.ldc = a.ldc;
  p.residual = static_cast<const __nv_bfloat16*>(a.residual);
  p.ldr = a.ldr;
  p.row_scale = a.row_scale;
  p.m_dev = a.m_dev;
  p.ln_w = static_cast<const __nv_bfloat16*>(a.ln_weight);
  p.ln_eps = a.ln_eps;
  p.M = a.M;
  p.N = a.N;
  p.K = a.K;
  p.act = a.act;
  p.out_f32 = a.out_dtype == MPL_DT_F32;
  return skinny_dispatch(p, w, nb, a.B2 != nullptr, stream);
}

// Experts of one MoE layer in one launch (one tile list over all experts), rows per expert read on the device.
int skinny_grouped_gemm_bf16(const mpl_grouped_gemm_args& a, cudaStream_t stream) {
  if (a.M <= 0 || a.N <= 0 || a.groups <= 0) return MPL_OK;
  if (a.M > 16 || a.groups > SK_MAXY) return MPL_ERR_UNSUPPORTED;
  if (a.K <= 0 || a.A == nullptr || a.C == nullptr) return MPL_ERR_ARG;
  SkinnyParams p;
  SkinnyWeights w;
  memset(&p, 0, sizeof(p));
  memset(&w, 0, sizeof(w));
  p.A = static_cast<const __nv_bfloat16*>(a.A);
  p.lda = a.lda;
  p.a_ystride = a.a_group_stride;
  const bool dual = a.B2[0] != nullptr;
  const long long csz = a.out_dtype == MPL_DT_F32 ? 4 : 2;
  for (int g = 0; g < a.groups; ++g) {
    if (a.B[g] == nullptr || (dual && a.B2[g] == nullptr)) return MPL_ERR_ARG;
    w.W[g] = a.B[g];
    w.W2[g] = a.B2[g];
    p.C[g] = static_cast<char*>(a.C) + (a.row_map ? 0 : static_cast<long long>(g) * a.c_group_stride * csz);
  }
  w.ldw = a.ldb;
  p.ldc = a.ldc;
  p.residual = static_cast<const __nv_bfloat16*>(a.residual);
  p.ldr = a.ldr;
  p.m_dev = a.m_dev;
  p.m_dev_ystride = 1;
  p.m_dev_stable = a.m_dev_stable;
  p.a_row_map = a.a_row_map;
  p.row_map = a.row_map;
  p.row_gate = a.row_gate;
  p.map_ystride = static_cast<int>(a.map_group_stride);
  p.M = a.M;
  p.N = a.N;
  p.K = a.K;
  p.act = a.act;
  p.out_f32 = a.out_dtype == MPL_DT_F32;
  return skinny_dispatch(p, w, a.groups, dual, stream);
}

// Dispatcher used by the host side: tensor-core tiles for M > 16, streaming kernel otherwise.
int linear_bf16(const mpl_gemm_args& a, cudaStream_t stream) {
  if (a.lora_r != 0 || a.ext_a != nullptr || a.dual_g != nullptr || a.silu_bwd_g != nullptr) return gemm_bf16(a, stream);  // fused LoRA: tensor-core path only
  if (a.M <= 16 && (a.K % 8) == 0) return skinny_gemm_bf16(a, stream);
  if (a.ln_weight != nullptr) return MPL_ERR_UNSUPPORTED;  // the RMSNorm prologue exists on the streaming path only
  return gemm_bf16(a, stream);
}

}  // namespace mpl

extern "C" int mpl_skinny_gemm_bf16(const mpl_gemm_args* args, void* stream) {
  if (args == nullptr) return MPL_ERR_ARG;
  return mpl::skinny_gemm_bf16(*args, static_cast<cudaStream_t>(stream));
}
extern "C" int mpl_linear_bf16(const mpl_gemm_args* args, void* stream) {
  if (args == nullptr) return MPL_ERR_ARG;
  return mpl::linear_bf16(*args, static_cast<cudaStream_t>(stream));
}
extern "C" int mpl_grouped_gemm_bf16(const mpl_grouped_gemm_args* args, void* stream) {
  if (args == nullptr) return MPL_ERR_ARG;
  if (args->M <= 16) return mpl::skinny_grouped_gemm_bf16(*args, static_cast<cudaStream_t>(stream));
  return mpl::grouped_gemm_bf16(*args, static_cast<cudaStream_t>(stream));
}
