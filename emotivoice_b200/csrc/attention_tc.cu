// Multi-head self-attention on the 5th-gen tensor cores (sm_100a): S = Q K^T and O = P V are tcgen05.mma (kind::tf32, fp32
// accumulation in TMEM), the softmax runs between two TMEM reads with thread = query row (a tcgen05.ld lane), so row maxima
// and sums are thread-local -- no shuffles, no (L x L) tensor in memory.  Replaces encoder.py:84-109
// (scores = q k^T / sqrt(d_k); key-padding mask; softmax; p v) for d_k = 48 (EmotiVoice: 384 / 8).
//
// One CTA = 128 queries of one (batch item, head); keys are walked ONCE in tiles of 64 with a lazily rescaled online softmax:
//   S_j = Q K_j^T  ->  P_j = exp(S_j / sqrt(d_k) - m), l += rowsum(P_j), O += P_j V_j
// m is the running row maximum; it only moves (and O, l are only rescaled by exp(m_old - m_new), a TMEM load / multiply / store of
// the row's 48 accumulator columns) when a tile's maximum exceeds it by more than 8: until then P_j <= e^8, harmless in fp32 and
// in the relative precision of the tf32 operand.  Softmax is shift invariant, so the result is the reference's up to rounding.
// Final: ctx = O / l.
//
// Operands are staged in shared memory in the no-swizzle K-major UMMA layout of conv1d_tc.cu (element (row r, 16-byte
// granule g) at (g * rows_pad + r) * 16):  Q [12 granules][128 rows], K_j [12][64 rows] (B operand of S), V_j^T [16 key
// granules][48 rows = d] (B operand of O: the loader warps transpose 4x4 blocks in registers), P_j [16 key granules][128 rows]
// (A operand of O, written by the softmax threads: thread = row, 16 B per granule -> conflict free).
// MODE 1 = 3xTF32 fp32 emulation (x = hi + lo, three MMAs per K step): the duration-critical encoder prefix and the "fp32"
// precision; MODE 0 = one tf32 MMA per K step (operands rounded to nearest).
//
// Roles (288 threads): warps 0-3 softmax (warp w <-> TMEM lanes 32w..), warps 4-7 loaders, warp 8 TMEM alloc + MMA issue.
// mbarriers: q_ready, kv_full/kv_empty[2], s_full/s_empty[2] (S is double buffered in TMEM: Q K_{j+1}^T runs under the
// softmax of tile j), p_full/p_empty (the O rescale sits between p_empty -- P V_{j-1} has completed -- and p_full), o_full.
#include "ev_common.cuh"
#include "tc_common.cuh"

namespace ev {
namespace atc {

using namespace tc;

constexpr int BQ = 128, BKT = 64;
constexpr int NSW = 4, NLW = 4;                  // softmax warps, loader warps
constexpr int W_LOAD = NSW, W_MMA = NSW + NLW;
constexpr int ATC_THREADS = (W_MMA + 1) * 32;    // 288
constexpr int QPAD = BQ + 1, KPAD = BKT + 1;     // rows_pad == 1 (mod 8): granule-fastest 16-byte stores are conflict free

template <int DK>
struct Smem {
  static constexpr int G = DK / 4;               // channel granules of Q / K
  static constexpr int GK = BKT / 4;             // key granules of V^T / P
  static constexpr int VPAD = DK + 1;
  static constexpr int q_plane = G * QPAD * 16;
  static constexpr int k_plane = G * KPAD * 16;
  static constexpr int v_plane = GK * VPAD * 16;
  static constexpr int p_plane = GK * BQ * 16;
  static constexpr int head = 256;
  static constexpr int total(int planes) { return head + planes * (q_plane + 2 * k_plane + 2 * v_plane + p_plane); }
};

template <int DK, int MODE>
__global__ void __launch_bounds__(ATC_THREADS, 1) attention_tc_kernel(const float* __restrict__ qkv, const int32_t* __restrict__ key_lens,
                                                                      float* __restrict__ ctx, int L, int H) {
  constexpr bool SPLIT3 = (MODE == 1);
  constexpr int PL = SPLIT3 ? 2 : 1;
  using S = Smem<DK>;
  constexpr int G = S::G, GK = S::GK, VPAD = S::VPAD;
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * BQ;
  const size_t ld = (size_t)3 * H;
  const float* base = qkv + (size_t)b * L * ld;

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem_raw + 128);
  uint8_t* q_s = smem_raw + S::head;                       // [plane][G][QPAD][16]
  uint8_t* k_s = q_s + PL * S::q_plane;                    // [stage][plane][G][KPAD][16]
  uint8_t* v_s = k_s + 2 * PL * S::k_plane;                // [stage][plane][GK][VPAD][16]
  uint8_t* p_s = v_s + 2 * PL * S::v_plane;                // [plane][GK][BQ][16]
  const uint32_t bar_base = smem_u32(bars);
  const uint32_t q_ready = bar_base, p_full = bar_base + 8, p_empty = bar_base + 16, o_full = bar_base + 24;
  auto kv_full = [&](int s) { return bar_base + 32u + 8u * s; };
  auto kv_empty = [&](int s) { return bar_base + 48u + 8u * s; };
  auto s_full = [&](int s) { return bar_base + 64u + 8u * s; };
  auto s_empty = [&](int s) { return bar_base + 80u + 8u * s; };

  if (tid == 0) {
    mbar_init(q_ready, NLW * 32); mbar_init(p_full, NSW * 32); mbar_init(p_empty, 1); mbar_init(o_full, 1);
    for (int s = 0; s < 2; ++s) { mbar_init(kv_full(s), NLW * 32); mbar_init(kv_empty(s), 1); mbar_init(s_full(s), 1); mbar_init(s_empty(s), NSW * 32); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  constexpr uint32_t TMEM_COLS = 256;           // S0 [0,64) | S1 [64,128) | O [128, 128+DK)
  if (warp == W_MMA) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // Programmatic dependent launch: the set-up above ran under the predecessor's tail; EVERY thread now waits for the preceding grids
  // before anything is read -- including key_lens, which in the encoder is written by validate_inputs_kernel a few launches upstream
  // (a role that decoded its tile count from stale lengths would desynchronise the pipeline).
  asm volatile("griddepcontrol.launch_dependents;\n\tgriddepcontrol.wait;" ::: "memory");
  const int klen = key_lens ? min(L, key_lens[b]) : L;
  const int nkt = (klen + BKT - 1) / BKT;

  if (warp < NSW) {
    // ================================ softmax: thread = query row ================================================
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const int r = warp * 32 + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);
    const float inv_sqrt_dk = 1.0f / sqrtf((float)DK);
    float m = -INFINITY, l = 0.f;
    float v[64];
    constexpr float RESCALE_AT = 8.0f;
    for (int j = 0; j < nkt; ++j) {
      const int sb = j & 1;
      mbar_wait(s_full(sb), (j >> 1) & 1);
      tc_fence_after();
      tmem_ld32(lane_addr + (uint32_t)(sb * BKT), 32, v);
      tmem_ld32(lane_addr + (uint32_t)(sb * BKT + 32), 32, v + 32);
      tc_fence_before();
      mbar_arrive(s_empty(sb));                  // S_j is in registers: the buffer may take Q K_{j+2}^T
      const int nvalid = min(BKT, klen - j * BKT);
      float mt = -INFINITY;
#pragma unroll
      for (int c = 0; c < BKT; ++c) {
        v[c] *= inv_sqrt_dk;
        if (c < nvalid) mt = fmaxf(mt, v[c]);
      }
      mbar_wait(p_empty, ((j & 1) ^ 1));         // the MMAs of tile j-1 have read P and have finished accumulating into O
      tc_fence_after();
      // lazy rescale: tcgen05.ld / st are warp collectives, so the warp decides together and rows that need nothing scale by 1
      const bool need = (j > 0) && (mt > m + RESCALE_AT);
      if (j == 0) m = mt;
      if (__any_sync(0xffffffffu, need)) {
        const float f = need ? __expf(m - mt) : 1.0f;
        if (need) { m = mt; l *= f; }
        float o[64];
        tmem_ld32(lane_addr + 128u, 32, o);
        tmem_ld32(lane_addr + 160u, 16, o + 32);      // columns 32..47 (the helper zero-fills o[48..63])
#pragma unroll
        for (int c = 0; c < DK; ++c) o[c] *= f;
        tmem_st32(lane_addr + 128u, o);
        tmem_st16(lane_addr + 160u, o + 32);
        tmem_st_wait();
      }
#pragma unroll
      for (int g = 0; g < GK; ++g) {
        float p[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int c = 4 * g + e;
          const float x = v[c] - m;
          p[e] = c < nvalid ? __expf(x) : 0.f;      // ex2.approx: relative error 2^-22, below the 3xTF32 product error; x <= 8
        }
        const float4 hi = make_float4(to_tf32(p[0]), to_tf32(p[1]), to_tf32(p[2]), to_tf32(p[3]));
        // the denominator sums exactly what the tensor core multiplies: p (= hi + lo) in the 3xTF32 mode, the rounded hi otherwise
        l += SPLIT3 ? (p[0] + p[1]) + (p[2] + p[3]) : (hi.x + hi.y) + (hi.z + hi.w);
        *reinterpret_cast<float4*>(p_s + ((size_t)g * BQ + r) * 16) = hi;
        if (SPLIT3) {
          const float4 lo = make_float4(to_tf32(p[0] - hi.x), to_tf32(p[1] - hi.y), to_tf32(p[2] - hi.z), to_tf32(p[3] - hi.w));
          *reinterpret_cast<float4*>(p_s + S::p_plane + ((size_t)g * BQ + r) * 16) = lo;
        }
      }
      fence_proxy_async();
      tc_fence_before();
      mbar_arrive(p_full);
    }
    // ctx = O / l   (an item without keys -- rejected by the module's input validation -- yields zeros instead of a hang)
    float o[64];
    if (nkt > 0) {
      mbar_wait(o_full, 0);
      tc_fence_after();
      tmem_ld32(lane_addr + 128u, 32, o);
      if (DK > 32) tmem_ld32(lane_addr + 160u, DK - 32, o + 32);
    } else {
#pragma unroll
      for (int c = 0; c < 64; ++c) o[c] = 0.f;
      l = 1.f;
    }
    const int row = q0 + r;
    if (row < L) {
      const float inv = 1.0f / l;
      float* ob = ctx + ((size_t)b * L + row) * H + h * DK;
#pragma unroll
      for (int c = 0; c < DK; c += 4)
        *reinterpret_cast<float4*>(ob + c) = make_float4(o[c] * inv, o[c + 1] * inv, o[c + 2] * inv, o[c + 3] * inv);
    }
    tc_fence_before();
  } else if (warp < W_MMA) {
    // ================================ loaders =====================================================================
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const int lt = (warp - W_LOAD) * 32 + lane;        // 0..127
    // Every tile is loaded in batches of independent 16-byte loads issued back to back (memory-level parallelism: one latency
    // per batch instead of one per element -- the first version of this loop serialised them and was loader bound), then
    // rounded / split and stored in operand layout.
    auto store_split = [&](uint8_t* dst, int plane_bytes, const float4& t) {
      const float4 hi = make_float4(to_tf32(t.x), to_tf32(t.y), to_tf32(t.z), to_tf32(t.w));
      *reinterpret_cast<float4*>(dst) = hi;
      if (SPLIT3) {
        const float4 lo = make_float4(to_tf32(t.x - hi.x), to_tf32(t.y - hi.y), to_tf32(t.z - hi.z), to_tf32(t.w - hi.w));
        *reinterpret_cast<float4*>(dst + plane_bytes) = lo;
      }
    };
    constexpr int NT = NLW * 32;
    constexpr int QB = 6;                                  // loads in flight per thread
    static_assert((BQ * G) % (NT * QB) == 0 && (BKT * G) == NT * QB, "loader batching assumes d_k = 48, 128 loader threads");
    // Q tile: (row, granule) pairs, granule fastest -> coalesced 192-byte head rows; rows >= L are zeros
    for (int base_idx = 0; base_idx < BQ * G; base_idx += NT * QB) {
      float4 t[QB];
#pragma unroll
      for (int u = 0; u < QB; ++u) {
        const int idx = base_idx + u * NT + lt;
        const int r = idx / G, g = idx - r * G;
        const int row = q0 + r;
        t[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row < L) t[u] = __ldg(reinterpret_cast<const float4*>(base + (size_t)row * ld + h * DK + g * 4));
      }
#pragma unroll
      for (int u = 0; u < QB; ++u) {
        const int idx = base_idx + u * NT + lt;
        const int r = idx / G, g = idx - r * G;
        store_split(q_s + ((size_t)g * QPAD + r) * 16, S::q_plane, t[u]);
      }
    }
    fence_proxy_async();
    mbar_arrive(q_ready);
    for (int j = 0; j < nkt; ++j) {
      const int s = j & 1;
      const int k0 = j * BKT;
      // all global loads of the tile first (K: 6 per thread; V: one or two 4-key x 4-channel blocks = 4 or 8 per thread) ...
      float4 kt[QB];
#pragma unroll
      for (int u = 0; u < QB; ++u) {
        const int idx = u * NT + lt;
        const int r = idx / G, g = idx - r * G;
        const int row = k0 + r;
        kt[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row < klen) kt[u] = __ldg(reinterpret_cast<const float4*>(base + (size_t)row * ld + H + h * DK + g * 4));
      }
      constexpr int VB = (GK * G + NT - 1) / NT;         // blocks per thread (2 for d_k = 48; the second only for lt < 64)
      float4 vt[VB][4];
#pragma unroll
      for (int u = 0; u < VB; ++u) {
        const int idx = u * NT + lt;
        const int gk = idx / G, d4 = idx - gk * G;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int row = k0 + 4 * gk + i;
          vt[u][i] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (idx < GK * G && row < klen) vt[u][i] = __ldg(reinterpret_cast<const float4*>(base + (size_t)row * ld + 2 * H + h * DK + d4 * 4));
        }
      }
      // ... then wait for the ring slot and store
      mbar_wait(kv_empty(s), ((j >> 1) & 1) ^ 1);
      uint8_t* kd = k_s + (size_t)s * PL * S::k_plane;
#pragma unroll
      for (int u = 0; u < QB; ++u) {
        const int idx = u * NT + lt;
        const int r = idx / G, g = idx - r * G;
        store_split(kd + ((size_t)g * KPAD + r) * 16, S::k_plane, kt[u]);
      }
      // V_j^T: a 4 keys x 4 channels block per (key granule gk, channel group d4), transposed in registers
      uint8_t* vd = v_s + (size_t)s * PL * S::v_plane;
#pragma unroll
      for (int u = 0; u < VB; ++u) {
        const int idx = u * NT + lt;
        if (idx < GK * G) {
          const int gk = idx / G, d4 = idx - gk * G;
          const float4* t = vt[u];
          store_split(vd + ((size_t)gk * VPAD + d4 * 4 + 0) * 16, S::v_plane, make_float4(t[0].x, t[1].x, t[2].x, t[3].x));
          store_split(vd + ((size_t)gk * VPAD + d4 * 4 + 1) * 16, S::v_plane, make_float4(t[0].y, t[1].y, t[2].y, t[3].y));
          store_split(vd + ((size_t)gk * VPAD + d4 * 4 + 2) * 16, S::v_plane, make_float4(t[0].z, t[1].z, t[2].z, t[3].z));
          store_split(vd + ((size_t)gk * VPAD + d4 * 4 + 3) * 16, S::v_plane, make_float4(t[0].w, t[1].w, t[2].w, t[3].w));
        }
      }
      fence_proxy_async();
      mbar_arrive(kv_full(s));
    }
  } else {
    // ================================ MMA issuer ===================================================================
    // all 32 lanes run the warp-uniform control flow and the barrier waits; one elected lane issues the tcgen05 instructions
    {
      // instruction descriptors: D=F32, A=B=TF32, K-major, N>>3 at [17,23), M>>4 at [24,29)
      const uint32_t idesc_s = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BKT >> 3) << 17) | ((uint32_t)(BQ >> 4) << 24);
      const uint32_t idesc_o = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(DK >> 3) << 17) | ((uint32_t)(BQ >> 4) << 24);
      constexpr uint32_t q_lbo = QPAD * 16, k_lbo = KPAD * 16, v_lbo = VPAD * 16, p_lbo = BQ * 16;
      const uint64_t q_desc = make_desc(smem_u32(q_s), q_lbo, 128u), p_desc = make_desc(smem_u32(p_s), p_lbo, 128u);
      const uint64_t k_desc0 = make_desc(0u, k_lbo, 128u), v_desc0 = make_desc(0u, v_lbo, 128u);
      mbar_wait(q_ready, 0);
      tc_fence_after();
      auto issue_qk = [&](int c, bool release_kv) {       // S[c & 1] = Q K^T with the K tile in stage c & 1
        const int s = c & 1;
        mbar_wait(kv_full(s), (c >> 1) & 1);
        mbar_wait(s_empty(s), ((c >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint64_t k_desc = desc_advance(k_desc0, smem_u32(k_s + (size_t)s * PL * S::k_plane));
        const uint32_t d = tmem_base + (uint32_t)(s * BKT);
        if (elect_one()) {
#pragma unroll
          for (int k8 = 0; k8 < G / 2; ++k8) {
            const uint64_t a_hi = desc_advance(q_desc, (uint32_t)(2 * k8) * q_lbo);
            const uint64_t b_hi = desc_advance(k_desc, (uint32_t)(2 * k8) * k_lbo);
            if (SPLIT3) {
              const uint64_t a_lo = desc_advance(a_hi, (uint32_t)S::q_plane);
              const uint64_t b_lo = desc_advance(b_hi, (uint32_t)S::k_plane);
              umma_tf32(d, a_lo, b_hi, idesc_s, k8 ? 1u : 0u);
              umma_tf32(d, a_hi, b_lo, idesc_s, 1u);
              umma_tf32(d, a_hi, b_hi, idesc_s, 1u);
            } else {
              umma_tf32(d, a_hi, b_hi, idesc_s, k8 ? 1u : 0u);
            }
          }
          umma_commit(s_full(s));
          if (release_kv) umma_commit(kv_empty(s));
        }
        __syncwarp();
      };
      int cnt = 0;
      if (nkt > 0) issue_qk(cnt, false);
      for (int j = 0; j < nkt; ++j) {
        const int c = cnt + j, s = c & 1;
        if (j + 1 < nkt) issue_qk(c + 1, false);          // runs under the softmax of tile j
        mbar_wait(p_full, j & 1);
        tc_fence_after();
        const uint64_t v_desc = desc_advance(v_desc0, smem_u32(v_s + (size_t)s * PL * S::v_plane));
        const uint32_t d = tmem_base + 128u;
        if (elect_one()) {
#pragma unroll
          for (int k8 = 0; k8 < GK / 2; ++k8) {
            const uint64_t a_hi = desc_advance(p_desc, (uint32_t)(2 * k8) * p_lbo);
            const uint64_t b_hi = desc_advance(v_desc, (uint32_t)(2 * k8) * v_lbo);
            const uint32_t acc = (j | k8) ? 1u : 0u;
            if (SPLIT3) {
              const uint64_t a_lo = desc_advance(a_hi, (uint32_t)S::p_plane);
              const uint64_t b_lo = desc_advance(b_hi, (uint32_t)S::v_plane);
              umma_tf32(d, a_lo, b_hi, idesc_o, acc);
              umma_tf32(d, a_hi, b_lo, idesc_o, 1u);
              umma_tf32(d, a_hi, b_hi, idesc_o, 1u);
            } else {
              umma_tf32(d, a_hi, b_hi, idesc_o, acc);
            }
          }
          umma_commit(p_empty);
          umma_commit(kv_empty(s));
          if (j == nkt - 1) umma_commit(o_full);
        }
        __syncwarp();
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == W_MMA) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS));
  }
}

}  // namespace atc

template <int DK, int MODE>
static int launch_atc(const float* qkv, const int32_t* key_lens, float* ctx, int B, int L, int H, int heads, cudaStream_t st) {
  const int smem = atc::Smem<DK>::total(MODE == 1 ? 2 : 1);
  static std::atomic<uint64_t> attr_devs{0};
  if (first_use_on_device(attr_devs))
    cudaFuncSetAttribute(atc::attention_tc_kernel<DK, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  dim3 grid((L + atc::BQ - 1) / atc::BQ, heads, B);
  if (pdl_mode() >= 2) {
    const cudaError_t e = launch_with_pdl(atc::attention_tc_kernel<DK, MODE>, grid, dim3(atc::ATC_THREADS), (size_t)smem, st, qkv, key_lens, ctx, L, H);
    if (e != cudaSuccess) { set_error("attention_tc_kernel (PDL launch): %s", cudaGetErrorString(e)); return EV_ECUDA; }
    count_launch();
    return EV_OK;
  }
  atc::attention_tc_kernel<DK, MODE><<<grid, atc::ATC_THREADS, smem, st>>>(qkv, key_lens, ctx, L, H);
  EV_CUDA_LAUNCH_CHECK("attention_tc_kernel");
  return EV_OK;
}

void preload_attention_tc() {      // see preload_conv1d_gp
  cudaFuncSetAttribute(atc::attention_tc_kernel<48, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, atc::Smem<48>::total(1));
  cudaFuncSetAttribute(atc::attention_tc_kernel<48, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, atc::Smem<48>::total(2));
  cudaGetLastError();
}

// tc_mode 1: 3xTF32 (fp32-accurate), 0: one tf32 MMA per K step.  Supported head size: 48.
int launch_attention_tc(const float* qkv, const int32_t* key_lens, float* ctx, int B, int L, int H, int heads, int tc_mode, cudaStream_t st) {
  EV_CHECK_ARG(B > 0 && L > 0 && heads > 0 && H % heads == 0 && H / heads == 48, "attention_tc: needs d_k = 48 (B=%d L=%d H=%d heads=%d)", B, L, H, heads);
  EV_CHECK_ARG(B <= 65535 && heads <= 65535, "attention_tc: grid too large");
  return tc_mode == 1 ? launch_atc<48, 1>(qkv, key_lens, ctx, B, L, H, heads, st) : launch_atc<48, 0>(qkv, key_lens, ctx, B, L, H, heads, st);
}

}  // namespace ev
