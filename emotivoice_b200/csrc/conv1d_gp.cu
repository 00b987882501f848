// HiFi-GAN convolutions on "granule-planar" activations (sm_100a): tcgen05.mma implicit GEMM whose A operand is fed by
// bulk async copies (cp.async.bulk -> mbarrier complete_tx) instead of a register round trip, and whose epilogue stores
// straight from the TMEM lane layout with fully coalesced 16-byte accesses -- no shared-memory transpose on either side.
//
// Layout (GP): an activation tensor (B, L, C) is stored as [b][g = C / cpg][l][cpg] with 16-byte granules
// (cpg = 4 fp32 or 8 bf16 channels): every granule is a contiguous plane of L x 16 bytes.  That is exactly the no-swizzle
// K-major UMMA operand layout of conv1d_tc.cu (element (row r, granule g) at (g * rows_pad + r) * 16), so
//   * the A stage of a tile = KBG contiguous runs of `rows x 16 B`  -> KBG bulk copies issued by ONE thread, as many stages
//     in flight as shared memory holds (the round-1 kernel had six warps doing ldg -> cvt -> st.shared, 86 % no-eligible);
//   * tap j of a dilated convolution is still the same staged tile with the descriptor start advanced by j*dil rows;
//   * tcgen05.ld gives thread = row, registers = consecutive channels: 4 (fp32) / 8 (bf16) consecutive registers are one
//     granule, consecutive lanes are consecutive rows -> each st.global.v4 of a warp covers 512 contiguous bytes.
// What still has to touch the operand between the copy and the MMA -- LeakyReLU of the input, zeroing of rows outside
// [0, len) (the convolution's zero padding and the batch-invariant contract), round-to-nearest tf32 and the hi/lo split of
// the 3xTF32 fp32 emulation -- is an IN-PLACE pass over the staged tile by four "transform" warps: shared memory to shared
// memory (29-cycle latency instead of a global-memory round trip), conflict-free (consecutive lanes = consecutive 16 B).
// HBM therefore holds plain activation values (fp32, or bf16 in the bf16 mode): residuals stay exact, one copy per tensor.
//
// Roles (480 threads): warps 0-7 epilogue, 8-11 transform, 12 A loader (one lane), 13 weight loader (one lane),
// 14 TMEM allocation + MMA issue (one lane).  mbarrier pipelines: a_full (copy landed) -> a_ready (transformed) -> a_empty
// (tcgen05.commit), b_full / b_empty, acc_full / acc_empty (two accumulator sets in TMEM: epilogue of tile i overlaps the
// main loop of tile i+1).  Persistent CTAs, static round-robin tile order.
//
// out[b, m*rate + n / CoutR, n % CoutR] = epi( bias[n] + sum_j sum_ci w[j][ci][n] * act_in(x[b, m + (j-(K-1)/2)*dil, ci]) )
// rate > 1 is the polyphase form of ConvTranspose1d (packing.polyphase_pack): the GEMM's N = rate * CoutR columns are the
// `rate` output phases of each input row.  Replaces hifigan/models.py:50-57 (ResBlock1 convs), :116 (conv_pre), :118-119 (ups).
#include "ev_common.cuh"
#include "tc_common.cuh"

namespace ev {
namespace gp {

using namespace tc;

constexpr int NEPI_WARPS = 8;
constexpr int NTW = 4;                       // transform warps
constexpr int W_XFORM = NEPI_WARPS;          // warps 8..11
constexpr int W_ALOAD = W_XFORM + NTW;       // 12
constexpr int W_BLOAD = W_ALOAD + 1;         // 13
constexpr int W_MMA = W_BLOAD + 1;           // 14
constexpr int GP_THREADS = (W_MMA + 1) * 32; // 480
constexpr int MAX_A = 8, MAX_B = 8;
constexpr int SMEM_HEAD = 1024;              // barriers + TMEM slot
constexpr int XF_UNROLL = 4;

struct GPlan {
  int BN, mt, kbg, planes;
  int rows_pad;
  int a_plane_bytes, b_plane_bytes, a_stage_bytes, b_stage_bytes;
  int a_stages, b_stages;
  int tmem_cols;
  int tiles_m, tiles_n, total_tiles;
  int smem_total;
};

// span = (K-1)*dil of the widest member and kmin = the fewest taps of a grouped launch (ng convolutions); -1 / 0: p's own
__host__ __device__ inline bool make_gplan(const GpConvParams& p, int mode, int BN, int mt, int kbg, GPlan* o, int span = -1, int kmin = 0, int ng = 1) {
  if (span < 0) span = (p.K - 1) * p.dil;
  if (kmin <= 0) kmin = p.K;
  GPlan q;
  q.planes = mode == 1 ? 2 : 1;          // A planes in shared memory (mode 3 keeps hi / lo interleaved in ONE plane, in place)
  const int b_planes = (mode == 1 || mode == 3) ? 2 : 1;
  const int cpg = mode == 2 ? 8 : 4;     // channels per 16-byte granule of the ACTIVATIONS
  const int wcpg = mode >= 2 ? 8 : 4;    // channels per 16-byte granule of the WEIGHTS (bf16 operands: 8)
  q.kbg = kbg; q.mt = mt; q.BN = BN;
  if (2 * mt * BN > 512) return false;
  q.tmem_cols = 32;
  while (q.tmem_cols < 2 * mt * BN) q.tmem_cols <<= 1;
  const int rows = BM * mt + span;
  q.rows_pad = (rows + 7) / 8 * 8;
  q.a_plane_bytes = kbg * q.rows_pad * 16;
  q.b_plane_bytes = (kbg * cpg / wcpg) * BN * 16;
  q.a_stage_bytes = q.planes * q.a_plane_bytes;
  q.b_stage_bytes = b_planes * q.b_plane_bytes;
  const int budget = 227 * 1024 - SMEM_HEAD;
  const int n_cb = (p.Cin + cpg * kbg - 1) / (cpg * kbg);
  const int b_max = n_cb * kmin < MAX_B ? n_cb * kmin : MAX_B;
  q.a_stages = 2;
  q.b_stages = b_max < 2 ? b_max : 2;
  if (q.a_stages * q.a_stage_bytes + q.b_stages * q.b_stage_bytes > budget) return false;
  auto fits = [&](int a, int b) { return a * q.a_stage_bytes + b * q.b_stage_bytes <= budget; };
  // the weight ring turns over K times per activation stage: first 4 weight stages, then up to 4 activation stages (they
  // prefetch ACROSS tiles: the ring is not bounded by the channel blocks of one tile), then whatever still fits
  while (q.b_stages < b_max && q.b_stages < 4 && fits(q.a_stages, q.b_stages + 1)) ++q.b_stages;
  while (q.a_stages < 4 && fits(q.a_stages + 1, q.b_stages)) ++q.a_stages;
  while (q.b_stages < b_max && fits(q.a_stages, q.b_stages + 1)) ++q.b_stages;
  while (q.a_stages < MAX_A && fits(q.a_stages + 1, q.b_stages)) ++q.a_stages;
  q.tiles_m = (p.L + BM * mt - 1) / (BM * mt);
  q.tiles_n = (p.Cout + BN - 1) / BN;
  q.total_tiles = ng * p.B * q.tiles_m * q.tiles_n;
  q.smem_total = SMEM_HEAD + q.a_stages * q.a_stage_bytes + q.b_stages * q.b_stage_bytes;
  *o = q;
  return true;
}

// LeakyReLU for 0 <= slope <= 1 as max(v, v*slope): two instructions (FMUL + FMNMX) instead of compare / multiply / select; same bits
__device__ __forceinline__ float lrelu_f(float v, float slope) { return fmaxf(v, v * slope); }

// MODE 0: one tf32 MMA per K step; 1: 3xTF32 fp32 emulation (hi/lo planes, three MMAs per K step); 2: bf16 operands,
// bf16 activations in HBM (kind::f16, 8 channels per granule); 3: "bf16x3": fp32 activations in HBM, every operand split
// into bf16 hi + lo (16 significant bits), three kind::f16 MMAs per K = 16 step -- an fp32-class result (~1e-5 relative) at
// half the tensor-core and shared-memory cost of 3xTF32.  Accumulation is fp32 in TMEM in every mode.
template <int MODE, int MT, int KBG>
__global__ void __launch_bounds__(GP_THREADS, 1) conv1d_gp_kernel(const __grid_constant__ GpConvParams p, const __grid_constant__ GPlan pl,
                                                                  const __grid_constant__ GpGroups gs) {
  constexpr bool SPLIT3 = (MODE == 1);     // 3xTF32: hi / lo tf32 planes
  constexpr bool BF16 = (MODE == 2);       // bf16 activations in HBM, bf16 operands
  constexpr bool X3B = (MODE == 3);        // fp32 activations in HBM, operands split into bf16 hi + lo: three kind::f16 MMAs per K=16 step
  constexpr bool OP16 = BF16 || X3B;       // the MMA operands are bf16
  constexpr int BPLANES = (SPLIT3 || X3B) ? 2 : 1;
  constexpr int CPG = BF16 ? 8 : 4;        // channels per 16-byte granule of the activations (HBM and the staged tile)
  constexpr int WCPG = OP16 ? 8 : 4;       // channels per 16-byte granule of the weights
  constexpr int KB = CPG * KBG;
  static_assert(!X3B || KBG % 4 == 0, "bf16x3 consumes four fp32 granules (16 channels) per MMA K step");
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const int BN = pl.BN;

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem_raw + 512);     // barriers occupy [0, 8 * 44) = 352 B
  uint8_t* a_tiles = smem_raw + SMEM_HEAD;
  uint8_t* b_tiles = a_tiles + pl.a_stages * pl.a_stage_bytes;
  const uint32_t bar_base = smem_u32(bars);
  auto a_full = [&](int s) { return bar_base + 8u * s; };
  auto a_ready = [&](int s) { return bar_base + 8u * (MAX_A + s); };
  auto a_empty = [&](int s) { return bar_base + 8u * (2 * MAX_A + s); };
  auto b_full = [&](int s) { return bar_base + 8u * (3 * MAX_A + s); };
  auto b_empty = [&](int s) { return bar_base + 8u * (3 * MAX_A + MAX_B + s); };
  auto acc_full = [&](int s) { return bar_base + 8u * (3 * MAX_A + 2 * MAX_B + s); };
  auto acc_empty = [&](int s) { return bar_base + 8u * (3 * MAX_A + 2 * MAX_B + 2 + s); };

  if (tid == 0) {
    for (int s = 0; s < pl.a_stages; ++s) { mbar_init(a_full(s), 1); mbar_init(a_ready(s), NTW * 32); mbar_init(a_empty(s), 1); }
    for (int s = 0; s < pl.b_stages; ++s) { mbar_init(b_full(s), 1); mbar_init(b_empty(s), 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(acc_full(s), 1); mbar_init(acc_empty(s), NEPI_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == W_MMA) {   // TMEM allocation by one full warp; the same warp frees it
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(pl.tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // Programmatic dependent launch: harmless without the launch attribute.  With it, the next kernel in the stream may start
  // its set-up (barriers, TMEM, first weight stages) while this grid's tail is still running; everything that touches
  // activations executes griddepcontrol.wait first (returns once the preceding grid has completed and flushed).
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  const int n_cb = (p.Cin + KB - 1) / KB;
  const int tiles_per_b = pl.tiles_m * pl.tiles_n;
  const int tiles_per_g = p.B * tiles_per_b;  // a launch may carry up to three convolutions of one shape (different taps, dilations,
  const int gin = p.Cin / CPG;                // weights and tensors): tile -> (convolution, item, row tile, column tile)

  auto decode = [&](int tile, int& gi, int& b, int& t0, int& n0, int& len) {
    gi = tile / tiles_per_g;
    tile -= gi * tiles_per_g;
    b = tile / tiles_per_b;
    const int r = tile - b * tiles_per_b;
    const int tm = r / pl.tiles_n, tn = r - tm * pl.tiles_n;
    t0 = tm * (BM * MT);
    n0 = tn * BN;
    len = p.lens ? min(p.L, p.lens[b] * p.lens_mul) : p.L;
  };

  if (warp < NEPI_WARPS) {
    // ============================ epilogue warps ==============================================
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const int quad = warp & 3, half = warp >> 2;
    const int nchunks = BN / 32;
    const int coutR = p.Cout / p.rate;
    const int gout = coutR / CPG;              // output granule planes per item
    const size_t Lout = (size_t)p.L * p.rate;
    const int accm = p.acc;
    int tile_cnt = 0;
    for (int tile = blockIdx.x; tile < pl.total_tiles; tile += gridDim.x) {
      int gi, b, t0, n0, len;
      decode(tile, gi, b, t0, n0, len);
      if (t0 >= len) continue;                 // padding tile: no MMA work was issued, nothing is stored (rows >= len are undefined)
      const int buf = tile_cnt & 1;
      const GpGroup& G = gs.g[gi];
      const bool has_res = G.res != nullptr;
      bool waited = false;
#pragma unroll 1
      for (int item = half; item < MT * nchunks; item += 2) {
        const int mt = item / nchunks, c = (item - mt * nchunks) * 32;
        const int row = t0 + mt * BM + quad * 32 + lane;         // this thread's GEMM row (input-resolution time step)
        const bool ok = row < len;
        const int n = n0 + c;                                     // first of this thread's 32 columns
        const int phase = n / coutR, co = n - phase * coutR;
        const size_t orow = (size_t)row * p.rate + phase;
        // 16-byte granule q of this chunk lives at plane (co/CPG + q), row orow
        const size_t gbase = ((size_t)b * gout + co / CPG) * Lout + orow;
        constexpr int NG = 32 / CPG;                              // granules per 32-column chunk: 8 (fp32) / 4 (bf16)
        uint4 rq[NG];
        if (has_res) {
#pragma unroll
          for (int q = 0; q < NG; ++q) {
            rq[q] = make_uint4(0u, 0u, 0u, 0u);
            if (ok) rq[q] = *(reinterpret_cast<const uint4*>(G.res) + gbase + (size_t)q * Lout);
          }
        }
        if (!waited) {
          mbar_wait(acc_full(buf), (tile_cnt >> 1) & 1);
          tc_fence_after();
          waited = true;
        }
        float v[32];
        tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(buf * MT * BN + mt * BN + c), 32, v);
        if (ok) {
          if (G.bias) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const float4 b4 = __ldg(reinterpret_cast<const float4*>(G.bias + n) + q);
              v[4 * q] += b4.x; v[4 * q + 1] += b4.y; v[4 * q + 2] += b4.z; v[4 * q + 3] += b4.w;
            }
          }
          if (has_res) {
#pragma unroll
            for (int q = 0; q < NG; ++q) {
              if (BF16) {
                const uint32_t w4[4] = {rq[q].x, rq[q].y, rq[q].z, rq[q].w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  v[8 * q + 2 * e] += __uint_as_float(w4[e] << 16);
                  v[8 * q + 2 * e + 1] += __uint_as_float(w4[e] & 0xffff0000u);
                }
              } else {
                v[4 * q] += __uint_as_float(rq[q].x); v[4 * q + 1] += __uint_as_float(rq[q].y);
                v[4 * q + 2] += __uint_as_float(rq[q].z); v[4 * q + 3] += __uint_as_float(rq[q].w);
              }
            }
          }
          if (accm != EV_ACC_STORE) {     // the xs += / xs /= n accumulation of the last layer of a ResBlock (2 launches in 18): loaded late to keep
            uint4 oq[NG];                 // the common path's register footprint small
#pragma unroll
            for (int q = 0; q < NG; ++q) oq[q] = *(reinterpret_cast<const uint4*>(G.out) + gbase + (size_t)q * Lout);
#pragma unroll
            for (int q = 0; q < NG; ++q) {
              if (BF16) {
                const uint32_t w4[4] = {oq[q].x, oq[q].y, oq[q].z, oq[q].w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  v[8 * q + 2 * e] += __uint_as_float(w4[e] << 16);
                  v[8 * q + 2 * e + 1] += __uint_as_float(w4[e] & 0xffff0000u);
                }
              } else {
                v[4 * q] += __uint_as_float(oq[q].x); v[4 * q + 1] += __uint_as_float(oq[q].y);
                v[4 * q + 2] += __uint_as_float(oq[q].z); v[4 * q + 3] += __uint_as_float(oq[q].w);
              }
            }
            if (accm == EV_ACC_ADD_DIV) {
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] /= p.div;
            }
          }
#pragma unroll
          for (int q = 0; q < NG; ++q) {
            uint4 o;
            if (BF16) {
              o.x = pack_bf16(v[8 * q], v[8 * q + 1]); o.y = pack_bf16(v[8 * q + 2], v[8 * q + 3]);
              o.z = pack_bf16(v[8 * q + 4], v[8 * q + 5]); o.w = pack_bf16(v[8 * q + 6], v[8 * q + 7]);
            } else {
              o.x = __float_as_uint(v[4 * q]); o.y = __float_as_uint(v[4 * q + 1]);
              o.z = __float_as_uint(v[4 * q + 2]); o.w = __float_as_uint(v[4 * q + 3]);
            }
            *(reinterpret_cast<uint4*>(G.out) + gbase + (size_t)q * Lout) = o;
          }
        }
      }
      if (!waited) {     // a warp without work items in this tile still follows the accumulator phases
        mbar_wait(acc_full(buf), (tile_cnt >> 1) & 1);
        tc_fence_after();
      }
      // all TMEM reads of this buffer are complete (tcgen05.wait::ld inside tmem_ld32): hand it back
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty(buf));
      ++tile_cnt;
    }
  } else if (warp < W_ALOAD) {
    // ============================ transform warps: in-place pass over the landed A stage =======================
    const int xt = (warp - W_XFORM) * 32 + lane;      // 0..127
    const bool lrelu = (p.in_act == EV_ACT_LRELU);
    const float slope = p.in_slope;
    int a_cnt = 0;
    for (int tile = blockIdx.x; tile < pl.total_tiles; tile += gridDim.x) {
      int gi, b, t0, n0, len;
      decode(tile, gi, b, t0, n0, len);
      if (t0 >= len) continue;
      const int span = (gs.g[gi].K - 1) * gs.g[gi].dil;
      const int rows_a = BM * MT + span;
      const int row0 = t0 - span / 2;
      for (int cb = 0; cb < n_cb; ++cb, ++a_cnt) {
        const int s = a_cnt % pl.a_stages;
        const int ngran = min(KB, p.Cin - cb * KB) / CPG;
        uint8_t* base = a_tiles + s * pl.a_stage_bytes;
        mbar_wait(a_full(s), (a_cnt / pl.a_stages) & 1);
        if (X3B) {
          // fp32 -> bf16 hi + lo in place: the granule pair (2q, 2q+1) = 8 channels becomes [hi of the 8 | lo of the 8], so the hi
          // plane is the even granule slots and the lo plane the odd ones (descriptor LBO = two slots)
          for (int q = 0; q < ngran / 2; ++q) {
            uint8_t* g0 = base + (size_t)(2 * q) * pl.rows_pad * 16;
            uint8_t* g1 = g0 + (size_t)pl.rows_pad * 16;
            for (int r = xt; r < rows_a; r += NTW * 32) {
              const int row = row0 + r;
              float4 u = make_float4(0.f, 0.f, 0.f, 0.f), w = u;
              if (row >= 0 && row < len) { u = *reinterpret_cast<const float4*>(g0 + r * 16); w = *reinterpret_cast<const float4*>(g1 + r * 16); }
              float f[8] = {u.x, u.y, u.z, u.w, w.x, w.y, w.z, w.w};
              uint32_t hi[4], lo[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                float a0 = f[2 * e], a1 = f[2 * e + 1];
                if (lrelu) { a0 = lrelu_f(a0, slope); a1 = lrelu_f(a1, slope); }
                hi[e] = pack_bf16(a0, a1);
                lo[e] = pack_bf16(a0 - __uint_as_float(hi[e] << 16), a1 - __uint_as_float(hi[e] & 0xffff0000u));
              }
              *reinterpret_cast<uint4*>(g0 + r * 16) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
              *reinterpret_cast<uint4*>(g1 + r * 16) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
            }
          }
        }
        for (int g = 0; g < (X3B ? 0 : ngran); ++g) {
          uint8_t* gb = base + (size_t)g * pl.rows_pad * 16;
          for (int r0 = 0; r0 < rows_a; r0 += NTW * 32 * XF_UNROLL) {
            uint4 v[XF_UNROLL];
#pragma unroll
            for (int u = 0; u < XF_UNROLL; ++u) {
              const int r = r0 + u * (NTW * 32) + xt;
              const int row = row0 + r;
              v[u] = make_uint4(0u, 0u, 0u, 0u);
              if (r < rows_a && row >= 0 && row < len) v[u] = *reinterpret_cast<const uint4*>(gb + r * 16);
            }
#pragma unroll
            for (int u = 0; u < XF_UNROLL; ++u) {
              const int r = r0 + u * (NTW * 32) + xt;
              if (r >= rows_a) continue;
              if (BF16) {
                if (lrelu) {
                  uint32_t w4[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
                  for (int e = 0; e < 4; ++e) {
                    const float lo = lrelu_f(__uint_as_float(w4[e] << 16), slope);
                    const float hi = lrelu_f(__uint_as_float(w4[e] & 0xffff0000u), slope);
                    w4[e] = pack_bf16(lo, hi);
                  }
                  v[u] = make_uint4(w4[0], w4[1], w4[2], w4[3]);
                }
                *reinterpret_cast<uint4*>(gb + r * 16) = v[u];
              } else {
                float4 t = make_float4(__uint_as_float(v[u].x), __uint_as_float(v[u].y), __uint_as_float(v[u].z), __uint_as_float(v[u].w));
                if (lrelu) { t.x = lrelu_f(t.x, slope); t.y = lrelu_f(t.y, slope); t.z = lrelu_f(t.z, slope); t.w = lrelu_f(t.w, slope); }
                // round to nearest tf32 (the MMA would otherwise truncate the low 13 mantissa bits)
                const float4 h = make_float4(to_tf32(t.x), to_tf32(t.y), to_tf32(t.z), to_tf32(t.w));
                *reinterpret_cast<float4*>(gb + r * 16) = h;
                if (SPLIT3) {
                  const float4 l = make_float4(to_tf32(t.x - h.x), to_tf32(t.y - h.y), to_tf32(t.z - h.z), to_tf32(t.w - h.w));
                  *reinterpret_cast<float4*>(gb + pl.a_plane_bytes + r * 16) = l;
                }
              }
            }
          }
        }
        fence_proxy_async();      // generic-proxy smem writes -> visible to the tensor core (async proxy)
        mbar_arrive(a_ready(s));
      }
    }
  } else if (warp == W_ALOAD) {
    // ============================ A loader: one thread, KBG bulk copies per stage ================================
    if (lane == 0) {
      asm volatile("griddepcontrol.wait;" ::: "memory");
      int a_cnt = 0;
      for (int tile = blockIdx.x; tile < pl.total_tiles; tile += gridDim.x) {
        int gi, b, t0, n0, len;
        decode(tile, gi, b, t0, n0, len);
        if (t0 >= len) continue;
        const int span = (gs.g[gi].K - 1) * gs.g[gi].dil;
        const int halo = span / 2, rows_a = BM * MT + span;
        const int r_lo = max(t0 - halo, 0);
        const int r_hi = min(t0 - halo + rows_a, len);      // len <= L: never past the plane
        const uint32_t nbytes = (uint32_t)(r_hi - r_lo) * 16u;
        const uint32_t roff = (uint32_t)(r_lo - (t0 - halo)) * 16u;
        const uint8_t* xb = reinterpret_cast<const uint8_t*>(gs.g[gi].x) + ((size_t)b * gin * p.L + r_lo) * 16;
        for (int cb = 0; cb < n_cb; ++cb, ++a_cnt) {
          const int s = a_cnt % pl.a_stages;
          const int ngran = min(KB, p.Cin - cb * KB) / CPG;
          mbar_wait(a_empty(s), ((a_cnt / pl.a_stages) & 1) ^ 1);
          mbar_expect_tx(a_full(s), (uint32_t)ngran * nbytes);
          const uint32_t dst = smem_u32(a_tiles + s * pl.a_stage_bytes) + roff;
          const uint8_t* src = xb + (size_t)(cb * KBG) * p.L * 16;
          for (int g = 0; g < ngran; ++g)
            bulk_g2s(dst + (uint32_t)(g * pl.rows_pad * 16), src + (size_t)g * p.L * 16, nbytes, a_full(s));
        }
      }
    }
    __syncwarp();
  } else if (warp == W_BLOAD) {
    // ============================ weight loader (weights are constants: no dependency wait) ========================
    if (lane == 0) {
      // w layout: [plane (hi, lo)][N tile of BNp = min(Cout,128)][tap][Cin/WCPG granules][BNp][16 bytes] (fp32 or bf16 granules)
      const int bnp = p.Cout < 128 ? p.Cout : 128;
      const int win = p.Cin / WCPG;                                  // weight granules along C_in
      constexpr int KBGW = KBG * CPG / WCPG;                         // weight granules per pipeline stage
      int b_cnt = 0;
      for (int tile = blockIdx.x; tile < pl.total_tiles; tile += gridDim.x) {
        int gi, b, t0, n0, len;
        decode(tile, gi, b, t0, n0, len);
        if (t0 >= len) continue;
        const int K = gs.g[gi].K;
        const size_t plane = (size_t)K * win * p.Cout * 4;          // 4-byte words per plane
        const size_t tile_stride = (size_t)K * win * bnp * 4;       // 4-byte words per packed N tile
        const float* wt = gs.g[gi].w + (size_t)(n0 / bnp) * tile_stride + (size_t)(n0 % bnp) * 4;
        for (int cb = 0; cb < n_cb; ++cb) {
          const int ngran = min(KB, p.Cin - cb * KB) / WCPG;          // weight granules of this stage
          for (int j = 0; j < K; ++j, ++b_cnt) {
            const int sb = b_cnt % pl.b_stages;
            mbar_wait(b_empty(sb), ((b_cnt / pl.b_stages) & 1) ^ 1);
            mbar_expect_tx(b_full(sb), (uint32_t)(BPLANES * ngran * BN * 16));
            const uint32_t dst = smem_u32(b_tiles + sb * pl.b_stage_bytes);
            const float* src = wt + ((size_t)j * win + (size_t)cb * KBGW) * bnp * 4;
            if (BN == bnp) {
              bulk_g2s(dst, src, (uint32_t)(ngran * BN * 16), b_full(sb));
              if (BPLANES == 2) bulk_g2s(dst + (uint32_t)pl.b_plane_bytes, src + plane, (uint32_t)(ngran * BN * 16), b_full(sb));
            } else {
              for (int g = 0; g < ngran; ++g) {
                bulk_g2s(dst + (uint32_t)(g * BN * 16), src + (size_t)g * bnp * 4, (uint32_t)(BN * 16), b_full(sb));
                if (BPLANES == 2) bulk_g2s(dst + (uint32_t)(pl.b_plane_bytes + g * BN * 16), src + plane + (size_t)g * bnp * 4, (uint32_t)(BN * 16), b_full(sb));
              }
            }
          }
        }
      }
    }
    __syncwarp();
  } else {
    // ============================ MMA issuer =====================================================
    // All 32 lanes run the (warp-uniform) control flow and the barrier waits; one elected lane issues the tcgen05 instructions.
    {
      const uint32_t a_lbo = (uint32_t)pl.rows_pad * 16u, b_lbo = (uint32_t)BN * 16u;
      const uint32_t fmt = OP16 ? 1u : 2u;      // instruction descriptor: D=F32 [4,6)=1, A/B format [7,10) / [10,13), N>>3 [17,23), M>>4 [24,29)
      const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      // bf16x3: the two K granules of one MMA are two slots apart (hi in the even slots, lo in the odd ones), one K step = 4 slots
      const uint64_t a_desc0 = make_desc(0u, X3B ? 2u * a_lbo : a_lbo, 128u), b_desc0 = make_desc(0u, b_lbo, 128u);
      const uint32_t a_k8 = (X3B ? 4u : 2u) * a_lbo, b_k8 = 2u * b_lbo;   // bytes per K step
      const uint32_t a_lo_off = X3B ? a_lbo : (uint32_t)pl.a_plane_bytes;
      int a_cnt = 0, b_cnt = 0, tile_cnt = 0;
      for (int tile = blockIdx.x; tile < pl.total_tiles; tile += gridDim.x) {
        int gi, b, t0, n0, len;
        decode(tile, gi, b, t0, n0, len);
        if (t0 >= len) continue;
        const int K = gs.g[gi].K;
        const uint32_t a_tap = (uint32_t)gs.g[gi].dil * 16u;          // bytes per tap shift
        const int buf = tile_cnt & 1;
        mbar_wait(acc_empty(buf), ((tile_cnt >> 1) & 1) ^ 1);     // epilogue has drained this accumulator set
        tc_fence_after();
        const uint32_t d_base = tmem_base + (uint32_t)(buf * MT * BN);
        for (int cb = 0; cb < n_cb; ++cb, ++a_cnt) {
          const int sa = a_cnt % pl.a_stages;
          const int nk8 = min(KB, p.Cin - cb * KB) / (2 * WCPG);  // MMA K steps: two 16-byte operand granules each
          mbar_wait(a_ready(sa), (a_cnt / pl.a_stages) & 1);
          const uint64_t a_hi0 = desc_advance(a_desc0, smem_u32(a_tiles + sa * pl.a_stage_bytes));
          for (int j = 0; j < K; ++j, ++b_cnt) {
            const int sb = b_cnt % pl.b_stages;
            mbar_wait(b_full(sb), (b_cnt / pl.b_stages) & 1);
            tc_fence_after();
            const uint64_t b_hi0 = desc_advance(b_desc0, smem_u32(b_tiles + sb * pl.b_stage_bytes));
            const uint64_t a_j = desc_advance(a_hi0, (uint32_t)j * a_tap);
            if (elect_one()) {
              for (int k8 = 0; k8 < nk8; ++k8) {
                const uint64_t b_hi = desc_advance(b_hi0, (uint32_t)k8 * b_k8);
                const uint64_t a_k = desc_advance(a_j, (uint32_t)k8 * a_k8);
                const uint32_t first = (cb | j | k8) != 0 ? 1u : 0u;
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) {      // one weight tile feeds MT accumulators
                  const uint64_t a_hi = desc_advance(a_k, (uint32_t)(mt * BM) * 16u);
                  const uint32_t d = d_base + (uint32_t)(mt * BN);
                  if (SPLIT3 || X3B) {
                    const uint64_t a_lo = desc_advance(a_hi, a_lo_off);
                    const uint64_t b_lo = desc_advance(b_hi, (uint32_t)pl.b_plane_bytes);
                    if (X3B) {
                      umma_bf16(d, a_lo, b_hi, idesc, first);     // small terms first
                      umma_bf16(d, a_hi, b_lo, idesc, 1u);
                      umma_bf16(d, a_hi, b_hi, idesc, 1u);
                    } else {
                      umma_tf32(d, a_lo, b_hi, idesc, first);
                      umma_tf32(d, a_hi, b_lo, idesc, 1u);
                      umma_tf32(d, a_hi, b_hi, idesc, 1u);
                    }
                  } else if (BF16) {
                    umma_bf16(d, a_hi, b_hi, idesc, first);
                  } else {
                    umma_tf32(d, a_hi, b_hi, idesc, first);
                  }
                }
              }
              umma_commit(b_empty(sb));               // weight stage free once these MMAs have read it
              if (j == K - 1) {
                umma_commit(a_empty(sa));             // activation stage free
                if (cb == n_cb - 1) umma_commit(acc_full(buf));   // accumulators of this tile complete -> epilogue
              }
            }
            __syncwarp();
          }
        }
        ++tile_cnt;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == W_MMA) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(pl.tmem_cols));
  }
}

// ---- layout conversion at the vocoder's boundary ---------------------------------------------------------------------
// in[b*sb + t*st + c*sc] fp32 (time-major (B,F,C): st = C, sc = 1; channels-first (B,C,F): st = 1, sc = F)  ->  GP.
template <bool BF16>
__global__ void __launch_bounds__(256) to_gp_kernel(const float* __restrict__ in, long long sb, long long st_, long long sc, void* __restrict__ out,
                                                    int B, int L, int C) {
  asm volatile("griddepcontrol.launch_dependents;\n\tgriddepcontrol.wait;" ::: "memory");
  constexpr int CPG = BF16 ? 8 : 4;
  const int G = C / CPG;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;      // (b, g, t), t fastest
  if (i >= (size_t)B * G * L) return;
  const int t = (int)(i % L);
  const int g = (int)((i / L) % G);
  const int b = (int)(i / ((size_t)L * G));
  const float* src = in + (size_t)b * sb + (size_t)t * st_ + (size_t)(g * CPG) * sc;
  float v[CPG];
#pragma unroll
  for (int e = 0; e < CPG; ++e) v[e] = src[(size_t)e * sc];
  uint4 o;
  if (BF16) {
    o.x = pack_bf16(v[0], v[1]); o.y = pack_bf16(v[2], v[3]); o.z = pack_bf16(v[CPG - 4], v[CPG - 3]); o.w = pack_bf16(v[CPG - 2], v[CPG - 1]);
  } else {
    o.x = __float_as_uint(v[0]); o.y = __float_as_uint(v[1]); o.z = __float_as_uint(v[2]); o.w = __float_as_uint(v[3]);
  }
  reinterpret_cast<uint4*>(out)[i] = o;
}

// out = ((b + a) [+ c]) / div over whole fp32 granule-planar tensors: the `xs += ...; x = xs / n` of a HiFi-GAN stage
// (hifigan/models.py:120-126) when the three ResBlocks' last layers ran as one grouped launch into their own tensors.  Same additions
// in the same order as the accumulate modes of the epilogue (a stored, then b + a, then c + that, then / div): identical bits.
template <int N>
__global__ void __launch_bounds__(256) gp_sum_div_kernel(const float4* __restrict__ a, const float4* __restrict__ b, const float4* __restrict__ c,
                                                         float4* __restrict__ out, size_t n4, float div) {
  asm volatile("griddepcontrol.launch_dependents;\n\tgriddepcontrol.wait;" ::: "memory");
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const float4 va = a[i], vb = b[i];
  float4 t = make_float4(vb.x + va.x, vb.y + va.y, vb.z + va.z, vb.w + va.w);
  if (N == 3) {
    const float4 vc = c[i];
    t = make_float4(vc.x + t.x, vc.y + t.y, vc.z + t.z, vc.w + t.w);
  }
  out[i] = make_float4(t.x / div, t.y / div, t.z / div, t.w / div);
}

// wav[b,t] = tanh( bias + sum_j sum_c w[j][c] * lrelu(x[b, t+j-(K-1)/2, c]) ) on a GP input (hifigan/models.py:127-129:
// F.leaky_relu default slope 0.01, Conv1d(C,1,7,pad 3), tanh).  HBM-bound: each CTA stages (256 + K - 1) rows once, already
// activated; rows >= len read as zero padding and are written as zeros.  Same summation order as conv_post_kernel.  Generic shapes;
// HiFi-GAN's own (K = 7, C <= 48) run conv_post_gp4_kernel below.
constexpr int GPP_BT = 256;
template <bool BF16>
__global__ void __launch_bounds__(GPP_BT) conv_post_gp_kernel(const void* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                                                              const int32_t* __restrict__ lens, int lens_mul, int L, int C, int K, float slope,
                                                              float* __restrict__ wav) {
  asm volatile("griddepcontrol.launch_dependents;\n\tgriddepcontrol.wait;" ::: "memory");
  constexpr int CPG = BF16 ? 8 : 4;
  extern __shared__ __align__(16) float gpp_smem[];
  const int rows = GPP_BT + K - 1;
  float* xs = gpp_smem;                 // [C][rows] channel-major: conflict-free for consecutive rows
  float* ws = gpp_smem + (size_t)C * rows;   // [K][C]
  const int b = blockIdx.y, t0 = blockIdx.x * GPP_BT;
  const int len = lens ? min(L, lens[b] * lens_mul) : L;
  const int halo = (K - 1) / 2;
  const int G = C / CPG;
  const uint4* xb = reinterpret_cast<const uint4*>(x) + (size_t)b * G * L;
  for (int i = threadIdx.x; i < G * rows; i += GPP_BT) {
    const int g = i / rows, r = i - g * rows;
    const int row = t0 - halo + r;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (row >= 0 && row < len) v = xb[(size_t)g * L + row];
    float f[CPG];
    if (BF16) {
      const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) { f[2 * e] = __uint_as_float(w4[e] << 16); f[2 * e + 1] = __uint_as_float(w4[e] & 0xffff0000u); }
    } else {
      f[0] = __uint_as_float(v.x); f[1] = __uint_as_float(v.y); f[2] = __uint_as_float(v.z); f[3] = __uint_as_float(v.w);
    }
#pragma unroll
    for (int e = 0; e < CPG; ++e) xs[(size_t)(g * CPG + e) * rows + r] = lrelu_f(f[e], slope);
  }
  for (int i = threadIdx.x; i < K * C; i += GPP_BT) ws[i] = w[i];
  __syncthreads();
  const int t = t0 + threadIdx.x;
  if (t >= L) return;
  float acc = 0.f;
  for (int c = 0; c < C; ++c) {       // channel-major reduction order (all three conv_post kernels share it)
    const float* xr = xs + (size_t)c * rows + threadIdx.x;
    for (int j = 0; j < K; ++j) acc = fmaf(xr[j], ws[j * C + c], acc);
  }
  wav[(size_t)b * L + t] = t < len ? tanhf(acc + bias[0]) : 0.f;
}

// The same operator for the shapes HiFi-GAN has (K = 7, C = 32): 128 threads x 4 consecutive outputs.  Per channel a thread reads its
// 4 + K - 1 inputs (three LDS.128) and the K taps (two broadcast LDS.128) for 4 K FMAs -- the one-output kernel above spends two
// shared-memory loads per FMA and measured 14x over its HBM time.  Tiles past the item's length only store zeros.  Same sums, same order.
constexpr int GP4_T = 128, GP4_O = 4, GP4_ROWS = GP4_T * GP4_O;
template <bool BF16, int K>
__global__ void __launch_bounds__(GP4_T) conv_post_gp4_kernel(const void* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                                                              const int32_t* __restrict__ lens, int lens_mul, int L, int C, float slope,
                                                              float* __restrict__ wav) {
  asm volatile("griddepcontrol.launch_dependents;\n\tgriddepcontrol.wait;" ::: "memory");
  constexpr int CPG = BF16 ? 8 : 4;
  constexpr int HALO = (K - 1) / 2;
  constexpr int NW = (GP4_O + K - 1 + 3) / 4;          // float4 loads per input window
  constexpr int RS = GP4_ROWS + (NW - 1) * 4;          // staged rows per channel (the last thread's window ends at 4*127 + 4*NW)
  constexpr int KP = (K + 3) / 4 * 4;
  extern __shared__ __align__(16) float gpp_smem[];
  float* xs = gpp_smem;                       // [C][RS]
  float* ws = gpp_smem + (size_t)C * RS;      // [C][KP]
  const int b = blockIdx.y, t0 = blockIdx.x * GP4_ROWS;
  const int len = lens ? min(L, lens[b] * lens_mul) : L;
  const int t = t0 + GP4_O * threadIdx.x;
  float* out = wav + (size_t)b * L + t;
  const bool vec = (L & 3) == 0 && t + GP4_O <= L;
  if (t0 >= len) {                            // padding tile
    if (vec) *reinterpret_cast<float4*>(out) = make_float4(0.f, 0.f, 0.f, 0.f);
    else for (int o = 0; o < GP4_O; ++o) if (t + o < L) out[o] = 0.f;
    return;
  }
  const int G = C / CPG;
  const uint4* xb = reinterpret_cast<const uint4*>(x) + (size_t)b * G * L;
#pragma unroll 4
  for (int i = threadIdx.x; i < G * RS; i += GP4_T) {
    const int g = i / RS, r = i - g * RS;
    const int row = t0 - HALO + r;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (row >= 0 && row < len) v = __ldg(xb + (size_t)g * L + row);
    float f[CPG];
    if (BF16) {
      const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) { f[2 * e] = __uint_as_float(w4[e] << 16); f[2 * e + 1] = __uint_as_float(w4[e] & 0xffff0000u); }
    } else {
      f[0] = __uint_as_float(v.x); f[1] = __uint_as_float(v.y); f[2] = __uint_as_float(v.z); f[3] = __uint_as_float(v.w);
    }
#pragma unroll
    for (int e = 0; e < CPG; ++e) xs[(size_t)(g * CPG + e) * RS + r] = lrelu_f(f[e], slope);
  }
  for (int i = threadIdx.x; i < C * KP; i += GP4_T) {
    const int c = i / KP, j = i - c * KP;
    ws[i] = j < K ? w[j * C + c] : 0.f;
  }
  __syncthreads();
  float acc[GP4_O];
#pragma unroll
  for (int o = 0; o < GP4_O; ++o) acc[o] = 0.f;
  for (int c = 0; c < C; ++c) {
    float xw[NW * 4], wv[KP];
    const float4* xr = reinterpret_cast<const float4*>(xs + (size_t)c * RS + GP4_O * threadIdx.x);
    const float4* wr = reinterpret_cast<const float4*>(ws + c * KP);
#pragma unroll
    for (int q = 0; q < NW; ++q) { const float4 v = xr[q]; xw[4 * q] = v.x; xw[4 * q + 1] = v.y; xw[4 * q + 2] = v.z; xw[4 * q + 3] = v.w; }
#pragma unroll
    for (int q = 0; q < KP / 4; ++q) { const float4 v = wr[q]; wv[4 * q] = v.x; wv[4 * q + 1] = v.y; wv[4 * q + 2] = v.z; wv[4 * q + 3] = v.w; }
#pragma unroll
    for (int j = 0; j < K; ++j)
#pragma unroll
      for (int o = 0; o < GP4_O; ++o) acc[o] = fmaf(xw[o + j], wv[j], acc[o]);
  }
  const float b0 = bias[0];
  float r[GP4_O];
#pragma unroll
  for (int o = 0; o < GP4_O; ++o) r[o] = t + o < len ? tanhf(acc[o] + b0) : 0.f;
  if (vec) *reinterpret_cast<float4*>(out) = make_float4(r[0], r[1], r[2], r[3]);
  else for (int o = 0; o < GP4_O; ++o) if (t + o < L) out[o] = r[o];
}

}  // namespace gp

static int validate_gp(const GpConvParams& p, int mode) {
  EV_CHECK_ARG(p.B > 0 && p.L > 0, "conv1d_gp: bad problem B=%d L=%d", p.B, p.L);
  EV_CHECK_ARG(mode >= 0 && mode <= 3, "conv1d_gp: mode %d", mode);
  EV_CHECK_ARG(p.Cin % (mode >= 2 ? 16 : 8) == 0, "conv1d_gp: Cin=%d must be a multiple of %d", p.Cin, mode >= 2 ? 16 : 8);
  EV_CHECK_ARG(p.Cout % 32 == 0 && (p.Cout <= 128 || p.Cout % 128 == 0), "conv1d_gp: Cout=%d must be a multiple of 32, and of 128 above 128", p.Cout);
  EV_CHECK_ARG(p.rate >= 1 && p.Cout % p.rate == 0 && (p.Cout / p.rate) % 32 == 0, "conv1d_gp: rate=%d does not split Cout=%d into multiples of 32", p.rate, p.Cout);
  EV_CHECK_ARG(p.K >= 1 && (p.K & 1) && p.dil >= 1, "conv1d_gp: K=%d must be odd, dil=%d >= 1", p.K, p.dil);
  EV_CHECK_ARG(p.in_act == EV_ACT_NONE || p.in_act == EV_ACT_LRELU, "conv1d_gp: unsupported input activation");
  EV_CHECK_ARG(p.rate == 1 || (!p.res && p.acc == EV_ACC_STORE), "conv1d_gp: residual / accumulate need rate == 1");
  EV_CHECK_ARG((long long)p.L * p.rate < (1ll << 31), "conv1d_gp: output too long");
  return EV_OK;
}

// K granules per stage decide the order of the (channel block, tap, k-step) reduction, so they are a function of the layer
// shape alone (never of batch or length): 8 outside the 3xTF32 mode if a one-accumulator tile fits with them, else 4.
static int gp_shape_kbg(const GpConvParams& p, int mode) {
  gp::GPlan pl;
  const int bn_max = p.Cout <= 128 ? p.Cout : 128;
  return (mode != 1 && gp::make_gplan(p, mode, bn_max, 1, 8, &pl)) ? 8 : 4;
}

// Tile shape: none of these choices changes the order in which any output element's K reduction is summed, so results are
// bitwise independent of batch size / sequence length (batch-invariant contract) and the shape can be picked by a cost model.
// Per tile of MT x 128 rows and BN columns, three engines run concurrently and the slowest one sets the pace:
//   tensor core + its shared-memory operand reads:  MT * m * max(BN/2, 32 + BN/4) cycles per (tap, 8 channels)   (m = 3 MMAs in 3xTF32; bf16: 16 channels)
//   weight stream L2 -> shared memory:              planes * 32 B * BN per (tap, 8 channels) at ~42 B/cycle/SM (6.3 KB/cycle chip-wide)
//   activations HBM -> shared memory -> HBM:        MT * 128 rows * (Cin + 2 Cout) * esize at ~23 B/cycle/SM
// plus a fixed pipeline fill / drain per tile; the launch takes ceil(tiles / SMs) such tile times.  More accumulators per tile
// (MT) amortise the weight stream and the halo rows, a narrower N tile fills idle SMs (HiFi-GAN stage 1 at batch 1).
// A grouped launch (ng convolutions of one shape, ksum = the sum of their taps, span / kmin as in make_gplan) is planned like one
// convolution with the mean number of taps and ng times the tiles; kbg must be the members' own (checked by the caller).
static int plan_gp(const GpConvParams& p, int mode, gp::GPlan* out, int ng = 1, int ksum = 0, int span = -1, int kmin = 0, int kbg_forced = 0,
                   const int* gK = nullptr) {
  EV_TRY(validate_gp(p, mode));
  const int nsm = sm_count();
  const int kbg = kbg_forced ? kbg_forced : gp_shape_kbg(p, mode);
  if (ksum <= 0) ksum = p.K;
  if (span < 0) span = (p.K - 1) * p.dil;
  const double kc8 = (double)ksum / ng * p.Cin / 8.0;
  const double n_mma = ((mode == 1 || mode == 3) ? 3.0 : 1.0) * (mode >= 2 ? 0.5 : 1.0) * kc8;     // MMA instructions per accumulator and tile
  const double w_per_n = ((mode == 1 || mode == 3) ? 2.0 : 1.0) * (mode >= 2 ? 16.0 : 32.0) / 42.0;
  const double esize = mode == 2 ? 2.0 : 4.0;
  double best = 1e300;
  bool found = false;
  gp::GPlan best_pl;
  const int bn_max = p.Cout <= 128 ? p.Cout : 128;
  // N tile = the weight packing tile (min(C_out, 128)): ONE bulk copy per weight stage and plane.  Narrower tiles would fill idle SMs
  // at batch 1 (stage 1: 68 tiles) but need one 1-KB copy per granule issued by a single thread -- measured 1.5x slower.
  for (int BN = bn_max; BN >= bn_max; BN /= 2) {
    if (p.Cout % BN || BN % 32) continue;
    for (int mt = 4; mt >= 1; mt >>= 1) {
      gp::GPlan pl;
      if (!gp::make_gplan(p, mode, BN, mt, kbg, &pl, span, kmin, ng)) continue;
      // per MMA instruction: BN/2 tensor cycles, but both operands come from shared memory (128 B/cycle): (128 + BN) rows x 32 B
      // = 32 + BN/4 cycles -- the binding term below BN = 128 (a 64-wide tile costs 1.5x per FLOP, a 32-wide one 2.5x)
      double c_mma = BN / 2.0;
      if (32.0 + BN / 4.0 > c_mma) c_mma = 32.0 + BN / 4.0;
      const double t_mma = mt * n_mma * c_mma;
      const double t_w = w_per_n * BN * kc8;
      const double t_hbm = (double)mt * tc::BM * ((double)p.Cin * (mt * tc::BM + span) / (mt * tc::BM) + 2.0 * BN) * esize / 23.0;
      double t = t_mma > t_w ? t_mma : t_w;
      if (t_hbm > t) t = t_hbm;
      double cost;
      if (ng > 1 && gK) {
        // grouped launch: the members' tiles cost in proportion to their taps and are dealt round-robin in member order (heaviest
        // first); the launch takes as long as its busiest CTA
        const int tg = pl.total_tiles / ng;
        const int ncta = pl.total_tiles < nsm ? pl.total_tiles : nsm;
        cost = 0.0;
        for (int c = 0; c < ncta; ++c) {
          double sum = 0.0;
          for (int i = c; i < pl.total_tiles; i += nsm) sum += t * gK[i / tg] * ng / (double)ksum + 3000.0;
          if (sum > cost) cost = sum;
        }
      } else {
        const double waves = (double)((pl.total_tiles + nsm - 1) / nsm);
        cost = waves * (t + 3000.0);
      }
      if (cost < best * 0.97) { best = cost; best_pl = pl; found = true; }     // widest N / most accumulators first; 3 % hysteresis
    }
  }
  if (!found) { set_error("conv1d_gp: tile does not fit in shared memory / TMEM (K=%d dil=%d Cout=%d)", p.K, p.dil, p.Cout); return EV_EINVAL; }
  *out = best_pl;
  return EV_OK;
}

int gp_solo_tiles(const GpConvParams& p, int mode) {
  gp::GPlan pl;
  return plan_gp(p, mode, &pl) == EV_OK ? pl.total_tiles : 0;
}

int debug_gp_plan(const GpConvParams& p, int mode, int* v) {
  gp::GPlan pl;
  const int rc = plan_gp(p, mode, &pl);
  if (rc != EV_OK) return rc;
  v[0] = pl.BN; v[1] = pl.mt; v[2] = pl.kbg; v[3] = pl.a_stages; v[4] = pl.b_stages; v[5] = gp::NTW;
  v[6] = pl.planes; v[7] = pl.tmem_cols; v[8] = pl.smem_total; v[9] = pl.total_tiles; v[10] = pl.rows_pad;
  return EV_OK;
}

template <int MODE, int MT, int KBG>
static int launch_gp_variant(const GpConvParams& p, const gp::GPlan& pl, const GpGroups& gs, cudaStream_t st) {
  static std::atomic<uint64_t> attr_devs{0};
  if (first_use_on_device(attr_devs))
    cudaFuncSetAttribute(gp::conv1d_gp_kernel<MODE, MT, KBG>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  const int nsm = sm_count();
  const int grid = pl.total_tiles < nsm ? pl.total_tiles : nsm;
  if (pdl_mode()) {
    const cudaError_t e = launch_with_pdl(gp::conv1d_gp_kernel<MODE, MT, KBG>, dim3((unsigned)grid), dim3(gp::GP_THREADS), (size_t)pl.smem_total, st, p, pl, gs);
    if (e != cudaSuccess) { set_error("conv1d_gp_kernel (PDL launch): %s", cudaGetErrorString(e)); return EV_ECUDA; }
    count_launch();
    return EV_OK;
  }
  gp::conv1d_gp_kernel<MODE, MT, KBG><<<grid, gp::GP_THREADS, pl.smem_total, st>>>(p, pl, gs);
  EV_CUDA_LAUNCH_CHECK("conv1d_gp_kernel");
  return EV_OK;
}

template <int MODE, int KBG>
static int launch_gp_mt(const GpConvParams& p, const gp::GPlan& pl, const GpGroups& gs, cudaStream_t st) {
  if (pl.mt == 4) return launch_gp_variant<MODE, 4, KBG>(p, pl, gs, st);
  if (pl.mt == 2) return launch_gp_variant<MODE, 2, KBG>(p, pl, gs, st);
  return launch_gp_variant<MODE, 1, KBG>(p, pl, gs, st);
}

static int dispatch_gp(const GpConvParams& p, const gp::GPlan& pl, const GpGroups& gs, int mode, cudaStream_t st) {
  if (mode == 1) return launch_gp_mt<1, 4>(p, pl, gs, st);
  if (mode == 3) return pl.kbg == 8 ? launch_gp_mt<3, 8>(p, pl, gs, st) : launch_gp_mt<3, 4>(p, pl, gs, st);
  if (mode == 2) return pl.kbg == 8 ? launch_gp_mt<2, 8>(p, pl, gs, st) : launch_gp_mt<2, 4>(p, pl, gs, st);
  return pl.kbg == 8 ? launch_gp_mt<0, 8>(p, pl, gs, st) : launch_gp_mt<0, 4>(p, pl, gs, st);
}

int launch_conv1d_gp(const GpConvParams& p, int mode, cudaStream_t st) {
  gp::GPlan pl;
  EV_TRY(plan_gp(p, mode, &pl));
  GpGroups gs{};
  gs.ng = 1;
  gs.g[0] = GpGroup{p.x, p.w, p.bias, p.res, p.out, p.K, p.dil};
  return dispatch_gp(p, pl, gs, mode, st);
}

// ---- grouped launch ------------------------------------------------------------------------------------------------------
static bool group_shapes_match(const GpConvParams* ps, int n, int mode, int* kbg) {
  if (n < 1 || n > 3) return false;
  const GpConvParams& a = ps[0];
  int kb = 0;
  for (int i = 0; i < n; ++i) {
    const GpConvParams& q = ps[i];
    if (q.B != a.B || q.L != a.L || q.Cin != a.Cin || q.Cout != a.Cout || q.rate != 1 || q.lens != a.lens || q.lens_mul != a.lens_mul ||
        q.in_act != a.in_act || q.in_slope != a.in_slope || q.acc != EV_ACC_STORE || !q.x || !q.w || !q.out)
      return false;
    if (q.K < 1 || !(q.K & 1) || q.dil < 1) return false;
    for (int j = 0; j < i; ++j)
      if (ps[j].out == q.out) return false;                 // every member writes its own tensor
    const int k = gp_shape_kbg(q, mode);                    // the reduction order of each member must be its own launch's
    if (i == 0) kb = k;
    else if (k != kb) return false;
  }
  *kbg = kb;
  return true;
}
bool gp_group_supported(const GpConvParams* ps, int n, int mode) {
  int kbg = 0;
  if (mode < 0 || mode > 3 || !group_shapes_match(ps, n, mode, &kbg)) return false;
  return validate_gp(ps[0], mode) == EV_OK;
}
static int plan_group(const GpConvParams* ps, int n, int mode, GpGroups* gs_out, gp::GPlan* pl_out);
int debug_gp_group_plan(const GpConvParams* ps, int n, int mode, int* v) {
  GpGroups gs{};
  gp::GPlan pl;
  EV_TRY(plan_group(ps, n, mode, &gs, &pl));
  v[0] = pl.BN; v[1] = pl.mt; v[2] = pl.kbg; v[3] = pl.a_stages; v[4] = pl.b_stages; v[5] = gp::NTW;
  v[6] = pl.planes; v[7] = pl.tmem_cols; v[8] = pl.smem_total; v[9] = pl.total_tiles; v[10] = pl.rows_pad;
  return EV_OK;
}
int launch_conv1d_gp_group(const GpConvParams* ps, int n, int mode, cudaStream_t st) {
  GpGroups gs{};
  gp::GPlan pl;
  EV_TRY(plan_group(ps, n, mode, &gs, &pl));
  return dispatch_gp(ps[0], pl, gs, mode, st);
}
static int plan_group(const GpConvParams* ps, int n, int mode, GpGroups* gs_out, gp::GPlan* pl_out) {
  GpGroups& gs = *gs_out;
  gp::GPlan& pl = *pl_out;
  int kbg = 0;
  EV_CHECK_ARG(ps && group_shapes_match(ps, n, mode, &kbg), "conv1d_gp group: the %d convolutions do not share a launch shape", n);
  // heaviest member first: with static round-robin tiles the CTAs that take a second (third) tile then take a light one
  int order[3] = {0, 1, 2};
  for (int i = 0; i < n; ++i)
    for (int j = i + 1; j < n; ++j)
      if (ps[order[j]].K > ps[order[i]].K) { const int t = order[i]; order[i] = order[j]; order[j] = t; }
  gs = GpGroups{};
  gs.ng = n;
  int ksum = 0, span = 0, kmin = 1 << 30;
  for (int i = 0; i < n; ++i) {
    const GpConvParams& q = ps[order[i]];
    gs.g[i] = GpGroup{q.x, q.w, q.bias, q.res, q.out, q.K, q.dil};
    ksum += q.K;
    span = (q.K - 1) * q.dil > span ? (q.K - 1) * q.dil : span;
    kmin = q.K < kmin ? q.K : kmin;
  }
  int gK[3] = {0, 0, 0};
  for (int i = 0; i < n; ++i) gK[i] = gs.g[i].K;
  return plan_gp(ps[0], mode, &pl, n, ksum, span, kmin, kbg, gK);
}

// Load every instantiation's code now (CUDA loads kernels lazily, at their first launch: tens of milliseconds for a kernel of this
// size, which would otherwise hit whichever utterance first needs a new tile shape) and set the shared-memory attribute.
template <int MODE, int KBG>
static void preload_gp_mode() {
  cudaFuncSetAttribute(gp::conv1d_gp_kernel<MODE, 1, KBG>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  cudaFuncSetAttribute(gp::conv1d_gp_kernel<MODE, 2, KBG>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  cudaFuncSetAttribute(gp::conv1d_gp_kernel<MODE, 4, KBG>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
}
void preload_conv1d_gp() {
  preload_gp_mode<0, 4>(); preload_gp_mode<0, 8>(); preload_gp_mode<1, 4>(); preload_gp_mode<2, 4>(); preload_gp_mode<2, 8>();
  preload_gp_mode<3, 4>(); preload_gp_mode<3, 8>();
  cudaFuncSetAttribute(gp::conv_post_gp_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
  cudaFuncSetAttribute(gp::conv_post_gp_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
  cudaFuncSetAttribute(gp::conv_post_gp4_kernel<false, 7>, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024);
  cudaFuncSetAttribute(gp::conv_post_gp4_kernel<true, 7>, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024);
  cudaFuncAttributes fa;
  cudaFuncGetAttributes(&fa, gp::to_gp_kernel<false>);
  cudaFuncGetAttributes(&fa, gp::to_gp_kernel<true>);
  cudaGetLastError();
}

int launch_to_gp(const float* in, long long sb, long long st_, long long sc, void* out, int B, int L, int C, int bf16, cudaStream_t st) {
  EV_CHECK_ARG(B > 0 && L > 0 && C % (bf16 ? 8 : 4) == 0, "to_gp: B=%d L=%d C=%d", B, L, C);
  const size_t n = (size_t)B * (C / (bf16 ? 8 : 4)) * L;
  const unsigned grid = (unsigned)((n + 255) / 256);
  cudaError_t e = cudaSuccess;
  if (pdl_mode()) {
    e = bf16 ? launch_with_pdl(gp::to_gp_kernel<true>, dim3(grid), dim3(256), 0, st, in, sb, st_, sc, out, B, L, C)
             : launch_with_pdl(gp::to_gp_kernel<false>, dim3(grid), dim3(256), 0, st, in, sb, st_, sc, out, B, L, C);
  } else if (bf16) {
    gp::to_gp_kernel<true><<<grid, 256, 0, st>>>(in, sb, st_, sc, out, B, L, C);
  } else {
    gp::to_gp_kernel<false><<<grid, 256, 0, st>>>(in, sb, st_, sc, out, B, L, C);
  }
  park_launch_error(e);
  EV_CUDA_LAUNCH_CHECK("to_gp_kernel");
  return EV_OK;
}

int launch_gp_sum_div(const float* a, const float* b, const float* c, float* out, size_t n_floats, float div, cudaStream_t st) {
  EV_CHECK_ARG(a && b && out && n_floats % 4 == 0, "gp_sum_div: bad arguments");
  const size_t n4 = n_floats / 4;
  const unsigned grid = (unsigned)((n4 + 255) / 256);
  const float4 *a4 = reinterpret_cast<const float4*>(a), *b4 = reinterpret_cast<const float4*>(b), *c4 = reinterpret_cast<const float4*>(c);
  float4* o4 = reinterpret_cast<float4*>(out);
  cudaError_t e = cudaSuccess;
  if (pdl_mode()) {
    e = c ? launch_with_pdl(gp::gp_sum_div_kernel<3>, dim3(grid), dim3(256), 0, st, a4, b4, c4, o4, n4, div)
          : launch_with_pdl(gp::gp_sum_div_kernel<2>, dim3(grid), dim3(256), 0, st, a4, b4, c4, o4, n4, div);
  } else if (c) {
    gp::gp_sum_div_kernel<3><<<grid, 256, 0, st>>>(a4, b4, c4, o4, n4, div);
  } else {
    gp::gp_sum_div_kernel<2><<<grid, 256, 0, st>>>(a4, b4, c4, o4, n4, div);
  }
  park_launch_error(e);
  EV_CUDA_LAUNCH_CHECK("gp_sum_div_kernel");
  return EV_OK;
}

int launch_conv_post_gp(const void* x, int bf16, const float* w, const float* bias, const int32_t* lens, int lens_mul, int B, int L, int C, int K,
                        float slope, float* wav, cudaStream_t st) {
  EV_CHECK_ARG(C % (bf16 ? 8 : 4) == 0 && C <= 128 && K <= 15 && (K & 1), "conv_post_gp: C=%d K=%d", C, K);
  EV_CHECK_ARG(B > 0 && B <= 65535 && L > 0, "conv_post_gp: bad shape");
  const size_t smem = (size_t)(C * (gp::GPP_BT + K - 1) + K * C) * sizeof(float);
  static std::atomic<uint64_t> attr_devs{0};
  if (first_use_on_device(attr_devs)) {
    cudaFuncSetAttribute(gp::conv_post_gp_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    cudaFuncSetAttribute(gp::conv_post_gp_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    cudaFuncSetAttribute(gp::conv_post_gp4_kernel<false, 7>, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024);
    cudaFuncSetAttribute(gp::conv_post_gp4_kernel<true, 7>, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024);
  }
  if (K == 7 && C <= 48) {
    constexpr int RS = gp::GP4_ROWS + 8;
    const size_t smem4 = (size_t)C * (RS + 8) * sizeof(float);
    dim3 grid4((L + gp::GP4_ROWS - 1) / gp::GP4_ROWS, B);
    cudaError_t e4 = cudaSuccess;
    if (pdl_mode()) {
      e4 = bf16 ? launch_with_pdl(gp::conv_post_gp4_kernel<true, 7>, grid4, dim3(gp::GP4_T), smem4, st, x, w, bias, lens, lens_mul, L, C, slope, wav)
                : launch_with_pdl(gp::conv_post_gp4_kernel<false, 7>, grid4, dim3(gp::GP4_T), smem4, st, x, w, bias, lens, lens_mul, L, C, slope, wav);
    } else if (bf16) {
      gp::conv_post_gp4_kernel<true, 7><<<grid4, gp::GP4_T, smem4, st>>>(x, w, bias, lens, lens_mul, L, C, slope, wav);
    } else {
      gp::conv_post_gp4_kernel<false, 7><<<grid4, gp::GP4_T, smem4, st>>>(x, w, bias, lens, lens_mul, L, C, slope, wav);
    }
    park_launch_error(e4);
    EV_CUDA_LAUNCH_CHECK("conv_post_gp4_kernel");
    return EV_OK;
  }
  dim3 grid((L + gp::GPP_BT - 1) / gp::GPP_BT, B);
  cudaError_t e = cudaSuccess;
  if (pdl_mode()) {
    e = bf16 ? launch_with_pdl(gp::conv_post_gp_kernel<true>, grid, dim3(gp::GPP_BT), smem, st, x, w, bias, lens, lens_mul, L, C, K, slope, wav)
             : launch_with_pdl(gp::conv_post_gp_kernel<false>, grid, dim3(gp::GPP_BT), smem, st, x, w, bias, lens, lens_mul, L, C, K, slope, wav);
  } else if (bf16) {
    gp::conv_post_gp_kernel<true><<<grid, gp::GPP_BT, smem, st>>>(x, w, bias, lens, lens_mul, L, C, K, slope, wav);
  } else {
    gp::conv_post_gp_kernel<false><<<grid, gp::GPP_BT, smem, st>>>(x, w, bias, lens, lens_mul, L, C, K, slope, wav);
  }
  park_launch_error(e);
  EV_CUDA_LAUNCH_CHECK("conv_post_gp_kernel");
  return EV_OK;
}

}  // namespace ev
