// One ResBlock1 layer of HiFi-GAN as ONE kernel on granule-planar activations (sm_100a):
//     out = [acc]( x + c2( lrelu( c1( lrelu(x), dilation d ) ), dilation 1 ) )            (hifigan/models.py:50-57)
// The intermediate xt = c1(lrelu(x)) never leaves the SM: c1's epilogue writes lrelu(acc1 + b1) -- rounded / split exactly like
// conv1d_gp.cu's transform warps would after a round trip through HBM -- straight from the TMEM lane layout (thread = row) into
// a shared-memory tile that already has the K-major UMMA operand layout, and c2's taps are descriptor shifts over that tile.
// Per layer the activation traffic drops from 5 passes (x, xt write, xt read, x residual, out) to 2 (x, out; the residual
// re-read hits L2) and the launches halve.  Same reduction orders and roundings as two conv1d_gp launches: BITWISE equal to them.
//
// Tile = R = 128*MT - (K-1) output rows.  c1 computes MT accumulators for the 128*MT xt rows [t0 - (K-1)/2, ...) from a staged x
// tile of 128*MT + (K-1)*d rows (bulk copies + transform warps, as in conv1d_gp.cu); c2 computes MT accumulators from the xt
// tile (its last K-1 rows per tile are surplus and discarded).  xt rows outside [0, len) are written as zeros: the reference
// pads c2's input with zeros, it does not convolve c1 over the padding.
//
// Unlike the round-1 fused kernel (which lost 10-21 %: one accumulator set, c1's epilogue, c2 and c2's epilogue serialised per
// tile) everything is double buffered and the two epilogues are separate warp groups:
//   MMA order   C1(0) | C1(i+1), C2(i) | ...        acc1[2], acc2[2] in TMEM (4*MT*C <= 512 columns)
//   warps 4-7   epi1: acc1 -> xt tile (shared memory)         warps 0-3   epi2: acc2 + b2 + x (+ accumulate modes) -> HBM
//   warps 8-11  transform the landed x stages in place; warp 12 x loader; warp 13 weight loader; warp 14 TMEM alloc + MMA issue
// so c1 of the next tile runs under c2 / both epilogues of the current one and the x ring prefetches across tiles.
#include "ev_common.cuh"
#include "tc_common.cuh"

namespace ev {
namespace gpp {

using namespace tc;

constexpr int NTW = 4;
constexpr int W_EPI1 = 4, W_XFORM = 8, W_ALOAD = 12, W_BLOAD = 13, W_MMA = 14;
constexpr int GPP_THREADS = 15 * 32;
constexpr int MAX_A = 6, MAX_B = 8;
constexpr int SMEM_HEAD = 1024;

struct PPlan {
  int mt, kbg;
  int rows1_pad, rows2_pad, R;
  int x_plane_bytes, x_stage_bytes, a2_plane_bytes, a2_bytes, b_plane_bytes, b_stage_bytes;
  int a_stages, b_stages;
  int tmem_cols;
  int tiles_m, total_tiles;
  int smem_total;
};

// span1 = (K-1)*dil of the widest member and kmax = the most taps of a grouped launch (-1 / 0: p's own)
__host__ __device__ inline bool make_pplan(const GpPairParams& p, int mode, int mt, int kbg, PPlan* o, int span1 = -1, int kmax = 0) {
  if (span1 < 0) span1 = (p.K - 1) * p.dil;
  if (kmax <= 0) kmax = p.K;
  PPlan q;
  q.mt = mt; q.kbg = kbg;
  const int C = p.C;
  if (C > 128 || C % 32 || 4 * mt * C > 512) return false;
  const int cpg = mode == 2 ? 8 : 4;            // activation granule (HBM / x tile)
  const int ocpg = mode >= 2 ? 8 : 4;           // operand granule of the xt tile and the weights
  const int xplanes = mode == 1 ? 2 : 1;
  const int splitp = (mode == 1 || mode == 3) ? 2 : 1;
  q.R = BM * mt - (p.K - 1);
  if (BM * mt - (kmax - 1) < 32) return false;
  const int rows1 = BM * mt + span1;
  q.rows1_pad = (rows1 + 7) / 8 * 8;
  q.rows2_pad = (BM * mt + (kmax - 1) + 7) / 8 * 8;
  q.x_plane_bytes = kbg * q.rows1_pad * 16;
  q.x_stage_bytes = xplanes * q.x_plane_bytes;
  q.a2_plane_bytes = (C / ocpg) * q.rows2_pad * 16;
  q.a2_bytes = splitp * q.a2_plane_bytes;
  q.b_plane_bytes = (kbg * cpg / ocpg) * C * 16;
  q.b_stage_bytes = splitp * q.b_plane_bytes;
  const int budget = 227 * 1024 - SMEM_HEAD - q.a2_bytes;
  q.a_stages = 2; q.b_stages = 2;
  auto fits = [&](int a, int b) { return a * q.x_stage_bytes + b * q.b_stage_bytes <= budget; };
  if (!fits(2, 2)) return false;
  while (q.b_stages < 4 && fits(q.a_stages, q.b_stages + 1)) ++q.b_stages;
  while (q.a_stages < 4 && fits(q.a_stages + 1, q.b_stages)) ++q.a_stages;
  while (q.b_stages < MAX_B && fits(q.a_stages, q.b_stages + 1)) ++q.b_stages;
  while (q.a_stages < MAX_A && fits(q.a_stages + 1, q.b_stages)) ++q.a_stages;
  q.tmem_cols = 32;
  while (q.tmem_cols < 4 * mt * C) q.tmem_cols <<= 1;
  q.tiles_m = (p.L + q.R - 1) / q.R;
  q.total_tiles = p.B * q.tiles_m;
  q.smem_total = SMEM_HEAD + q.a2_bytes + q.a_stages * q.x_stage_bytes + q.b_stages * q.b_stage_bytes;
  *o = q;
  return true;
}

// LeakyReLU for 0 <= slope <= 1 as max(v, v*slope): two instructions (FMUL + FMNMX) instead of compare / multiply / select; same bits
__device__ __forceinline__ float lrelu_f(float v, float slope) { return fmaxf(v, v * slope); }

// MODE as conv1d_gp.cu: 0 tf32, 1 3xTF32, 2 bf16 activations + operands, 3 bf16x3 on fp32 activations.
template <int MODE, int MT, int KBG>
__global__ void __launch_bounds__(GPP_THREADS, 1) resblock_gp_kernel(const __grid_constant__ GpPairParams p, const __grid_constant__ PPlan pl,
                                                                     const __grid_constant__ GpPairGroups gs) {
  constexpr bool SPLIT3 = (MODE == 1), BF16 = (MODE == 2), X3B = (MODE == 3);
  constexpr bool OP16 = BF16 || X3B;
  constexpr int SPL = (SPLIT3 || X3B) ? 2 : 1;
  constexpr int CPG = BF16 ? 8 : 4;
  constexpr int OCPG = OP16 ? 8 : 4;
  constexpr int KB = CPG * KBG;
  constexpr int KBGW = KBG * CPG / OCPG;        // operand granules (weights, xt tile) per channel block
  static_assert(!X3B || KBG % 4 == 0, "bf16x3 consumes four fp32 granules per MMA K step");
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int C = p.C;

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem_raw + 768);
  uint8_t* a2_tile = smem_raw + SMEM_HEAD;
  uint8_t* x_tiles = a2_tile + pl.a2_bytes;
  uint8_t* b_tiles = x_tiles + pl.a_stages * pl.x_stage_bytes;
  const uint32_t bar_base = smem_u32(bars);
  auto a_full = [&](int s) { return bar_base + 8u * s; };
  auto a_ready = [&](int s) { return bar_base + 8u * (MAX_A + s); };
  auto a_empty = [&](int s) { return bar_base + 8u * (2 * MAX_A + s); };
  auto b_full = [&](int s) { return bar_base + 8u * (3 * MAX_A + s); };
  auto b_empty = [&](int s) { return bar_base + 8u * (3 * MAX_A + MAX_B + s); };
  const uint32_t misc = bar_base + 8u * (3 * MAX_A + 2 * MAX_B);
  auto acc1_full = [&](int s) { return misc + 8u * s; };
  auto acc1_empty = [&](int s) { return misc + 8u * (2 + s); };
  auto acc2_full = [&](int s) { return misc + 8u * (4 + s); };
  auto acc2_empty = [&](int s) { return misc + 8u * (6 + s); };
  const uint32_t a2_full = misc + 64u, a2_empty = misc + 72u;

  if (tid == 0) {
    for (int s = 0; s < pl.a_stages; ++s) { mbar_init(a_full(s), 1); mbar_init(a_ready(s), NTW * 32); mbar_init(a_empty(s), 1); }
    for (int s = 0; s < pl.b_stages; ++s) { mbar_init(b_full(s), 1); mbar_init(b_empty(s), 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(acc1_full(s), 1); mbar_init(acc1_empty(s), 4); mbar_init(acc2_full(s), 1); mbar_init(acc2_empty(s), 4); }
    mbar_init(a2_full, 128); mbar_init(a2_empty, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == W_MMA) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(pl.tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  const int n_cb = (C + KB - 1) / KB;
  const int gC = C / CPG;                      // activation granule planes per item
  // TMEM columns: acc1[buf] at buf*MT*C, acc2[buf] at (2 + buf)*MT*C
  // A launch carries up to three layers of one shape (the same-index layers of HiFi-GAN's parallel ResBlocks: their own taps,
  // dilation, weights, tensors, hence their own rows per tile and tile count); tiles are numbered member after member.
  auto group_of = [&](int tile) { return (gs.ng > 1 && tile >= gs.g[1].tile0) ? ((gs.ng > 2 && tile >= gs.g[2].tile0) ? 2 : 1) : 0; };
  auto tile_len = [&](int tile, int& gi, int& b, int& t0) {
    gi = group_of(tile);
    const int local = tile - gs.g[gi].tile0;
    b = local / gs.g[gi].tiles_m;
    t0 = (local - b * gs.g[gi].tiles_m) * gs.g[gi].R;
    return p.lens ? min(p.L, p.lens[b] * p.lens_mul) : p.L;
  };
  auto active = [&](int tile) { int gi, b, t0; const int len = tile_len(tile, gi, b, t0); return t0 < len; };
  auto next_active = [&](int tile) {      // first active tile of this CTA at or after `tile` (stride gridDim.x); >= total when none
    while (tile < pl.total_tiles && !active(tile)) tile += gridDim.x;
    return tile;
  };

  if (warp < W_EPI1) {
    // ============================ epi2: acc2 + b2 + x (+ accumulate) -> out ================================
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const int quad = warp;
    const int nchunks = C / 32;
    constexpr int NG = 32 / CPG;
    int cnt = 0;
    for (int tile = next_active(blockIdx.x); tile < pl.total_tiles; tile = next_active(tile + gridDim.x), ++cnt) {
      int gi, b, t0;
      const int len = tile_len(tile, gi, b, t0);
      const GpPairGroup& G = gs.g[gi];
      const int R = G.R;
      const int buf = cnt & 1;
      bool waited = false;
#pragma unroll 1
      for (int item = 0; item < MT * nchunks; ++item) {
        const int mt = item / nchunks, c = (item - mt * nchunks) * 32;
        const int rl = mt * BM + quad * 32 + lane;              // row inside the tile
        const int row = t0 + rl;
        const bool ok = rl < R && row < len;
        const size_t gbase = ((size_t)b * gC + c / CPG) * p.L + row;
        uint4 rq[NG];
#pragma unroll
        for (int q = 0; q < NG; ++q) {
          rq[q] = make_uint4(0u, 0u, 0u, 0u);
          if (ok) rq[q] = *(reinterpret_cast<const uint4*>(G.x) + gbase + (size_t)q * p.L);       // the residual: L2 hit (the x tile was just staged)
        }
        if (!waited) {
          mbar_wait(acc2_full(buf), (cnt >> 1) & 1);
          tc_fence_after();
          waited = true;
        }
        float v[32];
        tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)((2 + buf) * MT * C + mt * C + c), 32, v);
        if (ok) {
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(G.b2 + c) + q);
            v[4 * q] += b4.x; v[4 * q + 1] += b4.y; v[4 * q + 2] += b4.z; v[4 * q + 3] += b4.w;
          }
#pragma unroll
          for (int q = 0; q < NG; ++q) {
            if (BF16) {
              const uint32_t w4[4] = {rq[q].x, rq[q].y, rq[q].z, rq[q].w};
#pragma unroll
              for (int e = 0; e < 4; ++e) { v[8 * q + 2 * e] += __uint_as_float(w4[e] << 16); v[8 * q + 2 * e + 1] += __uint_as_float(w4[e] & 0xffff0000u); }
            } else {
              v[4 * q] += __uint_as_float(rq[q].x); v[4 * q + 1] += __uint_as_float(rq[q].y);
              v[4 * q + 2] += __uint_as_float(rq[q].z); v[4 * q + 3] += __uint_as_float(rq[q].w);
            }
          }
          if (p.acc != EV_ACC_STORE) {
            uint4 oq[NG];
#pragma unroll
            for (int q = 0; q < NG; ++q) oq[q] = *(reinterpret_cast<const uint4*>(G.out) + gbase + (size_t)q * p.L);
#pragma unroll
            for (int q = 0; q < NG; ++q) {
              if (BF16) {
                const uint32_t w4[4] = {oq[q].x, oq[q].y, oq[q].z, oq[q].w};
#pragma unroll
                for (int e = 0; e < 4; ++e) { v[8 * q + 2 * e] += __uint_as_float(w4[e] << 16); v[8 * q + 2 * e + 1] += __uint_as_float(w4[e] & 0xffff0000u); }
              } else {
                v[4 * q] += __uint_as_float(oq[q].x); v[4 * q + 1] += __uint_as_float(oq[q].y);
                v[4 * q + 2] += __uint_as_float(oq[q].z); v[4 * q + 3] += __uint_as_float(oq[q].w);
              }
            }
            if (p.acc == EV_ACC_ADD_DIV) {
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
              o.x = __float_as_uint(v[4 * q]); o.y = __float_as_uint(v[4 * q + 1]); o.z = __float_as_uint(v[4 * q + 2]); o.w = __float_as_uint(v[4 * q + 3]);
            }
            *(reinterpret_cast<uint4*>(G.out) + gbase + (size_t)q * p.L) = o;
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc2_empty(buf));
    }
  } else if (warp < W_XFORM) {
    // ============================ epi1: acc1 + b1 -> lrelu -> operand format -> xt tile in shared memory ===================
    const int quad = warp - W_EPI1;
    const int nchunks = C / 32;
    const float slope = p.slope;
    int cnt = 0;
    for (int tile = next_active(blockIdx.x); tile < pl.total_tiles; tile = next_active(tile + gridDim.x), ++cnt) {
      int gi, b, t0;
      const int len = tile_len(tile, gi, b, t0);
      const int h2 = (gs.g[gi].K - 1) / 2;
      const float* b1 = gs.g[gi].b1;
      const int buf = cnt & 1;
      mbar_wait(acc1_full(buf), (cnt >> 1) & 1);
      mbar_wait(a2_empty, (cnt & 1) ^ 1);                 // c2 of the previous tile has read the xt tile
      tc_fence_after();
#pragma unroll 1
      for (int item = 0; item < MT * nchunks; ++item) {
        const int mt = item / nchunks, c = (item - mt * nchunks) * 32;
        const int r2 = mt * BM + quad * 32 + lane;
        const int row = t0 - h2 + r2;
        const bool ok = row >= 0 && row < len;            // xt outside the sequence is c2's ZERO padding
        float v[32];
        tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(buf * MT * C + mt * C + c), 32, v);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 b4 = __ldg(reinterpret_cast<const float4*>(b1 + c) + q);
          v[4 * q] += b4.x; v[4 * q + 1] += b4.y; v[4 * q + 2] += b4.z; v[4 * q + 3] += b4.w;
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          float t = v[i];
          if (BF16) t = __uint_as_float(pack_bf16(t, 0.f) << 16);      // the unfused path stores xt as bf16 before activating it
          v[i] = ok ? lrelu_f(t, slope) : 0.f;
        }
        uint8_t* dst = a2_tile + ((size_t)(c / OCPG) * pl.rows2_pad + r2) * 16;
        constexpr int NGO = 32 / OCPG;
#pragma unroll
        for (int q = 0; q < NGO; ++q) {
          uint8_t* d = dst + (size_t)q * pl.rows2_pad * 16;
          if (OP16) {
            uint32_t hi[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) hi[e] = pack_bf16(v[8 * q + 2 * e], v[8 * q + 2 * e + 1]);
            *reinterpret_cast<uint4*>(d) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            if (X3B) {
              uint32_t lo[4];
#pragma unroll
              for (int e = 0; e < 4; ++e)
                lo[e] = pack_bf16(v[8 * q + 2 * e] - __uint_as_float(hi[e] << 16), v[8 * q + 2 * e + 1] - __uint_as_float(hi[e] & 0xffff0000u));
              *reinterpret_cast<uint4*>(d + pl.a2_plane_bytes) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
            }
          } else {
            const float4 h = make_float4(to_tf32(v[4 * q]), to_tf32(v[4 * q + 1]), to_tf32(v[4 * q + 2]), to_tf32(v[4 * q + 3]));
            *reinterpret_cast<float4*>(d) = h;
            if (SPLIT3) {
              const float4 l = make_float4(to_tf32(v[4 * q] - h.x), to_tf32(v[4 * q + 1] - h.y), to_tf32(v[4 * q + 2] - h.z), to_tf32(v[4 * q + 3] - h.w));
              *reinterpret_cast<float4*>(d + pl.a2_plane_bytes) = l;
            }
          }
        }
      }
      fence_proxy_async();
      mbar_arrive(a2_full);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc1_empty(buf));
    }
  } else if (warp < W_ALOAD) {
    // ============================ transform warps: in-place pass over the landed x stage =======================
    const int xt = (warp - W_XFORM) * 32 + lane;
    const float slope = p.slope;
    int a_cnt = 0;
    for (int tile = next_active(blockIdx.x); tile < pl.total_tiles; tile = next_active(tile + gridDim.x)) {
      int gi, b, t0;
      const int len = tile_len(tile, gi, b, t0);
      const int span = (gs.g[gi].K - 1) * gs.g[gi].dil;
      const int rows1 = BM * MT + span;
      const int row0 = t0 - (gs.g[gi].K - 1) / 2 - span / 2;
      for (int cb = 0; cb < n_cb; ++cb, ++a_cnt) {
        const int s = a_cnt % pl.a_stages;
        const int ngran = min(KB, C - cb * KB) / CPG;
        uint8_t* base = x_tiles + s * pl.x_stage_bytes;
        mbar_wait(a_full(s), (a_cnt / pl.a_stages) & 1);
        if (X3B) {
          for (int q = 0; q < ngran / 2; ++q) {
            uint8_t* g0 = base + (size_t)(2 * q) * pl.rows1_pad * 16;
            uint8_t* g1 = g0 + (size_t)pl.rows1_pad * 16;
            for (int r = xt; r < rows1; r += NTW * 32) {
              const int row = row0 + r;
              float4 u = make_float4(0.f, 0.f, 0.f, 0.f), w = u;
              if (row >= 0 && row < len) { u = *reinterpret_cast<const float4*>(g0 + r * 16); w = *reinterpret_cast<const float4*>(g1 + r * 16); }
              const float f[8] = {u.x, u.y, u.z, u.w, w.x, w.y, w.z, w.w};
              uint32_t hi[4], lo[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float a0 = lrelu_f(f[2 * e], slope), a1 = lrelu_f(f[2 * e + 1], slope);
                hi[e] = pack_bf16(a0, a1);
                lo[e] = pack_bf16(a0 - __uint_as_float(hi[e] << 16), a1 - __uint_as_float(hi[e] & 0xffff0000u));
              }
              *reinterpret_cast<uint4*>(g0 + r * 16) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
              *reinterpret_cast<uint4*>(g1 + r * 16) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
            }
          }
        } else {
          for (int g = 0; g < ngran; ++g) {
            uint8_t* gb = base + (size_t)g * pl.rows1_pad * 16;
            for (int r = xt; r < rows1; r += NTW * 32) {
              const int row = row0 + r;
              uint4 v = make_uint4(0u, 0u, 0u, 0u);
              if (row >= 0 && row < len) v = *reinterpret_cast<const uint4*>(gb + r * 16);
              if (BF16) {
                uint32_t w4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int e = 0; e < 4; ++e)
                  w4[e] = pack_bf16(lrelu_f(__uint_as_float(w4[e] << 16), slope), lrelu_f(__uint_as_float(w4[e] & 0xffff0000u), slope));
                *reinterpret_cast<uint4*>(gb + r * 16) = make_uint4(w4[0], w4[1], w4[2], w4[3]);
              } else {
                float4 t = make_float4(lrelu_f(__uint_as_float(v.x), slope), lrelu_f(__uint_as_float(v.y), slope), lrelu_f(__uint_as_float(v.z), slope),
                                       lrelu_f(__uint_as_float(v.w), slope));
                const float4 h = make_float4(to_tf32(t.x), to_tf32(t.y), to_tf32(t.z), to_tf32(t.w));
                *reinterpret_cast<float4*>(gb + r * 16) = h;
                if (SPLIT3) {
                  const float4 l = make_float4(to_tf32(t.x - h.x), to_tf32(t.y - h.y), to_tf32(t.z - h.z), to_tf32(t.w - h.w));
                  *reinterpret_cast<float4*>(gb + pl.x_plane_bytes + r * 16) = l;
                }
              }
            }
          }
        }
        fence_proxy_async();
        mbar_arrive(a_ready(s));
      }
    }
  } else if (warp == W_ALOAD) {
    // ============================ x loader ===========================================================================
    if (lane == 0) {
      asm volatile("griddepcontrol.wait;" ::: "memory");
      int a_cnt = 0;
      for (int tile = next_active(blockIdx.x); tile < pl.total_tiles; tile = next_active(tile + gridDim.x)) {
        int gi, b, t0;
        const int len = tile_len(tile, gi, b, t0);
        const int span = (gs.g[gi].K - 1) * gs.g[gi].dil;
        const int rows1 = BM * MT + span;
        const int row0 = t0 - (gs.g[gi].K - 1) / 2 - span / 2;
        const int r_lo = max(row0, 0), r_hi = min(row0 + rows1, len);
        const uint32_t nbytes = (uint32_t)(r_hi - r_lo) * 16u, roff = (uint32_t)(r_lo - row0) * 16u;
        const uint8_t* xb = reinterpret_cast<const uint8_t*>(gs.g[gi].x) + ((size_t)b * gC * p.L + r_lo) * 16;
        for (int cb = 0; cb < n_cb; ++cb, ++a_cnt) {
          const int s = a_cnt % pl.a_stages;
          const int ngran = min(KB, C - cb * KB) / CPG;
          mbar_wait(a_empty(s), ((a_cnt / pl.a_stages) & 1) ^ 1);
          mbar_expect_tx(a_full(s), (uint32_t)ngran * nbytes);
          const uint32_t dst = smem_u32(x_tiles + s * pl.x_stage_bytes) + roff;
          const uint8_t* src = xb + (size_t)(cb * KBG) * p.L * 16;
          for (int g = 0; g < ngran; ++g) bulk_g2s(dst + (uint32_t)(g * pl.rows1_pad * 16), src + (size_t)g * p.L * 16, nbytes, a_full(s));
        }
      }
    }
    __syncwarp();
  } else if (warp == W_BLOAD) {
    // ============================ weight loader: the MMA issuer's order  C1(0) | C1(i+1), C2(i) ====================
    if (lane == 0) {
      const int win = C / OCPG;
      int b_cnt = 0;
      auto stream = [&](const float* w, int K) {
        const size_t plane = (size_t)K * win * C * 4;        // 4-byte words per plane
        for (int cb = 0; cb < n_cb; ++cb) {
          const int ngran = min(KB, C - cb * KB) / OCPG;
          for (int j = 0; j < K; ++j, ++b_cnt) {
            const int sb = b_cnt % pl.b_stages;
            mbar_wait(b_empty(sb), ((b_cnt / pl.b_stages) & 1) ^ 1);
            mbar_expect_tx(b_full(sb), (uint32_t)(SPL * ngran * C * 16));
            const uint32_t dst = smem_u32(b_tiles + sb * pl.b_stage_bytes);
            const float* src = w + ((size_t)j * win + (size_t)cb * KBGW) * C * 4;
            bulk_g2s(dst, src, (uint32_t)(ngran * C * 16), b_full(sb));
            if (SPL == 2) bulk_g2s(dst + (uint32_t)pl.b_plane_bytes, src + plane, (uint32_t)(ngran * C * 16), b_full(sb));
          }
        }
      };
      int tile = next_active(blockIdx.x);
      if (tile < pl.total_tiles) stream(gs.g[group_of(tile)].w1, gs.g[group_of(tile)].K);
      while (tile < pl.total_tiles) {
        const int nxt = next_active(tile + gridDim.x);
        if (nxt < pl.total_tiles) stream(gs.g[group_of(nxt)].w1, gs.g[group_of(nxt)].K);
        stream(gs.g[group_of(tile)].w2, gs.g[group_of(tile)].K);
        tile = nxt;
      }
    }
    __syncwarp();
  } else {
    // ============================ MMA issuer (all lanes in the control flow, one elected lane issues) =======================
    const uint32_t x_lbo = (uint32_t)pl.rows1_pad * 16u, a2_lbo = (uint32_t)pl.rows2_pad * 16u, b_lbo = (uint32_t)C * 16u;
    const uint32_t fmt = OP16 ? 1u : 2u;
    const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(C >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
    const uint64_t x_desc0 = make_desc(0u, X3B ? 2u * x_lbo : x_lbo, 128u), b_desc0 = make_desc(0u, b_lbo, 128u);
    const uint64_t a2_desc0 = make_desc(smem_u32(a2_tile), a2_lbo, 128u);
    const uint32_t x_k = (X3B ? 4u : 2u) * x_lbo, x_lo_off = X3B ? x_lbo : (uint32_t)pl.x_plane_bytes;
    int a_cnt = 0, b_cnt = 0;
    auto mma3 = [&](uint32_t d, uint64_t a_hi, uint64_t a_lo, uint64_t b_hi, uint64_t b_lo, uint32_t first) {
      if (SPLIT3) {
        umma_tf32(d, a_lo, b_hi, idesc, first); umma_tf32(d, a_hi, b_lo, idesc, 1u); umma_tf32(d, a_hi, b_hi, idesc, 1u);
      } else if (X3B) {
        umma_bf16(d, a_lo, b_hi, idesc, first); umma_bf16(d, a_hi, b_lo, idesc, 1u); umma_bf16(d, a_hi, b_hi, idesc, 1u);
      } else if (BF16) {
        umma_bf16(d, a_hi, b_hi, idesc, first);
      } else {
        umma_tf32(d, a_hi, b_hi, idesc, first);
      }
    };
    auto conv1 = [&](int cnt, int tile) {          // acc1[cnt & 1] = c1 over the staged x tile of the cnt-th active tile
      const int K = gs.g[group_of(tile)].K, dil = gs.g[group_of(tile)].dil;
      const int buf = cnt & 1;
      mbar_wait(acc1_empty(buf), ((cnt >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint32_t d_base = tmem_base + (uint32_t)(buf * MT * C);
      for (int cb = 0; cb < n_cb; ++cb, ++a_cnt) {
        const int sa = a_cnt % pl.a_stages;
        const int nk = min(KB, C - cb * KB) / (2 * OCPG);
        mbar_wait(a_ready(sa), (a_cnt / pl.a_stages) & 1);
        const uint64_t x0 = desc_advance(x_desc0, smem_u32(x_tiles + sa * pl.x_stage_bytes));
        for (int j = 0; j < K; ++j, ++b_cnt) {
          const int sb = b_cnt % pl.b_stages;
          mbar_wait(b_full(sb), (b_cnt / pl.b_stages) & 1);
          tc_fence_after();
          const uint64_t b0 = desc_advance(b_desc0, smem_u32(b_tiles + sb * pl.b_stage_bytes));
          const uint64_t xj = desc_advance(x0, (uint32_t)(j * dil) * 16u);
          if (elect_one()) {
            for (int k = 0; k < nk; ++k) {
              const uint64_t b_hi = desc_advance(b0, (uint32_t)k * 2u * b_lbo);
              const uint64_t b_lo = desc_advance(b_hi, (uint32_t)pl.b_plane_bytes);
              const uint64_t a_k = desc_advance(xj, (uint32_t)k * x_k);
              const uint32_t first = (cb | j | k) != 0 ? 1u : 0u;
#pragma unroll
              for (int mt = 0; mt < MT; ++mt) {
                const uint64_t a_hi = desc_advance(a_k, (uint32_t)(mt * BM) * 16u);
                mma3(d_base + (uint32_t)(mt * C), a_hi, desc_advance(a_hi, x_lo_off), b_hi, b_lo, first);
              }
            }
            umma_commit(b_empty(sb));
            if (j == K - 1) {
              umma_commit(a_empty(sa));
              if (cb == n_cb - 1) umma_commit(acc1_full(buf));
            }
          }
          __syncwarp();
        }
      }
    };
    auto conv2 = [&](int cnt, int tile) {          // acc2[cnt & 1] = c2 over the xt tile epi1 wrote for the cnt-th active tile
      const int K = gs.g[group_of(tile)].K;
      const int buf = cnt & 1;
      mbar_wait(a2_full, cnt & 1);
      mbar_wait(acc2_empty(buf), ((cnt >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint32_t d_base = tmem_base + (uint32_t)((2 + buf) * MT * C);
      for (int cb = 0; cb < n_cb; ++cb) {
        const int nk = min(KB, C - cb * KB) / (2 * OCPG);
        const uint64_t a0 = desc_advance(a2_desc0, (uint32_t)(cb * KBGW) * a2_lbo);
        for (int j = 0; j < K; ++j, ++b_cnt) {
          const int sb = b_cnt % pl.b_stages;
          mbar_wait(b_full(sb), (b_cnt / pl.b_stages) & 1);
          tc_fence_after();
          const uint64_t b0 = desc_advance(b_desc0, smem_u32(b_tiles + sb * pl.b_stage_bytes));
          const uint64_t aj = desc_advance(a0, (uint32_t)j * 16u);
          if (elect_one()) {
            for (int k = 0; k < nk; ++k) {
              const uint64_t b_hi = desc_advance(b0, (uint32_t)k * 2u * b_lbo);
              const uint64_t b_lo = desc_advance(b_hi, (uint32_t)pl.b_plane_bytes);
              const uint64_t a_k = desc_advance(aj, (uint32_t)k * 2u * a2_lbo);
              const uint32_t first = (cb | j | k) != 0 ? 1u : 0u;
#pragma unroll
              for (int mt = 0; mt < MT; ++mt) {
                const uint64_t a_hi = desc_advance(a_k, (uint32_t)(mt * BM) * 16u);
                mma3(d_base + (uint32_t)(mt * C), a_hi, desc_advance(a_hi, (uint32_t)pl.a2_plane_bytes), b_hi, b_lo, first);
              }
            }
            umma_commit(b_empty(sb));
            if (j == K - 1 && cb == n_cb - 1) { umma_commit(a2_empty); umma_commit(acc2_full(buf)); }
          }
          __syncwarp();
        }
      }
    };
    int tile = next_active(blockIdx.x), cnt = 0;
    if (tile < pl.total_tiles) conv1(0, tile);
    while (tile < pl.total_tiles) {
      const int nxt = next_active(tile + gridDim.x);
      if (nxt < pl.total_tiles) conv1(cnt + 1, nxt);     // runs under epi1 / c2 / epi2 of the current tile
      conv2(cnt, tile);
      tile = nxt;
      ++cnt;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == W_MMA) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(pl.tmem_cols));
  }
}

}  // namespace gpp

// K granules per stage: the same function of the layer shape as conv1d_gp.cu's (it fixes the reduction order, and the fused layer
// must stay bitwise equal to the two launches it replaces).
static int pair_shape_kbg(const GpPairParams& p, int mode) {
  GpConvParams q{};
  q.B = 1; q.L = 128; q.Cin = p.C; q.Cout = p.C; q.K = p.K; q.dil = p.dil; q.rate = 1;
  int v[11];
  return debug_gp_plan(q, mode, v) == EV_OK ? v[2] : 4;
}

static bool plan_pair(const GpPairParams& p, int mode, gpp::PPlan* out) {
  if (p.B <= 0 || p.L <= 0 || (p.C != 32 && p.C != 64 && p.C != 128) || !(p.K & 1) || p.dil < 1) return false;
  if (mode >= 2 && p.C % 16) return false;
  if (p.x == p.out) return false;
  const int kbg = pair_shape_kbg(p, mode);
  const int nsm = sm_count();
  // most accumulators per tile (the K-1 surplus rows and the halo are amortised) that still leave about a tile per SM
  for (int mt = 4; mt >= 1; mt >>= 1) {
    gpp::PPlan pl;
    if (!gpp::make_pplan(p, mode, mt, kbg, &pl)) continue;
    if (mt > 1 && pl.total_tiles < nsm) {
      gpp::PPlan smaller;
      if (gpp::make_pplan(p, mode, mt / 2, kbg, &smaller)) continue;
    }
    *out = pl;
    return true;
  }
  return false;
}

// What the ENGINE fuses: shapes whose tile keeps at least two accumulators (R >= 246 of 256 rows).  With one accumulator the K-1
// surplus rows of c2 and the halo re-reads of c1 cost ~9 % of a tile, which the compute-bound 128-channel layers do not win back.
bool gp_pair_supported(const GpPairParams& p, int mode) {
  gpp::PPlan pl;
  return plan_pair(p, mode, &pl) && pl.mt >= 2;
}

int debug_gp_pair_plan(const GpPairParams& p, int mode, int* v) {
  gpp::PPlan pl;
  if (!plan_pair(p, mode, &pl)) { set_error("resblock_gp: shape not supported (C=%d K=%d dil=%d mode=%d)", p.C, p.K, p.dil, mode); return EV_EINVAL; }
  v[0] = pl.mt; v[1] = pl.kbg; v[2] = pl.a_stages; v[3] = pl.b_stages; v[4] = gpp::NTW; v[5] = pl.tmem_cols; v[6] = pl.smem_total;
  v[7] = pl.total_tiles; v[8] = pl.R; v[9] = pl.rows1_pad; v[10] = pl.rows2_pad;
  return EV_OK;
}

template <int MODE, int MT, int KBG>
static int launch_pair_variant(const GpPairParams& p, const gpp::PPlan& pl, const GpPairGroups& gs, cudaStream_t st) {
  static std::atomic<uint64_t> attr_devs{0};
  if (first_use_on_device(attr_devs))
    cudaFuncSetAttribute(gpp::resblock_gp_kernel<MODE, MT, KBG>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  const int nsm = sm_count();
  const int grid = pl.total_tiles < nsm ? pl.total_tiles : nsm;
  if (pdl_mode()) {
    const cudaError_t e = launch_with_pdl(gpp::resblock_gp_kernel<MODE, MT, KBG>, dim3((unsigned)grid), dim3(gpp::GPP_THREADS), (size_t)pl.smem_total, st, p, pl, gs);
    if (e != cudaSuccess) { set_error("resblock_gp_kernel (PDL launch): %s", cudaGetErrorString(e)); return EV_ECUDA; }
    count_launch();
    return EV_OK;
  }
  gpp::resblock_gp_kernel<MODE, MT, KBG><<<grid, gpp::GPP_THREADS, pl.smem_total, st>>>(p, pl, gs);
  EV_CUDA_LAUNCH_CHECK("resblock_gp_kernel");
  return EV_OK;
}

template <int MODE, int KBG>
static int launch_pair_mt(const GpPairParams& p, const gpp::PPlan& pl, const GpPairGroups& gs, cudaStream_t st) {
  if (pl.mt == 4) return launch_pair_variant<MODE, 4, KBG>(p, pl, gs, st);
  if (pl.mt == 2) return launch_pair_variant<MODE, 2, KBG>(p, pl, gs, st);
  return launch_pair_variant<MODE, 1, KBG>(p, pl, gs, st);
}
static int dispatch_pair(const GpPairParams& p, const gpp::PPlan& pl, const GpPairGroups& gs, int mode, cudaStream_t st) {
  if (mode == 1) return launch_pair_mt<1, 4>(p, pl, gs, st);
  if (mode == 3) return pl.kbg == 8 ? launch_pair_mt<3, 8>(p, pl, gs, st) : launch_pair_mt<3, 4>(p, pl, gs, st);
  if (mode == 2) return pl.kbg == 8 ? launch_pair_mt<2, 8>(p, pl, gs, st) : launch_pair_mt<2, 4>(p, pl, gs, st);
  return pl.kbg == 8 ? launch_pair_mt<0, 8>(p, pl, gs, st) : launch_pair_mt<0, 4>(p, pl, gs, st);
}

template <int MODE, int KBG>
static void preload_pair_mode() {
  cudaFuncSetAttribute(gpp::resblock_gp_kernel<MODE, 1, KBG>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  cudaFuncSetAttribute(gpp::resblock_gp_kernel<MODE, 2, KBG>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  cudaFuncSetAttribute(gpp::resblock_gp_kernel<MODE, 4, KBG>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
}
void preload_resblock_gp() {      // see preload_conv1d_gp
  preload_pair_mode<0, 4>(); preload_pair_mode<0, 8>(); preload_pair_mode<1, 4>(); preload_pair_mode<2, 4>(); preload_pair_mode<2, 8>();
  preload_pair_mode<3, 4>(); preload_pair_mode<3, 8>();
  cudaGetLastError();
}

int launch_gp_pair(const GpPairParams& p, int mode, cudaStream_t st) {
  gpp::PPlan pl;
  if (!plan_pair(p, mode, &pl)) { set_error("resblock_gp: shape not supported (C=%d K=%d dil=%d mode=%d)", p.C, p.K, p.dil, mode); return EV_EINVAL; }
  GpPairGroups gs{};
  gs.ng = 1;
  gs.g[0] = GpPairGroup{p.x, p.w1, p.b1, p.w2, p.b2, p.out, p.K, p.dil, pl.R, pl.tiles_m, 0};
  return dispatch_pair(p, pl, gs, mode, st);
}

int gp_pair_solo_tiles(const GpPairParams& p, int mode) {
  gpp::PPlan pl;
  return plan_pair(p, mode, &pl) ? pl.total_tiles : 0;
}

// ---- grouped launch: members ordered heaviest first; one MT for all (the smallest any member's own plan takes, >= 2), stage sizes from
// ---- the widest halos; each member keeps its own rows per tile R = 128*MT - (K-1) and tile count ------------------------------------
static bool plan_pair_group(const GpPairParams* ps, int n, int mode, GpPairGroups* gs, gpp::PPlan* out) {
  if (!ps || n < 1 || n > 3) return false;
  const GpPairParams& a = ps[0];
  int kbg = 0, span1 = 0, kmax = 0;
  for (int i = 0; i < n; ++i) {
    const GpPairParams& q = ps[i];
    if (q.B != a.B || q.L != a.L || q.C != a.C || q.lens != a.lens || q.lens_mul != a.lens_mul || q.slope != a.slope || q.acc != EV_ACC_STORE) return false;
    if (!q.x || !q.out || !q.w1 || !q.w2 || !q.b1 || !q.b2 || q.x == q.out) return false;
    for (int j = 0; j < n; ++j)
      if (j != i && (ps[j].out == q.out || ps[j].out == q.x)) return false;       // a member's output is nobody's input
    gpp::PPlan solo;
    if (!plan_pair(q, mode, &solo) || solo.mt < 2) return false;
    if (i == 0) kbg = solo.kbg;
    else if (solo.kbg != kbg) return false;
    span1 = (q.K - 1) * q.dil > span1 ? (q.K - 1) * q.dil : span1;
    kmax = q.K > kmax ? q.K : kmax;
  }
  int order[3] = {0, 1, 2};
  for (int i = 0; i < n; ++i)
    for (int j = i + 1; j < n; ++j)
      if (ps[order[j]].K > ps[order[i]].K) { const int t = order[i]; order[i] = order[j]; order[j] = t; }
  const int nsm = sm_count();
  for (int mt = 4; mt >= 2; mt >>= 1) {
    gpp::PPlan pl;
    if (!gpp::make_pplan(a, mode, mt, kbg, &pl, span1, kmax)) continue;
    int total = 0;
    gs->ng = n;
    for (int i = 0; i < n; ++i) {
      const GpPairParams& q = ps[order[i]];
      const int R = tc::BM * mt - (q.K - 1);
      const int tm = (q.L + R - 1) / R;
      gs->g[i] = GpPairGroup{q.x, q.w1, q.b1, q.w2, q.b2, q.out, q.K, q.dil, R, tm, total};
      total += q.B * tm;
    }
    if (mt > 2 && total < 2 * nsm) continue;          // keep about two tiles per SM: prefer the smaller tile
    pl.total_tiles = total;
    *out = pl;
    return true;
  }
  return false;
}
int debug_gp_pair_group_plan(const GpPairParams* ps, int n, int mode, int* v) {      // v[16]
  GpPairGroups gs{};
  gpp::PPlan pl;
  if (!plan_pair_group(ps, n, mode, &gs, &pl)) { set_error("resblock_gp group: the %d layers do not share a launch shape", n); return EV_EINVAL; }
  v[0] = pl.mt; v[1] = pl.kbg; v[2] = pl.total_tiles; v[3] = pl.rows1_pad; v[4] = pl.rows2_pad; v[5] = pl.smem_total; v[6] = pl.tmem_cols;
  for (int i = 0; i < 3; ++i) {
    v[7 + 3 * i] = i < gs.ng ? gs.g[i].K : 0;
    v[8 + 3 * i] = i < gs.ng ? gs.g[i].tiles_m : 0;
    v[9 + 3 * i] = i < gs.ng ? gs.g[i].tile0 : 0;
  }
  return EV_OK;
}
bool gp_pair_group_supported(const GpPairParams* ps, int n, int mode) {
  GpPairGroups gs{};
  gpp::PPlan pl;
  return plan_pair_group(ps, n, mode, &gs, &pl);
}
int launch_gp_pair_group(const GpPairParams* ps, int n, int mode, cudaStream_t st) {
  GpPairGroups gs{};
  gpp::PPlan pl;
  if (!plan_pair_group(ps, n, mode, &gs, &pl)) { set_error("resblock_gp group: the %d layers do not share a launch shape", n); return EV_EINVAL; }
  return dispatch_pair(ps[0], pl, gs, mode, st);
}

}  // namespace ev
