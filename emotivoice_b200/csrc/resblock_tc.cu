// One ResBlock1 layer as ONE kernel (hifigan/models.py:50-57):
//
//   out = [acc-mode]( x + c2( lrelu( c1( lrelu(x) ) ) ) ),   c1: k taps, dilation d;  c2: k taps, dilation 1;  C -> C channels
//
// i.e. the two launches `xt = conv1d_tc(lrelu(x), w1) ; out = conv1d_tc(lrelu(xt), w2) + x` of the unfused path with the
// intermediate `xt` kept on the SM: c1 accumulates in TMEM exactly as conv1d_tc.cu does, its epilogue writes
// lrelu(acc + b1) -- zero outside [0, len), the reference pads c2's input with zeros -- rounded / hi-lo split straight into a
// second shared-memory tile in the same no-swizzle K-major UMMA layout (rows 16 B apart), and c2's taps are descriptor start
// address shifts over that tile.  Algorithmic HBM traffic per layer drops from 5 activation passes (x in, xt out, xt in, x
// residual in, out) to 2; launches halve.  OPT-IN (EV_FUSE_RES=1) until it has been validated on hardware.
//
// Every reduction runs in the same order as in the unfused pair (same K granules per stage, same (channel block, tap, k8) loop,
// same epilogue arithmetic), so the result must be BITWISE equal to the two-launch path: that is the test oracle.
//
// Tile: R = 128*MT - (k-1) output rows.  c1 produces MT accumulators = 128*MT rows starting (k-1)/2 rows before the tile; c2's
// accumulator mt reads A2 rows [128*mt + j, +128) for tap j, so the last k-1 rows of the last accumulator read rows c1 never
// produced: they are garbage and are not stored (each output row depends only on its own input rows).
// TMEM: c1 accumulators in columns [0, MT*C), c2's in [MT*C, 2*MT*C)  (2*MT*C <= 512).
// Roles and pipelines as in conv1d_tc.cu; extra barriers acc1_full/acc1_empty, a2_full/a2_empty, acc2_full/acc2_empty, each
// completing once per tile (parity = tile count & 1).
#include <cstdlib>

#include "ev_common.cuh"
#include "tc_common.cuh"

namespace ev {

namespace tc {

struct RPlan {
  int mt, kbg, planes;
  int rows1_pad, rows2_pad;
  int a1_plane_bytes, a1_stage_bytes, a2_plane_bytes, a2_bytes, b_plane_bytes, b_stage_bytes;
  int a_stages, b_stages, ngroups;
  int tmem_cols;
  int R;                      // output rows per tile
  int tiles_m, total_tiles;
  int smem_total;
};

// smem map: [0,512) barriers | [512,516) tmem base | 1024: epilogue staging (8 warps x 4 KB) | A1 ring | A2 tile | B ring
inline bool make_rplan(const ResPairParams& p, int mode, int mt, int kbg, RPlan* o) {
  RPlan q;
  q.mt = mt;
  q.kbg = kbg;
  q.planes = mode == 1 ? 2 : 1;
  const int cpg = mode == 2 ? 8 : 4;
  if (2 * mt * p.C > 512) return false;
  q.tmem_cols = 32;
  while (q.tmem_cols < 2 * mt * p.C) q.tmem_cols <<= 1;
  q.R = BM * mt - (p.K - 1);
  if (q.R < BM / 2) return false;
  const int rows1 = BM * mt + (p.K - 1) * p.dil;
  q.rows1_pad = ((rows1 + 7) / 8) * 8 + 8 / kbg;        // == 8/kbg (mod 8): conflict-free 16 B stores of the producers
  const int rows2 = BM * mt + (p.K - 1);
  q.rows2_pad = ((rows2 + 7) / 8) * 8;
  q.a1_plane_bytes = kbg * q.rows1_pad * 16;
  q.a1_stage_bytes = q.planes * q.a1_plane_bytes;
  q.a2_plane_bytes = (p.C / cpg) * q.rows2_pad * 16;
  q.a2_bytes = q.planes * q.a2_plane_bytes;
  q.b_plane_bytes = kbg * p.C * 16;
  q.b_stage_bytes = q.planes * q.b_plane_bytes;
  const int budget = 227 * 1024 - 1024 - STAGING_BYTES - q.a2_bytes;
  const int n_cb = (p.C + cpg * kbg - 1) / (cpg * kbg);
  int min_b = 4;
  if (min_b > n_cb * p.K) min_b = n_cb * p.K;
  if (min_b < 2) min_b = 2;
  if (2 * q.a1_stage_bytes + min_b * q.b_stage_bytes > budget) {
    min_b = 2;
    if (2 * q.a1_stage_bytes + min_b * q.b_stage_bytes > budget) return false;
  }
  q.a_stages = 2;
  q.b_stages = min_b;
  while (q.b_stages < MAX_B_STAGES && q.b_stages < n_cb * p.K && q.a_stages * q.a1_stage_bytes + (q.b_stages + 1) * q.b_stage_bytes <= budget &&
         q.b_stages < 4) ++q.b_stages;
  while (q.a_stages < MAX_A_STAGES && q.a_stages < n_cb && (q.a_stages + 1) * q.a1_stage_bytes + q.b_stages * q.b_stage_bytes <= budget) ++q.a_stages;
  while (q.b_stages < MAX_B_STAGES && q.b_stages < 2 * n_cb * p.K && q.a_stages * q.a1_stage_bytes + (q.b_stages + 1) * q.b_stage_bytes <= budget) ++q.b_stages;
  q.ngroups = q.a_stages >= 6 ? 6 : (q.a_stages >= 3 ? 3 : (q.a_stages >= 2 ? 2 : 1));      // <= a_stages (see tc_common.cuh)
  q.tiles_m = (p.L + q.R - 1) / q.R;
  q.total_tiles = p.B * q.tiles_m;
  q.smem_total = 1024 + STAGING_BYTES + q.a_stages * q.a1_stage_bytes + q.a2_bytes + q.b_stages * q.b_stage_bytes;
  *o = q;
  return true;
}

template <int MODE, int MT, int KBG>
__global__ void __launch_bounds__(NTHREADS, 1) resblock_pair_kernel(ResPairParams p, RPlan pl) {
  constexpr bool SPLIT3 = (MODE == 1);
  constexpr bool BF16 = (MODE == 2);
  constexpr int PLANES = SPLIT3 ? 2 : 1;
  constexpr int CPG = BF16 ? 8 : 4;
  constexpr int KB = CPG * KBG;
  constexpr int GSH = (KBG == 8 ? 3 : 2);
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const int C = p.C;

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem_raw + 512);
  uint8_t* staging = smem_raw + 1024;
  uint8_t* a1_tiles = staging + STAGING_BYTES;
  uint8_t* a2_tile = a1_tiles + pl.a_stages * pl.a1_stage_bytes;
  uint8_t* b_tiles = a2_tile + pl.a2_bytes;
  const uint32_t bar_base = smem_u32(bars);
  auto a_full = [&](int s) { return bar_base + 8u * s; };
  auto a_empty = [&](int s) { return bar_base + 8u * (MAX_A_STAGES + s); };
  auto b_full = [&](int s) { return bar_base + 8u * (2 * MAX_A_STAGES + s); };
  auto b_empty = [&](int s) { return bar_base + 8u * (2 * MAX_A_STAGES + MAX_B_STAGES + s); };
  const uint32_t x0 = bar_base + 8u * (2 * MAX_A_STAGES + 2 * MAX_B_STAGES);      // 32 ring barriers = 256 B, then six singles
  const uint32_t acc1_full = x0, acc1_empty = x0 + 8, a2_full = x0 + 16, a2_empty = x0 + 24, acc2_full = x0 + 32, acc2_empty = x0 + 40;

  if (tid == 0) {
    for (int s = 0; s < pl.a_stages; ++s) { mbar_init(a_full(s), (NPWARPS / pl.ngroups) * 32); mbar_init(a_empty(s), 1); }
    for (int s = 0; s < pl.b_stages; ++s) { mbar_init(b_full(s), 1); mbar_init(b_empty(s), 1); }
    mbar_init(acc1_full, 1);
    mbar_init(acc1_empty, NEPI / 32);
    mbar_init(a2_full, NEPI);            // every epilogue thread fences its own generic-proxy writes, then arrives
    mbar_init(a2_empty, 1);
    mbar_init(acc2_full, 1);
    mbar_init(acc2_empty, NEPI / 32);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == MMA_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(pl.tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int n_cb = (C + KB - 1) / KB;
  const int h2 = (p.K - 1) / 2;                 // halo of c2 (dilation 1)
  const int h1 = h2 * p.dil;                    // halo of c1
  const int rows_a1 = BM * MT + (p.K - 1) * p.dil;

  auto decode = [&](int tile, int& b, int& t0, int& len) {
    b = tile / pl.tiles_m;
    t0 = (tile - b * pl.tiles_m) * pl.R;
    len = p.lens ? min(p.L, p.lens[b] * p.lens_mul) : p.L;
  };

  if (warp < NEPI / 32) {
    // ============================ epilogue warps ==============================================
    const int quad = warp & 3, chalf = warp >> 2;
    float* stg = reinterpret_cast<float*>(staging + warp * (32 * 32 * 4));
    const int rr = lane >> 3, cq = lane & 7;
    const float slope = p.slope;
    int tile_cnt = 0;
    for (int tile = blockIdx.x; tile < pl.total_tiles; tile += gridDim.x) {
      int b, t0, len;
      decode(tile, b, t0, len);
      float* ob = p.out + (size_t)b * p.L * C;
      const float* rb = p.x + (size_t)b * p.L * C;       // the residual is the layer's own input
      if (t0 >= len) {   // padding tile: the batch-invariant contract stores zeros
        for (int mt = 0; mt < MT; ++mt)
          for (int c = chalf * 32; c < C; c += 64)
            for (int it = 0; it < 8; ++it) {
              const int rl = mt * BM + quad * 32 + it * 4 + rr;
              const int row = t0 + rl;
              if (rl < pl.R && row < p.L) *reinterpret_cast<float4*>(ob + (size_t)row * C + c + cq * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        continue;
      }
      const uint32_t par = tile_cnt & 1;
      // ---- epilogue 1: c1 accumulators -> + b1 -> lrelu -> zero outside the sequence -> operand tile of c2 ----
      mbar_wait(a2_empty, par ^ 1);          // c2 of the previous tile has finished reading the tile
      mbar_wait(acc1_full, par);
      tc_fence_after();
#pragma unroll 1
      for (int mt = 0; mt < MT; ++mt) {
        const int r = mt * BM + quad * 32 + lane;            // row of the c1 output tile == TMEM lane
        const int t = t0 - h2 + r;                           // its time index
        const bool valid = t >= 0 && t < len;
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(mt * C);
#pragma unroll 1
        for (int c = chalf * 32; c < C; c += 64) {
          float v[32];
          tmem_ld32(taddr + (uint32_t)c, 32, v);
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            float u = (v[i] + __ldg(p.b1 + c + i)) + 0.f;      // "+ 0" = the unfused c1 epilogue's (absent) residual: keeps -0 -> +0
            u = u > 0.f ? u : u * slope;
            v[i] = valid ? u : 0.f;
          }
          if (BF16) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              uint4 w;
              w.x = pack_bf16(v[8 * q + 0], v[8 * q + 1]); w.y = pack_bf16(v[8 * q + 2], v[8 * q + 3]);
              w.z = pack_bf16(v[8 * q + 4], v[8 * q + 5]); w.w = pack_bf16(v[8 * q + 6], v[8 * q + 7]);
              *reinterpret_cast<uint4*>(a2_tile + ((size_t)(c / 8 + q) * pl.rows2_pad + r) * 16) = w;
            }
          } else {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const float4 h = make_float4(to_tf32(v[4 * q]), to_tf32(v[4 * q + 1]), to_tf32(v[4 * q + 2]), to_tf32(v[4 * q + 3]));
              uint8_t* d = a2_tile + ((size_t)(c / 4 + q) * pl.rows2_pad + r) * 16;
              *reinterpret_cast<float4*>(d) = h;
              if (SPLIT3) {
                const float4 l = make_float4(to_tf32(v[4 * q] - h.x), to_tf32(v[4 * q + 1] - h.y), to_tf32(v[4 * q + 2] - h.z),
                                             to_tf32(v[4 * q + 3] - h.w));
                *reinterpret_cast<float4*>(d + pl.a2_plane_bytes) = l;
              }
            }
          }
        }
      }
      fence_proxy_async();                   // generic-proxy smem writes -> visible to the tensor core
      mbar_arrive(a2_full);
      tc_fence_before();                     // all TMEM reads of the c1 accumulators are complete (wait::ld inside tmem_ld32)
      __syncwarp();
      if (lane == 0) mbar_arrive(acc1_empty);

      // ---- epilogue 2: c2 accumulators -> + b2 + residual -> [accumulate modes] -> global ----
      bool waited = false;
#pragma unroll 1
      for (int mt = 0; mt < MT; ++mt) {
        const int rl_base = mt * BM + quad * 32;
        const int row_base = t0 + rl_base;
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(MT * C + mt * C);
#pragma unroll 1
        for (int c = chalf * 32; c < C; c += 64) {
          float4 rq[8], oq[8];
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int rl = rl_base + it * 4 + rr;
            const int row = row_base + it * 4 + rr;
            const size_t off = (size_t)row * C + c + cq * 4;
            const bool in = rl < pl.R && row < len;
            rq[it] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (in) rq[it] = *reinterpret_cast<const float4*>(rb + off);
            if (p.acc != EV_ACC_STORE) {
              oq[it] = make_float4(0.f, 0.f, 0.f, 0.f);
              if (in) oq[it] = *reinterpret_cast<const float4*>(ob + off);
            }
          }
          const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.b2 + c + cq * 4));
          if (!waited) {
            mbar_wait(acc2_full, par);
            tc_fence_after();
            waited = true;
          }
          float v[32];
          tmem_ld32(taddr + (uint32_t)c, 32, v);
#pragma unroll
          for (int q = 0; q < 8; ++q)
            *reinterpret_cast<float4*>(stg + lane * 32 + ((q ^ (lane & 7)) << 2)) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
          __syncwarp();
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int r = it * 4 + rr;
            const int rl = rl_base + r;
            const int row = row_base + r;
            float4 o = *reinterpret_cast<const float4*>(stg + r * 32 + ((cq ^ (r & 7)) << 2));
            if (row < len) {
              o.x += b4.x; o.y += b4.y; o.z += b4.z; o.w += b4.w;
              o.x += rq[it].x; o.y += rq[it].y; o.z += rq[it].z; o.w += rq[it].w;
              if (p.acc != EV_ACC_STORE) {
                o.x += oq[it].x; o.y += oq[it].y; o.z += oq[it].z; o.w += oq[it].w;
                if (p.acc == EV_ACC_ADD_DIV) { o.x /= p.div; o.y /= p.div; o.z /= p.div; o.w /= p.div; }
              }
            } else {
              o = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            if (rl < pl.R && row < p.L) *reinterpret_cast<float4*>(ob + (size_t)row * C + c + cq * 4) = o;
          }
          __syncwarp();
        }
      }
      if (!waited) {     // C == 32: the second half of the epilogue warps has no columns but follows the phases
        mbar_wait(acc2_full, par);
        tc_fence_after();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc2_empty);
      ++tile_cnt;
    }
  } else if (warp < MMA_WARP) {
    // ============================ A1 producers (the x tile of c1) ================================
    const int pwarp = warp - NEPI / 32;
    const int wpg = NPWARPS / pl.ngroups;
    const int grp = pwarp / wpg;
    const int gt = (pwarp - grp * wpg) * 32 + lane;
    const int GT = wpg * 32;
    const float slope = p.slope;
    const int total = rows_a1 * KBG;
    int a_cnt = 0;
    for (int tile = blockIdx.x; tile < pl.total_tiles; tile += gridDim.x) {
      int b, t0, len;
      decode(tile, b, t0, len);
      if (t0 >= len) continue;
      const float* __restrict__ xb = p.x + (size_t)b * p.L * C;
      for (int cb = 0; cb < n_cb; ++cb, ++a_cnt) {
        if (a_cnt % pl.ngroups != grp) continue;
        const int s = a_cnt % pl.a_stages;
        const int c0 = cb * KB;
        const int ngran = min(KB, C - c0) / CPG;
        uint8_t* dst = a1_tiles + s * pl.a1_stage_bytes;
        constexpr int ALD = BF16 ? A_LD / 2 : A_LD;
        constexpr int NW = BF16 ? 2 : 1;
        for (int base = 0; base < total; base += GT * ALD) {
          float4 v[ALD * NW];
#pragma unroll
          for (int u = 0; u < ALD; ++u) {
            const int idx = base + u * GT + gt;
            const int r = idx >> GSH, g = idx & (KBG - 1);
            const int row = t0 - h2 - h1 + r;
            const bool ok = idx < total && g < ngran && row >= 0 && row < len;
            const float* src = xb + (size_t)row * C + c0 + g * CPG;
#pragma unroll
            for (int w = 0; w < NW; ++w)
              v[u * NW + w] = ok ? __ldg(reinterpret_cast<const float4*>(src + 4 * w)) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
          if (base == 0) mbar_wait(a_empty(s), ((a_cnt / pl.a_stages) & 1) ^ 1);
#pragma unroll
          for (int u = 0; u < ALD; ++u) {
            const int idx = base + u * GT + gt;
            const int r = idx >> GSH, g = idx & (KBG - 1);
            if (idx < total && g < ngran) {
#pragma unroll
              for (int w = 0; w < NW; ++w) {
                float4& t = v[u * NW + w];
                t.x = t.x > 0.f ? t.x : t.x * slope;
                t.y = t.y > 0.f ? t.y : t.y * slope;
                t.z = t.z > 0.f ? t.z : t.z * slope;
                t.w = t.w > 0.f ? t.w : t.w * slope;
              }
              uint8_t* d = dst + ((size_t)g * pl.rows1_pad + r) * 16;
              if (BF16) {
                const float4 t0v = v[u * NW], t1v = v[u * NW + NW - 1];
                uint4 q;
                q.x = pack_bf16(t0v.x, t0v.y); q.y = pack_bf16(t0v.z, t0v.w);
                q.z = pack_bf16(t1v.x, t1v.y); q.w = pack_bf16(t1v.z, t1v.w);
                *reinterpret_cast<uint4*>(d) = q;
              } else {
                const float4 t = v[u * NW];
                const float4 h = make_float4(to_tf32(t.x), to_tf32(t.y), to_tf32(t.z), to_tf32(t.w));
                *reinterpret_cast<float4*>(d) = h;
                if (SPLIT3) {
                  const float4 l = make_float4(to_tf32(t.x - h.x), to_tf32(t.y - h.y), to_tf32(t.z - h.z), to_tf32(t.w - h.w));
                  *reinterpret_cast<float4*>(d + pl.a1_plane_bytes) = l;
                }
              }
            }
          }
        }
        fence_proxy_async();
        mbar_arrive(a_full(s));
      }
    }
  } else if (warp == MMA_WARP) {
    // ============================ MMA issuer =====================================================
    if (lane == 0) {
      const uint32_t a1_lbo = (uint32_t)pl.rows1_pad * 16u, a2_lbo = (uint32_t)pl.rows2_pad * 16u, b_lbo = (uint32_t)C * 16u;
      const uint32_t fmt = BF16 ? 1u : 2u;
      const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(C >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      const uint32_t a2_addr = smem_u32(a2_tile);
      int a_cnt = 0, b_cnt = 0, tile_cnt = 0;
      for (int tile = blockIdx.x; tile < pl.total_tiles; tile += gridDim.x) {
        int b, t0, len;
        decode(tile, b, t0, len);
        if (t0 >= len) continue;
        const uint32_t par = tile_cnt & 1;
        // ---- c1: dilated taps over the staged x tile -> accumulators [0, MT*C) ----
        mbar_wait(acc1_empty, par ^ 1);
        tc_fence_after();
        for (int cb = 0; cb < n_cb; ++cb, ++a_cnt) {
          const int sa = a_cnt % pl.a_stages;
          const int nk8 = min(KB, C - cb * KB) / (2 * CPG);
          mbar_wait(a_full(sa), (a_cnt / pl.a_stages) & 1);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(a1_tiles + sa * pl.a1_stage_bytes);
          for (int j = 0; j < p.K; ++j, ++b_cnt) {
            const int sb = b_cnt % pl.b_stages;
            mbar_wait(b_full(sb), (b_cnt / pl.b_stages) & 1);
            tc_fence_after();
            const uint32_t b_addr = smem_u32(b_tiles + sb * pl.b_stage_bytes);
            for (int k8 = 0; k8 < nk8; ++k8) {
              const uint32_t b_off = (uint32_t)(2 * k8) * b_lbo;
              const uint64_t b_hi = make_desc(b_addr + b_off, b_lbo, 128u);
              const uint64_t b_lo = make_desc(b_addr + pl.b_plane_bytes + b_off, b_lbo, 128u);
              const uint32_t first = (cb | j | k8) != 0 ? 1u : 0u;
#pragma unroll
              for (int mt = 0; mt < MT; ++mt) {
                const uint32_t a_off = (uint32_t)((2 * k8) * pl.rows1_pad + mt * BM + j * p.dil) * 16u;
                const uint64_t a_hi = make_desc(a_addr + a_off, a1_lbo, 128u);
                const uint32_t d = tmem_base + (uint32_t)(mt * C);
                if (SPLIT3) {
                  const uint64_t a_lo = make_desc(a_addr + pl.a1_plane_bytes + a_off, a1_lbo, 128u);
                  umma_tf32(d, a_lo, b_hi, idesc, first);
                  umma_tf32(d, a_hi, b_lo, idesc, 1u);
                  umma_tf32(d, a_hi, b_hi, idesc, 1u);
                } else if (BF16) {
                  umma_bf16(d, a_hi, b_hi, idesc, first);
                } else {
                  umma_tf32(d, a_hi, b_hi, idesc, first);
                }
              }
            }
            umma_commit(b_empty(sb));
          }
          umma_commit(a_empty(sa));
        }
        umma_commit(acc1_full);
        // ---- c2: taps (dilation 1) over the tile written by epilogue 1 -> accumulators [MT*C, 2*MT*C) ----
        mbar_wait(acc2_empty, par ^ 1);
        mbar_wait(a2_full, par);
        tc_fence_after();
        for (int cb = 0; cb < n_cb; ++cb) {
          const int nk8 = min(KB, C - cb * KB) / (2 * CPG);
          for (int j = 0; j < p.K; ++j, ++b_cnt) {
            const int sb = b_cnt % pl.b_stages;
            mbar_wait(b_full(sb), (b_cnt / pl.b_stages) & 1);
            tc_fence_after();
            const uint32_t b_addr = smem_u32(b_tiles + sb * pl.b_stage_bytes);
            for (int k8 = 0; k8 < nk8; ++k8) {
              const uint32_t b_off = (uint32_t)(2 * k8) * b_lbo;
              const uint64_t b_hi = make_desc(b_addr + b_off, b_lbo, 128u);
              const uint64_t b_lo = make_desc(b_addr + pl.b_plane_bytes + b_off, b_lbo, 128u);
              const uint32_t first = (cb | j | k8) != 0 ? 1u : 0u;
#pragma unroll
              for (int mt = 0; mt < MT; ++mt) {
                const uint32_t a_off = (uint32_t)((cb * KBG + 2 * k8) * pl.rows2_pad + mt * BM + j) * 16u;
                const uint64_t a_hi = make_desc(a2_addr + a_off, a2_lbo, 128u);
                const uint32_t d = tmem_base + (uint32_t)(MT * C + mt * C);
                if (SPLIT3) {
                  const uint64_t a_lo = make_desc(a2_addr + pl.a2_plane_bytes + a_off, a2_lbo, 128u);
                  umma_tf32(d, a_lo, b_hi, idesc, first);
                  umma_tf32(d, a_hi, b_lo, idesc, 1u);
                  umma_tf32(d, a_hi, b_hi, idesc, 1u);
                } else if (BF16) {
                  umma_bf16(d, a_hi, b_hi, idesc, first);
                } else {
                  umma_tf32(d, a_hi, b_hi, idesc, first);
                }
              }
            }
            umma_commit(b_empty(sb));
          }
        }
        umma_commit(a2_empty);            // the operand tile may be overwritten by the next tile's epilogue 1
        umma_commit(acc2_full);
        ++tile_cnt;
      }
    }
    __syncwarp();
  } else {
    // ============================ weight loader: W1 then W2 of every tile, one ring =================
    if (lane == 0) {
      const int cin4 = C / CPG;
      const size_t plane = (size_t)p.K * C * C;         // floats per plane (tf32 layouts; unused in bf16 mode)
      int b_cnt = 0;
      for (int tile = blockIdx.x; tile < pl.total_tiles; tile += gridDim.x) {
        int b, t0, len;
        decode(tile, b, t0, len);
        if (t0 >= len) continue;
        for (int conv = 0; conv < 2; ++conv) {
          const float* wt = conv == 0 ? p.w1 : p.w2;
          for (int cb = 0; cb < n_cb; ++cb) {
            const int ngran = min(KB, C - cb * KB) / CPG;
            for (int j = 0; j < p.K; ++j, ++b_cnt) {
              const int sb = b_cnt % pl.b_stages;
              mbar_wait(b_empty(sb), ((b_cnt / pl.b_stages) & 1) ^ 1);
              mbar_expect_tx(b_full(sb), (uint32_t)(PLANES * ngran * C * 16));
              const uint32_t dst = smem_u32(b_tiles + sb * pl.b_stage_bytes);
              const float* src = wt + ((size_t)j * cin4 + (size_t)cb * KBG) * C * 4;      // one N tile (C <= 128): granules adjacent
              bulk_g2s(dst, src, (uint32_t)(ngran * C * 16), b_full(sb));
              if (SPLIT3) bulk_g2s(dst + (uint32_t)pl.b_plane_bytes, src + plane, (uint32_t)(ngran * C * 16), b_full(sb));
            }
          }
        }
      }
    }
    __syncwarp();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(pl.tmem_cols));
  }
}

}  // namespace tc

static int rp_sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

template <int MODE, int MT, int KBG>
static int launch_rp_variant(const ResPairParams& p, const tc::RPlan& pl, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(tc::resblock_pair_kernel<MODE, MT, KBG>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    attr_set = true;
  }
  const int grid = pl.total_tiles < rp_sm_count() ? pl.total_tiles : rp_sm_count();
  tc::resblock_pair_kernel<MODE, MT, KBG><<<grid, tc::NTHREADS, pl.smem_total, st>>>(p, pl);
  EV_CUDA_LAUNCH_CHECK("resblock_pair_kernel");
  return EV_OK;
}

template <int MODE, int KBG>
static int launch_rp_mt(const ResPairParams& p, const tc::RPlan& pl, cudaStream_t st) {
  if (pl.mt == 4) return launch_rp_variant<MODE, 4, KBG>(p, pl, st);
  if (pl.mt == 2) return launch_rp_variant<MODE, 2, KBG>(p, pl, st);
  return launch_rp_variant<MODE, 1, KBG>(p, pl, st);
}

// K granules per stage of the two unfused convolutions (conv1d_tc.cu: a function of the layer shape only); the fused kernel
// must use the same value for both or its reduction order -- and the bits -- would differ from the two-launch path.
static int pair_kbg(const ResPairParams& p, int mode) {
  ConvParams c;
  c.x = nullptr; c.w = nullptr; c.bias = nullptr; c.res = nullptr; c.out = nullptr; c.bias_bs = 0;
  c.B = p.B; c.L = p.L; c.Cin = p.C; c.Cout = p.C; c.K = p.K; c.lens = nullptr; c.lens_mul = 1;
  c.in_act = EV_ACT_LRELU; c.in_slope = p.slope; c.out_act = EV_ACT_NONE; c.acc = EV_ACC_STORE; c.div = 1.f;
  c.dil = p.dil;
  const int k1 = tc_shape_kbg(c, mode);
  c.dil = 1;
  const int k2 = tc_shape_kbg(c, mode);
  return k1 == k2 ? k1 : 0;
}

static bool plan_resblock_pair(const ResPairParams& p, int mode, tc::RPlan* out) {
  if (p.B <= 0 || p.L <= 0 || !(p.C == 32 || p.C == 64 || p.C == 128)) return false;
  if (mode == 2 && p.C % 16) return false;
  if (p.K < 1 || !(p.K & 1) || p.K > 15 || p.dil < 1) return false;
  if (p.out == p.x) return false;                       // other CTAs read halo rows of x while this one writes out
  const int kbg = pair_kbg(p, mode);
  if (kbg == 0) return false;
  const long long tiles128 = (long long)((p.L + tc::BM - 1) / tc::BM) * p.B;
  static const int mt_thresh = [] { const char* e = getenv("EV_TC_MT_TILES"); return (e && *e) ? atoi(e) : 120; }();
  int mt = tiles128 >= 4 * mt_thresh ? 4 : (tiles128 >= 2 * mt_thresh ? 2 : 1);
  for (;; mt >>= 1) {
    if (tc::make_rplan(p, mode, mt, kbg, out)) return true;
    if (mt == 1) return false;
  }
}

bool resblock_pair_supported(const ResPairParams& p, int mode) {
  tc::RPlan pl;
  return plan_resblock_pair(p, mode, &pl);
}

int debug_resblock_plan(const ResPairParams& p, int mode, int* v) {
  tc::RPlan pl;
  if (!plan_resblock_pair(p, mode, &pl)) { set_error("resblock_pair: unsupported shape C=%d K=%d dil=%d mode=%d", p.C, p.K, p.dil, mode); return EV_EINVAL; }
  v[0] = pl.mt; v[1] = pl.kbg; v[2] = pl.a_stages; v[3] = pl.b_stages; v[4] = pl.ngroups; v[5] = pl.tmem_cols; v[6] = pl.smem_total;
  v[7] = pl.total_tiles; v[8] = pl.R; v[9] = pl.rows1_pad; v[10] = pl.rows2_pad;
  return EV_OK;
}

// w1 / w2 in the tensor-core layout of conv1d_tc (one N tile: C <= 128); mode 0: 1xTF32, 1: 3xTF32, 2: bf16.
int launch_resblock_pair(const ResPairParams& p, int mode, cudaStream_t st) {
  tc::RPlan pl;
  if (!plan_resblock_pair(p, mode, &pl)) { set_error("resblock_pair: unsupported shape C=%d K=%d dil=%d mode=%d", p.C, p.K, p.dil, mode); return EV_EINVAL; }
  if (mode == 1) return launch_rp_mt<1, 4>(p, pl, st);
  if (mode == 2) return pl.kbg == 8 ? launch_rp_mt<2, 8>(p, pl, st) : launch_rp_mt<2, 4>(p, pl, st);
  return pl.kbg == 8 ? launch_rp_mt<0, 8>(p, pl, st) : launch_rp_mt<0, 4>(p, pl, st);
}

}  // namespace ev
