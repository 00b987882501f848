// Time-major 1-D convolution as an implicit GEMM on the 5th-gen tensor cores (sm_100a):
// tcgen05.mma (kind::tf32, fp32 accumulate in TMEM), operands staged in shared memory, weights
// streamed by bulk async copies (cp.async.bulk -> mbarrier complete_tx), accumulators read back
// with tcgen05.ld by dedicated epilogue warps.  Same contract as conv1d_tm.cu (see ev_common.cuh):
//
//   out[b,t,co] = epi( bias[co] + sum_j sum_ci w[j][ci][co] * act_in( x[b, t + (j-(K-1)/2)*dil, ci] ) )
//
// GEMM view: M = time (128 rows per accumulator, MT accumulators per tile), N = C_out tile (<= 128),
// K = taps x C_in.
//
// * ONE activation fetch per tile for all k taps: the A operand lives in shared memory in the
//   no-swizzle K-major UMMA layout with the 8-row-group stride (SBO) set to 128 B, i.e. element
//   (row r, 16-byte K-granule g) sits at  A + (g * rows_pad + r) * 16 bytes.  Consecutive rows are
//   16 B apart for the whole tile, so tap j of a dilated convolution is the same staged tile with the
//   descriptor start address advanced by j*dil rows, and accumulator mt by 128*mt rows.  The producer
//   warps apply the LeakyReLU prologue, the zero padding at the sequence ends, the tf32 rounding /
//   hi-lo split and the layout change while staging, so activations stay plain fp32 time-major in HBM
//   and no tensor map is needed.
// * One weight tile from L2 feeds MT accumulators (MT x fewer weight bytes per output row).
// * Persistent CTAs (one per SM) loop over tiles; the accumulators are double buffered in TMEM so
//   the epilogue of tile i (TMEM -> registers -> swizzled smem -> fully coalesced 128-byte global
//   rows, residual/accumulate operands prefetched) overlaps the main loop of tile i+1.
//
// Roles (512 threads): warps 0-7 epilogue (warp e <-> TMEM lanes 32(e%4).., alternate 32-column
// chunks), warps 8-13 stage A, warp 14 allocates TMEM and its elected lane issues every tcgen05.mma,
// warp 15's elected lane streams the weight tiles.  mbarrier pipelines: A ring (a_full/a_empty), B ring (b_full/b_empty,
// released by tcgen05.commit), accumulators (acc_full/acc_empty).
#include <cstdio>
#include <cstdlib>

#include "ev_common.cuh"
#include "tc_common.cuh"

namespace ev {

namespace tc {

struct Plan {
  int BN, mt, kbg, planes;
  int rows_pad;        // staged rows per A granule; == 8/kbg (mod 8) -> conflict-free 16 B stores
  int a_plane_bytes, b_plane_bytes, a_stage_bytes, b_stage_bytes;
  int a_stages, b_stages;
  int ngroups;         // producer groups: largest of {6,3,2,1} that is <= a_stages
  int ksplit;          // K-split factor S: S CTAs share one output tile, each reducing a slice of the C_in blocks
                       // into a private partial buffer; splitk_reduce_kernel sums them in a fixed order
  int tmem_cols;
  int tiles_m, tiles_n, total_tiles;
  int smem_total;
};

// smem map: [0,288) barriers | [512,516) tmem base | 1024: epilogue staging (8 warps x 4 KB) | A ring | B ring
__host__ __device__ inline bool make_plan(const ConvParams& p, int mode, int BN, int mt, int kbg, int min_b_stages, Plan* o, int b_target = 4) {
  Plan q;
  q.planes = (mode == 1 || mode == 3) ? 2 : 1;
  const int cpg = mode >= 2 ? 8 : 4;      // channels per 16-byte operand granule (bf16 : tf32)
  q.kbg = kbg;
  q.mt = mt;
  q.BN = BN;
  if (2 * mt * q.BN > 512) return false;
  q.tmem_cols = 32;
  while (q.tmem_cols < 2 * mt * q.BN) q.tmem_cols <<= 1;
  const int rows = BM * mt + (p.K - 1) * p.dil;
  q.rows_pad = ((rows + 7) / 8) * 8 + 8 / q.kbg;
  q.a_plane_bytes = q.kbg * q.rows_pad * 16;
  q.b_plane_bytes = q.kbg * q.BN * 16;
  q.a_stage_bytes = q.planes * q.a_plane_bytes;
  q.b_stage_bytes = q.planes * q.b_plane_bytes;
  const int budget = 227 * 1024 - 1024 - STAGING_BYTES;
  const int n_cb = (p.Cin + cpg * q.kbg - 1) / (cpg * q.kbg);
  // at least 2 + 2 stages; then grow the weight ring first (it turns over K times per A stage)
  if (min_b_stages > n_cb * p.K) min_b_stages = n_cb * p.K;
  if (min_b_stages < 2) min_b_stages = 2;
  if (2 * q.a_stage_bytes + min_b_stages * q.b_stage_bytes > budget) return false;
  q.a_stages = 2;
  q.b_stages = 2;
  while (q.b_stages < MAX_B_STAGES && q.b_stages < n_cb * p.K &&
         q.a_stages * q.a_stage_bytes + (q.b_stages + 1) * q.b_stage_bytes <= budget && q.b_stages < b_target) ++q.b_stages;
  while (q.a_stages < MAX_A_STAGES && q.a_stages < n_cb &&
         (q.a_stages + 1) * q.a_stage_bytes + q.b_stages * q.b_stage_bytes <= budget && q.a_stages < 6) ++q.a_stages;
  while (q.b_stages < MAX_B_STAGES && q.b_stages < n_cb * p.K &&
         q.a_stages * q.a_stage_bytes + (q.b_stages + 1) * q.b_stage_bytes <= budget) ++q.b_stages;
  while (q.a_stages < MAX_A_STAGES && q.a_stages < n_cb &&
         (q.a_stages + 1) * q.a_stage_bytes + q.b_stages * q.b_stage_bytes <= budget) ++q.a_stages;
  q.ngroups = q.a_stages >= 6 ? 6 : (q.a_stages >= 3 ? 3 : (q.a_stages >= 2 ? 2 : 1));
  q.tiles_m = (p.L + BM * mt - 1) / (BM * mt);
  q.tiles_n = (p.Cout + q.BN - 1) / q.BN;
  q.ksplit = 1;
  q.total_tiles = p.B * q.tiles_m * q.tiles_n;
  q.smem_total = 1024 + STAGING_BYTES + q.a_stages * q.a_stage_bytes + q.b_stages * q.b_stage_bytes;
  *o = q;
  return true;
}

// out[e..e+3] = epi( sum_z partial[z][e..e+3] ): the slices added in the fixed order z = 0..S-1 starting from zero (deterministic,
// batch invariant), then bias / activation / residual / accumulate exactly like the fused epilogue; rows >= len are zeros.
__device__ __forceinline__ void splitk_reduce_store(const ConvParams& p, int S, size_t per, size_t e, int b, int row, int col, int len) {
  float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
  if (row < len) {
    for (int z = 0; z < S; ++z) {
      const float4 v = *reinterpret_cast<const float4*>(p.splitk_ws + (size_t)z * per + e);
      o.x += v.x; o.y += v.y; o.z += v.z; o.w += v.w;
    }
    if (p.bias) {
      const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + (size_t)b * p.bias_bs + col));
      o.x += b4.x; o.y += b4.y; o.z += b4.z; o.w += b4.w;
    }
    if (p.out_act != EV_ACT_NONE) {
      o.x = act_apply(o.x, p.out_act, 0.f); o.y = act_apply(o.y, p.out_act, 0.f);
      o.z = act_apply(o.z, p.out_act, 0.f); o.w = act_apply(o.w, p.out_act, 0.f);
    }
    if (p.res) {
      const float4 r4 = *reinterpret_cast<const float4*>(p.res + e);
      o.x += r4.x; o.y += r4.y; o.z += r4.z; o.w += r4.w;
    }
    if (p.acc != EV_ACC_STORE) {
      const float4 q4 = *reinterpret_cast<const float4*>(p.out + e);
      o.x += q4.x; o.y += q4.y; o.z += q4.z; o.w += q4.w;
      if (p.acc == EV_ACC_ADD_DIV) { o.x /= p.div; o.y /= p.div; o.z /= p.div; o.w /= p.div; }
    }
  }
  *reinterpret_cast<float4*>(p.out + e) = o;
}

// MODE 0: one tf32 MMA per K step (operands rounded to nearest tf32).
// MODE 2: bf16 operands (rounded to nearest even by the producers / the host), kind::f16, 8 channels per granule.
// MODE 3: "bf16x3" fp32-class emulation: x = hi + lo with hi = bf16(x), lo = bf16(x - hi) (16 significant bits), three kind::f16 MMAs per
//                 K = 16 step (a_lo*b_hi + a_hi*b_lo + a_hi*b_hi): ~1e-5 relative at half the tensor-core / weight-stream cost of 3xTF32.
// MODE 1: "3xTF32" fp32 emulation: x = hi + lo with hi = tf32(x), lo = tf32(x - hi);
//                 a*b ~= a_lo*b_hi + a_hi*b_lo + a_hi*b_hi (the dropped lo*lo term is 2^-22 relative),
//                 three MMAs per K step into the same fp32 TMEM accumulator.  Weights arrive pre-split
//                 (two planes, packing.to_tc_layout); activations are split by the producer warps.
// MT: 128-row accumulators per tile.  KBG: 16-byte K granules (4 tf32 or 8 bf16 channels each) per pipeline stage.
// PDLM: programmatic dependent launch mode (EV_PDL): 0 = plain launch (no extra instructions),
//       1 = convolutions only, 2 = every kernel of the engine launches this way (see the note after the set-up below).
template <int MODE, int MT, int KBG, int PDLM>
__global__ void __launch_bounds__(NTHREADS, 1) conv1d_tc_kernel(ConvParams p, Plan pl) {
  constexpr bool SPLIT3 = (MODE == 1);
  constexpr bool X3B = (MODE == 3);       // "bf16x3": fp32 operands split into bf16 hi + lo planes, three kind::f16 MMAs per K = 16 step
  constexpr bool BF16 = (MODE == 2) || X3B;      // the staged operands are bf16 (8 channels per granule)
  constexpr int PLANES = (SPLIT3 || X3B) ? 2 : 1;
  constexpr int CPG = BF16 ? 8 : 4;       // channels per 16-byte granule
  constexpr int KB = CPG * KBG;
  constexpr int GSH = (KBG == 8 ? 3 : 2);
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const int BN = pl.BN;

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem_raw + 512);   // barriers occupy [0, 8*(2*MAX_A+2*MAX_B+4)) = 288 B
  uint8_t* staging = smem_raw + 1024;
  uint8_t* a_tiles = staging + STAGING_BYTES;
  uint8_t* b_tiles = a_tiles + pl.a_stages * pl.a_stage_bytes;
  const uint32_t bar_base = smem_u32(bars);
  auto a_full = [&](int s) { return bar_base + 8u * s; };
  auto a_empty = [&](int s) { return bar_base + 8u * (MAX_A_STAGES + s); };
  auto b_full = [&](int s) { return bar_base + 8u * (2 * MAX_A_STAGES + s); };
  auto b_empty = [&](int s) { return bar_base + 8u * (2 * MAX_A_STAGES + MAX_B_STAGES + s); };
  auto acc_full = [&](int s) { return bar_base + 8u * (2 * MAX_A_STAGES + 2 * MAX_B_STAGES + s); };
  auto acc_empty = [&](int s) { return bar_base + 8u * (2 * MAX_A_STAGES + 2 * MAX_B_STAGES + 2 + s); };

  if (tid == 0) {
    for (int s = 0; s < pl.a_stages; ++s) { mbar_init(a_full(s), (NPWARPS / pl.ngroups) * 32); mbar_init(a_empty(s), 1); }
    for (int s = 0; s < pl.b_stages; ++s) { mbar_init(b_full(s), 1); mbar_init(b_empty(s), 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(acc_full(s), 1); mbar_init(acc_empty(s), NEPI / 32); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == MMA_WARP) {   // TMEM allocation by one full warp; the same warp frees it
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(pl.tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // Programmatic dependent launch (EV_PDL, default 2).  The grid is persistent (<= one CTA per SM, all resident), so it lets
  // the NEXT launch in the stream start as soon as SMs free up: that kernel's CTAs run their set-up (barriers, TMEM) while
  // this grid's tail is still running.  Everything that touches activations (producers: x; epilogue: res / out / split-K
  // partials) first executes griddepcontrol.wait, which returns once the preceding grid has completed and its writes are
  // visible.  PDLM == 1 (only the convolutions launch this way, so the launch before a convolution's predecessor has fully
  // completed): the loader and the MMA issuer, which read only weights and p.lens, do not wait and the first weight stages
  // are prefetched under the predecessor's tail.  PDLM == 2 (every kernel launches this way): p.lens may come from a grid that
  // is still running TWO launches upstream (validate_inputs_kernel writes the int32 lengths, LayerNorm starts early and waits,
  // this kernel starts early too) -- a role that decoded tiles from stale lengths would walk a different tile sequence than
  // the others and the pipeline would deadlock (seen once: a bench run whose batches had different lengths hung).  Every role waits.
  if (PDLM) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  const int n_cb = (p.Cin + KB - 1) / KB;
  const int halo = ((p.K - 1) / 2) * p.dil;
  const int rows_a = BM * MT + (p.K - 1) * p.dil;
  const int tiles_per_b = pl.tiles_m * pl.tiles_n;

  // tile -> (b, t0, n0, nt, len); every role walks the same sequence
  int z_cur = 0;     // K-split slice of the tile most recently decoded by this thread
  auto decode = [&](int tile, int& b, int& t0, int& n0, int& nt, int& len) {
    z_cur = tile % pl.ksplit;
    tile /= pl.ksplit;
    b = tile / tiles_per_b;
    const int r = tile - b * tiles_per_b;
    const int tm = r / pl.tiles_n, tn = r - tm * pl.tiles_n;
    t0 = tm * (BM * MT);
    n0 = tn * BN;
    nt = min(BN, p.Cout - n0);
    len = p.lens ? min(p.L, p.lens[b] * p.lens_mul) : p.L;
  };

  if (warp < NEPI / 32) {
    // ============================ epilogue warps ==============================================
    if (PDLM) asm volatile("griddepcontrol.wait;" ::: "memory");
    const int quad = warp & 3, chalf = warp >> 2;
    float* stg = reinterpret_cast<float*>(staging + warp * (32 * 32 * 4));
    const int rr = lane >> 3, cq = lane & 7;         // coalesced phase: 4 rows x 8 float4 per instruction
    const bool split = pl.ksplit > 1;           // K-split: raw partial sums, the fused epilogue runs in the reduce kernel
    const bool has_res = p.res != nullptr && !split;
    const int oact = split ? EV_ACT_NONE : p.out_act, accm = split ? EV_ACC_STORE : p.acc;
    int tile_cnt = 0;
    for (int tile = blockIdx.x; tile < pl.total_tiles; tile += gridDim.x) {
      int b, t0, n0, nt, len;
      decode(tile, b, t0, n0, nt, len);
      float* ob = (split ? p.splitk_ws + (size_t)z_cur * p.B * p.L * p.Cout : p.out) + (size_t)b * p.L * p.Cout;   // may alias p.res
      const float* rb = has_res ? p.res + (size_t)b * p.L * p.Cout : nullptr;
      if (t0 >= len) {   // padding tile: the batch-invariant contract stores zeros (no MMA work was issued)
        for (int mt = 0; mt < MT; ++mt)
          for (int c = chalf * 32; c < nt; c += 64)
            for (int it = 0; it < 8; ++it) {
              const int row = t0 + mt * BM + quad * 32 + it * 4 + rr;
              if (row < p.L && c + cq * 4 < nt)
                *reinterpret_cast<float4*>(ob + (size_t)row * p.Cout + n0 + c + cq * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        continue;
      }
      const int buf = tile_cnt & 1;
      const float* __restrict__ bias = (p.bias && !split) ? p.bias + (size_t)b * p.bias_bs + n0 : nullptr;
      bool waited = false;
#pragma unroll 1
      for (int mt = 0; mt < MT; ++mt) {
        const int row_base = t0 + mt * BM + quad * 32;
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(buf * MT * BN + mt * BN);
#pragma unroll 1
        for (int c = chalf * 32; c < nt; c += 64) {
          const int cw = min(32, nt - c);
          const bool col_ok = cq * 4 < cw;
          // residual / accumulate operands and the bias: coalesced loads issued before the accumulator is needed
          float4 rq[8], oq[8];
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int row = row_base + it * 4 + rr;
            const size_t off = (size_t)row * p.Cout + n0 + c + cq * 4;
            rq[it] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (has_res && col_ok && row < len) rq[it] = *reinterpret_cast<const float4*>(rb + off);
            if (accm != EV_ACC_STORE) {
              oq[it] = make_float4(0.f, 0.f, 0.f, 0.f);
              if (col_ok && row < len) oq[it] = *reinterpret_cast<const float4*>(ob + off);
            }
          }
          float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
          if (bias && col_ok) b4 = __ldg(reinterpret_cast<const float4*>(bias + c + cq * 4));
          if (!waited) {
            mbar_wait(acc_full(buf), (tile_cnt >> 1) & 1);
            tc_fence_after();
            waited = true;
          }
          float v[32];
          tmem_ld32(taddr + (uint32_t)c, cw, v);      // thread = row (TMEM lane), 32 consecutive columns
          // registers -> smem, 16-byte chunks XOR-swizzled by the row so both phases are conflict free
#pragma unroll
          for (int q = 0; q < 8; ++q)
            *reinterpret_cast<float4*>(stg + lane * 32 + ((q ^ (lane & 7)) << 2)) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
          __syncwarp();
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int r = it * 4 + rr;
            const int row = row_base + r;
            float4 o = *reinterpret_cast<const float4*>(stg + r * 32 + ((cq ^ (r & 7)) << 2));
            if (row < len) {
              o.x += b4.x; o.y += b4.y; o.z += b4.z; o.w += b4.w;
              if (oact != EV_ACT_NONE) {
                o.x = act_apply(o.x, oact, 0.f); o.y = act_apply(o.y, oact, 0.f);
                o.z = act_apply(o.z, oact, 0.f); o.w = act_apply(o.w, oact, 0.f);
              }
              o.x += rq[it].x; o.y += rq[it].y; o.z += rq[it].z; o.w += rq[it].w;
              if (accm != EV_ACC_STORE) {
                o.x += oq[it].x; o.y += oq[it].y; o.z += oq[it].z; o.w += oq[it].w;
                if (accm == EV_ACC_ADD_DIV) { o.x /= p.div; o.y /= p.div; o.z /= p.div; o.w /= p.div; }
              }
            } else {
              o = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            if (col_ok && row < p.L) *reinterpret_cast<float4*>(ob + (size_t)row * p.Cout + n0 + c + cq * 4) = o;
          }
          __syncwarp();
        }
      }
      if (!waited) {     // a warp without columns in this tile (N <= 32) still follows the accumulator phases
        mbar_wait(acc_full(buf), (tile_cnt >> 1) & 1);
        tc_fence_after();
      }
      // all TMEM reads of this buffer are complete (tcgen05.wait::ld inside tmem_ld32): hand it back
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty(buf));
      ++tile_cnt;
    }
  } else if (warp < MMA_WARP) {
    // ============================ A producers ===================================================
    if (PDLM) asm volatile("griddepcontrol.wait;" ::: "memory");
    const int pwarp = warp - NEPI / 32;
    const int wpg = NPWARPS / pl.ngroups;          // warps per group
    const int grp = pwarp / wpg;
    const int gt = (pwarp - grp * wpg) * 32 + lane;  // thread index inside the group
    const int GT = wpg * 32;
    const bool lrelu = (p.in_act == EV_ACT_LRELU);
    const float slope = p.in_slope;
    const int total = rows_a * KBG;   // (row, granule) pairs; granule fastest -> coalesced row segments
    int a_cnt = 0;
    for (int tile = blockIdx.x; tile < pl.total_tiles; tile += gridDim.x) {
      int b, t0, n0, nt, len;
      decode(tile, b, t0, n0, nt, len);
      if (t0 >= len) continue;
      const float* __restrict__ xb = p.x + (size_t)b * p.L * p.Cin;
      const int cb_lo = (z_cur * n_cb) / pl.ksplit, cb_hi = ((z_cur + 1) * n_cb) / pl.ksplit;
      for (int cb = cb_lo; cb < cb_hi; ++cb, ++a_cnt) {
        if (a_cnt % pl.ngroups != grp) continue;       // this stage belongs to another producer group
        const int s = a_cnt % pl.a_stages;
        const int c0 = cb * KB;
        const int ngran = min(KB, p.Cin - c0) / CPG;
        uint8_t* dst = a_tiles + s * pl.a_stage_bytes;
        constexpr int ALD = BF16 ? A_LD / 2 : A_LD;     // (row, granule) pairs per thread and batch
        constexpr int NW = BF16 ? 2 : 1;                // float4 loads per pair (8 : 4 channels)
        for (int base = 0; base < total; base += GT * ALD) {
          // a batch of global loads is issued before anything else (memory-level parallelism)
          float4 v[ALD * NW];
#pragma unroll
          for (int u = 0; u < ALD; ++u) {
            const int idx = base + u * GT + gt;
            const int r = idx >> GSH, g = idx & (KBG - 1);
            const int row = t0 - halo + r;
            const bool ok = idx < total && g < ngran && row >= 0 && row < len;
            const float* src = xb + (size_t)row * p.Cin + c0 + g * CPG;
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
                if (lrelu) {
                  t.x = t.x > 0.f ? t.x : t.x * slope;
                  t.y = t.y > 0.f ? t.y : t.y * slope;
                  t.z = t.z > 0.f ? t.z : t.z * slope;
                  t.w = t.w > 0.f ? t.w : t.w * slope;
                }
              }
              uint8_t* d = dst + ((size_t)g * pl.rows_pad + r) * 16;
              if (BF16) {
                const float4 t0 = v[u * NW], t1 = v[u * NW + NW - 1];
                uint4 q;
                q.x = pack_bf16(t0.x, t0.y); q.y = pack_bf16(t0.z, t0.w);
                q.z = pack_bf16(t1.x, t1.y); q.w = pack_bf16(t1.z, t1.w);
                *reinterpret_cast<uint4*>(d) = q;
                if (X3B) {      // lo plane: what the bf16 rounding dropped (8 more significant bits)
                  uint4 l;
                  l.x = pack_bf16(t0.x - __uint_as_float(q.x << 16), t0.y - __uint_as_float(q.x & 0xffff0000u));
                  l.y = pack_bf16(t0.z - __uint_as_float(q.y << 16), t0.w - __uint_as_float(q.y & 0xffff0000u));
                  l.z = pack_bf16(t1.x - __uint_as_float(q.z << 16), t1.y - __uint_as_float(q.z & 0xffff0000u));
                  l.w = pack_bf16(t1.z - __uint_as_float(q.w << 16), t1.w - __uint_as_float(q.w & 0xffff0000u));
                  *reinterpret_cast<uint4*>(d + pl.a_plane_bytes) = l;
                }
              } else {
                const float4 t = v[u * NW];
                // round-to-nearest tf32 (the MMA would otherwise truncate the low 13 mantissa bits)
                const float4 h = make_float4(to_tf32(t.x), to_tf32(t.y), to_tf32(t.z), to_tf32(t.w));
                *reinterpret_cast<float4*>(d) = h;
                if (SPLIT3) {
                  const float4 l = make_float4(to_tf32(t.x - h.x), to_tf32(t.y - h.y), to_tf32(t.z - h.z), to_tf32(t.w - h.w));
                  *reinterpret_cast<float4*>(d + pl.a_plane_bytes) = l;
                }
              }
            }
          }
        }
        fence_proxy_async();      // generic-proxy smem writes -> visible to the tensor core (async proxy)
        mbar_arrive(a_full(s));
      }
    }
  } else if (warp == MMA_WARP) {
    // ============================ MMA issuer =====================================================
    // All 32 lanes run the (warp-uniform) control flow and the barrier waits; one elected lane issues the tcgen05 instructions
    // (descriptors stay in uniform registers: no vote loop / R2UR per MMA, see tc_common.cuh: elect_one).
    {
      if (PDLM == 2) asm volatile("griddepcontrol.wait;" ::: "memory");      // p.lens (see the note after the set-up)
      const uint32_t a_lbo = (uint32_t)pl.rows_pad * 16u, b_lbo = (uint32_t)BN * 16u;
      const uint64_t a_desc0 = make_desc(0u, a_lbo, 128u), b_desc0 = make_desc(0u, b_lbo, 128u);
      const uint32_t a_tap = (uint32_t)p.dil * 16u, a_k8 = 2u * a_lbo, b_k8 = 2u * b_lbo;
      int a_cnt = 0, b_cnt = 0, tile_cnt = 0;
      for (int tile = blockIdx.x; tile < pl.total_tiles; tile += gridDim.x) {
        int b, t0, n0, nt, len;
        decode(tile, b, t0, n0, nt, len);
        if (t0 >= len) continue;
        // instruction descriptor (cute::UMMA::InstrDescriptor): D=F32 [4,6)=1, A=TF32 [7,10)=2, B=TF32 [10,13)=2,
        // A/B K-major (bits 15,16 = 0), N>>3 at [17,23), M>>4 at [24,29)
        // (bf16 mode: A = B = BF16, format code 1)
        const uint32_t fmt = BF16 ? 1u : 2u;
        const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(nt >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
        const int buf = tile_cnt & 1;
        mbar_wait(acc_empty(buf), ((tile_cnt >> 1) & 1) ^ 1);     // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t d_base = tmem_base + (uint32_t)(buf * MT * BN);
        const int cb_lo = (z_cur * n_cb) / pl.ksplit, cb_hi = ((z_cur + 1) * n_cb) / pl.ksplit;
        for (int cb = cb_lo; cb < cb_hi; ++cb, ++a_cnt) {
          const int sa = a_cnt % pl.a_stages;
          const int nk8 = min(KB, p.Cin - cb * KB) / (2 * CPG);   // MMA K steps: two 16-byte granules each
          mbar_wait(a_full(sa), (a_cnt / pl.a_stages) & 1);
          const uint64_t a_hi0 = desc_advance(a_desc0, smem_u32(a_tiles + sa * pl.a_stage_bytes));
          for (int j = 0; j < p.K; ++j, ++b_cnt) {
            const int sb = b_cnt % pl.b_stages;
            mbar_wait(b_full(sb), (b_cnt / pl.b_stages) & 1);
            tc_fence_after();
            const uint64_t b_hi0 = desc_advance(b_desc0, smem_u32(b_tiles + sb * pl.b_stage_bytes));
            const uint64_t a_j = desc_advance(a_hi0, (uint32_t)j * a_tap);
            if (elect_one()) {
              for (int k8 = 0; k8 < nk8; ++k8) {
                const uint64_t b_hi = desc_advance(b_hi0, (uint32_t)k8 * b_k8);
                const uint64_t a_k = desc_advance(a_j, (uint32_t)k8 * a_k8);
                const uint32_t first = ((cb - cb_lo) | j | k8) != 0 ? 1u : 0u;
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) {      // one weight tile feeds MT accumulators
                  const uint64_t a_hi = desc_advance(a_k, (uint32_t)(mt * BM) * 16u);
                  const uint32_t d = d_base + (uint32_t)(mt * BN);
                  if (SPLIT3) {
                    const uint64_t a_lo = desc_advance(a_hi, (uint32_t)pl.a_plane_bytes);
                    const uint64_t b_lo = desc_advance(b_hi, (uint32_t)pl.b_plane_bytes);
                    umma_tf32(d, a_lo, b_hi, idesc, first);     // small terms first
                    umma_tf32(d, a_hi, b_lo, idesc, 1u);
                    umma_tf32(d, a_hi, b_hi, idesc, 1u);
                  } else if (X3B) {
                    const uint64_t a_lo = desc_advance(a_hi, (uint32_t)pl.a_plane_bytes);
                    const uint64_t b_lo = desc_advance(b_hi, (uint32_t)pl.b_plane_bytes);
                    umma_bf16(d, a_lo, b_hi, idesc, first);
                    umma_bf16(d, a_hi, b_lo, idesc, 1u);
                    umma_bf16(d, a_hi, b_hi, idesc, 1u);
                  } else if (BF16) {
                    umma_bf16(d, a_hi, b_hi, idesc, first);
                  } else {
                    umma_tf32(d, a_hi, b_hi, idesc, first);
                  }
                }
              }
              umma_commit(b_empty(sb));     // weight stage free once these MMAs have read it
              if (j == p.K - 1) {
                umma_commit(a_empty(sa));   // activation stage free
                if (cb == cb_hi - 1) umma_commit(acc_full(buf));       // accumulators of this tile complete -> epilogue
              }
            }
            __syncwarp();
          }
        }
        ++tile_cnt;
      }
    }
  } else {
    // ============================ weight loader ==================================================
    if (lane == 0) {
      if (PDLM == 2) asm volatile("griddepcontrol.wait;" ::: "memory");      // p.lens (see the note after the set-up)
      // w_tc layout: [plane (hi, lo)][N tile of BNp = min(Cout,128)][tap][Cin/CPG granules][BNp][16 bytes]
      // (4 fp32 or 8 bf16 per granule; granule-major inside a tile)
      const int cin4 = p.Cin / CPG;
      const int bnp = p.Cout < 128 ? p.Cout : 128;
      const size_t plane = (size_t)p.K * cin4 * p.Cout * 4;         // 4-byte words per plane (fp32 or packed bf16 granules)
      const size_t tile_stride = (size_t)p.K * cin4 * bnp * 4;      // floats per packed N tile
      int b_cnt = 0;
      for (int tile = blockIdx.x; tile < pl.total_tiles; tile += gridDim.x) {
        int b, t0, n0, nt, len;
        decode(tile, b, t0, n0, nt, len);
        if (t0 >= len) continue;
        const float* wt = p.w + (size_t)(n0 / bnp) * tile_stride + (size_t)(n0 % bnp) * 4;
        const int cb_lo = (z_cur * n_cb) / pl.ksplit, cb_hi = ((z_cur + 1) * n_cb) / pl.ksplit;
        for (int cb = cb_lo; cb < cb_hi; ++cb) {
          const int ngran = min(KB, p.Cin - cb * KB) / CPG;
          for (int j = 0; j < p.K; ++j, ++b_cnt) {
            const int sb = b_cnt % pl.b_stages;
            mbar_wait(b_empty(sb), ((b_cnt / pl.b_stages) & 1) ^ 1);
            mbar_expect_tx(b_full(sb), (uint32_t)(PLANES * ngran * nt * 16));
            const uint32_t dst = smem_u32(b_tiles + sb * pl.b_stage_bytes);
            const float* src = wt + ((size_t)j * cin4 + (size_t)cb * KBG) * bnp * 4;
            if (nt == bnp) {
              // the CTA's N tile is a whole packed tile: the stage's granules are adjacent in memory -> ONE bulk copy per plane
              bulk_g2s(dst, src, (uint32_t)(ngran * nt * 16), b_full(sb));
              if (PLANES == 2) bulk_g2s(dst + (uint32_t)pl.b_plane_bytes, src + plane, (uint32_t)(ngran * nt * 16), b_full(sb));
            } else {
              for (int g = 0; g < ngran; ++g) {
                bulk_g2s(dst + (uint32_t)(g * BN * 16), src + (size_t)g * bnp * 4, (uint32_t)(nt * 16), b_full(sb));
                if (PLANES == 2) bulk_g2s(dst + (uint32_t)(pl.b_plane_bytes + g * BN * 16), src + plane + (size_t)g * bnp * 4, (uint32_t)(nt * 16), b_full(sb));
              }
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

// Second half of a K-split convolution: out = epi( sum_z partial[z] ) with the slices added in the
// fixed order z = 0..S-1 (deterministic, batch invariant), then bias / activation / residual /
// accumulate exactly like the fused epilogue.  One float4 per thread.
template <bool PDL>
__global__ void __launch_bounds__(256) splitk_reduce_kernel(ConvParams p, int S) {
  pdl_entry<PDL>();
  const size_t per = (size_t)p.B * p.L * p.Cout;
  const size_t i4 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i4 * 4 >= per) return;
  const size_t e = i4 * 4;
  const int col = (int)(e % p.Cout);
  const size_t rowg = e / p.Cout;
  const int b = (int)(rowg / p.L), row = (int)(rowg % p.L);
  const int len = p.lens ? min(p.L, p.lens[b] * p.lens_mul) : p.L;
  splitk_reduce_store(p, S, per, e, b, row, col, len);
}

}  // namespace tc

static int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return (e && *e) ? atoi(e) : dflt;
}

template <int MODE, int MT, int KBG, int PDLM>
static int launch_tc_pdl(const ConvParams& p, const tc::Plan& pl, cudaStream_t st) {
  static std::atomic<uint64_t> attr_devs{0};   // per instantiation; function attributes are per device
  if (first_use_on_device(attr_devs))
    cudaFuncSetAttribute(tc::conv1d_tc_kernel<MODE, MT, KBG, PDLM>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  const int grid = pl.total_tiles < sm_count() ? pl.total_tiles : sm_count();
  if (PDLM) {
    const cudaError_t e = launch_with_pdl(tc::conv1d_tc_kernel<MODE, MT, KBG, PDLM>, dim3((unsigned)grid), dim3(tc::NTHREADS),
                                          (size_t)pl.smem_total, st, p, pl);
    if (e != cudaSuccess) { set_error("conv1d_tc_kernel (PDL launch): %s", cudaGetErrorString(e)); return EV_ECUDA; }
    count_launch();
    return EV_OK;
  }
  tc::conv1d_tc_kernel<MODE, MT, KBG, PDLM><<<grid, tc::NTHREADS, pl.smem_total, st>>>(p, pl);
  EV_CUDA_LAUNCH_CHECK("conv1d_tc_kernel");
  return EV_OK;
}

// pdl: 0 for the default path, else pdl_mode()
template <int MODE, int MT, int KBG>
static int launch_tc_variant(const ConvParams& p, const tc::Plan& pl, cudaStream_t st, int pdl) {
  if (pdl >= 2) return launch_tc_pdl<MODE, MT, KBG, 2>(p, pl, st);
  if (pdl == 1) return launch_tc_pdl<MODE, MT, KBG, 1>(p, pl, st);
  return launch_tc_pdl<MODE, MT, KBG, 0>(p, pl, st);
}

template <int MODE, int KBG>
static int launch_tc_mt(const ConvParams& p, const tc::Plan& pl, cudaStream_t st, int pdl) {
  if (pl.mt == 4) return launch_tc_variant<MODE, 4, KBG>(p, pl, st, pdl);
  if (pl.mt == 2) return launch_tc_variant<MODE, 2, KBG>(p, pl, st, pdl);
  return launch_tc_variant<MODE, 1, KBG>(p, pl, st, pdl);
}

template <int MODE, int KBG>
static void preload_tc_mode() {
  cudaFuncSetAttribute(tc::conv1d_tc_kernel<MODE, 1, KBG, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  cudaFuncSetAttribute(tc::conv1d_tc_kernel<MODE, 2, KBG, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  cudaFuncSetAttribute(tc::conv1d_tc_kernel<MODE, 4, KBG, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
}
void preload_conv1d_tc() {      // see conv1d_gp.cu: preload_conv1d_gp (the default, non-PDL instantiations)
  preload_tc_mode<0, 4>(); preload_tc_mode<0, 8>(); preload_tc_mode<1, 4>(); preload_tc_mode<2, 4>(); preload_tc_mode<2, 8>();
  preload_tc_mode<3, 4>(); preload_tc_mode<3, 8>();
  cudaFuncAttributes fa;
  cudaFuncGetAttributes(&fa, tc::splitk_reduce_kernel<false>);
  cudaGetLastError();
}

static int validate_conv1d_tc(const ConvParams& p, int mode) {
  EV_CHECK_ARG(p.B > 0 && p.L > 0, "conv1d_tc: bad problem B=%d L=%d", p.B, p.L);
  EV_CHECK_ARG(p.Cin % (mode >= 2 ? 16 : 8) == 0, "conv1d_tc: Cin=%d must be a multiple of %d", p.Cin, mode >= 2 ? 16 : 8);
  EV_CHECK_ARG(p.Cout % 16 == 0 && (p.Cout <= 128 || p.Cout % 128 == 0), "conv1d_tc: Cout=%d must be a multiple of 16, and of 128 above 128", p.Cout);
  EV_CHECK_ARG(p.K >= 1 && (p.K & 1) && p.dil >= 1, "conv1d_tc: K=%d must be odd, dil=%d >= 1", p.K, p.dil);
  EV_CHECK_ARG(p.in_act == EV_ACT_NONE || p.in_act == EV_ACT_LRELU, "conv1d_tc: unsupported input activation");
  return EV_OK;
}

// K granules per stage decide how the (channel block, tap) reduction is ordered, so they must be a function
// of the layer shape alone: 8 if a (widest-N, one-accumulator) tile fits with them in 1x mode, else 4.
int tc_shape_kbg(const ConvParams& p, int mode) {
  tc::Plan pl;
  const int bn_max = p.Cout <= 128 ? p.Cout : 128;
  return (mode != 1 && tc::make_plan(p, mode, bn_max, 1, 8, 4, &pl)) ? 8 : 4;
}

// K-split: a long reduction at few output tiles (the acoustic model's GEMMs at batch 1; the conv-FFN's second
// conv has K = 3*1536) is one long serial chain per tile; share it among S CTAs with private partial buffers
// + a fixed-order reduce kernel.  S is requested per LAYER by the engine (never derived from batch or
// length), so the summation order -- and therefore every output bit -- is the same for a B=1 call and for
// the same utterance inside any batch.
static int apply_ksplit(const ConvParams& p, int mode, tc::Plan* pl) {
  const size_t per = (size_t)p.B * p.L * p.Cout;
  if (p.ksplit > 1 && (p.splitk_ws || p.splitk_cap == (size_t)-1)) {
    int S = p.ksplit;
    const int cpg = mode >= 2 ? 8 : 4;
    const int n_cb = (p.Cin + cpg * pl->kbg - 1) / (cpg * pl->kbg);
    if (S > n_cb) S = n_cb;
    if ((size_t)S * per > p.splitk_cap) { set_error("conv1d_tc: split-K scratch too small (%zu < %zu floats)", p.splitk_cap, (size_t)S * per); return EV_EWORKSPACE; }
    if (S > 1) {
      pl->ksplit = S;
      pl->total_tiles *= S;
    }
  }
  return EV_OK;
}

// Tile / pipeline plan of one launch (pure host arithmetic; also exported as ev_debug_tc_plan so the CPU tests
// can check the invariants the kernel's barrier protocol and the batch-invariance contract rely on).
static int plan_conv1d_tc(const ConvParams& p, int mode, tc::Plan* out) {
  EV_TRY(validate_conv1d_tc(p, mode));
  // Tile shape.  None of these choices changes the order in which any output element's K reduction is
  // summed, so results are bitwise independent of batch size / sequence length (batch-invariant contract).
  //  * N tile: min(C_out, 128) (the weight packing tile: one bulk copy per stage); halved (down to 32) only for
  //    launches with a handful of tiles (measured: finer N splitting costs more in per-granule copies and
  //    replicated A staging than it gains in parallelism).
  //  * rows per tile: as many 128-row accumulators as still leave about one tile per SM (every weight tile
  //    fetched from L2 then feeds MT MMAs), limited by TMEM (2 x MT x BN <= 512 columns) and smem.
  static const int bn_thresh = env_int("EV_TC_BN_TILES", 24);     // tuning knobs (tile shape only: results are unaffected)
  static const int mt_thresh = env_int("EV_TC_MT_TILES", 120);
  const long long tiles128 = (long long)((p.L + tc::BM - 1) / tc::BM) * p.B;
  int BN = p.Cout <= 128 ? p.Cout : 128;          // == the packing tile of packing.to_tc_layout
  while (BN >= 64 && (BN / 2) % 16 == 0 && tiles128 * ((p.Cout + BN - 1) / BN) < bn_thresh) BN /= 2;
  int mt = tiles128 >= 4 * mt_thresh ? 4 : (tiles128 >= 2 * mt_thresh ? 2 : 1);
  tc::Plan pl;
  const int kbg = tc_shape_kbg(p, mode);
  for (;; mt >>= 1) {
    if (tc::make_plan(p, mode, BN, mt, kbg, 4, &pl)) break;
    if (tc::make_plan(p, mode, BN, mt, kbg, 2, &pl)) break;
    if (mt == 1) { set_error("conv1d_tc: tile does not fit in shared memory / TMEM (K=%d dil=%d Cout=%d)", p.K, p.dil, p.Cout); return EV_EINVAL; }
  }
  EV_TRY(apply_ksplit(p, mode, &pl));
  *out = pl;
  return EV_OK;
}

int debug_tc_plan(const ConvParams& p, int mode, int* v) {
  tc::Plan pl;
  const int rc = plan_conv1d_tc(p, mode, &pl);
  if (rc != EV_OK) return rc;
  v[0] = pl.BN; v[1] = pl.mt; v[2] = pl.kbg; v[3] = pl.a_stages; v[4] = pl.b_stages; v[5] = pl.ngroups;
  v[6] = pl.ksplit; v[7] = pl.tmem_cols; v[8] = pl.smem_total; v[9] = pl.total_tiles; v[10] = pl.rows_pad;
  return EV_OK;
}

static int dispatch_tc(const ConvParams& p, int mode, const tc::Plan& pl, cudaStream_t st, int pdl = 0) {
  if (mode == 1) return launch_tc_mt<1, 4>(p, pl, st, pdl);
  if (mode == 3) return pl.kbg == 8 ? launch_tc_mt<3, 8>(p, pl, st, pdl) : launch_tc_mt<3, 4>(p, pl, st, pdl);
  if (mode == 2) return pl.kbg == 8 ? launch_tc_mt<2, 8>(p, pl, st, pdl) : launch_tc_mt<2, 4>(p, pl, st, pdl);
  return pl.kbg == 8 ? launch_tc_mt<0, 8>(p, pl, st, pdl) : launch_tc_mt<0, 4>(p, pl, st, pdl);
}

// p.w must be in the tensor-core layout [plane][Cout/BNp][K][Cin/4][BNp][4] (packing.py: to_tc_layout);
// mode 0: 1xTF32, 1: 3xTF32 fp32 emulation (reads both planes), 2: bf16 operands (p.w in the bf16 tc layout).
int launch_conv1d_tc(const ConvParams& p, int mode, cudaStream_t st) {
  tc::Plan pl;
  EV_TRY(plan_conv1d_tc(p, mode, &pl));
  const size_t per = (size_t)p.B * p.L * p.Cout;
  const int rc = dispatch_tc(p, mode, pl, st, pdl_mode());      // EV_PDL (default 2)
  if (rc != EV_OK || pl.ksplit == 1) return rc;
  const size_t n4 = per / 4;
  launch_k(tc::splitk_reduce_kernel<true>, tc::splitk_reduce_kernel<false>, (unsigned)((n4 + 255) / 256), 256, 0, st, p, pl.ksplit);
  EV_CUDA_LAUNCH_CHECK("splitk_reduce_kernel");
  return EV_OK;
}

}  // namespace ev
