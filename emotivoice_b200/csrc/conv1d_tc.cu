// Time-major 1-D convolution as an implicit GEMM on the 5th-gen tensor cores (sm_100a):
// tcgen05.mma (kind::tf32, fp32 accumulate in TMEM), operands staged in shared memory, weights
// streamed by bulk async copies (cp.async.bulk -> mbarrier complete_tx), accumulator read back
// with tcgen05.ld for the fused epilogue.  Same contract as conv1d_tm.cu (see ev_common.cuh):
//
//   out[b,t,co] = epi( bias[co] + sum_j sum_ci w[j][ci][co] * act_in( x[b, t + (j-(K-1)/2)*dil, ci] ) )
//
// GEMM view: M = 128 time steps per CTA, N = C_out tile (<= 256), K = taps x C_in.
//
// The trick that makes a dilated k-tap convolution cost ONE activation fetch per tile: the A
// operand lives in shared memory in the *no-swizzle K-major* UMMA layout with the 8-row-group
// stride (SBO) set to 128 B, i.e. element (row r, 16-byte K-granule g) sits at
//        A + (g * rows_pad + r) * 16 bytes
// so consecutive rows are exactly 16 B apart for the whole tile, and tap j of the convolution is
// the same tile with the descriptor start address advanced by j*dil rows (16 B granularity).
// One staged tile of BM + (K-1)*dil rows feeds all K taps.  The producer warps apply the
// LeakyReLU prologue, the zero padding at the sequence ends and the layout change while staging,
// so activations stay plain fp32 time-major in HBM and no tensor map is needed.
//
// Roles (320 threads): warps 0-7 stage A (then run the epilogue, one TMEM lane = one output row
// per thread, two warps per lane quadrant splitting the columns); warp 8 allocates TMEM and its
// elected lane issues every tcgen05.mma; warp 9's elected lane streams the weight tiles.  Pipelines: A 2 stages (a_full/a_empty), B 3 stages
// (b_full/b_empty, released by tcgen05.commit), accumulator (acc_full).
#include "ev_common.cuh"

namespace ev {

namespace tc {

constexpr int BM = 128;         // rows (time steps) per accumulator == TMEM lanes; a CTA owns MT of them
constexpr int A_STAGES = 2;
constexpr int B_STAGES = 3;
constexpr int NPRODUCER = 256;                 // warps 0-7 stage A, then run the epilogue
constexpr int MMA_WARP = NPRODUCER / 32;       // warp 8: TMEM alloc + MMA issue
constexpr int NTHREADS = NPRODUCER + 64;       // + warp 9: weight loader
constexpr int A_LD = 6;                        // float4 loads in flight per producer thread (rows_a <= 192)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "LAB_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE;\n\t"
      "bra LAB_WAIT;\n\t"
      "DONE:\n\t"
      "}" ::"r"(bar), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], tf32 inputs, fp32 accumulate, M=128, N from idesc, K=8
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
// no-swizzle K-major shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout):
// [0,14) start>>4 | [16,30) LBO>>4 (stride between the two 16-B K granules of one MMA)
// | [32,46) SBO>>4 (stride between 8-row groups) | [46,48) version = 1 | [61,64) layout = 0
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((addr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
__device__ __forceinline__ float to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

struct SmemLayout {
  int rows_pad;        // staged rows per A granule; == 8/KBG (mod 8) -> conflict-free 16 B stores
  int a_plane_bytes;   // KBG * rows_pad * 16   (one tf32 plane: hi, or lo in 3xTF32 mode)
  int b_plane_bytes;   // KBG * BN * 16
  int a_stage_bytes, b_stage_bytes;
  int total;
};
__host__ __device__ inline SmemLayout smem_layout(int K, int dil, int BN, int planes, int mt, int kbg) {
  SmemLayout s;
  const int rows = BM * mt + (K - 1) * dil;
  s.rows_pad = ((rows + 7) / 8) * 8 + 8 / kbg;
  s.a_plane_bytes = kbg * s.rows_pad * 16;
  s.b_plane_bytes = kbg * BN * 16;
  s.a_stage_bytes = planes * s.a_plane_bytes;
  s.b_stage_bytes = planes * s.b_plane_bytes;
  s.total = 1024 /*barriers + tmem ptr + alignment slack*/ + A_STAGES * s.a_stage_bytes + B_STAGES * s.b_stage_bytes;
  return s;
}

// SPLIT3 = false: one tf32 MMA per K step (operands rounded to nearest tf32).
// SPLIT3 = true : "3xTF32" fp32 emulation: x = hi + lo with hi = tf32(x), lo = tf32(x - hi);
//                 a*b ~= a_lo*b_hi + a_hi*b_lo + a_hi*b_hi (the dropped lo*lo term is 2^-22 relative),
//                 three MMAs per K step into the same fp32 TMEM accumulator.  Weights arrive pre-split
//                 (two planes, packing.to_tc_layout); activations are split by the producer warps.
// MT: 128-row accumulators per CTA (TMEM columns MT*BN).  Every weight tile fetched from L2 is used by
//     MT MMAs, so the weight stream -- the dominant L2->SM traffic of a k-tap conv at 128 rows per
//     CTA -- shrinks MT-fold; consecutive accumulators share the halo rows of one staged A tile.
// KBG: 16-byte K granules (4 fp32 channels) per pipeline stage (8 -> 32 channels, 4 -> 16 channels).
template <bool SPLIT3, int MT, int KBG>
__global__ void __launch_bounds__(NTHREADS, (MT == 1 ? 2 : 1)) conv1d_tc_kernel(ConvParams p, int BN, int tmem_cols) {
  constexpr int PLANES = SPLIT3 ? 2 : 1;
  constexpr int KB = 4 * KBG;
  constexpr int GSH = (KBG == 8 ? 3 : 2);
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const int b = blockIdx.z;
  const int t0 = blockIdx.x * (BM * MT);
  const int n0 = blockIdx.y * BN;
  const int nt = min(BN, p.Cout - n0);          // this tile's N (multiple of 16)
  const int len = p.lens ? min(p.L, p.lens[b] * p.lens_mul) : p.L;
  float* ob = p.out + (size_t)b * p.L * p.Cout;   // may alias p.res (in-place residual)

  if (t0 >= len) {   // whole tile is padding (uniform per CTA): the batch-invariant contract stores zeros
    for (int i = tid; i < BM * MT * (nt / 4); i += NTHREADS) {
      const int r = i / (nt / 4), c4 = i % (nt / 4);
      if (t0 + r < p.L) *reinterpret_cast<float4*>(ob + (size_t)(t0 + r) * p.Cout + n0 + c4 * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    return;
  }

  const SmemLayout sl = smem_layout(p.K, p.dil, BN, PLANES, MT, KBG);
  // carve: [0,128) barriers, [128,132) tmem base; tiles from 1024
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem_raw + 128);
  uint8_t* a_tiles = smem_raw + 1024;
  uint8_t* b_tiles = a_tiles + A_STAGES * sl.a_stage_bytes;
  const uint32_t bar_base = smem_u32(bars);
  auto a_full = [&](int s) { return bar_base + 8u * s; };
  auto a_empty = [&](int s) { return bar_base + 8u * (A_STAGES + s); };
  auto b_full = [&](int s) { return bar_base + 8u * (2 * A_STAGES + s); };
  auto b_empty = [&](int s) { return bar_base + 8u * (2 * A_STAGES + B_STAGES + s); };
  const uint32_t acc_full = bar_base + 8u * (2 * A_STAGES + 2 * B_STAGES);

  if (tid == 0) {
    for (int s = 0; s < A_STAGES; ++s) { mbar_init(a_full(s), NPRODUCER); mbar_init(a_empty(s), 1); }
    for (int s = 0; s < B_STAGES; ++s) { mbar_init(b_full(s), 1); mbar_init(b_empty(s), 1); }
    mbar_init(acc_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == MMA_WARP) {   // TMEM allocation by one full warp; the same warp frees it
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int n_cb = (p.Cin + KB - 1) / KB;
  const int halo = ((p.K - 1) / 2) * p.dil;
  const int rows_a = BM * MT + (p.K - 1) * p.dil;

  if (warp < NPRODUCER / 32) {
    // ------------------------------ A producers --------------------------------------------
    const float* __restrict__ xb = p.x + (size_t)b * p.L * p.Cin;
    const bool lrelu = (p.in_act == EV_ACT_LRELU);
    const float slope = p.in_slope;
    const int total = rows_a * KBG;   // (row, granule) pairs; granule fastest -> coalesced row segments
    for (int cb = 0; cb < n_cb; ++cb) {
      const int s = cb % A_STAGES;
      const int c0 = cb * KB;
      const int ngran = min(KB, p.Cin - c0) / 4;
      uint8_t* dst = a_tiles + s * sl.a_stage_bytes;
      for (int base = 0; base < total; base += NPRODUCER * A_LD) {
        // a batch of global loads is issued before anything else (memory-level parallelism: at
        // batch 1 the working set is L2 resident and the kernel is latency bound)
        float4 v[A_LD];
#pragma unroll
        for (int u = 0; u < A_LD; ++u) {
          const int idx = base + u * NPRODUCER + tid;
          const int r = idx >> GSH, g = idx & (KBG - 1);
          const int row = t0 - halo + r;
          v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (idx < total && g < ngran && row >= 0 && row < len)
            v[u] = __ldg(reinterpret_cast<const float4*>(xb + (size_t)row * p.Cin + c0 + g * 4));
        }
        if (base == 0) mbar_wait(a_empty(s), ((cb / A_STAGES) & 1) ^ 1);
#pragma unroll
        for (int u = 0; u < A_LD; ++u) {
          const int idx = base + u * NPRODUCER + tid;
          const int r = idx >> GSH, g = idx & (KBG - 1);
          if (idx < total && g < ngran) {
            float4 t = v[u];
            if (lrelu) {
              t.x = t.x > 0.f ? t.x : t.x * slope;
              t.y = t.y > 0.f ? t.y : t.y * slope;
              t.z = t.z > 0.f ? t.z : t.z * slope;
              t.w = t.w > 0.f ? t.w : t.w * slope;
            }
            // round-to-nearest tf32 (the MMA would otherwise truncate the low 13 mantissa bits)
            float4 h = make_float4(to_tf32(t.x), to_tf32(t.y), to_tf32(t.z), to_tf32(t.w));
            uint8_t* d = dst + ((size_t)g * sl.rows_pad + r) * 16;
            *reinterpret_cast<float4*>(d) = h;
            if (SPLIT3) {
              const float4 l = make_float4(to_tf32(t.x - h.x), to_tf32(t.y - h.y), to_tf32(t.z - h.z), to_tf32(t.w - h.w));
              *reinterpret_cast<float4*>(d + sl.a_plane_bytes) = l;
            }
          }
        }
      }
      fence_proxy_async();      // generic-proxy smem writes -> visible to the tensor core (async proxy)
      mbar_arrive(a_full(s));
    }
    // ------------------------------ epilogue ----------------------------------------------
    // warp w owns TMEM lanes 32*(w%4).. (hardware restriction) and every second 16-column chunk.
    const int quad = warp & 3, half = warp >> 2;
    const float* __restrict__ bias = p.bias ? p.bias + (size_t)b * p.bias_bs : nullptr;
    bool waited = false;
#pragma unroll 1
    for (int mt = 0; mt < MT; ++mt) {
      const int row = t0 + mt * BM + quad * 32 + lane;
      const bool row_ok = row < p.L, row_live = row < len;
      const float* rrow = (p.res && row_live) ? p.res + ((size_t)b * p.L + row) * p.Cout + n0 : nullptr;
      float* orow = ob + (size_t)row * p.Cout + n0;
      const bool acc_rd = (p.acc != EV_ACC_STORE) && row_live;
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(mt * BN);
      float4 rq[4], oq[4];
      auto prefetch = [&](int c) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (rrow) rq[q] = *reinterpret_cast<const float4*>(rrow + c + q * 4);
          if (acc_rd) oq[q] = *reinterpret_cast<const float4*>(orow + c + q * 4);
        }
      };
      int c = half * 16;
      if (c < nt) prefetch(c);               // residual / accumulate operands in flight while the MMAs finish
      if (!waited) {
        mbar_wait(acc_full, 0);
        tc_fence_after();
        waited = true;
      }
      for (; c < nt; c += 32) {
        float v[16];
        tmem_ld16(taddr + (uint32_t)c, v);     // warp-collective: every lane participates
        if (row_ok) {
          if (row_live) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              float t = v[i];
              if (bias) t += __ldg(bias + n0 + c + i);
              v[i] = act_apply(t, p.out_act, 0.f);
            }
            if (rrow) {
#pragma unroll
              for (int q = 0; q < 4; ++q) { v[q * 4 + 0] += rq[q].x; v[q * 4 + 1] += rq[q].y; v[q * 4 + 2] += rq[q].z; v[q * 4 + 3] += rq[q].w; }
            }
            if (acc_rd) {
#pragma unroll
              for (int q = 0; q < 4; ++q) { v[q * 4 + 0] += oq[q].x; v[q * 4 + 1] += oq[q].y; v[q * 4 + 2] += oq[q].z; v[q * 4 + 3] += oq[q].w; }
              if (p.acc == EV_ACC_ADD_DIV) {
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] /= p.div;
              }
            }
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = 0.f;
          }
          if (c + 32 < nt) prefetch(c + 32);   // next chunk's operands fly during these stores and the next TMEM load
#pragma unroll
          for (int q = 0; q < 4; ++q)
            *reinterpret_cast<float4*>(orow + c + q * 4) = make_float4(v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
        }
      }
    }
  } else if (warp == MMA_WARP) {
    // ------------------------------ MMA issuer ---------------------------------------------
    if (lane == 0) {
      // instruction descriptor (cute::UMMA::InstrDescriptor): D=F32 [4,6)=1, A=TF32 [7,10)=2, B=TF32 [10,13)=2,
      // A/B K-major (bits 15,16 = 0), N>>3 at [17,23), M>>4 at [24,29)
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(nt >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      const uint32_t a_lbo = (uint32_t)sl.rows_pad * 16u, b_lbo = (uint32_t)BN * 16u;
      int it = 0;
      for (int cb = 0; cb < n_cb; ++cb) {
        const int sa = cb % A_STAGES;
        const int nk8 = min(KB, p.Cin - cb * KB) / 8;
        mbar_wait(a_full(sa), (cb / A_STAGES) & 1);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(a_tiles + sa * sl.a_stage_bytes);
        for (int j = 0; j < p.K; ++j, ++it) {
          const int sb = it % B_STAGES;
          mbar_wait(b_full(sb), (it / B_STAGES) & 1);
          tc_fence_after();
          const uint32_t b_addr = smem_u32(b_tiles + sb * sl.b_stage_bytes);
          for (int k8 = 0; k8 < nk8; ++k8) {
            const uint32_t b_off = (uint32_t)(2 * k8) * b_lbo;
            const uint64_t b_hi = make_desc(b_addr + b_off, b_lbo, 128u);
            const uint64_t b_lo = make_desc(b_addr + sl.b_plane_bytes + b_off, b_lbo, 128u);
            const uint32_t first = (cb | j | k8) != 0 ? 1u : 0u;
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {      // one weight tile feeds MT accumulators
              const uint32_t a_off = (uint32_t)((2 * k8) * sl.rows_pad + mt * BM + j * p.dil) * 16u;
              const uint64_t a_hi = make_desc(a_addr + a_off, a_lbo, 128u);
              const uint32_t d = tmem_base + (uint32_t)(mt * BN);
              if (SPLIT3) {
                const uint64_t a_lo = make_desc(a_addr + sl.a_plane_bytes + a_off, a_lbo, 128u);
                umma_tf32(d, a_lo, b_hi, idesc, first);     // small terms first
                umma_tf32(d, a_hi, b_lo, idesc, 1u);
                umma_tf32(d, a_hi, b_hi, idesc, 1u);
              } else {
                umma_tf32(d, a_hi, b_hi, idesc, first);
              }
            }
          }
          umma_commit(b_empty(sb));     // weight stage free once these MMAs have read it
        }
        umma_commit(a_empty(sa));       // activation stage free
      }
      umma_commit(acc_full);            // accumulator complete -> epilogue
    }
    __syncwarp();
  } else {
    // ------------------------------ weight loader ------------------------------------------
    if (lane == 0) {
      // w_tc layout: [plane (hi, lo)][tap][Cin/4][Cout][4] fp32  (granule-major; one granule row = 16 B)
      const int cin4 = p.Cin / 4;
      const size_t plane = (size_t)p.K * p.Cin * p.Cout;
      int it = 0;
      for (int cb = 0; cb < n_cb; ++cb) {
        const int ngran = min(KB, p.Cin - cb * KB) / 4;
        for (int j = 0; j < p.K; ++j, ++it) {
          const int sb = it % B_STAGES;
          mbar_wait(b_empty(sb), ((it / B_STAGES) & 1) ^ 1);
          mbar_expect_tx(b_full(sb), (uint32_t)(PLANES * ngran * nt * 16));
          const uint32_t dst = smem_u32(b_tiles + sb * sl.b_stage_bytes);
          for (int g = 0; g < ngran; ++g) {
            const float* src = p.w + (((size_t)j * cin4 + (size_t)cb * KBG + g) * p.Cout + n0) * 4;
            bulk_g2s(dst + (uint32_t)(g * BN * 16), src, (uint32_t)(nt * 16), b_full(sb));
            if (SPLIT3) bulk_g2s(dst + (uint32_t)(sl.b_plane_bytes + g * BN * 16), src + plane, (uint32_t)(nt * 16), b_full(sb));
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
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols));
  }
}

}  // namespace tc

template <bool SPLIT3, int MT, int KBG>
static int launch_tc_variant(const ConvParams& p, int BN, cudaStream_t st) {
  int tmem_cols = 32;
  while (tmem_cols < MT * BN) tmem_cols <<= 1;
  const tc::SmemLayout sl = tc::smem_layout(p.K, p.dil, BN, SPLIT3 ? 2 : 1, MT, KBG);
  EV_CHECK_ARG(sl.total <= 227 * 1024 && tmem_cols <= 512, "conv1d_tc: tile does not fit (smem %d, tmem %d)", sl.total, tmem_cols);
  static bool attr_set = false;   // per instantiation
  if (!attr_set) {
    cudaFuncSetAttribute(tc::conv1d_tc_kernel<SPLIT3, MT, KBG>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    attr_set = true;
  }
  dim3 grid((p.L + tc::BM * MT - 1) / (tc::BM * MT), (p.Cout + BN - 1) / BN, p.B);
  tc::conv1d_tc_kernel<SPLIT3, MT, KBG><<<grid, tc::NTHREADS, sl.total, st>>>(p, BN, tmem_cols);
  EV_CUDA_LAUNCH_CHECK("conv1d_tc_kernel");
  return EV_OK;
}

// p.w must be in the tensor-core layout [plane][K][Cin/4][Cout][4] (packing.py: to_tc_layout);
// split3 selects the 3xTF32 fp32-emulation variant (reads both planes).
int launch_conv1d_tc(const ConvParams& p, bool split3, cudaStream_t st) {
  EV_CHECK_ARG(p.B > 0 && p.L > 0 && p.B <= 65535, "conv1d_tc: bad problem B=%d L=%d", p.B, p.L);
  EV_CHECK_ARG(p.Cin % 8 == 0, "conv1d_tc: Cin=%d must be a multiple of 8", p.Cin);
  EV_CHECK_ARG(p.Cout % 16 == 0, "conv1d_tc: Cout=%d must be a multiple of 16", p.Cout);
  EV_CHECK_ARG(p.K >= 1 && (p.K & 1) && p.dil >= 1, "conv1d_tc: K=%d must be odd, dil=%d >= 1", p.K, p.dil);
  EV_CHECK_ARG(p.in_act == EV_ACT_NONE || p.in_act == EV_ACT_LRELU, "conv1d_tc: unsupported input activation");
  // N tile: a single 256-wide tile when C_out == 256 in 1x mode (A staged once); otherwise <= 128
  const int BN = p.Cout <= 128 ? p.Cout : ((p.Cout == 256 && !split3) ? 256 : 128);
  // rows per CTA: as many 128-row accumulators as keep ~one CTA per SM busy (weight traffic / MT)
  const long long tiles = (long long)((p.L + tc::BM - 1) / tc::BM) * p.B;
  int mt = tiles >= 4 * 120 ? 4 : (tiles >= 2 * 120 ? 2 : 1);
  const int planes = split3 ? 2 : 1, kbg = split3 ? 4 : 8;
  while (mt > 1 && (mt * BN > 512 || tc::smem_layout(p.K, p.dil, BN, planes, mt, kbg).total > 227 * 1024)) mt >>= 1;
  if (split3) {
    if (mt == 4) return launch_tc_variant<true, 4, 4>(p, BN, st);
    if (mt == 2) return launch_tc_variant<true, 2, 4>(p, BN, st);
    return launch_tc_variant<true, 1, 4>(p, BN, st);
  }
  if (mt == 4) return launch_tc_variant<false, 4, 8>(p, BN, st);
  if (mt == 2) return launch_tc_variant<false, 2, 8>(p, BN, st);
  return launch_tc_variant<false, 1, 8>(p, BN, st);
}

}  // namespace ev
