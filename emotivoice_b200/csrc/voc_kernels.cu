// HiFi-GAN generator kernels that are not implicit GEMMs (sm_100a, fp32): the layout change at
// the Generator.forward boundary, the final conv_post + tanh, and the callers' PCM16 conversion.
#include "ev_common.cuh"

namespace ev {

// (B, C, L) channels-first (the reference's Generator.forward input, hifigan/models.py:115)
// -> (B, L, C) time-major (the engine's internal layout).  32x32 smem tile transpose.
template <bool PDL>
__global__ void transpose_cf_to_tm_kernel(const float* __restrict__ in, float* __restrict__ out, int C, int L) {
  pdl_entry<PDL>();
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int l0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const float* ib = in + (size_t)b * C * L;
  float* ob = out + (size_t)b * C * L;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, l = l0 + threadIdx.x;
    tile[i][threadIdx.x] = (c < C && l < L) ? ib[(size_t)c * L + l] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int l = l0 + i, c = c0 + threadIdx.x;
    if (l < L && c < C) ob[(size_t)l * C + c] = tile[threadIdx.x][i];
  }
}
int launch_transpose_cf_to_tm(const float* in, float* out, int B, int C, int L, cudaStream_t st) {
  EV_CHECK_ARG(B > 0 && C > 0 && L > 0 && B <= 65535, "transpose: bad shape");
  dim3 grid((L + 31) / 32, (C + 31) / 32, B), block(32, 8);
  launch_k(transpose_cf_to_tm_kernel<true>, transpose_cf_to_tm_kernel<false>, grid, block, 0, st, in, out, C, L);
  EV_CUDA_LAUNCH_CHECK("transpose_cf_to_tm_kernel");
  return EV_OK;
}

// wav[b,t] = tanh( bias + sum_j sum_c w[j][c] * lrelu(x[b, t+j-(K-1)/2, c]) )
// (hifigan/models.py:127-129: F.leaky_relu default slope 0.01, Conv1d(C,1,7,pad 3), tanh).
// HBM-bound: each CTA stages (256 + K - 1) rows once; rows >= len read as zero padding.
constexpr int CP_BT = 256;
template <bool PDL>
__global__ void __launch_bounds__(CP_BT) conv_post_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                          const float* __restrict__ bias, const int32_t* __restrict__ lens,
                                                          int lens_mul, int L, int C, int K, float slope,
                                                          float* __restrict__ wav) {
  pdl_entry<PDL>();
  extern __shared__ __align__(16) float cp_smem[];
  const int ld = C + 1;
  float* xs = cp_smem;                          // [(CP_BT + K - 1)][C + 1]
  float* ws = cp_smem + (CP_BT + K - 1) * ld;   // [K][C]
  const int b = blockIdx.y, t0 = blockIdx.x * CP_BT;
  const int len = lens ? min(L, lens[b] * lens_mul) : L;
  const int halo = (K - 1) / 2;
  const float* xb = x + (size_t)b * L * C;
  const int rows = CP_BT + K - 1;
  const int c4n = C / 4;
  for (int i = threadIdx.x; i < rows * c4n; i += CP_BT) {
    const int r = i / c4n, c4 = i % c4n;
    const int row = t0 - halo + r;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row >= 0 && row < len) v = __ldg(reinterpret_cast<const float4*>(xb + (size_t)row * C + c4 * 4));
    float* d = xs + r * ld + c4 * 4;
    d[0] = v.x > 0.f ? v.x : v.x * slope;
    d[1] = v.y > 0.f ? v.y : v.y * slope;
    d[2] = v.z > 0.f ? v.z : v.z * slope;
    d[3] = v.w > 0.f ? v.w : v.w * slope;
  }
  for (int i = threadIdx.x; i < K * C; i += CP_BT) ws[i] = w[i];
  __syncthreads();
  const int t = t0 + threadIdx.x;
  if (t >= L) return;
  float acc = 0.f;
  for (int c = 0; c < C; ++c) {       // channel-major reduction order, shared with the granule-planar conv_post kernels (conv1d_gp.cu)
    const float* xr = xs + threadIdx.x * ld + c;
    for (int j = 0; j < K; ++j) acc = fmaf(xr[j * ld], ws[j * C + c], acc);
  }
  wav[(size_t)b * L + t] = t < len ? tanhf(acc + bias[0]) : 0.f;
}
int launch_conv_post(const float* x, const float* w, const float* bias, const int32_t* lens, int lens_mul, int B,
                     int L, int C, int K, float slope, float* wav, cudaStream_t st) {
  EV_CHECK_ARG(C % 4 == 0 && C <= 128 && K <= 15 && (K & 1), "conv_post: C=%d K=%d", C, K);
  EV_CHECK_ARG(B > 0 && B <= 65535 && L > 0, "conv_post: bad shape");
  const size_t smem = (size_t)((CP_BT + K - 1) * (C + 1) + K * C) * sizeof(float);
  static std::atomic<uint64_t> attr_devs{0};
  if (first_use_on_device(attr_devs)) {
    cudaFuncSetAttribute(conv_post_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    cudaFuncSetAttribute(conv_post_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
  }
  dim3 grid((L + CP_BT - 1) / CP_BT, B);
  launch_k(conv_post_kernel<true>, conv_post_kernel<false>, grid, CP_BT, smem, st, x, w, bias, lens, lens_mul, L, C, K, slope, wav);
  EV_CUDA_LAUNCH_CHECK("conv_post_kernel");
  return EV_OK;
}

// pcm = (int16) trunc(wav * 32768): numpy astype('int16') of a float array is a C cast
// (inference_am_vocoder_joint.py:130-131).  Values are inside (-1, 1) after tanh.
template <bool PDL>
__global__ void pcm16_kernel(const float* __restrict__ wav, int16_t* __restrict__ pcm, size_t n) {
  pdl_entry<PDL>();
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    float v = truncf(wav[i] * 32768.0f);
    v = fminf(fmaxf(v, -32768.f), 32767.f);
    pcm[i] = (int16_t)(int)v;
  }
}
int launch_pcm16(const float* wav, int16_t* pcm, size_t n, cudaStream_t st) {
  if (n == 0) return EV_OK;
  launch_k(pcm16_kernel<true>, pcm16_kernel<false>, (unsigned)((n + 255) / 256), 256, 0, st, wav, pcm, n);
  EV_CUDA_LAUNCH_CHECK("pcm16_kernel");
  return EV_OK;
}

}  // namespace ev
