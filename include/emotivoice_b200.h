/*
 * emotivoice_b200.h -- C ABI of libemotivoice_b200.so (sm_100a).
 *
 * The reference (netease-youdao/EmotiVoice) has no FFI / plugin registry: its
 * boundary for this path is the Python torch.nn.Module API of
 *   JETSGenerator.forward      models/prompt_tts_modified/jets.py:50-71
 *   PromptTTS.forward          models/prompt_tts_modified/model_open_source.py:102-163
 *   Generator.forward          models/hifigan/models.py:115-131
 * The host-side mirror of those classes lives in emotivoice_b200/modules.py and
 * binds the entry points below with ctypes (see INTEGRATION.md for the stub a
 * reference maintainer would add).  Everything here is plain C: device pointers,
 * sizes, a CUDA stream handle passed as void*.  No torch types.
 *
 * Conventions
 *   - every function returns 0 on success or a negative EV_E* code; nothing throws
 *     or aborts across the ABI; ev_last_error() gives the message (thread local).
 *   - all work is enqueued asynchronously on the caller's stream; the library never
 *     synchronises and never allocates device memory: the caller (PyTorch's caching
 *     allocator) owns weights, workspace, inputs and outputs.
 *   - activations are fp32, TIME-MAJOR ("channels last"): a tensor of L steps and C
 *     channels of batch item b lives at base + b*L*C, element (t, c) at t*C + c.
 *   - `lens` arrays are int32 on the device; NULL means "every item is L long"
 *     (the reference's literal padded-batch semantics).  With `lens` given, item b
 *     is treated exactly as the reference treats a B=1 call of that length: rows
 *     >= lens[b] read as zero padding and are written as zeros.
 *   - an ev_ctx is immutable after ev_bind_*; concurrent calls are safe as long as
 *     each call has its own workspace (reference callers are multi-threaded:
 *     openaiapi.py:162-163).
 */
#ifndef EMOTIVOICE_B200_H_
#define EMOTIVOICE_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EV_API __attribute__((visibility("default")))
#define EV_ABI_VERSION 1

enum {
  EV_OK = 0,
  EV_EINVAL = -1,      /* bad argument / unsupported shape */
  EV_ENOWEIGHT = -2,   /* a required tensor is missing from the bound blob */
  EV_ECUDA = -3,       /* CUDA runtime error (message has the cudaError string) */
  EV_EWORKSPACE = -4,  /* workspace too small */
  EV_EPELEN = -5,      /* positional table shorter than the sequence: rebind a longer one */
  EV_EARCH = -6        /* device is not sm_100 */
};

/* activation / epilogue selectors of ev_op_conv1d */
enum { EV_ACT_NONE = 0, EV_ACT_LRELU = 1, EV_ACT_RELU = 2, EV_ACT_GELU = 3, EV_ACT_TANH = 4 };
enum { EV_ACC_STORE = 0, EV_ACC_ADD = 1, EV_ACC_ADD_DIV = 2 };

typedef struct ev_ctx ev_ctx;

/* The integers of config/joint/config.yaml:36-94 (+ n_vocab / n_speaker patched in by
 * every reference caller, inference_am_vocoder_joint.py:57-58). */
typedef struct ev_config {
  int32_t n_vocab, n_speaker;
  int32_t hidden;            /* encoder_n_hidden == decoder_n_hidden == variance_n_hidden (384) */
  int32_t n_heads;           /* 8 */
  int32_t enc_layers, dec_layers;
  int32_t ffn_kernel;        /* *_kernel_size_conv_mod (3) */
  int32_t bert_dim;          /* bert_embedding (768) */
  int32_t dur_layers, pitch_layers, energy_layers;
  int32_t pred_kernel;       /* duration/variance kernel size (3) */
  int32_t embed_kernel;      /* variance_embed_kernel_size (9) */
  int32_t n_mels;            /* 80 */
  int32_t voc_c0;            /* upsample_initial_channel (512) */
  int32_t n_ups;             /* len(upsample_rates) (4) */
  int32_t up_rates[8];
  int32_t up_kernels[8];
  int32_t n_resk;            /* len(resblock_kernel_sizes) (3) */
  int32_t res_kernels[4];
  int32_t n_dil;             /* dilations per ResBlock1 (3) */
  int32_t res_dils[4][4];
} ev_config;

/* One tensor of the packed weight blob (built by emotivoice_b200/packing.py). */
typedef struct ev_weight_entry {
  char name[56];
  uint64_t offset;   /* in floats from the blob base */
  uint64_t numel;
} ev_weight_entry;

EV_API int ev_abi_version(void);
EV_API const char* ev_last_error(void);

/* Replaces: JETSGenerator.__init__ (jets.py:27-47). */
EV_API int ev_create(ev_ctx** out, int device, const ev_config* cfg);
EV_API void ev_destroy(ev_ctx* ctx);

/* Replaces: module.load_state_dict + the per-forward weight_norm recomputation
 * (hifigan/models.py:31-46,96-111).  The blob holds folded / re-laid-out fp32 weights;
 * the engine borrows it (caller keeps it alive). */
EV_API int ev_bind_weights(ev_ctx* ctx, const float* blob, size_t n_floats,
                           const ev_weight_entry* index, int n_entries);
/* Replaces: PositionalEncoding.extend_pe (encoder.py:206-237).  pe is (pe_len, hidden)
 * fp32 on the device, built by the host with the reference's formula. */
EV_API int ev_bind_pe(ev_ctx* ctx, const float* pe, int pe_len);

/* Arithmetic of the GEMM-shaped layers (linear / conv / transposed conv):
 *   EV_PREC_FP32 (default): fp32-accurate on the tcgen05 tensor cores by 3xTF32 splitting (x = hi + lo,
 *     three tf32 MMAs per K step, fp32 accumulation in TMEM); error ~1e-6 relative, like an fp32 FFMA chain.
 *   EV_PREC_TF32: decoder + vocoder with ONE tf32 MMA per K step (operands rounded to nearest tf32) -- the
 *     arithmetic the reference's eager PyTorch uses for convolutions on a GPU (cudnn.allow_tf32 default);
 *     the duration-critical prefix (encoder, conditioning, predictors) stays 3xTF32.
 *   EV_PREC_BF16: decoder + vocoder with bf16 operands (tcgen05 kind::f16, fp32 accumulation; activations stay fp32
 *     in HBM and are rounded by the staging warps); prefix 3xTF32 like EV_PREC_TF32.  BASELINE.json configs[2].
 *   EV_PREC_FP32_FFMA: plain fp32 FFMA kernels everywhere (no tensor cores; the round-1 baseline path).
 * Attention, LayerNorm, softmax, upsampling and the heads are fp32 in every mode. */
enum { EV_PREC_FP32 = 0, EV_PREC_TF32 = 1, EV_PREC_FP32_FFMA = 2, EV_PREC_BF16 = 3 };
EV_API int ev_set_precision(ev_ctx* ctx, int precision);

/* Workspace sizes (bytes).  Phase 1 (encoder .. durations) is sized by (B, T); phase 2
 * (length regulator, decoder, to_mel) and the vocoder share one buffer sized by (B, F) --
 * F is only known after the host has read mel_lens_out[B] back. */
EV_API size_t ev_phase1_workspace_bytes(const ev_ctx* ctx, int B, int T);
EV_API size_t ev_phase2_workspace_bytes(const ev_ctx* ctx, int B, int F);

/* Replaces: PromptTTS.forward up to the duration prediction
 * (model_open_source.py:102-134; encoder.py:316-324; variance.py:36-56,101-124) plus
 * the cumsum / length bookkeeping of GaussianUpsampling (alignment.py:183-199).
 *   ling (B,T) i64; lens (B) i64 true phoneme counts (as the reference passes them); spk (B) i64;
 *   style, content (B,bert) f32
 *   dur_out (B,T) i64; pitch_out / energy_out (B,T) f32;
 *   lens32_out (B) i32: lens clamped to [0,T] (input of the later phases);
 *   mel_lens_out (B+2) i32: per-item frame counts, max over items in slot B, and in slot B+1 an input
 *     status word the host checks at the same read-back (the reference raises IndexError from nn.Embedding
 *     for these; the kernels clamp so nothing is read out of bounds): bit 0 = a token id outside
 *     [0, n_vocab), bit 1 = a speaker id outside [0, n_speaker), bit 2 = a length outside [1, T].
 *   invariant != 0: batch-invariant contract (each item == the reference's B=1 call);
 *   invariant == 0: literal padded-batch forward of the reference. */
EV_API int ev_am_phase1(ev_ctx* ctx, const int64_t* ling, const int64_t* lens, const int64_t* spk,
                        const float* style, const float* content, int B, int T, int invariant,
                        int64_t* dur_out, float* pitch_out, float* energy_out, int32_t* lens32_out,
                        int32_t* mel_lens_out, void* workspace, size_t workspace_bytes, void* stream);

/* Replaces: GaussianUpsampling.forward matmul (alignment.py:201-211), the decoder
 * (model_open_source.py:146) and to_mel (:147).  Must follow ev_am_phase1 on the same stream;
 * phase1_workspace is the (still live) buffer phase 1 ran on.  F = mel_lens_out[B] read back
 * by the host (the path's one sync).
 *   mel_out (B,F,n_mels) f32 time-major == outputs["dec_outputs"]. */
EV_API int ev_am_phase2(ev_ctx* ctx, const void* phase1_workspace, const int32_t* lens32, const int32_t* mel_lens,
                        int B, int T, int F, int invariant, float* mel_out, void* workspace,
                        size_t workspace_bytes, void* stream);

/* Replaces: Generator.forward (hifigan/models.py:115-131).
 *   mel (B,F,n_mels) time-major if mel_time_major else (B,n_mels,F) (the reference layout);
 *   mel_lens (B) i32 or NULL (literal padded semantics); wav_out (B, F*prod(up_rates)) f32.
 *   mel_lens must be COMPLETE when this is called (not pending in a kernel still running on the stream): the vocoder's kernels start
 *   under their predecessors' tails (programmatic dependent launch) and read the lengths before they wait.  The engine's own flow
 *   satisfies this by construction -- the host reads mel_lens_out back to learn F before it can call ev_am_phase2 / ev_vocoder.  The
 *   same holds for the `lens` argument of the ev_op_*_gp entry points below. */
EV_API int ev_vocoder(ev_ctx* ctx, const float* mel, int mel_time_major, const int32_t* mel_lens, int B, int F,
                      float* wav_out, void* workspace, size_t workspace_bytes, void* stream);

/* Replaces the callers' post-processing (inference_am_vocoder_joint.py:130-131):
 * pcm[i] = (int16) trunc(wav[i] * 32768), n elements. */
EV_API int ev_wav_to_pcm16(const float* wav, int16_t* pcm, size_t n, void* stream);

/* Number of kernel launches this library has enqueued in this process (bench.py's
 * `gpu_launches`). */
EV_API uint64_t ev_launch_count(void);

/* ---- single operators (used by the per-kernel parity tests) ------------------- */

/* Generic time-major 1-D convolution / linear layer (implicit GEMM):
 *   out[b,t,co] = epi( bias[co] + sum_{j<K} sum_{ci} w[j][ci][co] * act_in(x[b, t+(j-(K-1)/2)*dil, ci]) )
 * x rows outside [0, len_b) read as zero.  w is (K, Cin, Cout).  bias_bstride: floats between
 * per-item biases (0 = shared).  res (same layout as out) is added after out_act.
 * acc: EV_ACC_STORE out=v; EV_ACC_ADD out+=v; EV_ACC_ADD_DIV out=(out+v)/div.
 * Covers nn.Linear (K=1), the conv-FFN (encoder.py:50-52), predictor convs (variance.py:17-31),
 * conv_pre / ResBlock1 convs (hifigan/models.py:50-57,116) and, with polyphase-packed
 * weights, ConvTranspose1d (hifigan/models.py:100-103). */
EV_API int ev_op_conv1d(const float* x, const float* w, const float* bias, size_t bias_bstride,
                        const float* res, float* out, int B, int L, int Cin, int Cout, int K, int dil,
                        const int32_t* lens, int lens_mul, int in_act, float in_slope, int out_act,
                        int acc, float div, void* stream);
/* Same contract on the tensor cores (tcgen05.mma kind::tf32, accumulator in TMEM); w_tc is in the
 * tensor-core layout (2 planes hi|lo, Cout/BNp N tiles, K, Cin/4, BNp = min(Cout,128), 4; packing.to_tc_layout);
 * split3 = 0: 1xTF32, 1: 3xTF32 fp32 emulation, 2: bf16 operands (w_tc then in the bf16 layout of
 * packing.to_tc16_layout; Cin % 16 == 0).  Requires Cin % 8 == 0, Cout % 16 == 0 and Cout <= 128 or Cout % 128 == 0.  splitk_ws (optional, splitk_floats floats of scratch) lets a
 * launch with few output tiles and a long reduction be split along K (deterministic two-pass). */
EV_API int ev_op_conv1d_tc(const float* x, const float* w_tc, int split3, const float* bias, size_t bias_bstride,
                           const float* res, float* out, int B, int L, int Cin, int Cout, int K, int dil,
                           const int32_t* lens, int lens_mul, int in_act, float in_slope, int out_act,
                           int acc, float div, float* splitk_ws, size_t splitk_floats, void* stream);
/* Host-only introspection (no GPU needed): the tile / pipeline plan ev_op_conv1d_tc would use for a shape.
 * out11 = {BN, MT, KBG, a_stages, b_stages, producer groups, ksplit, tmem columns, smem bytes, tiles, rows_pad}.
 * The CPU tests check the invariants the kernel relies on (ring depth >= producer groups, TMEM/smem limits,
 * summation-order parameters independent of batch and length). */
EV_API int ev_debug_tc_plan(int B, int L, int Cin, int Cout, int K, int dil, int split3, int ksplit, int* out11);
/* HiFi-GAN convolution on GRANULE-PLANAR activations (csrc/conv1d_gp.cu; what ev_vocoder runs in every tensor-core mode).
 * Layout: a (B, L, C) tensor is stored [b][C/cpg][l][cpg] in 16-byte granules, cpg = 4 fp32 (mode 0 tf32, 1 3xTF32) or 8 bf16
 * (mode 2); mode 3 = "bf16x3": fp32 activations, every operand split into bf16 hi + lo, three bf16 MMAs per K step (fp32-class
 * result at half the cost of 3xTF32; w then holds two bf16 planes).  w: the ev_op_conv1d_tc weight layout of that mode.  rate > 1: polyphase ConvTranspose1d, the Cout GEMM columns
 * are `rate` phases of Cout/rate channels, out is (B, (Cout/rate)/cpg, L*rate, cpg).  res (rate == 1): same shape as out.
 * Rows >= lens[b]*lens_mul are neither read (they count as zero padding) nor written.  Replaces hifigan/models.py:50-57,
 * :116, :118-119. */
EV_API int ev_op_conv1d_gp(const void* x, const float* w, int mode, const float* bias, const void* res, void* out, int B, int L,
                           int Cin, int Cout, int K, int dil, int rate, const int32_t* lens, int lens_mul, int in_act,
                           float in_slope, int acc, float div, void* stream);
/* n <= 3 convolutions of ONE shape (B, L, Cin -> Cout, rate 1, plain store) but different taps / dilations / weights / tensors as ONE
 * launch: the same-index convolutions of the three parallel ResBlocks of a HiFi-GAN stage (hifigan/models.py:120-126), which at small
 * batch have too few tiles each to fill the machine.  Tables of n entries; bias / res may be null (or hold nulls).  Every tile is
 * computed as in the member's own ev_op_conv1d_gp launch: bitwise equal.  EV_EINVAL if the members cannot share a launch. */
EV_API int ev_op_conv1d_gp_group(int n, const void* const* x, const float* const* w, int mode, const float* const* bias,
                                 const void* const* res, void* const* out, const int* K, const int* dil, int B, int L, int Cin,
                                 int Cout, const int32_t* lens, int lens_mul, int in_act, float in_slope, void* stream);
/* Host-only: the plan of a grouped launch, out11 as ev_debug_gp_plan (tiles = all members'). */
EV_API int ev_debug_gp_group_plan(int n, const int* K, const int* dil, int B, int L, int Cin, int Cout, int mode, int* out11);
/* One ResBlock1 layer (hifigan/models.py:50-57) as ONE kernel on granule-planar activations (csrc/resblock_gp.cu):
 * out = [acc]( x + c2(lrelu(c1(lrelu(x), dil)), 1) ), C -> C channels (C in {32, 64, 128}), slope 0.1, both weights in the layout of
 * `mode` (as ev_op_conv1d_gp).  Bitwise equal to the two ev_op_conv1d_gp launches it replaces; EV_EINVAL for shapes it does not take
 * (the engine then runs the pair unfused). */
EV_API int ev_op_resblock_gp(const void* x, const float* w1, const float* b1, const float* w2, const float* b2, int mode, void* out, int B,
                             int L, int C, int K, int dil, const int32_t* lens, int lens_mul, int acc, float div, void* stream);
/* n <= 3 such layers of ONE shape (B, L, C; plain store) with their own taps / dilations / weights / tensors as ONE launch: the
 * same-index layers of HiFi-GAN's three parallel ResBlocks at small batch.  Bitwise equal to the members' own launches; EV_EINVAL if
 * they cannot share a launch (each member must itself be a shape ev_op_resblock_gp takes with at least two accumulators per tile). */
EV_API int ev_op_resblock_gp_group(int n, const void* const* x, const float* const* w1, const float* const* b1, const float* const* w2,
                                   const float* const* b2, int mode, void* const* out, int B, int L, int C, const int* K, const int* dil,
                                   const int32_t* lens, int lens_mul, void* stream);
/* Host-only: the plan of a grouped fused launch: out16 = {MT, KBG, tiles (all members), rows1_pad, rows2_pad, smem bytes, tmem columns,
 * then per member in launch order (heaviest first) {K, row tiles per item, first tile index}}. */
EV_API int ev_debug_resblock_gp_group_plan(int n, const int* K, const int* dil, int B, int L, int C, int mode, int* out16);
/* Host-only: out11 = {MT, KBG, x stages, weight stages, transform warps, tmem columns, smem bytes, tiles, rows per tile, rows1_pad, rows2_pad}. */
EV_API int ev_debug_resblock_gp_plan(int B, int L, int C, int K, int dil, int mode, int* out11);
/* Host-only: out11 = {BN, MT, KBG, a_stages, b_stages, transform warps, planes, tmem columns, smem bytes, tiles, rows_pad}. */
EV_API int ev_debug_gp_plan(int B, int L, int Cin, int Cout, int K, int dil, int rate, int mode, int* out11);
/* fp32 in[b*stride_b + t*stride_t + c*stride_c] -> granule-planar (fp32, or bf16 when bf16 != 0): the vocoder's input boundary
 * (hifigan/models.py:115 takes (B, n_mels, F); jets.py:62 hands over dec_outputs (B, F, n_mels) transposed). */
EV_API int ev_op_to_gp(const float* in, long long stride_b, long long stride_t, long long stride_c, void* out, int B, int L, int C,
                       int bf16, void* stream);
/* wav[b,t] = tanh(bias + conv_post(leaky_relu(x, slope))) on a granule-planar input (hifigan/models.py:127-129). */
EV_API int ev_op_conv_post_gp(const void* x, int bf16, const float* w, const float* bias, const int32_t* lens, int lens_mul, int B,
                              int L, int C, int K, float slope, float* wav, void* stream);
/* LayerNorm over the last dim, eps 1e-12 (encoder.py:112-127). rows x C. */
EV_API int ev_op_layernorm(const float* x, const float* w, const float* b, float* y, int rows, int C, void* stream);
/* Multi-head self-attention core (encoder.py:84-109) on a packed (B,L,3H) q|k|v buffer. */
EV_API int ev_op_attention(const float* qkv, const int32_t* key_lens, float* ctx_out, int B, int L, int H,
                           int n_heads, void* stream);
/* The same attention on the tensor cores (csrc/attention_tc.cu): QK^T and PV as tcgen05.mma with the softmax between two TMEM
 * reads; d_k = 48 only.  tc_mode 1 = 3xTF32 fp32 emulation (what the engine uses wherever a layer runs fp32-accurate),
 * 0 = one tf32 MMA per K step. */
EV_API int ev_op_attention_tc(const float* qkv, const int32_t* key_lens, float* ctx_out, int B, int L, int H, int n_heads,
                              int tc_mode, void* stream);
/* Gaussian upsampling (alignment.py:180-211) incl. cumsum; out (B,F,H); adds alpha*pe[f] when pe != NULL.
 * centers_tmp: 2*B*T floats of scratch; mel_lens_tmp: B+1 int32 (frame counts, max in slot B). */
EV_API int ev_op_gauss_upsample(const float* hs, const int64_t* dur, const int32_t* lens, int B, int T, int H,
                                int F, int invariant, const float* pe, const float* alpha, float* centers_tmp,
                                int32_t* mel_lens_tmp, float* out, void* stream);

/* ---- style encoder (the callers' prompt / content embedding: simbert.py:33-72 -> transformers BertModel) ---------------
 * Next-row widening: the reference runs this BERT-base on the CPU twice per utterance
 * (inference_am_vocoder_joint.py:25-38,106-107).  Separate context: it is a separate model with its own checkpoint. */
typedef struct ev_style_ctx ev_style_ctx;

/* The integers of the checkpoint's BertConfig + the width of the packed classification heads. */
typedef struct ev_style_config {
  int32_t vocab_size, max_position, type_vocab;
  int32_t hidden;            /* 768; multiple of 128, <= 768 */
  int32_t n_heads;           /* 12; hidden / n_heads in {32, 48, 64} */
  int32_t n_layers;          /* 12 */
  int32_t intermediate;      /* 3072; multiple of 128 */
  int32_t n_head_out;        /* columns of the packed [pitch|speed|energy|emotion] classifier (multiple of 8), 0 = none */
} ev_style_config;

/* StyleEncoder.__init__ (simbert.py:34-44). */
EV_API int ev_style_create(ev_style_ctx** out, int device, const ev_style_config* cfg);
EV_API void ev_style_destroy(ev_style_ctx* ctx);
/* load_state_dict (inference_am_vocoder_joint.py:60-65): same blob + index convention as ev_bind_weights; names are
 * the ones emotivoice_b200.packing.pack_style_state_dict emits ("sty.*"). */
EV_API int ev_style_bind_weights(ev_style_ctx* ctx, const float* blob, size_t blob_floats, const ev_weight_entry* index,
                                 int n_entries);
/* EV_PREC_FP32 (3xTF32 on tcgen05, default) or EV_PREC_TF32. */
EV_API int ev_style_set_precision(ev_style_ctx* ctx, int precision);
EV_API size_t ev_style_workspace_bytes(const ev_style_ctx* ctx, int B, int N);
/* StyleEncoder.forward (simbert.py:48-72): ids / type_ids (B,N) int64, lens (B,) int64 = number of leading tokens with
 * attention_mask == 1 (tokenizer padding is a suffix).  pooled (B, hidden) = BertModel's pooler_output;
 * heads (B, n_head_out) = the four classification heads side by side, or NULL to skip them. */
EV_API int ev_style_forward(ev_style_ctx* ctx, const int64_t* ids, const int64_t* type_ids, const int64_t* lens, int B, int N,
                            float* pooled, float* heads, void* workspace, size_t workspace_bytes, void* stream);

/* ---- training-mode alignment helpers (not used by any inference path) --------------------------------
 * viterbi_decode (alignment.py:124-142): monotonic alignment search per item on log_p_attn (B, T_mel, T_inp) float32 restricted
 * to [:feats_lens[b], :text_lens[b]].  path (B, T_mel) int32 (token per frame, -1 past the item), durations (B, T_inp) float32
 * = bincount(path), bin_loss (B) = -mean_j log_p[j, path[j]] per item (the reference averages them over B).
 * Bit-exact paths / durations vs the reference's numba loop (float64 scores, float32 row-0 sums, ties -> smaller index).
 * workspace: B*T_mel*T_inp bytes. */
EV_API int ev_op_mas(const float* log_p_attn, const int64_t* text_lens, const int64_t* feats_lens, int B, int T_mel, int T_inp,
                     int32_t* path, float* durations, float* bin_loss, uint8_t* workspace, size_t workspace_bytes, void* stream);
/* average_by_duration (alignment.py:145-177): out (B, T_inp) = per-token mean of xs (B, T_mel) over the token's frames. */
EV_API int ev_op_average_by_duration(const float* durations, const float* xs, const int64_t* text_lens, const int64_t* feats_lens,
                                     int B, int T_mel, int T_inp, float* out, void* stream);
/* AlignmentModule.forward after its convolutions (alignment.py:39-56): log_p_attn[b,f,t] = log_softmax_t( -||feats_feat[b,f,:] -
 * text_feat[b,t,:]||_2 ) with tokens t >= text_lens[b] masked to -inf, + prior (B,T_mel,T_inp; may be NULL).  text_feat (B,T_inp,A),
 * feats_feat (B,T_mel,A) time-major fp32, A a multiple of 128.  Training only. */
EV_API int ev_op_align_logp(const float* text_feat, const float* feats_feat, const int64_t* text_lens, const float* prior, int B, int T_mel,
                            int T_inp, int A, float* log_p_attn, void* stream);
/* get_segments (models/hifigan/get_random_segments.py:19-27; jets.py:55-60): out[b,c,i] = x[b,c,start_idxs[b]+i], zero past T.
 * x (B,C,T) channels-first, out (B,C,segment_size). */
EV_API int ev_op_get_segments(const float* x, const int64_t* start_idxs, int B, int C, int T, int segment_size, float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* EMOTIVOICE_B200_H_ */
