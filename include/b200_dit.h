/* b200_dit.h — C ABI of libb200dit.so: the sm_100a kernels behind LightX2V's DiT operator surface.
 *
 * Conventions:
 *   - plain device pointers + int64 shapes + a cudaStream_t; no torch types, no allocation, no implicit sync;
 *   - every kernel is launched on the stream passed in (the caller passes torch.cuda.current_stream(), because
 *     the reference runs blocks under `torch.cuda.stream(compute_stream)`, transformer_infer.py:92);
 *   - return 0 on success, a negative code otherwise; b200_last_error() returns a thread-local message
 *     (the reference's native ABI raises through TORCH_CHECK, lightx2v_kernel/csrc/gemm/nvfp4_scaled_mm_kernels_sm120.cu:253-273;
 *     the Python shim turns a non-zero return into RuntimeError);
 *   - bf16 tensors are row-major with an explicit leading dimension in ELEMENTS; pointers 16-byte aligned.
 *
 * Each entry point cites the reference interface it replaces (paths relative to the LightX2V tree).
 */
#ifndef B200_DIT_H_
#define B200_DIT_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* b200_stream_t; /* == cudaStream_t */

#define B200_OK 0
#define B200_ERR_INVALID (-1)
#define B200_ERR_CUDA (-2)
#define B200_ERR_UNSUPPORTED (-3)

/* Library / device introspection. */
const char* b200_last_error(void);
int b200_version(void);
int b200_num_sms(void);

/* Number of kernel launches this library has issued in this process so far (all entry points, incl. the ones issued below
 * b200_wan_block_fwd).  bench.py reports the difference over its timed region as `gpu_launches`. */
int64_t b200_launch_count(void);

/* Runtime A/B switches (measurements and tests; initial values from the environment variables of the same meaning): "conv_halo" (halo-staged
 * tiles for the wide 3x3x3 VAE convolutions, B200_CONV_HALO), "halo_base_offset" (B200_HALO_BASE_OFFSET), "conv_narrow" (round-1 32-channel-chunk
 * tiles, B200_CONV_NARROW).  b200_get_option returns the current value, -1 for an unknown name. */
int b200_set_option(const char* name, int value);
int b200_get_option(const char* name);

/* Attention launch profiler for bench.py's roofline: after b200_prof_fmha_begin(capacity) every b200_fmha_fwd_* launch (also the ones
 * issued inside b200_wan_block_fwd) is bracketed by a CUDA event pair on ITS stream, up to `capacity` launches.
 * b200_prof_fmha_end waits for them, writes ms[i] and meta[4 i + {0,1,2,3}] = {sq, sk, heads, head_dim} for the first `capacity`
 * launches, disables the profiler and returns the number written.  Off by default (no events, no overhead). */
int b200_prof_fmha_begin(int capacity);
int b200_prof_fmha_end(float* ms, int64_t* meta, int capacity);

/* GEMM epilogues (fusions of the elementwise passes that follow each linear in WanTransformerInfer). */
#define B200_EPI_BIAS 0          /* C = bf16(A B^T + bias)                         mm_weight.py:81-88            */
#define B200_EPI_BIAS_GELU 1     /* C = gelu_tanh(bf16(A B^T + bias))              + transformer_infer.py:492     */
#define B200_EPI_GATE_RESIDUAL 2 /* C = C + bf16(bf16(A B^T + bias) * gate[n])     + transformer_infer.py:402,503 */
#define B200_EPI_RESIDUAL 3      /* C = C + bf16(A B^T + bias)                     + transformer_infer.py:468     */

/* C[M,N] = epilogue(A[M,K] * B[N,K]^T + bias[N]).  A, B, C bf16 row-major (B is the checkpoint's [N,K] weight, the
 * storage MMWeight.load keeps as a transposed view, lightx2v/common/ops/mm/mm_weight.py:76).  bias / gate may be NULL
 * (gate required for B200_EPI_GATE_RESIDUAL).  K, N multiples of 8.  block_n: 0 = auto, 128 or 256.  max_ctas: 0 = one
 * per SM.  Replaces MMWeight.apply (mm_weight.py:81-88). */
int b200_gemm_bf16(const void* A, int64_t lda, const void* B, int64_t ldb, void* C, int64_t ldc, const void* bias,
                   const void* gate, int64_t M, int64_t N, int64_t K, int epilogue, int block_n, int max_ctas,
                   b200_stream_t stream);

/* y = LayerNorm(x; eps) [* ln_w + ln_b]  then optionally  y = bf16(bf16(y * bf16(1 + scale)) + shift).
 * ln_w/ln_b NULL -> no affine (norm1/norm2); scale/shift NULL -> no modulation (norm3).
 * Replaces LNWeight.apply (lightx2v/common/ops/norm/layer_norm_weight.py:100-111) + the in-place
 * `norm_out.mul_(1+scale).add_(shift)` (lightx2v/models/networks/wan/infer/transformer_infer.py:326-334,478-484). */
int b200_ln_modulate(const void* x, int64_t ldx, void* y, int64_t ldy, const void* ln_w, const void* ln_b,
                     const void* scale, const void* shift, int64_t rows, int D, float eps, b200_stream_t stream);

/* In-place RMSNorm over the full row (D) of x0 (and x1 if non-NULL, e.g. q and k inside one fused QKV buffer),
 * bf16 arithmetic as RMSWeightSgl.apply's torch fallback (lightx2v/common/ops/norm/rms_norm_weight.py:102-118), then,
 * if cos_sin != NULL, RoPE on adjacent pairs per 128-wide head with table cos_sin[(row + pos_offset), 64] float2
 * (apply_rotary_emb, lightx2v/models/networks/wan/infer/utils.py:107-115) for rows < rope_rows. */
int b200_rms_rope(void* x0, int64_t ld0, const void* w0, void* x1, int64_t ld1, const void* w1, int64_t rows, int D,
                  float eps, const void* cos_sin, int64_t rope_rows, int64_t pos_offset, b200_stream_t stream);

/* out[sq, H, 128] = softmax(q k^T * softmax_scale) v for one varlen segment, non-causal, head_dim 128.
 * q/k/v/out are [rows, H, 128] bf16 with row strides in elements (heads contiguous).
 * Replaces FlashAttn2Weight.apply / flash_attn_varlen_func (lightx2v/common/ops/attn/attn_weight.py:71-97). */
int b200_fmha_fwd_d128(const void* q, int64_t q_stride_s, const void* k, int64_t k_stride_s, const void* v,
                       int64_t v_stride_s, void* out, int64_t o_stride_s, int64_t sq, int64_t sk, int heads,
                       float softmax_scale, b200_stream_t stream);

/* Same kernel instantiated for head_dim 64: q/k/v/out [rows, H, 64].  CogVideoX's joint text+video attention (48 heads x 64,
 * lightx2v/models/networks/cogvideox/infer/transformer_infer.py:118-132: F.scaled_dot_product_attention over [text ; video] tokens). */
int b200_fmha_fwd_d64(const void* q, int64_t q_stride_s, const void* k, int64_t k_stride_s, const void* v, int64_t v_stride_s,
                      void* out, int64_t o_stride_s, int64_t sq, int64_t sk, int heads, float softmax_scale, b200_stream_t stream);

/* ---- w8a8-fp8 path (BASELINE config 3) ------------------------------------------------------------------------------- */

/* C[M,N] (bf16) = epilogue( a_scale[m] * (b_scale[n] * (A[M,K] * B[N,K]^T)) + bias[N] ),  A and B e4m3 (1 byte/element, leading
 * dimensions in elements, multiples of 16), a_scale [M] fp32 per-token, b_scale [N] fp32 per-out-channel.  Same epilogues as
 * b200_gemm_bf16.  Replaces `torch.ops._C.cutlass_scaled_mm(out, xq, Wq, sx, sw, bias)` in
 * MMWeightWfp8channelAfp8channeldynamicVllm.apply (lightx2v/common/ops/mm/mm_weight.py:304-319). */
int b200_gemm_fp8(const void* A, int64_t lda, const void* B, int64_t ldb, void* C, int64_t ldc, const float* a_scale,
                  const float* b_scale, const void* bias, const void* gate, int64_t M, int64_t N, int64_t K, int epilogue,
                  int block_n, int max_ctas, b200_stream_t stream);

/* Dynamic per-token e4m3 quantisation: scale[r] = max(absmax(x[r,:]) / 448, 1/(448*512)); q = e4m3_rn_sat(x / scale).
 * Replaces `ops.scaled_fp8_quant(x, None, scale_ub=None, use_per_token_if_dynamic=True)`
 * (act_quant_fp8_perchannel_sym_vllm, lightx2v/common/ops/mm/mm_weight.py:236-238). */
int b200_quant_fp8_per_token(const void* x, int64_t ldx, void* q8, int64_t ldq, float* q_scale, int64_t rows, int D,
                             b200_stream_t stream);

/* b200_ln_modulate whose output goes straight to e4m3 + per-token scale (the bf16 tensor the reference would quantise is
 * formed in registers and never written): LNWeight.apply + modulation + act quant in one pass. */
int b200_ln_modulate_fp8(const void* x, int64_t ldx, void* q8, int64_t ldq, float* q_scale, const void* ln_w,
                         const void* ln_b, const void* scale, const void* shift, int64_t rows, int D, float eps,
                         b200_stream_t stream);

/* In-place per-head RMSNorm of x0 (and x1) laid out [rows, H, 128] (weight [128], bf16 chain), then for rows < rope_rows
 * x*cos + rotate_half(x)*sin in bf16 arithmetic with table cos_sin[row, 64] float2 (bf16-representable values).  HunyuanVideo's
 * q/k path: RMSWeight over the head dim + apply_rotary_emb (lightx2v/models/networks/hunyuan/infer/transformer_infer.py:289-295,
 * 342-347; lightx2v/models/networks/hunyuan/infer/utils_bf16.py:5-31). */
int b200_rms_rope_heads(void* x0, int64_t ld0, const void* w0, void* x1, int64_t ld1, const void* w1, int64_t rows, int H,
                        float eps, const void* cos_sin, int64_t rope_rows, b200_stream_t stream);

/* In-place per-head affine LayerNorm over head_dim 64 of x0 (and x1) laid out [rows, H, 64] (weight / bias [64] bf16; F.layer_norm in
 * bf16), then for rows >= rope_start the pair rotation bf16(x * cos + rotate(x) * sin) in fp32 with table cos_sin[row - rope_start, 32]
 * float2.  CogVideoX's q/k path: attn1.norm_q / norm_k + apply_rotary_emb on the video tokens of the joint [text ; video] sequence
 * (lightx2v/models/networks/cogvideox/infer/transformer_infer.py:5-36, 99-103). */
int b200_ln_rope_heads64(void* x0, int64_t ld0, const void* w0, const void* b0, void* x1, int64_t ld1, const void* w1, const void* b1,
                         int64_t rows, int H, float eps, const void* cos_sin, int64_t rope_start, b200_stream_t stream);

/* ---- Ulysses sequence parallelism fused into the kernels (peer memory over NVLink / NVSwitch) ------------------------------ */

/* q/k RMSNorm + RoPE (and a copy of v) of the local token shard qkv[rows, 3*D], every 16-byte vector stored directly into the
 * peer GPU that owns its head: peers[r] is rank r's receive buffer [world*rows_per_rank, 3, H/world, 128] (peer-mapped device
 * pointers, host array of `world` entries).  Replaces all2all_seq2head x3 (+ transposes + cuda.synchronize) of ulysses_attn
 * (lightx2v/attentions/distributed/ulysses/attn.py:41-48, lightx2v/attentions/distributed/comm/all2all.py:7-44). */
int b200_rms_rope_scatter(const void* qkv, int64_t ld, const void* wq, const void* wk, int64_t rows, int D, float eps,
                          const void* cos_sin, int64_t rope_rows, void* const* peers, int world, int rank,
                          int64_t rows_per_rank, b200_stream_t stream);

/* b200_fmha_fwd_d128 whose epilogue scatters each query row to the rank that owns the token: row r -> peers[r / rows_per_rank]
 * at [(r % rows_per_rank), head_offset + head, :] with row stride peer_stride_s.  world <= 8 (one NVSwitch domain of a B200 box; the
 * kernel parameter block carries 8 peer pointers).  Replaces the attention + all2all_head2seq of
 * ulysses_attn (attn.py:51-88, all2all.py:48-89). */
int b200_fmha_fwd_d128_scatter(const void* q, int64_t q_stride_s, const void* k, int64_t k_stride_s, const void* v,
                               int64_t v_stride_s, void* const* peers, int world, int64_t rows_per_rank, int64_t peer_stride_s,
                               int head_offset, int64_t sq, int64_t sk, int heads, float softmax_scale, b200_stream_t stream);

/* ---- Wan 3-D causal VAE decoder (channels-last bf16 activations [T, H, W, C]) ---------------------------------------- */

/* out[t,h,w,n] = bias[n] + sum_tap sum_c in[t+dt, h+dh, w+dw, c] * wt[n, tap*cin + c] (+ residual[t,h,w,n]) (clamped to [-1,1] if
 * clamp_out).  in / out / residual are channels-last VIEWS given by a base pointer and element strides for (t, h, w); reads
 * outside [0,T)x[0,H)x[0,W) are zero (the convolution's zero padding, causal along t).  taps: ntaps x (dt, dh, dw) int32.
 * cin multiple of 32, cout 16 or a multiple of 64 / 96 / 192.  Replaces CausalConv3d.forward / nn.Conv2d of the decoder
 * (lightx2v/models/video_encoders/hf/wan/vae.py:19-44, 88-97, 185-223) incl. the two-frame cache protocol (:203-217). */
int b200_conv3d_cl(const void* in, int64_t in_st, int64_t in_sh, int64_t in_sw, const void* wt, const void* bias, void* out,
                   int64_t out_st, int64_t out_sh, int64_t out_sw, const void* residual, int64_t res_st, int64_t res_sh,
                   int64_t res_sw, int T, int H, int W, int cin, int cout, int ntaps, const int32_t* taps, int clamp_out,
                   b200_stream_t stream);

/* ---- one Wan DiT block as a single call ------------------------------------------------------------------------------------ */

/* Weights of one block, bf16, checkpoint layout [N, K] row-major (blocks.<i>.* keys of wan/weights/transformer_weights.py:62-250). */
typedef struct b200_wan_block_weights {
  const void *wqkv, *bqkv;           /* self_attn.{q,k,v}.weight concatenated along N: [3D, D], [3D] */
  const void *norm_q, *norm_k;       /* self_attn.norm_{q,k}.weight [D] */
  const void *wo, *bo;               /* self_attn.o [D, D] */
  const void *norm3_w, *norm3_b;     /* norm3 (affine LayerNorm) [D] */
  const void *wcq, *bcq, *cnorm_q;   /* cross_attn.q [D, D], cross_attn.norm_q.weight [D] */
  const void *wco, *bco;             /* cross_attn.o [D, D] */
  const void *w0, *b0;               /* ffn.0 [F, D] */
  const void *w2, *b2;               /* ffn.2 [D, F] */
} b200_wan_block_weights;

typedef struct b200_wan_block_args {
  void* x;                                                                            /* residual stream [S, D], updated in place */
  const void *shift_msa, *scale_msa, *gate_msa, *c_shift_msa, *c_scale_msa, *c_gate_msa; /* (modulation + embed0).chunk(6), [D] each */
  const void* cos_sin;                                                                /* RoPE table [rope_rows, 64] float2 */
  int64_t rope_rows;
  const void *ctx_k, *ctx_v;                                                          /* text K (post norm_k) / V [ctx_len, H, 128] */
  int64_t ctx_len;
  const void *img_k, *img_v;                                                          /* i2v: CLIP K / V [img_len, H, 128] or NULL */
  int64_t img_len;
  int64_t S;
  int D, H, F;
  float eps;
} b200_wan_block_args;

int64_t b200_wan_block_workspace_bytes(int64_t S, int D, int F);

/* x <- block(x): the launch schedule of WanTransformerInfer.infer_block (lightx2v/models/networks/wan/infer/transformer_infer.py
 * :289-508: infer_modulation's outputs are inputs here; infer_self_attn, infer_cross_attn, infer_ffn, post_process) issued natively on
 * `stream`: 13 kernel launches (15 for i2v), no allocation, no synchronisation; intermediates in `workspace`
 * (b200_wan_block_workspace_bytes).  bf16 linears, single GPU; the quantised and sequence-parallel variants are composed from the
 * per-op entry points by the host. */
int b200_wan_block_fwd(const b200_wan_block_weights* w, const b200_wan_block_args* a, void* workspace, int64_t workspace_bytes,
                       b200_stream_t stream);

/* ---- w4a4 NVFP4 linears ------------------------------------------------------------------------------------------------- */

/* bf16 x[rows, K] -> packed e2m1 q[rows, K/2] (two values per byte, low nibble = even element) + ue4m3 scale factors, one per 16
 * elements, in the 128x4 layout of the block-scaled MMA: byte ((m/128)*(K/64) + g/4)*512 + (m%32)*16 + ((m%128)/32)*4 + g%4 for row m,
 * group g; sf must hold roundup(rows,128)*K/16 bytes (padding rows are written as 0).  *global_scale (device fp32) = 448*6/amax(x).
 * Replaces scaled_fp4_quant (lightx2v_kernel/python/lightx2v_kernel/gemm.py:11-52, csrc/gemm/nvfp4_quant_kernels_sm120.cu:118-290);
 * arithmetic of lightx2v_kernel/test/nvfp4_nvfp4/fake_quant.py:37-51, bit for bit.  K multiple of 64. */
int b200_quant_nvfp4(const void* x, int64_t ldx, int64_t rows, int K, const float* global_scale, void* q, int64_t ldq, void* sf,
                     b200_stream_t stream);

/* Dynamic per-tensor scales on the device: *global_scale = 448*6 / max|x| (lightx2v_kernel/docs/en_US/nvfp4_quantization_basics.md:52,
 * test_bench1.py:120-121), *alpha = 1 / (*global_scale * *weight_global_scale) (test_bench1.py:126; alpha may be NULL).  scratch: 4 bytes. */
int b200_nvfp4_act_scale(const void* x, int64_t ldx, int64_t rows, int K, const float* weight_global_scale, float* global_scale,
                         float* alpha, void* scratch, b200_stream_t stream);

/* C[M,N] = epilogue( (*alpha) * sum_k (A_q[m,k] sfa[m,k/16]) (B_q[n,k] sfb[n,k/16]) + bias[n] ), bf16 out, fp32 accumulate in TMEM
 * (tcgen05.mma kind::mxf4nvf4.block_scale).  A [M,K/2], B [N,K/2] packed e2m1 (lda / ldb in bytes), scale factors as written by
 * b200_quant_nvfp4, *alpha = 1 / (global_scale_a * global_scale_b) (device fp32).  Epilogues as b200_gemm_bf16.  K multiple of 64.
 * Replaces cutlass_scaled_fp4_mm (lightx2v_kernel/python/lightx2v_kernel/gemm.py:4-8, csrc/gemm/nvfp4_scaled_mm_kernels_sm120.cu). */
int b200_gemm_nvfp4(const void* A, int64_t lda, const void* B, int64_t ldb, const void* sfa, const void* sfb, const float* alpha, void* C,
                    int64_t ldc, const void* bias, const void* gate, int64_t M, int64_t N, int64_t K, int epilogue, int block_n,
                    int max_ctas, b200_stream_t stream);

/* ---- HunyuanVideo causal 3-D VAE decoder ------------------------------------------------------------------------------ */

/* b200_conv3d_cl with a separate extent (in_T, in_H, in_W) for the input view: the producer has already written the
 * convolution's border into the tensor (replicate padding of CausalConv3d.forward,
 * lightx2v/models/video_encoders/hf/autoencoder_kl_causal_3d/unet_causal_3d_blocks.py:88-91), so `taps` are non-negative
 * offsets into the padded view and the output extent (T, H, W) is smaller than the input's.  cin multiple of 64 with
 * cout multiple of 128 / 256 selects the 128-byte-swizzle tiles used by the 128 / 256 / 512-channel stages. */
int b200_conv3d_cl_padded(const void* in, int64_t in_st, int64_t in_sh, int64_t in_sw, int in_T, int in_H, int in_W, const void* wt,
                          const void* bias, void* out, int64_t out_st, int64_t out_sh, int64_t out_sw, const void* residual,
                          int64_t res_st, int64_t res_sh, int64_t res_sw, int T, int H, int W, int cin, int cout, int ntaps,
                          const int32_t* taps, int clamp_out, b200_stream_t stream);

/* Size (in doubles) of the device workspace b200_gn_stats_cl needs: 64 result entries + room for the per-block partial sums. */
int64_t b200_gn_stats_workspace_doubles(void);

/* sums[0:32] = per-group sum, sums[32:64] = per-group sum of squares (fp64, device; `sums` is a workspace of
 * b200_gn_stats_workspace_doubles() entries, deterministic fixed-order reduction, no atomics) of a channels-last bf16
 * tensor [voxels, C], 32 groups of C/32 consecutive channels, C in {64,128,256,512}.  First half of torch.nn.GroupNorm as used by
 * ResnetBlockCausal3D.norm1/norm2, Attention.group_norm and DecoderCausal3D.conv_norm_out (unet_causal_3d_blocks.py:313,332;
 * vae.py:209). */
int b200_gn_stats_cl(const void* x, int64_t voxels, int C, double* sums, b200_stream_t stream);

/* y[T+pt, H+2ph, W+2pw, C] = replicate_pad( [silu]( (x - mean_g) * rsqrt(var_g + eps) * gamma + beta ) ), statistics from
 * b200_gn_stats_cl; sums == NULL -> plain replicate-pad copy.  GroupNorm -> SiLU -> F.pad(mode="replicate") of
 * ResnetBlockCausal3D.forward + CausalConv3d.forward (unet_causal_3d_blocks.py:364-412, 88-91). */
int b200_gn_apply_pad_cl(const void* x, void* y, const double* sums, const float* gamma, const float* beta, float eps, int T, int H,
                         int W, int C, int pt, int ph, int pw, int apply_silu, b200_stream_t stream);

/* y[v, :] = [silu] ( x[v, :] / max(||x[v, :]||_2, 1e-12) * sqrt(C) * gamma ) over `voxels` channels-last rows of C in {96,192,384}.
 * Replaces RMS_norm.forward + nn.SiLU (vae.py:47-59, 192-195, 430-433). */
int b200_rms_silu_cl(const void* x, void* y, const float* gamma, int64_t voxels, int C, int apply_silu, b200_stream_t stream);

/* z [CZ,T,H,W] fp32 -> channels-last bf16 [T,H,W,CP] of (z / inv_std + mean), channels >= CZ zero (WanVAE_.decode, vae.py:716-719). */
int b200_latent_to_cl(const float* z, void* out, const float* mean, const float* inv_std, int64_t voxels, int CZ, int CP,
                      b200_stream_t stream);

/* channels-last bf16 [voxels, CP] (3 valid channels) -> fp32 [3, voxels]  (the `.float()` video tensor, vae.py:951); the three channel planes
 * are out_channel_stride elements apart (0 = voxels), so a chunk of frames can be written into its slot of the full [3, T, H, W] video. */
int b200_cl_to_video(const void* in, float* out, int64_t voxels, int CP, int64_t out_channel_stride, b200_stream_t stream);

/* Hardware probe (debug; no reference counterpart, not on any product path): D[128, 64] fp32 = A[r0 : r0 + 128, :] * B^T for A [136, k] and
 * B [64, k] bf16 (k = 64: 128-byte rows / SWIZZLE_128B, k = 32: 64-byte rows / SWIZZLE_64B), with the UMMA descriptor of A starting r0 rows into
 * the TMA-written shared-memory tile; mode 1 also sets the descriptor's base_offset field to r0 % 8.  tools/probe_rowshift.py reports which
 * addressing the tensor core honours (the precondition for taking a convolution's dw taps as row-shifted views of one halo tile). */
int b200_debug_umma_rowshift(const void* A, const void* B, float* D, int k_elems, int r0, int mode, b200_stream_t stream);

/* Hardware probe (debug; no reference counterpart, not on any product path): clocks per SS-mode tcgen05.mma (M = 128, K = 16, bf16, operands
 * resident in shared memory, no TMA, no epilogue) at width n, rotating over n_acc TMEM accumulators, from `issuers` (1 | 2) warps, cycling
 * through a_tiles (1..4) A tiles, with `writers` (0..4) warps streaming shared-memory stores beside it.  out: uint64 [3 * grid] -
 * [2 * cta + issuer] clocks for iters x 4 MMAs, [2 * grid + cta] 512-byte store instructions retired by the writers.  tools/probe_umma_rate.py
 * prints the table DESIGN.md section 4 quotes as the per-shape floor of the conv / GEMM tiles. */
int b200_debug_umma_rate(int n, int iters, int n_acc, int issuers, int writers, int a_tiles, int grid, unsigned long long* out, b200_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* B200_DIT_H_ */
