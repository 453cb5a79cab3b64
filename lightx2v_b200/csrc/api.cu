// extern "C" surface of libb200dit.so — see include/b200_dit.h for the contract of every entry point.
#include "../../include/b200_dit.h"

#include "host_util.cuh"

#include <string.h>

namespace b200 {
const char* last_error();
int gemm_bf16(const void* A, long long lda, const void* B, long long ldb, void* C, long long ldc, const void* bias,
              const void* gate, long long M, long long N, long long K, int epilogue, int block_n, int max_ctas,
              cudaStream_t stream);
int ln_modulate(const void* x, long long ldx, void* y, long long ldy, const void* ln_w, const void* ln_b,
                const void* scale, const void* shift, long long rows, int D, float eps, cudaStream_t stream);
int rms_rope(void* x0, long long ld0, const void* w0, void* x1, long long ld1, const void* w1, long long rows, int D,
             float eps, const void* cos_sin, long long rope_rows, long long pos_offset, cudaStream_t stream);
int fmha_fwd_d128(const void* q, long long q_stride_s, const void* k, long long k_stride_s, const void* v,
                  long long v_stride_s, void* out, long long o_stride_s, long long sq, long long sk, int heads,
                  float softmax_scale, cudaStream_t stream);
int fmha_fwd_d64(const void* q, long long q_stride_s, const void* k, long long k_stride_s, const void* v, long long v_stride_s, void* out,
                 long long o_stride_s, long long sq, long long sk, int heads, float softmax_scale, cudaStream_t stream);
int prof_fmha_begin(int capacity);
int prof_fmha_end(float* ms, long long* meta, int capacity);
int gemm_fp8(const void* A, long long lda, const void* B, long long ldb, void* C, long long ldc, const float* scale_a,
             const float* scale_b, const void* bias, const void* gate, long long M, long long N, long long K,
             int epilogue, int block_n, int max_ctas, cudaStream_t stream);
int quant_fp8_per_token(const void* x, long long ldx, void* q8, long long ldq, float* q_scale, long long rows, int D,
                        cudaStream_t stream);
int ln_modulate_fp8(const void* x, long long ldx, void* q8, long long ldq, float* q_scale, const void* ln_w,
                    const void* ln_b, const void* scale, const void* shift, long long rows, int D, float eps,
                    cudaStream_t stream);
int conv3d_cl(const void* in, long long in_st, long long in_sh, long long in_sw, int in_T, int in_H, int in_W, const void* wt,
              const void* bias, void* out, long long out_st, long long out_sh, long long out_sw, const void* residual,
              long long res_st, long long res_sh, long long res_sw, int T, int H, int W, int cin, int cout, int ntaps,
              const int* taps, int clamp_out, cudaStream_t stream);
int gemm_nvfp4(const void* A, long long lda, const void* B, long long ldb, const void* sfa, const void* sfb, const float* alpha,
               void* C, long long ldc, const void* bias, const void* gate, long long M, long long N, long long K, int epilogue,
               int block_n, int max_ctas, cudaStream_t stream);
int quant_nvfp4(const void* x, long long ldx, long long rows, int K, const float* global_scale, void* q, long long ldq, void* sf,
                cudaStream_t stream);
int nvfp4_act_scale(const void* x, long long ldx, long long rows, int K, const float* weight_global_scale, float* global_scale,
                    float* alpha, void* scratch, cudaStream_t stream);
int gn_stats_cl(const void* x, long long voxels, int C, double* sums, cudaStream_t stream);
long long gn_stats_workspace_doubles();
int gn_apply_pad_cl(const void* x, void* y, const double* sums, const float* gamma, const float* beta, float eps, int T, int H, int W,
                    int C, int pt, int ph, int pw, int apply_silu, cudaStream_t stream);
int rms_silu_cl(const void* x, void* y, const float* gamma, long long voxels, int C, int apply_silu, cudaStream_t stream);
int latent_to_cl(const float* z, void* out, const float* mean, const float* inv_std, long long voxels, int CZ, int CP,
                 cudaStream_t stream);
int cl_to_video(const void* in, float* out, long long voxels, int CP, long long out_channel_stride, cudaStream_t stream);
int rms_rope_scatter(const void* qkv, long long ld, const void* wq, const void* wk, long long rows, int D, float eps,
                     const void* cos_sin, long long rope_rows, void* const* peers, int world, int rank,
                     long long rows_per_rank, cudaStream_t stream);
int fmha_fwd_d128_scatter(const void* q, long long q_stride_s, const void* k, long long k_stride_s, const void* v,
                          long long v_stride_s, void* const* peers, int world, long long rows_per_rank, long long peer_stride_s,
                          int head_offset, long long sq, long long sk, int heads, float softmax_scale, cudaStream_t stream);
int rms_rope_heads(void* x0, long long ld0, const void* w0, void* x1, long long ld1, const void* w1, long long rows,
                   int H, float eps, const void* cos_sin, long long rope_rows, cudaStream_t stream);
int ln_rope_heads64(void* x0, long long ld0, const void* w0, const void* b0, void* x1, long long ld1, const void* w1, const void* b1, long long rows,
                    int H, float eps, const void* cos_sin, long long rope_start, cudaStream_t stream);
int debug_umma_rowshift(const void* A, const void* B, float* D, int k_elems, int r0, int mode, cudaStream_t stream);
int debug_umma_rate(int n, int iters, int n_acc, int issuers, int writers, int a_tiles, int grid, unsigned long long* out, cudaStream_t stream);
}  // namespace b200

extern "C" {

const char* b200_last_error(void) { return b200::last_error(); }
int b200_version(void) { return 100; }
int b200_num_sms(void) { return b200::num_sms(); }
int64_t b200_launch_count(void) { return b200::launch_count(); }
int b200_set_option(const char* name, int value) { return b200::set_option(name, value); }
int b200_get_option(const char* name) {
  static const char* const names[b200::OPT_COUNT] = {"conv_halo", "halo_base_offset", "conv_narrow"};
  for (int i = 0; name && i < b200::OPT_COUNT; ++i)
    if (strcmp(name, names[i]) == 0) return b200::get_option(i);
  return -1;
}
int b200_prof_fmha_begin(int capacity) { return b200::prof_fmha_begin(capacity); }
int b200_prof_fmha_end(float* ms, int64_t* meta, int capacity) {
  static_assert(sizeof(long long) == sizeof(int64_t), "int64_t is long long on this ABI");
  return b200::prof_fmha_end(ms, reinterpret_cast<long long*>(meta), capacity);
}

int b200_gemm_bf16(const void* A, int64_t lda, const void* B, int64_t ldb, void* C, int64_t ldc, const void* bias,
                   const void* gate, int64_t M, int64_t N, int64_t K, int epilogue, int block_n, int max_ctas,
                   b200_stream_t stream) {
  return b200::gemm_bf16(A, lda, B, ldb, C, ldc, bias, gate, M, N, K, epilogue, block_n, max_ctas,
                         reinterpret_cast<cudaStream_t>(stream));
}

int b200_ln_modulate(const void* x, int64_t ldx, void* y, int64_t ldy, const void* ln_w, const void* ln_b,
                     const void* scale, const void* shift, int64_t rows, int D, float eps, b200_stream_t stream) {
  return b200::ln_modulate(x, ldx, y, ldy, ln_w, ln_b, scale, shift, rows, D, eps,
                           reinterpret_cast<cudaStream_t>(stream));
}

int b200_rms_rope(void* x0, int64_t ld0, const void* w0, void* x1, int64_t ld1, const void* w1, int64_t rows, int D,
                  float eps, const void* cos_sin, int64_t rope_rows, int64_t pos_offset, b200_stream_t stream) {
  return b200::rms_rope(x0, ld0, w0, x1, ld1, w1, rows, D, eps, cos_sin, rope_rows, pos_offset,
                        reinterpret_cast<cudaStream_t>(stream));
}

int b200_fmha_fwd_d128(const void* q, int64_t q_stride_s, const void* k, int64_t k_stride_s, const void* v,
                       int64_t v_stride_s, void* out, int64_t o_stride_s, int64_t sq, int64_t sk, int heads,
                       float softmax_scale, b200_stream_t stream) {
  return b200::fmha_fwd_d128(q, q_stride_s, k, k_stride_s, v, v_stride_s, out, o_stride_s, sq, sk, heads,
                             softmax_scale, reinterpret_cast<cudaStream_t>(stream));
}

int b200_fmha_fwd_d64(const void* q, int64_t q_stride_s, const void* k, int64_t k_stride_s, const void* v, int64_t v_stride_s,
                      void* out, int64_t o_stride_s, int64_t sq, int64_t sk, int heads, float softmax_scale, b200_stream_t stream) {
  return b200::fmha_fwd_d64(q, q_stride_s, k, k_stride_s, v, v_stride_s, out, o_stride_s, sq, sk, heads, softmax_scale,
                            reinterpret_cast<cudaStream_t>(stream));
}

int b200_gemm_fp8(const void* A, int64_t lda, const void* B, int64_t ldb, void* C, int64_t ldc, const float* a_scale,
                  const float* b_scale, const void* bias, const void* gate, int64_t M, int64_t N, int64_t K, int epilogue,
                  int block_n, int max_ctas, b200_stream_t stream) {
  return b200::gemm_fp8(A, lda, B, ldb, C, ldc, a_scale, b_scale, bias, gate, M, N, K, epilogue, block_n, max_ctas,
                        reinterpret_cast<cudaStream_t>(stream));
}

int b200_quant_fp8_per_token(const void* x, int64_t ldx, void* q8, int64_t ldq, float* q_scale, int64_t rows, int D,
                             b200_stream_t stream) {
  return b200::quant_fp8_per_token(x, ldx, q8, ldq, q_scale, rows, D, reinterpret_cast<cudaStream_t>(stream));
}

int b200_ln_modulate_fp8(const void* x, int64_t ldx, void* q8, int64_t ldq, float* q_scale, const void* ln_w,
                         const void* ln_b, const void* scale, const void* shift, int64_t rows, int D, float eps,
                         b200_stream_t stream) {
  return b200::ln_modulate_fp8(x, ldx, q8, ldq, q_scale, ln_w, ln_b, scale, shift, rows, D, eps,
                               reinterpret_cast<cudaStream_t>(stream));
}

int b200_conv3d_cl(const void* in, int64_t in_st, int64_t in_sh, int64_t in_sw, const void* wt, const void* bias, void* out,
                   int64_t out_st, int64_t out_sh, int64_t out_sw, const void* residual, int64_t res_st, int64_t res_sh,
                   int64_t res_sw, int T, int H, int W, int cin, int cout, int ntaps, const int32_t* taps, int clamp_out,
                   b200_stream_t stream) {
  return b200::conv3d_cl(in, in_st, in_sh, in_sw, T, H, W, wt, bias, out, out_st, out_sh, out_sw, residual, res_st, res_sh, res_sw,
                         T, H, W, cin, cout, ntaps, taps, clamp_out, reinterpret_cast<cudaStream_t>(stream));
}

int b200_conv3d_cl_padded(const void* in, int64_t in_st, int64_t in_sh, int64_t in_sw, int in_T, int in_H, int in_W, const void* wt,
                          const void* bias, void* out, int64_t out_st, int64_t out_sh, int64_t out_sw, const void* residual,
                          int64_t res_st, int64_t res_sh, int64_t res_sw, int T, int H, int W, int cin, int cout, int ntaps,
                          const int32_t* taps, int clamp_out, b200_stream_t stream) {
  return b200::conv3d_cl(in, in_st, in_sh, in_sw, in_T, in_H, in_W, wt, bias, out, out_st, out_sh, out_sw, residual, res_st, res_sh,
                         res_sw, T, H, W, cin, cout, ntaps, taps, clamp_out, reinterpret_cast<cudaStream_t>(stream));
}

int b200_gemm_nvfp4(const void* A, int64_t lda, const void* B, int64_t ldb, const void* sfa, const void* sfb, const float* alpha, void* C,
                    int64_t ldc, const void* bias, const void* gate, int64_t M, int64_t N, int64_t K, int epilogue, int block_n,
                    int max_ctas, b200_stream_t stream) {
  return b200::gemm_nvfp4(A, lda, B, ldb, sfa, sfb, alpha, C, ldc, bias, gate, M, N, K, epilogue, block_n, max_ctas,
                          reinterpret_cast<cudaStream_t>(stream));
}

int b200_quant_nvfp4(const void* x, int64_t ldx, int64_t rows, int K, const float* global_scale, void* q, int64_t ldq, void* sf,
                     b200_stream_t stream) {
  return b200::quant_nvfp4(x, ldx, rows, K, global_scale, q, ldq, sf, reinterpret_cast<cudaStream_t>(stream));
}

int b200_nvfp4_act_scale(const void* x, int64_t ldx, int64_t rows, int K, const float* weight_global_scale, float* global_scale,
                         float* alpha, void* scratch, b200_stream_t stream) {
  return b200::nvfp4_act_scale(x, ldx, rows, K, weight_global_scale, global_scale, alpha, scratch, reinterpret_cast<cudaStream_t>(stream));
}

int64_t b200_gn_stats_workspace_doubles(void) { return b200::gn_stats_workspace_doubles(); }

int b200_gn_stats_cl(const void* x, int64_t voxels, int C, double* sums, b200_stream_t stream) {
  return b200::gn_stats_cl(x, voxels, C, sums, reinterpret_cast<cudaStream_t>(stream));
}

int b200_gn_apply_pad_cl(const void* x, void* y, const double* sums, const float* gamma, const float* beta, float eps, int T, int H,
                         int W, int C, int pt, int ph, int pw, int apply_silu, b200_stream_t stream) {
  return b200::gn_apply_pad_cl(x, y, sums, gamma, beta, eps, T, H, W, C, pt, ph, pw, apply_silu, reinterpret_cast<cudaStream_t>(stream));
}

int b200_rms_silu_cl(const void* x, void* y, const float* gamma, int64_t voxels, int C, int apply_silu, b200_stream_t stream) {
  return b200::rms_silu_cl(x, y, gamma, voxels, C, apply_silu, reinterpret_cast<cudaStream_t>(stream));
}

int b200_latent_to_cl(const float* z, void* out, const float* mean, const float* inv_std, int64_t voxels, int CZ, int CP,
                      b200_stream_t stream) {
  return b200::latent_to_cl(z, out, mean, inv_std, voxels, CZ, CP, reinterpret_cast<cudaStream_t>(stream));
}

int b200_cl_to_video(const void* in, float* out, int64_t voxels, int CP, int64_t out_channel_stride, b200_stream_t stream) {
  return b200::cl_to_video(in, out, voxels, CP, out_channel_stride, reinterpret_cast<cudaStream_t>(stream));
}

int b200_rms_rope_scatter(const void* qkv, int64_t ld, const void* wq, const void* wk, int64_t rows, int D, float eps,
                          const void* cos_sin, int64_t rope_rows, void* const* peers, int world, int rank,
                          int64_t rows_per_rank, b200_stream_t stream) {
  return b200::rms_rope_scatter(qkv, ld, wq, wk, rows, D, eps, cos_sin, rope_rows, peers, world, rank, rows_per_rank,
                                reinterpret_cast<cudaStream_t>(stream));
}

int b200_fmha_fwd_d128_scatter(const void* q, int64_t q_stride_s, const void* k, int64_t k_stride_s, const void* v,
                               int64_t v_stride_s, void* const* peers, int world, int64_t rows_per_rank, int64_t peer_stride_s,
                               int head_offset, int64_t sq, int64_t sk, int heads, float softmax_scale, b200_stream_t stream) {
  return b200::fmha_fwd_d128_scatter(q, q_stride_s, k, k_stride_s, v, v_stride_s, peers, world, rows_per_rank, peer_stride_s,
                                     head_offset, sq, sk, heads, softmax_scale, reinterpret_cast<cudaStream_t>(stream));
}

int b200_rms_rope_heads(void* x0, int64_t ld0, const void* w0, void* x1, int64_t ld1, const void* w1, int64_t rows, int H,
                        float eps, const void* cos_sin, int64_t rope_rows, b200_stream_t stream) {
  return b200::rms_rope_heads(x0, ld0, w0, x1, ld1, w1, rows, H, eps, cos_sin, rope_rows, reinterpret_cast<cudaStream_t>(stream));
}

int b200_ln_rope_heads64(void* x0, int64_t ld0, const void* w0, const void* b0, void* x1, int64_t ld1, const void* w1, const void* b1,
                         int64_t rows, int H, float eps, const void* cos_sin, int64_t rope_start, b200_stream_t stream) {
  return b200::ln_rope_heads64(x0, ld0, w0, b0, x1, ld1, w1, b1, rows, H, eps, cos_sin, rope_start, reinterpret_cast<cudaStream_t>(stream));
}

int b200_debug_umma_rowshift(const void* A, const void* B, float* D, int k_elems, int r0, int mode, b200_stream_t stream) {
  return b200::debug_umma_rowshift(A, B, D, k_elems, r0, mode, reinterpret_cast<cudaStream_t>(stream));
}

int b200_debug_umma_rate(int n, int iters, int n_acc, int issuers, int writers, int a_tiles, int grid, unsigned long long* out, b200_stream_t stream) {
  return b200::debug_umma_rate(n, iters, n_acc, issuers, writers, a_tiles, grid, out, reinterpret_cast<cudaStream_t>(stream));
}

}  // extern "C"
