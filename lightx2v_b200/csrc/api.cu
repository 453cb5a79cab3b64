// extern "C" surface of libb200dit.so — see include/b200_dit.h for the contract of every entry point.
#include "../../include/b200_dit.h"

#include "host_util.cuh"

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
int gemm_fp8(const void* A, long long lda, const void* B, long long ldb, void* C, long long ldc, const float* scale_a,
             const float* scale_b, const void* bias, const void* gate, long long M, long long N, long long K,
             int epilogue, int block_n, int max_ctas, cudaStream_t stream);
int quant_fp8_per_token(const void* x, long long ldx, void* q8, long long ldq, float* q_scale, long long rows, int D,
                        cudaStream_t stream);
int ln_modulate_fp8(const void* x, long long ldx, void* q8, long long ldq, float* q_scale, const void* ln_w,
                    const void* ln_b, const void* scale, const void* shift, long long rows, int D, float eps,
                    cudaStream_t stream);
}  // namespace b200

extern "C" {

const char* b200_last_error(void) { return b200::last_error(); }
int b200_version(void) { return 100; }
int b200_num_sms(void) { return b200::num_sms(); }

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

}  // extern "C"
