// One Wan DiT block as a single native call: the launch schedule of host/wan_infer.py:WanTransformerInfer.infer_block (bf16 linears,
// single GPU) moved below the C ABI, so that a binding makes ONE call per block instead of ~13.
//
// Replaces WanTransformerInfer.infer_block = infer_modulation -> infer_self_attn -> infer_cross_attn -> infer_ffn -> post_process
// (lightx2v/models/networks/wan/infer/transformer_infer.py:289-508).  The caller passes the six modulation vectors
// (modulation + embed0, :308-319), the RoPE table and the step-invariant text K/V (norm_k(Linear(context)), Linear(context), :418-420),
// which the reference recomputes in every block of every step.  No allocation: all intermediates live in the workspace.
#include "../../include/b200_dit.h"

#include "host_util.cuh"
#include "ptx.cuh"

namespace b200 {
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

// a[i] += b[i] in bf16 (the reference sums the text and image cross-attention outputs in bf16, transformer_infer.py:454)
__global__ void __launch_bounds__(256) add_bf16_kernel(__nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ b, long long n8) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    uint4 va = reinterpret_cast<uint4*>(a)[i];
    const uint4 vb = reinterpret_cast<const uint4*>(b)[i];
    uint32_t* pa = reinterpret_cast<uint32_t*>(&va);
    const uint32_t* pb = reinterpret_cast<const uint32_t*>(&vb);
#pragma unroll
    for (int e = 0; e < 4; ++e) pa[e] = pack_bf16(bf16_lo(pa[e]) + bf16_lo(pb[e]), bf16_hi(pa[e]) + bf16_hi(pb[e]));
    reinterpret_cast<uint4*>(a)[i] = va;
  }
}

static long long big_cols(int D, int F) { return (3LL * D > F) ? 3LL * D : F; }

long long wan_block_workspace_bytes(long long S, int D, int F) {
  // n [S, D] (LN output, later the attention output) + one region of max(3D, F) columns: fused qkv, then cross-q (+ i2v scratch), then the FFN hidden
  return S * (D + big_cols(D, F)) * 2;
}

#define B200_TRY(expr)              \
  do {                              \
    int rc_ = (expr);               \
    if (rc_ != B200_OK) return rc_; \
  } while (0)

int wan_block_fwd(const b200_wan_block_weights* w, const b200_wan_block_args* a, void* workspace, long long workspace_bytes,
                  cudaStream_t stream) {
  B200_CHECK_ARG(w && a && workspace, "b200_wan_block_fwd: null pointer");
  const long long S = a->S;
  const int D = a->D, H = a->H, F = a->F;
  B200_CHECK_ARG(S > 0 && D > 0 && H > 0 && F > 0 && D == H * 128, "b200_wan_block_fwd: need D == 128 * H (got D=%d H=%d)", D, H);
  B200_CHECK_ARG(workspace_bytes >= wan_block_workspace_bytes(S, D, F), "b200_wan_block_fwd: workspace too small (%lld < %lld bytes)",
                 workspace_bytes, wan_block_workspace_bytes(S, D, F));
  B200_CHECK_ARG(a->x && a->shift_msa && a->scale_msa && a->gate_msa && a->c_shift_msa && a->c_scale_msa && a->c_gate_msa && a->ctx_k && a->ctx_v &&
                     a->ctx_len > 0,
                 "b200_wan_block_fwd: missing activation / modulation / text K,V pointer");
  B200_CHECK_ARG((a->img_k == nullptr) == (a->img_v == nullptr), "b200_wan_block_fwd: img_k and img_v go together");
  auto* x = reinterpret_cast<__nv_bfloat16*>(a->x);
  auto* n = reinterpret_cast<__nv_bfloat16*>(workspace);
  auto* big = n + S * D;
  const float eps = a->eps, sm = 0.08838834764831845f;   // 1 / sqrt(128)
  const long long D3 = 3LL * D;

  // ---- self-attention (transformer_infer.py:321-396) + gated residual (:402)
  B200_TRY(ln_modulate(x, D, n, D, nullptr, nullptr, a->scale_msa, a->shift_msa, S, D, eps, stream));
  B200_TRY(gemm_bf16(n, D, w->wqkv, D, big, D3, w->bqkv, nullptr, S, D3, D, 0 /*bias*/, 0, 0, stream));
  B200_TRY(rms_rope(big, D3, w->norm_q, big + D, D3, w->norm_k, S, D, eps, a->cos_sin, a->rope_rows, 0, stream));
  B200_TRY(fmha_fwd_d128(big, D3, big + D, D3, big + 2 * D, D3, n, D, S, S, H, sm, stream));
  B200_TRY(gemm_bf16(n, D, w->wo, D, x, D, w->bo, a->gate_msa, S, D, D, 2 /*gate residual*/, 0, 0, stream));

  // ---- cross-attention (:398-465): x += o(attn(norm_q(q(norm3(x))), K_text, V_text) [+ attn(., K_img, V_img)])
  __nv_bfloat16* cq = big;
  B200_TRY(ln_modulate(x, D, n, D, w->norm3_w, w->norm3_b, nullptr, nullptr, S, D, eps, stream));
  B200_TRY(gemm_bf16(n, D, w->wcq, D, cq, D, w->bcq, nullptr, S, D, D, 0, 0, 0, stream));
  B200_TRY(rms_rope(cq, D, w->cnorm_q, nullptr, 0, nullptr, S, D, eps, nullptr, 0, 0, stream));
  B200_TRY(fmha_fwd_d128(cq, D, a->ctx_k, D, a->ctx_v, D, n, D, S, a->ctx_len, H, sm, stream));
  if (a->img_k != nullptr) {
    __nv_bfloat16* tmp = big + S * D;
    B200_TRY(fmha_fwd_d128(cq, D, a->img_k, D, a->img_v, D, tmp, D, S, a->img_len, H, sm, stream));
    const long long n8 = S * D / 8;
    const long long want = (n8 + 255) / 256;
    const int blocks = (int)(want < num_sms() * 8 ? want : num_sms() * 8);
    add_bf16_kernel<<<blocks, 256, 0, stream>>>(n, tmp, n8); note_launch();
    B200_CHECK_CUDA(cudaGetLastError());
  }
  B200_TRY(gemm_bf16(n, D, w->wco, D, x, D, w->bco, nullptr, S, D, D, 3 /*residual*/, 0, 0, stream));

  // ---- FFN (:467-497) + gated residual (:499-508)
  __nv_bfloat16* hidden = big;
  B200_TRY(ln_modulate(x, D, n, D, nullptr, nullptr, a->c_scale_msa, a->c_shift_msa, S, D, eps, stream));
  B200_TRY(gemm_bf16(n, D, w->w0, D, hidden, F, w->b0, nullptr, S, F, D, 1 /*bias + gelu*/, 0, 0, stream));
  B200_TRY(gemm_bf16(hidden, F, w->w2, F, x, D, w->b2, a->c_gate_msa, S, D, F, 2, 0, 0, stream));
  return B200_OK;
}

}  // namespace b200

extern "C" {
int64_t b200_wan_block_workspace_bytes(int64_t S, int D, int F) { return b200::wan_block_workspace_bytes(S, D, F); }

int b200_wan_block_fwd(const b200_wan_block_weights* w, const b200_wan_block_args* a, void* workspace, int64_t workspace_bytes,
                       b200_stream_t stream) {
  return b200::wan_block_fwd(w, a, workspace, workspace_bytes, reinterpret_cast<cudaStream_t>(stream));
}
}
