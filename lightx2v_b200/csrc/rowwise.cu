// HBM-bound row-wise kernels of the DiT block: LayerNorm (+AdaLN modulation), full-width q/k RMSNorm (+3-axis RoPE).
// One CTA of 128 threads per token row; the row lives in registers between the statistics pass and the
// normalise/rotate pass, so every element is read once and written once (algorithmic bytes = 4 B per element).
//
// Rounding points mirror the reference's bf16 torch path exactly:
//   LayerNorm  : F.layer_norm(bf16) -> bf16, then .mul_(1+scale) -> bf16, .add_(shift) -> bf16
//                (lightx2v/models/networks/wan/infer/transformer_infer.py:326-334,478-484; layer_norm_weight.py:100-111)
//   RMSNorm    : x * rsqrt(x.pow(2).mean(-1) + eps) * w, every intermediate a bf16 tensor
//                (lightx2v/common/ops/norm/rms_norm_weight.py:111-113 — the fallback taken when sgl_kernel is absent)
//   RoPE       : complex multiply of adjacent pairs (2i, 2i+1) by freqs[pos, i], one rounding to bf16
//                (lightx2v/models/networks/wan/infer/utils.py:107-115); the reference does it in complex128,
//                here fp32 on a cos/sin table prepared in fp64 on the host.
#include "host_util.cuh"
#include "ptx.cuh"

#include <cuda_fp8.h>

namespace b200 {

constexpr int ROW_THREADS = 128;
constexpr int ROW_MAX_VEC = 8;  // up to 8 x 8 elements per thread -> D <= 8192

__device__ __forceinline__ float block_sum_128(float v, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) red[w] = v;
  __syncthreads();
  float t = red[0] + red[1] + red[2] + red[3];
  __syncthreads();
  return t;
}

__device__ __forceinline__ float block_max(float v, float* red, int nwarps) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  const int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) red[w] = v;
  __syncthreads();
  float t = red[0];
  for (int i = 1; i < nwarps; ++i) t = fmaxf(t, red[i]);
  __syncthreads();
  return t;
}

// 8 floats -> 8 e4m3 (round-to-nearest-even, saturate-to-finite), packed in 8 bytes
__device__ __forceinline__ uint2 pack8_e4m3(const float* f) {
  uint2 o;
  o.x = (uint32_t)__nv_cvt_float2_to_fp8x2(make_float2(f[0], f[1]), __NV_SATFINITE, __NV_E4M3) |
        ((uint32_t)__nv_cvt_float2_to_fp8x2(make_float2(f[2], f[3]), __NV_SATFINITE, __NV_E4M3) << 16);
  o.y = (uint32_t)__nv_cvt_float2_to_fp8x2(make_float2(f[4], f[5]), __NV_SATFINITE, __NV_E4M3) |
        ((uint32_t)__nv_cvt_float2_to_fp8x2(make_float2(f[6], f[7]), __NV_SATFINITE, __NV_E4M3) << 16);
  return o;
}

// dynamic per-token scale of vLLM's scaled_fp8_quant(use_per_token_if_dynamic=True) (mm_weight.py:236-238):
//   scale = max(absmax / 448, 1 / (448 * 512));  q = e4m3_rn_satfinite(x / scale)
__device__ __forceinline__ float fp8_token_scale(float absmax) { return fmaxf(absmax / 448.0f, 1.0f / (448.0f * 512.0f)); }

__device__ __forceinline__ void unpack8(const uint4& v, float* f) {
  f[0] = bf16_lo(v.x); f[1] = bf16_hi(v.x);
  f[2] = bf16_lo(v.y); f[3] = bf16_hi(v.y);
  f[4] = bf16_lo(v.z); f[5] = bf16_hi(v.z);
  f[6] = bf16_lo(v.w); f[7] = bf16_hi(v.w);
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  uint4 o;
  o.x = pack_bf16(f[0], f[1]);
  o.y = pack_bf16(f[2], f[3]);
  o.z = pack_bf16(f[4], f[5]);
  o.w = pack_bf16(f[6], f[7]);
  return o;
}

// ---------------------------------------------------------------------------------------------------------
// LayerNorm (eps, optional affine weight/bias) followed by optional AdaLN modulation  y = bf16(bf16(n * w1) + shift)
// with w1 = bf16(1 + scale).  `scale` / `shift` are [D] bf16 vectors (one per block and pass).
// ---------------------------------------------------------------------------------------------------------
template <bool kAffine, bool kModulate, bool kFp8Out>
__global__ void __launch_bounds__(ROW_THREADS)
ln_modulate_kernel(const __nv_bfloat16* __restrict__ x, long long ldx, __nv_bfloat16* __restrict__ y, long long ldy,
                   const __nv_bfloat16* __restrict__ ln_w, const __nv_bfloat16* __restrict__ ln_b,
                   const __nv_bfloat16* __restrict__ scale, const __nv_bfloat16* __restrict__ shift, int D,
                   float eps, uint8_t* __restrict__ q8, long long ldq, float* __restrict__ q_scale) {
  __shared__ float red[4];
  const long long row = blockIdx.x;
  const int nvec = D >> 3;
  const uint4* xr = reinterpret_cast<const uint4*>(x + row * ldx);
  uint4 raw[ROW_MAX_VEC];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < ROW_MAX_VEC; ++i) {
    const int v = threadIdx.x + i * ROW_THREADS;
    if (v < nvec) {
      raw[i] = xr[v];
      float f[8];
      unpack8(raw[i], f);
#pragma unroll
      for (int e = 0; e < 8; ++e) sum += f[e];
    }
  }
  const float mean = block_sum_128(sum, red) / (float)D;
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < ROW_MAX_VEC; ++i) {
    const int v = threadIdx.x + i * ROW_THREADS;
    if (v < nvec) {
      float f[8];
      unpack8(raw[i], f);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float d = f[e] - mean;
        sq += d * d;
      }
    }
  }
  const float rstd = rsqrtf(block_sum_128(sq, red) / (float)D + eps);
  uint4* yr = reinterpret_cast<uint4*>(y + row * ldy);
  float amax = 0.f;
#pragma unroll
  for (int i = 0; i < ROW_MAX_VEC; ++i) {
    const int v = threadIdx.x + i * ROW_THREADS;
    if (v < nvec) {
      float f[8];
      unpack8(raw[i], f);
#pragma unroll
      for (int e = 0; e < 8; ++e) f[e] = (f[e] - mean) * rstd;
      if constexpr (kAffine) {
        float w[8], b[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(ln_w) + v), w);
        unpack8(__ldg(reinterpret_cast<const uint4*>(ln_b) + v), b);
#pragma unroll
        for (int e = 0; e < 8; ++e) f[e] = fmaf(f[e], w[e], b[e]);
      }
      if constexpr (kModulate) {
        float sc[8], sh[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(scale) + v), sc);
        unpack8(__ldg(reinterpret_cast<const uint4*>(shift) + v), sh);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float w1 = bf16_round(1.0f + sc[e]);
          f[e] = bf16_round(bf16_round(f[e]) * w1) + sh[e];
        }
      }
      if constexpr (kFp8Out) {
        raw[i] = pack8(f);   // the bf16 tensor the reference would hand to the quantiser
        unpack8(raw[i], f);
#pragma unroll
        for (int e = 0; e < 8; ++e) amax = fmaxf(amax, fabsf(f[e]));
      } else {
        yr[v] = pack8(f);
      }
    }
  }
  if constexpr (kFp8Out) {
    const float sc = fp8_token_scale(block_max(amax, red, ROW_THREADS / 32));
    if (threadIdx.x == 0) q_scale[row] = sc;
    uint2* qr = reinterpret_cast<uint2*>(q8 + row * ldq);
#pragma unroll
    for (int i = 0; i < ROW_MAX_VEC; ++i) {
      const int v = threadIdx.x + i * ROW_THREADS;
      if (v < nvec) {
        float f[8];
        unpack8(raw[i], f);
#pragma unroll
        for (int e = 0; e < 8; ++e) f[e] = f[e] / sc;
        qr[v] = pack8_e4m3(f);
      }
    }
  }
}

static int ln_modulate_impl(const void* x, long long ldx, void* y, long long ldy, const void* ln_w, const void* ln_b,
                            const void* scale, const void* shift, long long rows, int D, float eps, void* q8, long long ldq,
                            float* q_scale, cudaStream_t stream) {
  const bool fp8 = q8 != nullptr;
  B200_CHECK_ARG(x && (y || fp8), "b200_ln_modulate: null pointer");
  B200_CHECK_ARG(rows > 0 && D > 0 && D % 8 == 0 && D <= ROW_THREADS * ROW_MAX_VEC * 8,
                 "b200_ln_modulate: D=%d must be a multiple of 8 and <= %d", D, ROW_THREADS * ROW_MAX_VEC * 8);
  B200_CHECK_ARG(ldx % 8 == 0 && ldx >= D && (fp8 ? (ldq % 8 == 0 && ldq >= D && q_scale) : (ldy % 8 == 0 && ldy >= D)),
                 "b200_ln_modulate: bad leading dimension");
  B200_CHECK_ARG((ln_w == nullptr) == (ln_b == nullptr), "b200_ln_modulate: affine weight and bias go together");
  B200_CHECK_ARG((scale == nullptr) == (shift == nullptr), "b200_ln_modulate: scale and shift go together");
  const auto* xp = reinterpret_cast<const __nv_bfloat16*>(x);
  auto* yp = reinterpret_cast<__nv_bfloat16*>(y);
  const auto* w = reinterpret_cast<const __nv_bfloat16*>(ln_w);
  const auto* b = reinterpret_cast<const __nv_bfloat16*>(ln_b);
  const auto* sc = reinterpret_cast<const __nv_bfloat16*>(scale);
  const auto* sh = reinterpret_cast<const __nv_bfloat16*>(shift);
  auto* q = reinterpret_cast<uint8_t*>(q8);
  dim3 grid((unsigned)rows);
#define B200_LN_LAUNCH(AFF, MOD, F8) \
  ln_modulate_kernel<AFF, MOD, F8><<<grid, ROW_THREADS, 0, stream>>>(xp, ldx, yp, ldy, w, b, sc, sh, D, eps, q, ldq, q_scale)
  if (fp8) {
    if (w && sc) B200_LN_LAUNCH(true, true, true);
    else if (w) B200_LN_LAUNCH(true, false, true);
    else if (sc) B200_LN_LAUNCH(false, true, true);
    else B200_LN_LAUNCH(false, false, true);
  } else {
    if (w && sc) B200_LN_LAUNCH(true, true, false);
    else if (w) B200_LN_LAUNCH(true, false, false);
    else if (sc) B200_LN_LAUNCH(false, true, false);
    else B200_LN_LAUNCH(false, false, false);
  }
#undef B200_LN_LAUNCH
  note_launch();
  B200_CHECK_CUDA(cudaGetLastError());
  return B200_OK;
}

int ln_modulate(const void* x, long long ldx, void* y, long long ldy, const void* ln_w, const void* ln_b,
                const void* scale, const void* shift, long long rows, int D, float eps, cudaStream_t stream) {
  return ln_modulate_impl(x, ldx, y, ldy, ln_w, ln_b, scale, shift, rows, D, eps, nullptr, 0, nullptr, stream);
}

int ln_modulate_fp8(const void* x, long long ldx, void* q8, long long ldq, float* q_scale, const void* ln_w,
                    const void* ln_b, const void* scale, const void* shift, long long rows, int D, float eps,
                    cudaStream_t stream) {
  B200_CHECK_ARG(q8 && q_scale, "b200_ln_modulate_fp8: null output pointer");
  return ln_modulate_impl(x, ldx, nullptr, 0, ln_w, ln_b, scale, shift, rows, D, eps, q8, ldq, q_scale, stream);
}

// ---------------------------------------------------------------------------------------------------------
// Dynamic per-token e4m3 quantisation of a bf16 [rows, D] tensor (act_quant_fp8_perchannel_sym_vllm, mm_weight.py:236-238).
// 256 threads per row, row in registers: read once (2 B/elem), write once (1 B/elem + 4 B/row).
// ---------------------------------------------------------------------------------------------------------
constexpr int QUANT_THREADS = 256;
constexpr int QUANT_MAX_VEC = 8;   // D <= 16384

__global__ void __launch_bounds__(QUANT_THREADS)
quant_fp8_kernel(const __nv_bfloat16* __restrict__ x, long long ldx, uint8_t* __restrict__ q8, long long ldq,
                 float* __restrict__ q_scale, int D) {
  __shared__ float red[QUANT_THREADS / 32];
  const long long row = blockIdx.x;
  const int nvec = D >> 3;
  const uint4* xr = reinterpret_cast<const uint4*>(x + row * ldx);
  uint4 raw[QUANT_MAX_VEC];
  float amax = 0.f;
#pragma unroll
  for (int i = 0; i < QUANT_MAX_VEC; ++i) {
    const int v = threadIdx.x + i * QUANT_THREADS;
    if (v < nvec) {
      raw[i] = xr[v];
      float f[8];
      unpack8(raw[i], f);
#pragma unroll
      for (int e = 0; e < 8; ++e) amax = fmaxf(amax, fabsf(f[e]));
    }
  }
  const float sc = fp8_token_scale(block_max(amax, red, QUANT_THREADS / 32));
  if (threadIdx.x == 0) q_scale[row] = sc;
  uint2* qr = reinterpret_cast<uint2*>(q8 + row * ldq);
#pragma unroll
  for (int i = 0; i < QUANT_MAX_VEC; ++i) {
    const int v = threadIdx.x + i * QUANT_THREADS;
    if (v < nvec) {
      float f[8];
      unpack8(raw[i], f);
#pragma unroll
      for (int e = 0; e < 8; ++e) f[e] = f[e] / sc;
      qr[v] = pack8_e4m3(f);
    }
  }
}

int quant_fp8_per_token(const void* x, long long ldx, void* q8, long long ldq, float* q_scale, long long rows, int D,
                        cudaStream_t stream) {
  B200_CHECK_ARG(x && q8 && q_scale, "b200_quant_fp8_per_token: null pointer");
  B200_CHECK_ARG(rows > 0 && D > 0 && D % 8 == 0 && D <= QUANT_THREADS * QUANT_MAX_VEC * 8,
                 "b200_quant_fp8_per_token: D=%d must be a multiple of 8 and <= %d", D, QUANT_THREADS * QUANT_MAX_VEC * 8);
  B200_CHECK_ARG(ldx % 8 == 0 && ldx >= D && ldq % 8 == 0 && ldq >= D, "b200_quant_fp8_per_token: bad leading dimension");
  quant_fp8_kernel<<<(unsigned)rows, QUANT_THREADS, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(x), ldx,
                                                               reinterpret_cast<uint8_t*>(q8), ldq, q_scale, D); note_launch();
  B200_CHECK_CUDA(cudaGetLastError());
  return B200_OK;
}

// ---------------------------------------------------------------------------------------------------------
// RMSNorm over `norm_dim` contiguous elements (Wan: the whole row, D; Hunyuan: one head, 128), bf16 arithmetic as in
// the reference fallback, then optional RoPE on adjacent pairs with a [rows, 64] (cos, sin) fp32 table.
// In place on up to two tensors at once (q and k of one fused QKV buffer): blockIdx.y selects the tensor.
// Rows >= rope_rows are normalised but not rotated (apply_rotary_emb keeps x[seq_len:] as is, utils.py:114).
// ---------------------------------------------------------------------------------------------------------
struct RmsRopeArgs {
  __nv_bfloat16* x[2];
  const __nv_bfloat16* w[2];
  long long ld[2];
};

template <bool kRope>
__global__ void __launch_bounds__(ROW_THREADS)
rms_rope_kernel(RmsRopeArgs a, int D, float eps, const float2* __restrict__ cs, long long rope_rows,
                long long pos_offset) {
  __shared__ float red[4];
  const long long row = blockIdx.x;
  const int which = blockIdx.y;
  const int nvec = D >> 3;
  uint4* xr = reinterpret_cast<uint4*>(a.x[which] + row * a.ld[which]);
  const uint4* wr = reinterpret_cast<const uint4*>(a.w[which]);
  uint4 raw[ROW_MAX_VEC];
  float ssq = 0.f;
#pragma unroll
  for (int i = 0; i < ROW_MAX_VEC; ++i) {
    const int v = threadIdx.x + i * ROW_THREADS;
    if (v < nvec) {
      raw[i] = xr[v];
      float f[8];
      unpack8(raw[i], f);
#pragma unroll
      for (int e = 0; e < 8; ++e) ssq += bf16_round(f[e] * f[e]);   // x.pow(2) is a bf16 tensor
    }
  }
  // mean(-1) accumulates in fp32 and rounds to bf16; + eps and rsqrt are bf16 tensor ops
  float ms = bf16_round(block_sum_128(ssq, red) / (float)D);
  ms = bf16_round(ms + eps);
  const float rinv = bf16_round(rsqrtf(ms));
  const bool rotate = kRope && row < rope_rows;
  const float2* csr = cs + (row + pos_offset) * 64;
#pragma unroll
  for (int i = 0; i < ROW_MAX_VEC; ++i) {
    const int v = threadIdx.x + i * ROW_THREADS;
    if (v < nvec) {
      float f[8], w[8];
      unpack8(raw[i], f);
      unpack8(__ldg(wr + v), w);
#pragma unroll
      for (int e = 0; e < 8; ++e) f[e] = bf16_round(bf16_round(f[e] * rinv) * w[e]);
      if (rotate) {
        const int pair0 = ((v * 8) & 127) >> 1;  // pair index inside the head (head_dim 128 -> 64 pairs)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 c = __ldg(csr + pair0 + e);
          const float re = f[2 * e], im = f[2 * e + 1];
          f[2 * e] = re * c.x - im * c.y;
          f[2 * e + 1] = re * c.y + im * c.x;
        }
      }
      xr[v] = pack8(f);
    }
  }
}

int rms_rope(void* x0, long long ld0, const void* w0, void* x1, long long ld1, const void* w1, long long rows, int D,
             float eps, const void* cos_sin, long long rope_rows, long long pos_offset, cudaStream_t stream) {
  B200_CHECK_ARG(x0 && w0, "b200_rms_rope: null pointer");
  B200_CHECK_ARG(rows > 0 && D > 0 && D % 128 == 0 && D <= ROW_THREADS * ROW_MAX_VEC * 8,
                 "b200_rms_rope: D=%d must be a multiple of 128 and <= %d", D, ROW_THREADS * ROW_MAX_VEC * 8);
  B200_CHECK_ARG(ld0 % 8 == 0 && ld0 >= D && (x1 == nullptr || (ld1 % 8 == 0 && ld1 >= D && w1)),
                 "b200_rms_rope: bad leading dimension / missing weight");
  RmsRopeArgs a;
  a.x[0] = reinterpret_cast<__nv_bfloat16*>(x0);
  a.w[0] = reinterpret_cast<const __nv_bfloat16*>(w0);
  a.ld[0] = ld0;
  a.x[1] = reinterpret_cast<__nv_bfloat16*>(x1);
  a.w[1] = reinterpret_cast<const __nv_bfloat16*>(w1);
  a.ld[1] = ld1;
  dim3 grid((unsigned)rows, x1 ? 2 : 1);
  if (cos_sin)
    rms_rope_kernel<true><<<grid, ROW_THREADS, 0, stream>>>(a, D, eps, reinterpret_cast<const float2*>(cos_sin),
                                                            rope_rows, pos_offset);
  else
    rms_rope_kernel<false><<<grid, ROW_THREADS, 0, stream>>>(a, D, eps, nullptr, 0, 0);
  note_launch();
  B200_CHECK_CUDA(cudaGetLastError());
  return B200_OK;
}

// ---------------------------------------------------------------------------------------------------------
// HunyuanVideo flavour: RMSNorm PER HEAD (over the 128 elements of one head, weight [128]) on q and k laid out [rows, H, 128]
// (lightx2v/models/networks/hunyuan/infer/transformer_infer.py:289-292, 342-343; bf16 chain of utils_bf16.py:5-8), then
// x*cos + rotate_half(x)*sin in bf16 arithmetic with bf16 cos/sin (utils_bf16.py:11-31) for rows < rope_rows (image tokens).
// 16 lanes own one head (8 elements each); a 128-thread CTA covers 8 heads per pass.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(ROW_THREADS)
rms_rope_heads_kernel(RmsRopeArgs a, int H, float eps, const float2* __restrict__ cs, long long rope_rows) {
  const long long row = blockIdx.x;
  const int which = blockIdx.y;
  const int l16 = threadIdx.x & 15;
  __nv_bfloat16* xrow = a.x[which] + row * a.ld[which];
  float w[8];
  unpack8(__ldg(reinterpret_cast<const uint4*>(a.w[which]) + l16), w);
  const bool rotate = cs != nullptr && row < rope_rows;
  float2 c[4];
  if (rotate) {
#pragma unroll
    for (int e = 0; e < 4; ++e) c[e] = __ldg(cs + row * 64 + l16 * 4 + e);
  }
  for (int head = threadIdx.x >> 4; head < H; head += ROW_THREADS / 16) {
    uint4* p = reinterpret_cast<uint4*>(xrow + head * 128) + l16;
    float f[8];
    unpack8(*p, f);
    float ssq = 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) ssq += bf16_round(f[e] * f[e]);
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) ssq += __shfl_xor_sync(0xffffffffu, ssq, o);
    float ms = bf16_round(ssq / 128.0f);
    ms = bf16_round(ms + eps);
    const float rinv = bf16_round(rsqrtf(ms));
#pragma unroll
    for (int e = 0; e < 8; ++e) f[e] = bf16_round(bf16_round(f[e] * rinv) * w[e]);
    if (rotate) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float re = f[2 * e], im = f[2 * e + 1];
        // x*cos + rotate_half(x)*sin with rotate_half = (-im, re); every product and the sum are bf16 tensors
        f[2 * e] = bf16_round(re * c[e].x) + bf16_round(-im * c[e].y);
        f[2 * e + 1] = bf16_round(im * c[e].x) + bf16_round(re * c[e].y);
      }
    }
    *p = pack8(f);
  }
}

int rms_rope_heads(void* x0, long long ld0, const void* w0, void* x1, long long ld1, const void* w1, long long rows,
                   int H, float eps, const void* cos_sin, long long rope_rows, cudaStream_t stream) {
  B200_CHECK_ARG(x0 && w0, "b200_rms_rope_heads: null pointer");
  B200_CHECK_ARG(rows > 0 && H > 0, "b200_rms_rope_heads: empty problem");
  B200_CHECK_ARG(ld0 % 8 == 0 && ld0 >= 128LL * H && (x1 == nullptr || (ld1 % 8 == 0 && ld1 >= 128LL * H && w1)),
                 "b200_rms_rope_heads: bad leading dimension / missing weight");
  RmsRopeArgs a;
  a.x[0] = reinterpret_cast<__nv_bfloat16*>(x0);
  a.w[0] = reinterpret_cast<const __nv_bfloat16*>(w0);
  a.ld[0] = ld0;
  a.x[1] = reinterpret_cast<__nv_bfloat16*>(x1);
  a.w[1] = reinterpret_cast<const __nv_bfloat16*>(w1);
  a.ld[1] = ld1;
  dim3 grid((unsigned)rows, x1 ? 2 : 1);
  rms_rope_heads_kernel<<<grid, ROW_THREADS, 0, stream>>>(a, H, eps, reinterpret_cast<const float2*>(cos_sin), rope_rows); note_launch();
  B200_CHECK_CUDA(cudaGetLastError());
  return B200_OK;
}

// ---------------------------------------------------------------------------------------------------------
// CogVideoX flavour: affine LayerNorm PER HEAD over head_dim 64 (weight / bias [64], F.layer_norm in bf16 -> bf16) on q and k laid
// out [rows, H, 64], then for rows >= rope_start (the video tokens behind the text tokens)
//   out = bf16( x.float() * cos + rotate(x).float() * sin ),  rotate(x)[2i] = -x[2i+1], rotate(x)[2i+1] = x[2i]
// (lightx2v/models/networks/cogvideox/infer/transformer_infer.py:99-103 and apply_rotary_emb :5-36; fp32 products and sum, one
// rounding).  Table cs[(row - rope_start), 32] float2 = (cos, sin) of pair i.  8 lanes own one head (8 elements = 4 pairs each).
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(ROW_THREADS)
ln_rope_heads64_kernel(RmsRopeArgs a, const __nv_bfloat16* __restrict__ b0, const __nv_bfloat16* __restrict__ b1, long long rows, int H, float eps,
                       const float2* __restrict__ cs, long long rope_start) {
  const int which = blockIdx.y;
  const int l8 = threadIdx.x & 7;
  float w[8], b[8];
  unpack8(__ldg(reinterpret_cast<const uint4*>(a.w[which]) + l8), w);
  unpack8(__ldg(reinterpret_cast<const uint4*>(which ? b1 : b0) + l8), b);
  const long long units = rows * H;
  for (long long u = (long long)blockIdx.x * (ROW_THREADS / 8) + (threadIdx.x >> 3); u < units; u += (long long)gridDim.x * (ROW_THREADS / 8)) {
    const long long row = u / H;
    const int head = (int)(u - row * H);
    uint4* p = reinterpret_cast<uint4*>(a.x[which] + row * a.ld[which] + head * 64) + l8;
    float f[8];
    unpack8(*p, f);
    float sum = 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) sum += f[e];
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float mean = sum * (1.0f / 64.0f);
    float sq = 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) sq += (f[e] - mean) * (f[e] - mean);
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    const float rstd = rsqrtf(sq * (1.0f / 64.0f) + eps);
#pragma unroll
    for (int e = 0; e < 8; ++e) f[e] = bf16_round(fmaf((f[e] - mean) * rstd, w[e], b[e]));
    if (cs != nullptr && row >= rope_start) {
      const float2* c = cs + (row - rope_start) * 32 + l8 * 4;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 t = __ldg(c + e);
        const float re = f[2 * e], im = f[2 * e + 1];
        f[2 * e] = __fadd_rn(__fmul_rn(re, t.x), __fmul_rn(-im, t.y));
        f[2 * e + 1] = __fadd_rn(__fmul_rn(im, t.x), __fmul_rn(re, t.y));
      }
    }
    *p = pack8(f);
  }
}

int ln_rope_heads64(void* x0, long long ld0, const void* w0, const void* b0, void* x1, long long ld1, const void* w1, const void* b1, long long rows,
                    int H, float eps, const void* cos_sin, long long rope_start, cudaStream_t stream) {
  B200_CHECK_ARG(x0 && w0 && b0, "b200_ln_rope_heads64: null pointer");
  B200_CHECK_ARG(rows > 0 && H > 0, "b200_ln_rope_heads64: empty problem");
  B200_CHECK_ARG(ld0 % 8 == 0 && ld0 >= 64LL * H && (x1 == nullptr || (ld1 % 8 == 0 && ld1 >= 64LL * H && w1 && b1)),
                 "b200_ln_rope_heads64: bad leading dimension / missing weight");
  B200_CHECK_ARG(rope_start >= 0, "b200_ln_rope_heads64: negative rope_start");
  RmsRopeArgs a;
  a.x[0] = reinterpret_cast<__nv_bfloat16*>(x0);
  a.w[0] = reinterpret_cast<const __nv_bfloat16*>(w0);
  a.ld[0] = ld0;
  a.x[1] = reinterpret_cast<__nv_bfloat16*>(x1);
  a.w[1] = reinterpret_cast<const __nv_bfloat16*>(w1);
  a.ld[1] = ld1;
  const long long want = (rows * H + ROW_THREADS / 8 - 1) / (ROW_THREADS / 8);
  const long long cap = (long long)num_sms() * 16;
  dim3 grid((unsigned)(want < cap ? want : cap), x1 ? 2 : 1);
  ln_rope_heads64_kernel<<<grid, ROW_THREADS, 0, stream>>>(a, reinterpret_cast<const __nv_bfloat16*>(b0), reinterpret_cast<const __nv_bfloat16*>(b1), rows, H,
                                                         eps, reinterpret_cast<const float2*>(cos_sin), rope_start); note_launch();
  B200_CHECK_CUDA(cudaGetLastError());
  return B200_OK;
}

// ---------------------------------------------------------------------------------------------------------
// Ulysses head scatter fused into the q/k RMSNorm + RoPE pass (and a plain copy for v): instead of normalising in
// place and then packing + all-to-all'ing, every 16-byte vector is stored STRAIGHT INTO THE PEER GPU that owns its head
// (NVLink P2P stores through NVSwitch), in the layout the attention kernel reads: recv[dest][token_global][which][h % hp][128].
// Replaces all2all_seq2head x3 + the .contiguous() transposes + torch.cuda.synchronize()
// (lightx2v/attentions/distributed/ulysses/attn.py:41-48, comm/all2all.py:7-44).
// ---------------------------------------------------------------------------------------------------------
struct ScatterArgs {
  __nv_bfloat16* peer[8];    // peer-mapped base of each rank's receive buffer [world*rows_per_rank, 3, hp, 128]
  int world, rank, hp;
  long long rows_per_rank;
};

__global__ void __launch_bounds__(ROW_THREADS)
rms_rope_scatter_kernel(const __nv_bfloat16* __restrict__ qkv, long long ld, const __nv_bfloat16* __restrict__ wq,
                        const __nv_bfloat16* __restrict__ wk, int D, float eps, const float2* __restrict__ cs,
                        long long rope_rows, ScatterArgs sc) {
  __shared__ float red[4];
  const long long row = blockIdx.x;
  const int which = blockIdx.y;                         // 0 = q, 1 = k, 2 = v
  const int nvec = D >> 3;
  const uint4* xr = reinterpret_cast<const uint4*>(qkv + row * ld + (long long)which * D);
  const uint4* wr = reinterpret_cast<const uint4*>(which == 0 ? wq : wk);
  const long long token = (long long)sc.rank * sc.rows_per_rank + row;
  uint4 raw[ROW_MAX_VEC];
  float ssq = 0.f;
#pragma unroll
  for (int i = 0; i < ROW_MAX_VEC; ++i) {
    const int v = threadIdx.x + i * ROW_THREADS;
    if (v < nvec) {
      raw[i] = xr[v];
      float f[8];
      unpack8(raw[i], f);
#pragma unroll
      for (int e = 0; e < 8; ++e) ssq += bf16_round(f[e] * f[e]);
    }
  }
  float rinv = 1.0f;
  if (which < 2) {                                      // block-uniform branch
    float ms = bf16_round(block_sum_128(ssq, red) / (float)D);
    ms = bf16_round(ms + eps);
    rinv = bf16_round(rsqrtf(ms));
  }
  const bool rotate = which < 2 && row < rope_rows;
  const float2* csr = cs + row * 64;
#pragma unroll
  for (int i = 0; i < ROW_MAX_VEC; ++i) {
    const int v = threadIdx.x + i * ROW_THREADS;
    if (v < nvec) {
      float f[8];
      unpack8(raw[i], f);
      if (which < 2) {
        float w[8];
        unpack8(__ldg(wr + v), w);
#pragma unroll
        for (int e = 0; e < 8; ++e) f[e] = bf16_round(bf16_round(f[e] * rinv) * w[e]);
        if (rotate) {
          const int pair0 = ((v * 8) & 127) >> 1;
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 c = __ldg(csr + pair0 + e);
            const float re = f[2 * e], im = f[2 * e + 1];
            f[2 * e] = re * c.x - im * c.y;
            f[2 * e + 1] = re * c.y + im * c.x;
          }
        }
      }
      const int head = v >> 4;                          // 16 vectors of 8 elements per 128-wide head
      const int dest = head / sc.hp;
      __nv_bfloat16* dst = sc.peer[dest] + ((token * 3 + which) * sc.hp + (head - dest * sc.hp)) * 128 + (v & 15) * 8;
      *reinterpret_cast<uint4*>(dst) = pack8(f);
    }
  }
}

int rms_rope_scatter(const void* qkv, long long ld, const void* wq, const void* wk, long long rows, int D, float eps,
                     const void* cos_sin, long long rope_rows, void* const* peers, int world, int rank,
                     long long rows_per_rank, cudaStream_t stream) {
  B200_CHECK_ARG(qkv && wq && wk && cos_sin && peers, "b200_rms_rope_scatter: null pointer");
  B200_CHECK_ARG(world >= 1 && world <= 8 && rank >= 0 && rank < world, "b200_rms_rope_scatter: world %d / rank %d out of range", world, rank);
  B200_CHECK_ARG(rows > 0 && rows <= rows_per_rank && D % 128 == 0 && D <= ROW_THREADS * ROW_MAX_VEC * 8 && (D / 128) % world == 0,
                 "b200_rms_rope_scatter: bad shape rows=%lld D=%d world=%d", rows, D, world);
  B200_CHECK_ARG(ld % 8 == 0 && ld >= 3LL * D, "b200_rms_rope_scatter: qkv leading dimension must cover 3*D");
  ScatterArgs sc;
  for (int i = 0; i < 8; ++i) sc.peer[i] = i < world ? reinterpret_cast<__nv_bfloat16*>(peers[i]) : nullptr;
  for (int i = 0; i < world; ++i) B200_CHECK_ARG(peers[i] != nullptr, "b200_rms_rope_scatter: null peer pointer %d", i);
  sc.world = world;
  sc.rank = rank;
  sc.hp = D / 128 / world;
  sc.rows_per_rank = rows_per_rank;
  dim3 grid((unsigned)rows, 3);
  rms_rope_scatter_kernel<<<grid, ROW_THREADS, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(qkv), ld,
                                                          reinterpret_cast<const __nv_bfloat16*>(wq),
                                                          reinterpret_cast<const __nv_bfloat16*>(wk), D, eps,
                                                          reinterpret_cast<const float2*>(cos_sin), rope_rows, sc); note_launch();
  B200_CHECK_CUDA(cudaGetLastError());
  return B200_OK;
}

}  // namespace b200
