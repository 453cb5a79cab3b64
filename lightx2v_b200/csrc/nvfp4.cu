// NVFP4 activation / weight quantiser: bf16 [rows, K] -> packed e2m1 [rows, K/2] + ue4m3 scale factors (one per 16 elements) in the
// 128x4 layout the block-scaled tcgen05 MMA reads: byte offset of (row m, K-group g) =
//   ((m / 128) * (K / 64) + g / 4) * 512 + (m % 32) * 16 + ((m % 128) / 32) * 4 + g % 4.
//
// Replaces lightx2v_kernel's scaled_fp4_quant (lightx2v_kernel/python/lightx2v_kernel/gemm.py:11-52; kernel
// csrc/gemm/nvfp4_quant_kernels_sm120.cu:118-290, built for sm_120a only).  Arithmetic follows the reference's own golden
// restatement lightx2v_kernel/test/nvfp4_nvfp4/fake_quant.py:37-51 operation for operation (exact reciprocal, not rcp.approx), so
// the packed bytes and scale bytes are bit-identical to it:
//   sf      = e4m3_rn(global_scale * (max|x_group| * (1/6)))
//   x_q     = e2m1_rn(clamp(x * (global_scale * (sf == 0 ? 0 : 1 / sf)), -6, 6))
// HBM-bound: 2 bytes read, 0.5 + 1/16 bytes written per element.
#include <cuda_fp8.h>

#include "host_util.cuh"
#include "ptx.cuh"

namespace b200 {

__device__ __forceinline__ uint32_t e2m1x8_pack(const float (&f)[8]) {
  uint32_t val;
  asm volatile(
      "{\n\t"
      ".reg .b8 b0, b1, b2, b3;\n\t"
      "cvt.rn.satfinite.e2m1x2.f32 b0, %2, %1;\n\t"
      "cvt.rn.satfinite.e2m1x2.f32 b1, %4, %3;\n\t"
      "cvt.rn.satfinite.e2m1x2.f32 b2, %6, %5;\n\t"
      "cvt.rn.satfinite.e2m1x2.f32 b3, %8, %7;\n\t"
      "mov.b32 %0, {b0, b1, b2, b3};\n\t"
      "}\n"
      : "=r"(val)
      : "f"(f[0]), "f"(f[1]), "f"(f[2]), "f"(f[3]), "f"(f[4]), "f"(f[5]), "f"(f[6]), "f"(f[7]));
  return val;
}

// One thread per group of 16 elements; consecutive threads walk along K (coalesced 32-byte reads, 8-byte writes); the four
// scale bytes of a 64-element span are gathered into one 32-bit store by the first lane of each quad.
__global__ void __launch_bounds__(256)
quant_nvfp4_kernel(const __nv_bfloat16* __restrict__ x, long long ldx, long long rows, long long rows_padded, int K,
                   const float* __restrict__ global_scale, uint8_t* __restrict__ q, long long ldq, uint8_t* __restrict__ sf) {
  const int groups = K / 16;
  const long long total = rows_padded * groups;
  const float gs = __ldg(global_scale);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long m = i / groups;
    const int g = (int)(i - m * groups);
    uint32_t sf_byte = 0;
    if (m < rows) {
      const uint4* src = reinterpret_cast<const uint4*>(x + m * ldx + (long long)g * 16);
      const uint4 r0 = src[0], r1 = src[1];
      const uint32_t w[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
      float f[16];
      float vmax = 0.f;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        f[2 * e] = bf16_lo(w[e]);
        f[2 * e + 1] = bf16_hi(w[e]);
        vmax = fmaxf(vmax, fmaxf(fabsf(f[2 * e]), fabsf(f[2 * e + 1])));
      }
      const float scale = __fmul_rn(gs, __fmul_rn(vmax, 0.16666666666666666f));
      const __nv_fp8_storage_t s8 = __nv_cvt_float_to_fp8(scale, __NV_SATFINITE, __NV_E4M3);
      sf_byte = s8;
      const float sq = __half2float(__half(__nv_cvt_fp8_to_halfraw(s8, __NV_E4M3)));
      const float out_scale = sq == 0.f ? 0.f : __fmul_rn(gs, __fdiv_rn(1.0f, sq));
      float lo[8], hi[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        lo[e] = fminf(fmaxf(__fmul_rn(f[e], out_scale), -6.0f), 6.0f);
        hi[e] = fminf(fmaxf(__fmul_rn(f[8 + e], out_scale), -6.0f), 6.0f);
      }
      uint2 o;
      o.x = e2m1x8_pack(lo);
      o.y = e2m1x8_pack(hi);
      *reinterpret_cast<uint2*>(q + m * ldq + (long long)g * 8) = o;
    }
    // gather the 4 scale bytes of this 64-element span (lanes 4j .. 4j+3 of the warp) into one word
    uint32_t word = sf_byte << (8 * (g & 3));
    word |= __shfl_xor_sync(0xffffffffu, word, 1);
    word |= __shfl_xor_sync(0xffffffffu, word, 2);
    if ((g & 3) == 0) {
      const long long off = ((m >> 7) * (K / 64) + (g >> 2)) * 512 + (m & 31) * 16 + ((m & 127) >> 5) * 4;
      *reinterpret_cast<uint32_t*>(sf + off) = word;
    }
  }
}

// q: [rows, K/2] bytes (row stride ldq); sf: [(rows rounded up to 128) * K/16] bytes, every byte written (padding rows get 0).
int quant_nvfp4(const void* x, long long ldx, long long rows, int K, const float* global_scale, void* q, long long ldq, void* sf,
                cudaStream_t stream) {
  B200_CHECK_ARG(x && q && sf && global_scale, "b200_quant_nvfp4: null pointer");
  B200_CHECK_ARG(rows > 0 && K > 0 && K % 64 == 0, "b200_quant_nvfp4: K (%d) must be a positive multiple of 64", K);
  B200_CHECK_ARG(ldx % 8 == 0 && ldx >= K && ldq % 8 == 0 && ldq >= K / 2, "b200_quant_nvfp4: bad leading dimensions");
  B200_CHECK_ARG(((uintptr_t)x % 16 == 0) && ((uintptr_t)q % 8 == 0) && ((uintptr_t)sf % 4 == 0), "b200_quant_nvfp4: misaligned pointer");
  const long long rows_padded = (rows + 127) / 128 * 128;
  const long long total = rows_padded * (K / 16);      // multiple of 4 groups per row -> quads never straddle rows
  const long long want = (total + 255) / 256;
  const int max_blocks = num_sms() * 16;
  const int blocks = (int)(want < max_blocks ? want : max_blocks);
  quant_nvfp4_kernel<<<blocks, 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(x), ldx, rows, rows_padded, K, global_scale,
                                                  reinterpret_cast<uint8_t*>(q), ldq, reinterpret_cast<uint8_t*>(sf)); note_launch();
  B200_CHECK_CUDA(cudaGetLastError());
  return B200_OK;
}

// ---- dynamic per-tensor activation scale -----------------------------------------------------------------------------------------
// global_scale = 448 * 6 / max|x| (docs/en_US/nvfp4_quantization_basics.md:52; test_bench1.py:120), alpha = 1 / (gs_x * gs_w), both
// left on the device so that quantiser and GEMM can consume them without a host round trip.  One read of x.
__global__ void __launch_bounds__(256)
absmax_kernel(const __nv_bfloat16* __restrict__ x, long long ldx, long long rows, int K, uint32_t* __restrict__ amax_bits) {
  const int vec_per_row = K / 8;
  const long long total = rows * vec_per_row;
  float m = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / vec_per_row;
    const int c = (int)(i - r * vec_per_row);
    const uint4 v = *reinterpret_cast<const uint4*>(x + r * ldx + (long long)c * 8);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) m = fmaxf(m, fmaxf(fabsf(bf16_lo(w[e])), fabsf(bf16_hi(w[e]))));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(amax_bits, __float_as_uint(m));     // non-negative floats order like their bit patterns
}

__global__ void nvfp4_scale_finalize_kernel(const uint32_t* __restrict__ amax_bits, const float* __restrict__ weight_global_scale,
                                            float* __restrict__ global_scale, float* __restrict__ alpha) {
  const float amax = __uint_as_float(*amax_bits);
  // torch evaluates `scalar / tensor` as reciprocal(tensor) * scalar (two roundings) - the form the reference's recipe is written in
  // (test_bench1.py:120-121: (FLOAT8_E4M3_MAX * FLOAT4_E2M1_MAX) / torch.amax(...)); reproduced so the scale is bit-identical
  const float gs = amax > 0.f ? __fmul_rn(__frcp_rn(amax), 448.0f * 6.0f) : 1.0f;
  *global_scale = gs;
  if (alpha != nullptr) *alpha = __fdiv_rn(1.0f, __fmul_rn(gs, weight_global_scale ? *weight_global_scale : 1.0f));
}

// scratch: one uint32 on the device (zeroed by the call).
int nvfp4_act_scale(const void* x, long long ldx, long long rows, int K, const float* weight_global_scale, float* global_scale,
                    float* alpha, void* scratch, cudaStream_t stream) {
  B200_CHECK_ARG(x && global_scale && scratch, "b200_nvfp4_act_scale: null pointer");
  B200_CHECK_ARG(rows > 0 && K > 0 && K % 8 == 0 && ldx % 8 == 0 && ldx >= K, "b200_nvfp4_act_scale: bad shape");
  B200_CHECK_CUDA(cudaMemsetAsync(scratch, 0, sizeof(uint32_t), stream));
  const long long want = (rows * (K / 8) + 255) / 256;
  const int max_blocks = num_sms() * 8;
  const int blocks = (int)(want < max_blocks ? want : max_blocks);
  absmax_kernel<<<blocks, 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(x), ldx, rows, K, reinterpret_cast<uint32_t*>(scratch)); note_launch();
  nvfp4_scale_finalize_kernel<<<1, 1, 0, stream>>>(reinterpret_cast<const uint32_t*>(scratch), weight_global_scale, global_scale, alpha); note_launch();
  B200_CHECK_CUDA(cudaGetLastError());
  return B200_OK;
}

}  // namespace b200
