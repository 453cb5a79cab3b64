// Element-wise / layout kernels of the Wan VAE decoder on channels-last bf16 activations.
//   rms_silu      : RMS_norm (per-voxel L2 normalisation over channels, * sqrt(C) * gamma) + SiLU
//                   (lightx2v/models/video_encoders/hf/wan/vae.py:47-59 RMS_norm, :192-195 the norm->SiLU->conv pattern)
//   latent_to_cl  : z[16,T,H,W] fp32 -> (z / inv_std + mean) -> channels-last bf16 [T,H,W,32] (zero-padded channels)  (vae.py:716-719)
//   cl_to_video   : channels-last bf16 [T,H,W,16] (3 valid) -> fp32 [3,T,H,W]                                           (vae.py:951)
#include "host_util.cuh"
#include "ptx.cuh"

namespace b200 {

// C = 96 / 192: 16 / 32 lanes reserved per voxel (12 / 24 of them own 8 channels = 16 bytes each), 2 / 1 voxels per warp pass, U = 4
// passes in flight.  C = 384: a warp takes two voxels per pass = 96 vectors of 16 bytes = exactly three per lane (vector v = lane + 32 j,
// voxel v / 48), so every lane is busy and 48 bytes per lane are in flight per pass.  Instruction budget: see silu_fast in ptx.cuh.
template <int C>
__global__ void __launch_bounds__(256)
rms_silu_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, const float* __restrict__ gamma,
                long long voxels, int apply_silu) {
  const int lane = threadIdx.x & 31;
  const long long warp_global = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  const float sqrt_c = sqrtf((float)C);
  if constexpr (C <= 256) {
    constexpr int LPV = C / 8;                 // lanes per voxel
    constexpr int G = (LPV <= 16) ? 16 : 32;   // lanes reserved per voxel (power of two)
    constexpr int VPW = 32 / G;                // voxels per warp pass
    const int sub = lane / G, l = lane % G;
    float g[8];
    if (l < LPV) {
#pragma unroll
      for (int e = 0; e < 8; ++e) g[e] = gamma[l * 8 + e] * sqrt_c;
    }
    constexpr int U = 4;   // independent 16-byte loads in flight per lane (HBM-bound: keep several MB outstanding chip-wide)
    const long long stride = nwarps * VPW;
    for (long long v0 = warp_global * VPW; v0 < voxels; v0 += U * stride) {
      uint4 raw[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long v = v0 + u * stride + sub;
        raw[u] = make_uint4(0, 0, 0, 0);
        if (v < voxels && l < LPV) raw[u] = ld_nc_v4(x + v * C + l * 8);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long v = v0 + u * stride + sub;
        const bool ok = v < voxels && l < LPV;
        float f[8] = {bf16_lo(raw[u].x), bf16_hi(raw[u].x), bf16_lo(raw[u].y), bf16_hi(raw[u].y),
                      bf16_lo(raw[u].z), bf16_hi(raw[u].z), bf16_lo(raw[u].w), bf16_hi(raw[u].w)};
        float2 s2 = make_float2(0.f, 0.f);
#pragma unroll
        for (int e = 0; e < 8; e += 2) s2 = __ffma2_rn(make_float2(f[e], f[e + 1]), make_float2(f[e], f[e + 1]), s2);
        float ss = s2.x + s2.y;
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
        const float inv = rcp_approx(fmaxf(sqrtf(ss), 1e-12f));        // F.normalize: x / max(||x||_2, eps)
        if (ok) {
          const float2 inv2 = make_float2(inv, inv);
#pragma unroll
          for (int e = 0; e < 8; e += 2) {
            const float2 t = __fmul2_rn(__fmul2_rn(make_float2(f[e], f[e + 1]), inv2), make_float2(g[e], g[e + 1]));
            f[e] = t.x;
            f[e + 1] = t.y;
          }
          if (apply_silu) silu8(f);
          uint4 o;
          o.x = pack_bf16(f[0], f[1]); o.y = pack_bf16(f[2], f[3]); o.z = pack_bf16(f[4], f[5]); o.w = pack_bf16(f[6], f[7]);
          *reinterpret_cast<uint4*>(y + v * C + l * 8) = o;
        }
      }
    }
  } else {
    static_assert(C == 384, "the three-vectors-per-lane mapping is written for C = 384");
    // vector v = lane + 32 j (j = 0, 1, 2) of a voxel PAIR: voxel v / 48, channels 8 (v % 48) ..
    int vox_of[3], c0_of[3];
    float g[3][8];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const int v = lane + 32 * j;
      vox_of[j] = v / 48;
      c0_of[j] = (v % 48) * 8;
#pragma unroll
      for (int e = 0; e < 8; ++e) g[j][e] = gamma[c0_of[j] + e] * sqrt_c;
    }
    constexpr int U = 2;
    const long long pairs = (voxels + 1) / 2;
    for (long long p0 = warp_global; p0 < pairs; p0 += U * nwarps) {
      uint4 raw[U][3];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long pr = p0 + u * nwarps;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const long long v = 2 * pr + vox_of[j];
          raw[u][j] = make_uint4(0, 0, 0, 0);
          if (pr < pairs && v < voxels) raw[u][j] = ld_nc_v4(x + v * C + c0_of[j]);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long pr = p0 + u * nwarps;
        float f[3][8];
        float s0 = 0.f, s1 = 0.f;     // partial sums of squares of voxel 0 / voxel 1 of the pair
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const uint32_t w[4] = {raw[u][j].x, raw[u][j].y, raw[u][j].z, raw[u][j].w};
          float ss = 0.f;
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            f[j][2 * e] = bf16_lo(w[e]);
            f[j][2 * e + 1] = bf16_hi(w[e]);
            ss += f[j][2 * e] * f[j][2 * e] + f[j][2 * e + 1] * f[j][2 * e + 1];
          }
          if (vox_of[j] == 0) s0 += ss; else s1 += ss;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          s0 += __shfl_xor_sync(0xffffffffu, s0, o);
          s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        }
        const float inv0 = rcp_approx(fmaxf(sqrtf(s0), 1e-12f)), inv1 = rcp_approx(fmaxf(sqrtf(s1), 1e-12f));
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const long long v = 2 * pr + vox_of[j];
          if (pr < pairs && v < voxels) {
            const float inv = vox_of[j] == 0 ? inv0 : inv1;
#pragma unroll
            for (int e = 0; e < 8; ++e) f[j][e] = f[j][e] * inv * g[j][e];
            if (apply_silu) silu8(f[j]);
            uint4 o;
            o.x = pack_bf16(f[j][0], f[j][1]); o.y = pack_bf16(f[j][2], f[j][3]); o.z = pack_bf16(f[j][4], f[j][5]); o.w = pack_bf16(f[j][6], f[j][7]);
            *reinterpret_cast<uint4*>(y + v * C + c0_of[j]) = o;
          }
        }
      }
    }
  }
}

int rms_silu_cl(const void* x, void* y, const float* gamma, long long voxels, int C, int apply_silu, cudaStream_t stream) {
  B200_CHECK_ARG(x && y && gamma, "b200_rms_silu_cl: null pointer");
  B200_CHECK_ARG(voxels > 0, "b200_rms_silu_cl: empty tensor");
  const auto* xp = reinterpret_cast<const __nv_bfloat16*>(x);
  auto* yp = reinterpret_cast<__nv_bfloat16*>(y);
  const int blocks = num_sms() * 8;
  switch (C) {
    case 96: rms_silu_kernel<96><<<blocks, 256, 0, stream>>>(xp, yp, gamma, voxels, apply_silu); note_launch(); break;
    case 192: rms_silu_kernel<192><<<blocks, 256, 0, stream>>>(xp, yp, gamma, voxels, apply_silu); note_launch(); break;
    case 384: rms_silu_kernel<384><<<blocks, 256, 0, stream>>>(xp, yp, gamma, voxels, apply_silu); note_launch(); break;
    default:
      set_last_error("b200_rms_silu_cl: unsupported channel count %d (96 / 192 / 384)", C);
      return B200_ERR_UNSUPPORTED;
  }
  B200_CHECK_CUDA(cudaGetLastError());
  return B200_OK;
}

// z [CZ, T, H, W] fp32 -> out [T, H, W, CP] bf16 with out[..., c] = z[c] / inv_std[c] + mean[c] for c < CZ, 0 otherwise
__global__ void latent_to_cl_kernel(const float* __restrict__ z, __nv_bfloat16* __restrict__ out, const float* __restrict__ mean,
                                    const float* __restrict__ inv_std, long long voxels, int CZ, int CP) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= voxels * CP) return;
  const long long v = i / CP;
  const int c = (int)(i - v * CP);
  float val = 0.f;
  if (c < CZ) val = z[(long long)c * voxels + v] / inv_std[c] + mean[c];
  out[i] = __float2bfloat16_rn(val);
}

int latent_to_cl(const float* z, void* out, const float* mean, const float* inv_std, long long voxels, int CZ, int CP,
                 cudaStream_t stream) {
  B200_CHECK_ARG(z && out && mean && inv_std && voxels > 0 && CZ > 0 && CP >= CZ, "b200_latent_to_cl: bad arguments");
  const long long n = voxels * CP;
  latent_to_cl_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(z, reinterpret_cast<__nv_bfloat16*>(out), mean, inv_std, voxels, CZ, CP); note_launch();
  B200_CHECK_CUDA(cudaGetLastError());
  return B200_OK;
}

// in [voxels, CP] bf16 (first 3 channels valid) -> out [3, voxels] fp32, channel planes `cstride` elements apart (>= voxels: a frame
// range of a larger [3, T, H, W] video when the decode is chunked along T)
__global__ void cl_to_video_kernel(const __nv_bfloat16* __restrict__ in, float* __restrict__ out, long long voxels, int CP, long long cstride) {
  const long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= voxels) return;
  const __nv_bfloat16* p = in + v * CP;
  out[v] = __bfloat162float(p[0]);
  out[cstride + v] = __bfloat162float(p[1]);
  out[2 * cstride + v] = __bfloat162float(p[2]);
}

int cl_to_video(const void* in, float* out, long long voxels, int CP, long long out_channel_stride, cudaStream_t stream) {
  B200_CHECK_ARG(in && out && voxels > 0 && CP >= 3, "b200_cl_to_video: bad arguments");
  const long long cstride = out_channel_stride > 0 ? out_channel_stride : voxels;
  B200_CHECK_ARG(cstride >= voxels, "b200_cl_to_video: channel stride %lld smaller than the plane (%lld voxels)", cstride, voxels);
  cl_to_video_kernel<<<(unsigned)((voxels + 255) / 256), 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(in), out, voxels, CP, cstride); note_launch();
  B200_CHECK_CUDA(cudaGetLastError());
  return B200_OK;
}

}  // namespace b200
