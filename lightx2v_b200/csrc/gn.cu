// GroupNorm(32 groups) + SiLU + replicate padding for the HunyuanVideo causal 3-D VAE decoder, channels-last bf16 [T, H, W, C].
//
// Replaces, per ResnetBlockCausal3D half (lightx2v/models/video_encoders/hf/autoencoder_kl_causal_3d/unet_causal_3d_blocks.py
// :364-412): torch.nn.GroupNorm -> SiLU -> F.pad(mode="replicate") of CausalConv3d.forward (:88-91).  Two HBM-bound passes:
//   gn_stats     : one read of the tensor -> per-group sum / sum of squares (fp32 inside a block, fp64 across blocks, fixed order)
//   gn_apply_pad : one read + one write: y = silu((x - mean) * rstd * gamma + beta) written into a tensor that already carries
//                  the convolution's replicate border (pt frames in front, ph / pw pixels around), so that the implicit-GEMM
//                  convolution (conv3d.cu) reads it with plain non-negative tap offsets and no padding logic of its own.
// With sums == nullptr gn_apply_pad is a pure replicate-pad copy (input of the phase-decomposed UpsampleCausal3D, :146-200).
#include "host_util.cuh"
#include "ptx.cuh"

namespace b200 {

constexpr int GN_GROUPS = 32;

// Deterministic two-level reduction (no atomics: the decode must be reproducible run to run and across tile-parallel ranks):
// thread -> warp shuffles -> fixed-order sum over the 8 warps -> one partial per block; a second tiny kernel sums the block
// partials in block order in fp64.
template <int C>
__global__ void __launch_bounds__(256)
gn_stats_kernel(const __nv_bfloat16* __restrict__ x, long long voxels, float* __restrict__ partials) {
  constexpr int LPV = C / 8;                    // threads per voxel (each owns 8 consecutive channels)
  constexpr int VPB = 256 / LPV;                // voxels per block pass
  constexpr int CPG = C / GN_GROUPS;            // channels per group
  constexpr int GPT = (CPG >= 8) ? 1 : 8 / CPG; // groups touched by one thread
  constexpr int EPG = 8 / GPT;                  // of its 8 elements, how many fall in one group
  __shared__ float part[2][8][GN_GROUPS];
  for (int i = threadIdx.x; i < 2 * 8 * GN_GROUPS; i += 256) (&part[0][0][0])[i] = 0.f;
  __syncthreads();
  const int l = threadIdx.x % LPV, sub = threadIdx.x / LPV;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float s[GPT], q[GPT];
#pragma unroll
  for (int g = 0; g < GPT; ++g) s[g] = q[g] = 0.f;
  const long long stride = (long long)gridDim.x * VPB;
  constexpr int U = 4;
  for (long long v0 = (long long)blockIdx.x * VPB + sub; v0 < voxels; v0 += U * stride) {
    uint4 raw[U];
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (v0 + u * stride < voxels) raw[u] = ld_nc_v4(x + (v0 + u * stride) * C + l * 8);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (v0 + u * stride < voxels) {
        const float f[8] = {bf16_lo(raw[u].x), bf16_hi(raw[u].x), bf16_lo(raw[u].y), bf16_hi(raw[u].y),
                            bf16_lo(raw[u].z), bf16_hi(raw[u].z), bf16_lo(raw[u].w), bf16_hi(raw[u].w)};
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          s[e / EPG] += f[e];
          q[e / EPG] += f[e] * f[e];
        }
      }
    }
  }
#pragma unroll
  for (int g = 0; g < GPT; ++g) {
    // lanes of a warp that own the same channel slice (LPV < 32)
#pragma unroll
    for (int o = 16; o >= LPV; o >>= 1) {
      s[g] += __shfl_xor_sync(0xffffffffu, s[g], o);
      q[g] += __shfl_xor_sync(0xffffffffu, q[g], o);
    }
    if constexpr (CPG == 16) {   // two neighbouring threads share a group
      s[g] += __shfl_xor_sync(0xffffffffu, s[g], 1);
      q[g] += __shfl_xor_sync(0xffffffffu, q[g], 1);
    }
  }
  if (CPG == 16 ? (lane & 1) == 0 : lane < LPV) {
#pragma unroll
    for (int g = 0; g < GPT; ++g) {
      const int grp = (l * 8) / CPG + g;
      part[0][warp][grp] = s[g];
      part[1][warp][grp] = q[g];
    }
  }
  __syncthreads();
  if (threadIdx.x < 2 * GN_GROUPS) {
    const int which = threadIdx.x / GN_GROUPS, grp = threadIdx.x % GN_GROUPS;
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += part[which][w][grp];
    partials[(long long)blockIdx.x * 2 * GN_GROUPS + threadIdx.x] = t;
  }
}

// 16 threads per output: thread j sums blocks j, j+16, ... in fp64, then the 16 partial sums are added in index order (fixed order ->
// reproducible); 1024 threads keep the 1184 dependent loads of the serial version off the critical path of every GroupNorm.
__global__ void __launch_bounds__(1024) gn_stats_finalize_kernel(const float* __restrict__ partials, int blocks, double* __restrict__ sums) {
  __shared__ double part[16][2 * GN_GROUPS];
  const int out = threadIdx.x % (2 * GN_GROUPS), j = threadIdx.x / (2 * GN_GROUPS);
  double t = 0.0;
  for (int b = j; b < blocks; b += 16) t += (double)partials[(long long)b * 2 * GN_GROUPS + out];
  part[j][out] = t;
  __syncthreads();
  if (threadIdx.x < 2 * GN_GROUPS) {
    double r = 0.0;
#pragma unroll
    for (int k = 0; k < 16; ++k) r += part[k][threadIdx.x];
    sums[threadIdx.x] = r;
  }
}

// sums: fp64 workspace of gn_stats_workspace_doubles() entries; the first 64 hold the result (sum[32], sum of squares[32]).
long long gn_stats_workspace_doubles() { return 2 * GN_GROUPS + (long long)num_sms() * 8 * GN_GROUPS; }   // 64 + float partials

int gn_stats_cl(const void* x, long long voxels, int C, double* sums, cudaStream_t stream) {
  B200_CHECK_ARG(x && sums && voxels > 0, "b200_gn_stats_cl: bad arguments");
  const auto* xp = reinterpret_cast<const __nv_bfloat16*>(x);
  const int lpv = C / 8;
  const long long passes = (voxels + (256 / (lpv > 0 ? lpv : 1)) - 1) / (256 / (lpv > 0 ? lpv : 1));
  const int max_blocks = num_sms() * 8;
  const int blocks = (int)(passes < max_blocks ? passes : max_blocks);
  float* partials = reinterpret_cast<float*>(sums + 2 * GN_GROUPS);
  switch (C) {
    case 64: gn_stats_kernel<64><<<blocks, 256, 0, stream>>>(xp, voxels, partials); note_launch(); break;
    case 128: gn_stats_kernel<128><<<blocks, 256, 0, stream>>>(xp, voxels, partials); note_launch(); break;
    case 256: gn_stats_kernel<256><<<blocks, 256, 0, stream>>>(xp, voxels, partials); note_launch(); break;
    case 512: gn_stats_kernel<512><<<blocks, 256, 0, stream>>>(xp, voxels, partials); note_launch(); break;
    default:
      set_last_error("b200_gn_stats_cl: unsupported channel count %d (64 / 128 / 256 / 512)", C);
      return B200_ERR_UNSUPPORTED;
  }
  gn_stats_finalize_kernel<<<1, 16 * 2 * GN_GROUPS, 0, stream>>>(partials, blocks, sums); note_launch();
  B200_CHECK_CUDA(cudaGetLastError());
  return B200_OK;
}

template <int C>
__global__ void __launch_bounds__(256)
gn_apply_pad_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, const double* __restrict__ sums,
                    const float* __restrict__ gamma, const float* __restrict__ beta, float eps, double inv_count, int T, int H, int W,
                    int pt, int ph, int pw, int apply_silu) {
  constexpr int LPV = C / 8;
  constexpr int VPB = 256 / LPV;
  constexpr int CPG = C / GN_GROUPS;
  const int l = threadIdx.x % LPV, sub = threadIdx.x / LPV;
  float a[8], b[8];
  if (sums != nullptr) {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int c = l * 8 + e;
      const int grp = c / CPG;
      const double mean = sums[grp] * inv_count;
      const double var = fmax(sums[GN_GROUPS + grp] * inv_count - mean * mean, 0.0);
      const float rstd = rsqrtf((float)var + eps);
      a[e] = rstd * gamma[c];
      b[e] = beta[c] - (float)mean * a[e];
    }
  } else {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      a[e] = 1.f;
      b[e] = 0.f;
    }
  }
  // One output row (tp, hp) of the padded tensor = Wp voxels = Wp * LPV vectors of 16 bytes, contiguous in y; the source row (t, h) is
  // contiguous in x.  Blocks take rows grid-stride; inside a row the 256 threads stride over the vectors (LPV divides 256, so every
  // thread keeps its 8 channels - and its a[], b[] - for the whole kernel).  Index arithmetic is per ROW, not per vector: the previous
  // version spent two 64-bit divisions per 16-byte vector, ~12 issue slots per element on a pass whose whole budget is ~14.
  const int Hp = H + 2 * ph, Wp = W + 2 * pw;
  const int rows = (T + pt) * Hp;
  constexpr int VSTEP = 256 / LPV;            // voxels advanced per thread step
  constexpr int U = 4;                        // independent 16-byte loads in flight per thread
  for (int r = blockIdx.x; r < rows; r += gridDim.x) {
    const int tp = r / Hp, hp = r - tp * Hp;
    const int t = max(tp - pt, 0), h = min(max(hp - ph, 0), H - 1);
    const __nv_bfloat16* xrow = x + ((long long)t * H + h) * (long long)W * C + l * 8;
    __nv_bfloat16* yrow = y + (long long)r * Wp * C + l * 8;
    for (int wp0 = sub; wp0 < Wp; wp0 += U * VSTEP) {
      uint4 raw[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int wp = wp0 + u * VSTEP;
        if (wp < Wp) raw[u] = ld_nc_v4(xrow + (long long)min(max(wp - pw, 0), W - 1) * C);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int wp = wp0 + u * VSTEP;
        if (wp < Wp) {
          float f[8] = {bf16_lo(raw[u].x), bf16_hi(raw[u].x), bf16_lo(raw[u].y), bf16_hi(raw[u].y),
                        bf16_lo(raw[u].z), bf16_hi(raw[u].z), bf16_lo(raw[u].w), bf16_hi(raw[u].w)};
          if (sums != nullptr) {
#pragma unroll
            for (int e = 0; e < 8; e += 2) {
              const float2 y_ = __ffma2_rn(make_float2(f[e], f[e + 1]), make_float2(a[e], a[e + 1]), make_float2(b[e], b[e + 1]));
              f[e] = y_.x;
              f[e + 1] = y_.y;
            }
            if (apply_silu) silu8(f);
          }
          uint4 o;
          o.x = pack_bf16(f[0], f[1]); o.y = pack_bf16(f[2], f[3]); o.z = pack_bf16(f[4], f[5]); o.w = pack_bf16(f[6], f[7]);
          *reinterpret_cast<uint4*>(yrow + (long long)wp * C) = o;
        }
      }
    }
  }
}

int gn_apply_pad_cl(const void* x, void* y, const double* sums, const float* gamma, const float* beta, float eps, int T, int H, int W,
                    int C, int pt, int ph, int pw, int apply_silu, cudaStream_t stream) {
  B200_CHECK_ARG(x && y && T > 0 && H > 0 && W > 0, "b200_gn_apply_pad_cl: bad arguments");
  B200_CHECK_ARG(sums == nullptr || (gamma && beta), "b200_gn_apply_pad_cl: gamma / beta required with sums");
  B200_CHECK_ARG(pt >= 0 && ph >= 0 && pw >= 0, "b200_gn_apply_pad_cl: negative padding");
  const auto* xp = reinterpret_cast<const __nv_bfloat16*>(x);
  auto* yp = reinterpret_cast<__nv_bfloat16*>(y);
  const double inv_count = 1.0 / ((double)T * H * W * (C / GN_GROUPS));
  const long long rows = (long long)(T + pt) * (H + 2 * ph);
  const int max_blocks = num_sms() * 8;
  const int blocks = (int)(rows < max_blocks ? rows : max_blocks);
#define B200_GN_LAUNCH(CC) \
  gn_apply_pad_kernel<CC><<<blocks, 256, 0, stream>>>(xp, yp, sums, gamma, beta, eps, inv_count, T, H, W, pt, ph, pw, apply_silu)
  switch (C) {
    case 64: B200_GN_LAUNCH(64); break;
    case 128: B200_GN_LAUNCH(128); break;
    case 256: B200_GN_LAUNCH(256); break;
    case 512: B200_GN_LAUNCH(512); break;
    default:
      set_last_error("b200_gn_apply_pad_cl: unsupported channel count %d (64 / 128 / 256 / 512)", C);
      return B200_ERR_UNSUPPORTED;
  }
#undef B200_GN_LAUNCH
  note_launch();
  B200_CHECK_CUDA(cudaGetLastError());
  return B200_OK;
}

}  // namespace b200
