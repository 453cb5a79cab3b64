// Implicit-GEMM causal 3-D convolution for the Wan VAE decoder on channels-last bf16 activations [T, H, W, C].
//
// Replaces CausalConv3d / nn.Conv2d as the reference runs them one latent frame at a time with a two-frame cache
// (lightx2v/models/video_encoders/hf/wan/vae.py:19-44 CausalConv3d, :185-223 ResidualBlock, :70-159 Resample, :436-489
// Decoder3d.forward): processed here over the WHOLE frame sequence at once — the cache protocol is exactly a causal
// zero-padded convolution along T (tests compare against a frame-by-frame restatement of the reference loop).
//
// out[t, h, w, n] = bias[n] + sum_{tap} sum_{c} in[t + dt(tap), h + dh(tap), w + dw(tap), c] * Wt[n, tap * Cin + c]   (+ residual)
//
// sm_100a design: same persistent warp-specialised skeleton as gemm_bf16.cu (TMA producer / tcgen05 issuer / TMEM
// double-buffered accumulators / epilogue warps), with the im2col gather done by TMA itself: the A tile of one (tap, 32-channel
// chunk) is a 4-D box {32 ch, 32 w, 4 h, 1 t} fetched at coordinates shifted by the tap offset; out-of-bounds coordinates
// (spatial border, t < 0) are zero-filled by the TMA unit, which IS the zero padding of the convolution — no halo copies,
// no padded tensors.  Input / output "views" (base pointer + strides chosen by the host) express the 2x nearest-upsample
// phase decomposition and the temporal interleave of the upsample3d time_conv without extra kernels.
#include "host_util.cuh"
#include "ptx.cuh"

#include <stdlib.h>

namespace b200 {

constexpr int CONV_BLOCK_M = 128;            // voxels per tile: 32 (w) x 4 (h)
constexpr int CONV_BW = 32, CONV_BH = 4;
constexpr int CONV_THREADS = 256;
constexpr int CONV_MAX_TAPS = 27;
constexpr int CONV_STAGING_BYTES = CONV_BLOCK_M * 64 * 2;        // 16 KB (64 output channels per store chunk)

// M_SUB sub-tiles of 128 voxels (32 w x 4 h each, stacked along h) share one weight tile: with the narrow N of the VAE
// (96 / 192 output channels) a 128-voxel tile gives only ~290 tensor cycles per pipeline stage, far below the TMA round trip;
// two sub-tiles double the MMA work per byte of weight traffic and per barrier.
//
// K-chunk geometries.  These per-tap tiles are bound by the TMA ingest into the SM (law 2): ~64 B/clk/SM with
// 128-byte box rows, and measurably less with 64-byte rows (A/B at the decoder's shapes: profiles/r02_conv_tiles_ab.txt, +7 % at 96 -> 96,
// +28 % at 192 -> 192), so a stage should be made of 128-byte rows wherever the channel count allows:
//   CONV_WIDE    64-channel chunks (128-byte rows, SWIZZLE_128B), one per stage, four stages: cin % 64 == 0 (192 / 384 channels of the
//                Wan decoder, 128 / 256 / 512 of the HunyuanVideo decoder);
//   CONV_MIXED96 cin % 96 == 0 but not % 64 (the 96-channel stage of the Wan decoder): a stage is one 96-channel group of one tap,
//                held as a 64-channel SWIZZLE_128B tile plus a 32-channel SWIZZLE_64B tile (two tensor maps over the same tensor);
//   CONV_NARROW  32-channel chunks (64-byte rows, SWIZZLE_64B), three per stage: everything else (cin = 32 latent stem).
enum : int { CONV_NARROW = 0, CONV_WIDE = 1, CONV_MIXED96 = 2 };

template <int BLOCK_N, int MODE>
struct ConvCfg {
  static constexpr int kMode = MODE;
  static constexpr int kStages = MODE == CONV_WIDE ? 4 : 3;
  static constexpr int kMSub = (BLOCK_N <= 128) ? 2 : 1;
  static constexpr int kKStage = MODE == CONV_NARROW ? 96 : (MODE == CONV_WIDE ? 64 : 96);   // channels of K consumed per pipeline stage
  static constexpr int kASubBytes = CONV_BLOCK_M * kKStage * 2;        // A bytes per sub-tile per stage (24 / 16 / 24 KB)
  static constexpr int kBBytes = BLOCK_N * kKStage * 2;                // B bytes per stage
  static constexpr int kStageBytes = kMSub * kASubBytes + kBBytes;
  static constexpr int kAccCols = kMSub * BLOCK_N;                       // TMEM columns of one accumulator stage
  static constexpr int kTmemCols = (2 * kAccCols <= 128) ? 128 : (2 * kAccCols <= 256 ? 256 : 512);
  // epilogue staging buffers: two unless that would exceed the 227 KB of shared memory (BLOCK_N = 96 with two sub-tiles)
  static constexpr int kNumStaging = (kStages * kStageBytes + 2 * CONV_STAGING_BYTES + 1280 <= 232448) ? 2 : 1;
  static constexpr int kSmemBytes = kStages * kStageBytes + kNumStaging * CONV_STAGING_BYTES + 1024 + 256;
  static_assert(kSmemBytes <= 232448, "shared memory budget exceeded");
  static_assert(2 * kAccCols <= 512, "two accumulator stages must fit the 512 TMEM columns");
  static_assert(kStageBytes % 1024 == 0, "stages must keep the 1024-byte alignment of the swizzle-128B tiles");
};

struct ConvParams {
  int T, H, W;                 // output extent (tile coordinate system); the input view may be larger (pre-padded input)
  int cin, cout;               // cin multiple of the K-chunk, cout multiple of 16
  int ntaps;
  int8_t dt[CONV_MAX_TAPS], dh[CONV_MAX_TAPS], dw[CONV_MAX_TAPS];
  int tiles_w, tiles_h, num_n_blocks;
  const __nv_bfloat16* bias;       // [cout] or null
  const __nv_bfloat16* residual;   // optional, added in the epilogue; element strides below (channels contiguous)
  long long res_st, res_sh, res_sw;
  int clamp_out;                   // head conv: clamp to [-1, 1]
};

// 64-byte swizzle, K-major: 8-row groups 512 B apart.
constexpr uint32_t kDescHiSw64 = (512u >> 4) | (1u << 14) | (4u << 29);

// tmIn / tmW: the 32-channel (NARROW) or 64-channel (WIDE, MIXED96) boxes; tmIn2 / tmW2: the 32-channel boxes of MIXED96 (unused otherwise).
template <int BLOCK_N, int MODE>
__global__ void __launch_bounds__(CONV_THREADS, 1)
conv3d_igemm_kernel(const __grid_constant__ CUtensorMap tmIn, const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmIn2,
                    const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ CUtensorMap tmOut, const ConvParams p) {
  using Cfg = ConvCfg<BLOCK_N, MODE>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int CONV_STAGES = Cfg::kStages;
  // sub-tile layout of one stage.  NARROW: three [128 x 32 ch] SW64 chunks per sub-tile (8 KB each), chunk-major: chunk c, sub-tile ms at
  // (c * kMSub + ms) * 8 KB; B chunks of BLOCK_N * 64 B behind them.  WIDE: one [128 x 64 ch] SW128 tile per sub-tile (16 KB).
  // MIXED96: [64-ch SW128 tiles of all sub-tiles][32-ch SW64 tiles of all sub-tiles][B 64-ch][B 32-ch].
  constexpr int A32 = CONV_BLOCK_M * 32 * 2, A64 = CONV_BLOCK_M * 64 * 2;     // 8 KB, 16 KB
  constexpr int B32 = BLOCK_N * 32 * 2, B64 = BLOCK_N * 64 * 2;
  constexpr int kAStage = Cfg::kMSub * Cfg::kASubBytes;
  uint8_t* sStage = smem + CONV_STAGES * Cfg::kStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sStage + Cfg::kNumStaging * CONV_STAGING_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + CONV_STAGES;
  uint64_t* tmem_full_bar = bars + 2 * CONV_STAGES;
  uint64_t* tmem_empty_bar = bars + 2 * CONV_STAGES + 2;
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(bars + 2 * CONV_STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int tiles_per_frame = p.tiles_w * p.tiles_h;
  const int num_m_tiles = p.T * tiles_per_frame;
  const int num_tiles = num_m_tiles * p.num_n_blocks;
  // K loop.  NARROW: stages of up to three consecutive 32-channel chunks of the flattened (tap, channel) axis.
  // WIDE / MIXED96: one stage per (tap, group of 64 / 96 channels).
  const int groups_per_tap = MODE == CONV_NARROW ? p.cin / 32 : p.cin / Cfg::kKStage;
  const int num_chunks = p.ntaps * groups_per_tap;                         // NARROW: 32-channel chunks; otherwise: stages
  const int num_stages_k = MODE == CONV_NARROW ? (num_chunks + 2) / 3 : num_chunks;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmIn);
    prefetch_tmap(&tmW);
    prefetch_tmap(&tmOut);
  }
  if (warp == 1 && lane == 0) {
    // With two sub-tiles the MMAs are issued by TWO warps (warp 1: sub-tile 0, warp 3: sub-tile 1), each committing its own MMAs, so the
    // slot-release and accumulator-ready barriers expect one arrival per issuer.  Measured (profiles/r02_conv_two_issuers.txt): one elected
    // lane spends ~17-22 SASS instructions (ELECT / R2UR / VOTEU / address adds) per UTCHMMA, but these per-tap tiles are bound by the TMA
    // ingest (~64 B/clk/SM, DESIGN.md section 4 law 2), not by the issue rate: the second issuer is neutral here (975 vs 979 TFLOP/s at
    // 96 -> 96) and is what lifts the halo-staged tiles of conv3d_halo.cu, which share this warp layout.
    constexpr int kIssuers = Cfg::kMSub == 2 ? 2 : 1;
    for (int i = 0; i < CONV_STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], kIssuers);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full_bar[i], kIssuers);
      mbar_init(&tmem_empty_bar[i], 128);
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc(tmem_base_slot, Cfg::kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;

  // tile -> (n block, frame, h tile, w tile); n fastest so the CTAs in flight share the same input tiles through L2
  auto tile_coords = [&](int tile, int& n_blk, int& t, int& h0, int& w0) {
    n_blk = tile % p.num_n_blocks;
    int m = tile / p.num_n_blocks;
    t = m / tiles_per_frame;
    m -= t * tiles_per_frame;
    h0 = (m / p.tiles_w) * (CONV_BH * Cfg::kMSub);
    w0 = (m % p.tiles_w) * CONV_BW;
  };

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        int n_blk, t, h0, w0;
        tile_coords(tile, n_blk, t, h0, w0);
        for (int ks = 0; ks < num_stages_k; ++ks) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sA = smem + stage * Cfg::kStageBytes;
          uint8_t* sB = sA + kAStage;
          if constexpr (MODE == CONV_NARROW) {
            const int c_begin = ks * 3;
            const int c_end = min(c_begin + 3, num_chunks);
            mbar_arrive_expect_tx(&full_bar[stage], (c_end - c_begin) * (Cfg::kMSub * A32 + B32));
            for (int c = c_begin; c < c_end; ++c) {
              const int tap = c / groups_per_tap;
              const int cc = (c - tap * groups_per_tap) * 32;
#pragma unroll
              for (int ms = 0; ms < Cfg::kMSub; ++ms)
                tma_load_4d(sA + ((c - c_begin) * Cfg::kMSub + ms) * A32, &tmIn, &full_bar[stage], cc, w0 + p.dw[tap], h0 + ms * CONV_BH + p.dh[tap],
                            t + p.dt[tap]);
              tma_load_2d(sB + (c - c_begin) * B32, &tmW, &full_bar[stage], c * 32, n_blk * BLOCK_N);
            }
          } else {
            const int tap = ks / groups_per_tap;
            const int cc = (ks - tap * groups_per_tap) * Cfg::kKStage;
            mbar_arrive_expect_tx(&full_bar[stage], Cfg::kStageBytes);
#pragma unroll
            for (int ms = 0; ms < Cfg::kMSub; ++ms)
              tma_load_4d(sA + ms * A64, &tmIn, &full_bar[stage], cc, w0 + p.dw[tap], h0 + ms * CONV_BH + p.dh[tap], t + p.dt[tap]);
            tma_load_2d(sB, &tmW, &full_bar[stage], tap * p.cin + cc, n_blk * BLOCK_N);
            if constexpr (MODE == CONV_MIXED96) {
#pragma unroll
              for (int ms = 0; ms < Cfg::kMSub; ++ms)
                tma_load_4d(sA + Cfg::kMSub * A64 + ms * A32, &tmIn2, &full_bar[stage], cc + 64, w0 + p.dw[tap], h0 + ms * CONV_BH + p.dh[tap],
                            t + p.dt[tap]);
              tma_load_2d(sB + B64, &tmW2, &full_bar[stage], tap * p.cin + cc + 64, n_blk * BLOCK_N);
            }
          }
          if (++stage == CONV_STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1 || (warp == 3 && Cfg::kMSub == 2)) {
    // ===================== MMA issuers (whole warp, warp-uniform control flow): warp 1 -> sub-tile 0, warp 3 -> sub-tile 1 =====================
    const int ms_begin = (Cfg::kMSub == 2 && warp == 3) ? 1 : 0;
    const int ms_end = Cfg::kMSub == 2 ? ms_begin + 1 : 1;
    constexpr uint32_t idesc = make_idesc(FMT_BF16, FMT_BF16, CONV_BLOCK_M, BLOCK_N, 0, 0);
    const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t s_lo0 = ((smem_u32(smem) & 0x3FFFF) >> 4) | (1u << 16);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tb + acc * Cfg::kAccCols;
      for (int ks = 0; ks < num_stages_k; ++ks) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t a_lo = s_lo0 + stage * (Cfg::kStageBytes >> 4);
        const uint32_t b_lo = a_lo + (kAStage >> 4);
        if constexpr (MODE == CONV_NARROW) {
          const int nch = min(3, num_chunks - ks * 3);
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            if (c < nch) {
              for (int ms = ms_begin; ms < ms_end; ++ms) {
#pragma unroll
                for (int k = 0; k < 2; ++k) {   // K=16 MMAs per chunk: +32 B inside the swizzled row
                  const uint32_t accum = (ks | c | k) != 0 ? 1u : 0u;
                  mma_f16_ss_w(d_tmem + ms * BLOCK_N, a_lo + (c * Cfg::kMSub + ms) * (A32 >> 4) + 2 * k, kDescHiSw64, b_lo + c * (B32 >> 4) + 2 * k,
                               kDescHiSw64, idesc, accum);
                }
              }
            }
          }
        } else {
          for (int ms = ms_begin; ms < ms_end; ++ms) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              mma_f16_ss_w(d_tmem + ms * BLOCK_N, a_lo + ms * (A64 >> 4) + 2 * k, kDescHiSw128, b_lo + 2 * k, kDescHiSw128, idesc, (ks | k) != 0 ? 1u : 0u);
            if constexpr (MODE == CONV_MIXED96) {
#pragma unroll
              for (int k = 0; k < 2; ++k)
                mma_f16_ss_w(d_tmem + ms * BLOCK_N, a_lo + ((Cfg::kMSub * A64 + ms * A32) >> 4) + 2 * k, kDescHiSw64, b_lo + (B64 >> 4) + 2 * k, kDescHiSw64,
                             idesc, 1u);
            }
          }
        }
        tc_commit_w(&empty_bar[stage]);
        if (++stage == CONV_STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
      tc_commit_w(&tmem_full_bar[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const int ewarp = warp - 4;
    const int row = ewarp * 32 + lane;      // voxel inside the tile: (row / 32) -> h, (row % 32) -> w
    const int et = threadIdx.x - 128;
    int acc = 0;
    uint32_t acc_phase = 0;
    int sbuf = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      int n_blk, t, h0, w0;
      tile_coords(tile, n_blk, t, h0, w0);
      const int n0 = n_blk * BLOCK_N;
      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tc_fence_after();
#pragma unroll 1
      for (int ms = 0; ms < Cfg::kMSub; ++ms) {
      const int hs = h0 + ms * CONV_BH;                     // first image row of this sub-tile
      if (hs >= p.H) break;                                 // uniform
      const int hh = hs + row / CONV_BW, ww = w0 + row % CONV_BW;
      const bool vox_ok = hh < p.H && ww < p.W;
      const uint32_t t_row = tmem_base + acc * Cfg::kAccCols + ms * BLOCK_N + (uint32_t(ewarp * 32) << 16);
      const __nv_bfloat16* res_row =
          p.residual ? p.residual + (long long)t * p.res_st + (long long)hh * p.res_sh + (long long)ww * p.res_sw : nullptr;

#pragma unroll 1
      for (int c0 = 0; c0 < BLOCK_N; c0 += 64) {
        const int ncol0 = n0 + c0;
        if (ncol0 >= p.cout) break;
        uint32_t v[64];
        if (c0 + 64 <= BLOCK_N) {
          tmem_ld_x32(t_row + c0, v);
          tmem_ld_x32(t_row + c0 + 32, v + 32);
        } else if (c0 + 32 <= BLOCK_N) {   // BLOCK_N % 64 == 32 (e.g. 96): last half chunk
          tmem_ld_x32(t_row + c0, v);
#pragma unroll
          for (int i = 32; i < 64; ++i) v[i] = 0;
        } else {                           // BLOCK_N == 16
          tmem_ld_x16(t_row + c0, v);
#pragma unroll
          for (int i = 16; i < 64; ++i) v[i] = 0;
        }
        if (et == 0) tma_store_wait_read<Cfg::kNumStaging - 1>();
        named_bar_sync(1, 128);
        tmem_ld_wait();
        uint8_t* stg = sStage + sbuf * CONV_STAGING_BYTES;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int col = ncol0 + j * 8;
          float f[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(v[j * 8 + e]);
          if (col < p.cout) {
            if (p.bias != nullptr) {
              uint4 bv = __ldg(reinterpret_cast<const uint4*>(p.bias + col));
              const uint32_t bw[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                f[2 * e] += bf16_lo(bw[e]);
                f[2 * e + 1] += bf16_hi(bw[e]);
              }
            }
            if (res_row != nullptr && vox_ok) {
              uint4 rv = *reinterpret_cast<const uint4*>(res_row + col);
              const uint32_t rw[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                f[2 * e] += bf16_lo(rw[e]);
                f[2 * e + 1] += bf16_hi(rw[e]);
              }
            }
            if (p.clamp_out) {
#pragma unroll
              for (int e = 0; e < 8; ++e) f[e] = fminf(fmaxf(f[e], -1.0f), 1.0f);
            }
          }
          uint4 o;
          o.x = pack_bf16(f[0], f[1]);
          o.y = pack_bf16(f[2], f[3]);
          o.z = pack_bf16(f[4], f[5]);
          o.w = pack_bf16(f[6], f[7]);
          if constexpr (BLOCK_N >= 64) {
            *reinterpret_cast<uint4*>(stg + row * 128 + ((j ^ (row & 7)) << 4)) = o;   // 128B-swizzled rows of 64 channels
          } else {
            if (j < BLOCK_N / 8) *reinterpret_cast<uint4*>(stg + row * (BLOCK_N * 2) + j * 16) = o;   // dense rows (no swizzle)
          }
        }
        fence_async_smem();
        named_bar_sync(1, 128);
        if (et == 0) {
          // output box {64 ch, 32 w, 4 h, 1 t}: rows of the staging tile are (h, w)-ordered like the input box
          asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                           reinterpret_cast<uint64_t>(&tmOut)),
                       "r"(smem_u32(stg)), "r"(ncol0), "r"(w0), "r"(hs), "r"(t)
                       : "memory");
          tma_store_commit();
        }
        sbuf = (sbuf + 1) % Cfg::kNumStaging;
      }
      }   // sub-tiles
      tc_fence_before();
      mbar_arrive(&tmem_empty_bar[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    if (et == 0) tma_store_wait_all<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

template <int BLOCK_N, int MODE>
static int launch_conv(const CUtensorMap& tmIn, const CUtensorMap& tmW, const CUtensorMap& tmIn2, const CUtensorMap& tmW2, const CUtensorMap& tmOut,
                       const ConvParams& p, cudaStream_t stream) {
  using Cfg = ConvCfg<BLOCK_N, MODE>;
  auto kern = conv3d_igemm_kernel<BLOCK_N, MODE>;
  static bool attr_set = false;
  if (!attr_set) {
    B200_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    attr_set = true;
  }
  const long long num_tiles = (long long)p.T * p.tiles_w * p.tiles_h * p.num_n_blocks;
  const int grid = (int)(num_tiles < num_sms() ? num_tiles : num_sms());
  kern<<<grid, CONV_THREADS, Cfg::kSmemBytes, stream>>>(tmIn, tmW, tmIn2, tmW2, tmOut, p); note_launch();
  B200_CHECK_CUDA(cudaGetLastError());
  return B200_OK;
}

int conv3d_cl_halo(const void* in, long long in_st, long long in_sh, long long in_sw, int in_T, int in_H, int in_W, const void* wt,
                   const void* bias, void* out, long long out_st, long long out_sh, long long out_sw, const void* residual, long long res_st,
                   long long res_sh, long long res_sw, int T, int H, int W, int cin, int cout, int ntaps, const int* taps, int clamp_out,
                   cudaStream_t stream);   // conv3d_halo.cu

// in  : channels-last view  [T, H, W, cin]  bf16, element strides (in_st, in_sh, in_sw), channels contiguous
// out : channels-last view  [T, H, W, cout] bf16, element strides (out_st, out_sh, out_sw)
// wt  : [cout, ntaps * cin] bf16 row-major (tap-major K order)
// taps: ntaps x (dt, dh, dw) input offsets
// in_T / in_H / in_W : extent of the input view (>= the output extent when the producer wrote a pre-padded tensor, e.g. the
//                       replicate padding of the HunyuanVideo CausalConv3d; taps are then non-negative offsets into it).
//                       Reads outside the input view return zero.
int conv3d_cl(const void* in, long long in_st, long long in_sh, long long in_sw, int in_T, int in_H, int in_W, const void* wt,
              const void* bias, void* out, long long out_st, long long out_sh, long long out_sw, const void* residual,
              long long res_st, long long res_sh, long long res_sw, int T, int H, int W, int cin, int cout, int ntaps,
              const int* taps, int clamp_out, cudaStream_t stream) {
  B200_CHECK_ARG(in && wt && out && taps, "b200_conv3d_cl: null pointer");
  B200_CHECK_ARG(T > 0 && H > 0 && W > 0 && in_T > 0 && in_H > 0 && in_W > 0, "b200_conv3d_cl: empty extent");
  B200_CHECK_ARG(cin % 32 == 0 && cin > 0, "b200_conv3d_cl: cin (%d) must be a multiple of 32 (pad with zero channels)", cin);
  B200_CHECK_ARG(cout % 16 == 0 && cout > 0, "b200_conv3d_cl: cout (%d) must be a multiple of 16 (pad with zero filters)", cout);
  B200_CHECK_ARG(ntaps >= 1 && ntaps <= CONV_MAX_TAPS, "b200_conv3d_cl: ntaps %d out of range", ntaps);
  B200_CHECK_ARG(in_sw % 8 == 0 && in_sh % 8 == 0 && in_st % 8 == 0 && out_sw % 8 == 0 && out_sh % 8 == 0 && out_st % 8 == 0,
                 "b200_conv3d_cl: strides must be multiples of 8 elements");
  B200_CHECK_ARG(((uintptr_t)in % 16 == 0) && ((uintptr_t)out % 16 == 0) && ((uintptr_t)wt % 16 == 0),
                 "b200_conv3d_cl: pointers must be 16-byte aligned");

  // wide 3 x 3 x 3 stages: halo-staged tiles (each input voxel fetched 3 times instead of 27), conv3d_halo.cu
  {
    const int hr = conv3d_cl_halo(in, in_st, in_sh, in_sw, in_T, in_H, in_W, wt, bias, out, out_st, out_sh, out_sw, residual, res_st, res_sh, res_sw, T, H,
                                  W, cin, cout, ntaps, taps, clamp_out, stream);
    if (hr != 1) return hr;
  }

  int block_n, mode;
  const bool c64 = cin % 64 == 0, c96 = cin % 96 == 0;
  if (cout % 256 == 0 && c64) block_n = 256, mode = CONV_WIDE;
  else if (cout % 128 == 0 && c64) block_n = 128, mode = CONV_WIDE;
  else if (cout % 192 == 0) block_n = 192, mode = c64 ? CONV_WIDE : (c96 ? CONV_MIXED96 : CONV_NARROW);
  else if (cout % 96 == 0) block_n = 96, mode = c64 ? CONV_WIDE : (c96 ? CONV_MIXED96 : CONV_NARROW);
  else if (cout <= 16) block_n = 16, mode = c64 ? CONV_WIDE : (c96 ? CONV_MIXED96 : CONV_NARROW);
  else if (cout % 64 == 0) block_n = 64, mode = CONV_NARROW;
  else {
    set_last_error("b200_conv3d_cl: unsupported cout %d (need 16, or a multiple of 64 / 96 / 192)", cout);
    return B200_ERR_UNSUPPORTED;
  }
  if (get_option(OPT_CONV_NARROW) && block_n != 128 && block_n != 256) mode = CONV_NARROW;   // round-1 tiles (A/B measurements)

  const int kc = mode == CONV_NARROW ? 32 : 64;
  const CUtensorMapSwizzle swz = mode == CONV_NARROW ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B;
  CUtensorMap tmIn, tmW, tmIn2, tmW2, tmOut;
  int rc;
  {
    uint64_t dims[4] = {(uint64_t)cin, (uint64_t)in_W, (uint64_t)in_H, (uint64_t)in_T};
    uint64_t strides[3] = {(uint64_t)in_sw * 2, (uint64_t)in_sh * 2, (uint64_t)in_st * 2};
    uint32_t box[4] = {(uint32_t)kc, CONV_BW, CONV_BH, 1};
    if ((rc = encode_tmap(&tmIn, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, in, dims, strides, box, swz))) return rc;
    tmIn2 = tmIn;
    if (mode == CONV_MIXED96) {
      uint32_t box2[4] = {32, CONV_BW, CONV_BH, 1};
      if ((rc = encode_tmap(&tmIn2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, in, dims, strides, box2, CU_TENSOR_MAP_SWIZZLE_64B))) return rc;
    }
  }
  {
    uint64_t dims[2] = {(uint64_t)ntaps * cin, (uint64_t)cout};
    uint64_t strides[1] = {(uint64_t)ntaps * cin * 2};
    uint32_t box[2] = {(uint32_t)kc, (uint32_t)block_n};
    if ((rc = encode_tmap(&tmW, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, wt, dims, strides, box, swz))) return rc;
    tmW2 = tmW;
    if (mode == CONV_MIXED96) {
      uint32_t box2[2] = {32, (uint32_t)block_n};
      if ((rc = encode_tmap(&tmW2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, wt, dims, strides, box2, CU_TENSOR_MAP_SWIZZLE_64B))) return rc;
    }
  }
  {
    uint64_t dims[4] = {(uint64_t)cout, (uint64_t)W, (uint64_t)H, (uint64_t)T};
    uint64_t strides[3] = {(uint64_t)out_sw * 2, (uint64_t)out_sh * 2, (uint64_t)out_st * 2};
    uint32_t box[4] = {(uint32_t)(cout < 64 ? cout : 64), CONV_BW, CONV_BH, 1};
    // staging rows are 128 bytes (64 channels) apart and 128B-swizzled; with cout < 64 only the first `cout` channels are stored
    if ((rc = encode_tmap(&tmOut, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, out, dims, strides, box,
                          cout < 64 ? CU_TENSOR_MAP_SWIZZLE_NONE : CU_TENSOR_MAP_SWIZZLE_128B)))
      return rc;
  }

  ConvParams p;
  p.T = T;
  p.H = H;
  p.W = W;
  p.cin = cin;
  p.cout = cout;
  p.ntaps = ntaps;
  for (int i = 0; i < ntaps; ++i) {
    p.dt[i] = (int8_t)taps[3 * i + 0];
    p.dh[i] = (int8_t)taps[3 * i + 1];
    p.dw[i] = (int8_t)taps[3 * i + 2];
  }
  p.tiles_w = (W + CONV_BW - 1) / CONV_BW;
  const int msub = (block_n <= 128) ? 2 : 1;
  p.tiles_h = (H + CONV_BH * msub - 1) / (CONV_BH * msub);
  p.num_n_blocks = (cout + block_n - 1) / block_n;
  p.bias = reinterpret_cast<const __nv_bfloat16*>(bias);
  p.residual = reinterpret_cast<const __nv_bfloat16*>(residual);
  p.res_st = res_st;
  p.res_sh = res_sh;
  p.res_sw = res_sw;
  p.clamp_out = clamp_out;
#define B200_CONV_CASE(BN, MD) \
  if (block_n == BN && mode == MD) return launch_conv<BN, MD>(tmIn, tmW, tmIn2, tmW2, tmOut, p, stream)
  B200_CONV_CASE(256, CONV_WIDE);
  B200_CONV_CASE(128, CONV_WIDE);
  B200_CONV_CASE(192, CONV_WIDE);
  B200_CONV_CASE(192, CONV_MIXED96);
  B200_CONV_CASE(192, CONV_NARROW);
  B200_CONV_CASE(96, CONV_WIDE);
  B200_CONV_CASE(96, CONV_MIXED96);
  B200_CONV_CASE(96, CONV_NARROW);
  B200_CONV_CASE(64, CONV_NARROW);
  B200_CONV_CASE(16, CONV_WIDE);
  B200_CONV_CASE(16, CONV_MIXED96);
  B200_CONV_CASE(16, CONV_NARROW);
#undef B200_CONV_CASE
  set_last_error("b200_conv3d_cl: no tile configuration for cout %d / cin %d", cout, cin);
  return B200_ERR_UNSUPPORTED;
}

}  // namespace b200
