// Halo-staged implicit-GEMM 3x3x3 convolution for the wide stages of the Wan VAE (W >= 512: the 96-channel 720 x 1280 stage incl. its
// 16-wide head, and the 192-channel 360 x 640 stage - 62 % of the decode).
//
// Why: conv3d.cu fetches the A operand of EVERY tap separately (a TMA box shifted by the tap offset), i.e. each input voxel crosses
// L2 -> shared memory 27 times per output tile; ncu (profiles/r01_misc_kernels_ncu_summary.txt, profiles/r02_*conv*) shows those tiles at
// 42-58 % tensor pipe with the L2 / TMA path saturated (lts throughput 56 % at 42 % tensor pipe, against 71 % at 88 % for the 128-wide
// tiles).  Here a CTA stages, per (input frame, 64-channel chunk), ONE halo tile of 4 image rows x 136 voxels (rows h0-1 .. h0+2, columns
// w0-1 .. w0+134) and takes all nine spatial taps of both output rows h0, h0+1 from it as ROW-SHIFTED views: the UMMA shared-memory
// descriptor of tap (dh, dw), output row ms starts (ms + dh + 1) * 136 + (dw + 1) rows into the tile.  A voxel is fetched 3 times (once per
// temporal tap) instead of 27, and 128-byte rows are used throughout (the 96-channel tensors are read as a 64-channel chunk plus a second
// 64-channel box whose upper 32 channels lie outside the tensor and are zero-filled by the TMA unit; only its two valid K-steps are issued).
//
// Row-shifted descriptors: measured on B200 with csrc/probe.cu / tools/probe_rowshift.py (profiles/r02_probe_rowshift.txt): a K-major
// swizzled operand whose descriptor start address is moved by whole rows (128-byte rows / SWIZZLE_128B and 64-byte rows / SWIZZLE_64B,
// shifts 0..8) is read correctly with base_offset = 0 - the tensor core, like the TMA unit, derives the swizzle phase from ABSOLUTE
// shared-memory address bits - and incorrectly with base_offset = row % 8.  So a row-shifted view is just "start address + rows * 128".
//
// Work per pipeline slot (N = 96): A halo 68 KB feeds 9 taps x 2 rows x 4 K-steps = 72 MMAs; the nine 12 KB weight tiles stream through
// their own ring: 2.5 KB of TMA ingest per MMA against 5.6 KB for conv3d.cu's MIXED96 tiles, which sit at the ~64 B/clk/SM ingest ceiling
// (law 2).  With the ingest out of the way the MMA warp's issue rate became the limit, hence the two issuing warps
// and the runs of four MMAs under one election below (law 3); what remains at N = 96 is the operand-read floor of law 1 (56 clocks per MMA).
//
// Warp roles: warp 0 loads halo tiles, warp 2 (after allocating TMEM) loads weight tiles - independent rings, so a stalled weight slot
// never delays the next halo prefetch -, warps 1 and 3 issue tcgen05.mma for output row 0 / row 1 (two issuers, each issuing a tap's K-steps as one run:
// ptx.cuh mma_f16_ss_w4 / _w2), warps 4-7 run the epilogue.
#include "host_util.cuh"
#include "ptx.cuh"

#include <stdlib.h>

namespace b200 {

constexpr int HALO_W = 136;                     // voxels per staged image row (128 + 2, rounded up to a multiple of 8 rows)
constexpr int HALO_ROWS = 4;                    // image rows per halo tile: two output rows + one above + one below
constexpr int HALO_A_BYTES = HALO_ROWS * HALO_W * 128;      // 69 632 B = 68 KB (1024-byte multiple)
constexpr int HALO_A_SLOTS = 2;
constexpr int HALO_THREADS = 256;
constexpr int HALO_STAGING_BYTES = 128 * 64 * 2;            // 16 KB: 128 voxels x 64 output channels

// HaloParams::base_offset_mode: 0 (default, correct) leaves the descriptor's base_offset field zero; 1 sets it to start row % 8 (kept as a
// switch for the probe experiment, option "halo_base_offset")

template <int BLOCK_N>
struct HaloCfg {
  static constexpr int kBBytes = BLOCK_N * 128;                                   // one (tap, 64-channel chunk) weight tile
  static constexpr int kBSlotBytes = (kBBytes + 1023) / 1024 * 1024;
  static constexpr int kBSlots = BLOCK_N <= 96 ? 4 : 3;
  static constexpr int kAccCols = 2 * BLOCK_N;                                    // two output rows
  static constexpr int kAccStages = (2 * kAccCols <= 512) ? 2 : 1;
  static constexpr int kTmemCols = (kAccStages * kAccCols <= 64) ? 64 : (kAccStages * kAccCols <= 128 ? 128 : (kAccStages * kAccCols <= 256 ? 256 : 512));
  static constexpr int kNumStaging = (HALO_A_SLOTS * HALO_A_BYTES + kBSlots * kBSlotBytes + 2 * HALO_STAGING_BYTES + 1280 <= 232448) ? 2 : 1;
  static constexpr int kSmemBytes = HALO_A_SLOTS * HALO_A_BYTES + kBSlots * kBSlotBytes + kNumStaging * HALO_STAGING_BYTES + 1024 + 256;
  static_assert(kSmemBytes <= 232448, "shared memory budget exceeded");
};

struct HaloParams {
  int T, H, W;                 // output extent
  int cin, cout;
  int nchunks;                 // 64-channel chunks per tap (cin = 96: 2, the second with 2 valid K-steps)
  int ksteps_last;             // K-steps (of 16 channels) of the last chunk: 4, or 2 when cin % 64 == 32
  int dt[3];                   // input-frame offset of temporal tap 0, 1, 2
  int tiles_w, tiles_h, num_n_blocks;
  const __nv_bfloat16* bias;
  const __nv_bfloat16* residual;
  long long res_st, res_sh, res_sw;
  int clamp_out;
  int base_offset_mode;
};

template <int BLOCK_N>
__global__ void __launch_bounds__(HALO_THREADS, 1)
conv3d_halo_kernel(const __grid_constant__ CUtensorMap tmIn, const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmOut,
                   const HaloParams p) {
  using Cfg = HaloCfg<BLOCK_N>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = sA + HALO_A_SLOTS * HALO_A_BYTES;
  uint8_t* sStage = sB + Cfg::kBSlots * Cfg::kBSlotBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sStage + Cfg::kNumStaging * HALO_STAGING_BYTES);
  uint64_t* a_full = bars;                                  // [2]
  uint64_t* a_empty = a_full + HALO_A_SLOTS;                // [2]
  uint64_t* b_full = a_empty + HALO_A_SLOTS;                // [kBSlots]
  uint64_t* b_empty = b_full + Cfg::kBSlots;                // [kBSlots]
  uint64_t* tmem_full_bar = b_empty + Cfg::kBSlots;         // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;             // [2]
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int tiles_per_frame = p.tiles_w * p.tiles_h;
  const int num_tiles = p.T * tiles_per_frame * p.num_n_blocks;
  const int slots_per_tile = 3 * p.nchunks;                 // (temporal tap, chunk) halo slots per output tile

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmIn);
    prefetch_tmap(&tmW);
    prefetch_tmap(&tmOut);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < HALO_A_SLOTS; ++i) {
      mbar_init(&a_full[i], 1);
      mbar_init(&a_empty[i], 2);          // one commit per MMA-issuing warp
    }
    for (int i = 0; i < Cfg::kBSlots; ++i) {
      mbar_init(&b_full[i], 1);
      mbar_init(&b_empty[i], 2);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full_bar[i], 2);
      mbar_init(&tmem_empty_bar[i], 128);
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc(tmem_base_slot, Cfg::kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;

  // tile -> (n block, frame, row pair, 128-column block); n fastest so CTAs in flight share input rows through L2
  auto tile_coords = [&](int tile, int& n_blk, int& t, int& h0, int& w0) {
    n_blk = tile % p.num_n_blocks;
    int m = tile / p.num_n_blocks;
    t = m / tiles_per_frame;
    m -= t * tiles_per_frame;
    h0 = (m / p.tiles_w) * 2;
    w0 = (m % p.tiles_w) * 128;
  };

  if (warp == 0) {
    // ===================== halo producer =====================
    if (lane == 0) {
      int slot = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        int n_blk, t, h0, w0;
        tile_coords(tile, n_blk, t, h0, w0);
        for (int s = 0; s < slots_per_tile; ++s) {
          const int kt = s / p.nchunks, ch = s - kt * p.nchunks;
          mbar_wait(&a_empty[slot], phase ^ 1);
          mbar_arrive_expect_tx(&a_full[slot], HALO_A_BYTES);
          // box {64 ch, 136 w, 4 h, 1 t}: columns / rows / frames (and, for cin = 96, channels) outside the tensor are zero-filled
          tma_load_4d(sA + slot * HALO_A_BYTES, &tmIn, &a_full[slot], ch * 64, w0 - 1, h0 - 1, t + p.dt[kt]);
          if (++slot == HALO_A_SLOTS) {
            slot = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 2) {
    // ===================== weight producer (the TMEM-allocating warp) =====================
    if (lane == 0) {
      int slot = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        int n_blk, t, h0, w0;
        tile_coords(tile, n_blk, t, h0, w0);
        for (int s = 0; s < slots_per_tile; ++s) {
          const int kt = s / p.nchunks, ch = s - kt * p.nchunks;
          for (int sp = 0; sp < 9; ++sp) {
            mbar_wait(&b_empty[slot], phase ^ 1);
            mbar_arrive_expect_tx(&b_full[slot], Cfg::kBBytes);
            tma_load_2d(sB + slot * Cfg::kBSlotBytes, &tmW, &b_full[slot], (kt * 9 + sp) * p.cin + ch * 64, n_blk * BLOCK_N);
            if (++slot == Cfg::kBSlots) {
              slot = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1 || warp == 3) {
    // ===================== MMA issuers (whole warp, warp-uniform control flow): warp 1 -> output row 0, warp 3 -> output row 1 =====================
    const int ms = warp == 3 ? 1 : 0;
    constexpr uint32_t idesc = make_idesc(FMT_BF16, FMT_BF16, 128, BLOCK_N, 0, 0);
    const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t a_base = smem_u32(sA), b_lo0 = desc_lo_kmajor(smem_u32(sB));
    int aslot = 0, bslot = 0;
    uint32_t aphase = 0, bphase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tb + acc * Cfg::kAccCols;
      for (int s = 0; s < slots_per_tile; ++s) {
        const int ch = s % p.nchunks;
        const int ksteps = (ch == p.nchunks - 1) ? p.ksteps_last : 4;
        mbar_wait(&a_full[aslot], aphase);
        tc_fence_after();
        // descriptor of this issuer's output row at tap (dh, dw) = (-1, -1); tap (dh, dw) starts (dh * HALO_W + dw) rows = that many
        // 128-byte units further on (the low word counts 16-byte units).  The 9 taps are unrolled so the offsets - and base_offset =
        // row % 8 = dw, as HALO_W % 8 == 0 - are immediates: the issue loop is what bounded this kernel (file header).
        const uint32_t a_row0 = desc_lo_kmajor(a_base + aslot * HALO_A_BYTES + (uint32_t)ms * HALO_W * 128);
        const bool full_chunk = ksteps == 4;
#pragma unroll
        for (int sp = 0; sp < 9; ++sp) {
          const int dh = sp / 3, dw = sp - dh * 3;                // 0..2 (= offset + 1)
          mbar_wait(&b_full[bslot], bphase);
          tc_fence_after();
          const uint32_t b_lo = b_lo0 + (uint32_t)bslot * (Cfg::kBSlotBytes >> 4);
          const uint32_t a_lo = a_row0 + (uint32_t)(dh * HALO_W + dw) * 8;
          const uint32_t a_hi = kDescHiSw128 | (p.base_offset_mode ? ((uint32_t)dw << 17) : 0u);
          const uint32_t accumulate = sp != 0 ? 1u : (s != 0 ? 1u : 0u);
          if (full_chunk) mma_f16_ss_w4(d_tmem + ms * BLOCK_N, a_lo, a_hi, b_lo, kDescHiSw128, idesc, accumulate);
          else mma_f16_ss_w2(d_tmem + ms * BLOCK_N, a_lo, a_hi, b_lo, kDescHiSw128, idesc, accumulate);
          tc_commit_w(&b_empty[bslot]);
          if (++bslot == Cfg::kBSlots) {
            bslot = 0;
            bphase ^= 1;
          }
        }
        tc_commit_w(&a_empty[aslot]);
        if (++aslot == HALO_A_SLOTS) {
          aslot = 0;
          aphase ^= 1;
        }
      }
      tc_commit_w(&tmem_full_bar[acc]);
      if (Cfg::kAccStages == 2) {
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      } else {
        acc_phase ^= 1;
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const int ewarp = warp - 4;
    const int row = ewarp * 32 + lane;      // voxel inside the 128-column tile
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
      for (int ms = 0; ms < 2; ++ms) {
        const int hh = h0 + ms, ww = w0 + row;
        if (hh >= p.H) break;                                 // uniform
        const bool vox_ok = ww < p.W;
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
          } else if (c0 + 32 <= BLOCK_N) {   // BLOCK_N % 64 == 32 (96): last half chunk
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
          uint8_t* stg = sStage + sbuf * HALO_STAGING_BYTES;
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
            // output box {64 ch, 128 w, 1 h, 1 t}
            asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                             reinterpret_cast<uint64_t>(&tmOut)),
                         "r"(smem_u32(stg)), "r"(ncol0), "r"(w0), "r"(hh), "r"(t)
                         : "memory");
            tma_store_commit();
          }
          sbuf = (sbuf + 1) % Cfg::kNumStaging;
        }
      }   // output rows
      tc_fence_before();
      mbar_arrive(&tmem_empty_bar[acc]);
      if (Cfg::kAccStages == 2) {
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      } else {
        acc_phase ^= 1;
      }
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

template <int BLOCK_N>
static int launch_halo(const CUtensorMap& tmIn, const CUtensorMap& tmW, const CUtensorMap& tmOut, const HaloParams& p, cudaStream_t stream) {
  using Cfg = HaloCfg<BLOCK_N>;
  auto kern = conv3d_halo_kernel<BLOCK_N>;
  static bool attr_set = false;
  if (!attr_set) {
    B200_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    attr_set = true;
  }
  const long long num_tiles = (long long)p.T * p.tiles_w * p.tiles_h * p.num_n_blocks;
  const int grid = (int)(num_tiles < num_sms() ? num_tiles : num_sms());
  kern<<<grid, HALO_THREADS, Cfg::kSmemBytes, stream>>>(tmIn, tmW, tmOut, p); note_launch();
  B200_CHECK_CUDA(cudaGetLastError());
  return B200_OK;
}

// Eligibility + launch.  Returns 1 when the shape is not handled here (the caller falls through to conv3d.cu's per-tap tiles),
// B200_OK / an error code otherwise.  Same argument meaning as conv3d_cl.
int conv3d_cl_halo(const void* in, long long in_st, long long in_sh, long long in_sw, int in_T, int in_H, int in_W, const void* wt,
                   const void* bias, void* out, long long out_st, long long out_sh, long long out_sw, const void* residual, long long res_st,
                   long long res_sh, long long res_sw, int T, int H, int W, int cin, int cout, int ntaps, const int* taps, int clamp_out,
                   cudaStream_t stream) {
  const int bo_mode = get_option(OPT_HALO_BASE_OFFSET);
  if (!get_option(OPT_CONV_HALO) || ntaps != 27 || W < 512) return 1;
  if (!(cin % 64 == 0 || cin % 64 == 32) || cin < 64) return 1;
  // Measured on B200 (profiles/r02_umma_rate_probe.txt, r02_conv_two_issuers.txt, r02_conv_halo_ab.txt):
  //  * an SS-mode 128 x N x 16 MMA costs max(N / 2, (4096 + 32 N) / 128) clocks - the tensor core reads its operands from shared memory
  //    at exactly 128 B/clk - so N = 96 tops out at 85.6 % of the pipe, N = 16 at 20 %; neither a second issuing warp nor concurrent
  //    shared-memory stores move that floor;
  //  * every im2col tile of conv3d.cu sits at ~64 B/clk/SM of TMA ingest (N = 96: 67.6 KB per 12 MMAs; 128 x 192: 40 KB per 4; 256 x 128:
  //    48 KB per 8), i.e. the L2 -> SM path, not the tensor pipe, bounds them at 45-68 %;
  //  * the halo tile cuts the ingest ~4x and leaves the MMA warp issue-bound (86 % of its samples in ELECT / R2UR / address code), which
  //    the second issuing warp relieves: 192 -> 192 1349 -> 1560 TFLOP/s (93 % of the measured cuBLAS burst peak), 96 -> 96 975 -> 1105,
  //    192 -> 96 1214 -> 1301;
  //  * issuing each (tap, chunk)'s K-steps as ONE run under one election (mma_f16_ss_w4 / _w2: ~7 instead of ~22 SASS instructions per
  //    MMA) lifts them again - 96 -> 96 1261, 192 -> 96 1487, 192 -> 192 1571 TFLOP/s - and makes the halo tiles the faster choice for
  //    the 96 -> 16 head too (302 vs 209 TFLOP/s; its floor is the 39-clock A read).  profiles/r02_mma_runs_perf.txt.
  int block_n;
  if (cout % 192 == 0) block_n = 192;
  else if (cout % 96 == 0) block_n = 96;
  else if (cout == 16) block_n = 16;
  else return 1;
  // taps must be the full 3 x 3 x 3 stencil in (temporal, dh, dw) order with dh, dw in {-1, 0, 1}
  for (int kt = 0; kt < 3; ++kt)
    for (int sp = 0; sp < 9; ++sp) {
      const int* tp = taps + 3 * (kt * 9 + sp);
      if (tp[0] != taps[3 * kt * 9] || tp[1] != sp / 3 - 1 || tp[2] != sp % 3 - 1) return 1;
    }

  CUtensorMap tmIn, tmW, tmOut;
  int rc;
  {
    uint64_t dims[4] = {(uint64_t)cin, (uint64_t)in_W, (uint64_t)in_H, (uint64_t)in_T};
    uint64_t strides[3] = {(uint64_t)in_sw * 2, (uint64_t)in_sh * 2, (uint64_t)in_st * 2};
    uint32_t box[4] = {64, HALO_W, HALO_ROWS, 1};
    if ((rc = encode_tmap(&tmIn, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, in, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  }
  {
    uint64_t dims[2] = {(uint64_t)ntaps * cin, (uint64_t)cout};
    uint64_t strides[1] = {(uint64_t)ntaps * cin * 2};
    uint32_t box[2] = {64, (uint32_t)block_n};
    if ((rc = encode_tmap(&tmW, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, wt, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  }
  {
    uint64_t dims[4] = {(uint64_t)cout, (uint64_t)W, (uint64_t)H, (uint64_t)T};
    uint64_t strides[3] = {(uint64_t)out_sw * 2, (uint64_t)out_sh * 2, (uint64_t)out_st * 2};
    uint32_t box[4] = {(uint32_t)(cout < 64 ? cout : 64), 128, 1, 1};
    if ((rc = encode_tmap(&tmOut, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, out, dims, strides, box,
                          cout < 64 ? CU_TENSOR_MAP_SWIZZLE_NONE : CU_TENSOR_MAP_SWIZZLE_128B)))
      return rc;
  }
  HaloParams p;
  p.T = T;
  p.H = H;
  p.W = W;
  p.cin = cin;
  p.cout = cout;
  p.nchunks = (cin + 63) / 64;
  p.ksteps_last = (cin % 64 == 32) ? 2 : 4;
  for (int kt = 0; kt < 3; ++kt) p.dt[kt] = taps[3 * kt * 9];
  p.tiles_w = (W + 127) / 128;
  p.tiles_h = (H + 1) / 2;
  p.num_n_blocks = (cout + block_n - 1) / block_n;
  p.bias = reinterpret_cast<const __nv_bfloat16*>(bias);
  p.residual = reinterpret_cast<const __nv_bfloat16*>(residual);
  p.res_st = res_st;
  p.res_sh = res_sh;
  p.res_sw = res_sw;
  p.clamp_out = clamp_out;
  p.base_offset_mode = bo_mode;
  switch (block_n) {
    case 192: return launch_halo<192>(tmIn, tmW, tmOut, p, stream);
    case 96: return launch_halo<96>(tmIn, tmW, tmOut, p, stream);
    default: return launch_halo<16>(tmIn, tmW, tmOut, p, stream);
  }
}

}  // namespace b200
