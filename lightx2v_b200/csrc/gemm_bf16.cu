// bf16 NT GEMM for the DiT linears:  C[M,N] = epilogue(A[M,K] * B[N,K]^T + bias)
//
// Replaces MMWeight.apply = torch.addmm(bias, x, W.t())  (reference: lightx2v/common/ops/mm/mm_weight.py:81-88;
// B is the checkpoint's [N,K] row-major weight, i.e. an NT GEMM) and absorbs the
// elementwise passes that follow it in WanTransformerInfer (transformer_infer.py:402,468,492,503).
//
// sm_100a design: persistent warp-specialised kernel, one CTA per SM.
//   warp 0      : TMA producer  (cp.async.bulk.tensor, 128B-swizzled K-major tiles, kStages-deep mbarrier ring)
//   warp 1      : MMA issuer    (tcgen05.mma cta_group::1 kind::f16, M=128, N=BLOCK_N, K=16; fp32 accumulators in TMEM,
//                                two accumulator stages so the epilogue of tile i overlaps the mainloop of tile i+1)
//   warp 2      : TMEM allocator
//   warps 4..7  : epilogue      (tcgen05.ld -> registers -> bias / GELU / gate*y+residual -> swizzled smem -> TMA store)
#include "host_util.cuh"
#include "ptx.cuh"

#include <stdlib.h>

namespace b200 {

enum GemmEpilogue : int {
  EPI_BIAS = 0,           // C = bf16(acc + bias)
  EPI_BIAS_GELU = 1,      // C = bf16(gelu_tanh(bf16(acc + bias)))
  EPI_GATE_RESIDUAL = 2,  // C = bf16(C + bf16(bf16(acc + bias) * gate[n]))      (in place on C)
  EPI_RESIDUAL = 3,       // C = bf16(C + bf16(acc + bias))                        (in place on C)
};

constexpr int GEMM_BLOCK_M = 128;
constexpr int GEMM_BLOCK_K = 64;   // 64 bf16 = 128 bytes = one swizzle-128B row
constexpr int GEMM_UMMA_K = 16;
constexpr int GEMM_THREADS = 256;
constexpr int GEMM_EPI_THREADS = 128;
constexpr int GEMM_EPI_CHUNK = 64;  // columns per epilogue store chunk (128 bytes of bf16)

enum GemmMode : int { MODE_BF16 = 0, MODE_FP8 = 1, MODE_NVFP4 = 2 };

// A K-block is always 128 bytes of K per row (one swizzle-128B row): 64 bf16 / 128 e4m3 / 256 e2m1 elements.
// MODE_NVFP4 adds per stage the ue4m3 scale factors of the block: 4 "SF tiles" of 512 B (128 rows x 4 K-groups of 16) per 128 rows,
// and keeps them in TMEM next to the accumulators - with BLOCK_N = 256 that leaves room for ONE accumulator stage only.
template <int BLOCK_N, int MODE>
struct GemmCfg {
  static constexpr bool kF4 = MODE == MODE_NVFP4;
  static constexpr int kAccStages = (kF4 && BLOCK_N == 256) ? 1 : 2;
  static constexpr int kStages = kF4 ? ((BLOCK_N == 256) ? 3 : 5) : ((BLOCK_N == 256) ? 4 : 6);
  static constexpr int kABytes = GEMM_BLOCK_M * GEMM_BLOCK_K * 2;       // 16 KB
  static constexpr int kBBytes = BLOCK_N * GEMM_BLOCK_K * 2;            // 32 / 16 KB
  static constexpr int kSfaBytes = kF4 ? 2048 : 0;
  static constexpr int kSfbBytes = kF4 ? (BLOCK_N / 128) * 2048 : 0;
  static constexpr int kStageBytes = kABytes + kBBytes + kSfaBytes + kSfbBytes;
  static constexpr int kStagingBytes = GEMM_BLOCK_M * GEMM_EPI_CHUNK * 2;  // 16 KB, x2 buffers
  static constexpr int kSfaCol = kAccStages * BLOCK_N;                  // TMEM columns of the scale factors (MODE_NVFP4)
  static constexpr int kSfbCol = kSfaCol + 16;
  static constexpr int kTmemCols = kF4 ? 512 : 2 * BLOCK_N;             // accumulator stages (+ scale factors)
  static constexpr int kSmemBytes = kStages * kStageBytes + 2 * kStagingBytes + 1024 /*align*/ + 256 /*barriers*/;
  static_assert(kSmemBytes <= 232448, "shared memory budget exceeded");
  static_assert(!kF4 || kSfbCol + 4 * (BLOCK_N / 32) <= 512, "TMEM budget exceeded");
};

struct GemmParams {
  int M, N, K;
  int num_m_blocks, num_n_blocks, num_k_blocks;
  const __nv_bfloat16* bias;   // [N] or null
  const __nv_bfloat16* gate;   // [N] (EPI_GATE_RESIDUAL)
  __nv_bfloat16* C;            // residual source (in-place epilogues)
  long long ldc;
  int group_m;                 // M-blocks per rasterisation group
  const float* scale_a;        // fp8 path: per-row (token) activation scale [M]
  const float* scale_b;        // fp8 path: per-column (out-channel) weight scale [N]
  const float* alpha;          // nvfp4 path: device scalar 1 / (global_scale_a * global_scale_b)
};

// Tile rasterisation: groups of kGroupM m-blocks, n fastest inside a group-row sweep, so the ~148 tiles in flight
// touch ~16 A panels and ~10 B panels (L2-resident) instead of 148 A panels.
__device__ __forceinline__ void tile_coords(int tile, const GemmParams& p, int& m_blk, int& n_blk) {
  const int kGroupM = p.group_m;
  const int tiles_per_group = kGroupM * p.num_n_blocks;
  const int group = tile / tiles_per_group;
  const int first_m = group * kGroupM;
  const int group_m = min(kGroupM, p.num_m_blocks - first_m);
  const int in_group = tile - group * tiles_per_group;
  m_blk = first_m + in_group % group_m;
  n_blk = in_group / group_m;
}

__device__ __forceinline__ float gelu_tanh_f(float x) {
  // torch.nn.functional.gelu(approximate="tanh")
  const float k0 = 0.7978845608028654f, k1 = 0.044715f;
  float inner = k0 * (x + k1 * x * x * x);
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(inner));
  return 0.5f * x * (1.0f + t);
}

// MODE_BF16 : bf16 x bf16 (kind::f16, 64 elements per 128-byte K-block)
// MODE_FP8  : e4m3 x e4m3 (kind::f8f6f4, 128 elements per 128-byte K-block), epilogue applies the per-token and
//             per-channel scales:  y = sa[m] * (sb[n] * acc) + bias   (vLLM cutlass_scaled_mm semantics, mm_weight.py:304-319)
// MODE_NVFP4: e2m1 x e2m1 with ue4m3 scales per 16 elements (kind::mxf4nvf4.block_scale, 256 elements per K-block; the scale
//             factors ride TMA -> smem -> tcgen05.cp -> TMEM), epilogue y = alpha * acc + bias
//             (lightx2v_kernel cutlass_scaled_fp4_mm semantics, lightx2v_kernel/python/lightx2v_kernel/gemm.py:4-8)
template <int BLOCK_N, int EPI, int MODE>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmSFA,
                 const __grid_constant__ CUtensorMap tmSFB, const GemmParams p) {
  using Cfg = GemmCfg<BLOCK_N, MODE>;
  constexpr int kStages = Cfg::kStages;
  constexpr int kAccStages = Cfg::kAccStages;
  constexpr bool kFp8 = MODE == MODE_FP8;
  constexpr bool kF4 = MODE == MODE_NVFP4;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;                                   // kStages x 16 KB
  uint8_t* sB = smem + kStages * Cfg::kABytes;          // kStages x kBBytes
  uint8_t* sSFA = sB + kStages * Cfg::kBBytes;          // kStages x 2 KB   (MODE_NVFP4)
  uint8_t* sSFB = sSFA + kStages * Cfg::kSfaBytes;      // kStages x 2 / 4 KB
  uint8_t* sStage = smem + kStages * Cfg::kStageBytes;  // 2 x 16 KB epilogue staging
  uint64_t* bars = reinterpret_cast<uint64_t*>(sStage + 2 * Cfg::kStagingBytes);
  uint64_t* full_bar = bars;                  // [kStages]
  uint64_t* empty_bar = bars + kStages;       // [kStages]
  uint64_t* tmem_full_bar = bars + 2 * kStages;       // [2]
  uint64_t* tmem_empty_bar = bars + 2 * kStages + 2;  // [2]
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_tiles = p.num_m_blocks * p.num_n_blocks;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    prefetch_tmap(&tmC);
    if constexpr (kF4) {
      prefetch_tmap(&tmSFA);
      prefetch_tmap(&tmSFB);
    }
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < kAccStages; ++i) {
      mbar_init(&tmem_full_bar[i], 1);
      mbar_init(&tmem_empty_bar[i], GEMM_EPI_THREADS);
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc(tmem_base_slot, Cfg::kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        int m_blk, n_blk;
        tile_coords(tile, p, m_blk, n_blk);
        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full_bar[stage], Cfg::kStageBytes);
          constexpr int kElemsPerKBlock = (kFp8 || kF4) ? 128 : 64;   // tensor-map elements (bytes for fp8 / packed fp4) per 128 B
          tma_load_2d(sA + stage * Cfg::kABytes, &tmA, &full_bar[stage], kb * kElemsPerKBlock, m_blk * GEMM_BLOCK_M);
          tma_load_2d(sB + stage * Cfg::kBBytes, &tmB, &full_bar[stage], kb * kElemsPerKBlock, n_blk * BLOCK_N);
          if constexpr (kF4) {   // scale factors: rows = 128-row tiles, 512 B per 64 elements of K
            tma_load_3d(sSFA + stage * Cfg::kSfaBytes, &tmSFA, &full_bar[stage], 0, kb * 4, m_blk);
            tma_load_3d(sSFB + stage * Cfg::kSfbBytes, &tmSFB, &full_bar[stage], 0, kb * 4, n_blk * (BLOCK_N / 128));
          }
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // Whole warp, warp-uniform control flow (descriptors and loop state in uniform registers); one elected lane per
    // tcgen05 instruction (the *_w wrappers in ptx.cuh).
    constexpr uint32_t idesc = kF4 ? make_idesc_f4(GEMM_BLOCK_M, BLOCK_N)
                                   : (kFp8 ? make_idesc(FMT_E4M3, FMT_E4M3, GEMM_BLOCK_M, BLOCK_N, 0, 0)
                                           : make_idesc(FMT_BF16, FMT_BF16, GEMM_BLOCK_M, BLOCK_N, 0, 0));
    const uint32_t sfa_lo0 = (smem_u32(sSFA) & 0x3FFFF) >> 4;
    const uint32_t sfb_lo0 = (smem_u32(sSFB) & 0x3FFFF) >> 4;
    const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t a_lo0 = desc_lo_kmajor(smem_u32(sA));
    const uint32_t b_lo0 = desc_lo_kmajor(smem_u32(sB));
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tb + acc * BLOCK_N;
      for (int kb = 0; kb < p.num_k_blocks; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t a_lo = a_lo0 + stage * (Cfg::kABytes >> 4);
        const uint32_t b_lo = b_lo0 + stage * (Cfg::kBBytes >> 4);
        if constexpr (kF4) {
          // scale factors of this K-block: smem -> TMEM, 512 B (one MMA's worth for 128 rows) per copy.  tcgen05.cp and
          // tcgen05.mma execute in issue order, so the copies need no barrier of their own and the single TMEM copy is safe.
          const uint32_t sfa_lo = sfa_lo0 + stage * (Cfg::kSfaBytes >> 4);
          const uint32_t sfb_lo = sfb_lo0 + stage * (Cfg::kSfbBytes >> 4);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            tmem_cp_32x128b_w(tb + Cfg::kSfaCol + 4 * k, sfa_lo + k * (512 >> 4), kDescHiSfNoSwz);
#pragma unroll
            for (int t = 0; t < BLOCK_N / 128; ++t)
              tmem_cp_32x128b_w(tb + Cfg::kSfbCol + (BLOCK_N / 32) * k + 4 * t, sfb_lo + t * (2048 >> 4) + k * (512 >> 4), kDescHiSfNoSwz);
          }
        }
        // one K-block = four MMAs, +32 bytes along K inside the 128B swizzle row = +2 in the (addr >> 4) field per step; the bf16 / fp8 runs are
        // issued under one election (ptx.cuh: ~7 instead of ~22 SASS instructions per MMA in this warp)
        static_assert(GEMM_BLOCK_K / GEMM_UMMA_K == 4, "MMA runs are four K-steps");
        if constexpr (kF4) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            mma_f4_bs_w(d_tmem, a_lo + 2 * k, kDescHiSw128, b_lo + 2 * k, kDescHiSw128, idesc, tb + Cfg::kSfaCol + 4 * k,
                        tb + Cfg::kSfbCol + (BLOCK_N / 32) * k, (kb | k) != 0 ? 1u : 0u);
        } else if constexpr (kFp8) {
          mma_f8_ss_w4(d_tmem, a_lo, kDescHiSw128, b_lo, kDescHiSw128, idesc, kb != 0 ? 1u : 0u);
        } else {
          mma_f16_ss_w4(d_tmem, a_lo, kDescHiSw128, b_lo, kDescHiSw128, idesc, kb != 0 ? 1u : 0u);
        }
        tc_commit_w(&empty_bar[stage]);  // frees the smem slot once these MMAs have read it
        if (++stage == kStages) {
          stage = 0;
          phase ^= 1;
        }
      }
      tc_commit_w(&tmem_full_bar[acc]);  // accumulator complete -> epilogue
      if (++acc == kAccStages) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const int ewarp = warp - 4;               // == warp % 4 -> TMEM lane group
    const int row = ewarp * 32 + lane;        // row inside the 128-row tile
    const int et = threadIdx.x - 128;         // 0..127
    int acc = 0;
    uint32_t acc_phase = 0;
    int sbuf = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      int m_blk, n_blk;
      tile_coords(tile, p, m_blk, n_blk);
      const int m0 = m_blk * GEMM_BLOCK_M;
      const int n0 = n_blk * BLOCK_N;
      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + acc * BLOCK_N + (uint32_t(ewarp * 32) << 16);
      const bool row_ok = (m0 + row) < p.M;
      float sa = 1.0f;
      if constexpr (kFp8) sa = row_ok ? __ldg(p.scale_a + m0 + row) : 0.0f;
      if constexpr (kF4) sa = __ldg(p.alpha);

#pragma unroll 1
      for (int c = 0; c < BLOCK_N / GEMM_EPI_CHUNK; ++c) {
        const int ncol0 = n0 + c * GEMM_EPI_CHUNK;
        if (ncol0 >= p.N) break;  // uniform across the CTA
        uint32_t v[64];
        tmem_ld_x32(t_row + c * GEMM_EPI_CHUNK, v);
        tmem_ld_x32(t_row + c * GEMM_EPI_CHUNK + 32, v + 32);

        // staging buffer `sbuf` must have been drained by the TMA store issued two chunks ago
        if (et == 0) tma_store_wait_read<1>();
        named_bar_sync(1, GEMM_EPI_THREADS);
        tmem_ld_wait();

        uint8_t* stg = sStage + sbuf * Cfg::kStagingBytes;
        const __nv_bfloat16* res_row = p.C + (long long)(m0 + row) * p.ldc + ncol0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {  // 8 columns (one 16-byte chunk) at a time
          const int col = ncol0 + j * 8;
          float f[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(v[j * 8 + e]);
          if constexpr (kF4) {
#pragma unroll
            for (int e = 0; e < 8; ++e) f[e] *= sa;
          }
          if constexpr (kFp8) {
            if (col < p.N) {
              const float4 s0 = __ldg(reinterpret_cast<const float4*>(p.scale_b + col));
              const float4 s1 = __ldg(reinterpret_cast<const float4*>(p.scale_b + col + 4));
              const float sb[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
#pragma unroll
              for (int e = 0; e < 8; ++e) f[e] = sa * (sb[e] * f[e]);
            }
          }
          if (p.bias != nullptr && col < p.N) {
            uint4 bv = __ldg(reinterpret_cast<const uint4*>(p.bias + col));
            const uint32_t bw[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              f[2 * e] += bf16_lo(bw[e]);
              f[2 * e + 1] += bf16_hi(bw[e]);
            }
          }
          if constexpr (EPI == EPI_BIAS_GELU) {
#pragma unroll
            for (int e = 0; e < 8; ++e) f[e] = gelu_tanh_f(bf16_round(f[e]));
          }
          if constexpr (EPI == EPI_GATE_RESIDUAL || EPI == EPI_RESIDUAL) {
            uint4 rv = make_uint4(0, 0, 0, 0);
            if (row_ok && col < p.N) rv = *reinterpret_cast<const uint4*>(res_row + j * 8);
            const uint32_t rw[4] = {rv.x, rv.y, rv.z, rv.w};
            if constexpr (EPI == EPI_GATE_RESIDUAL) {
              uint4 gv = make_uint4(0, 0, 0, 0);
              if (col < p.N) gv = __ldg(reinterpret_cast<const uint4*>(p.gate + col));
              const uint32_t gw[4] = {gv.x, gv.y, gv.z, gv.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                // reference rounding points: y -> bf16, y*gate -> bf16, x + (.) -> bf16
                f[2 * e] = bf16_lo(rw[e]) + bf16_round(bf16_round(f[2 * e]) * bf16_lo(gw[e]));
                f[2 * e + 1] = bf16_hi(rw[e]) + bf16_round(bf16_round(f[2 * e + 1]) * bf16_hi(gw[e]));
              }
            } else {
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                f[2 * e] = bf16_lo(rw[e]) + bf16_round(f[2 * e]);
                f[2 * e + 1] = bf16_hi(rw[e]) + bf16_round(f[2 * e + 1]);
              }
            }
          }
          uint4 o;
          o.x = pack_bf16(f[0], f[1]);
          o.y = pack_bf16(f[2], f[3]);
          o.z = pack_bf16(f[4], f[5]);
          o.w = pack_bf16(f[6], f[7]);
          // 128B-swizzled staging row: 16-byte chunk j of row r lives at chunk (j ^ (r & 7))
          *reinterpret_cast<uint4*>(stg + row * 128 + ((j ^ (row & 7)) << 4)) = o;
        }
        fence_async_smem();
        named_bar_sync(1, GEMM_EPI_THREADS);
        if (et == 0) {
          tma_store_2d(&tmC, stg, ncol0, m0);
          tma_store_commit();
        }
        sbuf ^= 1;
      }
      // accumulator stage fully read -> hand it back to the MMA warp
      tc_fence_before();
      mbar_arrive(&tmem_empty_bar[acc]);
      if (++acc == kAccStages) {
        acc = 0;
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

template <int BLOCK_N, int EPI, int MODE>
static int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const CUtensorMap& tmSFA,
                       const CUtensorMap& tmSFB, const GemmParams& p, int max_ctas, cudaStream_t stream) {
  using Cfg = GemmCfg<BLOCK_N, MODE>;
  auto kern = gemm_bf16_kernel<BLOCK_N, EPI, MODE>;
  static bool attr_set = false;
  if (!attr_set) {
    B200_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    attr_set = true;
  }
  const int num_tiles = p.num_m_blocks * p.num_n_blocks;
  int grid = num_tiles < max_ctas ? num_tiles : max_ctas;
  kern<<<grid, GEMM_THREADS, Cfg::kSmemBytes, stream>>>(tmA, tmB, tmC, tmSFA, tmSFB, p); note_launch();
  B200_CHECK_CUDA(cudaGetLastError());
  return B200_OK;
}

template <int BLOCK_N, int MODE>
static int dispatch_epi(int epi, const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC,
                        const GemmParams& p, int max_ctas, cudaStream_t stream, const CUtensorMap* tmSFA = nullptr,
                        const CUtensorMap* tmSFB = nullptr) {
  const CUtensorMap& sfa = tmSFA ? *tmSFA : tmA;   // unused outside MODE_NVFP4
  const CUtensorMap& sfb = tmSFB ? *tmSFB : tmB;
  switch (epi) {
    case EPI_BIAS: return launch_gemm<BLOCK_N, EPI_BIAS, MODE>(tmA, tmB, tmC, sfa, sfb, p, max_ctas, stream);
    case EPI_BIAS_GELU: return launch_gemm<BLOCK_N, EPI_BIAS_GELU, MODE>(tmA, tmB, tmC, sfa, sfb, p, max_ctas, stream);
    case EPI_GATE_RESIDUAL: return launch_gemm<BLOCK_N, EPI_GATE_RESIDUAL, MODE>(tmA, tmB, tmC, sfa, sfb, p, max_ctas, stream);
    case EPI_RESIDUAL: return launch_gemm<BLOCK_N, EPI_RESIDUAL, MODE>(tmA, tmB, tmC, sfa, sfb, p, max_ctas, stream);
  }
  set_last_error("b200_gemm_bf16: unknown epilogue %d", epi);
  return B200_ERR_INVALID;
}

int gemm_bf16(const void* A, long long lda, const void* B, long long ldb, void* C, long long ldc, const void* bias,
              const void* gate, long long M, long long N, long long K, int epilogue, int block_n, int max_ctas,
              cudaStream_t stream) {
  B200_CHECK_ARG(A && B && C, "b200_gemm_bf16: null operand pointer");
  B200_CHECK_ARG(M > 0 && N > 0 && K > 0, "b200_gemm_bf16: non-positive shape M=%lld N=%lld K=%lld", M, N, K);
  B200_CHECK_ARG(K % 8 == 0 && N % 8 == 0, "b200_gemm_bf16: K (%lld) and N (%lld) must be multiples of 8", K, N);
  B200_CHECK_ARG(lda % 8 == 0 && ldb % 8 == 0 && ldc % 8 == 0 && lda >= K && ldb >= K && ldc >= N,
                 "b200_gemm_bf16: leading dimensions must be multiples of 8 and cover the row (lda=%lld ldb=%lld ldc=%lld)",
                 lda, ldb, ldc);
  B200_CHECK_ARG(((uintptr_t)A % 16 == 0) && ((uintptr_t)B % 16 == 0) && ((uintptr_t)C % 16 == 0),
                 "b200_gemm_bf16: operands must be 16-byte aligned");
  B200_CHECK_ARG(epilogue != EPI_GATE_RESIDUAL || gate != nullptr, "b200_gemm_bf16: gate epilogue needs a gate vector");
  B200_CHECK_ARG((bias == nullptr || (uintptr_t)bias % 16 == 0) && (gate == nullptr || (uintptr_t)gate % 16 == 0),
                 "b200_gemm_bf16: bias / gate must be 16-byte aligned");
  if (block_n == 0) block_n = (N >= 256) ? 256 : 128;
  B200_CHECK_ARG(block_n == 128 || block_n == 256, "b200_gemm_bf16: block_n must be 128 or 256");
  if (max_ctas <= 0) max_ctas = num_sms();

  CUtensorMap tmA, tmB, tmC;
  int rc;
  if ((rc = encode_tmap_2d_bf16(&tmA, A, M, K, lda, GEMM_BLOCK_M, GEMM_BLOCK_K))) return rc;
  if ((rc = encode_tmap_2d_bf16(&tmB, B, N, K, ldb, block_n, GEMM_BLOCK_K))) return rc;
  if ((rc = encode_tmap_2d_bf16(&tmC, C, M, N, ldc, GEMM_BLOCK_M, GEMM_EPI_CHUNK))) return rc;

  GemmParams p;
  p.M = (int)M;
  p.N = (int)N;
  p.K = (int)K;
  p.num_m_blocks = (int)((M + GEMM_BLOCK_M - 1) / GEMM_BLOCK_M);
  p.num_n_blocks = (int)((N + block_n - 1) / block_n);
  p.num_k_blocks = (int)((K + GEMM_BLOCK_K - 1) / GEMM_BLOCK_K);
  p.bias = reinterpret_cast<const __nv_bfloat16*>(bias);
  p.gate = reinterpret_cast<const __nv_bfloat16*>(gate);
  p.C = reinterpret_cast<__nv_bfloat16*>(C);
  p.ldc = ldc;
  {
    static int group_m = 0;
    if (group_m == 0) {
      const char* e = getenv("B200_GEMM_GROUP_M");
      group_m = e ? atoi(e) : 16;
      if (group_m < 1) group_m = 16;
    }
    p.group_m = group_m;
  }
  p.scale_a = nullptr;
  p.scale_b = nullptr;
  p.alpha = nullptr;
  if (block_n == 256) return dispatch_epi<256, MODE_BF16>(epilogue, tmA, tmB, tmC, p, max_ctas, stream);
  return dispatch_epi<128, MODE_BF16>(epilogue, tmA, tmB, tmC, p, max_ctas, stream);
}

// e4m3 x e4m3 -> bf16 with per-token / per-channel scales.  A [M,K] e4m3, B [N,K] e4m3 (lda/ldb in elements = bytes).
int gemm_fp8(const void* A, long long lda, const void* B, long long ldb, void* C, long long ldc, const float* scale_a,
             const float* scale_b, const void* bias, const void* gate, long long M, long long N, long long K,
             int epilogue, int block_n, int max_ctas, cudaStream_t stream) {
  B200_CHECK_ARG(A && B && C && scale_a && scale_b, "b200_gemm_fp8: null operand / scale pointer");
  B200_CHECK_ARG(M > 0 && N > 0 && K > 0, "b200_gemm_fp8: non-positive shape M=%lld N=%lld K=%lld", M, N, K);
  B200_CHECK_ARG(K % 16 == 0 && N % 8 == 0, "b200_gemm_fp8: K (%lld) must be a multiple of 16 and N (%lld) of 8", K, N);
  B200_CHECK_ARG(lda % 16 == 0 && ldb % 16 == 0 && ldc % 8 == 0 && lda >= K && ldb >= K && ldc >= N,
                 "b200_gemm_fp8: bad leading dimensions (lda=%lld ldb=%lld ldc=%lld)", lda, ldb, ldc);
  B200_CHECK_ARG(((uintptr_t)A % 16 == 0) && ((uintptr_t)B % 16 == 0) && ((uintptr_t)C % 16 == 0) &&
                     ((uintptr_t)scale_b % 16 == 0),
                 "b200_gemm_fp8: operands must be 16-byte aligned");
  B200_CHECK_ARG(epilogue != EPI_GATE_RESIDUAL || gate != nullptr, "b200_gemm_fp8: gate epilogue needs a gate vector");
  if (block_n == 0) block_n = (N >= 256) ? 256 : 128;
  B200_CHECK_ARG(block_n == 128 || block_n == 256, "b200_gemm_fp8: block_n must be 128 or 256");
  if (max_ctas <= 0) max_ctas = num_sms();

  CUtensorMap tmA, tmB, tmC;
  int rc;
  {
    uint64_t dims[2] = {(uint64_t)K, (uint64_t)M};
    uint64_t strides[1] = {(uint64_t)lda};
    uint32_t box[2] = {128, GEMM_BLOCK_M};
    if ((rc = encode_tmap(&tmA, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, A, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  }
  {
    uint64_t dims[2] = {(uint64_t)K, (uint64_t)N};
    uint64_t strides[1] = {(uint64_t)ldb};
    uint32_t box[2] = {128, (uint32_t)block_n};
    if ((rc = encode_tmap(&tmB, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, B, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  }
  if ((rc = encode_tmap_2d_bf16(&tmC, C, M, N, ldc, GEMM_BLOCK_M, GEMM_EPI_CHUNK))) return rc;

  GemmParams p;
  p.M = (int)M;
  p.N = (int)N;
  p.K = (int)K;
  p.num_m_blocks = (int)((M + GEMM_BLOCK_M - 1) / GEMM_BLOCK_M);
  p.num_n_blocks = (int)((N + block_n - 1) / block_n);
  p.num_k_blocks = (int)((K + 127) / 128);
  p.bias = reinterpret_cast<const __nv_bfloat16*>(bias);
  p.gate = reinterpret_cast<const __nv_bfloat16*>(gate);
  p.C = reinterpret_cast<__nv_bfloat16*>(C);
  p.ldc = ldc;
  p.group_m = 16;
  p.scale_a = scale_a;
  p.scale_b = scale_b;
  p.alpha = nullptr;
  if (block_n == 256) return dispatch_epi<256, MODE_FP8>(epilogue, tmA, tmB, tmC, p, max_ctas, stream);
  return dispatch_epi<128, MODE_FP8>(epilogue, tmA, tmB, tmC, p, max_ctas, stream);
}

// e2m1 x e2m1 (two values per byte, low nibble first) with ue4m3 scale factors per 16 elements in the 128x4 "swizzled" layout
// [rows/128][K/64][32][4][4] (lightx2v_kernel scaled_fp4_quant, nvfp4_quant_kernels_sm120.cu:118-156) -> bf16.
// A [M, K/2] bytes, B [N, K/2] bytes (lda / ldb in bytes); sfa / sfb padded to whole 128-row tiles; alpha: device scalar.
int gemm_nvfp4(const void* A, long long lda, const void* B, long long ldb, const void* sfa, const void* sfb, const float* alpha,
               void* C, long long ldc, const void* bias, const void* gate, long long M, long long N, long long K, int epilogue,
               int block_n, int max_ctas, cudaStream_t stream) {
  B200_CHECK_ARG(A && B && C && sfa && sfb && alpha, "b200_gemm_nvfp4: null operand / scale pointer");
  B200_CHECK_ARG(M > 0 && N > 0 && K > 0, "b200_gemm_nvfp4: non-positive shape M=%lld N=%lld K=%lld", M, N, K);
  B200_CHECK_ARG(K % 64 == 0 && N % 8 == 0, "b200_gemm_nvfp4: K (%lld) must be a multiple of 64 and N (%lld) of 8", K, N);
  B200_CHECK_ARG(lda % 16 == 0 && ldb % 16 == 0 && ldc % 8 == 0 && lda >= K / 2 && ldb >= K / 2 && ldc >= N,
                 "b200_gemm_nvfp4: bad leading dimensions (lda=%lld ldb=%lld ldc=%lld)", lda, ldb, ldc);
  B200_CHECK_ARG(((uintptr_t)A % 16 == 0) && ((uintptr_t)B % 16 == 0) && ((uintptr_t)C % 16 == 0) && ((uintptr_t)sfa % 16 == 0) &&
                     ((uintptr_t)sfb % 16 == 0),
                 "b200_gemm_nvfp4: operands must be 16-byte aligned");
  B200_CHECK_ARG(epilogue != EPI_GATE_RESIDUAL || gate != nullptr, "b200_gemm_nvfp4: gate epilogue needs a gate vector");
  if (block_n == 0) block_n = (N >= 256) ? 256 : 128;   // measured: 256 (one accumulator stage) beats 128 (two stages) by 0-10 %
  B200_CHECK_ARG(block_n == 128 || block_n == 256, "b200_gemm_nvfp4: block_n must be 128 or 256");
  if (max_ctas <= 0) max_ctas = num_sms();

  CUtensorMap tmA, tmB, tmC, tmSFA, tmSFB;
  int rc;
  {
    uint64_t dims[2] = {(uint64_t)(K / 2), (uint64_t)M};
    uint64_t strides[1] = {(uint64_t)lda};
    uint32_t box[2] = {128, GEMM_BLOCK_M};
    if ((rc = encode_tmap(&tmA, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, A, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  }
  {
    uint64_t dims[2] = {(uint64_t)(K / 2), (uint64_t)N};
    uint64_t strides[1] = {(uint64_t)ldb};
    uint32_t box[2] = {128, (uint32_t)block_n};
    if ((rc = encode_tmap(&tmB, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, B, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  }
  // scale factors as [128-row tile][K/64][512 bytes]; one box = the 4 SF tiles of a 256-element K-block (x 1 or 2 row tiles);
  // K-blocks past the end of K are zero-filled by the TMA unit (zero scale, zero data -> contributes nothing)
  const uint64_t k_tiles = (uint64_t)(K / 64);
  {
    uint64_t dims[3] = {128, k_tiles, (uint64_t)((M + 127) / 128)};
    uint64_t strides[2] = {512, k_tiles * 512};
    uint32_t box[3] = {128, 4, 1};
    if ((rc = encode_tmap(&tmSFA, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, sfa, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE))) return rc;
  }
  {
    uint64_t dims[3] = {128, k_tiles, (uint64_t)((N + 127) / 128)};
    uint64_t strides[2] = {512, k_tiles * 512};
    uint32_t box[3] = {128, 4, (uint32_t)(block_n / 128)};
    if ((rc = encode_tmap(&tmSFB, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, sfb, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE))) return rc;
  }
  if ((rc = encode_tmap_2d_bf16(&tmC, C, M, N, ldc, GEMM_BLOCK_M, GEMM_EPI_CHUNK))) return rc;

  GemmParams p;
  p.M = (int)M;
  p.N = (int)N;
  p.K = (int)K;
  p.num_m_blocks = (int)((M + GEMM_BLOCK_M - 1) / GEMM_BLOCK_M);
  p.num_n_blocks = (int)((N + block_n - 1) / block_n);
  p.num_k_blocks = (int)((K + 255) / 256);
  p.bias = reinterpret_cast<const __nv_bfloat16*>(bias);
  p.gate = reinterpret_cast<const __nv_bfloat16*>(gate);
  p.C = reinterpret_cast<__nv_bfloat16*>(C);
  p.ldc = ldc;
  p.group_m = 16;
  p.scale_a = nullptr;
  p.scale_b = nullptr;
  p.alpha = alpha;
  if (block_n == 256) return dispatch_epi<256, MODE_NVFP4>(epilogue, tmA, tmB, tmC, p, max_ctas, stream, &tmSFA, &tmSFB);
  return dispatch_epi<128, MODE_NVFP4>(epilogue, tmA, tmB, tmC, p, max_ctas, stream, &tmSFA, &tmSFB);
}

}  // namespace b200
