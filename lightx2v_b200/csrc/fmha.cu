// Fused multi-head attention forward, head_dim 128 or 64, non-causal, one varlen segment per launch.
//
// Replaces FlashAttn2Weight.apply = flash_attn_varlen_func(q, k, v, cu_q, cu_k, max_q, max_k)
// (reference: lightx2v/common/ops/attn/attn_weight.py:71-97; softmax scale d^-0.5, no dropout, non-causal,
// SURVEY appendix A.8) for Wan self-attention (Sq = Sk = 75 600), Wan cross-attention (Sk = 512 / 257), the
// Hunyuan joint attention segments (d = 128) and CogVideoX's joint text+video attention (48 heads x d = 64,
// lightx2v/models/networks/cogvideox/infer/transformer_infer.py:85-145).
//
// sm_100a design (one CTA per SM, 2 x 128 query rows of one head per CTA, ping-pong):
//   warp 0        TMA producer: Q (once), then K_0 V_0 K_1 V_1 ... through a ring of [128 x d] tiles
//   warp 1        MMA issuer:   S_t = Q_t K_j^T   (tcgen05.mma SS, fp32 S in TMEM)
//                               O_t += P_t V_j    (tcgen05.mma TS: P read from TMEM, V MN-major from smem)
//   warp 2        TMEM allocator (512 columns: S0 | S1 | O0 | O1; P_t aliases the first 64 columns of S_t as bf16)
//   warps 4..7    softmax for query tile 0: tcgen05.ld S row -> online max/sum with lazy rescale -> exp2 ->
//   warps 8..11   softmax for query tile 1:   bf16 P -> tcgen05.st; O rescale in TMEM when the max moved by > 2^8;
//                                             final O / l -> bf16 -> global.
// While the softmax warpgroup of one tile works on S_{j+1}, the tensor pipe runs PV_j and QK_{j+1} of the other tile.
// P is published to the MMA warp in two 64-column halves (two mbarriers per tile) so that the first four PV MMAs run while the
// second half of the exponentials is still being computed.
//
// Measured-and-rejected variants (two softmax warpgroups per tile, speculative row max, 64-row K/V tiles with S and P un-aliased,
// early upper-half QK^T; profiles/r01_kernel_timings.jsonl) were removed from the shipped library in round 2; they are in the
// history of this file (commit de168c3).
#include "host_util.cuh"
#include "ptx.cuh"

#include <stdlib.h>

namespace b200 {

constexpr int FMHA_BLOCK_Q = 128;   // rows per query tile; two tiles per CTA
constexpr int FMHA_BLOCK_KV = 128;
constexpr int FMHA_THREADS = 384;

template <int kD>
struct FmhaCfg {
  static_assert(kD == 64 || kD == 128, "head_dim 64 or 128");
  static constexpr int kPanels = kD / 64;                       // [128 rows][64 d] swizzle-128B panels per tile
  static constexpr int kPanelBytes = 128 * 64 * 2;              // 16 KB
  static constexpr int kTileBytes = kPanels * kPanelBytes;      // 32 KB (d = 128) / 16 KB (d = 64)
  static constexpr int kStages = kD == 128 ? 4 : 8;             // K/V ring depth (128 KB either way)
  static constexpr int kSmemBytes = 2 * kTileBytes + kStages * kTileBytes + 1024 + 512;
};

struct FmhaParams {
  int sq, sk;               // rows of this segment
  int num_kv_tiles;
  float scale_log2;         // softmax_scale * log2(e)
  __nv_bfloat16* out;       // [sq, H, d] (+ stride)
  long long o_stride_s;     // elements between consecutive rows of out
  // Ulysses return exchange fused into the epilogue: when rows_per_rank > 0, query row r belongs to rank r / rows_per_rank and
  // is stored straight into that peer's buffer peer_out[rank] at local row r % rows_per_rank, head (head_offset + head).
  __nv_bfloat16* peer_out[8];
  long long rows_per_rank;
  int head_offset;
};

template <int kD>
__device__ __forceinline__ __nv_bfloat16* fmha_out_row(const FmhaParams& p, int q_row, int head) {
  if (p.rows_per_rank > 0) {
    const int dest = (int)(q_row / p.rows_per_rank);
    const long long local = q_row - (long long)dest * p.rows_per_rank;
    return p.peer_out[dest] + local * p.o_stride_s + (long long)(p.head_offset + head) * kD;
  }
  return p.out + (long long)q_row * p.o_stride_s + (long long)head * kD;
}

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int kD, int kPolyPairs>
__global__ void __launch_bounds__(FMHA_THREADS, 1)
fmha_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                     const __grid_constant__ CUtensorMap tmV, const FmhaParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  using Cfg = FmhaCfg<kD>;
  constexpr int FMHA_TILE_BYTES = Cfg::kTileBytes, FMHA_PANEL_BYTES = Cfg::kPanelBytes, FMHA_KV_STAGES = Cfg::kStages;
  constexpr int kPanels = Cfg::kPanels, kKSteps = kD / 16;
  uint8_t* sQ = smem;                             // 2 query tiles
  uint8_t* sKV = smem + 2 * FMHA_TILE_BYTES;      // K/V ring
  uint64_t* bars = reinterpret_cast<uint64_t*>(sKV + FMHA_KV_STAGES * FMHA_TILE_BYTES);
  uint64_t* q_full = bars;                          // [1]
  uint64_t* kv_full = bars + 1;                     // [kStages]
  uint64_t* kv_empty = kv_full + FMHA_KV_STAGES;    // [kStages]
  uint64_t* s_full = kv_empty + FMHA_KV_STAGES;     // [2]
  uint64_t* p_full = s_full + 2;                    // [2 tiles][2 halves]
  uint64_t* o_full = p_full + 4;                    // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int head = blockIdx.y;
  const int q0 = blockIdx.x * (2 * FMHA_BLOCK_Q);
  const int n_kv = p.num_kv_tiles;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmQ);
    prefetch_tmap(&tmK);
    prefetch_tmap(&tmV);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < FMHA_KV_STAGES; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    for (int t = 0; t < 2; ++t) {
      mbar_init(&s_full[t], 1);
      mbar_init(&p_full[2 * t], 128);
      mbar_init(&p_full[2 * t + 1], 128);
      mbar_init(&o_full[t], 1);
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // column offsets
  constexpr uint32_t COL_S0 = 0, COL_S1 = 128, COL_O0 = 256, COL_O1 = 384;

  if (warp < 4) {
   // warpgroup 0 (load / mma / alloc / idle) donates registers to the softmax warps
   reg_dealloc<88>();
   if (warp == 0) {
    // ============================== TMA producer ==============================
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, 2 * FMHA_TILE_BYTES);
      for (int t = 0; t < 2; ++t)
        for (int h = 0; h < kPanels; ++h)
          tma_load_3d(sQ + t * FMHA_TILE_BYTES + h * FMHA_PANEL_BYTES, &tmQ, q_full, h * 64, head,
                      q0 + t * FMHA_BLOCK_Q);
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0; it < 2 * n_kv; ++it) {   // K_0, V_0, K_1, V_1, ...
        const int j = it >> 1;
        const CUtensorMap* tm = (it & 1) ? &tmV : &tmK;
        mbar_wait(&kv_empty[stage], phase ^ 1);
        mbar_arrive_expect_tx(&kv_full[stage], FMHA_TILE_BYTES);
        for (int h = 0; h < kPanels; ++h)
          tma_load_3d(sKV + stage * FMHA_TILE_BYTES + h * FMHA_PANEL_BYTES, tm, &kv_full[stage], h * 64, head,
                      j * FMHA_BLOCK_KV);
        if (++stage == FMHA_KV_STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ============================== MMA issuer ==============================
    // The whole warp runs this (warp-uniform) control flow so that descriptors and loop state live in uniform registers;
    // each tcgen05.mma / tcgen05.commit is executed by one elected lane (the *_w wrappers).
    {
      constexpr uint32_t idesc_qk = make_idesc(FMT_BF16, FMT_BF16, 128, 128, 0, 0);  // A=Q K-major, B=K K-major
      constexpr uint32_t idesc_pv = make_idesc(FMT_BF16, FMT_BF16, 128, kD, 0, 1);   // A=P (TMEM), B=V MN-major, N = d
      const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);   // make the TMEM base provably warp-uniform (-> uniform register)
      const uint32_t tS0 = tb + COL_S0, tS1 = tb + COL_S1;
      const uint32_t tO0 = tb + COL_O0, tO1 = tb + COL_O1;
      const uint32_t q_lo = desc_lo_kmajor(smem_u32(sQ));
      const uint32_t kv_addr = smem_u32(sKV);

      // MMAs are issued in runs of four K-steps (one 64-wide swizzled panel) under one election (mma_f16_ss_w4 / mma_f16_ts_w4): with
      // one election per MMA this warp spent 74 % of its time in ELECT / R2UR / VOTEU code - 558 instructions per KV iteration for
      // 32 MMAs (round-2 ncu source view) - and paced the tensor pipe.
      auto issue_qk = [&](int t, int kstage) {
        const uint32_t a = q_lo + t * (FMHA_TILE_BYTES >> 4);
        const uint32_t b = desc_lo_kmajor(kv_addr + kstage * FMHA_TILE_BYTES);
        const uint32_t d = t ? tS1 : tS0;
#pragma unroll
        for (int pn = 0; pn < kKSteps / 4; ++pn) {  // head dim in steps of 16; 4 steps (32 B apart) per 64-wide panel
          const uint32_t off = (pn * FMHA_PANEL_BYTES) >> 4;
          mma_f16_ss_w4(d, a + off, kDescHiSw128, b + off, kDescHiSw128, idesc_qk, pn != 0 ? 1u : 0u);
        }
      };
      // kv = 128 in steps of 16 rows (16 x 128 B = 2048 B); P: 8 columns per step; half 0 = steps 0..3, half 1 = steps 4..7
      auto issue_pv_half = [&](int t, int vstage, uint32_t accumulate, int half) {
        const uint32_t b = desc_lo_mnmajor(kv_addr + vstage * FMHA_TILE_BYTES, FMHA_PANEL_BYTES);
        const uint32_t d = t ? tO1 : tO0;
        const uint32_t a = t ? tS1 : tS0;
        mma_f16_ts_w4(d, a + half * 32, b + half * (4 * 2048 >> 4), kDescHiSw128, idesc_pv, half ? 1u : accumulate);
      };
      auto issue_pv = [&](int t, int vstage, uint32_t accumulate, uint32_t parity) {
        mbar_wait(&p_full[2 * t], parity);
        tc_fence_after();
        issue_pv_half(t, vstage, accumulate, 0);
        mbar_wait(&p_full[2 * t + 1], parity);
        tc_fence_after();
        issue_pv_half(t, vstage, accumulate, 1);
      };

      int stage = 0;        // ring position of the next tile to consume
      uint32_t phase = 0;
      auto advance = [&]() {
        if (++stage == FMHA_KV_STAGES) {
          stage = 0;
          phase ^= 1;
        }
      };

      mbar_wait(q_full, 0);
      // K_0
      mbar_wait(&kv_full[stage], phase);
      tc_fence_after();
      issue_qk(0, stage);
      tc_commit_w(&s_full[0]);
      issue_qk(1, stage);
      tc_commit_w(&s_full[1]);
      tc_commit_w(&kv_empty[stage]);
      advance();

      for (int j = 0; j < n_kv; ++j) {
        const bool has_next = (j + 1) < n_kv;
        // V_j
        const int vstage = stage;
        mbar_wait(&kv_full[stage], phase);
        advance();
        // K_{j+1}
        const int kstage = stage;
        if (has_next) {
          mbar_wait(&kv_full[stage], phase);
          advance();
        }
        const uint32_t pj = j & 1;
        const uint32_t acc = j > 0 ? 1u : 0u;
        // ---- tile 0
        issue_pv(0, vstage, acc, pj);
        if (has_next) {
          issue_qk(0, kstage);
          tc_commit_w(&s_full[0]);
        } else {
          tc_commit_w(&o_full[0]);
        }
        // ---- tile 1
        issue_pv(1, vstage, acc, pj);
        tc_commit_w(&kv_empty[vstage]);
        if (has_next) {
          issue_qk(1, kstage);
          tc_commit_w(&s_full[1]);
          tc_commit_w(&kv_empty[kstage]);
        } else {
          tc_commit_w(&o_full[1]);
        }
      }
    }
   }
  } else {
    // ============================== softmax / correction / epilogue ==============================
    reg_alloc<208>();   // 128*88 + 256*208 == 384*168: the pool is what the CTA launched with
    const int t = (warp - 4) >> 2;             // query tile 0 / 1
    const int lg = warp & 3;                   // TMEM lane group of this warp
    const int row = lg * 32 + lane;            // row inside the query tile
    const uint32_t lane_off = uint32_t(lg * 32) << 16;
    const uint32_t tS = tmem_base + (t ? COL_S1 : COL_S0) + lane_off;
    const uint32_t tO = tmem_base + (t ? COL_O1 : COL_O0) + lane_off;
    const float sl2 = p.scale_log2;

    float m_used = -INFINITY;   // max the accumulated O / l are expressed against (raw score units)
    float l_sum = 0.f;

    for (int j = 0; j < n_kv; ++j) {
      mbar_wait(&s_full[t], j & 1);
      tc_fence_after();
      uint32_t s[128];
      tmem_ld_x32(tS + 0, s + 0);
      tmem_ld_x32(tS + 32, s + 32);
      tmem_ld_x32(tS + 64, s + 64);
      tmem_ld_x32(tS + 96, s + 96);
      tmem_ld_wait();

      const int kv_valid = p.sk - j * FMHA_BLOCK_KV;   // >= 1
      if (kv_valid < FMHA_BLOCK_KV) {
#pragma unroll
        for (int i = 0; i < 128; ++i)
          if (i >= kv_valid) s[i] = 0xff800000u;  // -inf
      }
      const float2 sl2v = make_float2(sl2, sl2);
      uint32_t pk[64];
      float2 acc[4] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f)};

      // exp2(s * scale_log2 - m * scale_log2) for 64 columns on packed f32x2 lanes (FFMA2 / FADD2 halve the fma-pipe instruction
      // count); kPolyPairs of every 4 pairs take the FMA-pipe polynomial instead of MUFU.EX2, the co-critical unit
      // (16 ex2/clk/SM vs 8192 tensor FLOP/clk/SM: a 128x128 tile costs 1024 MUFU cycles and 1024 MMA cycles).
      auto exp_half = [&](int half, float neg_m_) {
        const float2 negm = make_float2(neg_m_, neg_m_);
#pragma unroll
        for (int g = 0; g < 8; ++g) {
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) {
            const int i = half * 32 + g * 4 + jj;
            const float2 x = __ffma2_rn(make_float2(__uint_as_float(s[2 * i]), __uint_as_float(s[2 * i + 1])), sl2v, negm);
            float2 e;
            if (jj < kPolyPairs) {
              e = exp2_poly2(x);
            } else {
              e.x = ex2(x.x);
              e.y = ex2(x.y);
            }
            acc[jj] = __fadd2_rn(acc[jj], e);
            pk[i] = pack_bf16(e.x, e.y);
          }
        }
      };
      auto row_max = [&]() {
        float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
        for (int i = 0; i < 128; i += 4) {
          mx0 = fmaxf(mx0, __uint_as_float(s[i]));
          mx1 = fmaxf(mx1, __uint_as_float(s[i + 1]));
          mx2 = fmaxf(mx2, __uint_as_float(s[i + 2]));
          mx3 = fmaxf(mx3, __uint_as_float(s[i + 3]));
        }
        return fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
      };
      // lazy rescale: O and l follow the running max only when it moved by more than 2^8 in the exp2 domain
      auto rescale_if_needed = [&](float m_new) -> bool {
        const bool need = (m_new - m_used) * sl2 > 8.0f;
        if (!__any_sync(0xffffffffu, need)) return false;
        float alpha = 1.0f;
        if (need) {
          alpha = ex2((m_used - m_new) * sl2);
          m_used = m_new;
        }
        l_sum *= alpha;
#pragma unroll 1
        for (int c = 0; c < kD; c += 32) {
          uint32_t o[32];
          tmem_ld_x32(tO + c, o);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
          tmem_st_x32(tO + c, o);
        }
        tmem_st_wait();
        return true;
      };
      auto publish = [&](int half) {
        tmem_st_x32(tS + half * 32, pk + half * 32);   // P (bf16 pairs) overwrites the first 64 columns of S
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(&p_full[2 * t + half]);
      };

      {
        const float m_new = fmaxf(row_max(), m_used);
        if (j == 0) m_used = m_new;
        else rescale_if_needed(m_new);
        const float neg_m = -m_used * sl2;
        exp_half(0, neg_m);
        publish(0);
        exp_half(1, neg_m);
      }
      const float2 a01 = __fadd2_rn(acc[0], acc[1]), a23 = __fadd2_rn(acc[2], acc[3]);
      l_sum += (a01.x + a01.y) + (a23.x + a23.y);
      publish(1);
    }

    // ---- epilogue: O / l -> bf16 -> global
    mbar_wait(&o_full[t], 0);
    tc_fence_after();
    const float inv_l = 1.0f / l_sum;
    const int q_row = q0 + t * FMHA_BLOCK_Q + row;
    __nv_bfloat16* orow = q_row < p.sq ? fmha_out_row<kD>(p, q_row, head) : p.out;
#pragma unroll 1
    for (int c = 0; c < kD; c += 32) {
      uint32_t o[32];
      tmem_ld_x32(tO + c, o);
      tmem_ld_wait();
      if (q_row < p.sq) {
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          uint4 w;
          w.x = pack_bf16(__uint_as_float(o[v * 8 + 0]) * inv_l, __uint_as_float(o[v * 8 + 1]) * inv_l);
          w.y = pack_bf16(__uint_as_float(o[v * 8 + 2]) * inv_l, __uint_as_float(o[v * 8 + 3]) * inv_l);
          w.z = pack_bf16(__uint_as_float(o[v * 8 + 4]) * inv_l, __uint_as_float(o[v * 8 + 5]) * inv_l);
          w.w = pack_bf16(__uint_as_float(o[v * 8 + 6]) * inv_l, __uint_as_float(o[v * 8 + 7]) * inv_l);
          *reinterpret_cast<uint4*>(orow + c + v * 8) = w;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// q/k/v: [rows, H, d] bf16 with arbitrary row stride (elements), heads contiguous (stride d).
static int encode_qkv_map(CUtensorMap* tm, const void* base, long long rows, int heads, int head_dim, long long stride_s) {
  uint64_t dims[3] = {(uint64_t)head_dim, (uint64_t)heads, (uint64_t)rows};
  uint64_t strides[2] = {(uint64_t)head_dim * 2, (uint64_t)stride_s * 2};
  uint32_t box[3] = {64, 1, 128};
  return encode_tmap(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, base, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
}

template <int kD>
static int fmha_launch(const CUtensorMap& tmQ, const CUtensorMap& tmK, const CUtensorMap& tmV, const FmhaParams& p, dim3 grid, int poly,
                       cudaStream_t stream) {
  auto launch = [&](auto kern) -> int {
    B200_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, FmhaCfg<kD>::kSmemBytes));
    kern<<<grid, FMHA_THREADS, FmhaCfg<kD>::kSmemBytes, stream>>>(tmQ, tmK, tmV, p); note_launch();
    return B200_OK;
  };
  switch (poly) {
    case 0: return launch(fmha_fwd_kernel<kD, 0>);
    case 2: return launch(fmha_fwd_kernel<kD, 2>);
    default: return launch(fmha_fwd_kernel<kD, 1>);
  }
}

int fmha_fwd_impl(const void* q, long long q_stride_s, const void* k, long long k_stride_s, const void* v, long long v_stride_s, void* out,
                  long long o_stride_s, long long sq, long long sk, int heads, int head_dim, float softmax_scale, void* const* peers, int world,
                  long long rows_per_rank, int head_offset, cudaStream_t stream) {
  B200_CHECK_ARG(head_dim == 64 || head_dim == 128, "b200_fmha_fwd: head_dim must be 64 or 128 (got %d)", head_dim);
  if (peers != nullptr) {
    // FmhaParams::peer_out has 8 slots: one NVSwitch domain of a B200 box (include/b200_dit.h documents world <= 8)
    B200_CHECK_ARG(world >= 1 && world <= 8 && rows_per_rank > 0 && sq <= world * rows_per_rank,
                   "b200_fmha_fwd_d128_scatter: bad world %d / rows_per_rank %lld for sq %lld", world, rows_per_rank, sq);
    for (int i = 0; i < world; ++i) B200_CHECK_ARG(peers[i] != nullptr, "b200_fmha_fwd_d128_scatter: null peer pointer %d", i);
    out = peers[0];
  }
  B200_CHECK_ARG(q && k && v && out, "b200_fmha_fwd: null pointer");
  B200_CHECK_ARG(sq > 0 && sk > 0 && heads > 0, "b200_fmha_fwd: empty problem sq=%lld sk=%lld heads=%d", sq, sk, heads);
  B200_CHECK_ARG(q_stride_s % 8 == 0 && k_stride_s % 8 == 0 && v_stride_s % 8 == 0 && o_stride_s % 8 == 0,
                 "b200_fmha_fwd: row strides must be multiples of 8 elements");
  const long long hd = (long long)heads * head_dim;
  B200_CHECK_ARG(q_stride_s >= hd && k_stride_s >= hd && v_stride_s >= hd && o_stride_s >= hd, "b200_fmha_fwd: row stride smaller than heads*head_dim");
  B200_CHECK_ARG(((uintptr_t)q % 16 == 0) && ((uintptr_t)k % 16 == 0) && ((uintptr_t)v % 16 == 0) && ((uintptr_t)out % 16 == 0),
                 "b200_fmha_fwd: pointers must be 16-byte aligned");
  CUtensorMap tmQ, tmK, tmV;
  int rc;
  if ((rc = encode_qkv_map(&tmQ, q, sq, heads, head_dim, q_stride_s))) return rc;
  if ((rc = encode_qkv_map(&tmK, k, sk, heads, head_dim, k_stride_s))) return rc;
  if ((rc = encode_qkv_map(&tmV, v, sk, heads, head_dim, v_stride_s))) return rc;

  FmhaParams p;
  p.sq = (int)sq;
  p.sk = (int)sk;
  p.num_kv_tiles = (int)((sk + FMHA_BLOCK_KV - 1) / FMHA_BLOCK_KV);
  p.scale_log2 = softmax_scale * 1.4426950408889634f;
  p.out = reinterpret_cast<__nv_bfloat16*>(out);
  p.o_stride_s = o_stride_s;
  p.rows_per_rank = peers ? rows_per_rank : 0;
  p.head_offset = head_offset;
  for (int i = 0; i < 8; ++i) p.peer_out[i] = (peers && i < world) ? reinterpret_cast<__nv_bfloat16*>(peers[i]) : nullptr;

  // fraction of exp2 evaluated by the FMA-pipe polynomial: B200_FMHA_POLY = 0..2 pairs out of every 4 (default 1)
  static int poly = -1;
  if (poly < 0) {
    const char* e = getenv("B200_FMHA_POLY");
    poly = e ? atoi(e) : 1;   // measured on B200: 0 -> 1.19, 1 -> 1.23, 2 -> 1.13 PFLOP/s at S = 75 600 x 40 heads
    if (poly < 0 || poly > 2) poly = 1;
  }
  dim3 grid((unsigned)((sq + 2 * FMHA_BLOCK_Q - 1) / (2 * FMHA_BLOCK_Q)), (unsigned)heads, 1);
  int rc2;
  {
    ProfScope prof(sq, sk, heads, head_dim, stream);
    rc2 = head_dim == 128 ? fmha_launch<128>(tmQ, tmK, tmV, p, grid, poly, stream) : fmha_launch<64>(tmQ, tmK, tmV, p, grid, poly, stream);
  }
  if (rc2) return rc2;
  B200_CHECK_CUDA(cudaGetLastError());
  return B200_OK;
}

int fmha_fwd_d128(const void* q, long long q_stride_s, const void* k, long long k_stride_s, const void* v,
                  long long v_stride_s, void* out, long long o_stride_s, long long sq, long long sk, int heads,
                  float softmax_scale, cudaStream_t stream) {
  return fmha_fwd_impl(q, q_stride_s, k, k_stride_s, v, v_stride_s, out, o_stride_s, sq, sk, heads, 128, softmax_scale, nullptr, 0, 0, 0, stream);
}

int fmha_fwd_d64(const void* q, long long q_stride_s, const void* k, long long k_stride_s, const void* v, long long v_stride_s, void* out,
                 long long o_stride_s, long long sq, long long sk, int heads, float softmax_scale, cudaStream_t stream) {
  return fmha_fwd_impl(q, q_stride_s, k, k_stride_s, v, v_stride_s, out, o_stride_s, sq, sk, heads, 64, softmax_scale, nullptr, 0, 0, 0, stream);
}

int fmha_fwd_d128_scatter(const void* q, long long q_stride_s, const void* k, long long k_stride_s, const void* v,
                          long long v_stride_s, void* const* peers, int world, long long rows_per_rank, long long peer_stride_s,
                          int head_offset, long long sq, long long sk, int heads, float softmax_scale, cudaStream_t stream) {
  B200_CHECK_ARG(peers != nullptr, "b200_fmha_fwd_d128_scatter: null peer table");
  return fmha_fwd_impl(q, q_stride_s, k, k_stride_s, v, v_stride_s, nullptr, peer_stride_s, sq, sk, heads, 128, softmax_scale, peers, world,
                       rows_per_rank, head_offset, stream);
}

}  // namespace b200
