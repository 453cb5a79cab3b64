// Fused multi-head attention forward, head_dim 128, non-causal, one varlen segment per launch.
//
// Replaces FlashAttn2Weight.apply = flash_attn_varlen_func(q, k, v, cu_q, cu_k, max_q, max_k)
// (reference: lightx2v/common/ops/attn/attn_weight.py:71-97; softmax scale d^-0.5, no dropout, non-causal,
// SURVEY appendix A.8) for Wan self-attention (Sq = Sk = 75 600), Wan cross-attention (Sk = 512 / 257) and the
// Hunyuan joint attention segments.
//
// sm_100a design (one CTA per SM, 2 x 128 query rows of one head per CTA, ping-pong):
//   warp 0        TMA producer: Q (once), then K_0 V_0 K_1 V_1 ... through a 4-deep 32 KB ring
//   warp 1        MMA issuer:   S_t = Q_t K_j^T   (tcgen05.mma SS, fp32 S in TMEM)
//                               O_t += P_t V_j    (tcgen05.mma TS: P read from TMEM, V MN-major from smem)
//   warp 2        TMEM allocator (512 columns: S0 | S1 | O0 | O1; P_t aliases the first 64 columns of S_t as bf16)
//   warps 4..7    softmax for query tile 0: tcgen05.ld S row -> online max/sum with lazy rescale -> exp2 ->
//   warps 8..11   softmax for query tile 1:   bf16 P -> tcgen05.st; O rescale in TMEM when the max moved by > 2^8;
//                                             final O / l -> bf16 -> global.
// While the softmax warpgroup of one tile works on S_{j+1}, the tensor pipe runs PV_j and QK_{j+1} of the other tile.
#include "host_util.cuh"
#include "ptx.cuh"

#include <stdlib.h>

namespace b200 {

constexpr int FMHA_D = 128;
constexpr int FMHA_BLOCK_Q = 128;   // rows per query tile; two tiles per CTA
constexpr int FMHA_BLOCK_KV = 128;
constexpr int FMHA_THREADS = 384;
constexpr int FMHA_KV_STAGES = 4;
constexpr int FMHA_TILE_BYTES = 128 * 128 * 2;   // 32 KB: [128 rows][128 d] bf16 as two [128][64] swizzled panels
constexpr int FMHA_PANEL_BYTES = 128 * 64 * 2;   // 16 KB
constexpr int FMHA_SMEM_BYTES = 2 * FMHA_TILE_BYTES + FMHA_KV_STAGES * FMHA_TILE_BYTES + 1024 + 256;

struct FmhaParams {
  int sq, sk;               // rows of this segment
  int num_kv_tiles;
  float scale_log2;         // softmax_scale * log2(e)
  __nv_bfloat16* out;       // [sq, H, 128] (+ stride)
  long long o_stride_s;     // elements between consecutive rows of out
  // Ulysses return exchange fused into the epilogue: when rows_per_rank > 0, query row r belongs to rank r / rows_per_rank and
  // is stored straight into that peer's buffer peer_out[rank] at local row r % rows_per_rank, head (head_offset + head).
  __nv_bfloat16* peer_out[8];
  long long rows_per_rank;
  int head_offset;
};

__device__ __forceinline__ __nv_bfloat16* fmha_out_row(const FmhaParams& p, int q_row, int head) {
  if (p.rows_per_rank > 0) {
    const int dest = (int)(q_row / p.rows_per_rank);
    const long long local = q_row - (long long)dest * p.rows_per_rank;
    return p.peer_out[dest] + local * p.o_stride_s + (long long)(p.head_offset + head) * FMHA_D;
  }
  return p.out + (long long)q_row * p.o_stride_s + (long long)head * FMHA_D;
}

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// 2^x for x <= ~8 on the FMA pipe: n = round(x) by the 1.5*2^23 trick, f = x - n in [-0.5, 0.5], degree-3 minimax
// polynomial for 2^f (max relative error 7.5e-5, far below the 2^-9 of the bf16 P it feeds), exponent add in integer.
__device__ __forceinline__ float2 exp2_poly2(float2 x) {
  x.x = fmaxf(x.x, -126.f);
  x.y = fmaxf(x.y, -126.f);
  const float2 magic = make_float2(12582912.f, 12582912.f);
  const float2 fi = __fadd2_rn(x, magic);
  const float2 n = __fadd2_rn(fi, make_float2(-12582912.f, -12582912.f));
  const float2 f = __ffma2_rn(n, make_float2(-1.f, -1.f), x);
  float2 p = __ffma2_rn(f, make_float2(0.055171649903059006f, 0.055171649903059006f),
                        make_float2(0.2426111251115799f, 0.2426111251115799f));
  p = __ffma2_rn(p, f, make_float2(0.6932609677314758f, 0.6932609677314758f));
  p = __ffma2_rn(p, f, make_float2(0.9999280571937561f, 0.9999280571937561f));
  float2 r;
  r.x = __int_as_float(__float_as_int(p.x) + (__float_as_int(fi.x) << 23));
  r.y = __int_as_float(__float_as_int(p.y) + (__float_as_int(fi.y) << 23));
  return r;
}

// kSplitP: P is published to the MMA warp in two 64-column halves (two mbarriers per tile) so that the first four PV
// MMAs run while the second half of the exponentials is still being computed.
// kSpecMax: the exponentials of the first half start against the previous running max while the new row max is still being
// reduced; the result is validated (and in the rare > 2^8 jump recomputed) before P is published.
// kEarlyQK: the upper 64 score columns of S_{j+1} do not overlap P_j (columns 0..63), so that half of QK_{j+1} is issued as soon
// as the softmax warps have pulled S_j into registers (s_free), ahead of PV_j; only the lower half waits for PV_j to finish.
template <int kPolyPairs, bool kSplitP, bool kSpecMax, bool kEarlyQK>
__global__ void __launch_bounds__(FMHA_THREADS, 1)
fmha_fwd_d128_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                     const __grid_constant__ CUtensorMap tmV, const FmhaParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                             // 2 x 32 KB
  uint8_t* sKV = smem + 2 * FMHA_TILE_BYTES;      // 4 x 32 KB ring
  uint64_t* bars = reinterpret_cast<uint64_t*>(sKV + FMHA_KV_STAGES * FMHA_TILE_BYTES);
  uint64_t* q_full = bars;               // [1]
  uint64_t* kv_full = bars + 1;          // [4]
  uint64_t* kv_empty = bars + 5;         // [4]
  uint64_t* s_full = bars + 9;           // [2]
  uint64_t* p_full = bars + 11;          // [2 tiles][2 halves] (the second half only with kSplitP)
  uint64_t* o_full = bars + 15;          // [2]
  uint64_t* s_free = bars + 17;          // [2] (kEarlyQK)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 19);

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
      mbar_init(&s_free[t], 128);
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
        for (int h = 0; h < 2; ++h)
          tma_load_3d(sQ + t * FMHA_TILE_BYTES + h * FMHA_PANEL_BYTES, &tmQ, q_full, h * 64, head,
                      q0 + t * FMHA_BLOCK_Q);
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0; it < 2 * n_kv; ++it) {   // K_0, V_0, K_1, V_1, ...
        const int j = it >> 1;
        const CUtensorMap* tm = (it & 1) ? &tmV : &tmK;
        mbar_wait(&kv_empty[stage], phase ^ 1);
        mbar_arrive_expect_tx(&kv_full[stage], FMHA_TILE_BYTES);
        for (int h = 0; h < 2; ++h)
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
      constexpr uint32_t idesc_pv = make_idesc(FMT_BF16, FMT_BF16, 128, 128, 0, 1);  // A=P (TMEM), B=V MN-major
      const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);   // make the TMEM base provably warp-uniform (-> uniform register)
      const uint32_t tS0 = tb + COL_S0, tS1 = tb + COL_S1;
      const uint32_t tO0 = tb + COL_O0, tO1 = tb + COL_O1;
      const uint32_t q_lo = desc_lo_kmajor(smem_u32(sQ));
      const uint32_t kv_addr = smem_u32(sKV);

      auto issue_qk = [&](int t, int kstage) {
        const uint32_t a = q_lo + t * (FMHA_TILE_BYTES >> 4);
        const uint32_t b = desc_lo_kmajor(kv_addr + kstage * FMHA_TILE_BYTES);
        const uint32_t d = t ? tS1 : tS0;
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {  // d = 128 in steps of 16; 4 steps per 64-wide panel
          const uint32_t off = ((ks >> 2) * FMHA_PANEL_BYTES + (ks & 3) * 32) >> 4;
          mma_f16_ss_w(d, a + off, kDescHiSw128, b + off, kDescHiSw128, idesc_qk, ks != 0 ? 1u : 0u);
        }
      };
      // one 64-column half of S: K rows [64 half, 64 half + 64) (8192 B into each panel), TMEM columns 64 half ..
      auto issue_qk_half = [&](int t, int kstage, int half) {
        constexpr uint32_t idesc_qk64 = make_idesc(FMT_BF16, FMT_BF16, 128, 64, 0, 0);
        const uint32_t a = q_lo + t * (FMHA_TILE_BYTES >> 4);
        const uint32_t b = desc_lo_kmajor(kv_addr + kstage * FMHA_TILE_BYTES + half * 8192);
        const uint32_t d = (t ? tS1 : tS0) + half * 64;
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          const uint32_t off = ((ks >> 2) * FMHA_PANEL_BYTES + (ks & 3) * 32) >> 4;
          mma_f16_ss_w(d, a + off, kDescHiSw128, b + off, kDescHiSw128, idesc_qk64, ks != 0 ? 1u : 0u);
        }
      };
      // kv = 128 in steps of 16 rows (16 x 128 B = 2048 B); P: 8 columns per step; ks in [ks0, ks1)
      auto issue_pv_range = [&](int t, int vstage, uint32_t accumulate, int ks0, int ks1) {
        const uint32_t b = desc_lo_mnmajor(kv_addr + vstage * FMHA_TILE_BYTES, FMHA_PANEL_BYTES);
        const uint32_t d = t ? tO1 : tO0;
        const uint32_t a = t ? tS1 : tS0;
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          if (ks >= ks0 && ks < ks1) mma_f16_ts_w(d, a + ks * 8, b + ks * (2048 >> 4), kDescHiSw128, idesc_pv, ks != 0 ? 1u : accumulate);
        }
      };
      auto issue_pv = [&](int t, int vstage, uint32_t accumulate, uint32_t parity) {
        mbar_wait(&p_full[2 * t], parity);
        tc_fence_after();
        if constexpr (kSplitP) {
          issue_pv_range(t, vstage, accumulate, 0, 4);
          mbar_wait(&p_full[2 * t + 1], parity);
          tc_fence_after();
          issue_pv_range(t, vstage, accumulate, 4, 8);
        } else {
          issue_pv_range(t, vstage, accumulate, 0, 8);
        }
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
        if constexpr (kEarlyQK) {
          if (has_next) {
            mbar_wait(&s_free[0], pj);
            tc_fence_after();
            issue_qk_half(0, kstage, 1);
          }
        }
        issue_pv(0, vstage, acc, pj);
        if (has_next) {
          if constexpr (kEarlyQK) issue_qk_half(0, kstage, 0);
          else issue_qk(0, kstage);
          tc_commit_w(&s_full[0]);
        } else {
          tc_commit_w(&o_full[0]);
        }
        // ---- tile 1
        if constexpr (kEarlyQK) {
          if (has_next) {
            mbar_wait(&s_free[1], pj);
            tc_fence_after();
            issue_qk_half(1, kstage, 1);
          }
        }
        issue_pv(1, vstage, acc, pj);
        tc_commit_w(&kv_empty[vstage]);
        if (has_next) {
          if constexpr (kEarlyQK) issue_qk_half(1, kstage, 0);
          else issue_qk(1, kstage);
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
      if constexpr (kEarlyQK) {
        tc_fence_before();
        mbar_arrive(&s_free[t]);          // S_j is in registers: the upper half of S may be overwritten by QK_{j+1}
      }

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
        for (int c = 0; c < 128; c += 32) {
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

      if (!kSpecMax || j == 0) {
        const float m_new = fmaxf(row_max(), m_used);
        if (j == 0) m_used = m_new;
        else rescale_if_needed(m_new);
        const float neg_m = -m_used * sl2;
        exp_half(0, neg_m);
        if constexpr (kSplitP) publish(0);
        exp_half(1, neg_m);
      } else {
        // Speculative: start the exponentials of the first half against the stale max immediately (no dependency on the row
        // reduction, which the scheduler interleaves on the ALU pipe); validate before anything is published.
        const float2 acc_save[4] = {acc[0], acc[1], acc[2], acc[3]};
        exp_half(0, -m_used * sl2);
        const float m_new = fmaxf(row_max(), m_used);
        if (rescale_if_needed(m_new)) {               // rare: the max jumped by > 2^8 — redo the first half against the new max
          acc[0] = acc_save[0]; acc[1] = acc_save[1]; acc[2] = acc_save[2]; acc[3] = acc_save[3];
          exp_half(0, -m_used * sl2);
        }
        if constexpr (kSplitP) publish(0);
        exp_half(1, -m_used * sl2);
      }
      const float2 a01 = __fadd2_rn(acc[0], acc[1]), a23 = __fadd2_rn(acc[2], acc[3]);
      l_sum += (a01.x + a01.y) + (a23.x + a23.y);
      if constexpr (kSplitP) {
        publish(1);
      } else {
        tmem_st_x32(tS + 0, pk + 0);
        tmem_st_x32(tS + 32, pk + 32);
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(&p_full[2 * t]);
      }
    }

    // ---- epilogue: O / l -> bf16 -> global
    mbar_wait(&o_full[t], 0);
    tc_fence_after();
    const float inv_l = 1.0f / l_sum;
    const int q_row = q0 + t * FMHA_BLOCK_Q + row;
    __nv_bfloat16* orow = q_row < p.sq ? fmha_out_row(p, q_row, head) : p.out;
#pragma unroll 1
    for (int c = 0; c < 128; c += 32) {
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

// =====================================================================================================================
// v3: same tiling and TMEM plan, but TWO softmax warpgroups per query tile (each thread owns half a row: 64 score columns),
// and P is handed to the MMA warp in two halves so PV_j starts while the second half of the exponentials is still running.
// Rationale (ncu, v2): with P aliased onto S the per-tile loop is serial — softmax (Ts) then PV + QK (1024 tensor cycles) —
// so tensor utilisation = 2048 / (Ts + 1024 + latencies); v2 had Ts ~ 1600 cycles (one warp per scheduler, 128 elements per
// thread).  Halving the per-thread work and doubling the warps per scheduler brings Ts towards the MUFU floor.
//   warps 4..7  : tile 0, columns   0..63      warps  8..11 : tile 0, columns 64..127
//   warps 12..15: tile 1, columns   0..63      warps 16..19 : tile 1, columns 64..127
// Row max and the final row sum are exchanged between the two owners of a row through shared memory + a 256-thread named barrier.
// =====================================================================================================================
constexpr int FMHA3_THREADS = 640;
constexpr int FMHA3_SMEM_BYTES = FMHA_SMEM_BYTES + 2 * 2 * 2 * 128 * 4 + 2 * 2 * 128 * 4;   // + max exchange (double-buffered) + sum exchange

template <int kPolyPairs>
__global__ void __launch_bounds__(FMHA3_THREADS, 1)
fmha_fwd_d128_v3_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                        const __grid_constant__ CUtensorMap tmV, const FmhaParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sKV = smem + 2 * FMHA_TILE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sKV + FMHA_KV_STAGES * FMHA_TILE_BYTES);
  uint64_t* q_full = bars;               // [1]
  uint64_t* kv_full = bars + 1;          // [4]
  uint64_t* kv_empty = bars + 5;         // [4]
  uint64_t* s_full = bars + 9;           // [2]
  uint64_t* p_half = bars + 11;          // [2 tiles][2 halves]
  uint64_t* o_full = bars + 15;          // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 17);
  float* xmax = reinterpret_cast<float*>(bars + 20);      // [parity 2][tile 2][half 2][128]
  float* xsum = xmax + 2 * 2 * 2 * 128;                   // [tile 2][half 2][128]

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
      mbar_init(&p_half[t * 2 + 0], 128);
      mbar_init(&p_half[t * 2 + 1], 128);
      mbar_init(&o_full[t], 1);
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  constexpr uint32_t COL_S0 = 0, COL_S1 = 128, COL_O0 = 256, COL_O1 = 384;

  if (warp < 4) {
    // 640 threads launch with 96 registers each; 128*48 + 512*104 <= 640*96
    reg_dealloc<48>();
    if (warp == 0) {
      // ============================== TMA producer ==============================
      if (lane == 0) {
        mbar_arrive_expect_tx(q_full, 2 * FMHA_TILE_BYTES);
        for (int t = 0; t < 2; ++t)
          for (int h = 0; h < 2; ++h)
            tma_load_3d(sQ + t * FMHA_TILE_BYTES + h * FMHA_PANEL_BYTES, &tmQ, q_full, h * 64, head, q0 + t * FMHA_BLOCK_Q);
        int stage = 0;
        uint32_t phase = 0;
        for (int it = 0; it < 2 * n_kv; ++it) {   // K_0, V_0, K_1, V_1, ...
          const int j = it >> 1;
          const CUtensorMap* tm = (it & 1) ? &tmV : &tmK;
          mbar_wait(&kv_empty[stage], phase ^ 1);
          mbar_arrive_expect_tx(&kv_full[stage], FMHA_TILE_BYTES);
          for (int h = 0; h < 2; ++h)
            tma_load_3d(sKV + stage * FMHA_TILE_BYTES + h * FMHA_PANEL_BYTES, tm, &kv_full[stage], h * 64, head, j * FMHA_BLOCK_KV);
          if (++stage == FMHA_KV_STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    } else if (warp == 1) {
      // ============================== MMA issuer (whole warp, warp-uniform) ==============================
      constexpr uint32_t idesc_qk = make_idesc(FMT_BF16, FMT_BF16, 128, 128, 0, 0);
      constexpr uint32_t idesc_pv = make_idesc(FMT_BF16, FMT_BF16, 128, 128, 0, 1);
      const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint32_t tS0 = tb + COL_S0, tS1 = tb + COL_S1;
      const uint32_t tO0 = tb + COL_O0, tO1 = tb + COL_O1;
      const uint32_t q_lo = desc_lo_kmajor(smem_u32(sQ));
      const uint32_t kv_addr = smem_u32(sKV);

      auto issue_qk = [&](int t, int kstage) {
        const uint32_t a = q_lo + t * (FMHA_TILE_BYTES >> 4);
        const uint32_t b = desc_lo_kmajor(kv_addr + kstage * FMHA_TILE_BYTES);
        const uint32_t d = t ? tS1 : tS0;
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          const uint32_t off = ((ks >> 2) * FMHA_PANEL_BYTES + (ks & 3) * 32) >> 4;
          mma_f16_ss_w(d, a + off, kDescHiSw128, b + off, kDescHiSw128, idesc_qk, ks != 0 ? 1u : 0u);
        }
      };
      // half = 0: kv rows 0..63 (P columns 0..31), half = 1: kv rows 64..127
      auto issue_pv_half = [&](int t, int vstage, int half, uint32_t accumulate) {
        const uint32_t b = desc_lo_mnmajor(kv_addr + vstage * FMHA_TILE_BYTES, FMHA_PANEL_BYTES);
        const uint32_t d = t ? tO1 : tO0;
        const uint32_t a = t ? tS1 : tS0;
#pragma unroll
        for (int k4 = 0; k4 < 4; ++k4) {
          const int ks = half * 4 + k4;
          mma_f16_ts_w(d, a + ks * 8, b + ks * (2048 >> 4), kDescHiSw128, idesc_pv, (half | k4) != 0 ? 1u : accumulate);
        }
      };

      int stage = 0;
      uint32_t phase = 0;
      auto advance = [&]() {
        if (++stage == FMHA_KV_STAGES) {
          stage = 0;
          phase ^= 1;
        }
      };

      mbar_wait(q_full, 0);
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
        const int vstage = stage;
        mbar_wait(&kv_full[stage], phase);
        advance();
        const int kstage = stage;
        if (has_next) {
          mbar_wait(&kv_full[stage], phase);
          advance();
        }
        const uint32_t pj = j & 1;
        const uint32_t acc = j > 0 ? 1u : 0u;
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          mbar_wait(&p_half[t * 2 + 0], pj);
          tc_fence_after();
          issue_pv_half(t, vstage, 0, acc);
          mbar_wait(&p_half[t * 2 + 1], pj);
          tc_fence_after();
          issue_pv_half(t, vstage, 1, acc);
          if (t == 1) tc_commit_w(&kv_empty[vstage]);
          if (has_next) {
            issue_qk(t, kstage);
            tc_commit_w(&s_full[t]);
            if (t == 1) tc_commit_w(&kv_empty[kstage]);
          } else {
            tc_commit_w(&o_full[t]);
          }
        }
      }
    }
  } else {
    // ============================== softmax / correction / epilogue ==============================
    reg_alloc<104>();
    const int sw = warp - 4;
    const int t = sw >> 3;                     // query tile
    const int hf = (sw >> 2) & 1;              // column half owned by this thread
    const int lg = warp & 3;                   // TMEM lane group
    const int row = lg * 32 + lane;
    const uint32_t lane_off = uint32_t(lg * 32) << 16;
    const uint32_t tS = tmem_base + (t ? COL_S1 : COL_S0) + lane_off;
    const uint32_t tSh = tS + hf * 64;         // my 64 score columns
    const uint32_t tPh = tS + hf * 32;         // my 32 packed P columns
    const uint32_t tOh = tmem_base + (t ? COL_O1 : COL_O0) + lane_off + hf * 64;
    const float sl2 = p.scale_log2;
    const uint32_t bar_id = 2 + t;

    float m_used = -INFINITY;
    float l_sum = 0.f;                          // partial: my 64 columns only

    for (int j = 0; j < n_kv; ++j) {
      mbar_wait(&s_full[t], j & 1);
      tc_fence_after();
      uint32_t s[64];
      tmem_ld_x32(tSh, s);
      tmem_ld_x32(tSh + 32, s + 32);
      tmem_ld_wait();

      const int kv_valid = p.sk - j * FMHA_BLOCK_KV - hf * 64;   // valid columns among my 64 (may be <= 0)
      if (kv_valid < 64) {
#pragma unroll
        for (int i = 0; i < 64; ++i)
          if (i >= kv_valid) s[i] = 0xff800000u;  // -inf
      }
      float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
      for (int i = 0; i < 64; i += 2) {
        mx0 = fmaxf(mx0, __uint_as_float(s[i]));
        mx1 = fmaxf(mx1, __uint_as_float(s[i + 1]));
      }
      const float mloc = fmaxf(mx0, mx1);
      float* xm = xmax + (((j & 1) * 2 + t) * 2) * 128;
      xm[hf * 128 + row] = mloc;
      named_bar_sync(bar_id, 256);             // also: every S column of this tile has been read -> P may overwrite it
      const float m_new = fmaxf(fmaxf(mloc, xm[(hf ^ 1) * 128 + row]), m_used);

      if (j == 0) {
        m_used = m_new;
      } else {
        const bool need = (m_new - m_used) * sl2 > 8.0f;
        if (__any_sync(0xffffffffu, need)) {
          float alpha = 1.0f;
          if (need) {
            alpha = ex2((m_used - m_new) * sl2);
            m_used = m_new;
          }
          l_sum *= alpha;
#pragma unroll 1
          for (int c = 0; c < 64; c += 32) {
            uint32_t o[32];
            tmem_ld_x32(tOh + c, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st_x32(tOh + c, o);
          }
          tmem_st_wait();
        }
      }

      const float2 sl2v = make_float2(sl2, sl2);
      const float neg_m = -m_used * sl2;
      const float2 negm = make_float2(neg_m, neg_m);
      float2 acc[4] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
      uint32_t pk[32];
#pragma unroll
      for (int g = 0; g < 8; ++g) {
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          const int i = g * 4 + jj;
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
      const float2 a01 = __fadd2_rn(acc[0], acc[1]), a23 = __fadd2_rn(acc[2], acc[3]);
      l_sum += (a01.x + a01.y) + (a23.x + a23.y);
      tmem_st_x32(tPh, pk);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&p_half[t * 2 + hf]);
    }

    // ---- epilogue: exchange the partial row sums, then O / l -> bf16 -> global (my 64 of the 128 head-dim columns)
    xsum[(t * 2 + hf) * 128 + row] = l_sum;
    named_bar_sync(bar_id, 256);
    const float inv_l = 1.0f / (l_sum + xsum[(t * 2 + (hf ^ 1)) * 128 + row]);
    mbar_wait(&o_full[t], 0);
    tc_fence_after();
    const int q_row = q0 + t * FMHA_BLOCK_Q + row;
    __nv_bfloat16* orow = (q_row < p.sq ? fmha_out_row(p, q_row, head) : p.out) + hf * 64;
#pragma unroll 1
    for (int c = 0; c < 64; c += 32) {
      uint32_t o[32];
      tmem_ld_x32(tOh + c, o);
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

// =====================================================================================================================
// v6: 64-row K/V tiles with S and P in SEPARATE TMEM columns.
// With 128-row K/V tiles the TMEM budget (S0|S1|O0|O1 = 512 columns) forces P to alias S, so QK_{j+1} of a query tile cannot be
// issued before PV_j has consumed P_j: the per-tile loop is serial (softmax -> PV -> QK -> softmax) and neither the tensor pipe
// (72 %) nor the softmax warps (60 % busy) saturate.  With 64-row K/V tiles S_t (64 columns), P_t (32 columns) and O_t (128
// columns) all fit (448 of 512 columns) without aliasing: the softmax warps hand S back right after loading it into registers
// (s_free), so the tensor pipe runs QK_{j+1} while the exponentials of step j are still being computed, and PV_j as soon as
// P_j lands.  The chain per tile and step is max(softmax, MMA) instead of their sum.
//   TMEM: S0 @0, S1 @64, P0 @128, P1 @160, O0 @256, O1 @384.   smem: Q 2 x 32 KB, K/V ring 8 x 16 KB ([64 rows][128 d] tiles).
//   barriers per tile: s_full (MMA->softmax), s_free (softmax->MMA), p_full (softmax->MMA), p_free (MMA->softmax), o_full.
// =====================================================================================================================
constexpr int FMHA6_BLOCK_KV = 64;
constexpr int FMHA6_KV_STAGES = 8;
constexpr int FMHA6_KV_TILE_BYTES = 64 * 128 * 2;      // 16 KB: two [64 rows][64 d] swizzled panels of 8 KB
constexpr int FMHA6_KV_PANEL_BYTES = 64 * 64 * 2;      // 8 KB
constexpr int FMHA6_SMEM_BYTES = 2 * FMHA_TILE_BYTES + FMHA6_KV_STAGES * FMHA6_KV_TILE_BYTES + 1024 + 512;

template <int kPolyPairs>
__global__ void __launch_bounds__(FMHA_THREADS, 1)
fmha_fwd_d128_v6_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                        const __grid_constant__ CUtensorMap tmV, const FmhaParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sKV = smem + 2 * FMHA_TILE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sKV + FMHA6_KV_STAGES * FMHA6_KV_TILE_BYTES);
  uint64_t* q_full = bars;                  // [1]
  uint64_t* kv_full = bars + 1;             // [8]
  uint64_t* kv_empty = bars + 9;            // [8]
  uint64_t* s_full = bars + 17;             // [2]
  uint64_t* s_free = bars + 19;             // [2]
  uint64_t* p_full = bars + 21;             // [2]
  uint64_t* p_free = bars + 23;             // [2]
  uint64_t* o_full = bars + 25;             // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 27);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int head = blockIdx.y;
  const int q0 = blockIdx.x * (2 * FMHA_BLOCK_Q);
  const int n_kv = (p.sk + FMHA6_BLOCK_KV - 1) / FMHA6_BLOCK_KV;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmQ);
    prefetch_tmap(&tmK);
    prefetch_tmap(&tmV);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < FMHA6_KV_STAGES; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    for (int t = 0; t < 2; ++t) {
      mbar_init(&s_full[t], 1);
      mbar_init(&s_free[t], 128);
      mbar_init(&p_full[t], 128);
      mbar_init(&p_free[t], 1);
      mbar_init(&o_full[t], 1);
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  constexpr uint32_t COL_S0 = 0, COL_S1 = 64, COL_P0 = 128, COL_P1 = 160, COL_O0 = 256, COL_O1 = 384;

  if (warp < 4) {
    reg_dealloc<88>();
    if (warp == 0) {
      // ============================== TMA producer ==============================
      if (lane == 0) {
        mbar_arrive_expect_tx(q_full, 2 * FMHA_TILE_BYTES);
        for (int t = 0; t < 2; ++t)
          for (int h = 0; h < 2; ++h)
            tma_load_3d(sQ + t * FMHA_TILE_BYTES + h * FMHA_PANEL_BYTES, &tmQ, q_full, h * 64, head, q0 + t * FMHA_BLOCK_Q);
        int stage = 0;
        uint32_t phase = 0;
        for (int it = 0; it < 2 * n_kv; ++it) {   // K_0, V_0, K_1, V_1, ...
          const int j = it >> 1;
          const CUtensorMap* tm = (it & 1) ? &tmV : &tmK;
          mbar_wait(&kv_empty[stage], phase ^ 1);
          mbar_arrive_expect_tx(&kv_full[stage], FMHA6_KV_TILE_BYTES);
          for (int h = 0; h < 2; ++h)
            tma_load_3d(sKV + stage * FMHA6_KV_TILE_BYTES + h * FMHA6_KV_PANEL_BYTES, tm, &kv_full[stage], h * 64, head,
                        j * FMHA6_BLOCK_KV);
          if (++stage == FMHA6_KV_STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    } else if (warp == 1) {
      // ============================== MMA issuer (whole warp, warp-uniform) ==============================
      constexpr uint32_t idesc_qk = make_idesc(FMT_BF16, FMT_BF16, 128, 64, 0, 0);    // S[128 q, 64 kv]
      constexpr uint32_t idesc_pv = make_idesc(FMT_BF16, FMT_BF16, 128, 128, 0, 1);   // O[128 q, 128 d], V MN-major
      const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint32_t q_lo = desc_lo_kmajor(smem_u32(sQ));
      const uint32_t kv_addr = smem_u32(sKV);

      auto issue_qk = [&](int t, int kstage) {
        const uint32_t a = q_lo + t * (FMHA_TILE_BYTES >> 4);
        const uint32_t b = desc_lo_kmajor(kv_addr + kstage * FMHA6_KV_TILE_BYTES);
        const uint32_t d = tb + (t ? COL_S1 : COL_S0);
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {   // d = 128 in steps of 16: Q panels are 16 KB apart, K panels 8 KB apart
          const uint32_t offq = ((ks >> 2) * FMHA_PANEL_BYTES + (ks & 3) * 32) >> 4;
          const uint32_t offk = ((ks >> 2) * FMHA6_KV_PANEL_BYTES + (ks & 3) * 32) >> 4;
          mma_f16_ss_w(d, a + offq, kDescHiSw128, b + offk, kDescHiSw128, idesc_qk, ks != 0 ? 1u : 0u);
        }
      };
      auto issue_pv = [&](int t, int vstage, uint32_t accumulate) {
        const uint32_t b = desc_lo_mnmajor(kv_addr + vstage * FMHA6_KV_TILE_BYTES, FMHA6_KV_PANEL_BYTES);
        const uint32_t d = tb + (t ? COL_O1 : COL_O0);
        const uint32_t a = tb + (t ? COL_P1 : COL_P0);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {   // kv = 64 in steps of 16 rows (2048 B); P: 8 columns per step
          mma_f16_ts_w(d, a + ks * 8, b + ks * (2048 >> 4), kDescHiSw128, idesc_pv, ks != 0 ? 1u : accumulate);
        }
      };

      int stage = 0;
      uint32_t phase = 0;
      auto advance = [&]() {
        if (++stage == FMHA6_KV_STAGES) {
          stage = 0;
          phase ^= 1;
        }
      };

      mbar_wait(q_full, 0);
      mbar_wait(&kv_full[stage], phase);           // K_0
      tc_fence_after();
      issue_qk(0, stage);
      tc_commit_w(&s_full[0]);
      issue_qk(1, stage);
      tc_commit_w(&s_full[1]);
      tc_commit_w(&kv_empty[stage]);
      advance();

      for (int j = 0; j < n_kv; ++j) {
        const bool has_next = (j + 1) < n_kv;
        const int vstage = stage;
        const uint32_t vphase = phase;
        advance();
        const int kstage = stage;
        const uint32_t kphase = phase;
        if (has_next) advance();
        const uint32_t pj = j & 1;
        const uint32_t acc = j > 0 ? 1u : 0u;
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          if (has_next) {
            if (t == 0) mbar_wait(&kv_full[kstage], kphase);      // K_{j+1}
            mbar_wait(&s_free[t], pj);                             // S_j of this tile is in the softmax warps' registers
            tc_fence_after();
            issue_qk(t, kstage);
            tc_commit_w(&s_full[t]);
            if (t == 1) tc_commit_w(&kv_empty[kstage]);
          }
          if (t == 0) mbar_wait(&kv_full[vstage], vphase);        // V_j
          mbar_wait(&p_full[t], pj);
          tc_fence_after();
          issue_pv(t, vstage, acc);
          tc_commit_w(&p_free[t]);
          if (!has_next) tc_commit_w(&o_full[t]);
          if (t == 1) tc_commit_w(&kv_empty[vstage]);
        }
      }
    }
  } else {
    // ============================== softmax / correction / epilogue ==============================
    reg_alloc<208>();
    const int t = (warp - 4) >> 2;
    const int lg = warp & 3;
    const int row = lg * 32 + lane;
    const uint32_t lane_off = uint32_t(lg * 32) << 16;
    const uint32_t tS = tmem_base + (t ? COL_S1 : COL_S0) + lane_off;
    const uint32_t tP = tmem_base + (t ? COL_P1 : COL_P0) + lane_off;
    const uint32_t tO = tmem_base + (t ? COL_O1 : COL_O0) + lane_off;
    const float sl2 = p.scale_log2;
    const float2 sl2v = make_float2(sl2, sl2);

    float m_used = -INFINITY;
    float l_sum = 0.f;

    for (int j = 0; j < n_kv; ++j) {
      mbar_wait(&s_full[t], j & 1);
      tc_fence_after();
      uint32_t s[64];
      tmem_ld_x32(tS, s);
      tmem_ld_x32(tS + 32, s + 32);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(&s_free[t]);                       // the tensor pipe may overwrite S with QK_{j+1} from here on

      const int kv_valid = p.sk - j * FMHA6_BLOCK_KV;
      if (kv_valid < FMHA6_BLOCK_KV) {
#pragma unroll
        for (int i = 0; i < 64; ++i)
          if (i >= kv_valid) s[i] = 0xff800000u;
      }
      float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
      for (int i = 0; i < 64; i += 4) {
        mx0 = fmaxf(mx0, __uint_as_float(s[i]));
        mx1 = fmaxf(mx1, __uint_as_float(s[i + 1]));
        mx2 = fmaxf(mx2, __uint_as_float(s[i + 2]));
        mx3 = fmaxf(mx3, __uint_as_float(s[i + 3]));
      }
      const float m_new = fmaxf(fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)), m_used);
      float alpha = 1.0f;
      bool any_need = false;
      if (j == 0) {
        m_used = m_new;
      } else {
        const bool need = (m_new - m_used) * sl2 > 8.0f;     // lazy rescale threshold: 2^8 in the exp2 domain
        any_need = __any_sync(0xffffffffu, need);
        if (need) {
          alpha = ex2((m_used - m_new) * sl2);
          m_used = m_new;
        }
        l_sum *= alpha;
      }

      const float neg_m = -m_used * sl2;
      const float2 negm = make_float2(neg_m, neg_m);
      float2 acc[4] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
      uint32_t pk[32];
#pragma unroll
      for (int g = 0; g < 8; ++g) {
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          const int i = g * 4 + jj;
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
      const float2 a01 = __fadd2_rn(acc[0], acc[1]), a23 = __fadd2_rn(acc[2], acc[3]);
      l_sum += (a01.x + a01.y) + (a23.x + a23.y);

      // P_{j-1} must have been consumed (and O be quiescent) before P_j is written / O is rescaled
      if (j > 0) {
        mbar_wait(&p_free[t], (j - 1) & 1);
        tc_fence_after();
        if (any_need) {
#pragma unroll 1
          for (int c = 0; c < 128; c += 32) {
            uint32_t o[32];
            tmem_ld_x32(tO + c, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st_x32(tO + c, o);
          }
        }
      }
      tmem_st_x32(tP, pk);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&p_full[t]);
    }

    // ---- epilogue: O / l -> bf16 -> global (or the token owner's peer buffer)
    mbar_wait(&o_full[t], 0);
    tc_fence_after();
    const float inv_l = 1.0f / l_sum;
    const int q_row = q0 + t * FMHA_BLOCK_Q + row;
    __nv_bfloat16* orow = q_row < p.sq ? fmha_out_row(p, q_row, head) : p.out;
#pragma unroll 1
    for (int c = 0; c < 128; c += 32) {
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

// q/k/v: [rows, H, 128] bf16 with arbitrary row stride (elements), heads contiguous (stride 128).
static int encode_qkv_map(CUtensorMap* tm, const void* base, long long rows, int heads, long long stride_s, uint32_t box_rows = 128) {
  uint64_t dims[3] = {128, (uint64_t)heads, (uint64_t)rows};
  uint64_t strides[2] = {128 * 2, (uint64_t)stride_s * 2};
  uint32_t box[3] = {64, 1, box_rows};
  return encode_tmap(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, base, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
}

int fmha_fwd_d128_impl(const void* q, long long q_stride_s, const void* k, long long k_stride_s, const void* v,
                       long long v_stride_s, void* out, long long o_stride_s, long long sq, long long sk, int heads,
                       float softmax_scale, void* const* peers, int world, long long rows_per_rank, int head_offset,
                       cudaStream_t stream) {
  if (peers != nullptr) {
    B200_CHECK_ARG(world >= 1 && world <= 8 && rows_per_rank > 0 && sq <= world * rows_per_rank,
                   "b200_fmha_fwd_d128_scatter: bad world %d / rows_per_rank %lld for sq %lld", world, rows_per_rank, sq);
    for (int i = 0; i < world; ++i) B200_CHECK_ARG(peers[i] != nullptr, "b200_fmha_fwd_d128_scatter: null peer pointer %d", i);
    out = peers[0];
  }
  B200_CHECK_ARG(q && k && v && out, "b200_fmha_fwd_d128: null pointer");
  B200_CHECK_ARG(sq > 0 && sk > 0 && heads > 0, "b200_fmha_fwd_d128: empty problem sq=%lld sk=%lld heads=%d", sq, sk,
                 heads);
  B200_CHECK_ARG(q_stride_s % 8 == 0 && k_stride_s % 8 == 0 && v_stride_s % 8 == 0 && o_stride_s % 8 == 0,
                 "b200_fmha_fwd_d128: row strides must be multiples of 8 elements");
  B200_CHECK_ARG(q_stride_s >= heads * 128LL && k_stride_s >= heads * 128LL && v_stride_s >= heads * 128LL &&
                     o_stride_s >= heads * 128LL,
                 "b200_fmha_fwd_d128: row stride smaller than heads*128");
  B200_CHECK_ARG(((uintptr_t)q % 16 == 0) && ((uintptr_t)k % 16 == 0) && ((uintptr_t)v % 16 == 0) &&
                     ((uintptr_t)out % 16 == 0),
                 "b200_fmha_fwd_d128: pointers must be 16-byte aligned");
  CUtensorMap tmQ, tmK, tmV;
  int rc;
  if ((rc = encode_qkv_map(&tmQ, q, sq, heads, q_stride_s))) return rc;
  static int ver = -1;
  if (ver < 0) {
    const char* e = getenv("B200_FMHA_VER");
    ver = e ? atoi(e) : 4;     // 2: one softmax warpgroup per tile; 3: two per tile; 4: v2 + split-P publication; 5: 4 + speculative max; 6: 64-row K/V tiles, S/P un-aliased
    if (ver < 2 || ver > 7) ver = 4;
  }
  const uint32_t kv_box_rows = (ver == 6) ? 64 : 128;
  if ((rc = encode_qkv_map(&tmK, k, sk, heads, k_stride_s, kv_box_rows))) return rc;
  if ((rc = encode_qkv_map(&tmV, v, sk, heads, v_stride_s, kv_box_rows))) return rc;

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

  // fraction of exp2 evaluated by the FMA-pipe polynomial: B200_FMHA_POLY = 0..3 pairs out of every 4 (default 1)
  static int poly = -1;
  if (poly < 0) {
    const char* e = getenv("B200_FMHA_POLY");
    poly = e ? atoi(e) : 1;   // measured on B200: 0 -> 1.19, 1 -> 1.23, 2 -> 1.13 PFLOP/s at S = 75 600 x 40 heads
    if (poly < 0 || poly > 3) poly = 1;
  }
  dim3 grid((unsigned)((sq + 2 * FMHA_BLOCK_Q - 1) / (2 * FMHA_BLOCK_Q)), (unsigned)heads, 1);
  auto launch = [&](auto kern, int threads, int smem_bytes) -> int {
    B200_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    kern<<<grid, threads, smem_bytes, stream>>>(tmQ, tmK, tmV, p);
    return B200_OK;
  };
  int rc2;
  if (ver == 6) {
    switch (poly) {
      case 0: rc2 = launch(fmha_fwd_d128_v6_kernel<0>, FMHA_THREADS, FMHA6_SMEM_BYTES); break;
      case 2: rc2 = launch(fmha_fwd_d128_v6_kernel<2>, FMHA_THREADS, FMHA6_SMEM_BYTES); break;
      default: rc2 = launch(fmha_fwd_d128_v6_kernel<1>, FMHA_THREADS, FMHA6_SMEM_BYTES); break;
    }
  } else if (ver == 3) {
    switch (poly) {
      case 0: rc2 = launch(fmha_fwd_d128_v3_kernel<0>, FMHA3_THREADS, FMHA3_SMEM_BYTES); break;
      case 2: rc2 = launch(fmha_fwd_d128_v3_kernel<2>, FMHA3_THREADS, FMHA3_SMEM_BYTES); break;
      case 3: rc2 = launch(fmha_fwd_d128_v3_kernel<3>, FMHA3_THREADS, FMHA3_SMEM_BYTES); break;
      default: rc2 = launch(fmha_fwd_d128_v3_kernel<1>, FMHA3_THREADS, FMHA3_SMEM_BYTES); break;
    }
  } else if (ver == 7) {
    switch (poly) {
      case 0: rc2 = launch(fmha_fwd_d128_kernel<0, true, false, true>, FMHA_THREADS, FMHA_SMEM_BYTES); break;
      case 2: rc2 = launch(fmha_fwd_d128_kernel<2, true, false, true>, FMHA_THREADS, FMHA_SMEM_BYTES); break;
      default: rc2 = launch(fmha_fwd_d128_kernel<1, true, false, true>, FMHA_THREADS, FMHA_SMEM_BYTES); break;
    }
  } else if (ver == 5) {
    switch (poly) {
      case 0: rc2 = launch(fmha_fwd_d128_kernel<0, true, true, false>, FMHA_THREADS, FMHA_SMEM_BYTES); break;
      case 2: rc2 = launch(fmha_fwd_d128_kernel<2, true, true, false>, FMHA_THREADS, FMHA_SMEM_BYTES); break;
      default: rc2 = launch(fmha_fwd_d128_kernel<1, true, true, false>, FMHA_THREADS, FMHA_SMEM_BYTES); break;
    }
  } else if (ver == 4) {
    switch (poly) {
      case 0: rc2 = launch(fmha_fwd_d128_kernel<0, true, false, false>, FMHA_THREADS, FMHA_SMEM_BYTES); break;
      case 2: rc2 = launch(fmha_fwd_d128_kernel<2, true, false, false>, FMHA_THREADS, FMHA_SMEM_BYTES); break;
      default: rc2 = launch(fmha_fwd_d128_kernel<1, true, false, false>, FMHA_THREADS, FMHA_SMEM_BYTES); break;
    }
  } else {
    switch (poly) {
      case 0: rc2 = launch(fmha_fwd_d128_kernel<0, false, false, false>, FMHA_THREADS, FMHA_SMEM_BYTES); break;
      case 2: rc2 = launch(fmha_fwd_d128_kernel<2, false, false, false>, FMHA_THREADS, FMHA_SMEM_BYTES); break;
      case 3: rc2 = launch(fmha_fwd_d128_kernel<3, false, false, false>, FMHA_THREADS, FMHA_SMEM_BYTES); break;
      default: rc2 = launch(fmha_fwd_d128_kernel<1, false, false, false>, FMHA_THREADS, FMHA_SMEM_BYTES); break;
    }
  }
  if (rc2) return rc2;
  B200_CHECK_CUDA(cudaGetLastError());
  return B200_OK;
}

int fmha_fwd_d128(const void* q, long long q_stride_s, const void* k, long long k_stride_s, const void* v,
                  long long v_stride_s, void* out, long long o_stride_s, long long sq, long long sk, int heads,
                  float softmax_scale, cudaStream_t stream) {
  return fmha_fwd_d128_impl(q, q_stride_s, k, k_stride_s, v, v_stride_s, out, o_stride_s, sq, sk, heads, softmax_scale, nullptr,
                            0, 0, 0, stream);
}

int fmha_fwd_d128_scatter(const void* q, long long q_stride_s, const void* k, long long k_stride_s, const void* v,
                          long long v_stride_s, void* const* peers, int world, long long rows_per_rank, long long peer_stride_s,
                          int head_offset, long long sq, long long sk, int heads, float softmax_scale, cudaStream_t stream) {
  B200_CHECK_ARG(peers != nullptr, "b200_fmha_fwd_d128_scatter: null peer table");
  return fmha_fwd_d128_impl(q, q_stride_s, k, k_stride_s, v, v_stride_s, nullptr, peer_stride_s, sq, sk, heads, softmax_scale, peers,
                            world, rows_per_rank, head_offset, stream);
}

}  // namespace b200
