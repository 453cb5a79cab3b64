// Thin inline-PTX layer for sm_100a: mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (MMA / TMEM
// alloc / ld / st / commit / fences), UMMA shared-memory + instruction descriptors.
// Everything here is hand-written against the PTX ISA; no CUTLASS/CuTe types are used.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200 {

// ----------------------------------------------------------------------------------------------
// misc
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.b32 %0, 1, 0, P;\n\t"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ uint32_t lane_id() {
  uint32_t l;
  asm volatile("mov.u32 %0, %%laneid;" : "=r"(l));
  return l;
}

template <int kRegs>
__device__ __forceinline__ void reg_alloc() {
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;\n" ::"n"(kRegs));
}
template <int kRegs>
__device__ __forceinline__ void reg_dealloc() {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;\n" ::"n"(kRegs));
}

__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void named_bar_arrive(uint32_t id, uint32_t nthreads) {
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ----------------------------------------------------------------------------------------------
// mbarrier
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// generic-proxy writes to smem -> visible to the async proxy (TMA store / UMMA reads)
__device__ __forceinline__ void fence_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
// arrive on the same-offset barrier of another CTA in the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t cta_rank) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(bar)), "r"(cta_rank));
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, P;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Blocking wait with a watchdog: a protocol bug traps (launch error) instead of hanging the GPU.
#ifndef B200_WATCHDOG_SPINS
#define B200_WATCHDOG_SPINS (1u << 24)
#endif
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > B200_WATCHDOG_SPINS) {
      printf("b200 watchdog: mbarrier timeout block(%d,%d,%d) thread %d bar@%u parity %u\n", blockIdx.x,
             blockIdx.y, blockIdx.z, threadIdx.x, smem_u32(bar), parity);
      __trap();
    }
  }
}

// ----------------------------------------------------------------------------------------------
// TMA
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* smem, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* smem, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(smem)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2),
      "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_5d(void* smem, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2,
                                            int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], "
      "[%2];" ::"r"(smem_u32(smem)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
// 2-CTA pair variant: data lands in this CTA's smem, complete_tx is signalled on the barrier whose
// cluster-shared address is given (the leader CTA's barrier).
__device__ __forceinline__ void tma_load_2d_2sm(void* smem, const CUtensorMap* m, uint32_t bar_cluster_addr, int c0,
                                                int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], "
      "[%2];" ::"r"(smem_u32(smem)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* smem, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(smem)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* m, const void* smem, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(smem)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void tma_store_wait_all() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}

// ----------------------------------------------------------------------------------------------
// tcgen05: TMEM management
// ----------------------------------------------------------------------------------------------
// One full warp calls these (.sync.aligned). ncols: power of two in [32, 512].
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// MMA completion -> mbarrier arrive (implicitly fences before_thread_sync). One thread.
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// 2-CTA: arrive on the same barrier offset in every CTA of cta_mask.
__device__ __forceinline__ void tc_commit_2sm(uint64_t* bar, uint16_t cta_mask) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(cta_mask)
      : "memory");
}

// ----------------------------------------------------------------------------------------------
// tcgen05.mma  (D[tmem] (+)= A * B), issued by ONE thread.
// ----------------------------------------------------------------------------------------------
// A and B from shared memory descriptors.
__device__ __forceinline__ void mma_f16_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// A from tensor memory, B from shared memory.
__device__ __forceinline__ void mma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mma_f16_ss_2sm(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mma_f8_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// Warp-collective issue forms: called by ALL 32 lanes of a converged warp with warp-uniform operands; one elected lane executes the
// MMA.  Descriptors are passed as (lo, hi) 32-bit halves.  Measured cost in the issuing warp (round-2 ncu source views): the
// single-MMA forms below compile to ELECT + two VOTEU + five predicated R2UR.BROADCAST + moves, 17-22 SASS instructions per UTCHMMA (a
// divergent `if (lane == 0)` around a plain MMA was ~100); the run forms further down (mma_f16_ss_w4 / _w2, mma_f8_ss_w4,
// mma_f16_ts_w4) are what the hot loops use.
__device__ __forceinline__ void mma_f16_ss_w(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                             uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p, e;\n\t"
      ".reg .b64 da, db;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
// A run of 4 (or 2) MMAs over consecutive K-steps of one swizzled 64-element chunk (descriptor start + 32 B per step) under ONE
// elect.sync: per MMA the single-MMA wrapper above costs ~22 SASS instructions in the issuing warp (ELECT, two VOTEU, five R2UR, moves,
// and - when guarded by `if (k < ksteps)` - a BSSY / BSYNC pair), ~100 clocks, more than a 128 x 96 x 16 MMA's 56-clock floor
// (profiles/r02_umma_rate_probe.txt, r02_halo96_v2 source view).  The first MMA takes the accumulate flag, the others accumulate.
#define B200_MMA_STEP(kind, off)                                                \
  "add.u32 al, %1, " #off ";\n\t"                                           \
  "add.u32 bl, %3, " #off ";\n\t"                                           \
  "mov.b64 da, {al, %2};\n\t"                                               \
  "mov.b64 db, {bl, %4};\n\t"                                               \
  "tcgen05.mma.cta_group::1.kind::" kind " [%0], da, db, %5, t;\n\t"
// an explicit branch around the run: ptxas then knows exactly one lane executes it and moves each operand to a uniform register once
// (plain R2UR + UIADD3 per step) instead of re-broadcasting every operand of every MMA under the election predicate
#define B200_MMA_HEAD(kind)                                                 \
  "{\n\t"                                                                   \
  ".reg .pred p, e, t;\n\t"                                                 \
  ".reg .b64 da, db;\n\t"                                                   \
  ".reg .b32 al, bl;\n\t"                                                   \
  "elect.sync _|e, 0xffffffff;\n\t"                                         \
  "@!e bra MMA_RUN_DONE;\n\t"                                               \
  "setp.ne.b32 p, %6, 0;\n\t"                                               \
  "setp.eq.b32 t, 0, 0;\n\t"                                                \
  "mov.b64 da, {%1, %2};\n\t"                                               \
  "mov.b64 db, {%3, %4};\n\t"                                               \
  "tcgen05.mma.cta_group::1.kind::" kind " [%0], da, db, %5, p;\n\t"
#define B200_MMA_TAIL "MMA_RUN_DONE:\n\t}\n"
__device__ __forceinline__ void mma_f16_ss_w4(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(B200_MMA_HEAD("f16") B200_MMA_STEP("f16", 2) B200_MMA_STEP("f16", 4) B200_MMA_STEP("f16", 6) B200_MMA_TAIL ::"r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi),
               "r"(idesc), "r"(accumulate)
               : "memory");
}
__device__ __forceinline__ void mma_f16_ss_w2(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(B200_MMA_HEAD("f16") B200_MMA_STEP("f16", 2) B200_MMA_TAIL ::"r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
               : "memory");
}
// fp8 (e4m3 x e4m3, K = 32 per MMA: the same 32-byte K-step inside the 128-byte swizzle row)
__device__ __forceinline__ void mma_f8_ss_w4(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(B200_MMA_HEAD("f8f6f4") B200_MMA_STEP("f8f6f4", 2) B200_MMA_STEP("f8f6f4", 4) B200_MMA_STEP("f8f6f4", 6) B200_MMA_TAIL ::"r"(d_tmem),
               "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
               : "memory");
}
#undef B200_MMA_STEP
#undef B200_MMA_HEAD
#undef B200_MMA_TAIL
__device__ __forceinline__ void mma_f8_ss_w(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                            uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p, e;\n\t"
      ".reg .b64 da, db;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], da, db, %5, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Block-scaled NVFP4 MMA (e2m1 x e2m1, one ue4m3 scale per 16 elements along K, K = 64 per instruction).  The scale factors
// live in TMEM (4 bytes = 4 K-blocks per 32-bit column; rows r + 32 q of the operand in lane r, column q), put there by tcgen05.cp.
__device__ __forceinline__ void mma_f4_bs_w(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                            uint32_t idesc, uint32_t sfa_tmem, uint32_t sfb_tmem, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p, e;\n\t"
      ".reg .b64 da, db;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "@e tcgen05.mma.cta_group::1.kind::mxf4nvf4.block_scale.scale_vec::4X [%0], da, db, %5, [%7], [%8], p;\n\t"
      "}\n" ::"r"(d_tmem),
      "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate), "r"(sfa_tmem), "r"(sfb_tmem)
      : "memory");
}
// smem -> TMEM copy of 32 rows x 16 bytes, replicated into the four 32-lane sub-partitions (scale-factor staging).
// Source: no-swizzle K-major core matrices (8 rows x 16 B contiguous), 8-row groups `SBO` bytes apart.
__device__ __forceinline__ void tmem_cp_32x128b_w(uint32_t taddr, uint32_t desc_lo, uint32_t desc_hi) {
  asm volatile(
      "{\n\t"
      ".reg .pred e;\n\t"
      ".reg .b64 d;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "mov.b64 d, {%1, %2};\n\t"
      "@e tcgen05.cp.cta_group::1.32x128b.warpx4 [%0], d;\n\t"
      "}\n" ::"r"(taddr),
      "r"(desc_lo), "r"(desc_hi)
      : "memory");
}
// descriptor hi word for the scale-factor copy: SBO = 128 B (8-row groups are contiguous), version 1, no swizzle
constexpr uint32_t kDescHiSfNoSwz = (128u >> 4) | (1u << 14);
// instruction descriptor of kind::mxf4nvf4 (cute InstrDescriptorBlockScaled): a/b format E2M1 = 1, scale format ue4m3 = 0, K = 64
__host__ __device__ constexpr uint32_t make_idesc_f4(uint32_t M, uint32_t N) {
  return (1u << 7) | (1u << 10) | ((N >> 3) << 17) | (0u << 23) | ((M >> 4) << 24);
}
__device__ __forceinline__ void mma_f16_ts_w(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p, e;\n\t"
      ".reg .b64 db;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "mov.b64 db, {%2, %3};\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "r"(a_tmem), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Four TS-mode MMAs (A = P in TMEM, 8 packed columns per 16-row K-step; B = V MN-major, 2048 B per K-step) under one election with an
// explicit branch - see mma_f16_ss_w4.
__device__ __forceinline__ void mma_f16_ts_w4(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p, e, t;\n\t"
      ".reg .b64 db;\n\t"
      ".reg .b32 at, bl;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "@!e bra MMA_TS_RUN_DONE;\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "setp.eq.b32 t, 0, 0;\n\t"
      "mov.b64 db, {%2, %3};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;\n\t"
      "add.u32 at, %1, 8;\n\t"
      "add.u32 bl, %2, 128;\n\t"
      "mov.b64 db, {bl, %3};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [at], db, %4, t;\n\t"
      "add.u32 at, %1, 16;\n\t"
      "add.u32 bl, %2, 256;\n\t"
      "mov.b64 db, {bl, %3};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [at], db, %4, t;\n\t"
      "add.u32 at, %1, 24;\n\t"
      "add.u32 bl, %2, 384;\n\t"
      "mov.b64 db, {bl, %3};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [at], db, %4, t;\n\t"
      "MMA_TS_RUN_DONE:\n\t"
      "}\n" ::"r"(d_tmem),
      "r"(a_tmem), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_commit_w(uint64_t* bar) {
  asm volatile(
      "{\n\t"
      ".reg .pred e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t"
      "}\n" ::"r"(smem_u32(bar))
      : "memory");
}
// High 32 bits of a swizzle-128B shared-memory descriptor: SBO = 1024 B, version 1, layout SWIZZLE_128B.
constexpr uint32_t kDescHiSw128 = (1024u >> 4) | (1u << 14) | (2u << 29);
// Low 32 bits: (addr >> 4) | LBO field.  K-major: LBO ignored (1).  MN-major: LBO = byte distance between 64-element panels.
__device__ __forceinline__ uint32_t desc_lo_kmajor(uint32_t smem_addr) { return ((smem_addr & 0x3FFFF) >> 4) | (1u << 16); }
__device__ __forceinline__ uint32_t desc_lo_mnmajor(uint32_t smem_addr, uint32_t lbo_bytes) {
  return ((smem_addr & 0x3FFFF) >> 4) | ((lbo_bytes >> 4) << 16);
}

// ----------------------------------------------------------------------------------------------
// Descriptors
// ----------------------------------------------------------------------------------------------
// Instruction descriptor for kind::f16 / kind::tf32 / kind::f8f6f4 (dense, fp32 accumulate).
//   [4,6) c_format (1 = f32) | [7,10) a_format | [10,13) b_format | 15 a_major | 16 b_major (1 = MN-major)
//   [17,23) N>>3 | [24,29) M>>4
enum : uint32_t { FMT_F16 = 0, FMT_BF16 = 1, FMT_TF32 = 2, FMT_E4M3 = 0, FMT_E5M2 = 1 };
__host__ __device__ constexpr uint32_t make_idesc(uint32_t a_fmt, uint32_t b_fmt, uint32_t M, uint32_t N,
                                                  uint32_t a_mn_major, uint32_t b_mn_major) {
  return (1u << 4) | (a_fmt << 7) | (b_fmt << 10) | (a_mn_major << 15) | (b_mn_major << 16) | ((N >> 3) << 17) |
         ((M >> 4) << 24);
}

// Shared-memory matrix descriptor (64-bit):
//   [0,14) start>>4 | [16,30) LBO>>4 | [32,46) SBO>>4 | [46,48) version=1 | [49,52) base_offset | [61,64) swizzle
// 128-byte swizzle = 2.  The tile must be 1024-byte aligned (base_offset = 0).
//
// K-major operand, rows of exactly 128 bytes (64 bf16 / 128 fp8 along K), 8-row groups 1024 B apart:
//   SBO = 1024, LBO ignored (1).  Advancing K by one UMMA_K (32 bytes) = start address + 32 bytes.
__device__ __forceinline__ uint64_t make_desc_kmajor_sw128(uint32_t smem_addr) {
  return (uint64_t)((smem_addr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// MN-major operand (the contiguous axis is M/N): smem holds [K rows][64 elements = 128 B] swizzle-128B panels,
// 8-K-row groups 1024 B apart (SBO), successive 64-element MN panels `lbo_bytes` apart (LBO).
// Advancing K by one UMMA_K (16 rows of 128 B) = start address + 2048 bytes.
__device__ __forceinline__ uint64_t make_desc_mnmajor_sw128(uint32_t smem_addr, uint32_t lbo_bytes) {
  return (uint64_t)((smem_addr & 0x3FFFF) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) |
         ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}

// ----------------------------------------------------------------------------------------------
// tcgen05.ld / st, shape 32x32b: warp w (w % 4 selects TMEM lanes 32*(w%4) ..+31), thread = lane (row),
// N consecutive 32-bit columns per thread.
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld_x16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_x32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,"
      "%30,%31}, [%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_st_x16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};\n" ::"r"(
          taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_x32(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,"
      "%31,%32};\n" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
      "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
      "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
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

__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// SiLU t / (1 + e^-t) sized for HBM-bound element-wise passes.  At 4 bytes of traffic per element a B200 SM must retire ~9 elements per
// clock, i.e. the whole per-element instruction budget is ~14 issue slots; `t / (1.f + __expf(-t))` alone costs ~11 (the IEEE division
// expands into a Newton iteration with a slow path) plus two MUFU operations against a MUFU rate of 16 / clk / SM.
//   silu_fast : ex2.approx + rcp.approx, 5 instructions, 2 MUFU (relative error ~2^-21, far below the 2^-9 of the bf16 result);
//               saturates correctly: t -> -inf gives t * rcp(inf) = -0, t -> +inf gives t * rcp(1) = t.
//   silu2_poly: the exponential on the FMA pipe (exp2_poly2, relative error 7.5e-5), one MUFU per element; callers mix the two so that
//               the MUFU stays below its rate while the issue slots stay below theirs.
__device__ __forceinline__ float silu_fast(float t) { return t * rcp_approx(1.0f + ex2_approx(t * -1.4426950408889634f)); }
__device__ __forceinline__ float2 silu2_fast(float2 t) { return make_float2(silu_fast(t.x), silu_fast(t.y)); }
__device__ __forceinline__ float2 silu2_poly(float2 t) {
  float2 x = __fmul2_rn(t, make_float2(-1.4426950408889634f, -1.4426950408889634f));
  x.x = fminf(x.x, 126.f);
  x.y = fminf(x.y, 126.f);
  const float2 e = exp2_poly2(x);
  return make_float2(t.x * rcp_approx(1.0f + e.x), t.y * rcp_approx(1.0f + e.y));
}
// 8 values: one pair through the polynomial, three through the MUFU (1.75 MUFU operations per element)
__device__ __forceinline__ void silu8(float* f) {
  const float2 a = silu2_poly(make_float2(f[0], f[1]));
  f[0] = a.x; f[1] = a.y;
#pragma unroll
  for (int e = 2; e < 8; ++e) f[e] = silu_fast(f[e]);
}

// streaming 16-byte load that does not allocate in L1 (read-once data of the HBM-bound passes)
__device__ __forceinline__ uint4 ld_nc_v4(const __nv_bfloat16* p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}

// ----------------------------------------------------------------------------------------------
// bf16 helpers
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float bf16_lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t v) { return __uint_as_float(v & 0xFFFF0000u); }
__device__ __forceinline__ float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

}  // namespace b200
