// Hardware probe (debug entry point, not on any product path): does a K-major swizzled UMMA operand whose descriptor START ADDRESS is
// shifted by whole rows inside a TMA-written tile read the right data?  This decides whether a convolution can load one halo row of
// voxels once and take its (dw = -1, 0, +1) taps as row-shifted views of the same shared-memory tile (conv3d.cu round-2 notes).
//
//   A_full : [rows_a = 128 + 8, K] bf16 in global memory, loaded as ONE TMA box into shared memory (swizzle 128B for K = 64 elements per
//            row, swizzle 64B for K = 32), tile base 1024-byte aligned.
//   B      : [N = 64, K] bf16, one TMA box.
//   D[128, 64] = A_full[r0 : r0 + 128, :] * B^T  with the A descriptor start address = tile base + r0 * row_bytes and the descriptor's
//   base_offset field = (mode == 1) ? (r0 % 8) : 0.
// The host compares D with torch for r0 = 0..7 and both modes.
#include "host_util.cuh"
#include "ptx.cuh"

namespace b200 {

template <int KB>   // bytes per row: 128 (SW128) or 64 (SW64)
__global__ void __launch_bounds__(128, 1)
umma_rowshift_probe_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, float* __restrict__ D, int r0, int mode) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int ROWS_A = 136;
  uint8_t* sA = smem;                                   // 136 rows
  uint8_t* sB = smem + 18 * 1024;                       // 64 rows
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 28 * 1024);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 64);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(&bars[0], ROWS_A * KB + 64 * KB);
    tma_load_2d(sA, &tmA, &bars[0], 0, 0);
    tma_load_2d(sB, &tmB, &bars[0], 0, 0);
  }
  if (warp == 0) {
    mbar_wait(&bars[0], 0);
    tc_fence_after();
    constexpr uint32_t idesc = make_idesc(FMT_BF16, FMT_BF16, 128, 64, 0, 0);
    constexpr uint32_t hi_base = KB == 128 ? kDescHiSw128 : ((512u >> 4) | (1u << 14) | (4u << 29));
    const uint32_t a_addr = smem_u32(sA) + r0 * KB;
    const uint32_t hi_a = hi_base | (mode == 1 ? (uint32_t(r0 & 7) << 17) : 0u);       // base_offset: descriptor bits [49, 52) = hi word bits [17, 20)
    const uint32_t a_lo = desc_lo_kmajor(a_addr), b_lo = desc_lo_kmajor(smem_u32(sB));
#pragma unroll
    for (int k = 0; k < KB / 32; ++k) mma_f16_ss_w(tmem_base, a_lo + 2 * k, hi_a, b_lo + 2 * k, hi_base, idesc, k != 0 ? 1u : 0u);
    tc_commit_w(&bars[1]);
  }
  mbar_wait(&bars[1], 0);
  tc_fence_after();
  uint32_t v[64];
  const uint32_t t_row = tmem_base + (uint32_t(warp * 32) << 16);
  tmem_ld_x32(t_row, v);
  tmem_ld_x32(t_row + 32, v + 32);
  tmem_ld_wait();
  const int row = warp * 32 + lane;
#pragma unroll
  for (int j = 0; j < 64; ++j) D[row * 64 + j] = __uint_as_float(v[j]);
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 64);
  }
}

int debug_umma_rowshift(const void* A, const void* B, float* D, int k_elems, int r0, int mode, cudaStream_t stream) {
  B200_CHECK_ARG(A && B && D && (k_elems == 64 || k_elems == 32) && r0 >= 0 && r0 <= 8, "b200_debug_umma_rowshift: bad arguments");
  CUtensorMap tmA, tmB;
  const CUtensorMapSwizzle swz = k_elems == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  uint64_t dimsA[2] = {(uint64_t)k_elems, 136}, dimsB[2] = {(uint64_t)k_elems, 64};
  uint64_t strides[1] = {(uint64_t)k_elems * 2};
  uint32_t boxA[2] = {(uint32_t)k_elems, 136}, boxB[2] = {(uint32_t)k_elems, 64};
  int rc;
  if ((rc = encode_tmap(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, A, dimsA, strides, boxA, swz))) return rc;
  if ((rc = encode_tmap(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, B, dimsB, strides, boxB, swz))) return rc;
  const int smem_bytes = 30 * 1024;
  if (k_elems == 64) {
    B200_CHECK_CUDA(cudaFuncSetAttribute(umma_rowshift_probe_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    umma_rowshift_probe_kernel<128><<<1, 128, smem_bytes, stream>>>(tmA, tmB, D, r0, mode); note_launch();
  } else {
    B200_CHECK_CUDA(cudaFuncSetAttribute(umma_rowshift_probe_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    umma_rowshift_probe_kernel<64><<<1, 128, smem_bytes, stream>>>(tmA, tmB, D, r0, mode); note_launch();
  }
  B200_CHECK_CUDA(cudaGetLastError());
  return B200_OK;
}

// ---------------------------------------------------------------------------------------------------------------------------------
// UMMA issue-rate probe: how many clocks does one SS-mode tcgen05.mma (M = 128, K = 16, bf16, both operands K-major SW128 in shared
// memory) cost at a given N when NOTHING else limits it - no TMA, no epilogue, operands resident?  This is the floor every conv / GEMM
// tile shape in this library is measured against.  Knobs: number of TMEM accumulators the MMAs rotate over (a chain
// of MMAs into ONE accumulator is dependent), one or two issuing warps, how many distinct A tiles are cycled through, and `writers` warps
// that stream 16-byte shared-memory stores next to the MMAs (the bank traffic a TMA producer would add).
//   out[2 * cta + issuer] = clocks for `iters` x 4 MMAs (issue of the first to completion of the last); out[2 * grid + cta] = 512-byte
//   store instructions the writer warps retired in that time.
__global__ void __launch_bounds__(256, 1)
umma_rate_probe_kernel(int n, int iters, int n_acc, int issuers, int writers, int a_tiles, unsigned long long* __restrict__ out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;                                   // 4 tiles x 16 KB
  uint8_t* sB = smem + 64 * 1024;                       // 256 rows x 128 B
  uint8_t* sW = smem + 96 * 1024;                       // 16 KB writer scratch
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 112 * 1024);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
  volatile int* stop = reinterpret_cast<volatile int*>(bars + 3);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // finite, non-trivial bf16 operand bits (exponent field 0x3f / 0xbf: |x| in [1, 2)), so the datapath toggles like real data
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += 256) {
    const uint32_t h = (uint32_t)i * 2654435761u;
    reinterpret_cast<uint32_t*>(smem)[i] = (h & 0x807f807fu) | 0x3f803f80u;
  }
  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    *stop = 0;
    fence_mbar_init();
  }
  fence_async_smem();
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (warp < issuers) {
    const uint32_t idesc = make_idesc(FMT_BF16, FMT_BF16, 128, (uint32_t)n, 0, 0);
    const uint32_t a0 = desc_lo_kmajor(smem_u32(sA)), b_lo = desc_lo_kmajor(smem_u32(sB));
    const uint32_t d0 = tmem_base + (uint32_t)(warp * n_acc * n);
    const long long t0 = clock64();
    int acc = 0, at = 0;
    for (int it = 0; it < iters; ++it) {
      const uint32_t a_lo = a0 + (uint32_t)at * (16384u >> 4);
#pragma unroll
      for (int k = 0; k < 4; ++k) mma_f16_ss_w(d0 + (uint32_t)(acc * n), a_lo + 2 * k, kDescHiSw128, b_lo + 2 * k, kDescHiSw128, idesc, it >= n_acc || k ? 1u : 0u);
      if (++acc == n_acc) acc = 0;
      if (++at == a_tiles) at = 0;
    }
    tc_commit_w(&bars[warp]);
    mbar_wait(&bars[warp], 0);
    const long long t1 = clock64();
    if (lane == 0) {
      out[2 * blockIdx.x + warp] = (unsigned long long)(t1 - t0);
      if (warp == 0) *stop = 1;
    }
  } else if (warp >= 4 && warp < 4 + writers) {
    unsigned long long cnt = 0;
    uint4* dst = reinterpret_cast<uint4*>(sW) + (warp - 4) * 256 + lane;
    const uint4 v = make_uint4(lane, warp, 1, 2);
    while (!*stop) {
#pragma unroll
      for (int j = 0; j < 8; ++j) asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(smem_u32(dst + j * 32)), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
      cnt += 8;
    }
    if (lane == 0) atomicAdd(&out[2 * gridDim.x + blockIdx.x], cnt);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

int debug_umma_rate(int n, int iters, int n_acc, int issuers, int writers, int a_tiles, int grid, unsigned long long* out, cudaStream_t stream) {
  B200_CHECK_ARG(out && n >= 16 && n <= 256 && n % 16 == 0 && iters > 0 && n_acc >= 1 && issuers >= 1 && issuers <= 2 && writers >= 0 && writers <= 4 &&
                     a_tiles >= 1 && a_tiles <= 4 && grid >= 1 && issuers * n_acc * n <= 512,
                 "b200_debug_umma_rate: bad arguments");
  const int smem_bytes = 114 * 1024;
  B200_CHECK_CUDA(cudaFuncSetAttribute(umma_rate_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
  B200_CHECK_CUDA(cudaMemsetAsync(out, 0, sizeof(unsigned long long) * 3 * grid, stream));
  umma_rate_probe_kernel<<<grid, 256, smem_bytes, stream>>>(n, iters, n_acc, issuers, writers, a_tiles, out);
  note_launch();
  B200_CHECK_CUDA(cudaGetLastError());
  return B200_OK;
}

}  // namespace b200
