// Host-side helpers shared by the C-ABI entry points: error reporting, TMA tensor-map encoding
// (driver entry point fetched through the runtime, so the library does not link libcuda), SM count.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace b200 {

// status codes (same values as include/b200_dit.h)
#ifndef B200_OK
#define B200_OK 0
#define B200_ERR_INVALID (-1)      // bad argument (shape / alignment / null pointer)
#define B200_ERR_CUDA (-2)         // a CUDA runtime / driver call failed
#define B200_ERR_UNSUPPORTED (-3)
#endif

void set_last_error(const char* fmt, ...);

#define B200_CHECK_ARG(cond, ...)        \
  do {                                   \
    if (!(cond)) {                       \
      b200::set_last_error(__VA_ARGS__); \
      return B200_ERR_INVALID;     \
    }                                    \
  } while (0)

#define B200_CHECK_CUDA(expr)                                                                         \
  do {                                                                                                \
    cudaError_t _e = (expr);                                                                          \
    if (_e != cudaSuccess) {                                                                          \
      b200::set_last_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return B200_ERR_CUDA;                                                                     \
    }                                                                                                 \
  } while (0)

int num_sms();

// Runtime options (b200_set_option): A/B switches for measurements and tests.  Initial values come from the environment
// (B200_CONV_HALO, B200_HALO_BASE_OFFSET, B200_CONV_NARROW) the first time an option is read.
enum Option : int { OPT_CONV_HALO = 0, OPT_HALO_BASE_OFFSET = 1, OPT_CONV_NARROW = 2, OPT_COUNT = 3 };
int get_option(int opt);
int set_option(const char* name, int value);

// Launch accounting for bench.py's `gpu_launches` (b200_launch_count): every kernel launch of this library calls note_launch().
void note_launch(int n = 1);
long long launch_count();

// Optional CUDA-event bracket around the attention launches (b200_prof_fmha_begin / _end): the roofline of the dominant kernel is
// measured on the launching stream even when the launch happens below the C ABI (b200_wan_block_fwd).
struct ProfScope {
  int slot;
  cudaStream_t stream;
  ProfScope(long long sq, long long sk, int heads, int head_dim, cudaStream_t stream);
  ~ProfScope();
};

// Encode a tiled tensor map over a row-major tensor.  dims/box are innermost-first; strides_bytes
// has rank-1 entries (stride of dim 1..rank-1).  Returns 0 on success.
int encode_tmap(CUtensorMap* out, CUtensorMapDataType dtype, int rank, const void* base, const uint64_t* dims,
                const uint64_t* strides_bytes, const uint32_t* box, CUtensorMapSwizzle swizzle);

// 2-D bf16 row-major [rows, cols] with leading dimension ld (elements); box = {box_cols, box_rows}, 128B swizzle.
int encode_tmap_2d_bf16(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                        uint32_t box_rows, uint32_t box_cols);

}  // namespace b200
