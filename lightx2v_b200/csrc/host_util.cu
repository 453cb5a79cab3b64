#include "host_util.cuh"

#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <mutex>
#include <vector>

namespace b200 {

static thread_local char g_last_error[512] = "";

void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
  va_end(ap);
}

const char* last_error() { return g_last_error; }

int num_sms() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) sms = 148;
  }
  return sms;
}

// ---- runtime options ---------------------------------------------------------------------------------------------------------
#ifndef B200_CONV_HALO_DEFAULT
#define B200_CONV_HALO_DEFAULT 1     // conv3d_halo.cu: verified on hardware in round 2 (profiles/r02_probe_rowshift.txt, r02_halo_parity.txt)
#endif
static const char* const kOptNames[OPT_COUNT] = {"conv_halo", "halo_base_offset", "conv_narrow"};
static const char* const kOptEnv[OPT_COUNT] = {"B200_CONV_HALO", "B200_HALO_BASE_OFFSET", "B200_CONV_NARROW"};
static const int kOptDefault[OPT_COUNT] = {B200_CONV_HALO_DEFAULT, 0, 0};   // halo_base_offset 0: the tensor core swizzles by absolute address bits (probe)
static std::atomic<int> g_opts[OPT_COUNT];
static std::once_flag g_opts_once;

static void init_options() {
  for (int i = 0; i < OPT_COUNT; ++i) {
    const char* e = getenv(kOptEnv[i]);
    g_opts[i].store(e ? atoi(e) : kOptDefault[i]);
  }
}
int get_option(int opt) {
  std::call_once(g_opts_once, init_options);
  return (opt >= 0 && opt < OPT_COUNT) ? g_opts[opt].load(std::memory_order_relaxed) : 0;
}
int set_option(const char* name, int value) {
  std::call_once(g_opts_once, init_options);
  for (int i = 0; name && i < OPT_COUNT; ++i)
    if (strcmp(name, kOptNames[i]) == 0) {
      g_opts[i].store(value);
      return B200_OK;
    }
  set_last_error("b200_set_option: unknown option '%s' (conv_halo, halo_base_offset, conv_narrow)", name ? name : "(null)");
  return B200_ERR_INVALID;
}

static std::atomic<long long> g_launches{0};
void note_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
long long launch_count() { return g_launches.load(std::memory_order_relaxed); }

// ---- attention launch profiler (off unless b200_prof_fmha_begin was called) ------------------------------------------------
struct ProfRec {
  cudaEvent_t e0, e1;
  long long sq, sk;
  int heads, head_dim;
};
static std::mutex g_prof_mu;
static std::vector<ProfRec> g_prof;
static int g_prof_cap = 0;      // 0 = disabled

ProfScope::ProfScope(long long sq, long long sk, int heads, int head_dim, cudaStream_t s) : slot(-1), stream(s) {
  if (g_prof_cap == 0) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  if ((int)g_prof.size() >= g_prof_cap) return;
  ProfRec r{nullptr, nullptr, sq, sk, heads, head_dim};
  if (cudaEventCreate(&r.e0) != cudaSuccess || cudaEventCreate(&r.e1) != cudaSuccess) return;
  cudaEventRecord(r.e0, s);
  g_prof.push_back(r);
  slot = (int)g_prof.size() - 1;
}
ProfScope::~ProfScope() {
  if (slot < 0) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  if (slot < (int)g_prof.size()) cudaEventRecord(g_prof[slot].e1, stream);
}

int prof_fmha_begin(int capacity) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  for (auto& r : g_prof) {
    cudaEventDestroy(r.e0);
    cudaEventDestroy(r.e1);
  }
  g_prof.clear();
  g_prof_cap = capacity > 0 ? capacity : 0;
  return B200_OK;
}

// Blocks until the recorded launches have finished; ms[i] = duration, meta[4 i ..] = {sq, sk, heads, head_dim}.  Returns the count.
int prof_fmha_end(float* ms, long long* meta, int capacity) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  int n = 0;
  for (auto& r : g_prof) {
    float t = -1.f;
    if (cudaEventSynchronize(r.e1) == cudaSuccess) cudaEventElapsedTime(&t, r.e0, r.e1);
    if (n < capacity && ms && meta) {
      ms[n] = t;
      meta[4 * n + 0] = r.sq; meta[4 * n + 1] = r.sk; meta[4 * n + 2] = r.heads; meta[4 * n + 3] = r.head_dim;
      ++n;
    }
    cudaEventDestroy(r.e0);
    cudaEventDestroy(r.e1);
  }
  g_prof.clear();
  g_prof_cap = 0;
  return n;
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, []() {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
    if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) fn = reinterpret_cast<PFN_encodeTiled>(p);
  });
  return fn;
}

int encode_tmap(CUtensorMap* out, CUtensorMapDataType dtype, int rank, const void* base, const uint64_t* dims,
                const uint64_t* strides_bytes, const uint32_t* box, CUtensorMapSwizzle swizzle) {
  PFN_encodeTiled fn = get_encode_fn();
  if (!fn) {
    set_last_error("cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
    return B200_ERR_CUDA;
  }
  cuuint64_t gdims[5];
  cuuint64_t gstrides[4];
  cuuint32_t gbox[5];
  cuuint32_t estr[5];
  for (int i = 0; i < rank; ++i) {
    gdims[i] = dims[i];
    gbox[i] = box[i];
    estr[i] = 1;
  }
  for (int i = 0; i + 1 < rank; ++i) gstrides[i] = strides_bytes[i];
  CUresult r = fn(out, dtype, (cuuint32_t)rank, const_cast<void*>(base), gdims, gstrides, gbox, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("cuTensorMapEncodeTiled failed (CUresult %d): rank %d dims[0..1]=%llu,%llu stride1=%llu box=%u,%u base=%p",
                   (int)r, rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
                   (unsigned long long)(rank > 1 ? strides_bytes[0] : 0), box[0], rank > 1 ? box[1] : 0, base);
    return B200_ERR_CUDA;
  }
  return B200_OK;
}

int encode_tmap_2d_bf16(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                        uint32_t box_rows, uint32_t box_cols) {
  uint64_t dims[2] = {cols, rows};
  uint64_t strides[1] = {ld * 2};
  uint32_t box[2] = {box_cols, box_rows};
  return encode_tmap(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
}

}  // namespace b200
