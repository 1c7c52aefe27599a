// Host plumbing: last-error storage, device info cache, TMA tensor-map encoding, ABI bookkeeping.
#include "pio_host.h"

#include <stdlib.h>
#include <string.h>

#include <vector>

namespace pio {

thread_local char g_last_error[512] = "";
std::atomic<int64_t> g_launch_count{0};

int pdl_mode() {
  static const int mode = [] {
    const char* e = getenv("PIO_PDL");
    if (!e || !*e) return 1;
    return (e[0] == '0') ? 0 : (e[0] == '2' ? 2 : 1);
  }();
  return mode;
}

int pdl_sm_count() {
  DeviceInfo d;
  return get_device_info(&d) == PIO_OK ? d.sm_count : 0;
}

int get_device_info(DeviceInfo* out) {
  static std::mutex mu;
  static DeviceInfo cache[64];
  static bool have[64] = {false};
  int dev = 0;
  PIO_CUDA_OK(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) return fail(PIO_ERR_CUDA, "device index %d out of range", dev);
  std::lock_guard<std::mutex> lock(mu);
  if (!have[dev]) {
    DeviceInfo d;
    d.device = dev;
    PIO_CUDA_OK(cudaDeviceGetAttribute(&d.sm_count, cudaDevAttrMultiProcessorCount, dev));
    PIO_CUDA_OK(cudaDeviceGetAttribute(&d.cc_major, cudaDevAttrComputeCapabilityMajor, dev));
    PIO_CUDA_OK(cudaDeviceGetAttribute(&d.cc_minor, cudaDevAttrComputeCapabilityMinor, dev));
    PIO_CUDA_OK(cudaDeviceGetAttribute(&d.max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    cache[dev] = d;
    have[dev] = true;
  }
  *out = cache[dev];
  return PIO_OK;
}

// ---------------------------------------------------------------------------------------------------------
// profiling
// ---------------------------------------------------------------------------------------------------------
struct ProfRecord { int family; double flops, bytes; cudaEvent_t e0, e1; };
static std::mutex g_prof_mu;
static std::vector<ProfRecord> g_prof;
static std::atomic<int> g_prof_on{0};

ProfileScope::ProfileScope(int family, double flops, double bytes, cudaStream_t stream)
    : family_(family), flops_(flops), bytes_(bytes), stream_(stream) {
  if (!g_prof_on.load(std::memory_order_relaxed)) return;
  if (cudaEventCreate(&e0_) != cudaSuccess) return;
  on_ = true;
  cudaEventRecord(e0_, stream_);
}
ProfileScope::~ProfileScope() {
  if (!on_) return;
  cudaEvent_t e1;
  if (cudaEventCreate(&e1) != cudaSuccess) return;
  cudaEventRecord(e1, stream_);
  std::lock_guard<std::mutex> lock(g_prof_mu);
  g_prof.push_back({family_, flops_, bytes_, e0_, e1});
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

int encode_tmap_bf16(CUtensorMap* map, const void* base, int rank, const uint64_t* dims,
                     const uint64_t* strides_bytes, const uint32_t* box) {
  return encode_tmap(map, base, /*is_f32=*/false, rank, dims, strides_bytes, box, 128);
}

int encode_tmap(CUtensorMap* map, const void* base, bool is_f32, int rank, const uint64_t* dims,
                const uint64_t* strides_bytes, const uint32_t* box, int swizzle_bytes) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return fail(PIO_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t gdim[5];
  cuuint64_t gstr[5];
  cuuint32_t bx[5];
  cuuint32_t estr[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    estr[i] = 1;
    if (i > 0) gstr[i - 1] = strides_bytes[i - 1];
  }
  CUresult r = fn(map, is_f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16,
                  static_cast<cuuint32_t>(rank), const_cast<void*>(base),
                  gdim, gstr, bx, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                      : (swizzle_bytes == 0 ? CU_TENSOR_MAP_SWIZZLE_NONE : CU_TENSOR_MAP_SWIZZLE_128B),
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    return fail(PIO_ERR_CUDA,
                "cuTensorMapEncodeTiled failed (%d): rank %d dims [%llu,%llu,%llu] strides [%llu,%llu] box [%u,%u,%u] "
                "base %p",
                (int)r, rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
                (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 1 ? strides_bytes[0] : 0),
                (unsigned long long)(rank > 2 ? strides_bytes[1] : 0), box[0], rank > 1 ? box[1] : 0,
                rank > 2 ? box[2] : 0, base);
  }
  return PIO_OK;
}

}  // namespace pio

extern "C" {

int pio_abi_version(void) { return PIO_ABI_VERSION; }
const char* pio_last_error(void) { return pio::g_last_error; }
int64_t pio_launch_count(void) { return pio::g_launch_count.load(); }

void pio_profile_enable(int on) { pio::g_prof_on.store(on ? 1 : 0); }

// Drains the recorded launches: out[f*4 + {0,1,2,3}] = {milliseconds, flops, bytes, launches} per kernel family
// (pio::KernelFamily order: layernorm, gemm, softmax, attention, combine).  Synchronises on the recorded events.
int pio_profile_read(double* out, int n_families) {
  using namespace pio;
  std::lock_guard<std::mutex> lock(g_prof_mu);
  for (int i = 0; i < n_families * 4; ++i) out[i] = 0.0;
  for (auto& r : g_prof) {
    float ms = 0.f;
    cudaEventSynchronize(r.e1);
    if (cudaEventElapsedTime(&ms, r.e0, r.e1) == cudaSuccess && r.family < n_families) {
      out[r.family * 4 + 0] += ms;
      out[r.family * 4 + 1] += r.flops;
      out[r.family * 4 + 2] += r.bytes;
      out[r.family * 4 + 3] += 1.0;
    }
    cudaEventDestroy(r.e0);
    cudaEventDestroy(r.e1);
  }
  g_prof.clear();
  return PIO_OK;
}

int pio_check_device(void) {
  pio::DeviceInfo d;
  int rc = pio::get_device_info(&d);
  if (rc != PIO_OK) return rc;
  if (d.cc_major != 10)
    return pio::fail(PIO_ERR_ARCH, "device %d is sm_%d%d; this library contains sm_100a code only", d.device,
                     d.cc_major, d.cc_minor);
  return PIO_OK;
}

}  // extern "C"
