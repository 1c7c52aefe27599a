// Host-side plumbing shared by the C-ABI entry points: error reporting, driver entry points, TMA tensor maps.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>
#include <mutex>
#include <utility>

#include "../../include/pio_b200.h"

namespace pio {

extern thread_local char g_last_error[512];
extern std::atomic<int64_t> g_launch_count;

inline int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
  va_end(ap);
  return code;
}

#define PIO_CUDA_OK(expr)                                                                              \
  do {                                                                                                 \
    cudaError_t _e = (expr);                                                                           \
    if (_e != cudaSuccess)                                                                             \
      return ::pio::fail(PIO_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                         __LINE__);                                                                    \
  } while (0)

#define PIO_REQUIRE(cond, ...)                                          \
  do {                                                                  \
    if (!(cond)) return ::pio::fail(PIO_ERR_INVALID_ARGUMENT, __VA_ARGS__); \
  } while (0)

struct DeviceInfo {
  int device = -1;
  int sm_count = 0;
  int cc_major = 0;
  int cc_minor = 0;
  int max_smem_optin = 0;
};
// Cached properties of the current device (one lookup per device per process).
int get_device_info(DeviceInfo* out);

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time dependency on libcuda, so the
// library loads in the GPU-less build container for the symbol-export test).
// Encodes a bf16 tensor of rank `rank` (dims[0] innermost) with SWIZZLE_128B and zero OOB fill.
int encode_tmap_bf16(CUtensorMap* map, const void* base, int rank, const uint64_t* dims,
                     const uint64_t* strides_bytes /* rank-1 entries, for dims[1..] */, const uint32_t* box);
// Same for fp32 (is_f32) or bf16 elements; the box's inner extent must span exactly `swizzle_bytes` (128 or 64) bytes,
// or any multiple of 16 bytes with swizzle_bytes = 0 (no swizzle).
int encode_tmap(CUtensorMap* map, const void* base, bool is_f32, int rank, const uint64_t* dims,
                const uint64_t* strides_bytes, const uint32_t* box, int swizzle_bytes = 128);

// Optional per-launch profiling: when enabled (pio_profile_enable), every entry point brackets its kernel launch
// with CUDA events recorded on the launching stream *inside* the library, so the interval contains the kernel and
// nothing of the host-side Python gap.  Not usable during stream capture.
enum KernelFamily { KF_LAYERNORM = 0, KF_GEMM = 1, KF_SOFTMAX = 2, KF_FLASH = 3, KF_COMBINE = 4, KF_LINEAR_F32 = 5, KF_COUNT = 6 };
struct ProfileScope {
  ProfileScope(int family, double flops, double bytes, cudaStream_t stream);
  ~ProfileScope();
  int family_;
  double flops_, bytes_;
  cudaStream_t stream_;
  cudaEvent_t e0_ = nullptr;
  bool on_ = false;
};

// CTA-pair GEMM (pio_gemm2.cu)
bool gemm2_eligible(const pio_gemm_args* a);
int launch_gemm2(const pio_gemm_args* a, const DeviceInfo& dev, cudaStream_t stream);

// Persistent two-tile attention kernel (pio_flash2.cu)
bool flash2_eligible(const pio_attention_args* a);
int launch_flash2(const pio_attention_args* a, const DeviceInfo& dev, cudaStream_t stream);

// Query-tile-in-TMEM variant of the one-tile kernel for the folded encoder cross-attends (pio_flash_qt.cu)
int flash_qt_key_tile(int d);
bool flash_qt_eligible(const pio_attention_args* a);
int launch_flash_qt(const pio_attention_args* a, const DeviceInfo& dev, cudaStream_t stream);

// cudaLaunchKernelEx with an optional cluster width and programmatic dependent launch.
// PDL: a grid that leaves SMs idle (fewer CTAs than SMs: the batch-1 towers, 40 .. 128 CTAs per kernel, 7 kernels per
// layer) is launched with programmatic stream serialization, so its CTAs are scheduled and run their prologue (barrier
// init, TMEM allocation, descriptor prefetch) while the previous kernel drains; every kernel calls pdl_sync() before it
// touches global memory.  Grids that fill the machine keep plain serialization: for the batch-64 recipe, where every
// kernel occupies all SMs with one large-smem CTA, PDL measured slower (33.1 vs 31.7 ms per bench step).
// PIO_PDL=0 disables it, PIO_PDL=2 forces it for every launch.
int pdl_mode();
int pdl_sm_count();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                 int cluster_x, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int n = 0;
  if (cluster_x > 1) {
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = cluster_x;
    attr[n].val.clusterDim.y = 1;
    attr[n].val.clusterDim.z = 1;
    ++n;
  }
  const int mode = pdl_mode();
  const long long ctas = (long long)grid.x * grid.y * grid.z;
  if (mode == 2 || (mode == 1 && ctas < pdl_sm_count())) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

// cudaFuncAttributeMaxDynamicSharedMemorySize is a PER-DEVICE attribute: a process that drives several GPUs (a model moved
// with .to("cuda:1"), or one process per node) must set it on each device it launches on.  One bit per device ordinal.
struct PerDeviceOnce {
  std::atomic<uint64_t> done{0};
  template <typename F>
  cudaError_t run(int device, F&& f) {
    const uint64_t bit = 1ull << (device & 63);
    if (done.load(std::memory_order_acquire) & bit) return cudaSuccess;
    cudaError_t e = f();   // idempotent, so two racing threads may both run it
    if (e == cudaSuccess) done.fetch_or(bit, std::memory_order_release);
    return e;
  }
};

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace pio
