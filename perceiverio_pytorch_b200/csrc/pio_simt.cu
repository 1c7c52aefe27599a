// HBM-bound SIMT kernels: LayerNorm + bf16 cast, row softmax (materialised attention path) and the
// log-sum-exp merge of partial attention results (key splits / key-axis shards across GPUs).
#include "pio_common.cuh"
#include "pio_host.h"

namespace pio {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ------------------------------------------------------------------------------------------------------------
// LayerNorm + cast.  One warp per row, the row lives in registers (C <= 32 * MAXV), two-pass statistics.
// Lane l owns columns l, l+32, ... so that loads and bf16 stores are coalesced for any C (261, 322, 1026 ...).
// ------------------------------------------------------------------------------------------------------------
template <int MAXV>
__global__ void __launch_bounds__(256) pio_layernorm_kernel(const float* __restrict__ x, long long ldx,
                                                            __nv_bfloat16* __restrict__ y, long long ldy,
                                                            const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, long long rows, int C,
                                                            int normalize, float eps) {
  const int lane = threadIdx.x & 31;
  const long long warps_per_grid = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < rows;
       row += warps_per_grid) {
    const float* xr = x + row * ldx;
    float v[MAXV];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int c = lane + 32 * i;
      v[i] = (c < C) ? __ldg(xr + c) : 0.f;
      s += v[i];
    }
    float mean = 0.f, rstd = 1.f;
    if (normalize) {
      mean = warp_sum(s) / (float)C;
      float q = 0.f;
#pragma unroll
      for (int i = 0; i < MAXV; ++i) {
        const int c = lane + 32 * i;
        const float d = (c < C) ? v[i] - mean : 0.f;
        q += d * d;
      }
      rstd = rsqrtf(warp_sum(q) / (float)C + eps);
    }
    __nv_bfloat16* yr = y + row * ldy;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int c = lane + 32 * i;
      if (c < ldy) {
        float o = 0.f;
        if (c < C) {
          o = (v[i] - mean) * rstd;
          if (gamma) o = o * __ldg(gamma + c);
          if (beta) o += __ldg(beta + c);
        }
        yr[c] = __float2bfloat16_rn(o);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// Row softmax, one 256-thread block per row (rows can be 50k+ long): three passes over the row, which stays in
// L1/L2 (<= 208 KB).  Masked keys are excluded; an all-masked or wiped row is written as zeros.
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pio_softmax_kernel(pio_softmax_args a) {
  __shared__ float red[8];
  __shared__ float bcast;
  const long long row_id = blockIdx.x;
  const int b = (int)(row_id / a.rows);
  const int r = (int)(row_id % a.rows);
  const float* s = a.S + (long long)b * a.strideS + (long long)r * a.lds;
  __nv_bfloat16* p = reinterpret_cast<__nv_bfloat16*>(a.P) + (long long)b * a.strideP + (long long)r * a.ldp;
  const uint8_t* km = a.key_mask ? a.key_mask + (long long)b * a.stride_km : nullptr;
  const bool keep = a.row_keep ? a.row_keep[(long long)b * a.stride_rk + r] != 0 : true;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (!keep) {
    for (int c = tid; c < a.ldp; c += 256) p[c] = __float2bfloat16_rn(0.f);
    return;
  }
  float m = -INFINITY;
  for (int c = tid; c < a.cols; c += 256)
    if (!km || km[c]) m = fmaxf(m, __ldg(s + c) * a.scale);
  m = warp_max(m);
  if (lane == 0) red[warp] = m;
  __syncthreads();
  if (tid == 0) {
    float mm = red[0];
    for (int i = 1; i < 8; ++i) mm = fmaxf(mm, red[i]);
    bcast = mm;
  }
  __syncthreads();
  m = bcast;
  __syncthreads();
  if (m == -INFINITY) {  // every key masked
    for (int c = tid; c < a.ldp; c += 256) p[c] = __float2bfloat16_rn(0.f);
    return;
  }
  float l = 0.f;
  for (int c = tid; c < a.cols; c += 256)
    if (!km || km[c]) l += __expf(__ldg(s + c) * a.scale - m);
  l = warp_sum(l);
  if (lane == 0) red[warp] = l;
  __syncthreads();
  if (tid == 0) {
    float ll = 0.f;
    for (int i = 0; i < 8; ++i) ll += red[i];
    bcast = 1.0f / ll;
  }
  __syncthreads();
  const float inv = bcast;
  for (int c = tid; c < a.ldp; c += 256) {
    float o = 0.f;
    if (c < a.cols && (!km || km[c])) o = __expf(__ldg(s + c) * a.scale - m) * inv;
    p[c] = __float2bfloat16_rn(o);
  }
}

// ------------------------------------------------------------------------------------------------------------
// Merge partial attention results: one warp per (b, h, query row).
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pio_combine_kernel(pio_combine_args a) {
  const int lane = threadIdx.x & 31;
  const long long rows = (long long)a.B * a.H * a.Nq;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int q = (int)(row % a.Nq);
  const int h = (int)((row / a.Nq) % a.H);
  const int b = (int)(row / ((long long)a.Nq * a.H));
  __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(a.O) + (long long)b * a.strideO + (long long)q * a.ldo +
                     (long long)h * a.dv;
  const bool keep = a.row_keep ? a.row_keep[(long long)b * a.stride_rk + q] != 0 : true;
  float M = -INFINITY;
  for (int p = 0; p < a.parts; ++p) M = fmaxf(M, __ldg(a.m_part + p * rows + row));
  if (!keep || M == -INFINITY) {
    for (int c = lane; c < a.dv; c += 32) o[c] = __float2bfloat16_rn(0.f);
    return;
  }
  float L = 0.f;
  for (int p = 0; p < a.parts; ++p) {
    const float mp = __ldg(a.m_part + p * rows + row);
    if (mp != -INFINITY) L += __ldg(a.l_part + p * rows + row) * __expf(mp - M);
  }
  const float inv = 1.0f / L;
  for (int c = lane; c < a.dv; c += 32) {
    float acc = 0.f;
    for (int p = 0; p < a.parts; ++p) {
      const float mp = __ldg(a.m_part + p * rows + row);
      if (mp != -INFINITY) acc += __ldg(a.O_part + (p * rows + row) * a.dv + c) * __expf(mp - M);
    }
    o[c] = __float2bfloat16_rn(acc * inv);
  }
}

}  // namespace pio

extern "C" int pio_layernorm_bf16(const pio_layernorm_args* a, void* stream_) {
  using namespace pio;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  PIO_REQUIRE(a && a->x && a->y, "pio_layernorm_bf16: null pointer");
  PIO_REQUIRE(a->rows > 0 && a->C > 0, "pio_layernorm_bf16: bad shape rows=%lld C=%d", (long long)a->rows, a->C);
  PIO_REQUIRE(a->ldy >= a->C && a->ldx >= 0, "pio_layernorm_bf16: bad leading dimension");
  PIO_REQUIRE(a->ldy <= 2048, "pio_layernorm_bf16: C up to 2048 supported (got ldy=%lld)", (long long)a->ldy);
  DeviceInfo dev;
  int rc = get_device_info(&dev);
  if (rc != PIO_OK) return rc;
  const int warps_per_block = 8;
  long long blocks = (a->rows + warps_per_block - 1) / warps_per_block;
  const long long max_blocks = (long long)dev.sm_count * 32;
  if (blocks > max_blocks) blocks = max_blocks;
  __nv_bfloat16* y = reinterpret_cast<__nv_bfloat16*>(a->y);
#define PIO_LN_LAUNCH(MAXV)                                                                                      \
  pio_layernorm_kernel<MAXV><<<(unsigned)blocks, 256, 0, stream>>>(a->x, a->ldx, y, a->ldy, a->gamma, a->beta, \
                                                                   a->rows, a->C, a->normalize, a->eps)
  const int need = (int)((a->ldy + 31) / 32);
  if (need <= 4) PIO_LN_LAUNCH(4);
  else if (need <= 12) PIO_LN_LAUNCH(12);
  else if (need <= 24) PIO_LN_LAUNCH(24);
  else if (need <= 40) PIO_LN_LAUNCH(40);
  else PIO_LN_LAUNCH(64);
#undef PIO_LN_LAUNCH
  g_launch_count.fetch_add(1);
  PIO_CUDA_OK(cudaGetLastError());
  return PIO_OK;
}

extern "C" int pio_softmax_bf16(const pio_softmax_args* a, void* stream_) {
  using namespace pio;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  PIO_REQUIRE(a && a->S && a->P, "pio_softmax_bf16: null pointer");
  PIO_REQUIRE(a->batch > 0 && a->rows > 0 && a->cols > 0 && a->ldp >= a->cols && a->lds >= a->cols,
              "pio_softmax_bf16: bad shape");
  const long long blocks = (long long)a->batch * a->rows;
  PIO_REQUIRE(blocks < (1ll << 31), "pio_softmax_bf16: too many rows");
  pio_softmax_kernel<<<(unsigned)blocks, 256, 0, stream>>>(*a);
  g_launch_count.fetch_add(1);
  PIO_CUDA_OK(cudaGetLastError());
  return PIO_OK;
}

extern "C" int pio_attention_combine(const pio_combine_args* a, void* stream_) {
  using namespace pio;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  PIO_REQUIRE(a && a->O_part && a->m_part && a->l_part && a->O, "pio_attention_combine: null pointer");
  PIO_REQUIRE(a->parts > 0 && a->B > 0 && a->H > 0 && a->Nq > 0 && a->dv > 0, "pio_attention_combine: bad shape");
  const long long rows = (long long)a->B * a->H * a->Nq;
  const long long blocks = (rows + 7) / 8;
  pio_combine_kernel<<<(unsigned)blocks, 256, 0, stream>>>(*a);
  g_launch_count.fetch_add(1);
  PIO_CUDA_OK(cudaGetLastError());
  return PIO_OK;
}
