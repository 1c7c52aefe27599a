// Persistent, warp-specialised bf16 GEMM for sm_100a: TMA (SWIZZLE_128B) -> smem ring -> tcgen05.mma with
// double-buffered fp32 accumulators in TMEM -> fused epilogue (alpha, bias, exact GELU, fp32 residual, fp32 and/or
// bf16 stores).  One CTA per SM; tile 128 x BN x 64.
//
// Roles (256 threads): warp 0 = TMA producer, warp 1 = MMA issuer, warp 2 = TMEM allocator, warp 3 idle,
// warps 4..7 = epilogue (warp w owns TMEM lanes 32*(w%4) .. +31, i.e. accumulator rows).
#include "pio_common.cuh"
#include "pio_host.h"

namespace pio {

struct GemmEpilogue {
  int M, N, K, batch;
  int tiles_m, tiles_n;
  int a_bcast, b_bcast;  // operand shared by every batch entry (batch stride 0)
  const float* bias;
  int bias_mode;
  int act;
  float alpha;
  const float* residual;
  long long ldr, strideR;
  float* out_f32;
  long long ldo32, strideO32;
  __nv_bfloat16* out_bf16;
  long long ldo16, strideO16;
};

template <int BN>
struct GemmCfg {
  static constexpr int BM = 128;
  static constexpr int BK = 64;
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (BN == 256) ? 4 : (BN == 128 ? 6 : 8);
  static constexpr int TMEM_COLS = 2 * BN;  // 128, 256 or 512: powers of two
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*alignment slack*/ + 256 /*barriers*/;
};

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }

template <int BN, bool B_MN>
__global__ void __launch_bounds__(256, 1)
pio_gemm_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                const GemmEpilogue ep) {
  using Cfg = GemmCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B tiles need 1024-byte aligned bases.
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES);
  uint64_t* full_bar = bars;                      // [STAGES]  TMA -> MMA
  uint64_t* empty_bar = bars + Cfg::STAGES;       // [STAGES]  MMA -> TMA
  uint64_t* tmem_full = bars + 2 * Cfg::STAGES;   // [2]       MMA -> epilogue
  uint64_t* tmem_empty = tmem_full + 2;           // [2]       epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_k_chunks = (ep.K + Cfg::BK - 1) / Cfg::BK;
  const int total_tiles = ep.tiles_m * ep.tiles_n * ep.batch;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < Cfg::STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 128);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // ================= TMA producer =================
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const int nt = t % ep.tiles_n;
        const int mt = (t / ep.tiles_n) % ep.tiles_m;
        const int z = t / (ep.tiles_n * ep.tiles_m);
        const int m0 = mt * Cfg::BM, n0 = nt * BN;
        for (int kc = 0; kc < num_k_chunks; ++kc) {
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
          uint8_t* sb = sa + Cfg::A_BYTES;
          mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
          tma_load_3d(sa, &tmap_a, &full_bar[stage], kc * Cfg::BK, m0, ep.a_bcast ? 0 : z);
          if constexpr (!B_MN) {
            tma_load_3d(sb, &tmap_b, &full_bar[stage], kc * Cfg::BK, n0, ep.b_bcast ? 0 : z);
          } else {
#pragma unroll
            for (int j = 0; j < BN / 64; ++j)
              tma_load_3d(sb + j * (64 * 128), &tmap_b, &full_bar[stage], n0 + j * 64, kc * Cfg::BK, ep.b_bcast ? 0 : z);
          }
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ================= MMA issuer =================
      constexpr uint32_t idesc = make_idesc_f16(128, BN, /*bf16*/ 1, /*a K-major*/ 0, B_MN ? 1 : 0);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1u;
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kc = 0; kc < num_k_chunks; ++kc) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE_BYTES);
          const uint32_t sb = sa + Cfg::A_BYTES;
          const int krem = ep.K - kc * Cfg::BK;
          const int ksteps = krem >= Cfg::BK ? 4 : (krem + 15) / 16;
          for (int ks = 0; ks < ksteps; ++ks) {
            const uint64_t da = make_smem_desc_sw128(sa + ks * 32, 16, 1024);
            uint64_t db;
            if constexpr (!B_MN) db = make_smem_desc_sw128(sb + ks * 32, 16, 1024);
            else db = make_smem_desc_sw128(sb + ks * (16 * 128), 64 * 128, 1024);
            umma_ss(d_tmem, da, db, idesc, (kc | ks) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs have read it
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1u; }
        }
        umma_commit(&tmem_full[acc]);  // accumulator complete -> epilogue
      }
    }
  } else if (warp >= 4) {
    // ================= Epilogue =================
    const int quarter = warp & 3;
    int it = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
      const int nt = t % ep.tiles_n;
      const int mt = (t / ep.tiles_n) % ep.tiles_m;
      const int z = t / (ep.tiles_n * ep.tiles_m);
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1u;
      const int row = mt * Cfg::BM + quarter * 32 + lane;
      const bool row_ok = row < ep.M;
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + acc * BN + (static_cast<uint32_t>(quarter * 32) << 16);
      const float row_bias = (ep.bias_mode == 2 && row_ok) ? __ldg(ep.bias + row) : 0.0f;
      const float* res_row = ep.residual ? ep.residual + z * ep.strideR + static_cast<long long>(row) * ep.ldr : nullptr;
      float* o32_row = ep.out_f32 ? ep.out_f32 + z * ep.strideO32 + static_cast<long long>(row) * ep.ldo32 : nullptr;
      __nv_bfloat16* o16_row =
          ep.out_bf16 ? ep.out_bf16 + z * ep.strideO16 + static_cast<long long>(row) * ep.ldo16 : nullptr;
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        const int col0 = nt * BN + c * 32;
        if (col0 >= ep.N) break;  // warp-uniform
        uint32_t r[32];
        tmem_ld32(t_row + c * 32, r);
        tmem_wait_ld();
        if (row_ok) {
        const bool full = (col0 + 32 <= ep.N);
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) * ep.alpha + row_bias;
        if (ep.bias_mode == 1) {
          if (full) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] += __ldg(ep.bias + col0 + j);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (col0 + j < ep.N) v[j] += __ldg(ep.bias + col0 + j);
          }
        }
        if (ep.act == 1) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = gelu_erf(v[j]);
        }
        if (res_row) {
          const float* rp = res_row + col0;
          if (full && ((reinterpret_cast<uintptr_t>(rp) & 15u) == 0)) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 q = __ldg(reinterpret_cast<const float4*>(rp) + j);
              v[4 * j] += q.x; v[4 * j + 1] += q.y; v[4 * j + 2] += q.z; v[4 * j + 3] += q.w;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (col0 + j < ep.N) v[j] += __ldg(rp + j);
          }
        }
        if (o32_row) {
          float* op = o32_row + col0;
          if (full && ((reinterpret_cast<uintptr_t>(op) & 15u) == 0)) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
              reinterpret_cast<float4*>(op)[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (col0 + j < ep.N) op[j] = v[j];
          }
        }
        if (o16_row) {
          __nv_bfloat16* op = o16_row + col0;
          if (full && ((reinterpret_cast<uintptr_t>(op) & 15u) == 0)) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint4 q;
              q.x = pack_bf16x2(v[8 * j], v[8 * j + 1]);
              q.y = pack_bf16x2(v[8 * j + 2], v[8 * j + 3]);
              q.z = pack_bf16x2(v[8 * j + 4], v[8 * j + 5]);
              q.w = pack_bf16x2(v[8 * j + 6], v[8 * j + 7]);
              reinterpret_cast<uint4*>(op)[j] = q;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (col0 + j < ep.N) op[j] = __float2bfloat16_rn(v[j]);
          }
        }
        }  // row_ok
      }
      tc_fence_before();
      mbar_arrive(&tmem_empty[acc]);
    }
  }

  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

template <int BN, bool B_MN>
static int launch_gemm(const pio_gemm_args* a, const DeviceInfo& dev, cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
  CUtensorMap ta, tb;
  const bool a_bcast = a->batch > 1 && a->strideA == 0;
  const bool b_bcast = a->batch > 1 && a->strideB == 0;
  const uint64_t a_batch = a_bcast ? 1 : a->batch, b_batch = b_bcast ? 1 : a->batch;
  {
    const uint64_t dims[3] = {(uint64_t)a->K, (uint64_t)a->M, a_batch};
    const uint64_t strides[2] = {(uint64_t)a->lda * 2, (uint64_t)(a_batch > 1 ? a->strideA : a->lda * (int64_t)a->M) * 2};
    const uint32_t box[3] = {64, 128, 1};
    int rc = encode_tmap_bf16(&ta, a->A, 3, dims, strides, box);
    if (rc != PIO_OK) return rc;
  }
  if (!B_MN) {
    const uint64_t dims[3] = {(uint64_t)a->K, (uint64_t)a->N, b_batch};
    const uint64_t strides[2] = {(uint64_t)a->ldb * 2, (uint64_t)(b_batch > 1 ? a->strideB : a->ldb * (int64_t)a->N) * 2};
    const uint32_t box[3] = {64, (uint32_t)BN, 1};
    int rc = encode_tmap_bf16(&tb, a->B, 3, dims, strides, box);
    if (rc != PIO_OK) return rc;
  } else {
    const uint64_t dims[3] = {(uint64_t)a->N, (uint64_t)a->K, b_batch};
    const uint64_t strides[2] = {(uint64_t)a->ldb * 2, (uint64_t)(b_batch > 1 ? a->strideB : a->ldb * (int64_t)a->K) * 2};
    const uint32_t box[3] = {64, 64, 1};
    int rc = encode_tmap_bf16(&tb, a->B, 3, dims, strides, box);
    if (rc != PIO_OK) return rc;
  }
  GemmEpilogue ep;
  ep.M = a->M; ep.N = a->N; ep.K = a->K; ep.batch = a->batch;
  ep.a_bcast = a_bcast ? 1 : 0; ep.b_bcast = b_bcast ? 1 : 0;
  ep.tiles_m = (a->M + 127) / 128;
  ep.tiles_n = (a->N + BN - 1) / BN;
  ep.bias = a->bias; ep.bias_mode = a->bias ? a->bias_mode : 0;
  ep.act = a->act; ep.alpha = a->alpha;
  ep.residual = a->residual; ep.ldr = a->ldr; ep.strideR = a->strideR;
  ep.out_f32 = a->out_f32; ep.ldo32 = a->ldo32; ep.strideO32 = a->strideO32;
  ep.out_bf16 = reinterpret_cast<__nv_bfloat16*>(a->out_bf16); ep.ldo16 = a->ldo16; ep.strideO16 = a->strideO16;

  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(pio_gemm_kernel<BN, B_MN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    Cfg::SMEM_BYTES);
  });
  if (attr_err != cudaSuccess)
    return fail(PIO_ERR_CUDA, "cudaFuncSetAttribute(gemm<%d>) failed: %s", BN, cudaGetErrorString(attr_err));

  const long long total = (long long)ep.tiles_m * ep.tiles_n * ep.batch;
  int ctas = dev.sm_count;
  if (a->max_ctas > 0 && a->max_ctas < ctas) ctas = a->max_ctas;
  if (total < ctas) ctas = (int)total;
  pio_gemm_kernel<BN, B_MN><<<ctas, 256, Cfg::SMEM_BYTES, stream>>>(ta, tb, ep);
  g_launch_count.fetch_add(1);
  PIO_CUDA_OK(cudaGetLastError());
  return PIO_OK;
}

}  // namespace pio

extern "C" int pio_gemm_bf16(const pio_gemm_args* a, void* stream_) {
  using namespace pio;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  PIO_REQUIRE(a != nullptr, "pio_gemm_bf16: null args");
  PIO_REQUIRE(a->A && a->B, "pio_gemm_bf16: null operand");
  PIO_REQUIRE(a->M > 0 && a->N > 0 && a->K > 0 && a->batch > 0, "pio_gemm_bf16: bad shape M=%d N=%d K=%d batch=%d",
              a->M, a->N, a->K, a->batch);
  PIO_REQUIRE(a->out_f32 || a->out_bf16, "pio_gemm_bf16: no output");
  PIO_REQUIRE(aligned16(a->A) && aligned16(a->B), "pio_gemm_bf16: operand base not 16-byte aligned");
  PIO_REQUIRE(a->lda % 8 == 0 && a->ldb % 8 == 0, "pio_gemm_bf16: lda=%lld ldb=%lld must be multiples of 8",
              (long long)a->lda, (long long)a->ldb);
  PIO_REQUIRE(a->lda >= a->K, "pio_gemm_bf16: lda < K");
  PIO_REQUIRE(a->b_mn_major ? a->ldb >= a->N : a->ldb >= a->K, "pio_gemm_bf16: ldb too small");
  PIO_REQUIRE(a->batch == 1 || (a->strideA % 8 == 0 && a->strideB % 8 == 0),
              "pio_gemm_bf16: batch strides must be multiples of 8 elements");
  PIO_REQUIRE(a->bias_mode >= 0 && a->bias_mode <= 2 && (a->act == 0 || a->act == 1), "pio_gemm_bf16: bad epilogue");
  DeviceInfo dev;
  int rc = get_device_info(&dev);
  if (rc != PIO_OK) return rc;
  if (dev.cc_major != 10) return fail(PIO_ERR_ARCH, "pio_gemm_bf16 needs sm_100 (got sm_%d%d)", dev.cc_major, dev.cc_minor);
  int bn = a->tile_n;
  if (bn == 0) bn = a->N > 128 ? 256 : (a->N > 64 ? 128 : 64);
  if (a->b_mn_major) {
    switch (bn) {
      case 256: return launch_gemm<256, true>(a, dev, stream);
      case 128: return launch_gemm<128, true>(a, dev, stream);
      case 64: return launch_gemm<64, true>(a, dev, stream);
    }
  } else {
    switch (bn) {
      case 256: return launch_gemm<256, false>(a, dev, stream);
      case 128: return launch_gemm<128, false>(a, dev, stream);
      case 64: return launch_gemm<64, false>(a, dev, stream);
    }
  }
  return fail(PIO_ERR_INVALID_ARGUMENT, "pio_gemm_bf16: tile_n must be 0, 64, 128 or 256 (got %d)", a->tile_n);
}
