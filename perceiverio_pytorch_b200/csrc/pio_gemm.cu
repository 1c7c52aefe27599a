// Persistent, warp-specialised bf16 GEMM for sm_100a: TMA (SWIZZLE_128B) -> smem ring -> tcgen05.mma with
// double-buffered fp32 accumulators in TMEM -> fused epilogue (alpha, bias, exact GELU, fp32 residual, fp32 and/or
// bf16 stores).  One CTA per SM; tile 128 x BN x 64.
//
// L2 -> SM bandwidth is the binding resource of a 128 x 256 tile (48 KB of operands per 512 tensor cycles), so CTAs are
// launched as clusters of CL along M that work on the same N tile in lockstep: each CTA fetches 1/CL of the B tile and
// TMA-multicasts it into every CTA of the cluster, cutting the B traffic per CTA by CL.
//
// Roles (384 threads): warp 0 = TMA producer, warp 1 = MMA issuer, warp 2 = TMEM allocator, warp 3 idle,
// warps 4..11 = epilogue.  Epilogue warp w owns TMEM lanes 32*(w%4) .. +31 (accumulator rows) and the column half
// (w-4)/4 of the tile.  fp32 inputs/outputs (residual stream) go through a per-warp 32x32 shared-memory transpose
// so that every global access is a full 128-byte row segment; bf16-only outputs are stored straight from the
// accumulator rows (64 contiguous bytes per thread and chunk).
#include "pio_common.cuh"
#include "pio_host.h"

namespace pio {

struct GemmEpilogue {
  int M, N, K, batch;
  int fp16;              // 16-bit operand / output format: 0 = bf16, 1 = fp16
  int tiles_m, tiles_n;
  int m_groups;          // ceil(tiles_m / CL): M tiles are handed out to a cluster CL at a time
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
  // LayerNorm folded into the projections around it (pio_gemm_args; batch == 1): producer side ...
  float* row_stats_out;          // [M][stats_parts][2]: this row's (sum, sum of squares) over one half-tile of columns
  int stats_parts;
  // ... consumer side
  const float* row_stats_in;     // [M][stats_parts][2] of the A operand's rows
  const float* ln_colsum;        // [N]
  float ln_inv_c, ln_eps;
};

// Developer aid (compiled out unless -DPIO_GEMM_TRACE): CTA 0 records (tag, clock64, globaltimer) at the pipeline's
// hand-off points of its latest launch; pio_debug_gemm_trace() copies them out (tools/trace_gemm.py).
#ifdef PIO_GEMM_TRACE
__device__ unsigned long long g_gemm_trace[4 * 128 * 3];
#define GT(slot, tag)                                                                                         \
  do {                                                                                                        \
    if (blockIdx.x == 0 && lane == 0 && gtn < 128) {                                                          \
      unsigned long long gt_ = 0;                                                                             \
      if ((tag) < 10 || (tag) >= 400) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt_));                 \
      g_gemm_trace[((slot) * 128 + gtn) * 3] = (unsigned long long)(tag);                                     \
      g_gemm_trace[((slot) * 128 + gtn) * 3 + 1] = (unsigned long long)clock64();                             \
      g_gemm_trace[((slot) * 128 + gtn) * 3 + 2] = gt_;                                                       \
      ++gtn;                                                                                                  \
    }                                                                                                         \
  } while (0)
#else
#define GT(slot, tag)
#endif

template <int BN>
struct GemmCfg {
  static constexpr int BM = 128;
  static constexpr int BK = 64;
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (BN == 256) ? 4 : (BN == 128 ? 6 : 8);
  static constexpr int TMEM_COLS = 2 * BN;  // 128, 256 or 512: powers of two
  static constexpr int EPI_WARPS = 8;
  static constexpr int XPOSE_PITCH = 33;  // floats; conflict-free for row-wise writes and transposed reads
  static constexpr int XPOSE_BYTES = EPI_WARPS * 32 * XPOSE_PITCH * 4;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + XPOSE_BYTES + 256 /*barriers*/;
};

template <int BN, bool B_MN, int CL>
__global__ void __launch_bounds__(384, 1)
pio_gemm_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                const GemmEpilogue ep) {
  using Cfg = GemmCfg<BN>;
  // SWIZZLE_128B tiles need 1024-byte aligned bases; the kernel has no static shared memory, so the dynamic window
  // starts at the declared alignment.
  extern __shared__ __align__(1024) uint8_t smem[];
  float* xpose = reinterpret_cast<float*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES + Cfg::XPOSE_BYTES);
  uint64_t* full_bar = bars;                      // [STAGES]  TMA -> MMA
  uint64_t* empty_bar = bars + Cfg::STAGES;       // [STAGES]  MMA -> TMA
  uint64_t* tmem_full = bars + 2 * Cfg::STAGES;   // [2]       MMA -> epilogue
  uint64_t* tmem_empty = tmem_full + 2;           // [2]       epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);   // warp-uniform for the compiler too
  const int lane = threadIdx.x & 31;
#ifdef PIO_GEMM_TRACE
  int gtn = 0;
  if (warp == 0) GT(0, 1);
#endif
  if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) {
    printf("pio_gemm_kernel: dynamic shared memory base is not 1024-byte aligned\n");
    __trap();
  }
  const int num_k_chunks = (ep.K + Cfg::BK - 1) / Cfg::BK;
  // cluster-level tile schedule: cluster c works on (n tile, M group, batch) triples; CTA `crank` of the cluster takes
  // M tile group*CL + crank (possibly past the end: zero-filled loads, no stores)
  const int crank = (CL > 1) ? (int)cluster_ctarank() : 0;
  const int cluster_id = blockIdx.x / CL;
  const int num_clusters = gridDim.x / CL;
  const int total_tiles = ep.m_groups * ep.tiles_n * ep.batch;
  constexpr uint16_t kMask = (uint16_t)((1u << CL) - 1u);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < Cfg::STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], CL);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], Cfg::EPI_WARPS * 32);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (CL > 1) cluster_sync_all();  // peers' barriers are initialised before anyone multicasts into them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (warp == 0) GT(0, 2);
  // barriers / TMEM are set up: let the next grid start its prologue; every role that touches global memory waits for
  // the previous grid (pdl_wait) as late as it can, after the index arithmetic of its first tile
  pdl_launch();

  if (warp == 0) {
    // ================= TMA producer =================
    // All 32 lanes run the schedule and the barrier waits, one elected lane issues the copies: inside an
    // `if (lane == 0)` region every TMA / MMA instruction costs a divergent R2UR waterfall (~25 dependent instructions,
    // ~200 clk per MMA measured), which bounds the K loop of the narrow tiles.
    int stage = 0;
    uint32_t phase = 0;
    bool waited = false;
    for (int t = cluster_id; t < total_tiles; t += num_clusters) {
      const int nt = t % ep.tiles_n;
      const int mt = ((t / ep.tiles_n) % ep.m_groups) * CL + crank;
      const int z = t / (ep.tiles_n * ep.m_groups);
      const int m0 = mt * Cfg::BM, n0 = nt * BN;
      const int za = ep.a_bcast ? 0 : z, zb = ep.b_bcast ? 0 : z;
      if (!waited) {
        waited = true;
        pdl_wait();
        GT(0, 3);
      }
      for (int kc = 0; kc < num_k_chunks; ++kc) {
        mbar_wait(&empty_bar[stage], phase ^ 1u);
        GT(0, 100 + kc);
        if (elect_one()) {
          uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
          uint8_t* sb = sa + Cfg::A_BYTES;
          mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
          tma_load_3d(sa, &tmap_a, &full_bar[stage], kc * Cfg::BK, m0, za);
          if constexpr (CL == 1) {
            if constexpr (!B_MN) {
              tma_load_3d(sb, &tmap_b, &full_bar[stage], kc * Cfg::BK, n0, zb);
            } else {
#pragma unroll
              for (int j = 0; j < BN / 64; ++j)
                tma_load_3d(sb + j * (64 * 128), &tmap_b, &full_bar[stage], n0 + j * 64, kc * Cfg::BK, zb);
            }
          } else {
            // this CTA fetches its 1/CL slice of the B tile and multicasts it to the whole cluster
            if constexpr (!B_MN) {
              constexpr int SL = BN / CL;  // rows of B per CTA
              tma_load_3d_mcast(sb + crank * (SL * 128), &tmap_b, &full_bar[stage], kc * Cfg::BK, n0 + crank * SL, zb,
                                kMask);
            } else {
#pragma unroll
              for (int j = 0; j < BN / 64; ++j)
                if (j % CL == crank)
                  tma_load_3d_mcast(sb + j * (64 * 128), &tmap_b, &full_bar[stage], n0 + j * 64, kc * Cfg::BK, zb,
                                    kMask);
            }
          }
        }
        __syncwarp();
        if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    // Warp-converged like the producer; the descriptors advance as 32-bit low words (start address field), the high
    // words (stride / version / swizzle) are loop constants.
    const uint32_t idesc = make_idesc_f16(128, BN, idesc_fmt(ep.fp16), /*a K-major*/ 0, B_MN ? 1 : 0);
    const uint64_t da0 = make_smem_desc_sw128(smem_u32(smem), 16, 1024);
    const uint64_t db0 = B_MN ? make_smem_desc_sw128(smem_u32(smem) + Cfg::A_BYTES, 64 * 128, 1024)
                              : make_smem_desc_sw128(smem_u32(smem) + Cfg::A_BYTES, 16, 1024);
    const uint32_t a_lo0 = (uint32_t)da0, a_hi = (uint32_t)(da0 >> 32);
    const uint32_t b_lo0 = (uint32_t)db0, b_hi = (uint32_t)(db0 >> 32);
    constexpr uint32_t B_KSTEP = B_MN ? (16 * 128) >> 4 : 32 >> 4;   // low-word advance per 16-wide K step
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int t = cluster_id; t < total_tiles; t += num_clusters, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1u;
      mbar_wait(&tmem_empty[acc], acc_phase ^ 1u);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BN;
      for (int kc = 0; kc < num_k_chunks; ++kc) {
        mbar_wait(&full_bar[stage], phase);
        GT(1, 200 + kc);
        tc_fence_after();
        const uint32_t a_lo = a_lo0 + (uint32_t)((stage * Cfg::STAGE_BYTES) >> 4);
        const uint32_t b_lo = b_lo0 + (uint32_t)((stage * Cfg::STAGE_BYTES) >> 4);
        const int krem = ep.K - kc * Cfg::BK;
        if (krem >= Cfg::BK) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            if (elect_one())
              umma_ss_lh(d_tmem, a_lo + ks * 2, a_hi, b_lo + ks * B_KSTEP, b_hi, idesc, (kc | ks) != 0 ? 1u : 0u);
        } else {
          const int ksteps = (krem + 15) / 16;
          for (int ks = 0; ks < ksteps; ++ks)
            if (elect_one())
              umma_ss_lh(d_tmem, a_lo + ks * 2, a_hi, b_lo + ks * B_KSTEP, b_hi, idesc, (kc | ks) != 0 ? 1u : 0u);
        }
        // frees the smem slot (in every CTA that multicasts into it) once these MMAs have read it
        if (elect_one()) {
          if constexpr (CL == 1) umma_commit(&empty_bar[stage]);
          else umma_commit_mcast(&empty_bar[stage], kMask);
        }
        __syncwarp();
        if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1u; }
      }
      if (elect_one()) umma_commit(&tmem_full[acc]);  // accumulator complete -> epilogue
      __syncwarp();
      GT(1, 299);
    }
  } else if (warp >= 4) {
    // ================= Epilogue =================
    const int quarter = warp & 3;
    const int half = (warp - 4) >> 2;                 // which half of the tile's columns
    constexpr int HALF_COLS = BN / 2;
    float* xp = xpose + (warp - 4) * (32 * Cfg::XPOSE_PITCH);
    const bool via_smem = (ep.residual != nullptr) || (ep.out_f32 != nullptr);
    const bool vec32 = ((ep.ldo32 & 3) == 0) && ((reinterpret_cast<uintptr_t>(ep.out_f32) & 15u) == 0) &&
                       ((ep.strideO32 & 3) == 0);
    const bool vecr = ((ep.ldr & 3) == 0) && ((reinterpret_cast<uintptr_t>(ep.residual) & 15u) == 0) &&
                      ((ep.strideR & 3) == 0);
    const bool vec16 = ((ep.ldo16 & 3) == 0) && ((reinterpret_cast<uintptr_t>(ep.out_bf16) & 7u) == 0) &&
                       ((ep.strideO16 & 3) == 0);
    pdl_wait();
    int it = 0;
    for (int t = cluster_id; t < total_tiles; t += num_clusters, ++it) {
      const int nt = t % ep.tiles_n;
      const int mt = ((t / ep.tiles_n) % ep.m_groups) * CL + crank;
      const int z = t / (ep.tiles_n * ep.m_groups);
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1u;
      const int row_base = mt * Cfg::BM + quarter * 32;
      const int row = row_base + lane;
      const bool row_ok = row < ep.M;
      const float* res_z = ep.residual ? ep.residual + z * ep.strideR : nullptr;
      if (res_z != nullptr && row_ok) {
        // pull this warp's 32 x (BN/2) residual block towards L2 while the mainloop of the tile is still running
        const float* rrow = res_z + static_cast<long long>(row) * ep.ldr + nt * BN + half * HALF_COLS;
#pragma unroll
        for (int cc = 0; cc < HALF_COLS; cc += 32)
          if (nt * BN + half * HALF_COLS + cc < ep.N)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(rrow + cc));
      }
      if (warp == 4) GT(2, 300);
      // the epilogue's global inputs of a chunk (residual block in the transposed mapping: lane -> 4 columns of rows
      // i*4 + lane/8; bias and LayerNorm column sums of the chunk's 32 columns): the first chunk's are requested BEFORE
      // the accumulator wait, so a small problem (one chunk per warp, one tile per CTA) does not pay their latency
      // after its K loop
      // (the LayerNorm column sums share the residual's registers: a fused-LayerNorm consumer has no residual input)
      float4 resv[8], biasv[8];
      float4 (&colv)[8] = resv;
      bool res_fast = false, bias_fast = false, col_fast = false;
      auto load_inputs = [&](int c) {
        const int col0 = nt * BN + half * HALF_COLS + c * 32;
        const bool full = (col0 + 32 <= ep.N);
        res_fast = via_smem && res_z != nullptr && vecr && full;
        if (res_fast) {
          const float* rp0 = res_z + static_cast<long long>(row_base + (lane >> 3)) * ep.ldr + col0 + (lane & 7) * 4;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const bool ok = row_base + i * 4 + (lane >> 3) < ep.M;
            resv[i] = ok ? __ldg(reinterpret_cast<const float4*>(rp0 + static_cast<long long>(i * 4) * ep.ldr))
                         : make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
        bias_fast = ep.bias_mode == 1 && full && ((reinterpret_cast<uintptr_t>(ep.bias + col0) & 15u) == 0);
        if (bias_fast) {
#pragma unroll
          for (int j = 0; j < 8; ++j) biasv[j] = __ldg(reinterpret_cast<const float4*>(ep.bias + col0) + j);
        }
        col_fast = !res_fast && ep.row_stats_in != nullptr && full &&
                   ((reinterpret_cast<uintptr_t>(ep.ln_colsum + col0) & 15u) == 0);
        if (col_fast) {
#pragma unroll
          for (int j = 0; j < 8; ++j) colv[j] = __ldg(reinterpret_cast<const float4*>(ep.ln_colsum + col0) + j);
        }
      };
      if (nt * BN + half * HALF_COLS < ep.N) load_inputs(0);
      const float row_bias = (ep.bias_mode == 2 && row_ok) ? __ldg(ep.bias + row) : 0.0f;
      // fused LayerNorm, consumer side: this row's mean / rstd from the partial statistics its producer left (added in a
      // fixed order: bit-reproducible)
      float ln_mean = 0.f, ln_rstd = 1.f;
      if (ep.row_stats_in != nullptr && row_ok) {
        const float2* sp = reinterpret_cast<const float2*>(ep.row_stats_in) + (long long)row * ep.stats_parts;
        float s1 = 0.f, s2 = 0.f;
        for (int j = 0; j < ep.stats_parts; ++j) {
          const float2 st = __ldg(sp + j);
          s1 += st.x;
          s2 += st.y;
        }
        ln_mean = s1 * ep.ln_inv_c;
        ln_rstd = rsqrtf(fmaxf(s2 * ep.ln_inv_c - ln_mean * ln_mean, 0.f) + ep.ln_eps);
      }
      mbar_wait(&tmem_full[acc], acc_phase);
      if (warp == 4) GT(2, 301);
      tc_fence_after();
      const uint32_t t_row = tmem_base + acc * BN + half * HALF_COLS + (static_cast<uint32_t>(quarter * 32) << 16);
      // producer side: partial statistics of the final rows, in the transposed mapping (lane -> 4 columns of rows
      // i * 4 + lane / 8), accumulated over the warp's chunks and reduced over the 8 lanes of a row at the end
      float st_sum[8], st_sq[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) st_sum[i] = st_sq[i] = 0.f;
      float* o32_z = ep.out_f32 ? ep.out_f32 + z * ep.strideO32 : nullptr;
      __nv_bfloat16* o16_z = ep.out_bf16 ? ep.out_bf16 + z * ep.strideO16 : nullptr;
#pragma unroll 1
      for (int c = 0; c < HALF_COLS / 32; ++c) {
        const int col0 = nt * BN + half * HALF_COLS + c * 32;
        if (col0 >= ep.N) break;  // warp-uniform
        const bool full = (col0 + 32 <= ep.N);
        // later chunks: all loads are issued up front so their latency overlaps the TMEM read, the math and the transpose
        if (c > 0) load_inputs(c);
        uint32_t r[32];
        tmem_ld32(t_row + c * 32, r);
        tmem_wait_ld();
        if (warp == 4) GT(2, 310 + c);
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = fmaf(__uint_as_float(r[j]), ep.alpha, row_bias);
        if (ep.row_stats_in != nullptr) {
          // v = rstd * (acc - mean * colsum[n])   (the bias, which already holds W.beta, is added below)
          const float nm = -ln_mean;
          if (col_fast) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              v[4 * j] = ln_rstd * fmaf(nm, colv[j].x, v[4 * j]);
              v[4 * j + 1] = ln_rstd * fmaf(nm, colv[j].y, v[4 * j + 1]);
              v[4 * j + 2] = ln_rstd * fmaf(nm, colv[j].z, v[4 * j + 2]);
              v[4 * j + 3] = ln_rstd * fmaf(nm, colv[j].w, v[4 * j + 3]);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float cs = (col0 + j < ep.N) ? __ldg(ep.ln_colsum + col0 + j) : 0.f;
              v[j] = ln_rstd * fmaf(nm, cs, v[j]);
            }
          }
        }
        if (ep.bias_mode == 1) {
          if (bias_fast) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 bq = biasv[j];
              v[4 * j] += bq.x; v[4 * j + 1] += bq.y; v[4 * j + 2] += bq.z; v[4 * j + 3] += bq.w;
            }
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
        if (warp == 4) GT(2, 320 + c);
        if (!via_smem) {
          // bf16-only output: 64 contiguous bytes per thread
          if (row_ok) {
            __nv_bfloat16* op = o16_z + static_cast<long long>(row) * ep.ldo16 + col0;
            if (full && ((reinterpret_cast<uintptr_t>(op) & 15u) == 0)) {
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                uint4 q;
                q.x = pack16x2(v[8 * j], v[8 * j + 1], ep.fp16);
                q.y = pack16x2(v[8 * j + 2], v[8 * j + 3], ep.fp16);
                q.z = pack16x2(v[8 * j + 4], v[8 * j + 5], ep.fp16);
                q.w = pack16x2(v[8 * j + 6], v[8 * j + 7], ep.fp16);
                reinterpret_cast<uint4*>(op)[j] = q;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (col0 + j < ep.N) reinterpret_cast<uint16_t*>(op)[j] = cvt16(v[j], ep.fp16);
            }
          }
        } else if (full && row_base + 32 <= ep.M && (res_z == nullptr || res_fast) && (o32_z == nullptr || vec32) &&
                   (o16_z == nullptr || vec16)) {
          // whole 32 x 32 block inside the matrix, every pointer vector-aligned (the residual-stream GEMMs of the latent
          // towers): 16-byte shared-memory accesses in an XOR-swizzled layout (group g of row r at r * 8 + (g ^ (r & 7)):
          // conflict-free both ways) and no per-element bounds checks — this loop is on the critical path of the
          // one-tile-per-CTA problems
          float4* xq = reinterpret_cast<float4*>(xp);
#pragma unroll
          for (int g = 0; g < 8; ++g)
            xq[lane * 8 + (g ^ (lane & 7))] = make_float4(v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]);
          __syncwarp();
          const int g = lane & 7;
          float4 w[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int rr = i * 4 + (lane >> 3);
            w[i] = xq[rr * 8 + (g ^ (rr & 7))];
          }
          if (res_z != nullptr) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              w[i].x += resv[i].x; w[i].y += resv[i].y; w[i].z += resv[i].z; w[i].w += resv[i].w;
            }
          }
          if (ep.row_stats_out != nullptr) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              st_sum[i] += (w[i].x + w[i].y) + (w[i].z + w[i].w);
              st_sq[i] = fmaf(w[i].x, w[i].x, fmaf(w[i].y, w[i].y, fmaf(w[i].z, w[i].z, fmaf(w[i].w, w[i].w, st_sq[i]))));
            }
          }
          if (o32_z != nullptr) {
            float* op = o32_z + static_cast<long long>(row_base + (lane >> 3)) * ep.ldo32 + col0 + g * 4;
#pragma unroll
            for (int i = 0; i < 8; ++i) *reinterpret_cast<float4*>(op + static_cast<long long>(i * 4) * ep.ldo32) = w[i];
          }
          if (o16_z != nullptr) {
            __nv_bfloat16* op = o16_z + static_cast<long long>(row_base + (lane >> 3)) * ep.ldo16 + col0 + g * 4;
#pragma unroll
            for (int i = 0; i < 8; ++i)
              *reinterpret_cast<uint2*>(op + static_cast<long long>(i * 4) * ep.ldo16) =
                  make_uint2(pack16x2(w[i].x, w[i].y, ep.fp16), pack16x2(w[i].z, w[i].w, ep.fp16));
          }
          __syncwarp();
        } else {
          // transpose the warp's 32x32 block through shared memory: afterwards lane l holds 4 consecutive columns
          // (l%8)*4.. of rows i*4 + l/8, so a warp instruction touches 4 rows x 128 contiguous bytes
#pragma unroll
          for (int j = 0; j < 32; ++j) xp[lane * Cfg::XPOSE_PITCH + j] = v[j];
          __syncwarp();
          const int cq = (lane & 7) * 4;
          const int gcol = col0 + cq;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int rr = i * 4 + (lane >> 3);
            const int grow = row_base + rr;
            float4 w;
            w.x = xp[rr * Cfg::XPOSE_PITCH + cq];
            w.y = xp[rr * Cfg::XPOSE_PITCH + cq + 1];
            w.z = xp[rr * Cfg::XPOSE_PITCH + cq + 2];
            w.w = xp[rr * Cfg::XPOSE_PITCH + cq + 3];
            if (grow < ep.M && gcol < ep.N) {
              const bool quad = gcol + 4 <= ep.N;
              if (res_fast) {
                w.x += resv[i].x; w.y += resv[i].y; w.z += resv[i].z; w.w += resv[i].w;
              } else if (res_z) {
                const float* rp = res_z + static_cast<long long>(grow) * ep.ldr + gcol;
                {
                  w.x += __ldg(rp);
                  if (gcol + 1 < ep.N) w.y += __ldg(rp + 1);
                  if (gcol + 2 < ep.N) w.z += __ldg(rp + 2);
                  if (gcol + 3 < ep.N) w.w += __ldg(rp + 3);
                }
              }
              if (ep.row_stats_out != nullptr) {
                st_sum[i] += w.x;
                st_sq[i] = fmaf(w.x, w.x, st_sq[i]);
                if (gcol + 1 < ep.N) { st_sum[i] += w.y; st_sq[i] = fmaf(w.y, w.y, st_sq[i]); }
                if (gcol + 2 < ep.N) { st_sum[i] += w.z; st_sq[i] = fmaf(w.z, w.z, st_sq[i]); }
                if (gcol + 3 < ep.N) { st_sum[i] += w.w; st_sq[i] = fmaf(w.w, w.w, st_sq[i]); }
              }
              if (o32_z) {
                float* op = o32_z + static_cast<long long>(grow) * ep.ldo32 + gcol;
                if (quad && vec32) {
                  *reinterpret_cast<float4*>(op) = w;
                } else {
                  op[0] = w.x;
                  if (gcol + 1 < ep.N) op[1] = w.y;
                  if (gcol + 2 < ep.N) op[2] = w.z;
                  if (gcol + 3 < ep.N) op[3] = w.w;
                }
              }
              if (o16_z) {
                __nv_bfloat16* op = o16_z + static_cast<long long>(grow) * ep.ldo16 + gcol;
                if (quad && vec16) {
                  *reinterpret_cast<uint2*>(op) = make_uint2(pack16x2(w.x, w.y, ep.fp16), pack16x2(w.z, w.w, ep.fp16));
                } else {
                  uint16_t* oh = reinterpret_cast<uint16_t*>(op);
                  oh[0] = cvt16(w.x, ep.fp16);
                  if (gcol + 1 < ep.N) oh[1] = cvt16(w.y, ep.fp16);
                  if (gcol + 2 < ep.N) oh[2] = cvt16(w.z, ep.fp16);
                  if (gcol + 3 < ep.N) oh[3] = cvt16(w.w, ep.fp16);
                }
              }
            }
          }
          __syncwarp();
        }
        if (warp == 4) GT(2, 330 + c);
      }
      if (ep.row_stats_out != nullptr) {
        // one slot per (row, half-tile of columns): plain stores, no atomics; the warps of the first column tile also zero
        // the slots beyond the ones this tile width produces (the caller sizes the buffer for the narrowest tiles)
#pragma unroll
        for (int i = 0; i < 8; ++i) {
#pragma unroll
          for (int o = 1; o < 8; o <<= 1) {
            st_sum[i] += __shfl_xor_sync(0xffffffffu, st_sum[i], o);
            st_sq[i] += __shfl_xor_sync(0xffffffffu, st_sq[i], o);
          }
          const int grow = row_base + i * 4 + (lane >> 3);
          if ((lane & 7) == 0 && grow < ep.M) {
            float2* slots = reinterpret_cast<float2*>(ep.row_stats_out) + (long long)grow * ep.stats_parts;
            slots[nt * 2 + half] = make_float2(st_sum[i], st_sq[i]);
            if (nt == 0 && half == 0)
              for (int s = 2 * ep.tiles_n; s < ep.stats_parts; ++s) slots[s] = make_float2(0.f, 0.f);
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&tmem_empty[acc]);
      if (warp == 4) GT(2, 302);
    }
  }

  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if constexpr (CL > 1) cluster_sync_all();  // no CTA exits while a peer may still multicast into / arrive on it
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    GT(3, 400);
  }
}

template <int BN, bool B_MN, int CL>
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
    const uint32_t box[3] = {64, (uint32_t)(BN / CL), 1};
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
  ep.fp16 = a->fp16 ? 1 : 0;
  ep.a_bcast = a_bcast ? 1 : 0; ep.b_bcast = b_bcast ? 1 : 0;
  ep.tiles_m = (a->M + 127) / 128;
  ep.tiles_n = (a->N + BN - 1) / BN;
  ep.m_groups = (ep.tiles_m + CL - 1) / CL;
  ep.bias = a->bias; ep.bias_mode = a->bias ? a->bias_mode : 0;
  ep.act = a->act; ep.alpha = a->alpha;
  ep.residual = a->residual; ep.ldr = a->ldr; ep.strideR = a->strideR;
  ep.out_f32 = a->out_f32; ep.ldo32 = a->ldo32; ep.strideO32 = a->strideO32;
  ep.out_bf16 = reinterpret_cast<__nv_bfloat16*>(a->out_bf16); ep.ldo16 = a->ldo16; ep.strideO16 = a->strideO16;
  ep.row_stats_out = a->row_stats_out;
  ep.row_stats_in = a->row_stats_in;
  ep.stats_parts = a->row_stats_parts > 0 ? a->row_stats_parts : 1;
  ep.ln_colsum = a->ln_colsum;
  ep.ln_inv_c = a->ln_channels > 0 ? 1.0f / (float)a->ln_channels : 0.f;
  ep.ln_eps = a->ln_eps;
  if (a->row_stats_out && ep.stats_parts < 2 * ep.tiles_n)
    return fail(PIO_ERR_INVALID_ARGUMENT, "pio_gemm_bf16: row_stats_out needs row_stats_parts >= %d (got %d)", 2 * ep.tiles_n,
                a->row_stats_parts);

  static PerDeviceOnce once;
  const cudaError_t attr_err = once.run(dev.device, [] {
    return cudaFuncSetAttribute(pio_gemm_kernel<BN, B_MN, CL>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                Cfg::SMEM_BYTES);
  });
  if (attr_err != cudaSuccess)
    return fail(PIO_ERR_CUDA, "cudaFuncSetAttribute(gemm<%d>) failed: %s", BN, cudaGetErrorString(attr_err));

  const long long total = (long long)ep.m_groups * ep.tiles_n * ep.batch;  // cluster-level tiles
  int clusters = dev.sm_count / CL;
  if (a->max_ctas > 0 && a->max_ctas / CL < clusters) clusters = a->max_ctas / CL > 0 ? a->max_ctas / CL : 1;
  if (total < clusters) clusters = (int)total;
  {
    ProfileScope prof(KF_GEMM, 2.0 * a->M * a->N * (double)a->K * a->batch, 0.0, stream);
    PIO_CUDA_OK(launch_kernel(pio_gemm_kernel<BN, B_MN, CL>, dim3((unsigned)(clusters * CL), 1, 1), dim3(384, 1, 1),
                              Cfg::SMEM_BYTES, stream, CL, ta, tb, ep));
  }
  g_launch_count.fetch_add(1);
  PIO_CUDA_OK(cudaGetLastError());
  return PIO_OK;
}

}  // namespace pio

namespace pio {
// Tile selection of the automatic dispatch, shared by pio_gemm_bf16 and pio_gemm_stats_parts.
static bool auto_wants_pair(long long M, long long N, long long batch, int sm_count) {
  const long long pair_tiles = ((M + 255) / 256) * ((N + 255) / 256) * batch;
  return pair_tiles * 3 >= (long long)(sm_count / 2);
}
static int auto_tile_n(long long M, long long N, long long batch, int sm_count) {
  int bn = N > 128 ? 256 : (N > 64 ? 128 : 64);
  // small problems are bound by how many CTAs stream the weight matrix concurrently, not by tile efficiency: shrink
  // the tile until at least half of the SMs have one (narrower tiles also get deeper operand rings)
  const long long tiles_m = (M + 127) / 128 * batch;
  while (bn > 64 && tiles_m * ((N + bn - 1) / bn) * 2 < sm_count) bn >>= 1;
  return bn;
}
}  // namespace pio

// Number of (sum, sum of squares) slots per row that the fused-LayerNorm producer GEMM of this shape writes with the
// automatic kernel / tile choice (pio_gemm_args.row_stats_parts): two per column tile.
#ifdef PIO_GEMM_TRACE
extern "C" int pio_debug_gemm_trace(unsigned long long* out) {   // out: 4 * 128 * 3 words; clears the device buffer
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out, pio::g_gemm_trace, sizeof(pio::g_gemm_trace));
  static unsigned long long zeros[4 * 128 * 3];
  cudaMemcpyToSymbol(pio::g_gemm_trace, zeros, sizeof(zeros));
  return 0;
}
#endif

extern "C" int pio_gemm_stats_parts(int32_t M, int32_t N) {
  using namespace pio;
  DeviceInfo dev;
  int sm = 148;
  if (get_device_info(&dev) == PIO_OK && dev.sm_count > 0) sm = dev.sm_count;
  if (sm % 2 == 0 && auto_wants_pair(M, N, 1, sm)) return 4 * ((N + 255) / 256);   // the stream kernel's 64-column slices
  const int bn = auto_tile_n(M, N, 1, sm);
  return 2 * ((N + bn - 1) / bn);
}

extern "C" int pio_gemm_pair_kernel(int32_t M, int32_t N) {
  using namespace pio;
  DeviceInfo dev;
  int sm = 148;
  if (get_device_info(&dev) == PIO_OK && dev.sm_count > 0) sm = dev.sm_count;
  return (sm % 2 == 0 && auto_wants_pair(M, N, 1, sm)) ? 1 : 0;
}

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
  PIO_REQUIRE(a->kernel >= 0 && a->kernel <= 2, "pio_gemm_bf16: kernel must be 0 (auto), 1 or 2 (got %d)", a->kernel);
  {
    // CTA-pair kernel: when explicitly requested, or when there are enough 256 x 256 tiles to occupy a third of the SM
    // pairs (per tile it is ~1.4x faster than the single-CTA kernel, so it wins well before the grid is full)
    const bool elig = gemm2_eligible(a) && (dev.sm_count % 2 == 0);
    if (a->kernel == 2 && !elig)
      return fail(PIO_ERR_UNSUPPORTED, "pio_gemm_bf16: the CTA-pair kernel needs K-major B, exactly one output and "
                                       "16-byte aligned output / residual rows");
    const bool want = a->kernel == 2 ||
                      (a->kernel == 0 && a->tile_n == 0 && a->cluster_m == 0 && auto_wants_pair(a->M, a->N, a->batch, dev.sm_count));
    if (a->out_lo16 || a->residual_hi16 || a->residual_lo16) {
      // the 16-bit pair residual stream exists in the CTA-pair kernel only
      if (!elig || a->kernel == 1)
        return fail(PIO_ERR_UNSUPPORTED, "pio_gemm_bf16: out_lo16 / residual_hi16 / residual_lo16 need the CTA-pair kernel "
                                         "(batch 1, K-major B, 16-byte aligned rows, out_bf16 with out_lo16, both residual "
                                         "halves, no fp32 form of the same operand)");
      return launch_gemm2(a, dev, stream);
    }
    if (elig && want) return launch_gemm2(a, dev, stream);
    // the single-CTA kernel has the fused-LayerNorm epilogues too (small latent arrays), with these limits:
    if (a->row_stats_out || a->row_stats_in) {
      if (a->batch != 1 || a->b_mn_major)
        return fail(PIO_ERR_UNSUPPORTED, "pio_gemm_bf16: the fused-LayerNorm epilogues need batch == 1 and a K-major B");
      if (a->row_stats_out && !(a->out_f32 || a->residual))
        return fail(PIO_ERR_UNSUPPORTED, "pio_gemm_bf16: row_stats_out belongs to the GEMM that writes the fp32 residual stream");
      if (a->row_stats_in && (a->out_f32 || a->residual || !a->ln_colsum || a->ln_channels <= 0))
        return fail(PIO_ERR_UNSUPPORTED, "pio_gemm_bf16: row_stats_in needs a 16-bit-only output, ln_colsum and ln_channels");
    }
  }
  int bn = a->tile_n;
  if (bn == 0) bn = auto_tile_n(a->M, a->N, a->batch, dev.sm_count);
  // cluster width along M: multicast pays when there are at least two M tiles to pair up
  int cl = a->cluster_m;
  // (not for the narrow tiles of small problems: their K loop is bound by the MMA issue stream, not by operand traffic,
  // and a cluster costs ~1000 clk of start-up and ~1200 clk of tear-down synchronisation per launch)
  if (cl == 0) cl = (a->M > 128 && bn > 64) ? 2 : 1;
  PIO_REQUIRE(cl == 1 || cl == 2 || cl == 4, "pio_gemm_bf16: cluster_m must be 0, 1, 2 or 4 (got %d)", a->cluster_m);
  if (cl == 4 && bn == 64 && a->b_mn_major) cl = 2;  // a 64-column MN-major tile is a single TMA box
#define PIO_GEMM_DISPATCH(BN_, MN_)                                            \
  switch (cl) {                                                                \
    case 1: return launch_gemm<BN_, MN_, 1>(a, dev, stream);                   \
    case 2: return launch_gemm<BN_, MN_, 2>(a, dev, stream);                   \
    case 4: return launch_gemm<BN_, MN_, 4>(a, dev, stream);                   \
  }
  if (a->b_mn_major) {
    switch (bn) {
      case 256: PIO_GEMM_DISPATCH(256, true) break;
      case 128: PIO_GEMM_DISPATCH(128, true) break;
      case 64: if (cl == 4) cl = 2; PIO_GEMM_DISPATCH(64, true) break;
    }
  } else {
    switch (bn) {
      case 256: PIO_GEMM_DISPATCH(256, false) break;
      case 128: PIO_GEMM_DISPATCH(128, false) break;
      case 64: PIO_GEMM_DISPATCH(64, false) break;
    }
  }
#undef PIO_GEMM_DISPATCH
  return fail(PIO_ERR_INVALID_ARGUMENT, "pio_gemm_bf16: tile_n must be 0, 64, 128 or 256 (got %d)", a->tile_n);
}
