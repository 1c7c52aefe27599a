// CTA-pair bf16 GEMM for sm_100a: the fast path of pio_gemm_bf16 for the big projection / MLP products.
//
// Two CTAs of a 2-wide cluster (one TPC) compute one 256 x 256 output tile with tcgen05.mma.cta_group::2 (M = 256):
// each CTA stages its own 128 rows of A and only HALF of the B tile (128 of the 256 weight rows), so per 128-cycle MMA
// step an SM reads 8 KB of operands from shared memory and receives 8 KB from TMA instead of 12 KB + 12 KB for a
// single-CTA 128 x 256 tile.  Shared-memory bandwidth (128 B/clk/SM, shared by the tensor core's operand reads, the TMA
// fills and the epilogue staging) is what caps the single-CTA kernel at ~62 % of the tensor peak.
//
// The epilogue never touches global memory with per-thread accesses: the fp32 residual tile is prefetched by TMA into
// a per-warp shared-memory ring two chunks ahead (across tile boundaries, so HBM latency is hidden behind the
// mainloop), and results leave through a swizzled per-warp staging slot and TMA stores, which also clip the M / N
// tails.  Each epilogue warp runs its own pipeline (no CTA-wide barrier in the steady state).
//
// Roles (384 threads per CTA): warp 0 TMA producer (both CTAs), warp 1 MMA issuer (leader CTA only), warp 2 TMEM
// allocator, warp 3 idle, warps 4..11 epilogue: warp w owns TMEM lanes 32*(w%4).. (accumulator rows) and the
// 128-column half (w-4)/4 of the tile.
#include <stdlib.h>

#include <type_traits>

#include "pio_common.cuh"
#include "pio_host.h"

namespace pio {

struct Gemm2Params {
  int M, N, K, batch;
  int fp16;                      // 16-bit operand / output format: 0 = bf16, 1 = fp16
  int tiles_n, m_pairs;          // 256-column tiles, 256-row CTA-pair tiles
  int a_bcast, b_bcast, r_bcast; // operand shared by every batch entry
  const float* bias;
  int bias_mode;                 // 0 none, 1 per column, 2 per row
  int act;
  float alpha;
  int has_residual;
  int b_swap;                    // debug: which CTA of the pair stages which half of the B tile
  int reverse;                   // walk the tiles from the last M rows to the first (pio_gemm_args.reverse_tiles)
  // LayerNorm fusion (pio_gemm_args): producer side ...
  float* row_stats_out;          // [M][stats_parts][2]: (sum, sum of squares) of this row over one 128-column half-tile
  int stats_parts;               // partials per row: 2 * tiles_n on the producer side, whatever the producer wrote on the consumer side
  __nv_bfloat16* raw_bf16;       // bf16 copy of the fp32 output (un-normalised rows), row pitch ld_raw
  long long ld_raw;
  // residual stream as a pair of 16-bit arrays (pio_gemm_args.out_lo16 / residual_hi16 / residual_lo16)
  int out_split;                 // outputs: hi through tmap_raw, lo through tmap_out (no fp32 output)
  int res_split;                 // residual: hi through tmap_res, lo through tmap_res_lo
  // ... consumer side
  const float* row_stats_in;     // [M][stats_parts][2] of the A operand's rows
  const float* ln_colsum;        // [N]
  float ln_inv_c, ln_eps;
  // L2 eviction priorities ("what L2 holds"): the 16-bit hi half of the residual stream of a large
  // tower (67 MB at batch 64) is read four times and rewritten twice per layer — it is kept (evict_last) while the
  // streams that are produced once and consumed once by the next kernel (QKV, attention output, MLP hidden) pass through
  uint64_t hint_a, hint_b, hint_out, hint_res_hi, hint_res_lo, hint_out_hi, hint_out_lo;
};

enum { G2_BF16 = 0, G2_F32 = 1 };
#ifndef PIO_L2_HINTS_DEFAULT
#define PIO_L2_HINTS_DEFAULT 0
#endif

// fp32-output kind: operand stages vs depth of the per-warp residual prefetch ring (both live in shared memory)
// G2_F32_CHUNK: columns per epilogue chunk (16: 64-byte fp32 rows per TMA box, 32: 128-byte rows — half the TMA row
// requests per byte; the copy engine's request rate is what bounds this kind)
#ifndef G2_F32_CHUNK
#define G2_F32_CHUNK 32
#endif
#ifndef G2_F32_STAGES
#define G2_F32_STAGES 4
#endif
#ifndef G2_F32_RES_SLOTS
#define G2_F32_RES_SLOTS (G2_F32_CHUNK == 32 ? 1 : 2)
#endif
#ifndef G2_F32_OUT_SLOTS
#define G2_F32_OUT_SLOTS (G2_F32_CHUNK == 32 ? 1 : 2)
#endif

// Developer aid (compiled out unless -DPIO_GEMM2_TRACE): warp 4 of CTA 0 records (tag, clock64) pairs of its epilogue;
// pio_debug_gemm2_trace() copies them out (tools/trace_gemm2.py).
#ifdef PIO_GEMM2_TRACE
#ifndef PIO_G2T_WARP
#define PIO_G2T_WARP 4      // which epilogue warp records (4 .. 11; -DPIO_G2T_WARP=6: one that shares its scheduler with no role warp)
#endif
__device__ unsigned long long g_gemm2_trace[1024 * 2];
#define G2T(tag)                                                                       \
  do {                                                                                 \
    if (blockIdx.x == 0 && warp == PIO_G2T_WARP && lane == 0 && g2n < 512) {          \
      g_gemm2_trace[g2n * 2] = (unsigned long long)(tag);                              \
      g_gemm2_trace[g2n * 2 + 1] = (unsigned long long)clock64();                      \
      ++g2n;                                                                           \
    }                                                                                  \
  } while (0)
// ... and the MMA issuer of CTA 0 (second half of the buffer): 500 before / 501 after the accumulator-free wait, 502 all
// MMAs of the tile issued
#define G2TM(tag)                                                                      \
  do {                                                                                 \
    if (blockIdx.x == 0 && lane == 0 && g2n < 512) {                                   \
      g_gemm2_trace[(512 + g2n) * 2] = (unsigned long long)(tag);                      \
      g_gemm2_trace[(512 + g2n) * 2 + 1] = (unsigned long long)clock64();              \
      ++g2n;                                                                           \
    }                                                                                  \
  } while (0)
#else
#define G2T(tag)
#define G2TM(tag)
#endif

template <int KIND>
struct Gemm2Cfg {
  static constexpr int BM = 128;       // rows per CTA (256 per pair)
  static constexpr int BN = 256;       // columns per pair tile
  static constexpr int BK = 64;
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = (BN / 2) * BK * 2;   // this CTA's half of the B tile
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (KIND == G2_F32) ? G2_F32_STAGES : 5;
  static constexpr int EPI_WARPS = 8;
  // staging slots: bf16 output 32 rows x 64 columns (128-byte rows, SWIZZLE_128B); fp32 32 rows x 16 columns (64-byte
  // rows, SWIZZLE_64B) so that two residual + two output slots per warp leave room for the fifth operand stage
  static constexpr int SLOT_BYTES = (KIND == G2_F32) ? 32 * G2_F32_CHUNK * 4 : 4096;
  static constexpr int RES_SLOTS = (KIND == G2_F32) ? G2_F32_RES_SLOTS : 0;
  static constexpr int OUT_SLOTS = (KIND == G2_F32) ? G2_F32_OUT_SLOTS : 2;
  // fused-LayerNorm producer: raw bf16 copy of the fp32 output, 32 rows x 16 columns (32-byte rows, no swizzle)
  static constexpr int RAW_SLOT_BYTES = 32 * G2_F32_CHUNK * 2;
  // (with the pair stream the fp32 slot holds two lo buffers, so two raw (hi) slots make that output path double-buffered)
  static constexpr int RAW_SLOTS = (KIND == G2_F32) ? (G2_F32_OUT_SLOTS > 2 ? G2_F32_OUT_SLOTS : 2) : 0;
  static constexpr int WARP_EPI_BYTES = (RES_SLOTS + OUT_SLOTS) * SLOT_BYTES + RAW_SLOTS * RAW_SLOT_BYTES;
  static constexpr int EPI_BYTES = EPI_WARPS * WARP_EPI_BYTES;
  static constexpr int BAR_BYTES = 512;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_BYTES + BAR_BYTES;
  static constexpr int TMEM_COLS = 512;                // two 256-column fp32 accumulators
  static_assert(SMEM_BYTES <= 232448, "shared memory budget exceeded");
};

template <int KIND>
__global__ void __launch_bounds__(384, 1)
pio_gemm2_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                 const __grid_constant__ CUtensorMap tmap_out, const __grid_constant__ CUtensorMap tmap_res,
                 const __grid_constant__ CUtensorMap tmap_raw, const __grid_constant__ CUtensorMap tmap_res_lo,
                 const Gemm2Params p) {
  using Cfg = Gemm2Cfg<KIND>;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* epi_base = smem + Cfg::STAGES * Cfg::STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(epi_base + Cfg::EPI_BYTES);
  uint64_t* full_bar = bars;                        // [STAGES]  TMA (both CTAs) -> MMA; only the leader's are used
  uint64_t* empty_bar = full_bar + Cfg::STAGES;     // [STAGES]  MMA -> TMA producer of each CTA (multicast commit)
  uint64_t* tmem_full = empty_bar + Cfg::STAGES;    // [2]       MMA -> epilogue of each CTA (multicast commit)
  uint64_t* tmem_empty = tmem_full + 2;             // [2]       epilogue warps of both CTAs -> MMA (leader's)
  constexpr int RS = Cfg::RES_SLOTS > 0 ? Cfg::RES_SLOTS : 1;
  uint64_t* res_full = tmem_empty + 2;              // [EPI_WARPS][RES_SLOTS]  residual TMA -> epilogue warp
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_full + Cfg::EPI_WARPS * RS);

  // the shuffle makes the warp index provably warp-uniform, so the role code can use the uniform datapath
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
#ifdef PIO_GEMM2_TRACE
  int g2n = 0;
#endif
  if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) {
    printf("pio_gemm2_kernel: dynamic shared memory base is not 1024-byte aligned\n");
    __trap();
  }
  const uint32_t crank = cluster_ctarank();        // 0 = leader
  const int pair_id = blockIdx.x >> 1;
  const int num_pairs = gridDim.x >> 1;
  const int total_tiles = p.m_pairs * p.tiles_n * p.batch;
  const int num_k_chunks = (p.K + Cfg::BK - 1) / Cfg::BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    tma_prefetch_desc(&tmap_out);
    if (KIND == G2_F32 && p.has_residual) tma_prefetch_desc(&tmap_res);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < Cfg::STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 2 * Cfg::EPI_WARPS);   // one arrival per epilogue warp of either CTA
    }
    for (int i = 0; i < Cfg::EPI_WARPS * RS; ++i) mbar_init(&res_full[i], 1);
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc_2cta(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish_2cta();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // the peer's barriers are initialised before anything is signalled across the pair
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_sync();   // barriers / TMEM are set up; operands and residuals of the previous kernel are read from here on

  if (warp == 0) {
    // ================= TMA producer (both CTAs) =================
    // (the whole warp runs the schedule and waits; one elected lane issues — keeps coordinates in uniform registers)
    int stage = 0;
    uint32_t phase = 0;
    for (int seq = pair_id; seq < total_tiles; seq += num_pairs) {
      const int t = p.reverse ? total_tiles - 1 - seq : seq;
      const int nt = t % p.tiles_n;
      const int mp = (t / p.tiles_n) % p.m_pairs;
      const int z = t / (p.tiles_n * p.m_pairs);
      const int m0 = mp * 256 + (int)crank * Cfg::BM;
      const int n0 = nt * Cfg::BN + (int)(crank ^ (uint32_t)p.b_swap) * (Cfg::BN / 2);
      const int za = p.a_bcast ? 0 : z, zb = p.b_bcast ? 0 : z;
      for (int kc = 0; kc < num_k_chunks; ++kc) {
        mbar_wait(&empty_bar[stage], phase ^ 1u);
        if (elect_one()) {
          uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
          uint8_t* sb = sa + Cfg::A_BYTES;
          // both CTAs' bytes are credited to the leader's barrier, which expects the pair's total
          if (crank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * Cfg::STAGE_BYTES);
          const uint32_t leader_bar = mapa_u32(&full_bar[stage], 0);
          tma_load_3d_2cta(sa, &tmap_a, leader_bar, kc * Cfg::BK, m0, za, p.hint_a);
          tma_load_3d_2cta(sb, &tmap_b, leader_bar, kc * Cfg::BK, n0, zb, p.hint_b);
        }
        if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    if (crank == 0) {
      // ================= MMA issuer (leader CTA) =================
      // All 32 lanes run the schedule and the barrier waits; one elected lane issues each tcgen05 instruction, and the
      // descriptors advance as 32-bit low words (inside an `if (lane == 0)` region every MMA costs ~25 dependent
      // vector instructions + R2UR moves, about as long as the 128-cycle MMA itself).
      const uint32_t idesc = make_idesc_f16(256, Cfg::BN, idesc_fmt(p.fp16), 0, 0);
      const uint64_t d0 = make_smem_desc_sw128(smem_u32(smem), 16, 1024);
      const uint32_t d_lo = (uint32_t)d0, d_hi = (uint32_t)(d0 >> 32);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int t = pair_id; t < total_tiles; t += num_pairs, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1u;
        G2TM(500);
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1u);
        G2TM(501);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * Cfg::BN;
        for (int kc = 0; kc < num_k_chunks; ++kc) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_lo = d_lo + (uint32_t)((stage * Cfg::STAGE_BYTES) >> 4);
          const uint32_t b_lo = a_lo + (uint32_t)(Cfg::A_BYTES >> 4);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            if (elect_one())
              umma_ss_2cta_lh(d_tmem, a_lo + ks * 2, d_hi, b_lo + ks * 2, d_hi, idesc, (kc | ks) != 0 ? 1u : 0u);
          }
          if (elect_one()) umma_commit_2cta_mcast(&empty_bar[stage], 0x3);   // frees the stage in both CTAs
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1u; }
        }
        if (elect_one()) umma_commit_2cta_mcast(&tmem_full[acc], 0x3);       // accumulator halves complete in both CTAs
        G2TM(502);
      }
    }
  } else if (warp >= 4) {
    // ================= Epilogue (both CTAs) =================
    const int ew = warp - 4;
    const int quarter = warp & 3;
    const int half = ew >> 2;
    uint8_t* my = epi_base + ew * Cfg::WARP_EPI_BYTES;
    uint8_t* my_out = my + Cfg::RES_SLOTS * Cfg::SLOT_BYTES;
    uint8_t* my_raw = my_out + Cfg::OUT_SLOTS * Cfg::SLOT_BYTES;
    uint64_t* my_res_full = res_full + ew * RS;
    const bool has_res = (KIND == G2_F32) && p.has_residual;
    constexpr int CHUNK_COLS = (KIND == G2_F32) ? G2_F32_CHUNK : 64;
    constexpr int CHUNKS = 128 / CHUNK_COLS;           // chunks of this warp's column half per tile
    const uint32_t tmem_empty_leader0 = mapa_u32(&tmem_empty[0], 0);
    const uint32_t tmem_empty_leader1 = mapa_u32(&tmem_empty[1], 0);

    // residual prefetch cursor: the stream of (tile, chunk) boxes this warp will consume, RES_SLOTS ahead.  Run by the
    // whole warp with one elected lane per TMA instruction (inside an `if (lane == 0)` region the four copies of a chunk
    // and the cursor arithmetic cost 800 - 1900 clk of a 4100-clk chunk); the tile coordinates — three integer
    // divisions — are computed once per tile, not per chunk
    int pf_t = pair_id, pf_c = 0, pf_row = 0, pf_col = 0, pf_z = 0;
    uint32_t pf_idx = 0;
    auto issue_res = [&]() {
      if (pf_t >= total_tiles) return;
      if (pf_c == 0) {
        const int t = p.reverse ? total_tiles - 1 - pf_t : pf_t;
        const int nt = t % p.tiles_n;
        const int mp = (t / p.tiles_n) % p.m_pairs;
        pf_z = p.r_bcast ? 0 : t / (p.tiles_n * p.m_pairs);
        pf_row = mp * 256 + (int)crank * Cfg::BM + quarter * 32;
        pf_col = nt * Cfg::BN + half * 128;
      }
      const int col = pf_col + pf_c * CHUNK_COLS;
      const uint32_t slot = pf_idx % (uint32_t)RS;
      if (elect_one()) {
        mbar_arrive_expect_tx(&my_res_full[slot], Cfg::SLOT_BYTES);
        if (KIND == G2_F32 && p.res_split) {
          // hi and lo boxes (32 rows x CHUNK_COLS 16-bit values each) fill the two halves of the fp32-sized slot
          tma_load_3d(my + slot * Cfg::SLOT_BYTES, &tmap_res, &my_res_full[slot], col, pf_row, 0, p.hint_res_hi);
          tma_load_3d(my + slot * Cfg::SLOT_BYTES + Cfg::SLOT_BYTES / 2, &tmap_res_lo, &my_res_full[slot], col, pf_row, 0, p.hint_res_lo);
        } else {
          tma_load_3d(my + slot * Cfg::SLOT_BYTES, &tmap_res, &my_res_full[slot], col, pf_row, pf_z);
        }
      }
      __syncwarp();
      ++pf_idx;
      if (++pf_c == CHUNKS) { pf_c = 0; pf_t += num_pairs; }
    };
    if (has_res) {
      for (int i = 0; i < RS; ++i) issue_res();
    }
    uint32_t use_idx = 0;
    int it = 0;
    for (int seq = pair_id; seq < total_tiles; seq += num_pairs, ++it) {
      const int t = p.reverse ? total_tiles - 1 - seq : seq;
      const int nt = t % p.tiles_n;
      const int mp = (t / p.tiles_n) % p.m_pairs;
      const int z = t / (p.tiles_n * p.m_pairs);
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1u;
      const int row0 = mp * 256 + (int)crank * Cfg::BM + quarter * 32;   // first row of this warp's block
      const int row = row0 + lane;
      const float row_bias = (p.bias_mode == 2 && row < p.M) ? __ldg(p.bias + row) : 0.0f;
      // fused LayerNorm, consumer side: this row's mean / rstd from the statistics its producer accumulated
      float ln_mean = 0.f, ln_rstd = 1.f;
      if (KIND == G2_BF16 && p.row_stats_in != nullptr && row < p.M) {
        // the producer left one (sum, sum of squares) per 128-column half-tile: added up in a fixed order, so the
        // statistics (and with them the whole tower) are bit-reproducible from run to run
        const float2* sp = reinterpret_cast<const float2*>(p.row_stats_in) + (long long)row * p.stats_parts;
        float s1 = 0.f, s2 = 0.f;
        for (int j = 0; j < p.stats_parts; ++j) {
          const float2 st = __ldg(sp + j);
          s1 += st.x;
          s2 += st.y;
        }
        ln_mean = s1 * p.ln_inv_c;
        ln_rstd = rsqrtf(fmaxf(s2 * p.ln_inv_c - ln_mean * ln_mean, 0.f) + p.ln_eps);
      }
      float st_sum = 0.f, st_sq = 0.f;   // producer side: this row's partial statistics over the tile
      // 16-bit-output kind: the per-column epilogue parameters (bias, LayerNorm column sums) of this warp's 128 columns
      // are fetched ONCE per tile, four columns per lane, before the accumulator wait, and handed round by shuffles.
      // (Loaded per chunk — 32 broadcast float4 loads, in as many batches as the register file allows — they put
      // several global-memory round trips into every chunk: 2900 - 3400 clk per 64 columns in a clock64 trace of fc1,
      // and the MMA issuer waited 5000 clk per tile for a free accumulator.)
      float4 col_cs = make_float4(0.f, 0.f, 0.f, 0.f), col_b = make_float4(0.f, 0.f, 0.f, 0.f);
      if constexpr (KIND == G2_BF16) {
        const int cb = nt * Cfg::BN + half * 128 + 4 * lane;
        auto fetch4 = [&](const float* src) {
          float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
          if (cb + 3 < p.N && ((reinterpret_cast<uintptr_t>(src + cb) & 15u) == 0)) {
            r = __ldg(reinterpret_cast<const float4*>(src + cb));
          } else {
            if (cb < p.N) r.x = __ldg(src + cb);
            if (cb + 1 < p.N) r.y = __ldg(src + cb + 1);
            if (cb + 2 < p.N) r.z = __ldg(src + cb + 2);
            if (cb + 3 < p.N) r.w = __ldg(src + cb + 3);
          }
          return r;
        };
        if (p.bias_mode == 1) col_b = fetch4(p.bias);
        if (p.row_stats_in != nullptr) col_cs = fetch4(p.ln_colsum);
      }
      G2T(100);
      mbar_wait(&tmem_full[acc], acc_phase);
      G2T(101);
      tc_fence_after();
      const uint32_t t_row = tmem_base + acc * Cfg::BN + half * 128 + (static_cast<uint32_t>(quarter * 32) << 16);
#pragma unroll 1
      for (int c = 0; c < CHUNKS; ++c) {
        const int col0 = nt * Cfg::BN + half * 128 + c * CHUNK_COLS;
        float v[CHUNK_COLS];
        if constexpr (CHUNK_COLS == 64) {
          uint32_t r[32], r2[32];
          tmem_ld32(t_row + c * CHUNK_COLS, r);
          tmem_ld32(t_row + c * CHUNK_COLS + 32, r2);
          tmem_wait_ld();
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            v[j] = fmaf(__uint_as_float(r[j]), p.alpha, row_bias);
            v[32 + j] = fmaf(__uint_as_float(r2[j]), p.alpha, row_bias);
          }
        } else if constexpr (CHUNK_COLS == 32) {
          uint32_t r[32];
          tmem_ld32(t_row + c * CHUNK_COLS, r);
          tmem_wait_ld();
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = fmaf(__uint_as_float(r[j]), p.alpha, row_bias);
        } else {
          uint32_t r[16];
          tmem_ld16(t_row + c * CHUNK_COLS, r);
          tmem_wait_ld();
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = fmaf(__uint_as_float(r[j]), p.alpha, row_bias);
        }
        G2T(310 + c);
        if (c == CHUNKS - 1) {
          // this warp has read its whole share of the accumulator: hand the TMEM buffer back to the MMA issuer
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(acc ? tmem_empty_leader1 : tmem_empty_leader0);
        }
        bool bias_done = false;
        if constexpr (KIND == G2_BF16) {
          // v = acc * rstd + ((-mean * rstd) * colsum[n] + bias[n])  (bias already holds W.beta), or v = acc + bias[n]:
          // two packed fp32x2 FMAs per column pair, the column parameters out of the lanes that hold them
          const bool ln = p.row_stats_in != nullptr;
          if (ln || p.bias_mode == 1) {
            const float nmr = -ln_mean * ln_rstd;
            const uint64_t nmr2 = pack_f32x2(nmr, nmr), rstd2 = pack_f32x2(ln_rstd, ln_rstd);
#pragma unroll
            for (int j = 0; j < CHUNK_COLS / 4; ++j) {
              const int src = c * (CHUNK_COLS / 4) + j;      // the lane that holds columns 4 src .. 4 src + 3 of the warp's 128
              const float b0 = __shfl_sync(0xffffffffu, col_b.x, src), b1 = __shfl_sync(0xffffffffu, col_b.y, src);
              const float b2 = __shfl_sync(0xffffffffu, col_b.z, src), b3 = __shfl_sync(0xffffffffu, col_b.w, src);
              uint64_t add_lo = pack_f32x2(b0, b1), add_hi = pack_f32x2(b2, b3);
              if (ln) {
                const float c0 = __shfl_sync(0xffffffffu, col_cs.x, src), c1 = __shfl_sync(0xffffffffu, col_cs.y, src);
                const float c2 = __shfl_sync(0xffffffffu, col_cs.z, src), c3 = __shfl_sync(0xffffffffu, col_cs.w, src);
                add_lo = ffma2(nmr2, pack_f32x2(c0, c1), add_lo);
                add_hi = ffma2(nmr2, pack_f32x2(c2, c3), add_hi);
              }
              const uint64_t lo = ffma2(pack_f32x2(v[4 * j], v[4 * j + 1]), rstd2, add_lo);
              const uint64_t hi = ffma2(pack_f32x2(v[4 * j + 2], v[4 * j + 3]), rstd2, add_hi);
              unpack_f32x2(lo, v[4 * j], v[4 * j + 1]);
              unpack_f32x2(hi, v[4 * j + 2], v[4 * j + 3]);
            }
            bias_done = true;
          }
        }
        if (p.bias_mode == 1 && !bias_done) {
          if (col0 + CHUNK_COLS <= p.N && ((reinterpret_cast<uintptr_t>(p.bias + col0) & 15u) == 0)) {
#pragma unroll
            for (int j = 0; j < CHUNK_COLS / 4; ++j) {
              const float4 bq = __ldg(reinterpret_cast<const float4*>(p.bias + col0) + j);
              v[4 * j] += bq.x; v[4 * j + 1] += bq.y; v[4 * j + 2] += bq.z; v[4 * j + 3] += bq.w;
            }
          } else {
#pragma unroll
            for (int j = 0; j < CHUNK_COLS; ++j)
              if (col0 + j < p.N) v[j] += __ldg(p.bias + col0 + j);
          }
        }
        if (p.act == 1) {
#pragma unroll
          for (int j = 0; j < CHUNK_COLS / 2; ++j) {
            const uint64_t g2 = gelu_erf2(pack_f32x2(v[2 * j], v[2 * j + 1]));
            unpack_f32x2(g2, v[2 * j], v[2 * j + 1]);
          }
        }
        G2T(320 + c);
        if constexpr (KIND == G2_F32) {
          if (has_res) {
            const uint32_t slot = use_idx % (uint32_t)RS;
            mbar_wait(&my_res_full[slot], (use_idx / (uint32_t)RS) & 1u);
            G2T(330 + c);
            const uint8_t* rs = my + slot * Cfg::SLOT_BYTES;
            if (p.res_split) {
              // value = hi + lo, two 16-bit rows of CHUNK_COLS * 2 bytes each
              auto add_pair = [&](auto f16tag) {
                constexpr bool F16 = decltype(f16tag)::value;
#pragma unroll
                for (int j = 0; j < CHUNK_COLS / 8; ++j) {
                  const uint32_t off = CHUNK_COLS == 32 ? sw64_offset(lane, j) : lane * 32 + j * 16;
                  const uint4 h = *reinterpret_cast<const uint4*>(rs + off);
                  const uint4 l = *reinterpret_cast<const uint4*>(rs + Cfg::SLOT_BYTES / 2 + off);
                  const uint32_t hw[4] = {h.x, h.y, h.z, h.w}, lw[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
                  for (int k = 0; k < 4; ++k) {
                    float h0, h1, l0, l1;
                    unpack16x2<F16>(hw[k], h0, h1);
                    unpack16x2<F16>(lw[k], l0, l1);
                    v[8 * j + 2 * k] += h0 + l0;
                    v[8 * j + 2 * k + 1] += h1 + l1;
                  }
                }
              };
              if (p.fp16) add_pair(std::true_type{});
              else add_pair(std::false_type{});
            } else {
#pragma unroll
              for (int j = 0; j < CHUNK_COLS / 4; ++j) {
                const float4 rq = *reinterpret_cast<const float4*>(
                    rs + (CHUNK_COLS == 32 ? sw128_offset(lane, j) : sw64_offset(lane, j)));
                v[4 * j] += rq.x; v[4 * j + 1] += rq.y; v[4 * j + 2] += rq.z; v[4 * j + 3] += rq.w;
              }
            }
          }
          if (p.row_stats_out != nullptr) {
            // fused LayerNorm, producer side: statistics of the final row values (columns past N are zero-weighted)
#pragma unroll
            for (int j = 0; j < CHUNK_COLS; ++j)
              if (col0 + j < p.N) {
                st_sum += v[j];
                st_sq = fmaf(v[j], v[j], st_sq);
              }
          }
          G2T(340 + c);
          // staging slots: OUT_SLOTS fp32 buffers, or — pair stream — twice as many half-sized lo buffers
          uint8_t* slot_out;
          uint8_t* slot_raw;
          if (p.out_split) {
            constexpr uint32_t NS = 2 * Cfg::OUT_SLOTS < Cfg::RAW_SLOTS ? 2 * Cfg::OUT_SLOTS : Cfg::RAW_SLOTS;
            slot_out = my_out + (use_idx % NS) * (Cfg::SLOT_BYTES / 2);
            slot_raw = my_raw + (use_idx % NS) * Cfg::RAW_SLOT_BYTES;
            if (elect_one()) bulk_wait_read<NS - 1>();
          } else {
            slot_out = my_out + (use_idx % (uint32_t)Cfg::OUT_SLOTS) * Cfg::SLOT_BYTES;
            slot_raw = my_raw + (use_idx % (uint32_t)Cfg::OUT_SLOTS) * Cfg::RAW_SLOT_BYTES;
            if (elect_one()) bulk_wait_read<Cfg::OUT_SLOTS - 1>();   // the stores issued OUT_SLOTS chunks ago used these slots
          }
          __syncwarp();
          G2T(350 + c);
          if (p.raw_bf16 != nullptr) {
            // 16-bit copy of the un-normalised row segment for the fused LayerNorm of the consumer: 32-byte rows (no
            // swizzle) with 16-column chunks, 64-byte rows (SWIZZLE_64B) with 32-column chunks
            auto raw_out = [&](auto f16tag) {
              constexpr bool F16 = decltype(f16tag)::value;
#pragma unroll
              for (int j = 0; j < CHUNK_COLS / 8; ++j) {
                const uint32_t off = CHUNK_COLS == 32 ? sw64_offset(lane, j) : lane * 32 + j * 16;
                uint32_t hw[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) hw[k] = pack16x2<F16>(v[8 * j + 2 * k], v[8 * j + 2 * k + 1]);
                *reinterpret_cast<uint4*>(slot_raw + off) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
                if (p.out_split) {
                  // the remainder after the 16-bit rounding, in the (otherwise unused) fp32 staging slot
                  uint32_t lw[4];
#pragma unroll
                  for (int k = 0; k < 4; ++k) {
                    float h0, h1;
                    unpack16x2<F16>(hw[k], h0, h1);
                    lw[k] = pack16x2<F16>(v[8 * j + 2 * k] - h0, v[8 * j + 2 * k + 1] - h1);
                  }
                  *reinterpret_cast<uint4*>(slot_out + off) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
                }
              }
            };
            if (p.fp16) raw_out(std::true_type{});
            else raw_out(std::false_type{});
          }
          if (!p.out_split) {
#pragma unroll
            for (int j = 0; j < CHUNK_COLS / 4; ++j)
              *reinterpret_cast<float4*>(slot_out + (CHUNK_COLS == 32 ? sw128_offset(lane, j) : sw64_offset(lane, j))) =
                  make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          }
          G2T(360 + c);
          fence_proxy_async_smem();
          __syncwarp();   // all lanes have written the staging slot and finished reading the residual slot
          G2T(370 + c);
          if (elect_one()) {
            tma_store_3d_hint(&tmap_out, slot_out, col0, row0, z, p.out_split ? p.hint_out_lo : p.hint_out);
            if (p.raw_bf16 != nullptr) tma_store_3d_hint(&tmap_raw, slot_raw, col0, row0, z, p.hint_out_hi);
            bulk_commit();
          }
          __syncwarp();
          if (has_res) issue_res();   // refills the residual slot that was just consumed
          G2T(380 + c);
          ++use_idx;
        } else {
          uint8_t* slot_out = my_out + (use_idx & 1u) * Cfg::SLOT_BYTES;
          if (elect_one()) bulk_wait_read<1>();   // the store issued two chunks ago used this slot
          __syncwarp();
          G2T(350 + c);
          // one uniform branch per 64-value chunk picks the conversion (no per-element select in the issue stream)
          auto stage_out = [&](auto f16tag) {
            constexpr bool F16 = decltype(f16tag)::value;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              uint4 q;
              q.x = pack16x2<F16>(v[8 * j], v[8 * j + 1]);
              q.y = pack16x2<F16>(v[8 * j + 2], v[8 * j + 3]);
              q.z = pack16x2<F16>(v[8 * j + 4], v[8 * j + 5]);
              q.w = pack16x2<F16>(v[8 * j + 6], v[8 * j + 7]);
              *reinterpret_cast<uint4*>(slot_out + sw128_offset(lane, j)) = q;
            }
          };
          if (p.fp16) stage_out(std::true_type{});
          else stage_out(std::false_type{});
          fence_proxy_async_smem();
          __syncwarp();
          G2T(360 + c);
          if (elect_one()) {
            tma_store_3d_hint(&tmap_out, slot_out, col0, row0, z, p.hint_out);
            bulk_commit();
          }
          __syncwarp();
          G2T(380 + c);
          ++use_idx;
        }
      }
      if (KIND == G2_F32 && p.row_stats_out != nullptr && row < p.M) {
        // one slot per (row, 128-column half-tile): plain stores, no atomics, nothing to zero beforehand
        float2* slots = reinterpret_cast<float2*>(p.row_stats_out) + (long long)row * p.stats_parts;
        slots[nt * 2 + half] = make_float2(st_sum, st_sq);
        if (nt == 0 && half == 0)     // slots beyond this tile width's count (the caller sizes the buffer for 64-column tiles)
          for (int s = 2 * p.tiles_n; s < p.stats_parts; ++s) slots[s] = make_float2(0.f, 0.f);
      }
    }
    if (elect_one()) bulk_wait_read<0>();   // staging slots must outlive the stores that read them (same thread as the commits)
  }

  __syncwarp();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // neither CTA exits (or frees TMEM) while its peer may still read its smem / signal it
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_2cta(tmem_base, Cfg::TMEM_COLS);
  }
}


// ------------------------------------------------------------------------------------------------
// pio_gemm2_stream_kernel: the producer GEMMs of a large fused tower (out-projection, fc2) on the 16-bit (hi, lo)
// residual stream — the residual comes in as two 16-bit arrays, the sum leaves as two 16-bit arrays (in place if the
// caller passes the same arrays: every element is read and rewritten by the same epilogue warp) plus the per-row
// LayerNorm statistics.  Same mainloop as pio_gemm2_kernel; the epilogue differs in three ways:
//   * the residual slot (hi | lo, 2 + 2 KB) is overwritten in place by the output (hi | lo): every lane reads and
//     writes only its own row, so there is no hazard and no separate staging buffer.  Two slots per warp give a
//     residual prefetch one chunk ahead *and* a store in flight behind, in 8 KB per warp instead of 12 — which buys the
//     fifth operand stage;
//   * the bias row is requested before the accumulator load, the statistics run in independent chains;
//   * EW_ = 8 or 16 epilogue warps (32 rows x 128 or 64 columns of the tile each).  A clock64 trace of the 8-warp general
//     kernel showed ~16 400 clk of dependent epilogue chain per warp and tile against 8 200 clk of tensor time, which
//     suggested sixteen warps; with them the issuer never waits for an accumulator (tools/trace_gemm2.py, tags 500-502)
//     and the kernel is no faster: what the launch costs is set by the bytes it moves, not by who waits for whom
//     ("additive").
// Roles: warp 0 TMA producer, warp 1 MMA issuer (leader CTA), warp 2 TMEM allocator, warp 3 idle, warps 4.. epilogue
// (warp w: TMEM lanes 32*(w%4).., column slice (w-4)/4 of the 256-column tile).
// ------------------------------------------------------------------------------------------------
template <int EW_, int SLOTS_, int STAGES_>
struct Gemm2StreamCfg {
  static constexpr int BM = 128, BN = 256, BK = 64;
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = (BN / 2) * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = STAGES_;
  static constexpr int EPI_WARPS = EW_;            // 8: 128-column slices, 16: 64-column slices
  static constexpr int PARTS = EPI_WARPS / 4;
  static constexpr int PART_COLS = 256 / PARTS;
  static constexpr int THREADS = (4 + EPI_WARPS) * 32;
  static constexpr int CHUNK = 32;                       // columns per epilogue step
  static constexpr int HALF_BYTES = 32 * CHUNK * 2;      // one 16-bit box: 32 rows x 64 bytes (SWIZZLE_64B)
  static constexpr int SLOT_BYTES = 2 * HALF_BYTES;      // hi | lo
  static constexpr int SLOTS = SLOTS_;
  static constexpr int EPI_BYTES = EPI_WARPS * SLOTS * SLOT_BYTES;
  static constexpr int BAR_BYTES = 1024;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_BYTES + BAR_BYTES;
  static constexpr int TMEM_COLS = 512;
  static_assert(SMEM_BYTES <= 232448, "shared memory budget exceeded");
};

template <int EW_, int SLOTS_, int STAGES_>
__global__ void __launch_bounds__((Gemm2StreamCfg<EW_, SLOTS_, STAGES_>::THREADS), 1)
pio_gemm2_stream_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                        const __grid_constant__ CUtensorMap tmap_out_lo, const __grid_constant__ CUtensorMap tmap_out_hi,
                        const __grid_constant__ CUtensorMap tmap_res_hi, const __grid_constant__ CUtensorMap tmap_res_lo,
                        const Gemm2Params p) {
  using Cfg = Gemm2StreamCfg<EW_, SLOTS_, STAGES_>;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* epi_base = smem + Cfg::STAGES * Cfg::STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(epi_base + Cfg::EPI_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = full_bar + Cfg::STAGES;
  uint64_t* tmem_full = empty_bar + Cfg::STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint64_t* res_full = tmem_empty + 2;              // [EPI_WARPS][SLOTS]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_full + Cfg::EPI_WARPS * Cfg::SLOTS);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
#ifdef PIO_GEMM2_TRACE
  int g2n = 0;
#endif
  const uint32_t crank = cluster_ctarank();
  const int pair_id = blockIdx.x >> 1;
  const int num_pairs = gridDim.x >> 1;
  const int total_tiles = p.m_pairs * p.tiles_n;
  const int num_k_chunks = (p.K + Cfg::BK - 1) / Cfg::BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    tma_prefetch_desc(&tmap_out_lo);
    tma_prefetch_desc(&tmap_out_hi);
    tma_prefetch_desc(&tmap_res_hi);
    tma_prefetch_desc(&tmap_res_lo);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < Cfg::STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 2 * Cfg::EPI_WARPS);
    }
    for (int i = 0; i < Cfg::EPI_WARPS * Cfg::SLOTS; ++i) mbar_init(&res_full[i], 1);
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc_2cta(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish_2cta();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_sync();

  if (warp == 0) {
    // ================= TMA producer (both CTAs) =================
    int stage = 0;
    uint32_t phase = 0;
    for (int seq = pair_id; seq < total_tiles; seq += num_pairs) {
      const int t = p.reverse ? total_tiles - 1 - seq : seq;
      const int nt = t % p.tiles_n;
      const int mp = t / p.tiles_n;
      const int m0 = mp * 256 + (int)crank * Cfg::BM;
      const int n0 = nt * Cfg::BN + (int)crank * (Cfg::BN / 2);
      for (int kc = 0; kc < num_k_chunks; ++kc) {
        mbar_wait(&empty_bar[stage], phase ^ 1u);
        if (elect_one()) {
          uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
          uint8_t* sb = sa + Cfg::A_BYTES;
          if (crank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * Cfg::STAGE_BYTES);
          const uint32_t leader_bar = mapa_u32(&full_bar[stage], 0);
          tma_load_3d_2cta(sa, &tmap_a, leader_bar, kc * Cfg::BK, m0, 0, p.hint_a);
          tma_load_3d_2cta(sb, &tmap_b, leader_bar, kc * Cfg::BK, n0, 0, p.hint_b);
        }
        if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    if (crank == 0) {
      // ================= MMA issuer (leader CTA) =================
      const uint32_t idesc = make_idesc_f16(256, Cfg::BN, idesc_fmt(p.fp16), 0, 0);
      const uint64_t d0 = make_smem_desc_sw128(smem_u32(smem), 16, 1024);
      const uint32_t d_lo = (uint32_t)d0, d_hi = (uint32_t)(d0 >> 32);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int t = pair_id; t < total_tiles; t += num_pairs, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1u;
        G2TM(500);
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1u);
        G2TM(501);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * Cfg::BN;
        for (int kc = 0; kc < num_k_chunks; ++kc) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_lo = d_lo + (uint32_t)((stage * Cfg::STAGE_BYTES) >> 4);
          const uint32_t b_lo = a_lo + (uint32_t)(Cfg::A_BYTES >> 4);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            if (elect_one())
              umma_ss_2cta_lh(d_tmem, a_lo + ks * 2, d_hi, b_lo + ks * 2, d_hi, idesc, (kc | ks) != 0 ? 1u : 0u);
          }
          if (elect_one()) umma_commit_2cta_mcast(&empty_bar[stage], 0x3);
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1u; }
        }
        if (elect_one()) umma_commit_2cta_mcast(&tmem_full[acc], 0x3);
        G2TM(502);
      }
    }
  } else if (warp >= 4) {
    // ================= Epilogue (both CTAs) =================
    const int ew = warp - 4;
    const int quarter = warp & 3;
    const int part = ew >> 2;                          // column slice of the tile
    constexpr int CHUNKS = Cfg::PART_COLS / Cfg::CHUNK;
    uint8_t* my = epi_base + ew * (Cfg::SLOTS * Cfg::SLOT_BYTES);
    uint64_t* my_res_full = res_full + ew * Cfg::SLOTS;
    const uint32_t tmem_empty_leader0 = mapa_u32(&tmem_empty[0], 0);
    const uint32_t tmem_empty_leader1 = mapa_u32(&tmem_empty[1], 0);

    // residual prefetch cursor over this warp's (tile, chunk) stream
    int pf_t = pair_id, pf_c = 0, pf_row = 0, pf_col = 0;
    uint32_t pf_idx = 0;
    auto issue_res = [&]() {
      if (pf_t >= total_tiles) return;
      if (pf_c == 0) {
        const int t = p.reverse ? total_tiles - 1 - pf_t : pf_t;
        const int nt = t % p.tiles_n;
        const int mp = t / p.tiles_n;
        pf_row = mp * 256 + (int)crank * Cfg::BM + quarter * 32;
        pf_col = nt * Cfg::BN + part * Cfg::PART_COLS;
      }
      const int col = pf_col + pf_c * Cfg::CHUNK;
      const uint32_t slot = pf_idx % (uint32_t)Cfg::SLOTS;
      if (elect_one()) {
        // the store that last used this slot (two chunks ago) must have finished reading it
        bulk_wait_read<0>();
        mbar_arrive_expect_tx(&my_res_full[slot], Cfg::SLOT_BYTES);
        tma_load_3d(my + slot * Cfg::SLOT_BYTES, &tmap_res_hi, &my_res_full[slot], col, pf_row, 0, p.hint_res_hi);
        tma_load_3d(my + slot * Cfg::SLOT_BYTES + Cfg::HALF_BYTES, &tmap_res_lo, &my_res_full[slot], col, pf_row, 0, p.hint_res_lo);
      }
      __syncwarp();
      ++pf_idx;
      if (++pf_c == CHUNKS) { pf_c = 0; pf_t += num_pairs; }
    };
    issue_res();   // chunk 0; chunk g + 1 is requested at the start of chunk g

    auto run = [&](auto f16tag) {
      constexpr bool F16 = decltype(f16tag)::value;
      uint32_t use_idx = 0;
      int it = 0;
      for (int seq = pair_id; seq < total_tiles; seq += num_pairs, ++it) {
        const int t = p.reverse ? total_tiles - 1 - seq : seq;
        const int nt = t % p.tiles_n;
        const int mp = t / p.tiles_n;
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1u;
        const int row0 = mp * 256 + (int)crank * Cfg::BM + quarter * 32;
        const int row = row0 + lane;
        float s1[2] = {0.f, 0.f}, s2[2] = {0.f, 0.f};
        const uint32_t t_row = tmem_base + acc * Cfg::BN + part * Cfg::PART_COLS + (static_cast<uint32_t>(quarter * 32) << 16);
#pragma unroll 1
        for (int c = 0; c < CHUNKS; ++c) {
          const int col0 = nt * Cfg::BN + part * Cfg::PART_COLS + c * Cfg::CHUNK;
          const bool full = col0 + Cfg::CHUNK <= p.N;
          // bias row segment: requested first, so that its latency hides behind the accumulator load
          float4 bq[Cfg::CHUNK / 4];
          if (p.bias_mode == 1 && full) {
#pragma unroll
            for (int j = 0; j < Cfg::CHUNK / 4; ++j) bq[j] = __ldg(reinterpret_cast<const float4*>(p.bias + col0) + j);
          } else {
#pragma unroll
            for (int j = 0; j < Cfg::CHUNK / 4; ++j) {
              bq[j].x = (p.bias_mode == 1 && col0 + 4 * j + 0 < p.N) ? __ldg(p.bias + col0 + 4 * j + 0) : 0.f;
              bq[j].y = (p.bias_mode == 1 && col0 + 4 * j + 1 < p.N) ? __ldg(p.bias + col0 + 4 * j + 1) : 0.f;
              bq[j].z = (p.bias_mode == 1 && col0 + 4 * j + 2 < p.N) ? __ldg(p.bias + col0 + 4 * j + 2) : 0.f;
              bq[j].w = (p.bias_mode == 1 && col0 + 4 * j + 3 < p.N) ? __ldg(p.bias + col0 + 4 * j + 3) : 0.f;
            }
          }
          if (c == 0) {
            G2T(100);
            mbar_wait(&tmem_full[acc], acc_phase);
            G2T(101);
            tc_fence_after();
          }
          uint32_t r[Cfg::CHUNK];
          tmem_ld32(t_row + c * Cfg::CHUNK, r);
          tmem_wait_ld();
          G2T(310 + c);
          if (c == CHUNKS - 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(acc ? tmem_empty_leader1 : tmem_empty_leader0);
          }
          if (Cfg::SLOTS > 1) issue_res();   // next chunk's residual into the other slot (its last store was committed a chunk ago)
          G2T(320 + c);
          float v[Cfg::CHUNK];
#pragma unroll
          for (int j = 0; j < Cfg::CHUNK / 4; ++j) {
            v[4 * j + 0] = fmaf(__uint_as_float(r[4 * j + 0]), p.alpha, bq[j].x);
            v[4 * j + 1] = fmaf(__uint_as_float(r[4 * j + 1]), p.alpha, bq[j].y);
            v[4 * j + 2] = fmaf(__uint_as_float(r[4 * j + 2]), p.alpha, bq[j].z);
            v[4 * j + 3] = fmaf(__uint_as_float(r[4 * j + 3]), p.alpha, bq[j].w);
          }
          const uint32_t slot = use_idx % (uint32_t)Cfg::SLOTS;
          uint8_t* rs = my + slot * Cfg::SLOT_BYTES;
          mbar_wait(&my_res_full[slot], (use_idx / (uint32_t)Cfg::SLOTS) & 1u);
          G2T(330 + c);
#pragma unroll
          for (int j = 0; j < Cfg::CHUNK / 8; ++j) {
            const uint32_t off = sw64_offset(lane, j);
            const uint4 h = *reinterpret_cast<const uint4*>(rs + off);
            const uint4 l = *reinterpret_cast<const uint4*>(rs + Cfg::HALF_BYTES + off);
            const uint32_t hw[4] = {h.x, h.y, h.z, h.w}, lw[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              float h0, h1, l0, l1;
              unpack16x2<F16>(hw[k], h0, h1);
              unpack16x2<F16>(lw[k], l0, l1);
              v[8 * j + 2 * k] += h0 + l0;
              v[8 * j + 2 * k + 1] += h1 + l1;
            }
          }
          if (p.row_stats_out != nullptr) {
            if (full) {
#pragma unroll
              for (int j = 0; j < Cfg::CHUNK; j += 2) {
                s1[0] += v[j];
                s1[1] += v[j + 1];
                s2[0] = fmaf(v[j], v[j], s2[0]);
                s2[1] = fmaf(v[j + 1], v[j + 1], s2[1]);
              }
            } else {
#pragma unroll
              for (int j = 0; j < Cfg::CHUNK; ++j)
                if (col0 + j < p.N) {
                  s1[0] += v[j];
                  s2[0] = fmaf(v[j], v[j], s2[0]);
                }
            }
          }
          G2T(340 + c);
          // hi = round16(v) (the raw copy the next projection reads), lo = round16(v - hi): written over the residual
          // this lane has just read (its own row only)
#pragma unroll
          for (int j = 0; j < Cfg::CHUNK / 8; ++j) {
            const uint32_t off = sw64_offset(lane, j);
            uint32_t hw[4], lw[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              hw[k] = pack16x2<F16>(v[8 * j + 2 * k], v[8 * j + 2 * k + 1]);
              float h0, h1;
              unpack16x2<F16>(hw[k], h0, h1);
              lw[k] = pack16x2<F16>(v[8 * j + 2 * k] - h0, v[8 * j + 2 * k + 1] - h1);
            }
            *reinterpret_cast<uint4*>(rs + off) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
            *reinterpret_cast<uint4*>(rs + Cfg::HALF_BYTES + off) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
          }
          G2T(360 + c);
          fence_proxy_async_smem();
          __syncwarp();
          if (elect_one()) {
            tma_store_3d_hint(&tmap_out_hi, rs, col0, row0, 0, p.hint_out_hi);
            tma_store_3d_hint(&tmap_out_lo, rs + Cfg::HALF_BYTES, col0, row0, 0, p.hint_out_lo);
            bulk_commit();
          }
          __syncwarp();
          if (Cfg::SLOTS == 1) issue_res();  // single slot: the next residual follows the store through the same buffer
          G2T(380 + c);
          ++use_idx;
        }
        if (p.row_stats_out != nullptr && row < p.M) {
          // one slot per (row, column slice): plain stores, nothing to zero beforehand
          float2* slots = reinterpret_cast<float2*>(p.row_stats_out) + (long long)row * p.stats_parts;
          slots[nt * Cfg::PARTS + part] = make_float2(s1[0] + s1[1], s2[0] + s2[1]);
          if (nt == 0 && part == 0)
            for (int s = Cfg::PARTS * p.tiles_n; s < p.stats_parts; ++s) slots[s] = make_float2(0.f, 0.f);
        }
      }
    };
    if (p.fp16) run(std::true_type{});
    else run(std::false_type{});
    if (elect_one()) bulk_wait_read<0>();
  }

  __syncwarp();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_2cta(tmem_base, Cfg::TMEM_COLS);
  }
}

template <int EW_, int SLOTS_, int STAGES_>
static int launch_gemm2_stream(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& to_lo, const CUtensorMap& to_hi,
                               const CUtensorMap& tr_hi, const CUtensorMap& tr_lo, const Gemm2Params& p, int pairs,
                               const DeviceInfo& dev, cudaStream_t stream) {
  using SCfg = Gemm2StreamCfg<EW_, SLOTS_, STAGES_>;
  static PerDeviceOnce once_s;
  const cudaError_t e2 = once_s.run(dev.device, [] {
    return cudaFuncSetAttribute(pio_gemm2_stream_kernel<EW_, SLOTS_, STAGES_>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                SCfg::SMEM_BYTES);
  });
  if (e2 != cudaSuccess) return fail(PIO_ERR_CUDA, "cudaFuncSetAttribute(gemm2 stream) failed: %s", cudaGetErrorString(e2));
  ProfileScope prof(KF_GEMM, 2.0 * p.M * p.N * (double)p.K, 8.0 * p.M * (double)p.N, stream);
  PIO_CUDA_OK(launch_kernel(pio_gemm2_stream_kernel<EW_, SLOTS_, STAGES_>, dim3((unsigned)(pairs * 2), 1, 1),
                            dim3(SCfg::THREADS, 1, 1), SCfg::SMEM_BYTES, stream, 2, ta, tb, to_lo, to_hi, tr_hi, tr_lo, p));
  g_launch_count.fetch_add(1);
  PIO_CUDA_OK(cudaGetLastError());
  return PIO_OK;
}


template <int KIND>
static int launch_gemm2_kind(const pio_gemm_args* a, const DeviceInfo& dev, cudaStream_t stream) {
  using Cfg = Gemm2Cfg<KIND>;
  CUtensorMap ta, tb, to, tr, traw;
  const bool a_bcast = a->batch > 1 && a->strideA == 0;
  const bool b_bcast = a->batch > 1 && a->strideB == 0;
  const bool r_bcast = a->batch > 1 && a->strideR == 0;
  const uint64_t a_batch = a_bcast ? 1 : a->batch, b_batch = b_bcast ? 1 : a->batch;
  {
    const uint64_t dims[3] = {(uint64_t)a->K, (uint64_t)a->M, a_batch};
    const uint64_t strides[2] = {(uint64_t)a->lda * 2, (uint64_t)(a_batch > 1 ? a->strideA : a->lda * (int64_t)a->M) * 2};
    const uint32_t box[3] = {64, 128, 1};
    int rc = encode_tmap(&ta, a->A, false, 3, dims, strides, box);
    if (rc != PIO_OK) return rc;
  }
  {
    const uint64_t dims[3] = {(uint64_t)a->K, (uint64_t)a->N, b_batch};
    const uint64_t strides[2] = {(uint64_t)a->ldb * 2, (uint64_t)(b_batch > 1 ? a->strideB : a->ldb * (int64_t)a->N) * 2};
    const uint32_t box[3] = {64, 128, 1};
    int rc = encode_tmap(&tb, a->B, false, 3, dims, strides, box);
    if (rc != PIO_OK) return rc;
  }
  CUtensorMap trlo;
  if (KIND == G2_F32) {
    const uint32_t box[3] = {G2_F32_CHUNK, 32, 1};
    constexpr int F32_SWZ = G2_F32_CHUNK == 32 ? 128 : 64;
    constexpr int B16_SWZ = G2_F32_CHUNK == 32 ? 64 : 0;
    int rc;
    if (a->out_lo16) {
      // split output: the "out" map addresses the lo array (16-bit, the geometry of the raw copy)
      const uint64_t wdims[3] = {(uint64_t)a->N, (uint64_t)a->M, 1};
      const uint64_t wstrides[2] = {(uint64_t)a->ldo16 * 2, (uint64_t)a->ldo16 * (uint64_t)a->M * 2};
      rc = encode_tmap(&to, a->out_lo16, false, 3, wdims, wstrides, box, B16_SWZ);
    } else {
      const uint64_t dims[3] = {(uint64_t)a->N, (uint64_t)a->M, (uint64_t)a->batch};
      const uint64_t strides[2] = {(uint64_t)a->ldo32 * 4, (uint64_t)(a->batch > 1 ? a->strideO32 : a->ldo32 * (int64_t)a->M) * 4};
      rc = encode_tmap(&to, a->out_f32, true, 3, dims, strides, box, F32_SWZ);
    }
    if (rc != PIO_OK) return rc;
    tr = to;
    traw = to;
    trlo = to;
    if (a->residual_hi16) {
      const uint64_t rdims[3] = {(uint64_t)a->N, (uint64_t)a->M, 1};
      const uint64_t rstrides[2] = {(uint64_t)a->ldr16 * 2, (uint64_t)a->ldr16 * (uint64_t)a->M * 2};
      rc = encode_tmap(&tr, a->residual_hi16, false, 3, rdims, rstrides, box, B16_SWZ);
      if (rc != PIO_OK) return rc;
      rc = encode_tmap(&trlo, a->residual_lo16, false, 3, rdims, rstrides, box, B16_SWZ);
      if (rc != PIO_OK) return rc;
    }
    if (a->out_bf16) {
      const uint64_t wdims[3] = {(uint64_t)a->N, (uint64_t)a->M, 1};
      const uint64_t wstrides[2] = {(uint64_t)a->ldo16 * 2, (uint64_t)a->ldo16 * (uint64_t)a->M * 2};
      const uint32_t wbox[3] = {G2_F32_CHUNK, 32, 1};
      rc = encode_tmap(&traw, a->out_bf16, false, 3, wdims, wstrides, wbox, G2_F32_CHUNK == 32 ? 64 : 0);
      if (rc != PIO_OK) return rc;
    }
    if (a->residual) {
      const uint64_t rb = r_bcast ? 1 : a->batch;
      const uint64_t rdims[3] = {(uint64_t)a->N, (uint64_t)a->M, rb};
      const uint64_t rstrides[2] = {(uint64_t)a->ldr * 4, (uint64_t)(rb > 1 ? a->strideR : a->ldr * (int64_t)a->M) * 4};
      rc = encode_tmap(&tr, a->residual, true, 3, rdims, rstrides, box, F32_SWZ);
      if (rc != PIO_OK) return rc;
    }
  } else {
    const uint64_t dims[3] = {(uint64_t)a->N, (uint64_t)a->M, (uint64_t)a->batch};
    const uint64_t strides[2] = {(uint64_t)a->ldo16 * 2, (uint64_t)(a->batch > 1 ? a->strideO16 : a->ldo16 * (int64_t)a->M) * 2};
    const uint32_t box[3] = {64, 32, 1};
    int rc = encode_tmap(&to, a->out_bf16, false, 3, dims, strides, box);
    if (rc != PIO_OK) return rc;
    tr = to;
    traw = to;
    trlo = to;
  }
  Gemm2Params p;
  p.M = a->M; p.N = a->N; p.K = a->K; p.batch = a->batch;
  p.fp16 = a->fp16 ? 1 : 0;
  p.tiles_n = (a->N + Cfg::BN - 1) / Cfg::BN;
  p.m_pairs = (a->M + 255) / 256;
  p.a_bcast = a_bcast; p.b_bcast = b_bcast; p.r_bcast = r_bcast;
  p.bias = a->bias; p.bias_mode = a->bias ? a->bias_mode : 0;
  p.act = a->act; p.alpha = a->alpha;
  p.has_residual = a->residual != nullptr || a->residual_hi16 != nullptr;
  p.out_split = (KIND == G2_F32 && a->out_lo16 != nullptr) ? 1 : 0;
  p.res_split = (KIND == G2_F32 && a->residual_hi16 != nullptr) ? 1 : 0;
  p.reverse = a->reverse_tiles ? 1 : 0;
  p.row_stats_out = (KIND == G2_F32) ? a->row_stats_out : nullptr;
  p.raw_bf16 = (KIND == G2_F32) ? reinterpret_cast<__nv_bfloat16*>(a->out_bf16) : nullptr;
  p.ld_raw = a->ldo16;
  p.row_stats_in = (KIND == G2_BF16) ? a->row_stats_in : nullptr;
  p.stats_parts = a->row_stats_parts > 0 ? a->row_stats_parts : 1;
  if (p.row_stats_out != nullptr && p.stats_parts < 2 * p.tiles_n)
    return fail(PIO_ERR_INVALID_ARGUMENT, "pio_gemm_bf16: row_stats_out needs row_stats_parts >= 2 * ceil(N / 256) = %d (got %d)",
                2 * p.tiles_n, a->row_stats_parts);
  p.ln_colsum = a->ln_colsum;
  p.ln_inv_c = a->ln_channels > 0 ? 1.0f / (float)a->ln_channels : 0.f;
  p.ln_eps = a->ln_eps;
  {
    // PIO_L2_HINTS: bit mask of the policies below (0: every access with the default priority)
    static const int use_hints = [] { const char* e = getenv("PIO_L2_HINTS"); return e ? atoi(e) : PIO_L2_HINTS_DEFAULT; }();
    p.hint_a = p.hint_b = p.hint_out = p.hint_res_hi = p.hint_res_lo = p.hint_out_hi = p.hint_out_lo = kEvictNormal;
    // only the towers whose stream does not fit L2 next to their other arrays need a policy at all
    const bool big = (double)a->M * a->N * 2.0 >= 32e6 || (double)a->M * a->K * 2.0 >= 32e6;
    if (use_hints && big && a->batch == 1) {
      if (use_hints & 1) p.hint_b = kEvictLast;                // weights: a few MB, read by every CTA pair
      if (p.row_stats_in != nullptr) {                         // fused-LayerNorm consumer: A is the hi half of the stream
        if (use_hints & 2) p.hint_a = kEvictLast;
        if (use_hints & 4) p.hint_out = kEvictFirst;           // QKV / MLP hidden: read once by the next kernel
      }
      if (p.out_split || p.res_split) {                        // producer on the (hi, lo) stream
        if (use_hints & 8) p.hint_a = kEvictFirst;             // attention output / MLP hidden: this is its only read
        if (use_hints & 2) p.hint_res_hi = p.hint_out_hi = kEvictLast;
        if (use_hints & 16) p.hint_res_lo = p.hint_out_lo = kEvictFirst;   // lo is read here and overwritten at once
      }
    }
  }
  static const int b_swap = [] { const char* e = getenv("PIO_GEMM2_BSWAP"); return (e && e[0] == '1') ? 1 : 0; }();
  p.b_swap = b_swap;

  static PerDeviceOnce once;
  const cudaError_t attr_err = once.run(dev.device, [] {
    return cudaFuncSetAttribute(pio_gemm2_kernel<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
  });
  if (attr_err != cudaSuccess)
    return fail(PIO_ERR_CUDA, "cudaFuncSetAttribute(gemm2<%d>) failed: %s", KIND, cudaGetErrorString(attr_err));

  const long long total = (long long)p.m_pairs * p.tiles_n * p.batch;
  int pairs = dev.sm_count / 2;
  if (a->max_ctas > 0 && a->max_ctas / 2 < pairs) pairs = a->max_ctas / 2 > 0 ? a->max_ctas / 2 : 1;
  if (total < pairs) pairs = (int)total;
  if constexpr (KIND == G2_F32 && G2_F32_CHUNK == 32) {
    // the (hi, lo) residual stream in and out: the 16-warp epilogue kernel (PIO_GEMM2_STREAM=0 keeps the general one)
    static const int use_stream = [] { const char* e = getenv("PIO_GEMM2_STREAM"); return (e && e[0] == '0') ? 0 : 1; }();
    if (use_stream && p.out_split && p.res_split && a->out_bf16 && a->batch == 1 && p.bias_mode <= 1 && p.act == 0 &&
        (p.row_stats_out == nullptr || p.stats_parts >= 4 * p.tiles_n)) {
      // 8 epilogue warps, two in-place slots each, five operand stages; PIO_G2S_CFG=1: sixteen warps / three stages
      // (measured slower: the operand ring is what the freed shared memory is worth)
      static const int cfg = [] { const char* e = getenv("PIO_G2S_CFG"); return e ? atoi(e) : 0; }();
      if (cfg == 1) return launch_gemm2_stream<16, 2, 3>(ta, tb, to, traw, tr, trlo, p, pairs, dev, stream);
      return launch_gemm2_stream<8, 2, 5>(ta, tb, to, traw, tr, trlo, p, pairs, dev, stream);
    }
  }
  {
    const double bytes = (KIND == G2_F32 ? 4.0 * (p.has_residual ? 2 : 1) : 2.0) * a->M * (double)a->N * a->batch;
    ProfileScope prof(KF_GEMM, 2.0 * a->M * a->N * (double)a->K * a->batch, bytes, stream);
    PIO_CUDA_OK(launch_kernel(pio_gemm2_kernel<KIND>, dim3((unsigned)(pairs * 2), 1, 1), dim3(384, 1, 1), Cfg::SMEM_BYTES,
                              stream, 2, ta, tb, to, tr, traw, trlo, p));
  }
  g_launch_count.fetch_add(1);
  PIO_CUDA_OK(cudaGetLastError());
  return PIO_OK;
}

}  // namespace pio
#ifdef PIO_GEMM2_TRACE
extern "C" int pio_debug_gemm2_trace(unsigned long long* out) {   // out: 1024 * 2 words; clears the device buffer
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out, pio::g_gemm2_trace, sizeof(pio::g_gemm2_trace));
  static unsigned long long zeros[1024 * 2];
  cudaMemcpyToSymbol(pio::g_gemm2_trace, zeros, sizeof(zeros));
  return 0;
}
#endif
namespace pio {

// Whether the CTA-pair kernel can run this problem (layout / alignment rules of its TMA epilogue).
bool gemm2_eligible(const pio_gemm_args* a) {
  if (a->b_mn_major) return false;
  const bool split_out = a->out_lo16 != nullptr, split_res = a->residual_hi16 != nullptr || a->residual_lo16 != nullptr;
  if (split_out || split_res) {
    // 16-bit pair residual stream: batch 1, 16-byte aligned rows, never mixed with the fp32 form of the same operand
    if (a->batch != 1) return false;
    if (split_out && (a->out_f32 || !a->out_bf16 || !aligned16(a->out_lo16))) return false;
    if (split_res && (a->residual || !a->residual_hi16 || !a->residual_lo16 || !aligned16(a->residual_hi16) ||
                      !aligned16(a->residual_lo16) || a->ldr16 % 8 != 0 || a->ldr16 < a->N))
      return false;
    if (split_res && !split_out && !a->out_f32) return false;
  }
  const bool f32 = a->out_f32 != nullptr || split_out, b16 = a->out_bf16 != nullptr;
  if (!f32 && !b16) return false;
  if (f32 && b16) {
    // fp32 output plus its raw bf16 copy (fused-LayerNorm producer): staged per 16-column chunk, TMA-stored
    if (a->batch != 1 || a->ldo16 % 8 != 0 || a->ldo16 < a->N || !aligned16(a->out_bf16)) return false;
  }
  if (!f32 && a->residual) return false;
  if ((a->row_stats_out || a->row_stats_in) && a->batch != 1) return false;
  if (a->row_stats_out && !f32) return false;
  if (a->row_stats_in && (f32 || !a->ln_colsum || a->ln_channels <= 0)) return false;
  if (f32) {
    if (!split_out) {
      if (!aligned16(a->out_f32) || a->ldo32 % 4 != 0 || a->ldo32 < a->N) return false;
      if (a->batch > 1 && (a->strideO32 % 4 != 0 || a->strideO32 <= 0)) return false;
    }
    if (a->residual) {
      if (!aligned16(a->residual) || a->ldr % 4 != 0 || a->ldr < a->N) return false;
      if (a->batch > 1 && a->strideR % 4 != 0) return false;
    }
  }
  if (!f32) {
    if (!aligned16(a->out_bf16) || a->ldo16 % 8 != 0 || a->ldo16 < a->N) return false;
    if (a->batch > 1 && (a->strideO16 % 8 != 0 || a->strideO16 <= 0)) return false;
  }
  return true;
}

int launch_gemm2(const pio_gemm_args* a, const DeviceInfo& dev, cudaStream_t stream) {
  if (a->out_f32 || a->out_lo16) return launch_gemm2_kind<G2_F32>(a, dev, stream);
  return launch_gemm2_kind<G2_BF16>(a, dev, stream);
}

}  // namespace pio
