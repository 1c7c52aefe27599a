// Streaming attention with the QUERY TILE IN TENSOR MEMORY, for the folded single-head encoder cross-attends whose head
// size leaves room for it (K == V == LayerNorm(x), 192 < d <= 272: the ImageNet-pixels recipe, d = 261).
//
// Why: S = Q.K^T issued from shared memory (pio_flash.cu) re-reads the 128 x 16 query slice for every MMA — 4 KB of
// operand fetch against 1.5 .. 2 KB of keys, i.e. ~96 clk per MMA for a 32 clk tensor-pipe floor (DESIGN.md section 4.0:
// operands are fetched from shared memory at ~64 B/clk, operands in TMEM are free).  Here the softmax warps copy the query
// tile shared memory -> TMEM once per CTA and S is issued with the A operand in TMEM (as P.V already is), so both
// products only fetch the key tile.  TMEM: Q (dqk/2 columns, 16-bit pairs) + O (d fp32 columns) + two S buffers of BN
// keys; BN = 48 for d = 261 (144 + 272 + 96 = 512 columns).
//
// One CTA = one 128-query tile of one batch entry [and one key split].  Roles (128 + 32 * (4 + 4 * NW) threads): warp 0
// TMA producer, warp 1 MMA issuer, warp 2 TMEM allocator, warp 3 idle, then NW = BN / 16 softmax warps per lane quarter,
// each owning 16 of the tile's key columns (one tcgen05.ld.x16, eight packed P columns back).
#include <math.h>
#include <stdlib.h>

#include <type_traits>

#include "pio_common.cuh"
#include "pio_host.h"

namespace pio {

// Developer aid (compiled out unless -DPIO_FLASHQT_TRACE): CTA (0,0,0) records (tag, clock64) pairs; the first launch
// prints them (tools/trace_flash_qt.py).  Tags: 1000 + 10 j + e MMA issuer, 2000 + 10 j + e softmax warp 4.
#ifdef PIO_FLASHQT_TRACE
__device__ unsigned long long g_fqt_trace[2 * 2 * 1024];
#define QT_T(slot, tag)                                                                                 \
  do {                                                                                                  \
    if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && lane == 0 && qtn < 1024) {             \
      g_fqt_trace[((slot) * 1024 + qtn) * 2] = (unsigned long long)(tag);                               \
      g_fqt_trace[((slot) * 1024 + qtn) * 2 + 1] = (unsigned long long)clock64();                       \
      ++qtn;                                                                                            \
    }                                                                                                   \
  } while (0)
#else
#define QT_T(slot, tag)
#endif

struct FlashQtParams {
  int B, Nq, Nk, d;
  int fp16;
  int q_bcast;
  float scale_log2;
  const uint8_t* key_mask; long long stride_km;
  const uint8_t* row_keep; long long stride_rk;
  __nv_bfloat16* O; long long ldo, strideO;
  int num_splits, tiles_per_split, partial;
  float* O_part; float* m_part; float* l_part;
};

template <int NQC, int BN, int NW_>
struct FlashQtCfg {
  static constexpr int NW = NW_;                             // softmax warps per lane quarter, each owning BN / NW key columns
  static constexpr int CW = BN / NW;
  static constexpr int THREADS = 128 + 128 * NW;
  static constexpr int BAR_BYTES = 256 + 2 * NW * 128 * 4;   // mbarriers + two [NW][128] fp32 exchange slots
  static constexpr int Q_BYTES = NQC * 16384;
  static constexpr int CHUNK_BYTES = BN * 128;               // one 64-column chunk of a key tile
  static constexpr int STAGE_BYTES = NQC * CHUNK_BYTES;
  static constexpr int STAGES_RAW = (232448 - BAR_BYTES - Q_BYTES) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 4 ? 4 : STAGES_RAW;
  static constexpr int SMEM_BYTES = Q_BYTES + STAGES * STAGE_BYTES + BAR_BYTES;
  static_assert(BN % 16 == 0 && BN >= 32 && BN <= 64 && CW % 16 == 0, "key tile");
  static_assert(STAGES >= 2, "shared memory budget");
};

template <int NQC, int BN, int NW_>
__global__ void __launch_bounds__(FlashQtCfg<NQC, BN, NW_>::THREADS, 1)
pio_flash_qt_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                    const FlashQtParams p) {
  using Cfg = FlashQtCfg<NQC, BN, NW_>;
  constexpr int NW = Cfg::NW;
  constexpr int CW = Cfg::CW;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem;
  uint8_t* sKV = sQ + Cfg::Q_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sKV + Cfg::STAGES * Cfg::STAGE_BYTES);
  uint64_t* q_full = bars;                         // [1]  TMA -> softmax warps (they move Q to TMEM)
  uint64_t* qt_full = bars + 1;                    // [1]  Q is in TMEM (softmax warps -> MMA), 128 * NW arrivals
  uint64_t* kv_full = bars + 2;                    // [STAGES]
  uint64_t* kv_empty = kv_full + Cfg::STAGES;      // [STAGES]
  uint64_t* s_full = kv_empty + Cfg::STAGES;       // [2]
  uint64_t* p_full = s_full + 2;                   // [2]  128 * NW arrivals
  uint64_t* pv_done = p_full + 2;                  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + 2);
  float* xchg = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);   // [2 slots][NW][128 rows]

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
#ifdef PIO_FLASHQT_TRACE
  int qtn = 0;
#endif
  if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) {
    printf("pio_flash_qt_kernel: dynamic shared memory base is not 1024-byte aligned\n");
    __trap();
  }
  const int q0 = blockIdx.x * 128;
  const int b = blockIdx.y;
  const int split = blockIdx.z;
  const int total_tiles = (p.Nk + BN - 1) / BN;
  const int tile_begin = split * p.tiles_per_split;
  const int tile_end = min(total_tiles, tile_begin + p.tiles_per_split);
  const int ntiles = tile_end - tile_begin;  // host guarantees >= 1

  const int d_steps = (p.d + 15) / 16;            // 16-wide k-steps of Q.K^T == 16-column groups of O
  const int dv_n = d_steps * 16;
  const int q_cols = (d_steps * 8 + 15) & ~15;    // TMEM columns of the packed query tile, rounded to the store shape

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_k);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(q_full, 1);
    mbar_init(qt_full, 4 * NW * kArrivalsPerWarp);
    for (int s = 0; s < Cfg::STAGES; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 4 * NW * kArrivalsPerWarp);
      mbar_init(&pv_done[i], 1);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_o = tmem_base + 2 * BN;
  const uint32_t tmem_q = tmem_o + (uint32_t)dv_n;
  pdl_sync();

  if (warp == 0) {
    // ================= TMA producer =================
    const int bq = p.q_bcast ? 0 : b;
    if (elect_one()) {
      mbar_arrive_expect_tx(q_full, Cfg::Q_BYTES);
#pragma unroll
      for (int c = 0; c < NQC; ++c) tma_load_3d(sQ + c * 16384, &tmap_q, q_full, c * 64, q0, bq);
    }
    int stage = 0;
    uint32_t phase = 0;
    for (int j = 0; j < ntiles; ++j) {
      const int k0 = (tile_begin + j) * BN;
      mbar_wait(&kv_empty[stage], phase ^ 1u);
      if (elect_one()) {
        uint8_t* st = sKV + stage * Cfg::STAGE_BYTES;
        mbar_arrive_expect_tx(&kv_full[stage], Cfg::STAGE_BYTES);
#pragma unroll
        for (int c = 0; c < NQC; ++c) tma_load_3d(st + c * Cfg::CHUNK_BYTES, &tmap_k, &kv_full[stage], c * 64, k0, b);
      }
      if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1u; }
    }
  } else if (warp == 1) {
    // ================= MMA issuer: both products take their A operand from TMEM =================
    const uint32_t idesc_s = make_idesc_f16(128, BN, idesc_fmt(p.fp16), 0, 0);
    const uint64_t dk0 = make_smem_desc_sw128(smem_u32(sKV), 16, 1024);
    const uint64_t dv0 = make_smem_desc_sw128(smem_u32(sKV), Cfg::CHUNK_BYTES, 1024);
    const uint32_t k_lo = (uint32_t)dk0, k_hi = (uint32_t)(dk0 >> 32);
    const uint32_t v_lo = (uint32_t)dv0, v_hi = (uint32_t)(dv0 >> 32);
    const int n_rest = dv_n > 256 ? dv_n - 256 : 0;
    const uint32_t idesc_pv1 = make_idesc_f16(128, dv_n > 256 ? 256 : dv_n, idesc_fmt(p.fp16), 0, /*B MN-major*/ 1);
    const uint32_t idesc_pv2 = make_idesc_f16(128, n_rest > 0 ? n_rest : 16, idesc_fmt(p.fp16), 0, 1);
    auto issue_s = [&](int j, int stage) {
      const uint32_t dst = tmem_base + (uint32_t)((j & 1) * BN);
      const uint32_t b0 = k_lo + (uint32_t)((stage * Cfg::STAGE_BYTES) >> 4);
      // the issue rate of this thread bounds the kernel: the ImageNet-pixels width (261 -> 17 K
      // steps) gets a straight-line sequence without the per-step bound check (9 -> 6 uniform instructions per MMA)
      constexpr int FAST_STEPS = 4 * NQC - 3;
      if (d_steps == FAST_STEPS) {
#pragma unroll
        for (int ks = 0; ks < FAST_STEPS; ++ks) {
          const int c = ks >> 2, kk = ks & 3;
          if (elect_one())
            umma_ts_lh(dst, tmem_q + (uint32_t)(ks * 8), b0 + (uint32_t)((c * Cfg::CHUNK_BYTES + kk * 32) >> 4), k_hi, idesc_s,
                       ks != 0 ? 1u : 0u);
        }
      } else {
#pragma unroll
        for (int ks = 0; ks < 4 * NQC; ++ks) {
          if (ks < d_steps) {
            const int c = ks >> 2, kk = ks & 3;
            if (elect_one())
              umma_ts_lh(dst, tmem_q + (uint32_t)(ks * 8), b0 + (uint32_t)((c * Cfg::CHUNK_BYTES + kk * 32) >> 4), k_hi,
                         idesc_s, ks != 0 ? 1u : 0u);
          }
        }
      }
      if (elect_one()) umma_commit(&s_full[j & 1]);
    };
    auto issue_pv = [&](int j, int stage) {
      const uint32_t b0 = v_lo + (uint32_t)((stage * Cfg::STAGE_BYTES) >> 4);
      const uint32_t a0 = tmem_base + (uint32_t)((j & 1) * BN);   // P_j overlays the first BN/2 columns of S_j
#pragma unroll
      for (int ks = 0; ks < BN / 16; ++ks) {
        if (elect_one())
          umma_ts_lh(tmem_o, a0 + ks * 8, b0 + (uint32_t)((ks * 2048) >> 4), v_hi, idesc_pv1, (j | ks) != 0 ? 1u : 0u);
      }
      if (n_rest > 0) {
#pragma unroll
        for (int ks = 0; ks < BN / 16; ++ks) {
          if (elect_one())
            umma_ts_lh(tmem_o + 256, a0 + ks * 8, b0 + (uint32_t)((4 * Cfg::CHUNK_BYTES + ks * 2048) >> 4), v_hi, idesc_pv2,
                       (j | ks) != 0 ? 1u : 0u);
        }
      }
    };
    mbar_wait(qt_full, 0);
    int stage = 0;
    uint32_t phase = 0;
    mbar_wait(&kv_full[0], 0);
    tc_fence_after();
    issue_s(0, 0);
    for (int j = 0; j < ntiles; ++j) {
      int nstage = stage + 1;
      uint32_t nphase = phase;
      if (nstage == Cfg::STAGES) { nstage = 0; nphase ^= 1u; }
      if (j + 1 < ntiles) {
        mbar_wait(&kv_full[nstage], nphase);
        tc_fence_after();
        if (j < 40) QT_T(0, 1000 + 10 * j + 0);
        issue_s(j + 1, nstage);   // overwrites S_{j-1} / P_{j-1}: in order after PV_{j-1}, which consumed P_{j-1}
        if (j < 40) QT_T(0, 1000 + 10 * j + 1);
      }
      mbar_wait(&p_full[j & 1], (j >> 1) & 1);
      tc_fence_after();
      if (j < 40) QT_T(0, 1000 + 10 * j + 2);
      issue_pv(j, stage);
      if (elect_one()) {
        umma_commit(&kv_empty[stage]);
        umma_commit(&pv_done[j & 1]);
      }
      if (j < 40) QT_T(0, 1000 + 10 * j + 3);
      stage = nstage;
      phase = nphase;
    }
  } else if (warp >= 4) {
    // ================= softmax / correction / epilogue =================
    const int quarter = warp & 3;
    const int part = (warp - 4) >> 2;     // which CW key columns of a tile (0 .. NW-1)
    const int row = quarter * 32 + lane;  // row inside the tile == TMEM lane
    const int q = q0 + row;
    const uint32_t lane_off = static_cast<uint32_t>(quarter * 32) << 16;
    const uint8_t* km = p.key_mask ? p.key_mask + (long long)b * p.stride_km : nullptr;
    auto group_sync = [&]() {
      if constexpr (NW > 1) asm volatile("bar.sync %0, %1;" ::"r"(1 + quarter), "r"(32 * NW) : "memory");
    };

    // ---- the query tile moves shared memory -> TMEM once: this warp takes every NW-th 32-column group of the packed row
    //      (= a 64-element chunk of the swizzled tile: eight 16-byte pieces) ----
    mbar_wait(q_full, 0);
    for (int g = part; g * 32 < q_cols; g += NW) {
      uint32_t w[32];
#pragma unroll
      for (int c16 = 0; c16 < 8; ++c16) {
        const uint4 v = *reinterpret_cast<const uint4*>(sQ + g * 16384 + sw128_offset((uint32_t)row, (uint32_t)c16));
        w[4 * c16] = v.x; w[4 * c16 + 1] = v.y; w[4 * c16 + 2] = v.z; w[4 * c16 + 3] = v.w;
      }
      if ((g + 1) * 32 <= q_cols) {
        tmem_st32(tmem_q + lane_off + g * 32, w);
      } else {   // last group: q_cols is a multiple of 16
        uint32_t h16[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) h16[i] = w[i];
        tmem_st16(tmem_q + lane_off + g * 32, h16);
      }
    }
    tmem_wait_st();
    tc_fence_before();
    mbar_arrive_warp(qt_full);

    // O columns owned by this warp for the rescale and the epilogue (16-column groups, dealt round-robin)
    float m = -INFINITY;  // running max of scale_log2 * s (identical in all parts)
    float l = 0.f;        // running sum of exp2(t - m) over this part's columns
    for (int j = 0; j < ntiles; ++j) {
      const int k0 = (tile_begin + j) * BN + part * CW;
      const uint32_t t_s = tmem_base + (uint32_t)((j & 1) * BN + part * CW) + lane_off;
      mbar_wait(&s_full[j & 1], (j >> 1) & 1);
      tc_fence_after();
      if (warp == 4 && j < 40) QT_T(1, 2000 + 10 * j + 0);
      const bool tail = (k0 + CW > p.Nk) || (km != nullptr);
      uint32_t r[CW];
#pragma unroll
      for (int c = 0; c < CW / 16; ++c) tmem_ld16(t_s + c * 16, *reinterpret_cast<uint32_t(*)[16]>(&r[c * 16]));
      tmem_wait_ld();
      if (warp == 4 && j < 40) QT_T(1, 2000 + 10 * j + 1);
      if (tail) {
#pragma unroll
        for (int i = 0; i < CW; ++i) {
          const int k = k0 + i;
          const bool ok = (k < p.Nk) && (km == nullptr || km[k] != 0);
          if (!ok) r[i] = 0xff800000u;
        }
      }
      float tmax;
      {
        float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
        for (int i = 0; i < CW / 4; ++i) {
          mx0 = fmax3(mx0, __uint_as_float(r[4 * i]), __uint_as_float(r[4 * i + 1]));
          mx1 = fmax3(mx1, __uint_as_float(r[4 * i + 2]), __uint_as_float(r[4 * i + 3]));
        }
        tmax = fmaxf(mx0, mx1);
      }
      if constexpr (NW > 1) {
        float* slot = xchg + (j & 1) * (NW * 128);
        slot[part * 128 + row] = tmax;
        group_sync();
#pragma unroll
        for (int o = 0; o < NW; ++o) tmax = fmaxf(tmax, slot[o * 128 + row]);
      }
      tmax *= p.scale_log2;
      float m_use = m;
      const bool grow = tmax > m + 8.0f;
      float alpha = 1.0f;
      if (grow) {
        alpha = (m == -INFINITY) ? 0.0f : exp2f(m - tmax);
        m_use = tmax;
      }
      const bool any_grow = __any_sync(0xffffffffu, grow && j > 0 && m != -INFINITY);
      if (any_grow) {
        mbar_wait(&pv_done[(j - 1) & 1], ((j - 1) >> 1) & 1);
        tc_fence_after();
#pragma unroll 1
        for (int c = part * 16; c < dv_n; c += 16 * NW) {
          uint32_t o[16];
          tmem_ld16(tmem_o + lane_off + c, o);
          tmem_wait_ld();
#pragma unroll
          for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
          tmem_st16(tmem_o + lane_off + c, o);
        }
        tmem_wait_st();
      }
      l *= alpha;
      m = m_use;
      const float msub = (m == -INFINITY) ? 0.0f : m;
      const uint64_t sc2 = pack_f32x2(p.scale_log2, p.scale_log2);
      const uint64_t nm2 = pack_f32x2(-msub, -msub);
      uint64_t la = pack_f32x2(0.f, 0.f), lb = pack_f32x2(0.f, 0.f);
      uint32_t w[CW / 2];
      auto exp_block = [&](auto f16tag) {
        constexpr bool F16 = decltype(f16tag)::value;
#pragma unroll
        for (int i = 0; i < CW / 2; ++i) {
          const uint64_t t2 = ffma2(pack_f32x2(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1])), sc2, nm2);
          float t0, t1;
          unpack_f32x2(t2, t0, t1);
          const float e0 = ex2_approx(t0), e1 = ex2_approx(t1);
          w[i] = pack16x2<F16>(e0, e1);
          if (i & 1) lb = fadd2(lb, pack_f32x2(e0, e1));
          else la = fadd2(la, pack_f32x2(e0, e1));
        }
      };
      if (p.fp16) exp_block(std::true_type{});
      else exp_block(std::false_type{});
      // P overwrites S in place (two 16-bit values per 32-bit column).  Every warp of the group holds its S values in
      // registers (its own columns only when NW == 1; the group barrier above otherwise), so no unread S column is clobbered.
#pragma unroll
      for (int c = 0; c < CW / 16; ++c)
        tmem_st8(tmem_base + (uint32_t)((j & 1) * BN + part * (CW / 2) + c * 8) + lane_off,
                 *reinterpret_cast<uint32_t(*)[8]>(&w[c * 8]));
      {
        float a0, a1, b0, b1;
        unpack_f32x2(la, a0, a1);
        unpack_f32x2(lb, b0, b1);
        l += (a0 + a1) + (b0 + b1);
      }
      tmem_wait_st();
      tc_fence_before();
      if (warp == 4 && j < 40) QT_T(1, 2000 + 10 * j + 2);
      mbar_arrive_warp(&p_full[j & 1]);
    }
    // ---- epilogue: total row sum = all parts' partial sums ----
    if constexpr (NW > 1) {
      float* slot = xchg + (ntiles & 1) * (NW * 128);   // the slot the last tile did not use
      slot[part * 128 + row] = l;
      group_sync();
      l = 0.f;
#pragma unroll
      for (int o = 0; o < NW; ++o) l += slot[o * 128 + row];
    }
    mbar_wait(&pv_done[(ntiles - 1) & 1], ((ntiles - 1) >> 1) & 1);
    tc_fence_after();
    const bool keep = (q < p.Nq) && (p.row_keep == nullptr || p.row_keep[(long long)b * p.stride_rk + q] != 0);
    const bool emit_partial = p.partial || p.num_splits > 1;
    if (!emit_partial) {
      const float inv = (keep && l > 0.f) ? 1.0f / l : 0.0f;
      __nv_bfloat16* orow = p.O + (long long)b * p.strideO + (long long)q * p.ldo;
      for (int c = part * 16; c < dv_n; c += 16 * NW) {
        uint32_t r[16];
        tmem_ld16(tmem_o + lane_off + c, r);
        tmem_wait_ld();
        if (q < p.Nq) {
          __nv_bfloat16* op = orow + c;
          if (c + 16 <= p.d && ((reinterpret_cast<uintptr_t>(op) & 15u) == 0)) {
#pragma unroll
            for (int g = 0; g < 2; ++g) {
              uint4 v;
              v.x = pack16x2(__uint_as_float(r[8 * g]) * inv, __uint_as_float(r[8 * g + 1]) * inv, p.fp16);
              v.y = pack16x2(__uint_as_float(r[8 * g + 2]) * inv, __uint_as_float(r[8 * g + 3]) * inv, p.fp16);
              v.z = pack16x2(__uint_as_float(r[8 * g + 4]) * inv, __uint_as_float(r[8 * g + 5]) * inv, p.fp16);
              v.w = pack16x2(__uint_as_float(r[8 * g + 6]) * inv, __uint_as_float(r[8 * g + 7]) * inv, p.fp16);
              reinterpret_cast<uint4*>(op)[g] = v;
            }
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i)
              if (c + i < p.d) reinterpret_cast<uint16_t*>(op)[i] = cvt16(__uint_as_float(r[i]) * inv, p.fp16);
          }
        }
      }
    } else {
      const long long prow = ((long long)split * p.B + b) * p.Nq + q;
      if (q < p.Nq && part == 0) {
        p.m_part[prow] = (m == -INFINITY) ? -INFINITY : m * 0.69314718055994531f;  // back to natural-log units
        p.l_part[prow] = l;
      }
      float* orow = p.O_part + prow * p.d;
      for (int c = part * 16; c < dv_n; c += 16 * NW) {
        uint32_t r[16];
        tmem_ld16(tmem_o + lane_off + c, r);
        tmem_wait_ld();
        if (q < p.Nq) {
          if (c + 16 <= p.d && ((reinterpret_cast<uintptr_t>(orow + c) & 15u) == 0)) {
#pragma unroll
            for (int g = 0; g < 4; ++g)
              reinterpret_cast<float4*>(orow + c)[g] =
                  (l > 0.f) ? make_float4(__uint_as_float(r[4 * g]), __uint_as_float(r[4 * g + 1]), __uint_as_float(r[4 * g + 2]),
                                          __uint_as_float(r[4 * g + 3]))
                            : make_float4(0.f, 0.f, 0.f, 0.f);
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i)
              if (c + i < p.d) orow[c + i] = (l > 0.f) ? __uint_as_float(r[i]) : 0.0f;
          }
        }
      }
    }
  }

  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// Key-tile width for head size d with the query tile in TMEM, or 0 when it does not fit / is not worth it.
int flash_qt_key_tile(int d) {
  if (d <= 192 || d > 272) return 0;      // smaller heads fit the plain kernels' 128-key tiles; larger ones leave no room
  const int d_steps = (d + 15) / 16;
  const int q_cols = (d_steps * 8 + 15) & ~15;
  const int room = (512 - q_cols - d_steps * 16) / 2 / 16 * 16;
  return room >= 48 ? 48 : 0;
}

bool flash_qt_eligible(const pio_attention_args* a) {
  static const int off = [] { const char* e = getenv("PIO_FLASH_QT"); return (e && e[0] == '0') ? 1 : 0; }();
  if (off) return false;
  const bool same = (a->K == a->V) && (a->ldk == a->ldv) && (a->dqk == a->dv) && (a->strideK == a->strideV);
  return same && a->H == 1 && flash_qt_key_tile(a->dqk) == 48;
}

#ifndef PIO_FLASH_QT_NW
#define PIO_FLASH_QT_NW 1   // softmax warps per lane quarter: 1 = thread owns the whole 48-key row (no exchange), 3 = 16 keys each
#endif

int launch_flash_qt(const pio_attention_args* a, const DeviceInfo& dev, cudaStream_t stream) {
  constexpr int NQC = 5, BN = 48, NW = PIO_FLASH_QT_NW;
  using Cfg = FlashQtCfg<NQC, BN, NW>;
  CUtensorMap tq, tk;
  const int q_bcast = (a->strideQ == 0 && a->B > 1) ? 1 : 0;
  {
    const uint64_t dims[3] = {(uint64_t)a->dqk, (uint64_t)a->Nq, (uint64_t)(q_bcast ? 1 : a->B)};
    const uint64_t strides[2] = {(uint64_t)a->ldq * 2,
                                 (uint64_t)((q_bcast || a->B == 1) ? a->ldq * (int64_t)a->Nq : a->strideQ) * 2};
    const uint32_t box[3] = {64, 128, 1};
    int rc = encode_tmap_bf16(&tq, a->Q, 3, dims, strides, box);
    if (rc != PIO_OK) return rc;
  }
  {
    const uint64_t dims[3] = {(uint64_t)a->dqk, (uint64_t)a->Nk, (uint64_t)a->B};
    const uint64_t strides[2] = {(uint64_t)a->ldk * 2, (uint64_t)(a->B == 1 ? a->ldk * (int64_t)a->Nk : a->strideK) * 2};
    const uint32_t box[3] = {64, (uint32_t)BN, 1};
    int rc = encode_tmap_bf16(&tk, a->K, 3, dims, strides, box);
    if (rc != PIO_OK) return rc;
  }
  FlashQtParams p;
  p.B = a->B; p.Nq = a->Nq; p.Nk = a->Nk; p.d = a->dqk;
  p.fp16 = a->fp16 ? 1 : 0;
  p.q_bcast = q_bcast;
  p.scale_log2 = a->scale * 1.4426950408889634f;
  p.key_mask = a->key_mask; p.stride_km = a->stride_km;
  p.row_keep = a->row_keep; p.stride_rk = a->stride_rk;
  p.O = reinterpret_cast<__nv_bfloat16*>(a->O); p.ldo = a->ldo; p.strideO = a->strideO;
  const int total_tiles = (a->Nk + BN - 1) / BN;
  int splits = a->num_splits < 1 ? 1 : a->num_splits;
  if (splits > total_tiles) splits = total_tiles;
  p.tiles_per_split = (total_tiles + splits - 1) / splits;
  if (a->num_splits > 1 && (long long)(a->num_splits - 1) * p.tiles_per_split >= total_tiles)
    return fail(PIO_ERR_INVALID_ARGUMENT,
                "pio_attention_fwd: num_splits=%d leaves an empty split (Nk=%d, %d-key tiles: %d)", a->num_splits, a->Nk, BN,
                total_tiles);
  p.num_splits = a->num_splits < 1 ? 1 : a->num_splits;
  p.partial = a->partial;
  p.O_part = a->O_part; p.m_part = a->m_part; p.l_part = a->l_part;

  static PerDeviceOnce once;
  const cudaError_t attr_err = once.run(dev.device, [] {
    return cudaFuncSetAttribute(pio_flash_qt_kernel<NQC, BN, NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
  });
  if (attr_err != cudaSuccess)
    return fail(PIO_ERR_CUDA, "cudaFuncSetAttribute(flash_qt) failed: %s", cudaGetErrorString(attr_err));
  dim3 grid((a->Nq + 127) / 128, a->B, p.num_splits);
  {
    ProfileScope prof(KF_FLASH, 2.0 * a->B * (double)a->Nq * a->Nk * (a->dqk + a->dv), 0.0, stream);
    PIO_CUDA_OK(launch_kernel(pio_flash_qt_kernel<NQC, BN, NW>, grid, dim3(Cfg::THREADS, 1, 1), Cfg::SMEM_BYTES, stream, 1, tq, tk, p));
  }
#ifdef PIO_FLASHQT_TRACE
  {
    static bool dumped = false;
    if (!dumped) {
      dumped = true;
      cudaDeviceSynchronize();
      static unsigned long long host[2 * 2 * 1024];
      cudaMemcpyFromSymbol(host, g_fqt_trace, sizeof(host));
      unsigned long long t0 = ~0ull;
      for (int i = 0; i < 2 * 1024; ++i)
        if (host[2 * i] != 0 && host[2 * i + 1] < t0) t0 = host[2 * i + 1];
      for (int i = 0; i < 2 * 1024; ++i)
        if (host[2 * i] != 0) fprintf(stderr, "QT %d %llu %llu\n", i / 1024, host[2 * i], host[2 * i + 1] - t0);
    }
  }
#endif
  g_launch_count.fetch_add(1);
  PIO_CUDA_OK(cudaGetLastError());
  return PIO_OK;
}

}  // namespace pio
