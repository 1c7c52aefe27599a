// Streaming (flash-style) attention for sm_100a.
//
// One CTA = one 128-query tile of one (batch, head) [and one key split].  K/V tiles stream through a TMA ring
// (SWIZZLE_128B, 64-column chunks); S = Q.K^T is accumulated by tcgen05.mma into a double-buffered TMEM tile, the
// softmax warps read it with tcgen05.ld, keep running max / sum in registers (fp32, log2 domain), overwrite S in place
// with bf16 P (tcgen05.st), and O += P.V is accumulated in TMEM by a second tcgen05.mma whose A operand is P in TMEM
// and whose B operand is the V tile read MN-major — so when K and V are the same array (the folded single-head cross-attends) one TMA load
// feeds both products.  O is rescaled lazily (only when the running max grows by more than 2^8).
//
// Roles (384 threads): warp 0 TMA producer, warp 1 MMA issuer, warp 2 TMEM allocator, warp 3 idle,
// warps 4..11 softmax / correction / epilogue: thread = one query row (TMEM lane) x one half of the tile's key columns.
// The two warps that share a lane quarter exchange their partial row maxima through shared memory once per tile
// (a 64-thread named barrier), keep separate partial row sums, and split the O columns for the rescale and the
// epilogue.  With one warp per scheduler the softmax was latency-bound (ncu: the MMA issuer waiting on p_full); two
// warps per scheduler halve the per-tile softmax time.
#include <math.h>
#include <stdlib.h>

#include <type_traits>

#include "pio_common.cuh"
#include "pio_host.h"

namespace pio {

struct FlashParams {
  int B, H, Nq, Nk, dqk, dv;
  int fp16;               // 16-bit operand / output format: 0 = bf16, 1 = fp16
  int q_bcast;            // Q has a single batch entry shared by every b
  float scale_log2;       // scale * log2(e)
  const uint8_t* key_mask; long long stride_km;
  const uint8_t* row_keep; long long stride_rk;
  __nv_bfloat16* O; long long ldo, strideO;
  int num_splits, tiles_per_split, partial;
  float* O_part; float* m_part; float* l_part;
};

template <int NQC, int NVC, bool SAME>
struct FlashCfg {
  static constexpr int KV_CHUNKS = SAME ? NQC : (NQC + NVC);
  static constexpr int BAR_BYTES = 256 + 2048;     // mbarriers + the row-max / row-sum exchange buffer
  static constexpr int SMEM_LIMIT = 232448 - BAR_BYTES;  // 227 KB minus that block
  // BN = 128 needs Q + three K/V stages in shared memory and 2 x 128 + dv columns of TMEM (P lives in TMEM)
  static constexpr int SMEM128 = NQC * 16384 + 3 * KV_CHUNKS * 16384;
  static constexpr int TMEM128 = 2 * 128 + NVC * 64;
  static constexpr int BN = (SMEM128 <= SMEM_LIMIT && TMEM128 <= 512) ? 128 : 64;
  static constexpr int CHUNK_BYTES = BN * 128;            // one 64-column chunk of a K/V tile
  static constexpr int STAGE_BYTES = KV_CHUNKS * CHUNK_BYTES;
  static constexpr int Q_BYTES = NQC * 16384;
  static constexpr int AVAIL = SMEM_LIMIT - Q_BYTES;
  static constexpr int STAGES_RAW = AVAIL / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 4 ? 4 : STAGES_RAW;
  static constexpr int TMEM_NEED = 2 * BN + NVC * 64;
  static constexpr int TMEM_COLS = TMEM_NEED <= 128 ? 128 : (TMEM_NEED <= 256 ? 256 : 512);
  static constexpr int SMEM_USED = Q_BYTES + STAGES * STAGE_BYTES + BAR_BYTES;
  // request > half of the SM's shared memory so that exactly one CTA is resident per SM (TMEM is not oversubscribed)
  static constexpr int SMEM_BYTES = SMEM_USED < 120 * 1024 ? 120 * 1024 : SMEM_USED;
  static constexpr bool VALID = STAGES >= 2 && TMEM_NEED <= 512 && SMEM_USED <= 232448;
};

template <int NQC, int NVC, bool SAME>
__global__ void __launch_bounds__(384, 1)
pio_flash_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                 const __grid_constant__ CUtensorMap tmap_v, const FlashParams p) {
  using Cfg = FlashCfg<NQC, NVC, SAME>;
  constexpr int BN = Cfg::BN;
  extern __shared__ __align__(1024) uint8_t smem[];  // no static shared memory: the window starts 1024-aligned
  uint8_t* sQ = smem;
  uint8_t* sKV = sQ + Cfg::Q_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sKV + Cfg::STAGES * Cfg::STAGE_BYTES);
  uint64_t* q_full = bars;                         // [1]
  uint64_t* kv_full = bars + 1;                    // [STAGES]
  uint64_t* kv_empty = kv_full + Cfg::STAGES;      // [STAGES]
  uint64_t* s_full = kv_empty + Cfg::STAGES;       // [2]  S_j ready            (MMA -> softmax)
  uint64_t* p_full = s_full + 2;                   // [2]  P_j written          (softmax -> MMA, 128 arrivals)
  uint64_t* pv_done = p_full + 2;                  // [2]  O += P_j V_j retired (MMA -> softmax)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + 2);
  float* xchg = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);   // [2 slots][2 halves][128 rows]

  // the shuffle makes the warp index provably warp-uniform, so role code can use the uniform datapath
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) {
    printf("pio_flash_kernel: dynamic shared memory base is not 1024-byte aligned\n");
    __trap();
  }
  const int q0 = blockIdx.x * 128;
  const int b = blockIdx.y / p.H;
  const int h = blockIdx.y % p.H;
  const int split = blockIdx.z;
  const int total_tiles = (p.Nk + BN - 1) / BN;
  const int tile_begin = split * p.tiles_per_split;
  const int tile_end = min(total_tiles, tile_begin + p.tiles_per_split);
  const int ntiles = tile_end - tile_begin;  // host guarantees >= 1

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_k);
    if (!SAME) tma_prefetch_desc(&tmap_v);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(q_full, 1);
    for (int s = 0; s < Cfg::STAGES; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], 1);
    }
    mbar_init(&s_full[0], 1);
    mbar_init(&s_full[1], 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&p_full[i], 8 * kArrivalsPerWarp);
      mbar_init(&pv_done[i], 1);
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
  const uint32_t tmem_o = tmem_base + 2 * BN;
  pdl_sync();   // barriers / TMEM are set up; Q, K, V of the previous kernel are read from here on

  // number of 16-wide k-steps of the QK^T contraction in chunk c, and N extent of the PV product
  const int dqk_steps_total = (p.dqk + 15) / 16;
  const int dv_n = ((p.dv + 15) / 16) * 16;

  if (warp == 0) {
    // ================= TMA producer =================
    // (all 32 lanes run the schedule and the barrier waits, one elected lane issues: coordinates and addresses stay
    //  on the uniform datapath)
    const int bq = p.q_bcast ? 0 : b;
    if (elect_one()) {
      mbar_arrive_expect_tx(q_full, Cfg::Q_BYTES);
#pragma unroll
      for (int c = 0; c < NQC; ++c) tma_load_3d(sQ + c * 16384, &tmap_q, q_full, h * p.dqk + c * 64, q0, bq);
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
        for (int c = 0; c < NQC; ++c)
          tma_load_3d(st + c * Cfg::CHUNK_BYTES, &tmap_k, &kv_full[stage], h * p.dqk + c * 64, k0, b);
        if constexpr (!SAME) {
#pragma unroll
          for (int c = 0; c < NVC; ++c)
            tma_load_3d(st + (NQC + c) * Cfg::CHUNK_BYTES, &tmap_v, &kv_full[stage], h * p.dv + c * 64, k0, b);
        }
      }
      if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1u; }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    // All 32 lanes run the schedule and wait on the barriers; one elected lane issues each tcgen05 instruction and the
    // descriptors advance as 32-bit low words.  (Inside an `if (lane == 0)` region every MMA cost ~25 dependent vector
    // instructions + R2UR moves — several times the 32..128 cycles the MMA itself takes on the tensor pipe.)
    const uint32_t idesc_s = make_idesc_f16(128, BN, idesc_fmt(p.fp16), 0, 0);
    const uint64_t dq0 = make_smem_desc_sw128(smem_u32(sQ), 16, 1024);
    const uint64_t dk0 = make_smem_desc_sw128(smem_u32(sKV), 16, 1024);
    const uint64_t dv0 = make_smem_desc_sw128(smem_u32(sKV) + (SAME ? 0 : NQC * Cfg::CHUNK_BYTES), Cfg::CHUNK_BYTES, 1024);
    const uint32_t q_lo = (uint32_t)dq0, q_hi = (uint32_t)(dq0 >> 32);
    const uint32_t k_lo = (uint32_t)dk0, k_hi = (uint32_t)(dk0 >> 32);
    const uint32_t v_lo = (uint32_t)dv0, v_hi = (uint32_t)(dv0 >> 32);
    auto issue_s = [&](int j, int stage) {
      const uint32_t d = tmem_base + (j & 1) * BN;
      const uint32_t b0 = k_lo + (uint32_t)((stage * Cfg::STAGE_BYTES) >> 4);
      // this thread's issue rate bounds the kernel for the wide heads: widths that fill all but
      // the last three K steps of their last chunk (322 -> 21 steps, 261 -> 17) get a straight-line sequence without the
      // per-step bound check
      constexpr int FAST_STEPS = 4 * NQC - 3;
      if (dqk_steps_total == FAST_STEPS) {
#pragma unroll
        for (int ks = 0; ks < FAST_STEPS; ++ks) {
          const int c = ks >> 2, kk = ks & 3;
          if (elect_one())
            umma_ss_lh(d, q_lo + (uint32_t)((c * 16384 + kk * 32) >> 4), q_hi,
                       b0 + (uint32_t)((c * Cfg::CHUNK_BYTES + kk * 32) >> 4), k_hi, idesc_s, ks != 0 ? 1u : 0u);
        }
      } else {
#pragma unroll
        for (int ks = 0; ks < 4 * NQC; ++ks) {
          if (ks < dqk_steps_total) {
            const int c = ks >> 2, kk = ks & 3;
            if (elect_one())
              umma_ss_lh(d, q_lo + (uint32_t)((c * 16384 + kk * 32) >> 4), q_hi,
                         b0 + (uint32_t)((c * Cfg::CHUNK_BYTES + kk * 32) >> 4), k_hi, idesc_s, ks != 0 ? 1u : 0u);
          }
        }
      }
      if (elect_one()) umma_commit(&s_full[j & 1]);
    };
    auto issue_pv = [&](int j, int stage) {
      const uint32_t b0 = v_lo + (uint32_t)((stage * Cfg::STAGE_BYTES) >> 4);
      const uint32_t a0 = tmem_base + (j & 1) * BN;   // P_j overlays the first BN/2 columns of S_j (bf16 pairs)
      for (int nb = 0; nb * 256 < dv_n; ++nb) {
        const int n = min(256, dv_n - nb * 256);
        const uint32_t idesc_pv = make_idesc_f16(128, n, idesc_fmt(p.fp16), /*A (TMEM) K-major*/ 0, /*B MN-major*/ 1);
#pragma unroll
        for (int ks = 0; ks < BN / 16; ++ks) {
          if (elect_one())
            umma_ts_lh(tmem_o + nb * 256, a0 + ks * 8, b0 + (uint32_t)((nb * 4 * Cfg::CHUNK_BYTES + ks * 2048) >> 4), v_hi,
                       idesc_pv, (j | ks) != 0 ? 1u : 0u);
        }
      }
    };
    mbar_wait(q_full, 0);
    int stage = 0;       // stage of tile j
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
        issue_s(j + 1, nstage);   // overwrites S_{j-1} / P_{j-1}: in order after PV_{j-1}, which consumed P_{j-1}
      }
      mbar_wait(&p_full[j & 1], (j >> 1) & 1);
      tc_fence_after();
      issue_pv(j, stage);
      if (elect_one()) {
        umma_commit(&kv_empty[stage]);
        umma_commit(&pv_done[j & 1]);
      }
      stage = nstage;
      phase = nphase;
    }
  } else if (warp >= 4) {
    // ================= softmax / correction / epilogue =================
    constexpr int HC = BN / 2;            // key columns of a tile handled by this thread
    const int quarter = warp & 3;
    const int half = (warp - 4) >> 2;
    const int row = quarter * 32 + lane;  // row inside the tile == TMEM lane
    const int q = q0 + row;
    const uint32_t lane_off = static_cast<uint32_t>(quarter * 32) << 16;
    const uint8_t* km = p.key_mask ? p.key_mask + (long long)b * p.stride_km : nullptr;
    // O columns owned by this half for the rescale and the epilogue (32-column chunks)
    const int nchunks = (dv_n + 31) / 32;
    const int c_begin = half == 0 ? 0 : (nchunks + 1) / 2 * 32;
    const int c_end = half == 0 ? min(dv_n, (nchunks + 1) / 2 * 32) : dv_n;
    auto pair_sync = [&]() { asm volatile("bar.sync %0, 64;" ::"r"(1 + quarter) : "memory"); };
    float m = -INFINITY;  // running max of scale_log2 * s (identical in both halves)
    float l = 0.f;        // running sum of exp2(t - m) over this half's columns
    for (int j = 0; j < ntiles; ++j) {
      const int k0 = (tile_begin + j) * BN + half * HC;
      const uint32_t t_s = tmem_base + (j & 1) * BN + half * HC + lane_off;
      mbar_wait(&s_full[j & 1], (j >> 1) & 1);
      tc_fence_after();
      const bool tail = (k0 + HC > p.Nk) || (km != nullptr);
      // ---- this half of the S row moves to registers with one wait ----
      uint32_t r[HC];
#pragma unroll
      for (int c = 0; c < HC / 32; ++c) tmem_ld32(t_s + c * 32, *reinterpret_cast<uint32_t(*)[32]>(&r[c * 32]));
      tmem_wait_ld();
      if (tail) {
        // masked / out-of-range keys become -inf: they drop out of the max and exp2 turns them into exact zeros
#pragma unroll
        for (int i = 0; i < HC; ++i) {
          const int k = k0 + i;
          const bool ok = (k < p.Nk) && (km == nullptr || km[k] != 0);
          if (!ok) r[i] = 0xff800000u;
        }
      }
      float tmax;
      {
        float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
        for (int i = 0; i < HC / 8; ++i) {
          mx0 = fmax3(mx0, __uint_as_float(r[8 * i]), __uint_as_float(r[8 * i + 1]));
          mx1 = fmax3(mx1, __uint_as_float(r[8 * i + 2]), __uint_as_float(r[8 * i + 3]));
          mx2 = fmax3(mx2, __uint_as_float(r[8 * i + 4]), __uint_as_float(r[8 * i + 5]));
          mx3 = fmax3(mx3, __uint_as_float(r[8 * i + 6]), __uint_as_float(r[8 * i + 7]));
        }
        tmax = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
      }
      // ---- the row max over the whole tile: exchange with the partner warp (slots alternate with the tile parity, so
      //      one barrier per tile is enough) ----
      {
        float* slot = xchg + (j & 1) * 256;
        slot[half * 128 + row] = tmax;
        pair_sync();
        tmax = fmaxf(tmax, slot[(half ^ 1) * 128 + row]);
      }
      tmax *= p.scale_log2;  // scale > 0, so max commutes (an all-masked tile stays -inf)
      // ---- running max update, lazy rescale (both halves take identical decisions) ----
      float m_use = m;
      const bool grow = tmax > m + 8.0f;  // also true for the first valid tile (m == -inf)
      float alpha = 1.0f;
      if (grow) {
        alpha = (m == -INFINITY) ? 0.0f : exp2f(m - tmax);
        m_use = tmax;
      }
      const bool any_grow = __any_sync(0xffffffffu, grow && j > 0 && m != -INFINITY);
      if (any_grow) {
        // O may only be rescaled once PV_{j-1} has retired (rare: the running max grew by more than 2^8)
        mbar_wait(&pv_done[(j - 1) & 1], ((j - 1) >> 1) & 1);
        tc_fence_after();
#pragma unroll 1
        for (int c = c_begin; c < c_end; c += 16) {   // 16 columns at a time: the S row of this tile is live in registers
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
      // ---- p = exp2(scale * s - m) -> bf16 P written over S in TMEM; the row sum takes the un-rounded exponentials ----
      const uint64_t sc2 = pack_f32x2(p.scale_log2, p.scale_log2);
      const uint64_t nm2 = pack_f32x2(-msub, -msub);
      uint64_t la = pack_f32x2(0.f, 0.f), lb = pack_f32x2(0.f, 0.f);
#pragma unroll
      for (int c = 0; c < HC / 32; ++c) {
        uint32_t w[16];
        // one uniform branch per 32 keys selects the 16-bit format; the loop body itself stays select-free
        auto exp_block = [&](auto f16tag) {
          constexpr bool F16 = decltype(f16tag)::value;
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const uint64_t t2 = ffma2(pack_f32x2(__uint_as_float(r[32 * c + 2 * i]), __uint_as_float(r[32 * c + 2 * i + 1])),
                                      sc2, nm2);
            float t0, t1;
            unpack_f32x2(t2, t0, t1);
            const float e0 = ex2_approx(t0), e1 = ex2_approx(t1);
            w[i] = pack16x2<F16>(e0, e1);
            const uint64_t pr = pack_f32x2(e0, e1);   // un-rounded: RN is unbiased, see pio_flash2.cu
            if (i & 1) lb = fadd2(lb, pr);
            else la = fadd2(la, pr);
          }
        };
        if (p.fp16) exp_block(std::true_type{});
        else exp_block(std::false_type{});
        // P overwrites S in place (two bf16 values per 32-bit column).  Both warps of the pair have their S values in
        // registers (the pair barrier above), so no unread S column is clobbered.
        tmem_st16(tmem_base + (j & 1) * BN + lane_off + (half * HC + c * 32) / 2, w);
      }
      {
        float a0, a1, b0, b1;
        unpack_f32x2(la, a0, a1);
        unpack_f32x2(lb, b0, b1);
        l += (a0 + a1) + (b0 + b1);
      }
      tmem_wait_st();
      tc_fence_before();
      mbar_arrive_warp(&p_full[j & 1]);
    }
    // ---- epilogue: total row sum = both halves' partial sums ----
    {
      float* slot = xchg + (ntiles & 1) * 256;   // the slot the last tile did not use
      slot[half * 128 + row] = l;
      pair_sync();
      l += slot[(half ^ 1) * 128 + row];
    }
    mbar_wait(&pv_done[(ntiles - 1) & 1], ((ntiles - 1) >> 1) & 1);
    tc_fence_after();
    const bool keep = (q < p.Nq) && (p.row_keep == nullptr || p.row_keep[(long long)b * p.stride_rk + q] != 0);
    const bool emit_partial = p.partial || p.num_splits > 1;
    if (!emit_partial) {
      const float inv = (keep && l > 0.f) ? 1.0f / l : 0.0f;
      __nv_bfloat16* orow = p.O + (long long)b * p.strideO + (long long)q * p.ldo + (long long)h * p.dv;
      for (int c = c_begin; c < c_end; c += 32) {
        uint32_t r[32];
        tmem_ld32(tmem_o + lane_off + c, r);
        tmem_wait_ld();
        if (q < p.Nq) {
          __nv_bfloat16* op = orow + c;
          if (c + 32 <= p.dv && ((reinterpret_cast<uintptr_t>(op) & 15u) == 0)) {
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              uint4 w;
              w.x = pack16x2(__uint_as_float(r[8 * g]) * inv, __uint_as_float(r[8 * g + 1]) * inv, p.fp16);
              w.y = pack16x2(__uint_as_float(r[8 * g + 2]) * inv, __uint_as_float(r[8 * g + 3]) * inv, p.fp16);
              w.z = pack16x2(__uint_as_float(r[8 * g + 4]) * inv, __uint_as_float(r[8 * g + 5]) * inv, p.fp16);
              w.w = pack16x2(__uint_as_float(r[8 * g + 6]) * inv, __uint_as_float(r[8 * g + 7]) * inv, p.fp16);
              reinterpret_cast<uint4*>(op)[g] = w;
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (c + i < p.dv) reinterpret_cast<uint16_t*>(op)[i] = cvt16(__uint_as_float(r[i]) * inv, p.fp16);
          }
        }
      }
    } else {
      const long long prow = (((long long)split * p.B + b) * p.H + h) * p.Nq + q;
      if (q < p.Nq && half == 0) {
        p.m_part[prow] = (m == -INFINITY) ? -INFINITY : m * 0.69314718055994531f;  // back to natural-log units
        p.l_part[prow] = l;
      }
      float* orow = p.O_part + prow * p.dv;
      for (int c = c_begin; c < c_end; c += 32) {
        uint32_t r[32];
        tmem_ld32(tmem_o + lane_off + c, r);
        tmem_wait_ld();
        if (q < p.Nq) {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (c + i < p.dv) orow[c + i] = (l > 0.f) ? __uint_as_float(r[i]) : 0.0f;
        }
      }
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

template <int NQC, int NVC, bool SAME>
static int launch_flash(const pio_attention_args* a, const DeviceInfo& dev, cudaStream_t stream) {
  using Cfg = FlashCfg<NQC, NVC, SAME>;
  static_assert(Cfg::VALID, "flash configuration does not fit");
  constexpr int BN = Cfg::BN;
  CUtensorMap tq, tk, tv;
  const int q_bcast = (a->strideQ == 0 && a->B > 1) ? 1 : 0;
  {
    const uint64_t dims[3] = {(uint64_t)a->H * a->dqk, (uint64_t)a->Nq, (uint64_t)(q_bcast ? 1 : a->B)};
    const uint64_t strides[2] = {(uint64_t)a->ldq * 2,
                                 (uint64_t)((q_bcast || a->B == 1) ? a->ldq * (int64_t)a->Nq : a->strideQ) * 2};
    const uint32_t box[3] = {64, 128, 1};
    int rc = encode_tmap_bf16(&tq, a->Q, 3, dims, strides, box);
    if (rc != PIO_OK) return rc;
  }
  {
    const uint64_t dims[3] = {(uint64_t)a->H * a->dqk, (uint64_t)a->Nk, (uint64_t)a->B};
    const uint64_t strides[2] = {(uint64_t)a->ldk * 2, (uint64_t)(a->B == 1 ? a->ldk * (int64_t)a->Nk : a->strideK) * 2};
    const uint32_t box[3] = {64, (uint32_t)BN, 1};
    int rc = encode_tmap_bf16(&tk, a->K, 3, dims, strides, box);
    if (rc != PIO_OK) return rc;
  }
  {
    const uint64_t dims[3] = {(uint64_t)a->H * a->dv, (uint64_t)a->Nk, (uint64_t)a->B};
    const uint64_t strides[2] = {(uint64_t)a->ldv * 2, (uint64_t)(a->B == 1 ? a->ldv * (int64_t)a->Nk : a->strideV) * 2};
    const uint32_t box[3] = {64, (uint32_t)BN, 1};
    int rc = encode_tmap_bf16(&tv, a->V, 3, dims, strides, box);
    if (rc != PIO_OK) return rc;
  }
  FlashParams p;
  p.B = a->B; p.H = a->H; p.Nq = a->Nq; p.Nk = a->Nk; p.dqk = a->dqk; p.dv = a->dv;
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
  // the caller sized the partial buffers for a->num_splits; every split must own >= 1 tile
  if (a->num_splits > 1 && (long long)(a->num_splits - 1) * p.tiles_per_split >= total_tiles)
    return fail(PIO_ERR_INVALID_ARGUMENT,
                "pio_attention_fwd: num_splits=%d leaves an empty split (Nk=%d, %d-key tiles: %d); use <= %d splits that "
                "divide evenly", a->num_splits, a->Nk, BN, total_tiles, total_tiles);
  p.num_splits = a->num_splits < 1 ? 1 : a->num_splits;
  p.partial = a->partial;
  p.O_part = a->O_part; p.m_part = a->m_part; p.l_part = a->l_part;

  static PerDeviceOnce once;
  const cudaError_t attr_err = once.run(dev.device, [] {
    return cudaFuncSetAttribute(pio_flash_kernel<NQC, NVC, SAME>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                Cfg::SMEM_BYTES);
  });
  if (attr_err != cudaSuccess)
    return fail(PIO_ERR_CUDA, "cudaFuncSetAttribute(flash<%d,%d,%d>) failed: %s", NQC, NVC, (int)SAME,
                cudaGetErrorString(attr_err));
  dim3 grid((a->Nq + 127) / 128, a->B * a->H, p.num_splits);
  {
    ProfileScope prof(KF_FLASH, 2.0 * a->B * a->H * (double)a->Nq * a->Nk * (a->dqk + a->dv), 0.0, stream);
    PIO_CUDA_OK(launch_kernel(pio_flash_kernel<NQC, NVC, SAME>, grid, dim3(384, 1, 1), Cfg::SMEM_BYTES, stream, 1, tq, tk,
                              tv, p));
  }
  g_launch_count.fetch_add(1);
  PIO_CUDA_OK(cudaGetLastError());
  return PIO_OK;
}

static bool flash_shape_ok(int dqk, int dv, bool same) {
  const int nqc = (dqk + 63) / 64, nvc = (dv + 63) / 64;
  if (same) return nqc >= 1 && nqc <= 6 && dqk == dv;
  return nqc >= 1 && nqc <= 2 && nvc >= 1 && nvc <= 3;
}

}  // namespace pio

extern "C" int pio_attention_supported(int32_t dqk, int32_t dv) {
  return (pio::flash_shape_ok(dqk, dv, false) || pio::flash_shape_ok(dqk, dv, true)) ? PIO_OK : PIO_ERR_UNSUPPORTED;
}

// Tile width in keys the kernel uses for these head sizes (hosts use it to pick an even key split).
extern "C" int pio_attention_key_tile(int32_t dqk, int32_t dv, int32_t same_kv) {
  using namespace pio;
  const int nqc = (dqk + 63) / 64, nvc = (dv + 63) / 64;
  if (same_kv && dqk == dv && flash_qt_key_tile(dqk) > 0) {
    static const int off = [] { const char* e = getenv("PIO_FLASH_QT"); return (e && e[0] == '0') ? 1 : 0; }();
    if (!off) return flash_qt_key_tile(dqk);
  }
  if (same_kv && flash_shape_ok(dqk, dv, true)) {
    switch (nqc) {
      case 1: return FlashCfg<1, 1, true>::BN;
      case 2: return FlashCfg<2, 2, true>::BN;
      case 3: return FlashCfg<3, 3, true>::BN;
      case 4: return FlashCfg<4, 4, true>::BN;
      case 5: return FlashCfg<5, 5, true>::BN;
      case 6: return FlashCfg<6, 6, true>::BN;
    }
  }
  if (flash_shape_ok(dqk, dv, false)) {
    switch (nqc * 10 + nvc) {
      case 11: return FlashCfg<1, 1, false>::BN;
      case 12: return FlashCfg<1, 2, false>::BN;
      case 13: return FlashCfg<1, 3, false>::BN;
      case 21: return FlashCfg<2, 1, false>::BN;
      case 22: return FlashCfg<2, 2, false>::BN;
      case 23: return FlashCfg<2, 3, false>::BN;
    }
  }
  return PIO_ERR_UNSUPPORTED;
}

extern "C" int pio_attention_fwd(const pio_attention_args* a, void* stream_) {
  using namespace pio;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  PIO_REQUIRE(a != nullptr, "pio_attention_fwd: null args");
  PIO_REQUIRE(a->Q && a->K && a->V, "pio_attention_fwd: null operand");
  PIO_REQUIRE(a->B > 0 && a->H > 0 && a->Nq > 0 && a->Nk > 0 && a->dqk > 0 && a->dv > 0, "pio_attention_fwd: bad shape");
  PIO_REQUIRE(aligned16(a->Q) && aligned16(a->K) && aligned16(a->V), "pio_attention_fwd: operand base not 16-byte aligned");
  PIO_REQUIRE(a->ldq % 8 == 0 && a->ldk % 8 == 0 && a->ldv % 8 == 0, "pio_attention_fwd: leading dims must be multiples of 8");
  PIO_REQUIRE(a->strideQ % 8 == 0 && a->strideK % 8 == 0 && a->strideV % 8 == 0,
              "pio_attention_fwd: batch strides must be multiples of 8");
  PIO_REQUIRE(a->H == 1 || (a->dqk % 16 == 0),
              "pio_attention_fwd: multi-head needs dqk %% 16 == 0 (got %d); re-lay the heads out first", a->dqk);
  const bool emit_partial = a->partial || a->num_splits > 1;
  PIO_REQUIRE(emit_partial ? (a->O_part && a->m_part && a->l_part) : (a->O != nullptr),
              "pio_attention_fwd: missing output buffer");
  PIO_REQUIRE((long long)a->B * a->H < 65536 && a->num_splits < 65536, "pio_attention_fwd: grid too large");
  DeviceInfo dev;
  int rc = get_device_info(&dev);
  if (rc != PIO_OK) return rc;
  if (dev.cc_major != 10) return fail(PIO_ERR_ARCH, "pio_attention_fwd needs sm_100 (got sm_%d%d)", dev.cc_major, dev.cc_minor);
  {
    // short-key / many-head shapes (the latent tower): persistent two-tile kernel with P kept in TMEM
    static const int force = [] { const char* e = getenv("PIO_FLASH_KERNEL"); return e ? atoi(e) : 0; }();
    if (force != 1 && flash2_eligible(a)) {
      const long long items = (long long)a->B * a->H * ((a->Nq + 255) / 256);
      if (force == 2 || items * 2 >= dev.sm_count || a->Nk <= 4096) return launch_flash2(a, dev, stream);
    }
  }
  const bool same = (a->K == a->V) && (a->ldk == a->ldv) && (a->dqk == a->dv) && (a->strideK == a->strideV);
  const int nqc = (a->dqk + 63) / 64, nvc = (a->dv + 63) / 64;
  if (flash_qt_eligible(a)) return launch_flash_qt(a, dev, stream);   // query tile in TMEM (192 < d <= 272, K == V)
  if (same && flash_shape_ok(a->dqk, a->dv, true)) {
    switch (nqc) {
      case 1: return launch_flash<1, 1, true>(a, dev, stream);
      case 2: return launch_flash<2, 2, true>(a, dev, stream);
      case 3: return launch_flash<3, 3, true>(a, dev, stream);
      case 4: return launch_flash<4, 4, true>(a, dev, stream);
      case 5: return launch_flash<5, 5, true>(a, dev, stream);
      case 6: return launch_flash<6, 6, true>(a, dev, stream);
    }
  }
  if (flash_shape_ok(a->dqk, a->dv, false)) {
    switch (nqc * 10 + nvc) {
      case 11: return launch_flash<1, 1, false>(a, dev, stream);
      case 12: return launch_flash<1, 2, false>(a, dev, stream);
      case 13: return launch_flash<1, 3, false>(a, dev, stream);
      case 21: return launch_flash<2, 1, false>(a, dev, stream);
      case 22: return launch_flash<2, 2, false>(a, dev, stream);
      case 23: return launch_flash<2, 3, false>(a, dev, stream);
    }
  }
  return fail(PIO_ERR_UNSUPPORTED, "pio_attention_fwd: head sizes dqk=%d dv=%d (same_kv=%d) not covered by the streaming kernel",
              a->dqk, a->dv, (int)same);
}
