// Persistent two-tile streaming attention for sm_100a: the latent-tower (and other short-key, many-head) shapes.
//
// Each CTA loops over work items (batch, head, 256-query block).  An item is TWO 128-query tiles A and B that share
// every K/V tile: the tensor pipe alternates  S_A, S_B | PV_A, S_A' | PV_B, S_B' | ...  so that while the four
// softmax warps of one tile run, the MMAs of the other tile execute.  S (fp32) lives in TMEM; the softmax warps
// overwrite it in place with bf16 P (tcgen05.st, two values per 32-bit column) and O += P.V is issued with the A operand
// read from TMEM (tcgen05.mma [d], [a_tmem], b_desc) — P never touches shared memory, which is the binding resource of
// the one-tile kernel (pio_flash.cu: per key tile Q, K, P and V are all read from smem, plus the P writes).  K and V
// stream through separate TMA rings; the next item's Q, K and V are prefetched as soon as the last S MMAs of the
// current item retire, so the per-item prologue is hidden.
//
// Roles (384 threads = 3 warpgroups): warpgroup 0 = warp 0 TMA producer, warp 1 MMA issuer, warp 2 TMEM allocator,
// warp 3 idle; warpgroup 1 (warps 4..7) softmax / correction / epilogue of tile A, warpgroup 2 (warps 8..11) the same
// for tile B (thread = query row = TMEM lane).  The kernel launches with 168 registers per thread; warpgroup 0 then
// shrinks to 96 (setmaxnreg.dec) and the softmax warpgroups grow to 200 (setmaxnreg.inc) so that a softmax thread holds
// its whole 128-value S row in registers (one TMEM read per tile instead of a max pass plus an exp pass).
// TMEM columns: S_A | S_B | O_A | O_B  =  2 x BN + 2 x 64 NVC  <= 512.
#include <math.h>
#include <stdlib.h>

#include "pio_common.cuh"
#include "pio_host.h"

namespace pio {

struct Flash2Params {
  int B, H, Nq, Nk, dqk, dv;
  int q_bcast;
  float scale_log2;
  const uint8_t* key_mask; long long stride_km;
  const uint8_t* row_keep; long long stride_rk;
  __nv_bfloat16* O; long long ldo, strideO;
  int q_pairs;    // ceil(Nq / 256)
  int items;      // B * H * q_pairs
  int kv_tiles;   // ceil(Nk / BN)
};

template <int NQC, int NVC, int BN>
struct Flash2Cfg {
  static constexpr int Q_TILE_BYTES = NQC * 16384;            // 128 rows x 64-column chunks
  static constexpr int Q_BYTES = 2 * Q_TILE_BYTES;
  static constexpr int CHUNK_BYTES = BN * 128;                // one 64-column chunk of a K or V tile
  static constexpr int K_BYTES = NQC * CHUNK_BYTES;
  static constexpr int V_BYTES = NVC * CHUNK_BYTES;
  static constexpr int BAR_BYTES = 512;
  static constexpr int AVAIL = 232448 - BAR_BYTES - Q_BYTES;
  static constexpr int STAGES_RAW = AVAIL / (K_BYTES + V_BYTES);
  static constexpr int STAGES = STAGES_RAW > 4 ? 4 : STAGES_RAW;
  static constexpr int TMEM_NEED = 2 * BN + 2 * NVC * 64;
  static constexpr int TMEM_COLS = TMEM_NEED <= 128 ? 128 : (TMEM_NEED <= 256 ? 256 : 512);
  static constexpr int SMEM_USED = Q_BYTES + STAGES * (K_BYTES + V_BYTES) + BAR_BYTES;
  // more than half of the SM's shared memory: exactly one CTA per SM, TMEM is never oversubscribed
  static constexpr int SMEM_BYTES = SMEM_USED < 120 * 1024 ? 120 * 1024 : SMEM_USED;
  static constexpr bool VALID = STAGES >= 2 && TMEM_NEED <= 512;
};

template <int NQC, int NVC, int BN>
__global__ void __launch_bounds__(384, 1)
pio_flash2_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                  const __grid_constant__ CUtensorMap tmap_v, const Flash2Params p) {
  using Cfg = Flash2Cfg<NQC, NVC, BN>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem;                                   // [2 tiles][NQC chunks][128 x 128 B]
  uint8_t* sK = sQ + Cfg::Q_BYTES;                      // [STAGES][NQC][BN x 128 B]
  uint8_t* sV = sK + STAGES * Cfg::K_BYTES;             // [STAGES][NVC][BN x 128 B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + STAGES * Cfg::V_BYTES);
  uint64_t* q_full = bars;                              // TMA -> MMA
  uint64_t* q_empty = bars + 1;                         // MMA (last S of the item retired) -> TMA
  uint64_t* k_full = bars + 2;                          // [STAGES]
  uint64_t* k_empty = k_full + STAGES;
  uint64_t* v_full = k_empty + STAGES;
  uint64_t* v_empty = v_full + STAGES;
  uint64_t* s_full = v_empty + STAGES;                  // [2]  S_X ready                  (MMA -> softmax X)
  uint64_t* p_full = s_full + 2;                        // [2]  P_X in TMEM, O_X rescaled  (softmax X -> MMA), 128 arrivals
  uint64_t* pv_done = p_full + 2;                       // [2]  O_X += P_X V retired        (MMA -> softmax X)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + 2);

  // the shuffle makes the warp index provably warp-uniform, so role code can use the uniform datapath
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) {
    printf("pio_flash2_kernel: dynamic shared memory base is not 1024-byte aligned\n");
    __trap();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_v);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(q_full, 1);
    mbar_init(q_empty, 1);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&k_empty[s], 1);
      mbar_init(&v_full[s], 1);
      mbar_init(&v_empty[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 128);
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

  const int T = p.kv_tiles;
  const int dqk_steps = (p.dqk + 15) / 16;
  const int dv_n = ((p.dv + 15) / 16) * 16;

  // register re-balancing between the warpgroups: every warp of a warpgroup executes its setmaxnreg, and each role's
  // code is dominated by its own setmaxnreg
  if (warp < 4) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 96;");
  if (warp == 0) {
    {
      // ================= TMA producer =================
      // (all 32 lanes run the schedule so that addresses / coordinates stay in uniform registers; one lane issues)
      const bool leader = (lane == 0);
      uint32_t kv = 0;
      int it = 0;
      for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++it) {
        const int qp = item % p.q_pairs;
        const int bh = item / p.q_pairs;
        const int h = bh % p.H, b = bh / p.H;
        mbar_wait(q_empty, (uint32_t)(it & 1) ^ 1u);
        if (leader) {
          mbar_arrive_expect_tx(q_full, Cfg::Q_BYTES);
#pragma unroll
          for (int x = 0; x < 2; ++x)
#pragma unroll
            for (int c = 0; c < NQC; ++c)
              tma_load_3d(sQ + (x * NQC + c) * 16384, &tmap_q, q_full, h * p.dqk + c * 64, qp * 256 + x * 128,
                          p.q_bcast ? 0 : b);
        }
        for (int j = 0; j < T; ++j, ++kv) {
          const int stage = kv % STAGES;
          const uint32_t ph = (kv / STAGES) & 1u;
          mbar_wait(&k_empty[stage], ph ^ 1u);
          if (leader) {
            mbar_arrive_expect_tx(&k_full[stage], Cfg::K_BYTES);
#pragma unroll
            for (int c = 0; c < NQC; ++c)
              tma_load_3d(sK + (stage * NQC + c) * Cfg::CHUNK_BYTES, &tmap_k, &k_full[stage], h * p.dqk + c * 64, j * BN, b);
          }
          mbar_wait(&v_empty[stage], ph ^ 1u);
          if (leader) {
            mbar_arrive_expect_tx(&v_full[stage], Cfg::V_BYTES);
#pragma unroll
            for (int c = 0; c < NVC; ++c)
              tma_load_3d(sV + (stage * NVC + c) * Cfg::CHUNK_BYTES, &tmap_v, &v_full[stage], h * p.dv + c * 64, j * BN, b);
          }
        }
      }
    }
  } else if (warp == 1) {
    {
      // ================= MMA issuer =================
      // All 32 lanes run the schedule and wait on the barriers; one lane issues each tcgen05 instruction.  Inside an
      // `if (lane == 0)` region the compiler cannot use the uniform datapath, and every descriptor of every MMA costs
      // vector arithmetic plus R2UR moves (~20 dependent instructions per 64-cycle MMA: the issuer, not the tensor
      // pipe, paced the first version of this kernel).
      const bool leader = (lane == 0);
      constexpr uint32_t idesc_s = make_idesc_f16(128, BN, 1, 0, 0);
      const uint32_t idesc_pv = make_idesc_f16(128, dv_n, 1, /*A (TMEM) K-major*/ 0, /*B MN-major*/ 1);
      const uint64_t dq0 = make_smem_desc_sw128(smem_u32(sQ), 16, 1024);
      const uint64_t dk0 = make_smem_desc_sw128(smem_u32(sK), 16, 1024);
      const uint64_t dv0 = make_smem_desc_sw128(smem_u32(sV), Cfg::CHUNK_BYTES, 1024);
      const uint32_t q_lo = (uint32_t)dq0, q_hi = (uint32_t)(dq0 >> 32);
      const uint32_t k_lo = (uint32_t)dk0, k_hi = (uint32_t)(dk0 >> 32);
      const uint32_t v_lo = (uint32_t)dv0, v_hi = (uint32_t)(dv0 >> 32);
      auto issue_s = [&](int x, int stage) {
        const uint32_t d = tmem_base + x * BN;
        const uint32_t a0 = q_lo + (uint32_t)((x * Cfg::Q_TILE_BYTES) >> 4);
        const uint32_t b0 = k_lo + (uint32_t)((stage * Cfg::K_BYTES) >> 4);
        if (dqk_steps == 4 * NQC) {   // full-width heads: no per-instruction bound check in the issue stream
#pragma unroll
          for (int ks = 0; ks < 4 * NQC; ++ks) {
            const int c = ks >> 2, kk = ks & 3;
            if (elect_one())
              umma_ss_lh(d, a0 + (uint32_t)((c * 16384 + kk * 32) >> 4), q_hi,
                         b0 + (uint32_t)((c * Cfg::CHUNK_BYTES + kk * 32) >> 4), k_hi, idesc_s, ks != 0 ? 1u : 0u);
          }
        } else {
#pragma unroll
          for (int ks = 0; ks < 4 * NQC; ++ks) {
            if (ks < dqk_steps) {
              const int c = ks >> 2, kk = ks & 3;
              if (elect_one())
                umma_ss_lh(d, a0 + (uint32_t)((c * 16384 + kk * 32) >> 4), q_hi,
                           b0 + (uint32_t)((c * Cfg::CHUNK_BYTES + kk * 32) >> 4), k_hi, idesc_s, ks != 0 ? 1u : 0u);
            }
          }
        }
      };
      auto issue_pv = [&](int x, int stage, bool accum) {
        const uint32_t a = tmem_base + x * BN;          // P_x overlays the first BN/2 columns of S_x
        const uint32_t d = tmem_o + x * (NVC * 64);
        const uint32_t b0 = v_lo + (uint32_t)((stage * Cfg::V_BYTES) >> 4);
#pragma unroll
        for (int ks = 0; ks < BN / 16; ++ks) {
          if (elect_one())
            umma_ts_lh(d, a + ks * 8, b0 + (uint32_t)((ks * 2048) >> 4), v_hi, idesc_pv, (accum || ks != 0) ? 1u : 0u);
        }
      };
      uint32_t kv = 0;
      uint32_t pcount[2] = {0, 0};
      int it = 0;
      for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++it) {
        mbar_wait(q_full, (uint32_t)(it & 1));
        {
          const int stage = kv % STAGES;
          mbar_wait(&k_full[stage], (kv / STAGES) & 1u);
          tc_fence_after();
          issue_s(0, stage);
          if (leader) umma_commit(&s_full[0]);
          issue_s(1, stage);
          if (leader) umma_commit(&s_full[1]);
          if (leader) umma_commit(&k_empty[stage]);
          if (T == 1 && leader) umma_commit(q_empty);
        }
        for (int j = 0; j < T; ++j) {
          const uint32_t g = kv + j;
          const int stage = g % STAGES;
          mbar_wait(&v_full[stage], (g / STAGES) & 1u);
          const int nstage = (g + 1) % STAGES;
          const uint32_t nph = ((g + 1) / STAGES) & 1u;
#pragma unroll
          for (int x = 0; x < 2; ++x) {
            mbar_wait(&p_full[x], pcount[x] & 1u);
            ++pcount[x];
            tc_fence_after();
            issue_pv(x, stage, j > 0);
            if (leader) umma_commit(&pv_done[x]);
            if (x == 1 && leader) umma_commit(&v_empty[stage]);
            if (j + 1 < T) {
              if (x == 0) {
                mbar_wait(&k_full[nstage], nph);
                tc_fence_after();
              }
              issue_s(x, nstage);   // in-order after PV_x(j): P_x(j) has been consumed before S_x is overwritten
              if (leader) umma_commit(&s_full[x]);
              if (x == 1) {
                if (leader) umma_commit(&k_empty[nstage]);
                if (j + 2 == T && leader) umma_commit(q_empty);
              }
            }
          }
        }
        kv += T;
      }
    }
  }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 200;");
    // ================= softmax / correction / epilogue of tile X =================
    const int x = (warp - 4) >> 2;
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;  // row inside the tile == TMEM lane
    const uint32_t lane_off = static_cast<uint32_t>(quarter * 32) << 16;
    const uint32_t t_s = tmem_base + x * BN + lane_off;
    const uint32_t t_o = tmem_o + x * (NVC * 64) + lane_off;
    uint32_t g = 0;  // running key-tile index of this tile slot (phase bookkeeping across items)
    for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
      const int qp = item % p.q_pairs;
      const int bh = item / p.q_pairs;
      const int h = bh % p.H, b = bh / p.H;
      const int q = qp * 256 + x * 128 + row;
      const uint8_t* km = p.key_mask ? p.key_mask + (long long)b * p.stride_km : nullptr;
      float m = -INFINITY;  // running max of scale_log2 * s
      float l = 0.f;        // running sum of exp2(t - m)
      for (int j = 0; j < T; ++j, ++g) {
        const int k0 = j * BN;
        mbar_wait(&s_full[x], g & 1u);
        tc_fence_after();
        const bool tail = (k0 + BN > p.Nk) || (km != nullptr);
        // ---- the whole S row of this tile (BN fp32 values) moves to registers with one wait ----
        uint32_t r[BN];
#pragma unroll
        for (int c = 0; c < BN / 32; ++c) tmem_ld32(t_s + c * 32, *reinterpret_cast<uint32_t(*)[32]>(&r[c * 32]));
        tmem_wait_ld();
        if (tail) {
          // masked / out-of-range keys become -inf: they drop out of the max and exp2 turns them into exact zeros
#pragma unroll
          for (int i = 0; i < BN; ++i) {
            const int k = k0 + i;
            const bool ok = (k < p.Nk) && (km == nullptr || km[k] != 0);
            if (!ok) r[i] = 0xff800000u;
          }
        }
        float tmax;
        {
          float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
          for (int i = 0; i < BN / 8; ++i) {
            mx0 = fmax3(mx0, __uint_as_float(r[8 * i]), __uint_as_float(r[8 * i + 1]));
            mx1 = fmax3(mx1, __uint_as_float(r[8 * i + 2]), __uint_as_float(r[8 * i + 3]));
            mx2 = fmax3(mx2, __uint_as_float(r[8 * i + 4]), __uint_as_float(r[8 * i + 5]));
            mx3 = fmax3(mx3, __uint_as_float(r[8 * i + 6]), __uint_as_float(r[8 * i + 7]));
          }
          tmax = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
        }
        tmax *= p.scale_log2;  // scale > 0, so max commutes (an all-masked tile stays -inf)
        // ---- running max update, lazy rescale ----
        float m_use = m;
        const bool grow = tmax > m + 8.0f;  // also true for the first valid tile (m == -inf)
        float alpha = 1.0f;
        if (grow) {
          alpha = (m == -INFINITY) ? 0.0f : exp2f(m - tmax);
          m_use = tmax;
        }
        const bool any_grow = __any_sync(0xffffffffu, grow && j > 0 && m != -INFINITY);
        if (any_grow) {
          // O may only be rescaled once PV_x(j-1) has retired (rare: the running max grew by more than 2^8)
          mbar_wait(&pv_done[x], (g - 1u) & 1u);
          tc_fence_after();
          // (16 columns at a time: the S row of this tile is live in registers)
#pragma unroll 1
          for (int c = 0; c < dv_n; c += 16) {
            uint32_t o[16];
            tmem_ld16(t_o + c, o);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st16(t_o + c, o);
          }
        }
        l *= alpha;
        m = m_use;
        const float msub = (m == -INFINITY) ? 0.0f : m;
        // ---- p = exp2(scale * s - m) -> bf16 pairs written over S in TMEM (two values per 32-bit column); the row sum
        //      uses the bf16-rounded values the tensor core will multiply, so P and l stay consistent ----
        const uint64_t sc2 = pack_f32x2(p.scale_log2, p.scale_log2);
        const uint64_t nm2 = pack_f32x2(-msub, -msub);
        uint64_t la = pack_f32x2(0.f, 0.f), lb = pack_f32x2(0.f, 0.f);
#pragma unroll
        for (int c = 0; c < BN / 32; ++c) {
          uint32_t w[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const uint64_t t2 = ffma2(pack_f32x2(__uint_as_float(r[32 * c + 2 * i]), __uint_as_float(r[32 * c + 2 * i + 1])),
                                      sc2, nm2);
            float t0, t1;
            unpack_f32x2(t2, t0, t1);
            w[i] = pack_bf16x2(ex2_approx(t0), ex2_approx(t1));
            const uint64_t pr = pack_f32x2(__uint_as_float(w[i] << 16), __uint_as_float(w[i] & 0xffff0000u));
            if (i & 1) lb = fadd2(lb, pr);
            else la = fadd2(la, pr);
          }
          tmem_st16(t_s + c * 16, w);
        }
        float lsum;
        {
          float a0, a1, b0, b1;
          unpack_f32x2(la, a0, a1);
          unpack_f32x2(lb, b0, b1);
          lsum = (a0 + a1) + (b0 + b1);
        }
        l += lsum;
        tmem_wait_st();
        tc_fence_before();
        mbar_arrive(&p_full[x]);
      }
      // ---- epilogue of the item: O_x / l -> bf16 ----
      mbar_wait(&pv_done[x], (g - 1u) & 1u);
      tc_fence_after();
      const bool keep = (q < p.Nq) && (p.row_keep == nullptr || p.row_keep[(long long)b * p.stride_rk + q] != 0);
      const float inv = (keep && l > 0.f) ? 1.0f / l : 0.0f;
      __nv_bfloat16* orow = p.O + (long long)b * p.strideO + (long long)q * p.ldo + (long long)h * p.dv;
      for (int c = 0; c < dv_n; c += 32) {
        uint32_t r[32];
        tmem_ld32(t_o + c, r);
        tmem_wait_ld();
        if (q < p.Nq) {
          __nv_bfloat16* op = orow + c;
          if (c + 32 <= p.dv && ((reinterpret_cast<uintptr_t>(op) & 15u) == 0)) {
#pragma unroll
            for (int gq = 0; gq < 4; ++gq) {
              uint4 w;
              w.x = pack_bf16x2(__uint_as_float(r[8 * gq]) * inv, __uint_as_float(r[8 * gq + 1]) * inv);
              w.y = pack_bf16x2(__uint_as_float(r[8 * gq + 2]) * inv, __uint_as_float(r[8 * gq + 3]) * inv);
              w.z = pack_bf16x2(__uint_as_float(r[8 * gq + 4]) * inv, __uint_as_float(r[8 * gq + 5]) * inv);
              w.w = pack_bf16x2(__uint_as_float(r[8 * gq + 6]) * inv, __uint_as_float(r[8 * gq + 7]) * inv);
              reinterpret_cast<uint4*>(op)[gq] = w;
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (c + i < p.dv) op[i] = __float2bfloat16_rn(__uint_as_float(r[i]) * inv);
          }
        }
      }
      // the next item's first p_full arrival (after tc_fence_before) orders these O reads before PV_x overwrites O_x
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

template <int NQC, int NVC, int BN>
static int launch_flash2_cfg(const pio_attention_args* a, const DeviceInfo& dev, cudaStream_t stream) {
  using Cfg = Flash2Cfg<NQC, NVC, BN>;
  static_assert(Cfg::VALID, "flash2 configuration does not fit");
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
  Flash2Params p;
  p.B = a->B; p.H = a->H; p.Nq = a->Nq; p.Nk = a->Nk; p.dqk = a->dqk; p.dv = a->dv;
  p.q_bcast = q_bcast;
  p.scale_log2 = a->scale * 1.4426950408889634f;
  p.key_mask = a->key_mask; p.stride_km = a->stride_km;
  p.row_keep = a->row_keep; p.stride_rk = a->stride_rk;
  p.O = reinterpret_cast<__nv_bfloat16*>(a->O); p.ldo = a->ldo; p.strideO = a->strideO;
  p.q_pairs = (a->Nq + 255) / 256;
  p.items = a->B * a->H * p.q_pairs;
  p.kv_tiles = (a->Nk + BN - 1) / BN;

  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(pio_flash2_kernel<NQC, NVC, BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    Cfg::SMEM_BYTES);
  });
  if (attr_err != cudaSuccess)
    return fail(PIO_ERR_CUDA, "cudaFuncSetAttribute(flash2<%d,%d,%d>) failed: %s", NQC, NVC, BN,
                cudaGetErrorString(attr_err));
  const int grid = p.items < dev.sm_count ? p.items : dev.sm_count;
  {
    ProfileScope prof(KF_FLASH, 2.0 * a->B * a->H * (double)a->Nq * a->Nk * (a->dqk + a->dv), 0.0, stream);
    PIO_CUDA_OK(launch_kernel(pio_flash2_kernel<NQC, NVC, BN>, dim3((unsigned)grid, 1, 1), dim3(384, 1, 1), Cfg::SMEM_BYTES,
                              stream, 1, tq, tk, tv, p));
  }
  g_launch_count.fetch_add(1);
  PIO_CUDA_OK(cudaGetLastError());
  return PIO_OK;
}

// Shapes the two-tile kernel covers: head sizes d_qk <= 128, d_v <= 192, normalised bf16 output only.
bool flash2_eligible(const pio_attention_args* a) {
  if (a->partial || a->num_splits > 1 || a->O == nullptr) return false;
  if (a->dqk > 128 || a->dv > 192) return false;
  if (a->H > 1 && (a->dqk % 16 != 0)) return false;
  if ((long long)a->B * a->H * ((a->Nq + 255) / 256) > 0x7fffffffLL) return false;
  return true;
}

int launch_flash2(const pio_attention_args* a, const DeviceInfo& dev, cudaStream_t stream) {
  const int nqc = (a->dqk + 63) / 64, nvc = (a->dv + 63) / 64;
  switch (nqc * 10 + nvc) {
    case 11: return launch_flash2_cfg<1, 1, 128>(a, dev, stream);
    case 12: return launch_flash2_cfg<1, 2, 128>(a, dev, stream);
    case 13: return launch_flash2_cfg<1, 3, 64>(a, dev, stream);
    case 21: return launch_flash2_cfg<2, 1, 128>(a, dev, stream);
    case 22: return launch_flash2_cfg<2, 2, 128>(a, dev, stream);
    case 23: return launch_flash2_cfg<2, 3, 64>(a, dev, stream);
  }
  return fail(PIO_ERR_UNSUPPORTED, "flash2: head sizes dqk=%d dv=%d not covered", a->dqk, a->dv);
}

}  // namespace pio
