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

#include <type_traits>

#include "pio_common.cuh"
#include "pio_host.h"

namespace pio {

// Developer aid (compiled out unless -DPIO_FLASH2_TRACE): CTA 0 records (tag, clock64) pairs at the pipeline's
// hand-off points; the first launch prints them to stderr.  Tags: 1xx MMA issuer, 2xx softmax A, 3xx softmax B.
#ifdef PIO_FLASH2_TRACE
// fire-and-forget stores into a per-warp slice (no atomics: a returning atomic would stall the traced warp)
__device__ unsigned long long g_f2_trace[12 * 2 * 512];
#define F2T(tag)                                                                        \
  do {                                                                                  \
    if (blockIdx.x == 0 && lane == 0 && f2n < 512) {                                    \
      g_f2_trace[(warp * 512 + f2n) * 2] = (unsigned long long)(tag);                   \
      g_f2_trace[(warp * 512 + f2n) * 2 + 1] = (unsigned long long)clock64();           \
      ++f2n;                                                                            \
    }                                                                                   \
  } while (0)
#else
#define F2T(tag)
#endif

struct Flash2Params {
  int B, H, Nq, Nk, dqk, dv;
  int fp16;               // 16-bit operand / output format: 0 = bf16, 1 = fp16
  int q_bcast;
  float scale_log2;
  const uint8_t* key_mask; long long stride_km;
  const uint8_t* row_keep; long long stride_rk;
  __nv_bfloat16* O; long long ldo, strideO;
  int q_pairs;    // ceil(Nq / 256)
  int items;      // B * H * q_pairs
  int kv_tiles;   // ceil(Nk / BN)
  int staged;     // epilogue through shared memory + TMA store (see the epilogue)
};

template <int NQC, int NVC, int BN>
struct Flash2Cfg {
  static constexpr int Q_TILE_BYTES = NQC * 16384;            // 128 rows x 64-column chunks
  static constexpr int Q_SLOTS = 3;                           // tile t of this CTA (t = 2 * item + x) lives in slot t % 3
  static constexpr int Q_BYTES = Q_SLOTS * Q_TILE_BYTES;
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
                  const __grid_constant__ CUtensorMap tmap_v, const __grid_constant__ CUtensorMap tmap_o,
                  const Flash2Params p) {
  using Cfg = Flash2Cfg<NQC, NVC, BN>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem;                                   // [3 slots][NQC chunks][128 x 128 B]
  uint8_t* sK = sQ + Cfg::Q_BYTES;                      // [STAGES][NQC][BN x 128 B]
  uint8_t* sV = sK + STAGES * Cfg::K_BYTES;             // [STAGES][NVC][BN x 128 B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + STAGES * Cfg::V_BYTES);
  uint64_t* q_full = bars;                              // [3]  TMA -> MMA, one per Q slot
  uint64_t* q_empty = bars + 3;                         // [3]  MMA (last S of the tile retired) -> TMA
  uint64_t* k_full = bars + 6;                          // [STAGES]
  uint64_t* k_empty = k_full + STAGES;
  uint64_t* v_full = k_empty + STAGES;
  uint64_t* v_empty = v_full + STAGES;
  uint64_t* s_full = v_empty + STAGES;                  // [2]  S_X ready                  (MMA -> softmax X)
  uint64_t* p_full = s_full + 2;                        // [2 tiles][2 halves]  P_X (one half of the key tile) in TMEM, O_X
                                                        //      rescaled  (softmax X -> MMA), 128 arrivals each
  uint64_t* pv_done = p_full + 4;                       // [2]  O_X += P_X V retired        (MMA -> softmax X)
  uint64_t* o_ready = pv_done + 2;                      // [2]  staged output tile X is in shared memory (softmax X -> warp 3)
  uint64_t* epi_done = o_ready + 2;                     // [2]  ... and has left it again          (warp 3 -> TMA producer)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(epi_done + 2);

  // the shuffle makes the warp index provably warp-uniform, so role code can use the uniform datapath
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
#ifdef PIO_FLASH2_TRACE
  int f2n = 0;
#endif
  if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) {
    printf("pio_flash2_kernel: dynamic shared memory base is not 1024-byte aligned\n");
    __trap();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_v);
    tma_prefetch_desc(&tmap_o);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < Cfg::Q_SLOTS; ++s) {
      mbar_init(&q_full[s], 1);
      mbar_init(&q_empty[s], 1);
    }
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&k_empty[s], 1);
      mbar_init(&v_full[s], 1);
      mbar_init(&v_empty[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[2 * i], 4 * kArrivalsPerWarp);        // the four softmax warps of tile i
      mbar_init(&p_full[2 * i + 1], 4 * kArrivalsPerWarp);
      mbar_init(&pv_done[i], 1);
      mbar_init(&o_ready[i], 4 * kArrivalsPerWarp);
      mbar_init(&epi_done[i], 1);
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
      // Q tile t = 2 * it + x of this CTA goes to slot t % 3: tile A of the NEXT item is prefetched a whole item ahead
      // and tile B as soon as the slot of the previous item's tile A is released, so an item never starts with a Q wait.
      // Program order follows the order in which the MMA warp releases things (S_B(T-2), S_A(T-1), PV_B(T-2),
      // S_B(T-1), PV_B(T-1) of the previous item), so no wait here delays a load that could already go.
      auto load_q = [&](int item, int it_, int x) {
        const int qp = item % p.q_pairs;
        const int bh = item / p.q_pairs;
        const int h = bh % p.H, b = bh / p.H;
        const uint32_t t = 2u * (uint32_t)it_ + (uint32_t)x;
        const int slot = (int)(t % 3u);
        F2T(440 + x);
        mbar_wait(&q_empty[slot], ((t / 3u) & 1u) ^ 1u);
        F2T(450 + x);
        if (leader) {
          mbar_arrive_expect_tx(&q_full[slot], Cfg::Q_TILE_BYTES);
#pragma unroll
          for (int c = 0; c < NQC; ++c)
            tma_load_3d(sQ + (slot * NQC + c) * 16384, &tmap_q, &q_full[slot], h * p.dqk + c * 64, qp * 256 + x * 128,
                        p.q_bcast ? 0 : b);
        }
      };
      uint32_t kv = 0;
      int it = 0;
      if ((int)blockIdx.x < p.items) load_q(p.items - 1 - (int)blockIdx.x, 0, 0);
      for (int seq = blockIdx.x; seq < p.items; seq += gridDim.x, ++it) {
        const int item = p.items - 1 - seq;   // items are walked back to front, see the item decomposition below
        const int qp = item % p.q_pairs;
        const int bh = item / p.q_pairs;
        const int h = bh % p.H, b = bh / p.H;
        (void)qp;
        const int jq = T > 1 ? 1 : 0;   // the next item's tile A is requested after this key tile
        for (int j = 0; j < T; ++j, ++kv) {
          const int stage = kv % STAGES;
          const uint32_t ph = (kv / STAGES) & 1u;
          F2T(400 + j);
          // staged epilogue: tile A of the previous item parked its output in the K stage of that item's last key tile
          // (= this stage when j == STAGES - 1), tile B in the V stage
          const bool after_epilogue = p.staged && it > 0 && j == STAGES - 1;
          if (after_epilogue) mbar_wait(&epi_done[0], (uint32_t)(it - 1) & 1u);
          mbar_wait(&k_empty[stage], ph ^ 1u);
          F2T(410 + j);
          if (leader) {
            mbar_arrive_expect_tx(&k_full[stage], Cfg::K_BYTES);
#pragma unroll
            for (int c = 0; c < NQC; ++c)
              tma_load_3d(sK + (stage * NQC + c) * Cfg::CHUNK_BYTES, &tmap_k, &k_full[stage], h * p.dqk + c * 64, j * BN, b);
          }
          if (j == 0) load_q(item, it, 1);
          F2T(420 + j);
          if (after_epilogue) mbar_wait(&epi_done[1], (uint32_t)(it - 1) & 1u);
          mbar_wait(&v_empty[stage], ph ^ 1u);
          F2T(430 + j);
          if (leader) {
            mbar_arrive_expect_tx(&v_full[stage], Cfg::V_BYTES);
#pragma unroll
            for (int c = 0; c < NVC; ++c)
              tma_load_3d(sV + (stage * NVC + c) * Cfg::CHUNK_BYTES, &tmap_v, &v_full[stage], h * p.dv + c * 64, j * BN, b);
          }
          if (j == jq && seq + (int)gridDim.x < p.items) load_q(item - (int)gridDim.x, it + 1, 0);
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
      const uint32_t idesc_s = make_idesc_f16(128, BN, idesc_fmt(p.fp16), 0, 0);
      const uint32_t idesc_pv = make_idesc_f16(128, dv_n, idesc_fmt(p.fp16), /*A (TMEM) K-major*/ 0, /*B MN-major*/ 1);
      const uint64_t dq0 = make_smem_desc_sw128(smem_u32(sQ), 16, 1024);
      const uint64_t dk0 = make_smem_desc_sw128(smem_u32(sK), 16, 1024);
      const uint64_t dv0 = make_smem_desc_sw128(smem_u32(sV), Cfg::CHUNK_BYTES, 1024);
      const uint32_t q_lo = (uint32_t)dq0, q_hi = (uint32_t)(dq0 >> 32);
      const uint32_t k_lo = (uint32_t)dk0, k_hi = (uint32_t)(dk0 >> 32);
      const uint32_t v_lo = (uint32_t)dv0, v_hi = (uint32_t)(dv0 >> 32);
      auto issue_s = [&](int x, int stage, int qslot) {
        const uint32_t d = tmem_base + x * BN;
        const uint32_t a0 = q_lo + (uint32_t)((qslot * Cfg::Q_TILE_BYTES) >> 4);
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
      // P_x reaches the tensor pipe in two halves of BN / 2 keys: the PV MMAs of the first half run while the softmax
      // warps still exponentiate the second half (the chain S -> softmax -> PV -> S' of a tile is what bounds a step)
      auto issue_pv = [&](int x, int stage, bool accum, int half) {
        const uint32_t a = tmem_base + x * BN;          // P_x overlays the first BN/2 columns of S_x
        const uint32_t d = tmem_o + x * (NVC * 64);
        const uint32_t b0 = v_lo + (uint32_t)((stage * Cfg::V_BYTES) >> 4);
#pragma unroll
        for (int k2 = 0; k2 < BN / 32; ++k2) {
          const int ks = half * (BN / 32) + k2;
          if (elect_one())
            umma_ts_lh(d, a + ks * 8, b0 + (uint32_t)((ks * 2048) >> 4), v_hi, idesc_pv, (accum || ks != 0) ? 1u : 0u);
        }
      };
      // The CTA's items form ONE stream of key-tile steps g = it * T + j: S_x(g+1) is issued right behind PV_x(g) also
      // across an item boundary (the next item's Q tiles and first K tile are prefetched), so the tensor pipe never
      // drains between items; only the epilogue of a tile sits between its last PV and its next softmax.
      const int my_items = ((int)blockIdx.x < p.items) ? (p.items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
      const uint32_t G = (uint32_t)my_items * (uint32_t)T;
      // S_x of step g: waits for what it reads (K tile always; the tile's Q slot on the first key tile of an item)
      auto s_step = [&](int x, uint32_t g) {
        const uint32_t it = g / (uint32_t)T, j = g - it * (uint32_t)T;
        const int stage = (int)(g % STAGES);
        const uint32_t t = 2u * it + (uint32_t)x;
        const int qslot = (int)(t % 3u);
        if (x == 0) mbar_wait(&k_full[stage], (g / STAGES) & 1u);
        if (j == 0) mbar_wait(&q_full[qslot], (t / 3u) & 1u);
        tc_fence_after();
        issue_s(x, stage, qslot);
        if (leader) umma_commit(&s_full[x]);
        if ((int)j + 1 == T && leader) umma_commit(&q_empty[qslot]);   // the tile's last S: its Q slot is free
        if (x == 1 && leader) umma_commit(&k_empty[stage]);
      };
      if (G > 0) {
        F2T(100);
        s_step(0, 0);
        s_step(1, 0);
        F2T(103);
      }
      for (uint32_t g = 0; g < G; ++g) {
        const int stage = (int)(g % STAGES);
        const int j = (int)(g % (uint32_t)T);
        mbar_wait(&v_full[stage], (g / STAGES) & 1u);
        F2T(110 + j);
#pragma unroll
        for (int x = 0; x < 2; ++x) {
          mbar_wait(&p_full[2 * x], g & 1u);
          F2T(120 + 10 * x + j);
          tc_fence_after();
          issue_pv(x, stage, j > 0, 0);
          mbar_wait(&p_full[2 * x + 1], g & 1u);
          tc_fence_after();
          issue_pv(x, stage, j > 0, 1);
          if (leader) umma_commit(&pv_done[x]);
          if (x == 1 && leader) umma_commit(&v_empty[stage]);
          if (g + 1 < G) {
            s_step(x, g + 1);   // in-order after PV_x(g): P_x(g) has been consumed before S_x is overwritten
            F2T(140 + 10 * x + j);
          }
        }
      }
    }
  }
  else if (warp == 3 && p.staged) {
    // ================= output store (staged epilogue) =================
    const bool leader = (lane == 0);
    uint32_t g = 0;
    int it = 0;
    for (int seq = blockIdx.x; seq < p.items; seq += gridDim.x, ++it) {
      const int item = p.items - 1 - seq;
      const int qp = item % p.q_pairs;
      const int bh = item / p.q_pairs;
      const int h = bh % p.H, b = bh / p.H;
      g += (uint32_t)T;
      const uint32_t last_stage = (g - 1u) % STAGES;
#pragma unroll
      for (int x = 0; x < 2; ++x) {
        mbar_wait(&o_ready[x], (uint32_t)it & 1u);
        if (leader) {
          const uint8_t* stg = (x == 0) ? sK + last_stage * Cfg::K_BYTES : sV + last_stage * Cfg::V_BYTES;
#pragma unroll
          for (int c = 0; c < NVC; ++c)
            tma_store_3d(&tmap_o, stg + c * 16384, h * p.dv + c * 64, qp * 256 + x * 128, b);   // clips rows >= Nq
          bulk_commit();
          bulk_wait_read<0>();
          mbar_arrive(&epi_done[x]);
        }
        __syncwarp();
      }
    }
    if (leader) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
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
    for (int seq = blockIdx.x; seq < p.items; seq += gridDim.x) {
      // Items are walked from the END of the (batch, head, query block) order: Q/K/V were just written front to back by
      // the QKV projection, so the rows of the last batch entries are what is still resident in L2 when every CTA asks
      // for its first tiles at once (the first two items of a CTA took 20 000 cycles instead of 15 000), and the
      // output written last — the first batch entries — is what the out-projection asks for first.
      const int item = p.items - 1 - seq;
      const int qp = item % p.q_pairs;
      const int bh = item / p.q_pairs;
      const int h = bh % p.H, b = bh / p.H;
      const int q = qp * 256 + x * 128 + row;
      const uint8_t* km = p.key_mask ? p.key_mask + (long long)b * p.stride_km : nullptr;
      float m = -INFINITY;  // running max of scale_log2 * s
      float l = 0.f;        // running sum of exp2(t - m)
      for (int j = 0; j < T; ++j, ++g) {
        const int k0 = j * BN;
        if (quarter == 0) F2T(200 + 100 * x + 10 * j);
        mbar_wait(&s_full[x], g & 1u);
        if (quarter == 0) F2T(201 + 100 * x + 10 * j);
        tc_fence_after();
        const bool tail = (k0 + BN > p.Nk) || (km != nullptr);
        // ---- the whole S row of this tile (BN fp32 values) moves to registers with one wait ----
        uint32_t r[BN];
#pragma unroll
        for (int c = 0; c < BN / 32; ++c) tmem_ld32(t_s + c * 32, *reinterpret_cast<uint32_t(*)[32]>(&r[c * 32]));
        tmem_wait_ld();
        if (quarter == 0) F2T(202 + 100 * x + 10 * j);
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
        if (quarter == 0) F2T(203 + 100 * x + 10 * j);
        const float msub = (m == -INFINITY) ? 0.0f : m;
        // ---- p = exp2(scale * s - m) -> bf16 pairs written over S in TMEM (two values per 32-bit column); the row sum
        //      uses the bf16-rounded values the tensor core will multiply, so P and l stay consistent ----
        const uint64_t sc2 = pack_f32x2(p.scale_log2, p.scale_log2);
        const uint64_t nm2 = pack_f32x2(-msub, -msub);
        uint64_t la = pack_f32x2(0.f, 0.f), lb = pack_f32x2(0.f, 0.f);
#pragma unroll
        for (int c = 0; c < BN / 32; ++c) {
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
              // the row sum takes the un-rounded exponentials: round-to-nearest is unbiased, so l differs from the sum of
              // the 16-bit values the tensor core multiplies by ~2^-9 / sqrt(Nk) relative, and rebuilding the rounded
              // values cost two integer instructions per pair in a loop that is issue-bound (one warp per scheduler)
              const uint64_t pr = pack_f32x2(e0, e1);
              if (i & 1) lb = fadd2(lb, pr);
              else la = fadd2(la, pr);
            }
          };
          if (p.fp16) exp_block(std::true_type{});
          else exp_block(std::false_type{});
          tmem_st16(t_s + c * 16, w);
          if (c == BN / 64 - 1) {   // first half of the key tile is in TMEM
            tmem_wait_st();
            tc_fence_before();
            mbar_arrive_warp(&p_full[2 * x]);
          }
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
        mbar_arrive_warp(&p_full[2 * x + 1]);
        if (quarter == 0) F2T(204 + 100 * x + 10 * j);
      }
      // ---- epilogue of the item: O_x / l -> bf16 ----
      mbar_wait(&pv_done[x], (g - 1u) & 1u);
      if (quarter == 0) F2T(250 + 100 * x);
      tc_fence_after();
      const bool keep = (q < p.Nq) && (p.row_keep == nullptr || p.row_keep[(long long)b * p.stride_rk + q] != 0);
      const float inv = (keep && l > 0.f) ? 1.0f / l : 0.0f;
      __nv_bfloat16* orow = p.O + (long long)b * p.strideO + (long long)q * p.ldo + (long long)h * p.dv;
      if (p.staged) {
        // Every lane owns a different output row (rows are ldo apart), so direct stores cost one pass of the store
        // path per lane and instruction (~2200 cycles per tile, and they delay the MMA warp's barrier traffic).
        // Instead the tile is parked in shared memory that is idle right now — tile A: the K stage of the item's last
        // key tile (S_B of that tile retired before PV_A did), tile B: its V stage (PV_B retired) — in the 128-byte
        // swizzled layout, and one thread hands it to TMA; the producer refills those stages only after epi_done.
        const uint32_t last_stage = (g - 1u) % STAGES;
        uint8_t* stg = (x == 0) ? sK + last_stage * Cfg::K_BYTES : sV + last_stage * Cfg::V_BYTES;
#pragma unroll 1
        for (int c = 0; c < NVC * 64; c += 32) {
          uint32_t r[32];
          tmem_ld32(t_o + c, r);
          tmem_wait_ld();
          uint8_t* chunk = stg + (c >> 6) * 16384;
#pragma unroll
          for (int gq = 0; gq < 4; ++gq) {
            uint4 w;
            w.x = pack16x2(__uint_as_float(r[8 * gq]) * inv, __uint_as_float(r[8 * gq + 1]) * inv, p.fp16);
            w.y = pack16x2(__uint_as_float(r[8 * gq + 2]) * inv, __uint_as_float(r[8 * gq + 3]) * inv, p.fp16);
            w.z = pack16x2(__uint_as_float(r[8 * gq + 4]) * inv, __uint_as_float(r[8 * gq + 5]) * inv, p.fp16);
            w.w = pack16x2(__uint_as_float(r[8 * gq + 6]) * inv, __uint_as_float(r[8 * gq + 7]) * inv, p.fp16);
            *reinterpret_cast<uint4*>(chunk + sw128_offset((uint32_t)row, (uint32_t)(((c & 63) >> 3) + gq))) = w;
          }
        }
        fence_proxy_async_smem();
        mbar_arrive_warp(&o_ready[x]);   // warp 3 hands the tile to TMA; this thread goes straight on to the next item
        if (quarter == 0) F2T(251 + 100 * x);
        continue;
      }
      for (int c = 0; c < dv_n; c += 32) {
        uint32_t r[32];
        tmem_ld32(t_o + c, r);
        tmem_wait_ld();
        if (q < p.Nq) {
          __nv_bfloat16* op = orow + c;
          if (c + 32 <= p.dv && ((reinterpret_cast<uintptr_t>(op) & 31u) == 0)) {
            // 32-byte stores: every lane writes its own row (rows are ldo apart), so the store path handles one lane
            // per pass — halving the instruction count halves the epilogue
#pragma unroll
            for (int gq = 0; gq < 2; ++gq) {
              uint32_t w[8];
#pragma unroll
              for (int i = 0; i < 8; ++i)
                w[i] = pack16x2(__uint_as_float(r[16 * gq + 2 * i]) * inv, __uint_as_float(r[16 * gq + 2 * i + 1]) * inv, p.fp16);
              asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(op + 16 * gq), "r"(w[0]),
                           "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7])
                           : "memory");
            }
          } else if (c + 32 <= p.dv && ((reinterpret_cast<uintptr_t>(op) & 15u) == 0)) {
#pragma unroll
            for (int gq = 0; gq < 4; ++gq) {
              uint4 w;
              w.x = pack16x2(__uint_as_float(r[8 * gq]) * inv, __uint_as_float(r[8 * gq + 1]) * inv, p.fp16);
              w.y = pack16x2(__uint_as_float(r[8 * gq + 2]) * inv, __uint_as_float(r[8 * gq + 3]) * inv, p.fp16);
              w.z = pack16x2(__uint_as_float(r[8 * gq + 4]) * inv, __uint_as_float(r[8 * gq + 5]) * inv, p.fp16);
              w.w = pack16x2(__uint_as_float(r[8 * gq + 6]) * inv, __uint_as_float(r[8 * gq + 7]) * inv, p.fp16);
              reinterpret_cast<uint4*>(op)[gq] = w;
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (c + i < p.dv) reinterpret_cast<uint16_t*>(op)[i] = cvt16(__uint_as_float(r[i]) * inv, p.fp16);
          }
        }
      }
      if (quarter == 0) F2T(251 + 100 * x);
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
  // staged epilogue: whole 64-column chunks per head, an output the tensor map can address, O tile no larger than a
  // K stage, and at least STAGES key tiles per item (so the stage a tile parks its output in is refilled by a known
  // load of the next item)
  const int kv_tiles = (a->Nk + BN - 1) / BN;
  const bool staged = BN == 128 && NVC <= NQC && a->dv == NVC * 64 && kv_tiles >= Cfg::STAGES && aligned16(a->O) &&
                      a->ldo % 8 == 0 && (a->B == 1 || a->strideO % 8 == 0);
  CUtensorMap to = tq;
  if (staged) {
    const uint64_t dims[3] = {(uint64_t)a->H * a->dv, (uint64_t)a->Nq, (uint64_t)a->B};
    const uint64_t strides[2] = {(uint64_t)a->ldo * 2, (uint64_t)(a->B == 1 ? a->ldo * (int64_t)a->Nq : a->strideO) * 2};
    const uint32_t box[3] = {64, 128, 1};
    int rc = encode_tmap_bf16(&to, a->O, 3, dims, strides, box);
    if (rc != PIO_OK) return rc;
  }
  Flash2Params p;
  p.staged = staged ? 1 : 0;
  p.B = a->B; p.H = a->H; p.Nq = a->Nq; p.Nk = a->Nk; p.dqk = a->dqk; p.dv = a->dv;
  p.fp16 = a->fp16 ? 1 : 0;
  p.q_bcast = q_bcast;
  p.scale_log2 = a->scale * 1.4426950408889634f;
  p.key_mask = a->key_mask; p.stride_km = a->stride_km;
  p.row_keep = a->row_keep; p.stride_rk = a->stride_rk;
  p.O = reinterpret_cast<__nv_bfloat16*>(a->O); p.ldo = a->ldo; p.strideO = a->strideO;
  p.q_pairs = (a->Nq + 255) / 256;
  p.items = a->B * a->H * p.q_pairs;
  p.kv_tiles = (a->Nk + BN - 1) / BN;

  static PerDeviceOnce once;
  const cudaError_t attr_err = once.run(dev.device, [] {
    return cudaFuncSetAttribute(pio_flash2_kernel<NQC, NVC, BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                Cfg::SMEM_BYTES);
  });
  if (attr_err != cudaSuccess)
    return fail(PIO_ERR_CUDA, "cudaFuncSetAttribute(flash2<%d,%d,%d>) failed: %s", NQC, NVC, BN,
                cudaGetErrorString(attr_err));
  const int grid = p.items < dev.sm_count ? p.items : dev.sm_count;
  {
    ProfileScope prof(KF_FLASH, 2.0 * a->B * a->H * (double)a->Nq * a->Nk * (a->dqk + a->dv), 0.0, stream);
    PIO_CUDA_OK(launch_kernel(pio_flash2_kernel<NQC, NVC, BN>, dim3((unsigned)grid, 1, 1), dim3(384, 1, 1), Cfg::SMEM_BYTES,
                              stream, 1, tq, tk, tv, to, p));
  }
#ifdef PIO_FLASH2_TRACE
  {
    static bool dumped = false;
    if (!dumped) {
      dumped = true;
      cudaDeviceSynchronize();
      static unsigned long long host[12 * 2 * 512];
      cudaMemcpyFromSymbol(host, g_f2_trace, sizeof(host));
      unsigned long long t0 = ~0ull;
      for (unsigned i = 0; i < 12 * 512; ++i)
        if (host[2 * i] != 0 && host[2 * i + 1] < t0) t0 = host[2 * i + 1];
      for (unsigned i = 0; i < 12 * 512; ++i)
        if (host[2 * i] != 0) fprintf(stderr, "F2T %llu %llu %u\n", host[2 * i], host[2 * i + 1] - t0, i / 512);
    }
  }
#endif
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
