// Query-tiled decoder attention for sm_100a: very many output queries attend over a few thousand latent keys
// (PerceiverDecoder, perceiver.py:166-180 -> CrossAttention -> Attention.attend, transformer_primitives.py:117-180;
// optical flow: 182,528 queries x 2048 latents).
//
// The latent-side operands K and V are DIFFERENT matrices here (the host folds the query projection into K and the
// output projection into V, so both are Nk x ~330 wide) and every 128-query tile needs all of them: with one CTA per
// tile, Q (96 KB) plus one 64-key K/V stage (96 KB) already fill shared memory, and the 1426 tiles would pull 3.9 GB
// through L2.  So the kernel runs on CTA PAIRS (cluster of 2, tcgen05.mma.cta_group::2, M = 256): each CTA owns 128
// queries (its own Q tile, its own S / P / O in its own TMEM, its own softmax warps) but stages only HALF of every
// latent tile — 32 of the 64 keys of K (the N half of S = Q.K^T) and half of V's columns (the N half of O += P.V) — and
// the leader CTA issues one MMA for both.  Per CTA: Q + 3 K stages (24 KB) + 2 V stages (24 KB) = 216 KB; L2 traffic and
// the number of issued MMAs per SM are halved.
//
// As in pio_flash.cu: S accumulates in a double-buffered TMEM tile, the softmax warps (thread = query row x half of the
// tile's keys) keep running max / sum in fp32 in the log2 domain, overwrite S in place with 16-bit P (tcgen05.st), and
// O += P.V runs with the A operand in TMEM; O is rescaled lazily.  The epilogue writes the fp32 block output
// out = O / l + bias (+ residual), i.e. the attention output with the (folded) output projection already applied.
//
// Roles (384 threads per CTA): warp 0 TMA producer for Q and K, warp 3 TMA producer for V (both CTAs), warp 1 MMA
// issuer (leader CTA only), warp 2 TMEM allocator, warps 4..11 softmax / correction / epilogue.
#include <math.h>
#include <stdlib.h>

#include <type_traits>

#include "pio_common.cuh"
#include "pio_host.h"

namespace pio {

// Developer aid (compiled out unless -DPIO_DECODE_TRACE): CTAs 0 and 1 record (tag, clock64) pairs at the pipeline's
// hand-off points; the first launch prints them to stderr (tools/trace_decode.py).  Tags: 1000 + 10 j + e MMA issuer,
// 2000 + 10 j + e softmax warp 4 of the leader, 3000 + ... of its peer (clock64 is per SM: compare within one CTA only).
#ifdef PIO_DECODE_TRACE
__device__ unsigned long long g_dec_trace[3 * 2 * 512];
#define DT(slot, tag)                                                                     \
  do {                                                                                    \
    if (blockIdx.x < 2 && blockIdx.y == 0 && lane == 0 && dtn < 512) {                    \
      g_dec_trace[((slot) * 512 + dtn) * 2] = (unsigned long long)(tag);                  \
      g_dec_trace[((slot) * 512 + dtn) * 2 + 1] = (unsigned long long)clock64();          \
      ++dtn;                                                                              \
    }                                                                                     \
  } while (0)
#else
#define DT(slot, tag)
#endif

struct DecodeParams {
  int B, Nq, Nk, dqk, dv;
  int fp16;
  int nqc;                 // 64-column chunks of the contraction (dqk)
  int nv1, nv2;            // N of the two P.V MMAs: nv1 = min(256, dv_pad), nv2 = dv_pad - nv1 (0 or a multiple of 32)
  int vch1, vch2;          // 64-column V chunks a CTA stages for them: ceil(nv1 / 2 / 64), nv2 > 0
  int kst, vst;            // ring depths
  int q_bcast;
  float scale_log2;
  const uint8_t* key_mask; long long stride_km;
  const uint8_t* row_keep; long long stride_rk;
  const float* bias;
  const float* residual; long long ldr, strideR;
  float* out; long long ldo, strideO;
  uint16_t* out_ln; long long ld_ln, stride_ln;     // fused LayerNorm of the output rows (16-bit), or nullptr
  const float* ln_gamma; const float* ln_beta; float ln_eps;
};

constexpr int DEC_BN = 64;             // keys per tile
constexpr int DEC_KCHUNK = 32 * 128;   // one 64-column chunk of this CTA's 32 keys of a K tile
constexpr int DEC_VCHUNK = 64 * 128;   // one 64-column chunk of this CTA's V columns, 64 keys
constexpr int DEC_BAR_BYTES = 512 + 3072;   // mbarriers + three [2 halves][128 rows] fp32 exchange slots

// D[tmem] (+)= A[tmem] * B[smem] on a CTA pair (A: 16-bit pairs packed in 32-bit TMEM columns of either CTA)
__device__ __forceinline__ void umma_ts_2cta_lh(uint32_t tmem_d, uint32_t tmem_a, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                                uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 db;\n\tsetp.ne.b32 p, %5, 0;\n\t"
      "mov.b64 db, {%2, %3};\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], db, %4, p;\n\t}\n"
      ::"r"(tmem_d), "r"(tmem_a), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster_release(uint32_t bar_cluster) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster) : "memory");
}

__global__ void __launch_bounds__(384, 1)
pio_decode_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                  const __grid_constant__ CUtensorMap tmap_v, const DecodeParams p) {
  constexpr int BN = DEC_BN;
  extern __shared__ __align__(1024) uint8_t smem[];
  const int q_bytes = p.nqc * 16384;
  const int k_stage = p.nqc * DEC_KCHUNK;
  const int v_stage = (p.vch1 + p.vch2) * DEC_VCHUNK;
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + q_bytes;
  uint8_t* sV = sK + p.kst * k_stage;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + p.vst * v_stage);
  uint64_t* q_full = bars;                 // [1]    leader's: both CTAs' Q tiles have landed
  uint64_t* k_full = bars + 1;             // [4]    leader's: both halves of a K tile
  uint64_t* k_empty = k_full + 4;          // [4]    per CTA (multicast commit)
  uint64_t* v_full = k_empty + 4;          // [4]    leader's
  uint64_t* v_empty = v_full + 4;          // [4]    per CTA (multicast commit)
  uint64_t* s_full = v_empty + 4;          // [2]    per CTA (multicast commit): S_j is in this CTA's TMEM
  uint64_t* p_full = s_full + 2;           // [2]    leader's: 16 arrivals, one per softmax warp of either CTA
  uint64_t* pv_done = p_full + 2;          // [2]    per CTA (multicast commit)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + 2);
  float* xchg = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 512);   // [2 slots][2 halves][128 rows]

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
#ifdef PIO_DECODE_TRACE
  int dtn = 0;
#endif
  if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) {
    printf("pio_decode_kernel: dynamic shared memory base is not 1024-byte aligned\n");
    __trap();
  }
  const uint32_t crank = cluster_ctarank();          // 0 = leader
  const int q0 = (int)blockIdx.x * 128;              // blockIdx.x = 2 * pair + crank
  const int b = blockIdx.y;
  const int ntiles = (p.Nk + BN - 1) / BN;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_v);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(q_full, 1);
    for (int s = 0; s < 4; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&k_empty[s], 1);
      mbar_init(&v_full[s], 1);
      mbar_init(&v_empty[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 16);
      mbar_init(&pv_done[i], 1);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc_2cta(tmem_slot, 512);
    tmem_relinquish_2cta();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // the peer's barriers are initialised before anything is signalled across the pair
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_o = tmem_base + 2 * BN;
  pdl_sync();

  const int dqk_steps = (p.dqk + 15) / 16;
  const int dv_n = p.nv1 + p.nv2;

  if (warp == 0) {
    // ================= TMA producer: Q (once) and this CTA's 32 keys of every K tile =================
    const int bq = p.q_bcast ? 0 : b;
    if (elect_one()) {
      // both CTAs' bytes are credited to the leader's barrier, which expects the pair's total
      if (crank == 0) mbar_arrive_expect_tx(q_full, 2u * (uint32_t)q_bytes);
      const uint32_t lb = mapa_u32(q_full, 0);
      for (int c = 0; c < p.nqc; ++c) tma_load_3d_2cta(sQ + c * 16384, &tmap_q, lb, c * 64, q0, bq);
    }
    int stage = 0;
    uint32_t phase = 0;
    for (int j = 0; j < ntiles; ++j) {
      mbar_wait(&k_empty[stage], phase ^ 1u);
      if (elect_one()) {
        if (crank == 0) mbar_arrive_expect_tx(&k_full[stage], 2u * (uint32_t)k_stage);
        const uint32_t lb = mapa_u32(&k_full[stage], 0);
        uint8_t* st = sK + stage * k_stage;
        const int k0 = j * BN + (int)crank * 32;
        for (int c = 0; c < p.nqc; ++c) tma_load_3d_2cta(st + c * DEC_KCHUNK, &tmap_k, lb, c * 64, k0, b);
      }
      if (++stage == p.kst) { stage = 0; phase ^= 1u; }
    }
  } else if (warp == 3) {
    // ================= TMA producer: this CTA's column half of every V tile =================
    int stage = 0;
    uint32_t phase = 0;
    const int col1 = (int)crank * (p.nv1 / 2);               // first P.V MMA: columns [crank * nv1/2, + nv1/2)
    const int col2 = p.nv1 + (int)crank * (p.nv2 / 2);       // second: [nv1 + crank * nv2/2, + nv2/2)
    for (int j = 0; j < ntiles; ++j) {
      mbar_wait(&v_empty[stage], phase ^ 1u);
      if (elect_one()) {
        if (crank == 0) mbar_arrive_expect_tx(&v_full[stage], 2u * (uint32_t)v_stage);
        const uint32_t lb = mapa_u32(&v_full[stage], 0);
        uint8_t* st = sV + stage * v_stage;
        for (int c = 0; c < p.vch1; ++c) tma_load_3d_2cta(st + c * DEC_VCHUNK, &tmap_v, lb, col1 + c * 64, j * BN, b);
        if (p.vch2) tma_load_3d_2cta(st + p.vch1 * DEC_VCHUNK, &tmap_v, lb, col2, j * BN, b);
      }
      if (++stage == p.vst) { stage = 0; phase ^= 1u; }
    }
  } else if (warp == 1) {
    if (crank == 0) {
      // ================= MMA issuer (leader CTA) =================
      const uint32_t idesc_s = make_idesc_f16(256, BN, idesc_fmt(p.fp16), 0, 0);
      const uint32_t idesc_pv1 = make_idesc_f16(256, p.nv1, idesc_fmt(p.fp16), /*A (TMEM) K-major*/ 0, /*B MN-major*/ 1);
      const uint32_t idesc_pv2 = make_idesc_f16(256, p.nv2 > 0 ? p.nv2 : 32, idesc_fmt(p.fp16), 0, 1);
      const uint64_t dq0 = make_smem_desc_sw128(smem_u32(sQ), 16, 1024);
      const uint64_t dk0 = make_smem_desc_sw128(smem_u32(sK), 16, 1024);
      const uint64_t dv0 = make_smem_desc_sw128(smem_u32(sV), DEC_VCHUNK, 1024);
      const uint32_t q_lo = (uint32_t)dq0, q_hi = (uint32_t)(dq0 >> 32);
      const uint32_t k_lo = (uint32_t)dk0, k_hi = (uint32_t)(dk0 >> 32);
      const uint32_t v_lo = (uint32_t)dv0, v_hi = (uint32_t)(dv0 >> 32);
      int ks_stage = 0, vs_stage = 0;
      uint32_t k_phase = 0, v_phase = 0;
      auto issue_s = [&](int j) {
        mbar_wait(&k_full[ks_stage], k_phase);
        tc_fence_after();
        DT(0, 1000 + 10 * j + 0);
        const uint32_t d = tmem_base + (uint32_t)((j & 1) * BN);
        const uint32_t b0 = k_lo + (uint32_t)((ks_stage * k_stage) >> 4);
        for (int ks = 0; ks < dqk_steps; ++ks) {
          const int c = ks >> 2, kk = ks & 3;
          if (elect_one())
            umma_ss_2cta_lh(d, q_lo + (uint32_t)((c * 16384 + kk * 32) >> 4), q_hi,
                            b0 + (uint32_t)((c * DEC_KCHUNK + kk * 32) >> 4), k_hi, idesc_s, ks != 0 ? 1u : 0u);
        }
        if (elect_one()) {
          umma_commit_2cta_mcast(&k_empty[ks_stage], 0x3);   // this K stage is free in both CTAs
          umma_commit_2cta_mcast(&s_full[j & 1], 0x3);       // S_j is complete in both CTAs' TMEM
        }
        DT(0, 1000 + 10 * j + 1);
        if (++ks_stage == p.kst) { ks_stage = 0; k_phase ^= 1u; }
      };
      mbar_wait(q_full, 0);
      tc_fence_after();
      issue_s(0);
      for (int j = 0; j < ntiles; ++j) {
        if (j + 1 < ntiles) issue_s(j + 1);   // overwrites S_{j-1} / P_{j-1}: in order after PV_{j-1}, which consumed P_{j-1}
        mbar_wait(&p_full[j & 1], (j >> 1) & 1);
        DT(0, 1000 + 10 * j + 2);
        mbar_wait(&v_full[vs_stage], v_phase);
        tc_fence_after();
        DT(0, 1000 + 10 * j + 3);
        const uint32_t a0 = tmem_base + (uint32_t)((j & 1) * BN);   // P_j overlays the first BN/2 columns of S_j
        const uint32_t b0 = v_lo + (uint32_t)((vs_stage * v_stage) >> 4);
#pragma unroll
        for (int ks = 0; ks < BN / 16; ++ks) {
          if (elect_one())
            umma_ts_2cta_lh(tmem_o, a0 + ks * 8, b0 + (uint32_t)((ks * 2048) >> 4), v_hi, idesc_pv1, (j | ks) != 0 ? 1u : 0u);
        }
        if (p.nv2 > 0) {
          const uint32_t b1 = b0 + (uint32_t)((p.vch1 * DEC_VCHUNK) >> 4);
#pragma unroll
          for (int ks = 0; ks < BN / 16; ++ks) {
            if (elect_one())
              umma_ts_2cta_lh(tmem_o + (uint32_t)p.nv1, a0 + ks * 8, b1 + (uint32_t)((ks * 2048) >> 4), v_hi, idesc_pv2,
                              (j | ks) != 0 ? 1u : 0u);
          }
        }
        if (elect_one()) {
          umma_commit_2cta_mcast(&v_empty[vs_stage], 0x3);
          umma_commit_2cta_mcast(&pv_done[j & 1], 0x3);
        }
        DT(0, 1000 + 10 * j + 4);
        if (++vs_stage == p.vst) { vs_stage = 0; v_phase ^= 1u; }
      }
    }
  } else if (warp >= 4) {
    // ================= softmax / correction / epilogue (both CTAs, each on its own 128 queries) =================
    constexpr int HC = BN / 2;            // key columns of a tile handled by this thread
    const int quarter = warp & 3;
    const int half = (warp - 4) >> 2;
    const int row = quarter * 32 + lane;  // row inside the tile == TMEM lane
    const int q = q0 + row;
    const uint32_t lane_off = static_cast<uint32_t>(quarter * 32) << 16;
    const uint8_t* km = p.key_mask ? p.key_mask + (long long)b * p.stride_km : nullptr;
    const uint32_t p_full_leader0 = mapa_u32(&p_full[0], 0);
    const uint32_t p_full_leader1 = mapa_u32(&p_full[1], 0);
    // O columns owned by this half for the rescale and the epilogue (32-column chunks)
    const int nchunks = (dv_n + 31) / 32;
    const int c_begin = half == 0 ? 0 : (nchunks + 1) / 2 * 32;
    const int c_end = half == 0 ? min(dv_n, (nchunks + 1) / 2 * 32) : dv_n;
    auto pair_sync = [&]() { asm volatile("bar.sync %0, 64;" ::"r"(1 + quarter) : "memory"); };
    float m = -INFINITY;  // running max of scale_log2 * s (identical in both halves)
    float l = 0.f;        // running sum of exp2(t - m) over this half's columns
    for (int j = 0; j < ntiles; ++j) {
      const int k0 = j * BN + half * HC;
      const uint32_t t_s = tmem_base + (uint32_t)((j & 1) * BN + half * HC) + lane_off;
      mbar_wait(&s_full[j & 1], (j >> 1) & 1);
      tc_fence_after();
      if (warp == 4) DT(1 + (int)crank, 2000 + 1000 * (int)crank + 10 * j + 0);
      const bool tail = (k0 + HC > p.Nk) || (km != nullptr);
      uint32_t r[HC];
      tmem_ld32(t_s, r);
      tmem_wait_ld();
      if (warp == 4) DT(1 + (int)crank, 2000 + 1000 * (int)crank + 10 * j + 1);
      if (tail) {
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
      {
        float* slot = xchg + (j & 1) * 256;
        slot[half * 128 + row] = tmax;
        pair_sync();
        tmax = fmaxf(tmax, slot[(half ^ 1) * 128 + row]);
      }
      if (warp == 4) DT(1 + (int)crank, 2000 + 1000 * (int)crank + 10 * j + 2);
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
        for (int c = c_begin; c < c_end; c += 16) {
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
      uint32_t w[16];
      auto exp_block = [&](auto f16tag) {
        constexpr bool F16 = decltype(f16tag)::value;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const uint64_t t2 = ffma2(pack_f32x2(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1])), sc2, nm2);
          float t0, t1;
          unpack_f32x2(t2, t0, t1);
          const float e0 = ex2_approx(t0), e1 = ex2_approx(t1);
          w[i] = pack16x2<F16>(e0, e1);
          const uint64_t pr = pack_f32x2(e0, e1);
          if (i & 1) lb = fadd2(lb, pr);
          else la = fadd2(la, pr);
        }
      };
      if (p.fp16) exp_block(std::true_type{});
      else exp_block(std::false_type{});
      // P overwrites S in place (two 16-bit values per 32-bit column); both warps of the pair hold their S values in
      // registers (the pair barrier above), so no unread S column is clobbered
      tmem_st16(tmem_base + (uint32_t)((j & 1) * BN + (half * HC) / 2) + lane_off, w);
      {
        float a0, a1, b0, b1;
        unpack_f32x2(la, a0, a1);
        unpack_f32x2(lb, b0, b1);
        l += (a0 + a1) + (b0 + b1);
      }
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (warp == 4) DT(1 + (int)crank, 2000 + 1000 * (int)crank + 10 * j + 3);
      if (lane == 0) mbar_arrive_cluster_release((j & 1) ? p_full_leader1 : p_full_leader0);
    }
    // ---- epilogue: total row sum = both halves' partial sums ----
    {
      float* slot = xchg + (ntiles & 1) * 256;
      slot[half * 128 + row] = l;
      pair_sync();
      l += slot[(half ^ 1) * 128 + row];
    }
    mbar_wait(&pv_done[(ntiles - 1) & 1], ((ntiles - 1) >> 1) & 1);
    tc_fence_after();
    const bool keep = (q < p.Nq) && (p.row_keep == nullptr || p.row_keep[(long long)b * p.stride_rk + q] != 0);
    const float inv = (keep && l > 0.f) ? 1.0f / l : 0.0f;
    float* orow = p.out + (long long)b * p.strideO + (long long)q * p.ldo;
    const float* rrow = p.residual ? p.residual + (long long)b * p.strideR + (long long)q * p.ldr : nullptr;
    const bool al_out = (reinterpret_cast<uintptr_t>(orow) & 15u) == 0;
    const bool al_res = rrow == nullptr || (reinterpret_cast<uintptr_t>(rrow) & 15u) == 0;
    const bool al_bias = p.bias == nullptr || (reinterpret_cast<uintptr_t>(p.bias) & 15u) == 0;
    // v[0..31] = the block output for columns c .. c+31 of this thread's row (0 beyond dv): O / l + bias (+ residual)
    auto fill_chunk = [&](int c, const uint32_t (&r)[32], float (&v)[32]) {
      if (c + 32 <= p.dv && al_res && al_bias) {
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          float4 t = make_float4(__uint_as_float(r[4 * g]) * inv, __uint_as_float(r[4 * g + 1]) * inv,
                                 __uint_as_float(r[4 * g + 2]) * inv, __uint_as_float(r[4 * g + 3]) * inv);
          if (p.bias) {
            const float4 bb = __ldg(reinterpret_cast<const float4*>(p.bias + c) + g);
            t.x += bb.x; t.y += bb.y; t.z += bb.z; t.w += bb.w;
          }
          if (rrow) {
            const float4 rr = __ldg(reinterpret_cast<const float4*>(rrow + c) + g);
            t.x += rr.x; t.y += rr.y; t.z += rr.z; t.w += rr.w;
          }
          v[4 * g] = t.x; v[4 * g + 1] = t.y; v[4 * g + 2] = t.z; v[4 * g + 3] = t.w;
        }
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          float t = 0.f;
          if (c + i < p.dv) {
            t = __uint_as_float(r[i]) * inv;
            if (p.bias) t += __ldg(p.bias + c + i);
            if (rrow) t += __ldg(rrow + c + i);
          }
          v[i] = t;
        }
      }
    };
    float st1 = 0.f, st2 = 0.f;   // fused LayerNorm: this half's partial sum / sum of squares of the output row
    for (int c = c_begin; c < c_end; c += 32) {
      uint32_t r[32];
      tmem_ld32(tmem_o + lane_off + c, r);
      tmem_wait_ld();
      if (q < p.Nq) {
        float v[32];
        fill_chunk(c, r, v);
        if (p.out_ln != nullptr) {
#pragma unroll
          for (int i = 0; i < 32; ++i) {   // columns beyond dv hold zeros
            st1 += v[i];
            st2 = fmaf(v[i], v[i], st2);
          }
        }
        if (c + 32 <= p.dv && al_out) {
#pragma unroll
          for (int g = 0; g < 8; ++g)
            reinterpret_cast<float4*>(orow + c)[g] = make_float4(v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]);
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (c + i < p.dv) orow[c + i] = v[i];
        }
      }
    }
    if (p.out_ln != nullptr) {
      // ---- second pass: the row statistics of both halves, then the normalised 16-bit row (the MLP's operand) ----
      float* slot = xchg + ((ntiles + 1) & 1) * 256;   // the other exchange slot (the row-sum exchange used ntiles & 1)
      float* slot2 = xchg + 512;                       // a third [2][128] slot (DEC_BAR_BYTES)
      slot[half * 128 + row] = st1;
      slot2[half * 128 + row] = st2;
      pair_sync();
      st1 += slot[(half ^ 1) * 128 + row];
      st2 += slot2[(half ^ 1) * 128 + row];
      const float inv_c = 1.0f / (float)p.dv;
      const float mean = st1 * inv_c;
      const float rstd = rsqrtf(fmaxf(st2 * inv_c - mean * mean, 0.f) + p.ln_eps);
      const float nmr = -mean * rstd;
      uint16_t* lrow = p.out_ln + (long long)b * p.stride_ln + (long long)q * p.ld_ln;
      const bool al_ln = (reinterpret_cast<uintptr_t>(lrow) & 15u) == 0 &&
                         (p.ln_gamma == nullptr || (reinterpret_cast<uintptr_t>(p.ln_gamma) & 15u) == 0) &&
                         (p.ln_beta == nullptr || (reinterpret_cast<uintptr_t>(p.ln_beta) & 15u) == 0);
      const int ln_end = half == 0 ? c_end : (int)p.ld_ln;   // the second half also zeroes the pad columns
      for (int c = c_begin; c < ln_end; c += 32) {
        uint32_t r[32];
        tmem_ld32(tmem_o + lane_off + c, r);
        tmem_wait_ld();
        if (q < p.Nq) {
          float v[32];
          fill_chunk(c, r, v);
          if (c + 32 <= p.dv && al_ln) {
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              float y[8];
#pragma unroll
              for (int h2 = 0; h2 < 2; ++h2) {
                float4 gm = make_float4(1.f, 1.f, 1.f, 1.f), bt = make_float4(0.f, 0.f, 0.f, 0.f);
                if (p.ln_gamma) gm = __ldg(reinterpret_cast<const float4*>(p.ln_gamma + c) + 2 * g + h2);
                if (p.ln_beta) bt = __ldg(reinterpret_cast<const float4*>(p.ln_beta + c) + 2 * g + h2);
                y[4 * h2] = fmaf(fmaf(v[8 * g + 4 * h2], rstd, nmr), gm.x, bt.x);
                y[4 * h2 + 1] = fmaf(fmaf(v[8 * g + 4 * h2 + 1], rstd, nmr), gm.y, bt.y);
                y[4 * h2 + 2] = fmaf(fmaf(v[8 * g + 4 * h2 + 2], rstd, nmr), gm.z, bt.z);
                y[4 * h2 + 3] = fmaf(fmaf(v[8 * g + 4 * h2 + 3], rstd, nmr), gm.w, bt.w);
              }
              reinterpret_cast<uint4*>(lrow + c)[g] = make_uint4(pack16x2(y[0], y[1], p.fp16), pack16x2(y[2], y[3], p.fp16),
                                                                 pack16x2(y[4], y[5], p.fp16), pack16x2(y[6], y[7], p.fp16));
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const int cc = c + i;
              if (cc < p.ld_ln) {
                float y = 0.f;
                if (cc < p.dv) {
                  y = fmaf(v[i], rstd, nmr);
                  if (p.ln_gamma) y *= __ldg(p.ln_gamma + cc);
                  if (p.ln_beta) y += __ldg(p.ln_beta + cc);
                }
                lrow[cc] = cvt16(y, p.fp16);
              }
            }
          }
        }
      }
    }
  }

  __syncwarp();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // neither CTA exits (or frees TMEM) while its peer may still read its smem / signal it
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_2cta(tmem_base, 512);
  }
}

static bool decode_shape_ok(int dqk, int dv) {
  if (dqk < 1 || dqk > 384 || dv < 1) return false;
  const int dv_pad = (dv + 31) / 32 * 32;
  return dv_pad <= 384;      // 2 x 64 S columns + dv_pad O columns <= 512 TMEM columns
}

}  // namespace pio

extern "C" int pio_decoder_attention_supported(int32_t dqk, int32_t dv) {
  return pio::decode_shape_ok(dqk, dv) ? PIO_OK : PIO_ERR_UNSUPPORTED;
}

extern "C" int pio_decoder_attention_fwd(const pio_decoder_attention_args* a, void* stream_) {
  using namespace pio;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  PIO_REQUIRE(a != nullptr, "pio_decoder_attention_fwd: null args");
  PIO_REQUIRE(a->Q && a->K && a->V && a->out, "pio_decoder_attention_fwd: null operand / output");
  PIO_REQUIRE(a->B > 0 && a->Nq > 0 && a->Nk > 0 && a->dqk > 0 && a->dv > 0, "pio_decoder_attention_fwd: bad shape");
  PIO_REQUIRE(aligned16(a->Q) && aligned16(a->K) && aligned16(a->V), "pio_decoder_attention_fwd: operand base not 16-byte aligned");
  PIO_REQUIRE(a->ldq % 8 == 0 && a->ldk % 8 == 0 && a->ldv % 8 == 0 && a->ldq >= a->dqk && a->ldk >= a->dqk && a->ldv >= a->dv,
              "pio_decoder_attention_fwd: leading dims must be multiples of 8 and cover the head sizes");
  PIO_REQUIRE(a->strideQ % 8 == 0 && a->strideK % 8 == 0 && a->strideV % 8 == 0,
              "pio_decoder_attention_fwd: batch strides must be multiples of 8");
  PIO_REQUIRE(a->ldo >= a->dv && (!a->residual || a->ldr >= a->dv), "pio_decoder_attention_fwd: output / residual pitch < dv");
  PIO_REQUIRE(a->B < 65536, "pio_decoder_attention_fwd: grid too large");
  if (!decode_shape_ok(a->dqk, a->dv))
    return fail(PIO_ERR_UNSUPPORTED, "pio_decoder_attention_fwd: head sizes dqk=%d dv=%d not covered (dqk <= 384, dv <= 384)",
                a->dqk, a->dv);
  DeviceInfo dev;
  int rc = get_device_info(&dev);
  if (rc != PIO_OK) return rc;
  if (dev.cc_major != 10)
    return fail(PIO_ERR_ARCH, "pio_decoder_attention_fwd needs sm_100 (got sm_%d%d)", dev.cc_major, dev.cc_minor);

  DecodeParams p;
  p.B = a->B; p.Nq = a->Nq; p.Nk = a->Nk; p.dqk = a->dqk; p.dv = a->dv;
  p.fp16 = a->fp16 ? 1 : 0;
  p.nqc = (a->dqk + 63) / 64;
  const int dv_pad = (a->dv + 31) / 32 * 32;
  p.nv1 = dv_pad < 256 ? dv_pad : 256;
  p.nv2 = dv_pad - p.nv1;
  p.vch1 = (p.nv1 / 2 + 63) / 64;
  p.vch2 = p.nv2 > 0 ? 1 : 0;
  const int q_bytes = p.nqc * 16384, k_stage = p.nqc * DEC_KCHUNK, v_stage = (p.vch1 + p.vch2) * DEC_VCHUNK;
  const int budget = 232448 - DEC_BAR_BYTES - q_bytes;
  p.vst = 2;
  p.kst = (budget - p.vst * v_stage) / k_stage;
  if (p.kst > 4) p.kst = 4;
  if (p.kst >= 3 && budget - p.kst * k_stage - 3 * v_stage >= 0) p.vst = 3;
  if (p.kst < 2) return fail(PIO_ERR_UNSUPPORTED, "pio_decoder_attention_fwd: shared memory budget exceeded (dqk=%d dv=%d)", a->dqk, a->dv);
  int smem_bytes = q_bytes + p.kst * k_stage + p.vst * v_stage + DEC_BAR_BYTES;
  if (smem_bytes < 120 * 1024) smem_bytes = 120 * 1024;   // one CTA per SM: all 512 TMEM columns are allocated
  p.q_bcast = (a->strideQ == 0 && a->B > 1) ? 1 : 0;
  p.scale_log2 = a->scale * 1.4426950408889634f;
  p.key_mask = a->key_mask; p.stride_km = a->stride_km;
  p.row_keep = a->row_keep; p.stride_rk = a->stride_rk;
  p.bias = a->bias;
  p.residual = a->residual; p.ldr = a->ldr; p.strideR = a->strideR;
  p.out = a->out; p.ldo = a->ldo; p.strideO = a->strideO;
  p.out_ln = reinterpret_cast<uint16_t*>(a->out_ln); p.ld_ln = a->ld_ln; p.stride_ln = a->stride_ln;
  p.ln_gamma = a->ln_gamma; p.ln_beta = a->ln_beta; p.ln_eps = a->ln_eps;
  PIO_REQUIRE(!a->out_ln || (a->ld_ln >= a->dv && a->ld_ln % 2 == 0 && (reinterpret_cast<uintptr_t>(a->out_ln) & 3u) == 0 &&
                             a->stride_ln % 2 == 0),
              "pio_decoder_attention_fwd: out_ln needs an even, 4-byte aligned row pitch >= dv");

  CUtensorMap tq, tk, tv;
  {
    const uint64_t dims[3] = {(uint64_t)a->dqk, (uint64_t)a->Nq, (uint64_t)(p.q_bcast ? 1 : a->B)};
    const uint64_t strides[2] = {(uint64_t)a->ldq * 2,
                                 (uint64_t)((p.q_bcast || a->B == 1) ? a->ldq * (int64_t)a->Nq : a->strideQ) * 2};
    const uint32_t box[3] = {64, 128, 1};
    rc = encode_tmap_bf16(&tq, a->Q, 3, dims, strides, box);
    if (rc != PIO_OK) return rc;
  }
  {
    const uint64_t dims[3] = {(uint64_t)a->dqk, (uint64_t)a->Nk, (uint64_t)a->B};
    const uint64_t strides[2] = {(uint64_t)a->ldk * 2, (uint64_t)(a->B == 1 ? a->ldk * (int64_t)a->Nk : a->strideK) * 2};
    const uint32_t box[3] = {64, 32, 1};
    rc = encode_tmap_bf16(&tk, a->K, 3, dims, strides, box);
    if (rc != PIO_OK) return rc;
  }
  {
    const uint64_t dims[3] = {(uint64_t)a->dv, (uint64_t)a->Nk, (uint64_t)a->B};
    const uint64_t strides[2] = {(uint64_t)a->ldv * 2, (uint64_t)(a->B == 1 ? a->ldv * (int64_t)a->Nk : a->strideV) * 2};
    const uint32_t box[3] = {64, 64, 1};
    rc = encode_tmap_bf16(&tv, a->V, 3, dims, strides, box);
    if (rc != PIO_OK) return rc;
  }
  static PerDeviceOnce once;
  const cudaError_t attr_err = once.run(dev.device, [] {
    return cudaFuncSetAttribute(pio_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
  });
  if (attr_err != cudaSuccess)
    return fail(PIO_ERR_CUDA, "cudaFuncSetAttribute(decode) failed: %s", cudaGetErrorString(attr_err));
  const unsigned pairs = (unsigned)((a->Nq + 255) / 256);
  {
    ProfileScope prof(KF_FLASH, 2.0 * a->B * (double)a->Nq * a->Nk * (a->dqk + a->dv), 0.0, stream);
    PIO_CUDA_OK(launch_kernel(pio_decode_kernel, dim3(pairs * 2, (unsigned)a->B, 1), dim3(384, 1, 1), (size_t)smem_bytes, stream,
                              2, tq, tk, tv, p));
  }
#ifdef PIO_DECODE_TRACE
  {
    static bool dumped = false;
    if (!dumped) {
      dumped = true;
      cudaDeviceSynchronize();
      static unsigned long long host[3 * 2 * 512];
      cudaMemcpyFromSymbol(host, g_dec_trace, sizeof(host));
      for (int s = 0; s < 3; ++s) {
        unsigned long long t0 = ~0ull;
        for (int i = 0; i < 512; ++i)
          if (host[2 * (s * 512 + i)] != 0 && host[2 * (s * 512 + i) + 1] < t0) t0 = host[2 * (s * 512 + i) + 1];
        for (int i = 0; i < 512; ++i)
          if (host[2 * (s * 512 + i)] != 0)
            fprintf(stderr, "DT %d %llu %llu\n", s, host[2 * (s * 512 + i)], host[2 * (s * 512 + i) + 1] - t0);
      }
    }
  }
#endif
  g_launch_count.fetch_add(1);
  PIO_CUDA_OK(cudaGetLastError());
  return PIO_OK;
}
