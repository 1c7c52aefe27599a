// HBM-bound SIMT kernels: LayerNorm + bf16 cast, row softmax (materialised attention path) and the
// log-sum-exp merge of partial attention results (key splits / key-axis shards across GPUs).
#include <type_traits>

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
// LayerNorm + cast.  One warp per row, the row lives in registers, two-pass statistics (mean, then centred
// variance).  Two layouts:
//  * vector path (C % 4 == 0, 16-byte aligned rows: the 512/768/1024/1280-channel latent arrays): lane l owns the
//    float4 at columns 4*(l + 32 i); 16-byte loads, 8-byte bf16x4 stores, <= 64 registers so 32 warps per SM keep
//    >= 100 KB of loads in flight per SM;
//  * scalar path (odd widths 261 / 322 / 1026): lane l owns columns l + 32 i (coalesced for any C and any row
//    alignment) and each warp works on ROWS rows at once to double the loads in flight.
// ------------------------------------------------------------------------------------------------------------
template <int NV>
__global__ void __launch_bounds__(128) pio_layernorm_vec_kernel(const float* __restrict__ x, long long ldx,
                                                                __nv_bfloat16* __restrict__ y, long long ldy,
                                                                const float* __restrict__ gamma,
                                                                const float* __restrict__ beta, long long rows, int C,
                                                                int mode, float eps) {
  pdl_sync();
  const int normalize = mode & 1;
  const int f16 = (mode >> 3) & 1;   // 16-bit output format: 0 = bf16, 1 = fp16 (never with the split layouts)
  const int lane = threadIdx.x & 31;
  // Rows are walked from the END of the array: x was just written front-to-back by the producing GEMM (or H2D copy),
  // so its tail is what is still resident in the 126 MB L2; and the bf16 rows written last (the front) are the first
  // ones the consuming GEMM's TMA loads ask for.
  const long long row = (long long)(gridDim.x - 1 - blockIdx.x) * 4 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float4* xr = reinterpret_cast<const float4*>(x + row * ldx);
  const int nvec = C >> 2;
  float4 v[NV];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = lane + 32 * i;
    v[i] = (c < nvec) ? __ldg(xr + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  }
  // gamma / beta are requested here, before the reductions: left inside the output loop they cost a second memory round
  // trip after the statistics (4 x 1280 in a dependent chain: 4.7 us with the affine, 2.6 us without)
  float4 gv[NV], bv[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = lane + 32 * i;
    gv[i] = (gamma && c < nvec) ? __ldg(reinterpret_cast<const float4*>(gamma) + c) : make_float4(1.f, 1.f, 1.f, 1.f);
    bv[i] = (beta && c < nvec) ? __ldg(reinterpret_cast<const float4*>(beta) + c) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  float mean = 0.f, rstd = 1.f;
  if (normalize) {
    mean = warp_sum(s) / (float)C;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      if (lane + 32 * i < nvec) {
        const float a = v[i].x - mean, b = v[i].y - mean, c2 = v[i].z - mean, d = v[i].w - mean;
        q += (a * a + b * b) + (c2 * c2 + d * d);
      }
    }
    rstd = rsqrtf(warp_sum(q) / (float)C + eps);
  }
  uint2* yr = reinterpret_cast<uint2*>(y + row * ldy);
  // validation mode: three segments of ldy / 3 columns, [hi | lo | hi] (A side, bit 1) or [hi | hi | lo] (B side, bit 2)
  const bool split = (mode & 6) != 0;
  const bool b_side = (mode & 4) != 0;
  const int nvec_out = split ? (int)((ldy / 3) >> 2) : (int)(ldy >> 2);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = lane + 32 * i;
    if (c < nvec_out) {
      float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c < nvec) {
        o.x = (v[i].x - mean) * rstd; o.y = (v[i].y - mean) * rstd;
        o.z = (v[i].z - mean) * rstd; o.w = (v[i].w - mean) * rstd;
        o.x = fmaf(o.x, gv[i].x, bv[i].x); o.y = fmaf(o.y, gv[i].y, bv[i].y);
        o.z = fmaf(o.z, gv[i].z, bv[i].z); o.w = fmaf(o.w, gv[i].w, bv[i].w);
      }
      const uint2 hi = make_uint2(pack16x2(o.x, o.y, f16), pack16x2(o.z, o.w, f16));
      yr[c] = hi;
      if (split) {
        const uint2 lo = make_uint2(pack_bf16x2(o.x - __uint_as_float(hi.x << 16), o.y - __uint_as_float(hi.x & 0xffff0000u)),
                                    pack_bf16x2(o.z - __uint_as_float(hi.y << 16), o.w - __uint_as_float(hi.y & 0xffff0000u)));
        yr[nvec_out + c] = b_side ? hi : lo;
        yr[2 * nvec_out + c] = b_side ? lo : hi;
      }
    }
  }
}

template <int MAXV, int ROWS, bool SPLIT>
__global__ void __launch_bounds__(128) pio_layernorm_kernel(const float* __restrict__ x, long long ldx,
                                                            __nv_bfloat16* __restrict__ y, long long ldy,
                                                            const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, long long rows, int C,
                                                            int mode, float eps) {
  pdl_sync();
  // This path is instruction-issue bound (ncu: 72 % issue-active at 365 instructions per row in its first form), so
  // each warp walks many row groups with the affine parameters held in registers and as little per-element control
  // flow as possible.
  const int normalize = mode & 1;
  const int f16 = SPLIT ? 0 : (mode >> 3) & 1;   // 16-bit output format: 0 = bf16, 1 = fp16
  // validation mode: three segments of ldy / 3 columns, [hi | lo | hi] (A side, bit 1) or [hi | hi | lo] (B side, bit 2)
  const bool b_side = (mode & 4) != 0;
  const int seg = SPLIT ? (int)(ldy / 3) : (int)ldy;
  const int lane = threadIdx.x & 31;
  const long long wid = (long long)blockIdx.x * 4 + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * 4;
  const long long ngroups = (rows + ROWS - 1) / ROWS;
  const float inv_c = 1.0f / (float)C;
  float g[MAXV], bt[MAXV];
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int c = lane + 32 * i;
    g[i] = (gamma && c < C) ? __ldg(gamma + c) : 1.f;
    bt[i] = (beta && c < C) ? __ldg(beta + c) : 0.f;
  }
  // back to front: the tail of x is what is still resident in L2, and the front of y is what the consumer asks for first
  for (long long grp = ngroups - 1 - wid; grp >= 0; grp -= nwarps) {
    const long long row0 = grp * ROWS;
    float v[ROWS][MAXV];
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
      const bool rok = row0 + r < rows;
      const float* xr = x + (row0 + r) * ldx;
#pragma unroll
      for (int i = 0; i < MAXV; ++i) {
        const int c = lane + 32 * i;
        v[r][i] = (rok && c < C) ? __ldg(xr + c) : 0.f;
      }
    }
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
      if (row0 + r >= rows) break;  // warp-uniform
      float mean = 0.f, rstd = 1.f;
      if (normalize) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < MAXV; ++i) s += v[r][i];
        mean = warp_sum(s) * inv_c;
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < MAXV; ++i) {
          const float d = (lane + 32 * i < C) ? v[r][i] - mean : 0.f;
          q = fmaf(d, d, q);
        }
        rstd = rsqrtf(warp_sum(q) * inv_c + eps);
      }
      __nv_bfloat16* yr = y + (row0 + r) * ldy;
      const float nmr = -mean * rstd;
#pragma unroll
      for (int i = 0; i < MAXV; ++i) {
        const int c = lane + 32 * i;
        if (c < seg) {
          // (v - mean) * rstd * gamma + beta; pad columns (C <= c < seg) have v = 0, gamma = 1, beta = 0 -> forced to 0
          float o = fmaf(fmaf(v[r][i], rstd, nmr), g[i], bt[i]);
          if (c >= C) o = 0.f;
          if (!SPLIT) {
            reinterpret_cast<uint16_t*>(yr)[c] = cvt16(o, f16);
            continue;
          }
          const __nv_bfloat16 hi = __float2bfloat16_rn(o);
          yr[c] = hi;
          if (SPLIT) {
            const __nv_bfloat16 lo = __float2bfloat16_rn(o - __bfloat162float(hi));
            yr[seg + c] = b_side ? hi : lo;
            yr[2 * seg + c] = b_side ? lo : hi;
          }
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// LayerNorm + cast for CONTIGUOUS arrays of odd-width rows (the 261 / 322 / 1026-channel input and query arrays:
// rows are 1044 / 1288 / 4104 bytes, so neither vector loads nor a tiled tensor map can address them).  The array is
// treated as a flat byte stream: a group of R rows (R % 4 == 0 -> 16-byte aligned start and size) arrives in shared
// memory with ONE bulk copy (cp.async.bulk, mbarrier complete_tx), warps normalise rows shared -> registers -> shared,
// and the padded bf16 rows (pitch ldy, contiguous in global memory) leave with ONE bulk store.  No per-thread global
// access, NS input stages in flight per CTA, several CTAs per SM.
// ------------------------------------------------------------------------------------------------------------
constexpr int LNB_THREADS = 128;
constexpr int LNB_STAGES = 3;

// NFULL = C / 32 column slots of a lane are always inside the row (no predicate); one more slot covers the row's tail
// and the zero pad up to ldy (< 32 * (NFULL + 1)).  The row loop is instruction-issue bound, so this matters: the
// generic form spent more instructions on column predicates than on arithmetic.
template <int NFULL>
__global__ void __launch_bounds__(LNB_THREADS) pio_layernorm_bulk_kernel(const float* __restrict__ x,
                                                                         __nv_bfloat16* __restrict__ y, int ldy,
                                                                         const float* __restrict__ gamma,
                                                                         const float* __restrict__ beta,
                                                                         long long ngroups, int R, int C, int mode,
                                                                         float eps) {
  pdl_sync();
  const int normalize = mode & 1;
  const int f16 = (mode >> 3) & 1;   // 16-bit output format: 0 = bf16, 1 = fp16
  extern __shared__ __align__(128) uint8_t lnb_smem[];
  const uint32_t in_bytes = (uint32_t)R * (uint32_t)C * 4u;          // multiple of 16 (R % 4 == 0)
  const uint32_t in_pitch = (in_bytes + 127u) & ~127u;
  const uint32_t out_bytes = (uint32_t)R * (uint32_t)ldy * 2u;       // multiple of 16 (ldy % 8 == 0)
  const uint32_t out_pitch = (out_bytes + 127u) & ~127u;
  uint8_t* in_buf = lnb_smem;
  uint8_t* out_buf = lnb_smem + LNB_STAGES * in_pitch;
  uint64_t* full = reinterpret_cast<uint64_t*>(out_buf + 2 * out_pitch);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long my_groups = (ngroups - blockIdx.x + gridDim.x - 1) / gridDim.x;   // groups blockIdx.x + i * gridDim.x

  auto issue_load = [&](long long i) {
    const long long g = blockIdx.x + i * gridDim.x;
    const int s = (int)(i % LNB_STAGES);
    mbar_arrive_expect_tx(&full[s], in_bytes);
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(in_buf + s * in_pitch)), "l"(reinterpret_cast<uint64_t>(x) + (uint64_t)g * in_bytes),
                 "r"(in_bytes), "r"(smem_u32(&full[s]))
                 : "memory");
  };

  if (tid == 0) {
    for (int s = 0; s < LNB_STAGES; ++s) mbar_init(&full[s], 1);
    fence_mbar_init();
    for (long long i = 0; i < LNB_STAGES - 1 && i < my_groups; ++i) issue_load(i);
  }
  const int ct = 32 * NFULL + lane;        // this lane's tail column
  const bool t_in = ct < C, t_out = ct < ldy;
  float g[NFULL + 1], bt[NFULL + 1];
#pragma unroll
  for (int i = 0; i < NFULL; ++i) {
    g[i] = gamma ? __ldg(gamma + lane + 32 * i) : 1.f;
    bt[i] = beta ? __ldg(beta + lane + 32 * i) : 0.f;
  }
  g[NFULL] = (gamma && t_in) ? __ldg(gamma + ct) : (t_in ? 1.f : 0.f);   // pad columns: 0 * x + 0
  bt[NFULL] = (beta && t_in) ? __ldg(beta + ct) : 0.f;
  const float inv_c = 1.0f / (float)C;
  for (long long it = 0; it < my_groups; ++it) {
    const int s = (int)(it % LNB_STAGES);
    if (tid == 0) bulk_wait_read<1>();   // the store that read out_buf[it & 1] two iterations ago has drained
    __syncthreads();                     // ... and every warp is done with the input stage of iteration it - 1
    if (tid == 0 && it + LNB_STAGES - 1 < my_groups) issue_load(it + LNB_STAGES - 1);
    mbar_wait(&full[s], (uint32_t)((it / LNB_STAGES) & 1));
    const float* in = reinterpret_cast<const float*>(in_buf + s * in_pitch);
    __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(out_buf + (it & 1) * out_pitch);
    // two rows per warp at a time (R % 4 == 0 and 4 warps: rows warp*2 + {0,1} + 8 k)
    for (int r = warp * 2; r < R; r += 2 * (LNB_THREADS / 32)) {
      const float* x0 = in + r * C + lane;
      const float* x1 = x0 + C;
      float v0[NFULL + 1], v1[NFULL + 1];
      float s0 = 0.f, s1 = 0.f;
#pragma unroll
      for (int i = 0; i < NFULL; ++i) {
        v0[i] = x0[32 * i];
        v1[i] = x1[32 * i];
        s0 += v0[i];
        s1 += v1[i];
      }
      v0[NFULL] = t_in ? x0[32 * NFULL] : 0.f;
      v1[NFULL] = t_in ? x1[32 * NFULL] : 0.f;
      s0 += v0[NFULL];
      s1 += v1[NFULL];
      float mean0 = 0.f, rstd0 = 1.f, mean1 = 0.f, rstd1 = 1.f;
      if (normalize) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          s0 += __shfl_xor_sync(0xffffffffu, s0, o);
          s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        }
        mean0 = s0 * inv_c;
        mean1 = s1 * inv_c;
        float q0 = 0.f, q1 = 0.f;
#pragma unroll
        for (int i = 0; i < NFULL; ++i) {
          const float d0 = v0[i] - mean0, d1 = v1[i] - mean1;
          q0 = fmaf(d0, d0, q0);
          q1 = fmaf(d1, d1, q1);
        }
        {
          const float d0 = t_in ? v0[NFULL] - mean0 : 0.f, d1 = t_in ? v1[NFULL] - mean1 : 0.f;
          q0 = fmaf(d0, d0, q0);
          q1 = fmaf(d1, d1, q1);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          q0 += __shfl_xor_sync(0xffffffffu, q0, o);
          q1 += __shfl_xor_sync(0xffffffffu, q1, o);
        }
        rstd0 = rsqrtf(q0 * inv_c + eps);
        rstd1 = rsqrtf(q1 * inv_c + eps);
      }
      const float n0 = -mean0 * rstd0, n1 = -mean1 * rstd1;
      uint16_t* y0 = reinterpret_cast<uint16_t*>(out + r * ldy + lane);
      uint16_t* y1 = y0 + ldy;
      // one uniform branch per row pair selects the 16-bit format (the loop is issue-bound: no per-element select)
      auto store_rows = [&](auto f16tag) {
        constexpr bool F16 = decltype(f16tag)::value;
#pragma unroll
        for (int i = 0; i < NFULL; ++i) {
          y0[32 * i] = (uint16_t)pack16x2<F16>(fmaf(fmaf(v0[i], rstd0, n0), g[i], bt[i]), 0.f);
          y1[32 * i] = (uint16_t)pack16x2<F16>(fmaf(fmaf(v1[i], rstd1, n1), g[i], bt[i]), 0.f);
        }
        if (t_out) {   // tail of the row and the zero pad (g = bt = 0 there)
          y0[32 * NFULL] = (uint16_t)pack16x2<F16>(fmaf(fmaf(v0[NFULL], rstd0, n0), g[NFULL], bt[NFULL]), 0.f);
          y1[32 * NFULL] = (uint16_t)pack16x2<F16>(fmaf(fmaf(v1[NFULL], rstd1, n1), g[NFULL], bt[NFULL]), 0.f);
        }
      };
      if (f16) store_rows(std::true_type{});
      else store_rows(std::false_type{});
    }
    fence_proxy_async_smem();
    __syncthreads();
    if (tid == 0) {
      const long long gidx = blockIdx.x + it * gridDim.x;
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                   ::"l"(reinterpret_cast<uint64_t>(y) + (uint64_t)gidx * out_bytes),
                   "r"(smem_u32(out_buf + (it & 1) * out_pitch)), "r"(out_bytes)
                   : "memory");
      bulk_commit();
    }
  }
  if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// ------------------------------------------------------------------------------------------------------------
// LayerNorm of cat([features | position table]) without the concatenated array (pio_layernorm_concat_bf16).
// A CTA takes groups of R consecutive positions: the R table rows arrive with one bulk copy, the features of all B
// samples for those positions are gathered into shared memory (with their sums), and then every warp walks the samples
// for its two positions, writing each normalised row straight from registers.  The table is read once per position
// group, not once per sample; per row only the Cf feature values and a few scalars change:
//   y_c = x_c * (gamma_c * rstd) + (beta_c - mean * rstd * gamma_c),  x_c * gamma_c precomputed for the table part,
//   mean = (S_pos + S_feat) / C,  var = ((Q_pos + Q_feat) - mean (S_pos + S_feat)) / C   (S: sums, Q: sums of squares).
// HBM-write bound: 2 * ldy bytes per row out, the table and the image are read from L2.
// ------------------------------------------------------------------------------------------------------------
constexpr int LNC_MAX_FEAT_BYTES = 48 * 1024;
constexpr int LNC_THREADS = 256;     // 8 warps, two rows of a 16-position group each
constexpr int LNC_STAGES = 2;        // table stages: one load per R * B output rows, depth hardly matters
constexpr int LNC_MAX_NP = 18;       // up to 64 * 19 = 1216 columns

// NP = C / 64 pair slots of a lane are always inside the row: slot i holds columns 2 lane + 64 i and the next one, so a
// warp writes 128 contiguous bytes per instruction straight from registers (no staging, no barrier, no TMA descriptor
// per 1 KB of output — 1-KB bulk stores capped the first version at one store per ~80 cycles and SM).
template <int NP>
__global__ void __launch_bounds__(LNC_THREADS, 3) pio_layernorm_concat_kernel(pio_layernorm_concat_args a, int R,
                                                                              int ngroups) {
  pdl_sync();
  extern __shared__ __align__(128) uint8_t lnb_smem[];
  const int Cf = a.Cf, Cp = a.Cp, C = Cf + Cp, ldy = (int)a.ldy, B = a.B;
  const uint32_t in_bytes = (uint32_t)R * (uint32_t)Cp * 4u;         // multiple of 16 (R % 4 == 0)
  const uint32_t in_pitch = (in_bytes + 127u) & ~127u;
  const uint32_t feat_bytes = ((uint32_t)B * (uint32_t)R * (uint32_t)Cf * 4u + 127u) & ~127u;
  uint8_t* in_buf = lnb_smem;
  float* feat = reinterpret_cast<float*>(in_buf + LNC_STAGES * in_pitch);   // [B][R][Cf]
  float2* fstat = reinterpret_cast<float2*>(reinterpret_cast<uint8_t*>(feat) + feat_bytes);   // [B][R] (sum, sum of squares)
  uint64_t* full = reinterpret_cast<uint64_t*>(fstat + (size_t)B * R);
  __nv_bfloat16* y = reinterpret_cast<__nv_bfloat16*>(a.y);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int my_groups = (ngroups - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  auto issue_load = [&](int i) {
    const long long g = blockIdx.x + (long long)i * gridDim.x;
    const int s = i % LNC_STAGES;
    mbar_arrive_expect_tx(&full[s], in_bytes);
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(in_buf + s * in_pitch)), "l"(reinterpret_cast<uint64_t>(a.pos) + (uint64_t)g * in_bytes),
                 "r"(in_bytes), "r"(smem_u32(&full[s]))
                 : "memory");
  };
  if (tid == 0) {
    for (int s = 0; s < LNC_STAGES; ++s) mbar_init(&full[s], 1);
    fence_mbar_init();
    if (my_groups > 0) issue_load(0);
  }
  // pair slot i of a lane: columns c0 = 2 lane + 64 i and c0 + 1 of the concatenated row — a feature (c < Cf), table
  // column c - Cf, or zero pad (C <= c < ldy: gamma = beta = 0)
  float g[NP + 1][2], bt[NP + 1][2];
#pragma unroll
  for (int i = 0; i <= NP; ++i)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int c = 2 * lane + 64 * i + e;
      g[i][e] = (c < C) ? (a.gamma ? __ldg(a.gamma + c) : 1.f) : 0.f;
      bt[i][e] = (a.beta && c < C) ? __ldg(a.beta + c) : 0.f;
    }
  uint64_t g2[NP + 1], bt2[NP + 1];
#pragma unroll
  for (int i = 0; i <= NP; ++i) {
    g2[i] = pack_f32x2(g[i][0], g[i][1]);
    bt2[i] = pack_f32x2(bt[i][0], bt[i][1]);
  }
  const bool few_feat = Cf <= 64;          // the usual case (3 pixel channels): only slot 0 can hold features
  const float inv_c = 1.0f / (float)C;
  const bool t_out = 64 * NP + 2 * lane < ldy;
  __syncthreads();
  for (int it = 0; it < my_groups; ++it) {
    const int s = it % LNC_STAGES;
    const long long n0 = (blockIdx.x + (long long)it * gridDim.x) * R;
    // One thread per (sample, position): consecutive threads read consecutive positions of one channel, and the
    // feature part of the row statistics is formed here, off the sample loop.
    if (Cf <= 4) {
      // up to 16 loads of a thread are in flight before the first one is consumed
      for (int idx0 = tid; idx0 < B * R; idx0 += 4 * LNC_THREADS) {
        float f[4][4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int idx = idx0 + u * LNC_THREADS;
          const int r = idx % R, b = idx / R;
          const float* src = a.feat + (long long)b * a.feat_stride_b + (n0 + r) * a.feat_stride_n;
#pragma unroll
          for (int c = 0; c < 4; ++c)
            f[u][c] = (idx < B * R && c < Cf) ? __ldg(src + (long long)c * a.feat_stride_c) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int idx = idx0 + u * LNC_THREADS;
          if (idx < B * R) {
            float sf = 0.f, qf = 0.f;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              if (c < Cf) feat[(long long)idx * Cf + c] = f[u][c];
              sf += f[u][c];
              qf = fmaf(f[u][c], f[u][c], qf);
            }
            fstat[idx] = make_float2(sf, qf);
          }
        }
      }
    } else {
      for (int idx = tid; idx < B * R; idx += LNC_THREADS) {
        const int r = idx % R, b = idx / R;
        const float* src = a.feat + (long long)b * a.feat_stride_b + (n0 + r) * a.feat_stride_n;
        float* dst = feat + (long long)idx * Cf;
        float sf = 0.f, qf = 0.f;
        for (int c = 0; c < Cf; ++c) {
          const float f = __ldg(src + (long long)c * a.feat_stride_c);
          dst[c] = f;
          sf += f;
          qf = fmaf(f, f, qf);
        }
        fstat[idx] = make_float2(sf, qf);
      }
    }
    if (tid == 0 && it + 1 < my_groups) issue_load(it + 1);   // its stage was last read in iteration it - 1
    mbar_wait(&full[s], (uint32_t)((it / LNC_STAGES) & 1));
    __syncthreads();
    const float* in = reinterpret_cast<const float*>(in_buf + s * in_pitch);
    // this warp's rows of the group: r = 2 warp, 2 warp + 1 (R <= 16); the table part of the rows (times gamma) and its
    // statistics stay in registers across the sample loop
    constexpr int MAXROWS = 2;
    const int r0 = 2 * warp;
    const bool active = r0 < R;          // R is a multiple of 4
    float pg[MAXROWS][NP + 1][2];
    float s_pos[MAXROWS], q_pos[MAXROWS];
#pragma unroll
    for (int k = 0; k < MAXROWS; ++k) {
      const int r = r0 + k;
      float sp = 0.f, sq = 0.f;
#pragma unroll
      for (int i = 0; i <= NP; ++i)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int c = 2 * lane + 64 * i + e;
          const float p = (active && c >= Cf && c < C) ? in[r * Cp + (c - Cf)] : 0.f;
          sp += p;
          sq = fmaf(p, p, sq);
          pg[k][i][e] = p * g[i][e];
        }
      s_pos[k] = warp_sum(sp);
      q_pos[k] = warp_sum(sq);
    }
    uint64_t pg2[MAXROWS][NP + 1];
#pragma unroll
    for (int k = 0; k < MAXROWS; ++k)
#pragma unroll
      for (int i = 0; i <= NP; ++i) pg2[k][i] = pack_f32x2(pg[k][i][0], pg[k][i][1]);
    if (active) {
      for (int b = 0; b < B; ++b) {
#pragma unroll
        for (int k = 0; k < MAXROWS; ++k) {
          const int r = r0 + k;
          const float* fr = feat + ((long long)b * R + r) * Cf;
          // mean and variance from the two partial (sum, sum of squares) pairs: sum (x - mean)^2 = Q - mean S
          const float2 fs = fstat[b * R + r];
          const float ssum = s_pos[k] + fs.x;
          const float mean = ssum * inv_c;
          const float var = ((q_pos[k] + fs.y) - mean * ssum) * inv_c;
          const float rstd = rsqrtf(fmaxf(var, 0.f) + a.eps);
          const float nmr = -mean * rstd;
          uint32_t* yr = reinterpret_cast<uint32_t*>(y + ((long long)b * a.N + n0 + r) * ldy) + lane;
          const uint64_t nmr2 = pack_f32x2(nmr, nmr), rstd2 = pack_f32x2(rstd, rstd);
          // one uniform branch per row selects the 16-bit format (the loop is issue-bound: no per-element select)
          auto store_row = [&](auto f16tag) {
            constexpr bool F16 = decltype(f16tag)::value;
#pragma unroll
            for (int i = 0; i <= NP; ++i) {
              if (i < NP || t_out) {
                uint64_t x2 = pg2[k][i];
                if (i == 0 || !few_feat) {   // slots that can hold features
                  const int c = 2 * lane + 64 * i;
                  if (c < Cf) {
                    float x0, x1;
                    unpack_f32x2(x2, x0, x1);
                    x0 = fr[c] * g[i][0];
                    if (c + 1 < Cf) x1 = fr[c + 1] * g[i][1];
                    x2 = pack_f32x2(x0, x1);
                  }
                }
                // two packed fp32x2 FMAs per column pair: x * rstd + (-mean * rstd * gamma + beta)
                float y0, y1;
                unpack_f32x2(ffma2(x2, rstd2, ffma2(nmr2, g2[i], bt2[i])), y0, y1);
                yr[32 * i] = pack16x2<F16>(y0, y1);
              }
            }
          };
          if (a.fp16) store_row(std::true_type{});
          else store_row(std::false_type{});
        }
      }
    }
    __syncthreads();   // every warp is done with this group's table stage and features
  }
}

using lnc_kernel_t = void (*)(pio_layernorm_concat_args, int, int);
template <int... I>
static const lnc_kernel_t* lnc_kernel_table(std::integer_sequence<int, I...>) {
  static const lnc_kernel_t table[] = {pio_layernorm_concat_kernel<I>...};
  return table;
}

using lnb_kernel_t = void (*)(const float*, __nv_bfloat16*, int, const float*, const float*, long long, int, int, int, float);
constexpr int LNB_MAX_NFULL = 36;
template <int... I>
static const lnb_kernel_t* lnb_kernel_table(std::integer_sequence<int, I...>) {
  static const lnb_kernel_t table[] = {pio_layernorm_bulk_kernel<I>...};
  return table;
}

// ------------------------------------------------------------------------------------------------------------
// Row softmax, one 256-thread block per row (rows can be 50k+ long): three passes over the row, which stays in
// L1/L2 (<= 208 KB).  Masked keys are excluded; an all-masked or wiped row is written as zeros.
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pio_softmax_kernel(pio_softmax_args a) {
  pdl_sync();
  __shared__ float red[8];
  __shared__ float bcast;
  const long long row_id = blockIdx.x;
  const int b = (int)(row_id / a.rows);
  const int r = (int)(row_id % a.rows);
  const float* s = a.S + (long long)b * a.strideS + (long long)r * a.lds;
  __nv_bfloat16* p = reinterpret_cast<__nv_bfloat16*>(a.P) + (long long)b * a.strideP + (long long)r * a.ldp;
  const uint8_t* km = a.key_mask ? a.key_mask + (long long)b * a.stride_km : nullptr;
  const uint8_t* dm = a.dense_mask ? a.dense_mask + (long long)b * a.dm_stride_b + (long long)r * a.dm_stride_r : nullptr;
  const float* bias = a.bias ? a.bias + (long long)b * a.bias_stride_b + (long long)r * a.bias_stride_r : nullptr;
  float* pf = a.P_f32 ? a.P_f32 + (long long)b * a.stridePf + (long long)r * a.ldpf : nullptr;
  const bool keep = a.row_keep ? a.row_keep[(long long)b * a.stride_rk + r] != 0 : true;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int seg = a.split ? (int)(a.ldp / 3) : (int)a.ldp;   // validation mode: [hi | lo | hi] segments
  auto valid = [&](int c) { return (!km || km[c]) && (!dm || dm[c]); };
  auto logit = [&](int c) {
    float v = __ldg(s + c);
    if (bias) v += __ldg(bias + (long long)c * a.bias_stride_c);
    return v * a.scale;
  };
  float m = -INFINITY;
  if (keep)
    for (int c = tid; c < a.cols; c += 256)
      if (valid(c)) m = fmaxf(m, logit(c));
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
  if (m == -INFINITY) {  // wiped row or every key masked: zeros for P.V (:168-175); the returned matrix is uniform,
                         // which is what softmax makes of a row of -1e30 in the reference
    for (int c = tid; c < a.ldp; c += 256) p[c] = __float2bfloat16_rn(0.f);   // 0 has the same bits in fp16
    if (pf) {
      const float u = 1.0f / (float)a.cols;
      for (int c = tid; c < a.cols; c += 256) pf[c] = u;
    }
    return;
  }
  float l = 0.f;
  for (int c = tid; c < a.cols; c += 256)
    if (valid(c)) l += __expf(logit(c) - m);
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
  for (int c = tid; c < seg; c += 256) {
    float o = 0.f;
    if (c < a.cols && valid(c)) o = __expf(logit(c) - m) * inv;
    if (!a.split) {
      reinterpret_cast<uint16_t*>(p)[c] = cvt16(o, a.fp16);
    } else {
      const __nv_bfloat16 hi = __float2bfloat16_rn(o);
      p[c] = hi;
      p[seg + c] = __float2bfloat16_rn(o - __bfloat162float(hi));
      p[2 * seg + c] = hi;
    }
    if (pf && c < a.cols) pf[c] = o;
  }
}

// ------------------------------------------------------------------------------------------------------------
// Long rows without the general arguments (the multimodal encoder's explicit-S path: 784 rows x 52,097 keys): one
// 512-thread block per row, TWO passes with 16-byte loads — pass 1 keeps an online (max, sum) pair per thread (one read of
// the row instead of two), pass 2 re-reads the row (L2: a row is <= 208 KB) and writes four 16-bit probabilities per
// 8-byte store.  HBM/L2-bound: 8 cols bytes in + 2 ldp bytes out per row.
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(512) pio_softmax_long_kernel(pio_softmax_args a) {
  pdl_sync();
  __shared__ float red_m[16], red_l[16];
  const long long row_id = blockIdx.x;
  const int b = (int)(row_id / a.rows);
  const int r = (int)(row_id % a.rows);
  const float4* s4 = reinterpret_cast<const float4*>(a.S + (long long)b * a.strideS + (long long)r * a.lds);
  uint2* p2 = reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(a.P) + (long long)b * a.strideP + (long long)r * a.ldp);
  const uint8_t* km = a.key_mask ? a.key_mask + (long long)b * a.stride_km : nullptr;
  const bool keep = a.row_keep ? a.row_keep[(long long)b * a.stride_rk + r] != 0 : true;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nvec = (a.cols + 3) >> 2;
  const int nvec_out = (int)(a.ldp >> 2);
  auto load4 = [&](int c4, float (&v)[4]) {
    const float4 t = __ldg(s4 + c4);
    const int c = 4 * c4;
    v[0] = t.x * a.scale; v[1] = t.y * a.scale; v[2] = t.z * a.scale; v[3] = t.w * a.scale;
#pragma unroll
    for (int e = 0; e < 4; ++e)
      if (c + e >= a.cols || (km && !km[c + e])) v[e] = -INFINITY;
  };
  float m = -INFINITY, l = 0.f;
  if (keep) {
    for (int c4 = tid; c4 < nvec; c4 += 512) {
      float v[4];
      load4(c4, v);
      const float mx = fmaxf(fmaxf(v[0], v[1]), fmaxf(v[2], v[3]));
      if (mx > m) {
        l *= __expf(m - mx);      // m == -inf: l is 0 and exp(-inf) = 0
        m = mx;
      }
      if (m != -INFINITY) l += (__expf(v[0] - m) + __expf(v[1] - m)) + (__expf(v[2] - m) + __expf(v[3] - m));
    }
  }
  // block-wide merge of the (max, sum) pairs
  {
    const float wm = warp_max(m);
    l = (m == -INFINITY) ? 0.f : l * __expf(m - wm);
    l = warp_sum(l);
    if (lane == 0) { red_m[warp] = wm; red_l[warp] = l; }
  }
  __syncthreads();
  float M = -INFINITY;
#pragma unroll
  for (int i = 0; i < 16; ++i) M = fmaxf(M, red_m[i]);
  float L = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) L += (red_m[i] == -INFINITY) ? 0.f : red_l[i] * __expf(red_m[i] - M);
  const float inv = (M != -INFINITY && L > 0.f) ? 1.0f / L : 0.0f;   // wiped / fully masked rows are written as zeros
  for (int c4 = tid; c4 < nvec_out; c4 += 512) {
    float o[4] = {0.f, 0.f, 0.f, 0.f};
    if (inv != 0.f && c4 < nvec) {
      float v[4];
      load4(c4, v);
#pragma unroll
      for (int e = 0; e < 4; ++e) o[e] = __expf(v[e] - M) * inv;
    }
    p2[c4] = make_uint2(pack16x2(o[0], o[1], a.fp16), pack16x2(o[2], o[3], a.fp16));
  }
}

// ------------------------------------------------------------------------------------------------------------
// Row softmax for rows of up to 2048 columns (the decoders: 256 .. 2048 latents as keys): one warp per row, the row
// lives in registers (lane l owns the float4 at columns 4 (l + 32 i)), S is read exactly once with 16-byte loads and P
// is written with 8-byte stores.  HBM-bound: 4 cols bytes in + 2 ldp bytes out per row.
// ------------------------------------------------------------------------------------------------------------
template <int NV>
__global__ void __launch_bounds__(256) pio_softmax_warp_kernel(pio_softmax_args a) {
  pdl_sync();
  const int lane = threadIdx.x & 31;
  const long long row_id = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row_id >= (long long)a.batch * a.rows) return;
  const int b = (int)(row_id / a.rows);
  const int r = (int)(row_id % a.rows);
  const float4* s4 = reinterpret_cast<const float4*>(a.S + (long long)b * a.strideS + (long long)r * a.lds);
  uint2* p2 = reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(a.P) + (long long)b * a.strideP + (long long)r * a.ldp);
  const uint8_t* km = a.key_mask ? a.key_mask + (long long)b * a.stride_km : nullptr;
  const bool keep = a.row_keep ? a.row_keep[(long long)b * a.stride_rk + r] != 0 : true;
  const int nvec = (a.cols + 3) >> 2;      // float4s that hold at least one valid column
  const int nvec_out = (int)(a.ldp >> 2);
  float4 v[NV];
  float m = -INFINITY;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c4 = lane + 32 * i;
    v[i] = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    if (keep && c4 < nvec) {
      const float4 t = __ldg(s4 + c4);
      const int c = 4 * c4;
      const bool k0 = (c < a.cols) && (!km || km[c]), k1 = (c + 1 < a.cols) && (!km || km[c + 1]);
      const bool k2 = (c + 2 < a.cols) && (!km || km[c + 2]), k3 = (c + 3 < a.cols) && (!km || km[c + 3]);
      if (k0) v[i].x = t.x * a.scale;
      if (k1) v[i].y = t.y * a.scale;
      if (k2) v[i].z = t.z * a.scale;
      if (k3) v[i].w = t.w * a.scale;
      m = fmaxf(fmaxf(m, fmaxf(v[i].x, v[i].y)), fmaxf(v[i].z, v[i].w));
    }
  }
  m = warp_max(m);
  float l = 0.f;
  if (m != -INFINITY) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      v[i].x = __expf(v[i].x - m); v[i].y = __expf(v[i].y - m);
      v[i].z = __expf(v[i].z - m); v[i].w = __expf(v[i].w - m);
      l += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
  }
  l = warp_sum(l);
  const float inv = (m != -INFINITY && l > 0.f) ? 1.0f / l : 0.0f;   // wiped / fully masked rows are written as zeros
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c4 = lane + 32 * i;
    if (c4 < nvec_out) {
      float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
      if (inv != 0.f && c4 < nvec) o = make_float4(v[i].x * inv, v[i].y * inv, v[i].z * inv, v[i].w * inv);
      p2[c4] = make_uint2(pack16x2(o.x, o.y, a.fp16), pack16x2(o.z, o.w, a.fp16));
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// Merge partial attention results: one warp per (b, h, query row).
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pio_combine_kernel(pio_combine_args a) {
  pdl_sync();
  const int lane = threadIdx.x & 31;
  const long long rows = (long long)a.B * a.H * a.Nq;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int q = (int)(row % a.Nq);
  const int h = (int)((row / a.Nq) % a.H);
  const int b = (int)(row / ((long long)a.Nq * a.H));
  const long long so = a.part_stride_O ? a.part_stride_O : rows * a.dv;
  const long long sm = a.part_stride_ml ? a.part_stride_ml : rows;
  const bool keep = a.row_keep ? a.row_keep[(long long)b * a.stride_rk + q] != 0 : true;
  // Part p: either a slice of one local buffer, or rank p's packed partial in peer-mapped memory (plain loads: peer
  // lines are not cached in the local L2, and the L1 was invalidated when this kernel was launched).
  constexpr int MAXP = 16;
  const float* Op[MAXP];
  const float* mp_[MAXP];
  const float* lp_[MAXP];
  const int parts = a.parts;
#pragma unroll
  for (int p = 0; p < MAXP; ++p) {
    if (p < parts) {
      if (a.part_ptrs) {
        const float* base = a.part_ptrs[p];
        Op[p] = base + row * a.dv;
        mp_[p] = base + rows * a.dv + row;
        lp_[p] = base + rows * a.dv + rows + row;
      } else {
        Op[p] = a.O_part + p * so + row * a.dv;
        mp_[p] = a.m_part + p * sm + row;
        lp_[p] = a.l_part + p * sm + row;
      }
    }
  }
  float w[MAXP];
  float M = -INFINITY;
#pragma unroll
  for (int p = 0; p < MAXP; ++p)
    if (p < parts) {
      w[p] = *mp_[p];
      M = fmaxf(M, w[p]);
    }
  float L = 0.f;
#pragma unroll
  for (int p = 0; p < MAXP; ++p)
    if (p < parts) {
      w[p] = (M != -INFINITY && w[p] != -INFINITY) ? __expf(w[p] - M) : 0.f;
      if (w[p] != 0.f) L += *lp_[p] * w[p];
    }
  // whether this query row saw any valid key on any part (the reference wipes rows that did not, :168-175); the key
  // mask is shared by the heads, so head 0 speaks for the row
  if (a.row_alive && h == 0 && lane == 0)
    a.row_alive[(long long)b * a.stride_ra + q] = (keep && M != -INFINITY && L > 0.f) ? 1 : 0;
  if (a.O_out_part) {  // merged, still un-normalised partial (referenced to M)
    if (lane == 0) {
      a.m_out[row] = M;
      a.l_out[row] = L;
    }
    for (int c = lane; c < a.dv; c += 32) {
      float acc = 0.f;
#pragma unroll
      for (int p = 0; p < MAXP; ++p)
        if (p < parts && w[p] != 0.f) acc += Op[p][c] * w[p];
      a.O_out_part[row * a.dv + c] = acc;
    }
  }
  if (a.O) {
    __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(a.O) + (long long)b * a.strideO + (long long)q * a.ldo +
                       (long long)h * a.dv;
    if (!keep || M == -INFINITY || !(L > 0.f)) {
      for (int c = lane; c < a.dv; c += 32) o[c] = __float2bfloat16_rn(0.f);
      return;
    }
    const float inv = 1.0f / L;
    for (int c = lane; c < a.dv; c += 32) {
      float acc = 0.f;
#pragma unroll
      for (int p = 0; p < MAXP; ++p)
        if (p < parts && w[p] != 0.f) acc += Op[p][c] * w[p];
      reinterpret_cast<uint16_t*>(o)[c] = cvt16(acc * inv, a.fp16);
    }
  }
}

// More than 16 parts (deep local key splits): the general loop over one local buffer.
__global__ void __launch_bounds__(256) pio_combine_many_kernel(pio_combine_args a) {
  pdl_sync();
  const int lane = threadIdx.x & 31;
  const long long rows = (long long)a.B * a.H * a.Nq;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int q = (int)(row % a.Nq);
  const int h = (int)((row / a.Nq) % a.H);
  const int b = (int)(row / ((long long)a.Nq * a.H));
  const long long so = a.part_stride_O ? a.part_stride_O : rows * a.dv;
  const long long sm = a.part_stride_ml ? a.part_stride_ml : rows;
  const bool keep = a.row_keep ? a.row_keep[(long long)b * a.stride_rk + q] != 0 : true;
  float M = -INFINITY;
  for (int p = 0; p < a.parts; ++p) M = fmaxf(M, __ldg(a.m_part + p * sm + row));
  float L = 0.f;
  if (M != -INFINITY) {
    for (int p = 0; p < a.parts; ++p) {
      const float mp = __ldg(a.m_part + p * sm + row);
      if (mp != -INFINITY) L += __ldg(a.l_part + p * sm + row) * __expf(mp - M);
    }
  }
  if (a.row_alive && h == 0 && lane == 0)
    a.row_alive[(long long)b * a.stride_ra + q] = (keep && M != -INFINITY && L > 0.f) ? 1 : 0;
  if (a.O_out_part) {
    if (lane == 0) {
      a.m_out[row] = M;
      a.l_out[row] = L;
    }
    for (int c = lane; c < a.dv; c += 32) {
      float acc = 0.f;
      if (M != -INFINITY) {
        for (int p = 0; p < a.parts; ++p) {
          const float mp = __ldg(a.m_part + p * sm + row);
          if (mp != -INFINITY) acc += __ldg(a.O_part + p * so + row * a.dv + c) * __expf(mp - M);
        }
      }
      a.O_out_part[row * a.dv + c] = acc;
    }
  }
  if (a.O) {
    __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(a.O) + (long long)b * a.strideO + (long long)q * a.ldo +
                       (long long)h * a.dv;
    if (!keep || M == -INFINITY || !(L > 0.f)) {
      for (int c = lane; c < a.dv; c += 32) o[c] = __float2bfloat16_rn(0.f);
      return;
    }
    const float inv = 1.0f / L;
    for (int c = lane; c < a.dv; c += 32) {
      float acc = 0.f;
      for (int p = 0; p < a.parts; ++p) {
        const float mp = __ldg(a.m_part + p * sm + row);
        if (mp != -INFINITY) acc += __ldg(a.O_part + p * so + row * a.dv + c) * __expf(mp - M);
      }
      reinterpret_cast<uint16_t*>(o)[c] = cvt16(acc * inv, a.fp16);
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// fp32 linear with a very narrow output (N <= 16): one warp per row, lane l owns columns l, l+32, ... of x (coalesced),
// NOUT accumulators per lane, butterfly reduction.  W (N x K fp32, a few KB) is read through the read-only path.
// ------------------------------------------------------------------------------------------------------------
template <int NOUT>
__global__ void __launch_bounds__(256) pio_linear_f32_kernel(pio_linear_f32_args a) {
  pdl_sync();
  constexpr int ROWS = 4;   // rows per warp pass: their loads are issued together (the kernel is bound by loads in flight)
  const int lane = threadIdx.x & 31;
  const long long row0 = ((long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * ROWS;
  if (row0 >= a.M) return;
  float acc[ROWS][NOUT];
#pragma unroll
  for (int r = 0; r < ROWS; ++r)
#pragma unroll
    for (int n = 0; n < NOUT; ++n) acc[r][n] = 0.f;
  // rows beyond M are clamped to the last row (their results are not stored)
  const float* xr[ROWS];
#pragma unroll
  for (int r = 0; r < ROWS; ++r) xr[r] = a.x + (row0 + r < a.M ? row0 + r : a.M - 1) * a.ldx;
  int kdone = 0;
  if (((a.ldx | a.ldw) & 3) == 0 && ((reinterpret_cast<uintptr_t>(a.x) | reinterpret_cast<uintptr_t>(a.w)) & 15u) == 0) {
    // 16-byte loads: lane l owns the float4s l, l + 32, ... of each row (rows and weight rows are 16-byte aligned)
    const int nk4 = a.K >> 2;
#pragma unroll 2
    for (int i = lane; i < nk4; i += 32) {
      float4 xv[ROWS];
#pragma unroll
      for (int r = 0; r < ROWS; ++r) xv[r] = __ldg(reinterpret_cast<const float4*>(xr[r]) + i);
#pragma unroll
      for (int n = 0; n < NOUT; ++n) {
        if (n < a.N) {
          const float4 wv = __ldg(reinterpret_cast<const float4*>(a.w + (long long)n * a.ldw) + i);
#pragma unroll
          for (int r = 0; r < ROWS; ++r)
            acc[r][n] = fmaf(xv[r].x, wv.x, fmaf(xv[r].y, wv.y, fmaf(xv[r].z, wv.z, fmaf(xv[r].w, wv.w, acc[r][n]))));
        }
      }
    }
    kdone = nk4 << 2;
  }
  for (int k = kdone + lane; k < a.K; k += 32) {
    float xv[ROWS];
#pragma unroll
    for (int r = 0; r < ROWS; ++r) xv[r] = __ldg(xr[r] + k);
#pragma unroll
    for (int n = 0; n < NOUT; ++n) {
      if (n < a.N) {
        const float wv = __ldg(a.w + (long long)n * a.ldw + k);
#pragma unroll
        for (int r = 0; r < ROWS; ++r) acc[r][n] = fmaf(xv[r], wv, acc[r][n]);
      }
    }
  }
  if (a.x2 != nullptr) {
    // second operand: 16-bit rows times their own fp32 weights (the decoder tail: final_layer folded into fc2)
    const uint16_t* hr[ROWS];
#pragma unroll
    for (int r = 0; r < ROWS; ++r)
      hr[r] = reinterpret_cast<const uint16_t*>(a.x2) + (row0 + r < a.M ? row0 + r : a.M - 1) * a.ldx2;
    int k2done = 0;
    if ((a.ldx2 & 7) == 0 && (a.ldw2 & 3) == 0 &&
        ((reinterpret_cast<uintptr_t>(a.x2) | reinterpret_cast<uintptr_t>(a.w2)) & 15u) == 0) {
      const int nk8 = a.K2 >> 3;
#pragma unroll 2
      for (int i = lane; i < nk8; i += 32) {
        uint4 hv[ROWS];
#pragma unroll
        for (int r = 0; r < ROWS; ++r) hv[r] = __ldg(reinterpret_cast<const uint4*>(hr[r]) + i);
#pragma unroll
        for (int n = 0; n < NOUT; ++n) {
          if (n < a.N) {
            const float4* w4 = reinterpret_cast<const float4*>(a.w2 + (long long)n * a.ldw2) + 2 * i;
            const float4 w0 = __ldg(w4), w1 = __ldg(w4 + 1);
#pragma unroll
            for (int r = 0; r < ROWS; ++r) {
              const uint32_t hw[4] = {hv[r].x, hv[r].y, hv[r].z, hv[r].w};
              float xv[8];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                xv[2 * e] = cvt16_back((uint16_t)(hw[e] & 0xffffu), a.x2_fp16);
                xv[2 * e + 1] = cvt16_back((uint16_t)(hw[e] >> 16), a.x2_fp16);
              }
              acc[r][n] = fmaf(xv[0], w0.x, fmaf(xv[1], w0.y, fmaf(xv[2], w0.z, fmaf(xv[3], w0.w, acc[r][n]))));
              acc[r][n] = fmaf(xv[4], w1.x, fmaf(xv[5], w1.y, fmaf(xv[6], w1.z, fmaf(xv[7], w1.w, acc[r][n]))));
            }
          }
        }
      }
      k2done = nk8 << 3;
    }
    for (int k = k2done + lane; k < a.K2; k += 32) {
      float xv[ROWS];
#pragma unroll
      for (int r = 0; r < ROWS; ++r) xv[r] = cvt16_back(__ldg(hr[r] + k), a.x2_fp16);
#pragma unroll
      for (int n = 0; n < NOUT; ++n) {
        if (n < a.N) {
          const float wv = __ldg(a.w2 + (long long)n * a.ldw2 + k);
#pragma unroll
          for (int r = 0; r < ROWS; ++r) acc[r][n] = fmaf(xv[r], wv, acc[r][n]);
        }
      }
    }
  }
#pragma unroll
  for (int r = 0; r < ROWS; ++r)
#pragma unroll
    for (int n = 0; n < NOUT; ++n) acc[r][n] = warp_sum(acc[r][n]);
  if (lane < a.N) {
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
      if (row0 + r < a.M) {
        float v = 0.f;
#pragma unroll
        for (int n = 0; n < NOUT; ++n)
          if (n == lane) v = acc[r][n];
        a.y[(row0 + r) * a.ldy + lane] = v + (a.bias ? __ldg(a.bias + lane) : 0.f);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Content hash of a device buffer (the encode-once latent cache keys on it): every 32-bit word is mixed with its index
// (splitmix64 finaliser) and the mixed values are SUMMED in two independent 64-bit lanes — integer addition commutes, so
// the result does not depend on the order in which threads and blocks finish.  HBM-bound: 4 bytes read per word.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long mix64(unsigned long long z) {
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

__global__ void __launch_bounds__(256) pio_hash_kernel(const uint32_t* __restrict__ x, long long n_words,
                                                      unsigned long long seed, unsigned long long* __restrict__ out) {
  unsigned long long h0 = 0, h1 = 0;
  const long long n4 = n_words >> 2;
  const uint4* x4 = reinterpret_cast<const uint4*>(x);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const uint4 v = __ldg(x4 + i);
    const unsigned long long base = seed + (unsigned long long)i * 4ull;
    const unsigned long long a = mix64((base + 0) * 0x9E3779B97F4A7C15ull ^ v.x);
    const unsigned long long b = mix64((base + 1) * 0x9E3779B97F4A7C15ull ^ v.y);
    const unsigned long long c = mix64((base + 2) * 0x9E3779B97F4A7C15ull ^ v.z);
    const unsigned long long d = mix64((base + 3) * 0x9E3779B97F4A7C15ull ^ v.w);
    h0 += a + b + c + d;
    h1 += mix64(a ^ 0xD6E8FEB86659FD93ull) + mix64(b ^ 0xD6E8FEB86659FD93ull) + mix64(c ^ 0xD6E8FEB86659FD93ull) +
          mix64(d ^ 0xD6E8FEB86659FD93ull);
  }
  if (blockIdx.x == 0 && threadIdx.x < (n_words & 3)) {   // tail words
    const long long i = (n4 << 2) + threadIdx.x;
    const unsigned long long a = mix64((seed + (unsigned long long)i) * 0x9E3779B97F4A7C15ull ^ __ldg(x + i));
    h0 += a;
    h1 += mix64(a ^ 0xD6E8FEB86659FD93ull);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    h0 += __shfl_xor_sync(0xffffffffu, h0, o);
    h1 += __shfl_xor_sync(0xffffffffu, h1, o);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(out, h0);
    atomicAdd(out + 1, h1);
  }
}

}  // namespace pio

extern "C" int pio_hash_words(const void* data, int64_t n_words, uint64_t seed, uint64_t* out2, void* stream_) {
  using namespace pio;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  PIO_REQUIRE(data && out2 && n_words >= 0, "pio_hash_words: null pointer / negative size");
  PIO_REQUIRE(aligned16(data), "pio_hash_words: data must be 16-byte aligned");
  DeviceInfo dev;
  int rc = get_device_info(&dev);
  if (rc != PIO_OK) return rc;
  long long blocks = (n_words / 4 + 255) / 256;
  if (blocks > (long long)dev.sm_count * 8) blocks = (long long)dev.sm_count * 8;
  if (blocks < 1) blocks = 1;
  {
    ProfileScope prof(KF_LAYERNORM, 0.0, 4.0 * (double)n_words, stream);
    pio_hash_kernel<<<(unsigned)blocks, 256, 0, stream>>>(reinterpret_cast<const uint32_t*>(data), n_words,
                                                          (unsigned long long)seed,
                                                          reinterpret_cast<unsigned long long*>(out2));
  }
  g_launch_count.fetch_add(1);
  PIO_CUDA_OK(cudaGetLastError());
  return PIO_OK;
}

extern "C" int pio_linear_f32(const pio_linear_f32_args* a, void* stream_) {
  using namespace pio;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  PIO_REQUIRE(a && a->x && a->w && a->y, "pio_linear_f32: null pointer");
  PIO_REQUIRE(a->M > 0 && a->N > 0 && a->K > 0, "pio_linear_f32: bad shape");
  PIO_REQUIRE(a->ldx >= a->K && a->ldw >= a->K && a->ldy >= a->N, "pio_linear_f32: bad leading dimension");
  PIO_REQUIRE(!a->x2 || (a->w2 && a->K2 > 0 && a->ldx2 >= a->K2 && a->ldw2 >= a->K2),
              "pio_linear_f32: the second operand needs its weights and leading dimensions");
  if (a->N > 16) return fail(PIO_ERR_UNSUPPORTED, "pio_linear_f32 covers N <= 16 (got %d); use pio_gemm_bf16", a->N);
  const long long blocks = (a->M + 31) / 32;   // 8 warps x 4 rows per block
  PIO_REQUIRE(blocks < (1ll << 31), "pio_linear_f32: too many rows");
  {
    ProfileScope prof(KF_LINEAR_F32, 2.0 * a->M * a->N * (double)(a->K + (a->x2 ? a->K2 : 0)),
                      (double)a->M * (4.0 * a->K + 4.0 * a->N + (a->x2 ? 2.0 * a->K2 : 0.0)), stream);
    if (a->N <= 2) pio_linear_f32_kernel<2><<<(unsigned)blocks, 256, 0, stream>>>(*a);
    else if (a->N <= 4) pio_linear_f32_kernel<4><<<(unsigned)blocks, 256, 0, stream>>>(*a);
    else if (a->N <= 8) pio_linear_f32_kernel<8><<<(unsigned)blocks, 256, 0, stream>>>(*a);
    else pio_linear_f32_kernel<16><<<(unsigned)blocks, 256, 0, stream>>>(*a);
  }
  g_launch_count.fetch_add(1);
  PIO_CUDA_OK(cudaGetLastError());
  return PIO_OK;
}

// rows per bulk group: a multiple of 4 (16-byte granularity of the flat copy) with <= ~17 KB of fp32 per stage
static inline int lnb_rows_per_group(int C, int ldy) {
  (void)ldy;
  int R = (17 * 1024 / (4 * C)) & ~3;
  if (R > 32) R = 32;
  return R;   // 0 when a row is wider than 4 KB + ...: the caller falls back to the scalar kernel
}
static inline size_t lnb_smem_bytes(int R, int C, int ldy) {
  const size_t in_pitch = ((size_t)R * C * 4 + 127) & ~(size_t)127;
  const size_t out_pitch = ((size_t)R * ldy * 2 + 127) & ~(size_t)127;
  return pio::LNB_STAGES * in_pitch + 2 * out_pitch + 64;
}

extern "C" int pio_layernorm_bf16(const pio_layernorm_args* a, void* stream_) {
  using namespace pio;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  PIO_REQUIRE(a && a->x && a->y, "pio_layernorm_bf16: null pointer");
  PIO_REQUIRE(a->rows > 0 && a->C > 0, "pio_layernorm_bf16: bad shape rows=%lld C=%d", (long long)a->rows, a->C);
  PIO_REQUIRE(a->ldy >= a->C && a->ldx >= 0, "pio_layernorm_bf16: bad leading dimension");
  PIO_REQUIRE(a->split >= 0 && a->split <= 2, "pio_layernorm_bf16: split must be 0, 1 (A side) or 2 (B side)");
  const int split = a->split;
  const long long seg = split ? a->ldy / 3 : a->ldy;
  PIO_REQUIRE(!split || (a->ldy % 24 == 0 && seg >= a->C), "pio_layernorm_bf16: split output needs ldy = 3 * pad8(C)");
  PIO_REQUIRE(seg <= 2048, "pio_layernorm_bf16: C up to 2048 supported (got %lld columns)", (long long)seg);
  PIO_REQUIRE(!(a->fp16 && split), "pio_layernorm_bf16: the split (validation) layouts are bf16 only");
  const int mode = (a->normalize ? 1 : 0) | (split == 1 ? 2 : 0) | (split == 2 ? 4 : 0) | (a->fp16 ? 8 : 0);
  DeviceInfo dev;
  int rc = get_device_info(&dev);
  if (rc != PIO_OK) return rc;
  __nv_bfloat16* y = reinterpret_cast<__nv_bfloat16*>(a->y);
  const bool bulk = !split && (a->C % 4 != 0) && a->ldx == a->C && a->ldy % 8 == 0 && a->C / 32 <= LNB_MAX_NFULL && a->ldy < 32 * (a->C / 32 + 1) + 1 && aligned16(a->x) &&
                    aligned16(a->y) && lnb_rows_per_group(a->C, (int)a->ldy) > 0 &&
                    a->rows >= lnb_rows_per_group(a->C, (int)a->ldy);
  if (bulk && a->rows % lnb_rows_per_group(a->C, (int)a->ldy) != 0) {
    // whole groups of R rows go to the bulk-DMA kernel, the (< R rows) tail to the scalar kernel
    const long long done = a->rows / lnb_rows_per_group(a->C, (int)a->ldy) * lnb_rows_per_group(a->C, (int)a->ldy);
    pio_layernorm_args part = *a;
    part.rows = done;
    rc = pio_layernorm_bf16(&part, stream_);
    if (rc != PIO_OK) return rc;
    part.x = a->x + done * a->ldx;
    part.y = y + done * a->ldy;
    part.rows = a->rows - done;
    return pio_layernorm_bf16(&part, stream_);
  }
  ProfileScope prof(KF_LAYERNORM, 0.0, (double)a->rows * (4.0 * a->C + 2.0 * a->ldy), stream);
  (void)seg;
  const bool vec = (a->C % 4 == 0) && (a->ldx % 4 == 0) && aligned16(a->x) && aligned16(a->y) &&
                   (!a->gamma || aligned16(a->gamma)) && (!a->beta || aligned16(a->beta));
  if (vec) {
    const long long blocks = (a->rows + 3) / 4;
    PIO_REQUIRE(blocks < (1ll << 31), "pio_layernorm_bf16: too many rows");
    const int need = (int)((seg / 4 + 31) / 32);
#define PIO_LNV_LAUNCH(NV)                                                                                        \
  launch_kernel(pio_layernorm_vec_kernel<NV>, dim3((unsigned)blocks), dim3(128), 0, stream, 1, a->x, (long long)a->ldx, y, \
                (long long)a->ldy, a->gamma, a->beta, (long long)a->rows, (int)a->C, mode, a->eps)
    if (need <= 2) PIO_LNV_LAUNCH(2);
    else if (need <= 4) PIO_LNV_LAUNCH(4);
    else if (need <= 8) PIO_LNV_LAUNCH(8);
    else if (need <= 12) PIO_LNV_LAUNCH(12);
    else PIO_LNV_LAUNCH(16);
#undef PIO_LNV_LAUNCH
  } else if (bulk) {
    // contiguous odd-width rows: bulk-DMA kernel over whole groups of R rows
    const int R = lnb_rows_per_group(a->C, (int)a->ldy);
    const long long ngroups = a->rows / R;
    const size_t smem = lnb_smem_bytes(R, a->C, (int)a->ldy);
    const int nfull = a->C / 32;
    const lnb_kernel_t kern = lnb_kernel_table(std::make_integer_sequence<int, LNB_MAX_NFULL + 1>{})[nfull];
    static PerDeviceOnce once;
    const cudaError_t attr_err = once.run(dev.device, [] {
      const lnb_kernel_t* table = lnb_kernel_table(std::make_integer_sequence<int, LNB_MAX_NFULL + 1>{});
      cudaError_t e = cudaSuccess;
      for (int i = 0; i <= LNB_MAX_NFULL && e == cudaSuccess; ++i)
        e = cudaFuncSetAttribute(table[i], cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
      return e;
    });
    if (attr_err != cudaSuccess)
      return fail(PIO_ERR_CUDA, "cudaFuncSetAttribute(layernorm_bulk) failed: %s", cudaGetErrorString(attr_err));
    const int ctas_per_sm = (int)(200 * 1024 / (smem + 1024)) < 8 ? (int)(200 * 1024 / (smem + 1024)) : 8;
    long long blocks = (long long)dev.sm_count * (ctas_per_sm > 0 ? ctas_per_sm : 1);
    if (blocks > ngroups) blocks = ngroups;
    PIO_CUDA_OK(launch_kernel(kern, dim3((unsigned)blocks), dim3(LNB_THREADS), smem, stream, 1, a->x, y, (int)a->ldy,
                              a->gamma, a->beta, ngroups, R, (int)a->C, mode, a->eps));
  } else {
    const int need = (int)((seg + 31) / 32);
    const int rows_per_warp = need <= 12 ? 2 : 1;
    const long long groups = (a->rows + rows_per_warp - 1) / rows_per_warp;
    long long blocks = (groups + 3) / 4;
    const long long max_blocks = (long long)dev.sm_count * 16;   // persistent-style: each warp walks many row groups
    if (blocks > max_blocks) blocks = max_blocks;
#define PIO_LN_LAUNCH(MAXV, ROWS)                                                                                      \
  do {                                                                                                                 \
    if (split)                                                                                                         \
      launch_kernel(pio_layernorm_kernel<MAXV, ROWS, true>, dim3((unsigned)blocks), dim3(128), 0, stream, 1, a->x,      \
                    (long long)a->ldx, y, (long long)a->ldy, a->gamma, a->beta, (long long)a->rows, (int)a->C, mode,   \
                    a->eps);                                                                                           \
    else                                                                                                               \
      launch_kernel(pio_layernorm_kernel<MAXV, ROWS, false>, dim3((unsigned)blocks), dim3(128), 0, stream, 1, a->x,     \
                    (long long)a->ldx, y, (long long)a->ldy, a->gamma, a->beta, (long long)a->rows, (int)a->C, mode,   \
                    a->eps);                                                                                           \
  } while (0)
    if (need <= 4) PIO_LN_LAUNCH(4, 2);
    else if (need <= 12) PIO_LN_LAUNCH(12, 2);
    else if (need <= 24) PIO_LN_LAUNCH(24, 1);
    else if (need <= 40) PIO_LN_LAUNCH(40, 1);
    else PIO_LN_LAUNCH(64, 1);
#undef PIO_LN_LAUNCH
  }
  g_launch_count.fetch_add(1);
  PIO_CUDA_OK(cudaGetLastError());
  return PIO_OK;
}

extern "C" int pio_layernorm_concat_bf16(const pio_layernorm_concat_args* a, void* stream_) {
  using namespace pio;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  PIO_REQUIRE(a && a->feat && a->pos && a->y, "pio_layernorm_concat_bf16: null pointer");
  PIO_REQUIRE(a->B > 0 && a->N > 0 && a->Cf > 0 && a->Cp > 0, "pio_layernorm_concat_bf16: bad shape");
  const int C = a->Cf + a->Cp;
  PIO_REQUIRE(a->ldy == (C + 7) / 8 * 8, "pio_layernorm_concat_bf16: ldy must be pad8(Cf + Cp)");
  PIO_REQUIRE(C / 64 <= LNC_MAX_NP && a->ldy <= 64 * (C / 64 + 1),
              "pio_layernorm_concat_bf16: at most %d channels", 64 * (LNC_MAX_NP + 1) - 8);
  PIO_REQUIRE(a->N % 4 == 0, "pio_layernorm_concat_bf16: the number of positions must be a multiple of 4 (got %d)", a->N);
  PIO_REQUIRE(aligned16(a->pos) && aligned16(a->y), "pio_layernorm_concat_bf16: pos / y must be 16-byte aligned");
  DeviceInfo dev;
  int rc = get_device_info(&dev);
  if (rc != PIO_OK) return rc;
  // rows per position group: divides N, at most 16 (four per warp), table stage <= ~17 KB, feature gather <= 48 KB
  int R = 16;
  while (R > 4 && (a->N % R != 0 || (long long)R * a->Cp * 4 > 17 * 1024 ||
                   (long long)a->B * R * a->Cf * 4 > LNC_MAX_FEAT_BYTES))
    R -= 4;
  PIO_REQUIRE(a->N % R == 0 && (long long)a->B * R * a->Cf * 4 <= LNC_MAX_FEAT_BYTES && (long long)R * a->Cp * 4 <= 40 * 1024,
              "pio_layernorm_concat_bf16: batch x features (%d x %d) or table width (%d) too large for one position group",
              a->B, a->Cf, a->Cp);
  const int ngroups = a->N / R;
  const size_t in_pitch = ((size_t)R * a->Cp * 4 + 127) & ~(size_t)127;
  const size_t feat_bytes = ((size_t)a->B * R * a->Cf * 4 + 127) & ~(size_t)127;
  const size_t smem = LNC_STAGES * in_pitch + feat_bytes + (size_t)a->B * R * 8 + 64;
  PIO_REQUIRE(smem <= 200 * 1024, "pio_layernorm_concat_bf16: shared memory budget exceeded (%zu bytes)", smem);
  const lnc_kernel_t kern = lnc_kernel_table(std::make_integer_sequence<int, LNC_MAX_NP + 1>{})[C / 64];
  static PerDeviceOnce once;
  const cudaError_t attr_err = once.run(dev.device, [] {
    const lnc_kernel_t* table = lnc_kernel_table(std::make_integer_sequence<int, LNC_MAX_NP + 1>{});
    cudaError_t e = cudaSuccess;
    for (int i = 0; i <= LNC_MAX_NP && e == cudaSuccess; ++i)
      e = cudaFuncSetAttribute(table[i], cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    return e;
  });
  if (attr_err != cudaSuccess)
    return fail(PIO_ERR_CUDA, "cudaFuncSetAttribute(layernorm_concat) failed: %s", cudaGetErrorString(attr_err));
  int ctas_per_sm = (int)(200 * 1024 / (smem + 1024));
  if (ctas_per_sm > 8) ctas_per_sm = 8;
  if (ctas_per_sm < 1) ctas_per_sm = 1;
  long long blocks = (long long)dev.sm_count * ctas_per_sm;
  if (blocks > ngroups) blocks = ngroups;
  {
    ProfileScope prof(KF_LAYERNORM, 0.0,
                      (double)a->B * a->N * (4.0 * a->Cf + 2.0 * a->ldy) + (double)a->N * a->Cp * 4.0, stream);
    PIO_CUDA_OK(launch_kernel(kern, dim3((unsigned)blocks), dim3(LNC_THREADS), smem, stream, 1, *a, R, ngroups));
  }
  g_launch_count.fetch_add(1);
  PIO_CUDA_OK(cudaGetLastError());
  return PIO_OK;
}

extern "C" int pio_softmax_bf16(const pio_softmax_args* a, void* stream_) {
  using namespace pio;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  PIO_REQUIRE(a && a->S && a->P, "pio_softmax_bf16: null pointer");
  PIO_REQUIRE(a->batch > 0 && a->rows > 0 && a->cols > 0 && a->lds >= a->cols &&
                  (a->split ? (a->ldp % 3 == 0 && a->ldp / 3 >= a->cols) : a->ldp >= a->cols),
              "pio_softmax_bf16: bad shape");
  const long long blocks = (long long)a->batch * a->rows;
  PIO_REQUIRE(blocks < (1ll << 31), "pio_softmax_bf16: too many rows");
  {
    ProfileScope prof(KF_SOFTMAX, 0.0, (double)blocks * (4.0 * a->cols + 2.0 * a->ldp), stream);
    const bool warp_rows = !a->split && !a->dense_mask && !a->bias && !a->P_f32 && a->ldp <= 2048 && a->lds % 4 == 0 && a->strideS % 4 == 0 && a->ldp % 4 == 0 &&
                           a->strideP % 4 == 0 && aligned16(a->S) && (reinterpret_cast<uintptr_t>(a->P) & 7u) == 0;
    if (warp_rows) {
      const unsigned wblocks = (unsigned)((blocks + 7) / 8);
      const int need = (int)((a->ldp / 4 + 31) / 32);
      if (need <= 2) pio_softmax_warp_kernel<2><<<wblocks, 256, 0, stream>>>(*a);
      else if (need <= 4) pio_softmax_warp_kernel<4><<<wblocks, 256, 0, stream>>>(*a);
      else if (need <= 8) pio_softmax_warp_kernel<8><<<wblocks, 256, 0, stream>>>(*a);
      else pio_softmax_warp_kernel<16><<<wblocks, 256, 0, stream>>>(*a);
    } else if (!a->split && !a->dense_mask && !a->bias && !a->P_f32 && a->lds % 4 == 0 && a->strideS % 4 == 0 &&
               a->ldp % 4 == 0 && a->strideP % 4 == 0 && aligned16(a->S) && (reinterpret_cast<uintptr_t>(a->P) & 7u) == 0) {
      pio_softmax_long_kernel<<<(unsigned)blocks, 512, 0, stream>>>(*a);
    } else {
      pio_softmax_kernel<<<(unsigned)blocks, 256, 0, stream>>>(*a);
    }
  }
  g_launch_count.fetch_add(1);
  PIO_CUDA_OK(cudaGetLastError());
  return PIO_OK;
}

extern "C" int pio_attention_combine(const pio_combine_args* a, void* stream_) {
  using namespace pio;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  PIO_REQUIRE(a && (a->part_ptrs || (a->O_part && a->m_part && a->l_part)), "pio_attention_combine: null pointer");
  PIO_REQUIRE(!a->part_ptrs || a->parts <= 16, "pio_attention_combine: at most 16 peer parts (got %d)", a->parts);
  PIO_REQUIRE(a->O || (a->O_out_part && a->m_out && a->l_out), "pio_attention_combine: no output");
  PIO_REQUIRE(a->parts > 0 && a->B > 0 && a->H > 0 && a->Nq > 0 && a->dv > 0, "pio_attention_combine: bad shape");
  const long long rows = (long long)a->B * a->H * a->Nq;
  const long long blocks = (rows + 7) / 8;
  {
    ProfileScope prof(KF_COMBINE, 0.0, (double)rows * a->parts * (a->dv + 2) * 4.0 + (double)rows * a->dv * 2.0, stream);
    if (a->parts <= 16) pio_combine_kernel<<<(unsigned)blocks, 256, 0, stream>>>(*a);
    else pio_combine_many_kernel<<<(unsigned)blocks, 256, 0, stream>>>(*a);
  }
  g_launch_count.fetch_add(1);
  PIO_CUDA_OK(cudaGetLastError());
  return PIO_OK;
}
