// Shared device helpers for the sm_100a kernels: mbarrier, TMA, tcgen05 (MMA / TMEM) PTX wrappers and the
// UMMA shared-memory / instruction descriptor encodings.  Everything here is inline PTX; no CUTLASS.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace pio {

// ------------------------------------------------------------------------------------------------
// Generic helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31u; }

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred P1;\n\telect.sync _|P1, 0xffffffff;\n\tselp.b32 %0, 1, 0, P1;\n\t}\n"
      : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

// 16-bit operand / activation format of a launch (pio_*_args.fp16): bf16 (8-bit mantissa, fp32's range) or fp16 (11-bit
// mantissa: 8x less operand rounding, range 65504 — conversions saturate instead of producing inf).  tcgen05.mma
// kind::f16 runs both at the same rate; the format only changes the instruction descriptor and these conversions.
__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
template <bool F16>
__device__ __forceinline__ uint32_t pack16x2(float lo, float hi) {
  if constexpr (F16) return pack_f16x2(lo, hi);
  else return pack_bf16x2(lo, hi);
}
// run-time form for the kernels that are bound by memory, not by instruction issue
__device__ __forceinline__ uint32_t pack16x2(float lo, float hi, int f16) {
  return f16 ? pack_f16x2(lo, hi) : pack_bf16x2(lo, hi);
}
// the two values of a packed 16-bit pair as floats (exact)
template <bool F16>
__device__ __forceinline__ void unpack16x2(uint32_t w, float& lo, float& hi) {
  if constexpr (F16) {
    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w));
    lo = f.x;
    hi = f.y;
  } else {
    lo = __uint_as_float(w << 16);
    hi = __uint_as_float(w & 0xffff0000u);
  }
}
__device__ __forceinline__ uint16_t cvt16(float v, int f16) { return (uint16_t)(pack16x2(v, 0.f, f16) & 0xffffu); }
__device__ __forceinline__ float cvt16_back(uint16_t h, int f16) {
  return f16 ? __half2float(__ushort_as_half(h)) : __uint_as_float((uint32_t)h << 16);
}

// Packed fp32x2 arithmetic (sm_100: FFMA2 / FADD2 issue one instruction for two lanes) and 3-input max (FMNMX3).
__device__ __forceinline__ uint64_t pack_f32x2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack_f32x2(uint64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ uint64_t fadd2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}

// Exact (erf) GELU, GELU(x) = 0.5 x (1 + erf(x / sqrt 2)).  erf(z) = sign(z) (1 - 2^(-q(|z|))) with q a degree-5 polynomial
// without constant term fitted to -log2(erfc(z)) on [0, 4.2] (weighted for the absolute error of erf: 6.3e-7; beyond 4.2
// erf is 1 to 3e-9) — |abs error of GELU| < 1.3e-6 in fp32, three orders below the rounding of the stored 16-bit
// activation.  ONE MUFU op (ex2) and 9 FMA-pipe instructions per element: the epilogues that apply it are bound by the
// MUFU unit (4 lanes per clock and scheduler), and Abramowitz-Stegun 7.1.26, used before, needs rcp + ex2.
// The 1 / sqrt 2 is folded into the coefficients (q as a polynomial in |x|, |x| clamped at 4.2 sqrt 2).
__device__ __forceinline__ float ex2_approx(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float gelu_erf(float x) {
  const float a = fminf(fabsf(x), 5.9396970f);
  float q = fmaf(a, 0.0005204587359912694f, -0.007397511973977089f);
  q = fmaf(q, a, 0.052561238408088684f);
  q = fmaf(q, a, 0.4592546820640564f);
  q = fmaf(q, a, 1.1510913372039795f);
  const float erf_abs = 1.0f - ex2_approx(-(q * a));
  const float hx = 0.5f * x;
  return fmaf(hx, copysignf(erf_abs, x), hx);
}

// Two values at a time (the epilogues that apply GELU to a whole accumulator tile are bound by instruction issue): the
// polynomial, the products and the final blend as packed fp32x2 operations, 7 instead of 11 instructions per element.
__device__ __forceinline__ uint64_t gelu_erf2(uint64_t x2) {
  float x0, x1;
  unpack_f32x2(x2, x0, x1);
  const uint64_t a = pack_f32x2(fminf(fabsf(x0), 5.9396970f), fminf(fabsf(x1), 5.9396970f));
  uint64_t q = ffma2(a, pack_f32x2(0.0005204587359912694f, 0.0005204587359912694f),
                     pack_f32x2(-0.007397511973977089f, -0.007397511973977089f));
  q = ffma2(q, a, pack_f32x2(0.052561238408088684f, 0.052561238408088684f));
  q = ffma2(q, a, pack_f32x2(0.4592546820640564f, 0.4592546820640564f));
  q = ffma2(q, a, pack_f32x2(1.1510913372039795f, 1.1510913372039795f));
  float t0, t1;
  unpack_f32x2(ffma2(q, a, pack_f32x2(0.f, 0.f)), t0, t1);
  const float e0 = ex2_approx(-t0), e1 = ex2_approx(-t1);
  // erf(|x|) = 1 - e, with the sign of x
  float r0, r1;
  unpack_f32x2(ffma2(pack_f32x2(e0, e1), pack_f32x2(-1.f, -1.f), pack_f32x2(1.f, 1.f)), r0, r1);
  const uint64_t s = pack_f32x2(copysignf(r0, x0), copysignf(r1, x1));
  const uint64_t hx = ffma2(x2, pack_f32x2(0.5f, 0.5f), pack_f32x2(0.f, 0.f));
  return ffma2(hx, s, hx);
}

// ------------------------------------------------------------------------------------------------
// Programmatic dependent launch: every kernel lets the next grid in the stream start its prologue early
// (launch_dependents) and then waits for the previous grid's results before touching global memory (wait).
// Both are no-ops unless the launch carries the programmatic-serialization attribute (pio_host.h: launch_kernel).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void pdl_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_sync() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
}

// ------------------------------------------------------------------------------------------------
// mbarrier
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// One arrival per WARP: the lanes converge, lane 0 arrives for all of them (the barrier is initialised with the number of
// warps).  32 lanes arriving on one barrier are 32 serialised shared-memory atomics in front of the thread that waits.
#ifndef PIO_WARP_ARRIVE
#define PIO_WARP_ARRIVE 1
#endif
__device__ __forceinline__ void mbar_arrive_warp(uint64_t* bar) {
#if PIO_WARP_ARRIVE
  __syncwarp();
  if ((threadIdx.x & 31u) == 0) mbar_arrive(bar);
#else
  mbar_arrive(bar);
#endif
}
constexpr uint32_t kArrivalsPerWarp = PIO_WARP_ARRIVE ? 1u : 32u;
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P1;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, P1;\n\t}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug traps (reported as a launch failure through the C-ABI) instead of hanging
// the GPU box.  try_wait suspends in hardware, so the bound costs nothing on the fast path.
#ifndef PIO_MBAR_SPIN_LIMIT
#define PIO_MBAR_SPIN_LIMIT (1u << 26)
#endif
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > PIO_MBAR_SPIN_LIMIT) {
      printf("pio: mbarrier timeout block(%d,%d,%d) thread %d bar@%u parity %u\n", blockIdx.x, blockIdx.y,
             blockIdx.z, threadIdx.x, smem_u32(bar), parity);
      __trap();
    }
  }
}

// Generic-proxy writes to smem -> visible to the async proxy (TMA / tcgen05.mma operand reads).
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ------------------------------------------------------------------------------------------------
// TMA (cp.async.bulk.tensor), tile mode, arrives on an mbarrier with complete_tx::bytes
// ------------------------------------------------------------------------------------------------
constexpr uint64_t kEvictNormal = 0x1000000000000000ull;
constexpr uint64_t kEvictFirst = 0x12F0000000000000ull;
constexpr uint64_t kEvictLast = 0x14F0000000000000ull;

__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                            uint64_t hint = kEvictNormal) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1),
      "l"(hint)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                            int c2, uint64_t hint = kEvictNormal) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4, %5}], [%2], %6;"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1),
      "r"(c2), "l"(hint)
      : "memory");
}

// Multicast variant: the box lands at the same CTA-relative smem offset of every CTA in `cta_mask`, and each
// destination CTA's mbarrier (same offset) receives the complete_tx.
__device__ __forceinline__ void tma_load_3d_mcast(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                                  int c2, uint16_t cta_mask, uint64_t hint = kEvictNormal) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster.L2::cache_hint"
      " [%0], [%1, {%4, %5, %6}], [%2], %3, %7;"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "h"(cta_mask), "r"(c0),
      "r"(c1), "r"(c2), "l"(hint)
      : "memory");
}

// ------------------------------------------------------------------------------------------------
// thread-block clusters
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// ------------------------------------------------------------------------------------------------
// tcgen05: TMEM allocation, MMA issue, commit, TMEM <-> register moves
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result, uint32_t ncols) {  // whole warp, ncols pow2>=32
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // whole warp (the allocating one)
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem] * B[smem]; kind::f16 covers bf16 and fp16 inputs with fp32 accumulation.
__device__ __forceinline__ void umma_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                        uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Descriptors split into (low, high) words: only the 14-bit start-address field of the low word changes between the
// MMAs of a tile, so an issuer advances 32-bit values (and the compiler can keep them on the uniform datapath).
__device__ __forceinline__ void umma_ss_lh(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                           uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %6, 0;\n\t"
      "mov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}\n"
      ::"r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]: the A operand (bf16 pairs packed in 32-bit columns, K-major) comes from TMEM
__device__ __forceinline__ void umma_ts_lh(uint32_t tmem_d, uint32_t tmem_a, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 db;\n\tsetp.ne.b32 p, %5, 0;\n\t"
      "mov.b64 db, {%2, %3};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;\n\t}\n"
      ::"r"(tmem_d), "r"(tmem_a), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrive once all previously issued tcgen05.mma of this thread have completed
// (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// Same, arriving on the barrier at this offset in every CTA of `cta_mask` (frees a multicast-filled smem stage).
__device__ __forceinline__ void umma_commit_mcast(uint64_t* bar, uint16_t cta_mask) {
  asm volatile(
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(smem_u32(bar)), "h"(cta_mask)
      : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// 32 lanes x 32-bit, N consecutive columns: thread t of the warp reads TMEM lane (taddr.lane + t).
// A warp may only touch the lane quarter 32*(warp_id % 4) .. +31.
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
      "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]),
      "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]),
      "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
      "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}

// ------------------------------------------------------------------------------------------------
// CTA-pair (cta_group::2) variants: two CTAs of a 2-wide cluster run one M=256 MMA; each CTA stages its own 128 rows
// of A and its half of B, the leader (cluster rank 0) issues the MMAs and owns the full / tmem-empty barriers.
// ------------------------------------------------------------------------------------------------
// shared::cluster address of `p` (a shared::cta pointer of this CTA) as seen in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(const void* p, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(p)), "r"(rank));
  return r;
}
// TMA load into this CTA's smem whose completion bytes are credited to the mbarrier at cluster address `bar_cluster`
// (the leader CTA's barrier)
__device__ __forceinline__ void tma_load_3d_2cta(void* smem_dst, const CUtensorMap* m, uint32_t bar_cluster, int c0,
                                                 int c1, int c2, uint64_t hint = kEvictNormal) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4, %5}], [%2], %6;"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2),
      "l"(hint)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_cluster) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster) : "memory");
}
__device__ __forceinline__ void tmem_alloc_2cta(uint32_t* smem_result, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish_2cta() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2cta(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_ss_2cta(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Same with the descriptors split into (low, high) words: only the start-address field of the low word changes from
// one MMA to the next, so the issuer advances 32-bit values.
__device__ __forceinline__ void umma_ss_2cta_lh(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo,
                                                uint32_t b_hi, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %6, 0;\n\t"
      "mov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, p;\n\t}\n"
      ::"r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_2cta_mcast(uint64_t* bar, uint16_t cta_mask) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(smem_u32(bar)), "h"(cta_mask)
      : "memory");
}

// ------------------------------------------------------------------------------------------------
// TMA stores (smem -> global, bulk async group)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* m, const void* smem_src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
// ... with an L2 eviction-priority hint for the lines it writes (kEvictFirst / kEvictLast)
__device__ __forceinline__ void tma_store_3d_hint(const CUtensorMap* m, const void* smem_src, int c0, int c1, int c2,
                                                  uint64_t hint) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%2, %3, %4}], [%1], %5;"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2), "l"(hint)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// wait until at most N of this thread's committed bulk groups still have to READ their shared-memory source
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}

// ------------------------------------------------------------------------------------------------
// UMMA descriptors (sm_100): see DESIGN.md "descriptor encodings"
// ------------------------------------------------------------------------------------------------
// 64-bit shared-memory matrix descriptor: start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48),
// layout type [61,64) (2 = SWIZZLE_128B).
__device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFFu);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= 1ull << 46;
  d |= 2ull << 61;
  return d;
}
// 32-bit instruction descriptor for kind::f16, fp32 accumulate.  ab_fmt: 0 = fp16, 1 = bf16 (idesc_fmt(fp16 flag)).
// a_mn / b_mn: 0 = K-major operand, 1 = MN-major operand.
__host__ __device__ constexpr uint32_t make_idesc_f16(uint32_t M, uint32_t N, uint32_t ab_fmt, uint32_t a_mn,
                                                      uint32_t b_mn) {
  return (1u << 4) | (ab_fmt << 7) | (ab_fmt << 10) | (a_mn << 15) | (b_mn << 16) | ((N >> 3) << 17) |
         ((M >> 4) << 24);
}

__host__ __device__ constexpr uint32_t idesc_fmt(int fp16) { return fp16 ? 0u : 1u; }

// Byte offset of element (row r, 16-byte chunk c) inside a [rows x 128 B] SWIZZLE_128B tile whose base is
// 1024-byte aligned (the layout TMA produces and UMMA expects): chunk index XOR (row % 8).
__device__ __forceinline__ uint32_t sw128_offset(uint32_t r, uint32_t chunk16) {
  return r * 128u + ((chunk16 ^ (r & 7u)) << 4);
}

// Same for a [rows x 64 B] SWIZZLE_64B tile: chunk index (2 bits) XOR ((row / 2) % 4).
__device__ __forceinline__ uint32_t sw64_offset(uint32_t r, uint32_t chunk16) {
  return r * 64u + ((chunk16 ^ ((r >> 1) & 3u)) << 4);
}

}  // namespace pio
