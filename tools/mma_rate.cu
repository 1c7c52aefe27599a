// Developer aid: how long does ONE tcgen05.mma (kind::f16, M = 128, K = 16, cta_group::1) occupy the tensor pipe of an
// sm_100a SM as a function of N, of where the A operand lives (shared memory / TMEM) and of what the other warps do?
// One CTA per SM issues R back-to-back MMAs from one thread over a ring of SWIZZLE_128B operand stages, commits, waits.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I perceiverio_pytorch_b200/csrc tools/mma_rate.cu -o /tmp/mma_rate && /tmp/mma_rate
#include "pio_common.cuh"

using namespace pio;

struct Params { int N, R, a_tmem, pollers, accs, ctas_report, M; };

template <int A_TMEM, int ACCS>
__global__ void __launch_bounds__(384, 1) mma_rate_kernel(Params p, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t done_bar, never_bar;
  __shared__ uint32_t tmem_slot;
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  constexpr int STAGES = 4, A_BYTES = 128 * 128, B_BYTES = 256 * 128;
  for (int i = threadIdx.x; i < STAGES * (A_BYTES + B_BYTES) / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;   // finite values in either 16-bit format
  if (threadIdx.x == 0) {
    mbar_init(&done_bar, 1);
    mbar_init(&never_bar, 1);
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(&tmem_slot, 512);
    tmem_relinquish();
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (warp == 1) {
    // warp-converged issue, one elected lane per instruction, descriptors as 32-bit low words advanced by constants, the
    // TMEM base taken as 0 (the first 512-column allocation of an SM): everything stays on the uniform datapath
    const uint32_t idesc = make_idesc_f16(p.M, p.N, 1, 0, 0);
    const uint64_t da0 = make_smem_desc_sw128(smem_u32(smem), 16, 1024);
    const uint64_t db0 = make_smem_desc_sw128(smem_u32(smem) + A_BYTES, 16, 1024);
    const uint32_t a_lo0 = (uint32_t)da0, a_hi = (uint32_t)(da0 >> 32), b_lo0 = (uint32_t)db0, b_hi = (uint32_t)(db0 >> 32);
        const long long t0 = clock64();
    // identical operands for every MMA of the burst (the timing does not depend on the data): the loop body is 16
    // UTCHMMA instructions with nothing between them
    const uint32_t dA = 0u, dB = ACCS > 1 ? (uint32_t)p.N : 0u;
    for (int r = 0; r < p.R; r += 16) {
      if (elect_one()) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const uint32_t d = (ACCS == 3 && (i & 1)) || (ACCS == 2 && (i & 4)) ? dB : dA;
          if (A_TMEM) umma_ts_lh(d, 256, b_lo0, b_hi, idesc, 1u);
          else umma_ss_lh(d, a_lo0, a_hi, b_lo0, b_hi, idesc, 1u);
        }
      }
      __syncwarp();
    }
    const long long t1 = clock64();
    if (elect_one()) umma_commit(&done_bar);
    mbar_wait(&done_bar, 0);
    const long long t2 = clock64();
    if (blockIdx.x < p.ctas_report && lane == 0) {
      out[blockIdx.x * 2] = t1 - t0;
      out[blockIdx.x * 2 + 1] = t2 - t0;
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&never_bar);     // releases the pollers
  } else if (warp >= 4 && p.pollers) {
    // what idle epilogue warps do: every thread polls an mbarrier that completes when the MMAs are done
    if (p.pollers == 1 || lane == 0) mbar_wait(&never_bar, 0);
  }
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

int main() {
  long long* out;
  cudaMalloc(&out, 64 * sizeof(long long));
  const int smem = 4 * (128 * 128 + 256 * 128) + 1024;
  printf("tcgen05.mma kind::f16 M=128 K=16 cta_group::1, R = 512 back-to-back MMAs from one thread, 148 CTAs, 256 threads polling an mbarrier (clk per MMA: issue loop / until the commit arrives)\n");
  const int Ns[] = {32, 64, 128, 256};
  for (int M : {128})
    for (int a_tmem : {0, 1})
      for (int accs : {1, 2, 3})
        for (int N : Ns) {
          if ((accs > 1 ? 2 : 1) * N > 256) continue;
          Params p{N, 512, a_tmem, 1, accs, 1, M};
          long long h[2] = {0, 0};
          void (*kern)(Params, long long*) = nullptr;
          if (a_tmem == 0) kern = accs == 1 ? mma_rate_kernel<0, 1> : (accs == 2 ? mma_rate_kernel<0, 2> : mma_rate_kernel<0, 3>);
          else kern = accs == 1 ? mma_rate_kernel<1, 1> : (accs == 2 ? mma_rate_kernel<1, 2> : mma_rate_kernel<1, 3>);
          cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
          for (int rep = 0; rep < 2; ++rep) {
            kern<<<148, 384, smem>>>(p, out);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
          }
          cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
          printf("M %3d  A in %s  accumulator pattern %d  N %3d : issue %6.1f  complete %6.1f clk/MMA\n", M,
                 a_tmem ? "TMEM" : "smem", accs, N, (double)h[0] / p.R, (double)h[1] / p.R);
        }
  return 0;
}
