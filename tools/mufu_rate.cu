// Developer aid: how fast does ONE warp per scheduler get through a dense run of MUFU.EX2, alone and with a second warp
// on the same scheduler?  (The softmax warps of the attention kernels are one warp per scheduler and tile.)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mufu_rate tools/mufu_rate.cu && /tmp/mufu_rate
#include <cstdio>
#include <cuda_runtime.h>

template <int CHAINS, int EXTRA>
__global__ void mufu_kernel(float* out, long long* clk, int iters) {
  float v[CHAINS];
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) v[i] = 0.001f * (threadIdx.x + i);
  float acc = 0.f;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(v[i]));
#pragma unroll
    for (int e = 0; e < EXTRA; ++e) asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(acc) : "f"(v[e % CHAINS]));
  }
  const long long t1 = clock64();
  float s = acc;
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) s += v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

template <int CHAINS, int EXTRA>
void run(int threads, const char* what) {
  float* out;
  long long* clk;
  cudaMalloc(&out, 148 * 1024 * 4);
  cudaMalloc(&clk, 148 * 8);
  const int iters = 2000;
  mufu_kernel<CHAINS, EXTRA><<<148, threads>>>(out, clk, iters);
  mufu_kernel<CHAINS, EXTRA><<<148, threads>>>(out, clk, iters);
  cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0;
  for (int i = 0; i < 148; ++i) avg += h[i];
  avg /= 148;
  const double per = avg / (iters * (double)CHAINS);
  printf("%-44s threads/CTA %4d  chains %2d  extra FMA %d : %6.2f clk per MUFU.EX2 warp-instruction per warp, %5.2f lanes/clk/SM\n",
         what, threads, CHAINS, EXTRA, per, (threads / 32) * 32.0 / per);
  cudaFree(out);
  cudaFree(clk);
}

int main() {
  run<1, 0>(128, "one warp per scheduler, dependent chain");
  run<8, 0>(128, "one warp per scheduler, 8 independent");
  run<16, 0>(128, "one warp per scheduler, 16 independent");
  run<16, 0>(256, "two warps per scheduler, 16 independent");
  run<16, 0>(512, "four warps per scheduler, 16 independent");
  run<16, 8>(128, "one warp per scheduler, 16 + 8 FMA");
  run<16, 8>(256, "two warps per scheduler, 16 + 8 FMA");
  return 0;
}
