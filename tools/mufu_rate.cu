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

// MODE 1: cvt.rn.bf16x2.f32 only; MODE 2: per pair of values two ex2 and one cvt (the softmax loop's mix);
// MODE 3: two ex2 and the integer form of a truncating bf16 pack (2 x LOP3 + PRMT)
template <int MODE>
__global__ void mix_kernel(float* out, long long* clk, int iters) {
  float v[16];
  unsigned w[8];
  unsigned long long acc2 = 0;
  float accs = 0.001f, accs2 = 0.002f;
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = 0.001f * (threadIdx.x + i);
#pragma unroll
  for (int i = 0; i < 8; ++i) w[i] = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE != 1 && MODE != 6) {
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(v[2 * i]));
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(v[2 * i + 1]));
      }
      if (MODE == 1 || MODE == 2) {
        unsigned r;
        asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(v[2 * i + 1]), "f"(v[2 * i]));
        w[i] ^= r;
      }
      if (MODE == 4 || MODE == 5) {      // the softmax loop's full mix: scale-subtract, 2 x ex2, convert, row sum
        unsigned r;
        asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(v[2 * i + 1]), "f"(v[2 * i]));
        w[i] ^= r;
        if (MODE == 4) {
          unsigned long long t, u;
          asm volatile("mov.b64 %0, {%1, %2};" : "=l"(t) : "f"(v[2 * i]), "f"(v[2 * i + 1]));
          asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(t) : "l"(acc2));
          asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(acc2) : "l"(t));
          asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(v[2 * i]), "=f"(v[2 * i + 1]) : "l"(t));
          (void)u;
        } else {
          asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(v[2 * i]) : "f"(accs));
          asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(v[2 * i + 1]) : "f"(accs));
          asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(accs) : "f"(v[2 * i]));
          asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(accs2) : "f"(v[2 * i + 1]));
        }
      }
      if (MODE == 3) {
        unsigned a = __float_as_uint(v[2 * i]), b = __float_as_uint(v[2 * i + 1]), r;
        asm volatile("prmt.b32 %0, %1, %2, 0x7632;" : "=r"(r) : "r"(a), "r"(b));
        w[i] ^= r;
      }
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += v[i];
#pragma unroll
  for (int i = 0; i < 8; ++i) s += __uint_as_float(w[i]);
  s += accs + accs2 + __uint_as_float((unsigned)acc2);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run_mix(int threads, const char* what) {
  float* out;
  long long* clk;
  cudaMalloc(&out, 148 * 1024 * 4);
  cudaMalloc(&clk, 148 * 8);
  const int iters = 2000;
  mix_kernel<MODE><<<148, threads>>>(out, clk, iters);
  mix_kernel<MODE><<<148, threads>>>(out, clk, iters);
  cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0;
  for (int i = 0; i < 148; ++i) avg += h[i];
  avg /= 148;
  printf("%-60s threads/CTA %4d : %6.2f clk per PAIR of values per warp\n", what, threads, avg / (iters * 8.0));
  cudaFree(out);
  cudaFree(clk);
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
  run_mix<1>(128, "cvt.rn.bf16x2.f32 only, one warp per scheduler");
  run_mix<1>(256, "cvt.rn.bf16x2.f32 only, two warps per scheduler");
  run_mix<2>(128, "2 x ex2 + cvt.rn.bf16x2 per pair, one warp per scheduler");
  run_mix<2>(256, "2 x ex2 + cvt.rn.bf16x2 per pair, two warps per scheduler");
  run_mix<4>(128, "2 x ex2 + cvt + fma.f32x2 + add.f32x2 per pair, one warp");
  run_mix<4>(256, "2 x ex2 + cvt + fma.f32x2 + add.f32x2 per pair, two warps");
  run_mix<5>(128, "2 x ex2 + cvt + 2 fma.f32 + 2 add.f32 per pair, one warp");
  run_mix<5>(256, "2 x ex2 + cvt + 2 fma.f32 + 2 add.f32 per pair, two warps");
  run_mix<3>(128, "2 x ex2 + prmt (truncating pack) per pair, one warp");
  run_mix<3>(256, "2 x ex2 + prmt (truncating pack) per pair, two warps");
  return 0;
}
