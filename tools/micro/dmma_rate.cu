// dmma_rate.cu — issue-rate microbenchmark behind the C5 kernel design:
// fp64 mma.sync.m8n8k4 against plain DFMA on one B200, 148 CTAs x W warps, independent chains.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/dmma_rate tools/micro/dmma_rate.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int kChains>
__global__ void k_dmma(double* out, int iters) {
  double c[kChains][2];
  for (int i = 0; i < kChains; ++i) c[i][0] = c[i][1] = threadIdx.x * 1e-9;
  double a = 1.0 + threadIdx.x * 1e-12, b = 1.0 - threadIdx.x * 1e-12;
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int i = 0; i < kChains; ++i) dmma(c[i][0], c[i][1], a, b);
  double s = 0;
  for (int i = 0; i < kChains; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int kChains>
__global__ void k_dfma(double* out, int iters) {
  double c[kChains];
  for (int i = 0; i < kChains; ++i) c[i] = threadIdx.x * 1e-9;
  double a = 1.0 + threadIdx.x * 1e-12, b = 1e-9;
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int i = 0; i < kChains; ++i) c[i] = fma(c[i], a, b);
  double s = 0;
  for (int i = 0; i < kChains; ++i) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
float time_it(F f) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  return ms;
}

int main() {
  double* out; cudaMalloc(&out, 148 * 1024 * 8 * 8);
  int clk_khz; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  const int iters = 20000;
  for (int warps : {1, 2, 4, 8, 16, 32}) {
    float a = time_it([&] { k_dmma<1><<<148, warps * 32>>>(out, iters); });
    float b = time_it([&] { k_dmma<4><<<148, warps * 32>>>(out, iters); });
    float c = time_it([&] { k_dfma<1><<<148, warps * 32>>>(out, iters); });
    float d = time_it([&] { k_dfma<8><<<148, warps * 32>>>(out, iters); });
    auto tf = [&](double flops, float ms) { return flops / (ms * 1e-3) / 1e12; };
    const double nw = 148.0 * warps * iters;
    printf("warps/SM %2d | dmma x1 %.2f TF (%.1f clk/dep-mma) x4 %.2f TF | dfma x1 %.2f TF (%.1f clk/dep-fma) x8 %.2f TF\n", warps,
           tf(nw * 512, a), a * 1e-3 * clk_khz * 1e3 / iters, tf(nw * 4 * 512, b), tf(nw * 64, c),
           c * 1e-3 * clk_khz * 1e3 / iters, tf(nw * 8 * 64, d));
  }
  return 0;
}
