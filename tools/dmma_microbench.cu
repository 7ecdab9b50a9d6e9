// Micro-benchmark of the FP64 pipes on sm_100a (B200): DMMA.8x8x4 latency / throughput as a function of
// independent chains per warp and warps per SM sub-partition, DFMA ditto, and the two mixed.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_microbench dmma_microbench.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int CH, int FMA_PER>   // CH independent DMMA chains, FMA_PER independent DFMAs per DMMA round
__global__ void k_dmma(double* out, long long* cyc, int iters) {
  double c0[CH], c1[CH], f[FMA_PER > 0 ? FMA_PER : 1];
  const double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
#pragma unroll
  for (int i = 0; i < CH; ++i) c0[i] = c1[i] = i;
#pragma unroll
  for (int i = 0; i < (FMA_PER > 0 ? FMA_PER : 1); ++i) f[i] = i;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < CH; ++i) dmma(c0[i], c1[i], a, b);
#pragma unroll
    for (int i = 0; i < FMA_PER; ++i) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(f[i]) : "d"(a), "d"(b));
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < CH; ++i) s += c0[i] + c1[i];
#pragma unroll
  for (int i = 0; i < (FMA_PER > 0 ? FMA_PER : 1); ++i) s += f[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

template <int CH>
__global__ void k_dfma(double* out, long long* cyc, int iters) {
  double f[CH];
  const double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
#pragma unroll
  for (int i = 0; i < CH; ++i) f[i] = i;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < CH; ++i) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(f[i]) : "d"(a), "d"(b));
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < CH; ++i) s += f[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

template <typename K>
void run(const char* name, K kern, int warps_per_sm, int per_iter_dmma, int per_iter_fma, double* out, long long* cyc) {
  const int iters = 4096;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  kern<<<148, warps_per_sm * 32>>>(out, cyc, 16);
  cudaEventRecord(e0);
  kern<<<148, warps_per_sm * 32>>>(out, cyc, iters);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  double cyc_per_iter = (double)c / iters;
  double tf = (148.0 * warps_per_sm * iters * (per_iter_dmma * 512.0 + per_iter_fma * 64.0)) / (ms * 1e-3) / 1e12;
  printf("%-28s warps/SM %2d  cycles/iter %8.1f  cyc/dmma/smsp %6.2f  %6.2f TFLOP/s\n", name, warps_per_sm, cyc_per_iter,
         per_iter_dmma ? cyc_per_iter / (per_iter_dmma * (warps_per_sm / 4.0 > 1 ? warps_per_sm / 4.0 : 1)) : 0.0, tf);
}

int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&cyc, 8);
  for (int w : {1, 4, 8, 12, 16}) {
    run("dmma 1 chain", k_dmma<1, 0>, w, 1, 0, out, cyc);
    run("dmma 2 chains", k_dmma<2, 0>, w, 2, 0, out, cyc);
    run("dmma 4 chains", k_dmma<4, 0>, w, 4, 0, out, cyc);
    run("dmma 8 chains", k_dmma<8, 0>, w, 8, 0, out, cyc);
    run("dfma 1 chain", k_dfma<1>, w, 0, 1, out, cyc);
    run("dfma 8 chains", k_dfma<8>, w, 0, 8, out, cyc);
    run("dmma 4 ch + 4 dfma", k_dmma<4, 4>, w, 4, 4, out, cyc);
    run("dmma 4 ch + 16 dfma", k_dmma<4, 16>, w, 4, 16, out, cyc);
    run("dmma 8 ch + 8 dfma", k_dmma<8, 8>, w, 8, 8, out, cyc);
  }
  return 0;
}
