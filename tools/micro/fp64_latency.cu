// Micro-benchmark (B200): latency of dependent FP64 operations and how it changes with the number of active lanes
// and of warps per SM sub-partition.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o fp64_latency fp64_latency.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int OP>
__global__ void chain(double* out, const double* in, int iters, int activeLanes, long long* cycles) {
  int lane = threadIdx.x & 31;
  double a = in[0], b = in[1], x = in[2] + threadIdx.x;
  long long t0 = 0, t1 = 0;
  if (lane < activeLanes) {
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < iters; i++) {
#pragma unroll
      for (int k = 0; k < 16; k++) {
        if (OP == 0) x = x + a;                 // DADD
        if (OP == 1) x = x * b;                 // DMUL
        if (OP == 2) x = __fma_rn(x, b, a);     // DFMA
        if (OP == 3) x = fmax(x * b, a);        // DMUL + DMNMX
      }
    }
    t1 = clock64();
  }
  if (lane < activeLanes) out[blockIdx.x * blockDim.x + threadIdx.x] = x;
  if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
}
int main() {
  double *in, *out; long long* cyc;
  cudaMalloc(&in, 64); cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 64);
  double h[3] = {1e-9, 1.0000001, 1.0};
  cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
  const char* names[4] = {"DADD", "DMUL", "DFMA", "DMUL+DMNMX"};
  int iters = 2000;
  for (int op = 0; op < 4; op++)
    for (int threads : {32, 128, 256, 512, 1024})
      for (int lanes : {32, 16, 1}) {
        for (int rep = 0; rep < 2; rep++) {
          if (op == 0) chain<0><<<1, threads>>>(out, in, iters, lanes, cyc);
          if (op == 1) chain<1><<<1, threads>>>(out, in, iters, lanes, cyc);
          if (op == 2) chain<2><<<1, threads>>>(out, in, iters, lanes, cyc);
          if (op == 3) chain<3><<<1, threads>>>(out, in, iters, lanes, cyc);
          cudaDeviceSynchronize();
        }
        long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
        printf("%-11s warps/SM=%2d (per SMSP %.1f) active lanes=%2d : %.2f cycles per dependent op-group\n", names[op], threads / 32,
               threads / 128.0, lanes, (double)c / (iters * 16.0));
      }
  return 0;
}
