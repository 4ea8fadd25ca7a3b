// Micro-benchmark (B200): how fast ONE warp issues independent FP64 instructions (ILP = number of independent dependency
// chains interleaved in the instruction stream), alone on its SM sub-partition.  Answers whether a single-warp chain
// (k_pgs_giant) is bound by the 8-cycle latency, by the FP64 pipe (2 cycles per warp instruction) or by per-warp issue.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o fp64_ilp fp64_ilp.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP, int OP>
__global__ void k(double* out, const double* in, int iters, long long* cycles) {
  double a = in[0], b = in[1];
  double x[ILP];
#pragma unroll
  for (int j = 0; j < ILP; j++) x[j] = in[2] + threadIdx.x + j;
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int r = 0; r < 8; r++) {
#pragma unroll
      for (int j = 0; j < ILP; j++) {
        if (OP == 0) x[j] = x[j] + a;
        if (OP == 1) x[j] = x[j] * b;
        if (OP == 2) x[j] = __fma_rn(x[j], b, a);
        if (OP == 3) x[j] = x[j] * b + a;   // DMUL then DADD (fmad=false): the PGS pattern
      }
    }
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int j = 0; j < ILP; j++) s += x[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
}
template <int ILP, int OP>
void run(const char* name, double* out, double* in, long long* cyc) {
  int iters = 2000;
  for (int threads : {32, 128}) {
    for (int rep = 0; rep < 2; rep++) { k<ILP, OP><<<1, threads>>>(out, in, iters, cyc); cudaDeviceSynchronize(); }
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    double n = (double)iters * 8 * ILP * (OP == 3 ? 2 : 1);
    printf("%-9s ILP=%2d warps/SM=%d : %.2f cycles per FP64 instruction\n", name, ILP, threads / 32, c / n);
  }
}
int main() {
  double *in, *out; long long* cyc;
  cudaMalloc(&in, 64); cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 64);
  double h[3] = {1e-9, 1.0000001, 1.0};
  cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
#define ALL(OP, NAME) run<1, OP>(NAME, out, in, cyc); run<2, OP>(NAME, out, in, cyc); run<4, OP>(NAME, out, in, cyc); run<8, OP>(NAME, out, in, cyc); run<12, OP>(NAME, out, in, cyc);
  ALL(0, "DADD") ALL(1, "DMUL") ALL(2, "DFMA") ALL(3, "DMUL+DADD")
  return 0;
}
