// Host-side helpers shared by the orchestration code (launch macro, CUB wrappers, small read-backs).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <cub/cub.cuh>

#include "am3d_ctx.h"
#include "am3d_math.cuh"

static bool amTrace() { static int t = -1; if (t < 0) t = getenv("AM3D_TRACE") ? 1 : 0; return t == 1; }
#define LAUNCH(ctx, kernel, grid, block, ...)                  \
  do {                                                         \
    if (amTrace()) fprintf(stderr, "[am3d] %s grid=%d\n", #kernel, (int)(grid)); \
    if ((grid) > 0) {                                          \
      kernel<<<(grid), (block), 0, (ctx)->stream>>>(__VA_ARGS__); \
      (ctx)->kernelLaunches++;                                 \
      cudaError_t _le = cudaGetLastError();                    \
      if (_le != cudaSuccess) throw AmError(AM3D_ECUDA, std::string("launch of " #kernel ": ") + cudaGetErrorString(_le)); \
    }                                                          \
  } while (0)

template <class F>
static void cubRun(am3d_ctx* c, F f) {
  size_t bytes = 0;
  CK(f(nullptr, bytes));
  c->cubTemp.ensure(bytes + 16);
  CK(f(c->cubTemp.p, bytes));
  c->kernelLaunches++;
}

template <class T>
static void h2d(am3d_ctx* c, DevBuf<T>& d, const T* src, size_t n) {
  d.ensure(n ? n : 1);
  if (n) CK(cudaMemcpyAsync(d.p, src, n * sizeof(T), cudaMemcpyHostToDevice, c->stream));
}
template <class T>
static void h2dv(am3d_ctx* c, DevBuf<T>& d, const std::vector<T>& v) { h2d(c, d, v.data(), v.size()); }

// Small read-backs of the step (scan totals, phase tables, iteration counts) do NOT go through the copy engine: a kernel
// writes them into mapped pinned host memory.  A device-to-host memcpy - however small - queues behind whatever the
// copy engine is moving, and with am3d_download_bodies_async that is the previous step's body state (150 - 200 MB): the
// ~15 read-backs of a step then waited for it one after the other, which put the whole transfer (2.5 - 3 ms) back on
// the step's critical path.
#define AM3D_MAPPED_INTS 16384
__global__ void k_export_ints(const int* __restrict__ src, int* __restrict__ dstMapped, int n) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) dstMapped[i] = src[i];
}
// n 32-bit words from device memory to `dst`, synchronising the context's stream
static void readBack(am3d_ctx* c, void* dst, const void* devSrc, size_t words) {
  if (words == 0) { CK(cudaStreamSynchronize(c->stream)); return; }
  if (words > AM3D_MAPPED_INTS) {
    CK(cudaMemcpyAsync(dst, devSrc, words * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return;
  }
  if (!c->mappedHost) {
    CK(cudaHostAlloc((void**)&c->mappedHost, AM3D_MAPPED_INTS * sizeof(int), cudaHostAllocMapped));
    CK(cudaHostGetDevicePointer((void**)&c->mappedDev, c->mappedHost, 0));
  }
  int blocks = (int)std::min<size_t>((words + 255) / 256, 16);
  k_export_ints<<<blocks, 256, 0, c->stream>>>((const int*)devSrc, c->mappedDev, (int)words);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(c->stream));
  memcpy(dst, c->mappedHost, words * sizeof(int));
}
static int readInt(am3d_ctx* c, const int* p) {
  int v;
  readBack(c, &v, p, 1);
  return v;
}

static int bitsFor(unsigned long long v) {
  int b = 1;
  while ((v >> b) && b < 63) b++;
  return b;
}

// ------------------------------------------------------------------------------------------------
// scene upload / reset
