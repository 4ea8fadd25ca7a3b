// Hand-written device-wide primitives of the step: a stable LSD radix sort of (64-bit key, value) pairs and an exclusive
// prefix sum of ints.  They serve the sort-based broadphase (Morton cell codes, candidate pair keys), the solve order
// (layer | colour | size keys), the warm-start indices and the merge / unmerge bookkeeping.
//   am3d_set_option("own_primitives", 0) switches every call site back to cub::DeviceRadixSort / cub::DeviceScan (kept as the
//   cross-check: tests/test_gpu_primitives.py runs the same scenes both ways and compares bit for bit).
//
// Radix sort, 8 bits per pass, three launches per pass:
//   k_rs_hist     per tile of 2048 keys a 256-bin digit histogram in shared memory -> hist[digit][tile]
//   amExclusiveSum over hist in digit-major order -> where every (digit, tile) run starts in the output
//   k_rs_scatter  re-reads the tile; a warp owns 256 CONSECUTIVE keys and walks them in rounds of 32: lanes holding the same
//                 digit find each other with __match_any_sync, rank = (the warp's running count of that digit) + (lower lanes
//                 of the match); per-warp counts are then prefixed over the 8 warps - so equal digits keep their input order
//                 (stable, which LSD needs) - and every pair is written to its final place of the pass.
// Input arrays are left untouched; passes ping-pong between the output and a temporary so that the last pass lands in the output.
#pragma once
#include "am3d_host_util.cuh"

#define RS_THREADS 256
#define RS_ITEMS 8
#define RS_TILE (RS_THREADS * RS_ITEMS)
#define RS_RADIX 256
#define SC_THREADS 256
#define SC_ITEMS 8
#define SC_TILE (SC_THREADS * SC_ITEMS)

// ---------------------------------------------------------------- exclusive prefix sum ----------------------------------------
__global__ void __launch_bounds__(SC_THREADS) k_scan_reduce(const int* __restrict__ in, int n, int* __restrict__ sums) {
  __shared__ int ws[SC_THREADS / 32];
  long long base = (long long)blockIdx.x * SC_TILE;
  int s = 0;
#pragma unroll
  for (int j = 0; j < SC_ITEMS; j++) {
    long long i = base + (long long)j * SC_THREADS + threadIdx.x;
    if (i < n) s += in[i];
  }
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int w = 0; w < SC_THREADS / 32; w++) t += ws[w];
    sums[blockIdx.x] = t;
  }
}
// exclusive scan of the tile sums in place, one block (chunks of 1024 with a carry)
__global__ void __launch_bounds__(1024) k_scan_sums(int* __restrict__ sums, int nb) {
  __shared__ int ws[32];
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int base = 0; base < nb; base += 1024) {
    int i = base + threadIdx.x;
    int v = i < nb ? sums[i] : 0;
    int x = v;
    for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
    if (lane == 31) ws[warp] = x;
    __syncthreads();
    if (warp == 0) {
      int t = ws[lane];
      for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, t, o); if (lane >= o) t += y; }
      ws[lane] = t;  // inclusive over the warps
    }
    __syncthreads();
    int before = carry + (warp > 0 ? ws[warp - 1] : 0) + (x - v);
    if (i < nb) sums[i] = before;
    __syncthreads();
    if (threadIdx.x == 1023) carry = before + v;
    __syncthreads();
  }
}
// every thread owns SC_ITEMS consecutive ints of the tile
__global__ void __launch_bounds__(SC_THREADS) k_scan_apply(const int* __restrict__ in, int n, const int* __restrict__ sums, int* __restrict__ out) {
  __shared__ int ws[SC_THREADS / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  long long first = (long long)blockIdx.x * SC_TILE + (long long)threadIdx.x * SC_ITEMS;
  int v[SC_ITEMS];
  int s = 0;
#pragma unroll
  for (int j = 0; j < SC_ITEMS; j++) { v[j] = (first + j < n) ? in[first + j] : 0; s += v[j]; }
  int x = s;
  for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
  if (lane == 31) ws[warp] = x;
  __syncthreads();
  int wbefore = 0;
  for (int w = 0; w < warp; w++) wbefore += ws[w];
  int run = sums[blockIdx.x] + wbefore + (x - s);
#pragma unroll
  for (int j = 0; j < SC_ITEMS; j++) {
    if (first + j < n) out[first + j] = run;
    run += v[j];
  }
}
static void amExclusiveSum(am3d_ctx* c, const int* in, int* out, int n) {
  if (n <= 0) return;
  int nb = (int)(((long long)n + SC_TILE - 1) / SC_TILE);
  c->scanSums.ensure(nb + 1);
  k_scan_reduce<<<nb, SC_THREADS, 0, c->stream>>>(in, n, c->scanSums.p);
  k_scan_sums<<<1, 1024, 0, c->stream>>>(c->scanSums.p, nb);
  k_scan_apply<<<nb, SC_THREADS, 0, c->stream>>>(in, n, c->scanSums.p, out);
  CK(cudaGetLastError());
  c->kernelLaunches += 3;
}

// ---------------------------------------------------------------- radix sort ---------------------------------------------------
template <class K>
__global__ void __launch_bounds__(RS_THREADS) k_rs_hist(const K* __restrict__ keys, int n, int shift, unsigned mask, int nTiles,
                                                        int* __restrict__ hist) {
  __shared__ int h[RS_RADIX];
  h[threadIdx.x] = 0;
  __syncthreads();
  long long base = (long long)blockIdx.x * RS_TILE;
#pragma unroll
  for (int j = 0; j < RS_ITEMS; j++) {
    long long i = base + (long long)j * RS_THREADS + threadIdx.x;
    if (i < n) atomicAdd(&h[(unsigned)(keys[i] >> shift) & mask], 1);
  }
  __syncthreads();
  hist[(size_t)threadIdx.x * nTiles + blockIdx.x] = h[threadIdx.x];
}
template <class K, class V>
__global__ void __launch_bounds__(RS_THREADS) k_rs_scatter(const K* __restrict__ keys, const V* __restrict__ vals, int n, int shift,
                                                           unsigned mask, int nTiles, const int* __restrict__ offsets,
                                                           K* __restrict__ keysOut, V* __restrict__ valsOut) {
  __shared__ int cnt[RS_THREADS / 32][RS_RADIX];  // per warp: running count of every digit, later its start inside the tile's digit run
  __shared__ int gOff[RS_RADIX];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (RS_THREADS / 32) * RS_RADIX; i += RS_THREADS) (&cnt[0][0])[i] = 0;
  gOff[threadIdx.x] = offsets[(size_t)threadIdx.x * nTiles + blockIdx.x];
  __syncthreads();
  const long long first = (long long)blockIdx.x * RS_TILE + (long long)warp * (RS_TILE / (RS_THREADS / 32));
  K k[RS_ITEMS];
  V v[RS_ITEMS];
  int rank[RS_ITEMS];
  unsigned dig[RS_ITEMS];
#pragma unroll
  for (int r = 0; r < RS_ITEMS; r++) {
    long long i = first + r * 32 + lane;
    bool valid = i < n;
    if (valid) { k[r] = keys[i]; v[r] = vals[i]; }
    dig[r] = valid ? ((unsigned)(k[r] >> shift) & mask) : 0u;
    unsigned live = __ballot_sync(0xffffffffu, valid);
    rank[r] = 0;
    if (valid) {
      unsigned peers = __match_any_sync(live, dig[r]);
      int before = cnt[warp][dig[r]];
      __syncwarp(live);
      if (lane == __ffs(peers) - 1) cnt[warp][dig[r]] = before + __popc(peers);
      rank[r] = before + __popc(peers & ((1u << lane) - 1));
    }
    __syncwarp();
  }
  __syncthreads();
  {  // exclusive prefix of the per-warp counts of digit threadIdx.x over the warps
    int run = 0;
#pragma unroll
    for (int w = 0; w < RS_THREADS / 32; w++) { int t = cnt[w][threadIdx.x]; cnt[w][threadIdx.x] = run; run += t; }
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < RS_ITEMS; r++) {
    long long i = first + r * 32 + lane;
    if (i < n) {
      int pos = gOff[dig[r]] + cnt[warp][dig[r]] + rank[r];
      keysOut[pos] = k[r];
      valsOut[pos] = v[r];
    }
  }
}
// keysOut / valsOut <- (keysIn, valsIn) sorted by bits [beginBit, endBit) of the key, stable; inputs untouched
template <class K, class V>
static void amRadixSortPairs(am3d_ctx* c, const K* keysIn, K* keysOut, const V* valsIn, V* valsOut, int n,
                             int beginBit, int endBit) {
  if (n <= 0) return;
  int passes = std::max(1, (endBit - beginBit + 7) / 8);
  int nTiles = (int)(((long long)n + RS_TILE - 1) / RS_TILE);
  // (sized like the candidate-pair arrays, the largest thing sorted in a step, so that the temporaries do not grow step by step)
  size_t want = std::max<size_t>((size_t)n + 1, c->pairKey.cap);
  c->rsKeyTmp.ensure(want);
  c->rsValTmp.ensure(want);
  c->rsHist.ensure((size_t)RS_RADIX * nTiles + 2);
  c->rsOff.ensure((size_t)RS_RADIX * nTiles + 2);
  K* kt = reinterpret_cast<K*>(c->rsKeyTmp.p);  // (64-bit slots: hold either key / value type)
  V* vt = reinterpret_cast<V*>(c->rsValTmp.p);
  const K* kin = keysIn;
  const V* vin = valsIn;
  for (int p = 0; p < passes; p++) {
    int shift = beginBit + 8 * p;
    int bits = std::min(8, std::max(1, endBit - shift));
    unsigned mask = (1u << bits) - 1u;
    bool toOut = ((passes - 1 - p) % 2) == 0;  // the last pass writes the output, the ones before alternate
    K* kout = toOut ? keysOut : kt;
    V* vout = toOut ? valsOut : vt;
    k_rs_hist<K><<<nTiles, RS_THREADS, 0, c->stream>>>(kin, n, shift, mask, nTiles, c->rsHist.p);
    c->kernelLaunches++;
    amExclusiveSum(c, c->rsHist.p, c->rsOff.p, RS_RADIX * nTiles);
    k_rs_scatter<K, V><<<nTiles, RS_THREADS, 0, c->stream>>>(kin, vin, n, shift, mask, nTiles, c->rsOff.p, kout, vout);
    c->kernelLaunches++;
    CK(cudaGetLastError());
    kin = kout;
    vin = vout;
  }
}

// the two entry points the orchestration code calls
template <class K, class V>
static void sortPairs(am3d_ctx* c, const K* keysIn, K* keysOut, const V* valsIn, V* valsOut, int n, int beginBit,
                      int endBit) {
  if (c->ownPrimitives) { amRadixSortPairs<K, V>(c, keysIn, keysOut, valsIn, valsOut, n, beginBit, endBit); return; }
  cubRun(c, [&](void* t, size_t& b) { return cub::DeviceRadixSort::SortPairs(t, b, keysIn, keysOut, valsIn, valsOut, n, beginBit, endBit, c->stream); });
}
static void exclusiveSum(am3d_ctx* c, const int* in, int* out, int n) {
  if (c->ownPrimitives) { amExclusiveSum(c, in, out, n); return; }
  cubRun(c, [&](void* t, size_t& b) { return cub::DeviceScan::ExclusiveSum(t, b, in, out, n, c->stream); });
}

// exclusive scan of n ints (+ a trailing 0) so that out[n] is the total
static int scanTotal(am3d_ctx* c, DevBuf<int>& in, DevBuf<int>& out, int n) {
  out.ensure(n + 1);
  CK(cudaMemsetAsync(in.p + n, 0, sizeof(int), c->stream));
  exclusiveSum(c, in.p, out.p, n + 1);
  return readInt(c, out.p + n);
}
