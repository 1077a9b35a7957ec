// Microbenchmark (tooling, not product): TMEM -> register read throughput (tcgen05.ld 32x32b.x32) and MUFU ex2 rate per SM,
// the two figures that bound a flash-attention softmax on sm_100a.   nvcc -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../graph-gpt_b200/csrc/common.cuh"
using namespace ggpt;

__global__ void tmem_read_kernel(int iters, int nwarps_active, long long* cyc, float* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t base = slot;
  const uint32_t lane_addr = static_cast<uint32_t>((warp & 3) * 32) << 16;
  float acc = 0.f;
  __syncthreads();
  long long t0 = clock64();
  if (warp < nwarps_active) {
    for (int i = 0; i < iters; ++i) {
      uint32_t r[32];
      tmem_ld32(base + lane_addr + ((i * 32 + (warp >> 2) * 64) & 511 & ~31), r);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; j += 8) acc += __uint_as_float(r[j]);
    }
  }
  long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  if (acc == 123.456f) sink[0] = acc;
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(base, 512); }
}

// two loads in flight before the wait
__global__ void tmem_read2_kernel(int iters, int nwarps_active, long long* cyc, float* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t base = slot;
  const uint32_t lane_addr = static_cast<uint32_t>((warp & 3) * 32) << 16;
  float acc = 0.f;
  __syncthreads();
  long long t0 = clock64();
  if (warp < nwarps_active) {
    for (int i = 0; i < iters; i += 4) {
      uint32_t r0[32], r1[32], r2[32], r3[32];
      tmem_ld32(base + lane_addr + 0, r0);
      tmem_ld32(base + lane_addr + 32, r1);
      tmem_ld32(base + lane_addr + 64, r2);
      tmem_ld32(base + lane_addr + 96, r3);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; j += 8) acc += __uint_as_float(r0[j]) + __uint_as_float(r1[j]) + __uint_as_float(r2[j]) + __uint_as_float(r3[j]);
    }
  }
  long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  if (acc == 123.456f) sink[0] = acc;
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(base, 512); }
}

__global__ void ex2_kernel(int iters, long long* cyc, float* sink) {
  float x[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) x[j] = -0.001f * (threadIdx.x + j);
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] = fast_exp2(x[j]) - 1.5f;
  }
  long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  float s = 0.f;
  for (int j = 0; j < 8; ++j) s += x[j];
  if (s == 123.456f) sink[0] = s;
}

// FMA-pipe exp2 (Cody-Waite + degree-3 polynomial, as FA4 offloads a share of the exponentials)
__device__ __forceinline__ float poly_exp2(float x) {
  x = fmaxf(x, -126.f);
  const float fl = floorf(x);
  const float f = x - fl;
  float p = fmaf(f, 0.0555054f, 0.2402265f);
  p = fmaf(p, f, 0.6931472f);
  p = fmaf(p, f, 1.0f);
  return __int_as_float(__float_as_int(p) + (static_cast<int>(fl) << 23));
}
__global__ void poly_kernel(int iters, long long* cyc, float* sink) {
  float x[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) x[j] = -0.001f * (threadIdx.x + j);
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] = poly_exp2(x[j]) - 1.5f;
  }
  long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  float s = 0.f;
  for (int j = 0; j < 8; ++j) s += x[j];
  if (s == 123.456f) sink[0] = s;
}

int main() {
  long long* cyc; float* sink;
  cudaMalloc(&cyc, 1024 * sizeof(long long)); cudaMalloc(&sink, 4);
  long long h[8];
  const int iters = 4096;
  for (int nw : {1, 4, 8, 16}) {
    int threads = (nw < 4 ? 4 : nw) * 32;
    tmem_read_kernel<<<1, threads>>>(iters, nw, cyc, sink);
    cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("tmem_ld32 + wait per iter, %2d warps (1 CTA): %lld cyc total -> %.1f cyc per 4KB load per warp, SM aggregate %.1f B/cyc\n", nw, h[0],
           (double)h[0] / iters, (double)nw * iters * 4096.0 / h[0]);
    tmem_read2_kernel<<<1, threads>>>(iters, nw, cyc, sink);
    cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("  4 loads in flight,        %2d warps (1 CTA): %lld cyc total -> SM aggregate %.1f B/cyc\n", nw, h[0],
           (double)nw * iters * 4096.0 / h[0]);
  }
  for (int threads : {128, 256, 512, 1024}) {
    ex2_kernel<<<1, threads>>>(1024, cyc, sink);
    cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("ex2.approx: %4d threads: %.2f ex2/cyc/SM\n", threads, (double)threads * 1024 * 8 / h[0]);
    poly_kernel<<<1, threads>>>(1024, cyc, sink);
    cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("poly exp2 : %4d threads: %.2f exp2/cyc/SM\n", threads, (double)threads * 1024 * 8 / h[0]);
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return 0;
}
