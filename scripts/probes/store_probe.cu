// Micro-benchmark: how fast can the convexify kernel's row-store pattern go?  1024 CTAs x 256 threads, each CTA writes
// 1624 rows of 80 bytes (130 KB) into its own region, 55 KB of dynamic shared memory per CTA (4 CTAs per SM).
//   A: one lane per row, 5 x 16-byte stores at an 80-byte stride (what the kernel does)
//   B: warp-contiguous 16-byte stores (a warp covers 512 contiguous bytes per instruction)
//   C: rows staged in shared memory, then one cp.async.bulk (TMA 1D) store of 2560 bytes per 32 rows
//   D: like A but only ONE CTA per SM resident (227 KB of shared memory)
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
constexpr int ROWS = 1624, RD = 10;
__global__ void __launch_bounds__(256, 4) storeA(double* out, double v) {
  extern __shared__ double sm[];
  double* o = out + (size_t)blockIdx.x * ROWS * RD;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int c0 = warp * 32; c0 < ROWS; c0 += 256) {
    const int r = c0 + lane;
    if (r < ROWS) {
      double2* d = reinterpret_cast<double2*>(o + (size_t)r * RD);
#pragma unroll
      for (int i = 0; i < 5; ++i) d[i] = make_double2(v + r, v + i);
    }
  }
  if (v == -1.0) sm[threadIdx.x] = v;
}
__global__ void __launch_bounds__(256, 4) storeB(double* out, double v) {
  extern __shared__ double sm[];
  double2* o = reinterpret_cast<double2*>(out + (size_t)blockIdx.x * ROWS * RD);
  const int n2 = ROWS * RD / 2;
  for (int i = threadIdx.x; i < n2; i += 256) o[i] = make_double2(v + i, v);
  if (v == -1.0) sm[threadIdx.x] = v;
}
__global__ void __launch_bounds__(256, 4) storeC(double* out, double v) {
  extern __shared__ double sm[];
  double* o = out + (size_t)blockIdx.x * ROWS * RD;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* stage = sm + warp * 32 * RD;  // 2560 bytes per warp, 16-byte aligned
  for (int c0 = warp * 32; c0 < ROWS; c0 += 256) {
    const int r = c0 + lane, nrows = (ROWS - c0 < 32) ? ROWS - c0 : 32;
    if (r < ROWS) {
#pragma unroll
      for (int i = 0; i < RD; ++i) stage[lane * RD + i] = v + r + i;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0) {
      const uint32_t s = static_cast<uint32_t>(__cvta_generic_to_shared(stage));
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(o + (size_t)c0 * RD), "r"(s), "r"(nrows * RD * 8) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
    __syncwarp();
  }
  if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
int main() {
  const int B = 1024;
  double* out;
  const size_t bytes = (size_t)B * ROWS * RD * 8;
  cudaMalloc(&out, 2 * bytes);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  auto run = [&](const char* name, void (*k)(double*, double), int smem) {
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    float best = 1e9f;
    for (int it = 0; it < 6; ++it) {
      cudaEventRecord(e0);
      k<<<B, 256, smem>>>(out + (it & 1) * (bytes / 8), 1.0 + it);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (it >= 2 && ms < best) best = ms;
    }
    printf("%-40s %7.1f us  %6.0f GB/s  (%s)\n", name, best * 1e3, bytes / (best * 1e-3) / 1e9, cudaGetErrorString(cudaGetLastError()));
  };
  run("A lane-per-row 16B stores, 4 CTA/SM", storeA, 55 * 1024);
  run("B warp-contiguous 16B stores, 4 CTA/SM", storeB, 55 * 1024);
  run("C smem stage + TMA bulk store, 4 CTA/SM", storeC, 55 * 1024);
  run("A lane-per-row, 1 CTA/SM", storeA, 200 * 1024);
  run("B warp-contiguous, 1 CTA/SM", storeB, 200 * 1024);
  run("A lane-per-row, 8 CTA/SM (24 KB)", storeA, 24 * 1024);
  run("B warp-contiguous, 8 CTA/SM (24 KB)", storeB, 24 * 1024);
  return 0;
}
