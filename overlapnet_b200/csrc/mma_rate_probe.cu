// mma_rate_probe.cu -- cycles per tcgen05.mma (M=128, K=16) for TS / SS operand modes and N = 64..256,
// alone and with concurrent tcgen05.st traffic from 12 other warps.
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdio.h>
#include <stdint.h>
#include "umma.cuh"
using namespace umma;

__global__ void __launch_bounds__(512, 1) rate_kernel(long long* out, int n_mma, int N, int ts, int st_warps, int per_commit, int sw) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t s_tmem;
  __shared__ int stop;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); stop = 0; }
  if (warp == 0) tmem_alloc(&s_tmem, 512);
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  fence_proxy_async(); fence_before_sync(); __syncthreads(); fence_after_sync();
  const uint32_t tmem = s_tmem;
  if (warp == 1) {
    const uint32_t idesc = make_idesc_f16(128, N);
    const uint64_t ad = sw ? make_desc_kmajor_sw128(smem_u32(smem), 0) : make_desc_kmajor_noswizzle(smem_u32(smem), 128 * 16, 128);
    const uint64_t bd = sw ? make_desc_kmajor_sw128(smem_u32(smem + 16384), 0) : make_desc_kmajor_noswizzle(smem_u32(smem + 16384), N * 16, 128);
    const bool leader = elect_one() != 0;
    uint32_t phase = 0;
    const long long t0 = clock64();
    for (int i = 0; i < n_mma; i += per_commit) {
      if (leader) {
        for (int k = 0; k < per_commit; ++k) {
          if (ts) mma_ts(tmem, tmem + 256 + (k & 7) * 8, bd, idesc, 1);
          else mma_ss(tmem, ad, bd, idesc, 1);
        }
        commit(&bar);
      }
      __syncwarp();
      mbar_wait(&bar, phase, 1ll << 30);
      phase ^= 1;
    }
    const long long t1 = clock64();
    if (lane == 0) { out[0] = t1 - t0; stop = 1; }
  } else if (warp >= 4 && warp < 4 + st_warps) {
    uint32_t r[8];
    for (int i = 0; i < 8; ++i) r[i] = threadIdx.x + i;
    const uint32_t a = tmem + ((uint32_t)((warp & 3) * 32) << 16) + 320 + ((warp >> 2) & 3) * 32;
    while (!*(volatile int*)&stop) {
      for (int k = 0; k < 4; ++k) tmem_st_x8(a + k * 8, r);
      tmem_st_wait();
    }
  }
  fence_before_sync(); __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

int main() {
  long long* d; cudaMalloc(&d, 16);
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const int n = 6000;
  for (int sw = 0; sw < 2; ++sw)
    for (int ts = 0; ts < 2; ++ts)
      for (int N : {64, 128, 256})
        for (int pc : {6, 60}) {
          rate_kernel<<<1, 512, 64 * 1024>>>(d, n, N, ts, 0, pc, sw);
          cudaError_t e = cudaDeviceSynchronize();
          long long h = 0; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
          printf("%s %s N=%3d mma/commit=%2d : %7.1f clk per MMA (ideal %d)  %s\n", sw ? "SW128" : "NOSWZ", ts ? "TS" : "SS", N, pc,
                 (double)h / n, N / 2, cudaGetErrorString(e));
        }
  return 0;
}
