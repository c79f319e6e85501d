// ss_rate_probe.cu -- issue rate of SS-mode tcgen05.mma (M = 128, K = 16, fp16, SWIZZLE_NONE K-major operands) as a
// function of N and of the byte offset of the A operand's start address.  The batched leg shifts the start address
// of its activation window by 16 B per convolution tap (network_tc.cu, k_leg_batched_tc); this probe measures what
// a start address that is not a multiple of 128 B costs.
//   ./ss_rate_probe         one CTA, 6000 MMAs per case, clocks per MMA
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdio.h>

#include "umma.cuh"
using namespace umma;

__global__ void __launch_bounds__(128, 1)
probe(int N, int a_off, int n_mma, int per_commit, int n_acc, long long* __restrict__ clk_out, int* __restrict__ flag) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t s_tmem;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 96 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  if (warp == 0) tmem_alloc(&s_tmem, 512);
  fence_proxy_async();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = s_tmem;
  if (warp == 1) {
    const uint32_t idesc = make_idesc_f16(128, N);
    // A: [2 k-cores][144 rows][8] at smem + a_off (window pitch 144 rows like a one-tile leg window); B: [2][N][8] at +64 KB
    const uint64_t ad = make_desc_kmajor_noswizzle(smem_u32(smem) + a_off, 144 * 16, 128);
    const uint64_t bd = make_desc_kmajor_noswizzle(smem_u32(smem + 64 * 1024), N * 16, 128);
    const bool leader = elect_one() != 0;
    uint32_t phase = 0;
    const long long t0 = clock64();
    for (int i = 0; i < n_mma; i += per_commit) {
      if (leader) {
        for (int k = 0; k < per_commit; ++k) mma_ss(tmem + (k % n_acc) * (512 / n_acc), ad, bd, idesc, 1);
        commit(&bar);
      }
      __syncwarp();
      if (!mbar_wait(&bar, phase, 1ll << 30)) { if (lane == 0) atomicExch(flag, 2); break; }
      phase ^= 1;
    }
    if (lane == 0) clk_out[0] = clock64() - t0;
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

int main() {
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
  long long* dclk; int* dflag;
  cudaMalloc(&dclk, 8); cudaMalloc(&dflag, 4); cudaMemset(dflag, 0, 4);
  printf("SS-mode tcgen05.mma M=128 K=16 fp16, SWIZZLE_NONE; clocks per MMA (60 MMAs per commit)\n");
  printf("%5s %6s | %10s %10s %10s %10s\n", "N", "floor", "A+0 B", "A+16 B", "A+64 B", "A+128 B");
  for (int n_acc : {1, 2}) {
    printf("-- %d accumulator tile(s) in rotation\n", n_acc);
    for (int N : {16, 32, 64, 128, 256}) {
      if (n_acc * N > 512) continue;
      printf("%5d %6d |", N, N / 2);
      for (int off : {0, 16, 64, 128}) {
        probe<<<1, 128, 96 * 1024>>>(N, off, 6000, 60, n_acc, dclk, dflag);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 2; }
        long long h = 0; cudaMemcpy(&h, dclk, 8, cudaMemcpyDeviceToHost);
        printf(" %10.1f", (double)h / 6000);
      }
      printf("\n");
    }
  }
  int flag = 0; cudaMemcpy(&flag, dflag, 4, cudaMemcpyDeviceToHost);
  printf("flag %d\n", flag);
  return flag ? 1 : 0;
}
