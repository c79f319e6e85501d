// cta2_probe.cu -- known-answer + rate probe of tcgen05.mma.cta_group::2 (a CTA pair on one TPC computing
// one M = 256 tile, each CTA holding its 128 rows of A and HALF of B's N rows in its own shared memory).
// Validates, before the product kernels rely on them: cluster launch, tcgen05.alloc.cta_group::2, the M = 256
// instruction descriptor, which half of B each CTA must provide, the accumulator layout (CTA r holds rows
// r*128 + lane), tcgen05.commit ... multicast::cluster to both CTAs' barriers.
//   ./cta2_probe            known answer (N = 256 and N = 128, K = 64), then SS rates for N = 128 / 256
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#include "umma.cuh"
using namespace umma;

__device__ __forceinline__ void mma_ts_cta2(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
      :: "r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
}
// A: [256][K] row-major, B: [N][K] row-major (D = A B^T), D: [256][N] fp32.  mode 0 = known answer, 1 = rate.
__global__ void __launch_bounds__(128, 1)
probe(const __half* __restrict__ gA, const __half* __restrict__ gB, float* __restrict__ gD, int N, int K, int mode,
      int n_mma, int per_commit, long long* __restrict__ clk_out, int* __restrict__ flag) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t s_tmem;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  const int Nh = N / 2;                                   // B rows held by each CTA
  __half* sA = reinterpret_cast<__half*>(smem);           // [K/8][128][8]
  __half* sB = reinterpret_cast<__half*>(smem + 64 * 1024);   // [K/8][Nh][8]
  if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  if (warp == 0) tmem_alloc_cta2(&s_tmem, 512);
  for (int e = tid; e < 128 * K; e += 128) {
    const int r = e / K, k = e % K;
    sA[((size_t)(k / 8) * 128 + r) * 8 + (k % 8)] = gA[(size_t)(rank * 128 + r) * K + k];
  }
  for (int e = tid; e < Nh * K; e += 128) {
    const int n = e / K, k = e % K;
    sB[((size_t)(k / 8) * Nh + n) * 8 + (k % 8)] = gB[(size_t)(rank * Nh + n) * K + k];
  }
  fence_proxy_async();
  fence_before_sync();
  __syncthreads();
  cluster_sync_all();                                      // both CTAs' operands, barriers and TMEM are ready
  fence_after_sync();
  const uint32_t tmem = s_tmem;
  const bool ts = mode >= 2;
  if (ts) {
    // A operand into TMEM columns [256, 256 + K/2): lane = row, 32-bit column j holds k = 2j (low half), 2j + 1
    for (int j0 = 0; j0 < K / 2; j0 += 8) {
      uint32_t v[8];
      for (int j = 0; j < 8; ++j) {
        const int r = warp * 32 + lane, k = 2 * (j0 + j);
        const __half lo = gA[(size_t)(rank * 128 + r) * K + k], hi = gA[(size_t)(rank * 128 + r) * K + k + 1];
        v[j] = (uint32_t)__half_as_ushort(lo) | ((uint32_t)__half_as_ushort(hi) << 16);
      }
      tmem_st_x8(tmem + ((uint32_t)(warp * 32) << 16) + 256 + j0, v);
    }
    tmem_st_wait();
    fence_before_sync();
    __syncthreads();
    cluster_sync_all();
    fence_after_sync();
  }
  if (warp == 1 && rank == 0) {
    const uint32_t idesc = make_idesc_f16(256, N);
    const uint64_t ad = make_desc_kmajor_noswizzle(smem_u32(sA), 128 * 16, 128);
    const uint64_t bd = make_desc_kmajor_noswizzle(smem_u32(sB), Nh * 16, 128);
    const bool leader = elect_one() != 0;
    if (mode == 0 || mode == 2) {
      if (leader) {
        for (int k16 = 0; k16 < K / 16; ++k16) {
          if (ts) mma_ts_cta2(tmem, tmem + 256 + k16 * 8, bd + (uint64_t)((k16 * 2 * Nh * 16) >> 4), idesc, k16 != 0);
          else mma_ss_cta2(tmem, ad + (uint64_t)((k16 * 2 * 128 * 16) >> 4), bd + (uint64_t)((k16 * 2 * Nh * 16) >> 4), idesc, k16 != 0);
        }
        commit_cta2(&bar);
      }
      __syncwarp();
    } else {
      uint32_t phase = 0;
      const long long t0 = clock64();
      for (int i = 0; i < n_mma; i += per_commit) {
        if (leader) {
          for (int k = 0; k < per_commit; ++k) {
            if (ts) mma_ts_cta2(tmem + (k % 3) * 64 % 192, tmem + 256 + (k & 3) * 8, bd, idesc, 1);
            else mma_ss_cta2(tmem, ad, bd, idesc, 1);
          }
          commit_cta2(&bar);
        }
        __syncwarp();
        if (!mbar_wait(&bar, phase, 1ll << 30)) { if (lane == 0) atomicExch(flag, 2); break; }
        phase ^= 1;
      }
      if (lane == 0) clk_out[0] = clock64() - t0;
    }
  }
  if (mode == 0 || mode == 2) {
    if (!mbar_wait(&bar, 0, 1ll << 30)) { if (tid == 0) atomicExch(flag, 1); }
    fence_after_sync();
    for (int c0 = 0; c0 < N; c0 += 32) {
      uint32_t v[32];
      tmem_ld_x32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
      tmem_ld_wait();
      for (int j = 0; j < 32; ++j) gD[(size_t)(rank * 128 + warp * 32 + lane) * N + c0 + j] = __uint_as_float(v[j]);
    }
  }
  fence_before_sync();
  __syncthreads();
  cluster_sync_all();                                      // nobody deallocates while the peer still reads / is written to
  if (warp == 0) tmem_dealloc_cta2(tmem, 512);
}

static int launch(const __half* dA, const __half* dB, float* dD, int N, int K, int mode, int n_mma, int pc, long long* dclk, int* dflag) {
  cudaLaunchConfig_t lc = {};
  lc.gridDim = dim3(2); lc.blockDim = dim3(128); lc.dynamicSmemBytes = 128 * 1024;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  lc.attrs = at; lc.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&lc, probe, dA, dB, dD, N, K, mode, n_mma, pc, dclk, dflag);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
  return 0;
}

int main() {
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024);
  const int K = 64;
  int rc = 0;
  long long* dclk; int* dflag;
  cudaMalloc(&dclk, 8); cudaMalloc(&dflag, 4); cudaMemset(dflag, 0, 4);
  for (int tsm = 0; tsm < 2; ++tsm)
  for (int N : {256, 128, 64}) {
    if (!tsm && N == 64) continue;
    std::vector<__half> A(256 * K), B(N * K);
    std::vector<float> Af(256 * K), Bf(N * K);
    srand(7 + N);
    for (size_t i = 0; i < A.size(); ++i) { A[i] = __float2half((rand() % 201 - 100) / 64.0f); Af[i] = __half2float(A[i]); }
    for (size_t i = 0; i < B.size(); ++i) { B[i] = __float2half((rand() % 201 - 100) / 64.0f); Bf[i] = __half2float(B[i]); }
    __half *dA, *dB; float* dD;
    cudaMalloc(&dA, A.size() * 2); cudaMalloc(&dB, B.size() * 2); cudaMalloc(&dD, 256 * N * 4);
    cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice);
    cudaMemset(dD, 0xff, 256 * N * 4);
    if (launch(dA, dB, dD, N, K, tsm ? 2 : 0, 0, 0, dclk, dflag)) return 2;
    std::vector<float> D(256 * N);
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    int flag = 0; cudaMemcpy(&flag, dflag, 4, cudaMemcpyDeviceToHost);
    double maxerr = 0; int bad = 0;
    for (int m = 0; m < 256; ++m)
      for (int n = 0; n < N; ++n) {
        double acc = 0;
        for (int k = 0; k < K; ++k) acc += (double)Af[m * K + k] * Bf[n * K + k];
        const double err = fabs(acc - D[m * N + n]);
        if (!(err <= 1e-3)) { if (bad < 4) printf("  mismatch m=%d n=%d got %f want %f\n", m, n, D[m * N + n], acc); ++bad; }
        if (err > maxerr) maxerr = err;
      }
    printf("cta_group::2 %s M=256 N=%d K=%d: %s (max err %.2e, %d bad, flag %d)\n", tsm ? "TS" : "SS", N, K, bad ? "MISMATCH" : "match", maxerr, bad, flag);
    if (bad || flag) rc = 1;
    for (int pc : {6, 60}) {
      if (launch(dA, dB, dD, N, K, tsm ? 3 : 1, 6000, pc, dclk, dflag)) return 2;
      long long h = 0; cudaMemcpy(&h, dclk, 8, cudaMemcpyDeviceToHost);
      printf("  rate: cta_group::2 %s M=256 N=%3d mma/commit=%2d : %7.1f clk per MMA (ideal %d per SM)\n", tsm ? "TS" : "SS", N, pc, (double)h / 6000, N / 2);
    }
    cudaFree(dA); cudaFree(dB); cudaFree(dD);
  }
  return rc;
}
