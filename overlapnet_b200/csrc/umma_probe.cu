// umma_probe.cu -- standalone known-answer probe of the tcgen05 building blocks used by
// network_tc.cu.  Each mode runs one tiny GEMM D[128 x N] = A[128 x K] * B[N x K]^T through a
// different operand path and compares with a CPU result.  Run one mode per process:
//     ./umma_probe <mode>      (see kModes below; exit code 0 = match)
// Used on the GPU box to validate descriptor encodings before the product kernels rely on them.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#include "umma.cuh"

using namespace umma;

struct Params {
  int mode;      // operand path
  int N, K;      // M is always 128
  int shift;     // row shift for the sliding-window descriptor test
  int a_rows;    // rows of A held in shared memory (>= 128 + shift)
};

static const char* kModes[] = {
    "0: SS, uniform-pitch layout [K/8][rows][8], LBO=rows*16 SBO=128",
    "1: SS, same layout, LBO/SBO swapped in the descriptor (expected to FAIL; diagnostic)",
    "2: SS, sliding window: descriptor start advanced by 16*shift bytes (implicit im2col)",
    "3: TS, A written to TMEM with tcgen05.st, low half = even k",
    "4: TS, high half = even k (expected to FAIL; diagnostic)",
    "5: SS, canonical core-matrix layout [rows/8][K/8][8][8], LBO=128 SBO=(K/8)*128",
    "6: SS, N=256 (4 x tcgen05.ld.x32 per warp... full TMEM row)",
    "7: SS, N=192 (correlation head shape)",
    "8: TS, N=64, K=128, two A buffers at different TMEM columns (pipelining shape)",
    "9: SS, SWIZZLE_128B operands (chunk ^ (row&7)), aligned start",
    "10: SS, SWIZZLE_128B, A start shifted by 5 rows (640 B), base_offset = 0 (diagnostic)",
    "11: SS, SWIZZLE_128B, A start shifted by 5 rows, base_offset = (addr >> 7) & 7",
};

__global__ void __launch_bounds__(128, 1)
probe_kernel(const __half* __restrict__ gA, const __half* __restrict__ gB, float* __restrict__ gD, Params p,
             int* __restrict__ flag) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t s_tmem;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int N = p.N, K = p.K, R = p.a_rows;
  __half* sA = reinterpret_cast<__half*>(smem);
  __half* sB = reinterpret_cast<__half*>(smem + 48 * 1024);

  if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  if (warp == 0) tmem_alloc(&s_tmem, 512);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = s_tmem;
  const bool ts = (p.mode == 3 || p.mode == 4 || p.mode == 8);
  const bool canonical = (p.mode == 5);
  const bool sw = (p.mode >= 9);

  // ---- stage operands
  if (!ts) {
    for (int e = tid; e < R * K; e += 128) {
      const int r = e / K, k = e % K;
      size_t off = canonical ? ((size_t)(r / 8) * (K / 8) + k / 8) * 64 + (r % 8) * 8 + k % 8
                             : ((size_t)(k / 8) * R + r) * 8 + k % 8;
      if (sw) off = (size_t)r * 64 + (((k / 8) ^ (r & 7)) * 8) + k % 8;       // 128-byte rows, chunk swizzle (K == 64)
      sA[off] = gA[(size_t)r * K + k];
    }
  }
  for (int e = tid; e < N * K; e += 128) {
    const int r = e / K, k = e % K;
    size_t off = canonical ? ((size_t)(r / 8) * (K / 8) + k / 8) * 64 + (r % 8) * 8 + k % 8
                           : ((size_t)(k / 8) * N + r) * 8 + k % 8;
    if (sw) off = (size_t)r * 64 + (((k / 8) ^ (r & 7)) * 8) + k % 8;
    sB[off] = gB[(size_t)r * K + k];
  }
  const uint32_t a_col = 256;          // TMEM columns [256, 256 + K/2) hold A in TS mode
  if (ts) {
    const __half* row = gA + (size_t)tid * K;
    for (int c0 = 0; c0 < K / 2; c0 += 8) {
      uint32_t r[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const __half lo = row[2 * (c0 + c)], hi = row[2 * (c0 + c) + 1];
        const uint32_t l = __half_as_ushort(lo), h = __half_as_ushort(hi);
        r[c] = (p.mode == 4) ? ((l << 16) | h) : ((h << 16) | l);
      }
      uint32_t col = a_col + c0;
      if (p.mode == 8 && c0 >= K / 4) col += 64;     // second half of K in a second buffer
      tmem_st_x8(tmem + ((uint32_t)(warp * 32) << 16) + col, r);
    }
    tmem_st_wait();
  }
  fence_proxy_async();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();

  // ---- issue
  if (tid == 0) {
    const uint32_t idesc = make_idesc_f16(128, N);
    uint32_t a_lbo, a_sbo, b_lbo, b_sbo;
    if (canonical) { a_lbo = 128; a_sbo = (K / 8) * 128; b_lbo = 128; b_sbo = (K / 8) * 128; }
    else { a_lbo = R * 16; a_sbo = 128; b_lbo = N * 16; b_sbo = 128; }
    if (p.mode == 1) { uint32_t t = a_lbo; a_lbo = a_sbo; a_sbo = t; t = b_lbo; b_lbo = b_sbo; b_sbo = t; }
    const uint32_t a_base = smem_u32(sA) + (p.mode == 2 ? 16 * p.shift : 0);
    const uint32_t b_base = smem_u32(sB);
    for (int kb = 0; kb < K / 16; ++kb) {
      const uint32_t k_adv_a = canonical ? kb * 2 * 128 : kb * 2 * R * 16;
      const uint32_t k_adv_b = canonical ? kb * 2 * 128 : kb * 2 * N * 16;
      uint64_t bd = make_desc_kmajor_noswizzle(b_base + k_adv_b, b_lbo, b_sbo);
      if (sw) bd = make_desc_kmajor_sw128(b_base + kb * 32, 0);
      if (ts) {
        uint32_t col = a_col + kb * 8;
        if (p.mode == 8 && kb * 8 >= K / 4) col += 64;
        mma_ts(tmem, tmem + col, bd, idesc, kb > 0);
      } else {
        uint64_t ad = make_desc_kmajor_noswizzle(a_base + k_adv_a, a_lbo, a_sbo);
        if (sw) {
          const uint32_t a0 = smem_u32(sA) + p.shift * 128;
          ad = make_desc_kmajor_sw128(a0 + kb * 32, p.mode == 11 ? ((a0 >> 7) & 7) : 0);
        }
        mma_ss(tmem, ad, bd, idesc, kb > 0);
      }
    }
    commit(&bar);
  }
  const bool ok = mbar_wait(&bar, 0, 1ll << 28);
  if (!ok && tid == 0) atomicExch(flag, 1);
  fence_after_sync();
  if (ok) {
    for (int c0 = 0; c0 < N; c0 += 16) {
      uint32_t r[16];
      tmem_ld_x16(tmem + ((uint32_t)(warp * 32) << 16) + c0, r);
      tmem_ld_wait();
#pragma unroll
      for (int c = 0; c < 16; ++c) gD[(size_t)tid * N + c0 + c] = __uint_as_float(r[c]);
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d: %s\n", #x, __LINE__, cudaGetErrorString(e)); return 3; } } while (0)

int main(int argc, char** argv) {
  if (argc < 2) {
    for (const char* m : kModes) printf("%s\n", m);
    return 0;
  }
  Params p;
  p.mode = atoi(argv[1]);
  p.N = 64; p.K = 64; p.shift = 0; p.a_rows = 128;
  if (p.mode == 2 || p.mode == 10 || p.mode == 11) { p.shift = 5; p.a_rows = 144; }
  if (p.mode == 6) p.N = 256;
  if (p.mode == 7) p.N = 192;
  if (p.mode == 8) p.K = 128;
  printf("mode %s\n", p.mode >= 0 && p.mode <= 11 ? kModes[p.mode] : "?");
  const int R = p.a_rows, N = p.N, K = p.K;
  std::vector<__half> A((size_t)R * K), B((size_t)N * K);
  std::vector<float> Af((size_t)R * K), Bf((size_t)N * K);
  srand(1234 + p.mode);
  for (size_t i = 0; i < A.size(); ++i) { float v = (rand() % 33 - 16) / 8.0f; A[i] = __float2half(v); Af[i] = v; }
  for (size_t i = 0; i < B.size(); ++i) { float v = (rand() % 33 - 16) / 8.0f; B[i] = __float2half(v); Bf[i] = v; }
  __half *dA, *dB; float* dD; int* dflag;
  CK(cudaMalloc(&dA, A.size() * 2)); CK(cudaMalloc(&dB, B.size() * 2)); CK(cudaMalloc(&dD, (size_t)128 * N * 4));
  CK(cudaMalloc(&dflag, 4)); CK(cudaMemset(dflag, 0, 4)); CK(cudaMemset(dD, 0xFF, (size_t)128 * N * 4));
  CK(cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice));
  const int smem_bytes = 48 * 1024 + 40 * 1024;
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
  probe_kernel<<<1, 128, smem_bytes>>>(dA, dB, dD, p, dflag);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  int flag = 0; CK(cudaMemcpy(&flag, dflag, 4, cudaMemcpyDeviceToHost));
  std::vector<float> D((size_t)128 * N);
  CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
  if (flag) { printf("RESULT mode %d: TIMEOUT waiting for tcgen05.commit\n", p.mode); return 2; }
  double maxerr = 0; int bad = 0;
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < N; ++n) {
      double ref = 0;
      for (int k = 0; k < K; ++k) ref += (double)Af[(size_t)(m + p.shift) * K + k] * Bf[(size_t)n * K + k];
      double err = fabs(ref - D[(size_t)m * N + n]);
      if (!(err <= 1e-3)) { if (bad < 6) printf("  mismatch m=%d n=%d got %g want %g\n", m, n, D[(size_t)m * N + n], ref); ++bad; }
      if (err > maxerr || err != err) maxerr = err;
    }
  printf("RESULT mode %d: %s  max|err|=%g  mismatches=%d/%d\n", p.mode, bad ? "FAIL" : "PASS", maxerr, bad, 128 * N);
  return bad ? 1 : 0;
}
