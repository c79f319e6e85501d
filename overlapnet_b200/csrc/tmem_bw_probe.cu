// tmem_bw_probe.cu -- microbenchmark: tcgen05.st / tcgen05.ld throughput per SM for the shapes the
// producers could use.  ./tmem_bw_probe  prints bytes/clk/SM for 4..16 warps.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#include "umma.cuh"
using namespace umma;

template <int X>
__device__ __forceinline__ void st32(uint32_t taddr, const uint32_t* r);
template <> __device__ __forceinline__ void st32<8>(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
               :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
template <> __device__ __forceinline__ void st32<16>(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
               :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
                  "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
template <> __device__ __forceinline__ void st32<32>(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
               "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
               :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
                  "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
                  "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
                  "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]) : "memory");
}
// 16x256b.x1: each thread supplies 4 registers; the warp writes 16 lanes x 8 columns
__device__ __forceinline__ void st16x256(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.16x256b.x1.b32 [%0], {%1,%2,%3,%4};"
               :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]) : "memory");
}
__device__ __forceinline__ void st16x256x4(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.16x256b.x4.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
               :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
                  "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}

template <int MODE>
__global__ void __launch_bounds__(512, 1) bw_kernel(long long* out, int iters) {
  __shared__ uint32_t s_tmem;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc(&s_tmem, 512);
  fence_before_sync(); __syncthreads(); fence_after_sync();
  const uint32_t tmem = s_tmem + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * 64;
  uint32_t r[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) r[i] = threadIdx.x * 31 + i;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 8) { st32<8>(tmem, r); st32<8>(tmem + 8, r + 8); st32<8>(tmem + 16, r + 16); st32<8>(tmem + 24, r + 24); }
    if (MODE == 16) { st32<16>(tmem, r); st32<16>(tmem + 16, r + 16); }
    if (MODE == 32) { st32<32>(tmem, r); }
    if (MODE == 256) { for (int k = 0; k < 8; ++k) st16x256(tmem + k * 8, r + 4 * k); }       // 8 x (16 lanes x 8 cols) = 4 KB/2
    if (MODE == 257) { st16x256x4(tmem, r); st16x256x4(tmem + 32, r + 16); }
    if (MODE == 1) { uint32_t v[32]; tmem_ld_x32(tmem, v); tmem_ld_wait(); r[0] += v[5]; }
  }
  if (MODE != 1) tmem_st_wait();
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) out[0] = t1 - t0;
  if (r[0] == 0x12345) out[1] = r[0];
  fence_before_sync(); __syncthreads();
  if (warp == 0) tmem_dealloc(s_tmem, 512);
}

template <int MODE>
void run(const char* name, int bytes_per_warp_iter) {
  long long* d; cudaMalloc(&d, 16);
  for (int warps = 4; warps <= 16; warps += 4) {
    const int iters = 2000;
    bw_kernel<MODE><<<1, warps * 32>>>(d, iters);
    cudaError_t e = cudaDeviceSynchronize();
    long long h = 0; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    printf("%-28s warps=%2d  %8.1f B/clk/SM  (%s)\n", name, warps, (double)bytes_per_warp_iter * warps * iters / (double)h,
           cudaGetErrorString(e));
  }
  cudaFree(d);
}

int main() {
  run<8>("st 32x32b.x8 (x4 per iter)", 4096);
  run<16>("st 32x32b.x16 (x2 per iter)", 4096);
  run<32>("st 32x32b.x32", 4096);
  run<256>("st 16x256b.x1 (x8 per iter)", 8 * 16 * 32);
  run<257>("st 16x256b.x4 (x2 per iter)", 2 * 4 * 16 * 32);
  run<1>("ld 32x32b.x32 + wait", 4096);
  return 0;
}
