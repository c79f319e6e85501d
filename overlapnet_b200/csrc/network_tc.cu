// network_tc.cu -- tensor-core (tcgen05 / TMEM / bulk-async-copy) path of the two heads.
//
// Replaces: DeltaLayer + c_conv1 (generateNet.py:15-61,96-100)      -> k_delta_conv1_tc
//           c_conv2, c_conv3 (+ReLU) (generateNet.py:102-110)       -> k_gemm_stream_tc
//           Flatten + Dense(1, sigmoid) (generateNet.py:112-114)    -> fused epilogue + k_dense_finalize
//
// k_delta_conv1_tc (the kernel that decides scan-pairs/s; 83 % of the FLOPs of a pair)
//   GEMM view per pair:  o1[(i, jb), o] = b1[o] + sum_{dj<15, c<128} |L[i,c] - R[15 jb + dj, c]| W1[dj, c, o]
//   M = 360 x 24, N = 64, K = 1920.  The A operand (66 MB per pair in the reference) is never
//   materialised: 8 producer warps synthesise |l - r| as packed fp16 straight into TENSOR MEMORY
//   (tcgen05.st), one warp's single thread issues tcgen05.mma in TS mode (A from TMEM, B = W1
//   slice from shared memory), accumulators (3 row tiles x 64 fp32 columns) live in TMEM.
//   Mapping: TMEM lane = LEFT row i (3 tiles: i0 = 0, 128, 256), one CTA works through jb = 0..23
//   of a pair; a thread keeps its three LEFT rows' current 16 channels in registers and reads the
//   RIGHT row by broadcast LDS, so each synthesised element costs ~1 ALU instruction.
//   W1 (245 KB fp16) does not fit in shared memory next to L and R: it is streamed per stage
//   (4 KB slices, cp.async.bulk + mbarrier) through a 6-deep ring that is recycled by
//   tcgen05.commit.  Roofline: tensor pipe (co-limited by operand synthesis).
//
// k_gemm_stream_tc
//   D[512 rows x 128] += sum over K-slabs of A_slab[512 x 32] * B_slab[128 x 32]^T with both
//   operands streamed global -> shared by cp.async.bulk in the "C8-interleaved" layout
//   [channel/8][row][8] (16-byte core-matrix rows at uniform pitch, SWIZZLE_NONE descriptors).
//   A per-slab row shift turns the same kernel into the 3x3 convolution c_conv3 (implicit
//   im2col at the copy level: the window shift is just a different source row).
#include "common.cuh"
#include "umma.cuh"

#include <cuda_fp16.h>
#include <stdlib.h>

using namespace umma;

namespace ovn {

constexpr int WF = 360;                 // leg_output_width (the TC path is specialised to the
constexpr int CF = 128;                 //   reference geometry: 360 x 128 volumes, conv1size 15)
constexpr int S15 = 15;
constexpr int NB = 24;                  // 360 / 15
constexpr int PAIR_ROWS = NB * NB;      // 576 rows of c_conv2 output per pair
constexpr long long kWaitCycles = 1ll << 28;

struct TcState {
  __half* w1p = nullptr;        // [60 steps][4][64][8]
  __half* w2p = nullptr;        // [30 slabs][4][128][8]
  __half* w3p = nullptr;        // [2 halves][36 slabs][4][128][8]
  int* slab2_plane = nullptr; int* slab2_shift = nullptr;
  int* slab3_plane = nullptr; int* slab3_shift = nullptr;
  __half* l16 = nullptr;        // [max_pairs][360][128]
  __half* r16 = nullptr;        // [max_pairs][360][128] (pair mode) / [1][360][128] (query mode)
  __half* o1 = nullptr;         // [120 planes][rows_pad][8]
  __half* x3 = nullptr;         // [16 planes][rows_pad][8]
  float* partial = nullptr;     // [rows_pad][2]
  int* d_err = nullptr;
  int64_t rows_pad = 0;
};

// ------------------------------------------------------------------------------------------------
// fp32 feature volumes -> fp16 rows gathered by index (the tensor-core operands)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_gather_rows_f16(const float* __restrict__ bank, const int32_t* __restrict__ idx, int n, __half* __restrict__ out) {
  const int64_t per = (int64_t)WF * CF / 4;            // float4 per volume
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)n * per) return;
  const int p = (int)(i / per);
  const int64_t e = i % per;
  const int64_t row = idx ? idx[p] : p;
  const float4 v = __ldg(reinterpret_cast<const float4*>(bank + row * WF * CF) + e);
  __half2 a = __floats2half2_rn(v.x, v.y), b = __floats2half2_rn(v.z, v.w);
  uint2 o;
  o.x = *reinterpret_cast<uint32_t*>(&a);
  o.y = *reinterpret_cast<uint32_t*>(&b);
  reinterpret_cast<uint2*>(out + (int64_t)p * WF * CF)[e] = o;
}

// ------------------------------------------------------------------------------------------------
// k_delta_conv1_tc
// ------------------------------------------------------------------------------------------------
constexpr int K4_THREADS = 512;
constexpr int K4_STAGES = 6;
constexpr int K4_TILES = 3;
constexpr int K4_ACOL0 = 192;           // TMEM columns: D = [0,192), A stages = [192, 192 + 6*48)
constexpr int K4_STAGE_COLS = 48;
constexpr int K4_STEPS = 60;            // 4 channel chunks x 15 dj per jb
constexpr int K4_BSLICE = 4096;         // bytes of W1 per step: [4 k8][64 o][8]

struct K4Smem {
  __half R[WF * CF];
  __half L[WF * CF];
  __half B[K4_STAGES][K4_BSLICE / 2];
  float bias[64];
  uint64_t a_full[K4_STAGES], b_full[K4_STAGES], empty[K4_STAGES];
  uint64_t d_full, d_empty, l_full, l_empty, r_full;
  uint32_t tmem_base;
};

#define TC_WAIT(bar, parity, code)                       \
  if (!mbar_wait((bar), (parity), kWaitCycles)) {        \
    atomicExch(err, (code));                             \
    goto done;                                           \
  }

__global__ void __launch_bounds__(K4_THREADS, 1)
k_delta_conv1_tc(const __half* __restrict__ L16, const __half* __restrict__ R16, int r_per_pair,
                 const __half* __restrict__ W1p, const float* __restrict__ bias1, __half* __restrict__ o1,
                 int64_t rows_pad, int n_pairs, int* __restrict__ err) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  K4Smem& S = *reinterpret_cast<K4Smem*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int s = 0; s < K4_STAGES; ++s) { mbar_init(&S.a_full[s], 8); mbar_init(&S.b_full[s], 1); mbar_init(&S.empty[s], 1); }
    mbar_init(&S.d_full, 1); mbar_init(&S.d_empty, 4);
    mbar_init(&S.l_full, 1); mbar_init(&S.l_empty, 8); mbar_init(&S.r_full, 1);
    mbar_fence_init();
  }
  if (tid < 64) S.bias[tid] = bias1[tid];
  if (warp == 2) tmem_alloc(&S.tmem_base, 512);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = S.tmem_base;
  constexpr uint32_t VOL_BYTES = WF * CF * 2;

  if (warp == 0) {
    // ===================== loader: L (and R) per pair, W1 slices per step =====================
    if (lane == 0) {
      if (!r_per_pair) {
        mbar_arrive_expect_tx(&S.r_full, VOL_BYTES);
        bulk_g2s(S.R, R16, VOL_BYTES, &S.r_full);
      }
      uint32_t it = 0, pi = 0;
      for (int p = blockIdx.x; p < n_pairs; p += gridDim.x, ++pi) {
        TC_WAIT(&S.l_empty, (pi & 1) ^ 1, 101);
        mbar_arrive_expect_tx(&S.l_full, r_per_pair ? 2 * VOL_BYTES : VOL_BYTES);
        bulk_g2s(S.L, L16 + (size_t)p * WF * CF, VOL_BYTES, &S.l_full);
        if (r_per_pair) bulk_g2s(S.R, R16 + (size_t)p * WF * CF, VOL_BYTES, &S.l_full);
        for (int jb = 0; jb < NB; ++jb) {
          for (int st = 0; st < K4_STEPS; ++st, ++it) {
            const uint32_t s = it % K4_STAGES, ph = (it / K4_STAGES) & 1;
            TC_WAIT(&S.empty[s], ph ^ 1, 102);
            mbar_arrive_expect_tx(&S.b_full[s], K4_BSLICE);
            bulk_g2s(S.B[s], W1p + (size_t)st * (K4_BSLICE / 2), K4_BSLICE, &S.b_full[s]);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread) ============================================
    if (lane == 0) {
      const uint32_t idesc = make_idesc_f16(128, 64);
      uint32_t it = 0, jbit = 0;
      for (int p = blockIdx.x; p < n_pairs; p += gridDim.x) {
        for (int jb = 0; jb < NB; ++jb, ++jbit) {
          TC_WAIT(&S.d_empty, (jbit & 1) ^ 1, 201);
          fence_after_sync();
          for (int st = 0; st < K4_STEPS; ++st, ++it) {
            const uint32_t s = it % K4_STAGES, ph = (it / K4_STAGES) & 1;
            TC_WAIT(&S.a_full[s], ph, 202);
            TC_WAIT(&S.b_full[s], ph, 203);
            fence_after_sync();
            const uint32_t b_addr = smem_u32(S.B[s]);
#pragma unroll
            for (int t = 0; t < K4_TILES; ++t) {
#pragma unroll
              for (int kk = 0; kk < 2; ++kk) {
                const uint64_t bd = make_desc_kmajor_noswizzle(b_addr + kk * 2048, 1024, 128);
                mma_ts(tmem + t * 64, tmem + K4_ACOL0 + s * K4_STAGE_COLS + t * 16 + kk * 8, bd, idesc,
                       (st | kk) != 0);
              }
            }
            commit(&S.empty[s]);
          }
          commit(&S.d_full);
        }
      }
    }
  } else if (warp >= 4 && warp < 8) {
    // ===================== epilogue: D (TMEM) -> +bias -> fp16 -> o1 planes ====================
    const int q = warp & 3;
    uint32_t jbit = 0;
    for (int p = blockIdx.x; p < n_pairs; p += gridDim.x) {
      for (int jb = 0; jb < NB; ++jb, ++jbit) {
        TC_WAIT(&S.d_full, jbit & 1, 301);
        fence_after_sync();
#pragma unroll 1
        for (int t = 0; t < K4_TILES; ++t) {
          const int i = t * 128 + q * 32 + lane;
          const int ib = i / S15, di = i - ib * S15;
          const int64_t m = (int64_t)p * PAIR_ROWS + ib * NB + jb;
#pragma unroll
          for (int c0 = 0; c0 < 64; c0 += 16) {
            uint32_t v[16];
            tmem_ld_x16(tmem + ((uint32_t)(q * 32) << 16) + t * 64 + c0, v);
            tmem_ld_wait();
            if (i < WF) {
#pragma unroll
              for (int h8 = 0; h8 < 2; ++h8) {
                uint32_t pk[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const int o = c0 + h8 * 8 + 2 * j;
                  __half2 hh = __floats2half2_rn(__uint_as_float(v[h8 * 8 + 2 * j]) + S.bias[o],
                                                 __uint_as_float(v[h8 * 8 + 2 * j + 1]) + S.bias[o + 1]);
                  pk[j] = *reinterpret_cast<uint32_t*>(&hh);
                }
                const int k8 = di * 8 + (c0 >> 3) + h8;           // plane = (di, o/8)
                *reinterpret_cast<uint4*>(o1 + ((size_t)k8 * rows_pad + m) * 8) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
              }
            }
          }
        }
        fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&S.d_empty);
      }
    }
  } else if (warp >= 8) {
    // ===================== producers: |l - r| -> TMEM ==========================================
    const int pw = warp - 8, q = pw & 3, half = pw >> 2;
    const int row0 = q * 32 + lane;
    if (!r_per_pair) { TC_WAIT(&S.r_full, 0, 401); }
    uint32_t it = 0, pi = 0;
    for (int p = blockIdx.x; p < n_pairs; p += gridDim.x, ++pi) {
      TC_WAIT(&S.l_full, pi & 1, 402);
      for (int jb = 0; jb < NB; ++jb) {
#pragma unroll 1
        for (int cc = 0; cc < 4; ++cc) {
          const int ch = cc * 32 + half * 16;
          __half2 Lr[K4_TILES][8];
#pragma unroll
          for (int t = 0; t < K4_TILES; ++t) {
            const int i = t * 128 + row0;
            if (i < WF) {
              const uint4 a = *reinterpret_cast<const uint4*>(&S.L[i * CF + ch]);
              const uint4 b = *reinterpret_cast<const uint4*>(&S.L[i * CF + ch + 8]);
              const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
              for (int j = 0; j < 8; ++j) Lr[t][j] = *reinterpret_cast<const __half2*>(&w[j]);
            } else {
#pragma unroll
              for (int j = 0; j < 8; ++j) Lr[t][j] = __float2half2_rn(0.f);
            }
          }
#pragma unroll 1
          for (int dj = 0; dj < S15; ++dj, ++it) {
            const uint32_t s = it % K4_STAGES, ph = (it / K4_STAGES) & 1;
            const int rrow = jb * S15 + dj;
            const uint4 ra = *reinterpret_cast<const uint4*>(&S.R[rrow * CF + ch]);        // broadcast
            const uint4 rb = *reinterpret_cast<const uint4*>(&S.R[rrow * CF + ch + 8]);
            const uint32_t rw[8] = {ra.x, ra.y, ra.z, ra.w, rb.x, rb.y, rb.z, rb.w};
            TC_WAIT(&S.empty[s], ph ^ 1, 403);
            fence_after_sync();
#pragma unroll
            for (int t = 0; t < K4_TILES; ++t) {
              uint32_t o[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                // |l - r|: subtract on the FMA pipe, clear both sign bits on the ALU pipe (LOP3)
                const __half2 d = __hsub2(Lr[t][j], *reinterpret_cast<const __half2*>(&rw[j]));
                o[j] = *reinterpret_cast<const uint32_t*>(&d) & 0x7fff7fffu;
              }
              tmem_st_x8(tmem + ((uint32_t)(q * 32) << 16) + K4_ACOL0 + s * K4_STAGE_COLS + t * 16 + half * 8, o);
            }
            tmem_st_wait();
            fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&S.a_full[s]);
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&S.l_empty);
    }
  }
done:
  fence_before_sync();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem, 512);
}

// ------------------------------------------------------------------------------------------------
// k_gemm_stream_tc
// ------------------------------------------------------------------------------------------------
constexpr int G_THREADS = 256;
constexpr int G_STAGES = 4;
constexpr int G_ROWS = 512;                      // 4 row tiles of 128
constexpr int G_A_BYTES = 4 * G_ROWS * 16;       // 4 planes x 512 rows x 16 B = 32 KB
constexpr int G_B_BYTES = 4 * 128 * 16;          // 4 planes x 128 n x 16 B   =  8 KB

struct GSmem {
  uint8_t A[G_STAGES][G_A_BYTES];
  uint8_t B[G_STAGES][G_B_BYTES];
  float bias[128];
  uint64_t full[G_STAGES], empty[G_STAGES], d_full;
  uint32_t tmem_base;
};

struct GemmArgs {
  const __half* A;            // planes [n_planes][rows_pad][8]
  int64_t rows_pad;
  const int* slab_plane;      // first of the 4 consecutive planes of each slab
  const int* slab_shift;      // row shift of each slab (implicit im2col)
  int n_slabs;
  const __half* Bp;           // [n_half][n_slabs][4][128][8]
  const float* bias;          // [n_half*128]
  int64_t M;                  // valid rows
  // epilogue 1: relu -> fp16 planes
  __half* out_planes; int64_t out_rows_pad;
  // epilogue 2: relu -> dot with the Dense kernel -> per-row partial sums
  const float* wd; float* partial; int grid_w, valid_w, valid_h, n_total;
};

template <int EPI>
__global__ void __launch_bounds__(G_THREADS, 1)
k_gemm_stream_tc(GemmArgs g, int* __restrict__ err) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  GSmem& S = *reinterpret_cast<GSmem*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t row0 = (int64_t)blockIdx.x * G_ROWS;
  const int nh = blockIdx.y;

  if (tid == 0) {
    for (int s = 0; s < G_STAGES; ++s) { mbar_init(&S.full[s], 1); mbar_init(&S.empty[s], 1); }
    mbar_init(&S.d_full, 1);
    mbar_fence_init();
  }
  if (tid < 128) S.bias[tid] = g.bias[nh * 128 + tid];
  if (warp == 2) tmem_alloc(&S.tmem_base, 512);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = S.tmem_base;

  if (warp == 0) {
    if (lane == 0) {
      for (int sl = 0; sl < g.n_slabs; ++sl) {
        const uint32_t s = sl % G_STAGES, ph = (sl / G_STAGES) & 1;
        TC_WAIT(&S.empty[s], ph ^ 1, 501);
        mbar_arrive_expect_tx(&S.full[s], G_A_BYTES + G_B_BYTES);
        const int plane = g.slab_plane[sl];
        const int64_t r = row0 + g.slab_shift[sl];
#pragma unroll
        for (int j = 0; j < 4; ++j)
          bulk_g2s(S.A[s] + j * (G_ROWS * 16), g.A + ((size_t)(plane + j) * g.rows_pad + r) * 8, G_ROWS * 16, &S.full[s]);
        bulk_g2s(S.B[s], g.Bp + ((size_t)nh * g.n_slabs + sl) * (G_B_BYTES / 2), G_B_BYTES, &S.full[s]);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc_f16(128, 128);
      for (int sl = 0; sl < g.n_slabs; ++sl) {
        const uint32_t s = sl % G_STAGES, ph = (sl / G_STAGES) & 1;
        TC_WAIT(&S.full[s], ph, 502);
        fence_after_sync();
        const uint32_t a_addr = smem_u32(S.A[s]), b_addr = smem_u32(S.B[s]);
#pragma unroll
        for (int t = 0; t < 4; ++t) {
#pragma unroll
          for (int kk = 0; kk < 2; ++kk) {
            const uint64_t ad = make_desc_kmajor_noswizzle(a_addr + t * 128 * 16 + kk * 2 * (G_ROWS * 16), G_ROWS * 16, 128);
            const uint64_t bd = make_desc_kmajor_noswizzle(b_addr + kk * 2 * (128 * 16), 128 * 16, 128);
            mma_ss(tmem + t * 128, ad, bd, idesc, (sl | kk) != 0);
          }
        }
        commit(&S.empty[s]);
      }
      commit(&S.d_full);
    }
  } else if (warp >= 4) {
    const int q = warp & 3;
    TC_WAIT(&S.d_full, 0, 503);
    fence_after_sync();
#pragma unroll 1
    for (int t = 0; t < 4; ++t) {
      const int64_t r = row0 + t * 128 + q * 32 + lane;
      if (EPI == 1) {
#pragma unroll 1
        for (int c0 = 0; c0 < 128; c0 += 16) {
          uint32_t v[16];
          tmem_ld_x16(tmem + ((uint32_t)(q * 32) << 16) + t * 128 + c0, v);
          tmem_ld_wait();
          if (r < g.M) {
#pragma unroll
            for (int h8 = 0; h8 < 2; ++h8) {
              uint32_t pk[4];
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const int n = c0 + h8 * 8 + 2 * j;
                __half2 hh = __floats2half2_rn(fmaxf(__uint_as_float(v[h8 * 8 + 2 * j]) + S.bias[n], 0.f),
                                               fmaxf(__uint_as_float(v[h8 * 8 + 2 * j + 1]) + S.bias[n + 1], 0.f));
                pk[j] = *reinterpret_cast<uint32_t*>(&hh);
              }
              const int plane = nh * 16 + (c0 >> 3) + h8;
              *reinterpret_cast<uint4*>(g.out_planes + ((size_t)plane * g.out_rows_pad + r) * 8) =
                  make_uint4(pk[0], pk[1], pk[2], pk[3]);
            }
          }
        }
      } else {
        // row r = pair * (grid_w*grid_w) + y * grid_w + x ; valid output pixel iff y < valid_h, x < valid_w
        const int per = g.grid_w * g.grid_w;
        const int rem = (int)(r % per);
        const int y = rem / g.grid_w, x = rem - y * g.grid_w;
        const bool valid = (r < g.M) && (y < g.valid_h) && (x < g.valid_w);
        const float* wrow = g.wd + ((size_t)(valid ? (y * g.valid_w + x) : 0) * g.n_total + nh * 128);
        float acc = 0.f;
#pragma unroll 1
        for (int c0 = 0; c0 < 128; c0 += 16) {
          uint32_t v[16];
          tmem_ld_x16(tmem + ((uint32_t)(q * 32) << 16) + t * 128 + c0, v);
          tmem_ld_wait();
          if (valid) {
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) {
              const float4 w = __ldg(reinterpret_cast<const float4*>(wrow + c0) + j4);
              acc = fmaf(fmaxf(__uint_as_float(v[j4 * 4 + 0]) + S.bias[c0 + j4 * 4 + 0], 0.f), w.x, acc);
              acc = fmaf(fmaxf(__uint_as_float(v[j4 * 4 + 1]) + S.bias[c0 + j4 * 4 + 1], 0.f), w.y, acc);
              acc = fmaf(fmaxf(__uint_as_float(v[j4 * 4 + 2]) + S.bias[c0 + j4 * 4 + 2], 0.f), w.z, acc);
              acc = fmaf(fmaxf(__uint_as_float(v[j4 * 4 + 3]) + S.bias[c0 + j4 * 4 + 3], 0.f), w.w, acc);
            }
          }
        }
        if (r < g.M) g.partial[r * 2 + nh] = valid ? acc : 0.f;
      }
    }
  }
done:
  fence_before_sync();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem, 512);
}

// Dense bias + sigmoid: fixed-order reduction of the per-row partial sums of one pair
__global__ void __launch_bounds__(256)
k_dense_finalize(const float* __restrict__ partial, const float* __restrict__ bd, int rows_per_pair,
                 float* __restrict__ overlap) {
  __shared__ float red[256];
  const int p = blockIdx.x;
  const float* x = partial + (size_t)p * rows_per_pair * 2;
  float acc = 0.f;
  for (int i = threadIdx.x; i < rows_per_pair * 2; i += 256) acc += x[i];
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) overlap[p] = 1.0f / (1.0f + expf(-(red[0] + bd[0])));
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static int tc_supported(const ovn_handle* h) {
  return h->net_ok && h->cfg.leg_output_width == WF && h->cfg.conv1size == S15;
}

void tc_free(ovn_handle* h) {
  TcState* t = h->tc;
  if (!t) return;
  void* bufs[] = {t->w1p, t->w2p, t->w3p, t->slab2_plane, t->slab2_shift, t->slab3_plane, t->slab3_shift,
                  t->l16, t->r16, t->o1, t->x3, t->partial, t->d_err};
  for (void* b : bufs) if (b) cudaFree(b);
  delete t;
  h->tc = nullptr;
}

template <class T>
static int upload_vec(ovn_handle* h, T** dst, const std::vector<T>& v) {
  OVN_CUDA(h, cudaMalloc(dst, v.size() * sizeof(T)));
  OVN_CUDA(h, cudaMemcpy(*dst, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  return OVN_OK;
}

int tc_pack_weights(ovn_handle* h) {
  if (!tc_supported(h))
    OVN_SET_ERR(h, OVN_ERR_BAD_CONFIG, "precision f16_tc supports leg_output_width=360, conv1size=15 only");
  tc_free(h);
  TcState* t = new TcState();
  h->tc = t;
  const LayerWeights& w1 = h->host_w["c_conv1"];   // (1,15,128,64)
  const LayerWeights& w2 = h->host_w["c_conv2"];   // (15,1,64,128)
  const LayerWeights& w3 = h->host_w["c_conv3"];   // (3,3,128,256)
  // W1p[st = cc*15 + dj][k8][o][e] = W1[dj][c = cc*32 + k8*8 + e][o]
  std::vector<__half> p1((size_t)K4_STEPS * 4 * 64 * 8);
  for (int cc = 0; cc < 4; ++cc)
    for (int dj = 0; dj < S15; ++dj)
      for (int k8 = 0; k8 < 4; ++k8)
        for (int o = 0; o < 64; ++o)
          for (int e = 0; e < 8; ++e) {
            const int c = cc * 32 + k8 * 8 + e;
            p1[((((size_t)(cc * S15 + dj) * 4 + k8) * 64 + o) * 8) + e] =
                __float2half(w1.kernel[((size_t)dj * CF + c) * 64 + o]);
          }
  // c_conv2: K index k = di*64 + o, plane k8 = di*8 + o/8; slab = 4 planes; B[sl][j][n][e] = W2[di][o][n]
  std::vector<__half> p2((size_t)30 * 4 * 128 * 8);
  std::vector<int> s2p(30), s2s(30, 0);
  for (int sl = 0; sl < 30; ++sl) {
    s2p[sl] = sl * 4;
    for (int j = 0; j < 4; ++j) {
      const int k8 = sl * 4 + j, di = k8 / 8, o8 = k8 % 8;
      for (int n = 0; n < 128; ++n)
        for (int e = 0; e < 8; ++e)
          p2[(((size_t)sl * 4 + j) * 128 + n) * 8 + e] = __float2half(w2.kernel[((size_t)di * 64 + o8 * 8 + e) * 128 + n]);
    }
  }
  // c_conv3: slab = (dy, dx, channel group g of 32); planes c8 = g*4..g*4+3 of X3; shift = dy*24 + dx
  std::vector<__half> p3((size_t)2 * 36 * 4 * 128 * 8);
  std::vector<int> s3p(36), s3s(36);
  for (int dy = 0; dy < 3; ++dy)
    for (int dx = 0; dx < 3; ++dx)
      for (int gq = 0; gq < 4; ++gq) {
        const int sl = (dy * 3 + dx) * 4 + gq;
        s3p[sl] = gq * 4;
        s3s[sl] = dy * NB + dx;
        for (int nh = 0; nh < 2; ++nh)
          for (int j = 0; j < 4; ++j)
            for (int n = 0; n < 128; ++n)
              for (int e = 0; e < 8; ++e) {
                const int c = (gq * 4 + j) * 8 + e;
                p3[((((size_t)nh * 36 + sl) * 4 + j) * 128 + n) * 8 + e] =
                    __float2half(w3.kernel[(((size_t)dy * 3 + dx) * 128 + c) * 256 + nh * 128 + n]);
              }
      }
  int rc;
  if ((rc = upload_vec(h, &t->w1p, p1)) != OVN_OK) return rc;
  if ((rc = upload_vec(h, &t->w2p, p2)) != OVN_OK) return rc;
  if ((rc = upload_vec(h, &t->w3p, p3)) != OVN_OK) return rc;
  if ((rc = upload_vec(h, &t->slab2_plane, s2p)) != OVN_OK) return rc;
  if ((rc = upload_vec(h, &t->slab2_shift, s2s)) != OVN_OK) return rc;
  if ((rc = upload_vec(h, &t->slab3_plane, s3p)) != OVN_OK) return rc;
  if ((rc = upload_vec(h, &t->slab3_shift, s3s)) != OVN_OK) return rc;
  const int64_t maxp = h->cfg.max_batch_pairs;
  t->rows_pad = maxp * PAIR_ROWS + 1024;           // tile overrun (512) + window shift (50) slack
  OVN_CUDA(h, cudaMalloc(&t->l16, (size_t)maxp * WF * CF * sizeof(__half)));
  OVN_CUDA(h, cudaMalloc(&t->r16, (size_t)maxp * WF * CF * sizeof(__half)));
  OVN_CUDA(h, cudaMalloc(&t->o1, (size_t)120 * t->rows_pad * 8 * sizeof(__half)));
  OVN_CUDA(h, cudaMalloc(&t->x3, (size_t)16 * t->rows_pad * 8 * sizeof(__half)));
  OVN_CUDA(h, cudaMalloc(&t->partial, (size_t)t->rows_pad * 2 * sizeof(float)));
  OVN_CUDA(h, cudaMalloc(&t->d_err, sizeof(int)));
  OVN_CUDA(h, cudaMemset(t->d_err, 0, sizeof(int)));
  OVN_CUDA(h, cudaMemset(t->o1, 0, (size_t)120 * t->rows_pad * 8 * sizeof(__half)));
  OVN_CUDA(h, cudaMemset(t->x3, 0, (size_t)16 * t->rows_pad * 8 * sizeof(__half)));
  OVN_CUDA(h, cudaFuncSetAttribute(k_delta_conv1_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(K4Smem)));
  OVN_CUDA(h, cudaFuncSetAttribute(k_gemm_stream_tc<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(GSmem)));
  OVN_CUDA(h, cudaFuncSetAttribute(k_gemm_stream_tc<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(GSmem)));
  return OVN_OK;
}

int leg_forward_tc(ovn_handle* h, const float* d_input, int n, float* d_fv, cudaStream_t s) {
  // round 1: the leg runs on the fp32 SIMT kernels in both precision modes (one scan per query
  // in the 1 x N search; the tcgen05 leg is the next step)
  return leg_forward_fp32(h, d_input, n, d_fv, s);
}

int tc_check_error(ovn_handle* h, cudaStream_t s) {
  int e = 0;
  OVN_CUDA(h, cudaMemcpyAsync(&e, h->tc->d_err, sizeof(int), cudaMemcpyDeviceToHost, s));
  OVN_CUDA(h, cudaStreamSynchronize(s));
  if (e != 0) {
    cudaMemsetAsync(h->tc->d_err, 0, sizeof(int), s);
    OVN_SET_ERR(h, OVN_ERR_CUDA, "tensor-core pipeline barrier timed out (code %d)", e);
  }
  return OVN_OK;
}

int heads_forward_tc(ovn_handle* h, const float* d_bank, const float* d_query, const int32_t* d_left,
                     const int32_t* d_right, int n, float* d_overlap, int32_t* d_yaw, float* d_corr,
                     cudaStream_t s) {
  TcState* t = h->tc;
  if (!t) OVN_SET_ERR(h, OVN_ERR_WEIGHTS, "tensor-core weights not packed");
  const int maxp = h->cfg.max_batch_pairs;
  const int base = kMaxLegLayers;
  const int64_t per = (int64_t)WF * CF / 4;
  if (d_query) {
    k_gather_rows_f16<<<(unsigned)((per + 255) / 256), 256, 0, s>>>(d_query, nullptr, 1, t->r16);
    OVN_LAUNCH_CHECK(h);
  }
  for (int p0 = 0; p0 < n; p0 += maxp) {
    const int np = (n - p0 < maxp) ? n - p0 : maxp;
    const int32_t* left = d_left + p0;
    const int32_t* right = d_right ? d_right + p0 : nullptr;
    k_gather_rows_f16<<<(unsigned)((np * per + 255) / 256), 256, 0, s>>>(d_bank, left, np, t->l16);
    OVN_LAUNCH_CHECK(h);
    if (!d_query) {
      k_gather_rows_f16<<<(unsigned)((np * per + 255) / 256), 256, 0, s>>>(d_bank, right, np, t->r16);
      OVN_LAUNCH_CHECK(h);
    }
    const int grid4 = np < h->sm_count ? np : h->sm_count;
    prof_mark(h, PROF_DELTA, s);
    k_delta_conv1_tc<<<grid4, K4_THREADS, sizeof(K4Smem), s>>>(t->l16, t->r16, d_query ? 0 : 1, t->w1p,
                                                               h->d_b[base + 0], t->o1, t->rows_pad, np, t->d_err);
    prof_mark(h, PROF_DELTA, s);
    OVN_LAUNCH_CHECK(h);
    const int64_t M = (int64_t)np * PAIR_ROWS;
    const unsigned gx = (unsigned)((M + G_ROWS - 1) / G_ROWS);
    GemmArgs a2 = {};
    a2.A = t->o1; a2.rows_pad = t->rows_pad; a2.slab_plane = t->slab2_plane; a2.slab_shift = t->slab2_shift;
    a2.n_slabs = 30; a2.Bp = t->w2p; a2.bias = h->d_b[base + 1]; a2.M = M;
    a2.out_planes = t->x3; a2.out_rows_pad = t->rows_pad;
    prof_mark(h, PROF_CONV2, s);
    k_gemm_stream_tc<1><<<dim3(gx, 1), G_THREADS, sizeof(GSmem), s>>>(a2, t->d_err);
    prof_mark(h, PROF_CONV2, s);
    OVN_LAUNCH_CHECK(h);
    GemmArgs a3 = {};
    a3.A = t->x3; a3.rows_pad = t->rows_pad; a3.slab_plane = t->slab3_plane; a3.slab_shift = t->slab3_shift;
    a3.n_slabs = 36; a3.Bp = t->w3p; a3.bias = h->d_b[base + 2]; a3.M = M;
    a3.wd = h->d_w[base + 3]; a3.partial = t->partial; a3.grid_w = NB; a3.valid_w = NB - 2; a3.valid_h = NB - 2;
    a3.n_total = 256;
    prof_mark(h, PROF_CONV3, s);
    k_gemm_stream_tc<2><<<dim3(gx, 2), G_THREADS, sizeof(GSmem), s>>>(a3, t->d_err);
    prof_mark(h, PROF_CONV3, s);
    OVN_LAUNCH_CHECK(h);
    k_dense_finalize<<<np, 256, 0, s>>>(t->partial, h->d_b[base + 3], PAIR_ROWS, d_overlap + p0);
    OVN_LAUNCH_CHECK(h);
    prof_mark(h, PROF_CORR, s);
    int rc = corr_forward_fp32(h, d_bank, d_query, left, right, np, d_yaw + p0,
                               d_corr ? d_corr + (int64_t)p0 * WF : nullptr, s);
    prof_mark(h, PROF_CORR, s);
    if (rc != OVN_OK) return rc;
  }
  static const bool debug_sync = getenv("OVN_DEBUG_SYNC") != nullptr;
  if (debug_sync) return tc_check_error(h, s);
  return OVN_OK;
}

}  // namespace ovn
