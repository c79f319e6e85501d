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
constexpr int K4_PITCH = CF + 8;        // fp16 row pitch of L / R for k_delta_conv1_tc: 272 B => conflict-free LDS.128 across rows

struct TcState {
  __half* w1p = nullptr;        // [60 steps][4][64][8]
  __half* w2p = nullptr;        // [30 slabs][4][128][8]
  __half* w3p = nullptr;        // [2 halves][36 slabs][4][128][8]
  int* slab2_plane = nullptr; int* slab2_shift = nullptr;     // per-copy (n_slabs*4) tables
  int* slab3_plane = nullptr; int* slab3_shift = nullptr;
  // tensor-core leg (layers 2..): packed weights + copy tables per layer, ping-pong activation planes
  __half* wleg[kMaxLegLayers] = {};      // [1][n_slabs][4][NT][8]   (throughput mode)
  __half* wleg64[kMaxLegLayers] = {};    // [cout/64][n_slabs][4][64][8] (streamed, 64-wide)
  __half* wres[kMaxLegLayers] = {};      // [cout/64][kh*kw*3][C_in/8][64][8] (latency mode, resident activations)
  int* leg_plane[kMaxLegLayers] = {};
  int* leg_shift[kMaxLegLayers] = {};
  int leg_slabs[kMaxLegLayers] = {};
  int leg_nt[kMaxLegLayers] = {};
  __half* actp[2] = {nullptr, nullptr};
  __half* l16 = nullptr;        // [max_pairs][360][128]
  __half* r16 = nullptr;        // [max_pairs][360][128] (pair mode) / [1][360][128] (query mode)
  __half* o1 = nullptr;         // [120 planes][rows_pad][8]
  __half* x3 = nullptr;         // [16 planes][rows_pad][8]
  float* partial = nullptr;     // [rows_pad][2]
  __half* lc = nullptr;         // correlation operands: [max_pairs][3 tiles][2 k-halves][hi,lo][8][128][8]
  __half* rc = nullptr;         // [max_pairs or 1][2 n-halves][hi,lo][16][192][8]
  float* corr_part = nullptr;   // [max_pairs][2][360]
  // resident bank (ovn_bank_prepare): operand copies of the LEFT volumes, indexed by bank row
  const float* pb_key = nullptr;
  int64_t pb_cap = 0, pb_rows = 0;      // capacity / rows [0, pb_rows) prepared
  __half* pb_l16 = nullptr;             // [cap][360][K4_PITCH]
  __half* pb_lc = nullptr;              // [cap] x C6_VOL_L_BYTES
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
  const int64_t r = e / (CF / 4), c4 = e % (CF / 4);           // padded row pitch (K4_PITCH halves)
  reinterpret_cast<uint2*>(out + ((int64_t)p * WF + r) * K4_PITCH)[c4] = o;
}

// ------------------------------------------------------------------------------------------------
// k_delta_conv1_tc
// ------------------------------------------------------------------------------------------------
constexpr int K4_PROD_WARPS = 12;       // three groups of 4 warps (one per TMEM lane quarter); group g owns steps with step % 3 == g
constexpr int K4_THREADS = (8 + K4_PROD_WARPS) * 32;
constexpr int K4_STAGES = 6;            // A ring: TMEM column slots
constexpr int K4_BGROUPS = 4;           // B ring: 4 groups of 6 consecutive W1 slices (24 KB, one bulk copy, one barrier each)
constexpr int K4_BSLOTS = 24;
constexpr int K4_TILES = 3;
constexpr int K4_ACOL0 = 192;           // TMEM columns: D = [0,192), A stages = [192, 192 + 6*48)
constexpr int K4_STAGE_COLS = 48;
constexpr int K4_STEPS = 60;            // 4 channel chunks x 15 dj per jb
constexpr int K4_BSLICE = 4096;         // bytes of W1 per step: [4 k8][64 o][8]
constexpr int K4_RWIN_BYTES = S15 * K4_PITCH * 2;   // the 15 RIGHT rows one jb touches

struct K4Smem {
  __half L[WF * K4_PITCH];
  __half Rw[2][S15 * K4_PITCH];         // double-buffered RIGHT-row window (streamed per jb)
  __half B[K4_BSLOTS][K4_BSLICE / 2];
  float bias[64];
  uint64_t a_full[K4_STAGES], a_empty[K4_STAGES], b_full[K4_BGROUPS], b_empty[K4_BGROUPS];
  uint64_t d_full, d_empty, l_full, l_empty, rw_full[2], rw_empty[2];
  uint32_t tmem_base;
};

#define TC_WAIT(bar, parity, code)                       \
  if (!mbar_wait((bar), (parity), kWaitCycles)) {        \
    atomicExch(err, (code));                             \
    goto done;                                           \
  }

// Measured on B200 (profiles/r1_*): (1) with W1 slices sharing the 6-deep A ring the kernel was
// bound by the L2 -> shared round trip of a slice (slot turnaround ~4000 clk), so the W1 ring is
// separate and 24 deep; (2) that only fits next to the LEFT volume if the RIGHT volume is not
// resident: a jb touches just 15 RIGHT rows, which are streamed as a 4 KB double-buffered window;
// (3) a producer warp's LDS -> ALU -> tcgen05.st -> wait::st -> arrive chain runs at IPC ~0.2, so
// 16 producer warps in two groups work on alternating steps.
template <int PROD>   // producer organisation: 1 = 3 groups x 4 warps, full K per thread; 0 = 2 groups x 8 warps, half K per thread
__global__ void __launch_bounds__(PROD == 1 ? 640 : 768, 1)
k_delta_conv1_tc(const __half* __restrict__ L16, const int32_t* __restrict__ l_idx, const __half* __restrict__ R16,
                 int r_per_pair, const __half* __restrict__ W1p, const float* __restrict__ bias1, __half* __restrict__ o1,
                 int64_t rows_pad, int n_pairs, int* __restrict__ err) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  K4Smem& S = *reinterpret_cast<K4Smem*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int s = 0; s < K4_STAGES; ++s) { mbar_init(&S.a_full[s], PROD == 1 ? 4 : 8); mbar_init(&S.a_empty[s], 1); }
    for (int s = 0; s < K4_BGROUPS; ++s) { mbar_init(&S.b_full[s], 1); mbar_init(&S.b_empty[s], 1); }
    mbar_init(&S.d_full, 1); mbar_init(&S.d_empty, 4);
    mbar_init(&S.l_full, 1); mbar_init(&S.l_empty, PROD == 1 ? 12 : 16);
    for (int b = 0; b < 2; ++b) { mbar_init(&S.rw_full[b], 1); mbar_init(&S.rw_empty[b], PROD == 1 ? 12 : 16); }
    mbar_fence_init();
  }
  if (tid < 64) S.bias[tid] = bias1[tid];
  if (warp == 2) tmem_alloc(&S.tmem_base, 512);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = S.tmem_base;
  constexpr uint32_t VOL_BYTES = WF * K4_PITCH * 2;

  if (warp == 0) {
    // ===================== loader A: W1 through the deep ring, 6 consecutive slices per copy ====
    if (lane == 0) {
      uint32_t bg = 0, bph = 0;
      for (int p = blockIdx.x; p < n_pairs; p += gridDim.x) {
        for (int jb = 0; jb < NB; ++jb) {
          for (int o = 0; o < K4_STEPS / K4_STAGES; ++o) {
            TC_WAIT(&S.b_empty[bg], bph ^ 1, 102);
            mbar_arrive_expect_tx(&S.b_full[bg], K4_STAGES * K4_BSLICE);
            bulk_g2s(S.B[bg * K4_STAGES], W1p + (size_t)o * K4_STAGES * (K4_BSLICE / 2), K4_STAGES * K4_BSLICE, &S.b_full[bg]);
            if (++bg == K4_BGROUPS) { bg = 0; bph ^= 1; }
          }
        }
      }
    }
  } else if (warp == 3) {
    // ===================== loader B: LEFT volume per pair, RIGHT row window per jb ==============
    if (lane == 0) {
      uint32_t pi = 0, jbit = 0;
      for (int p = blockIdx.x; p < n_pairs; p += gridDim.x, ++pi) {
        TC_WAIT(&S.l_empty, (pi & 1) ^ 1, 101);
        mbar_arrive_expect_tx(&S.l_full, VOL_BYTES);
        bulk_g2s(S.L, L16 + (size_t)(l_idx ? l_idx[p] : p) * WF * K4_PITCH, VOL_BYTES, &S.l_full);
        const __half* Rp = R16 + (r_per_pair ? (size_t)p * WF * K4_PITCH : 0);
        for (int jb = 0; jb < NB; ++jb, ++jbit) {
          const uint32_t b = jbit & 1;
          TC_WAIT(&S.rw_empty[b], ((jbit >> 1) & 1) ^ 1, 103);
          mbar_arrive_expect_tx(&S.rw_full[b], K4_RWIN_BYTES);
          bulk_g2s(S.Rw[b], Rp + (size_t)jb * S15 * K4_PITCH, K4_RWIN_BYTES, &S.rw_full[b]);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer ==========================================================
    // Warp-uniform loop (addresses / descriptors stay in uniform registers), one elected lane
    // issues.  The A ring is unrolled (60 steps = 10 x 6 slots: slot offsets are immediates and the
    // phase is the parity of the outer counter); the B slot is a running counter.
    {
      const uint32_t idesc = make_idesc_f16(128, 64);
      const uint64_t bdesc0 = make_desc_kmajor_noswizzle(smem_u32(S.B[0]), 1024, 128);
      const uint32_t bd_hi = (uint32_t)(bdesc0 >> 32), bd_lo = (uint32_t)bdesc0;
      const bool leader = elect_one() != 0;
      // An mbarrier probe costs ~90 clk even when the phase is already complete, and this warp is the
      // pacemaker: W1 is waited for once per 6 steps (one barrier per 24 KB group) and the A barrier
      // of step s+1 is probed BEFORE the MMAs of step s are issued, so its latency hides behind them.
      const uint32_t a_full0 = smem_u32(&S.a_full[0]);
      uint32_t jbit = 0, bg = 0, bph = 0;
      for (int p = blockIdx.x; p < n_pairs; p += gridDim.x) {
        for (int jb = 0; jb < NB; ++jb, ++jbit) {
          TC_WAIT(&S.d_empty, (jbit & 1) ^ 1, 201);
          fence_after_sync();
#pragma unroll 1
          for (uint32_t o = 0; o < K4_STEPS / K4_STAGES; ++o) {
            const uint32_t ph = o & 1;            // (step / 6) & 1: steps per jb (60) and per pair are multiples of 12
            TC_WAIT(&S.b_full[bg], bph, 203);
            bool ready = mbar_try_wait_addr(a_full0, ph);
            const uint32_t b_lo = bd_lo + ((bg * K4_STAGES * K4_BSLICE) >> 4);
#pragma unroll
            for (int sg = 0; sg < K4_STAGES; ++sg) {
              if (!ready) { if (!mbar_wait_addr(a_full0 + sg * 8, ph, kWaitCycles)) { atomicExch(err, 202); goto done; } }
              if (sg + 1 < K4_STAGES) ready = mbar_try_wait_addr(a_full0 + (sg + 1) * 8, ph);   // probe ahead
              fence_after_sync();
              if (leader) {
                // consecutive MMAs go to different accumulator tiles (no back-to-back dependency on one D)
#pragma unroll
                for (int kk = 0; kk < 2; ++kk) {
                  const uint64_t bd = ((uint64_t)bd_hi << 32) | (uint64_t)(b_lo + ((sg * K4_BSLICE + kk * 2048) >> 4));
#pragma unroll
                  for (int t = 0; t < K4_TILES; ++t) {
                    mma_ts(tmem + t * 64, tmem + K4_ACOL0 + sg * K4_STAGE_COLS + t * 16 + kk * 8, bd, idesc,
                           (o | (uint32_t)sg | (uint32_t)kk) != 0);
                  }
                }
                commit(&S.a_empty[sg]);
                if (sg == K4_STAGES - 1) commit(&S.b_empty[bg]);
              }
              __syncwarp();
            }
            if (++bg == K4_BGROUPS) { bg = 0; bph ^= 1; }
          }
          if (leader) commit(&S.d_full);
          __syncwarp();
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
#pragma unroll 1
          for (int c = 0; c < 2; ++c) {
            uint32_t v[32];
            tmem_ld_x32(tmem + ((uint32_t)(q * 32) << 16) + t * 64 + c * 32, v);
            tmem_ld_wait();
            if (t == K4_TILES - 1 && c == 1) {   // everything is in registers: hand D back before the stores
              fence_before_sync();
              __syncwarp();
              if (lane == 0) mbar_arrive(&S.d_empty);
            }
            if (i < WF) {
#pragma unroll
              for (int h8 = 0; h8 < 4; ++h8) {
                uint32_t pk[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const int o = c * 32 + h8 * 8 + 2 * j;
                  __half2 hh = __floats2half2_rn(__uint_as_float(v[h8 * 8 + 2 * j]) + S.bias[o],
                                                 __uint_as_float(v[h8 * 8 + 2 * j + 1]) + S.bias[o + 1]);
                  pk[j] = *reinterpret_cast<uint32_t*>(&hh);
                }
                const int k8 = di * 8 + c * 4 + h8;               // plane = (di, o/8)
                *reinterpret_cast<uint4*>(o1 + ((size_t)k8 * rows_pad + m) * 8) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
              }
            }
          }
        }
      }
    }
  } else if (warp >= 8) {
    // ===================== producers: |l - r| -> TMEM ==========================================
    // 12 warps = 3 groups x 4 TMEM lane quarters.  A thread owns LEFT rows q*32+lane (+128, +256);
    // all 32 channels of the current chunk live in registers (48), the RIGHT row comes by broadcast
    // LDS.128 from the per-jb window.  Group g produces the steps with step % 3 == g (15 and 60 are
    // multiples of 3: dj = g, g+3, ...) into ring slots g, g+3: per synthesised element this costs
    // 1 ALU instruction + ~0.3 of loop / barrier overhead, and a group has three MMA stage-times
    // to hide its LDS -> ALU -> tcgen05.st -> wait::st -> arrive chain.
    if constexpr (PROD == 1) {
    const int pw = warp - 8, q = pw & 3, grp = pw >> 2;
    const int row0 = q * 32 + lane;
    const uint32_t lane_addr = tmem + ((uint32_t)(q * 32) << 16) + K4_ACOL0;
    const uint32_t a_empty0 = smem_u32(&S.a_empty[0]), a_full0 = smem_u32(&S.a_full[0]);
    uint32_t pi = 0, jbit = 0, n = 0;       // n: steps produced by this group: slot = grp + 3*(n&1), phase = (n>>1)&1
    bool slot_free = true;                  // result of the probe issued one step ahead (the first two steps find fresh slots)
    for (int p = blockIdx.x; p < n_pairs; p += gridDim.x, ++pi) {
      TC_WAIT(&S.l_full, pi & 1, 402);
      for (int jb = 0; jb < NB; ++jb, ++jbit) {
        const uint32_t wb = jbit & 1;
        TC_WAIT(&S.rw_full[wb], (jbit >> 1) & 1, 404);
#pragma unroll 1
        for (int cc = 0; cc < 4; ++cc) {
          const int ch = cc * 32;
          uint32_t Lr[K4_TILES][16];
#pragma unroll
          for (int t = 0; t < K4_TILES; ++t) {
            const int i = t * 128 + row0;
#pragma unroll
            for (int v4 = 0; v4 < 4; ++v4) {
              uint4 a = make_uint4(0u, 0u, 0u, 0u);
              if (i < WF) a = *reinterpret_cast<const uint4*>(&S.L[i * K4_PITCH + ch + v4 * 8]);
              Lr[t][v4 * 4 + 0] = a.x; Lr[t][v4 * 4 + 1] = a.y; Lr[t][v4 * 4 + 2] = a.z; Lr[t][v4 * 4 + 3] = a.w;
            }
          }
#pragma unroll 1
          for (int dj = grp; dj < S15; dj += 3, ++n) {
            const uint32_t sg = grp + 3 * (n & 1), ph = (n >> 1) & 1;
            const __half* rrow = &S.Rw[wb][dj * K4_PITCH + ch];
            if (!slot_free) { if (!mbar_wait_addr(a_empty0 + sg * 8, ph ^ 1, kWaitCycles)) { atomicExch(err, 403); goto done; } }
            {
              // probe the slot of this group's NEXT step now; the ~90 clk answer is consumed next iteration
              const uint32_t n1 = n + 1, sg1 = grp + 3 * (n1 & 1), ph1 = (n1 >> 1) & 1;
              slot_free = mbar_try_wait_addr(a_empty0 + sg1 * 8, ph1 ^ 1);
            }
            fence_after_sync();
#pragma unroll
            for (int hk = 0; hk < 2; ++hk) {          // two 16-channel halves: keeps the live set of r at 8 registers
              const uint4 ra = *reinterpret_cast<const uint4*>(rrow + hk * 16);        // broadcast LDS
              const uint4 rb = *reinterpret_cast<const uint4*>(rrow + hk * 16 + 8);
              const uint32_t rw[8] = {ra.x, ra.y, ra.z, ra.w, rb.x, rb.y, rb.z, rb.w};
#pragma unroll
              for (int t = 0; t < K4_TILES; ++t) {
                uint32_t o[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  // |l - r|: subtract on the FMA pipe, clear both sign bits on the ALU pipe (LOP3)
                  const __half2 d = __hsub2(*reinterpret_cast<const __half2*>(&Lr[t][hk * 8 + j]),
                                            *reinterpret_cast<const __half2*>(&rw[j]));
                  o[j] = *reinterpret_cast<const uint32_t*>(&d) & 0x7fff7fffu;
                }
                tmem_st_x8(lane_addr + sg * K4_STAGE_COLS + t * 16 + hk * 8, o);
              }
            }
            tmem_st_wait();
            fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive_addr(a_full0 + sg * 8);
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&S.rw_empty[wb]);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&S.l_empty);
    }
    } else {
      // 16 warps = 2 groups x (4 TMEM lane quarters x 2 K halves); group g produces steps with step % 2 == g
      const int pw = warp - 8, q = pw & 3, half = (pw >> 2) & 1, grp = pw >> 3;
      const int row0 = q * 32 + lane;
      const uint32_t lane_addr = tmem + ((uint32_t)(q * 32) << 16) + K4_ACOL0 + half * 8;
      const uint32_t a_empty0 = smem_u32(&S.a_empty[0]), a_full0 = smem_u32(&S.a_full[0]);
      uint32_t pi = 0, jbit = 0, sg = grp, ph = 0;
      for (int p = blockIdx.x; p < n_pairs; p += gridDim.x, ++pi) {
        TC_WAIT(&S.l_full, pi & 1, 402);
        for (int jb = 0; jb < NB; ++jb, ++jbit) {
          const uint32_t wb = jbit & 1;
          TC_WAIT(&S.rw_full[wb], (jbit >> 1) & 1, 404);
#pragma unroll 1
          for (int cc = 0; cc < 4; ++cc) {
            const int ch = cc * 32 + half * 16;
            uint32_t Lr[K4_TILES][8];
#pragma unroll
            for (int t = 0; t < K4_TILES; ++t) {
              const int i = t * 128 + row0;
              uint4 a = make_uint4(0u, 0u, 0u, 0u), b = a;
              if (i < WF) {
                a = *reinterpret_cast<const uint4*>(&S.L[i * K4_PITCH + ch]);
                b = *reinterpret_cast<const uint4*>(&S.L[i * K4_PITCH + ch + 8]);
              }
              Lr[t][0] = a.x; Lr[t][1] = a.y; Lr[t][2] = a.z; Lr[t][3] = a.w;
              Lr[t][4] = b.x; Lr[t][5] = b.y; Lr[t][6] = b.z; Lr[t][7] = b.w;
            }
#pragma unroll 1
            for (int dj = (grp + cc) & 1; dj < S15; dj += 2) {
              const __half* rrow = &S.Rw[wb][dj * K4_PITCH + ch];
              const uint4 ra = *reinterpret_cast<const uint4*>(rrow);
              const uint4 rb = *reinterpret_cast<const uint4*>(rrow + 8);
              const uint32_t rw[8] = {ra.x, ra.y, ra.z, ra.w, rb.x, rb.y, rb.z, rb.w};
              if (!mbar_wait_addr(a_empty0 + sg * 8, ph ^ 1, kWaitCycles)) { atomicExch(err, 403); goto done; }
              fence_after_sync();
#pragma unroll
              for (int t = 0; t < K4_TILES; ++t) {
                uint32_t o[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  const __half2 d = __hsub2(*reinterpret_cast<const __half2*>(&Lr[t][j]),
                                            *reinterpret_cast<const __half2*>(&rw[j]));
                  o[j] = *reinterpret_cast<const uint32_t*>(&d) & 0x7fff7fffu;
                }
                tmem_st_x8(lane_addr + sg * K4_STAGE_COLS + t * 16, o);
              }
              tmem_st_wait();
              fence_before_sync();
              __syncwarp();
              if (lane == 0) mbar_arrive_addr(a_full0 + sg * 8);
              sg += 2;
              if (sg >= K4_STAGES) { sg -= K4_STAGES; ph ^= 1; }
            }
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&S.rw_empty[wb]);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&S.l_empty);
      }
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

// TILES = 128-row tiles per CTA: 4 for throughput (W slabs are reused by four tiles, 4-deep ring
// of 40 KB stages), 1 for latency (single-scan leg: 4x more CTAs, 12-deep ring of 12-16 KB stages
// so that the L2 round trip of a slab is hidden by ring depth rather than by work per stage).
template <int NT, int TILES>
struct GCfg {
  static constexpr int ROWS = TILES * 128;
  static constexpr int A_BYTES = 4 * ROWS * 16;       // 4 planes x ROWS x 16 B
  static constexpr int B_BYTES = 4 * NT * 16;         // 4 planes x NT x 16 B
  static constexpr int STAGES = TILES == 4 ? 4 : 12;
};

template <int NT, int TILES>
struct GSmem {
  using C = GCfg<NT, TILES>;
  uint8_t A[C::STAGES][C::A_BYTES];
  uint8_t B[C::STAGES][C::B_BYTES];
  float bias[128];
  int2 tab[1024];               // (plane, shift) of every copy, staged once: the loader thread must not chase global loads
  uint64_t full[C::STAGES], empty[C::STAGES], d_full;
  uint32_t tmem_base;
};

// One "run" = a 1-D sequence of rows (pixels) whose channels live in C8-interleaved planes
// [plane][row][8] with `a_pitch` rows per plane.  blockIdx.x = 512-row tile inside the run,
// blockIdx.y = run (image row for the 2-D leg layers), blockIdx.z = 128-wide N slice.
// K loop = slabs of four (plane, row shift) copies: the row shift is the convolution tap along
// the run (implicit im2col at the copy level), the plane picks (input row tap, channel chunk).
struct GemmArgs {
  const __half* A;
  int64_t a_pitch;            // rows per input plane
  const int* copy_plane;      // [n_slabs*4] plane of each copy, relative to the run's first plane
  const int* copy_shift;      // [n_slabs*4] row shift of each copy
  int n_slabs;
  const __half* Bp;           // [gridDim.z][n_slabs][4][NT][8]
  const float* bias;          // [gridDim.z * NT]
  int64_t M;                  // valid rows per run
  int runs_per_img;           // blockIdx.y = img * runs_per_img + run
  int in_img_planes;          // planes per input image
  int in_run_planes;          // plane advance per run (stride_h * C8_in)
  // epilogue 1: bias + ReLU -> fp16 planes [y * out_run_planes + n/8][row][8]
  __half* out_planes; int64_t out_pitch; int out_run_planes;
  // epilogue 2: bias + ReLU -> dot with the Dense kernel -> per-row partial sums
  const float* wd; float* partial; int grid_w, valid_w, valid_h, n_total;
  // epilogue 3: bias + ReLU -> fp32 row-major [y][row][NT]
  float* out_f32;
  int n_valid;                // output channels actually present (<= NT); 0 = NT
};

template <int EPI, int NT, int TILES>
__global__ void __launch_bounds__(G_THREADS, 1)
k_gemm_stream_tc(GemmArgs g, int* __restrict__ err) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  using C = GCfg<NT, TILES>;
  GSmem<NT, TILES>& S = *reinterpret_cast<GSmem<NT, TILES>*>(smem_raw);
  constexpr int B_BYTES = C::B_BYTES, G_ROWS = C::ROWS, G_A_BYTES = C::A_BYTES, G_STAGES = C::STAGES;
  constexpr uint32_t TMEM_COLS = (TILES * NT) <= 64 ? 64 : ((TILES * NT) <= 128 ? 128 : ((TILES * NT) <= 256 ? 256 : 512));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t row0 = (int64_t)blockIdx.x * G_ROWS;
  const int y = blockIdx.y, nh = blockIdx.z;
  const int64_t in_base = (int64_t)(y / g.runs_per_img) * g.in_img_planes + (int64_t)(y % g.runs_per_img) * g.in_run_planes;

  if (tid == 0) {
    for (int s = 0; s < G_STAGES; ++s) { mbar_init(&S.full[s], 1); mbar_init(&S.empty[s], 1); }
    mbar_init(&S.d_full, 1);
    mbar_fence_init();
  }
  if (tid < NT) S.bias[tid] = (g.n_valid == 0 || nh * NT + tid < g.n_valid) ? g.bias[nh * NT + tid] : 0.f;
  for (int e = tid; e < g.n_slabs * 4; e += G_THREADS) S.tab[e] = make_int2(g.copy_plane[e], g.copy_shift[e]);
  if (warp == 2) tmem_alloc(&S.tmem_base, TMEM_COLS);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = S.tmem_base;

  if (warp == 0) {
    // the whole warp walks the ring; lanes 0..3 each issue one activation-plane copy and lane 4 the
    // weight slab, so a slab costs one bulk-copy issue latency instead of five in a row
    uint32_t s = 0, ph = 0;
    for (int sl = 0; sl < g.n_slabs; ++sl) {
      TC_WAIT(&S.empty[s], ph ^ 1, 501);
      if (lane == 0) mbar_arrive_expect_tx(&S.full[s], G_A_BYTES + B_BYTES);
      __syncwarp();
      if (lane < 4) {
        const int2 ps = S.tab[sl * 4 + lane];
        const int64_t plane = in_base + ps.x;
        const int64_t r = row0 + ps.y;
        bulk_g2s(S.A[s] + lane * (G_ROWS * 16), g.A + ((size_t)plane * g.a_pitch + r) * 8, G_ROWS * 16, &S.full[s]);
      } else if (lane == 4) {
        bulk_g2s(S.B[s], g.Bp + ((size_t)nh * g.n_slabs + sl) * (B_BYTES / 2), B_BYTES, &S.full[s]);
      }
      if (++s == G_STAGES) { s = 0; ph ^= 1; }
    }
  } else if (warp == 1) {
    // warp-uniform loop, one elected lane issues; descriptors advance by 32-bit adds
    const uint32_t idesc = make_idesc_f16(128, NT);
    const bool leader = elect_one() != 0;
    const uint64_t ad0 = make_desc_kmajor_noswizzle(smem_u32(S.A[0]), G_ROWS * 16, 128);
    const uint64_t bd0 = make_desc_kmajor_noswizzle(smem_u32(S.B[0]), NT * 16, 128);
    const uint32_t ad_hi = (uint32_t)(ad0 >> 32), ad_lo = (uint32_t)ad0;
    const uint32_t bd_hi = (uint32_t)(bd0 >> 32), bd_lo = (uint32_t)bd0;
    uint32_t sg = 0, ph = 0;
    for (int sl = 0; sl < g.n_slabs; ++sl) {
      TC_WAIT(&S.full[sg], ph, 502);
      fence_after_sync();
      if (leader) {
        const uint32_t a_off = (sg * G_A_BYTES) >> 4, b_off = (sg * B_BYTES) >> 4;
#pragma unroll
        for (int t = 0; t < TILES; ++t) {
#pragma unroll
          for (int kk = 0; kk < 2; ++kk) {
            const uint64_t ad = ((uint64_t)ad_hi << 32) | (uint64_t)(ad_lo + a_off + ((t * 128 * 16 + kk * 2 * (G_ROWS * 16)) >> 4));
            const uint64_t bd = ((uint64_t)bd_hi << 32) | (uint64_t)(bd_lo + b_off + ((kk * 2 * (NT * 16)) >> 4));
            mma_ss(tmem + t * NT, ad, bd, idesc, (sl | kk) != 0);
          }
        }
        commit(&S.empty[sg]);
      }
      __syncwarp();
      if (++sg == G_STAGES) { sg = 0; ph ^= 1; }
    }
    if (leader) commit(&S.d_full);
    __syncwarp();
  } else if (warp >= 4) {
    const int q = warp & 3;
    TC_WAIT(&S.d_full, 0, 503);
    fence_after_sync();
#pragma unroll 1
    for (int t = 0; t < TILES; ++t) {
      const int64_t r = row0 + t * 128 + q * 32 + lane;
      if (EPI == 1 || EPI == 3 || EPI == 4) {
#pragma unroll 1
        for (int c0 = 0; c0 < NT; c0 += 16) {
          uint32_t v[16];
          tmem_ld_x16(tmem + ((uint32_t)(q * 32) << 16) + t * NT + c0, v);
          tmem_ld_wait();
          if (r < g.M && (g.n_valid == 0 || nh * NT + c0 < g.n_valid)) {
            if (EPI == 4) {
              // bias + ReLU -> hi/lo fp16 split planes (x = hi + lo to 2^-22): the next layer's
              // three-term product keeps the leg at fp32-grade accuracy on the fp16 tensor pipe
#pragma unroll
              for (int h8 = 0; h8 < 2; ++h8) {
                uint32_t ph[4], pl[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const int n = c0 + h8 * 8 + 2 * j;
                  const float a = fmaxf(__uint_as_float(v[h8 * 8 + 2 * j]) + S.bias[n], 0.f);
                  const float b = fmaxf(__uint_as_float(v[h8 * 8 + 2 * j + 1]) + S.bias[n + 1], 0.f);
                  const __half2 hi = __floats2half2_rn(a, b);
                  const float2 hf = __half22float2(hi);
                  const __half2 lo = __floats2half2_rn(a - hf.x, b - hf.y);
                  ph[j] = *reinterpret_cast<const uint32_t*>(&hi);
                  pl[j] = *reinterpret_cast<const uint32_t*>(&lo);
                }
                const int c8 = nh * (NT / 8) + (c0 >> 3) + h8;
                const int64_t plane = (int64_t)y * g.out_run_planes + c8;
                *reinterpret_cast<uint4*>(g.out_planes + ((size_t)plane * g.out_pitch + r) * 8) =
                    make_uint4(ph[0], ph[1], ph[2], ph[3]);
                *reinterpret_cast<uint4*>(g.out_planes + ((size_t)(plane + g.out_run_planes / 2) * g.out_pitch + r) * 8) =
                    make_uint4(pl[0], pl[1], pl[2], pl[3]);
              }
            } else if (EPI == 1) {
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
                const int64_t plane = (int64_t)y * g.out_run_planes + nh * (NT / 8) + (c0 >> 3) + h8;
                *reinterpret_cast<uint4*>(g.out_planes + ((size_t)plane * g.out_pitch + r) * 8) =
                    make_uint4(pk[0], pk[1], pk[2], pk[3]);
              }
            } else {
              float* dst = g.out_f32 + ((size_t)y * g.M + r) * (g.n_valid ? g.n_valid : NT) + nh * NT + c0;
#pragma unroll
              for (int j4 = 0; j4 < 4; ++j4) {
                float4 o;
                o.x = fmaxf(__uint_as_float(v[j4 * 4 + 0]) + S.bias[c0 + j4 * 4 + 0], 0.f);
                o.y = fmaxf(__uint_as_float(v[j4 * 4 + 1]) + S.bias[c0 + j4 * 4 + 1], 0.f);
                o.z = fmaxf(__uint_as_float(v[j4 * 4 + 2]) + S.bias[c0 + j4 * 4 + 2], 0.f);
                o.w = fmaxf(__uint_as_float(v[j4 * 4 + 3]) + S.bias[c0 + j4 * 4 + 3], 0.f);
                reinterpret_cast<float4*>(dst)[j4] = o;
              }
            }
          }
        }
      } else {
        // row r = pair * (grid_w*grid_w) + y * grid_w + x ; valid output pixel iff y < valid_h, x < valid_w
        const int per = g.grid_w * g.grid_w;
        const int rem = (int)(r % per);
        const int yy = rem / g.grid_w, xx = rem - yy * g.grid_w;
        const bool valid = (r < g.M) && (yy < g.valid_h) && (xx < g.valid_w);
        const float* wrow = g.wd + ((size_t)(valid ? (yy * g.valid_w + xx) : 0) * g.n_total + nh * NT);
        float acc = 0.f;
#pragma unroll 1
        for (int c0 = 0; c0 < NT; c0 += 16) {
          uint32_t v[16];
          tmem_ld_x16(tmem + ((uint32_t)(q * 32) << 16) + t * NT + c0, v);
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
  if (warp == 2) tmem_dealloc(tmem, TMEM_COLS);
}

// ------------------------------------------------------------------------------------------------
// k_conv3_resident_tc -- c_conv3 (3x3, 128 -> 256, ReLU) + Flatten + Dense partial sums.
// The streamed GEMM re-reads the activation tile for each of the 9 taps and both N halves
// (3.6 GB of L2 traffic per 1101 pairs: L2-bound).  Here the 16 activation planes of a 512-row tile
// (+64 halo rows) are loaded ONCE (144 KB) and the 3x3 window is applied by the UMMA descriptor
// itself: tap (dy, dx) is a start-address offset of (dy*24 + dx) rows in the SWIZZLE_NONE layout
// (16-byte rows at uniform pitch -- probe mode 2).  Only the weights (8 KB per slab) are streamed.
// ------------------------------------------------------------------------------------------------
constexpr int C3_ROWS = 512, C3_WIN = 576, C3_PLANES = 16, C3_SLABS = 36, C3_STAGES = 4;
constexpr int C3_PLANE_BYTES = C3_WIN * 16;             // 9216
constexpr int C3_B_BYTES = 4 * 128 * 16;                // 8192

struct C3Smem {
  uint8_t A[C3_PLANES][C3_PLANE_BYTES];
  uint8_t B[C3_STAGES][C3_B_BYTES];
  float bias[128];
  uint64_t a_full, full[C3_STAGES], empty[C3_STAGES], d_full;
  uint32_t tmem_base;
};

__global__ void __launch_bounds__(G_THREADS, 1)
k_conv3_resident_tc(const __half* __restrict__ X3, int64_t a_pitch, const __half* __restrict__ Bp,
                    const float* __restrict__ bias, int64_t M, const float* __restrict__ wd, float* __restrict__ partial,
                    int* __restrict__ err) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  C3Smem& S = *reinterpret_cast<C3Smem*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t row0 = (int64_t)blockIdx.x * C3_ROWS;
  const int nh = blockIdx.z;
  if (tid == 0) {
    mbar_init(&S.a_full, 1);
    for (int s = 0; s < C3_STAGES; ++s) { mbar_init(&S.full[s], 1); mbar_init(&S.empty[s], 1); }
    mbar_init(&S.d_full, 1);
    mbar_fence_init();
  }
  if (tid < 128) S.bias[tid] = bias[nh * 128 + tid];
  if (warp == 2) tmem_alloc(&S.tmem_base, 512);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = S.tmem_base;

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(&S.a_full, C3_PLANES * C3_PLANE_BYTES);
      for (int pl = 0; pl < C3_PLANES; ++pl)
        bulk_g2s(S.A[pl], X3 + ((size_t)pl * a_pitch + row0) * 8, C3_PLANE_BYTES, &S.a_full);
      uint32_t s = 0, ph = 0;
      for (int sl = 0; sl < C3_SLABS; ++sl) {
        TC_WAIT(&S.empty[s], ph ^ 1, 701);
        mbar_arrive_expect_tx(&S.full[s], C3_B_BYTES);
        bulk_g2s(S.B[s], Bp + ((size_t)nh * C3_SLABS + sl) * (C3_B_BYTES / 2), C3_B_BYTES, &S.full[s]);
        if (++s == C3_STAGES) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = make_idesc_f16(128, 128);
    const bool leader = elect_one() != 0;
    const uint64_t ad0 = make_desc_kmajor_noswizzle(smem_u32(S.A[0]), C3_PLANE_BYTES, 128);
    const uint64_t bd0 = make_desc_kmajor_noswizzle(smem_u32(S.B[0]), 128 * 16, 128);
    const uint32_t ad_hi = (uint32_t)(ad0 >> 32), ad_lo = (uint32_t)ad0;
    const uint32_t bd_hi = (uint32_t)(bd0 >> 32), bd_lo = (uint32_t)bd0;
    TC_WAIT(&S.a_full, 0, 702);
    uint32_t sg = 0, ph = 0;
#pragma unroll 1
    for (int tap = 0; tap < 9; ++tap) {
      const uint32_t shift = (tap / 3) * NB + (tap % 3);                // rows
#pragma unroll 1
      for (int gq = 0; gq < 4; ++gq) {                                  // slab = (tap, 32-channel group)
        TC_WAIT(&S.full[sg], ph, 703);
        fence_after_sync();
        if (leader) {
          const uint32_t b_off = (sg * C3_B_BYTES) >> 4;
#pragma unroll
          for (int kk = 0; kk < 2; ++kk) {
            const uint64_t bd = ((uint64_t)bd_hi << 32) | (uint64_t)(bd_lo + b_off + ((kk * 2 * (128 * 16)) >> 4));
            const uint32_t a_k = ad_lo + (((gq * 4 + kk * 2) * C3_PLANE_BYTES + shift * 16) >> 4);
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              const uint64_t ad = ((uint64_t)ad_hi << 32) | (uint64_t)(a_k + ((t * 128 * 16) >> 4));
              mma_ss(tmem + t * 128, ad, bd, idesc, (tap | gq | kk) != 0);
            }
          }
          commit(&S.empty[sg]);
        }
        __syncwarp();
        if (++sg == C3_STAGES) { sg = 0; ph ^= 1; }
      }
    }
    if (leader) commit(&S.d_full);
    __syncwarp();
  } else if (warp >= 4) {
    const int q = warp & 3;
    TC_WAIT(&S.d_full, 0, 704);
    fence_after_sync();
#pragma unroll 1
    for (int t = 0; t < 4; ++t) {
      const int64_t r = row0 + t * 128 + q * 32 + lane;
      const int rem = (int)(r % PAIR_ROWS);
      const int yy = rem / NB, xx = rem - yy * NB;
      const bool valid = (r < M) && (yy < NB - 2) && (xx < NB - 2);
      const float* wrow = wd + ((size_t)(valid ? (yy * (NB - 2) + xx) : 0) * 256 + nh * 128);
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
      if (r < M) partial[r * 2 + nh] = valid ? acc : 0.f;
    }
  }
done:
  fence_before_sync();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem, 512);
}

// ------------------------------------------------------------------------------------------------
// k_leg_resident_tc -- one leg layer in latency mode (single query scan).
// Measured (profiles/r1_ncu_summary): the streamed GEMM needs 5 small bulk copies per K slab and a
// CTA cannot get more than ~1 bulk copy per ~150 clk through, so a 108-slab layer took 45 us at 9 %
// tensor activity.  Here the activation window of the CTA's 128 output pixels (kh input rows x
// [hi, lo] x C_in/8 planes x (128 + kw - 1) pixels, <= 108 KB) is loaded once and every (dh, dw) tap
// is a descriptor offset into it; only the weights move: one copy per (tap, split term) of
// C_in x 64 x 2 B.  Three-term hi/lo split product as in the streamed leg (fp32-grade accuracy).
// ------------------------------------------------------------------------------------------------
constexpr int LR_WIN = 144;                       // 128 + max(kw) - 1 = 142, rounded to a multiple of 8
constexpr int LR_A_MAX = 48 * LR_WIN * 16;        // 110 592 B (s_conv3a: 3 rows x 2 x 8 planes)
constexpr int LR_B_MAX = 16 * 64 * 16;            // 16 384 B  (C_in = 128)
constexpr int LR_STAGES = 5;

struct LRSmem {
  uint8_t A[LR_A_MAX];
  uint8_t B[LR_STAGES][LR_B_MAX];
  float bias[64];
  uint64_t a_full, full[LR_STAGES], empty[LR_STAGES], d_full;
  uint32_t tmem_base;
};

struct LegArgs {
  const __half* A; int64_t a_pitch;
  int runs_per_img, in_img_planes, in_run_planes;
  int kh, kw, c8in;
  const __half* Bp;           // [cout/64][kh*kw*3][c8in][64][8]
  const float* bias; int n_valid;
  int64_t M;                  // output pixels per run
  __half* out_planes; int64_t out_pitch; int out_run_planes;   // EPI 4
  float* out_f32;                                               // EPI 3
};

template <int EPI>
__global__ void __launch_bounds__(G_THREADS, 1)
k_leg_resident_tc(LegArgs g, int* __restrict__ err) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  LRSmem& S = *reinterpret_cast<LRSmem*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t row0 = (int64_t)blockIdx.x * 128;
  const int y = blockIdx.y, nh = blockIdx.z;
  const int64_t in_base = (int64_t)(y / g.runs_per_img) * g.in_img_planes + (int64_t)(y % g.runs_per_img) * g.in_run_planes;
  const int n_planes = g.kh * 2 * g.c8in;
  const int n_slabs = g.kh * g.kw * 3;
  const uint32_t b_bytes = (uint32_t)g.c8in * 64 * 16;

  if (tid == 0) {
    mbar_init(&S.a_full, 1);
    for (int s = 0; s < LR_STAGES; ++s) { mbar_init(&S.full[s], 1); mbar_init(&S.empty[s], 1); }
    mbar_init(&S.d_full, 1);
    mbar_fence_init();
  }
  if (tid < 64) S.bias[tid] = (nh * 64 + tid < g.n_valid) ? g.bias[nh * 64 + tid] : 0.f;
  if (warp == 2) tmem_alloc(&S.tmem_base, 64);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = S.tmem_base;

  if (warp == 0) {
    // activation window: one copy per plane, issued by the lanes in parallel
    if (lane == 0) mbar_arrive_expect_tx(&S.a_full, (uint32_t)n_planes * LR_WIN * 16);
    __syncwarp();
    for (int pl = lane; pl < n_planes; pl += 32)
      bulk_g2s(S.A + (size_t)pl * LR_WIN * 16, g.A + ((size_t)(in_base + pl) * g.a_pitch + row0) * 8, LR_WIN * 16, &S.a_full);
    uint32_t s = 0, ph = 0;
    for (int sl = 0; sl < n_slabs; ++sl) {
      TC_WAIT(&S.empty[s], ph ^ 1, 801);
      if (lane == 0) {
        mbar_arrive_expect_tx(&S.full[s], b_bytes);
        bulk_g2s(S.B[s], g.Bp + ((size_t)nh * n_slabs + sl) * (b_bytes / 2), b_bytes, &S.full[s]);
      }
      __syncwarp();
      if (++s == LR_STAGES) { s = 0; ph ^= 1; }
    }
  } else if (warp == 1) {
    const uint32_t idesc = make_idesc_f16(128, 64);
    const bool leader = elect_one() != 0;
    const uint64_t ad0 = make_desc_kmajor_noswizzle(smem_u32(S.A), LR_WIN * 16, 128);
    const uint64_t bd0 = make_desc_kmajor_noswizzle(smem_u32(S.B[0]), 64 * 16, 128);
    const uint32_t ad_hi = (uint32_t)(ad0 >> 32), ad_lo = (uint32_t)ad0;
    const uint32_t bd_hi = (uint32_t)(bd0 >> 32), bd_lo = (uint32_t)bd0;
    TC_WAIT(&S.a_full, 0, 802);
    uint32_t sg = 0, ph = 0, first = 1;
    for (int dh = 0; dh < g.kh; ++dh) {
      for (int dw = 0; dw < g.kw; ++dw) {
#pragma unroll 1
        for (int term = 0; term < 3; ++term) {            // x*w ~= xh*wh + xl*wh + xh*wl
          TC_WAIT(&S.full[sg], ph, 803);
          fence_after_sync();
          if (leader) {
            const uint32_t kind = (term == 1) ? 1u : 0u;
            const uint32_t a_k = ad_lo + ((((dh * 2 + kind) * g.c8in) * (LR_WIN * 16) + dw * 16) >> 4);
            const uint32_t b_k = bd_lo + ((sg * LR_B_MAX) >> 4);
            for (int c16 = 0; c16 < g.c8in / 2; ++c16) {
              const uint64_t ad = ((uint64_t)ad_hi << 32) | (uint64_t)(a_k + ((c16 * 2 * (LR_WIN * 16)) >> 4));
              const uint64_t bd = ((uint64_t)bd_hi << 32) | (uint64_t)(b_k + ((c16 * 2 * (64 * 16)) >> 4));
              mma_ss(tmem, ad, bd, idesc, first ? 0u : 1u);
              first = 0;
            }
            commit(&S.empty[sg]);
          }
          first = 0;
          __syncwarp();
          if (++sg == LR_STAGES) { sg = 0; ph ^= 1; }
        }
      }
    }
    if (leader) commit(&S.d_full);
    __syncwarp();
  } else if (warp >= 4) {
    const int q = warp & 3;
    TC_WAIT(&S.d_full, 0, 804);
    fence_after_sync();
    const int64_t r = row0 + q * 32 + lane;
#pragma unroll 1
    for (int c0 = 0; c0 < 64; c0 += 16) {
      uint32_t v[16];
      tmem_ld_x16(tmem + ((uint32_t)(q * 32) << 16) + c0, v);
      tmem_ld_wait();
      if (r < g.M && nh * 64 + c0 < g.n_valid) {
        if (EPI == 4) {
#pragma unroll
          for (int h8 = 0; h8 < 2; ++h8) {
            uint32_t phh[4], pll[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int n = c0 + h8 * 8 + 2 * j;
              const float a = fmaxf(__uint_as_float(v[h8 * 8 + 2 * j]) + S.bias[n], 0.f);
              const float b = fmaxf(__uint_as_float(v[h8 * 8 + 2 * j + 1]) + S.bias[n + 1], 0.f);
              const __half2 hi = __floats2half2_rn(a, b);
              const float2 hf = __half22float2(hi);
              const __half2 lo = __floats2half2_rn(a - hf.x, b - hf.y);
              phh[j] = *reinterpret_cast<const uint32_t*>(&hi);
              pll[j] = *reinterpret_cast<const uint32_t*>(&lo);
            }
            const int c8 = nh * 8 + (c0 >> 3) + h8;
            const int64_t plane = (int64_t)y * g.out_run_planes + c8;
            *reinterpret_cast<uint4*>(g.out_planes + ((size_t)plane * g.out_pitch + r) * 8) = make_uint4(phh[0], phh[1], phh[2], phh[3]);
            *reinterpret_cast<uint4*>(g.out_planes + ((size_t)(plane + g.out_run_planes / 2) * g.out_pitch + r) * 8) =
                make_uint4(pll[0], pll[1], pll[2], pll[3]);
          }
        } else {
          float* dst = g.out_f32 + ((size_t)y * g.M + r) * g.n_valid + nh * 64 + c0;
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) {
            float4 o;
            o.x = fmaxf(__uint_as_float(v[j4 * 4 + 0]) + S.bias[c0 + j4 * 4 + 0], 0.f);
            o.y = fmaxf(__uint_as_float(v[j4 * 4 + 1]) + S.bias[c0 + j4 * 4 + 1], 0.f);
            o.z = fmaxf(__uint_as_float(v[j4 * 4 + 2]) + S.bias[c0 + j4 * 4 + 2], 0.f);
            o.w = fmaxf(__uint_as_float(v[j4 * 4 + 3]) + S.bias[c0 + j4 * 4 + 3], 0.f);
            reinterpret_cast<float4*>(dst)[j4] = o;
          }
        }
      }
    }
  }
done:
  fence_before_sync();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem, 64);
}

// ------------------------------------------------------------------------------------------------
// Correlation (yaw) head on tensor cores.
//   G = L R^T (360 x 360, K = 128), corr[k] = sum_j G[(k + j + 180) mod 360, j]
//   (RangePadding2D.py:34 + NormalizedCorrelation2D.py:96-109), yaw = 180 - argmax (infer.py:158).
// fp16 alone (10-bit mantissa) is not enough to keep the argmax of a flat correlation curve, so the
// operands are split hi/lo (x = hi + lo exactly to 2^-22) and G = Lhi Rhi + Llo Rhi + Lhi Rlo is
// accumulated in fp32 in TMEM: three tcgen05.mma per K16 step, fp32-grade result, still < 5 % of
// c_conv1's tensor time.  One CTA owns one half of R's rows (N = 192, zero-padded past 360) for all
// of its pairs; L arrives as 32 KB stages (row tile x K half, hi+lo) through a 3-deep bulk-copy
// ring.  The diagonal sums never touch memory: each epilogue warp reads its 32 rows of a finished
// 128 x 192 tile from TMEM and rotates a running accumulator across lanes (bin(lane, col+1) ==
// bin(lane-1, col)), flushing one finished bin per column.
// ------------------------------------------------------------------------------------------------
constexpr int C6_THREADS = 256;
constexpr int C6_STAGES = 3;
constexpr int C6_STAGE_BYTES = 32768;            // [hi,lo][8 planes][128 rows][8] fp16
constexpr int C6_R_BYTES = 98304;                // [hi,lo][16 planes][192 rows][8] fp16
constexpr int C6_VOL_L_BYTES = 6 * C6_STAGE_BYTES;
constexpr int C6_VOL_R_BYTES = 2 * C6_R_BYTES;

struct C6Smem {
  uint8_t R[C6_R_BYTES];
  uint8_t A[C6_STAGES][C6_STAGE_BYTES];
  float corr[4][WF];
  uint64_t full[C6_STAGES], empty[C6_STAGES], d_full[2], d_empty[2], r_full, r_empty, epi;
  uint32_t tmem_base;
};

// fp32 volumes -> hi/lo fp16 split in the stage layout of k_corr_tc (zero rows past 360)
__global__ void __launch_bounds__(256)
k_pack_corr_L(const float* __restrict__ bank, const int32_t* __restrict__ idx, int n, __half* __restrict__ out) {
  // one thread per (pair, tile, khalf, plane j, row): 8 channels
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t per = 3 * 2 * 8 * 128;
  if (i >= (int64_t)n * per) return;
  const int p = (int)(i / per);
  int r = (int)(i % per);
  const int row = r % 128; r /= 128;
  const int j = r % 8; r /= 8;
  const int kh = r % 2; r /= 2;
  const int t = r;
  const int vrow = t * 128 + row;
  const int c = kh * 64 + j * 8;
  float v[8];
  if (vrow < WF) {
    const int64_t src = (idx ? idx[p] : p);
    const float4 a = __ldg(reinterpret_cast<const float4*>(bank + (src * WF + vrow) * CF + c));
    const float4 b = __ldg(reinterpret_cast<const float4*>(bank + (src * WF + vrow) * CF + c + 4));
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  } else {
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = 0.f;
  }
  __half hi[8], lo[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    hi[e] = __float2half_rn(v[e]);
    lo[e] = __float2half_rn(v[e] - __half2float(hi[e]));
  }
  __half* base = out + (size_t)p * (C6_VOL_L_BYTES / 2) + (size_t)(t * 2 + kh) * (C6_STAGE_BYTES / 2);
  *reinterpret_cast<uint4*>(base + ((size_t)j * 128 + row) * 8) = *reinterpret_cast<const uint4*>(hi);
  *reinterpret_cast<uint4*>(base + (C6_STAGE_BYTES / 4) + ((size_t)j * 128 + row) * 8) = *reinterpret_cast<const uint4*>(lo);
}

__global__ void __launch_bounds__(256)
k_pack_corr_R(const float* __restrict__ bank, const int32_t* __restrict__ idx, int n, __half* __restrict__ out) {
  // one thread per (pair, nhalf, plane, row): 8 channels
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t per = 2 * 16 * 192;
  if (i >= (int64_t)n * per) return;
  const int p = (int)(i / per);
  int r = (int)(i % per);
  const int row = r % 192; r /= 192;
  const int pl = r % 16; r /= 16;
  const int h = r;
  const int vrow = h * 192 + row;
  float v[8];
  if (vrow < WF) {
    const int64_t src = (idx ? idx[p] : p);
    const float4 a = __ldg(reinterpret_cast<const float4*>(bank + (src * WF + vrow) * CF + pl * 8));
    const float4 b = __ldg(reinterpret_cast<const float4*>(bank + (src * WF + vrow) * CF + pl * 8 + 4));
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  } else {
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = 0.f;
  }
  __half hi[8], lo[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    hi[e] = __float2half_rn(v[e]);
    lo[e] = __float2half_rn(v[e] - __half2float(hi[e]));
  }
  __half* base = out + (size_t)p * (C6_VOL_R_BYTES / 2) + (size_t)h * (C6_R_BYTES / 2);
  *reinterpret_cast<uint4*>(base + ((size_t)pl * 192 + row) * 8) = *reinterpret_cast<const uint4*>(hi);
  *reinterpret_cast<uint4*>(base + (C6_R_BYTES / 4) + ((size_t)pl * 192 + row) * 8) = *reinterpret_cast<const uint4*>(lo);
}

__device__ __forceinline__ int wrap360(int x) {
  x %= WF;
  return x < 0 ? x + WF : x;
}

__global__ void __launch_bounds__(C6_THREADS, 1)
k_corr_tc(const __half* __restrict__ Lc, const int32_t* __restrict__ l_idx, const __half* __restrict__ Rc,
          int r_per_pair, int n_pairs,
          float* __restrict__ corr_part, int* __restrict__ err) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  C6Smem& S = *reinterpret_cast<C6Smem*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int h = blockIdx.x & 1;
  const int cta = blockIdx.x >> 1, n_cta = gridDim.x >> 1;

  if (tid == 0) {
    for (int s = 0; s < C6_STAGES; ++s) { mbar_init(&S.full[s], 1); mbar_init(&S.empty[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&S.d_full[b], 1); mbar_init(&S.d_empty[b], 4); }
    mbar_init(&S.r_full, 1); mbar_init(&S.r_empty, 1); mbar_init(&S.epi, 4);
    mbar_fence_init();
  }
  if (warp == 2) tmem_alloc(&S.tmem_base, 512);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = S.tmem_base;

  if (warp == 0) {
    if (lane == 0) {
      if (!r_per_pair) {
        mbar_arrive_expect_tx(&S.r_full, C6_R_BYTES);
        bulk_g2s(S.R, reinterpret_cast<const uint8_t*>(Rc) + (size_t)h * C6_R_BYTES, C6_R_BYTES, &S.r_full);
      }
      uint32_t it = 0, pi = 0;
      for (int p = cta; p < n_pairs; p += n_cta, ++pi) {
        if (r_per_pair) {
          TC_WAIT(&S.r_empty, (pi & 1) ^ 1, 601);
          mbar_arrive_expect_tx(&S.r_full, C6_R_BYTES);
          bulk_g2s(S.R, reinterpret_cast<const uint8_t*>(Rc) + (size_t)p * C6_VOL_R_BYTES + (size_t)h * C6_R_BYTES,
                   C6_R_BYTES, &S.r_full);
        }
        const int64_t lrow = l_idx ? l_idx[p] : p;
        for (int st = 0; st < 6; ++st, ++it) {
          const uint32_t s = it % C6_STAGES, ph = (it / C6_STAGES) & 1;
          TC_WAIT(&S.empty[s], ph ^ 1, 602);
          mbar_arrive_expect_tx(&S.full[s], C6_STAGE_BYTES);
          bulk_g2s(S.A[s], reinterpret_cast<const uint8_t*>(Lc) + (size_t)lrow * C6_VOL_L_BYTES + (size_t)st * C6_STAGE_BYTES,
                   C6_STAGE_BYTES, &S.full[s]);
        }
      }
    }
  } else if (warp == 1) {
    {
      const uint32_t idesc = make_idesc_f16(128, 192);
      const bool leader = elect_one() != 0;
      const uint64_t ad0 = make_desc_kmajor_noswizzle(smem_u32(S.A[0]), 2048, 128);
      const uint64_t bd0 = make_desc_kmajor_noswizzle(smem_u32(S.R), 3072, 128);
      const uint32_t ad_hi = (uint32_t)(ad0 >> 32), ad_lo = (uint32_t)ad0;
      const uint32_t bd_hi = (uint32_t)(bd0 >> 32), bd_lo = (uint32_t)bd0;
      uint32_t sg = 0, ph = 0, tileit = 0, pi = 0;
      if (!r_per_pair) { TC_WAIT(&S.r_full, 0, 603); }
      for (int p = cta; p < n_pairs; p += n_cta, ++pi) {
        if (r_per_pair) { TC_WAIT(&S.r_full, pi & 1, 604); }
        for (int t = 0; t < 3; ++t, ++tileit) {
          const uint32_t buf = tileit & 1;
          TC_WAIT(&S.d_empty[buf], ((tileit >> 1) & 1) ^ 1, 605);
          fence_after_sync();
#pragma unroll
          for (int kh = 0; kh < 2; ++kh) {
            TC_WAIT(&S.full[sg], ph, 606);
            fence_after_sync();
            if (leader) {
              const uint32_t a_off = (sg * C6_STAGE_BYTES) >> 4;
#pragma unroll
              for (int kk = 0; kk < 4; ++kk) {
                const uint64_t a_hi = ((uint64_t)ad_hi << 32) | (uint64_t)(ad_lo + a_off + ((kk * 2 * 2048) >> 4));
                const uint64_t a_lo = ((uint64_t)ad_hi << 32) | (uint64_t)(ad_lo + a_off + ((16384 + kk * 2 * 2048) >> 4));
                const uint64_t b_hi = ((uint64_t)bd_hi << 32) | (uint64_t)(bd_lo + (((kh * 8 + kk * 2) * 3072) >> 4));
                const uint64_t b_lo = ((uint64_t)bd_hi << 32) | (uint64_t)(bd_lo + ((49152 + (kh * 8 + kk * 2) * 3072) >> 4));
                mma_ss(tmem + buf * 192, a_hi, b_hi, idesc, (kh | kk) != 0);
                mma_ss(tmem + buf * 192, a_lo, b_hi, idesc, 1);
                mma_ss(tmem + buf * 192, a_hi, b_lo, idesc, 1);
              }
              commit(&S.empty[sg]);
            }
            __syncwarp();
            if (++sg == C6_STAGES) { sg = 0; ph ^= 1; }
          }
          if (leader) commit(&S.d_full[buf]);
          __syncwarp();
        }
        if (r_per_pair && leader) commit(&S.r_empty);
        __syncwarp();
      }
    }
  } else if (warp >= 4) {
    const int q = warp & 3;
    float* my = S.corr[q];
    uint32_t tileit = 0, rz = 0;      // rz: rendezvous count of the 4 epilogue warps (bounded mbarrier, never bar.sync)
    for (int p = cta; p < n_pairs; p += n_cta) {
      for (int k = lane; k < WF; k += 32) my[k] = 0.f;
      __syncwarp();
      for (int t = 0; t < 3; ++t, ++tileit) {
        const uint32_t buf = tileit & 1;
        TC_WAIT(&S.d_full[buf], (tileit >> 1) & 1, 607);
        fence_after_sync();
        // bin(lane, col) = (i - j - 180) mod 360 with i = t*128 + q*32 + lane, j = h*192 + col
        const int b31 = t * 128 + q * 32 + 31 - h * 192 - 180;     // bin of lane 31 at col 0 (before wrap)
        float acc = 0.f, pend = 0.f;
#pragma unroll 1
        for (int ch = 0; ch < 6; ++ch) {
          uint32_t v[32];
          tmem_ld_x32(tmem + ((uint32_t)(q * 32) << 16) + buf * 192 + ch * 32, v);
          tmem_ld_wait();
          if (ch == 5) {                     // all of this tile is in registers / consumed: release the buffer
            fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&S.d_empty[buf]);
          }
#pragma unroll
          for (int cc = 0; cc < 32; ++cc) {
            if (ch > 0 || cc > 0) {
              // retire lane 31's accumulator (the bin that leaves this warp) into lane (col-1)&31's pend
              const float out = __shfl_sync(0xffffffffu, acc, 31);
              if (lane == ((cc + 31) & 31)) pend = out;
              if (cc == 0) {
                // cols (ch-1)*32 .. ch*32-1 retired: lane l holds the bin of lane 31 at col (ch-1)*32 + l
                const int bin = wrap360(b31 - ((ch - 1) * 32 + lane));
                my[bin] += pend;
                __syncwarp();
              }
              acc = __shfl_up_sync(0xffffffffu, acc, 1);
              if (lane == 0) acc = 0.f;
            }
            acc += __uint_as_float(v[cc]);
          }
        }
        // tail: cols 160..190 retired into lanes 0..30 of pend; col 191's accumulators still in acc
        if (lane < 31) my[wrap360(b31 - (160 + lane))] += pend;
        __syncwarp();
        my[wrap360(b31 - 31 + lane - 191)] += acc;
        __syncwarp();
      }
      // combine the four warps' private arrays in a fixed order (bit-reproducible)
      __syncwarp();
      if (lane == 0) mbar_arrive(&S.epi);
      TC_WAIT(&S.epi, rz & 1, 608);
      ++rz;
      for (int k = tid - 128; k < WF; k += 128)
        corr_part[((size_t)p * 2 + h) * WF + k] = ((S.corr[0][k] + S.corr[1][k]) + S.corr[2][k]) + S.corr[3][k];
      __syncwarp();
      if (lane == 0) mbar_arrive(&S.epi);
      TC_WAIT(&S.epi, rz & 1, 609);
      ++rz;
    }
  }
done:
  fence_before_sync();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem, 512);
}

__global__ void __launch_bounds__(384)
k_corr_finalize(const float* __restrict__ part, float* __restrict__ corr_out, int32_t* __restrict__ yaw) {
  __shared__ float s_corr[WF];
  const int p = blockIdx.x;
  for (int k = threadIdx.x; k < WF; k += blockDim.x) {
    const float c = part[((size_t)p * 2) * WF + k] + part[((size_t)p * 2 + 1) * WF + k];
    s_corr[k] = c;
    if (corr_out) corr_out[(size_t)p * WF + k] = c;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int best = 0;
    float bv = s_corr[0];
    for (int k = 1; k < WF; ++k)
      if (s_corr[k] > bv) { bv = s_corr[k]; best = k; }
    yaw[p] = WF / 2 - best;
  }
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
                  t->l16, t->r16, t->o1, t->x3, t->partial, t->d_err, t->lc, t->rc, t->corr_part};
  for (void* b : bufs) if (b) cudaFree(b);
  for (int l = 0; l < kMaxLegLayers; ++l) {
    if (t->wleg[l]) cudaFree(t->wleg[l]);
    if (t->wleg64[l]) cudaFree(t->wleg64[l]);
    if (t->wres[l]) cudaFree(t->wres[l]);
    if (t->leg_plane[l]) cudaFree(t->leg_plane[l]);
    if (t->leg_shift[l]) cudaFree(t->leg_shift[l]);
  }
  if (t->pb_l16) cudaFree(t->pb_l16);
  if (t->pb_lc) cudaFree(t->pb_lc);
  if (t->actp[0]) cudaFree(t->actp[0]);
  if (t->actp[1]) cudaFree(t->actp[1]);
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
  std::vector<int> s2p(30 * 4), s2s(30 * 4, 0);
  for (int sl = 0; sl < 30; ++sl) {
    for (int j = 0; j < 4; ++j) {
      s2p[sl * 4 + j] = sl * 4 + j;
      const int k8 = sl * 4 + j, di = k8 / 8, o8 = k8 % 8;
      for (int n = 0; n < 128; ++n)
        for (int e = 0; e < 8; ++e)
          p2[(((size_t)sl * 4 + j) * 128 + n) * 8 + e] = __float2half(w2.kernel[((size_t)di * 64 + o8 * 8 + e) * 128 + n]);
    }
  }
  // c_conv3: slab = (dy, dx, channel group g of 32); planes c8 = g*4..g*4+3 of X3; shift = dy*24 + dx
  std::vector<__half> p3((size_t)2 * 36 * 4 * 128 * 8);
  std::vector<int> s3p(36 * 4), s3s(36 * 4);
  for (int dy = 0; dy < 3; ++dy)
    for (int dx = 0; dx < 3; ++dx)
      for (int gq = 0; gq < 4; ++gq) {
        const int sl = (dy * 3 + dx) * 4 + gq;
        for (int j = 0; j < 4; ++j) { s3p[sl * 4 + j] = gq * 4 + j; s3s[sl * 4 + j] = dy * NB + dx; }
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
  // ---- leg layers 2.. : copy e = (dh, dw, c8) -> plane dh*C8in + c8 of the run, row shift dw
  size_t max_planes_bytes = 0;
  for (int l = 0; l < h->n_leg; ++l) {
    const ConvSpec& L = h->leg[l];
    const size_t out_bytes = (size_t)L.h_out * 2 * (L.cout / 8) * L.w_out * 16;   // hi + lo planes
    if (out_bytes > max_planes_bytes) max_planes_bytes = out_bytes;
    if (l == 0) continue;
    if (L.cin % 8 != 0 || L.cout % 8 != 0 || L.sw != 1 || L.w_out > 512)
      OVN_SET_ERR(h, OVN_ERR_BAD_CONFIG, "tensor-core leg: layer %s shape not supported", L.name);
    const LayerWeights& w = h->host_w[L.name];
    // three-term split product  x*w ~= xh*wh + xl*wh + xh*wl  (x = xh + xl, w = wh + wl in fp16):
    // every (dh, dw, c8) tap becomes three copies (A plane hi/lo/hi, B rows wh/wh/wl)
    const int c8in = L.cin / 8, nt = L.cout > 64 ? 128 : 64;
    const int E0 = L.kh * L.kw * c8in, E = 3 * E0, n_slabs = (E + 3) / 4;
    std::vector<int> cp(n_slabs * 4, 0), cs(n_slabs * 4, 0);
    std::vector<__half> bp((size_t)n_slabs * 4 * nt * 8, __float2half(0.f));
    for (int e0 = 0; e0 < E0; ++e0) {
      const int c8 = e0 % c8in, dw = (e0 / c8in) % L.kw, dh = e0 / (c8in * L.kw);
      for (int term = 0; term < 3; ++term) {
        const int e = e0 * 3 + term;
        cp[e] = dh * (2 * c8in) + (term == 1 ? c8in : 0) + c8;      // planes per input row: [hi c8in][lo c8in]
        cs[e] = dw;
        for (int n = 0; n < L.cout; ++n)
          for (int k = 0; k < 8; ++k) {
            const float wf = w.kernel[(((size_t)dh * L.kw + dw) * L.cin + c8 * 8 + k) * L.cout + n];
            const __half wh = __float2half(wf);
            const __half wl = __float2half(wf - __half2float(wh));
            bp[((size_t)e * nt + n) * 8 + k] = (term == 2) ? wl : wh;
          }
      }
    }
    // the same weights sliced into 64-channel halves for the latency variant
    const int nz = (L.cout + 63) / 64;
    std::vector<__half> bp64((size_t)nz * n_slabs * 4 * 64 * 8, __float2half(0.f));
    for (int z = 0; z < nz; ++z)
      for (int e = 0; e < n_slabs * 4; ++e)
        for (int n = 0; n < 64 && z * 64 + n < nt; ++n)
          for (int k = 0; k < 8; ++k)
            bp64[(((size_t)z * n_slabs * 4 + e) * 64 + n) * 8 + k] = bp[((size_t)e * nt + z * 64 + n) * 8 + k];
    t->leg_slabs[l] = n_slabs;
    t->leg_nt[l] = nt;
    int rc2;
    if ((rc2 = upload_vec(h, &t->wleg[l], bp)) != OVN_OK) return rc2;
    if ((rc2 = upload_vec(h, &t->wleg64[l], bp64)) != OVN_OK) return rc2;
    {
      // resident-activation layout: slab = (dh, dw, term), rows = all C_in/8 chunks, 64 output channels
      const int nsl = L.kh * L.kw * 3;
      std::vector<__half> br((size_t)nz * nsl * c8in * 64 * 8, __float2half(0.f));
      for (int z = 0; z < nz; ++z)
        for (int dh = 0; dh < L.kh; ++dh)
          for (int dw = 0; dw < L.kw; ++dw)
            for (int term = 0; term < 3; ++term) {
              const int sl = (dh * L.kw + dw) * 3 + term;
              for (int c8 = 0; c8 < c8in; ++c8)
                for (int n = 0; n < 64 && z * 64 + n < L.cout; ++n)
                  for (int k = 0; k < 8; ++k) {
                    const float wf = w.kernel[(((size_t)dh * L.kw + dw) * L.cin + c8 * 8 + k) * L.cout + z * 64 + n];
                    const __half wh = __float2half(wf);
                    const __half wl = __float2half(wf - __half2float(wh));
                    br[((((size_t)z * nsl + sl) * c8in + c8) * 64 + n) * 8 + k] = (term == 2) ? wl : wh;
                  }
            }
      if ((rc2 = upload_vec(h, &t->wres[l], br)) != OVN_OK) return rc2;
    }
    if ((rc2 = upload_vec(h, &t->leg_plane[l], cp)) != OVN_OK) return rc2;
    if ((rc2 = upload_vec(h, &t->leg_shift[l], cs)) != OVN_OK) return rc2;
  }
  for (int b = 0; b < 2; ++b) {
    const size_t bytes = max_planes_bytes * h->cfg.max_batch_scans + 32768;   // + tile overrun slack
    OVN_CUDA(h, cudaMalloc(&t->actp[b], bytes));
    OVN_CUDA(h, cudaMemset(t->actp[b], 0, bytes));
  }
  const int64_t maxp = h->cfg.max_batch_pairs;
  t->rows_pad = maxp * PAIR_ROWS + 1024;           // tile overrun (512) + window shift (50) slack
  OVN_CUDA(h, cudaMalloc(&t->l16, (size_t)maxp * WF * K4_PITCH * sizeof(__half)));
  OVN_CUDA(h, cudaMalloc(&t->r16, (size_t)maxp * WF * K4_PITCH * sizeof(__half)));
  OVN_CUDA(h, cudaMemset(t->l16, 0, (size_t)maxp * WF * K4_PITCH * sizeof(__half)));
  OVN_CUDA(h, cudaMemset(t->r16, 0, (size_t)maxp * WF * K4_PITCH * sizeof(__half)));
  OVN_CUDA(h, cudaMalloc(&t->o1, (size_t)120 * t->rows_pad * 8 * sizeof(__half)));
  OVN_CUDA(h, cudaMalloc(&t->x3, (size_t)16 * t->rows_pad * 8 * sizeof(__half)));
  OVN_CUDA(h, cudaMalloc(&t->partial, (size_t)t->rows_pad * 2 * sizeof(float)));
  OVN_CUDA(h, cudaMalloc(&t->lc, (size_t)maxp * C6_VOL_L_BYTES));
  OVN_CUDA(h, cudaMalloc(&t->rc, (size_t)maxp * C6_VOL_R_BYTES));
  OVN_CUDA(h, cudaMalloc(&t->corr_part, (size_t)maxp * 2 * WF * sizeof(float)));
  OVN_CUDA(h, cudaFuncSetAttribute(k_corr_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(C6Smem)));
  OVN_CUDA(h, cudaMalloc(&t->d_err, sizeof(int)));
  OVN_CUDA(h, cudaMemset(t->d_err, 0, sizeof(int)));
  OVN_CUDA(h, cudaMemset(t->o1, 0, (size_t)120 * t->rows_pad * 8 * sizeof(__half)));
  OVN_CUDA(h, cudaMemset(t->x3, 0, (size_t)16 * t->rows_pad * 8 * sizeof(__half)));
  OVN_CUDA(h, cudaFuncSetAttribute(k_delta_conv1_tc<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(K4Smem)));
  OVN_CUDA(h, cudaFuncSetAttribute(k_delta_conv1_tc<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(K4Smem)));
#define OVN_GEMM_ATTR(E, N, T)                                                                                  \
  OVN_CUDA(h, cudaFuncSetAttribute(k_gemm_stream_tc<E, N, T>, cudaFuncAttributeMaxDynamicSharedMemorySize,     \
                                   (int)sizeof(GSmem<N, T>)))
  OVN_GEMM_ATTR(1, 128, 4); OVN_GEMM_ATTR(2, 128, 4);
  OVN_CUDA(h, cudaFuncSetAttribute(k_conv3_resident_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(C3Smem)));
  OVN_GEMM_ATTR(4, 64, 4); OVN_GEMM_ATTR(4, 128, 4); OVN_GEMM_ATTR(3, 128, 4);
  OVN_GEMM_ATTR(4, 64, 1); OVN_GEMM_ATTR(3, 64, 1);
  OVN_CUDA(h, cudaFuncSetAttribute(k_leg_resident_tc<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(LRSmem)));
  OVN_CUDA(h, cudaFuncSetAttribute(k_leg_resident_tc<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(LRSmem)));
#undef OVN_GEMM_ATTR
  return OVN_OK;
}

// fp32 NHWC [n][H][W][C] -> hi/lo fp16 C8-interleaved planes [n][H][hi,lo][C/8][W][8]
__global__ void __launch_bounds__(256)
k_nhwc_to_planes(const float* __restrict__ x, int64_t total_chunks, int H, int W, int C8, __half* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;      // (img, h, c8, w)
  if (i >= total_chunks) return;
  const int w = (int)(i % W);
  int64_t r = i / W;
  const int c8 = (int)(r % C8); r /= C8;                                   // r = img*H + h
  const float* src = x + (r * W + w) * (int64_t)(C8 * 8) + c8 * 8;
  const float4 a = __ldg(reinterpret_cast<const float4*>(src));
  const float4 b = __ldg(reinterpret_cast<const float4*>(src + 4));
  const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
  __half hi[8], lo[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    hi[e] = __float2half_rn(v[e]);
    lo[e] = __float2half_rn(v[e] - __half2float(hi[e]));
  }
  const int64_t plane_hi = r * (2 * C8) + c8;
  *reinterpret_cast<uint4*>(out + ((size_t)plane_hi * W + w) * 8) = *reinterpret_cast<const uint4*>(hi);
  *reinterpret_cast<uint4*>(out + ((size_t)(plane_hi + C8) * W + w) * 8) = *reinterpret_cast<const uint4*>(lo);
}

int leg_forward_tc(ovn_handle* h, const float* d_input, int n, float* d_fv, cudaStream_t s) {
  // layer 1 (C_in = 4..25, stride (2,2)) stays on the fp32 SIMT kernel; layers 2..10 run on
  // k_gemm_stream_tc over C8-interleaved fp16 activation planes: one CTA per output image row,
  // the kw taps are row shifts of the bulk copies, the kh taps select the input row's planes.
  TcState* t = h->tc;
  if (!t) OVN_SET_ERR(h, OVN_ERR_WEIGHTS, "tensor-core weights not packed");
  prof_mark(h, PROF_LEG, s);
  int rc = leg_layer_fp32(h, 0, d_input, h->d_act[0], n, s);
  if (rc != OVN_OK) return rc;
  {
    const ConvSpec& L = h->leg[0];
    const int64_t chunks = (int64_t)n * L.h_out * (L.cout / 8) * L.w_out;
    k_nhwc_to_planes<<<(unsigned)((chunks + 255) / 256), 256, 0, s>>>(h->d_act[0], chunks, L.h_out, L.w_out,
                                                                     L.cout / 8, t->actp[0]);
    OVN_LAUNCH_CHECK(h);
  }
  int cur = 0;
  for (int l = 1; l < h->n_leg; ++l) {
    const ConvSpec& L = h->leg[l];
    const bool last = (l == h->n_leg - 1);
    GemmArgs a = {};
    a.A = t->actp[cur]; a.a_pitch = L.w_in;
    a.copy_plane = t->leg_plane[l]; a.copy_shift = t->leg_shift[l]; a.n_slabs = t->leg_slabs[l];
    a.Bp = t->wleg[l]; a.bias = h->d_b[l]; a.M = L.w_out;
    a.runs_per_img = L.h_out; a.in_img_planes = L.h_in * 2 * (L.cin / 8); a.in_run_planes = L.sh * 2 * (L.cin / 8);
    a.out_planes = t->actp[cur ^ 1]; a.out_pitch = L.w_out; a.out_run_planes = 2 * (L.cout / 8);
    a.out_f32 = d_fv;
    a.n_valid = L.cout;
    // latency mode (few scans): one 128-row tile and one 64-channel slice per CTA, 12-deep ring;
    // throughput mode (batched encode): four tiles per CTA share every weight slab
    const bool latency = (int64_t)n * L.h_out * 4 <= h->sm_count;
    if (last && (L.cout != 128 || L.h_out != 1)) OVN_SET_ERR(h, OVN_ERR_BAD_CONFIG, "tensor-core leg: unexpected last layer");
    if (latency) {
      const dim3 grid((unsigned)((L.w_out + 127) / 128), (unsigned)(n * L.h_out), (unsigned)((L.cout + 63) / 64));
      LegArgs la = {};
      la.A = a.A; la.a_pitch = a.a_pitch; la.runs_per_img = a.runs_per_img; la.in_img_planes = a.in_img_planes;
      la.in_run_planes = a.in_run_planes; la.kh = L.kh; la.kw = L.kw; la.c8in = L.cin / 8; la.Bp = t->wres[l];
      la.bias = h->d_b[l]; la.n_valid = L.cout; la.M = L.w_out; la.out_planes = a.out_planes; la.out_pitch = a.out_pitch;
      la.out_run_planes = a.out_run_planes; la.out_f32 = d_fv;
      if (last) k_leg_resident_tc<3><<<grid, G_THREADS, sizeof(LRSmem), s>>>(la, t->d_err);
      else k_leg_resident_tc<4><<<grid, G_THREADS, sizeof(LRSmem), s>>>(la, t->d_err);
    } else {
      const dim3 grid(1, (unsigned)(n * L.h_out), 1);
      if (last) k_gemm_stream_tc<3, 128, 4><<<grid, G_THREADS, sizeof(GSmem<128, 4>), s>>>(a, t->d_err);
      else if (t->leg_nt[l] == 128) k_gemm_stream_tc<4, 128, 4><<<grid, G_THREADS, sizeof(GSmem<128, 4>), s>>>(a, t->d_err);
      else k_gemm_stream_tc<4, 64, 4><<<grid, G_THREADS, sizeof(GSmem<64, 4>), s>>>(a, t->d_err);
    }
    OVN_LAUNCH_CHECK(h);
    cur ^= 1;
  }
  prof_mark(h, PROF_LEG, s);
  return OVN_OK;
}

int tc_bank_release(ovn_handle* h, const float* d_bank) {
  TcState* t = h->tc;
  if (!t || (d_bank && t->pb_key != d_bank)) return OVN_OK;
  OVN_CUDA(h, cudaDeviceSynchronize());
  if (t->pb_l16) cudaFree(t->pb_l16);
  if (t->pb_lc) cudaFree(t->pb_lc);
  t->pb_l16 = t->pb_lc = nullptr;
  t->pb_key = nullptr;
  t->pb_cap = t->pb_rows = 0;
  return OVN_OK;
}

int tc_bank_prepare(ovn_handle* h, const float* d_bank, int64_t capacity, int64_t first, int64_t count, cudaStream_t s) {
  TcState* t = h->tc;
  if (!t) OVN_SET_ERR(h, OVN_ERR_WEIGHTS, "tensor-core weights not packed");
  if (t->pb_key != d_bank || capacity > t->pb_cap) {
    if (t->pb_key != nullptr || t->pb_l16) { int rc = tc_bank_release(h, nullptr); if (rc != OVN_OK) return rc; }
    OVN_CUDA(h, cudaMalloc(&t->pb_l16, (size_t)capacity * WF * K4_PITCH * sizeof(__half)));
    OVN_CUDA(h, cudaMalloc(&t->pb_lc, (size_t)capacity * C6_VOL_L_BYTES));
    OVN_CUDA(h, cudaMemsetAsync(t->pb_l16, 0, (size_t)capacity * WF * K4_PITCH * sizeof(__half), s));
    t->pb_key = d_bank;
    t->pb_cap = capacity;
    t->pb_rows = 0;
    if (first != 0) OVN_SET_ERR(h, OVN_ERR_INVALID_ARG, "ovn_bank_prepare: a new bank must be prepared from row 0");
  }
  if (first > t->pb_rows) OVN_SET_ERR(h, OVN_ERR_INVALID_ARG, "ovn_bank_prepare: rows [%lld, %lld) were never prepared",
                                      (long long)t->pb_rows, (long long)first);
  const int64_t per = (int64_t)WF * CF / 4, perL = 3 * 2 * 8 * 128;
  const float* src = d_bank + (size_t)first * WF * CF;
  k_gather_rows_f16<<<(unsigned)((count * per + 255) / 256), 256, 0, s>>>(src, nullptr, (int)count,
                                                                         t->pb_l16 + (size_t)first * WF * K4_PITCH);
  OVN_LAUNCH_CHECK(h);
  k_pack_corr_L<<<(unsigned)((count * perL + 255) / 256), 256, 0, s>>>(src, nullptr, (int)count,
                                                                      t->pb_lc + (size_t)first * (C6_VOL_L_BYTES / 2));
  OVN_LAUNCH_CHECK(h);
  if (first + count > t->pb_rows) t->pb_rows = first + count;
  return OVN_OK;
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
    // resident bank: the LEFT operand copies already exist, the kernels index them through `left`
    const bool resident = (t->pb_key == d_bank) && t->pb_rows > 0;
    const __half* l16 = resident ? t->pb_l16 : t->l16;
    const __half* lc = resident ? t->pb_lc : t->lc;
    const int32_t* lidx = resident ? left : nullptr;
    if (!resident) {
      k_gather_rows_f16<<<(unsigned)((np * per + 255) / 256), 256, 0, s>>>(d_bank, left, np, t->l16);
      OVN_LAUNCH_CHECK(h);
    }
    if (!d_query) {
      k_gather_rows_f16<<<(unsigned)((np * per + 255) / 256), 256, 0, s>>>(d_bank, right, np, t->r16);
      OVN_LAUNCH_CHECK(h);
    }
    const int grid4 = np < h->sm_count ? np : h->sm_count;
    static const int prod = getenv("OVN_K4_PROD") ? atoi(getenv("OVN_K4_PROD")) : 1;
    prof_mark(h, PROF_DELTA, s);
    if (prod == 1)
      k_delta_conv1_tc<1><<<grid4, 640, sizeof(K4Smem), s>>>(l16, lidx, t->r16, d_query ? 0 : 1, t->w1p, h->d_b[base + 0],
                                                             t->o1, t->rows_pad, np, t->d_err);
    else
      k_delta_conv1_tc<0><<<grid4, 768, sizeof(K4Smem), s>>>(l16, lidx, t->r16, d_query ? 0 : 1, t->w1p, h->d_b[base + 0],
                                                             t->o1, t->rows_pad, np, t->d_err);
    prof_mark(h, PROF_DELTA, s);
    OVN_LAUNCH_CHECK(h);
    const int64_t M = (int64_t)np * PAIR_ROWS;
    const unsigned gx = (unsigned)((M + 511) / 512);
    GemmArgs a2 = {};
    a2.A = t->o1; a2.a_pitch = t->rows_pad; a2.copy_plane = t->slab2_plane; a2.copy_shift = t->slab2_shift;
    a2.n_slabs = 30; a2.Bp = t->w2p; a2.bias = h->d_b[base + 1]; a2.M = M;
    a2.runs_per_img = 1; a2.in_img_planes = 0; a2.in_run_planes = 0;
    a2.out_planes = t->x3; a2.out_pitch = t->rows_pad; a2.out_run_planes = 16;
    prof_mark(h, PROF_CONV2, s);
    k_gemm_stream_tc<1, 128, 4><<<dim3(gx, 1, 1), G_THREADS, sizeof(GSmem<128, 4>), s>>>(a2, t->d_err);
    prof_mark(h, PROF_CONV2, s);
    OVN_LAUNCH_CHECK(h);
    GemmArgs a3 = {};
    a3.A = t->x3; a3.a_pitch = t->rows_pad; a3.copy_plane = t->slab3_plane; a3.copy_shift = t->slab3_shift;
    a3.n_slabs = 36; a3.Bp = t->w3p; a3.bias = h->d_b[base + 2]; a3.M = M;
    a3.runs_per_img = 1; a3.in_img_planes = 0; a3.in_run_planes = 0;
    a3.wd = h->d_w[base + 3]; a3.partial = t->partial; a3.grid_w = NB; a3.valid_w = NB - 2; a3.valid_h = NB - 2;
    a3.n_total = 256;
    prof_mark(h, PROF_CONV3, s);
    k_conv3_resident_tc<<<dim3(gx, 1, 2), G_THREADS, sizeof(C3Smem), s>>>(t->x3, t->rows_pad, t->w3p, h->d_b[base + 2], M,
                                                                        h->d_w[base + 3], t->partial, t->d_err);
    prof_mark(h, PROF_CONV3, s);
    OVN_LAUNCH_CHECK(h);
    k_dense_finalize<<<np, 256, 0, s>>>(t->partial, h->d_b[base + 3], PAIR_ROWS, d_overlap + p0);
    OVN_LAUNCH_CHECK(h);
    // correlation head (tensor cores, hi/lo split operands)
    {
      const int64_t perL = 3 * 2 * 8 * 128, perR = 2 * 16 * 192;
      if (!resident) {
        k_pack_corr_L<<<(unsigned)((np * perL + 255) / 256), 256, 0, s>>>(d_bank, left, np, t->lc);
        OVN_LAUNCH_CHECK(h);
      }
      if (d_query) {
        if (p0 == 0) {
          k_pack_corr_R<<<(unsigned)((perR + 255) / 256), 256, 0, s>>>(d_query, nullptr, 1, t->rc);
          OVN_LAUNCH_CHECK(h);
        }
      } else {
        k_pack_corr_R<<<(unsigned)((np * perR + 255) / 256), 256, 0, s>>>(d_bank, right, np, t->rc);
        OVN_LAUNCH_CHECK(h);
      }
      int g6 = h->sm_count / 2;
      if (g6 > np) g6 = np;
      if (g6 < 1) g6 = 1;
      prof_mark(h, PROF_CORR, s);
      k_corr_tc<<<2 * g6, C6_THREADS, sizeof(C6Smem), s>>>(lc, lidx, t->rc, d_query ? 0 : 1, np, t->corr_part, t->d_err);
      prof_mark(h, PROF_CORR, s);
      OVN_LAUNCH_CHECK(h);
      k_corr_finalize<<<np, 384, 0, s>>>(t->corr_part, d_corr ? d_corr + (int64_t)p0 * WF : nullptr, d_yaw + p0);
      OVN_LAUNCH_CHECK(h);
    }
  }
  static const bool debug_sync = getenv("OVN_DEBUG_SYNC") != nullptr;
  if (debug_sync) return tc_check_error(h, s);
  return OVN_OK;
}

}  // namespace ovn
