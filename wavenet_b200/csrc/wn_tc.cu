// tcgen05 / TMEM / TMA kernels (kind::tf32, fp32 accumulate) for the training forward path.
//
//  tc_layer_kernel : one residual layer (wavenet.py:358-368 with the dilated conv of :294-342 in
//                    closed form) per launch, persistent over 128-position tiles:
//                      D1[128x128] = [x(t-d) | x(t)] . [Wf;Wg]^T        (tcgen05.mma, K=128)
//                      z = tanh(D1_f) * sigmoid(D1_g), zero prefix (Q1)  (epilogue, TMEM -> regs -> smem)
//                      D2[128x64]  = z . Wp^T                            (tcgen05.mma, K=64)
//                      x_out = D2 + x(t)                                 (epilogue)
//                    x tiles arrive by TMA (4-D tensor map, out-of-range rows zero-filled = causal pad),
//                    accumulators live in TMEM (double buffered), weights stay resident in shared memory.
//  tc_gemm_kernel  : Y = epi(A . W^T) with A gathered from `slabs` equally shaped tensors -- used for the
//                    skip sum  sum_l Ws_l z_l  (ONE GEMM with K = 64*L instead of L read-modify-write
//                    passes over the 256-channel skip tensor, wavenet.py:579) and for the head convs.
#include <cuda.h>

#include "wn_common.h"
#include "wn_tc.cuh"

namespace {

using namespace tc;

constexpr int TM = 128;        // positions per tile (UMMA M)
constexpr int SUBK = 32;       // tf32 elements per 128-byte swizzle row
constexpr int SUB_A = TM * 128;  // bytes of a [128 x 32] sub-tile

// ------------------------------------------------------------------------------------------
// weight preparation: TF32-rounded (RNA), K-major matrices the MMAs consume directly
//   w1[l][n][tap*R + c] = (n < G ? Wf : Wg)[n % G][c][tap]      n in [0, 2G)
//   w2[l][r][g]         = Wp[r][g]
//   ws[s][l*G + g]      = Ws_l[s][g]
struct TcTabEntry {
  int64_t wf, wg, wp, ws;
};

__global__ void tc_prep_kernel(const float* __restrict__ params, const TcTabEntry* __restrict__ tab, float* __restrict__ w1,
                               float* __restrict__ w2, float* __restrict__ wsc, int L, int R, int G, int S, int k) {
  const int l = blockIdx.y;
  const TcTabEntry e = tab[l];
  const int n1 = 2 * G * k * R, n2 = R * G, n3 = S * G;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n1 + n2 + n3; i += gridDim.x * blockDim.x) {
    if (i < n1) {
      const int n = i / (k * R), kk = i % (k * R);
      const int tap = kk / R, c = kk % R;
      const float* src = n < G ? params + e.wf : params + e.wg;
      w1[(int64_t)l * n1 + i] = tf32_rna(src[((int64_t)(n % G) * R + c) * k + tap]);
    } else if (i < n1 + n2) {
      const int j = i - n1;
      w2[(int64_t)l * n2 + j] = tf32_rna(params[e.wp + j]);
    } else {
      const int j = i - n1 - n2;
      const int sidx = j / G, g = j % G;
      wsc[(int64_t)sidx * (L * G) + (int64_t)l * G + g] = tf32_rna(params[e.ws + j]);
    }
  }
}

__global__ void tc_round_copy_kernel(const float* __restrict__ src, float* __restrict__ dst, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = tf32_rna(src[i]);
}

// in-place ReLU + TF32 rounding of the rows the head reads (wavenet.py:588)
__global__ void tc_relu_rows_kernel(float* __restrict__ a, int C, int rows_per_seq_in, int row_off, int rows_per_seq,
                                    int64_t rows) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int c4 = C / 4;
  if (i >= rows * c4) return;
  const int64_t row = i / c4;
  const int cc = (int)(i % c4);
  const int64_t seq = row / rows_per_seq;
  const int t = (int)(row % rows_per_seq);
  float4* p = reinterpret_cast<float4*>(a + ((seq * rows_per_seq_in + row_off + t) * (int64_t)C)) + cc;
  float4 v = *p;
  v.x = tf32_rna(fmaxf(v.x, 0.f));
  v.y = tf32_rna(fmaxf(v.y, 0.f));
  v.z = tf32_rna(fmaxf(v.z, 0.f));
  v.w = tf32_rna(fmaxf(v.w, 0.f));
  *p = v;
}

__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// ------------------------------------------------------------------------------------------
struct LayerArgs {
  float* x_out;            // [B][W][64]
  float* z_out;            // [B][W][64]
  const float* bias_fg;    // [128] or null
  const float* bias_p;     // [64] or null
  int W, d, zp, tiles_per_seq, num_tiles;
};

constexpr int L_B1 = 0;                       // 4 sub-tiles [128 x 32]  (64 KB)
constexpr int L_B2 = 65536;                   // 2 sub-tiles [64 x 32]   (16 KB)
constexpr int L_A = 81920;                    // 2 stages x 4 sub-tiles  (128 KB); Z aliases sub-tiles 0,1
constexpr int L_BAR = L_A + 2 * 65536;        // barriers
constexpr int L_SMEM = L_BAR + 256;
constexpr int L_THREADS = 64 + 256;           // producer warp, MMA warp, 8 epilogue warps

__global__ void __launch_bounds__(L_THREADS, 1)
tc_layer_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w1,
                const __grid_constant__ CUtensorMap tm_w2, const LayerArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bar0 = base + L_BAR;
  // barrier slots (8 bytes each)
  const uint32_t b_full = bar0;
  auto a_full = [&](int s) { return bar0 + 8 + 8 * s; };
  auto a_empty = [&](int s) { return bar0 + 24 + 8 * s; };
  auto d1_full = [&](int s) { return bar0 + 40 + 8 * s; };
  auto z_full = [&](int s) { return bar0 + 56 + 8 * s; };
  auto d2_full = [&](int s) { return bar0 + 72 + 8 * s; };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(gbase + L_BAR + 128);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(b_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(a_full(s), 1);
      mbar_init(a_empty(s), 256);
      mbar_init(d1_full(s), 1);
      mbar_init(z_full(s), 256);
      mbar_init(d2_full(s), 1);
    }
    fence_barrier_init();
    prefetch_tmap(&tm_x);
    prefetch_tmap(&tm_w1);
    prefetch_tmap(&tm_w2);
  }
  if (warp == 1) tmem_alloc<512>(smem_u32((const void*)tmem_slot));
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int n_local = (a.num_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(b_full, 81920);
      for (int j = 0; j < 4; ++j) tma_load_2d(base + L_B1 + j * SUB_A, &tm_w1, b_full, j * SUBK, 0);
      for (int j = 0; j < 2; ++j) tma_load_2d(base + L_B2 + j * 8192, &tm_w2, b_full, j * SUBK, 0);
      for (int j = 0; j < n_local; ++j) {
        const int tile = blockIdx.x + j * gridDim.x;
        const int s = j & 1, ph = (j >> 1) & 1;
        mbar_wait(a_empty(s), ph ^ 1);
        const int b = tile / a.tiles_per_seq, t0 = (tile % a.tiles_per_seq) * TM;
        const uint32_t as = base + L_A + s * 65536;
        mbar_arrive_expect_tx(a_full(s), 65536);
        tma_load_4d(as + 0 * SUB_A, &tm_x, a_full(s), 0, t0 - a.d, b, 0);
        tma_load_4d(as + 1 * SUB_A, &tm_x, a_full(s), SUBK, t0 - a.d, b, 0);
        tma_load_4d(as + 2 * SUB_A, &tm_x, a_full(s), 0, t0, b, 0);
        tma_load_4d(as + 3 * SUB_A, &tm_x, a_full(s), SUBK, t0, b, 0);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc1 = umma_idesc_tf32(128, 128);
      constexpr uint32_t idesc2 = umma_idesc_tf32(128, 64);
      mbar_wait(b_full, 0);
      auto issue1 = [&](int j) {
        const int s = j & 1, ph = (j >> 1) & 1;
        mbar_wait(a_full(s), ph);
        tcgen05_fence_after();
        const uint32_t as = base + L_A + s * 65536;
#pragma unroll
        for (int ks = 0; ks < 16; ++ks) {
          const uint32_t off = (ks >> 2) * SUB_A + (ks & 3) * 32;
          umma_tf32(tmem + s * 128, umma_desc_k_sw128(as + off), umma_desc_k_sw128(base + L_B1 + off), idesc1, ks > 0);
        }
        umma_commit(d1_full(s));
      };
      if (n_local > 0) issue1(0);
      for (int j = 0; j < n_local; ++j) {
        if (j + 1 < n_local) issue1(j + 1);
        const int s = j & 1, ph = (j >> 1) & 1;
        mbar_wait(z_full(s), ph);
        tcgen05_fence_after();
        const uint32_t as = base + L_A + s * 65536;
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          const uint32_t offa = (ks >> 2) * SUB_A + (ks & 3) * 32;
          const uint32_t offb = (ks >> 2) * 8192 + (ks & 3) * 32;
          umma_tf32(tmem + 256 + s * 64, umma_desc_k_sw128(as + offa), umma_desc_k_sw128(base + L_B2 + offb), idesc2,
                    ks > 0);
        }
        umma_commit(d2_full(s));
      }
    }
  } else {
    const int q = warp & 3;               // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;     // which 32 of the 64 channels this warp handles
    const int row = q * 32 + lane;
    for (int j = 0; j < n_local; ++j) {
      const int tile = blockIdx.x + j * gridDim.x;
      const int s = j & 1, ph = (j >> 1) & 1;
      const int b = tile / a.tiles_per_seq, t = (tile % a.tiles_per_seq) * TM + row;
      const bool valid = t < a.W;
      const int64_t grow = ((int64_t)b * a.W + t) * 64 + half * 32;
      uint8_t* as_g = gbase + L_A + s * 65536;
      const uint32_t trow = tmem + ((uint32_t)(q * 32) << 16);
      // ---- epilogue 1: gate ----
      mbar_wait(d1_full(s), ph);
      tcgen05_fence_after();
      uint32_t f[32], g[32];
      tmem_ld32(trow + s * 128 + half * 32, f);
      tmem_ld32(trow + s * 128 + 64 + half * 32, g);
      tmem_ld_wait();
      const bool live = valid && t >= a.zp;
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        float af = __uint_as_float(f[i]), ag = __uint_as_float(g[i]);
        if (a.bias_fg) {
          af += a.bias_fg[half * 32 + i];
          ag += a.bias_fg[64 + half * 32 + i];
        }
        const float zz = tanh_fast(af) * (0.5f * tanh_fast(0.5f * ag) + 0.5f);
        f[i] = __float_as_uint(live ? tf32_rna(zz) : 0.f);
      }
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const uint4 v = make_uint4(f[4 * c], f[4 * c + 1], f[4 * c + 2], f[4 * c + 3]);
        *reinterpret_cast<uint4*>(as_g + half * SUB_A + sw128_off(row, c)) = v;   // A operand of GEMM 2
        if (valid) *reinterpret_cast<uint4*>(a.z_out + grow + 4 * c) = v;
      }
      fence_proxy_async();
      tcgen05_fence_before();
      mbar_arrive(z_full(s));
      // ---- epilogue 2: projection + residual ----
      mbar_wait(d2_full(s), ph);
      tcgen05_fence_after();
      tmem_ld32(trow + 256 + s * 64 + half * 32, g);
      tmem_ld_wait();
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const float4 xv = *reinterpret_cast<const float4*>(as_g + (2 + half) * SUB_A + sw128_off(row, c));
        float4 o;
        o.x = __uint_as_float(g[4 * c]) + xv.x;
        o.y = __uint_as_float(g[4 * c + 1]) + xv.y;
        o.z = __uint_as_float(g[4 * c + 2]) + xv.z;
        o.w = __uint_as_float(g[4 * c + 3]) + xv.w;
        if (a.bias_p) {
          o.x += a.bias_p[half * 32 + 4 * c];
          o.y += a.bias_p[half * 32 + 4 * c + 1];
          o.z += a.bias_p[half * 32 + 4 * c + 2];
          o.w += a.bias_p[half * 32 + 4 * c + 3];
        }
        if (valid) *reinterpret_cast<float4*>(a.x_out + grow + 4 * c) = o;
      }
      tcgen05_fence_before();
      mbar_arrive(a_empty(s));
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem);
}

// ------------------------------------------------------------------------------------------
struct GemmTcArgs {
  float* Y;                // [num_seq * rows_out][ldy]
  int ldy;
  const float* bias;       // [N] or null
  int N;                   // valid output columns (<= BN)
  int relu, round_out;
  int rows_out;            // output rows per sequence (T)
  int a_row_off;           // first A row of a sequence that is used (W - T)
  int slabs, ksub;         // K = slabs * ksub * 32
  int tiles_per_seq, num_tiles;
};

template <int BN>
struct GemmCfg {
  static constexpr int STAGE = SUB_A + BN * 128;
  static constexpr int STAGES = (200 * 1024) / STAGE > 6 ? 6 : (200 * 1024) / STAGE;
  static constexpr int BAR = STAGES * STAGE;
  static constexpr int SMEM = BAR + 256 + 1024;
  static constexpr int ACC = 2;   // TMEM accumulator buffers (2 x BN <= 512 columns)
};

template <int BN>
__global__ void __launch_bounds__(L_THREADS, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b, const GemmTcArgs a) {
  using Cfg = GemmCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bar0 = base + Cfg::BAR;
  auto full = [&](int s) { return bar0 + 8 * s; };
  auto empty = [&](int s) { return bar0 + 64 + 8 * s; };
  auto acc_full = [&](int s) { return bar0 + 128 + 8 * s; };
  auto acc_empty = [&](int s) { return bar0 + 144 + 8 * s; };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(gbase + Cfg::BAR + 192);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < Cfg::STAGES; ++s) {
      mbar_init(full(s), 1);
      mbar_init(empty(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(acc_full(s), 1);
      mbar_init(acc_empty(s), 256);
    }
    fence_barrier_init();
    prefetch_tmap(&tm_a);
    prefetch_tmap(&tm_b);
  }
  if (warp == 1) tmem_alloc<512>(smem_u32((const void*)tmem_slot));
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int n_local = (a.num_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int ksteps = a.slabs * a.ksub;

  if (warp == 0) {
    if (lane == 0) {
      int it = 0;
      for (int j = 0; j < n_local; ++j) {
        const int tile = blockIdx.x + j * gridDim.x;
        const int b = tile / a.tiles_per_seq, t0 = (tile % a.tiles_per_seq) * TM;
        for (int sl = 0; sl < a.slabs; ++sl)
          for (int ks = 0; ks < a.ksub; ++ks, ++it) {
            const int s = it % Cfg::STAGES, ph = (it / Cfg::STAGES) & 1;
            mbar_wait(empty(s), ph ^ 1);
            const uint32_t st = base + s * Cfg::STAGE;
            mbar_arrive_expect_tx(full(s), Cfg::STAGE);
            tma_load_4d(st, &tm_a, full(s), ks * SUBK, a.a_row_off + t0, b, sl);
            tma_load_2d(st + SUB_A, &tm_b, full(s), (sl * a.ksub + ks) * SUBK, 0);
          }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_tf32(128, BN);
      int it = 0;
      for (int j = 0; j < n_local; ++j) {
        const int ab = j & 1, aph = (j >> 1) & 1;
        mbar_wait(acc_empty(ab), aph ^ 1);
        tcgen05_fence_after();
        for (int kk = 0; kk < ksteps; ++kk, ++it) {
          const int s = it % Cfg::STAGES, ph = (it / Cfg::STAGES) & 1;
          mbar_wait(full(s), ph);
          tcgen05_fence_after();
          const uint32_t st = base + s * Cfg::STAGE;
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4)
            umma_tf32(tmem + ab * BN, umma_desc_k_sw128(st + k4 * 32), umma_desc_k_sw128(st + SUB_A + k4 * 32), idesc,
                      (kk | k4) > 0);
          umma_commit(empty(s));
        }
        umma_commit(acc_full(ab));
      }
    }
  } else {
    const int q = warp & 3, half = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    constexpr int CH = BN / 64;   // 32-column chunks per warp
    for (int j = 0; j < n_local; ++j) {
      const int tile = blockIdx.x + j * gridDim.x;
      const int ab = j & 1, aph = (j >> 1) & 1;
      const int b = tile / a.tiles_per_seq, t = (tile % a.tiles_per_seq) * TM + row;
      const bool valid = t < a.rows_out;
      float* yrow = a.Y + ((int64_t)b * a.rows_out + t) * a.ldy;
      mbar_wait(acc_full(ab), aph);
      tcgen05_fence_after();
#pragma unroll 1
      for (int ch = 0; ch < CH; ++ch) {
        const int c0 = (half * CH + ch) * 32;
        uint32_t v[32];
        tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + ab * BN + c0, v);
        tmem_ld_wait();
        if (valid && c0 < a.N) {
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            float o[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              float x = __uint_as_float(v[4 * c + e]);
              if (a.bias) x += a.bias[c0 + 4 * c + e];
              if (a.relu) x = fmaxf(x, 0.f);
              if (a.round_out) x = tf32_rna(x);
              o[e] = x;
            }
            *reinterpret_cast<float4*>(yrow + c0 + 4 * c) = make_float4(o[0], o[1], o[2], o[3]);
          }
        }
      }
      tcgen05_fence_before();
      mbar_arrive(acc_empty(ab));
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem);
}

// ------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

// fp32 tensor [d3][d2][d1][d0] (d0 contiguous), box [1][1][box1][32], SWIZZLE_128B, zero fill out of range
int make_map_4d(CUtensorMap* m, const float* ptr, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t d3, uint64_t s1,
                uint64_t s2, uint64_t s3, uint32_t box1) {
  EncodeTiledFn enc = get_encode();
  WN_REQUIRE(enc, WN_ECUDA, "cuTensorMapEncodeTiled is unavailable");
  cuuint64_t dims[4] = {d0, d1, d2, d3};
  cuuint64_t strides[3] = {s1 * 4, s2 * 4, s3 * 4};
  cuuint32_t box[4] = {SUBK, box1, 1, 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)ptr, dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  WN_REQUIRE(r == CUDA_SUCCESS, WN_ECUDA, "cuTensorMapEncodeTiled(4d) failed: %d", (int)r);
  return WN_OK;
}

int make_map_2d(CUtensorMap* m, const float* ptr, uint64_t d0, uint64_t d1, uint64_t s1, uint32_t box1) {
  EncodeTiledFn enc = get_encode();
  WN_REQUIRE(enc, WN_ECUDA, "cuTensorMapEncodeTiled is unavailable");
  cuuint64_t dims[2] = {d0, d1};
  cuuint64_t strides[1] = {s1 * 4};
  cuuint32_t box[2] = {SUBK, box1};
  cuuint32_t es[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)ptr, dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  WN_REQUIRE(r == CUDA_SUCCESS, WN_ECUDA, "cuTensorMapEncodeTiled(2d) failed: %d", (int)r);
  return WN_OK;
}

template <int BN>
int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const GemmTcArgs& g, int sm_count, cudaStream_t s) {
  using Cfg = GemmCfg<BN>;
  static bool attr = false;
  if (!attr) {
    WN_CHECK_CUDA(cudaFuncSetAttribute(tc_gemm_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
    attr = true;
  }
  const int grid = g.num_tiles < sm_count ? g.num_tiles : sm_count;
  tc_gemm_kernel<BN><<<grid, L_THREADS, Cfg::SMEM, s>>>(ta, tb, g);
  WN_CHECK_LAUNCH();
  return WN_OK;
}

}  // namespace

// Y[(b, t)][0..N) = epi( sum_slab A_slab[b][a_row_off + t][:] . Wt[:, slab*K .. ]^T ),  Wt is [N][slabs*K] (TF32-rounded)
int tc_gemm(const wn_handle* h, const float* A, int K, int rows_in, int num_seq, int slabs, int64_t slab_stride,
            int a_row_off, int rows_out, const float* Wt, int N, const float* bias, int relu, int round_out, float* Y,
            int ldy, cudaStream_t s) {
  WN_REQUIRE(K % SUBK == 0 && N % 32 == 0 && N <= 256, WN_EINVAL, "tc_gemm: unsupported shape K=%d N=%d", K, N);
  CUtensorMap ta, tb;
  if (slabs == 1) slab_stride = (int64_t)rows_in * num_seq * K;
  WN_TRY(make_map_4d(&ta, A, K, rows_in, num_seq, slabs, K, (uint64_t)rows_in * K, slab_stride, TM));
  const int BN = N <= 64 ? 64 : (N <= 128 ? 128 : 256);
  WN_TRY(make_map_2d(&tb, Wt, (uint64_t)slabs * K, N, (uint64_t)slabs * K, BN));
  GemmTcArgs g;
  g.Y = Y;
  g.ldy = ldy;
  g.bias = bias;
  g.N = N;
  g.relu = relu;
  g.round_out = round_out;
  g.rows_out = rows_out;
  g.a_row_off = a_row_off;
  g.slabs = slabs;
  g.ksub = K / SUBK;
  g.tiles_per_seq = (rows_out + TM - 1) / TM;
  g.num_tiles = g.tiles_per_seq * num_seq;
  if (BN == 64) return launch_gemm<64>(ta, tb, g, h->sm_count, s);
  if (BN == 128) return launch_gemm<128>(ta, tb, g, h->sm_count, s);
  return launch_gemm<256>(ta, tb, g, h->sm_count, s);
}

bool tc_layer_supported(const wn_handle* h) {
  if (h->R != 64 || h->cfg.residual_filter_width != 2) return false;
  for (const ResLayer& l : h->layers)
    if (l.G != 64 || l.wf.b_off >= 0 || l.proj.b_off >= 0) return false;
  if (h->S % 32 != 0 || h->S > 256) return false;
  return get_encode() != nullptr;
}

bool tc_head_supported(const wn_handle* h) {
  for (const ConvParam& c : h->head)
    if (c.in_ch % 32 != 0 || c.out_ch % 32 != 0 || c.out_ch > 256) return false;
  return get_encode() != nullptr;
}

int tc_prepare_weights(wn_handle* h, const float* params, cudaStream_t s) {
  const Tape& t = h->tape;
  const int L = (int)h->layers.size();
  if (!h->tc_tab_uploaded) {
    std::vector<TcTabEntry> tab(L);
    for (int l = 0; l < L; ++l) {
      tab[l].wf = h->layers[l].wf.w_off;
      tab[l].wg = h->layers[l].wg.w_off;
      tab[l].wp = h->layers[l].proj.w_off;
      tab[l].ws = h->layers[l].skip.w_off;
    }
    WN_CHECK_CUDA(cudaMemcpyAsync(h->ws + t.tc_tab, tab.data(), sizeof(TcTabEntry) * L, cudaMemcpyHostToDevice, s));
    WN_CHECK_CUDA(cudaStreamSynchronize(s));   // tab is a stack temporary
    h->tc_tab_uploaded = true;
  }
  dim3 grid(16, L);
  tc_prep_kernel<<<grid, 256, 0, s>>>(params, (const TcTabEntry*)(h->ws + t.tc_tab), h->ws + t.tc_w1, h->ws + t.tc_w2,
                                      h->ws + t.tc_ws, L, h->R, 64, h->S, 2);
  WN_CHECK_LAUNCH();
  return WN_OK;
}

int tc_forward_residual(wn_handle* h, const float* params, cudaStream_t s) {
  const Tape& t = h->tape;
  const int L = (int)h->layers.size();
  const int R = 64, G = 64;
  WN_TRY(tc_prepare_weights(h, params, s));
  static bool attr = false;
  if (!attr) {
    WN_CHECK_CUDA(cudaFuncSetAttribute(tc_layer_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, L_SMEM + 1024));
    attr = true;
  }
  const int tiles_per_seq = (t.W + TM - 1) / TM;
  const int num_tiles = tiles_per_seq * t.B;
  const int grid = num_tiles < h->sm_count ? num_tiles : h->sm_count;
  for (int l = 0; l < L; ++l) {
    const ResLayer& ly = h->layers[l];
    WN_REQUIRE(ly.wf.b_off < 0 && ly.proj.b_off < 0, WN_EINVAL, "tc layer kernel: bias path goes through SIMT");
    CUtensorMap tx, tw1, tw2;
    WN_TRY(make_map_4d(&tx, h->ws + t.x[l], R, t.W, t.B, 1, R, (uint64_t)t.W * R, (uint64_t)t.P * R, TM));
    WN_TRY(make_map_2d(&tw1, h->ws + t.tc_w1 + (int64_t)l * 2 * G * 2 * R, 2 * R, 2 * G, 2 * R, 128));
    WN_TRY(make_map_2d(&tw2, h->ws + t.tc_w2 + (int64_t)l * R * G, G, R, G, 64));
    LayerArgs a;
    a.x_out = h->ws + t.x[l + 1];
    a.z_out = h->ws + t.z[l];
    a.bias_fg = nullptr;
    a.bias_p = nullptr;
    a.W = t.W;
    a.d = ly.dilation;
    a.zp = wn_zero_prefix(t.W, ly.dilation, 2);
    a.tiles_per_seq = tiles_per_seq;
    a.num_tiles = num_tiles;
    tc_layer_kernel<<<grid, L_THREADS, L_SMEM + 1024, s>>>(tx, tw1, tw2, a);
    WN_CHECK_LAUNCH();
  }
  // sum_skip = sum_l Ws_l z_l as one GEMM over K = L*G (z buffers are equally spaced slabs)
  const int64_t zstride = L > 1 ? t.z[1] - t.z[0] : 0;
  for (int l = 1; l < L; ++l)
    WN_REQUIRE(t.z[l] - t.z[l - 1] == zstride, WN_EINVAL, "z slabs are not equally spaced");
  return tc_gemm(h, h->ws + t.z[0], G, t.W, t.B, L, zstride, 0, t.W, h->ws + t.tc_ws, h->S, nullptr, 0, 0,
                 h->ws + t.skip, h->S, s);
}

// ReLU -> 1x1 conv per head layer (wavenet.py:587-590) on tensor cores.  Stored activations are
// post-ReLU (ReLU is idempotent, so the SIMT backward's a_relu / mask logic sees the same values).
int tc_forward_head(wn_handle* h, const float* params, int T, bool external, cudaStream_t s) {
  const Tape& t = h->tape;
  const int nh = (int)h->head.size();
  const int64_t rows = (int64_t)t.B * T;
  const int rows_in0 = external ? T : t.W, off0 = external ? 0 : t.W - T;
  {
    const int64_t n = rows * (h->S / 4);
    tc_relu_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(h->ws + t.skip, h->S, rows_in0, off0, T, rows);
    WN_CHECK_LAUNCH();
  }
  for (int i = 0; i < nh; ++i) {
    const ConvParam& cp = h->head[i];
    const int64_t n = (int64_t)cp.out_ch * cp.in_ch;
    tc_round_copy_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(params + cp.w_off, h->ws + t.tc_wh[i], n);
    WN_CHECK_LAUNCH();
    const float* A = i == 0 ? h->ws + t.skip : h->ws + t.hbuf[i - 1];
    const int rin = i == 0 ? rows_in0 : T, off = i == 0 ? off0 : 0;
    const bool last = i == nh - 1;
    WN_TRY(tc_gemm(h, A, cp.in_ch, rin, t.B, 1, 0, off, T, h->ws + t.tc_wh[i], cp.out_ch,
                   cp.b_off >= 0 ? params + cp.b_off : nullptr, last ? 0 : 1, last ? 0 : 1, h->ws + t.hbuf[i], cp.out_ch,
                   s));
  }
  return WN_OK;
}
