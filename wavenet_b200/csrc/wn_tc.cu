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
#include <cuda_fp16.h>
#include <stdlib.h>
#include <string.h>

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
                               float* __restrict__ w2, float* __restrict__ wsc, float* __restrict__ w1t,
                               float* __restrict__ wpt, float* __restrict__ wst, int L, int R, int G, int S, int k) {
  const int l = blockIdx.y;
  const TcTabEntry e = tab[l];
  const int n1 = 2 * G * k * R, n2 = R * G, n3 = S * G;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n1 + n2 + n3; i += gridDim.x * blockDim.x) {
    if (i < n1) {
      const int n = i / (k * R), kk = i % (k * R);
      const int tap = kk / R, c = kk % R;
      const float* src = n < G ? params + e.wf : params + e.wg;
      const float v = tf32_rna(src[((int64_t)(n % G) * R + c) * k + tap]);
      w1[(int64_t)l * n1 + i] = v;
      // backward data-gradient operand: w1t[c][slab*2G + n], slab 0 = current tap (k-1), slab 1 = past tap
      const int slab = (k - 1) - tap;
      w1t[(int64_t)l * n1 + (int64_t)c * (k * 2 * G) + slab * 2 * G + n] = v;
    } else if (i < n1 + n2) {
      const int j = i - n1;
      const float v = tf32_rna(params[e.wp + j]);
      w2[(int64_t)l * n2 + j] = v;
      const int r = j / G, g = j % G;
      wpt[(int64_t)l * n2 + (int64_t)g * R + r] = v;
    } else {
      const int j = i - n1 - n2;
      const int sidx = j / G, g = j % G;
      const float v = tf32_rna(params[e.ws + j]);
      wsc[(int64_t)sidx * (L * G) + (int64_t)l * G + g] = v;
      wst[(int64_t)l * n3 + (int64_t)g * S + sidx] = v;
    }
  }
}

// dst[i][o] = tf32(src[o][i])
__global__ void tc_transpose_round_kernel(const float* __restrict__ src, float* __restrict__ dst, int O, int I) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)O * I) return;
  const int o = (int)(idx / I), i = (int)(idx % I);
  dst[(int64_t)i * O + o] = tf32_rna(src[idx]);
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

// four consecutive sigmoid values of the tape: fp32, or fp16 when the fused layer kernel wrote them (the 10-bit
// mantissa equals what the TF32 MMAs keep of every other operand)
__device__ __forceinline__ float4 load_sg4(const float* sg, int64_t elem, int is_half) {
  if (is_half) {
    const uint2 raw = *reinterpret_cast<const uint2*>(reinterpret_cast<const __half*>(sg) + elem);
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&raw.x));
    const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&raw.y));
    return make_float4(a.x, a.y, b.x, b.y);
  }
  return *reinterpret_cast<const float4*>(sg + elem);
}

__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// sigmoid * (1 - tanh^2) rebuilt from z = tanh * sigmoid and the stored sigmoid.  A saturated gate stores sigmoid == 0
// (fp16 flushes below ~3e-8) while z may still be a tiny non-zero: the derivative is 0 there, not z*z/0.
__device__ __forceinline__ float gate_dtanh(float z, float sg) { return sg > 0.f ? sg - __fdividef(z * z, sg) : 0.f; }

// Coalesced store of a 32-row x 32-column fp32 block held row-per-lane (lane r owns row r): the block goes
// through a 2 KB per-warp XOR-swizzled staging buffer, 16 rows at a time, so that every STG.128 covers four
// complete 128-byte rows instead of 32 partial sectors.
__device__ __forceinline__ void store_block_coalesced(uint8_t* stg, const uint32_t (&v)[32], float* gblock, int64_t ld,
                                                      int rows_valid, int lane) {
#pragma unroll
  for (int hh = 0; hh < 2; ++hh) {
    if ((lane >> 4) == hh) {
      const int r = lane & 15;
#pragma unroll
      for (int c = 0; c < 8; ++c)
        *reinterpret_cast<uint4*>(stg + r * 128 + ((c ^ (r & 7)) << 4)) = make_uint4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
    }
    __syncwarp();
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      const int rr = jj * 4 + (lane >> 3);
      const int row = hh * 16 + rr;
      const uint4 val = *reinterpret_cast<const uint4*>(stg + rr * 128 + (((lane & 7) ^ (rr & 7)) << 4));
      if (row < rows_valid) *reinterpret_cast<uint4*>(gblock + (int64_t)row * ld + (lane & 7) * 4) = val;
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------
struct LayerArgs {
  float* x_out;            // [B][W][64]
  float* z_out;            // [B][W][64]
  float* sg_out;           // [B][W][64] fp16 sigmoid (backward rebuilds tanh = z / sigmoid), or null
  int W, d, zp, tiles_per_seq, num_tiles;
  int reverse;             // walk the tiles from the last to the first
};

constexpr int L_B1 = 0;                       // 4 sub-tiles [128 x 32]  (64 KB)
constexpr int L_B2 = 65536;                   // 2 sub-tiles [64 x 32]   (16 KB)
constexpr int L_A = 81920;                    // 2 stages x 4 sub-tiles  (128 KB); Z aliases sub-tiles 0,1
constexpr int L_BAR = L_A + 2 * 65536;        // barriers
constexpr int L_STG = L_BAR + 256;            // 8 epilogue warps x 2 KB store-transpose buffers
constexpr int L_SMEM = L_STG + 8 * 2048;
constexpr int L_THREADS = 64 + 256;           // producer warp, MMA warp, 8 epilogue warps

// Developer trace (make EXTRA=-DWN_LAYER_TRACE, tests/dev/trace_layer.py): clock64 stamps of CTA 0's pipeline events.
#ifdef WN_LAYER_TRACE
__device__ long long g_trace[64 * 32];
#define TR(j, e) do { if (blockIdx.x == 0 && (j) < 64) g_trace[(j) * 32 + (e)] = clock64(); } while (0)
#else
#define TR(j, e) do { } while (0)
#endif
__global__ void __launch_bounds__(L_THREADS, 1)
tc_layer_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w1,
                const __grid_constant__ CUtensorMap tm_w2, const __grid_constant__ CUtensorMap tm_z,
                const __grid_constant__ CUtensorMap tm_sg, const LayerArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bar0 = base + L_BAR;
  // barrier slots (8 bytes each)
  const uint32_t b_full = bar0;
  auto a_full = [&](int s) { return bar0 + 8 + 8 * s; };
  auto d1_full = [&](int s) { return bar0 + 40 + 8 * s; };
  auto z_full = [&](int s) { return bar0 + 56 + 8 * s; };
  auto d2_full = [&](int s) { return bar0 + 72 + 8 * s; };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(gbase + L_BAR + 128);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(b_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(a_full(s), 1);
      mbar_init(d1_full(s), 1);
      mbar_init(z_full(s), 256);
      mbar_init(d2_full(s), 1);
    }
    fence_barrier_init();
    prefetch_tmap(&tm_x);
    prefetch_tmap(&tm_w1);
    prefetch_tmap(&tm_w2);
    prefetch_tmap(&tm_z);
    prefetch_tmap(&tm_sg);
  }
  if (warp == 1) tmem_alloc<512>(smem_u32((const void*)tmem_slot));
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  pdl_launch_dependents();
  pdl_wait();
  const uint32_t tmem = *tmem_slot;
  const int n_local = (a.num_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  // Serpentine order: odd layers walk the tiles from the end, so a layer starts with the rows the previous layer wrote
  // last (still in L2) instead of the rows it wrote first (long evicted).
  auto tile_of = [&](int j) {
    const int i = (int)blockIdx.x + j * (int)gridDim.x;
    return a.reverse ? a.num_tiles - 1 - i : i;
  };

  if (warp == 0) {
    if (lane == 0) {
      // The producer thread also writes z and sigmoid back: both tiles sit in the stage in the 128B-swizzled layout
      // TMA understands (z = A operand of GEMM 2 in sub-tiles 0,1; sigmoid replaces x(t) in sub-tiles 2,3 once the
      // residual is in registers), so each is one pair of bulk tensor stores instead of a register -> smem ->
      // register -> global transpose in the epilogue warps.
      mbar_arrive_expect_tx(b_full, 81920);
      for (int j = 0; j < 4; ++j) tma_load_2d(base + L_B1 + j * SUB_A, &tm_w1, b_full, j * SUBK, 0);
      for (int j = 0; j < 2; ++j) tma_load_2d(base + L_B2 + j * 8192, &tm_w2, b_full, j * SUBK, 0);
      for (int j = 0; j <= n_local; ++j) {
        const int s = j & 1;
        const uint32_t as = base + L_A + s * 65536;
        if (j >= 2) {
          // tile j-2 used this stage.  Its last reader is GEMM 2 (z in sub-tiles 0,1); the residual was taken into
          // registers before z_full.  Also wait until the bulk stores of z / sigmoid have finished READING the stage.
          mbar_wait(d2_full(s), ((j - 2) >> 1) & 1);
          TR(j, 0);
          bulk_wait_group_read0();
          TR(j, 1);
        }
        if (j < n_local) {
          const int tile = tile_of(j);
          const int b = tile / a.tiles_per_seq, t0 = (tile % a.tiles_per_seq) * TM;
          mbar_arrive_expect_tx(a_full(s), 65536);
          tma_load_4d(as + 0 * SUB_A, &tm_x, a_full(s), 0, t0 - a.d, b, 0);
          tma_load_4d(as + 1 * SUB_A, &tm_x, a_full(s), SUBK, t0 - a.d, b, 0);
          tma_load_4d(as + 2 * SUB_A, &tm_x, a_full(s), 0, t0, b, 0);
          tma_load_4d(as + 3 * SUB_A, &tm_x, a_full(s), SUBK, t0, b, 0);
          TR(j, 2);
          if (j + 2 < n_local) {
            // the load of tile j+2 cannot be issued before this stage drains: pull its rows into L2 now so that it
            // costs an L2 hit instead of a DRAM round trip on the stage's critical path
            const int tp = tile_of(j + 2);
            const int bp = tp / a.tiles_per_seq, tp0 = (tp % a.tiles_per_seq) * TM;
            tma_prefetch_4d(&tm_x, 0, tp0, bp, 0);
            tma_prefetch_4d(&tm_x, SUBK, tp0, bp, 0);
          }
        }
        if (j >= 1) {
          const int jj = j - 1, s1 = jj & 1, tile = tile_of(jj);
          const int b = tile / a.tiles_per_seq, t0 = (tile % a.tiles_per_seq) * TM;
          const uint32_t zs = base + L_A + s1 * 65536;
          mbar_wait(z_full(s1), (jj >> 1) & 1);
          TR(j, 3);
          tma_store_4d(&tm_z, zs + 0 * SUB_A, 0, t0, b, 0);
          tma_store_4d(&tm_z, zs + 1 * SUB_A, SUBK, t0, b, 0);
          if (a.sg_out) tma_store_4d(&tm_sg, zs + 2 * SUB_A, 0, t0, b, 0);   // fp16 [128 x 64]
          bulk_commit_group();
        }
      }
      bulk_wait_group0();
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc1 = umma_idesc_tf32(128, 128);
      mbar_wait(b_full, 0);
      // GEMM 1 stream only: GEMM 2 of a tile is issued by an epilogue thread the moment z is complete (waiting here
      // for the next tile's load before GEMM 2 of the current tile would stall the epilogue warps on d2_full).
      const uint64_t db1 = umma_desc_k_sw128(base + L_B1);
      for (int j1 = 0; j1 < n_local; ++j1) {
        const int s = j1 & 1;
        // GEMM 1 of tile j1 overwrites the accumulator that epilogue 1 of tile j1-2 read: wait for its z_full
        if (j1 >= 2) mbar_wait(z_full(s), ((j1 - 2) >> 1) & 1);
        mbar_wait(a_full(s), (j1 >> 1) & 1);
        TR(j1, 4);
        tcgen05_fence_after();
        const uint64_t da = umma_desc_k_sw128(base + L_A + s * 65536);
#pragma unroll
        for (int ks = 0; ks < 16; ++ks) {
          const uint32_t off = (ks >> 2) * SUB_A + (ks & 3) * 32;
          umma_tf32(tmem + s * 128, da + (off >> 4), db1 + (off >> 4), idesc1, ks > 0);
        }
        umma_commit(d1_full(s));
        TR(j1, 5);
      }
    }
  } else {
    const int q = warp & 3;               // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;     // which 32 of the 64 channels this warp handles
    const int row = q * 32 + lane;
    uint8_t* stg = gbase + L_STG + (warp - 2) * 2048;
    if (threadIdx.x == 64) mbar_wait(b_full, 0);   // W2 resident before this thread issues the first GEMM 2
    for (int j = 0; j < n_local; ++j) {
      const int tile = tile_of(j);
      const int s = j & 1, ph = (j >> 1) & 1;
      const int b = tile / a.tiles_per_seq, t = (tile % a.tiles_per_seq) * TM + row;
      const bool valid = t < a.W;
      uint8_t* as_g = gbase + L_A + s * 65536;
      const uint32_t trow = tmem + ((uint32_t)(q * 32) << 16);
      // ---- epilogue 1: gate ----
      mbar_wait(d1_full(s), ph);
      if (threadIdx.x == 64) TR(j, 8);
      tcgen05_fence_after();
      uint32_t f[32], g[32];
      tmem_ld32(trow + s * 128 + half * 32, f);
      tmem_ld32(trow + s * 128 + 64 + half * 32, g);
      // residual x(t) of this thread's row: into registers now, so that the stage is free as soon as GEMM 2 is done
      float4 xr[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) xr[c] = *reinterpret_cast<const float4*>(as_g + (2 + half) * SUB_A + sw128_off(row, c));
      tmem_ld_wait();
      // the fp16 sigmoid tile below covers sub-tile 2 for BOTH channel halves: the partner warp (same rows, other half)
      // must have taken its residual before either of us overwrites it
      asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");
      const bool live = valid && t >= a.zp;
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        float af = __uint_as_float(f[i]), ag = __uint_as_float(g[i]);
        const float tf = tanh_fast(af), sg = 0.5f * tanh_fast(0.5f * ag) + 0.5f;
        g[i] = __float_as_uint(live ? sg : 0.5f);          // masked rows look like tanh(0) | sigmoid(0)
        f[i] = __float_as_uint(tf32_rna((live ? tf : 0.f) * sg));
      }
#pragma unroll
      for (int c = 0; c < 8; ++c)   // z: A operand of GEMM 2 and source of the z bulk store
        *reinterpret_cast<uint4*>(as_g + half * SUB_A + sw128_off(row, c)) =
            make_uint4(f[4 * c], f[4 * c + 1], f[4 * c + 2], f[4 * c + 3]);
      if (a.sg_out) {
        // sigmoid as fp16, one [128 x 64] tile of 128-byte rows over x(t) (already in xr): source of its bulk store
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint4 pk;
          __half2 h0 = __floats2half2_rn(__uint_as_float(g[8 * c]), __uint_as_float(g[8 * c + 1]));
          __half2 h1 = __floats2half2_rn(__uint_as_float(g[8 * c + 2]), __uint_as_float(g[8 * c + 3]));
          __half2 h2 = __floats2half2_rn(__uint_as_float(g[8 * c + 4]), __uint_as_float(g[8 * c + 5]));
          __half2 h3 = __floats2half2_rn(__uint_as_float(g[8 * c + 6]), __uint_as_float(g[8 * c + 7]));
          pk.x = *reinterpret_cast<uint32_t*>(&h0), pk.y = *reinterpret_cast<uint32_t*>(&h1);
          pk.z = *reinterpret_cast<uint32_t*>(&h2), pk.w = *reinterpret_cast<uint32_t*>(&h3);
          *reinterpret_cast<uint4*>(as_g + 2 * SUB_A + sw128_off(row, half * 4 + c)) = pk;
        }
      }
      fence_proxy_async();
      tcgen05_fence_before();
      mbar_arrive(z_full(s));
      if (threadIdx.x == 64) TR(j, 9);
      if (threadIdx.x == 64) {
        // one epilogue thread issues GEMM 2 as soon as every row of z is in shared memory
        constexpr uint32_t idesc2 = umma_idesc_tf32(128, 64);
        mbar_wait(z_full(s), ph);
        tcgen05_fence_after();
        const uint64_t da = umma_desc_k_sw128(base + L_A + s * 65536), db2 = umma_desc_k_sw128(base + L_B2);
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          const uint32_t offa = (ks >> 2) * SUB_A + (ks & 3) * 32;
          const uint32_t offb = (ks >> 2) * 8192 + (ks & 3) * 32;
          umma_tf32(tmem + 256 + s * 64, da + (offa >> 4), db2 + (offb >> 4), idesc2, ks > 0);
        }
        umma_commit(d2_full(s));
        TR(j, 6);
      }
      __syncwarp();
      // ---- epilogue 2: projection + residual ----
      mbar_wait(d2_full(s), ph);
      if (threadIdx.x == 64) TR(j, 10);
      tcgen05_fence_after();
      tmem_ld32(trow + 256 + s * 64 + half * 32, g);
      tmem_ld_wait();
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        g[4 * c] = __float_as_uint(__uint_as_float(g[4 * c]) + xr[c].x);
        g[4 * c + 1] = __float_as_uint(__uint_as_float(g[4 * c + 1]) + xr[c].y);
        g[4 * c + 2] = __float_as_uint(__uint_as_float(g[4 * c + 2]) + xr[c].z);
        g[4 * c + 3] = __float_as_uint(__uint_as_float(g[4 * c + 3]) + xr[c].w);
      }
      tcgen05_fence_before();
      {
        const int t_w0 = (tile % a.tiles_per_seq) * TM + q * 32;           // first row of the warp's block
        store_block_coalesced(stg, g, a.x_out + ((int64_t)b * a.W + t_w0) * 64 + half * 32, 64, a.W - t_w0, lane);
      }
      if (threadIdx.x == 64) TR(j, 11);
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem);
}

// ------------------------------------------------------------------------------------------
// Y[(b,t)][n] = epi( sum_slab  A[slab_idx][b][t + slab_row_off][:] . Wt[n][slab*K ...] )
constexpr int MAX_SLABS = 32;
struct GemmTcArgs {
  float* Y;                // [num_seq * rows_out][ldy]
  int ldy;
  const float* bias;       // [N] or null
  int N;                   // valid output columns (<= BN)
  int relu, round_out, accumulate;
  const float* Rsd;        // residual added to the result (same rows as Y), or null
  int ldr;
  const float* mask;       // result zeroed where mask[(b, mask_row_off + t)][n] <= 0, or null
  int ldm, mask_rows_in, mask_row_off;
  int rows_out;            // output rows per sequence
  int nslab, ksub;         // K = nslab * ksub * 32
  int slab_row_off[MAX_SLABS];
  int slab_idx[MAX_SLABS];
  int tiles_per_seq, num_tiles;
  int ngroups;             // BN-column groups of the output (N > BN): tile = row_tile * ngroups + group, so CTAs that run
                           // at the same time share the A rows through L2 (A is read from HBM once, not once per group)
  int y_slab_cols;         // >0: column block c goes to Y + (c / y_slab_cols) * y_slab_stride, column c % y_slab_cols
  int64_t y_slab_stride;
  // gate-backward epilogue (N == G): result is dz; writes da_f | da_g into dafg[row][0..2G) instead of Y
  const float* gate_sg;    // [rows][gate_sg_ld] sigmoid, or null
  const float* gate_z;     // [rows][G] z = tanh * sigmoid
  float* gate_dafg;
  int gate_zp, gate_sg_ld, gate_sg_half;   // gate_sg_half: the sigmoid tape is fp16 (ld in elements)
  int zero_rows_below;     // output rows with t < this are forced to 0 (quirk Q1 zero prefix)
  int reverse;             // walk the tiles from the last to the first (serpentine hand-off through L2)
  float* colsum_out;       // MODE 0: += column sums of the stored result (bias gradient of the producing conv), or null
};

template <int BN>
struct GemmCfg {
  static constexpr int STAGE = SUB_A + BN * 128;
  static constexpr int STAGES = (200 * 1024) / STAGE > 6 ? 6 : (200 * 1024) / STAGE;
  static constexpr int BAR = STAGES * STAGE;
  static constexpr int STG = BAR + 256;            // 8 epilogue warps x 4 KB transpose buffers
  static constexpr int SMEM = STG + 8 * 4096 + 1024 + 1024;   // + BN floats of column sums (MODE 3)
};

// MODE selects the epilogue at compile time (runtime-predicated feature code costs issue slots in the 8 epilogue warps,
// which are the critical resource of these HBM-bound kernels): 0 = generic (runtime flags), 1 = dz + gate derivative,
// 2 = Y = acc + Rsd only (the per-layer dx GEMM), 3 = generic + column sums of the result (bias gradient).
template <int BN, int MODE>
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
  if constexpr (MODE == 3) {
    if (threadIdx.x < BN) reinterpret_cast<float*>(gbase + Cfg::STG + 8 * 4096)[threadIdx.x] = 0.f;
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  pdl_launch_dependents();
  pdl_wait();
  const uint32_t tmem = *tmem_slot;
  const int n_local = (a.num_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int ksteps = a.nslab * a.ksub;

  if (warp == 0) {
    if (lane == 0) {
      int it = 0;
      for (int j = 0; j < n_local; ++j) {
        const int tile = a.reverse ? a.num_tiles - 1 - ((int)blockIdx.x + j * (int)gridDim.x) : (int)blockIdx.x + j * (int)gridDim.x;
        const int grp = tile % a.ngroups, rt = tile / a.ngroups;
        const int b = rt / a.tiles_per_seq, t0 = (rt % a.tiles_per_seq) * TM;
        for (int sl = 0; sl < a.nslab; ++sl)
          for (int ks = 0; ks < a.ksub; ++ks, ++it) {
            const int s = it % Cfg::STAGES, ph = (it / Cfg::STAGES) & 1;
            mbar_wait(empty(s), ph ^ 1);
            const uint32_t st = base + s * Cfg::STAGE;
            mbar_arrive_expect_tx(full(s), Cfg::STAGE);
            tma_load_4d(st, &tm_a, full(s), ks * SUBK, a.slab_row_off[sl] + t0, b, a.slab_idx[sl]);
            tma_load_2d(st + SUB_A, &tm_b, full(s), (sl * a.ksub + ks) * SUBK, grp * BN);
            if (BN == 256) TR(it, 0);
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
          if (BN == 256) TR(it, 1);
          tcgen05_fence_after();
          const uint32_t st = base + s * Cfg::STAGE;
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4)
            umma_tf32(tmem + ab * BN, umma_desc_k_sw128(st + k4 * 32), umma_desc_k_sw128(st + SUB_A + k4 * 32), idesc,
                      (kk | k4) > 0);
          umma_commit(empty(s));
        }
        umma_commit(acc_full(ab));
        if (BN == 256) TR(j, 2);
      }
    }
  } else {
    const int q = warp & 3, half = (warp - 2) >> 2;
    constexpr int CH = BN / 64;   // 32-column chunks per warp
    uint8_t* stg = gbase + Cfg::STG + (warp - 2) * 4096;   // [32 rows][128 B], 16-byte chunks XOR-swizzled by row
    const int cc4 = (lane & 7) * 4;                         // column (within the 32-col block) this lane owns when coalesced
    // MODE 3 (= MODE 0 + colsum_out): column sums of the stored result accumulate in a BN-float shared-memory array
    // (per-warp partial sums folded with shuffles, shared-memory atomics), flushed to global memory once per CTA
    float* cs_smem = reinterpret_cast<float*>(gbase + Cfg::STG + 8 * 4096);
    for (int j = 0; j < n_local; ++j) {
      const int tile = a.reverse ? a.num_tiles - 1 - ((int)blockIdx.x + j * (int)gridDim.x) : (int)blockIdx.x + j * (int)gridDim.x;
      const int ab = j & 1, aph = (j >> 1) & 1;
      const int grp = (MODE == 0 || MODE == 3) ? tile % a.ngroups : 0, rt = (MODE == 0 || MODE == 3) ? tile / a.ngroups : tile;
      const int b = rt / a.tiles_per_seq, t0 = (rt % a.tiles_per_seq) * TM + q * 32;
      mbar_wait(acc_full(ab), aph);
      if (BN == 256 && threadIdx.x == 64) TR(j, 3);
      tcgen05_fence_after();
#pragma unroll 1
      for (int ch = 0; ch < CH; ++ch) {
        const int ct = (half * CH + ch) * 32;          // column inside the TMEM tile
        const int c0 = grp * BN + ct;                  // output column
        uint32_t v[32];
        tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + ab * BN + ct, v);
        tmem_ld_wait();
        if (c0 >= a.N) continue;
        // row-per-lane -> smem -> 4 full 128-byte rows per instruction (coalesced global traffic)
#pragma unroll
        for (int c = 0; c < 8; ++c)
          *reinterpret_cast<uint4*>(stg + lane * 128 + ((c ^ (lane & 7)) << 4)) =
              make_uint4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
        __syncwarp();
        const int col = c0 + cc4;
        // phase 1: issue every global load of this 32x32 block (memory-level parallelism), phase 2: math + stores.
        // Mode-specific branches keep only the needed register arrays alive (no spills).
        const int rsub = lane >> 3;
        if constexpr (MODE == 2) {
          float4 r4[8];
#pragma unroll
          for (int jj = 0; jj < 8; ++jj) {
            const int t = min(t0 + jj * 4 + rsub, a.rows_out - 1);
            r4[jj] = *reinterpret_cast<const float4*>(a.Rsd + ((int64_t)b * a.rows_out + t) * a.ldr + col);
          }
#pragma unroll
          for (int jj = 0; jj < 8; ++jj) {
            const int rr = jj * 4 + rsub;
            const int t = t0 + rr;
            if (t >= a.rows_out) continue;
            float4 o = *reinterpret_cast<const float4*>(stg + rr * 128 + (((lane & 7) ^ (rr & 7)) << 4));
            o.x += r4[jj].x, o.y += r4[jj].y, o.z += r4[jj].z, o.w += r4[jj].w;
            *reinterpret_cast<float4*>(a.Y + ((int64_t)b * a.rows_out + t) * a.ldy + col) = o;
          }
        } else if constexpr (MODE == 1) {
          float4 r4[8], z4[8], sg4[8];
#pragma unroll
          for (int jj = 0; jj < 8; ++jj) {
            const int t = min(t0 + jj * 4 + rsub, a.rows_out - 1);
            const int64_t orow = (int64_t)b * a.rows_out + t;
            r4[jj] = *reinterpret_cast<const float4*>(a.Rsd + orow * a.ldr + col);
            z4[jj] = *reinterpret_cast<const float4*>(a.gate_z + orow * a.N + col);
            sg4[jj] = load_sg4(a.gate_sg, orow * a.gate_sg_ld + col, a.gate_sg_half);
          }
#pragma unroll
          for (int jj = 0; jj < 8; ++jj) {
            const int rr = jj * 4 + rsub;
            const int t = t0 + rr;
            if (t >= a.rows_out) continue;
            const int64_t orow = (int64_t)b * a.rows_out + t;
            float4 o = *reinterpret_cast<const float4*>(stg + rr * 128 + (((lane & 7) ^ (rr & 7)) << 4));
            o.x += r4[jj].x, o.y += r4[jj].y, o.z += r4[jj].z, o.w += r4[jj].w;
            // with z = tanh*sg:  da_f = dz*sg*(1-tanh^2) = dz*(sg - z*z/sg),  da_g = dz*tanh*sg*(1-sg) = dz*z*(1-sg);
            // rows inside the zero prefix get none (Q1)
            const float4 z = z4[jj], sg = sg4[jj];
            const float live = t >= a.gate_zp ? 1.f : 0.f;
            float4 df, dg;
            df.x = live * o.x * gate_dtanh(z.x, sg.x), dg.x = live * o.x * z.x * (1.f - sg.x);
            df.y = live * o.y * gate_dtanh(z.y, sg.y), dg.y = live * o.y * z.y * (1.f - sg.y);
            df.z = live * o.z * gate_dtanh(z.z, sg.z), dg.z = live * o.z * z.z * (1.f - sg.z);
            df.w = live * o.w * gate_dtanh(z.w, sg.w), dg.w = live * o.w * z.w * (1.f - sg.w);
            float* drow = a.gate_dafg + orow * (2 * a.N);
            *reinterpret_cast<float4*>(drow + col) = df;
            *reinterpret_cast<float4*>(drow + a.N + col) = dg;
          }
        } else {
          float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
          if (a.bias) bb = *reinterpret_cast<const float4*>(a.bias + col);
          float* ybase = a.Y;
          int ycol = col;
          if (a.y_slab_cols > 0) {
            ybase += (int64_t)(c0 / a.y_slab_cols) * a.y_slab_stride;
            ycol = col % a.y_slab_cols;
          }
          float4 r4[8], x4[8];   // x4: ReLU mask source or previous Y (never both)
          float4 csum = make_float4(0.f, 0.f, 0.f, 0.f);
          const bool has_aux = a.mask != nullptr || a.accumulate;
#pragma unroll
          for (int jj = 0; jj < 8; ++jj) {
            const int t = min(t0 + jj * 4 + rsub, a.rows_out - 1);
            const int64_t orow = (int64_t)b * a.rows_out + t;
            if (a.Rsd) r4[jj] = *reinterpret_cast<const float4*>(a.Rsd + orow * a.ldr + col);
            if (has_aux) {
              const float* ap = a.mask ? a.mask + ((int64_t)b * a.mask_rows_in + a.mask_row_off + t) * a.ldm + col
                                       : ybase + orow * a.ldy + ycol;
              x4[jj] = *reinterpret_cast<const float4*>(ap);
            }
          }
#pragma unroll
          for (int jj = 0; jj < 8; ++jj) {
            const int rr = jj * 4 + rsub;
            const int t = t0 + rr;
            if (t >= a.rows_out) continue;
            const int64_t orow = (int64_t)b * a.rows_out + t;
            float4 o = *reinterpret_cast<const float4*>(stg + rr * 128 + (((lane & 7) ^ (rr & 7)) << 4));
            o.x += bb.x, o.y += bb.y, o.z += bb.z, o.w += bb.w;
            if (a.relu) o.x = fmaxf(o.x, 0.f), o.y = fmaxf(o.y, 0.f), o.z = fmaxf(o.z, 0.f), o.w = fmaxf(o.w, 0.f);
            if (t < a.zero_rows_below) o = make_float4(0.f, 0.f, 0.f, 0.f);
            if (a.Rsd) o.x += r4[jj].x, o.y += r4[jj].y, o.z += r4[jj].z, o.w += r4[jj].w;
            if (a.mask) {
              o.x = x4[jj].x > 0.f ? o.x : 0.f;
              o.y = x4[jj].y > 0.f ? o.y : 0.f;
              o.z = x4[jj].z > 0.f ? o.z : 0.f;
              o.w = x4[jj].w > 0.f ? o.w : 0.f;
            } else if (a.accumulate) {
              o.x += x4[jj].x, o.y += x4[jj].y, o.z += x4[jj].z, o.w += x4[jj].w;
            }
            if (a.round_out) o.x = tf32_rna(o.x), o.y = tf32_rna(o.y), o.z = tf32_rna(o.z), o.w = tf32_rna(o.w);
            *reinterpret_cast<float4*>(ybase + orow * a.ldy + ycol) = o;
            if constexpr (MODE == 3) csum.x += o.x, csum.y += o.y, csum.z += o.z, csum.w += o.w;
          }
          if constexpr (MODE == 3) {
            // lanes l, l+8, l+16, l+24 own the same four columns (different rows)
#pragma unroll
            for (int o = 8; o < 32; o <<= 1) {
              csum.x += __shfl_xor_sync(0xffffffffu, csum.x, o);
              csum.y += __shfl_xor_sync(0xffffffffu, csum.y, o);
              csum.z += __shfl_xor_sync(0xffffffffu, csum.z, o);
              csum.w += __shfl_xor_sync(0xffffffffu, csum.w, o);
            }
            if (lane < 8) {
              float* c = cs_smem + ct + cc4;
              atomicAdd(c, csum.x), atomicAdd(c + 1, csum.y), atomicAdd(c + 2, csum.z), atomicAdd(c + 3, csum.w);
            }
          }
        }
        __syncwarp();
      }
      tcgen05_fence_before();
      mbar_arrive(acc_empty(ab));
      if (BN == 256 && threadIdx.x == 64) TR(j, 4);
    }
    if constexpr (MODE == 3) {
      asm volatile("bar.sync 1, 256;" ::: "memory");   // the eight epilogue warps
      const int c = threadIdx.x - 64;
      if (c < BN && c < a.N) atomicAdd(a.colsum_out + c, cs_smem[c]);
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem);
}

// ------------------------------------------------------------------------------------------
// A-resident slab GEMM for wide outputs: Y[(b,t)][0..Ntot) = A[b][t + row_off][0..K) . Wt[Ntot][K]^T with K <= 256.
// The 128 x K A tile is loaded ONCE per 128-row tile and stays in shared memory while the kernel walks the
// 256-column output groups (B streamed through a 2-stage ring from L2).  Used for dzs = dskip . [Ws_0 .. Ws_L-1]:
// dskip is read from HBM once instead of once per group of four layers.
struct GemmAresArgs {
  float* Y;
  int ldy;                 // row stride inside a slab
  int y_slab_cols;         // columns per output slab (G)
  int64_t y_slab_stride;
  int Ntot, ngroups;       // output columns, ceil(Ntot / 256)
  int ksub;                // K / 32  (<= 8)
  int rows_out, a_row_off;
  int tiles_per_seq, num_tiles;
};

constexpr int AR_A = 0;                        // up to 8 sub-tiles [128 x 32]  (128 KB)
constexpr int AR_B = 131072;                   // 2 stages x [256 x 32]          (64 KB)
constexpr int AR_BAR = AR_B + 2 * 32768;
constexpr int AR_STG = AR_BAR + 256;           // 8 x 4 KB transpose buffers
constexpr int AR_SMEM = AR_STG + 8 * 4096 + 1024;

__global__ void __launch_bounds__(L_THREADS, 1)
tc_gemm_ares_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                    const GemmAresArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bar0 = base + AR_BAR;
  const uint32_t a_full = bar0, a_empty = bar0 + 8;
  auto b_full = [&](int s) { return bar0 + 16 + 8 * s; };
  auto b_empty = [&](int s) { return bar0 + 32 + 8 * s; };
  auto acc_full = [&](int s) { return bar0 + 48 + 8 * s; };
  auto acc_empty = [&](int s) { return bar0 + 64 + 8 * s; };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(gbase + AR_BAR + 128);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(a_full, 1);
    mbar_init(a_empty, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(b_full(s), 1);
      mbar_init(b_empty(s), 1);
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

  if (warp == 0) {
    if (lane == 0) {
      int it = 0;
      for (int j = 0; j < n_local; ++j) {
        const int tile = blockIdx.x + j * gridDim.x;
        const int b = tile / a.tiles_per_seq, t0 = (tile % a.tiles_per_seq) * TM;
        mbar_wait(a_empty, (j & 1) ^ 1);
        mbar_arrive_expect_tx(a_full, a.ksub * SUB_A);
        for (int ks = 0; ks < a.ksub; ++ks) tma_load_4d(base + AR_A + ks * SUB_A, &tm_a, a_full, ks * SUBK, a.a_row_off + t0, b, 0);
        for (int ng = 0; ng < a.ngroups; ++ng)
          for (int ks = 0; ks < a.ksub; ++ks, ++it) {
            const int s = it & 1, ph = (it >> 1) & 1;
            mbar_wait(b_empty(s), ph ^ 1);
            mbar_arrive_expect_tx(b_full(s), 32768);
            tma_load_2d(base + AR_B + s * 32768, &tm_b, b_full(s), ks * SUBK, ng * 256);
          }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_tf32(128, 256);
      int it = 0, ia = 0;
      for (int j = 0; j < n_local; ++j) {
        mbar_wait(a_full, j & 1);
        tcgen05_fence_after();
        for (int ng = 0; ng < a.ngroups; ++ng, ++ia) {
          const int ab = ia & 1, aph = (ia >> 1) & 1;
          mbar_wait(acc_empty(ab), aph ^ 1);
          tcgen05_fence_after();
          for (int ks = 0; ks < a.ksub; ++ks, ++it) {
            const int s = it & 1, ph = (it >> 1) & 1;
            mbar_wait(b_full(s), ph);
            tcgen05_fence_after();
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4)
              umma_tf32(tmem + ab * 256, umma_desc_k_sw128(base + AR_A + ks * SUB_A + k4 * 32),
                        umma_desc_k_sw128(base + AR_B + s * 32768 + k4 * 32), idesc, (ks | k4) > 0);
            umma_commit(b_empty(s));
          }
          umma_commit(acc_full(ab));
        }
        umma_commit(a_empty);   // every MMA that reads this A tile has completed when this fires
      }
    }
  } else {
    const int q = warp & 3, half = (warp - 2) >> 2;
    uint8_t* stg = gbase + AR_STG + (warp - 2) * 4096;
    const int cc4 = (lane & 7) * 4, rsub = lane >> 3;
    int ia = 0;
    for (int j = 0; j < n_local; ++j) {
      const int tile = blockIdx.x + j * gridDim.x;
      const int b = tile / a.tiles_per_seq, t0 = (tile % a.tiles_per_seq) * TM + q * 32;
      for (int ng = 0; ng < a.ngroups; ++ng, ++ia) {
        const int ab = ia & 1, aph = (ia >> 1) & 1;
        mbar_wait(acc_full(ab), aph);
        tcgen05_fence_after();
#pragma unroll 1
        for (int ch = 0; ch < 4; ++ch) {
          const int c0 = (half * 4 + ch) * 32;
          uint32_t v[32];
          tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + ab * 256 + c0, v);
          tmem_ld_wait();
          const int gc = ng * 256 + c0;          // first global output column of this block
          if (gc >= a.Ntot) continue;
#pragma unroll
          for (int c = 0; c < 8; ++c)
            *reinterpret_cast<uint4*>(stg + lane * 128 + ((c ^ (lane & 7)) << 4)) =
                make_uint4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
          __syncwarp();
          float* ybase = a.Y + (int64_t)((gc + cc4) / a.y_slab_cols) * a.y_slab_stride + ((gc + cc4) % a.y_slab_cols);
#pragma unroll
          for (int jj = 0; jj < 8; ++jj) {
            const int rr = jj * 4 + rsub;
            const int t = t0 + rr;
            if (t >= a.rows_out) continue;
            const uint4 o = *reinterpret_cast<const uint4*>(stg + rr * 128 + (((lane & 7) ^ (rr & 7)) << 4));
            *reinterpret_cast<uint4*>(ybase + ((int64_t)b * a.rows_out + t) * a.ldy) = o;
          }
          __syncwarp();
        }
        tcgen05_fence_before();
        mbar_arrive(acc_empty(ab));
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem);
}

// ------------------------------------------------------------------------------------------
// Weight gradient: D[128 x NB] = sum over positions  dY[p][a_c0 + m] * X_slab[p + off_slab][c]
// The reduction dimension (positions) is the row index of both operands in memory, so both are
// MN-major UMMA operands: TMA deposits [32 positions x 32 channels] sub-tiles (128-byte rows,
// SWIZZLE_128B_ATOM_32B); one K=8 MMA step consumes two 4-row swizzle groups (SBO 512 B), sub-tiles
// are 4096 B apart (LBO).
struct WgradTcArgs {
  int rows_it, num_seq;    // iteration rows per sequence
  int a_row_off, a_c0;
  int nb_slab, nb_sub;     // B slabs (taps or layers) and 32-channel sub-tiles per slab
  int b_row_off[4];
  // Slab groups: CTA c works on group c % ngroups over the chunk range c / ngroups, so the CTAs that run at the same
  // time read the same dY rows for different B slabs (dY comes from HBM once, the other groups hit L2).
  int ngroups;
  int b_slab_idx[8][4];    // per group: 4-D map coordinate of each slab (< 0: padding slab, loaded but not reduced)
  float* dW0[8][4];        // per group and slab: rows [0, m_split)
  float* dW1[4];           // group 0 only, per slab: rows [m_split, m_valid)
  int m_split, m_valid;
  int64_t sn, sk;          // element (m, slab, c) -> dW[slab] + m*sn + c*sk
  int chunks_per_seq, num_chunks;
  int reverse;             // walk the chunks from the last to the first
  int run;                 // consecutive chunks per CTA visit (0: one contiguous range per CTA)
  int red_mode;            // 0 scalar atomics, 1 rows contiguous (sk == 1), 2 two taps interleaved (sk == 2); 1/2 need 16-B alignment
};

// MN-major tf32 operands only exist in the SWIZZLE_128B_BASE32B layout (layout_type 1): 128-byte rows,
// 4-row atoms, 32-byte chunks XOR-ed with (row & 3) -- what TMA's SWIZZLE_128B_ATOM_32B deposits.
// LBO = bytes between 32-element M atoms, SBO = bytes between 4-row K groups.
__device__ __forceinline__ uint64_t umma_desc_mn_sw128_32b(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(lbo_bytes >> 4) << 16;
  d |= (uint64_t)(sbo_bytes >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)1 << 61;
  return d;
}

// MH = number of 128-row halves of dY handled per CTA (2 lets a 256-channel gradient read X only once)
template <int NB, int MH>
struct WgradCfg {
  static constexpr int KC = MH == 2 ? 32 : 64;          // positions per pipeline stage
  static constexpr int SUB = KC * 128;                  // bytes of a [KC x 32] sub-tile
  static constexpr int STAGE = (4 * MH + NB / 32) * SUB;
  static constexpr int STAGES = (200 * 1024) / STAGE > 4 ? 4 : (200 * 1024) / STAGE;
  static constexpr int BAR = STAGES * STAGE;
  static constexpr int SMEM = BAR + 256 + 1024;
  static constexpr int TMEM_COLS = MH * NB > 256 ? 512 : 256;
};

template <int NB, int MH>
__global__ void __launch_bounds__(L_THREADS, 1)
tc_wgrad_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b, const WgradTcArgs a) {
  using Cfg = WgradCfg<NB, MH>;
  constexpr int KC = Cfg::KC, SUB = Cfg::SUB;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bar0 = base + Cfg::BAR;
  auto full = [&](int s) { return bar0 + 8 * s; };
  auto empty = [&](int s) { return bar0 + 64 + 8 * s; };
  const uint32_t acc_full = bar0 + 128;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(gbase + Cfg::BAR + 192);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < Cfg::STAGES; ++s) {
      mbar_init(full(s), 1);
      mbar_init(empty(s), 1);
    }
    mbar_init(acc_full, 1);
    fence_barrier_init();
    prefetch_tmap(&tm_a);
    prefetch_tmap(&tm_b);
  }
  if (warp == 1) tmem_alloc<Cfg::TMEM_COLS>(smem_u32((const void*)tmem_slot));
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  pdl_launch_dependents();
  pdl_wait();
  const uint32_t tmem = *tmem_slot;
  // contiguous chunk range per CTA (per group of CTAs when the B slabs are grouped)
  const int grp = (int)blockIdx.x % a.ngroups, nrange = (int)gridDim.x / a.ngroups;
  // chunk order: the CTAs of one group interleave runs of WG_RUN consecutive chunks (DRAM/L2 locality inside a run, and
  // the whole grid sweeps the rows front to back -- or back to front (serpentine hand-off through L2))
  const int WG_RUN = a.run;      // 0: one contiguous range per CTA
  const int r0 = (int)blockIdx.x / a.ngroups;
  int n_local, c_begin = 0;
  if (WG_RUN > 0) {
    const int nruns = (a.num_chunks + WG_RUN - 1) / WG_RUN;
    const int my_runs = r0 < nruns ? (nruns - r0 + nrange - 1) / nrange : 0;
    n_local = my_runs * WG_RUN;
    if (my_runs > 0 && r0 + (my_runs - 1) * nrange == nruns - 1) n_local -= nruns * WG_RUN - a.num_chunks;   // partial last run
  } else {
    const int per = (a.num_chunks + nrange - 1) / nrange;
    c_begin = r0 * per;
    n_local = max(0, min(a.num_chunks, c_begin + per) - c_begin);
  }

  if (warp == 0) {
    if (lane == 0) {
      for (int it = 0; it < n_local; ++it) {
        const int fwd = WG_RUN > 0 ? (r0 + (it / WG_RUN) * nrange) * WG_RUN + it % WG_RUN : c_begin + it;
        const int chunk = a.reverse ? a.num_chunks - 1 - fwd : fwd;
        const int b = chunk / a.chunks_per_seq, t0 = (chunk % a.chunks_per_seq) * KC;
        const int s = it % Cfg::STAGES, ph = (it / Cfg::STAGES) & 1;
        mbar_wait(empty(s), ph ^ 1);
        const uint32_t st = base + s * Cfg::STAGE;
        mbar_arrive_expect_tx(full(s), Cfg::STAGE);
#pragma unroll
        for (int i = 0; i < 4 * MH; ++i) tma_load_4d(st + i * SUB, &tm_a, full(s), a.a_c0 + i * SUBK, a.a_row_off + t0, b, 0);
        for (int sl = 0; sl < a.nb_slab; ++sl)
          for (int i = 0; i < a.nb_sub; ++i)
            tma_load_4d(st + (4 * MH + sl * a.nb_sub + i) * SUB, &tm_b, full(s), i * SUBK, a.b_row_off[sl] + t0, b,
                        max(a.b_slab_idx[grp][sl], 0));
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && n_local > 0) {
      constexpr uint32_t idesc = umma_idesc_tf32(128, NB) | (1u << 15) | (1u << 16);   // both operands MN-major
      constexpr uint32_t lbo = SUB, sbo = 512, kstep = 1024;
      for (int it = 0; it < n_local; ++it) {
        const int s = it % Cfg::STAGES, ph = (it / Cfg::STAGES) & 1;
        mbar_wait(full(s), ph);
        tcgen05_fence_after();
        const uint32_t st = base + s * Cfg::STAGE;
#pragma unroll
        for (int k8 = 0; k8 < KC / 8; ++k8) {
          const uint64_t bdesc = umma_desc_mn_sw128_32b(st + 4 * MH * SUB + k8 * kstep, lbo, sbo);
#pragma unroll
          for (int mh = 0; mh < MH; ++mh)
            umma_tf32(tmem + mh * NB, umma_desc_mn_sw128_32b(st + 4 * mh * SUB + k8 * kstep, lbo, sbo), bdesc, idesc,
                      (it | k8) > 0);
        }
        umma_commit(empty(s));
      }
      umma_commit(acc_full);
    }
  } else if (n_local > 0) {
    const int q = warp & 3, half = (warp - 2) >> 2;
    constexpr int CH = NB / 64;
    mbar_wait(acc_full, 0);
    tcgen05_fence_after();
    const int nb = a.nb_sub * 32;   // channels per slab
    // Every MMA has retired, so the pipeline stages are free: each epilogue warp turns its row-per-lane TMEM
    // blocks into column-per-lane order through 8 KB of that memory and issues 16-byte vector reductions that
    // cover whole 128/256-byte rows of dW (a strided scalar atomic per element costs 8x the L2 transactions).
    uint8_t* stg = gbase + (warp - 2) * 8192;
    auto row_ptr = [&](int m, int sl) {
      return m < a.m_split ? a.dW0[grp][sl] + (int64_t)m * a.sn : a.dW1[sl] + (int64_t)(m - a.m_split) * a.sn;
    };
    if (a.red_mode == 2) {
      // two taps interleaved in memory (sk == 2, dW[1] == dW[0] + 1): this warp owns channels [cw, cw+32) of both taps
#pragma unroll 1
      for (int mh = 0; mh < MH; ++mh) {
#pragma unroll 1
        for (int ch = 0; ch < nb / 64; ++ch) {
          const int cw = (half * (nb / 64) + ch) * 32;
          uint32_t v0[32], v1[32];
          tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + mh * NB + cw, v0);
          tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + mh * NB + nb + cw, v1);
          tmem_ld_wait();
#pragma unroll
          for (int k = 0; k < 16; ++k)
            *reinterpret_cast<uint4*>(stg + lane * 256 + ((k ^ (lane & 7)) << 4)) =
                make_uint4(v0[2 * k], v1[2 * k], v0[2 * k + 1], v1[2 * k + 1]);
          __syncwarp();
#pragma unroll 4
          for (int jj = 0; jj < 16; ++jj) {
            const int rr = jj * 2 + (lane >> 4), kk = lane & 15;
            const int m = mh * 128 + q * 32 + rr;
            const float4 o = *reinterpret_cast<const float4*>(stg + rr * 256 + ((kk ^ (rr & 7)) << 4));
            if (m < a.m_valid) red_add_v4(row_ptr(m, 0) + (int64_t)cw * 2 + kk * 4, o);
          }
          __syncwarp();
        }
      }
    } else if (a.red_mode == 1) {
      // contiguous channels (sk == 1)
#pragma unroll 1
      for (int mh = 0; mh < MH; ++mh) {
#pragma unroll 1
        for (int ch = 0; ch < CH; ++ch) {
          const int c0 = (half * CH + ch) * 32;
          uint32_t v[32];
          tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + mh * NB + c0, v);
          tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < 8; ++c)
            *reinterpret_cast<uint4*>(stg + lane * 128 + ((c ^ (lane & 7)) << 4)) =
                make_uint4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
          __syncwarp();
          const int sl = c0 / nb, cbase = c0 % nb;
          const bool slab_ok = a.b_slab_idx[grp][sl] >= 0;
#pragma unroll
          for (int jj = 0; jj < 8; ++jj) {
            const int rr = jj * 4 + (lane >> 3), kk = lane & 7;
            const int m = mh * 128 + q * 32 + rr;
            const float4 o = *reinterpret_cast<const float4*>(stg + rr * 128 + ((kk ^ (rr & 7)) << 4));
            if (m < a.m_valid && slab_ok) red_add_v4(row_ptr(m, sl) + cbase + kk * 4, o);
          }
          __syncwarp();
        }
      }
    } else {
#pragma unroll 1
      for (int mh = 0; mh < MH; ++mh) {
        const int m = mh * 128 + q * 32 + lane;
#pragma unroll 1
        for (int ch = 0; ch < CH; ++ch) {
          const int c0 = (half * CH + ch) * 32;
          uint32_t v[32];
          tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + mh * NB + c0, v);
          tmem_ld_wait();
          const int sl = c0 / nb, cbase = c0 % nb;
          if (m < a.m_valid && a.b_slab_idx[grp][sl] >= 0) {
            float* wrow = row_ptr(m, sl);
#pragma unroll
            for (int i = 0; i < 32; ++i) atomicAdd(wrow + (int64_t)(cbase + i) * a.sk, __uint_as_float(v[i]));
          }
        }
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<Cfg::TMEM_COLS>(tmem);
}

// ------------------------------------------------------------------------------------------
// Backward of one residual layer's gate + projection for the fused shape (R = G = 64), one pass over the tile:
//   dz   = dout . Wp + dzs_l                       (K-major GEMM, N = 64)
//   dafg = gate derivative of dz (z, sigmoid)      (epilogue, as tc_gemm_kernel MODE 1)
//   dWp += dout^T . z                              (MN-major GEMM over the same 128 rows, accumulated in TMEM for the
//                                                   whole CTA and reduced into the gradient buffer at the end)
// dout and z are already on their way through L2 for the first two, so folding dWp in removes the separate
// weight-gradient pass (a second HBM read of dout and z for every layer).
struct GateBwdArgs {
  const float* dzs;        // [rows][64]
  const float* z;          // [rows][64]
  const float* sg;         // [rows][sg_ld]
  float* dafg;             // [rows][128]
  float* dWp;              // [64 o][64 c]
  int sg_ld, sg_half, zp;  // sg_half: sigmoid tape is fp16 (ld in elements)
  int rows_out, tiles_per_seq, num_tiles;
  int reverse;
};
constexpr int GB_STAGE = SUB_A + 64 * 128;     // dout [128 x 32] + Wp^T [64 x 32]
constexpr int GB_STAGES = 4;
constexpr int GB_MN = GB_STAGES * GB_STAGE;    // dout^T atoms 0,1 | zero atoms 2,3 | z atoms 0,1  (6 x 16 KB)
constexpr int GB_BAR = GB_MN + 6 * SUB_A;
constexpr int GB_STG = GB_BAR + 256;
constexpr int GB_SMEM = GB_STG + 8 * 4096 + 1024;

__global__ void __launch_bounds__(L_THREADS, 1)
tc_gate_bwd_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                   const __grid_constant__ CUtensorMap tm_a_mn, const __grid_constant__ CUtensorMap tm_z_mn,
                   const GateBwdArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bar0 = base + GB_BAR;
  auto full = [&](int s) { return bar0 + 8 * s; };
  auto empty = [&](int s) { return bar0 + 64 + 8 * s; };
  auto acc_full = [&](int s) { return bar0 + 128 + 8 * s; };
  auto acc_empty = [&](int s) { return bar0 + 144 + 8 * s; };
  const uint32_t mn_full = bar0 + 160, mn_empty = bar0 + 168, wg_full = bar0 + 176;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(gbase + GB_BAR + 192);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < GB_STAGES; ++s) {
      mbar_init(full(s), 1);
      mbar_init(empty(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(acc_full(s), 1);
      mbar_init(acc_empty(s), 256);
    }
    mbar_init(mn_full, 1);
    mbar_init(mn_empty, 1);
    mbar_init(wg_full, 1);
    fence_barrier_init();
    prefetch_tmap(&tm_a);
    prefetch_tmap(&tm_b);
    prefetch_tmap(&tm_a_mn);
    prefetch_tmap(&tm_z_mn);
  }
  // channels 64..127 of the M = 128 weight-gradient MMA: two zero atoms, written once
  for (int i = threadIdx.x; i < 2 * SUB_A / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(gbase + GB_MN + 2 * SUB_A)[i] = make_uint4(0u, 0u, 0u, 0u);
  fence_proxy_async();
  if (warp == 1) tmem_alloc<256>(smem_u32((const void*)tmem_slot));
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  pdl_launch_dependents();
  pdl_wait();
  const uint32_t tmem = *tmem_slot;
  const int n_local = (a.num_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  auto tile_of = [&](int j) {
    const int i = (int)blockIdx.x + j * (int)gridDim.x;
    return a.reverse ? a.num_tiles - 1 - i : i;
  };

  if (warp == 0) {
    if (lane == 0) {
      int it = 0;
      for (int j = 0; j < n_local; ++j) {
        const int tile = tile_of(j);
        const int b = tile / a.tiles_per_seq, t0 = (tile % a.tiles_per_seq) * TM;
        for (int ks = 0; ks < 2; ++ks, ++it) {
          const int s = it % GB_STAGES, ph = (it / GB_STAGES) & 1;
          mbar_wait(empty(s), ph ^ 1);
          const uint32_t st = base + s * GB_STAGE;
          mbar_arrive_expect_tx(full(s), GB_STAGE);
          tma_load_4d(st, &tm_a, full(s), ks * SUBK, t0, b, 0);
          tma_load_2d(st + SUB_A, &tm_b, full(s), ks * SUBK, 0);
        }
        TR(j, 0);
        // MN-major copies of the same rows for the weight gradient (single buffer: freed by the MMA warp)
        mbar_wait(mn_empty, (j & 1) ^ 1);
        TR(j, 1);
        mbar_arrive_expect_tx(mn_full, 4 * SUB_A);
        tma_load_4d(base + GB_MN + 0 * SUB_A, &tm_a_mn, mn_full, 0, t0, b, 0);
        tma_load_4d(base + GB_MN + 1 * SUB_A, &tm_a_mn, mn_full, SUBK, t0, b, 0);
        tma_load_4d(base + GB_MN + 4 * SUB_A, &tm_z_mn, mn_full, 0, t0, b, 0);
        tma_load_4d(base + GB_MN + 5 * SUB_A, &tm_z_mn, mn_full, SUBK, t0, b, 0);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_tf32(128, 64);
      constexpr uint32_t idesc_mn = umma_idesc_tf32(128, 64) | (1u << 15) | (1u << 16);
      int it = 0;
      for (int j = 0; j < n_local; ++j) {
        const int ab = j & 1, aph = (j >> 1) & 1;
        mbar_wait(acc_empty(ab), aph ^ 1);
        TR(j, 2);
        tcgen05_fence_after();
        for (int kk = 0; kk < 2; ++kk, ++it) {
          const int s = it % GB_STAGES, ph = (it / GB_STAGES) & 1;
          mbar_wait(full(s), ph);
          tcgen05_fence_after();
          const uint32_t st = base + s * GB_STAGE;
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4)
            umma_tf32(tmem + ab * 64, umma_desc_k_sw128(st + k4 * 32), umma_desc_k_sw128(st + SUB_A + k4 * 32), idesc,
                      (kk | k4) > 0);
          umma_commit(empty(s));
        }
        umma_commit(acc_full(ab));
        TR(j, 3);
        mbar_wait(mn_full, j & 1);
        TR(j, 4);
        tcgen05_fence_after();
#pragma unroll
        for (int k8 = 0; k8 < TM / 8; ++k8)
          umma_tf32(tmem + 128, umma_desc_mn_sw128_32b(base + GB_MN + k8 * 1024, SUB_A, 512),
                    umma_desc_mn_sw128_32b(base + GB_MN + 4 * SUB_A + k8 * 1024, SUB_A, 512), idesc_mn, (j | k8) > 0);
        umma_commit(mn_empty);
        TR(j, 5);
      }
      umma_commit(wg_full);
    }
  } else {
    const int q = warp & 3, half = (warp - 2) >> 2;
    uint8_t* stg = gbase + GB_STG + (warp - 2) * 4096;
    const int cc4 = (lane & 7) * 4, rsub = lane >> 3;
    const int c0 = half * 32, col = c0 + cc4;
    for (int j = 0; j < n_local; ++j) {
      const int tile = tile_of(j);
      const int ab = j & 1, aph = (j >> 1) & 1;
      const int b = tile / a.tiles_per_seq, t0 = (tile % a.tiles_per_seq) * TM + q * 32;
      // the three epilogue inputs do not depend on the accumulator: issue their loads before waiting for it
      float4 r4[8], z4[8], sg4[8];
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) {
        const int t = min(t0 + jj * 4 + rsub, a.rows_out - 1);
        const int64_t orow = (int64_t)b * a.rows_out + t;
        r4[jj] = *reinterpret_cast<const float4*>(a.dzs + orow * 64 + col);
        z4[jj] = *reinterpret_cast<const float4*>(a.z + orow * 64 + col);
        sg4[jj] = load_sg4(a.sg, orow * a.sg_ld + col, a.sg_half);
      }
      if (threadIdx.x == 64) TR(j, 8);
      mbar_wait(acc_full(ab), aph);
      if (threadIdx.x == 64) TR(j, 9);
      tcgen05_fence_after();
      uint32_t v[32];
      tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + ab * 64 + c0, v);
      tmem_ld_wait();
      tcgen05_fence_before();
      mbar_arrive(acc_empty(ab));
#pragma unroll
      for (int c = 0; c < 8; ++c)
        *reinterpret_cast<uint4*>(stg + lane * 128 + ((c ^ (lane & 7)) << 4)) =
            make_uint4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
      __syncwarp();
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) {
        const int rr = jj * 4 + rsub;
        const int t = t0 + rr;
        if (t >= a.rows_out) continue;
        const int64_t orow = (int64_t)b * a.rows_out + t;
        float4 o = *reinterpret_cast<const float4*>(stg + rr * 128 + (((lane & 7) ^ (rr & 7)) << 4));
        o.x += r4[jj].x, o.y += r4[jj].y, o.z += r4[jj].z, o.w += r4[jj].w;
        // with z = tanh*sg:  da_f = dz*sg*(1-tanh^2) = dz*(sg - z*z/sg),  da_g = dz*tanh*sg*(1-sg) = dz*z*(1-sg);
        // rows inside the zero prefix get none (Q1)
        const float4 z = z4[jj], sg = sg4[jj];
        const float live = t >= a.zp ? 1.f : 0.f;
        float4 df, dg;
        df.x = live * o.x * gate_dtanh(z.x, sg.x), dg.x = live * o.x * z.x * (1.f - sg.x);
        df.y = live * o.y * gate_dtanh(z.y, sg.y), dg.y = live * o.y * z.y * (1.f - sg.y);
        df.z = live * o.z * gate_dtanh(z.z, sg.z), dg.z = live * o.z * z.z * (1.f - sg.z);
        df.w = live * o.w * gate_dtanh(z.w, sg.w), dg.w = live * o.w * z.w * (1.f - sg.w);
        float* drow = a.dafg + orow * 128;
        *reinterpret_cast<float4*>(drow + col) = df;
        *reinterpret_cast<float4*>(drow + 64 + col) = dg;
      }
      __syncwarp();
      if (threadIdx.x == 64) TR(j, 10);
    }
    if (n_local > 0 && q < 2) {
      // dWp rows (projection output channels) live in TMEM lanes 0..63
      mbar_wait(wg_full, 0);
      tcgen05_fence_after();
      uint32_t v[32];
      tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + 128 + c0, v);
      tmem_ld_wait();
#pragma unroll
      for (int c = 0; c < 8; ++c)
        *reinterpret_cast<uint4*>(stg + lane * 128 + ((c ^ (lane & 7)) << 4)) =
            make_uint4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
      __syncwarp();
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) {
        const int rr = jj * 4 + rsub;
        const float4 o = *reinterpret_cast<const float4*>(stg + rr * 128 + (((lane & 7) ^ (rr & 7)) << 4));
        red_add_v4(a.dWp + (int64_t)(q * 32 + rr) * 64 + col, o);
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<256>(tmem);
}

// ------------------------------------------------------------------------------------------
// Backward of one residual layer's dilated conv for the fused shape (R = G = 64, two taps), one pass over dafg:
//   dx[t]  = dout[t] + dafg[t] . W1(tap 1) + dafg[t+d] . W1(tap 0)       (K-major GEMM, K = 2 x 128, N = 64)
//   dW_f/g(o, c, tap) += sum_t dafg[t][o] * x[t - (1-tap) d][c]             (MN-major GEMM over the same rows, M = 128, N = 2 x 64)
// dafg is read from HBM once (the MN-major copy of the rows the dx GEMM just loaded comes from L2) instead of once per
// kernel; the weight-gradient accumulator stays in TMEM for the CTA's whole range and is reduced at the end.
struct DxwArgs {
  const float* rsd;        // dout of the layer above, [rows][64]; null for the top layer
  float* Y;                // new dout, [rows][64]
  float* dWf;              // (o, c, tap) with row stride 128: rows [0, 64) of the accumulator
  float* dWg;              // rows [64, 128)
  int d;
  int rows_out, tiles_per_seq, num_tiles;
  int reverse;
};
constexpr int DX_STAGE = SUB_A + 64 * 128;     // dafg [128 x 32] + W1t [64 x 32]
constexpr int DX_STAGES = 4;
constexpr int DX_MN = DX_STAGES * DX_STAGE;    // one 64-row MN-major chunk: dafg atoms 0..3 | x(t-d) atoms 0,1 | x(t) atoms 0,1
constexpr int DX_MN_SUB = 64 * 128;
constexpr int DX_BAR = DX_MN + 8 * DX_MN_SUB;
constexpr int DX_STG = DX_BAR + 256;
constexpr int DX_SMEM = DX_STG + 8 * 4096 + 1024;

__global__ void __launch_bounds__(L_THREADS, 1)
tc_dxw_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
              const __grid_constant__ CUtensorMap tm_a_mn, const __grid_constant__ CUtensorMap tm_x_mn, const DxwArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bar0 = base + DX_BAR;
  auto full = [&](int s) { return bar0 + 8 * s; };
  auto empty = [&](int s) { return bar0 + 32 + 8 * s; };
  auto acc_full = [&](int s) { return bar0 + 64 + 8 * s; };
  auto acc_empty = [&](int s) { return bar0 + 80 + 8 * s; };
  const uint32_t mn_full = bar0 + 96, mn_empty = bar0 + 104, wg_full = bar0 + 112;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(gbase + DX_BAR + 192);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < DX_STAGES; ++s) {
      mbar_init(full(s), 1);
      mbar_init(empty(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(acc_full(s), 1);
      mbar_init(acc_empty(s), 256);
    }
    mbar_init(mn_full, 1);
    mbar_init(mn_empty, 1);
    mbar_init(wg_full, 1);
    fence_barrier_init();
    prefetch_tmap(&tm_a);
    prefetch_tmap(&tm_b);
    prefetch_tmap(&tm_a_mn);
    prefetch_tmap(&tm_x_mn);
  }
  if (warp == 1) tmem_alloc<256>(smem_u32((const void*)tmem_slot));
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  pdl_launch_dependents();
  pdl_wait();
  const uint32_t tmem = *tmem_slot;
  const int n_local = (a.num_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  auto tile_of = [&](int j) {
    const int i = (int)blockIdx.x + j * (int)gridDim.x;
    return a.reverse ? a.num_tiles - 1 - i : i;
  };

  if (warp == 0) {
    if (lane == 0) {
      // two load streams, polled so that neither blocks the other: K-major k-steps (8 per tile through the ring) and the
      // single-buffered MN-major 64-row chunks (2 per tile).  A chunk is requested only after the K-major loads of its
      // tile are out (its dafg rows are then on their way into L2), and the x rows of the NEXT chunk are prefetched
      // into L2 because their load cannot start before the MMA thread releases the buffer.
      int uk = 0, um = 0;
      uint32_t spins = 0;
      auto chunk_coords = [&](int u, int& b, int& tc0) {
        const int tile = tile_of(u >> 1);
        b = tile / a.tiles_per_seq;
        tc0 = (tile % a.tiles_per_seq) * TM + (u & 1) * 64;
      };
      while (uk < 8 * n_local || um < 2 * n_local) {
        bool progressed = false;
        if (uk < 8 * n_local) {
          const int s = uk % DX_STAGES, ph = (uk / DX_STAGES) & 1;
          if (mbar_try_wait(empty(s), ph ^ 1)) {
            const int tile = tile_of(uk >> 3), sl = (uk >> 2) & 1, ks = uk & 3;
            const int b = tile / a.tiles_per_seq, t0 = (tile % a.tiles_per_seq) * TM;
            const uint32_t st = base + s * DX_STAGE;
            mbar_arrive_expect_tx(full(s), DX_STAGE);
            tma_load_4d(st, &tm_a, full(s), ks * SUBK, t0 + (sl ? a.d : 0), b, 0);
            tma_load_2d(st + SUB_A, &tm_b, full(s), (sl * 4 + ks) * SUBK, 0);
            ++uk;
            progressed = true;
          }
        }
        if (um < 2 * n_local && (um >> 1) * 8 + 4 <= uk && mbar_try_wait(mn_empty, (um & 1) ^ 1)) {
          int b, tc0;
          chunk_coords(um, b, tc0);
          const uint32_t st = base + DX_MN;
          mbar_arrive_expect_tx(mn_full, 8 * DX_MN_SUB);
#pragma unroll
          for (int i = 0; i < 4; ++i) tma_load_4d(st + i * DX_MN_SUB, &tm_a_mn, mn_full, i * SUBK, tc0, b, 0);
          tma_load_4d(st + 4 * DX_MN_SUB, &tm_x_mn, mn_full, 0, tc0 - a.d, b, 0);
          tma_load_4d(st + 5 * DX_MN_SUB, &tm_x_mn, mn_full, SUBK, tc0 - a.d, b, 0);
          tma_load_4d(st + 6 * DX_MN_SUB, &tm_x_mn, mn_full, 0, tc0, b, 0);
          tma_load_4d(st + 7 * DX_MN_SUB, &tm_x_mn, mn_full, SUBK, tc0, b, 0);
          if (um + 1 < 2 * n_local) {
            int bn, tn0;
            chunk_coords(um + 1, bn, tn0);
            tma_prefetch_4d(&tm_x_mn, 0, tn0, bn, 0);
            tma_prefetch_4d(&tm_x_mn, SUBK, tn0, bn, 0);
          }
          ++um;
          progressed = true;
        }
        if (progressed) {
          spins = 0;
        } else if (++spins > (1u << 26)) {
          printf("wavenet_b200: dxw producer timeout (block %d, uk %d, um %d)\n", (int)blockIdx.x, uk, um);
          __trap();
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_tf32(128, 64);
      constexpr uint32_t idesc_mn = umma_idesc_tf32(128, 128) | (1u << 15) | (1u << 16);
      // two MMA streams, polled: the dx GEMM (k-step units) and the weight gradient (64-row chunk units)
      int uk = 0, um = 0;
      uint32_t spins = 0;
      while (uk < 8 * n_local || um < 2 * n_local) {
        bool progressed = false;
        if (uk < 8 * n_local) {
          const int j = uk >> 3, kk = uk & 7, ab = j & 1, aph = (j >> 1) & 1;
          const int s = uk % DX_STAGES, ph = (uk / DX_STAGES) & 1;
          if ((kk > 0 || mbar_try_wait(acc_empty(ab), aph ^ 1)) && mbar_try_wait(full(s), ph)) {
            tcgen05_fence_after();
            const uint32_t st = base + s * DX_STAGE;
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4)
              umma_tf32(tmem + ab * 64, umma_desc_k_sw128(st + k4 * 32), umma_desc_k_sw128(st + SUB_A + k4 * 32), idesc,
                        (kk | k4) > 0);
            umma_commit(empty(s));
            if (kk == 7) umma_commit(acc_full(ab));
            ++uk;
            progressed = true;
          }
        }
        if (um < 2 * n_local && mbar_try_wait(mn_full, um & 1)) {
          tcgen05_fence_after();
#pragma unroll
          for (int k8 = 0; k8 < 8; ++k8)
            umma_tf32(tmem + 128, umma_desc_mn_sw128_32b(base + DX_MN + k8 * 1024, DX_MN_SUB, 512),
                      umma_desc_mn_sw128_32b(base + DX_MN + 4 * DX_MN_SUB + k8 * 1024, DX_MN_SUB, 512), idesc_mn,
                      (um | k8) > 0);
          umma_commit(mn_empty);
          ++um;
          progressed = true;
        }
        if (progressed) {
          spins = 0;
        } else if (++spins > (1u << 26)) {
          printf("wavenet_b200: dxw MMA timeout (block %d, uk %d, um %d)\n", (int)blockIdx.x, uk, um);
          __trap();
        }
      }
      umma_commit(wg_full);
    }
  } else {
    const int q = warp & 3, half = (warp - 2) >> 2;
    uint8_t* stg = gbase + DX_STG + (warp - 2) * 4096;
    const int cc4 = (lane & 7) * 4, rsub = lane >> 3;
    const int c0 = half * 32, col = c0 + cc4;
    for (int j = 0; j < n_local; ++j) {
      const int tile = tile_of(j);
      const int ab = j & 1, aph = (j >> 1) & 1;
      const int b = tile / a.tiles_per_seq, t0 = (tile % a.tiles_per_seq) * TM + q * 32;
      float4 r4[8];      // residual gradient: independent of the accumulator, so loaded before waiting for it
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) {
        const int t = min(t0 + jj * 4 + rsub, a.rows_out - 1);
        r4[jj] = a.rsd ? *reinterpret_cast<const float4*>(a.rsd + ((int64_t)b * a.rows_out + t) * 64 + col)
                       : make_float4(0.f, 0.f, 0.f, 0.f);   // top layer: nothing flows in from above
      }
      mbar_wait(acc_full(ab), aph);
      tcgen05_fence_after();
      uint32_t v[32];
      tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + ab * 64 + c0, v);
      tmem_ld_wait();
      tcgen05_fence_before();
      mbar_arrive(acc_empty(ab));
#pragma unroll
      for (int c = 0; c < 8; ++c)
        *reinterpret_cast<uint4*>(stg + lane * 128 + ((c ^ (lane & 7)) << 4)) =
            make_uint4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
      __syncwarp();
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) {
        const int rr = jj * 4 + rsub;
        const int t = t0 + rr;
        if (t >= a.rows_out) continue;
        float4 o = *reinterpret_cast<const float4*>(stg + rr * 128 + (((lane & 7) ^ (rr & 7)) << 4));
        o.x += r4[jj].x, o.y += r4[jj].y, o.z += r4[jj].z, o.w += r4[jj].w;
        *reinterpret_cast<float4*>(a.Y + ((int64_t)b * a.rows_out + t) * 64 + col) = o;
      }
      __syncwarp();
    }
    if (n_local > 0) {
      // dW_f | dW_g: TMEM lanes = gate channel, columns = tap*64 + c.  Both taps of 32 input channels are interleaved in
      // shared memory (the gradient layout is (o, c, tap)) and reduced with 16-byte vector REDs over whole 256-byte rows.
      mbar_wait(wg_full, 0);
      tcgen05_fence_after();
      uint8_t* stg2 = gbase + (warp - 2) * 8192;          // every MMA has retired: the K-major ring is free
      uint32_t v0[32], v1[32];
      tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + 128 + c0, v0);
      tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + 128 + 64 + c0, v1);
      tmem_ld_wait();
#pragma unroll
      for (int k = 0; k < 16; ++k)
        *reinterpret_cast<uint4*>(stg2 + lane * 256 + ((k ^ (lane & 7)) << 4)) =
            make_uint4(v0[2 * k], v1[2 * k], v0[2 * k + 1], v1[2 * k + 1]);
      __syncwarp();
#pragma unroll 4
      for (int jj = 0; jj < 16; ++jj) {
        const int rr = jj * 2 + (lane >> 4), kk = lane & 15;
        const int m = q * 32 + rr;
        const float4 o = *reinterpret_cast<const float4*>(stg2 + rr * 256 + ((kk ^ (rr & 7)) << 4));
        float* rowp = m < 64 ? a.dWf + (int64_t)m * 128 : a.dWg + (int64_t)(m - 64) * 128;
        red_add_v4(rowp + c0 * 2 + kk * 4, o);
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<256>(tmem);
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
                uint64_t s2, uint64_t s3, uint32_t box1, CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B) {
  EncodeTiledFn enc = get_encode();
  WN_REQUIRE(enc, WN_ECUDA, "cuTensorMapEncodeTiled is unavailable");
  cuuint64_t dims[4] = {d0, d1, d2, d3};
  cuuint64_t strides[3] = {s1 * 4, s2 * 4, s3 * 4};
  cuuint32_t box[4] = {SUBK, box1, 1, 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)ptr, dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  WN_REQUIRE(r == CUDA_SUCCESS, WN_ECUDA, "cuTensorMapEncodeTiled(4d) failed: %d", (int)r);
  return WN_OK;
}

// fp16 tensor [d2][d1][d0 = 64] viewed 4-D, box [1][1][box1][64] (128-byte rows), SWIZZLE_128B
int make_map_4d_f16(CUtensorMap* m, const void* ptr, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t s1, uint64_t s2,
                    uint32_t box1) {
  EncodeTiledFn enc = get_encode();
  WN_REQUIRE(enc, WN_ECUDA, "cuTensorMapEncodeTiled is unavailable");
  WN_REQUIRE(d0 == 64, WN_EINVAL, "fp16 tile maps are 64 channels wide");
  cuuint64_t dims[4] = {d0, d1, d2, 1};
  cuuint64_t strides[3] = {s1 * 2, s2 * 2, s2 * d2 * 2};
  cuuint32_t box[4] = {64, box1, 1, 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, (void*)ptr, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  WN_REQUIRE(r == CUDA_SUCCESS, WN_ECUDA, "cuTensorMapEncodeTiled(f16) failed: %d", (int)r);
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

// Launch with programmatic stream serialization (PDL): kernel N+1's CTAs may start while kernel N drains; the kernels
// call pdl_wait() before touching global memory.  Opt-in (WN_PDL=1): measured neutral on the config-C train step
// (18.58 vs 18.59 ms), where the inter-kernel gaps are not launch latency, so plain launches stay the default.
template <typename... KArgs, typename... Args>
cudaError_t launch_pdl(void (*kern)(KArgs...), int grid, int block, size_t smem, cudaStream_t s, Args&&... args) {
  static const bool pdl = getenv("WN_PDL") != nullptr;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3((unsigned)block);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}

template <int BN, int MODE>
int launch_gemm_mode(const CUtensorMap& ta, const CUtensorMap& tb, const GemmTcArgs& g, int sm_count, cudaStream_t s) {
  using Cfg = GemmCfg<BN>;
  static bool attr = false;
  if (!attr) {
    WN_CHECK_CUDA(cudaFuncSetAttribute(tc_gemm_kernel<BN, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
    attr = true;
  }
  const int grid = g.num_tiles < sm_count ? g.num_tiles : sm_count;
  WN_CHECK_CUDA(launch_pdl(tc_gemm_kernel<BN, MODE>, grid, L_THREADS, Cfg::SMEM, s, ta, tb, g));
  WN_CHECK_LAUNCH();
  return WN_OK;
}

template <int BN>
int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const GemmTcArgs& g, int sm_count, cudaStream_t s) {
  if (g.gate_sg) return launch_gemm_mode<BN, 1>(ta, tb, g, sm_count, s);
  const bool rsd_only = g.Rsd && !g.bias && !g.relu && !g.round_out && !g.accumulate && !g.mask && g.y_slab_cols == 0 &&
                        g.zero_rows_below <= 0 && g.N == BN;
  if (rsd_only) return launch_gemm_mode<BN, 2>(ta, tb, g, sm_count, s);
  if (g.colsum_out && g.ngroups == 1) return launch_gemm_mode<BN, 3>(ta, tb, g, sm_count, s);
  return launch_gemm_mode<BN, 0>(ta, tb, g, sm_count, s);
}

template <int NB, int MH>
int launch_wgrad(const CUtensorMap& ta, const CUtensorMap& tb, const WgradTcArgs& g, int sm_count, cudaStream_t s) {
  using Cfg = WgradCfg<NB, MH>;
  static bool attr = false;
  if (!attr) {
    WN_CHECK_CUDA(cudaFuncSetAttribute(tc_wgrad_kernel<NB, MH>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
    attr = true;
  }
  int grid = g.num_chunks * g.ngroups < sm_count ? g.num_chunks * g.ngroups : sm_count;
  grid -= grid % g.ngroups;
  WN_CHECK_CUDA(launch_pdl(tc_wgrad_kernel<NB, MH>, grid, L_THREADS, Cfg::SMEM, s, ta, tb, g));
  WN_CHECK_LAUNCH();
  return WN_OK;
}

}  // namespace

// One operand of a slab GEMM: rows of sequence b are A[slab][b][row_off + t][0..K)
struct TcOperand {
  const float* ptr;
  int K;            // channels (row length, contiguous)
  int rows_in;      // rows per sequence in memory
  int num_seq;
  int nslab;        // number of equally spaced slabs behind ptr (1 for a plain tensor)
  int64_t slab_stride;
};

struct TcEpilogue {
  const float* bias = nullptr;
  int relu = 0, round_out = 0, accumulate = 0;
  const float* Rsd = nullptr;
  int ldr = 0;
  const float* mask = nullptr;
  int ldm = 0, mask_rows_in = 0, mask_row_off = 0;
  int y_slab_cols = 0;
  int64_t y_slab_stride = 0;
  const float* gate_sg = nullptr;
  const float* gate_z = nullptr;
  float* gate_dafg = nullptr;
  int gate_zp = 0, gate_sg_ld = 0, gate_sg_half = 0;
  int zero_rows_below = 0;
  int reverse = 0;
  float* colsum_out = nullptr;
};

// Y[(b, t)][0..N) = epi( sum_s A[slab_idx[s]][b][t + row_off[s]][:] . Wt[:, s*K ..]^T ),  Wt is [N][ns*K] (TF32-rounded)
int tc_gemm(const wn_handle* h, const TcOperand& A, int ns, const int* slab_idx, const int* row_off, int rows_out,
            const float* Wt, int N, const TcEpilogue& e, float* Y, int ldy, cudaStream_t s) {
  const bool plain = !e.gate_sg && !e.Rsd && !e.mask && !e.accumulate;   // wide outputs: column groups (MODE 0 only)
  WN_REQUIRE(A.K % SUBK == 0 && N % 32 == 0 && (N <= 256 || plain) && ns <= MAX_SLABS, WN_EINVAL,
             "tc_gemm: unsupported shape K=%d N=%d slabs=%d", A.K, N, ns);
  CUtensorMap ta, tb;
  const int64_t sstride = A.nslab > 1 ? A.slab_stride : (int64_t)A.rows_in * A.num_seq * A.K;
  WN_TRY(make_map_4d(&ta, A.ptr, A.K, A.rows_in, A.num_seq, A.nslab, A.K, (uint64_t)A.rows_in * A.K, sstride, TM));
  const int BN = N <= 64 ? 64 : (N <= 128 ? 128 : 256);
  WN_TRY(make_map_2d(&tb, Wt, (uint64_t)ns * A.K, N, (uint64_t)ns * A.K, BN));
  GemmTcArgs g;
  memset(&g, 0, sizeof(g));
  g.Y = Y;
  g.ldy = ldy;
  g.bias = e.bias;
  g.N = N;
  g.relu = e.relu;
  g.round_out = e.round_out;
  g.accumulate = e.accumulate;
  g.Rsd = e.Rsd;
  g.ldr = e.ldr;
  g.mask = e.mask;
  g.ldm = e.ldm;
  g.mask_rows_in = e.mask_rows_in;
  g.mask_row_off = e.mask_row_off;
  g.y_slab_cols = e.y_slab_cols;
  g.y_slab_stride = e.y_slab_stride;
  g.gate_sg = e.gate_sg;
  g.gate_z = e.gate_z;
  g.gate_dafg = e.gate_dafg;
  g.gate_zp = e.gate_zp;
  g.gate_sg_ld = e.gate_sg_ld ? e.gate_sg_ld : N;
  g.gate_sg_half = e.gate_sg_half;
  g.zero_rows_below = e.zero_rows_below;
  g.reverse = e.reverse;
  g.colsum_out = e.colsum_out;
  g.rows_out = rows_out;
  g.nslab = ns;
  g.ksub = A.K / SUBK;
  for (int i = 0; i < ns; ++i) {
    g.slab_row_off[i] = row_off ? row_off[i] : 0;
    g.slab_idx[i] = slab_idx ? slab_idx[i] : 0;
  }
  g.tiles_per_seq = (rows_out + TM - 1) / TM;
  g.ngroups = (N + BN - 1) / BN;
  g.num_tiles = g.tiles_per_seq * A.num_seq * g.ngroups;
  if (BN == 64) return launch_gemm<64>(ta, tb, g, h->sm_count, s);
  if (BN == 128) return launch_gemm<128>(ta, tb, g, h->sm_count, s);
  return launch_gemm<256>(ta, tb, g, h->sm_count, s);
}

// dW[slab](m, c) += sum_{b,t} dY[b][a_row_off + t][a_c0 + m] * X[slab_idx[s]][b][b_row_off[s] + t][c]   for m < m_valid (<= 128)
int tc_wgrad(const wn_handle* h, const TcOperand& dY, int a_row_off, int a_c0, int m_valid, const TcOperand& X, int nb_slab,
             const int* b_row_off, const int* b_slab_idx, float* const* dW0, float* const* dW1, int m_split, int rows_it,
             int64_t sn, int64_t sk, cudaStream_t s, int ngroups = 1, int reverse = 0) {
  // ngroups > 1: b_slab_idx and dW0 hold ngroups x nb_slab entries (group-major); a negative slab index pads a group
  const int NB = nb_slab * X.K;
  WN_REQUIRE(X.K % 32 == 0 && (NB == 64 || NB == 128 || NB == 256) && nb_slab <= 4 && m_valid <= 256 && ngroups >= 1 &&
                 ngroups <= 8 && (ngroups == 1 || !dW1),
             WN_EINVAL, "tc_wgrad: unsupported shape X.K=%d slabs=%d M=%d groups=%d", X.K, nb_slab, m_valid, ngroups);
  const int MH = m_valid > 128 ? 2 : 1;
  const int WG_KC = MH == 2 ? 32 : 64;
  CUtensorMap ta, tb;
  WN_TRY(make_map_4d(&ta, dY.ptr, dY.K, dY.rows_in, dY.num_seq, 1, dY.K, (uint64_t)dY.rows_in * dY.K,
                     (uint64_t)dY.rows_in * dY.num_seq * dY.K, WG_KC, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B));
  const int64_t xstride = X.nslab > 1 ? X.slab_stride : (int64_t)X.rows_in * X.num_seq * X.K;
  WN_TRY(make_map_4d(&tb, X.ptr, X.K, X.rows_in, X.num_seq, X.nslab, X.K, (uint64_t)X.rows_in * X.K, xstride, WG_KC,
                     CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B));
  WgradTcArgs g;
  memset(&g, 0, sizeof(g));
  g.rows_it = rows_it;
  g.num_seq = dY.num_seq;
  g.a_row_off = a_row_off;
  g.a_c0 = a_c0;
  g.nb_slab = nb_slab;
  g.nb_sub = X.K / 32;
  g.ngroups = ngroups;
  g.reverse = reverse;
  {
    static const int run_env = getenv("WN_WG_RUN") ? atoi(getenv("WN_WG_RUN")) : 8;
    g.run = run_env;
  }
  bool aligned = sn % 4 == 0;
  for (int i = 0; i < nb_slab; ++i) {
    g.b_row_off[i] = b_row_off[i];
    g.dW1[i] = dW1 ? dW1[i] : nullptr;
    if (g.dW1[i]) aligned = aligned && ((uintptr_t)g.dW1[i] & 15) == 0;
    for (int gr = 0; gr < ngroups; ++gr) {
      g.b_slab_idx[gr][i] = b_slab_idx ? b_slab_idx[gr * nb_slab + i] : 0;
      g.dW0[gr][i] = dW0[gr * nb_slab + i];
      if (g.b_slab_idx[gr][i] >= 0) aligned = aligned && ((uintptr_t)g.dW0[gr][i] & 15) == 0;
    }
  }
  g.m_split = m_split;
  g.m_valid = m_valid;
  g.sn = sn;
  g.sk = sk;
  g.chunks_per_seq = (rows_it + WG_KC - 1) / WG_KC;
  g.num_chunks = g.chunks_per_seq * dY.num_seq;
  {
    const bool taps2 = ngroups == 1 && nb_slab == 2 && sk == 2 && X.K % 64 == 0 && g.dW0[0][1] == g.dW0[0][0] + 1 &&
                       (!g.dW1[0] || g.dW1[1] == g.dW1[0] + 1) && sn % 4 == 0 && ((uintptr_t)g.dW0[0][0] & 15) == 0 &&
                       (!g.dW1[0] || ((uintptr_t)g.dW1[0] & 15) == 0);
    g.red_mode = (sk == 1 && aligned) ? 1 : (taps2 ? 2 : 0);
  }
  if (MH == 1) {
    if (NB == 64) return launch_wgrad<64, 1>(ta, tb, g, h->sm_count, s);
    if (NB == 128) return launch_wgrad<128, 1>(ta, tb, g, h->sm_count, s);
    return launch_wgrad<256, 1>(ta, tb, g, h->sm_count, s);
  }
  if (NB == 64) return launch_wgrad<64, 2>(ta, tb, g, h->sm_count, s);
  if (NB == 128) return launch_wgrad<128, 2>(ta, tb, g, h->sm_count, s);
  return launch_wgrad<256, 2>(ta, tb, g, h->sm_count, s);
}

// Y slabs <- A . Wt^T with the A tile resident (see tc_gemm_ares_kernel); requires K <= 256, K % 32 == 0
int tc_gemm_ares(const wn_handle* h, const TcOperand& A, int row_off, int rows_out, const float* Wt, int Ntot, float* Y,
                 int ldy, int y_slab_cols, int64_t y_slab_stride, cudaStream_t s) {
  WN_REQUIRE(A.K % SUBK == 0 && A.K <= 256 && y_slab_cols % 32 == 0 && A.nslab == 1, WN_EINVAL,
             "tc_gemm_ares: unsupported shape K=%d", A.K);
  CUtensorMap ta, tb;
  WN_TRY(make_map_4d(&ta, A.ptr, A.K, A.rows_in, A.num_seq, 1, A.K, (uint64_t)A.rows_in * A.K,
                     (uint64_t)A.rows_in * A.num_seq * A.K, TM));
  WN_TRY(make_map_2d(&tb, Wt, A.K, Ntot, A.K, 256));
  GemmAresArgs g;
  memset(&g, 0, sizeof(g));
  g.Y = Y;
  g.ldy = ldy;
  g.y_slab_cols = y_slab_cols;
  g.y_slab_stride = y_slab_stride;
  g.Ntot = Ntot;
  g.ngroups = (Ntot + 255) / 256;
  g.ksub = A.K / SUBK;
  g.rows_out = rows_out;
  g.a_row_off = row_off;
  g.tiles_per_seq = (rows_out + TM - 1) / TM;
  g.num_tiles = g.tiles_per_seq * A.num_seq;
  static bool attr = false;
  if (!attr) {
    WN_CHECK_CUDA(cudaFuncSetAttribute(tc_gemm_ares_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AR_SMEM));
    attr = true;
  }
  const int grid = g.num_tiles < h->sm_count ? g.num_tiles : h->sm_count;
  tc_gemm_ares_kernel<<<grid, L_THREADS, AR_SMEM, s>>>(ta, tb, g);
  WN_CHECK_LAUNCH();
  return WN_OK;
}

// dz GEMM + gate derivative + dWp in one pass (fused shape R = G = 64); dWp must be 16-byte aligned
int tc_gate_bwd(const wn_handle* h, const float* dout, const float* wpt, const float* dzs, const float* z, const float* sg,
                int sg_ld, int sg_half, float* dafg, float* dWp, int zp, int rows, int num_seq, int reverse, cudaStream_t s) {
  CUtensorMap ta, tb, tam, tzm;
  const uint64_t seq = (uint64_t)rows * 64, all = seq * num_seq;
  WN_TRY(make_map_4d(&ta, dout, 64, rows, num_seq, 1, 64, seq, all, TM));
  WN_TRY(make_map_2d(&tb, wpt, 64, 64, 64, 64));
  WN_TRY(make_map_4d(&tam, dout, 64, rows, num_seq, 1, 64, seq, all, TM, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B));
  WN_TRY(make_map_4d(&tzm, z, 64, rows, num_seq, 1, 64, seq, all, TM, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B));
  GateBwdArgs g;
  memset(&g, 0, sizeof(g));
  g.dzs = dzs;
  g.z = z;
  g.sg = sg;
  g.sg_ld = sg_ld;
  g.sg_half = sg_half;
  g.dafg = dafg;
  g.dWp = dWp;
  g.zp = zp;
  g.reverse = reverse;
  g.rows_out = rows;
  g.tiles_per_seq = (rows + TM - 1) / TM;
  g.num_tiles = g.tiles_per_seq * num_seq;
  static bool attr = false;
  if (!attr) {
    WN_CHECK_CUDA(cudaFuncSetAttribute(tc_gate_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GB_SMEM));
    attr = true;
  }
  const int grid = g.num_tiles < h->sm_count ? g.num_tiles : h->sm_count;
  WN_CHECK_CUDA(launch_pdl(tc_gate_bwd_kernel, grid, L_THREADS, GB_SMEM, s, ta, tb, tam, tzm, g));
  WN_CHECK_LAUNCH();
  return WN_OK;
}

// dx GEMM + dW_f/dW_g in one pass over dafg (fused shape R = G = 64, two taps); gradient pointers 16-byte aligned
int tc_dxw(const wn_handle* h, const float* dafg, const float* w1t, const float* rsd, const float* x, float* Y, float* dWf,
           float* dWg, int d, int rows, int num_seq, int reverse, cudaStream_t s) {
  CUtensorMap ta, tb, tam, txm;
  const uint64_t seqa = (uint64_t)rows * 128, alla = seqa * num_seq, seqx = (uint64_t)rows * 64, allx = seqx * num_seq;
  WN_TRY(make_map_4d(&ta, dafg, 128, rows, num_seq, 1, 128, seqa, alla, TM));
  WN_TRY(make_map_2d(&tb, w1t, 256, 64, 256, 64));
  WN_TRY(make_map_4d(&tam, dafg, 128, rows, num_seq, 1, 128, seqa, alla, 64, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B));
  WN_TRY(make_map_4d(&txm, x, 64, rows, num_seq, 1, 64, seqx, allx, 64, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B));
  DxwArgs g;
  memset(&g, 0, sizeof(g));
  g.rsd = rsd;
  g.Y = Y;
  g.dWf = dWf;
  g.dWg = dWg;
  g.d = d;
  g.reverse = reverse;
  g.rows_out = rows;
  g.tiles_per_seq = (rows + TM - 1) / TM;
  g.num_tiles = g.tiles_per_seq * num_seq;
  static bool attr = false;
  if (!attr) {
    WN_CHECK_CUDA(cudaFuncSetAttribute(tc_dxw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DX_SMEM));
    attr = true;
  }
  const int grid = g.num_tiles < h->sm_count ? g.num_tiles : h->sm_count;
  WN_CHECK_CUDA(launch_pdl(tc_dxw_kernel, grid, L_THREADS, DX_SMEM, s, ta, tb, tam, txm, g));
  WN_CHECK_LAUNCH();
  return WN_OK;
}

// fused layer kernel: the benchmark shape (R = G = 64, k = 2, no biases)
bool tc_fused_supported(const wn_handle* h) {
  if (h->R != 64 || h->cfg.residual_filter_width != 2) return false;
  for (const ResLayer& l : h->layers)
    if (l.G != 64 || l.wf.b_off >= 0 || l.proj.b_off >= 0) return false;
  if (h->S % 32 != 0 || h->S > 256) return false;
  return get_encode() != nullptr;
}

// any network whose channel counts are multiples of 32 and fit one N tile: layers run as slab GEMMs
// (gate on SIMT), e.g. the reference default R=256/G=128 (train_audio/model.py:24-43) or Params() defaults
bool tc_layer_supported(const wn_handle* h) {
  if (tc_fused_supported(h)) return true;
  if (h->cfg.residual_filter_width != 2 || h->R % 32 != 0 || h->R > 256 || h->S % 32 != 0 || h->S > 256) return false;
  const int G = h->layers[0].G;
  if (G % 32 != 0 || 2 * G > 256 || (int)h->layers.size() > MAX_SLABS) return false;
  for (const ResLayer& l : h->layers)
    if (l.G != G || l.wf.b_off >= 0 || l.proj.b_off >= 0) return false;
  return get_encode() != nullptr;
}

bool tc_head_supported(const wn_handle* h) {
  for (const ConvParam& c : h->head)
    if ((c.in_ch != 64 && c.in_ch != 128 && c.in_ch != 256) || c.out_ch % 32 != 0 || c.out_ch > 256) return false;
  return h->S % 32 == 0 && get_encode() != nullptr;
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
                                      h->ws + t.tc_ws, h->ws + t.tc_w1t, h->ws + t.tc_wpt, h->ws + t.tc_wst, L, h->R,
                                      h->layers[0].G, h->S, 2);
  WN_CHECK_LAUNCH();
  return WN_OK;
}

int tc_layer_launch(wn_handle* h, int l, cudaStream_t s);
int tc_skip_gemm(wn_handle* h, cudaStream_t s);

int tc_forward_residual(wn_handle* h, const float* params, cudaStream_t s) {
  const Tape& t = h->tape;
  const int L = (int)h->layers.size();
  WN_TRY(tc_prepare_weights(h, params, s));
  if (tc_fused_supported(h)) {
    static bool attr = false;
    if (!attr) {
      WN_CHECK_CUDA(cudaFuncSetAttribute(tc_layer_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, L_SMEM + 1024));
      attr = true;
    }
    for (int l = 0; l < L; ++l) WN_TRY(tc_layer_launch(h, l, s));
    h->tape_gates_zs = true;     // tape keeps (z, sigmoid)
  } else {
    // generic shapes: a = [x(t-d) | x(t)] . W1^T as a 2-slab GEMM (zero prefix in the epilogue), SIMT gate,
    // projection GEMM with the residual in the epilogue
    const int R = h->R, G = h->layers[0].G;
    for (int l = 0; l < L; ++l) {
      const ResLayer& ly = h->layers[l];
      TcOperand X{h->ws + t.x[l], R, t.W, t.B, 1, 0};
      const int sidx[2] = {0, 0};
      const int roff[2] = {-ly.dilation, 0};
      TcEpilogue e;
      e.zero_rows_below = wn_zero_prefix(t.W, ly.dilation, 2);
      WN_TRY(tc_gemm(h, X, 2, sidx, roff, t.W, h->ws + t.tc_w1 + (int64_t)l * 2 * G * 2 * R, 2 * G, e, h->ws + t.tfsg[l],
                     2 * G, s));
      WN_TRY(simt_gate_forward(h->ws + t.tfsg[l], h->ws + t.z[l], t.P, G, s));
      TcOperand Z{h->ws + t.z[l], G, t.W, t.B, 1, 0};
      const int zero = 0;
      TcEpilogue e2;
      e2.Rsd = h->ws + t.x[l];
      e2.ldr = R;
      WN_TRY(tc_gemm(h, Z, 1, nullptr, &zero, t.W, h->ws + t.tc_w2 + (int64_t)l * R * G, R, e2, h->ws + t.x[l + 1], R, s));
    }
    h->tape_gates_zs = false;    // tape keeps (tanh | sigmoid) like the SIMT path
  }
  // sum_skip = sum_l Ws_l z_l as one GEMM over K = L*G (z buffers are equally spaced slabs)
  return tc_skip_gemm(h, s);
}

int tc_layer_launch(wn_handle* h, int l, cudaStream_t s) {
  const Tape& t = h->tape;
  const int R = 64, G = 64;
  const int tiles_per_seq = (t.W + TM - 1) / TM;
  const int num_tiles = tiles_per_seq * t.B;
  const int grid = num_tiles < h->sm_count ? num_tiles : h->sm_count;
  {
    const ResLayer& ly = h->layers[l];
    CUtensorMap tx, tw1, tw2, tz, tsg;
    WN_TRY(make_map_4d(&tx, h->ws + t.x[l], R, t.W, t.B, 1, R, (uint64_t)t.W * R, (uint64_t)t.P * R, TM));
    WN_TRY(make_map_4d(&tz, h->ws + t.z[l], G, t.W, t.B, 1, G, (uint64_t)t.W * G, (uint64_t)t.P * G, TM));
    WN_TRY(make_map_4d_f16(&tsg, h->ws + t.tfsg[l], G, t.W, t.B, G, (uint64_t)t.W * G, TM));
    WN_TRY(make_map_2d(&tw1, h->ws + t.tc_w1 + (int64_t)l * 2 * G * 2 * R, 2 * R, 2 * G, 2 * R, 128));
    WN_TRY(make_map_2d(&tw2, h->ws + t.tc_w2 + (int64_t)l * R * G, G, R, G, 64));
    LayerArgs a;
    a.x_out = h->ws + t.x[l + 1];
    a.z_out = h->ws + t.z[l];
    a.sg_out = h->save_gates ? h->ws + t.tfsg[l] : nullptr;   // TC tapes keep sigmoid only, row stride G
    a.W = t.W;
    a.d = ly.dilation;
    a.zp = wn_zero_prefix(t.W, ly.dilation, 2);
    a.tiles_per_seq = tiles_per_seq;
    a.num_tiles = num_tiles;
    a.reverse = getenv("WN_NO_SERP") ? 0 : (l & 1);
    WN_CHECK_CUDA(launch_pdl(tc_layer_kernel, grid, L_THREADS, L_SMEM + 1024, s, tx, tw1, tw2, tz, tsg, a));
    WN_CHECK_LAUNCH();
  }
  return WN_OK;
}

int tc_skip_gemm(wn_handle* h, cudaStream_t s) {
  const Tape& t = h->tape;
  const int L = (int)h->layers.size();
  const int G = h->layers[0].G;
  const int64_t zstride = L > 1 ? t.z[1] - t.z[0] : 0;
  for (int l = 1; l < L; ++l)
    WN_REQUIRE(t.z[l] - t.z[l - 1] == zstride, WN_EINVAL, "z slabs are not equally spaced");
  WN_REQUIRE(L <= MAX_SLABS, WN_EINVAL, "too many layers for the skip GEMM");
  int idx[MAX_SLABS], off[MAX_SLABS];
  for (int l = 0; l < L; ++l) {
    idx[l] = l;
    off[l] = 0;
  }
  TcOperand A{h->ws + t.z[0], G, t.W, t.B, L, zstride};
  TcEpilogue e;
  e.relu = h->fuse_head_relu;        // fused train path: the head's first ReLU (wavenet.py:588) rides on this epilogue
  e.round_out = h->fuse_head_relu;
  h->skip_is_relu = h->fuse_head_relu != 0;
  return tc_gemm(h, A, L, idx, off, t.W, h->ws + t.tc_ws, h->S, e, h->ws + t.skip, h->S, s);
}

// ReLU -> 1x1 conv per head layer (wavenet.py:587-590) on tensor cores.  Stored activations are
// post-ReLU (ReLU is idempotent, so mask / a_relu logic in backward sees the same values).
int tc_forward_head(wn_handle* h, const float* params, int T, bool external, cudaStream_t s) {
  const Tape& t = h->tape;
  const int nh = (int)h->head.size();
  const int64_t rows = (int64_t)t.B * T;
  const int rows_in0 = external ? T : t.W, off0 = external ? 0 : t.W - T;
  if (!h->skip_is_relu || external) {
    const int64_t n = rows * (h->S / 4);
    tc_relu_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(h->ws + t.skip, h->S, rows_in0, off0, T, rows);
    WN_CHECK_LAUNCH();
  }
  for (int i = 0; i < nh; ++i) {
    const ConvParam& cp = h->head[i];
    const int64_t n = (int64_t)cp.out_ch * cp.in_ch;
    tc_round_copy_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(params + cp.w_off, h->ws + t.tc_wh[i], n);
    WN_CHECK_LAUNCH();
    const bool last = i == nh - 1;
    TcOperand A{i == 0 ? h->ws + t.skip : h->ws + t.hbuf[i - 1], cp.in_ch, i == 0 ? rows_in0 : T, t.B, 1, 0};
    const int off = i == 0 ? off0 : 0;
    TcEpilogue e;
    e.bias = cp.b_off >= 0 ? params + cp.b_off : nullptr;
    e.relu = last ? 0 : 1;
    e.round_out = last ? 0 : 1;
    WN_TRY(tc_gemm(h, A, 1, nullptr, &off, T, h->ws + t.tc_wh[i], cp.out_ch, e, h->ws + t.hbuf[i], cp.out_ch, s));
  }
  return WN_OK;
}

// Backward of head + residual stack on tensor cores (embedding scatter stays SIMT).
// Requires a tape written by tc_forward_residual(save_gates) and tc_forward_head.
int tc_backward(wn_handle* h, const float* params, float* grads, cudaStream_t s) {
  const Tape& t = h->tape;
  const int T = h->T, W = t.W, B = t.B;
  const int64_t P = t.P;
  const int nh = (int)h->head.size();
  const int L = (int)h->layers.size();
  const int R = h->R, G = h->layers[0].G, S = h->S;
  float* ws = h->ws;
  const int zero = 0;
  // ---- head ----
  const float* d = ws + t.dlogits;
  int tog = 0;
  bool bias_done = false;
  for (int i = nh - 1; i >= 0; --i) {
    const ConvParam& cp = h->head[i];
    const bool first = i == 0;
    const int rin = first ? (h->head_external ? T : W) : T;
    const int aoff = first ? (h->head_external ? 0 : W - T) : 0;
    const float* Ain = first ? ws + t.skip : ws + t.hbuf[i - 1];
    TcOperand dY{d, cp.out_ch, T, B, 1, 0};
    TcOperand X{Ain, cp.in_ch, rin, B, 1, 0};
    for (int m0 = 0; m0 < cp.out_ch; m0 += 256) {
      const int mv = cp.out_ch - m0 < 256 ? cp.out_ch - m0 : 256;
      float* dw = grads + cp.w_off + (int64_t)m0 * cp.in_ch;
      WN_TRY(tc_wgrad(h, dY, 0, m0, mv, X, 1, &aoff, nullptr, &dw, nullptr, 256, T, cp.in_ch, 1, s));
    }
    // bias gradient = column sums of d: taken from the CE kernel (last conv) or from the epilogue of the GEMM that
    // produced d (previous iteration) when available, else one more pass over d
    if (cp.b_off >= 0 && !bias_done) {
      if (i == nh - 1 && h->ce_colsum_valid && d == ws + t.dlogits)
        WN_TRY(simt_add_vec(ws + t.ce_colsum, grads + cp.b_off, cp.out_ch, s));
      else
        WN_TRY(simt_colsum(d, (int64_t)B * T, cp.out_ch, grads + cp.b_off, s));
    }
    bias_done = false;
    if (first && h->head_external) return WN_OK;
    // d_prev = (d . W) masked by the stored (post-ReLU) input
    const int64_t n = (int64_t)cp.out_ch * cp.in_ch;
    tc_transpose_round_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(params + cp.w_off, ws + t.tc_wht[i], cp.out_ch,
                                                                         cp.in_ch);
    WN_CHECK_LAUNCH();
    TcEpilogue e;
    e.mask = Ain;
    e.ldm = cp.in_ch;
    e.mask_rows_in = rin;
    e.mask_row_off = aoff;
    if (i > 0 && h->head[i - 1].b_off >= 0 && cp.in_ch <= 256) {
      e.colsum_out = grads + h->head[i - 1].b_off;   // d of the next iteration is this GEMM's output
      bias_done = true;
    }
    WN_TRY(tc_gemm(h, dY, 1, nullptr, &zero, T, ws + t.tc_wht[i], cp.in_ch, e, ws + t.dh[tog], cp.in_ch, s));
    d = ws + t.dh[tog];
    tog ^= 1;
  }
  const float* dskip = d;   // [B*T][S]
  TcOperand DS{dskip, S, T, B, 1, 0};
  const int wt = W - T, nwt = -(W - T);
  const int64_t zstride = L > 1 ? t.z[1] - t.z[0] : (int64_t)P * G;
  auto nb_ok = [](int nb) { return nb == 64 || nb == 128 || nb == 256; };
  // generic weight-gradient helper: dW(m, c) += dY[.., m]^T X[.., c] over M in 128-row halves; rows below `split`
  // go to dWa, the rest to dWb (wf | wg); falls back to the SIMT kernel when the tile shape is not MMA-friendly
  auto wgrad_any = [&](const TcOperand& dY, int a_row_off, int Mtot, int split, const TcOperand& X, int x_row_off,
                       int rows_it, float* dWa, float* dWb, int64_t sn, int64_t sk) -> int {
    if (!nb_ok(X.K)) {
      WgradArgs wg;
      memset(&wg, 0, sizeof(wg));
      for (int part = 0; part < (dWb ? 2 : 1); ++part) {
        const int m0 = part == 0 ? 0 : split, m1 = part == 0 ? (dWb ? split : Mtot) : Mtot;
        wg.dY = dY.ptr + m0;
        wg.ldd = dY.K;
        wg.N = m1 - m0;
        wg.dy_rows_out = rows_it;
        wg.dy_rows_in = dY.rows_in;
        wg.dy_in_off = a_row_off;
        wg.A = X.ptr;
        wg.lda = X.K;
        wg.K = X.K;
        wg.ntaps = 1;
        wg.shift[0] = 0;
        wg.M = (int64_t)dY.num_seq * rows_it;
        wg.rows_out = rows_it;
        wg.rows_in = X.rows_in;
        wg.in_off = x_row_off;
        wg.dW = part == 0 ? dWa : dWb;
        wg.sn = sn;
        wg.sk = sk;
        wg.st = 0;
        WN_TRY(simt_wgrad(wg, h->sm_count, s));
      }
      return WN_OK;
    }
    for (int m0 = 0; m0 < Mtot; m0 += 256) {
      const int mv = Mtot - m0 < 256 ? Mtot - m0 : 256;
      float* d0;
      float* d1 = nullptr;
      int msplit = 256;
      if (!dWb || m0 + mv <= split) {
        d0 = dWa + (int64_t)m0 * sn;
      } else if (m0 >= split) {
        d0 = dWb + (int64_t)(m0 - split) * sn;
      } else {
        d0 = dWa + (int64_t)m0 * sn;
        d1 = dWb;
        msplit = split - m0;
      }
      WN_TRY(tc_wgrad(h, dY, a_row_off, m0, mv, X, 1, &x_row_off, nullptr, &d0, d1 ? &d1 : nullptr, msplit, rows_it, sn, sk, s));
    }
    return WN_OK;
  };
  // ---- skip path for ALL layers at once (mirror of the forward skip GEMM) ----
  //  dzs[l] = dskip . Ws_l   : GEMMs with N = nl layers x G, each column block written to its layer slab
  //  dWs_l  = dskip^T . z_l  : wgrad with the B operand gathered from nl z slabs per launch
  // all L slabs of dzs from one launch: 256-column groups of [Ws_0 .. Ws_L-1] interleaved per row tile, so dskip
  // comes from HBM once.  (WN_DZS_ARES=1 selects the A-resident kernel, WN_DZS_SPLIT=1 one launch per group.)
  const bool ares = S <= 256 && getenv("WN_DZS_ARES") != nullptr;
  const bool grouped = !ares && getenv("WN_DZS_SPLIT") == nullptr && 256 % G == 0;
  if (ares) {
    WN_TRY(tc_gemm_ares(h, DS, nwt, W, ws + t.tc_wst, L * G, ws + t.dzs, G, G, (int64_t)P * G, s));
  } else if (grouped) {
    TcEpilogue e;
    e.y_slab_cols = G;
    e.y_slab_stride = (int64_t)P * G;
    WN_TRY(tc_gemm(h, DS, 1, nullptr, &nwt, W, ws + t.tc_wst, L * G, e, ws + t.dzs, G, s));
  }
  int nl_max = 256 / G;
  if (nl_max > 4) nl_max = 4;
  if (const char* ev = getenv("WN_DZS_NL")) nl_max = atoi(ev) > 0 && atoi(ev) < nl_max ? atoi(ev) : nl_max;
  // dWs for all layers from one launch when the layers fit 8 groups of nl_max slabs (dskip read from HBM once)
  const int ngr = (L + nl_max - 1) / nl_max;
  const bool dws_grouped = nb_ok(nl_max * G) && nl_max > 1 && ngr > 1 && ngr <= 8 && S <= 256 && (grouped || ares) &&
                           getenv("WN_DWS_SPLIT") == nullptr;
  if (dws_grouped) {
    TcOperand Z{ws + t.z[0], G, W, B, L, zstride};
    int boff[4], bidx[32];
    float* dws[32];
    for (int j = 0; j < nl_max; ++j) boff[j] = wt;
    for (int gr = 0; gr < ngr; ++gr)
      for (int j = 0; j < nl_max; ++j) {
        const int l = gr * nl_max + j;
        bidx[gr * nl_max + j] = l < L ? l : -1;
        dws[gr * nl_max + j] = l < L ? grads + h->layers[l].skip.w_off : nullptr;
      }
    WN_TRY(tc_wgrad(h, DS, 0, 0, S, Z, nl_max, boff, bidx, dws, nullptr, 256, T, G, 1, s, ngr));
  }
  for (int l0 = 0; l0 < L && !(dws_grouped && (grouped || ares));) {
    int nl = L - l0 < nl_max ? L - l0 : nl_max;
    while (nl > 1 && !nb_ok(nl * G)) --nl;
    if (!ares && !grouped) {
      TcEpilogue e;
      e.y_slab_cols = G;
      e.y_slab_stride = (int64_t)P * G;
      WN_TRY(tc_gemm(h, DS, 1, nullptr, &nwt, W, ws + t.tc_wst + (int64_t)l0 * G * S, nl * G, e,
                     ws + t.dzs + (int64_t)l0 * P * G, G, s));
    }
    if (nb_ok(nl * G)) {
      TcOperand Z{ws + t.z[0], G, W, B, L, zstride};
      int boff[4], bidx[4];
      float* dws[4];
      for (int m0 = 0; m0 < S; m0 += 256) {
        const int mv = S - m0 < 256 ? S - m0 : 256;
        for (int j = 0; j < nl; ++j) {
          boff[j] = wt;
          bidx[j] = l0 + j;
          dws[j] = grads + h->layers[l0 + j].skip.w_off + (int64_t)m0 * G;
        }
        WN_TRY(tc_wgrad(h, DS, 0, m0, mv, Z, nl, boff, bidx, dws, nullptr, 256, T, G, 1, s));
      }
    } else {
      for (int j = 0; j < nl; ++j) {
        TcOperand Z{ws + t.z[l0 + j], G, W, B, 1, 0};
        WN_TRY(wgrad_any(DS, 0, S, S, Z, wt, T, grads + h->layers[l0 + j].skip.w_off, nullptr, G, 1));
      }
    }
    l0 += nl;
  }
  // ---- residual layers ----
  const bool serp = getenv("WN_NO_SERP") == nullptr;
  int dt = 0;
  const float* dout = nullptr;
  for (int l = L - 1; l >= 0; --l) {
    const ResLayer& ly = h->layers[l];
    const int zp = wn_zero_prefix(W, ly.dilation, 2);
    // serpentine hand-off: consecutive kernels walk the rows in opposite directions, so each one starts on the part
    // of its input the previous kernel touched last (still in L2): gate D -> dW1 !D -> dx D -> next layer's gate !D
    // (with the fused dx + dW1 kernel there are two kernels per layer: gate forwards, dxw backwards)
    float* dwf = grads + ly.wf.w_off;
    float* dwg = grads + ly.wg.w_off;
    const bool fuse_dxw = R == 64 && G == 64 && h->cfg.residual_filter_width == 2 &&
                          (((uintptr_t)dwf | (uintptr_t)dwg) & 15) == 0 && getenv("WN_NO_DXW_FUSE") == nullptr;
    const int dir = serp ? (fuse_dxw ? 0 : (l & 1)) : 0, ndir = serp ? !dir : 0;
    const float* dzs = ws + t.dzs + (int64_t)l * P * G;
    const float* sg = h->tape_gates_zs ? ws + t.tfsg[l] : ws + t.tfsg[l] + G;
    const int sg_ld = h->tape_gates_zs ? G : 2 * G;
    const int sg_half = h->tape_gates_zs ? 1 : 0;   // the fused layer kernel stores sigmoid as fp16
    TcOperand Z{ws + t.z[l], G, W, B, 1, 0};
    if (dout) {
      // dz = dout . Wp + dzs_l, gate derivative fused into the epilogue -> dafg
      TcOperand DO{dout, R, W, B, 1, 0};
      TcEpilogue e;
      e.Rsd = dzs;
      e.ldr = G;
      e.gate_sg = sg;
      e.gate_sg_ld = sg_ld;
      e.gate_sg_half = sg_half;
      e.gate_z = ws + t.z[l];
      e.gate_dafg = ws + t.dafg;
      e.gate_zp = zp;
      e.reverse = dir;
      float* dwp = grads + ly.proj.w_off;
      if (R == 64 && G == 64 && ((uintptr_t)dwp & 15) == 0 && getenv("WN_NO_GATE_FUSE") == nullptr) {
        WN_TRY(tc_gate_bwd(h, dout, ws + t.tc_wpt + (int64_t)l * G * R, dzs, ws + t.z[l], sg, sg_ld, sg_half, ws + t.dafg, dwp, zp, W, B, dir, s));
      } else {
        WN_TRY(tc_gemm(h, DO, 1, nullptr, &zero, W, ws + t.tc_wpt + (int64_t)l * G * R, G, e, ws + t.dz, G, s));
        WN_TRY(wgrad_any(DO, 0, R, R, Z, 0, W, grads + ly.proj.w_off, nullptr, G, 1));
      }
    } else if (h->tape_gates_zs) {
      WN_TRY(simt_gate_backward_zs(ws + t.z[l], ws + t.tfsg[l], 1, dzs, ws + t.dafg, P, W, G, zp, s));
    } else {
      WN_TRY(simt_gate_backward(ws + t.tfsg[l], dzs, ws + t.dafg, P, W, G, zp, s));
    }
    TcOperand DA{ws + t.dafg, 2 * G, W, B, 1, 0};
    if (fuse_dxw) {
      float* dnew = ws + t.dout[dt];
      WN_TRY(tc_dxw(h, ws + t.dafg, ws + t.tc_w1t + (int64_t)l * R * 4 * G, dout, ws + t.x[l], dnew, dwf, dwg, ly.dilation, W,
                    B, ndir, s));
      dout = dnew;
      dt ^= 1;
      continue;
    }
    {
      // dW_{f,g}(o, c, tap) += da[t][o] * x[t - (1-tap) d][c]; taps share a launch while 2R fits one N tile
      TcOperand X{ws + t.x[l], R, W, B, 1, 0};
      if (nb_ok(2 * R)) {
        const int boff[2] = {-ly.dilation, 0};
        for (int m0 = 0; m0 < 2 * G; m0 += 256) {
          const int mv = 2 * G - m0 < 256 ? 2 * G - m0 : 256;
          float* d0[2];
          float* d1[2] = {nullptr, nullptr};
          int msplit = 256;
          for (int tap = 0; tap < 2; ++tap) {
            if (m0 + mv <= G) {
              d0[tap] = grads + ly.wf.w_off + (int64_t)m0 * 2 * R + tap;
            } else if (m0 >= G) {
              d0[tap] = grads + ly.wg.w_off + (int64_t)(m0 - G) * 2 * R + tap;
            } else {
              d0[tap] = grads + ly.wf.w_off + (int64_t)m0 * 2 * R + tap;
              d1[tap] = grads + ly.wg.w_off + tap;
              msplit = G - m0;
            }
          }
          WN_TRY(tc_wgrad(h, DA, 0, m0, mv, X, 2, boff, nullptr, d0, d1[0] ? d1 : nullptr, msplit, W, 2 * R, 2, s, 1, ndir));
        }
      } else {
        for (int tap = 0; tap < 2; ++tap)
          WN_TRY(wgrad_any(DA, 0, 2 * G, G, X, tap == 0 ? -ly.dilation : 0, W, grads + ly.wf.w_off + tap,
                           grads + ly.wg.w_off + tap, 2 * R, 2));
      }
    }
    {
      float* dnew = ws + t.dout[dt];
      const int roff[2] = {0, ly.dilation};
      const int sidx[2] = {0, 0};
      TcEpilogue e;
      e.Rsd = dout;
      e.ldr = R;
      e.reverse = dir;
      WN_TRY(tc_gemm(h, DA, 2, sidx, roff, W, ws + t.tc_w1t + (int64_t)l * R * 4 * G, R, e, dnew, R, s));
      dout = dnew;
      dt ^= 1;
    }
  }
  h->bwd_dout = dout;
  return WN_OK;
}

// Profiling hooks (bench.py roofline): one launch of the fused residual-layer kernel / of the skip GEMM on
// the bound tape.  Require a previous wn_forward_residual_block under WN_PREC_TF32 (weights prepared).
extern "C" int wn_tc_layer_forward(wn_handle* h, int layer, void* stream) {
  WN_REQUIRE(h && h->ws && h->tape_tc, WN_ESTATE, "wn_tc_layer_forward: run a TF32 forward first");
  WN_REQUIRE(layer >= 0 && layer < (int)h->layers.size(), WN_EINVAL, "bad layer index");
  return tc_layer_launch(h, layer, (cudaStream_t)stream);
}
extern "C" int wn_tc_gate_backward_layer(wn_handle* h, int layer, float* grads, void* stream) {
  WN_REQUIRE(h && h->ws && h->tape_tc && h->tape_gates_zs && grads, WN_ESTATE,
             "wn_tc_gate_backward_layer: run a TF32 forward + backward of the fused shape first");
  WN_REQUIRE(layer >= 0 && layer + 1 < (int)h->layers.size() && h->R == 64 && h->layers[layer].G == 64, WN_EINVAL,
             "bad layer index / shape");
  const Tape& t = h->tape;
  const ResLayer& ly = h->layers[layer];
  float* dwp = grads + ly.proj.w_off;
  WN_REQUIRE(((uintptr_t)dwp & 15) == 0, WN_EINVAL, "grads_scratch must be 16-byte aligned");
  return tc_gate_bwd(h, h->ws + t.dout[0], h->ws + t.tc_wpt + (int64_t)layer * 64 * 64, h->ws + t.dzs + (int64_t)layer * t.P * 64,
                     h->ws + t.z[layer], h->ws + t.tfsg[layer], 64, 1, h->ws + t.dafg, dwp, wn_zero_prefix(t.W, ly.dilation, 2),
                     t.W, t.B, 0, (cudaStream_t)stream);
}
#ifdef WN_LAYER_TRACE
extern "C" int wn_debug_layer_trace(long long* out) {
  return cudaMemcpyFromSymbol(out, g_trace, sizeof(long long) * 64 * 32) == cudaSuccess ? 0 : -1;
}
#endif
extern "C" int wn_tc_skip_gemm(wn_handle* h, void* stream) {
  WN_REQUIRE(h && h->ws && h->tape_tc, WN_ESTATE, "wn_tc_skip_gemm: run a TF32 forward first");
  return tc_skip_gemm(h, (cudaStream_t)stream);
}
