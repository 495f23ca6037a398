// Exact-fp32 SIMT kernels: the parity path (WN_PREC_FP32) and every shape the
// tcgen05 path does not specialise.  All activations are channels-last fp32.
#include <cuda_fp16.h>
#include "wn_common.h"

namespace {

constexpr int BM = 64, BN = 64, BK = 16, NT = 256;

// Y[p, n] = epi( sum_tap sum_k A[row(p) - shift_tap, k] * w(n, tap, k) )
// One kernel covers: dilated causal convs (wavenet.py:294-342 in closed form), 1x1
// projections with residual add (wavenet.py:363-367), skip accumulation
// (wavenet.py:579), head convs with ReLU input (wavenet.py:587-590) and the
// data-gradient GEMMs of their backward passes.
__global__ void __launch_bounds__(NT) gemm_shift_kernel(GemmArgs g) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int tx = tid & 15, ty = tid >> 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int a_r = tid >> 2, a_k = (tid & 3) << 2;
  const int64_t p = m0 + a_r;
  const int64_t seq = p / g.rows_out;
  const int t_in = (int)(p - seq * g.rows_out) + g.in_off;
  const int b_n = tid >> 2, b_k = (tid & 3) << 2;
  const int n_ld = n0 + b_n;

  for (int tap = 0; tap < g.ntaps; ++tap) {
    const int ts = t_in - g.shift[tap];
    const bool rv = (p < g.M) && ts >= 0 && ts < g.rows_in;
    const float* arow = g.A + (seq * g.rows_in + ts) * (int64_t)g.lda;
    const float* wrow = g.Wt + (int64_t)n_ld * g.sn + (int64_t)tap * g.st;
    for (int k0 = 0; k0 < g.K; k0 += BK) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int k = k0 + a_k + i;
        float v = (rv && k < g.K) ? __ldg(arow + k) : 0.f;
        if (g.a_relu) v = fmaxf(v, 0.f);
        As[a_k + i][a_r] = v;
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int k = k0 + b_k + i;
        Bs[b_k + i][b_n] = (n_ld < g.N && k < g.K) ? __ldg(wrow + (int64_t)k * g.sk) : 0.f;
      }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < BK; ++kk) {
        float a[4], b[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
        for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
      __syncthreads();
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t row = m0 + ty * 4 + i;
    if (row >= g.M) continue;
    const int64_t seq_o = row / g.rows_out;
    const int t_out = (int)(row - seq_o * g.rows_out);
    const int64_t mrow = seq_o * g.mask_rows_in + t_out + g.mask_in_off;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= g.N) continue;
      float v = acc[i][j];
      if (g.bias) v += g.bias[n];
      if (t_out < g.zp) v = 0.f;
      if (g.Rsd) v += g.Rsd[row * g.ldr + n];
      if (g.mask_src && !(g.mask_src[mrow * g.ldm + n] > 0.f)) v = 0.f;
      float* y = g.Y + row * g.ldy + n;
      if (g.accumulate) v += *y;
      *y = v;
    }
  }
}

// dW(n, tap, k) += sum_p dY[rowY(p), n] * A[rowA(p) - shift_tap, k]   (+ dbias[n] += sum_p dY)
// grid: (ceil(N/64), ceil(K/64)*ntaps, splits over the row range)
__global__ void __launch_bounds__(NT) wgrad_kernel(WgradArgs g, int64_t rows_per_split) {
  __shared__ float Ds[BK][BN + 4];  // dY chunk  [row][n]
  __shared__ float Xs[BK][BM + 4];  // A chunk   [row][k]
  const int tid = threadIdx.x;
  const int n0 = blockIdx.x * BN;
  const int ktiles = (g.K + BM - 1) / BM;
  const int tap = blockIdx.y / ktiles;
  const int k0 = (blockIdx.y % ktiles) * BM;
  const int64_t r_begin = (int64_t)blockIdx.z * rows_per_split;
  const int64_t r_end = min(g.M, r_begin + rows_per_split);
  const int tx = tid & 15, ty = tid >> 4;  // ty -> n, tx -> k
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  float bsum[4] = {0.f, 0.f, 0.f, 0.f};
  const bool do_bias = g.dbias != nullptr && blockIdx.y == 0 && tx == 0;

  const int l_r = tid >> 4;          // 0..15 row in chunk
  const int l_c = (tid & 15) << 2;   // 0..60 column base
  for (int64_t r0 = r_begin; r0 < r_end; r0 += BK) {
    const int64_t p = r0 + l_r;
    const bool pv = p < r_end;
    {
      const int64_t seq = pv ? p / g.dy_rows_out : 0;
      const int t = pv ? (int)(p - seq * g.dy_rows_out) + g.dy_in_off : -1;
      const bool rv = pv && t >= 0 && t < g.dy_rows_in;
      const float* drow = g.dY + (seq * g.dy_rows_in + t) * (int64_t)g.ldd;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int n = n0 + l_c + i;
        Ds[l_r][l_c + i] = (rv && n < g.N) ? __ldg(drow + n) : 0.f;
      }
    }
    {
      const int64_t seq = pv ? p / g.rows_out : 0;
      const int t = pv ? (int)(p - seq * g.rows_out) + g.in_off - g.shift[tap] : -1;
      const bool rv = pv && t >= 0 && t < g.rows_in;
      const float* arow = g.A + (seq * g.rows_in + t) * (int64_t)g.lda;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int k = k0 + l_c + i;
        float v = (rv && k < g.K) ? __ldg(arow + k) : 0.f;
        if (g.a_relu) v = fmaxf(v, 0.f);
        Xs[l_r][l_c + i] = v;
      }
    }
    __syncthreads();
#pragma unroll
    for (int rr = 0; rr < BK; ++rr) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = Ds[rr][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Xs[rr][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        bsum[i] += a[i];
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int n = n0 + ty * 4 + i;
    if (n >= g.N) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = k0 + tx * 4 + j;
      if (k >= g.K) continue;
      atomicAdd(g.dW + (int64_t)n * g.sn + (int64_t)k * g.sk + (int64_t)tap * g.st, acc[i][j]);
    }
    if (do_bias) atomicAdd(g.dbias + n, bsum[i]);
  }
}

// emb[j][q][r] = Wc[r][q][j]   (first causal filter made gather-friendly)
__global__ void embed_prepare_kernel(const float* __restrict__ Wc, float* __restrict__ emb, int R, int Q, int kc) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= R * Q * kc) return;
  const int r = i % R, q = (i / R) % Q, j = i / (R * Q);
  emb[i] = Wc[((int64_t)r * Q + q) * kc + j];
}

// forward_causal_block on a one-hot input == gather-sum of filter columns
// (wavenet.py:565-570 with data.py:61-68 folded in).
__global__ void embed_forward_kernel(const float* __restrict__ emb, const float* __restrict__ bias,
                                     const int32_t* __restrict__ idx, float* __restrict__ out, int64_t P, int W, int R,
                                     int Q, int kc) {
  // one thread per (position, 4 channels) when R % 4 == 0 (float4 path), else per element; 32-bit index math
  const bool vec = (R & 3) == 0;
  const uint32_t rv = vec ? R >> 2 : R;
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (uint64_t)P * rv) return;
  const uint32_t p = (uint32_t)(i / rv);
  const uint32_t r = (uint32_t)(i - (uint64_t)p * rv) * (vec ? 4 : 1);
  const int t = (int)(p % (uint32_t)W);
  if (vec) {
    float4 v = bias ? *reinterpret_cast<const float4*>(bias + r) : make_float4(0.f, 0.f, 0.f, 0.f);
    for (int j = 0; j < kc; ++j) {
      const int s = kc - 1 - j;
      if (t - s >= 0) {
        const int q = idx[p - s];
        const float4 e = *reinterpret_cast<const float4*>(emb + ((int64_t)j * Q + q) * R + r);
        v.x += e.x, v.y += e.y, v.z += e.z, v.w += e.w;
      }
    }
    *reinterpret_cast<float4*>(out + (int64_t)p * R + r) = v;
  } else {
    float v = bias ? bias[r] : 0.f;
    for (int j = 0; j < kc; ++j) {
      const int s = kc - 1 - j;
      if (t - s >= 0) v += emb[((int64_t)j * Q + idx[p - s]) * R + r];
    }
    out[(int64_t)p * R + r] = v;
  }
}

__global__ void embed_backward_kernel(const float* __restrict__ dout, const int32_t* __restrict__ idx,
                                      float* __restrict__ demb, int64_t P, int W, int R, int Q, int kc) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P * R) return;
  const int r = (int)(i % R);
  const int64_t p = i / R;
  const int t = (int)(p % W);
  const float v = dout[i];
  for (int j = 0; j < kc; ++j) {
    const int s = kc - 1 - j;
    if (t - s >= 0) atomicAdd(demb + ((int64_t)j * Q + idx[p - s]) * R + r, v);
  }
}

__global__ void embed_unprepare_kernel(const float* __restrict__ demb, float* __restrict__ dWc, int R, int Q, int kc) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= R * Q * kc) return;
  const int r = i % R, q = (i / R) % Q, j = i / (R * Q);
  dWc[((int64_t)r * Q + q) * kc + j] += demb[i];
}

// Same scatter with a per-block copy of the whole table in shared memory (kc*Q*R floats <= 200 KB): shared-memory
// atomics absorb the 2*P*R updates, each block then adds its table to global memory once.
__global__ void __launch_bounds__(1024) embed_backward_smem_kernel(const float* __restrict__ dout,
                                                                   const int32_t* __restrict__ idx,
                                                                   float* __restrict__ demb, int64_t P, int W, int R, int Q,
                                                                   int kc) {
  extern __shared__ float tab[];
  const int n = kc * Q * R;
  for (int i = threadIdx.x; i < n; i += blockDim.x) tab[i] = 0.f;
  __syncthreads();
  const int rv = R >> 2;   // R % 4 == 0
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < P * rv; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p = i / rv;
    const int r = (int)(i - p * rv) * 4;
    const int t = (int)(p % W);
    const float4 v = *reinterpret_cast<const float4*>(dout + p * R + r);
    for (int j = 0; j < kc; ++j) {
      const int s = kc - 1 - j;
      if (t - s >= 0) {
        float* e = tab + (j * Q + idx[p - s]) * R + r;
        atomicAdd(e, v.x), atomicAdd(e + 1, v.y), atomicAdd(e + 2, v.z), atomicAdd(e + 3, v.w);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float v = tab[i];
    if (v != 0.f) atomicAdd(demb + i, v);
  }
}

__global__ void colsum_kernel(const float* __restrict__ a, int64_t rows, int C, float* __restrict__ out) {
  // grid.x over column blocks of 32, grid.y over row splits; block (32, 8)
  const int c = blockIdx.x * 32 + threadIdx.x;
  __shared__ float red[8][33];
  float s = 0.f;
  if (c < C)
    for (int64_t r = (int64_t)blockIdx.y * 8 + threadIdx.y; r < rows; r += (int64_t)gridDim.y * 8) s += a[r * C + c];
  red[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
    for (int i = 1; i < 8; ++i) s += red[i][threadIdx.x];
    atomicAdd(out + c, s);
  }
}

__device__ __forceinline__ float sigmoidf_exact(float x) { return 1.f / (1.f + expf(-x)); }

// z = tanh(a_f) * sigmoid(a_g)  (wavenet.py:360); afg holds [a_f | a_g] and is
// overwritten with [tanh | sigmoid] for backward.
__global__ void gate_forward_kernel(float* __restrict__ afg, float* __restrict__ z, int64_t P, int G) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P * G) return;
  const int64_t p = i / G;
  const int gch = (int)(i - p * G);
  float* row = afg + p * 2 * G;
  const float tf = tanhf(row[gch]);
  const float sg = sigmoidf_exact(row[G + gch]);
  row[gch] = tf;
  row[G + gch] = sg;
  z[i] = tf * sg;
}

__global__ void gate_backward_kernel(const float* __restrict__ tfsg, const float* __restrict__ dz,
                                     float* __restrict__ dafg, int64_t P, int W, int G, int zp) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P * G) return;
  const int64_t p = i / G;
  const int gch = (int)(i - p * G);
  const int t = (int)(p % W);
  const float tf = tfsg[p * 2 * G + gch];
  const float sg = tfsg[p * 2 * G + G + gch];
  const float d = dz[i];
  float df = d * sg * (1.f - tf * tf);
  float dg = d * tf * sg * (1.f - sg);
  if (t < zp) {  // a was a hard zero there (quirk Q1): no gradient
    df = 0.f;
    dg = 0.f;
  }
  dafg[p * 2 * G + gch] = df;
  dafg[p * 2 * G + G + gch] = dg;
}

// same derivative from (z, sigmoid) as stored by the tensor-core forward: tanh = z / sigmoid
__global__ void gate_backward_zs_kernel(const float* __restrict__ z, const float* __restrict__ sg, int sg_half,
                                        const float* __restrict__ dz, float* __restrict__ dafg, int64_t P, int W, int G,
                                        int zp) {
  // one thread per (position, 4 channels); G % 4 == 0 is guaranteed by the tensor-core layouts that store (z, sigmoid)
  const int gv = G >> 2;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P * gv) return;
  const int64_t p = i / gv;
  const int g4 = (int)(i - p * gv) * 4;
  const int t = (int)(p % W);
  const float4 zz = *reinterpret_cast<const float4*>(z + p * G + g4);
  float4 s;
  if (sg_half) {   // fp16 sigmoid tape written by the fused tensor-core layer kernel
    const uint2 raw = *reinterpret_cast<const uint2*>(reinterpret_cast<const __half*>(sg) + p * G + g4);
    const float2 lo = __half22float2(*reinterpret_cast<const __half2*>(&raw.x));
    const float2 hi = __half22float2(*reinterpret_cast<const __half2*>(&raw.y));
    s = make_float4(lo.x, lo.y, hi.x, hi.y);
  } else {
    s = *reinterpret_cast<const float4*>(sg + p * G + g4);
  }
  const float4 d = *reinterpret_cast<const float4*>(dz + p * G + g4);
  float4 df, dg;
  df.x = d.x * (s.x > 0.f ? s.x - zz.x * zz.x / s.x : 0.f), dg.x = d.x * zz.x * (1.f - s.x);
  df.y = d.y * (s.y > 0.f ? s.y - zz.y * zz.y / s.y : 0.f), dg.y = d.y * zz.y * (1.f - s.y);
  df.z = d.z * (s.z > 0.f ? s.z - zz.z * zz.z / s.z : 0.f), dg.z = d.z * zz.z * (1.f - s.z);
  df.w = d.w * (s.w > 0.f ? s.w - zz.w * zz.w / s.w : 0.f), dg.w = d.w * zz.w * (1.f - s.w);
  if (t < zp) df = dg = make_float4(0.f, 0.f, 0.f, 0.f);
  *reinterpret_cast<float4*>(dafg + p * 2 * G + g4) = df;
  *reinterpret_cast<float4*>(dafg + p * 2 * G + G + g4) = dg;
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// F.softmax over channels (wavenet.py:592): one warp per row.
__global__ void softmax_rows_kernel(const float* __restrict__ in, float* __restrict__ out, int64_t rows, int Q) {
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* x = in + row * Q;
  float m = -INFINITY;
  for (int q = lane; q < Q; q += 32) m = fmaxf(m, x[q]);
  m = warp_max(m);
  float s = 0.f;
  for (int q = lane; q < Q; q += 32) s += expf(x[q] - m);
  s = warp_sum(s);
  const float inv = 1.f / s;
  for (int q = lane; q < Q; q += 32) out[row * Q + q] = expf(x[q] - m) * inv;
}

// F.softmax_cross_entropy, mean over rows (wavenet.py:613-616) + its gradient.
__global__ void cross_entropy_kernel(const float* __restrict__ logits, const int32_t* __restrict__ target,
                                     int64_t rows, int Q, double* __restrict__ acc, float* __restrict__ dlogits) {
  const int warps = blockDim.x >> 5;
  const int64_t row = (int64_t)blockIdx.x * warps + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  __shared__ double part[32];
  double my = 0.0;
  if (row < rows) {
    const float* x = logits + row * Q;
    float m = -INFINITY;
    for (int q = lane; q < Q; q += 32) m = fmaxf(m, x[q]);
    m = warp_max(m);
    float s = 0.f;
    for (int q = lane; q < Q; q += 32) s += expf(x[q] - m);
    s = warp_sum(s);
    const int tg = target[row];
    const float inv = 1.f / s;
    const float invn = 1.f / (float)rows;
    for (int q = lane; q < Q; q += 32) {
      float d = expf(x[q] - m) * inv;
      if (q == tg) d -= 1.f;
      dlogits[row * Q + q] = d * invn;
    }
    if (lane == 0) my = (double)(m + logf(s)) - (double)x[tg];
  }
  if (lane == 0) part[threadIdx.x >> 5] = my;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < warps; ++i) t += part[i];
    atomicAdd(acc, t);
  }
}

__global__ void loss_finalize_kernel(const double* acc, int64_t rows, float* loss) { *loss = (float)(acc[0] / (double)rows); }

// inverse of data.onehot_pixel_image (data.py:61-68): (B,Q,1,W) -> (B,W)
__global__ void onehot_to_index_kernel(const float* __restrict__ oh, int B, int Q, int W, int32_t* __restrict__ idx) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)B * W) return;
  const int b = (int)(i / W), t = (int)(i % W);
  const float* col = oh + (int64_t)b * Q * W + t;
  int best = 0;
  float bv = col[0];
  for (int q = 1; q < Q; ++q) {
    const float v = col[(int64_t)q * W];
    if (v > bv) {
      bv = v;
      best = q;
    }
  }
  idx[i] = best;
}

inline unsigned blocks_for(int64_t n, int bs) { return (unsigned)((n + bs - 1) / bs); }

}  // namespace

int simt_gemm(const GemmArgs& g, cudaStream_t s) {
  if (g.M == 0 || g.N == 0) return WN_OK;
  dim3 grid(blocks_for(g.M, BM), blocks_for(g.N, BN));
  gemm_shift_kernel<<<grid, NT, 0, s>>>(g);
  WN_CHECK_LAUNCH();
  return WN_OK;
}

int simt_wgrad(const WgradArgs& g, int sm_count, cudaStream_t s) {
  if (g.M == 0 || g.N == 0 || g.K == 0) return WN_OK;
  const int ktiles = (g.K + BM - 1) / BM;
  const unsigned gx = blocks_for(g.N, BN), gy = ktiles * g.ntaps;
  int64_t splits = (4LL * sm_count + gx * gy - 1) / (gx * gy);
  const int64_t max_splits = (g.M + 4 * BK - 1) / (4 * BK);
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  if (splits > 65535) splits = 65535;
  int64_t rps = (g.M + splits - 1) / splits;
  rps = (rps + BK - 1) / BK * BK;
  splits = (g.M + rps - 1) / rps;
  dim3 grid(gx, gy, (unsigned)splits);
  wgrad_kernel<<<grid, NT, 0, s>>>(g, rps);
  WN_CHECK_LAUNCH();
  return WN_OK;
}

int simt_embed_prepare(const float* Wc, float* emb, int R, int Q, int kc, cudaStream_t s) {
  embed_prepare_kernel<<<blocks_for((int64_t)R * Q * kc, 256), 256, 0, s>>>(Wc, emb, R, Q, kc);
  WN_CHECK_LAUNCH();
  return WN_OK;
}

int simt_embed_forward(const float* emb, const float* bias, const int32_t* idx, float* out, int B, int W, int R, int Q,
                       int kc, cudaStream_t s) {
  const int64_t P = (int64_t)B * W;
  embed_forward_kernel<<<blocks_for((R & 3) == 0 ? P * (R / 4) : P * R, 256), 256, 0, s>>>(emb, bias, idx, out, P, W, R, Q, kc);
  WN_CHECK_LAUNCH();
  return WN_OK;
}

int simt_embed_backward(const float* dout, const int32_t* idx, float* demb, float* dWc, float* dbias, int B, int W,
                        int R, int Q, int kc, cudaStream_t s) {
  const int64_t P = (int64_t)B * W;
  WN_CHECK_CUDA(cudaMemsetAsync(demb, 0, sizeof(float) * R * Q * kc, s));
  const size_t tab_bytes = sizeof(float) * R * Q * kc;
  if (R % 4 == 0 && tab_bytes <= 200 * 1024 && P * R >= (int64_t)1 << 22) {
    static bool attr = false;
    if (!attr) {
      WN_CHECK_CUDA(cudaFuncSetAttribute(embed_backward_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      attr = true;
    }
    int dev = 0, sms = 0;
    WN_CHECK_CUDA(cudaGetDevice(&dev));
    WN_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    embed_backward_smem_kernel<<<sms, 1024, tab_bytes, s>>>(dout, idx, demb, P, W, R, Q, kc);
  } else {
    embed_backward_kernel<<<blocks_for(P * R, 256), 256, 0, s>>>(dout, idx, demb, P, W, R, Q, kc);
  }
  WN_CHECK_LAUNCH();
  embed_unprepare_kernel<<<blocks_for((int64_t)R * Q * kc, 256), 256, 0, s>>>(demb, dWc, R, Q, kc);
  WN_CHECK_LAUNCH();
  if (dbias) {
    dim3 grid(blocks_for(R, 32), 64), block(32, 8);
    colsum_kernel<<<grid, block, 0, s>>>(dout, P, R, dbias);
    WN_CHECK_LAUNCH();
  }
  return WN_OK;
}

int simt_colsum(const float* a, int64_t rows, int C, float* out, cudaStream_t s) {
  dim3 grid(blocks_for(C, 32), 64), block(32, 8);
  colsum_kernel<<<grid, block, 0, s>>>(a, rows, C, out);
  WN_CHECK_LAUNCH();
  return WN_OK;
}

int simt_gate_forward(float* afg, float* z, int64_t P, int G, cudaStream_t s) {
  gate_forward_kernel<<<blocks_for(P * G, 256), 256, 0, s>>>(afg, z, P, G);
  WN_CHECK_LAUNCH();
  return WN_OK;
}

int simt_gate_backward(const float* tfsg, const float* dz, float* dafg, int64_t P, int W, int G, int zp,
                       cudaStream_t s) {
  gate_backward_kernel<<<blocks_for(P * G, 256), 256, 0, s>>>(tfsg, dz, dafg, P, W, G, zp);
  WN_CHECK_LAUNCH();
  return WN_OK;
}

int simt_gate_backward_zs(const float* z, const float* sg, int sg_half, const float* dz, float* dafg, int64_t P, int W, int G, int zp,
                          cudaStream_t s) {
  WN_REQUIRE(G % 4 == 0, WN_EINVAL, "gate_backward_zs: G must be a multiple of 4");
  gate_backward_zs_kernel<<<blocks_for(P * (G / 4), 256), 256, 0, s>>>(z, sg, sg_half, dz, dafg, P, W, G, zp);
  WN_CHECK_LAUNCH();
  return WN_OK;
}

int simt_softmax_rows(const float* in, float* out, int64_t rows, int Q, cudaStream_t s) {
  softmax_rows_kernel<<<blocks_for(rows, 8), 256, 0, s>>>(in, out, rows, Q);
  WN_CHECK_LAUNCH();
  return WN_OK;
}

// Q == 256 fast path: one warp per row, the row stays in registers (one read of the logits), grid-stride over rows;
// also accumulates the column sums of dlogits (= bias gradient of the last head conv, wavenet.py:577-586 backward)
// so the backward pass does not have to read dlogits again for them.
__global__ void __launch_bounds__(256) cross_entropy256_kernel(const float* __restrict__ logits,
                                                               const int32_t* __restrict__ target, int64_t rows,
                                                               double* __restrict__ acc, float* __restrict__ dlogits,
                                                               float* __restrict__ colsum, float split_scale) {
  // split_scale > 0: dlogits leave as split fp16 rows [hi 256 | lo 256] scaled by split_scale (the fp16x2 backward's operand
  // format, wn_tcs.cu) -- the same 1 KB per row as the fp32 row they replace
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  __shared__ double part[8];
  __shared__ float cs[8][256];
  double my = 0.0;
  float c[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const float invn = 1.f / (float)rows;
  for (int64_t row = (int64_t)blockIdx.x * 8 + wid; row < rows; row += (int64_t)gridDim.x * 8) {
    const float4* x = reinterpret_cast<const float4*>(logits + row * 256);
    const float4 a = x[lane], b = x[32 + lane];
    float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    float m = v[0];
#pragma unroll
    for (int i = 1; i < 8; ++i) m = fmaxf(m, v[i]);
    m = warp_max(m);
    float e[8], sum = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      e[i] = expf(v[i] - m);
      sum += e[i];
    }
    sum = warp_sum(sum);
    const int tg = target[row];
    const float inv = 1.f / sum;
    float xt = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int q = (i < 4 ? 0 : 128) + lane * 4 + (i & 3);
      float d = e[i] * inv;
      if (q == tg) {
        d -= 1.f;
        xt = v[i];
      }
      e[i] = d * invn;
      c[i] += e[i];
    }
    xt = warp_sum(xt);
    if (split_scale > 0.f) {
      __half* hrow = reinterpret_cast<__half*>(dlogits) + row * 512;
#pragma unroll
      for (int part = 0; part < 2; ++part) {
        uint2 hi, lo;
        const float v0 = e[4 * part] * split_scale, v1 = e[4 * part + 1] * split_scale, v2 = e[4 * part + 2] * split_scale,
                    v3 = e[4 * part + 3] * split_scale;
        const __half2 h01 = __floats2half2_rn(v0, v1), h23 = __floats2half2_rn(v2, v3);
        const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
        const __half2 l01 = __floats2half2_rn(v0 - f01.x, v1 - f01.y), l23 = __floats2half2_rn(v2 - f23.x, v3 - f23.y);
        hi.x = *reinterpret_cast<const uint32_t*>(&h01), hi.y = *reinterpret_cast<const uint32_t*>(&h23);
        lo.x = *reinterpret_cast<const uint32_t*>(&l01), lo.y = *reinterpret_cast<const uint32_t*>(&l23);
        *reinterpret_cast<uint2*>(hrow + part * 128 + lane * 4) = hi;
        *reinterpret_cast<uint2*>(hrow + 256 + part * 128 + lane * 4) = lo;
      }
    } else {
      float4* o = reinterpret_cast<float4*>(dlogits + row * 256);
      o[lane] = make_float4(e[0], e[1], e[2], e[3]);
      o[32 + lane] = make_float4(e[4], e[5], e[6], e[7]);
    }
    if (lane == 0) my += (double)(m + logf(sum)) - (double)xt;
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) cs[wid][(i < 4 ? 0 : 128) + lane * 4 + (i & 3)] = c[i];
  if (lane == 0) part[wid] = my;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) t += cs[w][threadIdx.x];
  atomicAdd(colsum + threadIdx.x, t);
  if (threadIdx.x == 0) {
    double tt = 0.0;
    for (int i = 0; i < 8; ++i) tt += part[i];
    atomicAdd(acc, tt);
  }
}

__global__ void add_vec_kernel(const float* __restrict__ src, float* __restrict__ dst, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] += src[i];
}

int simt_add_vec(const float* src, float* dst, int n, cudaStream_t s) {
  add_vec_kernel<<<blocks_for(n, 256), 256, 0, s>>>(src, dst, n);
  WN_CHECK_LAUNCH();
  return WN_OK;
}

int simt_cross_entropy(const float* logits, const int32_t* target, int64_t rows, int Q, double* acc, float* loss,
                       float* dlogits, float* colsum, bool* colsum_written, int sm_count, cudaStream_t s, float split_scale,
                       bool* split_written) {
  WN_CHECK_CUDA(cudaMemsetAsync(acc, 0, 2 * sizeof(double), s));
  const bool fast = Q == 256 && colsum && (((uintptr_t)logits | (uintptr_t)dlogits) & 15) == 0;
  if (colsum_written) *colsum_written = fast;
  if (split_written) *split_written = fast && split_scale > 0.f;   // the generic kernel always writes fp32 rows
  if (fast) {
    WN_CHECK_CUDA(cudaMemsetAsync(colsum, 0, sizeof(float) * 256, s));
    const int64_t want = blocks_for(rows, 8);
    const int grid = (int)(want < (int64_t)sm_count * 8 ? want : (int64_t)sm_count * 8);
    cross_entropy256_kernel<<<grid, 256, 0, s>>>(logits, target, rows, acc, dlogits, colsum, split_scale);
  } else {
    cross_entropy_kernel<<<blocks_for(rows, 8), 256, 0, s>>>(logits, target, rows, Q, acc, dlogits);
  }
  WN_CHECK_LAUNCH();
  loss_finalize_kernel<<<1, 1, 0, s>>>(acc, rows, loss);
  WN_CHECK_LAUNCH();
  return WN_OK;
}

int simt_loss_finalize(const double* acc, int64_t rows, float* loss, cudaStream_t s) {
  loss_finalize_kernel<<<1, 1, 0, s>>>(acc, rows, loss);
  WN_CHECK_LAUNCH();
  return WN_OK;
}

int simt_onehot_to_index(const float* onehot, int B, int Q, int W, int32_t* idx, cudaStream_t s) {
  onehot_to_index_kernel<<<blocks_for((int64_t)B * W, 256), 256, 0, s>>>(onehot, B, Q, W, idx);
  WN_CHECK_LAUNCH();
  return WN_OK;
}
