// Split-fp16 tensor-core path (WN_PREC_F16X2): fp32-grade arithmetic on tcgen05.
//
// The reference computes in fp32 end to end (Chainer fp32, wavenet.py:515-519).  A single tensor-core pass in
// tf32/fp16 keeps 11 significand bits per operand, which misses the 1e-4 logit / 1e-3 gradient gates.  Here every
// MMA operand is held as TWO fp16 planes, v = hi + lo with hi = fp16(v), lo = fp16(v - hi) (22 significand bits),
// and every product is three kind::f16 MMAs into the same TMEM accumulator:
//     A.B  ~=  A_hi.B_hi + A_hi.B_lo + A_lo.B_hi          (the dropped A_lo.B_lo term is 2^-24 relative)
// A "split row" [hi(C) | lo(C)] is 4C bytes -- exactly the bytes of the fp32 row it replaces -- so HBM traffic,
// shared-memory footprint and the TMA tiling equal the tf32 path's; only the MMA count is 1.5x (kind::f16 runs at
// twice the tf32 rate).  Because both planes are plain fp16 [rows x 128 B] tiles, ONE shared-memory image serves as the
// K-major operand of a data GEMM and as the MN-major operand of the weight-gradient GEMM.
// Power-of-two scales keep both planes inside the fp16 NORMAL range (a subnormal lo plane would cap the precision at
// 2^-25 absolute instead of 2^-24 relative): stored activations carry ACT_SCALE, prepared weights W_SCALE, gradient
// tensors h->gscale (chosen from B*T by wn_cross_entropy).  Every epilogue / reduction multiplies the accumulator by the
// reciprocal product (exact), so the scales never show outside this file.
//
// Backward: the residual-gradient stream dout and every forward-tape operand (x, z, weights) stay split; the gradient
// tensors that are only ever the dY operand of a product (dlogits, dh, dskip, dzs, dafg) are ONE fp16 plane rounded to
// nearest (two MMAs per product) and the sigmoid tape is 16-bit fixed point -- together 6.5e-4 of the gradient norm against
// the 1e-3 gate (tests/dev/quant_sensitivity.py).  Where the hi and lo planes of an operand are adjacent in shared memory
// they are consumed by ONE MMA of twice the N (or M): small-N MMAs are bound by shared-memory operand reads, not by the
// tensor pipe.
//
//  tcs_layer_kernel : one residual layer (wavenet.py:358-368, dilated conv of :294-342 in closed form) per launch
//  tcs_gate_bwd_kernel / tcs_dxw_kernel : the two fused backward kernels of a residual layer
//  tcs_gemm_kernel  : Y = epi(sum_slab A_slab . W^T)   (skip sum, head, data gradients, gate derivative epilogue)
//  tcs_wgrad_kernel : dW = dY^T . X                    (all weight gradients)
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdlib.h>
#include <string.h>

#include "wn_common.h"
#include "wn_tc.cuh"

namespace {

using namespace tc;

constexpr int TM = 128;          // positions per tile (UMMA M)
constexpr int KB = 64;           // fp16 elements per 128-byte swizzle row
constexpr int SUB = TM * 128;    // bytes of a [128 x 64] fp16 sub-tile
constexpr int NTHREADS = 64 + 256;   // producer warp, MMA warp, 8 epilogue warps
constexpr int MAX_SLABS = 32;
constexpr float ACT_SCALE = 8.f;       // x, z, skip, head activations: |v| up to 8188 representable
constexpr float W_SCALE = 16.f;        // weights: |w| up to 4094 representable; lo planes are normal down to |w| = 0.016
constexpr float INV_ACT = 1.f / ACT_SCALE, INV_W = 1.f / W_SCALE, INV_ACT_W = 1.f / (ACT_SCALE * W_SCALE);

// Instruction descriptor for kind::f16 with fp16 operands, fp32 accumulate (cute InstrDescriptor: c_format [4,6) = 1,
// a_format [7,10) = 0 (F16), b_format [10,13) = 0, a_major bit 15, b_major bit 16, n>>3 [17,23), m>>4 [24,29)).
__host__ __device__ constexpr uint32_t idesc_f16(int M, int N) {
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
constexpr uint32_t IDESC_MN_MAJOR = (1u << 15) | (1u << 16);

__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// Deterministic mode (wn_set_deterministic): every CTA adds its weight-gradient contribution into ITS OWN copy ("slab") of the
// flat gradient buffer with plain stores -- each (CTA, element) pair is written at most once per backward pass -- and one
// kernel then sums the slabs in slab order.  `det_stride` (floats between slabs, 0 = atomic mode) travels in the kernel
// arguments; the destination pointers are rebased to slab 0 by the host.
__device__ __forceinline__ void wg_add_v4(float* p, const float4& v, bool det) {
  if (det)
    *reinterpret_cast<float4*>(p) = v;
  else
    red_add_v4(p, v);
}

// MN-major fp16 operand, SWIZZLE_128B (cute canonical layout ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units): 64
// contiguous M/N elements per 128-byte row, 8 K rows per 1024-byte group; LBO = bytes between 64-element M/N atoms,
// SBO = bytes between 8-row K groups.
__device__ __forceinline__ uint64_t desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(lbo_bytes >> 4) << 16;
  d |= (uint64_t)(sbo_bytes >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// One mbarrier arrival per WARP: hundreds of threads arriving on the same shared-memory word serialise (an arrive is an
// atomic); bar.warp.sync orders every lane's earlier shared-memory / TMEM accesses before lane 0's releasing arrive.
__device__ __forceinline__ void warp_arrive(uint32_t bar, int lane) {
  __syncwarp();
  if (lane == 0) mbar_arrive(bar);
}

// ---- hi/lo split of fp32 values ---------------------------------------------------------------
__device__ __forceinline__ __half2 bits_h2(uint32_t u) { return *reinterpret_cast<__half2*>(&u); }

// {a, b} -> packed fp16x2 (a in the low half), round to nearest, saturating to +-65504 instead of inf
__device__ __forceinline__ uint32_t pack_h2_sat(float a, float b) {
  uint32_t d;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(b), "f"(a));
  return d;
}
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
  hi = pack_h2_sat(a, b);
  const float2 f = __half22float2(bits_h2(hi));
  lo = pack_h2_sat(a - f.x, b - f.y);
}
__device__ __forceinline__ void split4(const float4& v, uint2& hi, uint2& lo) {
  split2(v.x, v.y, hi.x, lo.x);
  split2(v.z, v.w, hi.y, lo.y);
}
__device__ __forceinline__ float4 join4(const uint2& hi, const uint2& lo) {
  const float2 h0 = __half22float2(bits_h2(hi.x)), h1 = __half22float2(bits_h2(hi.y));
  const float2 l0 = __half22float2(bits_h2(lo.x)), l1 = __half22float2(bits_h2(lo.y));
  return make_float4(h0.x + l0.x, h0.y + l0.y, h1.x + l1.x, h1.y + l1.y);
}
// value(row, col..col+3) of a split tensor whose rows are [hi(C) | lo(C)]
__device__ __forceinline__ float4 load_split4(const __half* base, int64_t row, int C, int col) {
  const __half* p = base + row * (2 * (int64_t)C) + col;
  return join4(*reinterpret_cast<const uint2*>(p), *reinterpret_cast<const uint2*>(p + C));
}
__device__ __forceinline__ float4 scale4(const float4& v, float s) { return make_float4(v.x * s, v.y * s, v.z * s, v.w * s); }
__device__ __forceinline__ void store_split4(__half* base, int64_t row, int C, int col, const float4& v) {
  uint2 hi, lo;
  split4(v, hi, lo);
  __half* p = base + row * (2 * (int64_t)C) + col;
  *reinterpret_cast<uint2*>(p) = hi;
  *reinterpret_cast<uint2*>(p + C) = lo;
}

// ---- 256-bit global accesses (sm_100): one full 32-byte sector per lane and instruction -------------------
__device__ __forceinline__ void ld256(const void* p, uint32_t (&r)[8]) {
  asm volatile("ld.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "l"(p));
}
__device__ __forceinline__ void st256(void* p, const uint32_t (&r)[8]) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]),
               "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
// 16 consecutive channels of a split row (hi plane at p, lo plane at p + C) -> fp32
__device__ __forceinline__ void load_split16(const __half* p, int C, float (&o)[16]) {
  uint32_t h[8], l[8];
  ld256(p, h);
  ld256(p + C, l);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float2 a = __half22float2(bits_h2(h[i])), b = __half22float2(bits_h2(l[i]));
    o[2 * i] = a.x + b.x;
    o[2 * i + 1] = a.y + b.y;
  }
}
__device__ __forceinline__ void store_split16(__half* p, int C, const float (&o)[16]) {
  uint32_t h[8], l[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) split2(o[2 * i], o[2 * i + 1], h[i], l[i]);
  st256(p, h);
  st256(p + C, l);
}

// ------------------------------------------------------------------------------------------
// Y[(b,t)][n] = epi( sum_slab  A[slab_idx][b][t + slab_row_off][:] . W[n][slab*K ...] ), both operands split.
// The ring holds plane-stages [A_p 128x64 | B_p BNx64]; K block kb uses two consecutive stages (hi, lo).
struct SGemmArgs {
  void* Y;                 // fp32 [rows][ldy], split [rows][hi(ldy) | lo(ldy)] or (out_half) one fp16 plane [rows][ldy]
  int ldy, out_split, out_half;
  const float* bias;       // [N] or null
  int N;                   // valid output columns
  int relu;
  float acc_scale;         // multiplies the accumulator (undoes the operand scales)
  float out_scale;         // multiplies the final result before it is stored (ACT_SCALE for stored activations)
  float rsd_scale;         // multiplies the residual after loading
  const void* Rsd;         // residual added to the result (same rows as Y): fp32 [rows][ldr] or split [rows][2*ldr]
  int ldr, rsd_split;
  const __half* mask;      // split tensor: result zeroed where mask[(b, mask_row_off + t)][n] <= 0, or null
  int ldm, mask_rows_in, mask_row_off;
  int rows_out;
  int nslab, kblk;         // K = nslab * kblk * 64
  int a_planes;            // 2: A rows are split [hi | lo]; 1: ONE fp16 plane (gradient tensors), two MMAs per product
  int a_plane;             // elements between the hi and lo planes of an A row
  int b_plane;             // elements between the planes of a W row (= total K)
  int slab_row_off[MAX_SLABS];
  int slab_idx[MAX_SLABS];
  int tiles_per_seq, num_tiles, ngroups;
  int y_slab_cols;         // >0 (fp32 / fp16 output): column block c goes to Y + (c / y_slab_cols) * y_slab_stride elements
  int64_t y_slab_stride;
  // gate-backward epilogue (MODE 1, N == G): acc + Rsd is dz; writes da_f | da_g into the split tensor gate_dafg
  const float* gate_sg;    // [rows][gate_sg_ld] fp32 sigmoid
  const __half* gate_z;    // split [rows][2N]
  __half* gate_dafg;       // split [rows][2 * 2N]
  int gate_zp, gate_sg_ld;
  int zero_rows_below;
  int reverse;
  int flush;               // MODE 4 (forward GEMMs; no mask / y_slab / colsum): > 0 = K blocks per flush group -- every group
                           // accumulates from zero in TMEM and is ADDED TO REGISTERS by the epilogue (fp32, round to nearest)
  float* colsum_out;       // MODE 5: += colsum_scale * column sums of the stored result
  float colsum_scale;
  int64_t det_stride;      // deterministic mode: per-warp column-sum tables, fixed-order sum, plain store into the CTA's slab
  // MODE 6 (last head conv fused with softmax cross-entropy; N == 256): Y (optional) receives the fp32 logits
  const int32_t* ce_target;   // [rows]
  double* ce_acc;             // += sum over rows of (logsumexp - logit[target])
  __half* ce_dlogits;         // ONE fp16 plane [rows][256] = (softmax - onehot) * ce_invn * ce_gscale
  float ce_invn, ce_gscale;
};

template <int BN>
struct SGemmCfg {
  static constexpr int STAGE = SUB + BN * 128;
  static constexpr int STAGES = BN == 256 ? 4 : 6;
  static constexpr int BAR = STAGES * STAGE;
  static constexpr int STG = BAR + 256;
  static constexpr int SMEM = STG + 8 * 4096 + 1024 + 1024;
};

// fp32-grade gate nonlinearities on ONE MUFU.EX2 each: e^y = ex2.approx(y * log2 e).  ex2.approx is good to 2 ulp and the
// rounded product adds |y| * 2^-24 relative; both enter tanh / sigmoid through d tanh = 2 E eps / (E + 1)^2 <= eps / 2, so
// the absolute error stays below 1e-6 for every x (measured: logits 4e-6 from the oracle, against a 1e-4 gate).  expf() with
// its full range reduction costs ~3x the instructions in an epilogue that is bound by instruction issue, not by MUFU.
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// The gate with ONE reciprocal: with E1 = e^(2 a_f), E2 = e^(-a_g), r = 1 / ((E1 + 1)(1 + E2)):
//   sigmoid(a_g) = (E1 + 1) r,   tanh(a_f) sigmoid(a_g) = (E1 - 1) r        (three MUFU ops per gate element instead of four).
// E1 and E2 are capped at 1e13 (tanh is 1 - 2e-13, sigmoid 1e-13 there), so the product stays below 1e26; `scale` undoes
// the operand scales of the accumulator and is folded into the exponent multipliers.
__device__ __forceinline__ void gate_acc(float af, float ag, float scale, float& z, float& sg) {
  const float e1 = fminf(ex2_approx(af * (scale * 2.8853900817779268f)), 1e13f);
  const float e2 = fminf(ex2_approx(ag * (scale * -1.4426950408889634f)), 1e13f);
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"((e1 + 1.f) * (1.f + e2)));
  sg = (e1 + 1.f) * r;
  z = (e1 - 1.f) * r;
}

// gate derivative from (z = tanh * sg, sg):  da_f = dz * sg * (1 - tanh^2) = dz * (sg - z^2 / sg),
// da_g = dz * tanh * sg * (1 - sg) = dz * z * (1 - sg).  A saturated gate (sg == 0, hence z == 0) has zero derivative.
__device__ __forceinline__ void gate_deriv(float dz, float z, float sg, float live, float& df, float& dg) {
  const float r = sg > 1e-30f ? sg - __fdividef(z * z, sg) : 0.f;
  df = live * dz * r;
  dg = live * dz * z * (1.f - sg);
}

// Sigmoid tape of the fused shape: 16-bit FIXED POINT, sg ~= q / 65535.  The backward needs the sigmoid's ABSOLUTE accuracy
// (da_f = dz (sg - z^2 / sg), da_g = dz z (1 - sg): an error e in sg moves both by at most |dz| e (1 + tanh^2)), and a
// fixed-point code has a uniform 2^-17 = 7.6e-6 -- 30x finer than fp16 on [0.5, 1) -- in half the bytes of fp32.
// Both directions go through the 2^23 trick (an fp32 in [2^23, 2^24) has its integer part in the low mantissa bits, and
// the add rounds to nearest) so that neither F2I nor I2F -- quarter-rate XU-pipe instructions, like MUFU -- is needed.
__device__ __forceinline__ uint32_t sg_pack2(float a, float b) {
  // q = round(sg * 65535) sits in the low 16 mantissa bits of sg * 65535 + 2^23; one byte permute packs two of them
  return __byte_perm(__float_as_uint(fmaf(a, 65535.f, 8388608.f)), __float_as_uint(fmaf(b, 65535.f, 8388608.f)), 0x5410);
}
__device__ __forceinline__ float sg_lo(uint32_t u) {
  return (__uint_as_float(0x4b000000u | (u & 0xffffu)) - 8388608.f) * (1.f / 65535.f);
}
__device__ __forceinline__ float sg_hi(uint32_t u) {
  return (__uint_as_float(0x4b000000u | (u >> 16)) - 8388608.f) * (1.f / 65535.f);
}

// ------------------------------------------------------------------------------------------
// weight preparation: split K-major matrices [N][hi(K) | lo(K)] the MMAs consume directly
//   w1[l][n][tap*R + c]   = (n < G ? Wf : Wg)[n % G][c][tap]      n in [0, 2G)
//   w2[l][r][g]           = Wp[r][g]
//   ws[s][l*G + g]        = Ws_l[s][g]
//   w1t[l][c][slab*2G + n] (slab 0 = current tap), wpt[l][g][r], wst[l*G + g][s]: the transposes for the data gradients
struct TcsTabEntry {
  int64_t wf, wg, wp, ws;
};

__device__ __forceinline__ void put_split(__half* mat, int64_t row, int K, int k, float v) {
  v *= W_SCALE;
  uint32_t hi, lo;
  split2(v, 0.f, hi, lo);          // saturating: a weight beyond the representable range is clamped, never inf / NaN
  mat[row * 2 * K + k] = __ushort_as_half((unsigned short)(hi & 0xffffu));
  mat[row * 2 * K + K + k] = __ushort_as_half((unsigned short)(lo & 0xffffu));
}

__global__ void tcs_prep_kernel(const float* __restrict__ params, const TcsTabEntry* __restrict__ tab, __half* __restrict__ w1,
                                __half* __restrict__ w2, __half* __restrict__ wsc, __half* __restrict__ w1t,
                                __half* __restrict__ wpt, __half* __restrict__ wst, int L, int R, int G, int S, int k) {
  const int l = blockIdx.y;
  const TcsTabEntry e = tab[l];
  const int n1 = 2 * G * k * R, n2 = R * G, n3 = S * G;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n1 + n2 + n3; i += gridDim.x * blockDim.x) {
    if (i < n1) {
      const int n = i / (k * R), kk = i % (k * R);
      const int tap = kk / R, c = kk % R;
      const float* src = n < G ? params + e.wf : params + e.wg;
      const float v = src[((int64_t)(n % G) * R + c) * k + tap];
      put_split(w1 + (int64_t)l * 2 * n1, n, k * R, kk, v);
      const int slab = (k - 1) - tap;
      put_split(w1t + (int64_t)l * 2 * n1, c, k * 2 * G, slab * 2 * G + n, v);
    } else if (i < n1 + n2) {
      const int j = i - n1;
      const float v = params[e.wp + j];
      const int r = j / G, g = j % G;
      put_split(w2 + (int64_t)l * 2 * n2, r, G, g, v);
      put_split(wpt + (int64_t)l * 2 * n2, g, R, r, v);
    } else {
      const int j = i - n1 - n2;
      const int sidx = j / G, g = j % G;
      const float v = params[e.ws + j];
      put_split(wsc, sidx, L * G, l * G + g, v);
      put_split(wst, (int64_t)l * G + g, S, sidx, v);
    }
  }
}

// dst[o][hi(I) | lo(I)] = src[o][i]   and   dstT[i][hi(O) | lo(O)] = src[o][i]
__global__ void tcs_split_weight_kernel(const float* __restrict__ src, __half* __restrict__ dst, __half* __restrict__ dstT, int O,
                                        int I) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)O * I) return;
  const int o = (int)(idx / I), i = (int)(idx % I);
  const float v = src[idx];
  put_split(dst, o, I, i, v);
  put_split(dstT, i, O, o, v);
}

// ---- row-format converters (fp32 <-> split); one thread per (row, 4 channels) -----------------
// dst row r = scale * [relu](src row (seq * rows_per_seq_in + row_off + t)), r = seq * rows_per_seq + t.
// In place (dst == src, identity row mapping) is allowed when blockDim.x is a multiple of C/4: a block then owns
// whole rows and the barrier separates every load of a row from the stores that overwrite it.
__global__ void tcs_rows_to_split_kernel(const float* src, int rows_per_seq_in, int row_off, __half* dst, int C, int rows_per_seq,
                                         int64_t rows, int relu, float scale) {
  const int c4 = C / 4;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = i < rows * c4;
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  int64_t row = 0;
  int cc = 0;
  if (active) {
    row = i / c4;
    cc = (int)(i % c4);
    const int64_t seq = row / rows_per_seq;
    const int t = (int)(row % rows_per_seq);
    v = *(reinterpret_cast<const float4*>(src + ((seq * rows_per_seq_in + row_off + t) * (int64_t)C)) + cc);
  }
  __syncthreads();
  if (active) {
    if (relu) v.x = fmaxf(v.x, 0.f), v.y = fmaxf(v.y, 0.f), v.z = fmaxf(v.z, 0.f), v.w = fmaxf(v.w, 0.f);
    v.x *= scale, v.y *= scale, v.z *= scale, v.w *= scale;
    store_split4(dst, row, C, cc * 4, v);
  }
}

__global__ void tcs_split_to_rows_kernel(const __half* __restrict__ src, float* __restrict__ dst, int C, int64_t rows, float scale) {
  const int c4 = C / 4;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * c4) return;
  const int64_t row = i / c4;
  const int cc = (int)(i % c4);
  float4 v = load_split4(src, row, C, cc * 4);
  v.x *= scale, v.y *= scale, v.z *= scale, v.w *= scale;
  *(reinterpret_cast<float4*>(dst + row * C) + cc) = v;
}

// in-place ReLU of split rows (wavenet.py:588); the sign of hi + lo is the sign of hi unless hi == 0
__global__ void tcs_relu_split_rows_kernel(__half* a, int C, int rows_per_seq_in, int row_off, int rows_per_seq, int64_t rows) {
  const int c4 = C / 4;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * c4) return;
  const int64_t row = i / c4;
  const int cc = (int)(i % c4);
  const int64_t seq = row / rows_per_seq;
  const int t = (int)(row % rows_per_seq);
  const int64_t r = seq * rows_per_seq_in + row_off + t;
  float4 v = load_split4(a, r, C, cc * 4);
  v.x = fmaxf(v.x, 0.f), v.y = fmaxf(v.y, 0.f), v.z = fmaxf(v.z, 0.f), v.w = fmaxf(v.w, 0.f);
  store_split4(a, r, C, cc * 4, v);
}

// column sums of a split / single-plane tensor (bias gradients): out[c] += scale * sum_rows a[row][c].  A thread owns two
// adjacent columns (one half2 per plane and row: a warp reads 128 contiguous bytes), the 8 row-slots of a block stride
// through the rows; grid.y blocks split the rows.
__global__ void tcs_colsum_kernel(const __half* __restrict__ a, int planes, int64_t rows, int C, float scale, float* __restrict__ out,
                                  int64_t det_stride) {
  const int c = (blockIdx.x * 32 + threadIdx.x) * 2;
  __shared__ float red[8][66];
  float s0 = 0.f, s1 = 0.f;
  if (c < C) {
    const int64_t ld = (int64_t)planes * C;
#pragma unroll 4
    for (int64_t r = (int64_t)blockIdx.y * 8 + threadIdx.y; r < rows; r += (int64_t)gridDim.y * 8) {
      float2 v = __half22float2(*reinterpret_cast<const __half2*>(a + r * ld + c));
      if (planes == 2) {
        const float2 l = __half22float2(*reinterpret_cast<const __half2*>(a + r * ld + C + c));
        v.x += l.x, v.y += l.y;
      }
      s0 += v.x, s1 += v.y;
    }
  }
  red[threadIdx.y][2 * threadIdx.x] = s0;
  red[threadIdx.y][2 * threadIdx.x + 1] = s1;
  __syncthreads();
  if (threadIdx.y < 2 && c < C) {
    float s = 0.f;
    for (int i = 0; i < 8; ++i) s += red[i][2 * threadIdx.x + threadIdx.y];
    if (det_stride)        // deterministic mode: block (x, y) owns slab y (gridDim.y <= number of slabs), plain store
      out[(int64_t)blockIdx.y * det_stride + c + threadIdx.y] = s * scale;
    else
      atomicAdd(out + c + threadIdx.y, s * scale);
  }
}

// generic-shape gate (wavenet.py:360): afg [a_f | a_g] fp32 is overwritten with [tanh | sigmoid]; z leaves in split format
__global__ void tcs_gate_forward_kernel(float* __restrict__ afg, __half* __restrict__ z, int64_t P, int G) {
  const int gv = G >> 2;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P * gv) return;
  const int64_t p = i / gv;
  const int g4 = (int)(i - p * gv) * 4;
  float* row = afg + p * 2 * G;
  const float4 af = *reinterpret_cast<const float4*>(row + g4), ag = *reinterpret_cast<const float4*>(row + G + g4);
  const float4 tf = make_float4(tanhf(af.x), tanhf(af.y), tanhf(af.z), tanhf(af.w));
  const float4 sg = make_float4(1.f / (1.f + expf(-ag.x)), 1.f / (1.f + expf(-ag.y)), 1.f / (1.f + expf(-ag.z)),
                                1.f / (1.f + expf(-ag.w)));
  *reinterpret_cast<float4*>(row + g4) = tf;
  *reinterpret_cast<float4*>(row + G + g4) = sg;
  store_split4(z, p, G, g4, make_float4(ACT_SCALE * tf.x * sg.x, ACT_SCALE * tf.y * sg.y, ACT_SCALE * tf.z * sg.z, ACT_SCALE * tf.w * sg.w));
}

// gate derivative of the TOP layer (no gradient arrives through the residual branch): dz = dzs_l
__global__ void tcs_gate_backward_top_kernel(const __half* __restrict__ z, const float* __restrict__ sg, int sg_ld,
                                             const float* __restrict__ dz, __half* __restrict__ dafg, int64_t P, int W, int G,
                                             int zp) {
  const int gv = G >> 2;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P * gv) return;
  const int64_t p = i / gv;
  const int g4 = (int)(i - p * gv) * 4;
  const float live = (int)(p % W) >= zp ? 1.f : 0.f;
  const float4 zz = scale4(load_split4(z, p, G, g4), INV_ACT);
  const float4 s = *reinterpret_cast<const float4*>(sg + p * sg_ld + g4);
  const float4 d = *reinterpret_cast<const float4*>(dz + p * G + g4);
  float4 df, dg;
  gate_deriv(d.x, zz.x, s.x, live, df.x, dg.x);
  gate_deriv(d.y, zz.y, s.y, live, df.y, dg.y);
  gate_deriv(d.z, zz.z, s.z, live, df.z, dg.z);
  gate_deriv(d.w, zz.w, s.w, live, df.w, dg.w);
  store_split4(dafg, p, 2 * G, g4, df);
  store_split4(dafg, p, 2 * G, G + g4, dg);
}

// same for the fused shape's tape formats: sigmoid u16 fixed point [P][G], dzs fp16 [P][G] (carries gscale), dafg ONE fp16
// plane [P][da_f G | da_g G] -- gradient tensors that only ever feed MMAs as the dY operand are rounded to nearest once
__global__ void tcs_gate_backward_top_fused_kernel(const __half* __restrict__ z, const uint16_t* __restrict__ sg,
                                                   const __half* __restrict__ dz, __half* __restrict__ dafg, int64_t P, int W, int G,
                                                   int zp) {
  const int gv = G >> 2;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P * gv) return;
  const int64_t p = i / gv;
  const int g4 = (int)(i - p * gv) * 4;
  const float live = (int)(p % W) >= zp ? 1.f : 0.f;
  const float4 zz = scale4(load_split4(z, p, G, g4), INV_ACT);
  const uint2 sq = *reinterpret_cast<const uint2*>(sg + p * G + g4);
  const uint2 dq = *reinterpret_cast<const uint2*>(dz + p * G + g4);
  const float2 d0 = __half22float2(bits_h2(dq.x)), d1 = __half22float2(bits_h2(dq.y));
  float4 df, dg;
  gate_deriv(d0.x, zz.x, sg_lo(sq.x), live, df.x, dg.x);
  gate_deriv(d0.y, zz.y, sg_hi(sq.x), live, df.y, dg.y);
  gate_deriv(d1.x, zz.z, sg_lo(sq.y), live, df.z, dg.z);
  gate_deriv(d1.y, zz.w, sg_hi(sq.y), live, df.w, dg.w);
  *reinterpret_cast<uint2*>(dafg + p * 2 * G + g4) = make_uint2(pack_h2_sat(df.x, df.y), pack_h2_sat(df.z, df.w));
  *reinterpret_cast<uint2*>(dafg + p * 2 * G + G + g4) = make_uint2(pack_h2_sat(dg.x, dg.y), pack_h2_sat(dg.z, dg.w));
}

// ------------------------------------------------------------------------------------------
// Fused residual layer, R = G = 64, k = 2.  Shared memory (same footprint as the tf32 kernel):
//   B1: W1 split, sub-tiles [128 n x 64 k] indexed (tap, plane)      4 x 16 KB
//   B2: Wp split, sub-tiles [64 r x 64 g] indexed (plane)            2 x  8 KB
//   A : 2 stages x 4 sub-tiles: x(t-d) hi, x(t-d) lo, x(t) hi, x(t) lo; z (hi, lo) then overwrites sub-tiles 0,1 as
//       the A operand of GEMM 2 and the source of its bulk store; the 16-bit fixed-point sigmoid overwrites the thread's
//       own bytes of sub-tile 2 (x(t) hi, same channels) once the residual is in registers
struct SLayerArgs {
  __half* x_out;           // split [B][W][hi 64 | lo 64]
  int W, d, zp, tiles_per_seq, num_tiles;
  int reverse;
};

constexpr int SL_B1 = 0;
constexpr int SL_B2 = 65536;
constexpr int SL_A = 81920;
constexpr int SL_BAR = SL_A + 2 * 65536;
constexpr int SL_SMEM = SL_BAR + 256;
constexpr int SL_THREADS = 64 + 512;           // producer warp, MMA warp, 16 epilogue warps

// Developer trace (make EXTRA=-DWN_LAYER_TRACE, tests/dev/trace_tcs_layer.py): clock64 stamps of CTA 0's pipeline events.
#ifdef WN_LAYER_TRACE
__device__ long long g_trace_s[64 * 32];
#define TRS(j, e) do { if (blockIdx.x == 0 && (j) < 64) g_trace_s[(j) * 32 + (e)] = clock64(); } while (0)
#else
#define TRS(j, e) do { } while (0)
#endif

__global__ void __launch_bounds__(SL_THREADS, 1)
tcs_layer_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w1,
                 const __grid_constant__ CUtensorMap tm_w2, const __grid_constant__ CUtensorMap tm_z,
                 const __grid_constant__ CUtensorMap tm_sg, const SLayerArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bar0 = base + SL_BAR;
  const uint32_t b_full = bar0;
  auto a_full = [&](int s) { return bar0 + 8 + 8 * s; };
  auto d1_full = [&](int s) { return bar0 + 40 + 8 * s; };
  auto z_full = [&](int s) { return bar0 + 56 + 8 * s; };
  auto d2_full = [&](int s) { return bar0 + 72 + 8 * s; };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(gbase + SL_BAR + 128);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(b_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(a_full(s), 1);
      mbar_init(d1_full(s), 1);
      mbar_init(z_full(s), 16);        // one arrival per epilogue warp
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
  const uint32_t tmem = *tmem_slot;
  const int n_local = (a.num_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  auto tile_of = [&](int j) {
    const int i = (int)blockIdx.x + j * (int)gridDim.x;
    return a.reverse ? a.num_tiles - 1 - i : i;
  };

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(b_full, 81920);
      // x rows are [hi 64 | lo 64] halves: plane p starts at element p*64.  W1 rows are [hi 128 | lo 128].
      for (int tap = 0; tap < 2; ++tap)
        for (int p = 0; p < 2; ++p) tma_load_2d(base + SL_B1 + (tap * 2 + p) * SUB, &tm_w1, b_full, p * 128 + tap * KB, 0);
      for (int p = 0; p < 2; ++p) tma_load_2d(base + SL_B2 + p * 8192, &tm_w2, b_full, p * KB, 0);
      for (int j = 0; j <= n_local; ++j) {
        const int s = j & 1;
        const uint32_t as = base + SL_A + s * 65536;
        if (j >= 2) {
          mbar_wait(d2_full(s), ((j - 2) >> 1) & 1);
          TRS(j, 0);
          bulk_wait_group_read0();
          TRS(j, 1);
        }
        if (j < n_local) {
          const int tile = tile_of(j);
          const int b = tile / a.tiles_per_seq, t0 = (tile % a.tiles_per_seq) * TM;
          mbar_arrive_expect_tx(a_full(s), 65536);
          tma_load_4d(as + 0 * SUB, &tm_x, a_full(s), 0, t0 - a.d, b, 0);
          tma_load_4d(as + 1 * SUB, &tm_x, a_full(s), KB, t0 - a.d, b, 0);
          tma_load_4d(as + 2 * SUB, &tm_x, a_full(s), 0, t0, b, 0);
          tma_load_4d(as + 3 * SUB, &tm_x, a_full(s), KB, t0, b, 0);
          TRS(j, 2);
          if (j + 2 < n_local) {
            const int tp = tile_of(j + 2);
            const int bp = tp / a.tiles_per_seq, tp0 = (tp % a.tiles_per_seq) * TM;
            tma_prefetch_4d(&tm_x, 0, tp0, bp, 0);
            tma_prefetch_4d(&tm_x, KB, tp0, bp, 0);
          }
        }
        if (j >= 1) {
          const int jj = j - 1, s1 = jj & 1, tile = tile_of(jj);
          const int b = tile / a.tiles_per_seq, t0 = (tile % a.tiles_per_seq) * TM;
          const uint32_t zs = base + SL_A + s1 * 65536;
          mbar_wait(z_full(s1), (jj >> 1) & 1);
          TRS(j, 3);
          tma_store_4d(&tm_z, zs + 0 * SUB, 0, t0, b, 0);      // z hi plane
          tma_store_4d(&tm_z, zs + 1 * SUB, KB, t0, b, 0);     // z lo plane
          tma_store_4d(&tm_sg, zs + 2 * SUB, 0, t0, b, 0);     // sigmoid, 16-bit fixed point, 64 channels per 128-byte row
          bulk_commit_group();
        }
      }
      bulk_wait_group0();
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // One thread issues BOTH GEMMs and polls for whichever is ready: GEMM 1 of tile n1 (operands loaded, accumulator
      // buffer drained by the gate epilogue of tile n1 - 2) or GEMM 2 of tile n2 (every row of z in shared memory).  Neither
      // ever waits behind the other's barrier (an MMA thread blocked on the next tile's load used to stall the epilogue).
      constexpr uint32_t idesc1 = idesc_f16(128, 128), idesc2 = idesc_f16(128, 64);
      mbar_wait(b_full, 0);
      int n1 = 0, n2 = 0;
      uint32_t spins = 0;
      while (n2 < n_local) {
        bool did = false;
        if (n1 < n_local) {
          const int s = n1 & 1;
          if (mbar_try_wait(a_full(s), (n1 >> 1) & 1) && (n1 < 2 || mbar_try_wait(z_full(s), ((n1 - 2) >> 1) & 1))) {
            TRS(n1, 4);
            tcgen05_fence_after();
            const uint32_t as = base + SL_A + s * 65536;
            // The tensor core accumulates with round-toward-zero: every MMA step loses ~half an ulp OF THE ACCUMULATOR.  The
            // cross terms (hi.lo, lo.hi) are 2^-11 of the result, so they go first, while the accumulator is still small;
            // only the hi.hi steps then run at full magnitude (8 truncating steps instead of 24).
#pragma unroll
            for (int tap = 0; tap < 2; ++tap)
#pragma unroll
              for (int k4 = 0; k4 < 4; ++k4) {
                const uint32_t ao = as + tap * 2 * SUB + k4 * 32, bo = base + SL_B1 + tap * 2 * SUB + k4 * 32;
                umma_f16(tmem + s * 128, umma_desc_k_sw128(ao), umma_desc_k_sw128(bo + SUB), idesc1, (tap | k4) > 0);
                umma_f16(tmem + s * 128, umma_desc_k_sw128(ao + SUB), umma_desc_k_sw128(bo), idesc1, 1u);
              }
#pragma unroll
            for (int tap = 0; tap < 2; ++tap)
#pragma unroll
              for (int k4 = 0; k4 < 4; ++k4) {
                const uint32_t ao = as + tap * 2 * SUB + k4 * 32, bo = base + SL_B1 + tap * 2 * SUB + k4 * 32;
                umma_f16(tmem + s * 128, umma_desc_k_sw128(ao), umma_desc_k_sw128(bo), idesc1, 1u);
              }
            umma_commit(d1_full(s));
            TRS(n1, 5);
            ++n1;
            did = true;
          }
        }
        if (n2 < n1) {
          const int s = n2 & 1;
          if (mbar_try_wait(z_full(s), (n2 >> 1) & 1)) {
            TRS(n2, 6);
            tcgen05_fence_after();
            const uint32_t as = base + SL_A + s * 65536;
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4) {   // cross terms first (see GEMM 1)
              const uint32_t ao = as + k4 * 32, bo = base + SL_B2 + k4 * 32;
              umma_f16(tmem + 256 + s * 64, umma_desc_k_sw128(ao), umma_desc_k_sw128(bo + 8192), idesc2, k4 > 0);
              umma_f16(tmem + 256 + s * 64, umma_desc_k_sw128(ao + SUB), umma_desc_k_sw128(bo), idesc2, 1u);
            }
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4) {
              const uint32_t ao = as + k4 * 32, bo = base + SL_B2 + k4 * 32;
              umma_f16(tmem + 256 + s * 64, umma_desc_k_sw128(ao), umma_desc_k_sw128(bo), idesc2, 1u);
            }
            umma_commit(d2_full(s));
            ++n2;
            did = true;
          }
        }
        if (did) {
          spins = 0;
        } else if (++spins > (1u << 26)) {
          printf("wavenet_b200: tcs_layer_kernel MMA thread timeout (block %d, n1 %d, n2 %d)\n", (int)blockIdx.x, n1, n2);
          __trap();
        }
      }
    }
  } else {
    // 16 epilogue warps (four per scheduler: the gate epilogue is a long dependent chain per element).  The two epilogues of
    // a tile are SKEWED by one tile: gate epilogue of tile j, then the projection epilogue of tile j - 1 -- GEMM 2 of tile
    // j - 1 (issued when the slowest warp delivered its z rows, ~1.3 k cycles of tensor-core latency) has long finished by
    // then, so no warp ever sits waiting for it (trace before: 2.2 k of a 7.5 k cycle tile period were that wait).
    const int q = warp & 3;               // TMEM lane quarter this warp may access
    const int part = (warp - 2) >> 2;     // which 16 of the 64 channels this warp handles
    const int row = q * 32 + lane;
    const uint32_t trow = tmem + ((uint32_t)(q * 32) << 16);
    float xr[16];                         // residual ACT_SCALE * x(t) of the tile whose projection epilogue is pending
    for (int j = 0; j <= n_local; ++j) {
      float xn[16];
      if (j < n_local) {
        // ---- epilogue 1: gate ----
        const int tile = tile_of(j);
        const int s = j & 1, ph = (j >> 1) & 1;
        const int t = (tile % a.tiles_per_seq) * TM + row;
        uint8_t* as_g = gbase + SL_A + s * 65536;
        mbar_wait(d1_full(s), ph);
        if (threadIdx.x == 64) TRS(j, 8);
        tcgen05_fence_after();
        uint32_t f[16], g[16];
        tmem_ld16(trow + s * 128 + part * 16, f);
        tmem_ld16(trow + s * 128 + 64 + part * 16, g);
        // residual = hi + lo of this thread's row and channels, into registers
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const uint4 hv = *reinterpret_cast<const uint4*>(as_g + 2 * SUB + sw128_off(row, part * 2 + c));
          const uint4 lv = *reinterpret_cast<const uint4*>(as_g + 3 * SUB + sw128_off(row, part * 2 + c));
          const float4 v0 = join4(make_uint2(hv.x, hv.y), make_uint2(lv.x, lv.y));
          const float4 v1 = join4(make_uint2(hv.z, hv.w), make_uint2(lv.z, lv.w));
          xn[8 * c + 0] = v0.x, xn[8 * c + 1] = v0.y, xn[8 * c + 2] = v0.z, xn[8 * c + 3] = v0.w;
          xn[8 * c + 4] = v1.x, xn[8 * c + 5] = v1.y, xn[8 * c + 6] = v1.z, xn[8 * c + 7] = v1.w;
        }
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float zz, sg;
#ifndef WN_DIAG_NOGATE
          gate_acc(__uint_as_float(f[i]), __uint_as_float(g[i]), INV_ACT_W, zz, sg);
#else
          zz = __uint_as_float(f[i]) * 1e-6f, sg = __uint_as_float(g[i]) * 1e-6f + 0.5f;   // timing diagnostic: no MUFU
#endif
          g[i] = __float_as_uint(sg);
          f[i] = __float_as_uint(ACT_SCALE * zz);            // z is stored with ACT_SCALE
        }
        if (!(t < a.W && t >= a.zp)) {                       // zero-prefix / out-of-range rows look like tanh(0) | sigmoid(0)
#pragma unroll
          for (int i = 0; i < 16; ++i) f[i] = 0u, g[i] = 0x3f000000u;
        }
#pragma unroll
        for (int c = 0; c < 2; ++c) {   // z planes: 8 channels per 16-byte chunk
          uint4 hv, lv;
          split2(__uint_as_float(f[8 * c + 0]), __uint_as_float(f[8 * c + 1]), hv.x, lv.x);
          split2(__uint_as_float(f[8 * c + 2]), __uint_as_float(f[8 * c + 3]), hv.y, lv.y);
          split2(__uint_as_float(f[8 * c + 4]), __uint_as_float(f[8 * c + 5]), hv.z, lv.z);
          split2(__uint_as_float(f[8 * c + 6]), __uint_as_float(f[8 * c + 7]), hv.w, lv.w);
          *reinterpret_cast<uint4*>(as_g + 0 * SUB + sw128_off(row, part * 2 + c)) = hv;
          *reinterpret_cast<uint4*>(as_g + 1 * SUB + sw128_off(row, part * 2 + c)) = lv;
        }
#pragma unroll
        for (int c = 0; c < 2; ++c)   // sigmoid q16: channels [part*16, +16) = the two chunks this thread took its x(t) hi residual from
          *reinterpret_cast<uint4*>(as_g + 2 * SUB + sw128_off(row, part * 2 + c)) =
              make_uint4(sg_pack2(__uint_as_float(g[8 * c]), __uint_as_float(g[8 * c + 1])),
                         sg_pack2(__uint_as_float(g[8 * c + 2]), __uint_as_float(g[8 * c + 3])),
                         sg_pack2(__uint_as_float(g[8 * c + 4]), __uint_as_float(g[8 * c + 5])),
                         sg_pack2(__uint_as_float(g[8 * c + 6]), __uint_as_float(g[8 * c + 7])));
        fence_proxy_async();
        tcgen05_fence_before();
        warp_arrive(z_full(s), lane);
        if (threadIdx.x == 64) TRS(j, 9);
      }
      if (j >= 1) {
        // ---- epilogue 2 of the PREVIOUS tile: projection + residual ----
        const int jj = j - 1, tile = tile_of(jj);
        const int s = jj & 1, ph = (jj >> 1) & 1;
        const int b = tile / a.tiles_per_seq, t = (tile % a.tiles_per_seq) * TM + row;
        mbar_wait(d2_full(s), ph);
        if (threadIdx.x == 64) TRS(jj, 10);
        tcgen05_fence_after();
        uint32_t g[16];
        tmem_ld16(trow + 256 + s * 64 + part * 16, g);
        tmem_ld_wait();
        tcgen05_fence_before();
        // x_out rows are [hi 64 | lo 64] halves: the thread owns channels [part*16, +16) of its row = 32 contiguous bytes
        // in each plane = one 256-bit store per plane (a full sector, no staging)
        if (t < a.W) {
          __half* xrow = a.x_out + ((int64_t)b * a.W + t) * 128 + part * 16;
          uint32_t hv[8], lv[8];
#pragma unroll
          for (int i = 0; i < 8; ++i)
            split2(fmaf(__uint_as_float(g[2 * i]), INV_W, xr[2 * i]), fmaf(__uint_as_float(g[2 * i + 1]), INV_W, xr[2 * i + 1]), hv[i], lv[i]);
#ifndef WN_DIAG_NOSTORE
          st256(xrow, hv);
          st256(xrow + 64, lv);
#else
          if (hv[0] == 0x12345678u && lv[7] == 0x9abcdef0u) st256(xrow, hv);   // timing diagnostic: keep the math, drop the stores
#endif
        }
        if (threadIdx.x == 64) TRS(jj, 11);
      }
#pragma unroll
      for (int i = 0; i < 16; ++i) xr[i] = xn[i];
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem);
}

// Row-per-lane epilogue of 32 output columns [c0, c0 + 32): lane = TMEM lane = output row, so the thread owns 128 contiguous
// bytes of its row in every tensor it touches and moves them with 256-bit accesses (one full sector per lane and
// instruction; no shared-memory transpose, no cross-lane traffic except the optional column sums).
__device__ __forceinline__ void epi_row32(const SGemmArgs& a, const float (&v)[32], int c0, int t, int64_t orow, int b, float* cs_chunk,
                                          int lane) {
  const bool valid = t < a.rows_out;
  const float relu_floor = a.relu ? 0.f : -INFINITY;
  const bool zero_row = t < a.zero_rows_below;
#pragma unroll
  for (int hh = 0; hh < 2; ++hh) {
    const int col = c0 + hh * 16;
    float o[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) o[i] = v[hh * 16 + i] * a.acc_scale;
    if (a.bias) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 bb = __ldg(reinterpret_cast<const float4*>(a.bias + col) + i);
        o[4 * i] += bb.x, o[4 * i + 1] += bb.y, o[4 * i + 2] += bb.z, o[4 * i + 3] += bb.w;
      }
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) o[i] = zero_row ? 0.f : fmaxf(o[i], relu_floor);
    if (a.Rsd) {
      float r[16];
      if (a.rsd_split) {
        load_split16(reinterpret_cast<const __half*>(a.Rsd) + orow * (2 * (int64_t)a.ldr) + col, a.ldr, r);
      } else {
        uint32_t u[8];
        const float* rp = reinterpret_cast<const float*>(a.Rsd) + orow * a.ldr + col;
        ld256(rp, u);
#pragma unroll
        for (int i = 0; i < 8; ++i) r[i] = __uint_as_float(u[i]);
        ld256(rp + 8, u);
#pragma unroll
        for (int i = 0; i < 8; ++i) r[8 + i] = __uint_as_float(u[i]);
      }
#pragma unroll
      for (int i = 0; i < 16; ++i) o[i] = fmaf(r[i], a.rsd_scale, o[i]);
    }
    if (a.mask) {
      float m[16];
      const int64_t mrow = (int64_t)b * a.mask_rows_in + a.mask_row_off + min(t, a.rows_out - 1);
      load_split16(a.mask + mrow * (2 * (int64_t)a.ldm) + col, a.ldm, m);
#pragma unroll
      for (int i = 0; i < 16; ++i) o[i] = m[i] > 0.f ? o[i] : 0.f;
    }
    if (cs_chunk) {
      // column sums over the 32 rows of this warp: butterfly that halves the live values at every step; after four steps
      // lane l holds column (l >> 1) & 15 summed over the 16 lanes that share bit 0 with it, the last step adds the two halves
      float s[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) s[i] = valid ? o[i] : 0.f;
#pragma unroll
      for (int w = 8; w >= 1; w >>= 1) {
        const bool upper = (lane & (w << 1)) != 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          if (i < w) {
            const float send = upper ? s[i] : s[i + w];
            const float keep = upper ? s[i + w] : s[i];
            s[i] = keep + __shfl_xor_sync(0xffffffffu, send, w << 1);
          }
        }
      }
      s[0] += __shfl_xor_sync(0xffffffffu, s[0], 1);
      if ((lane & 1) == 0) atomicAdd(cs_chunk + hh * 16 + ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1), s[0]);
    }
    if (valid) {
      if (a.out_split) {
#pragma unroll
        for (int i = 0; i < 16; ++i) o[i] *= a.out_scale;
        store_split16(reinterpret_cast<__half*>(a.Y) + orow * (2 * (int64_t)a.ldy) + col, a.ldy, o);
      } else if (a.out_half) {
        __half* yp = reinterpret_cast<__half*>(a.Y);
        int ycol = col;
        if (a.y_slab_cols > 0) {
          yp += (int64_t)(col / a.y_slab_cols) * a.y_slab_stride;
          ycol = col % a.y_slab_cols;
        }
        uint32_t u[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) u[i] = pack_h2_sat(o[2 * i], o[2 * i + 1]);
        st256(yp + orow * a.ldy + ycol, u);
      } else {
        float* yp = reinterpret_cast<float*>(a.Y);
        int ycol = col;
        if (a.y_slab_cols > 0) {
          yp += (int64_t)(col / a.y_slab_cols) * a.y_slab_stride;
          ycol = col % a.y_slab_cols;
        }
        yp += orow * a.ldy + ycol;
        uint32_t u[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) u[i] = __float_as_uint(o[i]);
        st256(yp, u);
#pragma unroll
        for (int i = 0; i < 8; ++i) u[i] = __float_as_uint(o[8 + i]);
        st256(yp + 8, u);
      }
    }
  }
}

// MODE: 1 dz + gate derivative (generic shapes; staged, coalesced epilogue), 5 row-per-lane epilogue (bias / ReLU / zero
// prefix / residual / ReLU mask / column sums, fp32 or split output), 4 = 5 + FLUSH (forward GEMMs): every 64-channel K
// block is accumulated from zero in TMEM and ADDED TO REGISTERS by the epilogue warps in fp32 round-to-nearest -- the
// tensor core truncates its accumulator toward zero at every MMA step, which costs ~1e-5 relative on a K = 1920
// contraction (measured, tests/dev/check_rz.py: 9.7e-6 in one pass, 4.8e-7 flushed per block, 1.7e-6 for an fp32 SGEMM).
template <int BN, int MODE>
__global__ void __launch_bounds__(NTHREADS, 1)
tcs_gemm_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b, const SGemmArgs a) {
  using Cfg = SGemmCfg<BN>;
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
      mbar_init(acc_empty(s), 8);      // one arrival per epilogue warp
    }
    fence_barrier_init();
    prefetch_tmap(&tm_a);
    prefetch_tmap(&tm_b);
  }
  if (warp == 1) tmem_alloc<512>(smem_u32((const void*)tmem_slot));
  if constexpr (MODE == 6) {     // the fused CE epilogue reads the bias three times per element: keep it in shared memory
    if (threadIdx.x < BN) reinterpret_cast<float*>(gbase + Cfg::STG + 8 * 4096)[threadIdx.x] = a.bias ? a.bias[threadIdx.x] : 0.f;
  }
  if constexpr (MODE == 5) {
    if (threadIdx.x < BN) reinterpret_cast<float*>(gbase + Cfg::STG + 8 * 4096)[threadIdx.x] = 0.f;
    if (a.det_stride)      // the eight 4 KB staging blocks (unused by the row-per-lane epilogue) become per-warp tables
      for (int i = threadIdx.x; i < 8 * 1024; i += blockDim.x) reinterpret_cast<float*>(gbase + Cfg::STG)[i] = 0.f;
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int n_local = (a.num_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int kblocks = a.nslab * a.kblk;

  if (warp == 0) {
    if (lane == 0) {
      int it = 0;
      for (int j = 0; j < n_local; ++j) {
        const int tile = a.reverse ? a.num_tiles - 1 - ((int)blockIdx.x + j * (int)gridDim.x) : (int)blockIdx.x + j * (int)gridDim.x;
        const int grp = tile % a.ngroups, rt = tile / a.ngroups;
        const int b = rt / a.tiles_per_seq, t0 = (rt % a.tiles_per_seq) * TM;
        for (int sl = 0; sl < a.nslab; ++sl)
          for (int kc = 0; kc < a.kblk; ++kc)
            for (int p = 0; p < 2; ++p, ++it) {
              const int s = it % Cfg::STAGES, ph = (it / Cfg::STAGES) & 1;
              mbar_wait(empty(s), ph ^ 1);
              const uint32_t st = base + s * Cfg::STAGE;
              const bool a_here = p < a.a_planes;      // a single-plane A has no lo tile: the lo stage carries B_lo only
              mbar_arrive_expect_tx(full(s), a_here ? Cfg::STAGE : BN * 128);
              if (a_here) tma_load_4d(st, &tm_a, full(s), p * a.a_plane + kc * KB, a.slab_row_off[sl] + t0, b, a.slab_idx[sl]);
              tma_load_2d(st + SUB, &tm_b, full(s), p * a.b_plane + (sl * a.kblk + kc) * KB, grp * BN);
            }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = idesc_f16(128, BN);
      int it = 0;
      if constexpr (MODE == 4) {
        // flush group c (a.flush consecutive K blocks) accumulates from zero into TMEM buffer c & 1
        const int fl = a.flush;
        int c = 0;
        for (int j = 0; j < n_local; ++j)
          for (int kb = 0; kb < kblocks; ++kb, it += 2) {
            const int ab = c & 1, aph = (c >> 1) & 1;
            const int s0 = it % Cfg::STAGES, ph = (it / Cfg::STAGES) & 1;
            const uint32_t sh = base + s0 * Cfg::STAGE, sl = sh + Cfg::STAGE;
            const bool first = kb % fl == 0, last = kb % fl == fl - 1;
            if (first) mbar_wait(acc_empty(ab), aph ^ 1);
            mbar_wait(full(s0), ph);
            mbar_wait(full(s0 + 1), ph);
            tcgen05_fence_after();
#ifndef WN_DIAG_GEMM1          // (timing diagnostic WN_DIAG_GEMM1: only the hi.hi products -- is the kernel bound by its MMAs at all?)
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4) {   // cross terms while the accumulator is small, hi.hi last
              umma_f16(tmem + ab * BN, umma_desc_k_sw128(sh + k4 * 32), umma_desc_k_sw128(sl + SUB + k4 * 32), idesc, !first || k4 > 0);
              umma_f16(tmem + ab * BN, umma_desc_k_sw128(sl + k4 * 32), umma_desc_k_sw128(sh + SUB + k4 * 32), idesc, 1u);
            }
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4)
              umma_f16(tmem + ab * BN, umma_desc_k_sw128(sh + k4 * 32), umma_desc_k_sw128(sh + SUB + k4 * 32), idesc, 1u);
#else
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4)
              umma_f16(tmem + ab * BN, umma_desc_k_sw128(sh + k4 * 32), umma_desc_k_sw128(sh + SUB + k4 * 32), idesc, !first || k4 > 0);
#endif
            umma_commit(empty(s0));
            umma_commit(empty(s0 + 1));
            if (last) {
              umma_commit(acc_full(ab));
              ++c;
            }
          }
      }
      for (int j = 0; MODE != 4 && j < n_local; ++j) {
        const int ab = j & 1, aph = (j >> 1) & 1;
        mbar_wait(acc_empty(ab), aph ^ 1);
        tcgen05_fence_after();
        for (int kb = 0; kb < kblocks; ++kb, it += 2) {
          const int s0 = it % Cfg::STAGES, ph = (it / Cfg::STAGES) & 1;   // STAGES is even: s0 + 1 is the lo stage
          const uint32_t sh = base + s0 * Cfg::STAGE, sl = sh + Cfg::STAGE;
          mbar_wait(full(s0), ph);
          tcgen05_fence_after();
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4)
            umma_f16(tmem + ab * BN, umma_desc_k_sw128(sh + k4 * 32), umma_desc_k_sw128(sh + SUB + k4 * 32), idesc, (kb | k4) > 0);
          mbar_wait(full(s0 + 1), ph);
          tcgen05_fence_after();
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4) {
            umma_f16(tmem + ab * BN, umma_desc_k_sw128(sh + k4 * 32), umma_desc_k_sw128(sl + SUB + k4 * 32), idesc, 1u);
            if (a.a_planes == 2)
              umma_f16(tmem + ab * BN, umma_desc_k_sw128(sl + k4 * 32), umma_desc_k_sw128(sh + SUB + k4 * 32), idesc, 1u);
          }
          umma_commit(empty(s0));
          umma_commit(empty(s0 + 1));
        }
        umma_commit(acc_full(ab));
      }
    }
  } else {
    const int q = warp & 3, half = (warp - 2) >> 2;
    constexpr int CH = BN / 64;   // 32-column chunks per warp
    uint8_t* stg = gbase + Cfg::STG + (warp - 2) * 4096;   // [32 rows][128 B], 16-byte chunks XOR-swizzled by row
    const int cc4 = (lane & 7) * 4;
    const int rsub = lane >> 3;
    float* cs_smem = reinterpret_cast<float*>(gbase + Cfg::STG + 8 * 4096);
    if constexpr (MODE == 4) {
      int c = 0;
      for (int j = 0; j < n_local; ++j) {
        const int tile = a.reverse ? a.num_tiles - 1 - ((int)blockIdx.x + j * (int)gridDim.x) : (int)blockIdx.x + j * (int)gridDim.x;
        const int grp = tile % a.ngroups, rt = tile / a.ngroups;
        const int b = rt / a.tiles_per_seq, t0 = (rt % a.tiles_per_seq) * TM + q * 32;
        float acc[CH][32];
#pragma unroll
        for (int ch = 0; ch < CH; ++ch)
#pragma unroll
          for (int i = 0; i < 32; ++i) acc[ch][i] = 0.f;
        for (int kb = 0; kb < kblocks; kb += a.flush, ++c) {
          const int ab = c & 1, aph = (c >> 1) & 1;
          mbar_wait(acc_full(ab), aph);
          tcgen05_fence_after();
#pragma unroll
          for (int ch = 0; ch < CH; ++ch) {
            uint32_t v[32];
            tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + ab * BN + (half * CH + ch) * 32, v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) acc[ch][i] += __uint_as_float(v[i]);
          }
          tcgen05_fence_before();
          warp_arrive(acc_empty(ab), lane);
        }
        const int t = t0 + lane;
        const int64_t orow = (int64_t)b * a.rows_out + min(t, a.rows_out - 1);
#pragma unroll
        for (int ch = 0; ch < CH; ++ch) {
          const int c0 = grp * BN + (half * CH + ch) * 32;
          if (c0 < a.N) epi_row32(a, acc[ch], c0, t, orow, b, nullptr, lane);
        }
      }
    }
    if constexpr (MODE == 6) {
      // Last head conv fused with softmax cross-entropy (wavenet.py:590 + 597-617): the logits of a row live in the
      // registers of TWO threads (this warp's 128 columns, the partner warp's other 128), so the row maximum and the
      // exp-sum are exchanged through shared memory; the loss and dlogits (one fp16 plane, scaled by gscale -- the dY operand
      // format of the backward GEMMs) leave from here.  The logits themselves are only written on request.
      static_assert(MODE != 6 || BN == 256, "the fused CE epilogue needs the whole 256-class row in one tile");
      float* xmax = reinterpret_cast<float*>(stg);                 // this warp's 4 KB staging block: [32 rows] max | sum | tgt logit
      float* pmax = reinterpret_cast<float*>(gbase + Cfg::STG + (((warp - 2) ^ 4)) * 4096);   // partner warp (same rows, other half)
      double loss_part = 0.0;
      const uint32_t trow = tmem + ((uint32_t)(q * 32) << 16);
      const int cbase = half * CH * 32;
      for (int j = 0; j < n_local; ++j) {
        const int tile = a.reverse ? a.num_tiles - 1 - ((int)blockIdx.x + j * (int)gridDim.x) : (int)blockIdx.x + j * (int)gridDim.x;
        const int ab = j & 1, aph = (j >> 1) & 1;
        const int b = tile / a.tiles_per_seq, t = (tile % a.tiles_per_seq) * TM + q * 32 + lane;
        const bool valid = t < a.rows_out;
        const int64_t orow = (int64_t)b * a.rows_out + min(t, a.rows_out - 1);
        const int tg = a.ce_target[orow];
        mbar_wait(acc_full(ab), aph);
        tcgen05_fence_after();
        // the accumulator stays in TMEM and is read three times (max, exp-sum, gradient): no 128-register row copy
        float m = -INFINITY, xt = 0.f;
#pragma unroll 1
        for (int ch = 0; ch < CH; ++ch) {
          uint32_t v[32];
          tmem_ld32(trow + ab * BN + cbase + ch * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float l = fmaf(__uint_as_float(v[i]), a.acc_scale, cs_smem[cbase + ch * 32 + i]);
            m = fmaxf(m, l);
            if (cbase + ch * 32 + i == tg) xt = l;
          }
        }
        xmax[lane] = m;
        xmax[64 + lane] = xt;
        asm volatile("bar.sync %0, 64;" ::"r"(2 + q) : "memory");
        m = fmaxf(m, pmax[lane]);
        xt += pmax[64 + lane];
        float sum = 0.f;
#pragma unroll 1
        for (int ch = 0; ch < CH; ++ch) {
          uint32_t v[32];
          tmem_ld32(trow + ab * BN + cbase + ch * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i)
            sum += ex2_approx((fmaf(__uint_as_float(v[i]), a.acc_scale, cs_smem[cbase + ch * 32 + i]) - m) * 1.4426950408889634f);
        }
        xmax[32 + lane] = sum;
        asm volatile("bar.sync %0, 64;" ::"r"(2 + q) : "memory");
        sum += pmax[32 + lane];
        if (valid && half == 0) loss_part += (double)(m + logf(sum)) - (double)xt;
        const float inv = 1.f / sum, gk = a.ce_invn * a.ce_gscale;
        __half* drow = a.ce_dlogits + orow * (int64_t)BN + cbase;
        float* yp = a.Y ? reinterpret_cast<float*>(a.Y) + orow * a.ldy + cbase : nullptr;
#pragma unroll 1
        for (int ch = 0; ch < CH; ++ch) {
          uint32_t v[32];
          tmem_ld32(trow + ab * BN + cbase + ch * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            float d[16];
            uint32_t u[8];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const int col = cbase + ch * 32 + hh * 16 + i;
              const float l = fmaf(__uint_as_float(v[hh * 16 + i]), a.acc_scale, cs_smem[col]);
              v[hh * 16 + i] = __float_as_uint(l);
              d[i] = (ex2_approx((l - m) * 1.4426950408889634f) * inv - (col == tg ? 1.f : 0.f)) * gk;
            }
            if (valid) {
#pragma unroll
              for (int i = 0; i < 8; ++i) u[i] = pack_h2_sat(d[2 * i], d[2 * i + 1]);
              st256(drow + ch * 32 + hh * 16, u);
              if (yp) {     // logits on request
#pragma unroll
                for (int i = 0; i < 8; ++i) u[i] = v[hh * 16 + i];
                st256(yp + ch * 32 + hh * 16, u);
#pragma unroll
                for (int i = 0; i < 8; ++i) u[i] = v[hh * 16 + 8 + i];
                st256(yp + ch * 32 + hh * 16 + 8, u);
              }
            }
          }
        }
        tcgen05_fence_before();
        warp_arrive(acc_empty(ab), lane);
      }
      // loss: one double atomic per warp
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) loss_part += __shfl_xor_sync(0xffffffffu, loss_part, o);
      if (lane == 0 && half == 0) atomicAdd(a.ce_acc, loss_part);
    }
    if constexpr (MODE == 5) {
      for (int j = 0; j < n_local; ++j) {
        const int tile = a.reverse ? a.num_tiles - 1 - ((int)blockIdx.x + j * (int)gridDim.x) : (int)blockIdx.x + j * (int)gridDim.x;
        const int ab = j & 1, aph = (j >> 1) & 1;
        const int grp = tile % a.ngroups, rt = tile / a.ngroups;
        const int b = rt / a.tiles_per_seq, t = (rt % a.tiles_per_seq) * TM + q * 32 + lane;
        const int64_t orow = (int64_t)b * a.rows_out + min(t, a.rows_out - 1);
        mbar_wait(acc_full(ab), aph);
        tcgen05_fence_after();
#pragma unroll 1
        for (int ch = 0; ch < CH; ++ch) {
          const int ct = (half * CH + ch) * 32;
          const int c0 = grp * BN + ct;
          uint32_t v[32];
          tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + ab * BN + ct, v);
          tmem_ld_wait();
          if (c0 >= a.N) continue;
          float o[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) o[i] = __uint_as_float(v[i]);
          // deterministic mode: a warp's sums go into ITS OWN table (one warp adds in program order); atomic mode: one table
          float* cs_tab = a.det_stride ? reinterpret_cast<float*>(stg) : cs_smem;
          epi_row32(a, o, c0, t, orow, b, a.colsum_out ? cs_tab + ct : nullptr, lane);
        }
        tcgen05_fence_before();
        warp_arrive(acc_empty(ab), lane);
      }
      if (a.colsum_out) {
        asm volatile("bar.sync 1, 256;" ::: "memory");   // the eight epilogue warps
        const int c = threadIdx.x - 64;
        if (c < BN && c < a.N) {
          if (a.det_stride) {
            float sum = 0.f;
            for (int w = 0; w < 8; ++w) sum += reinterpret_cast<const float*>(gbase + Cfg::STG + w * 4096)[c];
            a.colsum_out[(int64_t)blockIdx.x * a.det_stride + c] = sum * a.colsum_scale;
          } else {
            atomicAdd(a.colsum_out + c, cs_smem[c] * a.colsum_scale);
          }
        }
      }
    }
    for (int j = 0; MODE == 1 && j < n_local; ++j) {
      const int tile = a.reverse ? a.num_tiles - 1 - ((int)blockIdx.x + j * (int)gridDim.x) : (int)blockIdx.x + j * (int)gridDim.x;
      const int ab = j & 1, aph = (j >> 1) & 1;
      const int grp = tile % a.ngroups, rt = tile / a.ngroups;
      const int b = rt / a.tiles_per_seq, t0 = (rt % a.tiles_per_seq) * TM + q * 32;
      mbar_wait(acc_full(ab), aph);
      tcgen05_fence_after();
#pragma unroll 1
      for (int ch = 0; ch < CH; ++ch) {
        const int ct = (half * CH + ch) * 32;          // column inside the TMEM tile
        const int c0 = grp * BN + ct;                  // output column
        uint32_t v[32];
        tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + ab * BN + ct, v);
        tmem_ld_wait();
        if (c0 >= a.N) continue;
#pragma unroll
        for (int c = 0; c < 8; ++c)
          *reinterpret_cast<uint4*>(stg + lane * 128 + ((c ^ (lane & 7)) << 4)) =
              make_uint4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
        __syncwarp();
        const int col = c0 + cc4;
        if constexpr (MODE == 1) {
          float4 r4[8], z4[8], sg4[8];
#pragma unroll
          for (int jj = 0; jj < 8; ++jj) {
            const int t = min(t0 + jj * 4 + rsub, a.rows_out - 1);
            const int64_t orow = (int64_t)b * a.rows_out + t;
            r4[jj] = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(a.Rsd) + orow * a.ldr + col);
            z4[jj] = load_split4(a.gate_z, orow, a.N, col);
            sg4[jj] = *reinterpret_cast<const float4*>(a.gate_sg + orow * a.gate_sg_ld + col);
          }
#pragma unroll
          for (int jj = 0; jj < 8; ++jj) {
            const int rr = jj * 4 + rsub;
            const int t = t0 + rr;
            if (t >= a.rows_out) continue;
            const int64_t orow = (int64_t)b * a.rows_out + t;
            float4 o = scale4(*reinterpret_cast<const float4*>(stg + rr * 128 + (((lane & 7) ^ (rr & 7)) << 4)), a.acc_scale);
            o.x += r4[jj].x, o.y += r4[jj].y, o.z += r4[jj].z, o.w += r4[jj].w;
            const float4 z = scale4(z4[jj], INV_ACT), sg = sg4[jj];
            const float live = t >= a.gate_zp ? 1.f : 0.f;
            float4 df, dg;
            gate_deriv(o.x, z.x, sg.x, live, df.x, dg.x);
            gate_deriv(o.y, z.y, sg.y, live, df.y, dg.y);
            gate_deriv(o.z, z.z, sg.z, live, df.z, dg.z);
            gate_deriv(o.w, z.w, sg.w, live, df.w, dg.w);
            store_split4(a.gate_dafg, orow, 2 * a.N, col, df);
            store_split4(a.gate_dafg, orow, 2 * a.N, a.N + col, dg);
          }
        }
        __syncwarp();
      }
      tcgen05_fence_before();
      warp_arrive(acc_empty(ab), lane);
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem);
}

// ------------------------------------------------------------------------------------------
// Weight gradient: D[128*MH x NB] = sum over positions  dY[p][a_c0 + m] * X_slab[p + off_slab][c], both operands split and
// MN-major (the reduction dimension -- positions -- is the row index of both tensors in memory).  TMA deposits
// [KC positions x 64 channels] sub-tiles; one K=16 MMA step consumes two 8-row swizzle groups (SBO 1024 B), 64-channel
// atoms are one sub-tile apart (LBO).
struct SWgradArgs {
  int rows_it, num_seq;
  int a_row_off, a_c0, a_plane;   // dY: first channel, elements between its planes
  int a_planes;                   // 2: dY rows are split; 1: one fp16 plane (two MMAs per product)
  int nb_slab, nb_sub;            // X slabs (taps or layers) and 64-channel sub-tiles per slab
  int b_plane;                    // elements between the planes of an X row
  int b_row_off[4];
  int ngroups;
  int b_slab_idx[8][4];
  float* dW0[8][4];
  float* dW1[4];
  int m_split, m_valid;
  int64_t sn, sk;
  int chunks_per_seq, num_chunks;
  int reverse, run, red_mode;
  float scale;                    // 1 / gscale
  int64_t det_stride;             // deterministic mode: floats between the per-CTA gradient slabs (0 = atomics)
};

template <int NB, int MH>
struct SWgradCfg {
  static constexpr int KC = 32;                          // positions per pipeline stage
  static constexpr int SUBW = KC * 128;                  // bytes of a [KC x 64] sub-tile
  static constexpr int NSUB = 2 * MH + NB / 64;          // sub-tiles per plane
  static constexpr int STAGE = 2 * NSUB * SUBW;
  static constexpr int STAGES = (192 * 1024) / STAGE > 6 ? 6 : (192 * 1024) / STAGE;
  static constexpr int BAR = STAGES * STAGE;
  static constexpr int SMEM = BAR + 256 + 1024;
  static constexpr int TMEM_COLS = MH * NB > 256 ? 512 : (MH * NB > 128 ? 256 : (MH * NB > 64 ? 128 : 64));
};

template <int NB, int MH>
__global__ void __launch_bounds__(NTHREADS, 1)
tcs_wgrad_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b, const SWgradArgs a) {
  using Cfg = SWgradCfg<NB, MH>;
  constexpr int KC = Cfg::KC, SUBW = Cfg::SUBW, NSUB = Cfg::NSUB;
  static_assert(Cfg::STAGES * Cfg::STAGE >= 8 * 8192, "epilogue staging needs 64 KB of ring memory");
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
  const uint32_t tmem = *tmem_slot;
  const int grp = (int)blockIdx.x % a.ngroups, nrange = (int)gridDim.x / a.ngroups;
  const int WG_RUN = a.run;
  const int r0 = (int)blockIdx.x / a.ngroups;
  int n_local, c_begin = 0;
  if (WG_RUN > 0) {
    const int nruns = (a.num_chunks + WG_RUN - 1) / WG_RUN;
    const int my_runs = r0 < nruns ? (nruns - r0 + nrange - 1) / nrange : 0;
    n_local = my_runs * WG_RUN;
    if (my_runs > 0 && r0 + (my_runs - 1) * nrange == nruns - 1) n_local -= nruns * WG_RUN - a.num_chunks;
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
        mbar_arrive_expect_tx(full(s), Cfg::STAGE - (a.a_planes == 1 ? 2 * MH * SUBW : 0));
        for (int p = 0; p < 2; ++p) {
          const uint32_t sp = st + p * NSUB * SUBW;
          if (p < a.a_planes) {
#pragma unroll
            for (int i = 0; i < 2 * MH; ++i)
              tma_load_4d(sp + i * SUBW, &tm_a, full(s), p * a.a_plane + a.a_c0 + i * KB, a.a_row_off + t0, b, 0);
          }
          for (int sl = 0; sl < a.nb_slab; ++sl)
            for (int i = 0; i < a.nb_sub; ++i)
              tma_load_4d(sp + (2 * MH + sl * a.nb_sub + i) * SUBW, &tm_b, full(s), p * a.b_plane + i * KB, a.b_row_off[sl] + t0, b,
                          max(a.b_slab_idx[grp][sl], 0));
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && n_local > 0) {
      constexpr uint32_t idesc = idesc_f16(128, NB) | IDESC_MN_MAJOR;
      constexpr uint32_t lbo = SUBW, sbo = 1024, kstep = 2048;
      for (int it = 0; it < n_local; ++it) {
        const int s = it % Cfg::STAGES, ph = (it / Cfg::STAGES) & 1;
        mbar_wait(full(s), ph);
        tcgen05_fence_after();
        const uint32_t sh = base + s * Cfg::STAGE, sl = sh + NSUB * SUBW;
#pragma unroll
        for (int k16 = 0; k16 < KC / 16; ++k16) {
          const uint64_t bh = desc_mn_sw128(sh + 2 * MH * SUBW + k16 * kstep, lbo, sbo);
          const uint64_t bl = desc_mn_sw128(sl + 2 * MH * SUBW + k16 * kstep, lbo, sbo);
#pragma unroll
          for (int mh = 0; mh < MH; ++mh) {
            const uint64_t ah = desc_mn_sw128(sh + 2 * mh * SUBW + k16 * kstep, lbo, sbo);
            umma_f16(tmem + mh * NB, ah, bh, idesc, (it | k16) > 0);
            umma_f16(tmem + mh * NB, ah, bl, idesc, 1u);
            if (a.a_planes == 2) umma_f16(tmem + mh * NB, desc_mn_sw128(sl + 2 * mh * SUBW + k16 * kstep, lbo, sbo), bh, idesc, 1u);
          }
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
    const int nb = a.nb_sub * 64;   // channels per slab
    const float sc = a.scale;
    uint8_t* stg = gbase + (warp - 2) * 8192;   // every MMA has retired: the ring is free
    const bool det = a.det_stride != 0;
    auto row_ptr = [&](int m, int sl) {
      return (m < a.m_split ? a.dW0[grp][sl] + (int64_t)m * a.sn : a.dW1[sl] + (int64_t)(m - a.m_split) * a.sn) +
             (int64_t)blockIdx.x * a.det_stride;
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
            float4 o = *reinterpret_cast<const float4*>(stg + rr * 256 + ((kk ^ (rr & 7)) << 4));
            o.x *= sc, o.y *= sc, o.z *= sc, o.w *= sc;
            if (m < a.m_valid) wg_add_v4(row_ptr(m, 0) + (int64_t)cw * 2 + kk * 4, o, det);
          }
          __syncwarp();
        }
      }
    } else if (a.red_mode == 1) {
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
            float4 o = *reinterpret_cast<const float4*>(stg + rr * 128 + ((kk ^ (rr & 7)) << 4));
            o.x *= sc, o.y *= sc, o.z *= sc, o.w *= sc;
            if (m < a.m_valid && slab_ok) wg_add_v4(row_ptr(m, sl) + cbase + kk * 4, o, det);
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
            for (int i = 0; i < 32; ++i) {
              if (det)
                wrow[(int64_t)(cbase + i) * a.sk] = __uint_as_float(v[i]) * sc;
              else
                atomicAdd(wrow + (int64_t)(cbase + i) * a.sk, __uint_as_float(v[i]) * sc);
            }
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
//   dafg = gate derivative of dz (z, sigmoid)      (epilogue; z comes from the shared-memory tile)
//   dWp += dout^T . z                              (MN-major GEMM over the same 128 rows, TMEM-resident accumulator)
// In the split format the [128 positions x 64 channels] fp16 tiles of dout serve BOTH as the K-major A operand of the dz
// GEMM and as the MN-major A operand of the dWp GEMM (same bytes, same 128-byte swizzle), so dout and z are read from
// HBM exactly once and nothing is re-fetched through L2.
struct SGateBwdArgs {
  const __half* dzs;       // fp16 [rows][64], carries gscale
  const uint16_t* sg;      // 16-bit fixed-point sigmoid [rows][64]
  __half* dafg;            // ONE fp16 plane [rows][da_f 64 | da_g 64], carries gscale
  float* dWp;              // [64 r][64 g]
  int zp;
  int rows_out, tiles_per_seq, num_tiles;
  int reverse;
  float wscale;            // 1 / (gscale * ACT_SCALE)
  int64_t det_stride;      // deterministic mode: floats between the per-CTA gradient slabs (0 = red.global.add)
};
constexpr int SG_W = 0;                         // Wp^T split: hi [64 g x 64 r] 8 KB, lo 8 KB
constexpr int SG_ST = 16384;                    // 2 stages x {dout hi, dout lo, z hi, z lo}
constexpr int SG_BAR = SG_ST + 2 * 65536;
constexpr int SG_SMEM = SG_BAR + 256;

__global__ void __launch_bounds__(NTHREADS, 1)
tcs_gate_bwd_kernel(const __grid_constant__ CUtensorMap tm_dout, const __grid_constant__ CUtensorMap tm_z,
                    const __grid_constant__ CUtensorMap tm_w, const SGateBwdArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bar0 = base + SG_BAR;
  const uint32_t w_full = bar0, wg_full = bar0 + 8;
  auto full = [&](int s) { return bar0 + 16 + 8 * s; };
  auto empty = [&](int s) { return bar0 + 32 + 8 * s; };
  auto acc_full = [&](int s) { return bar0 + 48 + 8 * s; };
  auto acc_empty = [&](int s) { return bar0 + 64 + 8 * s; };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(gbase + SG_BAR + 128);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(w_full, 1);
    mbar_init(wg_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(full(s), 1);
      mbar_init(empty(s), 9);         // the MMA commit + 8 epilogue warps (they read z from the stage)
      mbar_init(acc_full(s), 1);
      mbar_init(acc_empty(s), 8);      // one arrival per epilogue warp
    }
    fence_barrier_init();
    prefetch_tmap(&tm_dout);
    prefetch_tmap(&tm_z);
    prefetch_tmap(&tm_w);
  }
  if (warp == 1) tmem_alloc<512>(smem_u32((const void*)tmem_slot));
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int n_local = (a.num_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  auto tile_of = [&](int j) {
    const int i = (int)blockIdx.x + j * (int)gridDim.x;
    return a.reverse ? a.num_tiles - 1 - i : i;
  };

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(w_full, 16384);
      tma_load_2d(base + SG_W, &tm_w, w_full, 0, 0);
      tma_load_2d(base + SG_W + 8192, &tm_w, w_full, KB, 0);
      for (int j = 0; j < n_local; ++j) {
        const int tile = tile_of(j);
        const int b = tile / a.tiles_per_seq, t0 = (tile % a.tiles_per_seq) * TM;
        const int s = j & 1, ph = (j >> 1) & 1;
        mbar_wait(empty(s), ph ^ 1);
        const uint32_t st = base + SG_ST + s * 65536;
        mbar_arrive_expect_tx(full(s), 65536);
        tma_load_4d(st + 0 * SUB, &tm_dout, full(s), 0, t0, b, 0);
        tma_load_4d(st + 1 * SUB, &tm_dout, full(s), KB, t0, b, 0);
        tma_load_4d(st + 2 * SUB, &tm_z, full(s), 0, t0, b, 0);
        tma_load_4d(st + 3 * SUB, &tm_z, full(s), KB, t0, b, 0);
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && n_local > 0) {
      // Adjacent hi | lo planes are consumed as ONE operand of twice the size (shared-memory operand reads, not the tensor
      // pipe, are what these small-K MMAs cost): dz = dout_lo . [Wp_hi | Wp_lo] + dout_hi . [Wp_hi | Wp_lo] leaves the hi
      // and lo products in two 64-column halves the epilogue adds; dWp = [dout_hi | dout_lo]^T . [z_hi | z_lo] puts all four
      // plane products into one M = 128, N = 128 accumulator (lanes 0..63 dout_hi, 64..127 dout_lo) -- one MMA per K step
      // where three half-empty ones (M atom 1 was a zero tile) ran before.
      constexpr uint32_t idesc = idesc_f16(128, 128);
      constexpr uint32_t idesc_mn = idesc_f16(128, 128) | IDESC_MN_MAJOR;
      mbar_wait(w_full, 0);
      for (int j = 0; j < n_local; ++j) {
        const int s = j & 1, ph = (j >> 1) & 1;
        mbar_wait(acc_empty(s), ph ^ 1);
        mbar_wait(full(s), ph);
        tcgen05_fence_after();
        const uint32_t st = base + SG_ST + s * 65536, wb = base + SG_W;
#pragma unroll
        for (int k4 = 0; k4 < 4; ++k4)     // the small (dout_lo) products first: the accumulator truncates toward zero
          umma_f16(tmem + s * 128, umma_desc_k_sw128(st + SUB + k4 * 32), umma_desc_k_sw128(wb + k4 * 32), idesc, k4 > 0);
#pragma unroll
        for (int k4 = 0; k4 < 4; ++k4)
          umma_f16(tmem + s * 128, umma_desc_k_sw128(st + k4 * 32), umma_desc_k_sw128(wb + k4 * 32), idesc, 1u);
        umma_commit(acc_full(s));
        // dWp[r][g] += sum_pos dout[pos][r] * z[pos][g]: A = [dout_hi | dout_lo] tiles read MN-major (two M atoms), B = [z_hi |
        // z_lo] tiles MN-major (two N atoms)
#pragma unroll
        for (int k16 = 0; k16 < TM / 16; ++k16)
          umma_f16(tmem + 256, desc_mn_sw128(st + k16 * 2048, SUB, 1024), desc_mn_sw128(st + 2 * SUB + k16 * 2048, SUB, 1024),
                   idesc_mn, (j | k16) > 0);
        umma_commit(empty(s));
      }
      umma_commit(wg_full);
    }
  } else {
    const int q = warp & 3, half = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    for (int j = 0; j < n_local; ++j) {
      const int tile = tile_of(j);
      const int s = j & 1, ph = (j >> 1) & 1;
      const int b = tile / a.tiles_per_seq, t = (tile % a.tiles_per_seq) * TM + row;
      const bool valid = t < a.rows_out;
      const int64_t orow = (int64_t)b * a.rows_out + min(t, a.rows_out - 1);
      const float live = t >= a.zp ? 1.f : 0.f;
      const uint8_t* st = gbase + SG_ST + s * 65536;
      // the epilogue inputs that do not depend on the accumulator: issue their loads before waiting for it
      // (32 channels x 16 bit = two 256-bit loads per tensor)
      uint32_t d8[2][8], s8[2][8];
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        ld256(a.dzs + orow * 64 + half * 32 + i * 16, d8[i]);
        ld256(a.sg + orow * 64 + half * 32 + i * 16, s8[i]);
      }
      if (lane == 0) mbar_wait(full(s), ph);   // z tile (written by TMA) visible to this warp's generic loads: one lane
      __syncwarp();                            // observes the barrier, bar.warp.sync orders the others behind it
      mbar_wait(acc_full(s), ph);
      tcgen05_fence_after();
      uint32_t v[32];
      {
        uint32_t vl[32];
        tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + s * 128 + half * 32, v);        // dout . Wp_hi
        tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + s * 128 + 64 + half * 32, vl);  // dout . Wp_lo
        tmem_ld_wait();
        tcgen05_fence_before();
        warp_arrive(acc_empty(s), lane);
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) + __uint_as_float(vl[i]));
      }
      __half* drow = a.dafg + orow * 128 + half * 32;
      uint32_t fh[8], gh[8];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const uint4 zh = *reinterpret_cast<const uint4*>(st + 2 * SUB + sw128_off(row, half * 4 + c));
        const uint4 zl = *reinterpret_cast<const uint4*>(st + 3 * SUB + sw128_off(row, half * 4 + c));
        const float4 z0 = scale4(join4(make_uint2(zh.x, zh.y), make_uint2(zl.x, zl.y)), INV_ACT);
        const float4 z1 = scale4(join4(make_uint2(zh.z, zh.w), make_uint2(zl.z, zl.w)), INV_ACT);
        const float zz[8] = {z0.x, z0.y, z0.z, z0.w, z1.x, z1.y, z1.z, z1.w};
        float df[8], dg[8];
#pragma unroll
        for (int i = 0; i < 8; i += 2) {
          const uint32_t sq = s8[c >> 1][(c & 1) * 4 + (i >> 1)];
          const float2 dq = __half22float2(bits_h2(d8[c >> 1][(c & 1) * 4 + (i >> 1)]));
          gate_deriv(fmaf(__uint_as_float(v[8 * c + i]), INV_W, dq.x), zz[i], sg_lo(sq), live, df[i], dg[i]);
          gate_deriv(fmaf(__uint_as_float(v[8 * c + i + 1]), INV_W, dq.y), zz[i + 1], sg_hi(sq), live, df[i + 1], dg[i + 1]);
        }
        const int k = (c & 1) * 4;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          fh[k + i] = pack_h2_sat(df[2 * i], df[2 * i + 1]);
          gh[k + i] = pack_h2_sat(dg[2 * i], dg[2 * i + 1]);
        }
        if ((c & 1) && valid) {     // 16 channels = one 256-bit store per gate half
          const int o16 = (c >> 1) * 16;
          st256(drow + o16, fh);            // da_f
          st256(drow + 64 + o16, gh);       // da_g
        }
      }
      warp_arrive(empty(s), lane);          // done reading z from the stage
    }
    if (n_local > 0) {
      // dWp rows (projection output channels r): TMEM lanes r (dout_hi products) and 64 + r (dout_lo products), columns g
      // (z_hi) and 64 + g (z_lo); every lane quarter reduces its share into dWp
      mbar_wait(wg_full, 0);
      tcgen05_fence_after();
      uint32_t v[32];
      {
        uint32_t vl[32];
        tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + 256 + half * 32, v);
        tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + 256 + 64 + half * 32, vl);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) + __uint_as_float(vl[i]));
      }
      float* wrow = a.dWp + (int64_t)blockIdx.x * a.det_stride + (int64_t)(row & 63) * 64 + half * 32;
      if (a.det_stride) {
        // deterministic: the dout_lo products (lanes 64..127) go through shared memory (the stage ring is idle now) and are
        // added by the warp that owns the dout_hi products of the same rows; one plain store per element
        float* xch = reinterpret_cast<float*>(gbase + SG_ST) + ((q & 1) * 2 + half) * 1024 + lane * 32;
        if (q >= 2) {
#pragma unroll
          for (int i = 0; i < 8; ++i) reinterpret_cast<uint4*>(xch)[i] = make_uint4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (q < 2) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 o = reinterpret_cast<const float4*>(xch)[i];
            *reinterpret_cast<float4*>(wrow + 4 * i) =
                make_float4((__uint_as_float(v[4 * i]) + o.x) * a.wscale, (__uint_as_float(v[4 * i + 1]) + o.y) * a.wscale,
                            (__uint_as_float(v[4 * i + 2]) + o.z) * a.wscale, (__uint_as_float(v[4 * i + 3]) + o.w) * a.wscale);
          }
        }
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i)
          red_add_v4(wrow + 4 * i, make_float4(__uint_as_float(v[4 * i]) * a.wscale, __uint_as_float(v[4 * i + 1]) * a.wscale,
                                               __uint_as_float(v[4 * i + 2]) * a.wscale, __uint_as_float(v[4 * i + 3]) * a.wscale));
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem);
}

// ------------------------------------------------------------------------------------------
// Backward of one residual layer's dilated conv for the fused shape (R = G = 64, two taps), one pass over dafg:
//   dx[t]  = dout[t] + dafg[t] . W1(tap 1) + dafg[t+d] . W1(tap 0)       (K-major GEMM, K = 2 x 128, N = 64)
//   dW_f/g(o, c, tap) += sum_t dafg[t][o] * x[t - (1-tap) d][c]             (MN-major GEMM over the same rows)
// dafg is ONE fp16 plane (round-to-nearest of the gate derivative; it only ever is the dY operand of these two products and
// its rounding error is 2.5e-4 of the gradient norm, tests/dev/quant_sensitivity.py), so every product is TWO MMAs
// (dafg . B_hi + dafg . B_lo).  The ring streams "units" (row slab 0/1, gate half f/g): A = [128 positions x 64 channels] of
// dafg, B = the matching [64 c x 64 k] block of W1^T (hi, lo).  The two slab-0 units (da_f, da_g of the tile's own rows) sit
// in ring stages 2 and 3 and TOGETHER are the MN-major A operand of the weight gradient (M atom 0 = the 64 f channels,
// atom 1 = the 64 g channels, LBO = one ring stage): one M = 128 MMA per K step fills all 128 accumulator lanes (with one gate
// half per MMA and a zero tile as the other atom the tensor pipe did twice the work), B = [x(t-d) | x(t)].
struct SDxwArgs {
  const __half* rsd;       // dout of the layer above, split [rows][hi 64 | lo 64]; null for the top layer
  __half* Y;               // new dout, split
  float* dWf;              // (o, c, tap) with row stride 128
  float* dWg;
  int d;
  int rows_out, tiles_per_seq, num_tiles;
  int reverse;
  float wscale;            // 1 / (gscale * ACT_SCALE)
  int64_t det_stride;      // deterministic mode: floats between the per-CTA gradient slabs (0 = red.global.add)
};
// Shared-memory bandwidth (128 B/clk: TMA writes + MMA operand reads) is what bounds this kernel, so (1) W1^T (64 KB) is
// RESIDENT instead of travelling with every unit, (2) the hi and lo planes of a B operand are adjacent and consumed by ONE
// MMA of twice the N (dafg . [W_hi | W_lo] -> two 64-column halves the epilogue adds; dafg^T . [x_hi | x_lo] -> N = 256), so
// every A tile is read once per K step instead of twice: 352 KB of shared-memory traffic per tile instead of 512 KB.
constexpr int SD_W = 0;                         // W1^T resident: unit (slab, gate half) -> [hi 64 c x 64 k | lo] 16 KB
constexpr int SD_RING = 4 * 16384;
constexpr int SD_STAGE = SUB;                   // A [128 x 64] of dafg
constexpr int SD_STAGES = 4;
static_assert(SD_STAGES == 4, "a tile's four units must land in fixed stages (unit u in stage u)");
constexpr int SD_X = SD_RING + SD_STAGES * SD_STAGE;   // x(t-d) hi | x(t) hi | x(t-d) lo | x(t) lo
constexpr int SD_BAR = SD_X + 4 * SUB;
constexpr int SD_SMEM = SD_BAR + 256;

__global__ void __launch_bounds__(NTHREADS, 1)
tcs_dxw_kernel(const __grid_constant__ CUtensorMap tm_da, const __grid_constant__ CUtensorMap tm_w,
               const __grid_constant__ CUtensorMap tm_x, const SDxwArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bar0 = base + SD_BAR;
  auto full = [&](int s) { return bar0 + 8 * s; };
  auto empty = [&](int s) { return bar0 + 32 + 8 * s; };
  auto acc_full = [&](int s) { return bar0 + 64 + 8 * s; };
  auto acc_empty = [&](int s) { return bar0 + 80 + 8 * s; };
  const uint32_t x_full = bar0 + 96, x_empty = bar0 + 104, wg_full = bar0 + 112, w_full = bar0 + 120;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(gbase + SD_BAR + 128);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < SD_STAGES; ++s) {
      mbar_init(full(s), 1);
      mbar_init(empty(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(acc_full(s), 1);
      mbar_init(acc_empty(s), 8);      // one arrival per epilogue warp
    }
    mbar_init(x_full, 1);
    mbar_init(x_empty, 1);
    mbar_init(wg_full, 1);
    mbar_init(w_full, 1);
    fence_barrier_init();
    prefetch_tmap(&tm_da);
    prefetch_tmap(&tm_w);
    prefetch_tmap(&tm_x);
  }
  if (warp == 1) tmem_alloc<512>(smem_u32((const void*)tmem_slot));
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int n_local = (a.num_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  auto tile_of = [&](int j) {
    const int i = (int)blockIdx.x + j * (int)gridDim.x;
    return a.reverse ? a.num_tiles - 1 - i : i;
  };

  if (warp == 0) {
    if (lane == 0) {
      int it = 0;
      mbar_arrive_expect_tx(w_full, 4 * 16384);
      for (int u = 0; u < 4; ++u) {        // unit u = (slab u < 2 ? 1 : 0, gate half u & 1): W1^T[c][slab*128 + n], hi then lo
        const int sl = u < 2 ? 1 : 0, ch = u & 1;
        tma_load_2d(base + SD_W + u * 16384, &tm_w, w_full, sl * 128 + ch * KB, 0);
        tma_load_2d(base + SD_W + u * 16384 + 8192, &tm_w, w_full, 256 + sl * 128 + ch * KB, 0);
      }
      for (int j = 0; j < n_local; ++j) {
        const int tile = tile_of(j);
        const int b = tile / a.tiles_per_seq, t0 = (tile % a.tiles_per_seq) * TM;
        // Unit order: the two slab-1 units (dx only) first, the slab-0 units (dx + weight gradient) last -- the single x
        // buffer is released by the previous tile's LAST unit, so its reload (issued here after the first two unit loads,
        // rows already pulled into L2 one tile ahead) has two units of MMA work to hide behind.
        if (j + 1 < n_local) {
          const int tn = tile_of(j + 1);
          const int bn = tn / a.tiles_per_seq, tn0 = (tn % a.tiles_per_seq) * TM;
          tma_prefetch_4d(&tm_x, 0, tn0, bn, 0);
          tma_prefetch_4d(&tm_x, KB, tn0, bn, 0);
          if (a.d >= TM && tn0 >= a.d) {
            tma_prefetch_4d(&tm_x, 0, tn0 - a.d, bn, 0);
            tma_prefetch_4d(&tm_x, KB, tn0 - a.d, bn, 0);
          }
        }
        for (int u = 0; u < 4; ++u, ++it) {
          if (u == 2) {
            TRS(j, 12);
            mbar_wait(x_empty, (j & 1) ^ 1);
            TRS(j, 13);
            const uint32_t xs = base + SD_X;
            mbar_arrive_expect_tx(x_full, 4 * SUB);
            tma_load_4d(xs + 0 * SUB, &tm_x, x_full, 0, t0 - a.d, b, 0);
            tma_load_4d(xs + 1 * SUB, &tm_x, x_full, 0, t0, b, 0);
            tma_load_4d(xs + 2 * SUB, &tm_x, x_full, KB, t0 - a.d, b, 0);
            tma_load_4d(xs + 3 * SUB, &tm_x, x_full, KB, t0, b, 0);
          }
          const int sl = u < 2 ? 1 : 0, ch = u & 1;
          const int s = it % SD_STAGES, ph = (it / SD_STAGES) & 1;
          mbar_wait(empty(s), ph ^ 1);
          TRS(j, u);                        // producer: stage free, load of unit u issued
          const uint32_t st = base + SD_RING + s * SD_STAGE;
          mbar_arrive_expect_tx(full(s), SD_STAGE);
          tma_load_4d(st, &tm_da, full(s), ch * KB, t0 + sl * a.d, b, 0);                 // dafg, gate half ch
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && n_local > 0) {
      constexpr uint32_t idesc = idesc_f16(128, 128);                      // N = [W_hi 64 | W_lo 64]
      constexpr uint32_t idesc_mn = idesc_f16(128, 256) | IDESC_MN_MAJOR;   // N = [x(t-d) hi | x(t) hi | x(t-d) lo | x(t) lo]
      const uint32_t xs = base + SD_X;
      mbar_wait(w_full, 0);
      int it = 0;
      for (int j = 0; j < n_local; ++j) {
        const int ab = j & 1, aph = (j >> 1) & 1;
        for (int u = 0; u < 4; ++u, ++it) {
          const int s = it % SD_STAGES, ph = (it / SD_STAGES) & 1;
          if (u == 0) mbar_wait(acc_empty(ab), aph ^ 1);
          if (u == 2) {
            TRS(j, 14);
            mbar_wait(x_full, j & 1);
            TRS(j, 15);
          }
          mbar_wait(full(s), ph);
          TRS(j, 4 + u);                    // MMA thread: unit u loaded
          tcgen05_fence_after();
          const uint32_t st = base + SD_RING + s * SD_STAGE, wb = base + SD_W + u * 16384;
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4)   // columns [0, 64): dafg . W_hi, [64, 128): dafg . W_lo (separate fp32 accumulators)
            umma_f16(tmem + ab * 128, umma_desc_k_sw128(st + k4 * 32), umma_desc_k_sw128(wb + k4 * 32), idesc, (u | k4) > 0);
          if (u == 3) {
            // weight gradient of both gate halves: A atoms = the da_f tile (stage 2) and the da_g tile (this stage);
            // accumulator lanes = (f | g) channel, columns 256 + [plane 2][tap 2][c 64]
            const uint32_t sf = st - SD_STAGE;
#pragma unroll
            for (int k16 = 0; k16 < TM / 16; ++k16)
              umma_f16(tmem + 256, desc_mn_sw128(sf + k16 * 2048, SD_STAGE, 1024), desc_mn_sw128(xs + k16 * 2048, SUB, 1024), idesc_mn,
                       (j | k16) > 0);
            umma_commit(empty(s - 1));        // the da_f stage was held for the weight-gradient MMAs
            umma_commit(empty(s));
            umma_commit(x_empty);
            umma_commit(acc_full(ab));
            TRS(j, 8);                      // MMA thread: whole tile issued
          } else if (u != 2) {
            umma_commit(empty(s));
          }
        }
      }
      umma_commit(wg_full);
    }
  } else {
    const int q = warp & 3, half = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    for (int j = 0; j < n_local; ++j) {
      const int tile = tile_of(j);
      const int ab = j & 1, aph = (j >> 1) & 1;
      const int b = tile / a.tiles_per_seq, t = (tile % a.tiles_per_seq) * TM + row;
      const bool valid = t < a.rows_out;
      const int64_t orow = (int64_t)b * a.rows_out + min(t, a.rows_out - 1);
      uint32_t rh[2][8], rl[2][8];   // residual gradient (split): independent of the accumulator, loaded before waiting for it
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        if (a.rsd) {
          ld256(a.rsd + orow * 128 + half * 32 + c * 16, rh[c]);
          ld256(a.rsd + orow * 128 + 64 + half * 32 + c * 16, rl[c]);
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) rh[c][i] = rl[c][i] = 0u;   // top layer: nothing flows in from above
        }
      }
      mbar_wait(acc_full(ab), aph);
      if (threadIdx.x == 64) TRS(j, 9);     // epilogue: accumulator ready
      tcgen05_fence_after();
      uint32_t v[32], vl[32];
      tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + ab * 128 + half * 32, v);
      tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + ab * 128 + 64 + half * 32, vl);
      tmem_ld_wait();
      tcgen05_fence_before();
      warp_arrive(acc_empty(ab), lane);
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) + __uint_as_float(vl[i]));
      __half* yrow = a.Y + orow * 128 + half * 32;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t oh[8], ol[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float2 rhi = __half22float2(bits_h2(rh[c][i])), rlo = __half22float2(bits_h2(rl[c][i]));
          split2(fmaf(__uint_as_float(v[c * 16 + 2 * i]), INV_W, rhi.x + rlo.x),
                 fmaf(__uint_as_float(v[c * 16 + 2 * i + 1]), INV_W, rhi.y + rlo.y), oh[i], ol[i]);
        }
        if (valid) {
          st256(yrow + c * 16, oh);
          st256(yrow + 64 + c * 16, ol);
        }
      }
      if (threadIdx.x == 64) TRS(j, 10);    // epilogue: tile stored
    }
    if (n_local > 0) {
      // dW_f / dW_g: TMEM lanes 0..63 = f gate channel o, lanes 64..127 = g gate channel o, columns tap*64 + c; the
      // gradient layout is (o, c, tap)
      mbar_wait(wg_full, 0);
      tcgen05_fence_after();
      uint32_t v0[32], v1[32];
      {
        uint32_t t0[32];
        tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + 256 + half * 32, v0);              // tap 0, x hi plane, channels half*32..
        tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + 256 + 128 + half * 32, t0);        // tap 0, x lo plane
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) v0[i] = __float_as_uint(__uint_as_float(v0[i]) + __uint_as_float(t0[i]));
        tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + 256 + 64 + half * 32, v1);         // tap 1
        tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + 256 + 192 + half * 32, t0);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) v1[i] = __float_as_uint(__uint_as_float(v1[i]) + __uint_as_float(t0[i]));
      }
      float* wrow = (q < 2 ? a.dWf : a.dWg) + (int64_t)blockIdx.x * a.det_stride + (int64_t)(row & 63) * 128 + half * 64;
#pragma unroll
      for (int i = 0; i < 16; ++i)
        wg_add_v4(wrow + 4 * i, make_float4(__uint_as_float(v0[2 * i]) * a.wscale, __uint_as_float(v1[2 * i]) * a.wscale,
                                            __uint_as_float(v0[2 * i + 1]) * a.wscale, __uint_as_float(v1[2 * i + 1]) * a.wscale),
                  a.det_stride != 0);
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

// fp16 tensor [d3][d2][d1][d0] (d0 contiguous, strides in elements), box [1][1][box1][64], SWIZZLE_128B, zero fill
int map_h4d(CUtensorMap* m, const __half* ptr, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t d3, uint64_t s1, uint64_t s2,
            uint64_t s3, uint32_t box1) {
  EncodeTiledFn enc = get_encode();
  WN_REQUIRE(enc, WN_ECUDA, "cuTensorMapEncodeTiled is unavailable");
  cuuint64_t dims[4] = {d0, d1, d2, d3};
  cuuint64_t strides[3] = {s1 * 2, s2 * 2, s3 * 2};
  cuuint32_t box[4] = {KB, box1, 1, 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, (void*)ptr, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  WN_REQUIRE(r == CUDA_SUCCESS, WN_ECUDA, "cuTensorMapEncodeTiled(f16 4d) failed: %d", (int)r);
  return WN_OK;
}

int map_h2d(CUtensorMap* m, const __half* ptr, uint64_t d0, uint64_t d1, uint64_t s1, uint32_t box1) {
  EncodeTiledFn enc = get_encode();
  WN_REQUIRE(enc, WN_ECUDA, "cuTensorMapEncodeTiled is unavailable");
  cuuint64_t dims[2] = {d0, d1};
  cuuint64_t strides[1] = {s1 * 2};
  cuuint32_t box[2] = {KB, box1};
  cuuint32_t es[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (void*)ptr, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  WN_REQUIRE(r == CUDA_SUCCESS, WN_ECUDA, "cuTensorMapEncodeTiled(f16 2d) failed: %d", (int)r);
  return WN_OK;
}

template <int BN, int MODE>
int launch_sgemm_mode(const CUtensorMap& ta, const CUtensorMap& tb, const SGemmArgs& g, int sm_count, cudaStream_t s) {
  using Cfg = SGemmCfg<BN>;
  static bool attr = false;
  if (!attr) {
    WN_CHECK_CUDA(cudaFuncSetAttribute(tcs_gemm_kernel<BN, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
    attr = true;
  }
  const int grid = g.num_tiles < sm_count ? g.num_tiles : sm_count;
  tcs_gemm_kernel<BN, MODE><<<grid, NTHREADS, Cfg::SMEM, s>>>(ta, tb, g);
  WN_CHECK_LAUNCH();
  return WN_OK;
}

template <int BN>
int launch_sgemm(const CUtensorMap& ta, const CUtensorMap& tb, const SGemmArgs& g, int sm_count, cudaStream_t s) {
  if (g.gate_sg) return launch_sgemm_mode<BN, 1>(ta, tb, g, sm_count, s);
  if (g.ce_target) {
    if constexpr (BN == 256) return launch_sgemm_mode<256, 6>(ta, tb, g, sm_count, s);
    wn_set_error("fused cross-entropy epilogue needs 256 classes");
    return WN_EINVAL;
  }
  if (g.flush) return launch_sgemm_mode<BN, 4>(ta, tb, g, sm_count, s);
  return launch_sgemm_mode<BN, 5>(ta, tb, g, sm_count, s);
}

template <int NB, int MH>
int launch_swgrad(const CUtensorMap& ta, const CUtensorMap& tb, const SWgradArgs& g, int sm_count, cudaStream_t s) {
  using Cfg = SWgradCfg<NB, MH>;
  static bool attr = false;
  if (!attr) {
    WN_CHECK_CUDA(cudaFuncSetAttribute(tcs_wgrad_kernel<NB, MH>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
    attr = true;
  }
  int grid = g.num_chunks * g.ngroups < sm_count ? g.num_chunks * g.ngroups : sm_count;
  grid -= grid % g.ngroups;
  tcs_wgrad_kernel<NB, MH><<<grid, NTHREADS, Cfg::SMEM, s>>>(ta, tb, g);
  WN_CHECK_LAUNCH();
  return WN_OK;
}

// One split operand of a slab GEMM: rows of sequence b are A[slab][b][row_off + t][hi(C) | lo(C)]
struct SOperand {
  const __half* ptr;
  int C;            // channels per row (row = 2C halves)
  int rows_in;      // rows per sequence in memory
  int num_seq;
  int nslab;
  int64_t slab_stride;   // in halves
  int planes = 2;        // 2: rows are [hi(C) | lo(C)]; 1: one fp16 plane [C] (gradient tensors of the fused backward)
};

struct SEpilogue {
  const float* bias = nullptr;
  int relu = 0;
  float acc_scale = 1.f, out_scale = 1.f, rsd_scale = 1.f;
  const void* Rsd = nullptr;
  int ldr = 0, rsd_split = 0;
  const __half* mask = nullptr;
  int ldm = 0, mask_rows_in = 0, mask_row_off = 0;
  int y_slab_cols = 0;
  int64_t y_slab_stride = 0;
  int out_half = 0;
  const float* gate_sg = nullptr;
  const __half* gate_z = nullptr;
  __half* gate_dafg = nullptr;
  int gate_zp = 0, gate_sg_ld = 0;
  int zero_rows_below = 0;
  int reverse = 0;
  int flush = 0;
  float* colsum_out = nullptr;
  float colsum_scale = 1.f;
  const int32_t* ce_target = nullptr;
  double* ce_acc = nullptr;
  __half* ce_dlogits = nullptr;
  float ce_invn = 0.f, ce_gscale = 1.f;
  bool ngroups_ok_for_colsum(int N, int BN) const { return N <= BN; }   // the per-CTA column-sum table covers one column group
};

inline __half* HP(float* p) { return reinterpret_cast<__half*>(p); }
inline const __half* HP(const float* p) { return reinterpret_cast<const __half*>(p); }

// Y[(b, t)][0..N) = epi( sum_s A[slab_idx[s]][b][t + row_off[s]][0..K) . Wt[:, s*K ..]^T ); Wt is split [N][2 * ns*K]
int tcs_gemm(const wn_handle* h, const SOperand& A, int ns, const int* slab_idx, const int* row_off, int rows_out,
             const __half* Wt, int N, const SEpilogue& e, void* Y, int ldy, int out_split, cudaStream_t s) {
  const bool plain = !e.gate_sg && !e.Rsd && !e.mask;
  WN_REQUIRE(A.C % KB == 0 && N % 64 == 0 && (N <= 256 || plain) && ns <= MAX_SLABS && (e.y_slab_cols == 0 || !out_split) &&
                 !(e.out_half && out_split),
             WN_EINVAL, "tcs_gemm: unsupported shape K=%d N=%d slabs=%d", A.C, N, ns);
  CUtensorMap ta, tb;
  const int64_t rowe = A.planes * (int64_t)A.C;
  const int64_t sstride = A.nslab > 1 ? A.slab_stride : (int64_t)A.rows_in * A.num_seq * rowe;
  WN_TRY(map_h4d(&ta, A.ptr, rowe, A.rows_in, A.num_seq, A.nslab, rowe, (uint64_t)A.rows_in * rowe, sstride, TM));
  const int BN = N <= 64 ? 64 : (N <= 128 ? 128 : 256);
  const int64_t Ktot = (int64_t)ns * A.C;
  WN_TRY(map_h2d(&tb, Wt, 2 * Ktot, N, 2 * Ktot, BN));
  SGemmArgs g;
  memset(&g, 0, sizeof(g));
  g.Y = Y;
  g.ldy = ldy;
  g.out_split = out_split;
  g.out_half = e.out_half;
  g.bias = e.bias;
  g.N = N;
  g.relu = e.relu;
  g.acc_scale = e.acc_scale;
  g.out_scale = e.out_scale;
  g.rsd_scale = e.rsd_scale;
  g.Rsd = e.Rsd;
  g.ldr = e.ldr;
  g.rsd_split = e.rsd_split;
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
  g.zero_rows_below = e.zero_rows_below;
  g.reverse = e.reverse;
  g.flush = e.flush && !e.gate_sg && !e.colsum_out;
  // Two K blocks (128 channels) per flush group when the block count is even: reading a 128 x 256 fp32 accumulator out of
  // TMEM takes 2048 cycles (64 B/clk) against 1536 cycles of MMAs per K block, so one flush per block made the tensor pipe
  // wait for tcgen05.ld; the truncating accumulation over 128 instead of 64 channels costs < 1e-6 relative
  if (g.flush && (ns * (A.C / KB)) % 2 == 0) g.flush = 2;
  WN_REQUIRE(A.planes == 2 || (!g.flush && !e.gate_sg && !e.ce_target), WN_EINVAL, "tcs_gemm: single-plane A is a backward operand");
  g.colsum_out = e.ngroups_ok_for_colsum(N, BN) ? e.colsum_out : nullptr;
  if (g.colsum_out) g.colsum_out = wn_det_ptr(h, g.colsum_out);
  g.det_stride = wn_det_stride(h);
  g.ce_target = e.ce_target;
  g.ce_acc = e.ce_acc;
  g.ce_dlogits = e.ce_dlogits;
  g.ce_invn = e.ce_invn;
  g.ce_gscale = e.ce_gscale;
  g.colsum_scale = e.colsum_scale;
  g.rows_out = rows_out;
  g.nslab = ns;
  g.kblk = A.C / KB;
  g.a_planes = A.planes;
  g.a_plane = A.C;
  g.b_plane = (int)Ktot;
  for (int i = 0; i < ns; ++i) {
    g.slab_row_off[i] = row_off ? row_off[i] : 0;
    g.slab_idx[i] = slab_idx ? slab_idx[i] : 0;
  }
  g.tiles_per_seq = (rows_out + TM - 1) / TM;
  g.ngroups = (N + BN - 1) / BN;
  g.num_tiles = g.tiles_per_seq * A.num_seq * g.ngroups;
  if (BN == 64) return launch_sgemm<64>(ta, tb, g, h->sm_count, s);
  if (BN == 128) return launch_sgemm<128>(ta, tb, g, h->sm_count, s);
  return launch_sgemm<256>(ta, tb, g, h->sm_count, s);
}

// dW[slab](m, c) += scale * sum_{b,t} dY[b][a_row_off + t][a_c0 + m] * X[slab_idx[s]][b][b_row_off[s] + t][c]
int tcs_wgrad(const wn_handle* h, const SOperand& dY, int a_row_off, int a_c0, int m_valid, const SOperand& X, int nb_slab,
              const int* b_row_off, const int* b_slab_idx, float* const* dW0, float* const* dW1, int m_split, int rows_it,
              int64_t sn, int64_t sk, float scale, cudaStream_t s, int ngroups = 1, int reverse = 0) {
  const int NB = nb_slab * X.C;
  WN_REQUIRE(X.C % KB == 0 && dY.C % KB == 0 && a_c0 % KB == 0 && (NB == 64 || NB == 128 || NB == 256) && nb_slab <= 4 &&
                 m_valid <= 256 && ngroups >= 1 && ngroups <= 8 && (ngroups == 1 || !dW1),
             WN_EINVAL, "tcs_wgrad: unsupported shape X.C=%d slabs=%d M=%d groups=%d", X.C, nb_slab, m_valid, ngroups);
  const int MH = m_valid > 128 ? 2 : 1;
  // (the TMA boxes of dY cover 128*MH channels from a_c0; boxes past channel dY.C fetch lo-plane data or zeros into
  // accumulator rows >= m_valid, which the epilogue never stores)
  constexpr int WG_KC = 32;
  CUtensorMap ta, tb;
  const int64_t rowa = dY.planes * (int64_t)dY.C, rowx = 2 * (int64_t)X.C;
  WN_TRY(map_h4d(&ta, dY.ptr, rowa, dY.rows_in, dY.num_seq, 1, rowa, (uint64_t)dY.rows_in * rowa,
                 (uint64_t)dY.rows_in * dY.num_seq * rowa, WG_KC));
  const int64_t xstride = X.nslab > 1 ? X.slab_stride : (int64_t)X.rows_in * X.num_seq * rowx;
  WN_TRY(map_h4d(&tb, X.ptr, rowx, X.rows_in, X.num_seq, X.nslab, rowx, (uint64_t)X.rows_in * rowx, xstride, WG_KC));
  SWgradArgs g;
  memset(&g, 0, sizeof(g));
  g.rows_it = rows_it;
  g.num_seq = dY.num_seq;
  g.a_row_off = a_row_off;
  g.a_c0 = a_c0;
  g.a_plane = dY.C;
  g.a_planes = dY.planes;
  g.nb_slab = nb_slab;
  g.nb_sub = X.C / KB;
  g.b_plane = X.C;
  g.ngroups = ngroups;
  g.reverse = reverse;
  g.scale = scale;
  g.det_stride = wn_det_stride(h);
  {
    static const int run_env = getenv("WN_WG_RUN") ? atoi(getenv("WN_WG_RUN")) : 8;
    g.run = run_env;
  }
  bool aligned = sn % 4 == 0;
  for (int i = 0; i < nb_slab; ++i) {
    g.b_row_off[i] = b_row_off[i];
    g.dW1[i] = dW1 && dW1[i] ? wn_det_ptr(h, dW1[i]) : nullptr;
    if (g.dW1[i]) aligned = aligned && ((uintptr_t)g.dW1[i] & 15) == 0;
    for (int gr = 0; gr < ngroups; ++gr) {
      g.b_slab_idx[gr][i] = b_slab_idx ? b_slab_idx[gr * nb_slab + i] : 0;
      g.dW0[gr][i] = dW0[gr * nb_slab + i] ? wn_det_ptr(h, dW0[gr * nb_slab + i]) : nullptr;
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
    const bool taps2 = ngroups == 1 && nb_slab == 2 && sk == 2 && X.C % 64 == 0 && g.dW0[0][1] == g.dW0[0][0] + 1 &&
                       (!g.dW1[0] || g.dW1[1] == g.dW1[0] + 1) && sn % 4 == 0 && ((uintptr_t)g.dW0[0][0] & 15) == 0 &&
                       (!g.dW1[0] || ((uintptr_t)g.dW1[0] & 15) == 0);
    g.red_mode = (sk == 1 && aligned) ? 1 : (taps2 ? 2 : 0);
  }
  if (MH == 1) {
    if (NB == 64) return launch_swgrad<64, 1>(ta, tb, g, h->sm_count, s);
    if (NB == 128) return launch_swgrad<128, 1>(ta, tb, g, h->sm_count, s);
    return launch_swgrad<256, 1>(ta, tb, g, h->sm_count, s);
  }
  if (NB == 64) return launch_swgrad<64, 2>(ta, tb, g, h->sm_count, s);
  if (NB == 128) return launch_swgrad<128, 2>(ta, tb, g, h->sm_count, s);
  return launch_swgrad<256, 2>(ta, tb, g, h->sm_count, s);
}

// dz GEMM + gate derivative + dWp in one pass (fused shape R = G = 64); dWp must be 16-byte aligned
int tcs_gate_bwd(const wn_handle* h, const __half* dout, const __half* wpt, const __half* dzs, const __half* z, const uint16_t* sg,
                 __half* dafg, float* dWp, int zp, int rows, int num_seq, int reverse, float wscale, cudaStream_t s) {
  CUtensorMap td, tz, tw;
  const uint64_t seq = (uint64_t)rows * 128, all = seq * num_seq;
  WN_TRY(map_h4d(&td, dout, 128, rows, num_seq, 1, 128, seq, all, TM));
  WN_TRY(map_h4d(&tz, z, 128, rows, num_seq, 1, 128, seq, all, TM));
  WN_TRY(map_h2d(&tw, wpt, 128, 64, 128, 64));
  SGateBwdArgs g;
  memset(&g, 0, sizeof(g));
  g.dzs = dzs;
  g.sg = sg;
  g.dafg = dafg;
  g.dWp = wn_det_ptr(h, dWp);
  g.det_stride = wn_det_stride(h);
  g.zp = zp;
  g.reverse = reverse;
  g.wscale = wscale;
  g.rows_out = rows;
  g.tiles_per_seq = (rows + TM - 1) / TM;
  g.num_tiles = g.tiles_per_seq * num_seq;
  static bool attr = false;
  if (!attr) {
    WN_CHECK_CUDA(cudaFuncSetAttribute(tcs_gate_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SG_SMEM + 1024));
    attr = true;
  }
  const int grid = g.num_tiles < h->sm_count ? g.num_tiles : h->sm_count;
  tcs_gate_bwd_kernel<<<grid, NTHREADS, SG_SMEM + 1024, s>>>(td, tz, tw, g);
  WN_CHECK_LAUNCH();
  return WN_OK;
}

// dx GEMM + dW_f/dW_g in one pass over dafg (fused shape R = G = 64, two taps); gradient pointers 16-byte aligned
int tcs_dxw(const wn_handle* h, const __half* dafg, const __half* w1t, const __half* rsd, const __half* x, __half* Y, float* dWf,
            float* dWg, int d, int rows, int num_seq, int reverse, float wscale, cudaStream_t s) {
  CUtensorMap ta, tw, tx;
  const uint64_t seqa = (uint64_t)rows * 128, alla = seqa * num_seq, seqx = (uint64_t)rows * 128, allx = seqx * num_seq;
  WN_TRY(map_h4d(&ta, dafg, 128, rows, num_seq, 1, 128, seqa, alla, TM));
  WN_TRY(map_h2d(&tw, w1t, 512, 64, 512, 64));
  WN_TRY(map_h4d(&tx, x, 128, rows, num_seq, 1, 128, seqx, allx, TM));
  SDxwArgs g;
  memset(&g, 0, sizeof(g));
  g.rsd = rsd;
  g.Y = Y;
  g.dWf = wn_det_ptr(h, dWf);
  g.dWg = wn_det_ptr(h, dWg);
  g.det_stride = wn_det_stride(h);
  g.d = d;
  g.reverse = reverse;
  g.wscale = wscale;
  g.rows_out = rows;
  g.tiles_per_seq = (rows + TM - 1) / TM;
  g.num_tiles = g.tiles_per_seq * num_seq;
  static bool attr = false;
  if (!attr) {
    WN_CHECK_CUDA(cudaFuncSetAttribute(tcs_dxw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SD_SMEM + 1024));
    attr = true;
  }
  const int grid = g.num_tiles < h->sm_count ? g.num_tiles : h->sm_count;
  tcs_dxw_kernel<<<grid, NTHREADS, SD_SMEM + 1024, s>>>(ta, tw, tx, g);
  WN_CHECK_LAUNCH();
  return WN_OK;
}

inline unsigned nblk(int64_t n, int bs) { return (unsigned)((n + bs - 1) / bs); }

bool fused_shape(const wn_handle* h) {
  if (h->R != 64 || h->cfg.residual_filter_width != 2) return false;
  for (const ResLayer& l : h->layers)
    if (l.G != 64) return false;
  return true;
}

int tcs_prepare_weights(wn_handle* h, const float* params, cudaStream_t s) {
  const Tape& t = h->tape;
  const int L = (int)h->layers.size();
  if (!h->tc_tab_uploaded) {
    std::vector<TcsTabEntry> tab(L);
    for (int l = 0; l < L; ++l) {
      tab[l].wf = h->layers[l].wf.w_off;
      tab[l].wg = h->layers[l].wg.w_off;
      tab[l].wp = h->layers[l].proj.w_off;
      tab[l].ws = h->layers[l].skip.w_off;
    }
    WN_CHECK_CUDA(cudaMemcpyAsync(h->ws + t.tc_tab, tab.data(), sizeof(TcsTabEntry) * L, cudaMemcpyHostToDevice, s));
    WN_CHECK_CUDA(cudaStreamSynchronize(s));   // tab is a stack temporary
    h->tc_tab_uploaded = true;
  }
  dim3 grid(16, L);
  tcs_prep_kernel<<<grid, 256, 0, s>>>(params, (const TcsTabEntry*)(h->ws + t.tc_tab), HP(h->ws + t.tc_w1), HP(h->ws + t.tc_w2),
                                       HP(h->ws + t.tc_ws), HP(h->ws + t.tc_w1t), HP(h->ws + t.tc_wpt), HP(h->ws + t.tc_wst), L,
                                       h->R, h->layers[0].G, h->S, 2);
  WN_CHECK_LAUNCH();
  return WN_OK;
}

int tcs_layer_launch(wn_handle* h, int l, cudaStream_t s) {
  const Tape& t = h->tape;
  const int tiles_per_seq = (t.W + TM - 1) / TM;
  const int num_tiles = tiles_per_seq * t.B;
  const int grid = num_tiles < h->sm_count ? num_tiles : h->sm_count;
  const ResLayer& ly = h->layers[l];
  CUtensorMap tx, tw1, tw2, tz, tsg;
  WN_TRY(map_h4d(&tx, HP(h->ws + t.x[l]), 128, t.W, t.B, 1, 128, (uint64_t)t.W * 128, (uint64_t)t.P * 128, TM));
  WN_TRY(map_h4d(&tz, HP(h->ws + t.z[l]), 128, t.W, t.B, 1, 128, (uint64_t)t.W * 128, (uint64_t)t.P * 128, TM));
  WN_TRY(map_h4d(&tsg, HP(h->ws + t.tfsg[l]), 64, t.W, t.B, 1, 64, (uint64_t)t.W * 64, (uint64_t)t.P * 64, TM));   // u16 rows
  WN_TRY(map_h2d(&tw1, HP(h->ws + t.tc_w1) + (int64_t)l * 2 * 128 * 128, 256, 128, 256, 128));
  WN_TRY(map_h2d(&tw2, HP(h->ws + t.tc_w2) + (int64_t)l * 2 * 64 * 64, 128, 64, 128, 64));
  SLayerArgs a;
  a.x_out = HP(h->ws + t.x[l + 1]);
  a.W = t.W;
  a.d = ly.dilation;
  a.zp = wn_zero_prefix(t.W, ly.dilation, 2);
  a.tiles_per_seq = tiles_per_seq;
  a.num_tiles = num_tiles;
  a.reverse = getenv("WN_NO_SERP") ? 0 : (l & 1);
  static bool attr = false;
  if (!attr) {
    WN_CHECK_CUDA(cudaFuncSetAttribute(tcs_layer_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SL_SMEM + 1024));
    attr = true;
  }
  tcs_layer_kernel<<<grid, SL_THREADS, SL_SMEM + 1024, s>>>(tx, tw1, tw2, tz, tsg, a);
  WN_CHECK_LAUNCH();
  return WN_OK;
}

int tcs_skip_gemm(wn_handle* h, cudaStream_t s) {
  const Tape& t = h->tape;
  const int L = (int)h->layers.size();
  const int G = h->layers[0].G;
  const int64_t zstride = L > 1 ? t.z[1] - t.z[0] : 0;   // floats
  for (int l = 1; l < L; ++l)
    WN_REQUIRE(t.z[l] - t.z[l - 1] == zstride, WN_EINVAL, "z slabs are not equally spaced");
  int idx[MAX_SLABS], off[MAX_SLABS];
  for (int l = 0; l < L; ++l) {
    idx[l] = l;
    off[l] = 0;
  }
  SOperand A{HP(h->ws + t.z[0]), G, t.W, t.B, L, 2 * zstride};
  SEpilogue e;
  e.relu = h->fuse_head_relu;        // fused train path: the head's first ReLU (wavenet.py:588) rides on this epilogue
  e.acc_scale = INV_ACT_W;
  e.out_scale = ACT_SCALE;
  e.flush = 1;
  h->skip_is_relu = h->fuse_head_relu != 0;
  return tcs_gemm(h, A, L, idx, off, t.W, HP(h->ws + t.tc_ws), h->S, e, h->ws + t.skip, h->S, 1, s);
}

}  // namespace

// ---- entry points used by wn_api.cu -----------------------------------------------------------------------------
// Networks whose channel counts are multiples of 64 (one 128-byte row of fp16 per 64 channels): the fused layer kernel
// for R = G = 64 (BASELINE config C), slab GEMMs + SIMT gate otherwise (e.g. the reference default R256/G128).
float tcs_act_scale() { return ACT_SCALE; }

bool tcs_supported(const wn_handle* h) {
  if (h->cfg.residual_filter_width != 2 || h->R % 64 != 0 || h->R > 256 || h->S % 64 != 0 || h->S > 256) return false;
  const int G = h->layers[0].G;
  if (G % 64 != 0 || 2 * G > 256 || (int)h->layers.size() > MAX_SLABS) return false;
  for (const ResLayer& l : h->layers)
    if (l.G != G || l.wf.b_off >= 0 || l.proj.b_off >= 0 || l.skip.b_off >= 0) return false;
  for (const ConvParam& c : h->head)
    if (c.in_ch % 64 != 0 || c.in_ch > 256 || c.out_ch % 64 != 0 || c.out_ch > 256) return false;
  return get_encode() != nullptr;
}

// fp32 rows -> split rows in place, scaled (x[0] written by the embedding kernel; dlogits written by the CE kernel)
int tcs_split_rows_inplace(float* x, int C, int64_t rows, float scale, cudaStream_t s) {
  const int c4 = C / 4;
  WN_REQUIRE(C % 4 == 0 && c4 <= 1024 && rows < (1LL << 30), WN_EINVAL, "tcs_split_rows_inplace: bad shape C=%d", C);
  int bs = 256 - 256 % c4;      // whole rows per block (see the kernel)
  if (bs < c4) bs = c4;
  tcs_rows_to_split_kernel<<<nblk(rows * c4, bs), bs, 0, s>>>(x, 1 << 30, 0, HP(x), C, 1 << 30, rows, 0, scale);
  WN_CHECK_LAUNCH();
  return WN_OK;
}

int tcs_unsplit_rows(const float* src_split, float* dst, int C, int64_t rows, float scale, cudaStream_t s) {
  tcs_split_to_rows_kernel<<<nblk(rows * (C / 4), 256), 256, 0, s>>>(HP(src_split), dst, C, rows, scale);
  WN_CHECK_LAUNCH();
  return WN_OK;
}

// external fp32 head input [B*T][S] -> ReLU -> split rows at the start of the skip buffer
int tcs_import_head_input(wn_handle* h, const float* in, int T, cudaStream_t s) {
  const Tape& t = h->tape;
  const int64_t rows = (int64_t)t.B * T;
  tcs_rows_to_split_kernel<<<nblk(rows * (h->S / 4), 256), 256, 0, s>>>(in, T, 0, HP(h->ws + t.skip), h->S, T, rows, 1, ACT_SCALE);
  WN_CHECK_LAUNCH();
  return WN_OK;
}

int tcs_forward_residual(wn_handle* h, const float* params, cudaStream_t s) {
  const Tape& t = h->tape;
  const int L = (int)h->layers.size();
  WN_TRY(tcs_prepare_weights(h, params, s));
  if (fused_shape(h)) {
    for (int l = 0; l < L; ++l) WN_TRY(tcs_layer_launch(h, l, s));
    h->tape_gates_zs = true;     // tape keeps (z split, sigmoid 16-bit fixed point [P][G])
  } else {
    const int R = h->R, G = h->layers[0].G;
    for (int l = 0; l < L; ++l) {
      const ResLayer& ly = h->layers[l];
      SOperand X{HP(h->ws + t.x[l]), R, t.W, t.B, 1, 0};
      const int sidx[2] = {0, 0};
      const int roff[2] = {-ly.dilation, 0};
      SEpilogue e;
      e.zero_rows_below = wn_zero_prefix(t.W, ly.dilation, 2);
      e.acc_scale = INV_ACT_W;
      e.flush = 1;
      WN_TRY(tcs_gemm(h, X, 2, sidx, roff, t.W, HP(h->ws + t.tc_w1) + (int64_t)l * 2 * (2 * G * 2 * R), 2 * G, e,
                      h->ws + t.tfsg[l], 2 * G, 0, s));
      tcs_gate_forward_kernel<<<nblk(t.P * (G / 4), 256), 256, 0, s>>>(h->ws + t.tfsg[l], HP(h->ws + t.z[l]), t.P, G);
      WN_CHECK_LAUNCH();
      SOperand Z{HP(h->ws + t.z[l]), G, t.W, t.B, 1, 0};
      const int zero = 0;
      SEpilogue e2;
      e2.Rsd = h->ws + t.x[l];
      e2.ldr = R;
      e2.rsd_split = 1;
      e2.acc_scale = INV_ACT_W;
      e2.rsd_scale = INV_ACT;
      e2.out_scale = ACT_SCALE;
      e2.flush = 1;
      WN_TRY(tcs_gemm(h, Z, 1, nullptr, &zero, t.W, HP(h->ws + t.tc_w2) + (int64_t)l * 2 * R * G, R, e2, h->ws + t.x[l + 1], R, 1,
                      s));
    }
    h->tape_gates_zs = false;    // tape keeps (tanh | sigmoid) fp32 with row stride 2G, z split
  }
  return tcs_skip_gemm(h, s);
}

// ReLU -> 1x1 conv per head layer (wavenet.py:587-590).  Stored activations are post-ReLU split rows; the last conv
// writes fp32 logits.
int tcs_forward_head(wn_handle* h, const float* params, int T, bool external, cudaStream_t s) {
  const Tape& t = h->tape;
  const int nh = (int)h->head.size();
  const int64_t rows = (int64_t)t.B * T;
  const int rows_in0 = external ? T : t.W, off0 = external ? 0 : t.W - T;
  if (!h->skip_is_relu && !external) {
    tcs_relu_split_rows_kernel<<<nblk(rows * (h->S / 4), 256), 256, 0, s>>>(HP(h->ws + t.skip), h->S, rows_in0, off0, T, rows);
    WN_CHECK_LAUNCH();
  }
  for (int i = 0; i < nh; ++i) {
    const ConvParam& cp = h->head[i];
    const int64_t n = (int64_t)cp.out_ch * cp.in_ch;
    tcs_split_weight_kernel<<<nblk(n, 256), 256, 0, s>>>(params + cp.w_off, HP(h->ws + t.tc_wh[i]), HP(h->ws + t.tc_wht[i]),
                                                       cp.out_ch, cp.in_ch);
    WN_CHECK_LAUNCH();
    const bool last = i == nh - 1;
    SOperand A{i == 0 ? HP(h->ws + t.skip) : HP(h->ws + t.hbuf[i - 1]), cp.in_ch, i == 0 ? rows_in0 : T, t.B, 1, 0};
    const int off = i == 0 ? off0 : 0;
    SEpilogue e;
    e.bias = cp.b_off >= 0 ? params + cp.b_off : nullptr;
    e.relu = last ? 0 : 1;
    e.acc_scale = INV_ACT_W;
    e.out_scale = last ? 1.f : ACT_SCALE;
    e.flush = 1;
    if (last && h->fuse_ce_target && cp.out_ch == 256) {
      // wn_forward_loss: softmax cross-entropy rides on this GEMM's epilogue; logits are only written on request
      e.ce_target = h->fuse_ce_target;
      e.ce_acc = (double*)(h->ws + t.loss_acc);
      e.ce_dlogits = HP(h->ws + t.dlogits);
      e.ce_invn = 1.f / (float)rows;
      e.ce_gscale = h->gscale;
      WN_CHECK_CUDA(cudaMemsetAsync(e.ce_acc, 0, 2 * sizeof(double), s));
      WN_TRY(tcs_gemm(h, A, 1, nullptr, &off, T, HP(h->ws + t.tc_wh[i]), cp.out_ch, e, h->fuse_ce_logits, cp.out_ch, 0, s));
      h->ce_fused_done = true;
      h->dlogits_single = true;
      continue;
    }
    WN_TRY(tcs_gemm(h, A, 1, nullptr, &off, T, HP(h->ws + t.tc_wh[i]), cp.out_ch, e, h->ws + t.hbuf[i], cp.out_ch, last ? 0 : 1, s));
  }
  return WN_OK;
}

static int tcs_embed_backward_det(wn_handle* h, const __half* dout_split, float* grads, float inv, cudaStream_t s);

// Backward of head + residual stack; requires a tape written by tcs_forward_residual and tcs_forward_head and dlogits in
// split format scaled by h->gscale.
int tcs_backward(wn_handle* h, const float* params, float* grads, cudaStream_t s) {
  const Tape& t = h->tape;
  const int T = h->T, W = t.W, B = t.B;
  const int64_t P = t.P;
  const int nh = (int)h->head.size();
  const int L = (int)h->layers.size();
  const int R = h->R, G = h->layers[0].G, S = h->S;
  float* ws = h->ws;
  const float inv = 1.f / h->gscale;          // gradient tensors carry gscale
  const float inv_wg = inv * INV_ACT;         // weight gradients: dY (gscale) x activation (ACT_SCALE)
  const int zero = 0;
  // ---- head ----
  // head gradients: one fp16 plane when dlogits came from the fused cross-entropy epilogue (the train step), split otherwise
  const int dpl = h->dlogits_single ? 1 : 2;
  const __half* d = HP(ws + t.dlogits);
  int tog = 0;
  bool bias_done = false;
  for (int i = nh - 1; i >= 0; --i) {
    const ConvParam& cp = h->head[i];
    const bool first = i == 0;
    const int rin = first ? (h->head_external ? T : W) : T;
    const int aoff = first ? (h->head_external ? 0 : W - T) : 0;
    const __half* Ain = first ? HP(ws + t.skip) : HP(ws + t.hbuf[i - 1]);
    SOperand dY{d, cp.out_ch, T, B, 1, 0, dpl};
    SOperand X{Ain, cp.in_ch, rin, B, 1, 0};
    for (int m0 = 0; m0 < cp.out_ch; m0 += 256) {
      const int mv = cp.out_ch - m0 < 256 ? cp.out_ch - m0 : 256;
      float* dw = grads + cp.w_off + (int64_t)m0 * cp.in_ch;
      WN_TRY(tcs_wgrad(h, dY, 0, m0, mv, X, 1, &aoff, nullptr, &dw, nullptr, 256, T, cp.in_ch, 1, inv_wg, s));
    }
    if (cp.b_off >= 0 && !bias_done) {
      if (i == nh - 1 && h->ce_colsum_valid) {
        WN_TRY(simt_add_vec(ws + t.ce_colsum, grads + cp.b_off, cp.out_ch, s));   // unscaled column sums from the CE kernel
      } else {
        WN_REQUIRE(cp.out_ch % 2 == 0, WN_EINVAL, "tcs_colsum: odd channel count");
        dim3 grid(nblk(cp.out_ch, 64), 4 * h->sm_count / (int)nblk(cp.out_ch, 64) + 1), block(32, 8);
        if (h->deterministic) grid.y = h->det_nslab;
        tcs_colsum_kernel<<<grid, block, 0, s>>>(d, dpl, (int64_t)B * T, cp.out_ch, inv, wn_det_ptr(h, grads + cp.b_off),
                                                 wn_det_stride(h));
        WN_CHECK_LAUNCH();
      }
    }
    bias_done = false;
    if (first && h->head_external) return WN_OK;
    // d_prev = (d . W) masked by the stored (post-ReLU) input
    SEpilogue e;
    e.acc_scale = INV_W;
    e.mask = Ain;
    e.ldm = cp.in_ch;
    e.mask_rows_in = rin;
    e.mask_row_off = aoff;
    if (i > 0 && h->head[i - 1].b_off >= 0) {
      e.colsum_out = grads + h->head[i - 1].b_off;
      e.colsum_scale = inv;
      bias_done = true;
    }
    e.out_half = dpl == 1;
    WN_TRY(tcs_gemm(h, dY, 1, nullptr, &zero, T, HP(ws + t.tc_wht[i]), cp.in_ch, e, ws + t.dh[tog], cp.in_ch, dpl == 2, s));
    d = HP(ws + t.dh[tog]);
    tog ^= 1;
  }
  const __half* dskip = d;   // [B*T][S], dpl planes
  SOperand DS{dskip, S, T, B, 1, 0, dpl};
  const int wt = W - T, nwt = -(W - T);
  const int64_t zstride = L > 1 ? t.z[1] - t.z[0] : (int64_t)P * G;   // floats
  // The fused shape keeps the gradient tensors that are only ever read back as MMA dY operands / epilogue addends in ONE
  // fp16 plane (dzs, dafg; round to nearest, gscale keeps them in the normal range) and the sigmoid tape in 16-bit fixed
  // point; the residual-gradient stream dout, which accumulates over all layers, stays split.
  const bool fused = fused_shape(h);
  if (fused)
    for (const ResLayer& ly : h->layers)
      WN_REQUIRE((((uintptr_t)(grads + ly.proj.w_off) | (uintptr_t)(grads + ly.wf.w_off) | (uintptr_t)(grads + ly.wg.w_off)) & 15) == 0,
                 WN_EINVAL, "fp16x2 backward: gradient buffers of the residual layers must be 16-byte aligned");
  // ---- skip path for ALL layers at once ----
  //  dzs[l] = dskip . Ws_l  (slabs read by the gate epilogues only; fp16 for the fused shape, fp32 otherwise): one GEMM with
  //  N = L*G in 256-column groups
  {
    SEpilogue e;
    e.y_slab_cols = G;
    e.y_slab_stride = (int64_t)P * G;
    e.acc_scale = INV_W;
    e.out_half = fused ? 1 : 0;
    WN_TRY(tcs_gemm(h, DS, 1, nullptr, &nwt, W, HP(ws + t.tc_wst), L * G, e, ws + t.dzs, G, 0, s));
  }
  //  dWs_l = dskip^T . z_l : groups of nl layers per CTA group (dskip read from HBM once)
  {
    int nl = 256 / G;
    if (nl > 4) nl = 4;
    const int ngr_all = (L + nl - 1) / nl;
    SOperand Z{HP(ws + t.z[0]), G, W, B, L, 2 * zstride};
    for (int g0 = 0; g0 < ngr_all; g0 += 8) {
      const int ngr = ngr_all - g0 < 8 ? ngr_all - g0 : 8;
      int boff[4], bidx[32];
      float* dws[32];
      for (int j = 0; j < nl; ++j) boff[j] = wt;
      for (int gr = 0; gr < ngr; ++gr)
        for (int j = 0; j < nl; ++j) {
          const int l = (g0 + gr) * nl + j;
          bidx[gr * nl + j] = l < L ? l : -1;
          dws[gr * nl + j] = l < L ? grads + h->layers[l].skip.w_off : nullptr;
        }
      for (int m0 = 0; m0 < S; m0 += 256) {
        const int mv = S - m0 < 256 ? S - m0 : 256;
        float* dws_m[32];
        for (int i = 0; i < ngr * nl; ++i) dws_m[i] = dws[i] ? dws[i] + (int64_t)m0 * G : nullptr;
        WN_TRY(tcs_wgrad(h, DS, 0, m0, mv, Z, nl, boff, bidx, dws_m, nullptr, 256, T, G, 1, inv_wg, s, ngr));
      }
    }
  }
  // ---- residual layers ----
  const bool serp = getenv("WN_NO_SERP") == nullptr;
  int dt = 0;
  const __half* dout = nullptr;
  for (int l = L - 1; l >= 0; --l) {
    const ResLayer& ly = h->layers[l];
    const int zp = wn_zero_prefix(W, ly.dilation, 2);
    const int dir = serp ? (l & 1) : 0, ndir = serp ? !dir : 0;
    const float* dzs = ws + t.dzs + (int64_t)l * P * G;                      // generic shapes: fp32 slabs
    const __half* dzs_h = HP(ws + t.dzs) + (int64_t)l * P * G;               // fused shape: fp16 slabs
    const float* sg = ws + t.tfsg[l] + G;                                    // generic shapes: (tanh | sigmoid) fp32
    const int sg_ld = 2 * G;
    const uint16_t* sg_q = reinterpret_cast<const uint16_t*>(ws + t.tfsg[l]);   // fused shape: 16-bit fixed point [P][G]
    SOperand Z{HP(ws + t.z[l]), G, W, B, 1, 0};
    __half* dafg = HP(ws + t.dafg);
    float* dwp = grads + ly.proj.w_off;
    float* dwf = grads + ly.wf.w_off;
    float* dwg = grads + ly.wg.w_off;
    if (dout && fused) {
      // fused: gate forwards, dxw backwards (serpentine hand-off through L2)
      WN_TRY(tcs_gate_bwd(h, dout, HP(ws + t.tc_wpt) + (int64_t)l * 2 * G * R, dzs_h, HP(ws + t.z[l]), sg_q, dafg, dwp, zp, W, B, 0,
                          inv_wg, s));
    } else if (fused) {
      tcs_gate_backward_top_fused_kernel<<<nblk(P * (G / 4), 256), 256, 0, s>>>(HP(ws + t.z[l]), sg_q, dzs_h, dafg, P, W, G, zp);
      WN_CHECK_LAUNCH();
    } else if (dout) {
      // dz = dout . Wp + dzs_l, gate derivative fused into the epilogue -> dafg
      SOperand DO{dout, R, W, B, 1, 0};
      SEpilogue e;
      e.Rsd = dzs;
      e.ldr = G;
      e.acc_scale = INV_W;
      e.gate_sg = sg;
      e.gate_sg_ld = sg_ld;
      e.gate_z = HP(ws + t.z[l]);
      e.gate_dafg = dafg;
      e.gate_zp = zp;
      e.reverse = dir;
      WN_TRY(tcs_gemm(h, DO, 1, nullptr, &zero, W, HP(ws + t.tc_wpt) + (int64_t)l * 2 * G * R, G, e, nullptr, G, 0, s));
      // dWp += dout^T . z
      for (int m0 = 0; m0 < R; m0 += 256) {
        const int mv = R - m0 < 256 ? R - m0 : 256;
        float* dw = grads + ly.proj.w_off + (int64_t)m0 * G;
        WN_TRY(tcs_wgrad(h, DO, 0, m0, mv, Z, 1, &zero, nullptr, &dw, nullptr, 256, W, G, 1, inv_wg, s, 1, ndir));
      }
    } else {
      tcs_gate_backward_top_kernel<<<nblk(P * (G / 4), 256), 256, 0, s>>>(HP(ws + t.z[l]), sg, sg_ld, dzs, dafg, P, W, G, zp);
      WN_CHECK_LAUNCH();
    }
    SOperand DA{dafg, 2 * G, W, B, 1, 0};
    if (fused) {
      __half* dnew = HP(ws + t.dout[dt]);
      WN_TRY(tcs_dxw(h, dafg, HP(ws + t.tc_w1t) + (int64_t)l * 2 * (R * 4 * G), dout, HP(ws + t.x[l]), dnew, dwf, dwg, ly.dilation, W,
                     B, serp ? 1 : 0, inv_wg, s));
      dout = dnew;
      dt ^= 1;
      continue;
    }
    {
      // dW_{f,g}(o, c, tap) += da[t][o] * x[t - (1-tap) d][c]; both taps share a launch while 2R fits one N tile
      SOperand X{HP(ws + t.x[l]), R, W, B, 1, 0};
      if (2 * R <= 256) {
        const int boff[2] = {-ly.dilation, 0};
        float* d0[2];
        float* d1[2];
        for (int tap = 0; tap < 2; ++tap) {
          d0[tap] = grads + ly.wf.w_off + tap;
          d1[tap] = grads + ly.wg.w_off + tap;
        }
        // rows [0, G) of da are da_f, rows [G, 2G) da_g
        WN_TRY(tcs_wgrad(h, DA, 0, 0, 2 * G, X, 2, boff, nullptr, d0, d1, G, W, 2 * R, 2, inv_wg, s, 1, dir));
      } else {
        for (int tap = 0; tap < 2; ++tap) {
          const int boff = tap == 0 ? -ly.dilation : 0;
          float* d0 = grads + ly.wf.w_off + tap;
          float* d1 = grads + ly.wg.w_off + tap;
          WN_TRY(tcs_wgrad(h, DA, 0, 0, 2 * G, X, 1, &boff, nullptr, &d0, &d1, G, W, 2 * R, 2, inv_wg, s, 1, dir));
        }
      }
    }
    {
      // dx[t] = dout[t] + da[t] . W1(tap 1) + da[t + d] . W1(tap 0)
      __half* dnew = HP(ws + t.dout[dt]);
      const int roff[2] = {0, ly.dilation};
      const int sidx[2] = {0, 0};
      SEpilogue e;
      e.Rsd = dout;
      e.ldr = R;
      e.rsd_split = 1;
      e.acc_scale = INV_W;
      e.reverse = ndir;
      WN_TRY(tcs_gemm(h, DA, 2, sidx, roff, W, HP(ws + t.tc_w1t) + (int64_t)l * 2 * (R * 4 * G), R, e, dnew, R, 1, s));
      dout = dnew;
      dt ^= 1;
    }
  }
  if (h->deterministic) {
    h->bwd_dout = nullptr;
    return h->causal_from_idx ? tcs_embed_backward_det(h, dout, grads, inv, s) : WN_OK;
  }
  // gradient w.r.t. the causal output, back to unscaled fp32 rows for the SIMT embedding / causal-stack backward
  float* dfin = ws + t.dout[dt];
  WN_TRY(tcs_unsplit_rows(reinterpret_cast<const float*>(dout), dfin, R, P, inv, s));
  h->bwd_dout = dfin;
  return WN_OK;
}

// ---- deterministic mode: the embedding gradient as a tensor-core weight gradient ---------------------------------------------
// dWc(r, q, tap) = sum_p [idx[p - (kc-1-tap)] == q] * dout[p][r]  is  onehot^T . dout: the one-hot rows are written as ONE exact
// fp16 plane and the generic weight-gradient kernel (per-CTA accumulators in TMEM, per-CTA slabs) does the rest -- no
// shared-memory float atomics, whose order the SIMT scatter kernel cannot fix.
__global__ void tcs_onehot_rows_kernel(const int32_t* __restrict__ idx, __half* __restrict__ oh, int64_t P, int Q) {
  const int qv = Q >> 3;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P * qv) return;
  const int64_t p = i / qv;
  const int q0 = (int)(i - p * qv) * 8;
  const int k = idx[p] - q0;
  uint32_t w[4] = {0u, 0u, 0u, 0u};
  if (k >= 0 && k < 8) w[k >> 1] = (k & 1) ? 0x3c000000u : 0x00003c00u;   // fp16 1.0 in the upper / lower half
  *reinterpret_cast<uint4*>(oh + p * Q + q0) = make_uint4(w[0], w[1], w[2], w[3]);
}

bool tcs_det_supported(const wn_handle* h) {
  return fused_shape(h) && h->cfg.n_causal == 1 && h->causal[0].b_off < 0 && h->Q % 64 == 0 && h->Q <= 256 && h->layers.size() >= 2;
}

static int tcs_embed_backward_det(wn_handle* h, const __half* dout_split, float* grads, float inv, cudaStream_t s) {
  const Tape& t = h->tape;
  const int kc = h->cfg.causal_filter_width, Q = h->Q, R = h->R;
  __half* oh = HP(h->ws + t.dzs);                      // [P][Q] halves: the dzs slabs are dead by now
  tcs_onehot_rows_kernel<<<nblk(t.P * (Q / 8), 256), 256, 0, s>>>(h->x_idx, oh, t.P, Q);
  WN_CHECK_LAUNCH();
  SOperand OH{oh, Q, t.W, t.B, 1, 0, 1};
  SOperand X{dout_split, R, t.W, t.B, 1, 0};
  const int zero = 0;
  for (int tap = 0; tap < kc; ++tap) {
    float* dw = grads + h->causal[0].w_off + tap;      // element (r, q, tap) at (r * Q + q) * kc + tap
    WN_TRY(tcs_wgrad(h, OH, -(kc - 1 - tap), 0, Q, X, 1, &zero, nullptr, &dw, nullptr, 256, t.W, kc, (int64_t)Q * kc, inv, s));
  }
  return WN_OK;
}

// dlogits fp32 [B*T][Q] (written by the CE kernel) -> split rows scaled by gscale, in place
int tcs_scale_split_dlogits(wn_handle* h, int T, float gscale, cudaStream_t s) {
  const Tape& t = h->tape;
  return tcs_split_rows_inplace(h->ws + t.dlogits, h->Q, (int64_t)t.B * T, gscale, s);
}

// ---- developer hooks (tests/dev/check_tcs_kernels.py): the two generic split kernels on caller-provided split tensors ----
extern "C" int wn_tcs_debug_gemm(wn_handle* h, const void* a_split, int C, int rows_in, int num_seq, int ns, int row_off0,
                                 int row_off1, int rows_out, const void* w_split, int N, const void* rsd_split, const void* mask_split,
                                 int relu, void* y, int out_split, void* stream) {
  WN_REQUIRE(h && a_split && w_split && y && ns >= 1 && ns <= 2, WN_EINVAL, "wn_tcs_debug_gemm: bad argument");
  SOperand A{reinterpret_cast<const __half*>(a_split), C, rows_in, num_seq, 1, 0};
  const int sidx[2] = {0, 0};
  const int roff[2] = {row_off0, row_off1};
  SEpilogue e;
  e.relu = relu;
  if (rsd_split) {
    e.Rsd = rsd_split;
    e.ldr = N;
    e.rsd_split = 1;
  }
  if (mask_split) {
    e.mask = reinterpret_cast<const __half*>(mask_split);
    e.ldm = N;
    e.mask_rows_in = rows_out;
  }
  return tcs_gemm(h, A, ns, sidx, roff, rows_out, reinterpret_cast<const __half*>(w_split), N, e, y, N, out_split, (cudaStream_t)stream);
}

extern "C" int wn_tcs_debug_wgrad(wn_handle* h, const void* dy_split, int M, const void* x_split, int C, int rows, int num_seq,
                                  int x_row_off, float scale, float* dW, void* stream) {
  WN_REQUIRE(h && dy_split && x_split && dW && M <= 256, WN_EINVAL, "wn_tcs_debug_wgrad: bad argument");
  SOperand dY{reinterpret_cast<const __half*>(dy_split), M, rows, num_seq, 1, 0};
  SOperand X{reinterpret_cast<const __half*>(x_split), C, rows, num_seq, 1, 0};
  return tcs_wgrad(h, dY, 0, 0, M, X, 1, &x_row_off, nullptr, &dW, nullptr, 256, rows, C, 1, scale, (cudaStream_t)stream);
}

// Profiling hooks (bench.py roofline): ONE launch of the fused fp16x2 residual-layer kernel / of the skip GEMM on the bound
// tape; require a preceding fp16x2 wn_forward_residual_block (weights prepared, x[l] in split format).
extern "C" int wn_tcs_layer_forward(wn_handle* h, int layer, void* stream) {
  WN_REQUIRE(h && h->ws && h->tape_split && fused_shape(h), WN_ESTATE, "wn_tcs_layer_forward: run an fp16x2 forward of the fused shape first");
  WN_REQUIRE(layer >= 0 && layer < (int)h->layers.size(), WN_EINVAL, "bad layer index");
  return tcs_layer_launch(h, layer, (cudaStream_t)stream);
}
extern "C" int wn_tcs_skip_gemm(wn_handle* h, void* stream) {
  WN_REQUIRE(h && h->ws && h->tape_split, WN_ESTATE, "wn_tcs_skip_gemm: run an fp16x2 forward first");
  return tcs_skip_gemm(h, (cudaStream_t)stream);
}

#ifdef WN_LAYER_TRACE
extern "C" int wn_debug_tcs_trace(long long* out) {
  return cudaMemcpyFromSymbol(out, g_trace_s, sizeof(long long) * 64 * 32) == cudaSuccess ? 0 : -1;
}
#endif
