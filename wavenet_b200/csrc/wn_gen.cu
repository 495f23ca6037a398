// Incremental generator: FasterWaveNet._forward_one_step (faster_wavenet.py:50-113)
// and the sampling loop of train_audio/generate.py:24-43 as ONE persistent kernel.
//
// The reference keeps, per layer, a full receptive-field-wide window moved with
// xp.roll (faster_wavenet.py:72,90,94) and runs the head over the whole window
// (faster_wavenet.py:105-113) although only the last column is consumed
// (generate.py:38).  Here each layer keeps a (k-1)*d deep ring of its own input,
// only the last column is computed, sampling happens on device and the loop over
// audio samples never returns to the host.  Arithmetic is exact fp32 (FFMA) so
// greedy sequences can match the oracle.
#include <cuda_fp16.h>
#include <stdlib.h>
#include <string.h>

#include "wn_common.h"
#include "wn_tc.cuh"

static const int GT_HOST = 512;   // consumer threads per CTA (must equal GT below)
static const int V4_CS_HOST = 8;  // CTAs per cluster of the v4 generator (must equal V4_CS below)

struct GenLayerOff {
  int64_t wa, ba, wb, bb, ring;  // offsets (floats) into the state buffer
  int G, dilation, ring_len;
  int has_ba, has_bb, pad;         // biases present (absent ones are not even loaded)
};

struct GenLayout {
  int n = 0, Q = 0, R = 0, S = 0, k = 0, kc = 0, n_causal = 0, n_head = 0, L = 0;
  int causal_ch[WN_MAX_CAUSAL];
  int head_ch[WN_MAX_HEAD];
  int64_t emb = 0, emb_b = 0;
  int64_t cw[WN_MAX_CAUSAL], cb[WN_MAX_CAUSAL], chist[WN_MAX_CAUSAL];  // causal layers >= 1
  int64_t hw[WN_MAX_HEAD], hb[WN_MAX_HEAD];
  int has_hb = 0, has_cb = 0;
  int64_t idx_hist = 0;     // int32 [n][kc-1]
  int64_t cur_logits = 0;   // [n][Q]
  int64_t layers_dev = 0;   // GenLayerOff[L] copied to device (as raw bytes)
  int64_t chunks_dev = 0;   // GenChunk[] streaming schedule
  int64_t chunks3_dev = 0;  // GenChunk[] schedule of the packed (v3) weights
  int64_t wpk = 0;          // packed weights (v3)
  int64_t wpk4 = 0;         // per-cluster-rank packed weight slices (v4)
  int64_t wpk4_rank = 0;    // floats per rank
  int64_t wpk6 = 0;         // per-rank fp16 hi|lo B tiles of the tensor-core generator (v6)
  int64_t wpk6_rank_bytes = 0;
  int64_t ring6 = 0;        // v6 dilation rings: [cluster][rank][slot][32 KB operand tile]
  int64_t ring6_cta_bytes = 0;
  int64_t total = 0;
  int maxw = 0;             // widest vector anywhere (for smem sizing)
};

struct GenChunk {
  uint64_t off;     // byte offset from the state base (16-byte aligned)
  uint32_t bytes;
  uint32_t pad;
};

struct wn_gen {
  wn_handle* h = nullptr;
  int head_act = 1;
  GenLayout lay;
  std::vector<GenLayerOff> layers;
  float* state = nullptr;
  int64_t state_bytes = 0;
  bool primed = false;
  int64_t t = 0;         // absolute time of the next sample
  int64_t steps_done = 0;  // incremental steps since priming
  bool stream_ok = false;             // every matrix fits the streamed (cp.async.bulk ring) matvec
  std::vector<GenChunk> chunks;       // per-step streaming schedule
  bool v3_ok = false;                 // network tiles exactly over 16 warps (config C): packed-weight kernel
  std::vector<GenChunk> chunks3;      // schedule over the packed weights
  struct Pack3 { int64_t src, dst; int K, N, chunkK, mode; };
  std::vector<Pack3> packs3;
  bool v4_ok = false;                 // config-C shape: the cluster generators' packed weight slices exist (v4: one 8-CTA
                                      // cluster per stream while they are all co-resident; v5: 16 streams per cluster)
  bool v6_ok = false;                 // tensor-core generator (gen_kernel_v6): config-C shape without biases, >= 16 streams
  int v6_spc = 128, v6_clusters = 0, v6_cs = 8;  // streams per cluster, clusters, CTAs per cluster (8 or 4)
  bool ring6_valid = false;           // the v6 operand-tile rings hold the current state
  bool ringf_valid = true;            // the fp32 rings hold the current state
};

namespace {

constexpr int GT = 512;  // consumer threads per CTA (the streaming variant adds one producer warp)
constexpr int RING_STAGES = 4;
constexpr int RING_STAGE_BYTES = 32768;
constexpr int RING_GROUPS = 1;      // 4*nsl-row groups per chunk
constexpr int MAX_CHUNKS = 512;     // schedule entries kept in shared memory

// barrier among the GT consumer threads only (the producer warp never joins)
__device__ __forceinline__ void csync() { asm volatile("bar.sync 1, 512;" ::: "memory"); }
__device__ __forceinline__ void csync4() { asm volatile("bar.sync 1, 256;" ::: "memory"); }   // v4: 256 consumer threads

// consumer-side view of the weight ring
struct StreamCtx {
  uint32_t full0, empty0;   // smem addresses of the barrier arrays
  const float* ring;        // generic pointer to stage 0
  uint32_t it;              // chunks consumed so far
};

inline int64_t align_up64(int64_t v) { return (v + 63) / 64 * 64; }

// dst[(tap*C + c)*ldn + noff + o] = W[o][c][tap]
__global__ void gen_transpose_conv(const float* __restrict__ W, float* __restrict__ dst, int O, int C, int taps, int ldn,
                                   int noff) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= O * C * taps) return;
  const int o = i % O, c = (i / O) % C, tap = i / (O * C);
  dst[((int64_t)tap * C + c) * ldn + noff + o] = W[((int64_t)o * C + c) * taps + tap];
}

__global__ void gen_copy_bias(const float* __restrict__ b, float* __restrict__ dst, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = b ? b[i] : 0.f;
}

// ring[stream][slot][c] <- x[stream][tau][c] for the last ring_len window positions, slot = tau mod ring_len
__global__ void gen_fill_ring(const float* __restrict__ x, float* __restrict__ ring, int n, int Win, int C, int len,
                              int linear) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)n * len * C) return;
  const int c = (int)(i % C);
  const int j = (int)((i / C) % len);
  const int sidx = (int)(i / ((int64_t)C * len));
  const int tau = Win - len + j;
  const float v = tau >= 0 ? x[((int64_t)sidx * Win + tau) * C + c] : 0.f;
  const int slot = linear ? j : ((tau % len) + len) % len;   // linear: oldest first
  ring[((int64_t)sidx * len + slot) * C + c] = v;
}

// the same from the SPLIT tape of the fp16x2 forward: rows of [hi C | lo C] fp16, both carrying `inv_scale`^-1
__global__ void gen_fill_ring_split(const __half* __restrict__ x, float* __restrict__ ring, int n, int Win, int C, int len,
                                    float inv_scale) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)n * len * C) return;
  const int c = (int)(i % C);
  const int j = (int)((i / C) % len);
  const int sidx = (int)(i / ((int64_t)C * len));
  const int tau = Win - len + j;
  float v = 0.f;
  if (tau >= 0) {
    const __half* row = x + ((int64_t)sidx * Win + tau) * 2 * C;
    v = (__half2float(row[c]) + __half2float(row[C + c])) * inv_scale;
  }
  ring[((int64_t)sidx * len + ((tau % len) + len) % len) * C + c] = v;
}

__global__ void gen_fill_idx_hist(const int32_t* __restrict__ window, int32_t* __restrict__ hist, int n, int Win,
                                  int kc1) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * kc1) return;
  const int sidx = i / kc1, j = i % kc1;
  const int tau = Win - kc1 + j;
  hist[i] = tau >= 0 ? window[(int64_t)sidx * Win + tau] : -1;
}

struct GenArgs {
  float* state;
  GenLayout lay;
  const GenLayerOff* layers;
  const GenChunk* chunks;   // per-step weight streaming schedule (STREAM kernels)
  int n_chunks;
  int n_steps;
  int mode;            // WN_GEN_GREEDY / WN_GEN_SAMPLE
  int sample_first;    // 1: draw each step's input from cur_logits (run); 0: inputs forced (step API)
  const int32_t* forced;  // [n] when !sample_first
  uint64_t seed;
  int64_t t0;          // absolute time of the first processed sample
  int head_elu;        // ELU head (reference incremental steps) vs ReLU
  int32_t* out;        // [n][n_steps] or null
  float* probs;        // [n][Q] or null (written after the last step)
  int apply_softmax;
};

__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

__device__ __forceinline__ float gumbel(uint64_t seed, uint64_t stream, uint64_t t, uint32_t q) {
  const uint64_t h = splitmix64(seed ^ splitmix64(stream * 0x100000001B3ull + t) ^ ((uint64_t)q << 32 | q));
  const float u = ((float)(h >> 40) + 0.5f) * (1.0f / 16777216.0f);  // (0,1)
  return -logf(-logf(u));
}

// out[s][o] (+)= sum_kk Wt[kk][o] * xin[s][kk]; all GT threads cooperate; K split across
// thread groups when N is small.  part: smem scratch of >= GT*NS floats.
template <int NS>
__device__ __forceinline__ void matvec(const float* __restrict__ Wt, int K, int N, const float* xin, int ldx,
                                       float* part, float* out, int ldo, const float* __restrict__ bias,
                                       const float* addin, int lda) {
  const int tid = threadIdx.x;
  if (N <= GT) {
    const int ksplit = GT / N;  // >= 1
    const int Kc = (K + ksplit - 1) / ksplit;
    const int ks = tid / N, o = tid - ks * N;
    float acc[NS];
#pragma unroll
    for (int s = 0; s < NS; ++s) acc[s] = 0.f;
    if (ks < ksplit) {
      const int kb = ks * Kc, ke = min(K, kb + Kc);
      const float* wp = Wt + (int64_t)kb * N + o;
#pragma unroll 8
      for (int kk = kb; kk < ke; ++kk) {
        const float w = __ldg(wp);
        wp += N;
#pragma unroll
        for (int s = 0; s < NS; ++s) acc[s] = fmaf(w, xin[s * ldx + kk], acc[s]);
      }
#pragma unroll
      for (int s = 0; s < NS; ++s) part[(ks * NS + s) * N + o] = acc[s];
    }
    csync();
    for (int i = tid; i < NS * N; i += GT) {
      const int s = i / N, oo = i - s * N;
      float v = bias ? bias[oo] : 0.f;
      for (int q = 0; q < ksplit; ++q) v += part[(q * NS + s) * N + oo];
      if (addin) v += addin[s * lda + oo];
      out[s * ldo + oo] = v;
    }
    csync();
  } else {
    for (int o = tid; o < N; o += GT) {
      float acc[NS];
#pragma unroll
      for (int s = 0; s < NS; ++s) acc[s] = 0.f;
      const float* wp = Wt + o;
#pragma unroll 8
      for (int kk = 0; kk < K; ++kk) {
        const float w = __ldg(wp);
        wp += N;
#pragma unroll
        for (int s = 0; s < NS; ++s) acc[s] = fmaf(w, xin[s * ldx + kk], acc[s]);
      }
      const float b = bias ? bias[o] : 0.f;
#pragma unroll
      for (int s = 0; s < NS; ++s) out[s * ldo + o] = acc[s] + b + (addin ? addin[s * lda + o] : 0.f);
    }
    csync();
  }
}

// Streamed matvec (partials only): the [K][N] matrix arrives in row chunks through the shared-memory ring
// (cp.async.bulk issued by the producer warp in consumption order).  Each thread owns 4 consecutive outputs
// (LDS.128) for a 4-row slice of every row group; slice partial sums land in part[(slice*NS + s)*N + o] and the
// CALLER reduces them after a csync (so it can fuse its own epilogue).  Input rows [0,Ksplit) come from xa, the
// rest from xb (past taps | current sample) -- no gather pass.  Requires N%4==0, N<=1024, K%4==0, Ksplit%4==0.
// Returns the number of slices.
template <int NS>
__device__ __forceinline__ int matvec_stream(StreamCtx& cx, int K, int N, const float* xa, int lda, const float* xb,
                                             int ldb, int Ksplit, float* part) {
  const int tid = threadIdx.x;
  const int NV = N >> 2;
  const int nsl = GT / NV;
  const int rpc = nsl << 2;
  const int sl = tid / NV, o4 = tid - sl * NV;
  const bool active = sl < nsl;
  float acc[NS][4];
#pragma unroll
  for (int s = 0; s < NS; ++s) acc[s][0] = acc[s][1] = acc[s][2] = acc[s][3] = 0.f;
  for (int r0 = 0; r0 < K; r0 += RING_GROUPS * rpc) {
    const uint32_t stage = cx.it % RING_STAGES, parity = (cx.it / RING_STAGES) & 1;
    tc::mbar_wait(cx.full0 + 8 * stage, parity);
    const int rows = min(RING_GROUPS * rpc, K - r0);
    if (active) {
#pragma unroll
      for (int gp = 0; gp < RING_GROUPS; ++gp) {
        const int rr = gp * rpc + sl * 4;
        if (rr < rows) {
          const float* w = cx.ring + stage * (RING_STAGE_BYTES / 4) + rr * N + o4 * 4;
          const int kk = r0 + rr;
          const float* xp = kk < Ksplit ? xa + kk : xb + (kk - Ksplit);
          const int ld = kk < Ksplit ? lda : ldb;
          float4 wv[4];
#pragma unroll
          for (int r = 0; r < 4; ++r) wv[r] = *reinterpret_cast<const float4*>(w + r * N);
#pragma unroll
          for (int s = 0; s < NS; ++s) {
            const float4 x = *reinterpret_cast<const float4*>(xp + s * ld);
            acc[s][0] = fmaf(wv[0].x, x.x, acc[s][0]); acc[s][1] = fmaf(wv[0].y, x.x, acc[s][1]);
            acc[s][2] = fmaf(wv[0].z, x.x, acc[s][2]); acc[s][3] = fmaf(wv[0].w, x.x, acc[s][3]);
            acc[s][0] = fmaf(wv[1].x, x.y, acc[s][0]); acc[s][1] = fmaf(wv[1].y, x.y, acc[s][1]);
            acc[s][2] = fmaf(wv[1].z, x.y, acc[s][2]); acc[s][3] = fmaf(wv[1].w, x.y, acc[s][3]);
            acc[s][0] = fmaf(wv[2].x, x.z, acc[s][0]); acc[s][1] = fmaf(wv[2].y, x.z, acc[s][1]);
            acc[s][2] = fmaf(wv[2].z, x.z, acc[s][2]); acc[s][3] = fmaf(wv[2].w, x.z, acc[s][3]);
            acc[s][0] = fmaf(wv[3].x, x.w, acc[s][0]); acc[s][1] = fmaf(wv[3].y, x.w, acc[s][1]);
            acc[s][2] = fmaf(wv[3].z, x.w, acc[s][2]); acc[s][3] = fmaf(wv[3].w, x.w, acc[s][3]);
          }
        }
      }
    }
    __syncwarp();
    if ((tid & 31) == 0) tc::mbar_arrive(cx.empty0 + 8 * stage);
    ++cx.it;
  }
  if (active) {
#pragma unroll
    for (int s = 0; s < NS; ++s)
      *reinterpret_cast<float4*>(part + (sl * NS + s) * N + o4 * 4) = make_float4(acc[s][0], acc[s][1], acc[s][2], acc[s][3]);
  }
  return nsl;
}

template <int NS, bool STREAM>
__global__ void __launch_bounds__(GT + (STREAM ? 32 : 0)) gen_kernel(GenArgs a) {
  extern __shared__ float sm[];
  const GenLayout& L = a.lay;
  const int tid = threadIdx.x;
  const int s0 = blockIdx.x * NS;  // first stream of this CTA
  const int maxw = L.maxw;
  // smem carve-up
  float* xv = sm;                       // [NS][maxw] current trunk vector
  float* xin = xv + NS * maxw;          // [NS][k*maxw] conv input (past taps | current)
  float* av = xin + NS * L.k * maxw + NS * L.kc * maxw;  // [NS][2*maxw] pre-activations / z
  float* zv = av + NS * 2 * maxw;       // [NS][maxw]
  float* skipv = zv + NS * maxw;        // [NS][maxw] skip accumulator / head ping
  float* hv = skipv + NS * maxw;        // [NS][maxw] head pong
  float* part = hv + NS * maxw;         // [4*GT*NS]
  float* xpast = part + 4 * GT * NS;      // [NS][L][(k-1)R] past taps of every layer (STREAM only)
  __shared__ GenLayerOff s_layers[STREAM ? 128 : 1];
  __shared__ int s_sample[NS];
  __shared__ __align__(8) uint64_t s_bars[2 * RING_STAGES];
  __shared__ GenChunk s_sched[STREAM ? MAX_CHUNKS : 1];
  StreamCtx cx;
  cx.it = 0;
  if (STREAM) {
    uint8_t* ring_g = reinterpret_cast<uint8_t*>(xpast + NS * L.L * (L.k - 1) * L.R);
    for (int i = threadIdx.x; i < L.L; i += blockDim.x) s_layers[i] = a.layers[i];
    ring_g += (128 - (tc::smem_u32(ring_g) & 127)) & 127;
    cx.ring = reinterpret_cast<const float*>(ring_g);
    cx.full0 = tc::smem_u32(&s_bars[0]);
    cx.empty0 = tc::smem_u32(&s_bars[RING_STAGES]);
    if (threadIdx.x == 0) {
      for (int i = 0; i < RING_STAGES; ++i) {
        tc::mbar_init(cx.full0 + 8 * i, 1);
        tc::mbar_init(cx.empty0 + 8 * i, GT / 32);
      }
      tc::fence_barrier_init();
    }
    for (int i = threadIdx.x; i < a.n_chunks; i += blockDim.x) s_sched[i] = a.chunks[i];
    __syncthreads();   // all 288 threads, once
    if (threadIdx.x >= GT) {
      // ---- producer warp: stream every step's weights in consumption order ----
      if (threadIdx.x == GT) {
        const uint8_t* sbase = reinterpret_cast<const uint8_t*>(a.state);
        const uint32_t ring_s = tc::smem_u32(ring_g);
        uint32_t it = 0;
        for (int step = 0; step < a.n_steps; ++step)
          for (int c = 0; c < a.n_chunks; ++c, ++it) {
            const GenChunk ch = s_sched[c];
            const uint32_t stage = it % RING_STAGES;
            tc::mbar_wait(cx.empty0 + 8 * stage, ((it / RING_STAGES) & 1) ^ 1);
            tc::mbar_arrive_expect_tx(cx.full0 + 8 * stage, ch.bytes);
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             ring_s + stage * RING_STAGE_BYTES),
                         "l"(reinterpret_cast<uint64_t>(sbase + ch.off)), "r"(ch.bytes), "r"(cx.full0 + 8 * stage)
                         : "memory");
          }
      }
      return;
    }
  }
  __shared__ float s_redv[GT / 32];
  __shared__ int s_redi[GT / 32];

  float* st = a.state;
  int32_t* idx_hist = (int32_t*)(st + L.idx_hist);
  float* cur_logits = st + L.cur_logits;
  const int kc1 = L.kc - 1;

  for (int step = 0; step < a.n_steps; ++step) {
    const int64_t t = a.t0 + step;
    // ---- 1. choose this step's input sample --------------------------------------
    for (int s = 0; s < NS; ++s) {
      const int stream = s0 + s;
      if (stream >= L.n) {
        if (tid == 0) s_sample[s] = 0;
        continue;
      }
      if (!a.sample_first) {
        if (tid == 0) s_sample[s] = a.forced[stream];
        continue;
      }
      // argmax over Q of logits (+ Gumbel noise when sampling); ties -> lowest index (np.argmax)
      float bv = -INFINITY;
      int bi = 0x7fffffff;
      for (int q = tid; q < L.Q; q += GT) {
        float v = cur_logits[(int64_t)stream * L.Q + q];
        if (a.mode == WN_GEN_SAMPLE) v += gumbel(a.seed, (uint64_t)stream, (uint64_t)t, (uint32_t)q);
        if (v > bv || (v == bv && q < bi)) {
          bv = v;
          bi = q;
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > bv || (ov == bv && oi < bi)) {
          bv = ov;
          bi = oi;
        }
      }
      if ((tid & 31) == 0) {
        s_redv[tid >> 5] = bv;
        s_redi[tid >> 5] = bi;
      }
      csync();
      if (tid == 0) {
        for (int w = 1; w < GT / 32; ++w)
          if (s_redv[w] > bv || (s_redv[w] == bv && s_redi[w] < bi)) {
            bv = s_redv[w];
            bi = s_redi[w];
          }
        s_sample[s] = bi;
        if (a.out) a.out[(int64_t)stream * a.n_steps + step] = bi;
      }
      csync();
    }
    csync();

    // ---- 2. causal stack (wavenet.py:281-286 on one-hot taps == table gather) -----
    {
      const int R0 = L.causal_ch[0];
      const float* emb = st + L.emb;
      const float* eb = st + L.emb_b;
      for (int i = tid; i < NS * R0; i += GT) {
        const int s = i / R0, r = i - s * R0;
        const int stream = s0 + s;
        float v = eb[r];
        if (stream < L.n) {
          for (int j = 0; j < L.kc; ++j) {
            const int q = j == kc1 ? s_sample[s] : idx_hist[(int64_t)stream * kc1 + j];
            if (q >= 0) v += emb[((int64_t)j * L.Q + q) * R0 + r];
          }
        }
        xv[s * maxw + r] = v;
      }
      csync();
      if (kc1 > 0)
        for (int i = tid; i < NS; i += GT) {
          const int stream = s0 + i;
          if (stream < L.n) {
            for (int j = 0; j + 1 < kc1; ++j) idx_hist[(int64_t)stream * kc1 + j] = idx_hist[(int64_t)stream * kc1 + j + 1];
            idx_hist[(int64_t)stream * kc1 + kc1 - 1] = s_sample[i];
          }
        }
      for (int ci = 1; ci < L.n_causal; ++ci) {
        const int Cin = L.causal_ch[ci - 1], Cout = L.causal_ch[ci];
        float* hist = st + L.chist[ci];  // [n][kc-1][Cin], oldest first
        float* cin = xin;                // [NS][kc*Cin]
        for (int i = tid; i < NS * L.kc * Cin; i += GT) {
          const int s = i / (L.kc * Cin), rem = i - s * L.kc * Cin;
          const int j = rem / Cin, c = rem - j * Cin;
          const int stream = s0 + s;
          float v = 0.f;
          if (stream < L.n) v = j == kc1 ? xv[s * maxw + c] : hist[((int64_t)stream * kc1 + j) * Cin + c];
          cin[s * (L.kc * maxw) + j * Cin + c] = v;
        }
        csync();
        for (int i = tid; i < NS * kc1 * Cin; i += GT) {  // slide history
          const int s = i / (kc1 * Cin), rem = i - s * kc1 * Cin;
          const int stream = s0 + s;
          if (stream < L.n) hist[(int64_t)stream * kc1 * Cin + rem] = cin[s * (L.kc * maxw) + Cin + rem];
        }
        if (STREAM) {
          const int nsc = matvec_stream<NS>(cx, L.kc * Cin, Cout, cin, L.kc * maxw, cin, L.kc * maxw, L.kc * Cin, part);
          csync();
          for (int i = tid; i < NS * Cout; i += GT) {
            const int s = i / Cout, o = i - s * Cout;
            float v = L.has_cb ? st[L.cb[ci] + o] : 0.f;
            for (int q = 0; q < nsc; ++q) v += part[(q * NS + s) * Cout + o];
            xv[s * maxw + o] = v;
          }
          csync();
        } else matvec<NS>(st + L.cw[ci], L.kc * Cin, Cout, cin, L.kc * maxw, part, xv, maxw, st + L.cb[ci], nullptr, 0);
      }
    }

    float* hin;
    float* hout;
    if constexpr (STREAM) {
      // ---- 3s. residual layers, streamed weights (ResidualConvLayer._forward, wavenet.py:350-356) ----
      const int R = L.R, S = L.S;
      const int pastw = (L.k - 1) * R;                 // past-tap floats per layer per stream
      // all past taps x_l[t-(k-1-tap)d] of this step are already in the rings: fetch them in one go
      {
        const int per = pastw >> 2;                    // float4 per (stream, layer)
        for (int i = tid; i < NS * L.L * per; i += GT) {
          const int v4 = i % per, sl_ = i / per;
          const int l = sl_ % L.L, s = sl_ / L.L;
          const int stream = s0 + s;
          const GenLayerOff& ly = s_layers[l];
          const int tap = (v4 * 4) / R, c = v4 * 4 - tap * R;
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (stream < L.n) {
            const int len = ly.ring_len;
            const int64_t tau = t - (int64_t)(L.k - 1 - tap) * ly.dilation;
            const int slot = (int)(((tau % len) + len) % len);
            v = *reinterpret_cast<const float4*>(st + ly.ring + ((int64_t)stream * len + slot) * R + c);
          }
          *reinterpret_cast<float4*>(xpast + (s * L.L + l) * pastw + v4 * 4) = v;
        }
      }
      for (int i = tid; i < NS * S; i += GT) skipv[(i / S) * maxw + (i % S)] = 0.f;
      csync();
      for (int l = 0; l < L.L; ++l) {
        const GenLayerOff& ly = s_layers[l];
        const int len = ly.ring_len, G = ly.G;
        for (int i = tid; i < NS * R; i += GT) {        // push x[t] into the ring (roll, faster_wavenet.py:90-91)
          const int s = i / R, c = i - s * R;
          const int stream = s0 + s;
          if (stream < L.n) st[ly.ring + ((int64_t)stream * len + (int)(t % len)) * R + c] = xv[s * maxw + c];
        }
        const int nsa = matvec_stream<NS>(cx, L.k * R, 2 * G, xpast + l * pastw, L.L * pastw, xv, maxw, pastw, part);
        csync();
        for (int i = tid; i < NS * G; i += GT) {        // reduce slices + gate (wavenet.py:351)
          const int s = i / G, g = i - s * G;
          float f = ly.has_ba ? st[ly.ba + g] : 0.f, gg = ly.has_ba ? st[ly.ba + G + g] : 0.f;
          for (int q = 0; q < nsa; ++q) {
            f += part[(q * NS + s) * 2 * G + g];
            gg += part[(q * NS + s) * 2 * G + G + g];
          }
          zv[s * maxw + g] = tanhf(f) * (1.f / (1.f + expf(-gg)));
        }
        csync();
        const int N2 = R + S;
        const int nsb = matvec_stream<NS>(cx, G, N2, zv, maxw, zv, maxw, G, part);
        csync();
        for (int i = tid; i < NS * N2; i += GT) {       // reduce slices + residual / skip accumulation
          const int s = i / N2, o = i - s * N2;
          float v = ly.has_bb ? st[ly.bb + o] : 0.f;
          for (int q = 0; q < nsb; ++q) v += part[(q * NS + s) * N2 + o];
          if (o < R)
            xv[s * maxw + o] += v;                      // output = projection_block + x, wavenet.py:354
          else
            skipv[s * maxw + (o - R)] += v;             // sum_skip_connections += z, faster_wavenet.py:100
        }
        csync();
      }
      // ---- 4s. head (faster_wavenet.py:105-113: ELU; wavenet.py:584-593: ReLU) ----
      hin = skipv;
      hout = hv;
      for (int i = tid; i < NS * S; i += GT) {
        const int s = i / S, c = i - s * S;
        const float v = hin[s * maxw + c];
        hin[s * maxw + c] = a.head_elu ? (v > 0.f ? v : expm1f(v)) : fmaxf(v, 0.f);
      }
      csync();
      for (int hi = 0; hi < L.n_head; ++hi) {
        const int Cin = L.head_ch[hi], Cout = L.head_ch[hi + 1];
        const int nsh = matvec_stream<NS>(cx, Cin, Cout, hin, maxw, hin, maxw, Cin, part);
        csync();
        const bool last = hi == L.n_head - 1;
        for (int i = tid; i < NS * Cout; i += GT) {
          const int s = i / Cout, o = i - s * Cout;
          float v = L.has_hb ? st[L.hb[hi] + o] : 0.f;
          for (int q = 0; q < nsh; ++q) v += part[(q * NS + s) * Cout + o];
          if (!last) v = a.head_elu ? (v > 0.f ? v : expm1f(v)) : fmaxf(v, 0.f);   // activation of the next layer
          hout[s * maxw + o] = v;
        }
        csync();
        float* tmp = hin;
        hin = hout;
        hout = tmp;
      }
    } else {
      // ---- 3. residual layers (ResidualConvLayer._forward, wavenet.py:350-356) -------
      const int R = L.R;
      for (int i = tid; i < NS * L.S; i += GT) skipv[(i / L.S) * maxw + (i % L.S)] = 0.f;
      for (int l = 0; l < L.L; ++l) {
        const GenLayerOff ly = a.layers[l];
        const int len = ly.ring_len;
        float* ring = st + ly.ring;  // [n][len][R]
        const int kR = L.k * R;
        // gather taps: xin[s][tap*R + c], tap k-1 = current sample (wavenet.py:288-290)
        for (int i = tid; i < NS * kR; i += GT) {
          const int s = i / kR, rem = i - s * kR;
          const int tap = rem / R, c = rem - tap * R;
          const int stream = s0 + s;
          float v = 0.f;
          if (stream < L.n) {
            if (tap == L.k - 1) {
              v = xv[s * maxw + c];
            } else {
              const int64_t tau = t - (int64_t)(L.k - 1 - tap) * ly.dilation;
              const int slot = (int)(((tau % len) + len) % len);
              v = ring[((int64_t)stream * len + slot) * R + c];
            }
          }
          xin[s * (L.k * maxw) + rem] = v;
        }
        csync();
        if (len > 0)
          for (int i = tid; i < NS * R; i += GT) {  // push x[t] into the ring (roll, faster_wavenet.py:90-91)
            const int s = i / R, c = i - s * R;
            const int stream = s0 + s;
            if (stream < L.n) ring[((int64_t)stream * len + (int)(t % len)) * R + c] = xv[s * maxw + c];
          }
        const int G = ly.G;
        matvec<NS>(st + ly.wa, kR, 2 * G, xin, L.k * maxw, part, av, 2 * maxw, st + ly.ba, nullptr, 0);
        for (int i = tid; i < NS * G; i += GT) {  // z = tanh(a_f) * sigmoid(a_g), wavenet.py:351
          const int s = i / G, g = i - s * G;
          const float f = av[s * 2 * maxw + g], gg = av[s * 2 * maxw + G + g];
          zv[s * maxw + g] = tanhf(f) * (1.f / (1.f + expf(-gg)));
        }
        csync();
        // [x_next | skip] = WB^T z + b (+ x | + skip_acc)
        // proj and skip share one [G][R+S] matrix; addin differs, so run them as two calls on column ranges
        matvec<NS>(st + ly.wb, G, R + L.S, zv, maxw, part, av, 2 * maxw, st + ly.bb, nullptr, 0);
        for (int i = tid; i < NS * (R + L.S); i += GT) {
          const int s = i / (R + L.S), o = i - s * (R + L.S);
          const float v = av[s * 2 * maxw + o];
          if (o < R)
            xv[s * maxw + o] += v;                 // output = projection_block + x, wavenet.py:354
          else
            skipv[s * maxw + (o - R)] += v;        // sum_skip_connections += z, faster_wavenet.py:100
        }
        csync();
      }

      // ---- 4. head (faster_wavenet.py:105-113: ELU; wavenet.py:584-593: ReLU) --------
      hin = skipv;
      hout = hv;
      for (int hi = 0; hi < L.n_head; ++hi) {
        const int Cin = L.head_ch[hi], Cout = L.head_ch[hi + 1];
        for (int i = tid; i < NS * Cin; i += GT) {
          const int s = i / Cin, c = i - s * Cin;
          const float v = hin[s * maxw + c];
          hin[s * maxw + c] = a.head_elu ? (v > 0.f ? v : expm1f(v)) : fmaxf(v, 0.f);
        }
        csync();
        matvec<NS>(st + L.hw[hi], Cin, Cout, hin, maxw, part, hout, maxw, st + L.hb[hi], nullptr, 0);
        float* tmp = hin;
        hin = hout;
        hout = tmp;
      }
    }
    // hin now holds the logits for the next sample
    for (int i = tid; i < NS * L.Q; i += GT) {
      const int s = i / L.Q, q = i - s * L.Q;
      const int stream = s0 + s;
      if (stream < L.n) cur_logits[(int64_t)stream * L.Q + q] = hin[s * maxw + q];
    }
    csync();
  }

  if (a.probs) {
    for (int s = 0; s < NS; ++s) {
      const int stream = s0 + s;
      if (stream >= L.n) continue;
      const float* lg = cur_logits + (int64_t)stream * L.Q;
      if (!a.apply_softmax) {
        for (int q = tid; q < L.Q; q += GT) a.probs[(int64_t)stream * L.Q + q] = lg[q];
        continue;
      }
      float m = -INFINITY;
      for (int q = tid; q < L.Q; q += GT) m = fmaxf(m, lg[q]);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
      if ((tid & 31) == 0) s_redv[tid >> 5] = m;
      csync();
      m = s_redv[0];
      for (int w = 1; w < GT / 32; ++w) m = fmaxf(m, s_redv[w]);
      csync();
      float sum = 0.f;
      for (int q = tid; q < L.Q; q += GT) sum += expf(lg[q] - m);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      if ((tid & 31) == 0) s_redv[tid >> 5] = sum;
      csync();
      sum = 0.f;
      for (int w = 0; w < GT / 32; ++w) sum += s_redv[w];
      csync();
      for (int q = tid; q < L.Q; q += GT) a.probs[(int64_t)stream * L.Q + q] = expf(lg[q] - m) / sum;
    }
  }
}


// =============================================================================================
// Generator v3: specialised for networks whose matvecs tile exactly over 16 warps
// (2G = 128, R + S = 320, head widths 256 -- BASELINE config C).  Differences from the generic
// streamed kernel above:
//   * weights are pre-packed so that lane j of warp w reads ITS rows of ITS outputs with consecutive
//     LDS.128 (conflict-free), lanes split K, and partial sums are combined with a shuffle
//     reduce-scatter -- no shared-memory partials, no barrier between matvec and epilogue;
//   * the gate runs inside the warp that produced a_f/a_g, the skip accumulators stay in registers
//     for the whole step; two block barriers per layer remain (z visible, x visible).
constexpr int V3_STAGE_BYTES = 81920;   // one whole per-layer matrix per stage (WA 64 KB, WB 80 KB)
constexpr int V3_STAGES = 2;

// acc[s] += sum over this lane's rows of ITS output: per chunk NQ float4 of packed weights (4 consecutive rows of
// one output each) against NQ float4 of the input vector.  f4_off/q_stride locate the lane's float4s inside a
// chunk; rows [0,Ksplit) of the input come from xa, the rest from xb.
template <int NS, int NQ>
__device__ __forceinline__ void mv4(StreamCtx& cx, int nchunks, int f4_off, int q_stride, int chunkK, int row0,
                                    bool active, const float* xa, int lda, const float* xb, int ldb, int Ksplit,
                                    float (&acc)[NS]) {
  const int lane = threadIdx.x & 31;
  for (int c = 0; c < nchunks; ++c) {
    const uint32_t stage = cx.it % V3_STAGES, parity = (cx.it / V3_STAGES) & 1;
    tc::mbar_wait(cx.full0 + 8 * stage, parity);
    if (active) {
      const float4* wp = reinterpret_cast<const float4*>(cx.ring + stage * (V3_STAGE_BYTES / 4)) + f4_off;
      const int kk = c * chunkK + row0;
      const float* xp = kk < Ksplit ? xa + kk : xb + (kk - Ksplit);
      const int ld = kk < Ksplit ? lda : ldb;
#pragma unroll
      for (int q0 = 0; q0 < NQ; q0 += 4) {
        float4 wv[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) wv[q] = wp[(q0 + q) * q_stride];
#pragma unroll
        for (int s = 0; s < NS; ++s)
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 x = *reinterpret_cast<const float4*>(xp + s * ld + (q0 + q) * 4);
            acc[s] = fmaf(wv[q].x, x.x, acc[s]);
            acc[s] = fmaf(wv[q].y, x.y, acc[s]);
            acc[s] = fmaf(wv[q].z, x.z, acc[s]);
            acc[s] = fmaf(wv[q].w, x.w, acc[s]);
          }
      }
    }
    __syncwarp();
    if (lane == 0) tc::mbar_arrive(cx.empty0 + 8 * stage);
    ++cx.it;
  }
}

__device__ __forceinline__ float head_act(float v, int elu) { return elu ? (v > 0.f ? v : expm1f(v)) : fmaxf(v, 0.f); }

template <int NS>
__global__ void __launch_bounds__(GT + 32) gen_kernel_v3(GenArgs a) {
  extern __shared__ float sm[];
  const GenLayout& L = a.lay;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int s0 = blockIdx.x * NS;
  const int R = L.R, Q = L.Q;
  const int pastw = (L.k - 1) * R;
  // smem carve-up (floats)
  float* xv = sm;                               // [NS][R]
  float* zv = xv + NS * R;                      // [NS][128] (G <= 64 used)
  float* xpast = zv + NS * 128;                 // [NS][L][pastw]
  float* hA = xpast + NS * L.L * pastw;         // [NS][256]
  float* hB = hA + NS * 256;                    // [NS][256]
  float* hbias = hB + NS * 256;                 // [n_head][256]
  uint8_t* ring_g = reinterpret_cast<uint8_t*>(hbias + L.n_head * 256);
  ring_g += (128 - (tc::smem_u32(ring_g) & 127)) & 127;
  __shared__ int s_sample[NS];
  __shared__ __align__(8) uint64_t s_bars[2 * V3_STAGES];
  __shared__ GenChunk s_sched[MAX_CHUNKS];
  __shared__ GenLayerOff s_layers[128];
  __shared__ float s_redv[GT / 32];
  __shared__ int s_redi[GT / 32];
  StreamCtx cx;
  cx.it = 0;
  cx.ring = reinterpret_cast<const float*>(ring_g);
  cx.full0 = tc::smem_u32(&s_bars[0]);
  cx.empty0 = tc::smem_u32(&s_bars[V3_STAGES]);
  float* st = a.state;
  if (threadIdx.x == 0) {
    for (int i = 0; i < V3_STAGES; ++i) {
      tc::mbar_init(cx.full0 + 8 * i, 1);
      tc::mbar_init(cx.empty0 + 8 * i, GT / 32);
    }
    tc::fence_barrier_init();
  }
  for (int i = threadIdx.x; i < a.n_chunks; i += blockDim.x) s_sched[i] = a.chunks[i];
  for (int i = threadIdx.x; i < L.L; i += blockDim.x) s_layers[i] = a.layers[i];
  for (int i = threadIdx.x; i < L.n_head * 256; i += blockDim.x) hbias[i] = L.has_hb ? st[L.hb[i / 256] + (i % 256)] : 0.f;
  float* cur_logits = st + L.cur_logits;
  for (int i = threadIdx.x; i < NS * Q; i += blockDim.x) {
    const int stream = s0 + i / Q;
    hA[i] = stream < L.n ? cur_logits[(int64_t)stream * Q + (i % Q)] : 0.f;   // logits of the next sample
  }
  __syncthreads();
  if (threadIdx.x >= GT) {
    if (threadIdx.x == GT) {   // producer: stream the packed weights of every step in consumption order
      const uint8_t* sbase = reinterpret_cast<const uint8_t*>(a.state);
      const uint32_t ring_s = tc::smem_u32(ring_g);
      uint32_t it = 0;
      for (int step = 0; step < a.n_steps; ++step)
        for (int c = 0; c < a.n_chunks; ++c, ++it) {
          const GenChunk ch = s_sched[c];
          const uint32_t stage = it % V3_STAGES;
          tc::mbar_wait(cx.empty0 + 8 * stage, ((it / V3_STAGES) & 1) ^ 1);
          tc::mbar_arrive_expect_tx(cx.full0 + 8 * stage, ch.bytes);
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                           ring_s + stage * V3_STAGE_BYTES),
                       "l"(reinterpret_cast<uint64_t>(sbase + ch.off)), "r"(ch.bytes), "r"(cx.full0 + 8 * stage)
                       : "memory");
        }
    }
    return;
  }
  int32_t* idx_hist = (int32_t*)(st + L.idx_hist);
  const int kc1 = L.kc - 1;
  float* lg = hA;                       // logits of the next sample live in smem between steps
  // lane roles: WA  -> output ol = lane&7 of the warp's 8 (4 a_f + 4 a_g), K slice sl = lane>>3 (16 rows per chunk)
  //             WB  -> thread tid owns output tid (< R+S), all 32 rows of a chunk
  //             head-> output ol = lane&15 of the warp's 16, K slice lane>>4 (16 rows per chunk)
  const int olA = lane & 7, slA = lane >> 3;
  const int olH = lane & 15, slH = lane >> 4;
  const bool actB = tid < R + 256;

  for (int step = 0; step < a.n_steps; ++step) {
    const int64_t t = a.t0 + step;
    // ---- 1. sample (argmax, optionally Gumbel-perturbed; ties -> lowest index) ----
    for (int s = 0; s < NS; ++s) {
      const int stream = s0 + s;
      if (!a.sample_first || stream >= L.n) {
        if (tid == 0) s_sample[s] = (stream < L.n && !a.sample_first) ? a.forced[stream] : 0;
        continue;
      }
      float bv = -INFINITY;
      int bi = 0x7fffffff;
      for (int q = tid; q < Q; q += GT) {
        float v = lg[s * 256 + q];
        if (a.mode == WN_GEN_SAMPLE) v += gumbel(a.seed, (uint64_t)stream, (uint64_t)t, (uint32_t)q);
        if (v > bv || (v == bv && q < bi)) {
          bv = v;
          bi = q;
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > bv || (ov == bv && oi < bi)) {
          bv = ov;
          bi = oi;
        }
      }
      if (lane == 0) {
        s_redv[warp] = bv;
        s_redi[warp] = bi;
      }
      csync();
      if (tid == 0) {
        for (int w = 1; w < GT / 32; ++w)
          if (s_redv[w] > bv || (s_redv[w] == bv && s_redi[w] < bi)) {
            bv = s_redv[w];
            bi = s_redi[w];
          }
        s_sample[s] = bi;
        if (a.out) a.out[(int64_t)stream * a.n_steps + step] = bi;
      }
      csync();
    }
    // ---- 2. past taps of every layer (already in the rings) + embedding of the new sample ----
    {
      const int per = pastw >> 2;
      for (int i = tid; i < NS * L.L * per; i += GT) {
        const int v4 = i % per, sl_ = i / per;
        const int l = sl_ % L.L, s = sl_ / L.L;
        const int stream = s0 + s;
        const GenLayerOff& ly = s_layers[l];
        const int tap = (v4 * 4) / R, c = v4 * 4 - tap * R;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (stream < L.n) {
          const int len = ly.ring_len;
          const int64_t tau = t - (int64_t)(L.k - 1 - tap) * ly.dilation;
          const int slot = (int)(((tau % len) + len) % len);
          v = *reinterpret_cast<const float4*>(st + ly.ring + ((int64_t)stream * len + slot) * R + c);
        }
        *reinterpret_cast<float4*>(xpast + (s * L.L + l) * pastw + v4 * 4) = v;
      }
    }
    csync();   // s_sample visible
    {
      const float* emb = st + L.emb;
      const float* eb = st + L.emb_b;
      for (int i = tid; i < NS * R; i += GT) {
        const int s = i / R, r = i - s * R;
        const int stream = s0 + s;
        float v = L.has_cb ? eb[r] : 0.f;
        if (stream < L.n)
          for (int j = 0; j < L.kc; ++j) {
            const int q = j == kc1 ? s_sample[s] : idx_hist[(int64_t)stream * kc1 + j];
            if (q >= 0) v += emb[((int64_t)j * Q + q) * R + r];
          }
        xv[s * R + r] = v;
      }
    }
    csync();
    if (kc1 > 0 && tid < NS) {
      const int stream = s0 + tid;
      if (stream < L.n) {
        for (int j = 0; j + 1 < kc1; ++j) idx_hist[(int64_t)stream * kc1 + j] = idx_hist[(int64_t)stream * kc1 + j + 1];
        idx_hist[(int64_t)stream * kc1 + kc1 - 1] = s_sample[tid];
      }
    }
    // ---- 3. residual layers ----
    float skr[NS];   // this thread's skip-sum channel (tid - R) stays in a register for the whole step
#pragma unroll
    for (int s = 0; s < NS; ++s) skr[s] = 0.f;
    for (int l = 0; l < L.L; ++l) {
      const GenLayerOff& ly = s_layers[l];
      const int len = ly.ring_len, G = ly.G;
      for (int i = tid; i < NS * R; i += GT) {   // x[t] into the ring (roll, faster_wavenet.py:90-91)
        const int s = i / R, c = i - s * R;
        const int stream = s0 + s;
        if (stream < L.n) st[ly.ring + ((int64_t)stream * len + (int)(t % len)) * R + c] = xv[s * R + c];
      }
      {
        float acc[NS];
#pragma unroll
        for (int s = 0; s < NS; ++s) acc[s] = 0.f;
        mv4<NS, 8>(cx, 1, (warp * 8) * 32 + lane, 32, L.k * R, slA * 32, true, xpast + l * pastw, L.L * pastw, xv, R, pastw,
                   acc);
#pragma unroll
        for (int s = 0; s < NS; ++s) {
          float v = acc[s];
          v += __shfl_xor_sync(0xffffffffu, v, 8);
          v += __shfl_xor_sync(0xffffffffu, v, 16);           // every lane: full a_f / a_g of its output
          const float other = __shfl_xor_sync(0xffffffffu, v, 4);   // ol<4 holds a_f, ol+4 the matching a_g
          if (slA == 0 && olA < 4) {
            const int ch = 4 * warp + olA;
            float f = v, gg = other;
            if (ly.has_ba) {
              f += st[ly.ba + ch];
              gg += st[ly.ba + G + ch];
            }
            zv[s * 128 + ch] = tanhf(f) * (1.f / (1.f + expf(-gg)));   // wavenet.py:351
          }
        }
      }
      csync();
      {
        float acc[NS];
#pragma unroll
        for (int s = 0; s < NS; ++s) acc[s] = 0.f;
        mv4<NS, 16>(cx, 1, tid, R + 256, G, 0, actB, zv, 128, zv, 128, G, acc);
        if (actB) {
          const float bb = ly.has_bb ? st[ly.bb + tid] : 0.f;
#pragma unroll
          for (int s = 0; s < NS; ++s) {
            if (tid < R)
              xv[s * R + tid] += acc[s] + bb;      // output = projection_block + x, wavenet.py:354
            else
              skr[s] += acc[s] + bb;               // sum_skip_connections += z, faster_wavenet.py:100
          }
        }
      }
      csync();
    }
    // ---- 4. head (faster_wavenet.py:105-113: ELU on incremental steps; ReLU variant) ----
    if (actB && tid >= R) {
#pragma unroll
      for (int s = 0; s < NS; ++s) hB[s * 256 + (tid - R)] = head_act(skr[s], a.head_elu);
    }
    csync();
    float* hin = hB;
    float* hout = hA;
    for (int hi = 0; hi < L.n_head; ++hi) {
      float acc[NS];
#pragma unroll
      for (int s = 0; s < NS; ++s) acc[s] = 0.f;
      mv4<NS, 8>(cx, L.head_ch[hi] / 64, (warp * 8) * 32 + lane, 32, 64, slH * 32, true, hin, 256, hin, 256, L.head_ch[hi],
                 acc);
      const bool last = hi == L.n_head - 1;
#pragma unroll
      for (int s = 0; s < NS; ++s) {
        float v = acc[s] + __shfl_xor_sync(0xffffffffu, acc[s], 16);
        if (slH == 0) {
          const int o = 16 * warp + olH;
          v += hbias[hi * 256 + o];
          if (!last) v = head_act(v, a.head_elu);
          hout[s * 256 + o] = v;
        }
      }
      csync();
      float* tmp = hin;
      hin = hout;
      hout = tmp;
    }
    lg = hin;   // logits for the next sample
  }
  // ---- epilogue: publish logits (and probabilities for the step API) ----
  for (int i = tid; i < NS * Q; i += GT) {
    const int s = i / Q, q = i - s * Q;
    const int stream = s0 + s;
    if (stream < L.n) cur_logits[(int64_t)stream * Q + q] = lg[s * 256 + q];
  }
  if (a.probs) {
    for (int s = 0; s < NS; ++s) {
      const int stream = s0 + s;
      if (stream >= L.n) continue;
      const float* lgs = lg + s * 256;
      if (!a.apply_softmax) {
        for (int q = tid; q < Q; q += GT) a.probs[(int64_t)stream * Q + q] = lgs[q];
        continue;
      }
      float m = -INFINITY;
      for (int q = tid; q < Q; q += GT) m = fmaxf(m, lgs[q]);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
      if (lane == 0) s_redv[warp] = m;
      csync();
      m = s_redv[0];
      for (int w = 1; w < GT / 32; ++w) m = fmaxf(m, s_redv[w]);
      csync();
      float sum = 0.f;
      for (int q = tid; q < Q; q += GT) sum += expf(lgs[q] - m);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      if (lane == 0) s_redv[warp] = sum;
      csync();
      sum = 0.f;
      for (int w = 0; w < GT / 32; ++w) sum += s_redv[w];
      csync();
      for (int q = tid; q < Q; q += GT) a.probs[(int64_t)stream * Q + q] = expf(lgs[q] - m) / sum;
    }
  }
}

// dst (packed so that every lane reads consecutive float4s) <- src [K][N].  One float4 = 4 consecutive rows of one
// output.  mode 1 (gated conv, 8 outputs/warp = 4 a_f + 4 a_g, 4 K slices/warp), mode 2 (one output per thread),
// mode 3 (16 outputs/warp, 2 K slices/warp).
__global__ void gen_pack_v3(const float* __restrict__ src, float* __restrict__ dst, int K, int N, int chunkK, int mode) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= K * N) return;
  const int k = idx / N, n = idx % N;
  const int c = k / chunkK, kr = k % chunkK, comp = kr & 3;
  int64_t f4;
  if (mode == 1) {          // 4 K slices per warp, 8 outputs per warp (4 a_f + 4 a_g)
    const int G = N / 2, ch = n < G ? n : n - G;
    const int w = ch / 4, ol = (n < G ? 0 : 4) + ch % 4;
    const int rps = chunkK / 4, nq = rps / 4;
    const int sl = kr / rps, q = (kr % rps) / 4;
    f4 = (int64_t)(w * nq + q) * 32 + sl * 8 + ol;
  } else if (mode == 2) {   // one output per thread
    f4 = (int64_t)(kr / 4) * N + n;
  } else {                  // 2 K slices per warp, 16 outputs per warp
    const int w = n / 16, ol = n % 16;
    const int rps = chunkK / 2, nq = rps / 4;
    const int sl = kr / rps, q = (kr % rps) / 4;
    f4 = (int64_t)(w * nq + q) * 32 + sl * 16 + ol;
  }
  dst[(int64_t)c * chunkK * N + f4 * 4 + comp] = src[idx];
}

size_t gen_smem_bytes_v3(const GenLayout& L, int NS) {
  size_t f = (size_t)NS * L.R + NS * 128 + (size_t)NS * L.L * (L.k - 1) * L.R + 2 * NS * 256 + L.n_head * 256 + 64;
  return f * sizeof(float) + V3_STAGES * V3_STAGE_BYTES + 128;
}

template <int NS>
int launch_gen_v3(const GenArgs& a, cudaStream_t s) {
  const size_t smem = gen_smem_bytes_v3(a.lay, NS);
  WN_CHECK_CUDA(cudaFuncSetAttribute(gen_kernel_v3<NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = (a.lay.n + NS - 1) / NS;
  gen_kernel_v3<NS><<<grid, GT + 32, smem, s>>>(a);
  WN_CHECK_LAUNCH();
  return WN_OK;
}

// =============================================================================================
// Generator v4: ONE STREAM PER 8-CTA CLUSTER (latency path: batch 1 .. sm_count/8 streams).
// At batch 1 the v3 kernel is bound by what one SM can pull from L2 (5 MB of weights per audio sample).  Here the
// eight CTAs of a cluster split every matvec by OUTPUT rows -- each streams only its 1/8 slice of the packed weights
// (604 KB per sample) -- and exchange the 64-float activations through distributed shared memory: a thread that
// finishes an output sends it to all eight CTAs with st.async (a remote store that completes 4 bytes on the destination's
// mbarrier); consumers wait for the byte count.  ONE exchange per layer (z) + one per head conv: the residual update
// x += Wp z is computed redundantly by every CTA from the full z (16 KB more weights per layer, but no second exchange).
// Sampling and the embedding of the next sample are computed redundantly by every CTA (same logits, same counter RNG),
// so no broadcast of the sample is needed.
constexpr int V4_CS = 8;                  // CTAs per cluster
constexpr int V4_T = 256;                 // consumer threads per CTA (+ one producer warp)
constexpr int V4_STAGE_BYTES = 32768;     // one chunk per stage: a layer's WA|WB slice or a head slice (32 KB each)
constexpr int V4_STAGES = 4;
constexpr int V4_WA_F = 2 * 256 * 4;      // floats of a packed WA slice: [q 2][thread 256] float4
constexpr int V4_WB_F = 8 * 192 * 4;      // WB slice: ALL 64 residual rows + this rank's 32 skip rows: [q 8][thread 192] float4
constexpr int V4_WH_F = 8 * 256 * 4;      // head slice: [q 8][thread 256] float4

// tanh on the dependency chain of the cluster generator: 1 - 2 / (e^(2x) + 1) with one MUFU.EX2 and one MUFU.RCP (five
// dependent instructions; tanhf() is ~20).  Absolute error < 2e-7 for every x (saturates cleanly: e = inf -> 1, e = 0 -> -1).
__device__ __forceinline__ float tanh_ex2(float x) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * 2.8853900817779268f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(e + 1.f));
  return fmaf(-2.f, r, 1.f);
}

__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t map_to_cta(uint32_t smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}
// remote store whose completion is signalled on the DESTINATION CTA's mbarrier (complete_tx of 4 bytes): no separate
// arrive and no release round trip per peer -- eight fire-and-forget messages per value
__device__ __forceinline__ void st_async_f32(uint32_t addr, float v, uint32_t bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(addr),
               "r"(__float_as_uint(v)), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// one thread per CTA opens the phase (expects `bytes` from the cluster), everybody waits for it
// The wait acquires at CLUSTER scope: besides the st.async payload it orders this CTA's later reads of the global rings
// behind the ring stores the peers made before they sent (store -> block barrier -> st.async/complete_tx -> this wait).
__device__ __forceinline__ void exchange_wait(uint32_t bar, uint32_t parity, uint32_t bytes, int tid) {
  if (tid == 0) tc::mbar_arrive_expect_tx(bar, bytes);
  uint32_t spins = 0;
  for (;;) {
    uint32_t ok;
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (ok) return;
    if (++spins > (1u << 24)) {
      printf("wavenet_b200: generator v4 exchange timeout (block %d thread %d bar 0x%x parity %u)\n", (int)blockIdx.x,
             (int)threadIdx.x, bar, parity);
      __trap();
    }
  }
}

// Developer trace (make EXTRA=-DWN_LAYER_TRACE, tests/dev/trace_gen.py): clock64 stamps of CTA 0 / thread 0 in step 2
#ifdef WN_LAYER_TRACE
__device__ long long g_trace_gen[64 * 16];
#define TRG(l, e) do { if (blockIdx.x == 0 && tid == 0 && step == 2 && (l) < 64) g_trace_gen[(l) * 16 + (e)] = clock64(); } while (0)
#else
#define TRG(l, e) do { } while (0)
#endif
__global__ void __cluster_dims__(V4_CS, 1, 1) __launch_bounds__(V4_T + 32) gen_kernel_v4(GenArgs a) {
  extern __shared__ float sm[];
  const GenLayout& L = a.lay;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int rank = (int)cluster_rank();
  const int stream = blockIdx.x / V4_CS;
  constexpr int R = 64, G = 64, Q = 256;
  float* xv = sm;                       // [64]   layer input (all channels, filled by the exchange)
  float* zv = xv + 64;                  // [2][64]  double-buffered by layer parity (skip-only threads read z without
                                        //          sending anything afterwards, so the next z must not land on top of it)
  float* hA = zv + 128;                 // [256]
  float* hB = hA + 256;                 // [256]
  float* xpast = hB + 256;              // [L][64]
  float* hbias = xpast + L.L * 64;      // [n_head][256]
  uint8_t* ring_g = reinterpret_cast<uint8_t*>(hbias + L.n_head * 256);
  ring_g += (128 - (tc::smem_u32(ring_g) & 127)) & 127;
  __shared__ __align__(8) uint64_t s_bars[2 * V4_STAGES + 4];
  __shared__ GenLayerOff s_layers[128];
  __shared__ float s_redv[V4_T / 32];
  __shared__ int s_redi[V4_T / 32];
  __shared__ int s_sample;
  __shared__ int s_pos[128];            // ring slot of time t per layer (t mod ring_len), advanced once per step: no 64-bit
                                        // modulo on the per-layer critical path
  const uint32_t full0 = tc::smem_u32(&s_bars[0]), empty0 = tc::smem_u32(&s_bars[V4_STAGES]);
  // exchange barriers.  Consecutive exchanges never use the same barrier (the layers alternate zbar0/zbar1, the head
  // hbar0/hbar1): a peer that is one exchange ahead then completes its bytes on a barrier whose previous phase is over
  // in every CTA, so byte counts of different phases cannot mix.
  const uint32_t zbar0 = tc::smem_u32(&s_bars[2 * V4_STAGES]), zbar1 = zbar0 + 8, hbar0 = zbar0 + 16, hbar1 = zbar0 + 24;
  const uint32_t xv_s = tc::smem_u32(xv), zv_s = tc::smem_u32(zv), hA_s = tc::smem_u32(hA), hB_s = tc::smem_u32(hB);
  float* st = a.state;
  if (tid == 0) {
    for (int i = 0; i < V4_STAGES; ++i) {
      tc::mbar_init(full0 + 8 * i, 1);
      tc::mbar_init(empty0 + 8 * i, V4_T / 32);
    }
    tc::mbar_init(zbar0, 1);            // phases are opened by tid 0 with the expected byte count
    tc::mbar_init(zbar1, 1);
    tc::mbar_init(hbar0, 1);
    tc::mbar_init(hbar1, 1);
    tc::fence_barrier_init();
  }
  for (int i = tid; i < L.L; i += blockDim.x) {
    s_layers[i] = a.layers[i];
    s_pos[i] = (int)(a.t0 % a.layers[i].ring_len);
  }
  for (int i = tid; i < L.n_head * 256; i += blockDim.x) hbias[i] = L.has_hb ? st[L.hb[i / 256] + (i % 256)] : 0.f;
  float* cur_logits = st + L.cur_logits;
  for (int i = tid; i < Q; i += blockDim.x) hA[i] = cur_logits[(int64_t)stream * Q + i];
  __syncthreads();
  cluster_sync_all();                   // every CTA's barriers exist before anybody arrives on them
  const int n_chunks = L.L + L.n_head;
  if (tid >= V4_T) {
    if (tid == V4_T) {                  // producer: this rank's slices of every step, in consumption order
      const uint8_t* wbase = reinterpret_cast<const uint8_t*>(st + L.wpk4 + (int64_t)rank * L.wpk4_rank);
      const uint32_t ring_s = tc::smem_u32(ring_g);
      uint32_t it = 0;
      for (int step = 0; step < a.n_steps; ++step) {
        uint64_t off = 0;
        for (int c = 0; c < n_chunks; ++c, ++it) {
          const uint32_t bytes = (c < L.L ? (V4_WA_F + V4_WB_F) : V4_WH_F) * 4;   // 32 KB either way
          const uint32_t stage = it % V4_STAGES;
          tc::mbar_wait(empty0 + 8 * stage, ((it / V4_STAGES) & 1) ^ 1);
          tc::mbar_arrive_expect_tx(full0 + 8 * stage, bytes);
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                           ring_s + stage * V4_STAGE_BYTES),
                       "l"(reinterpret_cast<uint64_t>(wbase + off)), "r"(bytes), "r"(full0 + 8 * stage)
                       : "memory");
          off += bytes;
        }
      }
    }
    __syncwarp();
    cluster_sync_all();                 // matches the consumers' final cluster barrier
    return;
  }
  int32_t* idx_hist = (int32_t*)(st + L.idx_hist);
  const int kc1 = L.kc - 1;             // 0 or 1 previous samples feed the causal (embedding) layer
  int prev_q = kc1 > 0 ? idx_hist[(int64_t)stream * kc1 + kc1 - 1] : -1;
  float* lg = hA;
  uint32_t it = 0, zph[2] = {0, 0}, hph[2] = {0, 0}, zsel = 0;
  // phase A roles: warp w -> gate channel 8*rank + w; lanes 0..15 a_f, 16..31 a_g; K slice = lane & 15 (8 rows)
  // phase B roles: thread t < 192 -> output t >> 1 (0..63 residual channel o -- every CTA computes all of them --,
  //                64..95 skip channel 32*rank + o - 64), K half t & 1
  // head roles   : output tid >> 3 (32*rank + o), K slice tid & 7 (32 rows)
  const int ksA = lane & 15;
  const int oB = tid >> 1, ksB = tid & 1;
  const int oH = tid >> 3, ksH = tid & 7;

  for (int step = 0; step < a.n_steps; ++step) {
    const int64_t t = a.t0 + step;
    // ---- 1. sample (every CTA computes the same value) ----
    TRG(41, 0);
    if (!a.sample_first) {
      if (tid == 0) s_sample = a.forced[stream];
    } else {
      float bv = lg[tid];
      int bi = tid;
      if (a.mode == WN_GEN_SAMPLE) bv += gumbel(a.seed, (uint64_t)stream, (uint64_t)t, (uint32_t)tid);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > bv || (ov == bv && oi < bi)) {
          bv = ov;
          bi = oi;
        }
      }
      if (lane == 0) {
        s_redv[warp] = bv;
        s_redi[warp] = bi;
      }
      csync4();
      if (tid == 0) {
        for (int w = 1; w < V4_T / 32; ++w)
          if (s_redv[w] > bv || (s_redv[w] == bv && s_redi[w] < bi)) {
            bv = s_redv[w];
            bi = s_redi[w];
          }
        s_sample = bi;
        if (a.out && rank == 0) a.out[(int64_t)stream * a.n_steps + step] = bi;
      }
    }
    // ---- 2. past taps of every layer (rings live in global memory; written at least one step ago) ----
    for (int i = tid; i < L.L * 16; i += V4_T) {
      const int l = i >> 4, v4 = i & 15;
      const GenLayerOff& ly = s_layers[l];
      const int len = ly.ring_len;      // == dilation (k = 2): x[t - d] sits in the slot x[t] is about to take
      const float4 v = __ldcg(reinterpret_cast<const float4*>(st + ly.ring + ((int64_t)stream * len + s_pos[l]) * R) + v4);
      *reinterpret_cast<float4*>(xpast + l * 64 + v4 * 4) = v;
    }
    csync4();                           // s_sample visible
    if (tid < R) {
      const float* emb = st + L.emb;
      float v = L.has_cb ? st[L.emb_b + tid] : 0.f;
      const int q_new = s_sample;
      if (kc1 > 0 && prev_q >= 0) v += emb[((int64_t)0 * Q + prev_q) * R + tid];
      v += emb[((int64_t)kc1 * Q + q_new) * R + tid];
      xv[tid] = v;
    }
    prev_q = s_sample;
    csync4();
    // ---- 3. residual layers ----
    TRG(41, 1);
    float skr = 0.f;                    // skip-sum channel 32*rank + oB - 64 (threads with ksB == 0, oB >= 64)
    for (int l = 0; l < L.L; ++l, ++it) {
      // layer constants into registers before the weights wait (shared-memory loads off the critical path)
      const GenLayerOff& ly = s_layers[l];
      const int64_t ring_slot = ly.ring + ((int64_t)stream * ly.ring_len + s_pos[l]) * R;
      const bool has_ba = ly.has_ba != 0, has_bb = ly.has_bb != 0;
      const int64_t ba_off = ly.ba, bb_off = ly.bb;
      const uint32_t stage = it % V4_STAGES;
      TRG(l, 0);
      tc::mbar_wait(full0 + 8 * stage, (it / V4_STAGES) & 1);
      TRG(l, 1);
      const float4* wst = reinterpret_cast<const float4*>(ring_g + stage * V4_STAGE_BYTES);
      {
        const float* xin = ksA < 8 ? xpast + l * 64 + ksA * 8 : xv + (ksA - 8) * 8;
        const float4 w0 = wst[tid], w1 = wst[256 + tid];
        const float4 x0 = *reinterpret_cast<const float4*>(xin), x1 = *reinterpret_cast<const float4*>(xin + 4);
        float acc = w0.x * x0.x, acc1 = w1.x * x1.x;    // two chains: the reduction below is latency-bound
        acc = fmaf(w0.y, x0.y, acc), acc = fmaf(w0.z, x0.z, acc), acc = fmaf(w0.w, x0.w, acc);
        acc1 = fmaf(w1.y, x1.y, acc1), acc1 = fmaf(w1.z, x1.z, acc1), acc1 = fmaf(w1.w, x1.w, acc1);
        acc += acc1;
        acc += __shfl_xor_sync(0xffffffffu, acc, 1);
        acc += __shfl_xor_sync(0xffffffffu, acc, 2);
        acc += __shfl_xor_sync(0xffffffffu, acc, 4);
        acc += __shfl_xor_sync(0xffffffffu, acc, 8);
        // lanes 0..15 hold a_f, lanes 16..31 a_g.  z = tanh(a_f) * sigmoid(a_g) (wavenet.py:351) with
        // sigmoid(g) = 0.5 + 0.5 tanh(g/2): both halves then run the SAME tanhf at the same time instead of tanhf
        // followed by expf and a division on one thread; lanes 0..7 each send the result to one CTA.
        const int ch = 8 * rank + warp;
        float v = lane < 16 ? acc : 0.5f * acc;
        if (has_ba) v += lane < 16 ? st[ba_off + ch] : 0.5f * st[ba_off + G + ch];
        const float th = tanh_ex2(v);
        const float zval = __shfl_sync(0xffffffffu, th, 0) * (0.5f + 0.5f * __shfl_sync(0xffffffffu, th, 16));
        if (lane < V4_CS) {
          const uint32_t zb = zsel ? zbar1 : zbar0;
          st_async_f32(map_to_cta(zv_s + zsel * 256 + 4u * (uint32_t)ch, (uint32_t)lane), zval, map_to_cta(zb, (uint32_t)lane));
        }
      }
      const float xring = tid < 8 ? xv[8 * rank + tid] : 0.f;   // this CTA's ring slice, read before xv is updated below
      TRG(l, 2);
      exchange_wait(zsel ? zbar1 : zbar0, zph[zsel], 64 * 4, tid);
      TRG(l, 3);
      zph[zsel] ^= 1;
      if (tid < 192) {
        const float4* wb = wst + V4_WA_F / 4;
        float a4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int q = 0; q < 8; ++q) {                   // four independent chains
          const float4 w = wb[q * 192 + tid];
          const float4 z = *reinterpret_cast<const float4*>(zv + zsel * 64 + ksB * 32 + q * 4);
          a4[q & 3] = fmaf(w.x, z.x, a4[q & 3]), a4[q & 3] = fmaf(w.y, z.y, a4[q & 3]);
          a4[q & 3] = fmaf(w.z, z.z, a4[q & 3]), a4[q & 3] = fmaf(w.w, z.w, a4[q & 3]);
        }
        float acc = (a4[0] + a4[1]) + (a4[2] + a4[3]);
        acc += __shfl_xor_sync(0xffffffffu, acc, 1);
        if (ksB == 0) {
          if (oB < 64) {
            const float bb = has_bb ? st[bb_off + oB] : 0.f;
            xv[oB] += acc + bb;                                                // output = projection + x, wavenet.py:354
          } else {
            const float bb = has_bb ? st[bb_off + R + 32 * rank + oB - 64] : 0.f;
            skr += acc + bb;                                                   // faster_wavenet.py:100
          }
        }
      }
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(empty0 + 8 * stage);
      // this CTA's slice of x[t] into the ring (roll, faster_wavenet.py:90-91).  The slot is the one every CTA read as
      // x[t-d] at the top of the step: this point is behind the z exchange, hence behind all those reads.
      if (tid < 8) st[ring_slot + 8 * rank + tid] = xring;
      TRG(l, 4);
      csync4();                           // the new x (identical in every CTA) is complete
      TRG(l, 5);
      zsel ^= 1;
    }
    // ---- 4. head (faster_wavenet.py:105-113: ELU on incremental steps; ReLU variant) ----
    TRG(40, 0);
    int hsel = 0;
    if (tid < 192) {                    // the two lanes of a skip group send to four CTAs each
      const float hv = head_act(__shfl_sync(0xffffffffu, skr, lane & ~1), a.head_elu);
      if (oB >= 64) {
#pragma unroll
        for (uint32_t r = 0; r < 4; ++r) {
          const uint32_t dst = 4u * (uint32_t)ksB + r;
          st_async_f32(map_to_cta(hB_s + 4u * (uint32_t)(32 * rank + oB - 64), dst), hv, map_to_cta(hbar0, dst));
        }
      }
    }
    exchange_wait(hbar0, hph[0], 256 * 4, tid);
    hph[0] ^= 1;
    hsel = 1;
    float* hin = hB;
    float* hout = hA;
    uint32_t hout_s = hA_s, hin_s = hB_s;
    for (int hi = 0; hi < L.n_head; ++hi, ++it) {
      const uint32_t stage = it % V4_STAGES;
      tc::mbar_wait(full0 + 8 * stage, (it / V4_STAGES) & 1);
      const float4* wst = reinterpret_cast<const float4*>(ring_g + stage * V4_STAGE_BYTES);
      float acc = 0.f;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float4 w = wst[q * 256 + tid];
        const float4 x = *reinterpret_cast<const float4*>(hin + ksH * 32 + q * 4);
        acc = fmaf(w.x, x.x, acc), acc = fmaf(w.y, x.y, acc), acc = fmaf(w.z, x.z, acc), acc = fmaf(w.w, x.w, acc);
      }
      acc += __shfl_xor_sync(0xffffffffu, acc, 1);
      acc += __shfl_xor_sync(0xffffffffu, acc, 2);
      acc += __shfl_xor_sync(0xffffffffu, acc, 4);
      const bool last = hi == L.n_head - 1;
      {                                 // every lane of the 8-lane group holds the sum: lane ks sends to CTA ks
        const int o = 32 * rank + oH;
        float v = acc + hbias[hi * 256 + o];
        if (!last) v = head_act(v, a.head_elu);
        st_async_f32(map_to_cta(hout_s + 4u * (uint32_t)o, (uint32_t)ksH), v, map_to_cta(hsel ? hbar1 : hbar0, (uint32_t)ksH));
      }
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(empty0 + 8 * stage);
      exchange_wait(hsel ? hbar1 : hbar0, hph[hsel], 256 * 4, tid);
      hph[hsel] ^= 1;
      hsel ^= 1;
      float* tmp = hin;
      hin = hout;
      hout = tmp;
      const uint32_t ts = hin_s;
      hin_s = hout_s;
      hout_s = ts;
    }
    lg = hin;   // logits for the next sample
    if (tid < L.L) {                    // every thread is past this step's ring accesses (the head exchanges came after)
      const int p = s_pos[tid] + 1;
      s_pos[tid] = p == s_layers[tid].ring_len ? 0 : p;
    }
    TRG(40, 1);
  }
  // ---- epilogue (rank 0 publishes the stream's state) ----
  if (rank == 0) {
    for (int q = tid; q < Q; q += V4_T) cur_logits[(int64_t)stream * Q + q] = lg[q];
    if (kc1 > 0 && tid == 0) idx_hist[(int64_t)stream * kc1 + kc1 - 1] = prev_q;
    if (a.probs) {
      if (!a.apply_softmax) {
        for (int q = tid; q < Q; q += V4_T) a.probs[(int64_t)stream * Q + q] = lg[q];
      } else {
        float m = lg[tid];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        if (lane == 0) s_redv[warp] = m;
        csync4();
        m = s_redv[0];
        for (int w = 1; w < V4_T / 32; ++w) m = fmaxf(m, s_redv[w]);
        csync4();
        float sum = expf(lg[tid] - m);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        if (lane == 0) s_redv[warp] = sum;
        csync4();
        sum = 0.f;
        for (int w = 0; w < V4_T / 32; ++w) sum += s_redv[w];
        a.probs[(int64_t)stream * Q + tid] = expf(lg[tid] - m) / sum;
      }
    }
  }
  cluster_sync_all();   // nobody leaves while a peer may still store into its shared memory
}

// ------------------------------------------------------------------------------------------
// gen_kernel_v5: MANY streams on the cluster layout of v4 -- one 8-CTA cluster per V5_NS = 16 streams.
// v3 gives every CTA whole streams, so every SM pulls all 5 MB of weights through its shared memory each step (128 CTAs x
// 5 MB = 7.9 TB/s of L2 -> SMEM traffic at 256 streams, 82 us per step).  Here the eight CTAs of a cluster split every
// matrix by output rows exactly like v4 (same packed slices, 1 MB per CTA and step), keep each thread's weights in
// REGISTERS and sweep them over the cluster's 16 streams.  Differences from v4:
//  * exchanged activations are CHANNEL-major [channel][stream], so a CTA's slice of an exchange (its 8 gate channels or
//    32 head outputs x 16 streams) is one contiguous block: it is staged in local shared memory and sent with ONE
//    cp.async.bulk per destination CTA (shared::cta -> shared::cluster, complete_tx on the destination's mbarrier) -- 8
//    bulk copies per CTA and exchange.  (Per-value st.async as in v4 costs ~9 cycles per 4-byte message: 47 k messages
//    per step and CTA made the first version of this kernel 222 us per step.)
//  * phase A: a thread accumulates its K slice for all 16 streams, then a TRANSPOSING butterfly (15 shuffles instead of
//    16 x 4) leaves lane j of each half-warp with the complete sum of stream j -> ONE tanhf per lane serves 16 streams;
//  * the x(t-d) taps are prefetched three layers ahead into a 4-layer shared-memory ring (the all-layer table of v4 would
//    need 123 KB for 16 streams);
//  * sampling is warp-parallel over streams (warp w finishes the arg-max of streams 2w, 2w+1).
constexpr int V5_NS = 16;
constexpr int V5_XP = 4;                   // layers of x(t-d) taps resident in shared memory

__device__ __forceinline__ void bulk_copy_to_cta(uint32_t dst_cluster, uint32_t src_cta, uint32_t bytes, uint32_t bar_cluster) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_cluster),
               "r"(src_cta), "r"(bytes), "r"(bar_cluster)
               : "memory");
}

__global__ void __cluster_dims__(V4_CS, 1, 1) __launch_bounds__(V4_T + 32) gen_kernel_v5(GenArgs a) {
  extern __shared__ float sm[];
  const GenLayout& L = a.lay;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int rank = (int)cluster_rank();
  const int stream0 = (blockIdx.x / V4_CS) * V5_NS;
  const int ns = min(V5_NS, L.n - stream0);          // live streams of this cluster (the same in its eight CTAs); dead ones
                                                     // compute on mirrored inputs and are never stored
  constexpr int R = 64, G = 64, Q = 256, NS = V5_NS;
  float* xv = sm;                            // [NS][64]      layer input of every stream (local)
  float* zv = xv + NS * 64;                  // [2][64][NS]   gate outputs, channel-major, double-buffered by layer parity
  float* hA = zv + 2 * 64 * NS;              // [256][NS]     head activations / logits, channel-major
  float* hB = hA + 256 * NS;                 // [256][NS]
  float* xpast = hB + 256 * NS;              // [V5_XP][NS][64]
  float* stg = xpast + V5_XP * NS * 64;      // [2][32][NS]   this CTA's slice of an exchange (source of the bulk copies)
  float* hbias = stg + 2 * 32 * NS;          // [n_head][256]
  uint8_t* ring_g = reinterpret_cast<uint8_t*>(hbias + L.n_head * 256);
  ring_g += (128 - (tc::smem_u32(ring_g) & 127)) & 127;
  __shared__ __align__(8) uint64_t s_bars[2 * V4_STAGES + 4];
  __shared__ GenLayerOff s_layers[128];
  __shared__ float s_redv[NS][V4_T / 32];
  __shared__ int s_redi[NS][V4_T / 32];
  __shared__ int s_sample[NS], s_prev[NS];
  __shared__ int s_pos[128];
  const uint32_t full0 = tc::smem_u32(&s_bars[0]), empty0 = tc::smem_u32(&s_bars[V4_STAGES]);
  const uint32_t zbar0 = tc::smem_u32(&s_bars[2 * V4_STAGES]), zbar1 = zbar0 + 8, hbar0 = zbar0 + 16, hbar1 = zbar0 + 24;
  const uint32_t zv_s = tc::smem_u32(zv), hA_s = tc::smem_u32(hA), hB_s = tc::smem_u32(hB), stg_s = tc::smem_u32(stg);
  float* st = a.state;
  if (tid == 0) {
    for (int i = 0; i < V4_STAGES; ++i) {
      tc::mbar_init(full0 + 8 * i, 1);
      tc::mbar_init(empty0 + 8 * i, V4_T / 32);
    }
    tc::mbar_init(zbar0, 1);
    tc::mbar_init(zbar1, 1);
    tc::mbar_init(hbar0, 1);
    tc::mbar_init(hbar1, 1);
    tc::fence_barrier_init();
  }
  for (int i = tid; i < L.L; i += blockDim.x) {
    s_layers[i] = a.layers[i];
    s_pos[i] = (int)(a.t0 % a.layers[i].ring_len);
  }
  for (int i = tid; i < L.n_head * 256; i += blockDim.x) hbias[i] = L.has_hb ? st[L.hb[i / 256] + (i % 256)] : 0.f;
  float* cur_logits = st + L.cur_logits;
  for (int i = tid; i < NS * Q; i += blockDim.x) {
    const int q = i % Q, s = i / Q;
    hA[q * NS + s] = cur_logits[(int64_t)(stream0 + min(s, ns - 1)) * Q + q];
  }
  int32_t* idx_hist = (int32_t*)(st + L.idx_hist);
  const int kc1 = L.kc - 1;
  if (tid < NS) s_prev[tid] = kc1 > 0 ? idx_hist[(int64_t)(stream0 + min(tid, ns - 1)) * kc1 + kc1 - 1] : -1;
  __syncthreads();
  cluster_sync_all();
  const int n_chunks = L.L + L.n_head;
  if (tid >= V4_T) {
    if (tid == V4_T) {                       // producer: identical to v4 (this rank's packed slices, every step)
      const uint8_t* wbase = reinterpret_cast<const uint8_t*>(st + L.wpk4 + (int64_t)rank * L.wpk4_rank);
      const uint32_t ring_s = tc::smem_u32(ring_g);
      uint32_t it = 0;
      for (int step = 0; step < a.n_steps; ++step) {
        uint64_t off = 0;
        for (int c = 0; c < n_chunks; ++c, ++it) {
          const uint32_t bytes = (c < L.L ? (V4_WA_F + V4_WB_F) : V4_WH_F) * 4;
          const uint32_t stage = it % V4_STAGES;
          tc::mbar_wait(empty0 + 8 * stage, ((it / V4_STAGES) & 1) ^ 1);
          tc::mbar_arrive_expect_tx(full0 + 8 * stage, bytes);
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                           ring_s + stage * V4_STAGE_BYTES),
                       "l"(reinterpret_cast<uint64_t>(wbase + off)), "r"(bytes), "r"(full0 + 8 * stage)
                       : "memory");
          off += bytes;
        }
      }
    }
    __syncwarp();
    cluster_sync_all();
    return;
  }
  float* lg = hA;
  uint32_t it = 0, zph[2] = {0, 0}, hph[2] = {0, 0}, zsel = 0, xsel = 0;
  // phase A roles: warp w -> gate channel 8*rank + w; lanes 0..15 a_f, 16..31 a_g; K slice = lane & 15 (8 rows); after the
  //                transposing reduction lane j of each half owns stream j
  // phase B roles: thread t < 192 -> output t >> 1 (0..63 residual channel -- every CTA computes all --, 64..95 skip channel
  //                32*rank + o - 64), K half t & 1
  // head roles   : output tid >> 3 (32*rank + o), K slice tid & 7 (32 rows); after the reduction lane ks owns streams 2ks, 2ks+1
  const int ksA = lane & 15, jA = lane & 15;
  const int oB = tid >> 1, ksB = tid & 1;
  const int oH = tid >> 3, ksH = tid & 7;
  const int ps = tid >> 4, pv = tid & 15;    // x(t-d) prefetch: stream ps, float4 pv of the 64 channels
  const int psc = min(ps, ns - 1);
  auto past_ptr = [&](int l) {
    const GenLayerOff& ly = s_layers[l];
    return reinterpret_cast<const float4*>(st + ly.ring + ((int64_t)(stream0 + psc) * ly.ring_len + s_pos[l]) * R) + pv;
  };
  // this CTA's staged slice (rows x NS floats) -> the same rows of `dst_local` in every CTA of the cluster; one thread per
  // destination.  The staging buffer alternates between consecutive exchanges: a copy out of it has landed everywhere before
  // the exchange after next can start (peers send exchange e+1 only after they received all of e).
  auto send_slice = [&](uint32_t dst_local, int rows, uint32_t bar_local) {
    tc::fence_proxy_async();                 // staged values (generic stores) -> visible to the bulk-copy engine
    csync4();
    if (tid < V4_CS)
      bulk_copy_to_cta(map_to_cta(dst_local, (uint32_t)tid), stg_s + xsel * (32 * NS * 4), (uint32_t)(rows * NS * 4),
                       map_to_cta(bar_local, (uint32_t)tid));
    xsel ^= 1;
  };

  for (int step = 0; step < a.n_steps; ++step) {
    const int64_t t = a.t0 + step;
    TRG(41, 0);
    // ---- 1. sample: thread q holds logit q of every stream; warp-level arg-max per stream, then warp w finishes streams 2w
    //         and 2w+1 (every CTA of the cluster computes the same values) ----
    if (!a.sample_first) {
      if (tid < NS) s_sample[tid] = a.forced[stream0 + min(tid, ns - 1)];
    } else {
      float lv[NS];
#pragma unroll
      for (int g4 = 0; g4 < NS / 4; ++g4) {
        const float4 v = *reinterpret_cast<const float4*>(lg + tid * NS + g4 * 4);
        lv[4 * g4] = v.x, lv[4 * g4 + 1] = v.y, lv[4 * g4 + 2] = v.z, lv[4 * g4 + 3] = v.w;
      }
#pragma unroll 4
      for (int s = 0; s < NS; ++s) {
        float bv = lv[s];
        int bi = tid;
        if (a.mode == WN_GEN_SAMPLE) bv += gumbel(a.seed, (uint64_t)(stream0 + s), (uint64_t)t, (uint32_t)tid);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
          const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
          if (ov > bv || (ov == bv && oi < bi)) {
            bv = ov;
            bi = oi;
          }
        }
        if (lane == 0) {
          s_redv[s][warp] = bv;
          s_redi[s][warp] = bi;
        }
      }
      csync4();
      if (lane < 2) {
        const int s = 2 * warp + lane;
        float bv = s_redv[s][0];
        int bi = s_redi[s][0];
        for (int w = 1; w < V4_T / 32; ++w)
          if (s_redv[s][w] > bv || (s_redv[s][w] == bv && s_redi[s][w] < bi)) {
            bv = s_redv[s][w];
            bi = s_redi[s][w];
          }
        s_sample[s] = bi;
        if (a.out && rank == 0 && s < ns) a.out[(int64_t)(stream0 + s) * a.n_steps + step] = bi;
      }
    }
    // ---- 2. x(t-d) taps of the first three layers (written at least one step ago) ----
    {
      const float4 p0 = __ldcg(past_ptr(0));
      const float4 p1 = L.L > 1 ? __ldcg(past_ptr(1)) : p0;
      const float4 p2 = L.L > 2 ? __ldcg(past_ptr(2)) : p0;
      *reinterpret_cast<float4*>(xpast + (0 * NS + ps) * 64 + pv * 4) = p0;
      *reinterpret_cast<float4*>(xpast + (1 * NS + ps) * 64 + pv * 4) = p1;
      *reinterpret_cast<float4*>(xpast + (2 * NS + ps) * 64 + pv * 4) = p2;
    }
    csync4();                                // s_sample visible
    {
      const float* emb = st + L.emb;
      const int c = tid & 63;
      const float bias = L.has_cb ? st[L.emb_b + c] : 0.f;
      for (int s = tid >> 6; s < NS; s += 4) {
        const int q_new = s_sample[s], q_old = s_prev[s];
        float v = bias;
        if (kc1 > 0 && q_old >= 0) v += emb[((int64_t)0 * Q + q_old) * R + c];
        v += emb[((int64_t)kc1 * Q + q_new) * R + c];
        xv[s * 64 + c] = v;
      }
    }
    csync4();
    if (tid < NS) s_prev[tid] = s_sample[tid];
    // ---- 3. residual layers ----
    TRG(41, 1);
    float skr[NS];
#pragma unroll
    for (int s = 0; s < NS; ++s) skr[s] = 0.f;
    for (int l = 0; l < L.L; ++l, ++it) {
      const GenLayerOff& ly = s_layers[l];
      const bool has_ba = ly.has_ba != 0, has_bb = ly.has_bb != 0;
      const int64_t ba_off = ly.ba, bb_off = ly.bb;
      const uint32_t stage = it % V4_STAGES;
      // prefetch the taps of layer l + 3 (its ring slot is overwritten only when that layer runs)
      const bool pf = l + 3 < L.L;
      float4 pnext = make_float4(0.f, 0.f, 0.f, 0.f);
      if (pf) pnext = __ldcg(past_ptr(l + 3));
      // this CTA's slice of x[t] for the ring (8 channels x 16 streams), read before phase B updates xv
      const int rs = tid >> 3, rc = tid & 7;
      const float xring = tid < 8 * NS ? xv[rs * 64 + 8 * rank + rc] : 0.f;
      const int64_t ring_slot = ly.ring + ((int64_t)(stream0 + min(rs, ns - 1)) * ly.ring_len + s_pos[l]) * R;
      TRG(l, 0);
      tc::mbar_wait(full0 + 8 * stage, (it / V4_STAGES) & 1);
      TRG(l, 1);
      const float4* wst = reinterpret_cast<const float4*>(ring_g + stage * V4_STAGE_BYTES);
      const uint32_t zb = zsel ? zbar1 : zbar0;
      {
        const float4 w0 = wst[tid], w1 = wst[256 + tid];
        const float* xin = ksA < 8 ? xpast + ((l & (V5_XP - 1)) * NS) * 64 + ksA * 8 : xv + (ksA - 8) * 8;
        float v[NS];
#pragma unroll
        for (int s = 0; s < NS; ++s) {
          const float4 x0 = *reinterpret_cast<const float4*>(xin + s * 64), x1 = *reinterpret_cast<const float4*>(xin + s * 64 + 4);
          float acc = w0.x * x0.x, acc1 = w1.x * x1.x;
          acc = fmaf(w0.y, x0.y, acc), acc = fmaf(w0.z, x0.z, acc), acc = fmaf(w0.w, x0.w, acc);
          acc1 = fmaf(w1.y, x1.y, acc1), acc1 = fmaf(w1.z, x1.z, acc1), acc1 = fmaf(w1.w, x1.w, acc1);
          v[s] = acc + acc1;
        }
        // transposing butterfly over the 16 lanes of a half-warp: every step halves the live values; lane j ends with the
        // sum of stream j
#pragma unroll
        for (int w = 8; w >= 1; w >>= 1) {
          const bool upper = (lane & w) != 0;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            if (i < w) {
              const float send = upper ? v[i] : v[i + w];
              const float keep = upper ? v[i + w] : v[i];
              v[i] = keep + __shfl_xor_sync(0xffffffffu, send, w);
            }
          }
        }
        const int ch = 8 * rank + warp;
        float g = lane < 16 ? v[0] : 0.5f * v[0];
        if (has_ba) g += lane < 16 ? st[ba_off + ch] : 0.5f * st[ba_off + G + ch];
        const float th = tanhf(g);
        const float zval = th * (0.5f + 0.5f * __shfl_down_sync(0xffffffffu, th, 16));   // lanes 0..15: stream jA
        if (lane < 16) stg[(xsel * 32 + warp) * NS + jA] = zval;
      }
      TRG(l, 6);
      send_slice(zv_s + 4u * (uint32_t)((zsel * 64 + 8 * rank) * NS), 8, zb);
      TRG(l, 2);
      exchange_wait(zb, zph[zsel], (uint32_t)(64 * NS * 4), tid);
      TRG(l, 3);
      zph[zsel] ^= 1;
      if (tid < 192) {
        const float4* wb = wst + V4_WA_F / 4;
        const float* zk = zv + (zsel * 64 + ksB * 32) * NS;
        float acc[NS];
#pragma unroll
        for (int s = 0; s < NS; ++s) acc[s] = 0.f;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 w = wb[q * 192 + tid];
          const float wk[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
          for (int c = 0; c < 4; ++c) {
#pragma unroll
            for (int g4 = 0; g4 < NS / 4; ++g4) {
              const float4 z = *reinterpret_cast<const float4*>(zk + (q * 4 + c) * NS + g4 * 4);
              acc[4 * g4] = fmaf(wk[c], z.x, acc[4 * g4]), acc[4 * g4 + 1] = fmaf(wk[c], z.y, acc[4 * g4 + 1]);
              acc[4 * g4 + 2] = fmaf(wk[c], z.z, acc[4 * g4 + 2]), acc[4 * g4 + 3] = fmaf(wk[c], z.w, acc[4 * g4 + 3]);
            }
          }
        }
        const float bb = has_bb ? st[bb_off + (oB < 64 ? oB : R + 32 * rank + oB - 64)] : 0.f;
#pragma unroll
        for (int s = 0; s < NS; ++s) acc[s] += __shfl_xor_sync(0xffffffffu, acc[s], 1) + bb;
        if (oB < 64) {
          if (ksB == 0) {
#pragma unroll
            for (int s = 0; s < NS; ++s) xv[s * 64 + oB] += acc[s];      // output = projection + x, wavenet.py:354
          }
        } else {
#pragma unroll
          for (int s = 0; s < NS; ++s) skr[s] += acc[s];                 // faster_wavenet.py:100 (both lanes of the pair hold it)
        }
      }
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(empty0 + 8 * stage);
      // ring roll (faster_wavenet.py:90-91): the slot is the one read as x[t-d] >= 3 layers ago by every CTA
      if (tid < 8 * NS && rs < ns) st[ring_slot + 8 * rank + rc] = xring;
      if (pf) *reinterpret_cast<float4*>(xpast + (((l + 3) & (V5_XP - 1)) * NS + ps) * 64 + pv * 4) = pnext;
      TRG(l, 4);
      csync4();
      TRG(l, 5);
      zsel ^= 1;
    }
    // ---- 4. head ----
    TRG(40, 0);
    int hsel = 0;
    if (tid >= 128 && tid < 192) {           // skip sums: lane ksB of a pair stages streams 8 ksB .. 8 ksB + 7
      float* row = stg + (xsel * 32 + oB - 64) * NS + ksB * 8;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float v0 = ksB ? skr[8 + i] : skr[i];
        row[i] = head_act(v0, a.head_elu);
      }
    }
    send_slice(hB_s + 4u * (uint32_t)(32 * rank * NS), 32, hbar0);
    exchange_wait(hbar0, hph[0], (uint32_t)(256 * NS * 4), tid);
    hph[0] ^= 1;
    hsel = 1;
    float* hin = hB;
    float* hout = hA;
    uint32_t hout_s = hA_s, hin_s = hB_s;
    for (int hi = 0; hi < L.n_head; ++hi, ++it) {
      const uint32_t stage = it % V4_STAGES;
      tc::mbar_wait(full0 + 8 * stage, (it / V4_STAGES) & 1);
      const float4* wst = reinterpret_cast<const float4*>(ring_g + stage * V4_STAGE_BYTES);
      const bool last = hi == L.n_head - 1;
      const float hb = hbias[hi * 256 + 32 * rank + oH];
      const uint32_t hbar = hsel ? hbar1 : hbar0;
      const float* xk = hin + (ksH * 32) * NS;
      float v[NS];
#pragma unroll
      for (int s = 0; s < NS; ++s) v[s] = 0.f;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float4 w = wst[q * 256 + tid];
        const float wk[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
        for (int c = 0; c < 4; ++c) {
#pragma unroll
          for (int g4 = 0; g4 < NS / 4; ++g4) {
            const float4 x = *reinterpret_cast<const float4*>(xk + (q * 4 + c) * NS + g4 * 4);
            v[4 * g4] = fmaf(wk[c], x.x, v[4 * g4]), v[4 * g4 + 1] = fmaf(wk[c], x.y, v[4 * g4 + 1]);
            v[4 * g4 + 2] = fmaf(wk[c], x.z, v[4 * g4 + 2]), v[4 * g4 + 3] = fmaf(wk[c], x.w, v[4 * g4 + 3]);
          }
        }
      }
      // transposing butterfly over the 8 K-slice lanes: 16 -> 8 -> 4 -> 2 live values; lane ks ends with streams 2ks, 2ks+1
#pragma unroll
      for (int w = 4, half = 8; w >= 1; w >>= 1, half >>= 1) {
        const bool upper = (lane & w) != 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          if (i < half) {
            const float send = upper ? v[i] : v[i + half];
            const float keep = upper ? v[i + half] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, w);
          }
        }
      }
      float o0 = v[0] + hb, o1 = v[1] + hb;
      if (!last) o0 = head_act(o0, a.head_elu), o1 = head_act(o1, a.head_elu);
      *reinterpret_cast<float2*>(stg + (xsel * 32 + oH) * NS + 2 * ksH) = make_float2(o0, o1);
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(empty0 + 8 * stage);
      send_slice(hout_s + 4u * (uint32_t)(32 * rank * NS), 32, hbar);
      exchange_wait(hbar, hph[hsel], (uint32_t)(256 * NS * 4), tid);
      hph[hsel] ^= 1;
      hsel ^= 1;
      float* tmp = hin;
      hin = hout;
      hout = tmp;
      const uint32_t ts = hin_s;
      hin_s = hout_s;
      hout_s = ts;
    }
    lg = hin;
    if (tid < L.L) {
      const int p = s_pos[tid] + 1;
      s_pos[tid] = p == s_layers[tid].ring_len ? 0 : p;
    }
    csync4();                                // s_pos settled before the next step's tap addresses
    TRG(40, 1);
  }
  // ---- epilogue (rank 0 publishes the streams' state) ----
  if (rank == 0) {
    for (int i = tid; i < ns * Q; i += V4_T) {
      const int q = i % Q, s = i / Q;
      cur_logits[(int64_t)(stream0 + s) * Q + q] = lg[q * NS + s];
    }
    if (kc1 > 0 && tid < ns) idx_hist[(int64_t)(stream0 + tid) * kc1 + kc1 - 1] = s_prev[tid];
    if (a.probs) {
      for (int s = 0; s < ns; ++s) {
        float* pr = a.probs + (int64_t)(stream0 + s) * Q;
        const float mine = lg[tid * NS + s];
        if (!a.apply_softmax) {
          pr[tid] = mine;
        } else {
          float m = mine;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
          if (lane == 0) s_redv[0][warp] = m;
          csync4();
          m = s_redv[0][0];
          for (int w = 1; w < V4_T / 32; ++w) m = fmaxf(m, s_redv[0][w]);
          csync4();
          float sum = expf(mine - m);
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
          if (lane == 0) s_redv[0][warp] = sum;
          csync4();
          sum = 0.f;
          for (int w = 0; w < V4_T / 32; ++w) sum += s_redv[0][w];
          pr[tid] = expf(mine - m) / sum;
          csync4();
        }
      }
    }
  }
  cluster_sync_all();
}

size_t gen_smem_bytes_v5(const GenLayout& L) {
  const size_t f = (size_t)V5_NS * (64 + 128 + 256 + 256 + V5_XP * 64 + 64) + (size_t)L.n_head * 256 + 64;
  return f * sizeof(float) + V4_STAGES * V4_STAGE_BYTES + 128;
}

// dst[rank][...] <- src [K][N] (generator layout), cut into the per-rank, per-thread order gen_kernel_v4 reads:
// mode 1: WA (K = 128, N = 128 = a_f | a_g), mode 2: WB (K = 64, N = 320 = residual (all ranks) | skip), mode 3: head.
__global__ void gen_pack_v4(const float* __restrict__ src, float* __restrict__ dst, int64_t rank_stride, int K, int N,
                            int mode) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= K * N) return;
  const int k = idx / N, n = idx % N, comp = k & 3;
  int r, f4;
  if (mode == 1) {
    const int ch = n & 63, half = n >> 6;
    r = ch >> 3;
    const int t = ((ch & 7) * 2 + half) * 16 + (k >> 3);
    f4 = ((k & 7) >> 2) * 256 + t;
  } else if (mode == 2) {
    // residual rows (n < 64) go to EVERY rank (each CTA updates the whole x), skip rows to their owner
    const int o = n < 64 ? n : 64 + ((n - 64) & 31);
    const int t = o * 2 + (k >> 5);
    f4 = ((k & 31) >> 2) * 192 + t;
    if (n < 64) {
      for (int rr = 0; rr < 8; ++rr) dst[(int64_t)rr * rank_stride + (int64_t)f4 * 4 + comp] = src[idx];
      return;
    }
    r = (n - 64) >> 5;
  } else {
    r = n >> 5;
    const int t = (n & 31) * 8 + (k >> 5);
    f4 = ((k & 31) >> 2) * 256 + t;
  }
  dst[(int64_t)r * rank_stride + (int64_t)f4 * 4 + comp] = src[idx];
}

size_t gen_smem_bytes_v4(const GenLayout& L) {
  const size_t f = 64 + 128 + 256 + 256 + (size_t)L.L * 64 + (size_t)L.n_head * 256 + 64;
  return f * sizeof(float) + V4_STAGES * V4_STAGE_BYTES + 128;
}

// how many 8-CTA clusters of gen_kernel_v4 can be resident at once (they must all be: the kernel is persistent over
// every audio sample, a second wave would only start when the first has finished)
int gen_v4_max_streams(const GenLayout& L) {
  // cached per shared-memory size: the footprint depends on the number of layers / head convs of THIS network (a process
  // may hold several networks, e.g. the test suite)
  static int cached = -1;
  static size_t cached_smem = 0;
  const size_t smem = gen_smem_bytes_v4(L);
  if (cached >= 0 && cached_smem == smem) return cached;
  cached_smem = smem;
  if (cudaFuncSetAttribute(gen_kernel_v4, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return cached = 0;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(V4_CS * 32);
  cfg.blockDim = dim3(V4_T + 32);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = V4_CS;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, gen_kernel_v4, &cfg) != cudaSuccess) {
    cudaGetLastError();
    n = 0;
  }
  return cached = n;
}

int launch_gen_v4(const GenArgs& a, cudaStream_t s) {
  const size_t smem = gen_smem_bytes_v4(a.lay);
  static size_t attr_smem = 0;       // the opt-in limit must cover the largest network launched so far
  if (smem > attr_smem) {
    WN_CHECK_CUDA(cudaFuncSetAttribute(gen_kernel_v4, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_smem = smem;
  }
  gen_kernel_v4<<<a.lay.n * V4_CS, V4_T + 32, smem, s>>>(a);
  WN_CHECK_LAUNCH();
  return WN_OK;
}

// co-resident 8-CTA clusters of gen_kernel_v5 (16 streams each)
int gen_v5_max_clusters(const GenLayout& L) {
  static int cached = -1;
  static size_t cached_smem = 0;
  const size_t smem = gen_smem_bytes_v5(L);
  if (cached >= 0 && cached_smem == smem) return cached;
  cached_smem = smem;
  if (smem > 227 * 1024 ||
      cudaFuncSetAttribute(gen_kernel_v5, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
    cudaGetLastError();
    return cached = 0;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(V4_CS * 32);
  cfg.blockDim = dim3(V4_T + 32);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = V4_CS;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, gen_kernel_v5, &cfg) != cudaSuccess) {
    cudaGetLastError();
    n = 0;
  }
  return cached = n;
}

int launch_gen_v5(const GenArgs& a, cudaStream_t s) {
  const size_t smem = gen_smem_bytes_v5(a.lay);
  static size_t attr_smem = 0;
  if (smem > attr_smem) {
    WN_CHECK_CUDA(cudaFuncSetAttribute(gen_kernel_v5, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_smem = smem;
  }
  const int clusters = (a.lay.n + V5_NS - 1) / V5_NS;
  gen_kernel_v5<<<clusters * V4_CS, V4_T + 32, smem, s>>>(a);
  WN_CHECK_LAUNCH();
  return WN_OK;
}

#include "wn_gen_mma.cuh"

size_t gen_smem_bytes(const GenLayout& L, int NS, bool stream) {
  const size_t maxw = L.maxw;
  size_t f = NS * maxw                               // xv
             + NS * (L.k + L.kc) * maxw              // xin
             + NS * 2 * maxw                         // av
             + NS * maxw * 3                         // zv, skipv, hv
             + (size_t)4 * GT * NS + 64;             // part
  if (stream) f += (size_t)NS * L.L * (L.k - 1) * L.R;   // xpast
  return f * sizeof(float) + (stream ? RING_STAGES * RING_STAGE_BYTES + 128 : 0);
}

template <int NS, bool STREAM>
int launch_gen(const GenArgs& a, cudaStream_t s) {
  const size_t smem = gen_smem_bytes(a.lay, NS, STREAM);
  WN_REQUIRE(smem <= 227 * 1024, WN_EINVAL, "generator: network too wide for shared memory (%zu bytes)", smem);
  WN_CHECK_CUDA(cudaFuncSetAttribute(gen_kernel<NS, STREAM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = (a.lay.n + NS - 1) / NS;
  gen_kernel<NS, STREAM><<<grid, GT + (STREAM ? 32 : 0), smem, s>>>(a);
  WN_CHECK_LAUNCH();
  return WN_OK;
}

// fp32 rings <-> operand-tile rings of gen_kernel_v6 (lazily, whenever the kernel family changes between calls)
int gen_v6_convert(wn_gen* g, bool to_v6, cudaStream_t s) {
  const GenLayout& L = g->lay;
  uint8_t* ring6 = reinterpret_cast<uint8_t*>(g->state + L.ring6);
  int base = 0;
  for (int l = 0; l < L.L; ++l) {
    const GenLayerOff& o = g->layers[l];
    const int64_t items = (int64_t)g->v6_clusters * o.ring_len * 1024;
    const unsigned nb = (unsigned)((items + 255) / 256);
    if (to_v6)
      gen_ring_to_v6<<<nb, 256, 0, s>>>(g->state + o.ring, ring6, L.ring6_cta_bytes, base, o.ring_len, L.n, g->v6_spc, g->v6_clusters,
                                        g->v6_cs);
    else
      gen_ring_from_v6<<<nb, 256, 0, s>>>(g->state + o.ring, ring6, L.ring6_cta_bytes, base, o.ring_len, L.n, g->v6_spc,
                                          g->v6_clusters, g->v6_cs);
    base += o.ring_len;
  }
  WN_CHECK_LAUNCH();
  return WN_OK;
}

// co-resident clusters of gen_kernel_v6<CS> (the kernel is persistent over every audio sample: a second wave of clusters
// would only start when the first has finished)
template <int CS>
int gen_v6_max_clusters_cs() {
  static int cached = -1;
  if (cached >= 0) return cached;
  const int smem = V6C<CS>::SMEM + 128;
  if (cudaFuncSetAttribute(gen_kernel_v6<CS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) {
    cudaGetLastError();
    return cached = 0;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(CS * 64);
  cfg.blockDim = dim3(V6_THREADS);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CS;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, gen_kernel_v6<CS>, &cfg) != cudaSuccess) {
    cudaGetLastError();
    n = 0;
  }
  return cached = n;
}
int gen_v6_max_clusters(int cs) { return cs == 4 ? gen_v6_max_clusters_cs<4>() : gen_v6_max_clusters_cs<8>(); }

template <int CS>
int launch_gen_v6_cs(wn_gen* g, const GenArgs& a, const V6Args& v, int n_clusters, cudaStream_t s) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(n_clusters * CS);
  cfg.blockDim = dim3(V6_THREADS);
  cfg.dynamicSmemBytes = V6C<CS>::SMEM + 128;
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CS;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  WN_CHECK_CUDA(cudaLaunchKernelEx(&cfg, gen_kernel_v6<CS>, a, v));
  return WN_OK;
}

int launch_gen_v6(wn_gen* g, const GenArgs& a, cudaStream_t s) {
  if (!g->ring6_valid) WN_TRY(gen_v6_convert(g, true, s));
  V6Args v;
  v.wpk = reinterpret_cast<const uint8_t*>(g->state + g->lay.wpk6);
  v.wpk_rank_bytes = g->lay.wpk6_rank_bytes;
  v.ring = reinterpret_cast<uint8_t*>(g->state + g->lay.ring6);
  v.ring_cta_bytes = g->lay.ring6_cta_bytes;
  v.spc = g->v6_spc;
  // every cluster of a launch must be resident (the kernel is persistent over all audio samples): more clusters than fit
  // run as consecutive launches over the same samples
  int wave = gen_v6_max_clusters(g->v6_cs);              // also sets the shared-memory attribute
  if (const char* e = getenv("WN_GEN_V6_WAVE"))           // developer switch: smaller waves (tests)
    if (atoi(e) > 0 && atoi(e) < wave) wave = atoi(e);
  WN_REQUIRE(wave > 0, WN_ESTATE, "gen_kernel_v6 cannot be resident on this device");
  for (int c0 = 0; c0 < g->v6_clusters; c0 += wave) {
    const int nc = g->v6_clusters - c0 < wave ? g->v6_clusters - c0 : wave;
    v.cluster0 = c0;
    if (g->v6_cs == 4)
      WN_TRY(launch_gen_v6_cs<4>(g, a, v, nc, s));
    else
      WN_TRY(launch_gen_v6_cs<8>(g, a, v, nc, s));
    WN_CHECK_LAUNCH();
  }
  g->ring6_valid = true;
  g->ringf_valid = false;
  return WN_OK;
}

int pick_ns(const wn_gen* g) {
  // streams per CTA: spread over the SMs first, then stack streams to amortise weight reads
  const int n = g->lay.n, sms = g->h->sm_count;
  if (n <= sms) return 1;
  if (n <= 2 * sms) return 2;
  return 4;
}

int run_gen(wn_gen* g, GenArgs& a, cudaStream_t s) {
  const int ns = pick_ns(g);
  // tensor-core generator: >= 16 streams of the config-C shape, free-running sampling loop (not the one-step API)
  if (g->v6_ok && a.sample_first && a.n_steps >= 8 && !a.probs) {
    const char* e = getenv("WN_GEN_V6");
    if (!e || atoi(e) != 0) return launch_gen_v6(g, a, s);
  }
  if (!g->ringf_valid) {                     // the last call ran on the v6 rings: bring the fp32 rings up to date
    WN_TRY(gen_v6_convert(g, false, s));
    g->ringf_valid = true;
  }
  g->ring6_valid = false;
  if (g->v4_ok && g->lay.n <= gen_v4_max_streams(g->lay)) {
    const char* e = getenv("WN_GEN_V4");
    if (!e || atoi(e) != 0) return launch_gen_v4(a, s);
  }
  // gen_kernel_v5 (16 streams per cluster) is OPT-IN: it is correct (tests) but measured 164 us per step against 59 us for
  // one stream per CTA -- with FP32 SIMT math every FMA needs a shared-memory operand and 16 streams per cluster put ~4 k
  // shared-memory wavefronts on each layer; kept as the measured record of that design.
  if (g->v4_ok && g->lay.n >= 2 && (g->lay.n + V5_NS - 1) / V5_NS <= gen_v5_max_clusters(g->lay)) {
    const char* e = getenv("WN_GEN_V5");
    if (e && atoi(e) != 0) return launch_gen_v5(a, s);
  }
  if (g->v3_ok) {
    int ns3 = ns;
    if (const char* e = getenv("WN_GEN_NS")) ns3 = atoi(e);
    if (ns3 != 1 && ns3 != 2 && ns3 != 4) ns3 = 2;
    while (ns3 > 1 && gen_smem_bytes_v3(g->lay, ns3) > 208 * 1024) ns3 /= 2;
    a.chunks = (const GenChunk*)(g->state + g->lay.chunks3_dev);
    a.n_chunks = (int)g->chunks3.size();
    if (ns3 == 1) return launch_gen_v3<1>(a, s);
    if (ns3 == 2) return launch_gen_v3<2>(a, s);
    return launch_gen_v3<4>(a, s);
  }
  if (g->stream_ok && gen_smem_bytes(g->lay, ns, true) <= 208 * 1024) {   // + ~16 KB static (schedule, layer table)
    if (ns == 1) return launch_gen<1, true>(a, s);
    if (ns == 2) return launch_gen<2, true>(a, s);
    return launch_gen<4, true>(a, s);
  }
  if (ns == 1) return launch_gen<1, false>(a, s);
  if (ns == 2) return launch_gen<2, false>(a, s);
  return launch_gen<4, false>(a, s);
}

}  // namespace

extern "C" int wn_gen_create(wn_handle* h, int n_streams, int head_act, wn_gen** out) {
  WN_REQUIRE(h && out, WN_EINVAL, "null argument");
  WN_REQUIRE(n_streams >= 1, WN_EINVAL, "n_streams must be >= 1");
  WN_REQUIRE(head_act == 0 || head_act == 1, WN_EINVAL, "head_act must be 0 (relu) or 1 (reference)");
  wn_gen* g = new wn_gen();
  g->h = h;
  g->head_act = head_act;
  GenLayout& L = g->lay;
  const wn_config& c = h->cfg;
  L.n = n_streams;
  L.Q = h->Q;
  L.R = h->R;
  L.S = h->S;
  L.k = c.residual_filter_width;
  L.kc = c.causal_filter_width;
  L.n_causal = c.n_causal;
  L.n_head = (int)h->head.size();
  L.L = (int)h->layers.size();
  L.has_hb = !c.softmax_no_bias;
  L.has_cb = !c.causal_no_bias;
  int maxw = h->Q > h->S ? h->Q : h->S;
  for (int i = 0; i < c.n_causal; ++i) {
    L.causal_ch[i] = c.causal_channels[i];
    maxw = maxw > c.causal_channels[i] ? maxw : c.causal_channels[i];
  }
  for (int i = 0; i < c.n_softmax; ++i) {
    L.head_ch[i] = c.softmax_channels[i];
    maxw = maxw > c.softmax_channels[i] ? maxw : c.softmax_channels[i];
  }
  int64_t off = 0;
  auto take = [&](int64_t nfl) {
    int64_t o = off;
    off = align_up64(off + nfl);
    return o;
  };
  L.emb = take((int64_t)L.kc * L.Q * L.causal_ch[0]);
  L.emb_b = take(L.causal_ch[0]);
  for (int i = 1; i < c.n_causal; ++i) {
    L.cw[i] = take((int64_t)L.kc * L.causal_ch[i - 1] * L.causal_ch[i]);
    L.cb[i] = take(L.causal_ch[i]);
    L.chist[i] = take((int64_t)n_streams * (L.kc - 1) * L.causal_ch[i - 1] + 1);
  }
  for (int l = 0; l < L.L; ++l) {
    const ResLayer& ly = h->layers[l];
    GenLayerOff o;
    o.G = ly.G;
    o.dilation = ly.dilation;
    o.ring_len = (L.k - 1) * ly.dilation;
    o.has_ba = ly.wf.b_off >= 0;
    o.has_bb = ly.proj.b_off >= 0;
    o.pad = 0;
    o.wa = take((int64_t)L.k * L.R * 2 * ly.G);
    o.ba = take(2 * ly.G);
    o.wb = take((int64_t)ly.G * (L.R + L.S));
    o.bb = take(L.R + L.S);
    o.ring = take((int64_t)n_streams * o.ring_len * L.R + 1);
    g->layers.push_back(o);
    maxw = maxw > 2 * ly.G ? maxw : 2 * ly.G;
    maxw = maxw > L.R + L.S ? maxw : L.R + L.S;
  }
  for (int i = 0; i < L.n_head; ++i) {
    L.hw[i] = take((int64_t)L.head_ch[i] * L.head_ch[i + 1]);
    L.hb[i] = take(L.head_ch[i + 1]);
  }
  L.idx_hist = take((int64_t)n_streams * (L.kc - 1) + 1);
  L.cur_logits = take((int64_t)n_streams * L.Q);
  L.layers_dev = take((int64_t)(sizeof(GenLayerOff) * L.L + 3) / 4);
  // streaming schedule: every matrix of one step, in consumption order, cut into row chunks of <= 16 KB
  g->stream_ok = true;
  auto add_matrix = [&](int64_t off_floats, int K, int N) {
    if (N % 4 != 0 || K % 4 != 0 || N > 2048) {
      g->stream_ok = false;
      return;
    }
    const int rpc = 1 * 4 * (GT_HOST / (N / 4));   // RING_GROUPS groups of 4*nsl rows
    for (int r0 = 0; r0 < K; r0 += rpc) {
      const int rows = K - r0 < rpc ? K - r0 : rpc;
      GenChunk c;
      c.off = (uint64_t)(off_floats + (int64_t)r0 * N) * 4;
      c.bytes = (uint32_t)(rows * N * 4);
      c.pad = 0;
      g->chunks.push_back(c);
    }
  };
  for (int i = 1; i < c.n_causal; ++i) add_matrix(L.cw[i], L.kc * L.causal_ch[i - 1], L.causal_ch[i]);
  for (int l = 0; l < L.L; ++l) {
    add_matrix(g->layers[l].wa, L.k * L.R, 2 * g->layers[l].G);
    add_matrix(g->layers[l].wb, g->layers[l].G, L.R + L.S);
  }
  for (int i = 0; i < L.n_head; ++i) add_matrix(L.hw[i], L.head_ch[i], L.head_ch[i + 1]);
  if (g->chunks.size() > 512 || L.L > 128 || L.R % 4 != 0) g->stream_ok = false;   // MAX_CHUNKS, s_layers, float4 taps
  L.chunks_dev = take((int64_t)(sizeof(GenChunk) * g->chunks.size() + 3) / 4 + 4);
  // v3 (packed weights, warp-shuffle reductions): needs 2G = 128, R + S = 320, head widths 256, one causal layer
  g->v3_ok = c.n_causal == 1 && L.R + L.S == 320 && L.S == 256 && L.R % 4 == 0 && L.k * L.R == 128 && ((L.k - 1) * L.R) % 32 == 0 &&
             L.L <= 128 && L.Q == 256;
  for (int l = 0; l < L.L && g->v3_ok; ++l) g->v3_ok = g->layers[l].G == 64;
  for (int i = 0; i < L.n_head && g->v3_ok; ++i) g->v3_ok = L.head_ch[i + 1] == 256 && L.head_ch[i] % 64 == 0 && L.head_ch[i] <= 256;
  if (g->v3_ok) {
    int64_t pk = 0;
    auto add3 = [&](int64_t src, int K, int N, int chunkK, int mode) {
      wn_gen::Pack3 p{src, pk, K, N, chunkK, mode};
      g->packs3.push_back(p);
      for (int r0 = 0; r0 < K; r0 += chunkK) {
        GenChunk ch;
        ch.off = (uint64_t)(pk + (int64_t)r0 * N) * 4;   // relative to wpk, fixed up below
        ch.bytes = (uint32_t)(chunkK * N * 4);
        ch.pad = 0;
        g->chunks3.push_back(ch);
      }
      pk += (int64_t)K * N;
    };
    for (int l = 0; l < L.L; ++l) {
      add3(g->layers[l].wa, L.k * L.R, 128, L.k * L.R, 1);     // whole matrix = one chunk (64 KB)
      add3(g->layers[l].wb, g->layers[l].G, 320, g->layers[l].G, 2);   // 80 KB
    }
    for (int i = 0; i < L.n_head; ++i) add3(L.hw[i], L.head_ch[i], 256, 64, 3);   // 64 KB chunks
    L.wpk = take(pk);
    for (auto& ch : g->chunks3) ch.off += (uint64_t)L.wpk * 4;
    for (auto& p : g->packs3) p.dst += L.wpk;
    if (g->chunks3.size() > 512) g->v3_ok = false;
    L.chunks3_dev = take((int64_t)(sizeof(GenChunk) * g->chunks3.size() + 3) / 4 + 4);
  }
  // v4 (one 8-CTA cluster per stream): the v3 shape with 64 residual channels and 256-wide head inputs, few streams
  g->v4_ok = g->v3_ok && L.R == 64 && L.k == 2 && L.kc <= 2;
  for (int i = 0; i < L.n_head && g->v4_ok; ++i) g->v4_ok = L.head_ch[i] == 256;
  for (int l = 0; l < L.L && g->v4_ok; ++l) g->v4_ok = g->layers[l].ring_len > 0;
  if (g->v4_ok) {
    L.wpk4_rank = (int64_t)L.L * (2 * 256 * 4 + 8 * 192 * 4) + (int64_t)L.n_head * (8 * 256 * 4);
    L.wpk4 = take(L.wpk4_rank * V4_CS_HOST);
  }
  // v6 (streams = MMA M dimension, <= 128 per 8-CTA cluster): the v4 shape with biases in the head only, two head convs
  // It pays as soon as gen_kernel_v3 needs four streams per CTA, i.e. beyond 2 x sm_count streams (measured: 107-116 us per
  // step at any stream count against 59 / 82 / 132 us for 1 / 2 / 4 streams per CTA of gen_kernel_v3); WN_GEN_V6=1 forces it
  // from 16 streams (tests), WN_GEN_V6=0 disables it.
  const char* e6 = getenv("WN_GEN_V6");
  const int min6 = e6 && atoi(e6) != 0 ? 16 : 2 * h->sm_count + 1;
  g->v6_ok = g->v4_ok && L.n_head == 2 && !L.has_cb && L.L >= 16 && L.L <= 64 && n_streams >= min6 && !(e6 && atoi(e6) == 0);
  for (int l = 0; l < L.L && g->v6_ok; ++l) g->v6_ok = !g->layers[l].has_ba && !g->layers[l].has_bb;
  if (g->v6_ok) {
    int spc = 128;
    if (const char* e = getenv("WN_GEN_V6_SPC")) spc = atoi(e);
    if (spc < 8 || spc > 128) spc = 128;
    g->v6_spc = spc;
    g->v6_clusters = (n_streams + spc - 1) / spc;
    // 8-CTA clusters while they are all co-resident (15 on a B200), else 4-CTA clusters (twice the streams per SM, a slightly
    // longer step); WN_GEN_V6_CS = 4 / 8 pins the choice
    int cs = g->v6_clusters <= gen_v6_max_clusters(8) ? 8 : 4;
    if (const char* e = getenv("WN_GEN_V6_CS")) cs = atoi(e) == 4 ? 4 : 8;
    g->v6_cs = cs;
    if (gen_v6_max_clusters(cs) <= 0) g->v6_ok = false;
    // every CTA keeps its own operand-tile copy of the dilation rings (~100 MB for config C): stay below 48 GB of state
    int64_t slots6 = 0;
    for (int l = 0; l < L.L; ++l) slots6 += g->layers[l].ring_len;
    if (slots6 * 32768 * cs * g->v6_clusters > ((int64_t)48 << 30)) g->v6_ok = false;
  }
  if (g->v6_ok) {
    const int cs = g->v6_cs;
    int64_t slots = 0;
    for (int l = 0; l < L.L; ++l) slots += g->layers[l].ring_len;
    const int64_t chunk = cs == 4 ? V6C<4>::CHUNK : V6C<8>::CHUNK;
    L.wpk6_rank_bytes = (int64_t)L.L * chunk + (int64_t)L.n_head * (256 / cs / 32) * 32768;
    L.wpk6 = take(L.wpk6_rank_bytes * cs / 4);
    L.ring6_cta_bytes = slots * 32768;
    L.ring6 = take(L.ring6_cta_bytes / 4 * cs * g->v6_clusters);
  }
  L.maxw = (maxw + 3) / 4 * 4;
  L.total = off;
  *out = g;
  return WN_OK;
}

#ifdef WN_LAYER_TRACE
extern "C" int wn_debug_gen_trace(long long* out) {
  return cudaMemcpyFromSymbol(out, g_trace_gen, sizeof(long long) * 64 * 16) == cudaSuccess ? 0 : -1;
}
#endif
extern "C" int wn_gen_mma_capacity(int cluster_size) {
  if (cluster_size != 4 && cluster_size != 8) return WN_EINVAL;
  return gen_v6_max_clusters(cluster_size) * 128;
}

extern "C" int wn_gen_destroy(wn_gen* g) {
  delete g;
  return WN_OK;
}

extern "C" int64_t wn_gen_state_bytes(const wn_gen* g) { return g ? g->lay.total * (int64_t)sizeof(float) : WN_EINVAL; }

extern "C" int wn_gen_bind_state(wn_gen* g, void* state, int64_t bytes) {
  WN_REQUIRE(g && state, WN_EINVAL, "null argument");
  WN_REQUIRE(bytes >= g->lay.total * (int64_t)sizeof(float), WN_ENOMEM, "generator state too small");
  WN_REQUIRE(((uintptr_t)state & 255) == 0, WN_EINVAL, "generator state must be 256-byte aligned");
  g->state = (float*)state;
  g->state_bytes = bytes;
  g->primed = false;
  return WN_OK;
}

extern "C" int wn_gen_prime(wn_gen* g, const float* params, const int32_t* window, float* probs_opt, wn_stream_t st) {
  WN_REQUIRE(g, WN_EINVAL, "null argument");
  return wn_gen_prime_part(g, params, window, 0, g->lay.n, probs_opt, st);
}

// Priming of the streams [stream0, stream0 + count): the full pass needs a training workspace of `count` sequences only, so
// thousands of streams can be primed in slices (the tape of one stream of config C is ~130 MB).  The generator is primed
// when the slice that ends at n_streams has run; slices may come in any order before that one.
extern "C" int wn_gen_prime_part(wn_gen* g, const float* params, const int32_t* window, int stream0, int count, float* probs_opt,
                                 wn_stream_t st) {
  WN_REQUIRE(g && params && window, WN_EINVAL, "null argument");
  WN_REQUIRE(g->state, WN_ESTATE, "no generator state bound");
  wn_handle* h = g->h;
  const GenLayout& L = g->lay;
  const int Win = wn_input_width(h);
  WN_REQUIRE(stream0 >= 0 && count >= 1 && stream0 + count <= L.n, WN_EINVAL, "stream slice [%d, %d) outside 0..%d", stream0,
             stream0 + count, L.n);
  WN_REQUIRE(h->ws && h->tape.B == count && h->tape.W == Win, WN_ESTATE,
             "wn_gen_prime: bind a training workspace for (B=%d, W=%d) first", count, Win);
  cudaStream_t s = (cudaStream_t)st;
  const wn_config& c6 = h->cfg;
  // full pass over the window (faster_wavenet.py:13-47); the priming head is ReLU (Q2).
  // The SIMT generators are exact fp32, and so is their priming pass (greedy sequences match the reference arithmetic).
  // The tensor-core generator computes in split fp16 (fp32-grade) anyway: its priming pass runs on the fp16x2 training
  // kernels -- 7x faster, which matters when thousands of streams are primed (WN_GEN_TC_PRIME=0: fp32 pass).
  bool tc_prime = g->v6_ok && c6.n_causal == 1 && tcs_supported(h);
  if (const char* e = getenv("WN_GEN_TC_PRIME")) tc_prime = tc_prime && atoi(e) != 0;
  const int saved_prec = h->prec;
  h->prec = tc_prime ? WN_PREC_F16X2 : WN_PREC_FP32;
  int rc = wn_forward_causal_block(h, params, window, nullptr, st);
  if (rc == WN_OK) rc = wn_forward_residual_block(h, params, nullptr, nullptr, nullptr, st);
  if (rc == WN_OK) rc = wn_forward_softmax_block(h, params, nullptr, 1, 1, probs_opt, st);
  h->prec = saved_prec;
  WN_TRY(rc);
  float* S = g->state;
  const Tape& t = h->tape;
  WN_CHECK_CUDA(cudaMemcpyAsync(S + L.cur_logits + (int64_t)stream0 * L.Q, h->ws + t.hbuf.back(), sizeof(float) * count * L.Q,
                                cudaMemcpyDeviceToDevice, s));
  auto nb = [](int64_t n) { return (unsigned)((n + 255) / 256); };
  // generator-layout weights
  const wn_config& c = h->cfg;
  WN_CHECK_CUDA(cudaMemcpyAsync(S + L.emb, h->ws + t.emb, sizeof(float) * L.kc * L.Q * L.causal_ch[0],
                                cudaMemcpyDeviceToDevice, s));
  gen_copy_bias<<<nb(L.causal_ch[0]), 256, 0, s>>>(h->causal[0].b_off >= 0 ? params + h->causal[0].b_off : nullptr,
                                                   S + L.emb_b, L.causal_ch[0]);
  for (int i = 1; i < c.n_causal; ++i) {
    const ConvParam& cp = h->causal[i];
    gen_transpose_conv<<<nb((int64_t)cp.out_ch * cp.in_ch * cp.taps), 256, 0, s>>>(params + cp.w_off, S + L.cw[i],
                                                                                    cp.out_ch, cp.in_ch, cp.taps,
                                                                                    cp.out_ch, 0);
    gen_copy_bias<<<nb(cp.out_ch), 256, 0, s>>>(cp.b_off >= 0 ? params + cp.b_off : nullptr, S + L.cb[i], cp.out_ch);
    if (L.kc > 1)
      gen_fill_ring<<<nb((int64_t)count * (L.kc - 1) * cp.in_ch), 256, 0, s>>>(
          h->ws + t.cx[i - 1], S + L.chist[i] + (int64_t)stream0 * (L.kc - 1) * cp.in_ch, count, Win, cp.in_ch, L.kc - 1, 1);
  }
  for (int l = 0; l < L.L; ++l) {
    const ResLayer& ly = h->layers[l];
    const GenLayerOff& o = g->layers[l];
    const int64_t nw = (int64_t)ly.G * L.R * L.k;
    gen_transpose_conv<<<nb(nw), 256, 0, s>>>(params + ly.wf.w_off, S + o.wa, ly.G, L.R, L.k, 2 * ly.G, 0);
    gen_transpose_conv<<<nb(nw), 256, 0, s>>>(params + ly.wg.w_off, S + o.wa, ly.G, L.R, L.k, 2 * ly.G, ly.G);
    gen_copy_bias<<<nb(ly.G), 256, 0, s>>>(ly.wf.b_off >= 0 ? params + ly.wf.b_off : nullptr, S + o.ba, ly.G);
    gen_copy_bias<<<nb(ly.G), 256, 0, s>>>(ly.wg.b_off >= 0 ? params + ly.wg.b_off : nullptr, S + o.ba + ly.G, ly.G);
    gen_transpose_conv<<<nb((int64_t)L.R * ly.G), 256, 0, s>>>(params + ly.proj.w_off, S + o.wb, L.R, ly.G, 1,
                                                                L.R + L.S, 0);
    gen_transpose_conv<<<nb((int64_t)L.S * ly.G), 256, 0, s>>>(params + ly.skip.w_off, S + o.wb, L.S, ly.G, 1,
                                                                L.R + L.S, L.R);
    gen_copy_bias<<<nb(L.R), 256, 0, s>>>(ly.proj.b_off >= 0 ? params + ly.proj.b_off : nullptr, S + o.bb, L.R);
    gen_copy_bias<<<nb(L.S), 256, 0, s>>>(ly.skip.b_off >= 0 ? params + ly.skip.b_off : nullptr, S + o.bb + L.R, L.S);
    if (o.ring_len > 0 && tc_prime)
      gen_fill_ring_split<<<nb((int64_t)count * o.ring_len * L.R), 256, 0, s>>>(
          reinterpret_cast<const __half*>(h->ws + t.x[l]), S + o.ring + (int64_t)stream0 * o.ring_len * L.R, count, Win, L.R,
          o.ring_len, 1.f / tcs_act_scale());
    else if (o.ring_len > 0)
      gen_fill_ring<<<nb((int64_t)count * o.ring_len * L.R), 256, 0, s>>>(
          h->ws + t.x[l], S + o.ring + (int64_t)stream0 * o.ring_len * L.R, count, Win, L.R, o.ring_len, 0);
  }
  for (int i = 0; i < L.n_head; ++i) {
    const ConvParam& cp = h->head[i];
    gen_transpose_conv<<<nb((int64_t)cp.out_ch * cp.in_ch), 256, 0, s>>>(params + cp.w_off, S + L.hw[i], cp.out_ch,
                                                                          cp.in_ch, 1, cp.out_ch, 0);
    gen_copy_bias<<<nb(cp.out_ch), 256, 0, s>>>(cp.b_off >= 0 ? params + cp.b_off : nullptr, S + L.hb[i], cp.out_ch);
  }
  if (L.kc > 1)
    gen_fill_idx_hist<<<nb((int64_t)count * (L.kc - 1)), 256, 0, s>>>(
        window, (int32_t*)(S + L.idx_hist) + (int64_t)stream0 * (L.kc - 1), count, Win, L.kc - 1);
  WN_CHECK_LAUNCH();
  WN_CHECK_CUDA(cudaMemcpyAsync(S + L.layers_dev, g->layers.data(), sizeof(GenLayerOff) * L.L, cudaMemcpyHostToDevice,
                                s));
  if (!g->chunks.empty())
    WN_CHECK_CUDA(cudaMemcpyAsync(S + L.chunks_dev, g->chunks.data(), sizeof(GenChunk) * g->chunks.size(),
                                  cudaMemcpyHostToDevice, s));
  if (g->v3_ok) {
    for (const auto& p : g->packs3) {
      gen_pack_v3<<<nb((int64_t)p.K * p.N), 256, 0, s>>>(S + p.src, S + p.dst, p.K, p.N, p.chunkK, p.mode);
      WN_CHECK_LAUNCH();
    }
    WN_CHECK_CUDA(cudaMemcpyAsync(S + L.chunks3_dev, g->chunks3.data(), sizeof(GenChunk) * g->chunks3.size(),
                                  cudaMemcpyHostToDevice, s));
  }
  if (g->v4_ok) {
    int64_t o4 = 0;
    for (int l = 0; l < L.L; ++l) {
      gen_pack_v4<<<nb(128 * 128), 256, 0, s>>>(S + g->layers[l].wa, S + L.wpk4 + o4, L.wpk4_rank, 128, 128, 1);
      o4 += V4_WA_F;
      gen_pack_v4<<<nb(64 * 320), 256, 0, s>>>(S + g->layers[l].wb, S + L.wpk4 + o4, L.wpk4_rank, 64, 320, 2);
      o4 += V4_WB_F;
    }
    for (int i = 0; i < L.n_head; ++i) {
      gen_pack_v4<<<nb(256 * 256), 256, 0, s>>>(S + L.hw[i], S + L.wpk4 + o4, L.wpk4_rank, 256, 256, 3);
      o4 += V4_WH_F;
    }
    WN_CHECK_LAUNCH();
  }
  if (g->v6_ok) {
    uint8_t* w6 = reinterpret_cast<uint8_t*>(S + L.wpk6);
    const int cs = g->v6_cs, gc = 64 / cs, psn = 64 + 256 / cs, nsub = 256 / cs / 32;
    const int64_t chunk = cs == 4 ? V6C<4>::CHUNK : V6C<8>::CHUNK;
    const int64_t gbytes = 16 * 4 * gc * 16;
    for (int l = 0; l < L.L; ++l) {
      uint8_t* d = w6 + (int64_t)l * chunk;
      gen_pack_v6<<<nb((int64_t)cs * 16 * 4 * gc * 8), 256, 0, s>>>(S + g->layers[l].wa, d, L.wpk6_rank_bytes, 1, 128, cs, 0);
      gen_pack_v6<<<nb((int64_t)cs * 8 * 2 * psn * 8), 256, 0, s>>>(S + g->layers[l].wb, d + gbytes, L.wpk6_rank_bytes, 2, 320, cs, 0);
    }
    for (int i = 0; i < L.n_head; ++i)
      for (int j = 0; j < nsub; ++j)
        gen_pack_v6<<<nb((int64_t)cs * 32 * 64 * 8), 256, 0, s>>>(S + L.hw[i], w6 + (int64_t)L.L * chunk + (int64_t)(i * nsub + j) * 32768,
                                                                  L.wpk6_rank_bytes, 4, 256, cs, j);
    WN_CHECK_LAUNCH();
  }
  g->ring6_valid = false;
  g->ringf_valid = true;
  g->primed = stream0 + count == L.n;
  g->t = Win;
  g->steps_done = 0;
  return WN_OK;
}

static void fill_args(wn_gen* g, GenArgs* a) {
  *a = GenArgs{};
  a->state = g->state;
  a->lay = g->lay;
  a->layers = (const GenLayerOff*)(g->state + g->lay.layers_dev);
  a->chunks = (const GenChunk*)(g->state + g->lay.chunks_dev);
  a->n_chunks = (int)g->chunks.size();
  a->t0 = g->t;
  a->head_elu = g->head_act == 1;
}

extern "C" int wn_gen_step(wn_gen* g, const float* params, const int32_t* x_new, int apply_softmax, float* probs,
                           wn_stream_t st) {
  (void)params;
  WN_REQUIRE(g && x_new, WN_EINVAL, "null argument");
  WN_REQUIRE(g->primed, WN_ESTATE, "wn_gen_step: call wn_gen_prime first");
  GenArgs a;
  fill_args(g, &a);
  a.n_steps = 1;
  a.sample_first = 0;
  a.forced = x_new;
  a.probs = probs;
  a.apply_softmax = apply_softmax;
  WN_TRY(run_gen(g, a, (cudaStream_t)st));
  g->t += 1;
  g->steps_done += 1;
  return WN_OK;
}

// logits of the NEXT sample as left by the last priming call / step (before softmax): what the reference returns from
// _forward_one_step(apply_softmax=False) (faster_wavenet.py:50-63, last column)
extern "C" int wn_gen_logits(wn_gen* g, float* logits, wn_stream_t st) {
  WN_REQUIRE(g && logits, WN_EINVAL, "null argument");
  WN_REQUIRE(g->primed, WN_ESTATE, "wn_gen_logits: call wn_gen_prime first");
  const GenLayout& L = g->lay;
  WN_CHECK_CUDA(cudaMemcpyAsync(logits, g->state + L.cur_logits, sizeof(float) * L.n * L.Q, cudaMemcpyDeviceToDevice,
                                (cudaStream_t)st));
  return WN_OK;
}

extern "C" int wn_gen_run(wn_gen* g, const float* params, int n_steps, int mode, uint64_t seed, int32_t* out,
                          wn_stream_t st) {
  (void)params;
  WN_REQUIRE(g && out, WN_EINVAL, "null argument");
  WN_REQUIRE(g->primed, WN_ESTATE, "wn_gen_run: call wn_gen_prime first");
  WN_REQUIRE(n_steps >= 1, WN_EINVAL, "n_steps must be >= 1");
  WN_REQUIRE(mode == WN_GEN_GREEDY || mode == WN_GEN_SAMPLE, WN_EINVAL, "unknown sampling mode");
  GenArgs a;
  fill_args(g, &a);
  a.n_steps = n_steps;
  a.mode = mode;
  a.sample_first = 1;
  a.seed = seed;
  a.out = out;
  WN_TRY(run_gen(g, a, (cudaStream_t)st));
  g->t += n_steps;
  g->steps_done += n_steps;
  return WN_OK;
}

// ---- data.py on device -------------------------------------------------------------
namespace {
__global__ void crop_batch_kernel(const int32_t* __restrict__ signal, int64_t len, const int32_t* __restrict__ starts,
                                  int B, int iw, int tw, int32_t* __restrict__ x, int32_t* __restrict__ tgt) {
  const int n = blockIdx.y;
  const int64_t st = starts[n];
  const int wx = iw + tw;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < wx + tw; i += gridDim.x * blockDim.x) {
    if (i < wx) {
      const int64_t p = st + i;
      x[(int64_t)n * wx + i] = p < len ? signal[p] : 127;
    } else {
      const int j = i - wx;
      const int64_t p = st + iw + 1 + j;
      tgt[(int64_t)n * tw + j] = p < len ? signal[p] : 127;
    }
  }
}
}  // namespace

extern "C" int wn_crop_batch(const int32_t* signal, int64_t signal_len, const int32_t* starts, int B, int input_width,
                             int target_width, int32_t* x, int32_t* tgt, wn_stream_t s) {
  WN_REQUIRE(signal && starts && x && tgt && B >= 1 && input_width >= 0 && target_width >= 1, WN_EINVAL, "bad argument");
  dim3 grid((unsigned)((input_width + 2 * target_width + 255) / 256), (unsigned)B);
  crop_batch_kernel<<<grid, 256, 0, (cudaStream_t)s>>>(signal, signal_len, starts, B, input_width, target_width, x, tgt);
  WN_CHECK_LAUNCH();
  return WN_OK;
}

namespace {
__global__ void mulaw_encode_kernel(const double* __restrict__ sig, int64_t n, int Q, int32_t* __restrict__ q) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double mu = (double)(Q - 1);
  const double x = sig[i];
  const double sgn = x > 0.0 ? 1.0 : (x < 0.0 ? -1.0 : 0.0);
  const double y = sgn * log(1.0 + mu * fabs(x)) / log(1.0 + mu);   // data.py:20
  double c = y * 0.5 + 0.5;
  c = c < 0.0 ? 0.0 : (c > 1.0 ? 1.0 : c);
  q[i] = (int32_t)(c * mu);                                            // truncation, data.py:23
}
__global__ void mulaw_decode_kernel(const int32_t* __restrict__ q, int64_t n, int Q, double scale,
                                    double* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double nrm = ((double)q[i] / (double)Q - 0.5) * 2.0;           // data.py:39
  const double mu = (double)(Q - 1);
  const double sgn = nrm > 0.0 ? 1.0 : (nrm < 0.0 ? -1.0 : 0.0);
  out[i] = sgn * pow(1.0 + mu, fabs(nrm)) / mu * scale;               // data.py:43,54
}
}  // namespace

extern "C" int wn_mulaw_encode(const double* signal, int64_t n, int quantization_steps, int32_t* q, wn_stream_t s) {
  WN_REQUIRE(signal && q && n >= 0, WN_EINVAL, "bad argument");
  if (n == 0) return WN_OK;
  mulaw_encode_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)s>>>(signal, n, quantization_steps, q);
  WN_CHECK_LAUNCH();
  return WN_OK;
}

extern "C" int wn_mulaw_decode(const int32_t* q, int64_t n, int quantization_steps, double scale, double* out,
                               wn_stream_t s) {
  WN_REQUIRE(q && out && n >= 0, WN_EINVAL, "bad argument");
  if (n == 0) return WN_OK;
  mulaw_decode_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)s>>>(q, n, quantization_steps, scale, out);
  WN_CHECK_LAUNCH();
  return WN_OK;
}
