// gen_kernel_v6<CS>: the MANY-stream generator on the tensor cores (included by wn_gen.cu inside its anonymous namespace).
//
// FasterWaveNet._forward_one_step (faster_wavenet.py:50-113) + the sampling loop of train_audio/generate.py:24-43 for up
// to 128 streams per cluster of CS = 8 or 4 CTAs.  The SIMT generators feed every FMA one operand from shared memory (v3:
// 5 MB of weights per CTA and step; v5: ~4 k LDS wavefronts per layer), so they stop scaling at one or two streams per SM.
// Here the STREAMS are the M dimension of tcgen05.mma (M = 128, one TMEM lane per stream) and the CTAs of a cluster split
// every weight matrix by OUTPUT rows like v4/v5 (packed fp16 hi|lo B tiles streamed through a 2-stage cp.async.bulk ring:
// 1 MB per CTA and step with 8 CTAs, 1.5 MB with 4).  Per layer and CTA (CS = 8 / 4):
//   gate      D[128 x 16 / 32]  = [x(t-d) | x(t)] (K = 128) . Wf/Wg rows of this rank.  The x(t-d) half is issued one layer
//                                 ahead into the other gate accumulator; the x(t) half takes its A operand FROM TENSOR MEMORY
//                                 (written by the previous epilogue with tcgen05.st): no shared-memory read on the chain
//   project   D[128 x 64]       = z (K = 64) . Wp (all 64 rows, computed redundantly by every CTA: ONE exchange per layer)
//   skip      D[128 x 32 / 64] += z . Ws rows of this rank -- one N = 96 / 128 GEMM with the projection; both accumulate in
//                                 TMEM over V6_FLUSH layers (the tensor core accumulates round-toward-zero), then move to
//                                 fp32 registers
//   head      D[128 x 32]       = h (K = 256) . W rows of this rank, twice (CS = 4: two 32-output sub-chunks per conv)
// Arithmetic is the fp16x2 split of the training path (wn_tcs.cu): A.B ~ A_hi.B_hi + A_hi.B_lo + A_lo.B_hi with fp32
// accumulation.  For the gate and the head the hi and lo weight planes are adjacent rows of one B tile, so a product is TWO
// MMAs (A_hi . [B_hi|B_lo] with twice the N, A_lo . B_hi) and the epilogue adds the two column halves; the project + skip
// GEMM keeps three MMAs into one accumulator (its epilogue is bound by the 64 B/clk TMEM read).
// Operand tiles use the NO-SWIZZLE K-major canonical layout (8-row x 16-byte core matrices): tile[k_core][plane][128 rows]
// [16 B] -- a CTA's slice of an exchanged activation (its 8 or 16 gate channels = one or two k_cores, both planes, all 128
// streams) is then one contiguous block and travels with ONE cp.async.bulk shared::cta -> shared::cluster per destination
// (complete_tx on the destination's mbarrier, the data path the MMA's async proxy reads without a cross-proxy fence).
// The all-gathers are the long pole (a DSMEM port moves 17-21 B/clk), so consumers are pipelined against them: one barrier
// per K step of the z tile / per source rank of the head tile, and the MMA thread walks them in arrival order.
// The dilation rings hold whole operand tiles (x(t) of a layer for all 128 streams, 32 KB): one bulk store per layer and
// step, one bulk load d steps later, private to each CTA (no cross-CTA ordering of global memory on the chain).
// Warp roles: 8 epilogue warps (TMEM lane quarter = warp & 3, column half = warp >> 2), one producer thread (weights,
// ring loads / stores), one MMA-issuing thread.  Measured: 98 us per step (CS = 8), 101-108 us (CS = 4) at any stream count.

constexpr int V6_T = 256;                       // epilogue threads
constexpr int V6_THREADS = V6_T + 64;           // + producer warp + MMA warp
constexpr int V6_TILE = 32768;                  // [8 k_cores][2 planes][128 rows][16 B]
constexpr int V6_FLUSH = 5;                     // layers per TMEM accumulation group (projection and skip sums)
constexpr int V6_MAXL = 64;
constexpr float V6_ACT = 8.f, V6_WSC = 16.f, V6_INV = 1.f / (V6_ACT * V6_WSC);
constexpr uint32_t V6_TM_XH = 256, V6_TM_XL = 288;   // x(t) as the A operand IN TMEM: fp16 pairs, 32 columns per plane

// Everything that depends on the cluster size CS (8: up to 15 clusters = 1920 streams per GPU; 4: every CTA owns twice the
// output rows -- same MMA count, the MMAs are bound by their 128-row A operand -- and ~36 clusters = 4608 streams fit)
template <int CS>
struct V6C {
  static constexpr int GC = 64 / CS;            // gate channels per CTA
  static constexpr int GH = GC / 2;             // ... per epilogue thread (column half)
  static constexpr int SKC = 256 / CS;          // skip channels = head outputs per CTA
  static constexpr int NLG = SKC / 2;           // ... per epilogue thread
  static constexpr int NSUB = SKC / 32;         // 32-output sub-chunks per head conv
  static constexpr int NXD = CS == 8 ? 2 : 1;   // x(t-d) buffers
  static constexpr int G_ROWS = 4 * GC;         // gate B rows per k_core: hi (2 halves x (f GH | g GH)) | lo (same)
  static constexpr int G_KS = G_ROWS * 16;      // bytes between k_cores
  static constexpr int G_BYTES = 16 * G_KS;
  static constexpr int PS_N = 64 + SKC;         // project (all 64 rows, every CTA) | skip (this CTA's rows)
  static constexpr int PS_KS = PS_N * 16;
  static constexpr int PS_PLANE = 8 * PS_KS;
  static constexpr int CHUNK = G_BYTES + 2 * PS_PLANE;     // one layer: 32 KB (CS = 8) / 48 KB (CS = 4) = one weight stage
  static constexpr int HCHUNK = 32768;                     // 32 head outputs: [k_core 32][hi 32 | lo 32][16 B]
  static constexpr int OFF_X = 0;                          // x(t) tile (source of the ring store)
  static constexpr int OFF_XD = V6_TILE;                   // x(t-d) tile(s)
  static constexpr int OFF_Z = OFF_XD + NXD * V6_TILE;     // 2 z tiles
  static constexpr int OFF_CAND = OFF_Z + V6_TILE;         // sampling candidates live in z tile 1 (idle between head and layer 1)
  static constexpr int OFF_W = OFF_Z + 2 * V6_TILE;        // 2 weight stages
  static constexpr int SMEM = OFF_W + 2 * CHUNK;           // 224 KB either way; the head's K = 256 tile overlays [0, 128 KB)
  static constexpr bool CAND_IN_H = OFF_CAND < 4 * V6_TILE;
  static constexpr int Z_SLICE = GC / 8 * 4096, H_SLICE = SKC / 8 * 4096;
  static constexpr uint32_t TM_GATE = 0, TM_PS = 128, TM_HEAD = 320;   // gate: 2 x G_ROWS columns (layer parity)
};

__host__ __device__ constexpr uint32_t v6_idesc(int M, int N) {   // kind::f16, fp16 operands, fp32 accumulate, K-major A and B
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void v6_umma(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// K-major operand without swizzle (cute canonical layout ((8,n),2):((1,SBO),LBO) in 16-byte units): LBO = bytes between the
// two 8-element K chunks of one MMA, SBO = bytes between 8-row groups.
__device__ __forceinline__ uint64_t v6_desc(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(lbo >> 4) << 16;
  d |= (uint64_t)(sbo >> 4) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
__device__ __forceinline__ void v6_ld8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ uint32_t v6_pack(float a, float b) {
  uint32_t d;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(b), "f"(a));
  return d;
}
__device__ __forceinline__ void v6_split2(float a, float b, uint32_t& hi, uint32_t& lo) {
  hi = v6_pack(a, b);
  const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&hi));
  lo = v6_pack(a - f.x, b - f.y);
}
// eight consecutive channels of one row -> the 16-byte hi chunk and the 16-byte lo chunk (2048 bytes further) of a tile
__device__ __forceinline__ void v6_store8(uint32_t addr, const float* v, float scale) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) v6_split2(v[2 * i] * scale, v[2 * i + 1] * scale, h[i], l[i]);
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(h[0]), "r"(h[1]), "r"(h[2]), "r"(h[3]) : "memory");
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr + 2048), "r"(l[0]), "r"(l[1]), "r"(l[2]), "r"(l[3]) : "memory");
}
// A operand from TENSOR MEMORY (128 lanes x 8 columns of fp16 pairs per K = 16 step): no shared-memory read at all
__device__ __forceinline__ void v6_umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void v6_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]),
      "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void v6_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// 32 channels of one stream -> (1) the TMEM A operand of the gate GEMM, signalled to the MMA thread at once; (2) the
// shared-memory tile the producer rolls into the dilation ring (off the chain)
__device__ __forceinline__ void v6_publish_x(const float* x, uint32_t tm_lane, int hh, uint32_t xs, uint32_t bar_xt, uint32_t bar_xr, int lane) {
  uint32_t h[16], l[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) v6_split2(x[2 * i] * V6_ACT, x[2 * i + 1] * V6_ACT, h[i], l[i]);
  v6_st16(tm_lane + V6_TM_XH + 16 * hh, h);
  v6_st16(tm_lane + V6_TM_XL + 16 * hh, l);
  v6_st_wait();
  tc::tcgen05_fence_before();
  __syncwarp();
  if (lane == 0) tc::mbar_arrive(bar_xt);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(xs + j * 4096), "r"(h[4 * j]), "r"(h[4 * j + 1]), "r"(h[4 * j + 2]),
                 "r"(h[4 * j + 3])
                 : "memory");
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(xs + j * 4096 + 2048), "r"(l[4 * j]), "r"(l[4 * j + 1]), "r"(l[4 * j + 2]),
                 "r"(l[4 * j + 3])
                 : "memory");
  }
  tc::fence_proxy_async();
  __syncwarp();
  if (lane == 0) tc::mbar_arrive(bar_xr);
}

// waits that acquire at cluster scope (payload written by a peer's bulk copy / a peer's remote arrive)
__device__ __forceinline__ void v6_wait_cluster(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  for (;;) {
    uint32_t ok;
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (ok) return;
    if (++spins > (1u << 25)) {
      printf("wavenet_b200: generator v6 exchange timeout (block %d thread %d bar 0x%x parity %u)\n", (int)blockIdx.x,
             (int)threadIdx.x, bar, parity);
      __trap();
    }
  }
}
__device__ __forceinline__ void v6_remote_arrive(uint32_t bar_cluster) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster) : "memory");
}
__device__ __forceinline__ void v6_esync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
__device__ __forceinline__ void v6_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void v6_bulk_s2g(void* dst, uint32_t src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(reinterpret_cast<uint64_t>(dst)), "r"(src),
               "r"(bytes)
               : "memory");
}

#ifdef WN_LAYER_TRACE
#define TR6(l, e) do { if (blockIdx.x == 0 && step == 2 && (l) < 64) g_trace_gen[(l) * 16 + (e)] = clock64(); } while (0)
#else
#define TR6(l, e) do { } while (0)
#endif

// ELU / ReLU of the head on the per-step chain: e^v - 1 through one MUFU.EX2 (absolute error < 2e-7; expm1f is ~40
// instructions, 16 values per thread and head conv)
__device__ __forceinline__ float v6_head_act(float v, int elu) {
  if (!elu) return fmaxf(v, 0.f);
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fminf(v, 0.f) * 1.4426950408889634f));
  return v > 0.f ? v : e - 1.f;
}

template <int N>
__device__ __forceinline__ void v6_ldn(uint32_t taddr, uint32_t (&r)[N]) {
  static_assert(N == 8 || N == 16 || N == 32, "tcgen05.ld width");
  if constexpr (N == 8) v6_ld8(taddr, r);
  else if constexpr (N == 16) tc::tmem_ld16(taddr, r);
  else tc::tmem_ld32(taddr, r);
}

struct V6Args {
  const uint8_t* wpk;       // [rank][L x CHUNK | n_head x NSUB x 32 KB]
  int64_t wpk_rank_bytes;
  uint8_t* ring;            // [cluster][rank][slot][32 KB]
  int64_t ring_cta_bytes;
  int spc;                  // streams per cluster (<= 128)
  int cluster0;             // first cluster of this launch (more clusters than fit at once run as consecutive launches)
};

template <int CS>
__global__ void __launch_bounds__(V6_THREADS, 1) gen_kernel_v6(GenArgs a, V6Args v) {
  using C = V6C<CS>;
  extern __shared__ uint8_t sm6_raw[];
  const GenLayout& L = a.lay;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int rank = (int)cluster_rank();
  const int cluster = blockIdx.x / CS + v.cluster0;
  const int stream0 = cluster * v.spc;
  const int ns = min(v.spc, L.n - stream0);
  const int NL = L.L;
  constexpr int Q = 256, R = 64;
  uint8_t* smp = sm6_raw + ((128 - (tc::smem_u32(sm6_raw) & 127)) & 127);
  const uint32_t sb = tc::smem_u32(smp);
  // barriers
  // B_ZF: [z buffer 2][K step 4] -- the project + skip GEMM consumes z K step by K step as the slices land;  B_HF: one per
  // source rank of a head-tile slice
  enum { B_WF = 0, B_WE = 2, B_XDF = 4, B_XDE = 6, B_XR = 8, B_XF = 9, B_GD = 10, B_PD = 11, B_HD = 12, B_HFREE = 13, B_CF = 14,
         B_XT = 15, B_CFREE = 16, B_ZF = 17, B_HF = 25, B_N = 33 };
  __shared__ __align__(8) uint64_t s_bar[B_N];
  __shared__ uint32_t s_tmem;
  __shared__ int s_len[V6_MAXL], s_base[V6_MAXL];
  const uint32_t bar0 = tc::smem_u32(&s_bar[0]);
  auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      tc::mbar_init(BAR(B_WF + i), 1);
      tc::mbar_init(BAR(B_WE + i), 1);
      tc::mbar_init(BAR(B_XDF + i), 1);
      tc::mbar_init(BAR(B_XDE + i), 1);
    }
    for (int i = 0; i < 8; ++i) {
      tc::mbar_init(BAR(B_ZF + i), 1);
      tc::mbar_init(BAR(B_HF + i), 1);
    }
    tc::mbar_init(BAR(B_XR), V6_T / 32);
    tc::mbar_init(BAR(B_XT), V6_T / 32);
    tc::mbar_init(BAR(B_XF), 1);
    tc::mbar_init(BAR(B_GD), 1);
    tc::mbar_init(BAR(B_PD), 1);
    tc::mbar_init(BAR(B_HD), 1);
    tc::mbar_init(BAR(B_HFREE), CS);
    tc::mbar_init(BAR(B_CFREE), CS);
    tc::mbar_init(BAR(B_CF), 1);
    tc::fence_barrier_init();
    int base = 0;
    for (int l = 0; l < NL; ++l) {
      s_len[l] = a.layers[l].ring_len;
      s_base[l] = base;
      base += a.layers[l].ring_len;
    }
  }
  if (warp == 9) tc::tmem_alloc<512>(tc::smem_u32(&s_tmem));
  tc::tcgen05_fence_before();
  __syncthreads();
  tc::tcgen05_fence_after();
  const uint32_t tmem = s_tmem;
  cluster_sync_all();

  const int n_chunks = NL + 2 * C::NSUB;
  const uint8_t* wbase = v.wpk + (int64_t)rank * v.wpk_rank_bytes;
  uint8_t* ring = v.ring + (int64_t)(cluster * CS + rank) * v.ring_cta_bytes;

  if (warp == 8) {
    // ================= producer: weights ring, x(t-d) tile loads, x(t) tile stores =================
    if (lane == 0) {
      uint32_t n_w[2] = {0, 0};      // loads issued per weight stage
      uint32_t n_xd[2] = {0, 0};     // loads issued per x(t-d) buffer
      uint32_t n_xr = 0;             // x tiles seen
      uint32_t it = 0;
      auto load_w = [&](int chunk) {
        const uint32_t st = it & 1;
        if (n_w[st] > 0) tc::mbar_wait(BAR(B_WE + st), (n_w[st] - 1) & 1);
        const uint32_t bytes = chunk < NL ? C::CHUNK : C::HCHUNK;
        const int64_t off = chunk < NL ? (int64_t)chunk * C::CHUNK : (int64_t)NL * C::CHUNK + (int64_t)(chunk - NL) * C::HCHUNK;
        tc::mbar_arrive_expect_tx(BAR(B_WF + st), bytes);
        v6_bulk_g2s(sb + C::OFF_W + st * C::CHUNK, wbase + off, bytes, BAR(B_WF + st));
        ++n_w[st];
        ++it;
      };
      auto load_xd = [&](int l, int64_t t) {
        const uint32_t b = C::NXD == 2 ? (l & 1) : 0;
        if (n_xd[b] > 0) tc::mbar_wait(BAR(B_XDE + b), (n_xd[b] - 1) & 1);
        asm volatile("cp.async.bulk.wait_group 2;" ::: "memory");      // the slot's last store (>= 3 groups ago) has landed
        tc::mbar_arrive_expect_tx(BAR(B_XDF + b), V6_TILE);
        const int pos = (int)(t % s_len[l]);
        v6_bulk_g2s(sb + C::OFF_XD + b * V6_TILE, ring + (int64_t)(s_base[l] + pos) * V6_TILE, V6_TILE, BAR(B_XDF + b));
        ++n_xd[b];
      };
      load_w(0);
      load_w(1);
      load_xd(0, a.t0);
      if (C::NXD == 2 && NL > 1) load_xd(1, a.t0);
      for (int step = 0; step < a.n_steps; ++step) {
        const int64_t t = a.t0 + step;
        for (int l = 0; l < NL; ++l) {
          const uint32_t b = C::NXD == 2 ? (l & 1) : 0;
          tc::mbar_wait(BAR(B_XR), n_xr & 1);                       // x(t) tile of layer l complete
          ++n_xr;
          TR6(l, 10);
          tc::mbar_wait(BAR(B_XDF + b), (n_xd[b] - 1) & 1);         // its slot has been read: roll the ring (faster_wavenet.py:90-91)
          TR6(l, 11);
          v6_bulk_s2g(ring + (int64_t)(s_base[l] + (int)(t % s_len[l])) * V6_TILE, sb + C::OFF_X, V6_TILE);
          tc::bulk_commit_group();
          tc::bulk_wait_group_read0();
          tc::mbar_arrive(BAR(B_XF));
          TR6(l, 12);
          if (l + C::NXD < NL) load_xd(l + C::NXD, t);
          TR6(l, 13);
          load_w(l + 2);                                            // chunk l + 2 of this step (the head's chunks follow the layers')
          TR6(l, 14);
        }
        const bool more = step + 1 < a.n_steps;
        for (int hc = 0; hc < 2 * C::NSUB; ++hc) {                  // as each head chunk releases its stage: the chunk two ahead
          const int nxt = NL + hc + 2;
          if (nxt < n_chunks) {
            load_w(nxt);
          } else if (more) {
            load_w(nxt - n_chunks);
          } else {
            const uint32_t st = it & 1;
            tc::mbar_wait(BAR(B_WE + st), (n_w[st] - 1) & 1);
            ++it;
          }
        }
        if (more) {                                                 // the head tile overlaid the x(t-d) buffers until now
          load_xd(0, t + 1);
          if (C::NXD == 2 && NL > 1) load_xd(1, t + 1);
        }
      }
      tc::bulk_wait_group0();
    }
  } else if (warp == 9) {
    // ================= MMA issuer =================
    if (lane == 0) {
      uint32_t n_w = 0, n_xr = 0, n_gate[2] = {0, 0}, n_z[2] = {0, 0}, n_h = 0;
      constexpr uint32_t ID_G2 = v6_idesc(128, C::G_ROWS), ID_G1 = v6_idesc(128, C::G_ROWS / 2), ID_S2 = v6_idesc(128, 64),
                         ID_S1 = v6_idesc(128, 32), ID_PS = v6_idesc(128, C::PS_N);
      // Slices of an all-gather arrive from rank - 1, rank - 2, ... (every sender serves rank + 1 first); the own slice is local.
      // K step ks of the z tile holds the slices of ranks [ks CS / 4, (ks + 1) CS / 4): order the K steps by the arrival of
      // their last slice.
      int zorder[4];
      {
        int ready[4];
        for (int ks = 0; ks < 4; ++ks) {
          int last = 0;
          for (int m = ks * CS / 4; m < (ks + 1) * CS / 4; ++m) {
            const int pos = m == rank ? 0 : ((rank - m - 1) & (CS - 1)) + 1;
            last = pos > last ? pos : last;
          }
          ready[ks] = last;
          zorder[ks] = ks;
        }
        for (int i = 0; i < 4; ++i)
          for (int j = i + 1; j < 4; ++j)
            if (ready[zorder[j]] < ready[zorder[i]]) {
              const int tmp = zorder[i];
              zorder[i] = zorder[j];
              zorder[j] = tmp;
            }
      }
      // x(t-d) half of a layer's gate GEMM: independent of the sample in flight, issued one layer ahead into the other
      // gate accumulator
      auto gate_early = [&](int l, uint32_t chunk) {
        const uint32_t st = chunk & 1, b = C::NXD == 2 ? (l & 1) : 0;
        const uint32_t wst = sb + C::OFF_W + st * C::CHUNK;
        const uint32_t dg = tmem + C::TM_GATE + C::G_ROWS * (l & 1);
        tc::mbar_wait(BAR(B_WF + st), (chunk >> 1) & 1);
        tc::mbar_wait(BAR(B_XDF + b), n_gate[b] & 1);
        ++n_gate[b];
        tc::tcgen05_fence_after();
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const uint32_t at = sb + C::OFF_XD + b * V6_TILE + (uint32_t)ks * 8192;
          const uint64_t bd = v6_desc(wst + ks * 2 * C::G_KS, C::G_KS, 128);
          v6_umma(dg, v6_desc(at, 4096, 128), bd, ID_G2, ks > 0);
          v6_umma(dg, v6_desc(at + 2048, 4096, 128), bd, ID_G1, 1u);
        }
        tc::umma_commit(BAR(B_XDE + b));       // the x(t-d) buffer is free as soon as these have read it
      };
      for (int step = 0; step < a.n_steps; ++step) {
        for (int l = 0; l < NL; ++l, ++n_w) {
          const uint32_t st = n_w & 1, b = l & 1;
          const uint32_t wst = sb + C::OFF_W + st * C::CHUNK;
          const uint32_t dg = tmem + C::TM_GATE + C::G_ROWS * b;
          if (l == 0) gate_early(0, n_w);
          TR6(l, 6);
          tc::mbar_wait(BAR(B_XT), n_xr & 1);
          ++n_xr;
          TR6(l, 7);
          tc::tcgen05_fence_after();
#pragma unroll
          for (int ks = 4; ks < 8; ++ks) {
            const uint64_t bd = v6_desc(wst + ks * 2 * C::G_KS, C::G_KS, 128);
            v6_umma_ts(dg, tmem + V6_TM_XH + 8 * (ks & 3), bd, ID_G2, 1u);
            v6_umma_ts(dg, tmem + V6_TM_XL + 8 * (ks & 3), bd, ID_G1, 1u);
          }
          tc::umma_commit(BAR(B_GD));
          const uint32_t zt = sb + C::OFF_Z + b * V6_TILE;
          const uint32_t grp_acc = (l % V6_FLUSH) != 0;
#pragma unroll
          for (int i = 0; i < 4; ++i) {              // K steps in the order their z slices land (the DSMEM all-gather is the long pole)
            const int ks = zorder[i];
            v6_wait_cluster(BAR(B_ZF + 4 * b + ks), n_z[b] & 1);
            if (i == 0) TR6(l, 8);
            tc::tcgen05_fence_after();
            const uint64_t ah = v6_desc(zt + ks * 8192, 4096, 128), al = v6_desc(zt + ks * 8192 + 2048, 4096, 128);
            const uint64_t bh = v6_desc(wst + C::G_BYTES + ks * 2 * C::PS_KS, C::PS_KS, 128);
            const uint64_t bl = v6_desc(wst + C::G_BYTES + C::PS_PLANE + ks * 2 * C::PS_KS, C::PS_KS, 128);
            v6_umma(tmem + C::TM_PS, ah, bl, ID_PS, grp_acc | (i > 0));      // cross terms first, hi.hi last (round-toward-zero)
            v6_umma(tmem + C::TM_PS, al, bh, ID_PS, 1u);
            v6_umma(tmem + C::TM_PS, ah, bh, ID_PS, 1u);
          }
          ++n_z[b];
          tc::umma_commit(BAR(B_PD));
          tc::umma_commit(BAR(B_WE + st));
          TR6(l, 9);
          if (l + 1 < NL) gate_early(l + 1, n_w + 1);
        }
        for (int hi = 0; hi < 2; ++hi) {
          for (int j = 0; j < C::NSUB; ++j, ++n_w) {
            const uint32_t st = n_w & 1;
            const uint32_t wst = sb + C::OFF_W + st * C::CHUNK;
            tc::mbar_wait(BAR(B_WF + st), (n_w >> 1) & 1);
            for (int i = 0; i < CS; ++i) {           // slices in arrival order: own, rank - 1, rank - 2, ...
              const int src = (rank - i) & (CS - 1);
              if (j == 0) v6_wait_cluster(BAR(B_HF + src), n_h & 1);
              tc::tcgen05_fence_after();
#pragma unroll
              for (int k = 0; k < C::SKC / 16; ++k) {
                const int ks = src * (C::SKC / 16) + k;
                const uint64_t bd = v6_desc(wst + ks * 2048, 1024, 128);
                v6_umma(tmem + C::TM_HEAD + 64 * j, v6_desc(sb + ks * 8192, 4096, 128), bd, ID_S2, (i | k) > 0);
                v6_umma(tmem + C::TM_HEAD + 64 * j, v6_desc(sb + ks * 8192 + 2048, 4096, 128), bd, ID_S1, 1u);
              }
            }
            if (j == C::NSUB - 1) {
              tc::umma_commit(BAR(B_HD));
              ++n_h;
            }
            tc::umma_commit(BAR(B_WE + st));
          }
        }
      }
    }
  } else {
    // ================= epilogue warps: stream s = TMEM lane, column half hh =================
    constexpr int GH = C::GH, NLG = C::NLG, SKC = C::SKC;
    const int q4 = warp & 3, hh = warp >> 2;
    const int s = 32 * q4 + lane;
    const int sg = stream0 + min(s, ns - 1);           // dead lanes mirror the last live stream and are never stored
    const uint32_t tl = tmem + ((uint32_t)(32 * q4) << 16);
    const uint32_t row_off = (uint32_t)((s >> 3) * 128 + (s & 7) * 16);
    // exchanges: thread j < CS - 1 copies this CTA's slice to rank + 1 + j (rotated, so no destination is everybody's first
    // target: a DSMEM port moves ~17-21 B/clk), thread CS - 1 opens the local phase
    const uint32_t dst_rank = (uint32_t)((rank + 1 + tid) & (CS - 1));
    float* st = a.state;
    float* cur_logits = st + L.cur_logits;
    int32_t* idx_hist = (int32_t*)(st + L.idx_hist);
    const int kc1 = L.kc - 1;
    const float* emb = st + L.emb;
    // this thread's head outputs / classes: sub-chunk j, i < 16  ->  SKC * rank + 32 j + 16 hh + i   (array index 16 j + i)
    auto cls = [&](int j, int i) { return SKC * rank + 32 * j + 16 * hh + i; };
    const bool noisy = a.mode == WN_GEN_SAMPLE;
    // lg: logits of the next sample PLUS its Gumbel noise.  The noise does not depend on the data: it is written into lg
    // inside the layer loop's waits (16 x (3 splitmix64 + 2 logf) would otherwise sit on the per-step chain) and the head's
    // last epilogue adds the logits; no noise behind the last step (the stored state is the plain logits).
    float lg[NLG], xr[32], sk[NLG];
#pragma unroll
    for (int j = 0; j < C::NSUB; ++j) {
      const float4* p = reinterpret_cast<const float4*>(cur_logits + (int64_t)sg * Q + cls(j, 0));
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 v4 = p[i];
        lg[16 * j + 4 * i] = v4.x, lg[16 * j + 4 * i + 1] = v4.y, lg[16 * j + 4 * i + 2] = v4.z, lg[16 * j + 4 * i + 3] = v4.w;
      }
#pragma unroll
      for (int i = 0; i < 16; ++i)
        if (noisy) lg[16 * j + i] += gumbel(a.seed, (uint64_t)sg, (uint64_t)a.t0, (uint32_t)cls(j, i));
    }
    int prev = kc1 > 0 ? idx_hist[(int64_t)sg * kc1 + kc1 - 1] : -1;
    uint32_t n_gd = 0, n_pd = 0, n_xf = 0, n_hd = 0, n_hfree = 0, n_cf = 0, n_cfree = 0;
    for (int step = 0; step < a.n_steps; ++step) {
      const int64_t t = a.t0 + step;
      const bool more = step + 1 < a.n_steps;
      TRG(41, 0);
      // ---- 1. sample (generate.py:38-43): partial arg-max over this thread's classes, candidates to every CTA ----
      {
        float bv = -INFINITY;
        int bi = 0;
#pragma unroll
        for (int j = 0; j < C::NSUB; ++j) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float vv = lg[16 * j + i];
            if (vv > bv) bv = vv, bi = cls(j, i);
          }
        }
#pragma unroll
        for (int i = 0; i < NLG; ++i) lg[i] = 0.f;
        const uint32_t ca = sb + C::OFF_CAND + (uint32_t)(((rank * 2 + hh) * 128 + s) * 8);
        asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(ca), "r"(__float_as_uint(bv)), "r"((uint32_t)bi) : "memory");
        TRG(42, 0);
        tc::fence_proxy_async();
        TRG(45, 0);
        v6_esync();
        TRG(45, 1);
        if (tid < CS) {
          // CS = 4: the candidate buffer lies inside the head tile, so the peers must have finished their last head GEMM
          if (C::CAND_IN_H && step > 0) v6_wait_cluster(BAR(B_CFREE), n_cfree & 1);
          if (tid == CS - 1)
            tc::mbar_arrive_expect_tx(BAR(B_CF), (CS - 1) * 2048);
          else
            bulk_copy_to_cta(map_to_cta(sb + C::OFF_CAND + rank * 2048, dst_rank), sb + C::OFF_CAND + rank * 2048, 2048,
                             map_to_cta(BAR(B_CF), dst_rank));
        }
        if (C::CAND_IN_H && step > 0) ++n_cfree;
        v6_wait_cluster(BAR(B_CF), n_cf & 1);
        ++n_cf;
        TRG(42, 1);
        bv = -INFINITY, bi = 0;
#pragma unroll
        for (int j = 0; j < 2 * CS; ++j) {           // np.argmax: the first maximum wins -> lowest class index on ties
          uint32_t cv, ci;
          asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(cv), "=r"(ci) : "r"(sb + C::OFF_CAND + (uint32_t)((j * 128 + s) * 8)));
          const float fv = __uint_as_float(cv);
          if (fv > bv || (fv == bv && (int)ci < bi)) bv = fv, bi = (int)ci;
        }
        TRG(45, 2);
        if (a.out && rank == 0 && hh == 0 && s < ns) a.out[(int64_t)(stream0 + s) * a.n_steps + step] = bi;
        // ---- 2. embedding of the new sample = causal conv of the one-hot pair (wavenet.py:565-570) ----
        // Every thread needs 2 x 128 B of the embedding table (its stream's half rows of both taps).  Read row-per-lane that is
        // 32 different lines per load instruction (4 k LSU wavefronts per step and CTA, 7 k cycles); instead eight lanes share
        // a row (4 full lines per instruction), add the two taps and hand the sums over through the warp's own 4 KB of the
        // x tile (the bytes this warp overwrites with x(t) right after; nobody else touches them now).
        {
          const uint32_t stg = sb + C::OFF_X + (uint32_t)(4 * hh) * 4096 + 512u * (uint32_t)q4;
          const int kk = lane & 7;
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int i = 4 * it + (lane >> 3);                  // stream of this lane group in this round
            const int bi_i = __shfl_sync(0xffffffffu, bi, i), prev_i = __shfl_sync(0xffffffffu, prev, i);
            float4 v1 = __ldg(reinterpret_cast<const float4*>(emb + ((int64_t)kc1 * Q + bi_i) * R + 32 * hh) + kk);
            if (kc1 > 0 && prev_i >= 0) {
              const float4 v0 = __ldg(reinterpret_cast<const float4*>(emb + (int64_t)prev_i * R + 32 * hh) + kk);
              v1.x += v0.x, v1.y += v0.y, v1.z += v0.z, v1.w += v0.w;
            }
            const uint32_t ad = stg + (uint32_t)(it >> 1) * 4096 + (uint32_t)(it & 1) * 2048 + (uint32_t)(i & 3) * 128 +
                                (uint32_t)((kk ^ (i & 7)) << 4);
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(ad), "f"(v1.x), "f"(v1.y), "f"(v1.z), "f"(v1.w) : "memory");
          }
          __syncwarp();
          const uint32_t rd = stg + (uint32_t)(lane >> 3) * 4096 + (uint32_t)((lane >> 2) & 1) * 2048 + (uint32_t)(lane & 3) * 128;
#pragma unroll
          for (int k = 0; k < 8; ++k)
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                         : "=f"(xr[4 * k]), "=f"(xr[4 * k + 1]), "=f"(xr[4 * k + 2]), "=f"(xr[4 * k + 3])
                         : "r"(rd + (uint32_t)((k ^ (lane & 7)) << 4)));
          __syncwarp();                                          // all rows read before any lane publishes x(t) over them
        }
        prev = bi;
        TRG(42, 2);
      }
      v6_publish_x(xr, tl, hh, sb + C::OFF_X + (uint32_t)(4 * hh) * 4096 + row_off, BAR(B_XT), BAR(B_XR), lane);
#pragma unroll
      for (int i = 0; i < NLG; ++i) sk[i] = 0.f;
      TRG(41, 1);
      // ---- 3. residual layers ----
      for (int l = 0; l < NL; ++l) {
        const uint32_t b = l & 1;
        TRG(l, 0);
        tc::mbar_wait(BAR(B_GD), n_gd & 1);
        ++n_gd;
        tc::tcgen05_fence_after();
        TRG(l, 1);
        {
          // gate accumulator columns: hi.[hi | lo] -> this half's (f GH | g GH) at 2 GH hh, the A_hi . B_lo part 2 GC further
          uint32_t gh[2 * GH], gl[2 * GH];
          v6_ldn<2 * GH>(tl + C::TM_GATE + C::G_ROWS * b + 2 * GH * hh, gh);
          v6_ldn<2 * GH>(tl + C::TM_GATE + C::G_ROWS * b + 2 * C::GC + 2 * GH * hh, gl);
          tc::tmem_ld_wait();
          float zv[GH];
#pragma unroll
          for (int c = 0; c < GH; ++c) {
            const float af = (__uint_as_float(gh[c]) + __uint_as_float(gl[c])) * V6_INV;
            const float ag = (__uint_as_float(gh[GH + c]) + __uint_as_float(gl[GH + c])) * V6_INV;
            zv[c] = tanh_ex2(af) * (0.5f + 0.5f * tanh_ex2(0.5f * ag)) * V6_ACT;      // wavenet.py:362-364
          }
          // channels GC rank + GH hh .. + GH of the z tile: k_core (GC rank + GH hh) / 8, 2 bytes per channel inside it
          const uint32_t ch0 = (uint32_t)(C::GC * rank + GH * hh);
          const uint32_t za = sb + C::OFF_Z + b * V6_TILE + (ch0 >> 3) * 4096 + row_off + (ch0 & 7) * 2;
          if constexpr (GH == 4) {
            uint32_t h0, h1, l0, l1;
            v6_split2(zv[0], zv[1], h0, l0);
            v6_split2(zv[2], zv[3], h1, l1);
            asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(za), "r"(h0), "r"(h1) : "memory");
            asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(za + 2048), "r"(l0), "r"(l1) : "memory");
          } else {
            v6_store8(za, zv, 1.f);
          }
        }
        if (l == NL - 1) {                     // peers answer z of the last layer with head slices that land on the x tile:
          tc::mbar_wait(BAR(B_XF), n_xf & 1);  // the ring store must have read it first
          ++n_xf;
        }
        tc::fence_proxy_async();
        v6_esync();
        if (tid < CS) {
          const uint32_t zs = sb + C::OFF_Z + b * V6_TILE + (uint32_t)rank * C::Z_SLICE;
          if (tid < CS - 1)                    // lands on the destination's barrier of the K step this rank's channels belong to
            bulk_copy_to_cta(map_to_cta(zs, dst_rank), zs, C::Z_SLICE, map_to_cta(BAR(B_ZF + 4 * b + rank * 4 / CS), dst_rank));
          if (tid < 4) {                       // open the local phase of K step `tid`: its remote slices (the own one is in place)
            const int own = rank * 4 / CS == tid ? 1 : 0;
            const uint32_t bytes = (uint32_t)((CS / 4 - own) * C::Z_SLICE);
            if (bytes)
              tc::mbar_arrive_expect_tx(BAR(B_ZF + 4 * b + tid), bytes);
            else
              tc::mbar_arrive(BAR(B_ZF + 4 * b + tid));
          }
        }
        TRG(l, 2);
        if (noisy && more && l < 16) {
#pragma unroll
          for (int j = 0; j < C::NSUB; ++j) {
            const float g = gumbel(a.seed, (uint64_t)sg, (uint64_t)(t + 1), (uint32_t)cls(j, l));
#pragma unroll
            for (int i = 0; i < 16; ++i)
              if (i == l) lg[16 * j + i] = g;
          }
        }
        tc::mbar_wait(BAR(B_PD), n_pd & 1);
        ++n_pd;
        tc::tcgen05_fence_after();
        TRG(l, 3);
        float xn[32];
        {
          uint32_t ph[32];
          tc::tmem_ld32(tl + C::TM_PS + 32 * hh, ph);
          const bool flush = (l % V6_FLUSH) == V6_FLUSH - 1 || l == NL - 1;
          if (flush) {
            uint32_t sh[NLG];
            v6_ldn<NLG>(tl + C::TM_PS + 64 + NLG * hh, sh);
            tc::tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < NLG; ++i) sk[i] = fmaf(__uint_as_float(sh[i]), V6_INV, sk[i]);     // faster_wavenet.py:100
          } else {
            tc::tmem_ld_wait();
          }
          // the projection accumulates in TMEM over a group of V6_FLUSH layers: x = x(group start) + sum (wavenet.py:354)
#pragma unroll
          for (int i = 0; i < 32; ++i) xn[i] = fmaf(__uint_as_float(ph[i]), V6_INV, xr[i]);
          if (flush) {
#pragma unroll
            for (int i = 0; i < 32; ++i) xr[i] = xn[i];
          }
        }
        if (l + 1 < NL) {
          tc::mbar_wait(BAR(B_XF), n_xf & 1);
          ++n_xf;
          v6_publish_x(xn, tl, hh, sb + C::OFF_X + (uint32_t)(4 * hh) * 4096 + row_off, BAR(B_XT), BAR(B_XR), lane);
        }
        TRG(l, 4);
      }
      // ---- 4. head (faster_wavenet.py:105-113): activation of the skip sum, two 256 x 256 convs split by output rows ----
      TRG(40, 0);
      const uint32_t hs = sb + (uint32_t)rank * C::H_SLICE;          // this CTA's k_cores of the K = 256 tile
#pragma unroll
      for (int i = 0; i < NLG; ++i) sk[i] = v6_head_act(sk[i], a.head_elu);
      // skip channels SKC rank + NLG hh + i: NLG / 8 consecutive k_cores
#pragma unroll
      for (int k = 0; k < NLG / 8; ++k) v6_store8(hs + (uint32_t)(NLG / 8 * hh + k) * 4096 + row_off, sk + 8 * k, V6_ACT);
      for (int hi = 0; hi < 2; ++hi) {
        TRG(44, 4 * hi);
        tc::fence_proxy_async();
        tc::tcgen05_fence_before();
        TRG(44, 4 * hi + 1);
        v6_esync();
        TRG(44, 4 * hi + 2);
        if (tid < CS) {
          // every CTA must be done reading the first head tile before anyone overwrites it
          if (hi == 1) v6_wait_cluster(BAR(B_HFREE), n_hfree & 1);
          TRG(44, 4 * hi + 3);
          if (tid < CS - 1) bulk_copy_to_cta(map_to_cta(hs, dst_rank), hs, C::H_SLICE, map_to_cta(BAR(B_HF + rank), dst_rank));
          if (tid == rank)                     // local phase of source rank `tid`
            tc::mbar_arrive(BAR(B_HF + tid));
          else
            tc::mbar_arrive_expect_tx(BAR(B_HF + tid), C::H_SLICE);
        }
        if (hi == 1) ++n_hfree;
        TRG(43, 2 * hi);
        float hb[NLG];                         // head bias of this thread's outputs (loaded under the MMA wait)
#pragma unroll
        for (int j = 0; j < C::NSUB; ++j) {
          const float4* bp = reinterpret_cast<const float4*>(st + L.hb[hi] + cls(j, 0));
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 b4 = L.has_hb ? __ldg(bp + i) : make_float4(0.f, 0.f, 0.f, 0.f);
            hb[16 * j + 4 * i] = b4.x, hb[16 * j + 4 * i + 1] = b4.y, hb[16 * j + 4 * i + 2] = b4.z, hb[16 * j + 4 * i + 3] = b4.w;
          }
        }
        tc::mbar_wait(BAR(B_HD), n_hd & 1);
        ++n_hd;
        TRG(43, 2 * hi + 1);
        tc::tcgen05_fence_after();
        if (hi == 0) {
          if (tid < CS) v6_remote_arrive(map_to_cta(BAR(B_HFREE), (uint32_t)tid));     // my first head GEMM has read the tile
        } else if (C::CAND_IN_H && more) {
          if (tid < CS) v6_remote_arrive(map_to_cta(BAR(B_CFREE), (uint32_t)tid));     // ... and my second one
        }
#pragma unroll
        for (int j = 0; j < C::NSUB; ++j) {
          uint32_t dh[16], dl[16];
          tc::tmem_ld16(tl + C::TM_HEAD + 64 * j + 16 * hh, dh);
          tc::tmem_ld16(tl + C::TM_HEAD + 64 * j + 32 + 16 * hh, dl);
          tc::tmem_ld_wait();
          if (hi == 0) {
            float h1[16];
#pragma unroll
            for (int i = 0; i < 16; ++i)
              h1[i] = v6_head_act(fmaf(__uint_as_float(dh[i]) + __uint_as_float(dl[i]), V6_INV, hb[16 * j + i]), a.head_elu);
            // outputs SKC rank + 32 j + 16 hh + i of the next tile: k_cores 4 j + 2 hh, + 1 of this CTA's slice
            v6_store8(hs + (uint32_t)(4 * j + 2 * hh) * 4096 + row_off, h1, V6_ACT);
            v6_store8(hs + (uint32_t)(4 * j + 2 * hh + 1) * 4096 + row_off, h1 + 8, V6_ACT);
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i)
              lg[16 * j + i] += fmaf(__uint_as_float(dh[i]) + __uint_as_float(dl[i]), V6_INV, hb[16 * j + i]);
          }
        }
        tc::tcgen05_fence_before();
      }
      TRG(40, 1);
    }
    // ---- state for the next call: logits of the next sample, last index ----
    if (s < ns) {
#pragma unroll
      for (int j = 0; j < C::NSUB; ++j) {
        float4* p = reinterpret_cast<float4*>(cur_logits + (int64_t)(stream0 + s) * Q + cls(j, 0));
#pragma unroll
        for (int i = 0; i < 4; ++i)
          p[i] = make_float4(lg[16 * j + 4 * i], lg[16 * j + 4 * i + 1], lg[16 * j + 4 * i + 2], lg[16 * j + 4 * i + 3]);
      }
      if (kc1 > 0 && rank == 0 && hh == 0) idx_hist[(int64_t)(stream0 + s) * kc1 + kc1 - 1] = prev;
    }
  }
  __syncwarp();
  tc::tcgen05_fence_before();
  __syncthreads();
  if (warp == 9) tc::tmem_dealloc<512>(tmem);
  cluster_sync_all();
}

// ---- packing: generator-layout fp32 matrices [K][N] -> per-rank fp16 hi|lo B tiles (x V6_WSC); cs = CTAs per cluster ----
// mode 1: gate (K = 128, N = 128 = f | g): rows of rank r = plane x [half x (f GH | g GH)];  mode 2: project (64 rows, every
// rank) | skip (this rank's 256 / cs rows) as one tile per plane;  mode 4: 32 outputs `sub` of a head conv (rows = plane x 32)
__global__ void gen_pack_v6(const float* __restrict__ src, uint8_t* __restrict__ dst, int64_t rank_bytes, int mode, int ldn, int cs,
                            int sub) {
  const int GC = 64 / cs, GH = GC / 2, SKC = 256 / cs, PSN = 64 + SKC;
  const int K = mode == 1 ? 128 : (mode == 4 ? 256 : 64);
  const int rows = mode == 1 ? 4 * GC : (mode == 2 ? 2 * PSN : 64);
  const int total = cs * (K / 8) * rows * 8;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int e = i & 7, row = (i >> 3) % rows, kc = (i / (8 * rows)) % (K / 8), r = i / (8 * rows * (K / 8));
  const int k = kc * 8 + e;
  int plane, n;
  int64_t off;                                 // element offset inside the rank's block
  if (mode == 1) {                             // [k_core 16][hi 2 GC | lo 2 GC]
    plane = row / (2 * GC);
    const int j = row % (2 * GC), half = j / GC, jj = j % GC, isg = jj / GH, c = jj % GH;
    n = (isg ? 64 : 0) + GC * r + GH * half + c;
    off = ((int64_t)kc * rows + row) * 8 + e;
  } else if (mode == 2) {                      // hi block [k_core 8][project 64 | skip SKC], then the lo block
    plane = row / PSN;
    const int j = row % PSN;
    n = j < 64 ? j : 64 + SKC * r + (j - 64);
    off = (int64_t)plane * (8 * PSN * 8) + ((int64_t)kc * PSN + j) * 8 + e;
  } else {                                     // head outputs SKC r + 32 sub + ..: [k_core 32][hi 32 | lo 32]
    plane = row >> 5;
    n = SKC * r + 32 * sub + (row & 31);
    off = ((int64_t)kc * rows + row) * 8 + e;
  }
  const float w = src[(int64_t)k * ldn + n] * V6_WSC;
  const __half hi = __float2half_rn(w);
  const __half val = plane ? __float2half_rn(w - __half2float(hi)) : hi;
  reinterpret_cast<__half*>(dst + (int64_t)r * rank_bytes)[off] = val;
}

// fp32 rings [stream][slot][64] -> operand tiles, one copy per CTA of the cluster;  and back (from rank 0's copy)
__global__ void gen_ring_to_v6(const float* __restrict__ ringf, uint8_t* __restrict__ ring6, int64_t cta_bytes, int slot_base,
                               int len, int n, int spc, int clusters, int cs) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;      // (cluster, slot, row, k_core)
  if (i >= (int64_t)clusters * len * 128 * 8) return;
  const int kc = (int)(i & 7), row = (int)((i >> 3) & 127);
  const int slot = (int)((i >> 10) % len), cl = (int)(i / ((int64_t)len * 1024));
  const int ns = min(spc, n - cl * spc);
  const int stream = cl * spc + min(row, ns - 1);
  const float* p = ringf + ((int64_t)stream * len + slot) * 64 + kc * 8;
  uint32_t h[4], l[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) v6_split2(p[2 * j] * V6_ACT, p[2 * j + 1] * V6_ACT, h[j], l[j]);
  const int64_t off = (int64_t)(slot_base + slot) * V6_TILE + kc * 4096 + (row >> 3) * 128 + (row & 7) * 16;
  for (int r = 0; r < cs; ++r) {
    uint8_t* d = ring6 + (int64_t)(cl * cs + r) * cta_bytes + off;
    *reinterpret_cast<uint4*>(d) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4*>(d + 2048) = make_uint4(l[0], l[1], l[2], l[3]);
  }
}
__global__ void gen_ring_from_v6(float* __restrict__ ringf, const uint8_t* __restrict__ ring6, int64_t cta_bytes, int slot_base,
                                 int len, int n, int spc, int clusters, int cs) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)clusters * len * 128 * 8) return;
  const int kc = (int)(i & 7), row = (int)((i >> 3) & 127);
  const int slot = (int)((i >> 10) % len), cl = (int)(i / ((int64_t)len * 1024));
  const int ns = min(spc, n - cl * spc);
  if (row >= ns) return;
  const uint8_t* d = ring6 + (int64_t)(cl * cs) * cta_bytes + (int64_t)(slot_base + slot) * V6_TILE + kc * 4096 +
                     (row >> 3) * 128 + (row & 7) * 16;
  const uint4 h = *reinterpret_cast<const uint4*>(d), l = *reinterpret_cast<const uint4*>(d + 2048);
  const uint32_t hw[4] = {h.x, h.y, h.z, h.w}, lw[4] = {l.x, l.y, l.z, l.w};
  float* p = ringf + ((int64_t)(cl * spc + row) * len + slot) * 64 + kc * 8;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&hw[j])), b = __half22float2(*reinterpret_cast<const __half2*>(&lw[j]));
    p[2 * j] = (a.x + b.x) * (1.f / V6_ACT);
    p[2 * j + 1] = (a.y + b.y) * (1.f / V6_ACT);
  }
}
