// C-ABI entry points of the training path: parameter layout, tape planning and
// the phase orchestration that mirrors train_audio/train.py:66-80.
#include <math.h>
#include <atomic>
#include <stdarg.h>
#include <string.h>

#include "wn_common.h"

static thread_local char g_err[1024] = "";

void wn_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static std::atomic<long long> g_launches{0};
void wn_count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
extern "C" int64_t wn_launch_count(int reset) {
  return reset ? g_launches.exchange(0) : g_launches.load();
}

extern "C" int64_t wn_launch_count_add(int64_t n) { return g_launches.fetch_add(n) + n; }
extern "C" const char* wn_last_error(void) { return g_err; }
extern "C" int wn_version(void) { return 100; }

static inline int64_t align_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }

static int ipow(int b, int e) {
  int r = 1;
  for (int i = 0; i < e; ++i) r *= b;
  return r;
}

extern "C" int wn_zero_prefix(int width, int dilation, int filter_width) {
  // DilatedConvolution1D.__call__, wavenet.py:304-340 (quirk Q1)
  if (dilation == 1) return 0;
  int pad = ((-width) % dilation + dilation) % dilation;
  int height = (width + pad) / dilation;
  if (height < filter_width) pad += (filter_width - height) * dilation;
  int zp = (filter_width - 1) * dilation - pad;
  return zp > 0 ? zp : 0;
}

static void add_desc(wn_handle* h, const std::string& name, int ndim, int s0, int s1, int s2, int s3, int64_t* off) {
  wn_param_desc d;
  memset(&d, 0, sizeof(d));
  snprintf(d.name, WN_NAME_LEN, "%s", name.c_str());
  d.ndim = ndim;
  d.shape[0] = s0;
  d.shape[1] = s1;
  d.shape[2] = s2;
  d.shape[3] = s3;
  d.numel = (int64_t)s0 * (ndim > 1 ? s1 : 1) * (ndim > 2 ? s2 : 1) * (ndim > 3 ? s3 : 1);
  d.offset = h->flat_size;
  *off = d.offset;
  h->flat_size = align_up(h->flat_size + d.numel, 64);
  h->param_elems += d.numel;
  h->descs.push_back(d);
}

static void add_conv(wn_handle* h, const std::string& link, int out_ch, int in_ch, int kh, int kw, bool bias,
                     ConvParam* cp) {
  cp->out_ch = out_ch;
  cp->in_ch = in_ch;
  cp->taps = kh * kw;
  add_desc(h, link + "/W", 4, out_ch, in_ch, kh, kw, &cp->w_off);
  if (bias) add_desc(h, link + "/b", 1, out_ch, 1, 1, 1, &cp->b_off);
}

extern "C" int wn_create(const wn_config* c, wn_handle** out) {
  WN_REQUIRE(c && out, WN_EINVAL, "wn_create: null argument");
  WN_REQUIRE(c->n_causal >= 1 && c->n_causal <= WN_MAX_CAUSAL, WN_EINVAL, "causal_conv_channels: need 1..%d entries",
             WN_MAX_CAUSAL);
  WN_REQUIRE(c->n_res_layers >= 1 && c->n_res_layers * c->residual_num_blocks <= 4 * WN_MAX_LAYERS &&
                 c->n_res_layers <= WN_MAX_LAYERS,
             WN_EINVAL, "residual_conv_channels: need 1..%d entries", WN_MAX_LAYERS);
  WN_REQUIRE(c->residual_num_blocks >= 1, WN_EINVAL, "residual_num_blocks must be >= 1");
  WN_REQUIRE(c->n_softmax >= 2 && c->n_softmax <= WN_MAX_HEAD, WN_EINVAL,
             "softmax_conv_channels: need 2..%d entries", WN_MAX_HEAD);
  WN_REQUIRE(c->causal_filter_width >= 1 && c->causal_filter_width <= 4 && c->residual_filter_width >= 1 &&
                 c->residual_filter_width <= 4,
             WN_EINVAL, "filter widths must be in 1..4");
  // Params.check, wavenet.py:172-173
  WN_REQUIRE(c->quantization_steps == c->softmax_channels[c->n_softmax - 1], WN_EINVAL,
             "quantization_steps != softmax_conv_channels[-1]");
  WN_REQUIRE(c->quantization_steps >= 2, WN_EINVAL, "quantization_steps must be >= 2");
  int64_t maxd = ipow(c->residual_filter_width, c->n_res_layers - 1);
  WN_REQUIRE(maxd <= (1 << 20), WN_EINVAL, "dilation too large");
  for (int i = 0; i < c->n_causal; ++i) WN_REQUIRE(c->causal_channels[i] >= 1, WN_EINVAL, "bad causal channel count");
  for (int i = 0; i < c->n_res_layers; ++i)
    WN_REQUIRE(c->residual_channels[i] >= 1, WN_EINVAL, "bad residual channel count");
  for (int i = 0; i < c->n_softmax; ++i) WN_REQUIRE(c->softmax_channels[i] >= 1, WN_EINVAL, "bad softmax channel count");

  wn_handle* h = new wn_handle();
  h->cfg = *c;
  h->Q = c->quantization_steps;
  h->R = c->causal_channels[c->n_causal - 1];
  h->S = c->softmax_channels[0];
  char buf[96];
  // causal stack, wavenet.py:388-396
  h->causal.resize(c->n_causal);
  for (int i = 0; i < c->n_causal; ++i) {
    const int n_in = i == 0 ? h->Q : c->causal_channels[i - 1];
    snprintf(buf, sizeof(buf), "causal_%d", i);
    add_conv(h, buf, c->causal_channels[i], n_in, 1, c->causal_filter_width, !c->causal_no_bias, &h->causal[i]);
  }
  // residual blocks, wavenet.py:412-444
  const int k = c->residual_filter_width;
  for (int j = 0; j < c->residual_num_blocks; ++j) {
    for (int i = 0; i < c->n_res_layers; ++i) {
      ResLayer L;
      L.G = c->residual_channels[i];
      L.dilation = ipow(k, i);
      const int kh = i == 0 ? 1 : k, kw = i == 0 ? k : 1;   // wavenet.py:418-424
      snprintf(buf, sizeof(buf), "residual_%d_block_%d_", j, i);
      const std::string base(buf);
      add_conv(h, base + "wf", L.G, h->R, kh, kw, !c->residual_dilation_no_bias, &L.wf);
      add_conv(h, base + "wg", L.G, h->R, kh, kw, !c->residual_dilation_no_bias, &L.wg);
      add_conv(h, base + "projection_block", h->R, L.G, 1, 1, !c->residual_projection_no_bias, &L.proj);
      add_conv(h, base + "projection_softmax", h->S, L.G, 1, 1, !c->residual_projection_no_bias, &L.skip);
      h->layers.push_back(L);
    }
  }
  // softmax block, wavenet.py:451-455
  h->head.resize(c->n_softmax - 1);
  for (int i = 0; i + 1 < c->n_softmax; ++i) {
    snprintf(buf, sizeof(buf), "softmax_%d", i);
    add_conv(h, buf, c->softmax_channels[i + 1], c->softmax_channels[i], 1, 1, !c->softmax_no_bias, &h->head[i]);
  }
  // a handle may be created on a host without a GPU (layout queries only); when a device is visible it must be sm_100
  const int rc = wn_set_device_info(h);
  if (rc == WN_EARCH) {
    delete h;
    return rc;
  }
  cudaGetLastError();
  *out = h;
  return WN_OK;
}

extern "C" int wn_set_device_info(wn_handle* h) {
  WN_REQUIRE(h, WN_EINVAL, "null handle");
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    cudaGetLastError();
    wn_set_error("no CUDA device is visible");
    return WN_ECUDA;
  }
  int major = 0, sm = 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) {
    cudaGetLastError();
    wn_set_error("cannot query the compute capability of device %d", dev);
    return WN_ECUDA;
  }
  WN_REQUIRE(major == 10, WN_EARCH, "device %d has compute capability %d.x; libwavenet_b200.so holds sm_100a code only", dev,
             major);
  if (cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && sm > 0) h->sm_count = sm;
  return WN_OK;
}

extern "C" int wn_destroy(wn_handle* h) {
  if (h && h->comm) wn_comm_destroy(h);
  delete h;
  return WN_OK;
}

extern "C" int wn_set_precision(wn_handle* h, int prec) {
  WN_REQUIRE(h, WN_EINVAL, "null handle");
  WN_REQUIRE(prec == WN_PREC_FP32 || prec == WN_PREC_TF32 || prec == WN_PREC_F16X2, WN_EINVAL, "unknown precision %d", prec);
  h->prec = prec;
  return WN_OK;
}
extern "C" int wn_get_precision(const wn_handle* h) { return h ? h->prec : WN_EINVAL; }

extern "C" int64_t wn_flat_size(const wn_handle* h) { return h ? h->flat_size : WN_EINVAL; }
extern "C" int64_t wn_param_elems(const wn_handle* h) { return h ? h->param_elems : WN_EINVAL; }
extern "C" int wn_num_params(const wn_handle* h) { return h ? (int)h->descs.size() : WN_EINVAL; }
extern "C" int wn_param_layout(const wn_handle* h, wn_param_desc* out, int max_out) {
  WN_REQUIRE(h && out, WN_EINVAL, "null argument");
  const int n = (int)h->descs.size() < max_out ? (int)h->descs.size() : max_out;
  for (int i = 0; i < n; ++i) out[i] = h->descs[i];
  return n;
}
extern "C" int wn_receptive_width(const wn_handle* h) {
  if (!h) return WN_EINVAL;
  return (ipow(h->cfg.residual_filter_width, h->cfg.n_res_layers) - 1) * h->cfg.residual_num_blocks + 1;
}
extern "C" int wn_input_width(const wn_handle* h) { return h ? wn_receptive_width(h) + h->cfg.n_causal : WN_EINVAL; }

// ---- tape ------------------------------------------------------------------------
static void plan_tape(const wn_handle* h, int B, int W, Tape* t) {
  const wn_config& c = h->cfg;
  int64_t off = 0;
  auto take = [&](int64_t n) {
    int64_t o = off;
    off = align_up(off + n, 64);
    return o;
  };
  t->B = B;
  t->W = W;
  t->P = (int64_t)B * W;
  const int64_t P = t->P;
  const int64_t embn = (int64_t)c.causal_filter_width * h->Q * c.causal_channels[0];
  t->emb = take(embn);
  t->demb = take(embn);
  t->cx.clear();
  for (int i = 0; i < c.n_causal; ++i) t->cx.push_back(take(P * c.causal_channels[i]));
  const int L = (int)h->layers.size();
  t->x.assign(L + 1, 0);
  t->x[0] = t->cx.back();
  for (int l = 1; l <= L; ++l) t->x[l] = take(P * h->R);
  t->tfsg.clear();
  t->z.clear();
  int gmax = 0;
  for (int l = 0; l < L; ++l) {
    t->tfsg.push_back(take(P * 2 * h->layers[l].G));
    t->z.push_back(take(P * h->layers[l].G));
    gmax = gmax > h->layers[l].G ? gmax : h->layers[l].G;
  }
  t->skip = take(P * h->S);
  t->hbuf.clear();
  int hmax = 0;
  for (int i = 0; i < c.n_softmax; ++i) hmax = hmax > c.softmax_channels[i] ? hmax : c.softmax_channels[i];
  for (int i = 1; i < c.n_softmax; ++i) t->hbuf.push_back(take(P * c.softmax_channels[i]));
  t->dlogits = take(P * h->Q);
  t->dh[0] = take(P * hmax);
  t->dh[1] = take(P * hmax);
  t->dout[0] = take(P * h->R);
  t->dout[1] = take(P * h->R);
  t->dz = take(P * gmax);
  t->dafg = take(P * 2 * gmax);
  t->dzs = take((int64_t)L * P * gmax);
  int cmax = 0;
  for (int i = 0; i < c.n_causal; ++i) cmax = cmax > c.causal_channels[i] ? cmax : c.causal_channels[i];
  if (c.n_causal > 1) {
    t->dcx[0] = take(P * cmax);
    t->dcx[1] = take(P * cmax);
  }
  t->loss_acc = take(16);
  t->ce_colsum = take(c.quantization_steps > 256 ? c.quantization_steps : 256);
  {
    int gm = 0;
    for (int l = 0; l < L; ++l) gm = gm > h->layers[l].G ? gm : h->layers[l].G;
    const int kk = c.residual_filter_width;
    t->tc_tab = take((int64_t)L * 8);
    t->tc_w1 = take((int64_t)L * 2 * gm * kk * h->R);
    t->tc_w2 = take((int64_t)L * h->R * gm);
    t->tc_ws = take((int64_t)L * h->S * gm);
    t->tc_w1t = take((int64_t)L * 2 * gm * kk * h->R);
    t->tc_wpt = take((int64_t)L * h->R * gm);
    t->tc_wst = take((int64_t)L * h->S * gm);
    t->tc_wh.clear();
    t->tc_wht.clear();
    for (int i = 0; i + 1 < c.n_softmax; ++i) {
      t->tc_wh.push_back(take((int64_t)c.softmax_channels[i] * c.softmax_channels[i + 1]));
      t->tc_wht.push_back(take((int64_t)c.softmax_channels[i] * c.softmax_channels[i + 1]));
    }
  }
  t->total = off;
}

extern "C" int64_t wn_workspace_bytes(const wn_handle* h, int B, int W) {
  if (!h || B < 1 || W < 1) return WN_EINVAL;
  Tape t;
  plan_tape(h, B, W, &t);
  return t.total * (int64_t)sizeof(float);
}

extern "C" int wn_bind_workspace(wn_handle* h, void* ws, int64_t bytes, int B, int W) {
  WN_REQUIRE(h && ws, WN_EINVAL, "null argument");
  WN_REQUIRE(B >= 1 && W >= 1, WN_EINVAL, "bad batch/width");
  WN_REQUIRE(((uintptr_t)ws & 255) == 0, WN_EINVAL, "workspace must be 256-byte aligned");
  Tape t;
  plan_tape(h, B, W, &t);
  WN_REQUIRE(bytes >= t.total * (int64_t)sizeof(float), WN_ENOMEM, "workspace too small: %lld < %lld", (long long)bytes,
             (long long)(t.total * sizeof(float)));
  h->ws = (float*)ws;
  h->ws_bytes = bytes;
  h->tape = t;
  h->phase = PH_NONE;
  h->tc_tab_uploaded = false;
  h->ce_colsum_valid = false;
  return WN_OK;
}

#define WS(off) (h->ws + (off))
#define PRM(off) (params + (off))
#define BIAS(cp) ((cp).b_off >= 0 ? params + (cp).b_off : nullptr)

static GemmArgs base_gemm(int64_t M, int rows) {
  GemmArgs g;
  memset(&g, 0, sizeof(g));
  g.M = M;
  g.rows_out = rows;
  g.rows_in = rows;
  g.in_off = 0;
  g.ntaps = 1;
  g.mask_rows_in = rows;
  return g;
}

// dilated / causal conv forward as a shifted GEMM: Y = conv(A) (+bias), zero prefix zp
static int conv_forward(const wn_handle* h, const float* params, const ConvParam& cp, const float* A, int lda,
                        int dilation, int zp, float* Y, int ldy, cudaStream_t s) {
  GemmArgs g = base_gemm(h->tape.P, h->tape.W);
  g.A = A;
  g.lda = lda;
  g.K = cp.in_ch;
  g.ntaps = cp.taps;
  for (int i = 0; i < cp.taps; ++i) g.shift[i] = (cp.taps - 1 - i) * dilation;
  g.Wt = PRM(cp.w_off);
  g.sn = (int64_t)cp.in_ch * cp.taps;
  g.sk = cp.taps;
  g.st = 1;
  g.N = cp.out_ch;
  g.bias = BIAS(cp);
  g.zp = zp;
  g.Y = Y;
  g.ldy = ldy;
  return simt_gemm(g, s);
}

extern "C" int wn_forward_causal_block(wn_handle* h, const float* params, const int32_t* x, float* out, wn_stream_t st) {
  WN_REQUIRE(h && params && x, WN_EINVAL, "null argument");
  WN_REQUIRE(h->ws, WN_ESTATE, "no workspace bound");
  cudaStream_t s = (cudaStream_t)st;
  const wn_config& c = h->cfg;
  const Tape& t = h->tape;
  const ConvParam& c0 = h->causal[0];
  WN_TRY(simt_embed_prepare(PRM(c0.w_off), WS(t.emb), c0.out_ch, h->Q, c.causal_filter_width, s));
  WN_TRY(simt_embed_forward(WS(t.emb), BIAS(c0), x, WS(t.cx[0]), t.B, t.W, c0.out_ch, h->Q, c.causal_filter_width, s));
  for (int i = 1; i < c.n_causal; ++i)  // no activation between causal convs (wavenet.py:567-569)
    WN_TRY(conv_forward(h, params, h->causal[i], WS(t.cx[i - 1]), h->causal[i].in_ch, 1, 0, WS(t.cx[i]),
                        h->causal[i].out_ch, s));
  if (out)
    WN_CHECK_CUDA(cudaMemcpyAsync(out, WS(t.x[0]), sizeof(float) * t.P * h->R, cudaMemcpyDeviceToDevice, s));
  h->x_idx = x;
  h->causal_from_idx = true;
  h->x0_split = false;
  h->phase = PH_CAUSAL;
  return WN_OK;
}

static int residual_forward_simt(wn_handle* h, const float* params, cudaStream_t s) {
  const Tape& t = h->tape;
  const int L = (int)h->layers.size();
  const int k = h->cfg.residual_filter_width;
  for (int l = 0; l < L; ++l) {
    const ResLayer& ly = h->layers[l];
    const int zp = wn_zero_prefix(t.W, ly.dilation, k);
    const int G = ly.G;
    // a_f | a_g  (wavenet.py:360 via DilatedConvolution1D.__call__ :294-342)
    WN_TRY(conv_forward(h, params, ly.wf, WS(t.x[l]), h->R, ly.dilation, zp, WS(t.tfsg[l]), 2 * G, s));
    WN_TRY(conv_forward(h, params, ly.wg, WS(t.x[l]), h->R, ly.dilation, zp, WS(t.tfsg[l]) + G, 2 * G, s));
    WN_TRY(simt_gate_forward(WS(t.tfsg[l]), WS(t.z[l]), t.P, G, s));
    // output = projection_block(z) + x   (wavenet.py:363,367)
    GemmArgs g = base_gemm(t.P, t.W);
    g.A = WS(t.z[l]);
    g.lda = G;
    g.K = G;
    g.Wt = PRM(ly.proj.w_off);
    g.sn = G;
    g.sk = 1;
    g.N = h->R;
    g.bias = BIAS(ly.proj);
    g.Rsd = WS(t.x[l]);
    g.ldr = h->R;
    g.Y = WS(t.x[l + 1]);
    g.ldy = h->R;
    WN_TRY(simt_gemm(g, s));
    // sum_skip_connections += projection_softmax(z)   (wavenet.py:364,579)
    g.Wt = PRM(ly.skip.w_off);
    g.N = h->S;
    g.bias = BIAS(ly.skip);
    g.Rsd = nullptr;
    g.accumulate = l > 0;
    g.Y = WS(t.skip);
    g.ldy = h->S;
    WN_TRY(simt_gemm(g, s));
  }
  return WN_OK;
}

extern "C" int wn_forward_residual_block(wn_handle* h, const float* params, const float* in, float* out,
                                         float* sum_skip, wn_stream_t st) {
  WN_REQUIRE(h && params, WN_EINVAL, "null argument");
  WN_REQUIRE(h->ws, WN_ESTATE, "no workspace bound");
  cudaStream_t s = (cudaStream_t)st;
  const Tape& t = h->tape;
  if (in) {
    WN_CHECK_CUDA(cudaMemcpyAsync(WS(t.x[0]), in, sizeof(float) * t.P * h->R, cudaMemcpyDeviceToDevice, s));
    h->causal_from_idx = false;
    h->x0_split = false;
  } else {
    WN_REQUIRE(h->phase >= PH_CAUSAL, WN_ESTATE, "forward_residual_block: no causal output on the tape");
  }
  const int L = (int)h->layers.size();
  if (h->prec == WN_PREC_F16X2 && tcs_supported(h)) {
    // split-fp16 tensor-core path: MMA operands on the tape are [hi | lo] fp16 rows of the same byte size
    if (!h->x0_split) WN_TRY(tcs_split_rows_inplace(WS(t.x[0]), h->R, t.P, tcs_act_scale(), s));
    h->x0_split = true;
    WN_TRY(tcs_forward_residual(h, params, s));
    h->tape_has_tfsg = false;
    h->tape_tc = true;
    h->tape_split = true;
    if (out) WN_TRY(tcs_unsplit_rows(WS(t.x[L]), out, h->R, t.P, 1.f / tcs_act_scale(), s));
    if (sum_skip) WN_TRY(tcs_unsplit_rows(WS(t.skip), sum_skip, h->S, t.P, 1.f / tcs_act_scale(), s));
    h->head_external = false;
    h->phase = PH_RESIDUAL;
    return WN_OK;
  }
  WN_REQUIRE(!h->x0_split, WN_ESTATE, "forward_residual_block: the causal output on the tape was converted by the fp16x2 path");
  h->tape_split = false;
  if (h->prec == WN_PREC_TF32 && tc_layer_supported(h)) {
    WN_TRY(tc_forward_residual(h, params, s));
    h->tape_has_tfsg = false;      // tensor-core tapes keep (z, sigmoid) only; the SIMT backward would recompute
    h->tape_tc = true;
  } else {
    WN_TRY(residual_forward_simt(h, params, s));
    h->tape_has_tfsg = true;
    h->tape_tc = false;
    h->skip_is_relu = false;
  }
  if (out) WN_CHECK_CUDA(cudaMemcpyAsync(out, WS(t.x[L]), sizeof(float) * t.P * h->R, cudaMemcpyDeviceToDevice, s));
  if (sum_skip)
    WN_CHECK_CUDA(cudaMemcpyAsync(sum_skip, WS(t.skip), sizeof(float) * t.P * h->S, cudaMemcpyDeviceToDevice, s));
  h->head_external = false;
  h->phase = PH_RESIDUAL;
  return WN_OK;
}

extern "C" int wn_forward_softmax_block(wn_handle* h, const float* params, const float* in, int T, int apply_softmax,
                                        float* out, wn_stream_t st) {
  WN_REQUIRE(h && params, WN_EINVAL, "null argument");
  WN_REQUIRE(h->ws, WN_ESTATE, "no workspace bound");
  cudaStream_t s = (cudaStream_t)st;
  const Tape& t = h->tape;
  WN_REQUIRE(T >= 1 && T <= t.W, WN_EINVAL, "softmax block width %d outside 1..%d", T, t.W);
  const int64_t rows = (int64_t)t.B * T;
  const bool split_head = h->prec == WN_PREC_F16X2 && tcs_supported(h) && (in || h->tape_split);
  WN_REQUIRE(in || split_head || !h->tape_split, WN_ESTATE, "forward_softmax_block: the tape was written by the fp16x2 path");
  if (in) {  // external (already sliced) input: kept compact at the start of the skip buffer
    if (split_head)
      WN_TRY(tcs_import_head_input(h, in, T, s));   // ReLU + hi/lo split on the way in
    else
      WN_CHECK_CUDA(cudaMemcpyAsync(WS(t.skip), in, sizeof(float) * rows * h->S, cudaMemcpyDeviceToDevice, s));
    h->head_external = true;
  } else {
    WN_REQUIRE(h->phase >= PH_RESIDUAL, WN_ESTATE, "forward_softmax_block: no skip sum on the tape");
    WN_REQUIRE(!h->head_external, WN_ESTATE, "skip buffer was overwritten by an external head input");
  }
  h->T = T;
  const int nh = (int)h->head.size();
  const bool tc_head = split_head || (h->prec == WN_PREC_TF32 && tc_head_supported(h) && h->S % 32 == 0);
  if (split_head)
    WN_TRY(tcs_forward_head(h, params, T, h->head_external, s));
  else if (tc_head)
    WN_TRY(tc_forward_head(h, params, T, h->head_external, s));
  h->head_tc = tc_head;
  h->head_split = split_head;
  for (int i = 0; i < nh && !tc_head; ++i) {  // ReLU -> 1x1 conv per head layer (wavenet.py:587-590)
    const ConvParam& cp = h->head[i];
    GemmArgs g = base_gemm(rows, T);
    if (i == 0) {
      g.A = WS(t.skip);
      g.lda = h->S;
      if (!h->head_external) {  // slice_1d to the last T columns, train.py:72-73
        g.rows_in = t.W;
        g.in_off = t.W - T;
      }
    } else {
      g.A = WS(t.hbuf[i - 1]);
      g.lda = cp.in_ch;
    }
    g.a_relu = 1;
    g.K = cp.in_ch;
    g.Wt = PRM(cp.w_off);
    g.sn = cp.in_ch;
    g.sk = 1;
    g.N = cp.out_ch;
    g.bias = BIAS(cp);
    g.Y = WS(t.hbuf[i]);
    g.ldy = cp.out_ch;
    WN_TRY(simt_gemm(g, s));
  }
  if (out) {
    if (apply_softmax)
      WN_TRY(simt_softmax_rows(WS(t.hbuf[nh - 1]), out, rows, h->Q, s));
    else
      WN_CHECK_CUDA(cudaMemcpyAsync(out, WS(t.hbuf[nh - 1]), sizeof(float) * rows * h->Q, cudaMemcpyDeviceToDevice, s));
  }
  h->phase = PH_HEAD;
  return WN_OK;
}

extern "C" int wn_cross_entropy(wn_handle* h, const int32_t* target, float* loss, wn_stream_t st) {
  WN_REQUIRE(h && target && loss, WN_EINVAL, "null argument");
  WN_REQUIRE(h->phase >= PH_HEAD, WN_ESTATE, "cross_entropy: no logits on the tape");
  cudaStream_t s = (cudaStream_t)st;
  const Tape& t = h->tape;
  const int64_t rows = (int64_t)t.B * h->T;
  // the split backward keeps every gradient tensor in fp16 planes: dlogits (<= 1/rows in magnitude) are scaled by a power of
  // two into [8, 16) so that the whole gradient stream sits in the fp16 normal range; weight-gradient reductions undo it
  float gscale = 0.f;
  if (h->head_split) {
    int k = 3;
    while (((int64_t)1 << (k - 3)) < rows) ++k;
    gscale = ldexpf(1.f, k);
    h->gscale = gscale;
  }
  bool split_written = false;
  h->dlogits_single = false;
  WN_TRY(simt_cross_entropy(WS(t.hbuf.back()), target, rows, h->Q, (double*)WS(t.loss_acc), loss, WS(t.dlogits),
                            WS(t.ce_colsum), &h->ce_colsum_valid, h->sm_count, s, gscale, &split_written));
  if (h->head_split && !split_written) WN_TRY(tcs_scale_split_dlogits(h, h->T, h->gscale, s));   // generic Q: convert in place
  h->phase = PH_LOSS;
  return WN_OK;
}

static WgradArgs base_wgrad(int64_t M, int rows) {
  WgradArgs g;
  memset(&g, 0, sizeof(g));
  g.M = M;
  g.dy_rows_out = g.dy_rows_in = rows;
  g.rows_out = g.rows_in = rows;
  g.ntaps = 1;
  return g;
}

// ---- causal stack (backward of wavenet.py:565-570) ----
static int causal_backward(wn_handle* h, const float* params, float* grads, const float* dout, cudaStream_t s) {
  if (!h->causal_from_idx) return WN_OK;
  const Tape& t = h->tape;
  const wn_config& c = h->cfg;
  const int W = t.W;
  const int64_t P = t.P;
  const float* dcur = dout;
  int ct = 0;
  for (int i = c.n_causal - 1; i >= 1; --i) {
    const ConvParam& cp = h->causal[i];
    WgradArgs wg = base_wgrad(P, W);
    wg.dY = dcur;
    wg.ldd = cp.out_ch;
    wg.N = cp.out_ch;
    wg.A = WS(t.cx[i - 1]);
    wg.lda = cp.in_ch;
    wg.K = cp.in_ch;
    wg.ntaps = cp.taps;
    for (int j = 0; j < cp.taps; ++j) wg.shift[j] = cp.taps - 1 - j;
    wg.dW = grads + cp.w_off;
    wg.sn = (int64_t)cp.in_ch * cp.taps;
    wg.sk = cp.taps;
    wg.st = 1;
    wg.dbias = cp.b_off >= 0 ? grads + cp.b_off : nullptr;
    WN_TRY(simt_wgrad(wg, h->sm_count, s));
    GemmArgs g = base_gemm(P, W);
    g.A = dcur;
    g.lda = cp.out_ch;
    g.K = cp.out_ch;
    g.ntaps = cp.taps;
    for (int j = 0; j < cp.taps; ++j) g.shift[j] = -(cp.taps - 1 - j);
    g.Wt = PRM(cp.w_off);
    g.sn = cp.taps;
    g.sk = (int64_t)cp.in_ch * cp.taps;
    g.st = 1;
    g.N = cp.in_ch;
    g.Y = WS(t.dcx[ct]);
    g.ldy = cp.in_ch;
    WN_TRY(simt_gemm(g, s));
    dcur = WS(t.dcx[ct]);
    ct ^= 1;
  }
  const ConvParam& c0 = h->causal[0];
  return simt_embed_backward(dcur, h->x_idx, WS(t.demb), grads + c0.w_off, c0.b_off >= 0 ? grads + c0.b_off : nullptr,
                             t.B, W, c0.out_ch, h->Q, c.causal_filter_width, s);
}

// ---- deterministic mode ---------------------------------------------------------------------------------------------------------
static __global__ void det_reduce_kernel(float* __restrict__ grads, const float* __restrict__ slab, int nslab, int64_t n) {
  for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += (int64_t)gridDim.x * blockDim.x * 4) {
    float4 a = *reinterpret_cast<const float4*>(grads + i);
    for (int c = 0; c < nslab; ++c) {             // fixed order: slab 0, 1, 2, ...
      const float4 v = *reinterpret_cast<const float4*>(slab + (int64_t)c * n + i);
      a.x += v.x, a.y += v.y, a.z += v.z, a.w += v.w;
    }
    *reinterpret_cast<float4*>(grads + i) = a;
  }
}

extern "C" int64_t wn_det_scratch_bytes(const wn_handle* h) {
  return h ? ((int64_t)h->sm_count * h->flat_size * (int64_t)sizeof(float) + 4096 * (int64_t)sizeof(double)) : WN_EINVAL;
}

extern "C" int wn_set_deterministic(wn_handle* h, int on, void* scratch) {
  WN_REQUIRE(h, WN_EINVAL, "null handle");
  if (!on) {
    h->deterministic = false;
    h->det_slab = nullptr;
    return WN_OK;
  }
  WN_REQUIRE(scratch && ((uintptr_t)scratch & 15) == 0, WN_EINVAL, "wn_set_deterministic: 16-byte aligned scratch of wn_det_scratch_bytes() needed");
  WN_REQUIRE(h->flat_size % 4 == 0, WN_EINVAL, "wn_set_deterministic: flat size must be a multiple of 4");
  WN_REQUIRE(tcs_supported(h) && tcs_det_supported(h), WN_EINVAL,
             "deterministic mode covers the fused fp16x2 shape (R = G = 64, k = 2, one bias-free causal layer)");
  h->deterministic = true;
  h->det_slab = (float*)scratch;
  h->det_nslab = h->sm_count;
  return WN_OK;
}

extern "C" int wn_accumulate_grads(wn_handle* h, const float* grads, float* acc, wn_stream_t st) {
  WN_REQUIRE(h && grads && acc, WN_EINVAL, "null argument");
  return simt_add_vec(grads, acc, (int)h->flat_size, (cudaStream_t)st);
}

extern "C" int wn_backward(wn_handle* h, const float* params, float* grads, wn_stream_t st) {
  WN_REQUIRE(h && params && grads, WN_EINVAL, "null argument");
  WN_REQUIRE(h->phase == PH_LOSS, WN_ESTATE, "backward: run forward and cross_entropy first");
  cudaStream_t s = (cudaStream_t)st;
  const Tape& t = h->tape;
  const wn_config& c = h->cfg;
  const int T = h->T, W = t.W;
  const int64_t P = t.P, rows = (int64_t)t.B * T;
  const int nh = (int)h->head.size();
  WN_CHECK_CUDA(cudaMemsetAsync(grads, 0, sizeof(float) * h->flat_size, s));
  if (h->deterministic) {
    WN_REQUIRE(h->head_split && h->tape_split && !h->head_external && h->prec == WN_PREC_F16X2, WN_ESTATE,
               "deterministic backward needs a tape written by the fp16x2 forward of the whole network");
    h->det_grads = grads;
    WN_CHECK_CUDA(cudaMemsetAsync(h->det_slab, 0, sizeof(float) * h->flat_size * h->det_nslab, s));
    WN_TRY(tcs_backward(h, params, grads, s));
    det_reduce_kernel<<<h->sm_count * 4, 256, 0, s>>>(grads, h->det_slab, h->det_nslab, h->flat_size);
    WN_CHECK_LAUNCH();
    return WN_OK;
  }
  if (h->head_split) {
    WN_REQUIRE(h->prec == WN_PREC_F16X2 && (h->tape_split || h->head_external), WN_ESTATE,
               "backward: the tape was written by the fp16x2 path; keep that precision selected");
    WN_TRY(tcs_backward(h, params, grads, s));
    if (h->head_external) return WN_OK;
    return causal_backward(h, params, grads, h->bwd_dout, s);
  }
  WN_REQUIRE(!h->tape_split, WN_ESTATE, "backward: the tape was written by the fp16x2 path; keep that precision selected");
  if (h->prec == WN_PREC_TF32 && h->tape_tc && h->head_tc && h->save_gates) {
    WN_TRY(tc_backward(h, params, grads, s));
    if (h->head_external) return WN_OK;
    return causal_backward(h, params, grads, h->bwd_dout, s);
  }

  // ---- head (backward of wavenet.py:584-593) ----
  const float* d = WS(t.dlogits);
  int tog = 0;
  for (int i = nh - 1; i >= 0; --i) {
    const ConvParam& cp = h->head[i];
    WgradArgs wg = base_wgrad(rows, T);
    wg.dY = d;
    wg.ldd = cp.out_ch;
    wg.N = cp.out_ch;
    if (i == 0) {
      wg.A = WS(t.skip);
      wg.lda = h->S;
      if (!h->head_external) {
        wg.rows_in = W;
        wg.in_off = W - T;
      }
    } else {
      wg.A = WS(t.hbuf[i - 1]);
      wg.lda = cp.in_ch;
    }
    wg.a_relu = 1;
    wg.K = cp.in_ch;
    wg.dW = grads + cp.w_off;
    wg.sn = cp.in_ch;
    wg.sk = 1;
    wg.dbias = cp.b_off >= 0 ? grads + cp.b_off : nullptr;
    WN_TRY(simt_wgrad(wg, h->sm_count, s));
    if (i == 0 && h->head_external) break;
    GemmArgs g = base_gemm(rows, T);
    g.A = d;
    g.lda = cp.out_ch;
    g.K = cp.out_ch;
    g.Wt = PRM(cp.w_off);
    g.sn = 1;
    g.sk = cp.in_ch;
    g.N = cp.in_ch;
    if (i == 0) {
      g.mask_src = WS(t.skip);
      g.ldm = h->S;
      g.mask_rows_in = W;
      g.mask_in_off = W - T;
    } else {
      g.mask_src = WS(t.hbuf[i - 1]);
      g.ldm = cp.in_ch;
    }
    g.Y = WS(t.dh[tog]);
    g.ldy = cp.in_ch;
    WN_TRY(simt_gemm(g, s));
    d = WS(t.dh[tog]);
    tog ^= 1;
  }
  if (h->head_external) return WN_OK;
  const float* dskip = d;  // [B*T][S], gradient w.r.t. the sliced skip sum

  // ---- residual layers (backward of wavenet.py:358-368, 572-582) ----
  const int L = (int)h->layers.size();
  const int k = c.residual_filter_width;
  int dt = 0;
  const float* dout = nullptr;  // gradient w.r.t. x[l+1]; the final output is unused (train.py:72) -> zero
  for (int l = L - 1; l >= 0; --l) {
    const ResLayer& ly = h->layers[l];
    const int G = ly.G;
    const int zp = wn_zero_prefix(W, ly.dilation, k);
    // dz = Wp^T dout + Ws^T dskip
    if (dout) {
      GemmArgs g = base_gemm(P, W);
      g.A = dout;
      g.lda = h->R;
      g.K = h->R;
      g.Wt = PRM(ly.proj.w_off);
      g.sn = 1;
      g.sk = G;
      g.N = G;
      g.Y = WS(t.dz);
      g.ldy = G;
      WN_TRY(simt_gemm(g, s));
      WgradArgs wg = base_wgrad(P, W);
      wg.dY = dout;
      wg.ldd = h->R;
      wg.N = h->R;
      wg.A = WS(t.z[l]);
      wg.lda = G;
      wg.K = G;
      wg.dW = grads + ly.proj.w_off;
      wg.sn = G;
      wg.sk = 1;
      wg.dbias = ly.proj.b_off >= 0 ? grads + ly.proj.b_off : nullptr;
      WN_TRY(simt_wgrad(wg, h->sm_count, s));
    }
    {
      GemmArgs g = base_gemm(P, W);
      g.A = dskip;
      g.lda = h->S;
      g.K = h->S;
      g.rows_in = T;
      g.in_off = -(W - T);
      g.Wt = PRM(ly.skip.w_off);
      g.sn = 1;
      g.sk = G;
      g.N = G;
      g.accumulate = dout != nullptr;
      g.Y = WS(t.dz);
      g.ldy = G;
      WN_TRY(simt_gemm(g, s));
      WgradArgs wg = base_wgrad(rows, T);
      wg.dY = dskip;
      wg.ldd = h->S;
      wg.N = h->S;
      wg.A = WS(t.z[l]);
      wg.lda = G;
      wg.K = G;
      wg.rows_in = W;
      wg.in_off = W - T;
      wg.dW = grads + ly.skip.w_off;
      wg.sn = G;
      wg.sk = 1;
      wg.dbias = ly.skip.b_off >= 0 ? grads + ly.skip.b_off : nullptr;
      WN_TRY(simt_wgrad(wg, h->sm_count, s));
    }
    if (!h->tape_has_tfsg) {  // tensor-core forward keeps only x and z: recompute tanh | sigmoid from x[l]
      WN_TRY(conv_forward(h, params, ly.wf, WS(t.x[l]), h->R, ly.dilation, zp, WS(t.tfsg[l]), 2 * G, s));
      WN_TRY(conv_forward(h, params, ly.wg, WS(t.x[l]), h->R, ly.dilation, zp, WS(t.tfsg[l]) + G, 2 * G, s));
      WN_TRY(simt_gate_forward(WS(t.tfsg[l]), WS(t.dafg), P, G, s));   // z recomputed into scratch (dafg)
    }
    WN_TRY(simt_gate_backward(WS(t.tfsg[l]), WS(t.dz), WS(t.dafg), P, W, G, zp, s));
    float* dnew = WS(t.dout[dt]);
    for (int part = 0; part < 2; ++part) {
      const ConvParam& cp = part == 0 ? ly.wf : ly.wg;
      const float* da = WS(t.dafg) + part * G;
      WgradArgs wg = base_wgrad(P, W);
      wg.dY = da;
      wg.ldd = 2 * G;
      wg.N = G;
      wg.A = WS(t.x[l]);
      wg.lda = h->R;
      wg.K = h->R;
      wg.ntaps = cp.taps;
      for (int i = 0; i < cp.taps; ++i) wg.shift[i] = (cp.taps - 1 - i) * ly.dilation;
      wg.dW = grads + cp.w_off;
      wg.sn = (int64_t)h->R * cp.taps;
      wg.sk = cp.taps;
      wg.st = 1;
      wg.dbias = cp.b_off >= 0 ? grads + cp.b_off : nullptr;
      WN_TRY(simt_wgrad(wg, h->sm_count, s));
      // dx[s] += sum_tap W_tap^T da[s + (k-1-tap)d]
      GemmArgs g = base_gemm(P, W);
      g.A = da;
      g.lda = 2 * G;
      g.K = G;
      g.ntaps = cp.taps;
      for (int i = 0; i < cp.taps; ++i) g.shift[i] = -(cp.taps - 1 - i) * ly.dilation;
      g.Wt = PRM(cp.w_off);
      g.sn = cp.taps;
      g.sk = (int64_t)h->R * cp.taps;
      g.st = 1;
      g.N = h->R;
      if (part == 0) {
        g.Rsd = dout;  // residual branch (wavenet.py:367); null for the last layer
        g.ldr = h->R;
      } else {
        g.accumulate = 1;
      }
      g.Y = dnew;
      g.ldy = h->R;
      WN_TRY(simt_gemm(g, s));
    }
    dout = dnew;
    dt ^= 1;
  }
  return causal_backward(h, params, grads, dout, s);
}

extern "C" int wn_forward_loss(wn_handle* h, const float* params, const int32_t* x, const int32_t* target, int T,
                               float* loss, float* logits_opt, wn_stream_t st) {
  WN_TRY(wn_forward_causal_block(h, params, x, nullptr, st));
  h->fuse_head_relu = 1;
  int rc = wn_forward_residual_block(h, params, nullptr, nullptr, nullptr, st);
  h->fuse_head_relu = 0;
  WN_TRY(rc);
  if (target && h->prec == WN_PREC_F16X2 && h->tape_split && h->Q == 256) {
    // fp16x2 path: the last head conv carries the softmax cross-entropy in its epilogue (wavenet.py:590 + 597-617 in one
    // kernel): the logits never reach HBM unless logits_opt asks for them, and no separate loss kernel runs
    WN_REQUIRE(loss, WN_EINVAL, "null loss pointer");
    const int64_t rows = (int64_t)h->tape.B * T;
    int k = 3;
    while (((int64_t)1 << (k - 3)) < rows) ++k;
    h->gscale = ldexpf(1.f, k);
    h->fuse_ce_target = target;
    h->fuse_ce_logits = logits_opt;
    h->ce_fused_done = false;
    rc = wn_forward_softmax_block(h, params, nullptr, T, 0, nullptr, st);
    h->fuse_ce_target = nullptr;
    h->fuse_ce_logits = nullptr;
    WN_TRY(rc);
    if (h->ce_fused_done) {
      WN_TRY(simt_loss_finalize((const double*)(h->ws + h->tape.loss_acc), rows, loss, (cudaStream_t)st));
      h->ce_colsum_valid = false;      // the bias gradient of the last conv comes from a column-sum pass over dlogits
      h->phase = PH_LOSS;
      return WN_OK;
    }
    if (logits_opt)
      WN_CHECK_CUDA(cudaMemcpyAsync(logits_opt, h->ws + h->tape.hbuf.back(), sizeof(float) * rows * h->Q, cudaMemcpyDeviceToDevice,
                                    (cudaStream_t)st));
    return wn_cross_entropy(h, target, loss, st);
  }
  WN_TRY(wn_forward_softmax_block(h, params, nullptr, T, 0, logits_opt, st));
  if (target) WN_TRY(wn_cross_entropy(h, target, loss, st));
  return WN_OK;
}

extern "C" int wn_tc_active(const wn_handle* h) {
  if (!h) return 0;
  if (h->prec == WN_PREC_F16X2) return tcs_supported(h) ? 1 : 0;
  return h->prec == WN_PREC_TF32 && tc_layer_supported(h) ? 1 : 0;
}

extern "C" int64_t wn_optim_scratch_bytes(const wn_handle* h) { return h ? 256 : WN_EINVAL; }

extern "C" int wn_clip_adam_step(wn_handle* h, float* params, float* grads, float* m, float* v, int t, float lr,
                                 float beta1, float beta2, float eps, float weight_decay, float clip, float grad_scale,
                                 void* scratch, float* norm_out, wn_stream_t st) {
  WN_REQUIRE(h && params && grads && m && v && scratch, WN_EINVAL, "null argument");
  WN_REQUIRE(t >= 1, WN_EINVAL, "Adam step count must be >= 1");
  double* det_partials = h->deterministic ? reinterpret_cast<double*>(h->det_slab + (int64_t)h->det_nslab * h->flat_size) : nullptr;
  return optim_clip_adam(params, grads, m, v, h->flat_size, t, lr, beta1, beta2, eps, weight_decay, clip, grad_scale,
                         (double*)scratch, norm_out, h->sm_count, (cudaStream_t)st, det_partials);
}

extern "C" int wn_onehot_to_index(const float* onehot, int B, int Q, int W, int32_t* idx, wn_stream_t s) {
  WN_REQUIRE(onehot && idx, WN_EINVAL, "null argument");
  return simt_onehot_to_index(onehot, B, Q, W, idx, (cudaStream_t)s);
}
