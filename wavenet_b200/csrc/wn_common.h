// Internal definitions shared by the translation units of libwavenet_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>

#include "../../include/wavenet_b200.h"

void wn_set_error(const char* fmt, ...);
void wn_count_launch();

#define WN_CHECK_CUDA(expr)                                                               \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess) {                                                              \
      wn_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return WN_ECUDA;                                                                    \
    }                                                                                     \
  } while (0)

#define WN_CHECK_LAUNCH()                                                               \
  do {                                                                                  \
    wn_count_launch();                                                                  \
    cudaError_t _e = cudaGetLastError();                                                \
    if (_e != cudaSuccess) {                                                            \
      wn_set_error("%s:%d: kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(_e)); \
      return WN_ECUDA;                                                                  \
    }                                                                                   \
  } while (0)

#define WN_REQUIRE(cond, code, ...) \
  do {                              \
    if (!(cond)) {                  \
      wn_set_error(__VA_ARGS__);    \
      return (code);                \
    }                               \
  } while (0)

#define WN_TRY(expr)        \
  do {                      \
    int _r = (expr);        \
    if (_r != WN_OK) return _r; \
  } while (0)

// ---- parameter table ---------------------------------------------------------------
struct ConvParam {
  int64_t w_off = -1;  // offset (floats) of W (out, in, taps) -- Chainer (O,C,1,k)/(O,C,k,1) share this element order
  int64_t b_off = -1;  // offset of bias or -1
  int out_ch = 0, in_ch = 0, taps = 1;
};

struct ResLayer {
  ConvParam wf, wg, proj, skip;
  int dilation = 1;
  int G = 0;
};

// ---- training tape (offsets in floats into the bound workspace) ------------------
struct Tape {
  int B = 0, W = 0;
  int64_t P = 0;
  int64_t emb = 0, demb = 0;          // [kc][Q][R0] transposed first causal filter / its gradient
  std::vector<int64_t> cx;            // causal layer outputs [P][C_i]
  std::vector<int64_t> x;             // x[l]: input of residual layer l, x[L]: final output   [P][R]
  std::vector<int64_t> tfsg;          // [P][2G]  tanh | sigmoid
  std::vector<int64_t> z;             // [P][G]
  int64_t skip = 0;                   // [P][S]
  std::vector<int64_t> hbuf;          // head activations h[i] (i>=1) [B*T][c_i]; logits = last
  int64_t dlogits = 0;                // [B*T][Q]
  int64_t dh[2] = {0, 0};             // [B*T][max head width]
  int64_t dout[2] = {0, 0};           // [P][R]
  int64_t dz = 0;                     // [P][Gmax]
  int64_t dafg = 0;                   // [P][2Gmax]
  int64_t dzs = 0;                    // [L][P][Gmax]  dskip . Ws_l for every layer (tensor-core backward)
  int64_t dcx[2] = {0, 0};            // [P][max causal width] (only when n_causal > 1)
  int64_t loss_acc = 0;               // 2 doubles
  int64_t ce_colsum = 0;              // [Q] column sums of dlogits (bias gradient of the last head conv), by the CE kernel
  // tensor-core path: TF32-rounded, K-major weight copies (rebuilt every forward)
  int64_t tc_tab = 0, tc_w1 = 0, tc_w2 = 0, tc_ws = 0, tc_w1t = 0, tc_wpt = 0, tc_wst = 0;
  std::vector<int64_t> tc_wh, tc_wht;
  int64_t total = 0;
};

enum Phase { PH_NONE = 0, PH_CAUSAL = 1, PH_RESIDUAL = 2, PH_HEAD = 3, PH_LOSS = 4 };

struct wn_handle {
  wn_config cfg;
  int prec = WN_PREC_FP32;
  std::vector<wn_param_desc> descs;
  int64_t flat_size = 0, param_elems = 0;
  std::vector<ConvParam> causal;
  std::vector<ResLayer> layers;  // residual_num_blocks * n_res_layers
  std::vector<ConvParam> head;
  int R = 0, S = 0, Q = 0;
  // bound workspace
  float* ws = nullptr;
  int64_t ws_bytes = 0;
  Tape tape;
  int phase = PH_NONE;
  bool causal_from_idx = false;   // gradient can reach the causal stack
  bool residual_external = false;
  bool head_external = false;
  const int32_t* x_idx = nullptr;  // device pointer given to the causal phase (must stay alive until backward)
  int T = 0;                       // columns seen by head/loss
  int sm_count = 148;
  bool tc_tab_uploaded = false;
  bool tape_has_tfsg = false;      // tanh | sigmoid of every layer are on the tape
  bool head_tc = false;            // head activations on the tape came from the tensor-core head
  bool tape_gates_zs = false;      // tensor-core tape stores (z, sigmoid) [fused kernel] instead of (tanh | sigmoid)
  bool tape_tc = false;            // tape written by the tensor-core forward (stored head activations are post-ReLU)
  int fuse_head_relu = 0;          // set by wn_forward_loss: nobody reads the raw skip sum, ReLU it in the skip GEMM
  bool skip_is_relu = false;
  bool save_gates = true;          // tensor-core forward also stores tanh | sigmoid (needed by backward)
  const float* bwd_dout = nullptr; // gradient w.r.t. the causal output after the residual backward
  // split-fp16 path (wn_tcs.cu): MMA operands on the tape are [hi | lo] fp16 rows
  bool tape_split = false;         // x / z / skip on the tape are split rows (written by tcs_forward_residual)
  bool head_split = false;         // head activations + dlogits are split rows (tcs_forward_head)
  bool x0_split = false;           // x[0] has already been converted in place
  float gscale = 1.f;              // power-of-two scale of every gradient tensor of the split backward
  const int32_t* fuse_ce_target = nullptr;   // set by wn_forward_loss: the head's last conv runs the cross-entropy epilogue
  float* fuse_ce_logits = nullptr;           //   optional fp32 logits output of that fused kernel
  bool ce_fused_done = false;
  bool dlogits_single = false;     // dlogits on the tape are ONE fp16 plane (fused CE epilogue) instead of split rows
  // deterministic mode (wn_set_deterministic): per-CTA gradient slabs, summed in slab order by wn_backward
  bool deterministic = false;
  float* det_slab = nullptr;       // [det_nslab][flat_size] floats + [4096] doubles of reduction partials, caller-owned
  int det_nslab = 0;
  float* det_grads = nullptr;      // the gradient buffer of the backward pass in flight (slab pointers are rebased from it)
  // data-parallel communicator (wn_comm.cu): an ncclComm_t owned by the handle
  void* comm = nullptr;
  int comm_rank = 0, comm_world = 1;
  void* peer = nullptr;            // PeerState* (wn_comm.cu): CUDA-IPC exchange buffers of the fused one-shot all-reduce
  bool ce_colsum_valid = false;    // tape.ce_colsum matches tape.dlogits (set by wn_cross_entropy)
};

inline float* wn_det_ptr(const wn_handle* h, float* p) { return h->deterministic ? h->det_slab + (p - h->det_grads) : p; }
inline int64_t wn_det_stride(const wn_handle* h) { return h->deterministic ? h->flat_size : 0; }

// ---- SIMT fp32 kernels (wn_simt.cu) ---------------------------------------------
struct GemmArgs {
  const float* A;
  int lda;
  int K;
  int ntaps;
  int shift[4];       // A row = in_row - shift[tap] (negative shift = look ahead)
  int64_t M;          // output rows
  int rows_out;       // output rows per sequence
  int rows_in;        // A rows per sequence
  int in_off;         // t_in = t_out + in_off
  int a_relu;         // apply max(.,0) to A elements
  const float* Wt;    // w(n, tap, k) = Wt[n*sn + k*sk + tap*st]
  int64_t sn, sk, st;
  int N;
  const float* bias;  // [N] or null
  int zp;             // rows with t_out < zp are forced to 0 (after bias)
  const float* Rsd;   // residual added after the mask, or null
  int ldr;
  const float* mask_src;  // multiply by (mask_src[mrow][n] > 0), or null;
  int ldm;                //   mrow = seq*mask_rows_in + t_out + mask_in_off
  int mask_rows_in;
  int mask_in_off;
  int accumulate;     // Y += result
  float* Y;
  int ldy;
};
int simt_gemm(const GemmArgs& g, cudaStream_t s);

struct WgradArgs {
  const float* dY;    // [M][ldd], N columns used
  int ldd;
  int N;
  int dy_rows_out;    // dY row mapping (same scheme as A): rows of the iteration space per sequence
  int dy_rows_in;
  int dy_in_off;
  const float* A;
  int lda;
  int K;
  int ntaps;
  int shift[4];
  int64_t M;          // iteration rows
  int rows_out, rows_in, in_off;   // A row mapping
  int a_relu;
  float* dW;          // dW(n, tap, k) += ... at dW[n*sn + k*sk + tap*st]
  int64_t sn, sk, st;
  float* dbias;       // [N] += column sums of dY, or null
};
int simt_wgrad(const WgradArgs& g, int sm_count, cudaStream_t s);

int simt_embed_prepare(const float* Wc, float* emb, int R, int Q, int kc, cudaStream_t s);
int simt_embed_forward(const float* emb, const float* bias, const int32_t* idx, float* out, int B, int W, int R, int Q,
                       int kc, cudaStream_t s);
int simt_embed_backward(const float* dout, const int32_t* idx, float* demb, float* dWc, float* dbias, int B, int W,
                        int R, int Q, int kc, cudaStream_t s);
int simt_gate_forward(float* afg, float* z, int64_t P, int G, cudaStream_t s);
int simt_gate_backward(const float* tfsg, const float* dz, float* dafg, int64_t P, int W, int G, int zp,
                       cudaStream_t s);
int simt_softmax_rows(const float* in, float* out, int64_t rows, int Q, cudaStream_t s);
int simt_cross_entropy(const float* logits, const int32_t* target, int64_t rows, int Q, double* acc, float* loss,
                       float* dlogits, float* colsum, bool* colsum_written, int sm_count, cudaStream_t s,
                       float split_scale = 0.f, bool* split_written = nullptr);
int simt_add_vec(const float* src, float* dst, int n, cudaStream_t s);
int simt_loss_finalize(const double* acc, int64_t rows, float* loss, cudaStream_t s);
int simt_onehot_to_index(const float* onehot, int B, int Q, int W, int32_t* idx, cudaStream_t s);

// ---- optimiser (wn_optim.cu) ---------------------------------------------------
int optim_clip_adam(float* params, float* grads, float* m, float* v, int64_t n, int t, float lr, float beta1,
                    float beta2, float eps, float wd, float clip, float grad_scale, double* scratch, float* norm_out,
                    int sm_count, cudaStream_t s, double* det_partials = nullptr);
// the second half of optim_clip_adam for callers that produced the squared norm themselves (scratch[0], or det_partials[0..n))
int optim_adam_after_norm(float* params, float* grads, float* m, float* v, int64_t n, int t, float lr, float beta1, float beta2,
                          float eps, float clip, double* scratch, float* norm_out, int sm_count, cudaStream_t s,
                          double* det_partials, int det_n);

// ---- tcgen05 TF32 kernels (wn_tc.cu) ---------------------------------------------
bool tc_layer_supported(const wn_handle* h);
bool tc_fused_supported(const wn_handle* h);
bool tc_head_supported(const wn_handle* h);
int tc_forward_residual(wn_handle* h, const float* params, cudaStream_t s);
int tc_forward_head(wn_handle* h, const float* params, int T, bool external, cudaStream_t s);
int tc_backward(wn_handle* h, const float* params, float* grads, cudaStream_t s);
int simt_colsum(const float* a, int64_t rows, int C, float* out, cudaStream_t s);
int simt_gate_backward_zs(const float* z, const float* sg, int sg_half, const float* dz, float* dafg, int64_t P, int W, int G, int zp,
                          cudaStream_t s);

// ---- tcgen05 split-fp16 kernels (wn_tcs.cu): fp32-grade arithmetic, three kind::f16 MMAs per product ----------
bool tcs_supported(const wn_handle* h);
float tcs_act_scale();   // power-of-two scale carried by every stored activation of the split tape
int tcs_split_rows_inplace(float* x, int C, int64_t rows, float scale, cudaStream_t s);
int tcs_unsplit_rows(const float* src_split, float* dst, int C, int64_t rows, float scale, cudaStream_t s);
int tcs_import_head_input(wn_handle* h, const float* in, int T, cudaStream_t s);
int tcs_forward_residual(wn_handle* h, const float* params, cudaStream_t s);
int tcs_forward_head(wn_handle* h, const float* params, int T, bool external, cudaStream_t s);
int tcs_scale_split_dlogits(wn_handle* h, int T, float gscale, cudaStream_t s);
int tcs_backward(wn_handle* h, const float* params, float* grads, cudaStream_t s);
bool tcs_det_supported(const wn_handle* h);   // deterministic mode covers the fused fp16x2 shape with a bias-free single causal layer
