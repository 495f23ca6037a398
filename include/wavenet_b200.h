/* wavenet_b200.h -- C ABI of libwavenet_b200.so (sm_100a).
 *
 * Drop-in GPU backend for the two hot paths of musyoku/wavenet:
 *   (1) the dilated causal convolution training stack  (reference wavenet.py)
 *   (2) the cached incremental generator               (reference faster_wavenet.py)
 *
 * The reference has no FFI of its own (it is pure Python over Chainer); the
 * functions below are what a ctypes binding placed under the reference's
 * WaveNet / FasterWaveNet objects calls.  Each entry point
 * cites the reference code it replaces as file:line relative to the reference
 * repository root.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host;
 *   - activations are channels-last fp32: [B][W][C] (the reference's
 *     (B, C, 1, W) with the last axis made outermost), audio samples are int32
 *     class indices [B][W] (the one-hot tensor of data.py:61-68 is never built);
 *   - parameters, gradients and Adam moments are ONE flat fp32 buffer each, laid
 *     out by wn_param_layout() under the reference's Chainer link names
 *     (wavenet.py:461-472) and in Chainer's (out, in, kh, kw) element order;
 *   - all work is enqueued on the caller's stream; nothing synchronises unless
 *     stated; the caller owns every buffer (torch allocates them);
 *   - every function returns 0 on success or a negative WN_E* code, and
 *     wn_last_error() returns a thread-local message.
 */
#ifndef WAVENET_B200_H_
#define WAVENET_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WN_OK 0
#define WN_EINVAL (-1)   /* bad argument / hyper-parameter (reference raises Exception) */
#define WN_ECUDA (-2)    /* CUDA runtime or driver error */
#define WN_ESTATE (-3)   /* call order violated (e.g. backward before forward) */
#define WN_ENOMEM (-4)   /* workspace too small */
#define WN_EARCH (-5)    /* device is not sm_100 */

#define WN_MAX_CAUSAL 8
#define WN_MAX_LAYERS 32
#define WN_MAX_HEAD 8
#define WN_NAME_LEN 64

/* Hyper-parameters: the fields of reference Params (wavenet.py:100-146) that
 * shape the network.  residual_channels has one entry per layer of ONE block;
 * dilation of layer i is residual_filter_width**i (wavenet.py:406-410,429). */
typedef struct wn_config {
  int32_t quantization_steps;
  int32_t n_causal;
  int32_t causal_channels[WN_MAX_CAUSAL];
  int32_t causal_filter_width;
  int32_t causal_no_bias;
  int32_t n_res_layers;
  int32_t residual_channels[WN_MAX_LAYERS];
  int32_t residual_num_blocks;
  int32_t residual_filter_width;
  int32_t residual_dilation_no_bias;
  int32_t residual_projection_no_bias;
  int32_t n_softmax;                     /* len(softmax_conv_channels) */
  int32_t softmax_channels[WN_MAX_HEAD];
  int32_t softmax_no_bias;
} wn_config;

typedef struct wn_param_desc {
  char name[WN_NAME_LEN];   /* "<link name>/W" or "<link name>/b", wavenet.py:461-472 */
  int64_t offset;           /* in floats, into the flat buffers */
  int64_t numel;
  int32_t ndim;
  int32_t shape[4];
} wn_param_desc;

typedef struct wn_handle wn_handle;
typedef struct wn_gen wn_gen;
typedef void* wn_stream_t;   /* cudaStream_t */

/* arithmetic of the GEMM-shaped kernels */
#define WN_PREC_FP32 0   /* SIMT FFMA, exact fp32: the parity path (1e-4 logits) */
#define WN_PREC_TF32 1   /* tcgen05 kind::tf32, fp32 accumulate in TMEM: the fast path (1e-2 logits) */
#define WN_PREC_F16X2 2  /* tcgen05 kind::f16 on split operands (v = hi + lo, two fp16 planes, three MMAs per product,
                          * fp32 accumulate in TMEM): fp32-grade tensor-core path (1e-4 logits, 1e-3 gradients);
                          * channel counts must be multiples of 64, otherwise the SIMT fp32 kernels run */

const char* wn_last_error(void);
int wn_version(void);
/* number of kernels this library has launched in this process (reset!=0: read and zero) */
int64_t wn_launch_count(int reset);
/* a CUDA-graph replay re-launches kernels this library recorded at capture time: the caller adds them here */
int64_t wn_launch_count_add(int64_t n);

/* WaveNet.__init__ / create_network (wavenet.py:371-455): validates the
 * hyper-parameters (Params.check, wavenet.py:167-173) and builds the layout. */
int wn_create(const wn_config* cfg, wn_handle** out);
int wn_destroy(wn_handle* h);
/* Re-reads the CURRENT CUDA device into the handle (SM count for the persistent grids) -- the counterpart of
 * cuda.get_device(n).use() + chain.to_gpu() (train_audio/model.py:57-59).  WN_EARCH unless the device is sm_100;
 * wn_create performs the same check when a device is visible. */
int wn_set_device_info(wn_handle* h);
int wn_set_precision(wn_handle* h, int prec);
int wn_get_precision(const wn_handle* h);
/* 1 when the residual stack of this network runs on the tcgen05 path under the current precision */
int wn_tc_active(const wn_handle* h);

/* Flat parameter layout (replaces chain.add_link registration, wavenet.py:461-472). */
int64_t wn_flat_size(const wn_handle* h);      /* floats, including alignment padding */
int64_t wn_param_elems(const wn_handle* h);    /* floats, excluding padding */
int wn_num_params(const wn_handle* h);
int wn_param_layout(const wn_handle* h, wn_param_desc* out, int max_out);
int wn_receptive_width(const wn_handle* h);    /* train_audio/train.py:36-38 */
int wn_input_width(const wn_handle* h);        /* train_audio/train.py:41-44 */
int wn_zero_prefix(int width, int dilation, int filter_width);  /* wavenet.py:304-340 (quirk Q1) */

/* ---- training path ------------------------------------------------------
 * The workspace is the tape: each phase stores what backward needs.  Phases
 * mirror the calls train_audio/train.py:66-80 makes on the WaveNet object. */
int64_t wn_workspace_bytes(const wn_handle* h, int B, int W);
int wn_bind_workspace(wn_handle* h, void* ws, int64_t bytes, int B, int W);

/* forward_causal_block (wavenet.py:565-570) on int32 samples x[B][W];
 * out (optional) receives [B][W][R]. */
int wn_forward_causal_block(wn_handle* h, const float* params, const int32_t* x, float* out, wn_stream_t s);
/* forward_residual_block (wavenet.py:572-582): in==NULL continues from the
 * causal phase; otherwise in[B][W][R] is copied in (no gradient reaches the
 * causal stack then).  out[B][W][R] / sum_skip[B][W][S] optional. */
int wn_forward_residual_block(wn_handle* h, const float* params, const float* in, float* out, float* sum_skip,
                              wn_stream_t s);
/* slice_1d to the last T columns (train.py:72-73) + forward_softmax_block
 * (wavenet.py:584-593).  in==NULL continues from the residual phase.
 * out[B][T][Q]: logits, or probabilities when apply_softmax. */
int wn_forward_softmax_block(wn_handle* h, const float* params, const float* in, int T, int apply_softmax,
                             float* out, wn_stream_t s);
/* cross_entropy (wavenet.py:597-617) of the logits of the softmax phase against
 * target[B][T]; loss is one device float.  Also stores dlogits for backward. */
int wn_cross_entropy(wn_handle* h, const int32_t* target, float* loss, wn_stream_t s);
/* loss.backward() of backprop (wavenet.py:515-519): writes every gradient into
 * grads (flat layout); parameters not reached get zeros (Chainer
 * reallocate_cleared_grads). */
int wn_backward(wn_handle* h, const float* params, float* grads, wn_stream_t s);
/* forward_one_step(apply_softmax=False)+cross_entropy in one call
 * (wavenet.py:556-563,597): all phases above back to back. */
int wn_forward_loss(wn_handle* h, const float* params, const int32_t* x, const int32_t* target, int T, float* loss,
                    float* logits_opt, wn_stream_t s);
/* Deterministic training (bit-reproducible gradients and weights run to run): every CTA of the weight-gradient kernels adds into
 * its own copy of the flat gradient buffer with plain stores, wn_backward sums the copies in a fixed order, bias column sums
 * and the optimiser's norm use fixed-order reductions, the embedding gradient runs as a tensor-core weight gradient.  Covers the
 * fused fp16x2 shape (R = G = 64, k = 2, one bias-free causal layer: BASELINE config C); scratch: wn_det_scratch_bytes() device
 * bytes owned by the caller for as long as the mode is on (~150 x the flat buffer).  The reported loss still sums per-warp
 * partials with double-precision atomics.  The reference (Chainer/cuDNN) makes no such promise. */
int64_t wn_det_scratch_bytes(const wn_handle* h);
int wn_set_deterministic(wn_handle* h, int on, void* scratch);
/* Gradient accumulation over micro-batches (a global batch larger than one tape, e.g. BASELINE config 5's 256 x 16000 on
 * one GPU): acc[i] += grads[i] over the flat layout.  wn_backward always starts from zeros (Chainer cleargrads,
 * wavenet.py:516), so the caller adds each micro-batch's gradient to its accumulator and hands the accumulator to
 * wn_clip_adam_step with grad_scale = 1 / (micro_batches * world). */
int wn_accumulate_grads(wn_handle* h, const float* grads, float* acc, wn_stream_t s);

/* optimizer.update hooks + Adam (wavenet.py:175-199,477-480 and Chainer-2
 * Adam selected at wavenet.py:83): [g += wd*p] -> global L2 norm -> clip ->
 * Adam with bias-corrected step; t is the 1-based update count.
 * scratch: >= wn_optim_scratch_bytes() device bytes.  norm_out (optional) gets
 * the pre-clip norm as one device float.  grad_scale multiplies the gradient
 * first (1/world_size after a sum all-reduce). */
int64_t wn_optim_scratch_bytes(const wn_handle* h);
int wn_clip_adam_step(wn_handle* h, float* params, float* grads, float* m, float* v, int t, float lr, float beta1,
                      float beta2, float eps, float weight_decay, float clip, float grad_scale, void* scratch,
                      float* norm_out, wn_stream_t s);

/* ---- data-parallel training (the reference is single-device) ---------------------------------------
 * The batch of train_audio/train.py:58-80 is sharded over one process per GPU; the only exchange step is ONE sum
 * all-reduce of the flat gradient buffer between wn_backward and wn_clip_adam_step(grad_scale = 1/world), so the clip
 * hook acts on the averaged gradient like wavenet.py:477-480.  NCCL (libnccl.so.2, resolved at run time) lives behind
 * these calls: rank 0 makes a 128-byte id (host memory), the caller ships it to the other ranks, every rank calls
 * wn_comm_init with the device it trains on current. */
int wn_comm_available(void);                                   /* 1 when libnccl.so.2 could be loaded */
int wn_comm_unique_id(char* id_out_host /* [128] */);
int wn_comm_init(wn_handle* h, const char* id_host /* [128] */, int rank, int world);
int wn_comm_world(const wn_handle* h);
int wn_allreduce_grads(wn_handle* h, float* grads, wn_stream_t s);   /* in-place sum over ranks; no-op without a communicator */
int wn_comm_destroy(wn_handle* h);
/* One-shot peer-memory all-reduce fused with the optimiser's first pass (opt-in: WN_FUSED_ALLREDUCE=1 in the environment when
 * wn_comm_init runs; <= 8 ranks of one node): every rank publishes its gradient in a CUDA-IPC buffer, ONE kernel sums all
 * buffers in rank order over NVLink peer loads, applies 1/N and the weight-decay hook and accumulates the squared norm; the
 * clip + Adam kernel follows.  Without the peer path the call is wn_allreduce_grads + wn_clip_adam_step.  grad_scale multiplies
 * the SUMMED gradient. */
int wn_comm_peer_enabled(const wn_handle* h);
int wn_allreduce_clip_adam_step(wn_handle* h, float* params, float* grads, float* m, float* v, int t, float lr, float beta1,
                                float beta2, float eps, float weight_decay, float clip, float grad_scale, void* scratch,
                                float* norm_out, wn_stream_t s);

/* Profiling hooks used by bench.py's roofline leg: ONE launch of the fused residual-layer kernel (layer l)
 * or of the skip-sum GEMM on the bound tape; need a preceding TF32 wn_forward_residual_block. */
int wn_tc_layer_forward(wn_handle* h, int layer, wn_stream_t s);
int wn_tc_skip_gemm(wn_handle* h, wn_stream_t s);
/* ONE launch of the fused gate-backward kernel of layer l (dz GEMM + gate derivative + dWp) on the buffers a
 * preceding TF32 wn_backward left on the tape; grads_scratch (flat-size floats) receives the dWp accumulation. */
int wn_tc_gate_backward_layer(wn_handle* h, int layer, float* grads_scratch, wn_stream_t s);

/* Same hooks for the fp16x2 path (need a preceding fp16x2 wn_forward_residual_block). */
int wn_tcs_layer_forward(wn_handle* h, int layer, wn_stream_t s);
int wn_tcs_skip_gemm(wn_handle* h, wn_stream_t s);
/* Developer hooks (tests/dev/check_tcs_kernels.py): the generic fp16x2 GEMM / weight-gradient kernels on caller-made
 * split tensors ([rows][hi(C) | lo(C)] fp16). */
int wn_tcs_debug_gemm(wn_handle* h, const void* a_split, int C, int rows_in, int num_seq, int ns, int row_off0, int row_off1,
                      int rows_out, const void* w_split, int N, const void* rsd_split, const void* mask_split, int relu,
                      void* y, int out_split, wn_stream_t s);
int wn_tcs_debug_wgrad(wn_handle* h, const void* dy_split, int M, const void* x_split, int C, int rows, int num_seq,
                       int x_row_off, float scale, float* dW, wn_stream_t s);

/* ---- incremental generation (faster_wavenet.py) ----------------------------
 * n_streams independent utterances (the reference hard-codes one, wavenet.py:286).
 * head_act: 0 = ReLU always; 1 = reference (ReLU on the priming call,
 * faster_wavenet.py:51-52, ELU afterwards, faster_wavenet.py:108). */
#define WN_GEN_GREEDY 0   /* np.argmax, lowest index wins ties */
#define WN_GEN_SAMPLE 1   /* categorical sampling from softmax (generate.py:39), on-device Gumbel-max */

int wn_gen_create(wn_handle* h, int n_streams, int head_act, wn_gen** out);
int wn_gen_destroy(wn_gen* g);
/* Streams the tensor-core generator (gen_kernel_v6, automatic beyond 2 x sm_count = 296 streams) keeps resident at once on the current
 * device with clusters of `cluster_size` (4 or 8) CTAs: co-resident clusters x 128.  More streams run as consecutive
 * launches over the same samples.  (B200: 1920 with 8-CTA clusters, 4224 with 4-CTA clusters.) */
int wn_gen_mma_capacity(int cluster_size);
int64_t wn_gen_state_bytes(const wn_gen* g);
int wn_gen_bind_state(wn_gen* g, void* state, int64_t bytes);
/* Priming call (faster_wavenet.py:13-47): full pass over window[n][Win],
 * Win = wn_input_width(); fills the dilation rings; probs_opt[n][Q] gets the
 * last-column softmax; needs a bound training workspace for (n, Win). */
int wn_gen_prime(wn_gen* g, const float* params, const int32_t* window, float* probs_opt, wn_stream_t s);
/* The same for the streams [stream0, stream0 + count) only: window[count][Win], probs_opt[count][Q], a training workspace
 * bound for (count, Win).  Thousands of streams are primed in slices (the full pass keeps ~130 MB of activations per
 * stream of BASELINE config C); the generator counts as primed once the slice ending at n_streams has run. */
int wn_gen_prime_part(wn_gen* g, const float* params, const int32_t* window, int stream0, int count, float* probs_opt,
                      wn_stream_t s);
/* One _forward_one_step (faster_wavenet.py:50-63) for new samples x_new[n]:
 * probs[n][Q] = last-column softmax (logits when !apply_softmax). */
int wn_gen_step(wn_gen* g, const float* params, const int32_t* x_new, int apply_softmax, float* probs,
                wn_stream_t s);
/* logits[n][Q] of the NEXT sample as the last priming call / step left them (before softmax): the last column of what
 * the reference returns from _forward_one_step(apply_softmax=False), faster_wavenet.py:50-63. */
int wn_gen_logits(wn_gen* g, float* logits, wn_stream_t s);
/* The whole generate.py:24-43 loop on device: draws the first sample from the
 * priming distribution already stored by wn_gen_prime, then n_steps-1
 * incremental steps, each feeding its own sample back.  out[n][n_steps]. */
int wn_gen_run(wn_gen* g, const float* params, int n_steps, int mode, uint64_t seed, int32_t* out, wn_stream_t s);

/* create_batch of train_audio/train.py:14-22 on a device-resident quantised signal: for every start index
 * (host-drawn with np.random.randint exactly like the reference, then uploaded: B ints) gathers
 *   x[n]   = signal[start : start + input_width + target_width]
 *   tgt[n] = signal[start + input_width + 1 : start + input_width + target_width + 1]. */
int wn_crop_batch(const int32_t* signal, int64_t signal_len, const int32_t* starts, int B, int input_width,
                  int target_width, int32_t* x, int32_t* tgt, wn_stream_t s);

/* ---- data.py helpers on device ----------------------------------------------
 * onehot_pixel_image inverse: argmax over Q of a (B,Q,1,W) one-hot -> int32. */
int wn_onehot_to_index(const float* onehot_bq1w, int B, int Q, int W, int32_t* idx, wn_stream_t s);
/* mu-law quantiser (data.py:19-23) / inverse (data.py:38-43,54) on normalised
 * fp64 signals. */
int wn_mulaw_encode(const double* signal, int64_t n, int quantization_steps, int32_t* q, wn_stream_t s);
int wn_mulaw_decode(const int32_t* q, int64_t n, int quantization_steps, double scale, double* out, wn_stream_t s);

#ifdef __cplusplus
}
#endif
#endif /* WAVENET_B200_H_ */
