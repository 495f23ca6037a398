// Fused optimizer.update(): [WeightDecay] -> GradientClipping -> Adam over the flat
// parameter buffer.  Replaces the per-parameter cuBLAS sdot + host sync of
// sum_sqnorm (wavenet.py:175-182), the per-parameter `grad *= rate`
// (wavenet.py:196-199) and Chainer-2 Adam's per-parameter elementwise kernel.
#include "wn_common.h"

namespace {

__global__ void sqnorm_kernel(const float* __restrict__ params, float* __restrict__ grads, int64_t n, float wd,
                              float grad_scale, double* __restrict__ acc, double* __restrict__ det_partials) {
  double s = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float g = grads[i] * grad_scale;
    if (wd > 0.f) g += wd * params[i];   // chainer.optimizer.WeightDecay hook, wavenet.py:477-478
    grads[i] = g;
    s += (double)g * (double)g;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  __shared__ double part[32];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += part[i];
    if (det_partials)
      det_partials[blockIdx.x] = t;      // deterministic mode: summed in block order by sqnorm_finalize_kernel
    else
      atomicAdd(acc, t);
  }
}

__global__ void sqnorm_finalize_kernel(const double* __restrict__ partials, int n, double* __restrict__ acc) {
  double t = 0.0;
  for (int i = 0; i < n; ++i) t += partials[i];
  acc[0] = t;
}

__global__ void clip_adam_kernel(float* __restrict__ params, float* __restrict__ grads, float* __restrict__ m,
                                 float* __restrict__ v, int64_t n, float step, float one_minus_b1,
                                 float one_minus_b2, float eps, float clip, const double* __restrict__ acc,
                                 float* __restrict__ norm_out) {
  const double norm = sqrt(acc[0]);
  float rate = 1.f;
  if (clip > 0.f && norm != 0.0) {       // GradientClipping.__call__, wavenet.py:190-199
    const double r = (double)clip / norm;
    if (r < 1.0) rate = (float)r;
  }
  if (norm_out && blockIdx.x == 0 && threadIdx.x == 0) *norm_out = (float)norm;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float g = grads[i] * rate;
    grads[i] = g;
    float mi = m[i], vi = v[i];
    mi += one_minus_b1 * (g - mi);
    vi += one_minus_b2 * (g * g - vi);
    m[i] = mi;
    v[i] = vi;
    params[i] -= step * mi / (sqrtf(vi) + eps);
  }
}

}  // namespace

int optim_clip_adam(float* params, float* grads, float* m, float* v, int64_t n, int t, float lr, float beta1,
                    float beta2, float eps, float wd, float clip, float grad_scale, double* scratch, float* norm_out,
                    int sm_count, cudaStream_t s, double* det_partials) {
  WN_CHECK_CUDA(cudaMemsetAsync(scratch, 0, sizeof(double), s));
  int blocks = (int)((n + 255) / 256);
  if (blocks > sm_count * 8) blocks = sm_count * 8;
  if (blocks < 1) blocks = 1;
  sqnorm_kernel<<<blocks, 256, 0, s>>>(params, grads, n, wd, grad_scale, scratch, det_partials);
  WN_CHECK_LAUNCH();
  return optim_adam_after_norm(params, grads, m, v, n, t, lr, beta1, beta2, eps, clip, scratch, norm_out, sm_count, s,
                               det_partials, blocks);
}

int optim_adam_after_norm(float* params, float* grads, float* m, float* v, int64_t n, int t, float lr, float beta1, float beta2,
                          float eps, float clip, double* scratch, float* norm_out, int sm_count, cudaStream_t s,
                          double* det_partials, int det_n) {
  int blocks = (int)((n + 255) / 256);
  if (blocks > sm_count * 8) blocks = sm_count * 8;
  if (blocks < 1) blocks = 1;
  if (det_partials) {
    sqnorm_finalize_kernel<<<1, 1, 0, s>>>(det_partials, det_n, scratch);
    WN_CHECK_LAUNCH();
  }
  const double fix1 = 1.0 - pow((double)beta1, (double)t);
  const double fix2 = 1.0 - pow((double)beta2, (double)t);
  const float step = (float)((double)lr * sqrt(fix2) / fix1);
  clip_adam_kernel<<<blocks, 256, 0, s>>>(params, grads, m, v, n, step, 1.f - beta1, 1.f - beta2, eps, clip, scratch,
                                          norm_out);
  WN_CHECK_LAUNCH();
  return WN_OK;
}
