// SeLA self-labelling step - replaces the inner loop of SeLA.self_label_step
// (reference models/sela.py:146-166, one call per batch):
//     P      = pow(log_softmax(logits, -1), lambda)^T                     [K x B]      (:152)
//     repeat num_iters:  alpha = 1 / (P beta);  beta = 1 / (alpha^T P)^T               (:154-156)
//     labels = argmax_k  alpha_k P_kb beta_b                                            (:158-160)
// alpha [K] and beta [B] are STATE carried from batch to batch (sela.py:72-73), updated in place.
// Same alternating-scaling matvec pattern as the Sinkhorn kernels, but tiny (K = 128, B = 500 in configs/sela.yaml)
// and strictly sequential (2 * num_iters = 160 dependent matvecs): the reference issues ~500 eager launches per batch,
// here ONE persistent CTA runs the whole step with P kept in a 256 KB L2-resident scratch matrix.
#include "host_util.h"
#include "common.cuh"

using namespace ssvb;

namespace {

__global__ void __launch_bounds__(1024, 1)
sela_kernel(const float* __restrict__ logits, int64_t ld, int b, int k, float lambda, int num_iters,
            float* __restrict__ alpha, float* __restrict__ beta, float* __restrict__ P /* [k][b] */,
            int64_t* __restrict__ labels) {
  extern __shared__ float sh[];  // alpha_s [k] | beta_s [b]
  float* alpha_s = sh;
  float* beta_s = sh + k;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = blockDim.x >> 5;

  // ---- P = pow(log_softmax(logits), lambda)^T : one warp per sample row
  for (int r = warp; r < b; r += nwarp) {
    const float* x = logits + static_cast<int64_t>(r) * ld;
    float m = -INFINITY;
    for (int c = lane; c < k; c += 32) m = fmaxf(m, x[c]);
    m = warp_max(m);
    float z = 0.f;
    for (int c = lane; c < k; c += 32) z += expf(x[c] - m);
    z = warp_sum(z);
    const float lse = m + logf(z);
    for (int c = lane; c < k; c += 32) P[static_cast<int64_t>(c) * b + r] = powf(x[c] - lse, lambda);
  }
  for (int i = tid; i < k; i += blockDim.x) alpha_s[i] = alpha[i];
  for (int i = tid; i < b; i += blockDim.x) beta_s[i] = beta[i];
  __syncthreads();

  for (int it = 0; it < num_iters; ++it) {
    // alpha_k = 1 / sum_b P_kb beta_b : one warp per cluster row (coalesced along b)
    for (int c = warp; c < k; c += nwarp) {
      const float* row = P + static_cast<int64_t>(c) * b;
      float s = 0.f;
      for (int j = lane; j < b; j += 32) s = fmaf(row[j], beta_s[j], s);
      s = warp_sum(s);
      if (lane == 0) alpha_s[c] = 1.f / s;
    }
    __syncthreads();
    // beta_b = 1 / sum_k alpha_k P_kb : one thread per sample column (coalesced along b)
    for (int j = tid; j < b; j += blockDim.x) {
      float s = 0.f;
      for (int c = 0; c < k; ++c) s = fmaf(alpha_s[c], P[static_cast<int64_t>(c) * b + j], s);
      beta_s[j] = 1.f / s;
    }
    __syncthreads();
  }
  // ---- labels = argmax_k alpha_k P_kb beta_b (first maximum, like torch.argmax)
  for (int j = tid; j < b; j += blockDim.x) {
    const float bj = beta_s[j];
    float best = -INFINITY;
    int arg = 0;
    for (int c = 0; c < k; ++c) {
      const float v = alpha_s[c] * P[static_cast<int64_t>(c) * b + j] * bj;
      // torch.argmax semantics: the first maximum, and a NaN counts as the maximum (the reference's own configuration,
      // lambda = 25 with 80 iterations, drives alpha to 0 and beta to inf in fp32, so NaN scores do occur)
      const bool v_nan = v != v, best_nan = best != best;
      if (!best_nan && (v_nan || v > best)) { best = v; arg = c; }
    }
    labels[j] = arg;
    beta[j] = bj;
  }
  for (int i = tid; i < k; i += blockDim.x) alpha[i] = alpha_s[i];
}

}  // namespace

extern "C" {

size_t ssvb_sela_workspace_bytes(int64_t b, int64_t k) {
  return (b > 0 && k > 0) ? static_cast<size_t>(b) * k * sizeof(float) + 256 : 0;
}

int ssvb_sela_self_label(const float* logits, int64_t b, int64_t k, int64_t ld, float lambda, int64_t num_iters,
                         float* alpha, float* beta, int64_t* labels, void* workspace, size_t workspace_bytes,
                         void* stream) {
  SSVB_TRY(check_device_sm100());
  if (!logits || !alpha || !beta || !labels || !workspace || b <= 0 || k <= 0 || ld < k || num_iters < 0)
    return SSVB_ERR_INVALID;
  if (workspace_bytes < ssvb_sela_workspace_bytes(b, k)) return SSVB_ERR_WORKSPACE;
  if ((b + k) * sizeof(float) > 200 * 1024 || b > (1 << 24) || k > (1 << 20)) return SSVB_ERR_UNSUPPORTED;
  const int smem = static_cast<int>((b + k) * sizeof(float));
  if (smem > 48 * 1024)
    SSVB_CUDA(cudaFuncSetAttribute(sela_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  sela_kernel<<<1, 1024, smem, static_cast<cudaStream_t>(stream)>>>(
      logits, ld, static_cast<int>(b), static_cast<int>(k), lambda, static_cast<int>(num_iters), alpha, beta,
      static_cast<float*>(workspace), labels);
  SSVB_LAUNCH_CHECK();
  return SSVB_OK;
}

}  // extern "C"
