// MoCo InfoNCE loss, forward + backward — replaces MocoLoss.forward (reference utils/losses.py:56-72,
// call site models/moco.py:117).
//
//   l_a0 = qh_a . kh_a / tau          (positive: row dot, NOT the diagonal of a full N x N GEMM)
//   l_aj = qh_a . m_j / tau           (queue rows used as stored)
//   loss = mean_a [ LSE_{j=0..K} l_aj - l_a0 ]
//   d qh_a = [(p_a0 - 1) kh_a + sum_j p_aj m_j] / (N tau),  d kh_a = (p_a0 - 1) qh_a / (N tau)
// The N x K logits never reach HBM: the queue axis is split across CTAs (sim_fwd_kernel / sim_bwd_kernel
// in SIM_MOCO mode, online-max because stored queue rows are not guaranteed unit-norm).
#include "sim_host.cuh"

using namespace ssvb;

namespace {

struct MocoSaved {
  __nv_bfloat16* qhat;  // [npad x dpad]
  float *inv_q, *inv_k, *pos, *lse2;
  size_t bytes;
};
MocoSaved moco_saved(void* base, int64_t n, int64_t dpad) {
  Carver c(base);
  MocoSaved s;
  const int64_t npad = round_up(n, 128);
  s.qhat = c.take<__nv_bfloat16>(npad * dpad);
  s.inv_q = c.take<float>(npad);
  s.inv_k = c.take<float>(npad);
  s.pos = c.take<float>(npad);
  s.lse2 = c.take<float>(npad);
  s.bytes = c.used();
  return s;
}

struct MocoWs {
  __nv_bfloat16* queue_bf16;  // [k x dpad] (only when no shadow is supplied)
  float *part_m, *part_l, *block_sums, *dacc;
  unsigned int* counter;
  size_t bytes;
};
void moco_plan(SimParams& p, int64_t n, int64_t k, float c, int BN, int min_tiles) {
  p = SimParams{};
  p.nseg = 1;
  p.seg_rows = static_cast<int>(n);
  p.seg_start[0] = 0;
  p.bps = static_cast<int>(ceil_div(n, 128));
  p.row_blocks = p.bps;
  p.cols = static_cast<int>(k);
  p.c = c;
  p.shift = 0.f;
  plan_chunks(p, BN, min_tiles);
}
MocoWs moco_ws(void* base, int64_t n, int64_t k, int64_t dpad) {
  Carver c(base);
  MocoWs w;
  const int64_t npad = round_up(n, 128);
  SimParams p;
  moco_plan(p, n, k, 1.f, 256, 2);
  w.queue_bf16 = c.take<__nv_bfloat16>(k * dpad);
  w.part_m = c.take<float>(static_cast<size_t>(4 * p.nchunks) * npad);
  w.part_l = c.take<float>(static_cast<size_t>(4 * p.nchunks) * npad);
  w.block_sums = c.take<float>(ceil_div(npad, 8) + 8);
  w.counter = c.take<unsigned int>(4);
  w.dacc = c.take<float>(npad * dpad);
  w.bytes = c.used();
  return w;
}

// fp32 [k x d] -> bf16 [k x dpad] (zero padded), one warp per row
__global__ void queue_to_bf16_kernel(const float* __restrict__ q, int64_t k, int d, int64_t ld, int dpad,
                                     __nv_bfloat16* __restrict__ out) {
  const int64_t row = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= k) return;
  for (int c = lane * 4; c < dpad; c += 128) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c < d) v = __ldg(reinterpret_cast<const float4*>(q + row * ld + c));
    __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
    uint2 pk;
    pk.x = *reinterpret_cast<uint32_t*>(&lo);
    pk.y = *reinterpret_cast<uint32_t*>(&hi);
    *reinterpret_cast<uint2*>(out + row * dpad + c) = pk;
  }
}

__global__ void moco_grad_finish_kernel(const float* __restrict__ q, const float* __restrict__ kk, int64_t ldq,
                                        int64_t ldk, int n, int d, const float* __restrict__ dacc, int ld_dacc,
                                        const MocoSaved sv, int normalize, float c, float inv_n_tau,
                                        const float* __restrict__ grad_out, float* __restrict__ dq,
                                        float* __restrict__ dk, int64_t lddq, int64_t lddk) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= n) return;
  const float scale = inv_n_tau * __ldg(grad_out);
  const float p0m1 = exp2f(sv.pos[row] * c - sv.lse2[row]) - 1.f;
  const float iq = normalize ? sv.inv_q[row] : 1.f, ik = normalize ? sv.inv_k[row] : 1.f;
  const int k4 = lane * 4;
  float gq[4] = {0.f, 0.f, 0.f, 0.f}, gk[4] = {0.f, 0.f, 0.f, 0.f}, qh[4] = {0.f, 0.f, 0.f, 0.f},
        kh[4] = {0.f, 0.f, 0.f, 0.f};
  if (k4 < d) {
    const float4 acc = *reinterpret_cast<const float4*>(dacc + static_cast<int64_t>(row) * ld_dacc + k4);
    const float4 vq = *reinterpret_cast<const float4*>(q + static_cast<int64_t>(row) * ldq + k4);
    const float4 vk = *reinterpret_cast<const float4*>(kk + static_cast<int64_t>(row) * ldk + k4);
    qh[0] = vq.x * iq; qh[1] = vq.y * iq; qh[2] = vq.z * iq; qh[3] = vq.w * iq;
    kh[0] = vk.x * ik; kh[1] = vk.y * ik; kh[2] = vk.z * ik; kh[3] = vk.w * ik;
    const float a[4] = {acc.x, acc.y, acc.z, acc.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      gq[e] = (p0m1 * kh[e] + a[e]) * scale;
      gk[e] = p0m1 * qh[e] * scale;
    }
  }
  if (normalize) {
    float dq_dot = gq[0] * qh[0] + gq[1] * qh[1] + gq[2] * qh[2] + gq[3] * qh[3];
    float dk_dot = gk[0] * kh[0] + gk[1] * kh[1] + gk[2] * kh[2] + gk[3] * kh[3];
    dq_dot = warp_sum(dq_dot);
    dk_dot = warp_sum(dk_dot);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      gq[e] = (gq[e] - dq_dot * qh[e]) * iq;
      gk[e] = (gk[e] - dk_dot * kh[e]) * ik;
    }
  }
  if (k4 < d) {
    if (dq) *reinterpret_cast<float4*>(dq + static_cast<int64_t>(row) * lddq + k4) = make_float4(gq[0], gq[1], gq[2], gq[3]);
    if (dk) *reinterpret_cast<float4*>(dk + static_cast<int64_t>(row) * lddk + k4) = make_float4(gk[0], gk[1], gk[2], gk[3]);
  }
}

int check_rows(const void* p, int64_t ld) {
  if (!p) return SSVB_ERR_INVALID;
  if ((reinterpret_cast<uintptr_t>(p) & 15) || (ld & 3)) return SSVB_ERR_ALIGNMENT;
  return SSVB_OK;
}
int check_shape(int64_t n, int64_t k, int64_t d, float temperature) {
  if (n <= 0 || k <= 0 || d <= 0 || !(temperature > 0.f)) return SSVB_ERR_INVALID;
  if (d % 4) return SSVB_ERR_ALIGNMENT;
  if (d > 128 || k > (1 << 30) || n > (1 << 30)) return SSVB_ERR_UNSUPPORTED;
  return SSVB_OK;
}

int get_queue_bf16(const float* queue, const void* shadow, int64_t k, int64_t d, int64_t ld, int64_t dpad,
                   MocoWs& ws, cudaStream_t s, const __nv_bfloat16** out) {
  if (shadow) {
    if (reinterpret_cast<uintptr_t>(shadow) & 15) return SSVB_ERR_ALIGNMENT;
    *out = static_cast<const __nv_bfloat16*>(shadow);
    return SSVB_OK;
  }
  SSVB_TRY(check_rows(queue, ld));
  queue_to_bf16_kernel<<<static_cast<unsigned>(ceil_div(k, 8)), 256, 0, s>>>(queue, k, static_cast<int>(d), ld,
                                                                             static_cast<int>(dpad), ws.queue_bf16);
  SSVB_LAUNCH_CHECK();
  *out = ws.queue_bf16;
  return SSVB_OK;
}

}  // namespace

extern "C" {

size_t ssvb_moco_saved_bytes(int64_t n, int64_t k, int64_t d) {
  (void)k;
  if (n <= 0 || d <= 0) return 0;
  return moco_saved(nullptr, n, sim_dpad(d)).bytes;
}
size_t ssvb_moco_workspace_bytes(int64_t n, int64_t k, int64_t d) {
  if (n <= 0 || k <= 0 || d <= 0) return 0;
  return moco_ws(nullptr, n, k, sim_dpad(d)).bytes;
}

int ssvb_moco_fwd(const float* query, const float* keys, const float* queue, const void* queue_bf16, int64_t n,
                  int64_t k, int64_t d, int64_t ld_q, int64_t ld_k, int64_t ld_queue, int normalize,
                  float temperature, float* loss, void* saved, void* workspace, size_t workspace_bytes,
                  void* stream) {
  SSVB_TRY(check_device_sm100());
  SSVB_TRY(check_shape(n, k, d, temperature));
  SSVB_TRY(check_rows(query, ld_q));
  SSVB_TRY(check_rows(keys, ld_k));
  if (!loss || !saved || !workspace || (!queue && !queue_bf16)) return SSVB_ERR_INVALID;
  if (workspace_bytes < ssvb_moco_workspace_bytes(n, k, d)) return SSVB_ERR_WORKSPACE;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int64_t dpad = sim_dpad(d), npad = round_up(n, 128);
  MocoSaved sv = moco_saved(saved, n, dpad);
  MocoWs ws = moco_ws(workspace, n, k, dpad);
  const float c = SSVB_LOG2E / temperature;

  if (npad > n) SSVB_CUDA(cudaMemsetAsync(sv.qhat + n * dpad, 0, (npad - n) * dpad * sizeof(__nv_bfloat16), s));
  SSVB_CUDA(cudaMemsetAsync(ws.counter, 0, 16, s));
  pair_prep_kernel<<<static_cast<unsigned>(ceil_div(n, 8)), 256, 0, s>>>(
      query, keys, static_cast<int>(n), static_cast<int>(d), ld_q, ld_k, normalize, 0, sv.qhat, nullptr,
      static_cast<int>(dpad), sv.inv_q, sv.inv_k, sv.pos, nullptr);
  SSVB_LAUNCH_CHECK();
  const __nv_bfloat16* qb = nullptr;
  SSVB_TRY(get_queue_bf16(queue, queue_bf16, k, d, ld_queue, dpad, ws, s, &qb));

  SimParams p;
  moco_plan(p, n, k, c, 256, 2);
  p.part_m = ws.part_m;
  p.part_l = ws.part_l;
  p.part_stride = static_cast<int>(npad);
  SSVB_TRY(launch_sim_fwd(SIM_MOCO, sv.qhat, npad, qb, k, dpad, p, s));
  lse_finalize_wide_kernel<SIM_MOCO><<<static_cast<unsigned>(ceil_div(n, 8)), 256, 0, s>>>(
      ws.part_m, ws.part_l, 4 * p.nchunks, p.part_stride, static_cast<int>(n), sv.pos, c, sv.lse2, ws.block_sums,
      ws.counter, 1.f / static_cast<float>(n), loss);
  SSVB_LAUNCH_CHECK();
  return SSVB_OK;
}

int ssvb_moco_bwd(const float* query, const float* keys, const float* queue, const void* queue_bf16, int64_t n,
                  int64_t k, int64_t d, int64_t ld_q, int64_t ld_k, int64_t ld_queue, int normalize,
                  float temperature, const float* grad_out, const void* saved, float* dquery, float* dkeys,
                  int64_t ld_dq, int64_t ld_dk, void* workspace, size_t workspace_bytes, void* stream) {
  SSVB_TRY(check_device_sm100());
  SSVB_TRY(check_shape(n, k, d, temperature));
  SSVB_TRY(check_rows(query, ld_q));
  SSVB_TRY(check_rows(keys, ld_k));
  if (dquery) SSVB_TRY(check_rows(dquery, ld_dq));
  if (dkeys) SSVB_TRY(check_rows(dkeys, ld_dk));
  if (!grad_out || !saved || !workspace || (!queue && !queue_bf16) || (!dquery && !dkeys)) return SSVB_ERR_INVALID;
  if (workspace_bytes < ssvb_moco_workspace_bytes(n, k, d)) return SSVB_ERR_WORKSPACE;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int64_t dpad = sim_dpad(d), npad = round_up(n, 128);
  MocoSaved sv = moco_saved(const_cast<void*>(saved), n, dpad);
  MocoWs ws = moco_ws(workspace, n, k, dpad);
  const float c = SSVB_LOG2E / temperature;
  const __nv_bfloat16* qb = nullptr;
  SSVB_TRY(get_queue_bf16(queue, queue_bf16, k, d, ld_queue, dpad, ws, s, &qb));

  SimParams p;
  moco_plan(p, n, k, c, 128, 4);
  p.rowstat = sv.lse2;
  p.colstat = nullptr;
  p.dacc = ws.dacc;
  p.ld_dacc = static_cast<int>(dpad);
  p.use_atomic = p.nchunks > 1;
  if (p.use_atomic) SSVB_CUDA(cudaMemsetAsync(ws.dacc, 0, npad * dpad * sizeof(float), s));
  SSVB_TRY(launch_sim_bwd(SIM_MOCO, sv.qhat, npad, qb, k, dpad, p, s));
  moco_grad_finish_kernel<<<static_cast<unsigned>(ceil_div(n, 8)), 256, 0, s>>>(
      query, keys, ld_q, ld_k, static_cast<int>(n), static_cast<int>(d), ws.dacc, static_cast<int>(dpad), sv,
      normalize, c, 1.f / (static_cast<float>(n) * temperature), grad_out, dquery, dkeys, ld_dq, ld_dk);
  SSVB_LAUNCH_CHECK();
  return SSVB_OK;
}

}  // extern "C"
