// NT-Xent (SimCLR) loss, forward + backward, single GPU and row-sharded multi-GPU pieces.
// Replaces SimclrLoss.forward (reference utils/losses.py:15-46) and the contrastive term of
// RelicLoss.forward (utils/losses.py:163-194).
//
// Math: Z = [zi; zj] (M = 2N rows), s_ab = zh_a.zh_b / tau, partner(a) = a +- N,
//   loss = 1/M sum_a [ LSE_{b != a} s_ab - s_{a,partner(a)} ]
//   d zh_a = [ sum_{b != a} (P_ab + P_ba) zh_b - 2 zh_partner(a) ] / (M tau),  P_ab = exp(s_ab - lse_a)
//   d z_a  = (d zh_a - (d zh_a . zh_a) zh_a) / max(||z_a||, 1e-12)              (normalize=True)
// All exponentials run in the log2 domain: t = s * log2(e)/tau.
#include "sim_host.cuh"

using namespace ssvb;

namespace {

struct NtxPlan {
  int64_t n_glob, m, mpad, d, dpad;
  int mode;
  int normalize;
  float c, shift;   // what the kernels apply to the accumulator: x = s * c - shift
  float prescale;   // what pair_prep multiplies the staged rows with
  float wscale;     // FIXED mode: backward weights carry this power of two (fp16 normal range), undone in grad_finish
  float pos_c;      // log2-domain positive logit = pos * pos_c
};

int make_plan(NtxPlan& pl, int64_t n_glob, int64_t d, int normalize, float temperature) {
  if (n_glob <= 0 || d <= 0 || !(temperature > 0.f)) return SSVB_ERR_INVALID;
  if (d % 4) return SSVB_ERR_ALIGNMENT;
  if (d > 128) return SSVB_ERR_UNSUPPORTED;
  if (2 * n_glob > (1 << 30)) return SSVB_ERR_UNSUPPORTED;
  pl.n_glob = n_glob;
  pl.m = 2 * n_glob;
  pl.mpad = sim_mpad(pl.m);
  pl.d = d;
  pl.dpad = sim_dpad(d);
  pl.normalize = normalize ? 1 : 0;
  const float c = SSVB_LOG2E / temperature;
  // unit-norm rows bound |s| <= 1/tau, i.e. the log2-domain logits lie in [-c, c]: for c <= 32 (tau >= ~0.045)
  // exp2 needs neither a running max nor a shift (sums stay below 2^(32+30)); otherwise (or for raw inputs) online max.
  pl.mode = (normalize && 2.f * c <= 64.f) ? SIM_NTX_FIXED : SIM_NTX_ONLINE;
  if (pl.mode == SIM_NTX_FIXED) {
    // the staged fp16 rows are pre-scaled by sqrt(c): the tensor-core accumulator IS the log2-domain logit and the
    // exp loops carry no scale/shift instruction.
    pl.prescale = sqrtf(c);
    pl.c = 1.f;
    pl.shift = 0.f;
    pl.pos_c = 1.f;
    // W_ab = P_ab + P_ba is ~2/M: below fp16's normal range for M > 2^15.  Carry 2^k with k = floor(log2 M) - 4
    // (typical W -> 2^-3, largest possible W = 2 -> 2^(k+1) <= 32768 < 65504).
    int k = 0;
    while ((int64_t{2} << k) <= pl.m) ++k;  // k = floor(log2 M)
    k -= 4;
    if (k < 0) k = 0;
    if (k > 14) k = 14;
    pl.wscale = static_cast<float>(1 << k);
  } else {
    pl.prescale = 1.f;
    pl.c = c;
    pl.shift = c;
    pl.pos_c = c;
    pl.wscale = 1.f;
  }
  return SSVB_OK;
}

struct SavedLayout {  // single-GPU saved blob
  __nv_bfloat16* zhat;
  float* inv_norm;
  float* stat;
  size_t bytes;
};
SavedLayout saved_layout(void* base, int64_t mpad, int64_t dpad) {
  Carver c(base);
  SavedLayout s;
  s.zhat = c.take<__nv_bfloat16>(mpad * dpad);
  s.inv_norm = c.take<float>(mpad);
  s.stat = c.take<float>(mpad);
  s.bytes = c.used();
  return s;
}

struct WsLayout {
  float* pos;
  float* part_m;
  float* part_l;
  float* block_sums;
  unsigned int* counter;
  float* dacc;
  size_t bytes;
};
// `cols` = 2 * n_global: the partial buffers are sized from the same chunk plan the launch uses, and are
// at least mpad floats so rows_bwd can stage the per-column statistics in part_m.
WsLayout ws_layout(void* base, int64_t local_rows, int64_t row_blocks, int64_t cols, int64_t dpad) {
  Carver c(base);
  WsLayout w;
  const int64_t lr = round_up(local_rows, 256);
  SimParams p{};
  p.row_blocks = static_cast<int>(row_blocks);
  p.cols = static_cast<int>(cols);
  plan_chunks(p, 256, 4);
  size_t part_elems = static_cast<size_t>(4 * p.nchunks + 2) * lr;
  if (part_elems < static_cast<size_t>(sim_mpad(cols))) part_elems = sim_mpad(cols);
  w.pos = c.take<float>(lr);
  w.part_m = c.take<float>(part_elems);
  w.part_l = c.take<float>(part_elems);
  w.block_sums = c.take<float>(ceil_div(lr, 256) + 8);
  w.counter = c.take<unsigned int>(4);
  w.dacc = c.take<float>(lr * dpad);
  w.bytes = c.used();
  return w;
}

// ---- gradient finish: partner term, 1/(M tau) * grad_out, normalise-backward --------------------------
__global__ void ntx_grad_finish_kernel(const float* __restrict__ zi, const float* __restrict__ zj, int64_t ldi,
                                       int64_t ldj, int n_view, int d, int row0,
                                       const float* __restrict__ dacc, int ld_dacc,
                                       const __nv_bfloat16* __restrict__ zhat, int dpad,
                                       const float* __restrict__ inv_norm, int normalize, float inv_m_tau,
                                       const float* __restrict__ grad_out, float* __restrict__ dzi,
                                       float* __restrict__ dzj, int64_t ld_dzi, int64_t ld_dzj,
                                       float acc_scale, float zp_scale) {
  // acc_scale / zp_scale undo the staging scales: dacc = wscale * prescale * sum_b W_ab zh_b, staged rows = prescale * zh
  const int lrow = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (lrow >= 2 * n_view) return;
  const int view = lrow >= n_view;
  const int r = lrow - view * n_view;
  const int a_glob = row0 + lrow;  // rank-major layout: [rank][view][row]
  const int partner = view ? a_glob - n_view : a_glob + n_view;
  const float* z = view ? zj + static_cast<int64_t>(r) * ldj : zi + static_cast<int64_t>(r) * ldi;
  float* out = view ? dzj + static_cast<int64_t>(r) * ld_dzj : dzi + static_cast<int64_t>(r) * ld_dzi;
  const float scale = inv_m_tau * __ldg(grad_out);
  const float inv = normalize ? inv_norm[lrow] : 1.f;
  const int k = lane * 4;
  float g[4] = {0.f, 0.f, 0.f, 0.f}, zh[4] = {0.f, 0.f, 0.f, 0.f};
  if (k < d) {
    const float4 acc = *reinterpret_cast<const float4*>(dacc + static_cast<int64_t>(lrow) * ld_dacc + k);
    const uint2 pz = *reinterpret_cast<const uint2*>(zhat + static_cast<int64_t>(partner) * dpad + k);
    const float2 p01 = unpack_h2(pz.x, normalize != 0), p23 = unpack_h2(pz.y, normalize != 0);  // fp16 iff normalised
    const float4 zz = *reinterpret_cast<const float4*>(z + k);
    const float za = -2.f * zp_scale;
    g[0] = fmaf(acc.x, acc_scale, za * p01.x) * scale;
    g[1] = fmaf(acc.y, acc_scale, za * p01.y) * scale;
    g[2] = fmaf(acc.z, acc_scale, za * p23.x) * scale;
    g[3] = fmaf(acc.w, acc_scale, za * p23.y) * scale;
    zh[0] = zz.x * inv; zh[1] = zz.y * inv; zh[2] = zz.z * inv; zh[3] = zz.w * inv;
  }
  if (normalize) {
    float dot = g[0] * zh[0] + g[1] * zh[1] + g[2] * zh[2] + g[3] * zh[3];
    dot = warp_sum(dot);
#pragma unroll
    for (int i = 0; i < 4; ++i) g[i] = (g[i] - dot * zh[i]) * inv;
  }
  if (k < d) *reinterpret_cast<float4*>(out + k) = make_float4(g[0], g[1], g[2], g[3]);
}

int check_rows(const void* p, int64_t ld) {
  if (!p) return SSVB_ERR_INVALID;
  if ((reinterpret_cast<uintptr_t>(p) & 15) || (ld & 3)) return SSVB_ERR_ALIGNMENT;
  return SSVB_OK;
}

void fill_sim_params_rows(SimParams& p, const NtxPlan& pl, int nseg, int64_t seg_rows, int64_t s0, int64_t s1) {
  p = SimParams{};
  p.opf16 = pl.normalize ? 1 : 0;  // fp16 staging for unit-norm rows (3 more mantissa bits than bf16), bf16 otherwise
  p.nseg = nseg;
  p.seg_rows = static_cast<int>(seg_rows);
  p.seg_start[0] = static_cast<int>(s0);
  p.seg_start[1] = static_cast<int>(s1);
  p.bps = static_cast<int>(ceil_div(seg_rows, 128));
  p.row_blocks = nseg * p.bps;
  p.cols = static_cast<int>(pl.m);
  p.c = pl.c;
  p.shift = pl.shift;
}

}  // namespace

extern "C" {

int64_t ssvb_ntxent_dpad(int64_t d) { return sim_dpad(d); }
int64_t ssvb_ntxent_mpad(int64_t n_global) { return sim_mpad(2 * n_global); }

size_t ssvb_ntxent_saved_bytes(int64_t n, int64_t d) {
  if (n <= 0 || d <= 0) return 0;
  return saved_layout(nullptr, sim_mpad(2 * n), sim_dpad(d)).bytes;
}
size_t ssvb_ntxent_workspace_bytes(int64_t n, int64_t d) {
  if (n <= 0 || d <= 0) return 0;
  return ws_layout(nullptr, 2 * n, ceil_div(2 * n, 128), 2 * n, sim_dpad(d)).bytes;
}

int ssvb_ntxent_fwd(const float* zi, const float* zj, int64_t n, int64_t d, int64_t ld_zi, int64_t ld_zj,
                    int normalize, float temperature, float* loss, void* saved, void* workspace,
                    size_t workspace_bytes, void* stream) {
  SSVB_TRY(check_device_sm100());
  NtxPlan pl;
  SSVB_TRY(make_plan(pl, n, d, normalize, temperature));
  SSVB_TRY(check_rows(zi, ld_zi));
  SSVB_TRY(check_rows(zj, ld_zj));
  if (!loss || !saved || !workspace) return SSVB_ERR_INVALID;
  if (workspace_bytes < ssvb_ntxent_workspace_bytes(n, d)) return SSVB_ERR_WORKSPACE;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  SavedLayout sv = saved_layout(saved, pl.mpad, pl.dpad);
  WsLayout ws = ws_layout(workspace, pl.m, ceil_div(pl.m, 128), pl.m, pl.dpad);

  // zero the padding rows of the bf16 staging matrix and the reduction counter
  if (pl.mpad > pl.m)
    SSVB_CUDA(cudaMemsetAsync(sv.zhat + pl.m * pl.dpad, 0, (pl.mpad - pl.m) * pl.dpad * sizeof(__nv_bfloat16), s));
  SSVB_CUDA(cudaMemsetAsync(ws.counter, 0, 16, s));
  {
    const int wpb = 8;
    pair_prep_kernel<<<static_cast<unsigned>(ceil_div(n, wpb)), wpb * 32, 0, s>>>(
        zi, zj, static_cast<int>(n), static_cast<int>(d), ld_zi, ld_zj, normalize, normalize ? 1 : 0, sv.zhat,
        sv.zhat + n * pl.dpad,
        static_cast<int>(pl.dpad), sv.inv_norm, sv.inv_norm + n, ws.pos, ws.pos + n, -1, pl.prescale);
    SSVB_LAUNCH_CHECK();
  }
  SimParams p;
  fill_sim_params_rows(p, pl, 1, pl.m, 0, 0);
  plan_chunks(p, 256, 4);
  p.part_m = ws.part_m;
  p.part_l = ws.part_l;
  p.part_stride = static_cast<int>(round_up(pl.m, 256));
#ifdef SSVB_DBG_TIMING
  if (const char* e = getenv("SSVB_DBG_PTR")) p.dbg = reinterpret_cast<unsigned long long*>(strtoull(e, nullptr, 0));
#endif
  SSVB_TRY(launch_sim_fwd(pl.mode, sv.zhat, pl.mpad, sv.zhat, pl.mpad, pl.dpad, p, s));
  {
    const unsigned grid = static_cast<unsigned>(ceil_div(pl.m, 256));
    const float scale = 1.f / static_cast<float>(pl.m);
    if (pl.mode == SIM_NTX_FIXED)
      lse_finalize_kernel<SIM_NTX_FIXED><<<grid, 256, 0, s>>>(ws.part_m, ws.part_l, 4 * p.nchunks, p.part_stride,
                                                            static_cast<int>(pl.m), ws.pos, pl.c, pl.shift,
                                                            sv.stat, nullptr, ws.block_sums, ws.counter, scale, loss,
                                                            nullptr, nullptr, 0, 0, pl.wscale);
    else
      lse_finalize_kernel<SIM_NTX_ONLINE><<<grid, 256, 0, s>>>(ws.part_m, ws.part_l, 4 * p.nchunks, p.part_stride,
                                                             static_cast<int>(pl.m), ws.pos, pl.c, pl.shift,
                                                             sv.stat, nullptr, ws.block_sums, ws.counter, scale, loss);
    SSVB_LAUNCH_CHECK();
    if (pl.mpad > pl.m) {
      const float padv = pl.mode == SIM_NTX_FIXED ? 0.f : 1e30f;
      fill_kernel<<<static_cast<unsigned>(ceil_div(pl.mpad - pl.m, 256)), 256, 0, s>>>(sv.stat + pl.m,
                                                                                       pl.mpad - pl.m, padv);
      SSVB_LAUNCH_CHECK();
    }
  }
  return SSVB_OK;
}

int ssvb_ntxent_bwd(const float* zi, const float* zj, int64_t n, int64_t d, int64_t ld_zi, int64_t ld_zj,
                    int normalize, float temperature, const float* grad_out, const void* saved, float* dzi,
                    float* dzj, int64_t ld_dzi, int64_t ld_dzj, void* workspace, size_t workspace_bytes,
                    void* stream) {
  SSVB_TRY(check_device_sm100());
  NtxPlan pl;
  SSVB_TRY(make_plan(pl, n, d, normalize, temperature));
  SSVB_TRY(check_rows(zi, ld_zi));
  SSVB_TRY(check_rows(zj, ld_zj));
  SSVB_TRY(check_rows(dzi, ld_dzi));
  SSVB_TRY(check_rows(dzj, ld_dzj));
  if (!grad_out || !saved || !workspace) return SSVB_ERR_INVALID;
  if (workspace_bytes < ssvb_ntxent_workspace_bytes(n, d)) return SSVB_ERR_WORKSPACE;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  SavedLayout sv = saved_layout(const_cast<void*>(saved), pl.mpad, pl.dpad);
  WsLayout ws = ws_layout(workspace, pl.m, ceil_div(pl.m, 128), pl.m, pl.dpad);

  SimParams p;
  fill_sim_params_rows(p, pl, 1, pl.m, 0, 0);
  plan_chunks(p, 128, 8);
  p.rowstat = sv.stat;
  p.colstat = sv.stat;
  p.dacc = ws.dacc;
  p.ld_dacc = static_cast<int>(pl.dpad);
  p.use_atomic = p.nchunks > 1;
#ifdef SSVB_DBG_TIMING
  if (const char* e = getenv("SSVB_DBG_PTR")) p.dbg = reinterpret_cast<unsigned long long*>(strtoull(e, nullptr, 0));
#endif
  if (p.use_atomic) SSVB_CUDA(cudaMemsetAsync(ws.dacc, 0, pl.m * pl.dpad * sizeof(float), s));
  SSVB_TRY(launch_sim_bwd(pl.mode, sv.zhat, pl.mpad, sv.zhat, pl.mpad, pl.dpad, p, s));
  {
    const int wpb = 8;
    ntx_grad_finish_kernel<<<static_cast<unsigned>(ceil_div(pl.m, wpb)), wpb * 32, 0, s>>>(
        zi, zj, ld_zi, ld_zj, static_cast<int>(n), static_cast<int>(d), 0, ws.dacc,
        static_cast<int>(pl.dpad), sv.zhat, static_cast<int>(pl.dpad), sv.inv_norm, normalize,
        1.f / (static_cast<float>(pl.m) * temperature), grad_out, dzi, dzj, ld_dzi, ld_dzj,
        1.f / (pl.wscale * pl.prescale), 1.f / pl.prescale);
    SSVB_LAUNCH_CHECK();
  }
  return SSVB_OK;
}

// ------------------------------------------------------------------------------------ multi-GPU pieces
// Gathered layout is RANK-MAJOR: global row of (rank r, view v, local row i) = r*2L + v*L + i with L = n_local,
// so every rank owns ONE contiguous slot of zhat_all / stat_all (a single all-gather each) and the positive
// partner of a row is always on the same rank (a +- L).  NT-Xent is invariant to this row permutation.
size_t ssvb_ntxent_dist_workspace_bytes(int64_t world, int64_t n_local, int64_t d) {
  if (world <= 0 || n_local <= 0 || d <= 0) return 0;
  return ws_layout(nullptr, 2 * n_local, ceil_div(2 * n_local, 128), 2 * n_local * world, sim_dpad(d)).bytes;
}

namespace {
int dist_check(int64_t world, int64_t rank, int64_t n_local) {
  if (world <= 0 || rank < 0 || rank >= world || n_local <= 0) return SSVB_ERR_INVALID;
  return SSVB_OK;
}
}  // namespace

int ssvb_ntxent_dist_prep(const float* zi, const float* zj, int64_t n_local, int64_t d, int64_t ld_zi,
                          int64_t ld_zj, int normalize, float temperature, int64_t world, int64_t rank,
                          void* zhat_all, float* inv_norm_local, float* pos_local, void* stream) {
  SSVB_TRY(check_device_sm100());
  SSVB_TRY(dist_check(world, rank, n_local));
  NtxPlan pl;
  SSVB_TRY(make_plan(pl, n_local * world, d, normalize, temperature));
  SSVB_TRY(check_rows(zi, ld_zi));
  SSVB_TRY(check_rows(zj, ld_zj));
  if (!zhat_all || !inv_norm_local || !pos_local) return SSVB_ERR_INVALID;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  __nv_bfloat16* zh = static_cast<__nv_bfloat16*>(zhat_all);
  if (pl.mpad > pl.m)
    SSVB_CUDA(cudaMemsetAsync(zh + pl.m * pl.dpad, 0, (pl.mpad - pl.m) * pl.dpad * sizeof(__nv_bfloat16), s));
  const int64_t row0 = rank * 2 * n_local;
  const int wpb = 8;
  pair_prep_kernel<<<static_cast<unsigned>(ceil_div(n_local, wpb)), wpb * 32, 0, s>>>(
      zi, zj, static_cast<int>(n_local), static_cast<int>(d), ld_zi, ld_zj, normalize, normalize ? 1 : 0,
      zh + row0 * pl.dpad,
      zh + (row0 + n_local) * pl.dpad, static_cast<int>(pl.dpad), inv_norm_local, inv_norm_local + n_local,
      pos_local, pos_local + n_local, -1, pl.prescale);
  SSVB_LAUNCH_CHECK();
  return SSVB_OK;
}

namespace {
int rows_fwd_impl(const void* zhat_all, int64_t world, int64_t rank, int64_t n_local, int64_t d, int normalize,
                  float temperature, const float* pos_local, float* stat_local, float* const* peer_stat,
                  float* loss_sum, void* workspace, size_t workspace_bytes, void* stream);
}
int ssvb_ntxent_dist_rows_fwd(const void* zhat_all, int64_t world, int64_t rank, int64_t n_local, int64_t d,
                              int normalize, float temperature, const float* pos_local, float* stat_local,
                              float* loss_sum, void* workspace, size_t workspace_bytes, void* stream) {
  if (!stat_local) return SSVB_ERR_INVALID;
  return rows_fwd_impl(zhat_all, world, rank, n_local, d, normalize, temperature, pos_local, stat_local, nullptr,
                       loss_sum, workspace, workspace_bytes, stream);
}
int ssvb_ntxent_dist_rows_fwd_push(const void* zhat_all, int64_t world, int64_t rank, int64_t n_local, int64_t d,
                                   int normalize, float temperature, const float* pos_local,
                                   void* const* peer_stat, float* loss_sum, void* workspace,
                                   size_t workspace_bytes, void* stream) {
  if (!peer_stat) return SSVB_ERR_INVALID;
  return rows_fwd_impl(zhat_all, world, rank, n_local, d, normalize, temperature, pos_local, nullptr,
                       reinterpret_cast<float* const*>(peer_stat), loss_sum, workspace, workspace_bytes, stream);
}
namespace {
int rows_fwd_impl(const void* zhat_all, int64_t world, int64_t rank, int64_t n_local, int64_t d, int normalize,
                  float temperature, const float* pos_local, float* stat_local, float* const* peer_stat,
                  float* loss_sum, void* workspace, size_t workspace_bytes, void* stream) {
  SSVB_TRY(check_device_sm100());
  SSVB_TRY(dist_check(world, rank, n_local));
  NtxPlan pl;
  SSVB_TRY(make_plan(pl, n_local * world, d, normalize, temperature));
  if (!zhat_all || !pos_local || !loss_sum || !workspace) return SSVB_ERR_INVALID;
  if (workspace_bytes < ssvb_ntxent_dist_workspace_bytes(world, n_local, d)) return SSVB_ERR_WORKSPACE;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int64_t lr = 2 * n_local;
  WsLayout ws = ws_layout(workspace, lr, ceil_div(lr, 128), pl.m, pl.dpad);
  SSVB_CUDA(cudaMemsetAsync(ws.counter, 0, 16, s));
  // local outputs either go to the caller's [2][2L] block or (push mode) into slot `rank` of every peer's buffer
  float* lse_out = stat_local;
  float* term_out = stat_local ? stat_local + lr : nullptr;
  const size_t peer_off = static_cast<size_t>(rank) * 2 * lr;
  SimParams p;
  fill_sim_params_rows(p, pl, 1, lr, rank * lr, 0);
  plan_chunks(p, 256, 4);
  p.part_m = ws.part_m;
  p.part_l = ws.part_l;
  p.part_stride = static_cast<int>(round_up(lr, 256));
  SSVB_TRY(launch_sim_fwd(pl.mode, zhat_all, pl.mpad, zhat_all, pl.mpad, pl.dpad, p, s));
  const unsigned grid = static_cast<unsigned>(ceil_div(lr, 256));
  // stat_local always carries the log2-domain LSE (what gets all-gathered); the FIXED-mode 1/L' is
  // re-derived from it in rows_bwd.
  if (pl.mode == SIM_NTX_FIXED)
    lse_finalize_kernel<SIM_NTX_FIXED><<<grid, 256, 0, s>>>(ws.part_m, ws.part_l, 4 * p.nchunks, p.part_stride,
                                                          static_cast<int>(lr), pos_local, pl.c, pl.shift,
                                                          ws.dacc /*scratch*/, lse_out, ws.block_sums, ws.counter,
                                                          1.f, loss_sum, term_out, peer_stat, static_cast<int>(world),
                                                          peer_off);
  else
    lse_finalize_kernel<SIM_NTX_ONLINE><<<grid, 256, 0, s>>>(ws.part_m, ws.part_l, 4 * p.nchunks, p.part_stride,
                                                           static_cast<int>(lr), pos_local, pl.c, pl.shift,
                                                           ws.dacc /*scratch*/, lse_out, ws.block_sums,
                                                           ws.counter, 1.f, loss_sum, term_out, peer_stat,
                                                           static_cast<int>(world), peer_off);
  SSVB_LAUNCH_CHECK();
  return SSVB_OK;
}
}  // namespace

// Fused normalise + all-gather: like dist_prep, but every bf16 row goes to the same slot of EVERY rank's gathered
// matrix through `peer_zhat` (a DEVICE array of `world` peer-mapped base pointers, e.g. torch symmetric memory).
int ssvb_ntxent_dist_prep_push(const float* zi, const float* zj, int64_t n_local, int64_t d, int64_t ld_zi,
                               int64_t ld_zj, int normalize, float temperature, int64_t world, int64_t rank,
                               void* const* peer_zhat, float* inv_norm_local, float* pos_local, void* stream) {
  SSVB_TRY(check_device_sm100());
  SSVB_TRY(dist_check(world, rank, n_local));
  NtxPlan pl;
  SSVB_TRY(make_plan(pl, n_local * world, d, normalize, temperature));
  SSVB_TRY(check_rows(zi, ld_zi));
  SSVB_TRY(check_rows(zj, ld_zj));
  if (!peer_zhat || !inv_norm_local || !pos_local) return SSVB_ERR_INVALID;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int64_t row0 = rank * 2 * n_local;
  pair_prep_push_kernel<<<static_cast<unsigned>(ceil_div(n_local, 8)), 256, 0, s>>>(
      zi, zj, static_cast<int>(n_local), static_cast<int>(d), ld_zi, ld_zj, normalize, normalize ? 1 : 0, peer_zhat,
      static_cast<int>(world), row0, row0 + n_local, static_cast<int>(pl.dpad), inv_norm_local,
      inv_norm_local + n_local, pos_local, pos_local + n_local, pl.prescale);
  SSVB_LAUNCH_CHECK();
  return SSVB_OK;
}

// global loss from the gathered per-row terms: loss = sum(stat_all[:, 1, :]) / (2 * world * L); fixed order
namespace {
__global__ void dist_loss_kernel(const float* __restrict__ g, int world, int lr, float scale, float* __restrict__ out) {
  // sum of the gathered per-row loss terms ([world][2][lr], second half of every block) in a FIXED order (identical on
  // every rank).  8 independent loads in flight per thread: the serial version (64 dependent-latency loads per thread)
  // took 29 us at world*lr = 65536.
  __shared__ float red[32];
  float acc[8];
#pragma unroll
  for (int u = 0; u < 8; ++u) acc[u] = 0.f;
  for (int r = 0; r < world; ++r) {
    const float* t = g + static_cast<size_t>(r) * 2 * lr + lr;
    int i = threadIdx.x;
    for (; i + 7 * static_cast<int>(blockDim.x) < lr; i += 8 * blockDim.x) {
#pragma unroll
      for (int u = 0; u < 8; ++u) acc[u] += t[i + u * blockDim.x];
    }
    for (; i < lr; i += blockDim.x) acc[0] += t[i];
  }
  float v = ((acc[0] + acc[1]) + (acc[2] + acc[3])) + ((acc[4] + acc[5]) + (acc[6] + acc[7]));
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    t = warp_sum(t);
    if (threadIdx.x == 0) out[0] = t * scale;
  }
}
}  // namespace
int ssvb_ntxent_dist_loss(const float* stat_all, int64_t world, int64_t n_local, float* loss, void* stream) {
  SSVB_TRY(check_device_sm100());
  if (!stat_all || !loss || world <= 0 || n_local <= 0) return SSVB_ERR_INVALID;
  dist_loss_kernel<<<1, 1024, 0, static_cast<cudaStream_t>(stream)>>>(
      stat_all, static_cast<int>(world), static_cast<int>(2 * n_local), 1.f / static_cast<float>(2 * world * n_local),
      loss);
  SSVB_LAUNCH_CHECK();
  return SSVB_OK;
}

namespace {
// lse2 (all rows) -> the column statistic the backward kernel consumes, with finite padding
// gathered layout: [world][2][2L] (per rank: 2L lse2 values, then 2L per-row loss terms)
__global__ void dist_stat_kernel(const float* __restrict__ gathered, float* __restrict__ stat, int m, int mpad, int lr,
                                 int fixed, float shift, float wscale) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= mpad) return;
  if (i < m) {
    const int r = i / lr;
    const float l2 = gathered[static_cast<size_t>(r) * 2 * lr + (i - r * lr)];
    stat[i] = fixed ? wscale * exp2f(shift - l2) : l2;
  } else {
    stat[i] = fixed ? 0.f : 1e30f;
  }
}
}  // namespace

int ssvb_ntxent_dist_rows_bwd(const float* zi, const float* zj, int64_t n_local, int64_t d, int64_t ld_zi,
                              int64_t ld_zj, int normalize, float temperature, int64_t world, int64_t rank,
                              const void* zhat_all, const float* stat_all, const float* inv_norm_local,
                              const float* grad_out, float* dzi, float* dzj, int64_t ld_dzi, int64_t ld_dzj,
                              void* workspace, size_t workspace_bytes, void* stream) {
  SSVB_TRY(check_device_sm100());
  SSVB_TRY(dist_check(world, rank, n_local));
  NtxPlan pl;
  SSVB_TRY(make_plan(pl, n_local * world, d, normalize, temperature));
  SSVB_TRY(check_rows(zi, ld_zi));
  SSVB_TRY(check_rows(zj, ld_zj));
  SSVB_TRY(check_rows(dzi, ld_dzi));
  SSVB_TRY(check_rows(dzj, ld_dzj));
  if (!zhat_all || !stat_all || !inv_norm_local || !grad_out || !workspace) return SSVB_ERR_INVALID;
  if (workspace_bytes < ssvb_ntxent_dist_workspace_bytes(world, n_local, d)) return SSVB_ERR_WORKSPACE;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int64_t lr = 2 * n_local;
  WsLayout ws = ws_layout(workspace, lr, ceil_div(lr, 128), pl.m, pl.dpad);
  // column statistics for all M rows live in the (otherwise unused here) partial buffer
  float* stat = ws.part_m;
  dist_stat_kernel<<<static_cast<unsigned>(ceil_div(pl.mpad, 256)), 256, 0, s>>>(
      stat_all, stat, static_cast<int>(pl.m), static_cast<int>(pl.mpad), static_cast<int>(lr), pl.mode == SIM_NTX_FIXED, pl.shift, pl.wscale);
  SSVB_LAUNCH_CHECK();

  SimParams p;
  fill_sim_params_rows(p, pl, 1, lr, rank * lr, 0);
  plan_chunks(p, 128, 8);
  p.rowstat = stat;
  p.colstat = stat;
  p.dacc = ws.dacc;
  p.ld_dacc = static_cast<int>(pl.dpad);
  p.use_atomic = p.nchunks > 1;
  if (p.use_atomic) SSVB_CUDA(cudaMemsetAsync(ws.dacc, 0, lr * pl.dpad * sizeof(float), s));
  SSVB_TRY(launch_sim_bwd(pl.mode, zhat_all, pl.mpad, zhat_all, pl.mpad, pl.dpad, p, s));
  const int wpb = 8;
  ntx_grad_finish_kernel<<<static_cast<unsigned>(ceil_div(lr, wpb)), wpb * 32, 0, s>>>(
      zi, zj, ld_zi, ld_zj, static_cast<int>(n_local), static_cast<int>(d), static_cast<int>(rank * lr), ws.dacc,
      static_cast<int>(pl.dpad), static_cast<const __nv_bfloat16*>(zhat_all), static_cast<int>(pl.dpad),
      inv_norm_local, normalize, 1.f / (static_cast<float>(pl.m) * temperature), grad_out, dzi, dzj, ld_dzi, ld_dzj,
        1.f / (pl.wscale * pl.prescale), 1.f / pl.prescale);
  SSVB_LAUNCH_CHECK();
  return SSVB_OK;
}

}  // extern "C"
