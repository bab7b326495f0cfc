// Host-side launch of the tcgen05 GEMM (gemm_kernels.cuh) + small shared bandwidth kernels.
#pragma once
#include "host_util.h"
#include "gemm_kernels.cuh"

namespace ssvb {

struct GemmOperand {
  const void* ptr;   // bf16
  int64_t ld;        // elements
  bool mn_major;     // false: global [rows x K]; true: global [K x rows]
};

template <int BN, bool A_MN, bool B_MN, int EPI>
int launch_gemm_t(const CUtensorMap& tmA, const CUtensorMap& tmB, GemmParams p, int max_ctas, cudaStream_t s) {
  auto kern = gemm_kernel<BN, A_MN, B_MN, EPI>;
  constexpr int smem = GemmCfg<BN>::SMEM;
  SSVB_TRY((set_smem_once<gemm_kernel<BN, A_MN, B_MN, EPI>>(smem)));
  const int ntiles = p.tiles_m * p.tiles_n;
  int grid = ntiles < num_sms() ? ntiles : num_sms();
  if (max_ctas > 0 && grid > max_ctas) grid = max_ctas;
  const int slot = prof_begin(PROF_GEMM, s);
  kern<<<grid, 192, smem, s>>>(tmA, tmB, p);
  prof_end(slot, s);
  SSVB_LAUNCH_CHECK();
  return SSVB_OK;
}

// C[M x N] = alpha * A * B^T-style contraction over K (see gemm_kernels.cuh).  bn = 128 or 256.
// For EPI_BARLOW the grid is capped at `max_ctas` (= size of p.loss_partials).
inline int launch_gemm(const GemmOperand& A, const GemmOperand& B, GemmParams p, int bn, int epi, int max_ctas,
                       cudaStream_t s) {
  if (p.M <= 0 || p.N <= 0 || p.K <= 0) return SSVB_ERR_INVALID;
  CUtensorMap tmA, tmB;
  if (A.mn_major)
    SSVB_TRY(make_tmap_bf16(&tmA, A.ptr, p.K, p.M, A.ld, 64));
  else
    SSVB_TRY(make_tmap_bf16(&tmA, A.ptr, p.M, p.K, A.ld, 128));
  if (B.mn_major)
    SSVB_TRY(make_tmap_bf16(&tmB, B.ptr, p.K, p.N, B.ld, 64));
  else
    SSVB_TRY(make_tmap_bf16(&tmB, B.ptr, p.N, p.K, B.ld, bn));
  p.tiles_m = static_cast<int>(ceil_div(p.M, 128));
  p.tiles_n = static_cast<int>(ceil_div(p.N, bn));
#define SSVB_G(BNV, AM, BM_, E) return launch_gemm_t<BNV, AM, BM_, E>(tmA, tmB, p, max_ctas, s)
  if (epi == EPI_BARLOW) {
    if (bn == 256 && A.mn_major && B.mn_major) SSVB_G(256, true, true, EPI_BARLOW);
    return SSVB_ERR_UNSUPPORTED;
  }
  if (bn == 256) {
    if (!A.mn_major && !B.mn_major) SSVB_G(256, false, false, EPI_STORE_F32);
    if (!A.mn_major && B.mn_major) SSVB_G(256, false, true, EPI_STORE_F32);
    if (A.mn_major && B.mn_major) SSVB_G(256, true, true, EPI_STORE_F32);
  } else if (bn == 128) {
    if (!A.mn_major && !B.mn_major) SSVB_G(128, false, false, EPI_STORE_F32);
    if (!A.mn_major && B.mn_major) SSVB_G(128, false, true, EPI_STORE_F32);
    if (A.mn_major && B.mn_major) SSVB_G(128, true, true, EPI_STORE_F32);
  }
#undef SSVB_G
  return SSVB_ERR_UNSUPPORTED;
}

// fp32 [rows x d] (ld) -> bf16 [rows x dpad] zero padded; optional per-row scale.  One warp per row.
static __global__ void rows_to_bf16_kernel(const float* __restrict__ x, int64_t rows, int d, int64_t ld, int dpad,
                                           const float* __restrict__ row_scale, __nv_bfloat16* __restrict__ out) {
  const int64_t row = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float sc = row_scale ? row_scale[row] : 1.f;
  for (int c = lane * 4; c < dpad; c += 128) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c < d) v = __ldg(reinterpret_cast<const float4*>(x + row * ld + c));
    __nv_bfloat162 lo = __floats2bfloat162_rn(v.x * sc, v.y * sc), hi = __floats2bfloat162_rn(v.z * sc, v.w * sc);
    uint2 pk;
    pk.x = *reinterpret_cast<uint32_t*>(&lo);
    pk.y = *reinterpret_cast<uint32_t*>(&hi);
    *reinterpret_cast<uint2*>(out + row * dpad + c) = pk;
  }
}

// sum of `n` floats in index order by one block -> out[0] = scale * sum (deterministic)
static __global__ void sum_partials_kernel(const float* __restrict__ part, int n, float scale, float* __restrict__ out) {
  __shared__ float red[32];
  float v = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) v += part[i];
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    t = warp_sum(t);
    if (threadIdx.x == 0) out[0] = t * scale;
  }
}

}  // namespace ssvb
