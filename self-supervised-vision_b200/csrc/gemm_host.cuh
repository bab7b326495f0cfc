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

// grid of a persistent GEMM launch (also the number of EPI_BARLOW loss partials written)
inline int gemm_grid(int64_t nunits, int max_ctas) {
  int grid = nunits < num_sms() ? static_cast<int>(nunits) : num_sms();
  if (max_ctas > 0 && grid > max_ctas) grid = max_ctas;
  return grid;
}

// split-K factor for a GEMM with `tiles` output tiles and `nkb` 64-wide K blocks: only when the tiles fill less than
// half of the SMs, every slice keeps >= 4 K blocks, and no slice is empty
inline void gemm_plan_splits(GemmParams& p, int64_t tiles, bool allow) {
  const int nkb = static_cast<int>(ceil_div(p.K, 64));
  int splits = 1;
  if (allow && tiles * 2 <= num_sms()) {
    splits = static_cast<int>(num_sms() / tiles);
    if (splits > nkb / 4) splits = nkb / 4;
    if (splits < 1) splits = 1;
  }
  p.kb_per_split = static_cast<int>(ceil_div(nkb, splits));
  p.splits = static_cast<int>(ceil_div(nkb, p.kb_per_split));
}

inline bool gemm_tma_store_allowed() {
  static int v = -1;
  if (v < 0) v = getenv("SSVB_GEMM_NO_TMA_STORE") ? 0 : 1;  // A/B switch: per-thread row stores instead
  return v != 0;
}

template <int BN, bool A_MN, bool B_MN, int EPI, bool DUAL>
int launch_gemm_t(const CUtensorMap* tm, GemmParams p, int max_ctas, cudaStream_t s) {
  auto kern = gemm_kernel<BN, A_MN, B_MN, EPI, DUAL>;
  constexpr int smem = GemmCfg<BN>::SMEM;
  SSVB_TRY((set_smem_once<gemm_kernel<BN, A_MN, B_MN, EPI, DUAL>>(smem)));
  const int64_t nunits = static_cast<int64_t>(p.tiles_m) * p.tiles_n * (DUAL ? 2 : 1) * p.splits;
  const int grid = gemm_grid(nunits, max_ctas);
  const int slot = prof_begin(PROF_GEMM, s);
  kern<<<grid, 192, smem, s>>>(tm[0], tm[1], tm[2], tm[3], tm[4], tm[5], p);
  prof_end(slot, s);
  SSVB_LAUNCH_CHECK();
  return SSVB_OK;
}

// C[M x N] = alpha * A * B^T-style contraction over K (see gemm_kernels.cuh).  bn = 128 or 256.
// For EPI_BARLOW the grid is capped at `max_ctas` (>= the number of p.loss_partials written = gemm_grid(tiles, max_ctas)).
// split_k: allow the split-K (add) epilogue - the caller has ZEROED p.out on the stream.
// A2 / B2 (with p.out2): a second problem of the same shape in the same launch (B2 MN-major, B K-major).
inline int launch_gemm(const GemmOperand& A, const GemmOperand& B, GemmParams p, int bn, int epi, int max_ctas,
                       cudaStream_t s, bool split_k = false, const GemmOperand* A2 = nullptr,
                       const GemmOperand* B2 = nullptr) {
  if (p.M <= 0 || p.N <= 0 || p.K <= 0) return SSVB_ERR_INVALID;
  const bool dual = A2 != nullptr;
  if (dual && (!B2 || A.mn_major || A2->mn_major || B.mn_major || !B2->mn_major || bn != 256 ||
               (epi != EPI_STORE_F32 && epi != EPI_BARLOW_BWD) || !p.out2))
    return SSVB_ERR_UNSUPPORTED;
  if (epi == EPI_BARLOW_BWD && (!dual || !p.vm || !p.vr || !p.vq || !p.vm2 || !p.vr2 || !p.vq2 || !p.xf || !p.xf2 || !p.go ||
                                (p.ldxf & 3) || (p.ldxf2 & 3) || (reinterpret_cast<uintptr_t>(p.xf) & 15) ||
                                (reinterpret_cast<uintptr_t>(p.xf2) & 15)))
    return SSVB_ERR_INVALID;
  const int64_t ldc2 = p.ldc2 > 0 ? p.ldc2 : p.ldc;
  CUtensorMap tm[6];
  auto operand_maps = [&](const GemmOperand& a, const GemmOperand& b, CUtensorMap* ta, CUtensorMap* tb) -> int {
    if (a.mn_major)
      SSVB_TRY(make_tmap_bf16(ta, a.ptr, p.K, p.M, a.ld, 64));
    else
      SSVB_TRY(make_tmap_bf16(ta, a.ptr, p.M, p.K, a.ld, 128));
    if (b.mn_major)
      SSVB_TRY(make_tmap_bf16(tb, b.ptr, p.K, p.N, b.ld, 64));
    else
      SSVB_TRY(make_tmap_bf16(tb, b.ptr, p.N, p.K, b.ld, bn));
    return SSVB_OK;
  };
  SSVB_TRY(operand_maps(A, B, &tm[0], &tm[1]));
  tm[3] = tm[0];
  tm[4] = tm[1];
  if (dual) SSVB_TRY(operand_maps(*A2, *B2, &tm[3], &tm[4]));
  p.tiles_m = static_cast<int>(ceil_div(p.M, 128));
  p.tiles_n = static_cast<int>(ceil_div(p.N, bn));
  // staged TMA-store epilogue whenever the output rows are 16-byte aligned; otherwise per-thread row stores
  if (epi == EPI_BARLOW) {
    p.tma_store = gemm_tma_store_allowed() && !(reinterpret_cast<uintptr_t>(p.dC) & 15) && (p.ld_dc % 8 == 0);
    if (p.tma_store) SSVB_TRY(make_tmap_out(&tm[2], p.dC, p.M, p.N, p.ld_dc, 2));
  } else {
    p.tma_store = gemm_tma_store_allowed() && !(reinterpret_cast<uintptr_t>(p.out) & 15) && (p.ldc % 4 == 0) &&
                  (!dual || (!(reinterpret_cast<uintptr_t>(p.out2) & 15) && (ldc2 % 4 == 0)));
    if (p.tma_store) SSVB_TRY(make_tmap_out(&tm[2], p.out, p.M, p.N, p.ldc, 4));
  }
  if (!p.tma_store) {
    tm[2] = tm[0];  // never dereferenced
    if (p.colpart || epi == EPI_BARLOW_BWD) return SSVB_ERR_ALIGNMENT;  // these live in the staged epilogue only
  }
  tm[5] = tm[2];
  if (dual && p.tma_store) SSVB_TRY(make_tmap_out(&tm[5], p.out2, p.M, p.N, ldc2, 4));
  gemm_plan_splits(p, static_cast<int64_t>(p.tiles_m) * p.tiles_n,
                   split_k && p.tma_store && epi == EPI_STORE_F32 && !dual && !p.colpart);
#define SSVB_G(BNV, AM, BM_, E, D) return launch_gemm_t<BNV, AM, BM_, E, D>(tm, p, max_ctas, s)
  if (epi == EPI_BARLOW) {
    if (bn == 256 && A.mn_major && B.mn_major) SSVB_G(256, true, true, EPI_BARLOW, false);
    return SSVB_ERR_UNSUPPORTED;
  }
  if (dual && epi == EPI_BARLOW_BWD) SSVB_G(256, false, false, EPI_BARLOW_BWD, true);
  if (epi == EPI_BARLOW_BWD) return SSVB_ERR_UNSUPPORTED;
  if (dual) SSVB_G(256, false, false, EPI_STORE_F32, true);
  if (bn == 256) {
    if (!A.mn_major && !B.mn_major) SSVB_G(256, false, false, EPI_STORE_F32, false);
    if (!A.mn_major && B.mn_major) SSVB_G(256, false, true, EPI_STORE_F32, false);
    if (A.mn_major && B.mn_major) SSVB_G(256, true, true, EPI_STORE_F32, false);
  } else if (bn == 128) {
    if (!A.mn_major && !B.mn_major) SSVB_G(128, false, false, EPI_STORE_F32, false);
    if (!A.mn_major && B.mn_major) SSVB_G(128, false, true, EPI_STORE_F32, false);
    if (A.mn_major && B.mn_major) SSVB_G(128, true, true, EPI_STORE_F32, false);
  }
#undef SSVB_G
  return SSVB_ERR_UNSUPPORTED;
}

// whether launch_gemm(..., split_k = true) will use the add epilogue for this problem (the caller then zeroes the output)
inline bool gemm_will_split(int64_t M, int64_t N, int64_t K, int bn, const float* out, int64_t ldc) {
  if (!gemm_tma_store_allowed() || (reinterpret_cast<uintptr_t>(out) & 15) || (ldc % 4)) return false;
  GemmParams p{};
  p.K = static_cast<int>(K);
  gemm_plan_splits(p, ceil_div(M, 128) * ceil_div(N, bn), true);
  return p.splits > 1;
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
