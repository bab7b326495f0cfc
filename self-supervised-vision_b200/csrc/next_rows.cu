// SURVEY.md §8(f) "next" rows built to the same bar as the hot path:
//   * multi-tensor EMA of the target / key / teacher network — models/moco.py:108-111, byol.py:120-123,
//     relic.py:119-122, dino.py:129-134:  t = m * t + (1 - m) * s   (bit-exact with the reference's three eager ops)
//   * DinoLoss forward + backward — utils/losses.py:75-89 (call site models/dino.py:161-162) and the centre EMA
//     (models/dino.py:136-141)
// Both are pure HBM-bandwidth kernels: coalesced float4 traffic, warp-shuffle / block reductions, fixed-order sums.
#include "gemm_host.cuh"

using namespace ssvb;

namespace {

// ------------------------------------------------------------------------------------------------ EMA
struct EmaChunk {
  float* t;
  const float* s;
  long long n;
};
constexpr int kEmaChunk = 8192;  // elements per table entry (host side splits tensors into chunks of this size)

// one block per chunk.  __fmul_rn / __fadd_rn keep the reference's rounding: m*t and (1-m)*s are rounded separately and
// then added (three eager ATen kernels), an FMA contraction would differ in the last bit.
__global__ void ema_kernel(const EmaChunk* __restrict__ table, float m, float om) {
  const EmaChunk c = table[blockIdx.x];
  const bool vec = ((reinterpret_cast<uintptr_t>(c.t) | reinterpret_cast<uintptr_t>(c.s)) & 15) == 0;
  if (vec) {
    const long long n4 = c.n >> 2;
    float4* t4 = reinterpret_cast<float4*>(c.t);
    const float4* s4 = reinterpret_cast<const float4*>(c.s);
    // 4 float4 of each operand per thread and trip: 8 independent 16-byte loads in flight before the first use
    // (r1 ncu: 54.7 % of DRAM throughput with long_scoreboard the top stall = too few bytes in flight)
    constexpr int U = 4;
    long long i = threadIdx.x;
    for (; i + (U - 1) * static_cast<long long>(blockDim.x) < n4; i += U * static_cast<long long>(blockDim.x)) {
      float4 a[U], b[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        a[u] = t4[i + u * blockDim.x];
        b[u] = __ldg(s4 + i + u * blockDim.x);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        float4 r;
        r.x = __fadd_rn(__fmul_rn(m, a[u].x), __fmul_rn(om, b[u].x));
        r.y = __fadd_rn(__fmul_rn(m, a[u].y), __fmul_rn(om, b[u].y));
        r.z = __fadd_rn(__fmul_rn(m, a[u].z), __fmul_rn(om, b[u].z));
        r.w = __fadd_rn(__fmul_rn(m, a[u].w), __fmul_rn(om, b[u].w));
        t4[i + u * blockDim.x] = r;
      }
    }
    for (; i < n4; i += blockDim.x) {
      const float4 a = t4[i];
      const float4 b = __ldg(s4 + i);
      float4 r;
      r.x = __fadd_rn(__fmul_rn(m, a.x), __fmul_rn(om, b.x));
      r.y = __fadd_rn(__fmul_rn(m, a.y), __fmul_rn(om, b.y));
      r.z = __fadd_rn(__fmul_rn(m, a.z), __fmul_rn(om, b.z));
      r.w = __fadd_rn(__fmul_rn(m, a.w), __fmul_rn(om, b.w));
      t4[i] = r;
    }
    for (long long i = (n4 << 2) + threadIdx.x; i < c.n; i += blockDim.x)
      c.t[i] = __fadd_rn(__fmul_rn(m, c.t[i]), __fmul_rn(om, __ldg(c.s + i)));
  } else {
    for (long long i = threadIdx.x; i < c.n; i += blockDim.x)
      c.t[i] = __fadd_rn(__fmul_rn(m, c.t[i]), __fmul_rn(om, __ldg(c.s + i)));
  }
}

// ------------------------------------------------------------------------------------------------ DINO
// One block per sample b.  teacher [bs][2][K], student [bs][nv][K], center [K].
//   T_g = softmax((teacher[b, g] - center) / temp_t)           g = 0, 1
//   logp_v = log_softmax(student[b, v] / temp_s)
//   loss = -(1 / (bs nv)) sum_b sum_v sum_k (T_0 + T_1) logp_v                   (utils/losses.py:86-89)
//   d student[b, v] = (2 softmax(student/temp_s) - (T_0 + T_1)) * grad_out / (temp_s bs nv)   (sum_k T_g = 1)
// The block first builds Tsum = T_0 + T_1 (in registers), then walks the nv student rows.
__device__ __forceinline__ float block_reduce_f(float v, bool is_max, float* red) {
  v = is_max ? warp_max(v) : warp_sum(v);
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  float t = lane < (blockDim.x >> 5) ? red[lane] : (is_max ? -INFINITY : 0.f);
  return is_max ? warp_max(t) : warp_sum(t);
}

// Every row (K floats) is loaded ONCE into registers (EPT values per thread, coalesced: column = tid + e * 256) and the
// max / sum / dot / gradient passes run on the registers; one block reduction per statistic.  The 2 + nv rows of a
// sample form one sequence and the loads of row r + 1 are issued BEFORE the reductions of row r (register double
// buffer), so every CTA keeps a row of HBM traffic in flight through its barrier / exp phases.  Tsum lives in shared
// memory (each thread only ever touches its own columns: no barrier needed), which keeps the kernel at <= 51 registers
// = 5 CTAs per SM for K <= 4096 (5 x 16 KB of loads in flight per SM).
template <bool BWD, int EPT>
__global__ void __launch_bounds__(256, (EPT <= 16 ? 5 : (EPT <= 32 ? 2 : 1)))
dino_kernel(const float* __restrict__ teacher, int64_t ld_tb, int64_t ld_tv, const float* __restrict__ student,
            int64_t ld_sb, int64_t ld_sv, int nv, int k, const float* __restrict__ center, float inv_ts, float inv_tt,
            float* __restrict__ loss_part /* [bs] */, const float* __restrict__ grad_out, float coef,
            float* __restrict__ dstudent, int64_t ld_db, int64_t ld_dv) {
  __shared__ float red[32];
  const int64_t b = blockIdx.x;
  const int nrows = 2 + nv;
  auto row_ptr = [&](int r) -> const float* {
    return r < 2 ? teacher + b * ld_tb + r * ld_tv : student + b * ld_sb + static_cast<int64_t>(r - 2) * ld_sv;
  };
  extern __shared__ float tsum[];  // [EPT * 256]: column c = threadIdx.x + e * 256 belongs to this thread
  float nxt[EPT];
#pragma unroll
  for (int e = 0; e < EPT; ++e) {
    const int c = threadIdx.x + e * 256;
    nxt[e] = c < k ? row_ptr(0)[c] : 0.f;
  }
  float acc = 0.f;
  const float go = BWD ? __ldg(grad_out) * coef : 0.f;
  for (int r = 0; r < nrows; ++r) {
    float x[EPT];
    const bool is_teacher = r < 2;
    float m = -INFINITY;
#pragma unroll
    for (int e = 0; e < EPT; ++e) {
      const int c = threadIdx.x + e * 256;
      const float raw = nxt[e];
      x[e] = c < k ? (is_teacher ? (raw - __ldg(center + c)) * inv_tt : raw * inv_ts) : -INFINITY;
      m = fmaxf(m, x[e]);
    }
    if (r + 1 < nrows) {  // block-uniform: next row's loads go out before this row's reductions
      const float* np = row_ptr(r + 1);
#pragma unroll
      for (int e = 0; e < EPT; ++e) {
        const int c = threadIdx.x + e * 256;
        nxt[e] = c < k ? np[c] : 0.f;
      }
    }
    m = block_reduce_f(m, true, red);
    float z = 0.f, dot = 0.f;
#pragma unroll
    for (int e = 0; e < EPT; ++e) {
      if (!BWD && !is_teacher && threadIdx.x + e * 256 < k) dot = fmaf(tsum[threadIdx.x + e * 256], x[e], dot);
      x[e] = __expf(x[e] - m);  // exp(-inf) = 0 for the padding
      z += x[e];
    }
    z = block_reduce_f(z, false, red);
    if (is_teacher) {
      const float iz = 1.f / z;
#pragma unroll
      for (int e = 0; e < EPT; ++e)
        tsum[threadIdx.x + e * 256] = r == 0 ? x[e] * iz : fmaf(x[e], iz, tsum[threadIdx.x + e * 256]);
    } else if (!BWD) {
      dot = block_reduce_f(dot, false, red);
      acc += dot - 2.f * (m + __logf(z));  // sum_k Tsum (x - lse) with sum_k Tsum = 2
    } else {
      const float iz = 2.f / z;
      float* d = dstudent + b * ld_db + static_cast<int64_t>(r - 2) * ld_dv;
#pragma unroll
      for (int e = 0; e < EPT; ++e) {
        const int c = threadIdx.x + e * 256;
        if (c < k) d[c] = (x[e] * iz - tsum[c]) * go;
      }
    }
  }
  if (!BWD && threadIdx.x == 0) loss_part[b] = -acc;
}

template <bool BWD>
int dino_launch(const float* teacher, const float* student, const float* center, int64_t bs, int64_t nv, int64_t k,
                float temp_s, float temp_t, float* part, const float* grad_out, float coef, float* dstudent,
                cudaStream_t s) {
  const unsigned grid = static_cast<unsigned>(bs);
  const int ki = static_cast<int>(k), nvi = static_cast<int>(nv);
  const float its = 1.f / temp_s, itt = 1.f / temp_t;
#define SSVB_DINO(E)                                                                                                   \
  dino_kernel<BWD, E><<<grid, 256, E * 256 * sizeof(float), s>>>(teacher, 2 * k, k, student, nv * k, k, nvi, ki, center, its, itt, part, grad_out, \
                                           coef, dstudent, nv * k, k)
  if (k <= 4 * 256) SSVB_DINO(4);
  else if (k <= 8 * 256) SSVB_DINO(8);
  else if (k <= 16 * 256) SSVB_DINO(16);
  else if (k <= 32 * 256) SSVB_DINO(32);
  else return SSVB_ERR_UNSUPPORTED;
#undef SSVB_DINO
  SSVB_LAUNCH_CHECK();
  return SSVB_OK;
}

// centre EMA (models/dino.py:136-141): center = m * center + (1 - m) * mean_rows(teacher_fvecs); first call: plain mean.
// block (32 columns x 8 row phases), fixed-order combine -> deterministic.
__global__ void dino_center_kernel(const float* __restrict__ t, int64_t rows, int k, int64_t ld, float m, float om,
                                   int first, float* __restrict__ center) {
  __shared__ float sh[8][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  float s = 0.f;
  if (c < k)
    for (int64_t r = threadIdx.y; r < rows; r += 8) s += t[r * ld + c];
  sh[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && c < k) {
    float a = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) a += sh[i][threadIdx.x];
    const float mean = a / static_cast<float>(rows);
    center[c] = first ? mean : __fadd_rn(__fmul_rn(m, center[c]), __fmul_rn(om, mean));
  }
}

}  // namespace

extern "C" {

int64_t ssvb_ema_chunk_elems(void) { return kEmaChunk; }

int ssvb_ema_update(const void* chunk_table, int64_t n_chunks, float m, float one_minus_m, void* stream) {
  SSVB_TRY(check_device_sm100());
  if (n_chunks < 0 || (n_chunks > 0 && !chunk_table)) return SSVB_ERR_INVALID;
  if (n_chunks == 0) return SSVB_OK;
  if (n_chunks > 0x7fffffffLL) return SSVB_ERR_UNSUPPORTED;
  if (reinterpret_cast<uintptr_t>(chunk_table) & 7) return SSVB_ERR_ALIGNMENT;
  ema_kernel<<<static_cast<unsigned>(n_chunks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const EmaChunk*>(chunk_table), m, one_minus_m);
  SSVB_LAUNCH_CHECK();
  return SSVB_OK;
}

size_t ssvb_dino_workspace_bytes(int64_t bs) { return bs > 0 ? static_cast<size_t>(bs) * sizeof(float) + 256 : 0; }

int ssvb_dino_fwd(const float* teacher, const float* student, const float* center, int64_t bs, int64_t nv, int64_t k,
                  float temp_s, float temp_t, float* loss, void* workspace, size_t workspace_bytes, void* stream) {
  SSVB_TRY(check_device_sm100());
  if (!teacher || !student || !center || !loss || !workspace || bs <= 0 || nv <= 0 || k <= 0 || !(temp_s > 0.f) ||
      !(temp_t > 0.f))
    return SSVB_ERR_INVALID;
  if (k > 8192 || bs > 0x7fffffffLL) return SSVB_ERR_UNSUPPORTED;
  if (workspace_bytes < ssvb_dino_workspace_bytes(bs)) return SSVB_ERR_WORKSPACE;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  float* part = static_cast<float*>(workspace);
  SSVB_TRY(dino_launch<false>(teacher, student, center, bs, nv, k, temp_s, temp_t, part, nullptr, 0.f, nullptr, s));
  sum_partials_kernel<<<1, 1024, 0, s>>>(part, static_cast<int>(bs), 1.f / static_cast<float>(bs * nv), loss);
  SSVB_LAUNCH_CHECK();
  return SSVB_OK;
}

int ssvb_dino_bwd(const float* teacher, const float* student, const float* center, int64_t bs, int64_t nv, int64_t k,
                  float temp_s, float temp_t, const float* grad_out, float* dstudent, void* stream) {
  SSVB_TRY(check_device_sm100());
  if (!teacher || !student || !center || !grad_out || !dstudent || bs <= 0 || nv <= 0 || k <= 0 || !(temp_s > 0.f) ||
      !(temp_t > 0.f))
    return SSVB_ERR_INVALID;
  if (k > 8192 || bs > 0x7fffffffLL) return SSVB_ERR_UNSUPPORTED;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  return dino_launch<true>(teacher, student, center, bs, nv, k, temp_s, temp_t, nullptr, grad_out,
                           1.f / (temp_s * static_cast<float>(bs * nv)), dstudent, s);
}

int ssvb_dino_center_update(const float* teacher_rows, int64_t rows, int64_t k, int64_t ld, float momentum,
                            float one_minus_m, int first, float* center, void* stream) {
  SSVB_TRY(check_device_sm100());
  if (!teacher_rows || !center || rows <= 0 || k <= 0 || ld < k) return SSVB_ERR_INVALID;
  dino_center_kernel<<<static_cast<unsigned>(ceil_div(k, 32)), dim3(32, 8), 0, static_cast<cudaStream_t>(stream)>>>(
      teacher_rows, rows, static_cast<int>(k), ld, momentum, one_minus_m, first, center);
  SSVB_LAUNCH_CHECK();
  return SSVB_OK;
}

}  // extern "C"
