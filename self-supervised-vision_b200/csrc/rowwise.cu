// Bandwidth-bound row-wise kernels: BYOL MSE / SimSiam neg-cosine (a8, a9), row L2-normalise fwd/bwd,
// ring-buffer enqueue (a3, a7) and the ReLIC KL term (a10).  Coalesced float4 accesses, warp-shuffle
// reductions, deterministic scalar reductions.
#include "sim_host.cuh"

using namespace ssvb;

namespace {

// fixed-order sum of the per-block partials by one block -> out = scale * sum.  The single-kernel losses use this
// second tiny launch instead of a last-block counter: a counter needs a memset node in front of the kernel, and a
// memset -> kernel edge costs 2-4 us in a CUDA graph where a kernel -> kernel edge costs ~0.3 us.
__global__ void block_sums_finish_kernel(const float* __restrict__ part, int n, float scale, float* __restrict__ out) {
  __shared__ float red[8];
  float v = 0.f;
  for (int i = threadIdx.x; i < n; i += 256) v += part[i];
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = threadIdx.x < 8 ? red[threadIdx.x] : 0.f;
    t = warp_sum(t);
    if (threadIdx.x == 0) out[0] = t * scale;
  }
}

int check_rows(const void* p, int64_t ld) {
  if (!p) return SSVB_ERR_INVALID;
  if ((reinterpret_cast<uintptr_t>(p) & 15) || (ld & 3)) return SSVB_ERR_ALIGNMENT;
  return SSVB_OK;
}

// =========================================================================================== rowdot
// kind 0: sum (o-t)^2 ; kind 1: sum o*t          (models/byol.py:89 nn.MSELoss, utils/losses.py:150-151)
// FLAT: all rows contiguous (ld == d, the usual case): one flat float4 index, no 64-bit division per element.
// Every thread keeps U float4 of each operand in flight (2U independent 16-byte loads before the first use).
template <bool FLAT>
__device__ __forceinline__ int64_t rowdot_off(int64_t j, int d4, int64_t ld4) {
  if (FLAT) return j;
  const int64_t r = j / d4;
  return r * ld4 + (j - r * d4);
}
template <int KIND, bool FLAT>
__global__ void __launch_bounds__(256)
rowdot_fwd_kernel(const float* __restrict__ o, const float* __restrict__ t, int64_t n, int d4, int64_t ldo,
                  int64_t ldt, float* block_sums, unsigned int* counter, float scale, float* loss) {
  const int64_t total = n * d4;
  const float4* o4 = reinterpret_cast<const float4*>(o);
  const float4* t4 = reinterpret_cast<const float4*>(t);
  const int64_t lo4 = ldo >> 2, lt4 = ldt >> 2;
  float acc0 = 0.f, acc1 = 0.f;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  constexpr int U = 4;
  int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  for (; i + (U - 1) * stride < total; i += U * stride) {
    float4 a[U], b[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      a[u] = __ldg(o4 + rowdot_off<FLAT>(i + u * stride, d4, lo4));
      b[u] = __ldg(t4 + rowdot_off<FLAT>(i + u * stride, d4, lt4));
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (KIND == 0) {
        const float dx = a[u].x - b[u].x, dy = a[u].y - b[u].y, dz = a[u].z - b[u].z, dw = a[u].w - b[u].w;
        acc0 += dx * dx + dy * dy;
        acc1 += dz * dz + dw * dw;
      } else {
        acc0 += a[u].x * b[u].x + a[u].y * b[u].y;
        acc1 += a[u].z * b[u].z + a[u].w * b[u].w;
      }
    }
  }
  for (; i < total; i += stride) {
    const float4 a = __ldg(o4 + rowdot_off<FLAT>(i, d4, lo4));
    const float4 b = __ldg(t4 + rowdot_off<FLAT>(i, d4, lt4));
    if (KIND == 0) {
      const float dx = a.x - b.x, dy = a.y - b.y, dz = a.z - b.z, dw = a.w - b.w;
      acc0 += dx * dx + dy * dy;
      acc1 += dz * dz + dw * dw;
    } else {
      acc0 += a.x * b.x + a.y * b.y;
      acc1 += a.z * b.z + a.w * b.w;
    }
  }
  const float bt = block_sum_256(acc0 + acc1);
  grid_sum_finish(bt, block_sums, counter, scale, loss, false);
}

template <int KIND, bool FLAT>
__global__ void __launch_bounds__(256)
rowdot_bwd_kernel(const float* __restrict__ o, const float* __restrict__ t, int64_t n, int d4, int64_t ldo,
                  int64_t ldt, const float* __restrict__ grad_out, float coef, float* __restrict__ d_o,
                  float* __restrict__ d_t, int64_t lddo, int64_t lddt) {
  const int64_t total = n * d4;
  const float g = __ldg(grad_out) * coef;
  const float4* o4 = reinterpret_cast<const float4*>(o);
  const float4* t4 = reinterpret_cast<const float4*>(t);
  float4* do4 = reinterpret_cast<float4*>(d_o);
  float4* dt4 = reinterpret_cast<float4*>(d_t);
  const int64_t lo4 = ldo >> 2, lt4 = ldt >> 2, ldo4 = lddo >> 2, ldt4 = lddt >> 2;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  constexpr int U = 4;
  const bool need_o = (KIND == 0) || d_t != nullptr;   // kind 1: d_t = -o/N needs o, d_o = -t/N needs t
  const bool need_t = (KIND == 0) || d_o != nullptr;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += U * stride) {
    float4 a[U], b[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t j = i + u * stride;
      a[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      b[u] = a[u];
      if (j < total) {
        if (need_o) a[u] = __ldg(o4 + rowdot_off<FLAT>(j, d4, lo4));
        if (need_t) b[u] = __ldg(t4 + rowdot_off<FLAT>(j, d4, lt4));
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t j = i + u * stride;
      if (j >= total) break;
      float4 go, gt;
      if (KIND == 0) {  // d/do mean((o-t)^2) = 2 (o-t) / (N D)
        go = make_float4((a[u].x - b[u].x) * g, (a[u].y - b[u].y) * g, (a[u].z - b[u].z) * g, (a[u].w - b[u].w) * g);
        gt = make_float4(-go.x, -go.y, -go.z, -go.w);
      } else {  // d/do -(1/N) sum o t = -t / N
        go = make_float4(b[u].x * g, b[u].y * g, b[u].z * g, b[u].w * g);
        gt = make_float4(a[u].x * g, a[u].y * g, a[u].z * g, a[u].w * g);
      }
      if (d_o) do4[rowdot_off<FLAT>(j, d4, ldo4)] = go;
      if (d_t) dt4[rowdot_off<FLAT>(j, d4, ldt4)] = gt;
    }
  }
}

// =========================================================================================== rowdot on RAW rows (f1)
// The reference's BYOL / SimSiam heads end in F.normalize (models/byol.py:47,59; simsiam.py:48,69) and the loss then
// reads the unit rows again: normalise fwd (r + w per operand), loss fwd (2 r), loss bwd (2 r + w), normalise bwd
// (2 r + w per operand).  Fused: the loss takes the RAW head outputs, one pass computes |o|^2, |t|^2, o.t per row (warp
// per row) -> loss; the backward is one pass as well (normalise-backward projection from the saved row scalars).
struct RowdotNormSaved {
  float *inv_o, *inv_t, *cosv;
  size_t bytes;
};
RowdotNormSaved rowdot_norm_saved(void* base, int64_t n) {
  Carver c(base);
  RowdotNormSaved s;
  s.inv_o = c.take<float>(n);
  s.inv_t = c.take<float>(n);
  s.cosv = c.take<float>(n);
  s.bytes = c.used();
  return s;
}

template <int KIND>
__global__ void __launch_bounds__(256)
rowdot_norm_fwd_kernel(const float* __restrict__ o, const float* __restrict__ t, int64_t n, int d4, int64_t ldo,
                       int64_t ldt, int norm_o, int norm_t, RowdotNormSaved sv, float* block_sums,
                       unsigned int* counter, float scale, float* loss) {
  const int lane = threadIdx.x & 31;
  const int64_t wstride = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  float acc = 0.f;
  for (int64_t row = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5; row < n; row += wstride) {
    const float4* po = reinterpret_cast<const float4*>(o + row * ldo);
    const float4* pt = reinterpret_cast<const float4*>(t + row * ldt);
    float so = 0.f, st = 0.f, sot = 0.f;
    for (int c = lane; c < d4; c += 32) {
      const float4 a = __ldg(po + c), b = __ldg(pt + c);
      so += a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w;
      st += b.x * b.x + b.y * b.y + b.z * b.z + b.w * b.w;
      sot += a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w;
    }
    so = warp_sum(so); st = warp_sum(st); sot = warp_sum(sot);
    const float io = norm_o ? 1.f / fmaxf(sqrtf(so), 1e-12f) : 1.f;
    const float it = norm_t ? 1.f / fmaxf(sqrtf(st), 1e-12f) : 1.f;
    const float cs = sot * io * it;                 // o^ . t^
    if (lane == 0) {
      sv.inv_o[row] = io; sv.inv_t[row] = it; sv.cosv[row] = cs;
      acc += (KIND == 0) ? (so * io * io + st * it * it - 2.f * cs) : cs;
    }
  }
  const float bt = block_sum_256(acc);
  grid_sum_finish(bt, block_sums, counter, scale, loss, false);
}

template <int KIND>
__global__ void __launch_bounds__(256)
rowdot_norm_bwd_kernel(const float* __restrict__ o, const float* __restrict__ t, int64_t n, int d4, int64_t ldo,
                       int64_t ldt, int norm_o, int norm_t, const RowdotNormSaved sv,
                       const float* __restrict__ grad_out, float coef, float* __restrict__ d_o,
                       float* __restrict__ d_t, int64_t lddo, int64_t lddt) {
  const int lane = threadIdx.x & 31;
  const int64_t wstride = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  const float g = __ldg(grad_out) * coef;
  for (int64_t row = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5; row < n; row += wstride) {
    const float io = sv.inv_o[row], it = sv.inv_t[row], cs = sv.cosv[row];
    const float4* po = reinterpret_cast<const float4*>(o + row * ldo);
    const float4* pt = reinterpret_cast<const float4*>(t + row * ldt);
    float4* qo = d_o ? reinterpret_cast<float4*>(d_o + row * lddo) : nullptr;
    float4* qt = d_t ? reinterpret_cast<float4*>(d_t + row * lddt) : nullptr;
    // squared norms of the normalised rows (1 for ordinary rows, 0 for zero rows, |x|^2 when not normalised): needed for
    // the exact projection; recomputed from the row below
    float so = 0.f, st = 0.f;
    if (KIND == 0) {
      for (int c = lane; c < d4; c += 32) {
        const float4 a = __ldg(po + c), b = __ldg(pt + c);
        so += a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w;
        st += b.x * b.x + b.y * b.y + b.z * b.z + b.w * b.w;
      }
      so = warp_sum(so) * io * io;
      st = warp_sum(st) * it * it;
    }
    for (int c = lane; c < d4; c += 32) {
      const float4 a = __ldg(po + c), b = __ldg(pt + c);
      const float oh[4] = {a.x * io, a.y * io, a.z * io, a.w * io};
      const float th[4] = {b.x * it, b.y * it, b.z * it, b.w * it};
      float go[4], gt[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float dgo, dgt, po_dot, pt_dot;  // gradient w.r.t. the normalised rows and its dot with them
        if (KIND == 0) {  // d/do^ mean((o^-t^)^2) = 2 (o^-t^) g
          dgo = 2.f * (oh[e] - th[e]) * g;
          dgt = -dgo;
          po_dot = 2.f * (so - cs) * g;
          pt_dot = 2.f * (st - cs) * g;
        } else {          // d/do^ -(1/N) sum o^.t^ = -t^ g ... (g carries the sign and 1/N)
          dgo = th[e] * g;
          dgt = oh[e] * g;
          po_dot = cs * g;
          pt_dot = cs * g;
        }
        go[e] = norm_o ? (dgo - po_dot * oh[e]) * io : dgo;
        gt[e] = norm_t ? (dgt - pt_dot * th[e]) * it : dgt;
      }
      if (qo) qo[c] = make_float4(go[0], go[1], go[2], go[3]);
      if (qt) qt[c] = make_float4(gt[0], gt[1], gt[2], gt[3]);
    }
  }
}

// =========================================================================================== l2norm
__global__ void l2norm_fwd_kernel(const float* __restrict__ x, int64_t n, int d, int64_t ldx, float* __restrict__ y,
                                  int64_t ldy, float* __restrict__ inv_norm) {
  const int64_t row = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= n) return;
  const float4* xr = reinterpret_cast<const float4*>(x + row * ldx);
  float s = 0.f;
  for (int c = lane; c < d / 4; c += 32) {
    const float4 v = __ldg(xr + c);
    s += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  s = warp_sum(s);
  const float inv = 1.f / fmaxf(sqrtf(s), 1e-12f);
  float4* yr = reinterpret_cast<float4*>(y + row * ldy);
  for (int c = lane; c < d / 4; c += 32) {
    const float4 v = __ldg(xr + c);
    yr[c] = make_float4(v.x * inv, v.y * inv, v.z * inv, v.w * inv);
  }
  if (lane == 0 && inv_norm) inv_norm[row] = inv;
}

__global__ void l2norm_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y,
                                  const float* __restrict__ inv_norm, int64_t n, int d, int64_t lddy, int64_t ldy,
                                  float* __restrict__ dx, int64_t lddx) {
  const int64_t row = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= n) return;
  const float4* gr = reinterpret_cast<const float4*>(dy + row * lddy);
  const float4* yr = reinterpret_cast<const float4*>(y + row * ldy);
  float s = 0.f;
  for (int c = lane; c < d / 4; c += 32) {
    const float4 g = __ldg(gr + c), v = __ldg(yr + c);
    s += g.x * v.x + g.y * v.y + g.z * v.z + g.w * v.w;
  }
  s = warp_sum(s);
  const float inv = inv_norm[row];
  float4* o = reinterpret_cast<float4*>(dx + row * lddx);
  for (int c = lane; c < d / 4; c += 32) {
    const float4 g = __ldg(gr + c), v = __ldg(yr + c);
    o[c] = make_float4((g.x - s * v.x) * inv, (g.y - s * v.y) * inv, (g.z - s * v.z) * inv, (g.w - s * v.w) * inv);
  }
}

// =========================================================================================== ring enqueue
// one warp per batch row; rows that a later row of the same batch would overwrite (n > size) are skipped,
// which is exactly "last writer wins" of the reference's sequential loop.
__global__ void ring_enqueue_kernel(float* __restrict__ bank, __nv_bfloat16* __restrict__ bank_bf16, int64_t size,
                                    int d, int dpad, int64_t ld_bank, const float* __restrict__ batch, int64_t n,
                                    int64_t ld_batch, int64_t ptr, int normalize, int64_t shard_lo,
                                    int64_t shard_rows) {
  const int64_t i = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (i >= n || i < n - size) return;
  int64_t slot = (ptr + i) % size;
  // sharded ring: `bank` holds global rows [shard_lo, shard_lo + shard_rows); other slots belong to other ranks
  if (slot < shard_lo || slot >= shard_lo + shard_rows) return;
  slot -= shard_lo;
  const float4* src = reinterpret_cast<const float4*>(batch + i * ld_batch);
  float inv = 1.f;
  if (normalize) {
    float s = 0.f;
    for (int c = lane; c < d / 4; c += 32) {
      const float4 v = __ldg(src + c);
      s += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
    s = warp_sum(s);
    inv = fmaxf(sqrtf(s), 1e-12f);
  }
  float4* dst = reinterpret_cast<float4*>(bank + slot * ld_bank);
  for (int c = lane; c < dpad / 4; c += 32) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c < d / 4) {
      v = __ldg(src + c);
      if (normalize) {  // x / max(||x||, eps): true division, like F.normalize
        v.x = v.x / inv; v.y = v.y / inv; v.z = v.z / inv; v.w = v.w / inv;
      }
      dst[c] = v;
    }
    if (bank_bf16) {
      __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
      uint2 pk;
      pk.x = *reinterpret_cast<uint32_t*>(&lo);
      pk.y = *reinterpret_cast<uint32_t*>(&hi);
      *reinterpret_cast<uint2*>(bank_bf16 + slot * dpad + c * 4) = pk;
    }
  }
}

// =========================================================================================== PIRL indexed bank
// models/pirl.py:22-46: per-sample momentum bank.  mode 0 (initialize_vectors :32-34): bank[idx] = normalize(v);
// mode 1 (update_vectors :36-38): bank[idx] = m * bank[idx] + (1 - m) * normalize(v) (products and sum rounded
// separately, like the eager expression).  One warp per row; duplicate indices are not supported (as undefined in the
// reference's index_put).
__global__ void bank_scatter_kernel(float* __restrict__ bank, int64_t size, int d, int64_t ld_bank,
                                    const long long* __restrict__ idx, int64_t n, const float* __restrict__ v,
                                    int64_t ld_v, float m, float om, int mode) {
  const int64_t i = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (i >= n) return;
  const long long slot = idx[i];
  if (slot < 0 || slot >= size) return;
  const float4* src = reinterpret_cast<const float4*>(v + i * ld_v);
  float s = 0.f;
  for (int c = lane; c < d / 4; c += 32) {
    const float4 x = __ldg(src + c);
    s += x.x * x.x + x.y * x.y + x.z * x.z + x.w * x.w;
  }
  s = warp_sum(s);
  const float den = fmaxf(sqrtf(s), 1e-12f);
  float4* dst = reinterpret_cast<float4*>(bank + slot * ld_bank);
  for (int c = lane; c < d / 4; c += 32) {
    float4 x = __ldg(src + c);
    x.x = x.x / den; x.y = x.y / den; x.z = x.z / den; x.w = x.w / den;
    if (mode == 1) {
      const float4 b = dst[c];
      x.x = __fadd_rn(__fmul_rn(m, b.x), __fmul_rn(om, x.x));
      x.y = __fadd_rn(__fmul_rn(m, b.y), __fmul_rn(om, x.y));
      x.z = __fadd_rn(__fmul_rn(m, b.z), __fmul_rn(om, x.z));
      x.w = __fadd_rn(__fmul_rn(m, b.w), __fmul_rn(om, x.w));
    }
    dst[c] = x;
  }
}
// out[i] = bank[idx[i]]  (get_positives :40-41, get_negatives :43-45)
__global__ void bank_gather_kernel(const float* __restrict__ bank, int64_t size, int d, int64_t ld_bank,
                                   const long long* __restrict__ idx, int64_t n, float* __restrict__ out, int64_t ld_out) {
  const int64_t i = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (i >= n) return;
  const long long slot = idx[i];
  float4* dst = reinterpret_cast<float4*>(out + i * ld_out);
  for (int c = lane; c < d / 4; c += 32)
    dst[c] = (slot >= 0 && slot < size) ? __ldg(reinterpret_cast<const float4*>(bank + slot * ld_bank) + c)
                                        : make_float4(0.f, 0.f, 0.f, 0.f);
}

// =========================================================================================== ReLIC KL
struct RelicSaved {
  float *a, *b, *inv_i, *inv_j, *inv_o, *scal;  // scal: [lse_a, lse_b, sum(p*q), kl]
  size_t bytes;
};
RelicSaved relic_saved(void* base, int64_t n) {
  Carver c(base);
  RelicSaved s;
  s.a = c.take<float>(n);
  s.b = c.take<float>(n);
  s.inv_i = c.take<float>(n);
  s.inv_j = c.take<float>(n);
  s.inv_o = c.take<float>(n);
  s.scal = c.take<float>(8);
  s.bytes = c.used();
  return s;
}

__global__ void relic_dots_kernel(const float* __restrict__ zi, const float* __restrict__ zj,
                                  const float* __restrict__ zo, int64_t n, int d, int64_t ldi, int64_t ldj,
                                  int64_t ldo, int normalize, float inv_tau, RelicSaved sv) {
  const int64_t row = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= n) return;
  const float4* pi = reinterpret_cast<const float4*>(zi + row * ldi);
  const float4* pj = reinterpret_cast<const float4*>(zj + row * ldj);
  const float4* po = reinterpret_cast<const float4*>(zo + row * ldo);
  float sii = 0.f, sjj = 0.f, soo = 0.f, sio = 0.f, sjo = 0.f;
  for (int c = lane; c < d / 4; c += 32) {
    const float4 a = __ldg(pi + c), b = __ldg(pj + c), o = __ldg(po + c);
    sii += a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w;
    sjj += b.x * b.x + b.y * b.y + b.z * b.z + b.w * b.w;
    soo += o.x * o.x + o.y * o.y + o.z * o.z + o.w * o.w;
    sio += a.x * o.x + a.y * o.y + a.z * o.z + a.w * o.w;
    sjo += b.x * o.x + b.y * o.y + b.z * o.z + b.w * o.w;
  }
  sii = warp_sum(sii); sjj = warp_sum(sjj); soo = warp_sum(soo); sio = warp_sum(sio); sjo = warp_sum(sjo);
  if (lane == 0) {
    float ii = 1.f, ij = 1.f, io = 1.f;
    if (normalize) {
      ii = 1.f / fmaxf(sqrtf(sii), 1e-12f);
      ij = 1.f / fmaxf(sqrtf(sjj), 1e-12f);
      io = 1.f / fmaxf(sqrtf(soo), 1e-12f);
    }
    sv.inv_i[row] = ii; sv.inv_j[row] = ij; sv.inv_o[row] = io;
    sv.a[row] = sio * ii * io * inv_tau;
    sv.b[row] = sjo * ij * io * inv_tau;
  }
}

__device__ float block_reduce_1024(float v, bool is_max) {
  __shared__ float red[32];
  v = is_max ? warp_max(v) : warp_sum(v);
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  float t = lane < (blockDim.x >> 5) ? red[lane] : (is_max ? -INFINITY : 0.f);
  t = is_max ? warp_max(t) : warp_sum(t);
  return t;  // valid in every warp
}

// single block: softmax over the batch axis of a, log-softmax of b, quirk-KL (utils/losses.py:198-200)
// `gathered` (multi-GPU): the N-vectors are read from the all-gathered buffer [world][2][n_local] (n = world * n_local,
// rank-order concatenation = the single-process batch axis); otherwise from the saved blob.
struct RelicVec {
  const float* a;
  const float* b;
  int64_t n_local;  // 0: plain vectors
  __device__ __forceinline__ float A(int64_t i) const {
    if (!n_local) return a[i];
    const int64_t r = i / n_local;
    return a[r * 2 * n_local + (i - r * n_local)];
  }
  __device__ __forceinline__ float B(int64_t i) const {
    if (!n_local) return b[i];
    const int64_t r = i / n_local;
    return a[r * 2 * n_local + n_local + (i - r * n_local)];
  }
};
__global__ void relic_softmax_kernel(int64_t n, RelicVec v, RelicSaved sv, float alpha, float* kl_out,
                                     const float* add_in = nullptr /* may alias kl_out: *kl_out = *add_in + alpha KL */) {
  float ma = -INFINITY, mb = -INFINITY;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    ma = fmaxf(ma, v.A(i));
    mb = fmaxf(mb, v.B(i));
  }
  ma = block_reduce_1024(ma, true);
  mb = block_reduce_1024(mb, true);
  float sa = 0.f, sb = 0.f;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    sa += expf(v.A(i) - ma);
    sb += expf(v.B(i) - mb);
  }
  sa = block_reduce_1024(sa, false);
  sb = block_reduce_1024(sb, false);
  const float lse_a = ma + logf(sa), lse_b = mb + logf(sb);
  float kl = 0.f, spq = 0.f;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    const float p = expf(v.A(i) - lse_a);
    const float lq = v.B(i) - lse_b;
    const float q = expf(lq);
    kl += q * (lq - p);
    spq += p * q;
  }
  kl = block_reduce_1024(kl, false);
  spq = block_reduce_1024(spq, false);
  if (threadIdx.x == 0) {
    sv.scal[0] = lse_a; sv.scal[1] = lse_b; sv.scal[2] = spq; sv.scal[3] = kl;
    *kl_out = (add_in ? *add_in : 0.f) + alpha * kl;
  }
}

__global__ void relic_bwd_kernel(const float* __restrict__ zi, const float* __restrict__ zj,
                                 const float* __restrict__ zo, int64_t n, int d, int64_t ldi, int64_t ldj,
                                 int64_t ldo, int normalize, float inv_tau, float alpha,
                                 const float* __restrict__ grad_out, RelicSaved sv, float* __restrict__ dzi,
                                 float* __restrict__ dzj, float* __restrict__ dzo, int64_t lddi, int64_t lddj,
                                 int64_t lddo) {
  const int64_t row = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= n) return;
  const float lse_a = sv.scal[0], lse_b = sv.scal[1], spq = sv.scal[2], kl = sv.scal[3];
  const float a = sv.a[row], b = sv.b[row];
  const float p = expf(a - lse_a), lq = b - lse_b, q = expf(lq);
  const float g = __ldg(grad_out) * alpha * inv_tau;
  const float da = -(p * q - p * spq) * g;          // through softmax of the kl_div *input*
  const float h = q * (lq - p) + q;
  const float db = (h - q * (kl + 1.f)) * g;        // through log_softmax of the target; sum(h) = kl + 1
  const float ii = sv.inv_i[row], ij = sv.inv_j[row], io = sv.inv_o[row];
  // cosines (needed for the normalise-backward projections): a*tau, b*tau
  const float cio = a / inv_tau, cjo = b / inv_tau;
  const float4* pi = reinterpret_cast<const float4*>(zi + row * ldi);
  const float4* pj = reinterpret_cast<const float4*>(zj + row * ldj);
  const float4* po = reinterpret_cast<const float4*>(zo + row * ldo);
  float4* oi = reinterpret_cast<float4*>(dzi + row * lddi);
  float4* oj = reinterpret_cast<float4*>(dzj + row * lddj);
  float4* oo = reinterpret_cast<float4*>(dzo + row * lddo);
  for (int c = lane; c < d / 4; c += 32) {
    const float4 vi = __ldg(pi + c), vj = __ldg(pj + c), vo = __ldg(po + c);
    const float xi[4] = {vi.x * ii, vi.y * ii, vi.z * ii, vi.w * ii};
    const float xj[4] = {vj.x * ij, vj.y * ij, vj.z * ij, vj.w * ij};
    const float xo[4] = {vo.x * io, vo.y * io, vo.z * io, vo.w * io};
    float gi[4], gj[4], go[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      if (normalize) {
        // d zhat_i = da * zhat_o ; projected: (da*zo - (da*zo.zi) zi) * inv_i, with zo.zi = cos_io
        gi[e] = da * (xo[e] - cio * xi[e]) * ii;
        gj[e] = db * (xo[e] - cjo * xj[e]) * ij;
        // d zhat_o = da*zi + db*zj ; (g.zo) = da*cio + db*cjo
        go[e] = (da * xi[e] + db * xj[e] - (da * cio + db * cjo) * xo[e]) * io;
      } else {
        gi[e] = da * xo[e];
        gj[e] = db * xo[e];
        go[e] = da * xi[e] + db * xj[e];
      }
    }
    float4 ci = oi[c], cj = oj[c];
    ci.x += gi[0]; ci.y += gi[1]; ci.z += gi[2]; ci.w += gi[3];
    cj.x += gj[0]; cj.y += gj[1]; cj.z += gj[2]; cj.w += gj[3];
    oi[c] = ci;
    oj[c] = cj;
    oo[c] = make_float4(go[0], go[1], go[2], go[3]);
  }
}

struct RowdotWs {
  float* block_sums;
  unsigned int* counter;
  size_t bytes;
};
RowdotWs rowdot_ws(void* base) {
  Carver c(base);
  RowdotWs w;
  w.block_sums = c.take<float>(4096);
  w.counter = c.take<unsigned int>(4);
  w.bytes = c.used();
  return w;
}

}  // namespace

extern "C" {

size_t ssvb_rowdot_workspace_bytes(int64_t n, int64_t d) {
  (void)n; (void)d;
  return rowdot_ws(nullptr).bytes;
}

int ssvb_rowdot_fwd(int kind, const float* o, const float* t, int64_t n, int64_t d, int64_t ld_o, int64_t ld_t,
                    float* loss, void* workspace, size_t workspace_bytes, void* stream) {
  SSVB_TRY(check_device_sm100());
  if (n <= 0 || d <= 0 || !loss || !workspace || (kind != 0 && kind != 1)) return SSVB_ERR_INVALID;
  if (d % 4) return SSVB_ERR_ALIGNMENT;
  SSVB_TRY(check_rows(o, ld_o));
  SSVB_TRY(check_rows(t, ld_t));
  if (workspace_bytes < rowdot_ws(nullptr).bytes) return SSVB_ERR_WORKSPACE;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  RowdotWs ws = rowdot_ws(workspace);
  const int64_t total = n * (d / 4);
  int64_t grid = ceil_div(total, 256 * 4);
  const int64_t cap = static_cast<int64_t>(num_sms()) * 8;
  if (grid > cap) grid = cap;
  if (grid < 1) grid = 1;
  const bool flat = (ld_o == d) && (ld_t == d);
  const int d4 = static_cast<int>(d / 4);
  const float sc = kind == 0 ? 1.f / (static_cast<float>(n) * d) : -1.f / static_cast<float>(n);
  const unsigned g = static_cast<unsigned>(grid);
#define SSVB_RD(K, F) rowdot_fwd_kernel<K, F><<<g, 256, 0, s>>>(o, t, n, d4, ld_o, ld_t, ws.block_sums, nullptr, sc, loss)
  if (kind == 0) { if (flat) SSVB_RD(0, true); else SSVB_RD(0, false); }
  else           { if (flat) SSVB_RD(1, true); else SSVB_RD(1, false); }
#undef SSVB_RD
  SSVB_LAUNCH_CHECK();
  block_sums_finish_kernel<<<1, 256, 0, s>>>(ws.block_sums, static_cast<int>(grid), sc, loss);
  SSVB_LAUNCH_CHECK();
  return SSVB_OK;
}

size_t ssvb_rowdot_norm_saved_bytes(int64_t n) { return n > 0 ? rowdot_norm_saved(nullptr, n).bytes : 0; }

int ssvb_rowdot_norm_fwd(int kind, const float* o, const float* t, int64_t n, int64_t d, int64_t ld_o, int64_t ld_t,
                         int normalize_o, int normalize_t, float* loss, void* saved, void* workspace,
                         size_t workspace_bytes, void* stream) {
  SSVB_TRY(check_device_sm100());
  if (n <= 0 || d <= 0 || !loss || !saved || !workspace || (kind != 0 && kind != 1)) return SSVB_ERR_INVALID;
  if (d % 4) return SSVB_ERR_ALIGNMENT;
  SSVB_TRY(check_rows(o, ld_o));
  SSVB_TRY(check_rows(t, ld_t));
  if (workspace_bytes < rowdot_ws(nullptr).bytes) return SSVB_ERR_WORKSPACE;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  RowdotWs ws = rowdot_ws(workspace);
  RowdotNormSaved sv = rowdot_norm_saved(saved, n);
  int64_t grid = ceil_div(n, 8);
  const int64_t cap = static_cast<int64_t>(num_sms()) * 8;
  if (grid > cap) grid = cap;
  const int d4 = static_cast<int>(d / 4);
  if (kind == 0)
    rowdot_norm_fwd_kernel<0><<<static_cast<unsigned>(grid), 256, 0, s>>>(
        o, t, n, d4, ld_o, ld_t, normalize_o, normalize_t, sv, ws.block_sums, nullptr,
        1.f / (static_cast<float>(n) * d), loss);
  else
    rowdot_norm_fwd_kernel<1><<<static_cast<unsigned>(grid), 256, 0, s>>>(
        o, t, n, d4, ld_o, ld_t, normalize_o, normalize_t, sv, ws.block_sums, nullptr, -1.f / static_cast<float>(n),
        loss);
  SSVB_LAUNCH_CHECK();
  block_sums_finish_kernel<<<1, 256, 0, s>>>(ws.block_sums, static_cast<int>(grid),
                                             kind == 0 ? 1.f / (static_cast<float>(n) * d) : -1.f / static_cast<float>(n),
                                             loss);
  SSVB_LAUNCH_CHECK();
  return SSVB_OK;
}

int ssvb_rowdot_norm_bwd(int kind, const float* o, const float* t, int64_t n, int64_t d, int64_t ld_o, int64_t ld_t,
                         int normalize_o, int normalize_t, const float* grad_out, const void* saved, float* d_o,
                         float* d_t, int64_t ld_do, int64_t ld_dt, void* stream) {
  SSVB_TRY(check_device_sm100());
  if (n <= 0 || d <= 0 || !grad_out || !saved || (kind != 0 && kind != 1) || (!d_o && !d_t)) return SSVB_ERR_INVALID;
  if (d % 4) return SSVB_ERR_ALIGNMENT;
  SSVB_TRY(check_rows(o, ld_o));
  SSVB_TRY(check_rows(t, ld_t));
  if (d_o) SSVB_TRY(check_rows(d_o, ld_do));
  if (d_t) SSVB_TRY(check_rows(d_t, ld_dt));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  RowdotNormSaved sv = rowdot_norm_saved(const_cast<void*>(saved), n);
  int64_t grid = ceil_div(n, 8);
  const int64_t cap = static_cast<int64_t>(num_sms()) * 8;
  if (grid > cap) grid = cap;
  const int d4 = static_cast<int>(d / 4);
  if (kind == 0)
    rowdot_norm_bwd_kernel<0><<<static_cast<unsigned>(grid), 256, 0, s>>>(
        o, t, n, d4, ld_o, ld_t, normalize_o, normalize_t, sv, grad_out, 1.f / (static_cast<float>(n) * d), d_o, d_t,
        ld_do, ld_dt);
  else
    rowdot_norm_bwd_kernel<1><<<static_cast<unsigned>(grid), 256, 0, s>>>(
        o, t, n, d4, ld_o, ld_t, normalize_o, normalize_t, sv, grad_out, -1.f / static_cast<float>(n), d_o, d_t, ld_do,
        ld_dt);
  SSVB_LAUNCH_CHECK();
  return SSVB_OK;
}

int ssvb_rowdot_bwd(int kind, const float* o, const float* t, int64_t n, int64_t d, int64_t ld_o, int64_t ld_t,
                    const float* grad_out, float* d_o, float* d_t, int64_t ld_do, int64_t ld_dt, void* stream) {
  SSVB_TRY(check_device_sm100());
  if (n <= 0 || d <= 0 || !grad_out || (kind != 0 && kind != 1) || (!d_o && !d_t)) return SSVB_ERR_INVALID;
  if (d % 4) return SSVB_ERR_ALIGNMENT;
  SSVB_TRY(check_rows(o, ld_o));
  SSVB_TRY(check_rows(t, ld_t));
  if (d_o) SSVB_TRY(check_rows(d_o, ld_do));
  if (d_t) SSVB_TRY(check_rows(d_t, ld_dt));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int64_t total = n * (d / 4);
  int64_t grid = ceil_div(total, 256 * 4);
  const int64_t cap = static_cast<int64_t>(num_sms()) * 8;
  if (grid > cap) grid = cap;
  if (grid < 1) grid = 1;
  const bool flat = (ld_o == d) && (ld_t == d) && (!d_o || ld_do == d) && (!d_t || ld_dt == d);
  const int d4 = static_cast<int>(d / 4);
  const float cf = kind == 0 ? 2.f / (static_cast<float>(n) * d) : -1.f / static_cast<float>(n);
  const unsigned g = static_cast<unsigned>(grid);
#define SSVB_RD(K, F) \
  rowdot_bwd_kernel<K, F><<<g, 256, 0, s>>>(o, t, n, d4, ld_o, ld_t, grad_out, cf, d_o, d_t, ld_do, ld_dt)
  if (kind == 0) { if (flat) SSVB_RD(0, true); else SSVB_RD(0, false); }
  else           { if (flat) SSVB_RD(1, true); else SSVB_RD(1, false); }
#undef SSVB_RD
  SSVB_LAUNCH_CHECK();
  return SSVB_OK;
}

int ssvb_l2norm_fwd(const float* x, int64_t n, int64_t d, int64_t ld_x, float* y, int64_t ld_y, float* inv_norm,
                    void* stream) {
  SSVB_TRY(check_device_sm100());
  if (n <= 0 || d <= 0) return SSVB_ERR_INVALID;
  if (d % 4) return SSVB_ERR_ALIGNMENT;
  SSVB_TRY(check_rows(x, ld_x));
  SSVB_TRY(check_rows(y, ld_y));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  l2norm_fwd_kernel<<<static_cast<unsigned>(ceil_div(n, 8)), 256, 0, s>>>(x, n, static_cast<int>(d), ld_x, y, ld_y,
                                                                          inv_norm);
  SSVB_LAUNCH_CHECK();
  return SSVB_OK;
}

int ssvb_l2norm_bwd(const float* dy, const float* y, const float* inv_norm, int64_t n, int64_t d, int64_t ld_dy,
                    int64_t ld_y, float* dx, int64_t ld_dx, void* stream) {
  SSVB_TRY(check_device_sm100());
  if (n <= 0 || d <= 0 || !inv_norm) return SSVB_ERR_INVALID;
  if (d % 4) return SSVB_ERR_ALIGNMENT;
  SSVB_TRY(check_rows(dy, ld_dy));
  SSVB_TRY(check_rows(y, ld_y));
  SSVB_TRY(check_rows(dx, ld_dx));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  l2norm_bwd_kernel<<<static_cast<unsigned>(ceil_div(n, 8)), 256, 0, s>>>(dy, y, inv_norm, n, static_cast<int>(d),
                                                                          ld_dy, ld_y, dx, ld_dx);
  SSVB_LAUNCH_CHECK();
  return SSVB_OK;
}

int ssvb_ring_enqueue(float* bank, void* bank_bf16, int64_t size, int64_t d, int64_t ld_bank, const float* batch,
                      int64_t n, int64_t ld_batch, int64_t ptr, int normalize, int64_t* new_ptr, void* stream) {
  SSVB_TRY(check_device_sm100());
  if (size <= 0 || d <= 0 || n < 0 || ptr < 0 || ptr >= size || !new_ptr) return SSVB_ERR_INVALID;
  if (d % 4) return SSVB_ERR_ALIGNMENT;
  SSVB_TRY(check_rows(bank, ld_bank));
  if (n == 0) {
    *new_ptr = ptr;
    return SSVB_OK;
  }
  SSVB_TRY(check_rows(batch, ld_batch));
  if (bank_bf16 && (reinterpret_cast<uintptr_t>(bank_bf16) & 15)) return SSVB_ERR_ALIGNMENT;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  ring_enqueue_kernel<<<static_cast<unsigned>(ceil_div(n, 8)), 256, 0, s>>>(
      bank, static_cast<__nv_bfloat16*>(bank_bf16), size, static_cast<int>(d), static_cast<int>(sim_dpad(d)), ld_bank,
      batch, n, ld_batch, ptr, normalize, 0, size);
  SSVB_LAUNCH_CHECK();
  *new_ptr = (ptr + n) % size;
  return SSVB_OK;
}

// sharded ring (MoCo queue range-partitioned over ranks): `bank_shard` holds global rows
// [shard_lo, shard_lo + shard_rows) of a ring of `size` rows; `batch` is the GLOBAL batch (all ranks' keys in rank
// order, all-gathered by the caller); only the rows whose slot falls into this shard are written.  ptr arithmetic is
// the single-process ring's, identical on every rank.
int ssvb_ring_enqueue_shard(float* bank_shard, void* bank_shard_bf16, int64_t size, int64_t shard_lo, int64_t shard_rows,
                            int64_t d, int64_t ld_bank, const float* batch, int64_t n, int64_t ld_batch, int64_t ptr,
                            int normalize, int64_t* new_ptr, void* stream) {
  SSVB_TRY(check_device_sm100());
  if (size <= 0 || d <= 0 || n < 0 || ptr < 0 || ptr >= size || !new_ptr || shard_lo < 0 || shard_rows <= 0 ||
      shard_lo + shard_rows > size)
    return SSVB_ERR_INVALID;
  if (d % 4) return SSVB_ERR_ALIGNMENT;
  SSVB_TRY(check_rows(bank_shard, ld_bank));
  if (n == 0) {
    *new_ptr = ptr;
    return SSVB_OK;
  }
  SSVB_TRY(check_rows(batch, ld_batch));
  if (bank_shard_bf16 && (reinterpret_cast<uintptr_t>(bank_shard_bf16) & 15)) return SSVB_ERR_ALIGNMENT;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  ring_enqueue_kernel<<<static_cast<unsigned>(ceil_div(n, 8)), 256, 0, s>>>(
      bank_shard, static_cast<__nv_bfloat16*>(bank_shard_bf16), size, static_cast<int>(d), static_cast<int>(sim_dpad(d)),
      ld_bank, batch, n, ld_batch, ptr, normalize, shard_lo, shard_rows);
  SSVB_LAUNCH_CHECK();
  *new_ptr = (ptr + n) % size;
  return SSVB_OK;
}

size_t ssvb_relic_kl_saved_bytes(int64_t n) { return n > 0 ? relic_saved(nullptr, n).bytes : 0; }
size_t ssvb_relic_kl_workspace_bytes(int64_t n) {
  (void)n;
  return 256;
}

int ssvb_relic_kl_fwd(const float* zi, const float* zj, const float* zo, int64_t n, int64_t d, int64_t ld_zi,
                      int64_t ld_zj, int64_t ld_zo, int normalize, float temperature, float alpha, float* kl,
                      void* saved, void* workspace, size_t workspace_bytes, void* stream) {
  (void)workspace; (void)workspace_bytes;
  SSVB_TRY(check_device_sm100());
  if (n <= 0 || d <= 0 || !kl || !saved || !(temperature > 0.f)) return SSVB_ERR_INVALID;
  if (d % 4) return SSVB_ERR_ALIGNMENT;
  SSVB_TRY(check_rows(zi, ld_zi));
  SSVB_TRY(check_rows(zj, ld_zj));
  SSVB_TRY(check_rows(zo, ld_zo));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  RelicSaved sv = relic_saved(saved, n);
  relic_dots_kernel<<<static_cast<unsigned>(ceil_div(n, 8)), 256, 0, s>>>(zi, zj, zo, n, static_cast<int>(d), ld_zi,
                                                                          ld_zj, ld_zo, normalize, 1.f / temperature,
                                                                          sv);
  SSVB_LAUNCH_CHECK();
  relic_softmax_kernel<<<1, 1024, 0, s>>>(n, RelicVec{sv.a, sv.b, 0}, sv, alpha, kl);
  SSVB_LAUNCH_CHECK();
  return SSVB_OK;
}

// The whole RelicLoss.forward / backward in ONE call each (utils/losses.py:162-201): NT-Xent on (zi, zj), the KL term on
// top, `loss` = contrastive + alpha * KL written by the last kernel (no host-side add, one library crossing).
int ssvb_relic_fwd(const float* zi, const float* zj, const float* zo, int64_t n, int64_t d, int64_t ld_zi, int64_t ld_zj,
                   int64_t ld_zo, int normalize, float temperature, float alpha, float* loss, void* saved_ntxent,
                   void* saved_kl, void* workspace, size_t workspace_bytes, void* stream) {
  if (!saved_kl) return SSVB_ERR_INVALID;
  SSVB_TRY(check_rows(zo, ld_zo));
  SSVB_TRY(ssvb_ntxent_fwd(zi, zj, n, d, ld_zi, ld_zj, normalize, temperature, loss, saved_ntxent, workspace,
                           workspace_bytes, stream));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  RelicSaved sv = relic_saved(saved_kl, n);
  relic_dots_kernel<<<static_cast<unsigned>(ceil_div(n, 8)), 256, 0, s>>>(zi, zj, zo, n, static_cast<int>(d), ld_zi,
                                                                          ld_zj, ld_zo, normalize, 1.f / temperature,
                                                                          sv);
  SSVB_LAUNCH_CHECK();
  relic_softmax_kernel<<<1, 1024, 0, s>>>(n, RelicVec{sv.a, sv.b, 0}, sv, alpha, loss, loss);
  SSVB_LAUNCH_CHECK();
  return SSVB_OK;
}

// ---- multi-GPU ReLIC-KL: the KL's two softmaxes run over the BATCH axis (utils/losses.py:196-200), so the
// per-row logits a_n, b_n of all ranks are all-gathered (2 * n_local floats per rank) between two stages:
//   dist_dots   : this rank's a, b (+ inverse norms) into `saved`, and into ab_local [2][n_local] for the all-gather
//   dist_reduce : global softmax statistics + alpha * KL from the gathered [world][2][n_local] buffer (fixed order,
//                 identical on every rank); the statistics land in this rank's `saved`, so the unchanged
//                 ssvb_relic_kl_bwd produces the gradient rows of this rank's inputs.
int ssvb_relic_kl_dist_dots(const float* zi, const float* zj, const float* zo, int64_t n_local, int64_t d,
                            int64_t ld_zi, int64_t ld_zj, int64_t ld_zo, int normalize, float temperature,
                            void* saved, float* ab_local, void* stream) {
  SSVB_TRY(check_device_sm100());
  if (n_local <= 0 || d <= 0 || !saved || !ab_local || !(temperature > 0.f)) return SSVB_ERR_INVALID;
  if (d % 4) return SSVB_ERR_ALIGNMENT;
  SSVB_TRY(check_rows(zi, ld_zi));
  SSVB_TRY(check_rows(zj, ld_zj));
  SSVB_TRY(check_rows(zo, ld_zo));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  RelicSaved sv = relic_saved(saved, n_local);
  relic_dots_kernel<<<static_cast<unsigned>(ceil_div(n_local, 8)), 256, 0, s>>>(
      zi, zj, zo, n_local, static_cast<int>(d), ld_zi, ld_zj, ld_zo, normalize, 1.f / temperature, sv);
  SSVB_LAUNCH_CHECK();
  SSVB_CUDA(cudaMemcpyAsync(ab_local, sv.a, n_local * sizeof(float), cudaMemcpyDeviceToDevice, s));
  SSVB_CUDA(cudaMemcpyAsync(ab_local + n_local, sv.b, n_local * sizeof(float), cudaMemcpyDeviceToDevice, s));
  return SSVB_OK;
}
int ssvb_relic_kl_dist_reduce(const float* ab_all, int64_t world, int64_t n_local, float alpha, void* saved, float* kl,
                              void* stream) {
  SSVB_TRY(check_device_sm100());
  if (!ab_all || world <= 0 || n_local <= 0 || !saved || !kl) return SSVB_ERR_INVALID;
  RelicSaved sv = relic_saved(saved, n_local);
  relic_softmax_kernel<<<1, 1024, 0, static_cast<cudaStream_t>(stream)>>>(world * n_local, RelicVec{ab_all, ab_all, n_local},
                                                                          sv, alpha, kl);
  SSVB_LAUNCH_CHECK();
  return SSVB_OK;
}

int ssvb_relic_kl_bwd(const float* zi, const float* zj, const float* zo, int64_t n, int64_t d, int64_t ld_zi,
                      int64_t ld_zj, int64_t ld_zo, int normalize, float temperature, float alpha,
                      const float* grad_out, const void* saved, float* dzi, float* dzj, float* dzo, int64_t ld_dzi,
                      int64_t ld_dzj, int64_t ld_dzo, void* stream) {
  SSVB_TRY(check_device_sm100());
  if (n <= 0 || d <= 0 || !grad_out || !saved || !(temperature > 0.f)) return SSVB_ERR_INVALID;
  if (d % 4) return SSVB_ERR_ALIGNMENT;
  SSVB_TRY(check_rows(zi, ld_zi));
  SSVB_TRY(check_rows(zj, ld_zj));
  SSVB_TRY(check_rows(zo, ld_zo));
  SSVB_TRY(check_rows(dzi, ld_dzi));
  SSVB_TRY(check_rows(dzj, ld_dzj));
  SSVB_TRY(check_rows(dzo, ld_dzo));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  RelicSaved sv = relic_saved(const_cast<void*>(saved), n);
  relic_bwd_kernel<<<static_cast<unsigned>(ceil_div(n, 8)), 256, 0, s>>>(
      zi, zj, zo, n, static_cast<int>(d), ld_zi, ld_zj, ld_zo, normalize, 1.f / temperature, alpha, grad_out, sv, dzi,
      dzj, dzo, ld_dzi, ld_dzj, ld_dzo);
  SSVB_LAUNCH_CHECK();
  return SSVB_OK;
}

int ssvb_bank_scatter(float* bank, int64_t size, int64_t d, int64_t ld_bank, const int64_t* indices, int64_t n,
                      const float* vectors, int64_t ld_vectors, float momentum, float one_minus_m, int mode,
                      void* stream) {
  SSVB_TRY(check_device_sm100());
  if (size <= 0 || d <= 0 || n < 0 || (mode != 0 && mode != 1)) return SSVB_ERR_INVALID;
  if (d % 4) return SSVB_ERR_ALIGNMENT;
  SSVB_TRY(check_rows(bank, ld_bank));
  if (n == 0) return SSVB_OK;
  if (!indices) return SSVB_ERR_INVALID;
  SSVB_TRY(check_rows(vectors, ld_vectors));
  bank_scatter_kernel<<<static_cast<unsigned>(ceil_div(n, 8)), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      bank, size, static_cast<int>(d), ld_bank, reinterpret_cast<const long long*>(indices), n, vectors, ld_vectors,
      momentum, one_minus_m, mode);
  SSVB_LAUNCH_CHECK();
  return SSVB_OK;
}

int ssvb_bank_gather(const float* bank, int64_t size, int64_t d, int64_t ld_bank, const int64_t* indices, int64_t n,
                     float* out, int64_t ld_out, void* stream) {
  SSVB_TRY(check_device_sm100());
  if (size <= 0 || d <= 0 || n < 0) return SSVB_ERR_INVALID;
  if (d % 4) return SSVB_ERR_ALIGNMENT;
  SSVB_TRY(check_rows(bank, ld_bank));
  if (n == 0) return SSVB_OK;
  if (!indices) return SSVB_ERR_INVALID;
  SSVB_TRY(check_rows(out, ld_out));
  bank_gather_kernel<<<static_cast<unsigned>(ceil_div(n, 8)), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      bank, size, static_cast<int>(d), ld_bank, reinterpret_cast<const long long*>(indices), n, out, ld_out);
  SSVB_LAUNCH_CHECK();
  return SSVB_OK;
}

int ssvb_relic_bwd(const float* zi, const float* zj, const float* zo, int64_t n, int64_t d, int64_t ld_zi, int64_t ld_zj,
                   int64_t ld_zo, int normalize, float temperature, float alpha, const float* grad_out,
                   const void* saved_ntxent, const void* saved_kl, float* dzi, float* dzj, float* dzo, int64_t ld_dzi,
                   int64_t ld_dzj, int64_t ld_dzo, void* workspace, size_t workspace_bytes, void* stream) {
  SSVB_TRY(ssvb_ntxent_bwd(zi, zj, n, d, ld_zi, ld_zj, normalize, temperature, grad_out, saved_ntxent, dzi, dzj, ld_dzi,
                           ld_dzj, workspace, workspace_bytes, stream));
  return ssvb_relic_kl_bwd(zi, zj, zo, n, d, ld_zi, ld_zj, ld_zo, normalize, temperature, alpha, grad_out, saved_kl, dzi,
                           dzj, dzo, ld_dzi, ld_dzj, ld_dzo, stream);
}

}  // extern "C"
