// Persistent warp-specialised tcgen05 GEMM with fused epilogues:  C[m, n] = sum_k A(m, k) * B(n, k)
//   bf16 operands via TMA (128-byte swizzle), fp32 accumulators double-buffered in TMEM (2 x BN columns) so the
//   epilogue of tile i overlaps the main loop of tile i+1.
//   Each operand is either K-major (global [rows x K], K contiguous) or MN-major (global [K x rows], rows
//   contiguous) — both are consumed in place through the matching UMMA shared-memory descriptor, so no
//   transposed copies are ever materialised:
//     Barlow  C = Xi^T Xj         : A MN-major (Xi [n x D]), B MN-major (Xj [n x D])
//     Barlow  dXi = Xj dC^T       : A K-major,  B K-major (dC [a x b])
//     Barlow  dXj = Xi dC         : A K-major,  B MN-major
//     SwAV    scores = z C^T      : A K-major,  B K-major;   dz = ds C : B MN-major;   dC = ds^T z : both MN-major
// Warp roles (192 threads): warp 0 TMA producer, warp 1 MMA issuer, warps 2..5 epilogue (TMEM lane quarter = warp%4).
//
// Epilogue: every epilogue warp drains its 32 accumulator rows 32 columns at a time (tcgen05.ld), stages the
// [32 rows x 128 bytes] block in shared memory in the 128-byte-swizzle layout (conflict-free 16-byte stores: the chunk
// index is XORed with row & 7) and hands it to the TMA unit (`cp.async.bulk.tensor` store, two staging buffers per
// warp): full 128-byte lines leave the SM instead of 32 scattered 16-byte row pieces per store instruction, and the
// ragged M / N tails are clipped by the tensor map.  The same path with `cp.reduce.async.bulk.tensor ... .add` is the
// split-K epilogue (work unit = tile x K-slice; partial products are added into a zeroed output), used when a GEMM has
// fewer tiles than SMs (SwAV's prototype / embedding gradients: 24 and 55 tiles for 148 SMs).
// DUAL: two problems of identical shape (second one with B MN-major) share one launch and one persistent tile queue -
// Barlow's two backward GEMMs are 512 tiles each = 3.46 waves of 148 CTAs separately, 6.92 waves together.
// Fused column partials (Barlow backward): while a block sits in the staging buffer each lane owns one column and
// accumulates sum_r out[r, c] and sum_r out[r, c] * x~[r, c] over the 32 rows (conflict-free shared loads, coalesced
// bf16 loads of x~), one partial row per (row tile, warp) - the separate column-reduction pass over dT is gone.
#pragma once
#include "common.cuh"

namespace ssvb {

enum GemmEpiMode { EPI_STORE_F32 = 0, EPI_BARLOW = 1, EPI_BARLOW_BWD = 2 };

struct GemmParams {
  int M, N, K;  // logical sizes (tails are zero-filled by TMA and predicated / clipped in the epilogue)
  int tiles_m, tiles_n;
  int splits, kb_per_split;  // split-K: unit = (tile, K slice of kb_per_split 64-wide blocks); splits > 1 -> add epilogue
  int tma_store;             // 1: staged TMA-store epilogue (needs 16-byte aligned rows); 0: per-thread row stores
  // EPI_STORE_F32: out[m * ldc + n] = alpha * acc
  float alpha;
  float* out;
  float* out2;  // DUAL: second problem's output (same ldc)
  int64_t ldc;
  // optional fused column partials (EPI_STORE_F32, tma_store, splits == 1): colpart[(t * 2 + {0,1}) * N + n] with
  // t = tile_m * 4 + lane quarter; xt = bf16 [M x ldx] multiplied into the second sum
  float* colpart;
  float* colpart2;
  const __nv_bfloat16* xt;
  const __nv_bfloat16* xt2;
  int64_t ldx;
  // EPI_BARLOW: c = acc * alpha; loss += (m==n) ? (c-1)^2 : lambda c^2 ; dC = (m==n) ? 2(c-1) : 2 lambda c  (bf16)
  float lambda;
  __nv_bfloat16* dC;
  int64_t ld_dc;
  float* loss_partials;  // [gridDim.x]
  int diag_off;          // the output is a column slab of the full matrix: global column = n + diag_off
  // EPI_BARLOW, optional (nullptr = off): partial sums of dC .* C for the closed-form backward of the standardisation
  // (csrc/barlow.cu): bl_rowpart[tile_n * M + m] = sum over the tile's columns, bl_colpart[(tile_m * 4 + q) * N + n] = sum
  // over the 32 rows of epilogue warp q
  float* bl_rowpart;
  float* bl_colpart;
  // EPI_BARLOW_BWD (staged TMA-store epilogue only): out[m, n] = (acc * alpha - (xf[m, n] - vm[n]) * vq[n]) * vr[n] * go[0],
  // i.e. the standardisation backward applied to the dT tile while it leaves TMEM (xf = the fp32 input rows, vm / vr =
  // column mean / 1/std, vq = rstd * sum_n(dT x~)/(n-1)); second problem: xf2 / ldxf2 / vm2 / vr2 / vq2 / out2 / ldc2
  const float *xf, *xf2;
  int64_t ldxf, ldxf2;
  const float *vm, *vr, *vq, *vm2, *vr2, *vq2;
  const float* go;
  int64_t ldc2;
};

template <int BN>
struct GemmCfg {
  static constexpr int BM = 128;
  static constexpr int BK = 64;
  static constexpr int A_BYTES = BM * BK * 2;  // 16 KB
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int NSTAGE = (BN == 256) ? 4 : 6;
  static constexpr int STAGE_BYTES = 4 * 2 * 4096;  // 4 epilogue warps x 2 staging buffers x [32 rows x 128 B]
  static constexpr int NBARS = 2 * NSTAGE + 4;
  static constexpr int SMEM = 1024 + NSTAGE * (A_BYTES + B_BYTES) + STAGE_BYTES + NBARS * 8 + 16;
};

__device__ __forceinline__ void tmem_ld_wait_regs32(uint32_t (&v)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]),
                 "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]),
                 "+r"(v[15]), "+r"(v[16]), "+r"(v[17]), "+r"(v[18]), "+r"(v[19]), "+r"(v[20]), "+r"(v[21]),
                 "+r"(v[22]), "+r"(v[23]), "+r"(v[24]), "+r"(v[25]), "+r"(v[26]), "+r"(v[27]), "+r"(v[28]),
                 "+r"(v[29]), "+r"(v[30]), "+r"(v[31])::"memory");
}

// ---- TMA store side (shared -> global through a tensor map; bulk async-groups are per issuing thread) ----
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, uint32_t src_smem, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(src_smem), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* m, uint32_t src_smem, int c0, int c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(src_smem), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void bulk_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_group_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void bulk_wait_group_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void sts_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ float lds_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
  return v;
}

template <int BN, bool A_MN, bool B_MN, int EPI, bool DUAL>
__global__ void __launch_bounds__(192, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
            const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmA2,
            const __grid_constant__ CUtensorMap tmB2, const __grid_constant__ CUtensorMap tmC2, const GemmParams p) {
  using C = GemmCfg<BN>;
  constexpr int NSTAGE = C::NSTAGE;
  constexpr bool B2_MN = !B_MN;  // DUAL: the second problem consumes its B operand in the other major-ness
  extern __shared__ uint8_t gemm_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(gemm_smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + NSTAGE * C::A_BYTES;
  uint8_t* sStage = sB + NSTAGE * C::B_BYTES;  // 1024-byte aligned (all sizes above are multiples of 1024)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sStage + C::STAGE_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + NSTAGE;
  uint64_t* acc_full = empty + NSTAGE;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  __shared__ float loss_red[4];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (p.tma_store) tma_prefetch_desc(&tmC);
    if (DUAL) {
      tma_prefetch_desc(&tmA2);
      tma_prefetch_desc(&tmB2);
      tma_prefetch_desc(&tmC2);
    }
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < NSTAGE; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], 4);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<2 * BN>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int ntiles1 = p.tiles_m * p.tiles_n;            // tiles of one problem
  const int ntiles = DUAL ? 2 * ntiles1 : ntiles1;      // tile index >= ntiles1 -> second problem
  const int nunits = ntiles * p.splits;                 // unit = split * ntiles + tile
  const int nkb = (p.K + C::BK - 1) / C::BK;

  // producer / issuer loops run on the whole warp (uniform operands stay in uniform registers); elect_one() issues
  if (warp == 0) {
    int it = 0;
    for (int unit = blockIdx.x; unit < nunits; unit += gridDim.x) {
      const int split = unit / ntiles;
      int tile = unit - split * ntiles;
      const bool second = DUAL && tile >= ntiles1;
      if (second) tile -= ntiles1;
      const int tm = tile % p.tiles_m, tn = tile / p.tiles_m;
      const int kb0 = split * p.kb_per_split;
      const int kb1 = min(nkb, kb0 + p.kb_per_split);
      const CUtensorMap* ma = second ? &tmA2 : &tmA;
      const CUtensorMap* mb = second ? &tmB2 : &tmB;
      for (int kb = kb0; kb < kb1; ++kb, ++it) {
        const int st = it % NSTAGE;
        mbar_wait(&empty[st], ((it / NSTAGE) & 1) ^ 1);
        if (elect_one()) {
          mbar_expect_tx(&full[st], C::A_BYTES + C::B_BYTES);
          uint8_t* a = sA + st * C::A_BYTES;
          uint8_t* b = sB + st * C::B_BYTES;
          if (A_MN) {
#pragma unroll
            for (int blk = 0; blk < C::BM / 64; ++blk)
              tma_load_2d(a + blk * 8192, ma, &full[st], tm * C::BM + blk * 64, kb * C::BK);
          } else {
            tma_load_2d(a, ma, &full[st], kb * C::BK, tm * C::BM);
          }
          if (second ? B2_MN : B_MN) {
#pragma unroll
            for (int blk = 0; blk < BN / 64; ++blk)
              tma_load_2d(b + blk * 8192, mb, &full[st], tn * BN + blk * 64, kb * C::BK);
          } else {
            tma_load_2d(b, mb, &full[st], kb * C::BK, tn * BN);
          }
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t IDESC = make_idesc(C::BM, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
    constexpr uint32_t IDESC2 = make_idesc(C::BM, BN, A_MN ? 1 : 0, B2_MN ? 1 : 0);
    int it = 0, tcount = 0;
    for (int unit = blockIdx.x; unit < nunits; unit += gridDim.x, ++tcount) {
      const int split = unit / ntiles;
      const int tile = unit - split * ntiles;
      const bool second = DUAL && tile >= ntiles1;
      const bool bmn = second ? B2_MN : B_MN;
      const uint32_t idesc = second ? IDESC2 : IDESC;
      const int kb0 = split * p.kb_per_split;
      const int kb1 = min(nkb, kb0 + p.kb_per_split);
      const int ab = tcount & 1;
      mbar_wait(&acc_empty[ab], ((tcount >> 1) & 1) ^ 1);
      tc_fence_after();
      for (int kb = kb0; kb < kb1; ++kb, ++it) {
        const int st = it % NSTAGE;
        mbar_wait(&full[st], (it / NSTAGE) & 1);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t abase = smem_u32(sA + st * C::A_BYTES), bbase = smem_u32(sB + st * C::B_BYTES);
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4) {
            const uint64_t ad = A_MN ? desc_mnmajor(abase + k4 * 2048, 8192) : desc_kmajor(abase + k4 * 32);
            const uint64_t bd = bmn ? desc_mnmajor(bbase + k4 * 2048, 8192) : desc_kmajor(bbase + k4 * 32);
            umma_ss(tmem + ab * BN, ad, bd, idesc, (kb != kb0) || (k4 != 0));
          }
          umma_commit(&empty[st]);
          if (kb == kb1 - 1) umma_commit(&acc_full[ab]);
        }
        __syncwarp();
      }
    }
  } else {
    const int q = warp & 3;
    const int row_l = q * 32 + lane;
    const uint32_t tlane = static_cast<uint32_t>(q * 32) << 16;
    const uint32_t stage0 = smem_u32(sStage) + static_cast<uint32_t>(q) * 8192u;  // this warp's two 4 KB buffers
    const uint32_t my_row = stage0 + static_cast<uint32_t>(lane) * 128u;        // + buffer * 4096 + swizzled chunk
    const uint32_t swz = static_cast<uint32_t>(lane & 7);
    int nstore = 0;  // staged blocks issued so far by this warp (buffer = nstore & 1)
    float loss_acc = 0.f;
    int tcount = 0;
    for (int unit = blockIdx.x; unit < nunits; unit += gridDim.x, ++tcount) {
      const int split = unit / ntiles;
      int tile = unit - split * ntiles;
      const bool second = DUAL && tile >= ntiles1;
      if (second) tile -= ntiles1;
      const int tm = tile % p.tiles_m, tn = tile / p.tiles_m;
      const int ab = tcount & 1;
      const int row = tm * C::BM + row_l;
      const bool row_ok = row < p.M;
      const CUtensorMap* mc = second ? &tmC2 : &tmC;
      mbar_wait(&acc_full[ab], (tcount >> 1) & 1);
      tc_fence_after();
      if (EPI == EPI_STORE_F32) {
        float* outp = second ? p.out2 : p.out;
        float* cpart = second ? p.colpart2 : p.colpart;
        const __nv_bfloat16* xt = second ? p.xt2 : p.xt;
#pragma unroll 1
        for (int cc = 0; cc < BN / 32; ++cc) {
          uint32_t v[32];
          tmem_ld_x32(tmem + tlane + ab * BN + cc * 32, v);
          tmem_ld_wait_regs32(v);
          const int col0 = tn * BN + cc * 32;
          if (col0 >= p.N) continue;  // warp-uniform
          if (p.tma_store) {
            const uint32_t buf = my_row + static_cast<uint32_t>(nstore & 1) * 4096u;
            if (nstore >= 2) {  // the store that last read this buffer (two blocks ago) must have drained it
              if (lane == 0) bulk_wait_group_read<1>();
              __syncwarp();
            }
#pragma unroll
            for (int i = 0; i < 8; ++i)
              sts_v4(buf + ((static_cast<uint32_t>(i) ^ swz) << 4),
                     __float_as_uint(__uint_as_float(v[4 * i]) * p.alpha),
                     __float_as_uint(__uint_as_float(v[4 * i + 1]) * p.alpha),
                     __float_as_uint(__uint_as_float(v[4 * i + 2]) * p.alpha),
                     __float_as_uint(__uint_as_float(v[4 * i + 3]) * p.alpha));
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
              const uint32_t src = stage0 + static_cast<uint32_t>(nstore & 1) * 4096u;
              if (p.splits > 1)
                tma_reduce_add_2d(mc, src, col0, tm * C::BM + q * 32);
              else
                tma_store_2d(mc, src, col0, tm * C::BM + q * 32);
              bulk_commit_group();
            }
            if (cpart != nullptr) {
              // lane <-> column col0 + lane: sum over the 32 staged rows of out and out * x~ (rows beyond M hold zeros)
              const int col = col0 + lane;
              const uint32_t base = stage0 + static_cast<uint32_t>(nstore & 1) * 4096u + static_cast<uint32_t>(lane & 3) * 4u;
              const uint32_t chunk = static_cast<uint32_t>(lane >> 2);
              const int r0 = tm * C::BM + q * 32;
              float s1 = 0.f, s2 = 0.f;
              if (col < p.N) {
                const __nv_bfloat16* xp = xt + static_cast<int64_t>(r0) * p.ldx + col;
                float xv[32];
#pragma unroll
                for (int r = 0; r < 32; ++r)
                  xv[r] = (r0 + r < p.M) ? __bfloat162float(xp[static_cast<int64_t>(r) * p.ldx]) : 0.f;
#pragma unroll
                for (int r = 0; r < 32; ++r) {
                  const float o = lds_f32(base + static_cast<uint32_t>(r) * 128u + ((chunk ^ static_cast<uint32_t>(r & 7)) << 4));
                  s1 += o;
                  s2 = fmaf(o, xv[r], s2);
                }
                const int64_t t = static_cast<int64_t>(tm) * 4 + q;
                cpart[(t * 2 + 0) * p.N + col] = s1;
                cpart[(t * 2 + 1) * p.N + col] = s2;
              }
            }
            ++nstore;
          } else if (row_ok) {
            float* dst = outp + static_cast<int64_t>(row) * p.ldc + col0;
            if (col0 + 32 <= p.N && (p.ldc & 3) == 0) {
#pragma unroll
              for (int i = 0; i < 8; ++i)
                reinterpret_cast<float4*>(dst)[i] =
                    make_float4(__uint_as_float(v[4 * i]) * p.alpha, __uint_as_float(v[4 * i + 1]) * p.alpha,
                                __uint_as_float(v[4 * i + 2]) * p.alpha, __uint_as_float(v[4 * i + 3]) * p.alpha);
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (col0 + i < p.N) dst[i] = __uint_as_float(v[i]) * p.alpha;
            }
          }
        }
      } else if (EPI == EPI_BARLOW_BWD) {
        // dz tile = (dT - x~ * q) * rstd * grad_out with x~ = (x - mean) * rstd from the fp32 input rows, straight from the
        // accumulator to the caller's gradient (TMA store)
        const float* xf = second ? p.xf2 : p.xf;
        const int64_t ldxf = second ? p.ldxf2 : p.ldxf;
        const float* vm = second ? p.vm2 : p.vm;
        const float* vr = second ? p.vr2 : p.vr;
        const float* vq = second ? p.vq2 : p.vq;
        const float go = __ldg(p.go);
        const float ago = p.alpha * go;
#pragma unroll 1
        for (int cc = 0; cc < BN / 32; ++cc) {
          const int col0 = tn * BN + cc * 32;
          uint32_t v[32];
          tmem_ld_x32(tmem + tlane + ab * BN + cc * 32, v);
          // this thread's 32 input values (128 contiguous bytes) while the TMEM load is in flight
          float4 xv[8];
          const bool full = col0 + 32 <= p.N;  // warp-uniform
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            xv[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (row_ok && (full || col0 + 4 * i < p.N))  // (N % 4 == 0: whole 16-byte pieces)
              xv[i] = __ldg(reinterpret_cast<const float4*>(xf + static_cast<int64_t>(row) * ldxf + col0) + i);
          }
          tmem_ld_wait_regs32(v);
          if (col0 >= p.N) continue;  // warp-uniform
          const uint32_t buf = my_row + static_cast<uint32_t>(nstore & 1) * 4096u;
          if (nstore >= 2) {
            if (lane == 0) bulk_wait_group_read<1>();
            __syncwarp();
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float4 m4 = make_float4(0.f, 0.f, 0.f, 0.f), r4 = m4, q4 = m4;
            if (full || col0 + 4 * i < p.N) {
              m4 = __ldg(reinterpret_cast<const float4*>(vm + col0) + i);
              r4 = __ldg(reinterpret_cast<const float4*>(vr + col0) + i);
              q4 = __ldg(reinterpret_cast<const float4*>(vq + col0) + i);
            }
            // (acc * alpha - (x - mean) * q) * rstd * go   [q already carries one rstd: x~ * m2 = (x - mean) * rstd * m2]
            const float o0 = fmaf(__uint_as_float(v[4 * i]) * ago, r4.x, -((xv[i].x - m4.x) * q4.x) * (r4.x * go));
            const float o1 = fmaf(__uint_as_float(v[4 * i + 1]) * ago, r4.y, -((xv[i].y - m4.y) * q4.y) * (r4.y * go));
            const float o2 = fmaf(__uint_as_float(v[4 * i + 2]) * ago, r4.z, -((xv[i].z - m4.z) * q4.z) * (r4.z * go));
            const float o3 = fmaf(__uint_as_float(v[4 * i + 3]) * ago, r4.w, -((xv[i].w - m4.w) * q4.w) * (r4.w * go));
            sts_v4(buf + ((static_cast<uint32_t>(i) ^ swz) << 4), __float_as_uint(o0), __float_as_uint(o1), __float_as_uint(o2),
                   __float_as_uint(o3));
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(mc, stage0 + static_cast<uint32_t>(nstore & 1) * 4096u, col0, tm * C::BM + q * 32);
            bulk_commit_group();
          }
          ++nstore;
        }
      } else {
        // EPI_BARLOW: loss terms + dC (bf16); 64 columns (128 bytes of bf16 per row) per staged block
        float rs_acc = 0.f;  // sum over this tile's columns of dC .* C (row = this thread)
        // (no `unroll 1` here: the compiler's own unrolling of these BN / 64 = 4 blocks measured 193 vs 202 us at cfg3)
        for (int cc = 0; cc < BN / 64; ++cc) {
          const int col0 = tn * BN + cc * 64;
          uint32_t pk[32];
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            uint32_t v[32];
            tmem_ld_x32(tmem + tlane + ab * BN + cc * 64 + h * 32, v);
            tmem_ld_wait_regs32(v);
            float gc[32];  // dC .* C of this thread's row, 32 columns
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              float g[2];
#pragma unroll
              for (int e = 0; e < 2; ++e) {
                const int col = col0 + h * 32 + 2 * i + e;
                const float c = __uint_as_float(v[2 * i + e]) * p.alpha;
                const bool diag = (col + p.diag_off == row);
                const float r = diag ? (c - 1.f) : c;
                const float w = diag ? 1.f : p.lambda;
                const bool ok = row_ok && col < p.N;
                if (ok) loss_acc = fmaf(w * r, r, loss_acc);
                g[e] = 2.f * w * r;
                gc[2 * i + e] = ok ? c : 0.f;
              }
              const uint32_t pkv = pack_bf16x2(g[0], g[1]);
              pk[h * 16 + i] = pkv;
              // dC .* C with dC AS STORED (bf16): the backward GEMM multiplies by the rounded dC, and sum_n(dT x~) must carry
              // the same rounding - dT and x~ * sum_n(dT x~)/(n-1) largely cancel, an inconsistent dC would be amplified
              gc[2 * i] *= __uint_as_float(pkv << 16);
              gc[2 * i + 1] *= __uint_as_float(pkv & 0xffff0000u);
            }
            if (p.bl_rowpart != nullptr) {
#pragma unroll
              for (int i = 0; i < 32; ++i) rs_acc += gc[i];
            }
            if (p.bl_colpart != nullptr) {
              // transpose-reduce over the warp's 32 rows: 31 shuffles leave the sum of column `lane` in gc[0]
#pragma unroll
              for (int sft = 16; sft >= 1; sft >>= 1) {
                const bool upper = (lane & sft) != 0;
#pragma unroll
                for (int j = 0; j < sft; ++j) {
                  const float keep = upper ? gc[j + sft] : gc[j];
                  const float send = upper ? gc[j] : gc[j + sft];
                  gc[j] = keep + __shfl_xor_sync(0xffffffffu, send, sft);
                }
              }
              const int col = col0 + h * 32 + lane;
              if (col < p.N) p.bl_colpart[(static_cast<int64_t>(tm) * 4 + q) * p.N + col] = gc[0];
            }
          }
          if (col0 >= p.N) continue;  // warp-uniform
          if (p.tma_store) {
            const uint32_t buf = my_row + static_cast<uint32_t>(nstore & 1) * 4096u;
            if (nstore >= 2) {
              if (lane == 0) bulk_wait_group_read<1>();
              __syncwarp();
            }
#pragma unroll
            for (int i = 0; i < 8; ++i)
              sts_v4(buf + ((static_cast<uint32_t>(i) ^ swz) << 4), pk[4 * i], pk[4 * i + 1], pk[4 * i + 2], pk[4 * i + 3]);
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
              tma_store_2d(mc, stage0 + static_cast<uint32_t>(nstore & 1) * 4096u, col0, tm * C::BM + q * 32);
              bulk_commit_group();
            }
            ++nstore;
          } else if (row_ok) {
            __nv_bfloat16* dst = p.dC + static_cast<int64_t>(row) * p.ld_dc + col0;
            if (col0 + 64 <= p.N) {
#pragma unroll
              for (int i = 0; i < 8; ++i)
                reinterpret_cast<uint4*>(dst)[i] = make_uint4(pk[4 * i], pk[4 * i + 1], pk[4 * i + 2], pk[4 * i + 3]);
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (col0 + 2 * i + 1 < p.N) reinterpret_cast<uint32_t*>(dst)[i] = pk[i];
            }
          }
        }
        if (p.bl_rowpart != nullptr && row_ok) p.bl_rowpart[static_cast<int64_t>(tn) * p.M + row] = rs_acc;
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[ab]);
    }
    if (lane == 0) bulk_wait_group_all();  // staging buffers stay valid until the TMA unit has read (and written) them
    if (EPI == EPI_BARLOW) {
      loss_acc = warp_sum(loss_acc);
      if (lane == 0) loss_red[q] = loss_acc;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (EPI == EPI_BARLOW && threadIdx.x == 0)
    p.loss_partials[blockIdx.x] = (loss_red[0] + loss_red[1]) + (loss_red[2] + loss_red[3]);
  if (warp == 2) tmem_dealloc<2 * BN>(tmem);
}

}  // namespace ssvb
