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
#pragma once
#include "common.cuh"

namespace ssvb {

enum GemmEpiMode { EPI_STORE_F32 = 0, EPI_BARLOW = 1 };

struct GemmParams {
  int M, N, K;  // logical sizes (tails are zero-filled by TMA and predicated in the epilogue)
  int tiles_m, tiles_n;
  // EPI_STORE_F32: out[m * ldc + n] = alpha * acc
  float alpha;
  float* out;
  int64_t ldc;
  // EPI_BARLOW: c = acc * alpha; loss += (m==n) ? (c-1)^2 : lambda c^2 ; dC = (m==n) ? 2(c-1) : 2 lambda c  (bf16)
  float lambda;
  __nv_bfloat16* dC;
  int64_t ld_dc;
  float* loss_partials;  // [gridDim.x]
  int diag_off;          // the output is a column slab of the full matrix: global column = n + diag_off
};

template <int BN>
struct GemmCfg {
  static constexpr int BM = 128;
  static constexpr int BK = 64;
  static constexpr int A_BYTES = BM * BK * 2;  // 16 KB
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int NSTAGE = (BN == 256) ? 4 : 6;
  static constexpr int NBARS = 2 * NSTAGE + 4;
  static constexpr int SMEM = 1024 + NSTAGE * (A_BYTES + B_BYTES) + NBARS * 8 + 16;
};

__device__ __forceinline__ void tmem_ld_wait_regs32(uint32_t (&v)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]),
                 "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]),
                 "+r"(v[15]), "+r"(v[16]), "+r"(v[17]), "+r"(v[18]), "+r"(v[19]), "+r"(v[20]), "+r"(v[21]),
                 "+r"(v[22]), "+r"(v[23]), "+r"(v[24]), "+r"(v[25]), "+r"(v[26]), "+r"(v[27]), "+r"(v[28]),
                 "+r"(v[29]), "+r"(v[30]), "+r"(v[31])::"memory");
}

template <int BN, bool A_MN, bool B_MN, int EPI>
__global__ void __launch_bounds__(192, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmParams p) {
  using C = GemmCfg<BN>;
  constexpr int NSTAGE = C::NSTAGE;
  extern __shared__ uint8_t gemm_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(gemm_smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + NSTAGE * C::A_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + NSTAGE * C::B_BYTES);
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
  const int ntiles = p.tiles_m * p.tiles_n;
  const int nkb = (p.K + C::BK - 1) / C::BK;

  // producer / issuer loops run on the whole warp (uniform operands stay in uniform registers); elect_one() issues
  if (warp == 0) {
    int it = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const int tm = tile % p.tiles_m, tn = tile / p.tiles_m;
      for (int kb = 0; kb < nkb; ++kb, ++it) {
        const int st = it % NSTAGE;
        mbar_wait(&empty[st], ((it / NSTAGE) & 1) ^ 1);
        if (elect_one()) {
          mbar_expect_tx(&full[st], C::A_BYTES + C::B_BYTES);
          uint8_t* a = sA + st * C::A_BYTES;
          uint8_t* b = sB + st * C::B_BYTES;
          if (A_MN) {
#pragma unroll
            for (int blk = 0; blk < C::BM / 64; ++blk)
              tma_load_2d(a + blk * 8192, &tmA, &full[st], tm * C::BM + blk * 64, kb * C::BK);
          } else {
            tma_load_2d(a, &tmA, &full[st], kb * C::BK, tm * C::BM);
          }
          if (B_MN) {
#pragma unroll
            for (int blk = 0; blk < BN / 64; ++blk)
              tma_load_2d(b + blk * 8192, &tmB, &full[st], tn * BN + blk * 64, kb * C::BK);
          } else {
            tma_load_2d(b, &tmB, &full[st], kb * C::BK, tn * BN);
          }
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t IDESC = make_idesc(C::BM, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
    int it = 0, tcount = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++tcount) {
      const int ab = tcount & 1;
      mbar_wait(&acc_empty[ab], ((tcount >> 1) & 1) ^ 1);
      tc_fence_after();
      for (int kb = 0; kb < nkb; ++kb, ++it) {
        const int st = it % NSTAGE;
        mbar_wait(&full[st], (it / NSTAGE) & 1);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t abase = smem_u32(sA + st * C::A_BYTES), bbase = smem_u32(sB + st * C::B_BYTES);
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4) {
            const uint64_t ad = A_MN ? desc_mnmajor(abase + k4 * 2048, 8192) : desc_kmajor(abase + k4 * 32);
            const uint64_t bd = B_MN ? desc_mnmajor(bbase + k4 * 2048, 8192) : desc_kmajor(bbase + k4 * 32);
            umma_ss(tmem + ab * BN, ad, bd, IDESC, (kb | k4) != 0);
          }
          umma_commit(&empty[st]);
          if (kb == nkb - 1) umma_commit(&acc_full[ab]);
        }
        __syncwarp();
      }
    }
  } else {
    const int q = warp & 3;
    const int row_l = q * 32 + lane;
    const uint32_t tlane = static_cast<uint32_t>(q * 32) << 16;
    float loss_acc = 0.f;
    int tcount = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++tcount) {
      const int tm = tile % p.tiles_m, tn = tile / p.tiles_m;
      const int ab = tcount & 1;
      const int row = tm * C::BM + row_l;
      const bool row_ok = row < p.M;
      mbar_wait(&acc_full[ab], (tcount >> 1) & 1);
      tc_fence_after();
#pragma unroll 1
      for (int cc = 0; cc < BN / 32; ++cc) {
        uint32_t v[32];
        tmem_ld_x32(tmem + tlane + ab * BN + cc * 32, v);
        tmem_ld_wait_regs32(v);
        const int col0 = tn * BN + cc * 32;
        if (col0 >= p.N) continue;  // warp-uniform
        if (EPI == EPI_STORE_F32) {
          if (row_ok) {
            float* dst = p.out + static_cast<int64_t>(row) * p.ldc + col0;
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
        } else {
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            float g[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int col = col0 + 2 * i + e;
              const float c = __uint_as_float(v[2 * i + e]) * p.alpha;
              const bool diag = (col + p.diag_off == row);
              const float r = diag ? (c - 1.f) : c;
              const float w = diag ? 1.f : p.lambda;
              if (row_ok && col < p.N) loss_acc = fmaf(w * r, r, loss_acc);
              g[e] = 2.f * w * r;
            }
            pk[i] = pack_bf16x2(g[0], g[1]);
          }
          if (row_ok) {
            __nv_bfloat16* dst = p.dC + static_cast<int64_t>(row) * p.ld_dc + col0;
            if (col0 + 32 <= p.N) {
#pragma unroll
              for (int i = 0; i < 4; ++i)
                reinterpret_cast<uint4*>(dst)[i] = make_uint4(pk[4 * i], pk[4 * i + 1], pk[4 * i + 2], pk[4 * i + 3]);
            } else {
#pragma unroll
              for (int i = 0; i < 16; ++i)
                if (col0 + 2 * i + 1 < p.N) reinterpret_cast<uint32_t*>(dst)[i] = pk[i];
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[ab]);
    }
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
