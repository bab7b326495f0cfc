// Flash-style similarity-matrix kernels for NT-Xent (SimCLR / ReLIC) and MoCo InfoNCE.
//
//   sim_fwd_kernel : S tile = A_I * B_J^T on tcgen05 (bf16 operands via TMA, fp32 accumulators in TMEM),
//                    streamed through exp2 / row-sum in registers -> per-(row, column-chunk) partial
//                    (max, sum) pairs.  S never touches HBM.
//   sim_bwd_kernel : recompute the S tile, turn it into the softmax weight tile W (bf16, written back to
//                    TMEM), second tcgen05 GEMM dA_I += W * B_J with W as the TMEM A-operand and the SAME
//                    smem B tile read MN-major.  Accumulator stays in TMEM for the whole row block.
//
// Warp roles (320 threads, 1 CTA / SM): warp 0 = TMA producer, warp 1 = MMA issuer, warps 2..9 = two
// softmax warpgroups (TMEM lane quarter = warp % 4).  All hand-offs are mbarriers; every wait is bounded.
#pragma once
#include "common.cuh"

namespace ssvb {

enum SimMode { SIM_NTX_FIXED = 0, SIM_NTX_ONLINE = 1, SIM_MOCO = 2 };

struct SimParams {
  // A rows: `nseg` segments of `seg_rows` rows starting at global rows seg_start[s] of the A tensor map.
  // Local row index (partials / dacc / MoCo rowstat) = seg * seg_rows + row-in-segment.
  int nseg, seg_rows, seg_start[2];
  int bps;         // 128-row blocks per segment
  int row_blocks;  // nseg * bps
  int cols;        // valid columns (rows of B)
  int col_tiles;   // ceil(cols / BN)
  int nchunks, tiles_per_chunk;
  float c;      // log2(e) / temperature
  float shift;  // FIXED mode: constant log2-domain shift (= c, the largest possible logit)
  // forward outputs: [4 * nchunks][part_stride] (one partial per softmax warpgroup and column chunk)
  float* part_m;
  float* part_l;
  int part_stride;
  // backward inputs / outputs
  const float* rowstat;  // NT-Xent: indexed by global row; MoCo: by local row
  const float* colstat;  // NT-Xent: fp32 column statistic (FIXED: 1/L', ONLINE: lse2), padded to a multiple of 128
  float* dacc;           // [local rows x ld_dacc] fp32
  int ld_dacc;
  int use_atomic;  // nchunks > 1: red.add into a zeroed dacc
  unsigned long long* dbg;  // SSVB_DBG_TIMING builds only
  int opf16;  // similarity operands staged as fp16 (normalised rows) instead of bf16
  int dhalf;  // backward, dpad = 256 only: which 128-column half of dZ this launch produces
};

__device__ __forceinline__ uint8_t* align1024(uint8_t* p) {
  return reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(p) + 1023) & ~static_cast<uintptr_t>(1023));
}

// wait::ld with the destination registers as in/out operands so no consumer is scheduled above it
__device__ __forceinline__ void tmem_ld_wait_regs(uint32_t (&v)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]),
                 "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]),
                 "+r"(v[15]), "+r"(v[16]), "+r"(v[17]), "+r"(v[18]), "+r"(v[19]), "+r"(v[20]), "+r"(v[21]),
                 "+r"(v[22]), "+r"(v[23]), "+r"(v[24]), "+r"(v[25]), "+r"(v[26]), "+r"(v[27]), "+r"(v[28]),
                 "+r"(v[29]), "+r"(v[30]), "+r"(v[31])::"memory");
}

struct UnitInfo {
  int g0;      // global A row of the block's first row
  int lrow0;   // local row index of the block's first row
  int nvalid;  // valid rows in the block
  int t0, t1;  // column-tile range of the chunk
  int ch;
};
__device__ __forceinline__ UnitInfo decode_unit(const SimParams& p, int u) {
  UnitInfo ui;
  const int rb = u / p.nchunks;
  ui.ch = u - rb * p.nchunks;
  const int seg = rb / p.bps;
  const int rbi = rb - seg * p.bps;
  ui.g0 = p.seg_start[seg] + rbi * 128;
  ui.lrow0 = seg * p.seg_rows + rbi * 128;
  ui.nvalid = min(128, p.seg_rows - rbi * 128);
  ui.t0 = ui.ch * p.tiles_per_chunk;
  ui.t1 = min(p.col_tiles, ui.t0 + p.tiles_per_chunk);
  return ui;
}

// phase timing (timing experiments only): block 0, first lane of the first math warp / of the MMA warp accumulate
// clock64 deltas per phase into p.dbg[0..15]
#ifdef SSVB_DBG_TIMING
#define SSVB_T0() long long _t_prev = clock64()
#define SSVB_TP(slot) do { const long long _t_now = clock64(); _t_acc[slot] += _t_now - _t_prev; _t_prev = _t_now; } while (0)
#else
#define SSVB_T0() do {} while (0)
#define SSVB_TP(slot) do {} while (0)
#endif

__device__ __forceinline__ float ex2_poly3(float x);
// ---- timing-experiment knobs (tools/ab_bench.sh): each one REMOVES a piece of work, results become wrong ----
#ifdef SSVB_DBG_NOEXP
#define SSVB_EX2(x) (x)
#else
#define SSVB_EX2(x) ex2f(x)
#endif
#ifndef SSVB_POLY_MOD_FWD
#define SSVB_POLY_MOD_FWD 3   // masked (rare) tiles only: every 3rd element on the scalar polynomial
#endif
// FIXED mode hot loops: of every SSVB_POLY_*_MOD consecutive PAIRS, the first SSVB_POLY_*_CNT take the packed
// polynomial exp2 and the rest two MUFU.EX2 (CNT = 0: MUFU only).  MUFU issues through the MIO queue at 16 / clk / SM
// and is co-critical with the tensor pipe in both kernels (ncu: mio_throttle on MUFU.EX2), so a share of the
// exponentials moves to the FMA pipe; tuned by A/B builds (profiles/r2_tuning_log.md).
#ifndef SSVB_POLY_FWD_MOD
#define SSVB_POLY_FWD_MOD 2
#endif
#ifndef SSVB_POLY_FWD_CNT
#define SSVB_POLY_FWD_CNT 1
#endif
#ifndef SSVB_POLY_BWD_MOD
#define SSVB_POLY_BWD_MOD 3
#endif
#ifndef SSVB_POLY_BWD_CNT
#define SSVB_POLY_BWD_CNT 1
#endif

// exp2 of two values on the FMA pipe with packed instructions (Cody-Waite split + degree-3 minimax, max rel. error
// 7.5e-5): 6 FADD2/FFMA2 + 2 IMAD per pair = 4 issue slots per element and no MUFU slot.  Valid for x > -126.
__device__ __forceinline__ unsigned long long ex2_poly3_x2(unsigned long long x2) {
  const unsigned long long magic2 = pack_f32x2(12582912.f, 12582912.f);
  const unsigned long long nmagic2 = pack_f32x2(-12582912.f, -12582912.f);
  const unsigned long long r2 = add_f32x2(x2, magic2);      // low mantissa bits = round(x)
  const unsigned long long t2 = add_f32x2(r2, nmagic2);     // round(x) as a float (exact)
  const unsigned long long f2 = fma_f32x2(t2, pack_f32x2(-1.f, -1.f), x2);  // f in [-0.5, 0.5]
  unsigned long long q2 = fma_f32x2(f2, pack_f32x2(5.517166745e-2f, 5.517166745e-2f),
                                    pack_f32x2(2.426111221e-1f, 2.426111221e-1f));
  q2 = fma_f32x2(q2, f2, pack_f32x2(6.932609858e-1f, 6.932609858e-1f));
  q2 = fma_f32x2(q2, f2, pack_f32x2(9.999280736e-1f, 9.999280736e-1f));
  float q0, q1, r0, r1;
  unpack_f32x2(q2, q0, q1);
  unpack_f32x2(r2, r0, r1);
  const float e0 = __int_as_float(__float_as_int(q0) + (__float_as_int(r0) << 23));
  const float e1 = __int_as_float(__float_as_int(q1) + (__float_as_int(r1) << 23));
  return pack_f32x2(e0, e1);
}
// 16-byte shared-memory load from a 32-bit shared address (the generic-pointer form costs two address instructions
// per load plus a generic->shared window computation per call site)
__device__ __forceinline__ void lds_v4(uint32_t saddr, unsigned long long& lo, unsigned long long& hi) {
  asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(lo), "=l"(hi) : "r"(saddr));
}

// =====================================================================================================
// forward
// =====================================================================================================
// Forward tile width.  256: two 256-column S buffers, a warpgroup PAIR shares a tile (one 128-column half each) and the
// buffer is handed back after 3/4 of the pair's math (default).  128: FOUR 128-column S buffers, each softmax warpgroup
// owns a whole tile in its own buffer (more independent MMA -> softmax streams); kept as a build option, measured slower.
#ifndef SSVB_FWD_BN
#define SSVB_FWD_BN 256   // measured (profiles/r2_tuning_log.md): 128 is 4 % slower (16 instead of 8 MMAs + commits per 256 columns on the one issuing thread)
#endif
constexpr int kFwdBN = SSVB_FWD_BN;
template <int KB>
struct FwdCfg {
  // d <= 128 (KB <= 2): BN = kFwdBN.  128 < d <= 256 (KB = 4, dpad = 256): 128-column tiles (four S buffers) and a
  // 2-stage ring - a 64 KB A tile + 2 x 64 KB B tiles is what fits in 227 KB; functional coverage, not the tuned path.
  static constexpr int BN = (KB > 2) ? 128 : kFwdBN;
  static constexpr int NBUF = 512 / BN;  // S buffers in TMEM
  static constexpr int A_BYTES = 128 * 128 * KB;
  static constexpr int B_BYTES = BN * 128 * KB;
  static constexpr int NSTAGE = (KB > 2) ? 2 : ((KB == 1) ? 6 : 3) * (256 / BN);  // deep enough to cover the TMA latency
  static constexpr int NBARS = 4 + 2 * NSTAGE + 2 * NBUF;
  static constexpr int SMEM = 1024 + A_BYTES + NSTAGE * B_BYTES + NBARS * 8 + 16;
};

template <int MODE, bool MASKED>
__device__ __forceinline__ void fwd_tile(uint32_t taddr, const SimParams& p, int a_glob, int j0, float& m,
                                         float (&l)[4], uint32_t s_empty_bar, int lane
#ifdef SSVB_DBG_TIMING
                                         , long long (&_t_acc)[8], long long& _t_prev
#endif
                                         ) {
  uint32_t v[2][32];
  unsigned long long l2[2] = {pack_f32x2(0.f, 0.f), pack_f32x2(0.f, 0.f)};  // packed row-sum accumulators
#ifndef SSVB_DBG_NOLD
  tmem_ld_x32(taddr, v[0]);
#else
#pragma unroll
  for (int i = 0; i < 32; ++i) { v[0][i] = __float_as_uint(1e-3f * (i + lane)); v[1][i] = __float_as_uint(2e-3f * (i + lane)); }
#endif
#pragma unroll
  for (int cc = 0; cc < 4; ++cc) {
    uint32_t(&cur)[32] = v[cc & 1];
    SSVB_TP(3);  // exp / sum work of the previous chunk
#ifndef SSVB_DBG_NOLD
    tmem_ld_wait_regs(cur);
#endif
    SSVB_TP(2);  // waiting for the TMEM load
    if (cc < 3) {
#ifndef SSVB_DBG_NOLD
      tmem_ld_x32(taddr + (cc + 1) * 32, v[(cc + 1) & 1]);
#endif
    } else {
      // the S buffer is fully in registers: hand it back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_a(s_empty_bar);
    }
    const int colbase = j0 + cc * 32;
    if (MODE == SIM_NTX_FIXED && !MASKED) {
      // FIXED mode operands are pre-scaled by sqrt(log2(e)/tau) (pair_prep), so the accumulator IS the log2-domain
      // logit: no scale/shift FFMA.  Per pair: two MUFU.EX2 or one packed polynomial, one packed row-sum FADD2.
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const bool poly = (i % SSVB_POLY_FWD_MOD) < SSVB_POLY_FWD_CNT;
        unsigned long long e2;
        if (poly) {
          e2 = ex2_poly3_x2(pack_f32x2(__uint_as_float(cur[2 * i]), __uint_as_float(cur[2 * i + 1])));
        } else {
          e2 = pack_f32x2(SSVB_EX2(__uint_as_float(cur[2 * i])), SSVB_EX2(__uint_as_float(cur[2 * i + 1])));
        }
        l2[i & 1] = add_f32x2(l2[i & 1], e2);
      }
    } else if (MODE == SIM_NTX_FIXED) {
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        float t = fmaf(__uint_as_float(cur[i]), p.c, -p.shift);
        if (MASKED) {
          const int col = colbase + i;
          if (col == a_glob || col >= p.cols) t = -INFINITY;
        }
        const bool poly = !MASKED && SSVB_POLY_MOD_FWD > 0 && (i % (SSVB_POLY_MOD_FWD > 0 ? SSVB_POLY_MOD_FWD : 1)) == 1;
        l[i & 3] += poly ? ex2_poly3(t) : SSVB_EX2(t);
      }
    } else {
      float x[32];
      float cm = -INFINITY;
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        float t = __uint_as_float(cur[i]) * p.c;
        if (MASKED) {
          const int col = colbase + i;
          if ((MODE != SIM_MOCO && col == a_glob) || col >= p.cols) t = -INFINITY;
        }
        x[i] = t;
        cm = fmaxf(cm, t);
      }
      const float mn = fmaxf(m, cm);
      const float sc = ex2f(m - mn);
      m = mn;
#pragma unroll
      for (int i = 0; i < 4; ++i) l[i] *= sc;
#pragma unroll
      for (int i = 0; i < 32; ++i) l[i & 3] += ex2f(x[i] - mn);
    }
  }
  if (MODE == SIM_NTX_FIXED && !MASKED) {
    float a0, a1, b0, b1;
    unpack_f32x2(l2[0], a0, a1);
    unpack_f32x2(l2[1], b0, b1);
    l[0] += a0; l[1] += a1; l[2] += b0; l[3] += b1;
  }
}


// 640 threads: warpgroup 0 = {TMA producer, MMA issuer, 2 idle warps} shrinks to 40 registers/thread; warps 4..19 are
// FOUR softmax warpgroups (TMEM lane quarter = warp % 4): tile t belongs to the warpgroup pair (t & 1) (= its S
// buffer), and within the pair each warpgroup streams one 128-column half.  Four math warps per scheduler keep the
// MUFU and FMA pipes (polynomial exp2) busy at the same time.
template <int KB, int MODE, bool OPF16 = false>
__global__ void __launch_bounds__(640, 1)
sim_fwd_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const SimParams p) {
  using C = FwdCfg<KB>;
  constexpr int BN = C::BN, NSTAGE = C::NSTAGE, NBUF = C::NBUF;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align1024(smem_raw);
  uint8_t* sA = smem;
  uint8_t* sB = smem + C::A_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + NSTAGE * C::B_BYTES);
  uint64_t* a_full = bars;
  uint64_t* a_empty = bars + 2;
  uint64_t* b_full = bars + 4;
  uint64_t* b_empty = b_full + NSTAGE;
  uint64_t* s_full = b_empty + NSTAGE;
  uint64_t* s_empty = s_full + NBUF;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_empty + NBUF);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&a_full[i], 1);
      mbar_init(&a_empty[i], 1);
    }
    for (int i = 0; i < NBUF; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&s_empty[i], 16 / NBUF);  // the warps that drain one S buffer: a warpgroup pair (8) or one warpgroup (4)
    }
    for (int i = 0; i < NSTAGE; ++i) {
      mbar_init(&b_full[i], 1);
      mbar_init(&b_empty[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int nunits = p.row_blocks * p.nchunks;

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
  // Producer / issuer loops are executed by the WHOLE warp (warp-uniform control flow and operands, so the TMA /
  // UMMA operands stay in uniform registers); only the issue instructions sit behind elect_one().
  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    int ucount = 0, gt = 0;
    for (int u = blockIdx.x; u < nunits; u += gridDim.x, ++ucount) {
      const UnitInfo ui = decode_unit(p, u);
      mbar_wait(&a_empty[0], (ucount & 1) ^ 1);
      if (elect_one()) {
        mbar_expect_tx(&a_full[0], C::A_BYTES);
#pragma unroll
        for (int kb = 0; kb < KB; ++kb) tma_load_2d(sA + kb * (128 * 128), &tmA, &a_full[0], kb * 64, ui.g0);
      }
      __syncwarp();
      for (int t = ui.t0; t < ui.t1; ++t, ++gt) {
        const int st = gt % NSTAGE;
        mbar_wait(&b_empty[st], ((gt / NSTAGE) & 1) ^ 1);
        if (elect_one()) {
          mbar_expect_tx(&b_full[st], C::B_BYTES);
#pragma unroll
          for (int kb = 0; kb < KB; ++kb)
            tma_load_2d(sB + st * C::B_BYTES + kb * (BN * 128), &tmB, &b_full[st], kb * 64, t * BN);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    constexpr uint32_t IDESC = make_idesc(128, BN, 0, 0, OPF16 ? 0 : 1);
    const uint32_t abase = smem_u32(sA);
    int ucount = 0, gt = 0;
    for (int u = blockIdx.x; u < nunits; u += gridDim.x, ++ucount) {
      const UnitInfo ui = decode_unit(p, u);
      mbar_wait(&a_full[0], ucount & 1);
      for (int t = ui.t0; t < ui.t1; ++t, ++gt) {
        const int st = gt % NSTAGE, buf = gt % NBUF;
        mbar_wait(&b_full[st], (gt / NSTAGE) & 1);
        mbar_wait(&s_empty[buf], ((gt / NBUF) & 1) ^ 1);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t bbase = smem_u32(sB + st * C::B_BYTES);
#ifndef SSVB_DBG_FWD_NOS
#pragma unroll
          for (int kb = 0; kb < KB; ++kb)
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4)
              umma_ss(tmem + buf * BN, desc_kmajor(abase + kb * (128 * 128) + k4 * 32),
                      desc_kmajor(bbase + kb * (BN * 128) + k4 * 32), IDESC, (kb | k4) != 0);
#endif
          umma_commit(&b_empty[st]);
          umma_commit(&s_full[buf]);
          if (t == ui.t1 - 1) umma_commit(&a_empty[0]);
        }
        __syncwarp();
      }
    }
  }
  } else {
    // ------------------------------------------------------------------ softmax warpgroups
    asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
    const int wgi = (warp - 4) >> 2;  // 0..3
    const int pair = wgi >> 1, half = wgi & 1;
    const int q = warp & 3;
    const int row_l = q * 32 + lane;
    const uint32_t tlane = static_cast<uint32_t>(q * 32) << 16;
    int gt = 0;
    const int my_buf = NBUF == 2 ? pair : wgi;
    const uint32_t a_s_full = smem_u32(&s_full[my_buf]), a_s_empty = smem_u32(&s_empty[my_buf]);  // hoisted (see bwd)
#ifdef SSVB_DBG_TIMING
    long long _t_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#endif
    SSVB_T0();
    for (int u = blockIdx.x; u < nunits; u += gridDim.x) {
      const UnitInfo ui = decode_unit(p, u);
      const int a_glob = ui.g0 + row_l;
      float m = -1e30f;
      float l[4] = {0.f, 0.f, 0.f, 0.f};
      for (int t = ui.t0; t < ui.t1; ++t, ++gt) {
        // BN = 256: tile t belongs to the warpgroup pair (t & 1), each warpgroup of the pair streams one 128-column half;
        // BN = 128: tile t belongs to warpgroup (t & 3) alone (its own S buffer)
        if (NBUF == 2 ? ((gt & 1) != pair) : ((gt & 3) != wgi)) continue;
        const int buf = NBUF == 2 ? pair : wgi;
        SSVB_TP(0);  // loop / other
        mbar_wait_a(a_s_full, (gt / NBUF) & 1);
        tc_fence_after();
        SSVB_TP(1);  // wait s_full
        const int j0 = t * BN + (NBUF == 2 ? half * 128 : 0);
        const bool special = (MODE != SIM_MOCO && j0 < ui.g0 + 128 && j0 + 128 > ui.g0) || (j0 + 128 > p.cols);
        const uint32_t taddr = tmem + tlane + buf * BN + (NBUF == 2 ? half * 128 : 0);
#ifdef SSVB_DBG_TIMING
        if (special)
          fwd_tile<MODE, true>(taddr, p, a_glob, j0, m, l, a_s_empty, lane, _t_acc, _t_prev);
        else
          fwd_tile<MODE, false>(taddr, p, a_glob, j0, m, l, a_s_empty, lane, _t_acc, _t_prev);
#else
        if (special)
          fwd_tile<MODE, true>(taddr, p, a_glob, j0, m, l, a_s_empty, lane);
        else
          fwd_tile<MODE, false>(taddr, p, a_glob, j0, m, l, a_s_empty, lane);
#endif
      }
      if (row_l < ui.nvalid) {
        const size_t o = static_cast<size_t>(ui.ch * 4 + wgi) * p.part_stride + ui.lrow0 + row_l;
        p.part_l[o] = (l[0] + l[1]) + (l[2] + l[3]);
        if (MODE != SIM_NTX_FIXED) p.part_m[o] = m;
      }
    }
#ifdef SSVB_DBG_TIMING
    if (p.dbg && blockIdx.x == 0 && warp == 4 && lane == 0)
      for (int i = 0; i < 8; ++i) p.dbg[16 + i] = static_cast<unsigned long long>(_t_acc[i]);
#endif
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<512>(tmem);
}

// =====================================================================================================
// backward
// =====================================================================================================
// TMEM plan (512 columns): S0 @0, S1 @128 (fp32 similarity tiles), dZ @256 (DP-column gradient accumulator),
// W0 @384, W1 @448 (bf16 weight tiles, the TMEM A-operand of the second GEMM).
// Decoupling: a weight warpgroup pulls its WHOLE S tile into registers first and hands the S buffer straight back,
// and the MMA thread never blocks on a single barrier — it polls "next S issuable" and "next dZ issuable" and issues
// whichever is ready — so the similarity GEMM of tile t+2 runs while W(t) is still being computed and the
// W -> dZ latency is off the critical path of the exp warps.
template <int KB>
struct BwdCfg {
  static constexpr int BN = 128;
  // output columns of one launch: the whole row for d <= 128; for KB = 4 (dpad = 256) the dZ accumulator holds ONE
  // 128-column half (TMEM: S0 S1 dZ W0 W1 = 512 columns) and the host launches the kernel once per half (p.dhalf) -
  // S and the weights are recomputed, the second GEMM reads k-blocks {2 dhalf, 2 dhalf + 1} of the same B tile
  static constexpr int DP = (KB >= 2) ? 128 : 64;
  static constexpr int A_BYTES = 128 * 128 * KB;
  static constexpr int B_BYTES = BN * 128 * KB;
  static constexpr int CS_BYTES = BN * 4;  // per-stage column statistics (fp32)
  static constexpr int NSTAGE = (KB == 1) ? 8 : (KB == 2 ? 5 : 2);
  static constexpr int NBARS = 2 + 2 * NSTAGE + 8 + 2;
  static constexpr int SMEM = 1024 + A_BYTES + NSTAGE * (B_BYTES + CS_BYTES) + NBARS * 8 + 16;
  static constexpr int T_S = 0;
  static constexpr int T_DZ = 256;
  static constexpr int T_W = 384;
};

// exp2 on the FMA/ALU pipes (Cody-Waite split + degree-3 minimax, max rel. error 7.5e-5): takes a share of the
// exponentials off the 16-op/clk MUFU unit, which is what bounds the forward kernel at d = 128.  Valid for x > -126.
__device__ __forceinline__ float ex2_poly3(float x) {
  const float magic = 12582912.f;  // 1.5 * 2^23: x + magic rounds x to the nearest integer in the low mantissa bits
  const float r = x + magic;
  const float f = x - (r - magic);  // f in [-0.5, 0.5]
  float p = fmaf(f, 5.517166745e-2f, 2.426111221e-1f);
  p = fmaf(p, f, 6.932609858e-1f);
  p = fmaf(p, f, 9.999280736e-1f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(r) << 23));
}
#ifndef SSVB_POLY_MOD_BWD
#define SSVB_POLY_MOD_BWD 0  // share of backward exponentials on the polynomial path (measured: no gain, off)
#endif

// weights of 32 consecutive columns [cb, cb+32) of one row: sv = S values, pk = packed bf16 pairs out
template <int MODE, bool MASKED, bool OPF16>
__device__ __forceinline__ void bwd_weights32(const uint32_t (&sv)[32], uint32_t (&pk)[16], const SimParams& p,
                                              int a_glob, int colbase, float rs, uint32_t cs32, float& wsum) {
#ifndef SSVB_BWD_SCALAR_MATH
  if (MODE == SIM_NTX_FIXED && !MASKED) {
    // issue-bound loop.  Operands are pre-scaled (the S accumulator is the log2-domain logit): per pair two MUFU.EX2
    // (or one packed polynomial), FADD2 (row + column factor), FMUL2, one pack -> 2.5 issue slots per element.
    // Column factors: fp32 pairs straight from shared memory (LDS.128 = 2 pairs, 32-bit shared address).
    const unsigned long long rs2 = pack_f32x2(rs, rs);
#pragma unroll
    for (int i2 = 0; i2 < 8; ++i2) {
      unsigned long long cu[2];  // column factors of 4 consecutive columns
      lds_v4(cs32 + i2 * 16, cu[0], cu[1]);
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int i = i2 * 2 + e;  // pair index: columns 2i, 2i+1
        const bool poly = (i % SSVB_POLY_BWD_MOD) < SSVB_POLY_BWD_CNT;
        unsigned long long e2;
        if (poly) {
          e2 = ex2_poly3_x2(pack_f32x2(__uint_as_float(sv[2 * i]), __uint_as_float(sv[2 * i + 1])));
        } else {
          e2 = pack_f32x2(SSVB_EX2(__uint_as_float(sv[2 * i])), SSVB_EX2(__uint_as_float(sv[2 * i + 1])));
        }
        const unsigned long long w2 = mul_f32x2(e2, add_f32x2(rs2, cu[e]));
        float w0, w1;
        unpack_f32x2(w2, w0, w1);
        pk[i] = pack_h2<OPF16>(w0, w1);
      }
    }
    return;
  }
#endif
#pragma unroll
  for (int i4 = 0; i4 < 8; ++i4) {
    float csv[4] = {0.f, 0.f, 0.f, 0.f};
    if (MODE != SIM_MOCO) {
      unsigned long long c01, c23;
      lds_v4(cs32 + i4 * 16, c01, c23);
      unpack_f32x2(c01, csv[0], csv[1]);
      unpack_f32x2(c23, csv[2], csv[3]);
    }
    float w[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int i = i4 * 4 + e;
      const float s = __uint_as_float(sv[i]);
      float wv;
      if (MODE == SIM_NTX_FIXED) {
        const float x = fmaf(s, p.c, -p.shift);
        const bool poly = !MASKED && SSVB_POLY_MOD_BWD > 0 && (i % (SSVB_POLY_MOD_BWD > 0 ? SSVB_POLY_MOD_BWD : 1)) == 1;
        wv = (poly ? ex2_poly3(x) : SSVB_EX2(x)) * (rs + csv[e]);
      } else if (MODE == SIM_NTX_ONLINE) {
        const float t = s * p.c;
        wv = ex2f(t - rs) + ex2f(t - csv[e]);
      } else {
        wv = ex2f(fmaf(s, p.c, -rs));
      }
      if (MASKED && MODE != SIM_MOCO && (colbase + i == a_glob)) wv = 0.f;
      if (MASKED && MODE == SIM_MOCO && (colbase + i >= p.cols)) wv = 0.f;  // queue tail (zero-filled B rows)
      if (MODE == SIM_MOCO) wsum += wv;  // fused MoCo forward: row sums of the un-normalised weights
      w[e] = wv;
    }
    pk[i4 * 2] = pack_h2<OPF16>(w[0], w[1]);
    pk[i4 * 2 + 1] = pack_h2<OPF16>(w[2], w[3]);
  }
}

// 640 threads: warpgroup 0 = {TMA producer, MMA issuer, 2 idle warps} at 40 registers/thread; warps 4..19 = FOUR weight
// warpgroups.  Tile t is processed by the warpgroup pair (t & 1); within the pair each warpgroup owns one 64-column
// half of the 128 x 128 tile (thread = row, TMEM lane quarter = warp % 4).  Four resident math warps per scheduler
// (instead of two) hide the TMEM-load / barrier / MUFU latencies of each other.
template <int KB, int MODE, bool OPF16 = false>
__global__ void __launch_bounds__(640, 1)
sim_bwd_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const SimParams p) {
  using C = BwdCfg<KB>;
  constexpr int BN = C::BN, NSTAGE = C::NSTAGE, DP = C::DP;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align1024(smem_raw);
  uint8_t* sA = smem;
  uint8_t* sB = smem + C::A_BYTES;
  uint8_t* sC = sB + NSTAGE * C::B_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sC + NSTAGE * C::CS_BYTES);
  uint64_t* a_full = bars;
  uint64_t* a_empty = bars + 1;
  uint64_t* b_full = bars + 2;
  uint64_t* b_empty = b_full + NSTAGE;
  uint64_t* s_full = b_empty + NSTAGE;
  uint64_t* s_empty = s_full + 2;
  uint64_t* w_full = s_empty + 2;
  uint64_t* w_empty = w_full + 2;
  uint64_t* dz_full = w_empty + 2;
  uint64_t* dz_empty = dz_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(dz_empty + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(a_full, 1);
    mbar_init(a_empty, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&s_empty[i], 8);
      mbar_init(&w_full[i], 8);
      mbar_init(&w_empty[i], 1);
    }
    for (int i = 0; i < NSTAGE; ++i) {
      mbar_init(&b_full[i], 1);
      mbar_init(&b_empty[i], 1);
    }
    mbar_init(dz_full, 1);
    mbar_init(dz_empty, 16);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int nunits = p.row_blocks * p.nchunks;

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (warp == 0) {
      // ---------------------------------------------------------------- TMA producer (whole warp, elect_one issues)
      int ucount = 0, gt = 0;
      for (int u = blockIdx.x; u < nunits; u += gridDim.x, ++ucount) {
        const UnitInfo ui = decode_unit(p, u);
        mbar_wait(a_empty, (ucount & 1) ^ 1);
        if (elect_one()) {
          mbar_expect_tx(a_full, C::A_BYTES);
#pragma unroll
          for (int kb = 0; kb < KB; ++kb) tma_load_2d(sA + kb * (128 * 128), &tmA, a_full, kb * 64, ui.g0);
        }
        __syncwarp();
        for (int t = ui.t0; t < ui.t1; ++t, ++gt) {
          const int st = gt % NSTAGE;
          mbar_wait(&b_empty[st], ((gt / NSTAGE) & 1) ^ 1);
          if (elect_one()) {
            constexpr int CSB = (MODE != SIM_MOCO) ? BN * 4 : 0;
            mbar_expect_tx(&b_full[st], C::B_BYTES + CSB);
#pragma unroll
            for (int kb = 0; kb < KB; ++kb)
              tma_load_2d(sB + st * C::B_BYTES + kb * (BN * 128), &tmB, &b_full[st], kb * 64, t * BN);
            if (MODE != SIM_MOCO) bulk_load_1d(sC + st * C::CS_BYTES, p.colstat + t * BN, CSB, &b_full[st]);
          }
          __syncwarp();
        }
      }
    } else if (warp == 1) {
      // ---------------------------------------------------------------- S-GEMM issuer (whole warp, elect_one issues).
      // The two GEMMs have their own issuer warps: a tile needs 16 tcgen05.mma of only 64 tensor-cycles each, and one
      // thread cannot issue them (plus commits and barrier waits) fast enough to keep the pipe full.
      constexpr uint32_t IDESC_S = make_idesc(128, BN, 0, 0, OPF16 ? 0 : 1);
      const uint32_t abase = smem_u32(sA);
      int ucount = 0, gt = 0;
      for (int u = blockIdx.x; u < nunits; u += gridDim.x, ++ucount) {
        const UnitInfo ui = decode_unit(p, u);
        mbar_wait(a_full, ucount & 1);
        for (int t = ui.t0; t < ui.t1; ++t, ++gt) {
          const int st = gt % NSTAGE, buf = gt & 1;
          mbar_wait(&b_full[st], (gt / NSTAGE) & 1);
          mbar_wait(&s_empty[buf], ((gt >> 1) & 1) ^ 1);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t bbase = smem_u32(sB + st * C::B_BYTES);
#ifndef SSVB_DBG_NOS
#pragma unroll
            for (int kb = 0; kb < KB; ++kb)
#pragma unroll
              for (int k4 = 0; k4 < 4; ++k4)
                umma_ss(tmem + C::T_S + buf * BN, desc_kmajor(abase + kb * (128 * 128) + k4 * 32),
                        desc_kmajor(bbase + kb * (BN * 128) + k4 * 32), IDESC_S, (kb | k4) != 0);
#endif
            umma_commit(&s_full[buf]);
            if (t == ui.t1 - 1) umma_commit(a_empty);  // last reader of this unit's A tile
          }
          __syncwarp();
        }
      }
    } else if (warp == 2) {
      // ---------------------------------------------------------------- dZ-GEMM issuer
      constexpr uint32_t IDESC_D = make_idesc(128, DP, 0, 1, OPF16 ? 0 : 1);  // A = W from TMEM, B = smem tile MN-major
      int ucount = 0, gt = 0;
      for (int u = blockIdx.x; u < nunits; u += gridDim.x, ++ucount) {
        const UnitInfo ui = decode_unit(p, u);
        mbar_wait(dz_empty, (ucount & 1) ^ 1);
        for (int t = ui.t0; t < ui.t1; ++t, ++gt) {
          const int st = gt % NSTAGE, buf = gt & 1;
          mbar_wait(&w_full[buf], (gt >> 1) & 1);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t bbase = smem_u32(sB + st * C::B_BYTES);
#ifndef SSVB_DBG_NOD
#pragma unroll
            for (int k = 0; k < BN / 16; ++k)
              umma_ts(tmem + C::T_DZ, tmem + C::T_W + buf * 64 + k * 8,
                      desc_mnmajor(bbase + (KB > 2 ? p.dhalf * 2 * (BN * 128) : 0) + k * (16 * 128), BN * 128), IDESC_D,
                      (t > ui.t0 || k > 0) ? 1u : 0u);
#endif
            umma_commit(&b_empty[st]);
            umma_commit(&w_empty[buf]);
            if (t == ui.t1 - 1) umma_commit(dz_full);
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ weight warpgroups + epilogue
    asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
    const int wgi = (warp - 4) >> 2;  // 0..3
    const int pair = wgi >> 1;        // which tiles (parity) / which S and W buffer
    const int half = wgi & 1;         // which 64-column half of the tile
    const int q = warp & 3;
    const int row_l = q * 32 + lane;
    const uint32_t tlane = static_cast<uint32_t>(q * 32) << 16;
    int gt = 0, ucount = 0;
    // barrier addresses in the shared window, computed once (a generic pointer costs a cvta sequence per use)
    const uint32_t a_s_full = smem_u32(&s_full[pair]), a_s_empty = smem_u32(&s_empty[pair]);
    const uint32_t a_w_full = smem_u32(&w_full[pair]), a_w_empty = smem_u32(&w_empty[pair]);
    const uint32_t a_dz_full = smem_u32(dz_full), a_dz_empty = smem_u32(dz_empty);
    const uint32_t a_sC = smem_u32(sC) + half * 64 * 4;
#ifdef SSVB_DBG_TIMING
    long long _t_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#endif
    SSVB_T0();
    for (int u = blockIdx.x; u < nunits; u += gridDim.x, ++ucount) {
      const UnitInfo ui = decode_unit(p, u);
      const int a_glob = ui.g0 + row_l;
      const bool valid = row_l < ui.nvalid;
      float rs = 0.f;
      if (valid) rs = (MODE == SIM_MOCO) ? (p.rowstat ? p.rowstat[ui.lrow0 + row_l] : p.shift) : p.rowstat[a_glob];
      float lsum = 0.f;  // SIM_MOCO with p.part_l: sum over this unit's columns of exp2(l - shift) (fused forward)
      for (int t = ui.t0; t < ui.t1; ++t, ++gt) {
        if ((gt & 1) != pair) continue;
        const int st = gt % NSTAGE;
        SSVB_TP(0);  // loop overhead / other
        // no separate wait on b_full: S(t) was issued only after the MMA warp observed b_full[st], so observing
        // s_full(t) also orders this warp after the TMA / bulk-copy writes of the B tile and its column statistics
        SSVB_TP(1);
        mbar_wait_a(a_s_full, (gt >> 1) & 1);
        tc_fence_after();
        SSVB_TP(2);  // wait s_full
        const uint32_t t_s = tmem + tlane + C::T_S + pair * BN + half * 64;
        uint32_t sv[2][32];
#ifndef SSVB_DBG_NOLD
        tmem_ld_x32(t_s, sv[0]);
        tmem_ld_x32(t_s + 32, sv[1]);
        tmem_ld_wait_regs(sv[0]);
        tmem_ld_wait_regs(sv[1]);
#else
#pragma unroll
        for (int i = 0; i < 32; ++i) { sv[0][i] = __float_as_uint(1e-3f * (i + lane)); sv[1][i] = sv[0][i]; }
#endif
        // this half of the S tile is in registers: give the buffer back so S(t+2) can be issued right away
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_a(a_s_empty);
        SSVB_TP(3);  // tmem loads + release
        const int j0 = t * BN + half * 64;
        const bool special = (MODE != SIM_MOCO) ? ((j0 < ui.g0 + 128) && (j0 + 64 > ui.g0)) : (j0 + 64 > p.cols);
        const uint32_t t_w = tmem + tlane + C::T_W + pair * 64 + half * 32;
        const uint32_t cs = a_sC + st * C::CS_BYTES;
        constexpr int CSTEP = 32 * 4;
        // all 64 weights are computed and packed before waiting for the W buffer (the dZ GEMM of tile t-2 may still
        // be reading it): that wait is off the critical path unless the tensor pipe is the bottleneck
        uint32_t pk[2][16];
        if (special) {
          bwd_weights32<MODE, true, OPF16>(sv[0], pk[0], p, a_glob, j0, rs, cs, lsum);
          bwd_weights32<MODE, true, OPF16>(sv[1], pk[1], p, a_glob, j0 + 32, rs, cs + CSTEP, lsum);
        } else {
          bwd_weights32<MODE, false, OPF16>(sv[0], pk[0], p, a_glob, j0, rs, cs, lsum);
          bwd_weights32<MODE, false, OPF16>(sv[1], pk[1], p, a_glob, j0 + 32, rs, cs + CSTEP, lsum);
        }
        SSVB_TP(4);  // 64 weights
        mbar_wait_a(a_w_empty, ((gt >> 1) & 1) ^ 1);
        tc_fence_after();
        SSVB_TP(5);  // wait w_empty
        tmem_st_x16(t_w, pk[0]);
        tmem_st_x16(t_w + 16, pk[1]);
        SSVB_TP(6);  // stores issued
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_a(a_w_full);
        SSVB_TP(7);  // st drain + arrive
      }
      if (MODE == SIM_MOCO && p.part_l != nullptr && valid)
        p.part_l[static_cast<size_t>(ui.ch * 4 + wgi) * p.part_stride + ui.lrow0 + row_l] = lsum;
      // ---- epilogue: the four warpgroups drain DP/4 accumulator columns each (DP = 64: only two of them have work)
      mbar_wait_a(a_dz_full, ucount & 1);
      tc_fence_after();
      if (DP == 128 || wgi < 2) {
        const int col0 = wgi * 32;
        uint32_t v[32];
        tmem_ld_x32(tmem + tlane + C::T_DZ + col0, v);
        tmem_ld_wait_regs(v);
        if (valid) {
          float* dst = p.dacc + static_cast<size_t>(ui.lrow0 + row_l) * p.ld_dacc + col0 + (KB > 2 ? p.dhalf * 128 : 0);
          if (p.use_atomic) {
#pragma unroll
            for (int i = 0; i < 8; ++i)
              red_add_v4(dst + 4 * i, __uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]),
                         __uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3]));
          } else {
#pragma unroll
            for (int i = 0; i < 8; ++i)
              reinterpret_cast<float4*>(dst)[i] =
                  make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]),
                              __uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3]));
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_a(a_dz_empty);
    }
#ifdef SSVB_DBG_TIMING
    if (p.dbg && blockIdx.x == 0 && warp == 4 && lane == 0)
      for (int i = 0; i < 8; ++i) p.dbg[i] = static_cast<unsigned long long>(_t_acc[i]);
#endif
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<512>(tmem);
}

}  // namespace ssvb
