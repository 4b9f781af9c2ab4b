// gemm_tc.cu -- persistent, warp-specialised, grouped bf16 GEMM on the 5th-gen tensor cores (sm_100a):
//   TMA (cp.async.bulk.tensor, 128B swizzle) -> shared memory ring -> tcgen05.mma (one issuing thread, fp32
//   accumulators in TMEM, double-buffered) -> tcgen05.ld -> fused PSGD epilogue (common.cuh:Epi) -> global.
//
// One launch runs up to TC_MAX_PROBLEMS independent GEMMs ("group"): the left- and right-factor Grams /
// Newton-Schulz products of one Kron update, or the same product of several parameters.  Their tiles share one
// flat index space that the persistent CTAs (one per SM) walk round-robin, so small problems fill each other's
// wave-quantisation bubbles.
//
// Operand transposition is done by the tensor core, not by a copy: a row-major matrix whose contraction index
// runs along its rows ("K-major") and one whose contraction index runs down its columns ("MN-major") are both
// fed through 128B-swizzled smem tiles; the UMMA instruction descriptor carries the per-operand major bit and
// the smem descriptor the matching (LBO, SBO) strides.
//
// Tile: 128 x BN x 64 (BN = 256 or 128), cta_group::1, UMMA 128 x BN x 16, 4 (BN=256) / 6 (BN=128) smem stages.
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = MMA issuer + TMEM owner, warps 2-5 = epilogue
// (warp w may only touch TMEM lanes 32*(w%4) .. +31).
#include <cuda.h>
#include <stdio.h>
#include <string.h>

#include <stdlib.h>
#include "common.cuh"
#include "tc_ptx.cuh"

namespace psgd {

constexpr int TC_MAX_PROBLEMS = 4;
constexpr int TC_BM = 128;
constexpr int TC_BK = 64;
constexpr int TC_THREADS = 224;   // warp 0: TMA producer of A, warp 1: MMA issuer, warps 2-5: epilogue, warp 6: TMA producer of B
constexpr int TC2_THREADS = 192;

struct alignas(64) TcProblem {
  CUtensorMap map_a;
  CUtensorMap map_b;
  Epi epi;
  int M, N, K;
  int a_mn, b_mn;  // operand is MN-major in memory
  int a_3d, b_3d;  // MN-major operand whose MN extent is a multiple of 64: the whole tile is ONE 3-D TMA box {64, 64 k-rows, chunks}
  int sym;         // symmetric output: only tiles that reach the upper triangle are enumerated
  int tiles_m, tiles_n;
  int tile_start;  // first flat work-unit index of this problem
  int tile_first;  // first tile (in this problem's own enumeration) this entry covers (tail entries start past 0)
  int splits;      // K partitions per tile (1 = none); partial accumulators meet in the fp32 workspace
  int kb_split;    // k-blocks per partition
  int ws_slot0;    // first workspace tile slot of this entry
  int nsplit;      // BN / bn_eff: a full-width tile position is covered by nsplit consecutive units
  int bn_eff;      // tile width of this entry: BN, or BN/2 = 128 for tail-wave entries (half tiles fill the last wave without a reduction)
  CUtensorMap map_b_half;  // K-major B with a 128-row box (only used when bn_eff < BN)
};

struct alignas(64) TcGroup {
  TcProblem p[TC_MAX_PROBLEMS];
  int num_problems;
  int total_tiles;
  int mn_lbo, mn_sbo;
  int* error_flag;
  float* ws;        // split-K workspace: slots of 128 x BN fp32, zero between launches
  int* ws_count;    // arrival counter per slot, zero between launches
};

// flat work unit -> (problem, tile_m, tile_n in units of bn_eff, K slice).  A unit is (full-width tile position, N half, K slice); tile
// positions follow an m-grouped rasterisation (8 row-tiles per group) for L2 reuse, symmetric problems enumerate the tiles that reach
// the upper triangle only.
__device__ __forceinline__ void locate_tile(const TcGroup& g, int t, int bn_full, int& pi, int& tm, int& tn, int& kslice) {
  pi = 0;
#pragma unroll
  for (int i = 1; i < TC_MAX_PROBLEMS; ++i)
    if (i < g.num_problems && t >= g.p[i].tile_start) pi = i;
  const TcProblem& p = g.p[pi];
  int u = t - p.tile_start;
  kslice = u % p.splits;
  u /= p.splits;
  const int nhalf = u % p.nsplit;
  int lt = p.tile_first + u / p.nsplit;
  if (p.sym) {
    // row tm owns the tiles tn >= tn_min(tm) = floor(tm * 128 / BN): those whose last column reaches the diagonal block of row tm
    tm = 0;
    while (true) {
      const int cnt = p.tiles_n - (tm * TC_BM) / bn_full;
      if (lt < cnt) break;
      lt -= cnt;
      ++tm;
    }
    tn = ((tm * TC_BM) / bn_full + lt) * p.nsplit + nhalf;
    return;
  }
  const int GROUP = 8;
  int per_group = GROUP * p.tiles_n;
  int gidx = lt / per_group;
  int first_m = gidx * GROUP;
  int gsz = min(GROUP, p.tiles_m - first_m);
  int in_g = lt - gidx * per_group;
  tm = first_m + in_g % gsz;
  tn = (in_g / gsz) * p.nsplit + nhalf;
}

// ------------------------------------------------------------------------------------------------
// epilogue for one 32-column chunk of one accumulator row
// ------------------------------------------------------------------------------------------------
struct RowAcc {
  float row_sumsq, tot, amax, tr, dmax, dot;
};

// mirror: 0 = plain block, 1 = strictly-upper block of a symmetric output (also writes C[col][row] and credits the column sums to
// row_sumsq[col], i.e. the row norms of the mirrored block)
// v[j] += f * X[row, col0 + j] (32 columns, 16-byte loads; N % 8 == 0 so every vector is fully in or out of range)
__device__ __forceinline__ void epi_add_tile(float* v, const void* X, int x_dtype, int ldx, int row, int col0, int N, float f) {
  if (x_dtype == PSGD_BF16) {
    const bf16* xp = reinterpret_cast<const bf16*>(X) + (size_t)row * ldx + col0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if (col0 + q * 8 < N) {
        uint4 u = *reinterpret_cast<const uint4*>(xp + q * 8);
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          float2 x = __bfloat1622float2(h[t]);
          v[q * 8 + 2 * t] = fmaf(f, x.x, v[q * 8 + 2 * t]);
          v[q * 8 + 2 * t + 1] = fmaf(f, x.y, v[q * 8 + 2 * t + 1]);
        }
      }
    }
  } else {
    const float* xp = reinterpret_cast<const float*>(X) + (size_t)row * ldx + col0;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      if (col0 + q * 4 < N) {
        float4 x = *reinterpret_cast<const float4*>(xp + q * 4);
        v[q * 4] = fmaf(f, x.x, v[q * 4]); v[q * 4 + 1] = fmaf(f, x.y, v[q * 4 + 1]);
        v[q * 4 + 2] = fmaf(f, x.z, v[q * 4 + 2]); v[q * 4 + 3] = fmaf(f, x.w, v[q * 4 + 3]);
      }
    }
  }
}

__device__ __forceinline__ void epilogue_chunk(const Epi& e, int M, int N, int row, int col0, uint32_t* raw, float alpha,
                                               float beta, float beta2, int lane, RowAcc& ra, int mirror) {
  float v[32];
  const bool row_ok = row < M;
  float rs = (e.row_scale && row_ok) ? e.row_scale[row] : 1.f;
  if (e.norm_axis == 1 && row_ok) rs *= epi_norm_factor(e, row);
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(raw[j]) * alpha * rs;
  if (e.col_scale) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] *= (col0 + j < N) ? e.col_scale[col0 + j] : 0.f;
  }
  if (e.norm_axis == 2) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] *= (col0 + j < N) ? epi_norm_factor(e, col0 + j) : 0.f;
  }
  if (e.D && row_ok) {
    const float drs = beta * (e.d_row_scale ? e.d_row_scale[row] : 1.f);
    if (e.d_dtype == PSGD_BF16) {
      const bf16* dp = reinterpret_cast<const bf16*>(e.D) + (size_t)row * e.ldd + col0;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (col0 + q * 8 < N) {
          uint4 u = *reinterpret_cast<const uint4*>(dp + q * 8);
          const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            float2 f = __bfloat1622float2(h[t]);
            const int cc = col0 + q * 8 + 2 * t;
            v[q * 8 + 2 * t] += drs * f.x * (e.d_col_scale ? e.d_col_scale[cc] : 1.f);
            v[q * 8 + 2 * t + 1] += drs * f.y * (e.d_col_scale ? e.d_col_scale[cc + 1] : 1.f);
          }
        }
      }
    } else {
      const float* dp = reinterpret_cast<const float*>(e.D) + (size_t)row * e.ldd + col0;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        if (col0 + q * 4 < N) {
          float4 f = *reinterpret_cast<const float4*>(dp + q * 4);
          const int cc = col0 + q * 4;
          v[q * 4] += drs * f.x * (e.d_col_scale ? e.d_col_scale[cc] : 1.f); v[q * 4 + 1] += drs * f.y * (e.d_col_scale ? e.d_col_scale[cc + 1] : 1.f);
          v[q * 4 + 2] += drs * f.z * (e.d_col_scale ? e.d_col_scale[cc + 2] : 1.f); v[q * 4 + 3] += drs * f.w * (e.d_col_scale ? e.d_col_scale[cc + 3] : 1.f);
        }
      }
    }
  }
  if (e.D2 && row_ok) epi_add_tile(v, e.D2, e.d_dtype, e.ldd2, row, col0, N, beta2);
  float diag_unrounded = 0.f;
  const bool on_diag = row >= col0 && row < col0 + 32;
  if (e.diag_resid && on_diag) {
#pragma unroll
    for (int j = 0; j < 32; ++j)
      asm("{\n\t.reg .pred p;\n\tsetp.eq.s32 p, %2, %3;\n\tselp.f32 %0, %1, %0, p;\n\t}" : "+f"(diag_unrounded) : "f"(v[j]), "r"(row - col0), "r"(j));
  }
  // round + store (N is a multiple of 8, so every 8-column vector is fully in or fully out of range)
  if (e.out_dtype == PSGD_BF16) {
    bf16* cp = reinterpret_cast<bf16*>(e.C) + (size_t)row * e.ldc + col0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      uint4 u;
      __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        h[t] = __floats2bfloat162_rn(v[q * 8 + 2 * t], v[q * 8 + 2 * t + 1]);
        float2 f = __bfloat1622float2(h[t]);
        v[q * 8 + 2 * t] = f.x;
        v[q * 8 + 2 * t + 1] = f.y;
      }
      if (row_ok && col0 + q * 8 < N) *reinterpret_cast<uint4*>(cp + q * 8) = u;
    }
  } else {
    float* cp = reinterpret_cast<float*>(e.C) + (size_t)row * e.ldc + col0;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      if (row_ok && col0 + q * 4 < N)
        *reinterpret_cast<float4*>(cp + q * 4) = make_float4(v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
    }
  }
  if (e.diag_resid && on_diag && row_ok) {
    float dr = 0.f;
#pragma unroll
    for (int j = 0; j < 32; ++j)
      asm("{\n\t.reg .pred p;\n\tsetp.eq.s32 p, %2, %3;\n\tselp.f32 %0, %1, %0, p;\n\t}" : "+f"(dr) : "f"(v[j]), "r"(row - col0), "r"(j));
    e.diag_resid[row] = diag_unrounded - dr;
  }
  if (mirror) {  // C[col0 + j][row] = v[j]: for a fixed j the 32 lanes (consecutive rows) write 64 contiguous bytes
    if (e.out_dtype == PSGD_BF16) {
      bf16* cp = reinterpret_cast<bf16*>(e.C) + (size_t)col0 * e.ldc + row;
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (row_ok && col0 + j < N) cp[(size_t)j * e.ldc] = __float2bfloat16_rn(v[j]);
    } else {
      float* cp = reinterpret_cast<float*>(e.C) + (size_t)col0 * e.ldc + row;
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (row_ok && col0 + j < N) cp[(size_t)j * e.ldc] = v[j];
    }
  }
  // reductions over the rounded values; out-of-range elements contribute 0
  const bool need_red = e.row_sumsq || e.col_sumsq || e.total_sumsq || e.abs_max || e.trace || e.diag_max || e.dot_out;
  if (!need_red) return;
#pragma unroll
  for (int j = 0; j < 32; ++j)
    if (!row_ok || col0 + j >= N) v[j] = 0.f;
  if (e.dot_out && row_ok) {
    float m[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) m[j] = 0.f;
    epi_add_tile(m, e.dotm, e.d_dtype, e.ld_dot, row, col0, N, 1.f);
    float d = 0.f;
#pragma unroll
    for (int j = 0; j < 32; ++j) d = fmaf(v[j], m[j], d);
    ra.dot += d;
  }
  float s = 0.f, am = 0.f;
#pragma unroll
  for (int j = 0; j < 32; ++j) { s = fmaf(v[j], v[j], s); am = fmaxf(am, fabsf(v[j])); }
  ra.row_sumsq += s;
  ra.tot += mirror ? 2.f * s : s;
  ra.amax = fmaxf(ra.amax, am);
  if ((e.trace || e.diag_max) && row >= col0 && row < col0 + 32) {
    float dv = 0.f;
#pragma unroll
    for (int j = 0; j < 32; ++j)  // opaque selp: nvcc turns a C++ select chain into a local-memory indexed load
      asm("{\n\t.reg .pred p;\n\tsetp.eq.s32 p, %2, %3;\n\tselp.f32 %0, %1, %0, p;\n\t}" : "+f"(dv) : "f"(v[j]), "r"(row - col0), "r"(j));
    ra.tr += dv;
    ra.dmax = fmaxf(ra.dmax, dv);
  }
  float* csq = mirror ? e.row_sumsq : e.col_sumsq;
  if (csq) {
    // transpose-reduce: after the 5 halving steps lane l holds sum over the warp's 32 rows of column col0 + l
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = v[j] * v[j];
#define PSGD_TR_STEP(OFF)                                              \
    {                                                                  \
      const bool upper = (lane & (OFF)) != 0;                          \
      _Pragma("unroll") for (int i = 0; i < (OFF); ++i) {              \
        float send = upper ? v[i] : v[i + (OFF)];                      \
        float keep = upper ? v[i + (OFF)] : v[i];                      \
        float recv = __shfl_xor_sync(0xffffffffu, send, (OFF));        \
        v[i] = keep + recv;                                            \
      }                                                                \
    }
    PSGD_TR_STEP(16) PSGD_TR_STEP(8) PSGD_TR_STEP(4) PSGD_TR_STEP(2) PSGD_TR_STEP(1)
#undef PSGD_TR_STEP
    if (col0 + lane < N) atomicAdd(&csq[col0 + lane], v[0]);
  }
}

// ------------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------------
template <int BN>
struct TcCfg {
  static constexpr int STAGES = (BN == 256) ? 4 : 6;
  static constexpr int A_BYTES = TC_BM * TC_BK * 2;
  static constexpr int B_BYTES = BN * TC_BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BAR_BYTES = 256;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + BAR_BYTES + 1024;  // +1024 for manual alignment
  static constexpr int TMEM_COLS = 2 * BN;
};

template <int BN>
__global__ void __launch_bounds__(TC_THREADS, 1) gemm_tc_kernel(const __grid_constant__ TcGroup g) {
  using Cfg = TcCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar_base = smem_base + Cfg::STAGES * Cfg::STAGE_BYTES;
  // barrier layout (8 B each): full[S], empty[S], tfull[2], tempty[2], then tmem ptr (4 B)
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (Cfg::STAGES + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * Cfg::STAGES + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * Cfg::STAGES + 2 + s); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * Cfg::STAGES + 4);
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + Cfg::STAGES * Cfg::STAGE_BYTES + 8 * (2 * Cfg::STAGES + 4));
  volatile int* epi_flag = reinterpret_cast<volatile int*>(smem_gen + Cfg::STAGES * Cfg::STAGE_BYTES + 8 * (2 * Cfg::STAGES + 4) + 8);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < Cfg::STAGES; ++s) { mbar_init(full_bar(s), 2); mbar_init(empty_bar(s), 1); }   // full: one arrival per producer warp
    for (int s = 0; s < 2; ++s) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), 4); }
    fence_barrier_init();
    for (int i = 0; i < g.num_problems; ++i) { prefetch_tmap(&g.p[i].map_a); prefetch_tmap(&g.p[i].map_b); }
  }
  if (warp == 1) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;

  if (warp == 0 || warp == 6) {
    // ===================== TMA producers: warp 0 streams A, warp 6 streams B =====================
    // (the issue rate of a single producer thread is a measurable limit: every TMA op costs tens of cycles next to the ~90-cycle
    //  barrier probe, and a 128 x 256 x 64 k-block leaves only 512 cycles)
    {
      const bool do_a = warp == 0;
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < g.total_tiles; t += gridDim.x) {
        int pi, tm, tn, kslice;
        locate_tile(g, t, BN, pi, tm, tn, kslice);
        const TcProblem& p = g.p[pi];
        const int num_kb = (p.K + TC_BK - 1) / TC_BK;
        const int kb0 = kslice * p.kb_split, kb1 = min(num_kb, kb0 + p.kb_split);
        const int bn = p.bn_eff;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u, g.error_flag);
          const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
          const uint32_t sb = sa + Cfg::A_BYTES;
          if (!elect_one()) {
            // not the issuing lane
          } else if (do_a) {
            mbar_arrive_expect_tx(full_bar(stage), Cfg::A_BYTES);
            if (!p.a_mn) {
              tma_load_2d(&p.map_a, full_bar(stage), sa, kb * TC_BK, tm * TC_BM);
            } else if (p.a_3d) {
              tma_load_3d(&p.map_a, full_bar(stage), sa, 0, kb * TC_BK, tm * (TC_BM / 64));
            } else {
#pragma unroll
              for (int c = 0; c < TC_BM / 64; ++c)
                tma_load_2d(&p.map_a, full_bar(stage), sa + c * (TC_BK * 128), tm * TC_BM + c * 64, kb * TC_BK);
            }
          } else {
            mbar_arrive_expect_tx(full_bar(stage), bn * TC_BK * 2);
            if (!p.b_mn) {
              tma_load_2d(bn == BN ? &p.map_b : &p.map_b_half, full_bar(stage), sb, kb * TC_BK, tn * bn);
            } else if (p.b_3d) {
              tma_load_3d(bn == BN ? &p.map_b : &p.map_b_half, full_bar(stage), sb, 0, kb * TC_BK, tn * (bn / 64));
            } else {
              for (int c = 0; c < bn / 64; ++c)
                tma_load_2d(&p.map_b, full_bar(stage), sb + c * (TC_BK * 128), tn * bn + c * 64, kb * TC_BK);
            }
          }
          __syncwarp();
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (the whole warp runs the loop; one elected lane issues) =====================
    {
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int t = blockIdx.x; t < g.total_tiles; t += gridDim.x) {
        int pi, tm, tn, kslice;
        locate_tile(g, t, BN, pi, tm, tn, kslice);
        const TcProblem& p = g.p[pi];
        const int num_kb = (p.K + TC_BK - 1) / TC_BK;
        // instruction descriptor: D=f32 (bit4), A=bf16 (bit7), B=bf16 (bit10), a_major bit15, b_major bit16,
        // N>>3 at bits 17-22, M>>4 at bits 24-28
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (uint32_t(p.a_mn) << 15) | (uint32_t(p.b_mn) << 16) |
                               (uint32_t(p.bn_eff >> 3) << 17) | (uint32_t(TC_BM >> 4) << 24);
        const uint32_t a_lbo = p.a_mn ? (uint32_t)g.mn_lbo : 0u, a_sbo = p.a_mn ? (uint32_t)g.mn_sbo : 1024u;
        const uint32_t b_lbo = p.b_mn ? (uint32_t)g.mn_lbo : 0u, b_sbo = p.b_mn ? (uint32_t)g.mn_sbo : 1024u;
        const uint32_t a_kstep = p.a_mn ? 16u * 128u : 32u;  // bytes to advance per UMMA_K=16
        const uint32_t b_kstep = p.b_mn ? 16u * 128u : 32u;
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u, g.error_flag);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + uint32_t(acc * BN);
        const int kb0 = kslice * p.kb_split, kb1 = min(num_kb, kb0 + p.kb_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(full_bar(stage), phase, g.error_flag);
          int nstage = stage + 1; uint32_t nphase = phase;
          if (nstage == Cfg::STAGES) { nstage = 0; nphase ^= 1u; }
          tc_fence_after();
          const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
          const uint32_t sb = sa + Cfg::A_BYTES;
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < TC_BK / 16; ++k) {
              const uint64_t adesc = make_smem_desc(sa + k * a_kstep, a_lbo, a_sbo);
              const uint64_t bdesc = make_smem_desc(sb + k * b_kstep, b_lbo, b_sbo);
              umma_bf16(d_tmem, adesc, bdesc, idesc, ((kb - kb0) | k) != 0 ? 1u : 0u);
            }
            umma_commit(empty_bar(stage));  // smem slot is free once these MMAs have read it
            if (kb + 1 == kb1) umma_commit(tfull_bar(acc));  // accumulator complete -> epilogue
          }
          __syncwarp();
          stage = nstage; phase = nphase;
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    const int quarter = warp & 3;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int t = blockIdx.x; t < g.total_tiles; t += gridDim.x) {
      int pi, tm, tn, kslice;
      locate_tile(g, t, BN, pi, tm, tn, kslice);
      const TcProblem& p = g.p[pi];
      const Epi& e = p.epi;
      float alpha = e.alpha * (e.alpha_ptr ? *e.alpha_ptr : 1.f);
      const float beta = e.beta * (e.beta_ptr ? *e.beta_ptr : 1.f);
      float beta2 = e.beta2;
      if (e.pro_fs) { const float a = epi_procrustes_step(e); alpha *= 0.5f * a * a; beta2 = a; }
      mbar_wait_relaxed(tfull_bar(acc), acc_phase, g.error_flag);
      tc_fence_after();
      const int row = tm * TC_BM + quarter * 32 + lane;
      RowAcc ra = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      const int bn = p.bn_eff;
      const int nchunks = bn / 32;
      bool run_epilogue = true;
      const float4* ws_tile = nullptr;   // partial tiles of this output tile, fragment layout: [split][chunk][q][warp][lane] float4
      const int my_split = kslice;
      const size_t part_f4 = (size_t)TC_BM * BN / 4;   // float4 per partial tile slot
      if (p.splits > 1) {
        // split-K unit: publish the partial accumulator (every warp-level store is 512 contiguous bytes), then count arrivals; the unit
        // that arrives last sums the other partials with its own accumulator (still in TMEM) in split order and runs the epilogue
        const int tile_local = (t - p.tile_start) / p.splits;   // (tile position, N half) index inside this entry
        float4* tbase = reinterpret_cast<float4*>(g.ws) + (size_t)(p.ws_slot0 + tile_local * p.splits) * part_f4;
        ws_tile = tbase;
        float4* mine = tbase + (size_t)my_split * part_f4;
#pragma unroll 1
        for (int c = 0; c < nchunks; ++c) {
          const int col0 = tn * bn + c * 32;
          if (col0 >= p.N) break;
          if (p.sym && (col0 >> 7) < tm) continue;
          uint32_t raw[32];
          const uint32_t taddr = tmem_base + (uint32_t(quarter * 32) << 16) + uint32_t(acc * BN + c * 32);
          tmem_ld_32x32(taddr, raw);
          tmem_ld_wait();
#pragma unroll
          for (int q = 0; q < 8; ++q)
            __stcg(mine + ((c * 8 + q) * 4 + quarter) * 32 + lane,
                   make_float4(__uint_as_float(raw[4 * q]), __uint_as_float(raw[4 * q + 1]), __uint_as_float(raw[4 * q + 2]), __uint_as_float(raw[4 * q + 3])));
        }
        __threadfence();
        asm volatile("bar.sync 1, 128;" ::: "memory");  // the four epilogue warps
        if (threadIdx.x == 64) {
          const int cslot = p.ws_slot0 + tile_local;
          const int old = atomicAdd(&g.ws_count[cslot], 1);
          *epi_flag = (old == p.splits - 1) ? 1 : 0;
          if (old == p.splits - 1) g.ws_count[cslot] = 0;
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        run_epilogue = (*epi_flag != 0);
        if (run_epilogue) __threadfence();
      }
      if (run_epilogue) {
#pragma unroll 1
        for (int c = 0; c < nchunks; ++c) {
          const int col0 = tn * bn + c * 32;
          int mirror = 0;
          if (p.sym) {  // 128-block classification: below the diagonal block -> produced by the mirror of its transpose, skip
            const int cb = col0 >> 7;
            if (cb < tm) continue;
            mirror = cb > tm;
          }
          if (p.splits > 1 && col0 >= p.N) break;
          uint32_t raw[32];
          const uint32_t taddr = tmem_base + (uint32_t(quarter * 32) << 16) + uint32_t(acc * BN + c * 32);
          tmem_ld_32x32(taddr, raw);
          if (p.splits > 1) {
            // other splits' partials: issue all loads of a half chunk before touching them (latency, not bandwidth, is the cost here)
            float sum[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) sum[j] = 0.f;
            bool own_added = false;
#pragma unroll
            for (int half = 0; half < 2; ++half) {
              float4 buf[3][4];
              for (int s0 = 0; s0 < p.splits; s0 += 3) {     // up to 3 foreign partials per batch
#pragma unroll
                for (int u = 0; u < 3; ++u) {
                  const int sp = s0 + u;
                  if (sp < p.splits && sp != my_split) {
                    const float4* src = ws_tile + (size_t)sp * part_f4;
#pragma unroll
                    for (int q = 0; q < 4; ++q) buf[u][q] = __ldcg(src + ((c * 8 + half * 4 + q) * 4 + quarter) * 32 + lane);
                  }
                }
#pragma unroll
                for (int u = 0; u < 3; ++u) {
                  const int sp = s0 + u;
                  if (sp < p.splits && sp != my_split) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                      const int j = (half * 4 + q) * 4;
                      sum[j] += buf[u][q].x; sum[j + 1] += buf[u][q].y; sum[j + 2] += buf[u][q].z; sum[j + 3] += buf[u][q].w;
                    }
                  }
                }
              }
            }
            (void)own_added;
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) raw[j] = __float_as_uint(sum[j] + __uint_as_float(raw[j]));
          } else {
            tmem_ld_wait();
          }
          if (col0 < p.N) epilogue_chunk(e, p.M, p.N, row, col0, raw, alpha, beta, beta2, lane, ra, mirror);
        }
      }
      // release the accumulator buffer to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));
      // per-row / per-warp reductions
      if (e.row_sumsq && row < p.M) atomicAdd(&e.row_sumsq[row], ra.row_sumsq);
      if (e.total_sumsq) { float s = warp_sum(ra.tot); if (lane == 0) atomicAdd(e.total_sumsq, s); }
      if (e.dot_out) { float s = warp_sum(ra.dot); if (lane == 0 && s != 0.f) atomicAdd(e.dot_out, s); }
      if (e.abs_max) { float s = warp_max(ra.amax); if (lane == 0) atomic_max_nonneg(e.abs_max, s); }
      if (e.trace) { float s = warp_sum(ra.tr); if (lane == 0 && s != 0.f) atomicAdd(e.trace, s); }
      if (e.diag_max) { float s = warp_max(ra.dmax); if (lane == 0) atomic_max_nonneg(e.diag_max, s); }
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, Cfg::TMEM_COLS); }
}


// ================================================================================================================
// 2-CTA variant (cta_group::2): a cluster of two CTAs (one TPC) owns one 256 x 256 output tile.  Each CTA stages only ITS 128 rows
// of A and ITS half (128 rows) of the B tile -- 32 KB per k-block per SM instead of 48 KB, which is what the 1-CTA kernel is
// bound by (L2->SM and smem-read bandwidth; profiles/r01_ncu_full_gemm_tc_raw.csv: tensor pipe 75 %) -- and the leader CTA's
// single thread issues tcgen05.mma.cta_group::2 with M = 256 whose accumulator halves live in the two CTAs' TMEM.
//   full[s]   (leader): armed with both CTAs' bytes; both CTAs' TMA loads complete_tx on it
//   empty[s]  (each CTA): tcgen05.commit multicast from the leader frees the slot in both CTAs
//   tfull[a]  (each CTA): commit multicast publishes the accumulator to both epilogues
//   tempty[a] (leader): 8 arrivals, the 4 epilogue warps of each CTA (the peer's arrive remotely through mapa)
// ================================================================================================================
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(const CUtensorMap* map, uint32_t leader_bar, uint32_t dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(leader_bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_bf16_2sm(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_2sm(uint32_t bar) {  // arrives on the barrier at this offset in BOTH CTAs of the pair
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"((uint16_t)3)
               : "memory");
}

constexpr int TC2_BM = 256;   // rows of the pair tile
constexpr int TC2_BN = 256;

// flat pair-tile -> (problem, pair row pm, tn)
__device__ __forceinline__ void locate_tile2(const TcGroup& g, int t, int& pi, int& pm, int& tn) {
  pi = 0;
#pragma unroll
  for (int i = 1; i < TC_MAX_PROBLEMS; ++i)
    if (i < g.num_problems && t >= g.p[i].tile_start) pi = i;
  const TcProblem& p = g.p[pi];
  int lt = t - p.tile_start;
  if (p.sym) {   // 256 x 256 pair tiles: row pm owns tn >= pm
    pm = 0;
    while (true) {
      const int cnt = p.tiles_n - pm;
      if (lt < cnt) break;
      lt -= cnt;
      ++pm;
    }
    tn = pm + lt;
    return;
  }
  const int GROUP = 4;
  int per_group = GROUP * p.tiles_n;
  int gidx = lt / per_group;
  int first_m = gidx * GROUP;
  int gsz = min(GROUP, p.tiles_m - first_m);
  int in_g = lt - gidx * per_group;
  pm = first_m + in_g % gsz;
  tn = in_g / gsz;
}

template <int BN_>
struct Tc2Cfg {
  static constexpr int BN = BN_;                           // 256, or 128 for launches whose last 256-wide wave would be mostly empty
  static constexpr int STAGES = BN_ == 256 ? 6 : 8;
  static constexpr int A_BYTES = 128 * TC_BK * 2;          // this CTA's 128 rows of A
  static constexpr int B_BYTES = (BN / 2) * TC_BK * 2;     // this CTA's half of B
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;    // 32 KB
  static constexpr int BAR_BYTES = 256;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + BAR_BYTES + 1024;
  static constexpr int TMEM_COLS = 2 * BN;
};

template <int BN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC2_THREADS, 1) gemm_tc2_kernel(const __grid_constant__ TcGroup g) {
  using Cfg = Tc2Cfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar_base = smem_base + Cfg::STAGES * Cfg::STAGE_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (Cfg::STAGES + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * Cfg::STAGES + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * Cfg::STAGES + 2 + s); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * Cfg::STAGES + 4);
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + Cfg::STAGES * Cfg::STAGE_BYTES + 8 * (2 * Cfg::STAGES + 4));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < Cfg::STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), 8); }
    fence_barrier_init();
    for (int i = 0; i < g.num_problems; ++i) { prefetch_tmap(&g.p[i].map_a); prefetch_tmap(&g.p[i].map_b); }
  }
  if (warp == 1) tmem_alloc_2sm(tmem_slot, Cfg::TMEM_COLS);
  tc_fence_before();
  cluster_sync_all();     // both CTAs' barriers are initialised before any remote arrive / multicast commit / peer TMA
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs; bytes are credited to the leader's full barrier) =====================
    {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = pair; t < g.total_tiles; t += npairs) {
        int pi, pm, tn;
        locate_tile2(g, t, pi, pm, tn);
        const TcProblem& p = g.p[pi];
        const int num_kb = (p.K + TC_BK - 1) / TC_BK;
        const int row0 = pm * TC2_BM + (int)rank * 128;          // this CTA's rows of A / C
        const int ncol0 = tn * BN + (int)rank * (BN / 2);         // this CTA's half of the B tile
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u, g.error_flag);
          const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
          const uint32_t sb = sa + Cfg::A_BYTES;
          const uint32_t lbar = mapa_u32(full_bar(stage), 0);
          if (elect_one()) {
          if (leader) mbar_arrive_expect_tx(full_bar(stage), 2 * Cfg::STAGE_BYTES);
          if (!p.a_mn) {
            tma_load_2d_2sm(&p.map_a, lbar, sa, kb * TC_BK, row0);
          } else {
#pragma unroll
            for (int c = 0; c < 2; ++c) tma_load_2d_2sm(&p.map_a, lbar, sa + c * (TC_BK * 128), row0 + c * 64, kb * TC_BK);
          }
          if (!p.b_mn) {
            tma_load_2d_2sm(&p.map_b, lbar, sb, kb * TC_BK, ncol0);
          } else {
#pragma unroll
            for (int c = 0; c < BN / 128; ++c) tma_load_2d_2sm(&p.map_b, lbar, sb + c * (TC_BK * 128), ncol0 + c * 64, kb * TC_BK);
          }
          }
          __syncwarp();
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (leader) {
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int t = pair; t < g.total_tiles; t += npairs) {
        int pi, pm, tn;
        locate_tile2(g, t, pi, pm, tn);
        const TcProblem& p = g.p[pi];
        const int num_kb = (p.K + TC_BK - 1) / TC_BK;
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (uint32_t(p.a_mn) << 15) | (uint32_t(p.b_mn) << 16) |
                               (uint32_t(BN >> 3) << 17) | (uint32_t(TC2_BM >> 4) << 24);
        const uint32_t a_lbo = p.a_mn ? (uint32_t)g.mn_lbo : 0u, a_sbo = p.a_mn ? (uint32_t)g.mn_sbo : 1024u;
        const uint32_t b_lbo = p.b_mn ? (uint32_t)g.mn_lbo : 0u, b_sbo = p.b_mn ? (uint32_t)g.mn_sbo : 1024u;
        const uint32_t a_kstep = p.a_mn ? 16u * 128u : 32u;
        const uint32_t b_kstep = p.b_mn ? 16u * 128u : 32u;
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u, g.error_flag);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + uint32_t(acc * BN);
        uint32_t ready = mbar_try(full_bar(stage), phase);
        for (int kb = 0; kb < num_kb; ++kb) {
          if (!ready) mbar_wait(full_bar(stage), phase, g.error_flag);
          int nstage = stage + 1; uint32_t nphase = phase;
          if (nstage == Cfg::STAGES) { nstage = 0; nphase ^= 1u; }
          ready = (kb + 1 < num_kb) ? mbar_try(full_bar(nstage), nphase) : 0u;
          tc_fence_after();
          const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
          const uint32_t sb = sa + Cfg::A_BYTES;
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < TC_BK / 16; ++k) {
              const uint64_t adesc = make_smem_desc(sa + k * a_kstep, a_lbo, a_sbo);
              const uint64_t bdesc = make_smem_desc(sb + k * b_kstep, b_lbo, b_sbo);
              umma_bf16_2sm(d_tmem, adesc, bdesc, idesc, (kb | k) != 0 ? 1u : 0u);
            }
            umma_commit_2sm(empty_bar(stage));
            if (kb + 1 == num_kb) umma_commit_2sm(tfull_bar(acc));
          }
          __syncwarp();
          stage = nstage; phase = nphase;
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
    }
  } else {
    // ===================== epilogue (warps 2..5 of both CTAs; each CTA drains its own 128 accumulator rows) =====================
    const int quarter = warp & 3;
    int acc = 0;
    uint32_t acc_phase = 0;
    const uint32_t leader_tempty0 = mapa_u32(tempty_bar(0), 0), leader_tempty1 = mapa_u32(tempty_bar(1), 0);
    for (int t = pair; t < g.total_tiles; t += npairs) {
      int pi, pm, tn;
      locate_tile2(g, t, pi, pm, tn);
      const TcProblem& p = g.p[pi];
      const Epi& e = p.epi;
      float alpha = e.alpha * (e.alpha_ptr ? *e.alpha_ptr : 1.f);
      const float beta = e.beta * (e.beta_ptr ? *e.beta_ptr : 1.f);
      float beta2 = e.beta2;
      if (e.pro_fs) { const float a = epi_procrustes_step(e); alpha *= 0.5f * a * a; beta2 = a; }
      mbar_wait_relaxed(tfull_bar(acc), acc_phase, g.error_flag);
      tc_fence_after();
      const int tm = pm * 2 + (int)rank;                 // this CTA's 128-row block
      const int row = tm * TC_BM + quarter * 32 + lane;
      RowAcc ra = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        const int col0 = tn * BN + c * 32;
        int mirror = 0;
        if (p.sym) {
          const int cb = col0 >> 7;
          if (cb < tm) continue;
          mirror = cb > tm;
        }
        uint32_t raw[32];
        const uint32_t taddr = tmem_base + (uint32_t(quarter * 32) << 16) + uint32_t(acc * BN + c * 32);
        tmem_ld_32x32(taddr, raw);
        tmem_ld_wait();
        if (col0 < p.N) epilogue_chunk(e, p.M, p.N, row, col0, raw, alpha, beta, beta2, lane, ra, mirror);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(acc == 0 ? leader_tempty0 : leader_tempty1);
      if (e.row_sumsq && row < p.M) atomicAdd(&e.row_sumsq[row], ra.row_sumsq);
      if (e.total_sumsq) { float s = warp_sum(ra.tot); if (lane == 0) atomicAdd(e.total_sumsq, s); }
      if (e.dot_out) { float s = warp_sum(ra.dot); if (lane == 0 && s != 0.f) atomicAdd(e.dot_out, s); }
      if (e.abs_max) { float s = warp_max(ra.amax); if (lane == 0) atomic_max_nonneg(e.abs_max, s); }
      if (e.trace) { float s = warp_sum(ra.tr); if (lane == 0 && s != 0.f) atomicAdd(e.trace, s); }
      if (e.diag_max) { float s = warp_max(ra.dmax); if (lane == 0) atomic_max_nonneg(e.diag_max, s); }
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
  }

  tc_fence_before();
  cluster_sync_all();    // neither CTA may free TMEM / exit while the peer's MMAs, multicast commits or remote arrives are in flight
  if (warp == 1) { tc_fence_after(); tmem_dealloc_2sm(tmem_base, Cfg::TMEM_COLS); }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// row-major bf16 matrix (rows x cols, leading dimension ld elements), box = {64 cols, box_rows}
int make_tmap(Ctx* ctx, CUtensorMap* out, const void* ptr, int rows, int cols, int ld, int box_rows) {
  EncodeTiledFn fn = reinterpret_cast<EncodeTiledFn>(ctx->encode_tiled);
  if (!fn) return PSGD_ERR_CUDA;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    snprintf(ctx->last_error, sizeof(ctx->last_error), "cuTensorMapEncodeTiled failed (%d) rows=%d cols=%d ld=%d", (int)r, rows, cols, ld);
    return PSGD_ERR_CUDA;
  }
  return PSGD_OK;
}

// MN-major operand (rows = K, cols = MN, cols % 64 == 0) as a 3-D tensor {64 elems, rows, cols/64}: box {64, 64, chunks} lands in smem as
// [chunk][k row][64 elems] -- the same layout the per-chunk 2-D boxes produce -- with a single TMA operation
int make_tmap_mn3(Ctx* ctx, CUtensorMap* out, const void* ptr, int rows, int cols, int ld, int chunks, int k_rows) {
  EncodeTiledFn fn = reinterpret_cast<EncodeTiledFn>(ctx->encode_tiled);
  if (!fn) return PSGD_ERR_CUDA;
  cuuint64_t dims[3] = {64, (cuuint64_t)rows, (cuuint64_t)(cols / 64)};
  cuuint64_t strides[2] = {(cuuint64_t)ld * 2, 128};
  cuuint32_t box[3] = {64, (cuuint32_t)(k_rows > 0 ? k_rows : TC_BK), (cuuint32_t)chunks};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    snprintf(ctx->last_error, sizeof(ctx->last_error), "cuTensorMapEncodeTiled(3d) failed (%d) rows=%d cols=%d ld=%d", (int)r, rows, cols, ld);
    return PSGD_ERR_CUDA;
  }
  return PSGD_OK;
}

bool tc_eligible(const GemmDesc& g) {
  if (g.in_dtype != PSGD_BF16) return false;
  if (g.M < 128 || g.N < 8 || g.K < 64) return false;  // narrow N (the 32-probe norm-bound products) rides on TMA zero fill
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
  if (!al16(g.A) || !al16(g.B) || !al16(g.epi.C)) return false;
  if (g.lda % 8 || g.ldb % 8) return false;
  if (g.N % 8) return false;
  const int ovec = g.epi.out_dtype == PSGD_BF16 ? 8 : 4;
  if (g.epi.ldc % ovec) return false;
  const int dvec = g.epi.d_dtype == PSGD_BF16 ? 8 : 4;
  if (g.epi.D && (!al16(g.epi.D) || g.epi.ldd % dvec)) return false;
  if (g.epi.D2 && (!al16(g.epi.D2) || g.epi.ldd2 % dvec)) return false;
  if (g.epi.dotm && (!al16(g.epi.dotm) || g.epi.ld_dot % dvec)) return false;
  return true;
}

template <int BN>
static int launch_tc(Ctx* ctx, TcGroup& grp, const GemmDesc* gs, int n, cudaStream_t st) {
  using Cfg = TcCfg<BN>;
  int ntiles[TC_MAX_PROBLEMS];
  int T = 0;
  for (int i = 0; i < n; ++i) {
    const GemmDesc& g = gs[i];
    TcProblem& p = grp.p[i];
    p.epi = g.epi;
    p.M = g.M; p.N = g.N; p.K = g.K;
    p.a_mn = g.ta ? 1 : 0;        // A stored K x M: M runs along the contiguous dimension
    p.b_mn = g.tb ? 0 : 1;        // B stored K x N: N runs along the contiguous dimension
    int rc;
    p.a_3d = (g.ta && g.M % 64 == 0 && !(ctx->debug_flags & 32)) ? 1 : 0;
    p.b_3d = (!g.tb && g.N % 64 == 0 && !(ctx->debug_flags & 32)) ? 1 : 0;
    if (!g.ta) rc = make_tmap(ctx, &p.map_a, g.A, g.M, g.K, g.lda, TC_BM);
    else if (p.a_3d) rc = make_tmap_mn3(ctx, &p.map_a, g.A, g.K, g.M, g.lda, TC_BM / 64);
    else rc = make_tmap(ctx, &p.map_a, g.A, g.K, g.M, g.lda, TC_BK);
    if (rc) return rc;
    if (g.tb) rc = make_tmap(ctx, &p.map_b, g.B, g.N, g.K, g.ldb, BN);
    else if (p.b_3d) rc = make_tmap_mn3(ctx, &p.map_b, g.B, g.K, g.N, g.ldb, BN / 64);
    else rc = make_tmap(ctx, &p.map_b, g.B, g.K, g.N, g.ldb, TC_BK);
    if (rc) return rc;
    p.tiles_m = (g.M + TC_BM - 1) / TC_BM;
    p.tiles_n = (g.N + BN - 1) / BN;
    p.sym = (!(ctx->debug_flags & 1) && g.sym && g.M == g.N && !g.epi.D && !g.epi.D2 && !g.epi.dotm && !g.epi.row_scale && !g.epi.col_scale && !g.epi.col_sumsq && !g.epi.norm_axis) ? 1 : 0;
    if (p.sym) {
      ntiles[i] = 0;
      for (int tm = 0; tm < p.tiles_m; ++tm) ntiles[i] += p.tiles_n - (tm * TC_BM) / BN;
    } else {
      ntiles[i] = p.tiles_m * p.tiles_n;
    }
    p.tile_first = 0; p.splits = 1; p.kb_split = (g.K + TC_BK - 1) / TC_BK; p.ws_slot0 = 0; p.nsplit = 1; p.bn_eff = BN;
    if (BN == 256 && g.tb) { rc = make_tmap(ctx, &p.map_b_half, g.B, g.N, g.K, g.ldb, 128); if (rc) return rc; }
    if (BN == 256 && p.b_3d) { rc = make_tmap_mn3(ctx, &p.map_b_half, g.B, g.K, g.N, g.ldb, 2); if (rc) return rc; }
    T += ntiles[i];
  }
  // ---- filling the machine ----
  //  * under-filled launch (T*2 <= SMs): first halve the tile width (128 x 128 tiles, twice as many, no reduction needed), then, if still
  //    under-filled (the 32-probe norm-bound products, small Grams), split K: partial accumulators meet in the fp32 workspace and the
  //    unit that arrives last runs the epilogue
  //  * otherwise, if the last wave would use at most half of the SMs, its tiles are issued as half-width tiles (2 units each)
  const int sms = ctx->num_sms;
  int nent = n;
  int slots = 0;
  const bool can_split = !(ctx->debug_flags & 16);
  if (can_split && T * 2 <= sms) {
    int T2 = 0;
    for (int i = 0; i < n; ++i) {
      TcProblem& p = grp.p[i];
      if (BN == 256 && p.N > 128) { p.nsplit = 2; p.bn_eff = 128; }
      T2 += ntiles[i] * p.nsplit;
    }
    if (T2 * 2 <= sms && ctx->ws) {
      for (int i = 0; i < n; ++i) {
        TcProblem& p = grp.p[i];
        const int num_kb = (p.K + TC_BK - 1) / TC_BK;
        int sp = sms / T2;
        if (sp > 8) sp = 8;
        if (sp > num_kb / 4) sp = num_kb / 4;
        if (sp < 2) continue;
        const int kbs = (num_kb + sp - 1) / sp;
        const int spl = (num_kb + kbs - 1) / kbs;
        if (spl < 2 || slots + ntiles[i] * p.nsplit * spl > ctx->ws_slots) continue;
        p.kb_split = kbs;
        p.splits = spl;
        p.ws_slot0 = slots;
        slots += ntiles[i] * p.nsplit * spl;
      }
    }
  } else if (can_split && BN == 256 && T > sms && n < TC_MAX_PROBLEMS) {
    const int frac = T % sms;
    if (frac > 0 && frac * 2 <= sms && frac <= ntiles[n - 1] && grp.p[n - 1].N > 128) {
      grp.p[n] = grp.p[n - 1];
      TcProblem& tail = grp.p[n];
      tail.tile_first = ntiles[n - 1] - frac;
      tail.nsplit = 2;
      tail.bn_eff = 128;
      ntiles[n] = frac;
      ntiles[n - 1] -= frac;
      nent = n + 1;
    }
  }
  int units = 0;
  for (int i = 0; i < nent; ++i) { grp.p[i].tile_start = units; units += ntiles[i] * grp.p[i].nsplit * grp.p[i].splits; }
  grp.num_problems = nent;
  grp.total_tiles = units;
  grp.mn_lbo = ctx->mn_lbo;
  grp.mn_sbo = ctx->mn_sbo;
  grp.error_flag = nullptr;
  grp.ws = ctx->ws;
  grp.ws_count = ctx->ws_count;
  static PerDeviceOnce attr_set;   // one per instantiation
  if (attr_set.need(ctx->device)) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return check_cuda(ctx, e, "cudaFuncSetAttribute(gemm_tc)");
  }
  int grid = units < sms ? units : sms;
  const bool timed = ctx->timing_on && ctx->timing_count < ctx->ev_capacity;
  if (ctx->timing_on) ctx->timing_seen++;
  if (timed) cudaEventRecord(ctx->ev_begin[ctx->timing_count], st);
  gemm_tc_kernel<BN><<<grid, TC_THREADS, Cfg::SMEM_BYTES, st>>>(grp);
  if (timed) {
    cudaEventRecord(ctx->ev_end[ctx->timing_count], st);
    ctx->timing_count++;
    for (int i = 0; i < n; ++i) ctx->timing_flops += 2.0 * gs[i].M * (double)gs[i].N * gs[i].K;
    for (int i = 0; i < nent; ++i) ctx->timing_flops_exec += 2.0 * ntiles[i] * (double)TC_BM * BN * grp.p[i].K;
  }
  ctx->launches++;
  return check_cuda(ctx, cudaGetLastError(), "gemm_tc");
}

// 2-CTA launch (256 x 256 pair tiles, B split between the two SMs: 64 B/clk/SM of operand traffic instead of 96).  Measured on B200 the
// same work costs a 1-CTA wave ~1.2x a pair wave, but the 1-CTA kernel has finer tiles (half-width tail wave, split-K), so the choice is
// made on modelled waves: pair waves are whole, 1-CTA waves round up to the next half.
// Returns 0 (use the 1-CTA kernel), 256 or 128 (pair-tile width of the 2-CTA kernel).  128-wide pair tiles halve the wave granularity (a
// 4096^3 product is 256 tiles of 256 x 256 = 3.46 waves of 74 pairs -> 4 waves, but 512 tiles of 256 x 128 = 6.92 -> 7 half-size waves), yet
// measured on B200 a 256 x 128 wave costs 0.70 of a 256 x 256 wave, not 0.5 (4096^3: 122 us against 100 us, profiles/r01_gemm_tile_n128.log):
// the relative cost PSGD_B200_TC2_N128_PCT defaults to 140, which keeps the narrow tile for launches that would otherwise waste most
// of their last wave only.
static int tc2_choice(const Ctx* ctx, const GemmDesc* gs, int n) {
  if (ctx->debug_flags & 8) return 0;
  const int pairs = ctx->num_sms / 2;
  long t2 = 0, t2n = 0, t1 = 0;
  bool any_sym = false;
  for (int i = 0; i < n; ++i) {
    const GemmDesc& g = gs[i];
    if (g.M < 256 || g.N <= 128) return 0;
    const bool sym = (!(ctx->debug_flags & 1) && g.sym && g.M == g.N && !g.epi.D && !g.epi.D2 && !g.epi.dotm && !g.epi.row_scale && !g.epi.col_scale && !g.epi.col_sumsq && !g.epi.norm_axis);
    const long pm = (g.M + TC2_BM - 1) / TC2_BM, pn = (g.N + TC2_BN - 1) / TC2_BN;
    const long tm = (g.M + TC_BM - 1) / TC_BM;
    if (sym) {
      any_sym = true;
      for (long r = 0; r < pm; ++r) t2 += pn - r;
      for (long r = 0; r < tm; ++r) t1 += pn - (r * TC_BM) / 256;
    } else {
      t2 += pm * pn;
      t2n += pm * ((g.N + 127) / 128);
      t1 += tm * pn;
    }
  }
  if (t2 < pairs) return 0;                     // under-filled: the 1-CTA kernel's narrow tiles / split-K fill the machine better
  const double w2 = (double)((t2 + pairs - 1) / pairs);
  const long full = t1 / ctx->num_sms, frac = t1 % ctx->num_sms;
  const double w1 = (double)full + (frac == 0 ? 0.0 : (frac * 2 <= ctx->num_sms ? 0.5 : 1.0));
  static int wave_pct = -1;                            // experiment knob: relative cost of a 1-CTA wave in percent (default 118)
  if (wave_pct < 0) { const char* e = getenv("PSGD_B200_TC1_WAVE_PCT"); wave_pct = e ? atoi(e) : 0; if (wave_pct < 0 || wave_pct > 1000) wave_pct = 0; }
  const int pct = wave_pct;
  const double c1 = w1 * (pct ? pct * 0.01 : 1.18);
  double c2 = w2;
  int bn = 256;
  if (!any_sym && !(ctx->debug_flags & 128)) {
    static int n128_pct = -1;
    if (n128_pct < 0) { const char* e = getenv("PSGD_B200_TC2_N128_PCT"); n128_pct = e ? atoi(e) : 140; if (n128_pct < 50) n128_pct = 140; }
    const double c2n = (double)((t2n + pairs - 1) / pairs) * 0.5 * (n128_pct * 0.01);
    if (c2n < c2 || (ctx->debug_flags & 65536)) { c2 = c2n; bn = 128; }
  }
  return c2 <= c1 ? bn : 0;
}

template <int BN>
static int launch_tc2(Ctx* ctx, TcGroup& grp, const GemmDesc* gs, int n, cudaStream_t st) {
  using Cfg = Tc2Cfg<BN>;
  int tiles = 0;
  double exec_flops = 0.0;
  for (int i = 0; i < n; ++i) {
    const GemmDesc& g = gs[i];
    const int tiles_before = tiles;
    TcProblem& p = grp.p[i];
    p.epi = g.epi;
    p.M = g.M; p.N = g.N; p.K = g.K;
    p.a_mn = g.ta ? 1 : 0;
    p.b_mn = g.tb ? 0 : 1;
    p.a_3d = 0; p.b_3d = 0;
    int rc;
    if (!g.ta) rc = make_tmap(ctx, &p.map_a, g.A, g.M, g.K, g.lda, 128);
    else rc = make_tmap(ctx, &p.map_a, g.A, g.K, g.M, g.lda, TC_BK);
    if (rc) return rc;
    if (g.tb) rc = make_tmap(ctx, &p.map_b, g.B, g.N, g.K, g.ldb, Cfg::BN / 2);
    else rc = make_tmap(ctx, &p.map_b, g.B, g.K, g.N, g.ldb, TC_BK);
    if (rc) return rc;
    p.tiles_m = (g.M + TC2_BM - 1) / TC2_BM;
    p.tiles_n = (g.N + Cfg::BN - 1) / Cfg::BN;
    p.tile_start = tiles; p.tile_first = 0; p.splits = 1; p.kb_split = 0; p.ws_slot0 = 0; p.nsplit = 1; p.bn_eff = Cfg::BN;
    p.sym = (!(ctx->debug_flags & 1) && g.sym && g.M == g.N && !g.epi.D && !g.epi.D2 && !g.epi.dotm && !g.epi.row_scale && !g.epi.col_scale && !g.epi.col_sumsq && !g.epi.norm_axis) ? 1 : 0;
    if (p.sym) {
      for (int pm = 0; pm < p.tiles_m; ++pm) tiles += p.tiles_n - pm;
    } else {
      tiles += p.tiles_m * p.tiles_n;
    }
    exec_flops += 2.0 * (tiles - tiles_before) * (double)TC2_BM * Cfg::BN * g.K;
  }
  grp.num_problems = n;
  grp.total_tiles = tiles;
  grp.mn_lbo = ctx->mn_lbo;
  grp.mn_sbo = ctx->mn_sbo;
  grp.error_flag = nullptr;
  static PerDeviceOnce attr_set;   // one per instantiation
  if (attr_set.need(ctx->device)) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc2_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return check_cuda(ctx, e, "cudaFuncSetAttribute(gemm_tc2)");
  }
  const int max_pairs = ctx->num_sms / 2;
  const int pairs = tiles < max_pairs ? tiles : max_pairs;
  const bool timed = ctx->timing_on && ctx->timing_count < ctx->ev_capacity;
  if (ctx->timing_on) ctx->timing_seen++;
  if (timed) cudaEventRecord(ctx->ev_begin[ctx->timing_count], st);
  gemm_tc2_kernel<BN><<<2 * pairs, TC2_THREADS, Cfg::SMEM_BYTES, st>>>(grp);
  if (timed) {
    cudaEventRecord(ctx->ev_end[ctx->timing_count], st);
    ctx->timing_count++;
    for (int i = 0; i < n; ++i) ctx->timing_flops += 2.0 * gs[i].M * (double)gs[i].N * gs[i].K;
    ctx->timing_flops_exec += exec_flops;
  }
  ctx->launches++;
  return check_cuda(ctx, cudaGetLastError(), "gemm_tc2");
}

int launch_gemm_tc_group(Ctx* ctx, const GemmDesc* gs, int n, cudaStream_t st) {
  if (n < 1 || n > TC_MAX_PROBLEMS) return PSGD_ERR_INVALID_ARG;
  int max_n = 0;
  for (int i = 0; i < n; ++i) {
    if (!tc_eligible(gs[i])) return PSGD_ERR_UNSUPPORTED;
    if (gs[i].N > max_n) max_n = gs[i].N;
  }
  TcGroup grp;
  memset(&grp, 0, sizeof(grp));
  if (ctx->force_bn == 0) {
    const int bn2 = tc2_choice(ctx, gs, n);
    if (bn2 == 256) return launch_tc2<256>(ctx, grp, gs, n, st);
    if (bn2 == 128) return launch_tc2<128>(ctx, grp, gs, n, st);
  }
  if (ctx->force_bn == 128) return launch_tc<128>(ctx, grp, gs, n, st);
  if (ctx->force_bn == 256) return launch_tc<256>(ctx, grp, gs, n, st);
  if (max_n > 128) return launch_tc<256>(ctx, grp, gs, n, st);
  return launch_tc<128>(ctx, grp, gs, n, st);
}

}  // namespace psgd
