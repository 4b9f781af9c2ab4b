// lra_tc.cuh -- sweep 1 of the LRA update (psgd.py:1006 and the r-sized right-hand sides of 1017-1052) on the 5th-generation tensor cores.
//
// The mma.sync sweeps of lra_mma.cuh are bound by the issue rate of HMMA.16816 on this part (one per ~25-30 cycles per SM sub-partition:
// 24 of them per 16-row chunk = 4.4 TB/s at rank 32, and rank 64 does not fit a warp's registers at all).  tcgen05 has two orders of
// magnitude more throughput, but wants 128-byte operand rows.  Trick: view U (n x r, r = 16 / 32 / 64) as Upk (n / PACK x 64) with
// PACK = 64 / r consecutive rows per 128-byte line.  Then with W = [Upk | Vpk] (n / PACK x 128, two MN-major 64-column chunks: exactly what
// two 2-D TMA boxes deliver),
//     D = W^T W   (128 x 128, one UMMA chain with M = N = 128, fp32 in TMEM)
// holds U^T U, V^T V and V^T U as the sums of the PACK diagonal r x r sub-blocks of its 64 x 64 blocks (the off-diagonal sub-blocks pair
// different rows of a line and are ignored: the price of the packing, paid in tensor-core time that is not the bound).
// The projections U^T x, V^T x (x1 = d.h, x2 = v / d, psgd.py:1017 / 1022) ride on a second, 16-column UMMA per k step: X is a K-major
// operand tile that four builder warps form in shared memory from the d / h / v slices of the stage (row c: x1 of sub-row c, row PACK + c: x2).
//   warp 0: producer (TMA 2-D boxes of Upk, Vpk + bulk copies of d, h, v into a 5-stage mbarrier ring), warp 1: tcgen05.mma issuer,
//   warps 2-5: X builders during the sweep, readers of TMEM at the end (atomics into the same accumulator block as k_lra_gram_mma).
// Handles the whole 128 * PACK-row blocks of the vector; the caller sends the remainder through the mma.sync kernel.
#pragma once
#include "common.cuh"
#include "lra_mma.cuh"
#include "tc_ptx.cuh"

namespace psgd {

constexpr int LT_KR = 128;                 // packed rows (128-byte lines) per stage and factor
constexpr int LT_STAGES = 5;
constexpr int LT_THREADS = 192;
constexpr int LT_TILE = LT_KR * 128;       // 16 KB: one factor's tile
constexpr int LT_X_BYTES = 2 * 16 * 128;   // X tile: two 64-k chunks of [16 rows][128 B]

template <int RP> struct LtCfg {
  static constexpr int PACK = 64 / RP;
  static constexpr int ROWS = LT_KR * PACK;            // original rows per stage
  static constexpr int VEC_BYTES = ROWS * 2;           // one vector slice (bf16)
  static constexpr int STAGE_BYTES = 2 * LT_TILE + 3 * VEC_BYTES;
  static constexpr int SMEM_BYTES = LT_STAGES * (STAGE_BYTES + LT_X_BYTES) + 1024 + 256;
};

struct alignas(64) LtParams {
  CUtensorMap map_u;     // Upk: 2-D {64, n / PACK}, box {64, LT_KR}, 128-byte swizzle
  CUtensorMap map_v;
  const bf16* d; const bf16* h; const bf16* v;
  long long nblocks;     // whole stages
  float* acc_out;        // [UU][VV][VU][Utx1][Vtx1][Utx2][Vtx2][|x1|^2, |x2|^2] (lra_acc_floats)
};

__device__ __forceinline__ void lt_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ float lt_rbf(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }
__device__ __forceinline__ void lt_fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

template <int RP>
__global__ void __launch_bounds__(LT_THREADS, 1) k_lra_gram_tc(const __grid_constant__ LtParams P) {
  using Cfg = LtCfg<RP>;
  constexpr int PACK = Cfg::PACK;
  extern __shared__ uint8_t lt_smem_raw[];
  const uint32_t smem_base = (smem_u32(lt_smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = lt_smem_raw + (smem_base - smem_u32(lt_smem_raw));
  // layout: [stage tiles: U 16 KB | V 16 KB] x S (1024-aligned), then [X tiles 4 KB] x S, then [vectors d | h | v] x S, then barriers
  const uint32_t x_base = smem_base + LT_STAGES * 2 * LT_TILE;
  const uint32_t vec_base = x_base + LT_STAGES * LT_X_BYTES;
  const uint32_t bar_base = vec_base + LT_STAGES * 3 * Cfg::VEC_BYTES;
  uint8_t* x_gen = smem_gen + LT_STAGES * 2 * LT_TILE;
  uint8_t* vec_gen = x_gen + LT_STAGES * LT_X_BYTES;
  auto full_bar = [&](int s_) { return bar_base + 8u * s_; };
  auto xfull_bar = [&](int s_) { return bar_base + 8u * (LT_STAGES + s_); };
  auto empty_bar = [&](int s_) { return bar_base + 8u * (2 * LT_STAGES + s_); };
  const uint32_t done_bar = bar_base + 8u * (3 * LT_STAGES);
  const uint32_t tmem_slot = bar_base + 8u * (3 * LT_STAGES + 1);
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(vec_gen + LT_STAGES * 3 * Cfg::VEC_BYTES + 8 * (3 * LT_STAGES + 1));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int i = 0; i < LT_STAGES; ++i) { mbar_init(full_bar(i), 1); mbar_init(xfull_bar(i), 4); mbar_init(empty_bar(i), 1); }
    mbar_init(done_bar, 1);
    fence_barrier_init();
    prefetch_tmap(&P.map_u); prefetch_tmap(&P.map_v);
  }
  // X tiles: rows that no builder writes must read as zero
  for (int i = tid; i < LT_STAGES * LT_X_BYTES / 16; i += LT_THREADS) reinterpret_cast<uint4*>(x_gen)[i] = make_uint4(0u, 0u, 0u, 0u);
  if (warp == 1) tmem_alloc(tmem_slot, 256);
  lt_fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;
  const long long nblk = P.nblocks;

  if (warp == 0) {
    // ===================== producer =====================
    int stage = 0; uint32_t phase = 0;
    for (long long blk = blockIdx.x; blk < nblk; blk += gridDim.x) {
      mbar_wait(empty_bar(stage), phase ^ 1u, nullptr);
      if (elect_one()) {
        const uint32_t su = smem_base + stage * 2 * LT_TILE, fb = full_bar(stage);
        const uint32_t sv = vec_base + stage * 3 * Cfg::VEC_BYTES;
        const long long r0 = blk * Cfg::ROWS;
        mbar_arrive_expect_tx(fb, 2 * LT_TILE + 3 * Cfg::VEC_BYTES);
        tma_load_2d(&P.map_u, fb, su, 0, (int)(blk * LT_KR));
        tma_load_2d(&P.map_v, fb, su + LT_TILE, 0, (int)(blk * LT_KR));
        lt_bulk_g2s(sv, P.d + r0, Cfg::VEC_BYTES, fb);
        lt_bulk_g2s(sv + Cfg::VEC_BYTES, P.h + r0, Cfg::VEC_BYTES, fb);
        lt_bulk_g2s(sv + 2 * Cfg::VEC_BYTES, P.v + r0, Cfg::VEC_BYTES, fb);
      }
      __syncwarp();
      if (++stage == LT_STAGES) { stage = 0; phase ^= 1u; }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // D (TMEM columns 0..127) += W^T W: both operands MN-major, M = N = 128;  D2 (columns 128..143) += W^T X: X K-major, N = 16
    const uint32_t idesc_g = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | (uint32_t(128 >> 3) << 17) | (uint32_t(128 >> 4) << 24);
    const uint32_t idesc_x = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (uint32_t(16 >> 3) << 17) | (uint32_t(128 >> 4) << 24);
    const uint64_t wdesc0 = make_smem_desc(smem_base, (uint32_t)LT_TILE, 1024u);     // chunk stride = U tile -> V tile
    const uint64_t xdesc0 = make_smem_desc(x_base, 0u, 1024u);
    int stage = 0; uint32_t phase = 0;
    bool first = true;
    for (long long blk = blockIdx.x; blk < nblk; blk += gridDim.x) {
      mbar_wait(full_bar(stage), phase, nullptr);
      mbar_wait(xfull_bar(stage), phase, nullptr);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t wo = (uint64_t)((stage * 2 * LT_TILE) >> 4), xo = (uint64_t)((stage * LT_X_BYTES) >> 4);
#pragma unroll
        for (int k = 0; k < LT_KR / 16; ++k) {
          const uint64_t wd = wdesc0 + wo + (uint64_t)((k * 16 * 128) >> 4);
          const uint64_t xd = xdesc0 + xo + (uint64_t)(((k >> 2) * 2048 + (k & 3) * 32) >> 4);
          const uint32_t accf = (first && k == 0) ? 0u : 1u;
          umma_bf16(tmem_base, wd, wd, idesc_g, accf);
          umma_bf16(tmem_base + 128u, wd, xd, idesc_x, accf);
        }
        umma_commit(empty_bar(stage));
      }
      __syncwarp();
      first = false;
      if (++stage == LT_STAGES) { stage = 0; phase ^= 1u; }
    }
    if (elect_one()) umma_commit(done_bar);
    __syncwarp();
  } else {
    // ===================== X builders (warps 2..5 = 128 threads: thread t owns line t of the stage) =====================
    const int t = tid - 64;
    float sq1 = 0.f, sq2 = 0.f;
    int stage = 0; uint32_t phase = 0;
    for (long long blk = blockIdx.x; blk < nblk; blk += gridDim.x) {
      mbar_wait(full_bar(stage), phase, nullptr);
      const bf16* dv = reinterpret_cast<const bf16*>(vec_gen + stage * 3 * Cfg::VEC_BYTES);
      const bf16* hv = dv + Cfg::ROWS;
      const bf16* vv = hv + Cfg::ROWS;
      uint8_t* xt = x_gen + stage * LT_X_BYTES + (t >> 6) * 2048;      // 64-k chunk of this line
      const int kin = t & 63;
#pragma unroll
      for (int c = 0; c < PACK; ++c) {
        const float dd = __bfloat162float(dv[PACK * t + c]);
        const float x1 = lt_rbf(dd * __bfloat162float(hv[PACK * t + c]));   // d.h   psgd.py:1017
        const float x2 = lt_rbf(__bfloat162float(vv[PACK * t + c]) / dd);   // v/d   psgd.py:1022
        sq1 = fmaf(x1, x1, sq1); sq2 = fmaf(x2, x2, sq2);
        // K-major tile [16 rows][128 B], 128-byte swizzle: element (row j, k) at j * 128 + (((k >> 3) ^ (j & 7)) << 4) + (k & 7) * 2
        const int j1 = c, j2 = PACK + c;
        *reinterpret_cast<bf16*>(xt + j1 * 128 + (((kin >> 3) ^ (j1 & 7)) << 4) + (kin & 7) * 2) = __float2bfloat16_rn(x1);
        *reinterpret_cast<bf16*>(xt + j2 * 128 + (((kin >> 3) ^ (j2 & 7)) << 4) + (kin & 7) * 2) = __float2bfloat16_rn(x2);
      }
      lt_fence_async_smem();       // generic-proxy stores -> tensor-core (async proxy) reads
      __syncwarp();
      if (lane == 0) mbar_arrive(xfull_bar(stage));
      if (++stage == LT_STAGES) { stage = 0; phase ^= 1u; }
    }
    // ---- read-out: warp w owns TMEM lanes 32 (w % 4) .. + 31 = rows m of D ----
    mbar_wait_relaxed(done_bar, 0u, nullptr);
    tc_fence_after();
    const int quarter = warp & 3;
    const int m = quarter * 32 + lane;
    const bool is_v = m >= 64;
    const int c = (m & 63) / RP, a = (m & 63) % RP;
    float* UU = P.acc_out; float* VV = UU + RP * RP; float* VU = VV + RP * RP; float* PR = VU + RP * RP;
    if (nblk > (long long)blockIdx.x) {
#pragma unroll 1
      for (int cb = 0; cb < 4; ++cb) {          // 32-column chunks of D: columns 32 cb .. + 31
        uint32_t raw[32];
        tmem_ld_32x32(tmem_base + (uint32_t(quarter * 32) << 16) + uint32_t(cb * 32), raw);
        tmem_ld_wait();
        // wanted: columns [c RP, c RP + RP) of the U block (cb < 2) for U^T U (U rows) / V^T U (V rows); of the V block (cb >= 2) for V^T V (V rows)
#pragma unroll
        for (int jx = 0; jx < 32; ++jx) {
          const int col = cb * 32 + jx;
          const int cc = (col & 63) / RP, b = (col & 63) % RP;
          if (cc != c) continue;
          const float val = __uint_as_float(raw[jx]);
          if (col < 64) atomicAdd((is_v ? VU : UU) + a * RP + b, val);
          else if (is_v) atomicAdd(VV + a * RP + b, val);
        }
      }
      {
        uint32_t raw[32];
        tmem_ld_32x32(tmem_base + (uint32_t(quarter * 32) << 16) + 128u, raw);
        tmem_ld_wait();
        // [Utx1][Vtx1][Utx2][Vtx2]: x1 of sub-row c is column c, x2 is column PACK + c
        atomicAdd(PR + (is_v ? RP : 0) + a, __uint_as_float(raw[c]));
        atomicAdd(PR + 2 * RP + (is_v ? RP : 0) + a, __uint_as_float(raw[PACK + c]));
      }
    }
    sq1 = warp_sum(sq1); sq2 = warp_sum(sq2);
    if (lane == 0) { atomicAdd(PR + 4 * RP, sq1); atomicAdd(PR + 4 * RP + 1, sq2); }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 256); }
}

}  // namespace psgd
