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
__device__ __forceinline__ void lt_tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void lt_bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void lt_bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
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

namespace psgd {

// =====================================================================================================================================
// sweep 2 of the LRA update on tcgen05: balancing rotation U' = (U - U Eu) / rho, V' = (V + V Ev) rho (psgd.py:1012-1015), the per-row terms
// of the d update (psgd.py:1017-1029) and the rank-2 update of U or V (psgd.py:1036-1052), written back in place.
//   packed lines as in sweep 1; per tile of 128 lines:  D_U[128 x 80] = Upk_tile (K-major, K = 64) x B_U,  D_V[128 x 80 / 96] = Vpk_tile x B_V
//   B_U = [blockdiag(Eu, .., Eu) | hi / lo bf16 columns of the vectors whose row dot products are needed (Au c1, Au s2), one set per
//   sub-row] -- K-major operands built once per CTA from the parameter block of k_lra_small.  The identity part of the rotation is
//   applied exactly in fp32 by the epilogue (thread = line: original line from the shared-memory tile, correction and dots from TMEM).
//   TMEM holds two accumulator sets so that the products of tile t + 1 overlap the epilogue of tile t.
//   warp 0: producer, warp 1: MMA issuer, warps 2-9 / 10-17: two epilogue groups that take alternate tiles (group g owns accumulator set
//   g); inside a group four warps (one per TMEM lane quarter) do the U side and four the V side.  The epilogue is ~6000 warp
//   instructions per tile (ncu: issue 29 % active with one group, 5300 cycles per tile against 2800 at the HBM rate), so it needs
//   warps, not bandwidth.  The new lines are written back INTO the shared-memory tile (each thread owns its line) and leave through one
//   TMA store per factor and tile.
// =====================================================================================================================================
// packed fp32x2 arithmetic (FFMA2 / FMUL2 on sm_100a): the rotate epilogue is instruction-bound, these halve its FMA count
typedef unsigned long long lt_f2;
__device__ __forceinline__ lt_f2 lt_pack2(float a, float b) { lt_f2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float2 lt_unpack2(lt_f2 v) { float2 r; asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v)); return r; }
__device__ __forceinline__ lt_f2 lt_fma2(lt_f2 a, lt_f2 b, lt_f2 c) { lt_f2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ lt_f2 lt_mul2(lt_f2 a, lt_f2 b) { lt_f2 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }

constexpr int LR_THREADS = 576;      // warp 0 producer, warp 1 MMA, then two epilogue groups of 8 warps (4 for the U side, 4 for the V side)

template <int RP> struct LrCfg {
  static constexpr int PACK = 64 / RP;
  static constexpr int ROWS = LT_KR * PACK;
  static constexpr int VEC_BYTES = ROWS * 2;
  static constexpr int STAGES = RP == 16 ? 5 : 6;         // a stage lives for load latency + up to two tile times of epilogue + the store drain
  static constexpr int NU = 80;                           // 64 rotation columns + 4 PACK dot columns, padded to a multiple of 16
  static constexpr int NV = PACK == 4 ? 96 : 80;          // 64 + 8 PACK
  static constexpr int BU_BYTES = NU * 128, BV_BYTES = NV * 128;
  static constexpr int SMEM_BYTES = STAGES * (2 * LT_TILE + 3 * VEC_BYTES) + BU_BYTES + BV_BYTES + 1024 + 512;
};

struct alignas(64) LrParams {
  CUtensorMap map_u;
  CUtensorMap map_v;
  bf16* U; bf16* V;
  const bf16* d; const bf16* h; const bf16* v;
  long long nblocks;
  const float* par;
  int update_U;
  float* dd_out;
  float* scal_out;
};

template <int RP>
__global__ void __launch_bounds__(LR_THREADS, 1) k_lra_rotate_tc(const __grid_constant__ LrParams P) {
  using Cfg = LrCfg<RP>;
  constexpr int PACK = Cfg::PACK;
  constexpr int LR_STAGES = Cfg::STAGES;
  extern __shared__ uint8_t lr_smem_raw[];
  const uint32_t smem_base = (smem_u32(lr_smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = lr_smem_raw + (smem_base - smem_u32(lr_smem_raw));
  // layout: [U tile | V tile] x S, B_U, B_V (all 1024-aligned), vectors [d | h | v] x S, barriers
  const uint32_t bu_base = smem_base + LR_STAGES * 2 * LT_TILE;
  const uint32_t bv_base = bu_base + Cfg::BU_BYTES;
  const uint32_t vec_base = bv_base + Cfg::BV_BYTES;
  const uint32_t bar_base = vec_base + LR_STAGES * 3 * Cfg::VEC_BYTES;
  uint8_t* bu_gen = smem_gen + LR_STAGES * 2 * LT_TILE;
  uint8_t* bv_gen = bu_gen + Cfg::BU_BYTES;
  uint8_t* vec_gen = bv_gen + Cfg::BV_BYTES;
  auto full_bar = [&](int s_) { return bar_base + 8u * s_; };
  auto empty_bar = [&](int s_) { return bar_base + 8u * (LR_STAGES + s_); };
  auto tfull_bar = [&](int a_) { return bar_base + 8u * (2 * LR_STAGES + a_); };
  auto tempty_bar = [&](int a_) { return bar_base + 8u * (2 * LR_STAGES + 2 + a_); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * LR_STAGES + 4);
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(vec_gen + LR_STAGES * 3 * Cfg::VEC_BYTES + 8 * (2 * LR_STAGES + 4));
  __shared__ __align__(16) float wvec[2][RP];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const float* pvec = P.par + lra_par_vec_off(RP);
  const float* pscal = P.par + lra_par_scal_off(RP);
  const bf16* EuT = reinterpret_cast<const bf16*>(P.par + lra_par_et_off(RP));
  const bf16* EvT = EuT + RP * RP;
  const int update_U = P.update_U;

  if (tid == 0) {
    for (int i = 0; i < LR_STAGES; ++i) { mbar_init(full_bar(i), 1); mbar_init(empty_bar(i), 2); }   // MMA commit + the storing thread
    for (int i = 0; i < 2; ++i) { mbar_init(tfull_bar(i), 1); mbar_init(tempty_bar(i), 8); }     // 8 warps of the group that owns set i
    fence_barrier_init();
    prefetch_tmap(&P.map_u); prefetch_tmap(&P.map_v);
  }
  for (int e = tid; e < RP; e += LR_THREADS) {
    wvec[0][e] = update_U ? pvec[LV_WA * RP + e] : pvec[LV_ATU * RP + e];
    wvec[1][e] = update_U ? pvec[LV_WB * RP + e] : pvec[LV_BTU * RP + e];
  }
  // K-major operand tiles B_U / B_V: row n (output column), 64 k elements, 128-byte swizzle: element (n, k) at n * 128 + (((k >> 3) ^ (n & 7)) << 4) + (k & 7) * 2
  auto b_put = [&](uint8_t* base, int n, int k, float val) {
    *reinterpret_cast<bf16*>(base + n * 128 + (((k >> 3) ^ (n & 7)) << 4) + (k & 7) * 2) = __float2bfloat16_rn(val);
  };
  for (int e = tid; e < (Cfg::NU + Cfg::NV) * 64; e += LR_THREADS) {
    const bool isv = e >= Cfg::NU * 64;
    const int ee = isv ? e - Cfg::NU * 64 : e;
    const int n = ee >> 6, k = ee & 63;
    const int kc = k / RP, kk = k % RP;            // sub-row and column of the contraction index
    float val = 0.f;
    if (n < 64) {
      const int nc = n / RP, nn = n % RP;
      if (nc == kc) val = __bfloat162float((isv ? EvT : EuT)[nn * RP + kk]);
    } else {
      const int idx = n - 64;
      const int nvec = isv ? 4 : 2;
      const int c = idx / (2 * nvec), vi = (idx / 2) % nvec, part = idx & 1;
      if (c < PACK && c == kc) {
        const int id = isv ? (vi == 0 ? LV_AVC2 : (vi == 1 ? LV_AVS1 : (vi == 2 ? LV_AVATU : LV_AVBTU))) : (vi == 0 ? LV_AUC1 : LV_AUS2);
        const float x = (isv && vi >= 2 && update_U) ? 0.f : pvec[id * RP + kk];
        const float hi = lt_rbf(x);
        val = part ? x - hi : hi;
      }
    }
    b_put(isv ? bv_gen : bu_gen, n, k, val);
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
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
      if (++stage == LR_STAGES) { stage = 0; phase ^= 1u; }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: both operands K-major, M = 128 lines, K = 64 =====================
    const uint32_t idesc_u = (1u << 4) | (1u << 7) | (1u << 10) | (uint32_t(Cfg::NU >> 3) << 17) | (uint32_t(128 >> 4) << 24);
    const uint32_t idesc_v = (1u << 4) | (1u << 7) | (1u << 10) | (uint32_t(Cfg::NV >> 3) << 17) | (uint32_t(128 >> 4) << 24);
    const uint64_t adesc0 = make_smem_desc(smem_base, 0u, 1024u);
    const uint64_t budesc = make_smem_desc(bu_base, 0u, 1024u), bvdesc = make_smem_desc(bv_base, 0u, 1024u);
    int stage = 0; uint32_t phase = 0;
    int acc = 0; uint32_t acc_phase = 0;
    for (long long blk = blockIdx.x; blk < nblk; blk += gridDim.x) {
      mbar_wait(tempty_bar(acc), acc_phase ^ 1u, nullptr);
      mbar_wait(full_bar(stage), phase, nullptr);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t so = (uint64_t)((stage * 2 * LT_TILE) >> 4);
        const uint32_t du = tmem_base + uint32_t(acc * 256), dv = du + 96u;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          umma_bf16(du, adesc0 + so + (uint64_t)((k * 32) >> 4), budesc + (uint64_t)((k * 32) >> 4), idesc_u, k != 0 ? 1u : 0u);
          umma_bf16(dv, adesc0 + so + (uint64_t)((LT_TILE + k * 32) >> 4), bvdesc + (uint64_t)((k * 32) >> 4), idesc_v, k != 0 ? 1u : 0u);
        }
        umma_commit(tfull_bar(acc));
        umma_commit(empty_bar(stage));
      }
      __syncwarp();
      if (++stage == LR_STAGES) { stage = 0; phase ^= 1u; }
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
  } else {
    // ===================== epilogue: warp w owns TMEM lanes 32 (w % 4) .. + 31 = lines of the tile =====================
    const int quarter = warp & 3;
    const int grp = (warp - 2) >> 3;                 // epilogue group: tiles it = grp, grp + 2, ... of this CTA; accumulator set = grp
    const int side = ((warp - 2) >> 2) & 1;          // 0: U, 1: V
    const int ln = quarter * 32 + lane;
    const float step = pscal[LS_STEP], inv_rho = pscal[LS_INV_RHO], rho = pscal[LS_RHO];
    float mx1 = 0.f, mx2 = 0.f;
    const int acc = grp;
    int prev_stage = -1;                                         // stage whose TMA stores are still reading shared memory (storing thread)
    long long it = grp;
    for (long long blk = blockIdx.x + (long long)grp * gridDim.x; blk < nblk; blk += 2LL * gridDim.x, it += 2) {
      const int stage = (int)(it % LR_STAGES);
      const uint32_t phase = (uint32_t)((it / LR_STAGES) & 1);
      const uint32_t acc_phase = (uint32_t)((it >> 1) & 1);
      mbar_wait(full_bar(stage), phase, nullptr);                // the tile and the vectors (read below through the generic proxy)
      mbar_wait_relaxed(tfull_bar(acc), acc_phase, nullptr);
      tc_fence_after();
      const uint32_t trow = tmem_base + (uint32_t(quarter * 32) << 16) + uint32_t(acc * 256);
      uint8_t* tile_u = smem_gen + stage * 2 * LT_TILE + ln * 128;
      uint8_t* tile_v = tile_u + LT_TILE;
      const bf16* dvp = reinterpret_cast<const bf16*>(vec_gen + stage * 3 * Cfg::VEC_BYTES);
      const bf16* hvp = dvp + Cfg::ROWS;
      const bf16* vvp = hvp + Cfg::ROWS;
      const bool upd = side ? !update_U : (update_U != 0);        // does this side get the rank-2 update?
      lt_f2 nca2[PACK], cb2[PACK];
      if (side == 0 || upd) {                                      // dot columns: the U side writes dd_out, the updated side needs ca / cb
        uint32_t du_raw[32], dv_raw[32];
        tmem_ld_32x32(trow + 64u, du_raw);
        tmem_ld_32x32(trow + 96u + 64u, dv_raw);
        tmem_ld_wait();
        const long long row0 = blk * Cfg::ROWS + (long long)PACK * ln;
#pragma unroll
        for (int c = 0; c < PACK; ++c) {
          const float duc1 = __uint_as_float(du_raw[(c * 2 + 0) * 2]) + __uint_as_float(du_raw[(c * 2 + 0) * 2 + 1]);
          const float dus2 = __uint_as_float(du_raw[(c * 2 + 1) * 2]) + __uint_as_float(du_raw[(c * 2 + 1) * 2 + 1]);
          const float dvc2 = __uint_as_float(dv_raw[(c * 4 + 0) * 2]) + __uint_as_float(dv_raw[(c * 4 + 0) * 2 + 1]);
          const float dvs1 = __uint_as_float(dv_raw[(c * 4 + 1) * 2]) + __uint_as_float(dv_raw[(c * 4 + 1) * 2 + 1]);
          const float dva = __uint_as_float(dv_raw[(c * 4 + 2) * 2]) + __uint_as_float(dv_raw[(c * 4 + 2) * 2 + 1]);
          const float dvb = __uint_as_float(dv_raw[(c * 4 + 3) * 2]) + __uint_as_float(dv_raw[(c * 4 + 3) * 2 + 1]);
          const float dd = __bfloat162float(dvp[PACK * ln + c]), hh = __bfloat162float(hvp[PACK * ln + c]), vv = __bfloat162float(vvp[PACK * ln + c]);
          const float x1 = lt_rbf(dd * hh), x2 = lt_rbf(vv / dd);
          const float a = x1 + duc1;                    // Qh_i          psgd.py:1017
          const float b = x2 - dvs1;                    // invQtv_i      psgd.py:1024
          if (side == 0) {
            const float Ph = dd * (a + dvc2);           // Ph_i          psgd.py:1018
            const float invPv = __fdividef(b - dus2, dd);   // invPv_i   psgd.py:1025-1026
            const float Phh = Ph * hh, vinv = vv * invPv;
            mx1 = fmaxf(mx1, fabsf(Phh)); mx2 = fmaxf(mx2, fabsf(vinv));
            P.dd_out[row0 + c] = Phh - vinv;
          }
          const float ca = update_U ? step * a : step * (a + dva);
          const float cb = update_U ? step * b : step * (b + dvb);
          nca2[c] = lt_pack2(-ca, -ca);
          cb2[c] = lt_pack2(cb, cb);
        }
      }
      // rotation (identity part exact, correction from the MMA) + rank-2 update, 32 columns at a time; whole 16-byte pieces of the line;
      // two columns per FFMA2
      {
        uint8_t* tl = side ? tile_v : tile_u;
        const lt_f2 scale2 = side ? lt_pack2(rho, rho) : lt_pack2(inv_rho, inv_rho);
        const lt_f2 sgn2 = side ? lt_pack2(1.f, 1.f) : lt_pack2(-1.f, -1.f);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          uint32_t raw[32];
          tmem_ld_32x32(trow + (side ? 96u : 0u) + uint32_t(half * 32), raw);
          tmem_ld_wait();
#pragma unroll
          for (int q = 0; q < 4; ++q) {            // 16-byte piece j = 4 half + q of the line: columns 8 j .. 8 j + 7
            const int j = half * 4 + q;
            const uint4 ow = *reinterpret_cast<const uint4*>(tl + ((j ^ (ln & 7)) << 4));
            const uint32_t w[4] = {ow.x, ow.y, ow.z, ow.w};
            uint32_t o[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int col = 8 * j + 2 * e;        // column of the line; sub-row c = col / RP, rank index nn = col % RP
              const int c = col / RP, nn = col % RP;
              const lt_f2 x2 = lt_pack2(__uint_as_float(w[e] << 16), __uint_as_float(w[e] & 0xffff0000u));
              const lt_f2 r2 = lt_pack2(__uint_as_float(raw[8 * q + 2 * e]), __uint_as_float(raw[8 * q + 2 * e + 1]));
              const lt_f2 t = lt_fma2(sgn2, r2, x2);
              lt_f2 y;
              if (upd) {
                lt_f2 z = lt_mul2(cb2[c], *reinterpret_cast<const lt_f2*>(&wvec[1][nn]));
                z = lt_fma2(nca2[c], *reinterpret_cast<const lt_f2*>(&wvec[0][nn]), z);
                y = lt_fma2(t, scale2, z);          // (x + sgn raw) scale - (ca wa - cb wb)
              } else {
                y = lt_mul2(t, scale2);
              }
              const float2 yy = lt_unpack2(y);
              o[e] = pack_bf16(yy.x, yy.y);
            }
            *reinterpret_cast<uint4*>(tl + ((j ^ (ln & 7)) << 4)) = make_uint4(o[0], o[1], o[2], o[3]);   // in place: this thread owns the line
          }
        }
      }
      tc_fence_before();
      lt_fence_async_smem();                                      // the rewritten lines -> TMA store (async proxy)
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));
      if (grp == 0) asm volatile("bar.sync 1, 256;" ::: "memory");   // all eight warps of the group are done with the tile and its vectors
      else asm volatile("bar.sync 2, 256;" ::: "memory");
      if (tid == 64 + grp * 256) {
        const uint32_t su = smem_base + stage * 2 * LT_TILE;
        lt_tma_store_2d(&P.map_u, su, 0, (int)(blk * LT_KR));
        lt_tma_store_2d(&P.map_v, su + LT_TILE, 0, (int)(blk * LT_KR));
        lt_bulk_commit();
        if (prev_stage >= 0) { lt_bulk_wait_read<1>(); mbar_arrive(empty_bar(prev_stage)); }   // the previous tile's stores have left shared memory
        prev_stage = stage;
      }
    }
    if (tid == 64 + grp * 256 && prev_stage >= 0) { lt_bulk_wait_read<0>(); mbar_arrive(empty_bar(prev_stage)); }
    mx1 = warp_max(mx1); mx2 = warp_max(mx2);
    if (lane == 0 && side == 0) { atomic_max_nonneg(&P.scal_out[LS_MAX_PHH], mx1); atomic_max_nonneg(&P.scal_out[LS_MAX_VINV], mx2); }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

}  // namespace psgd
