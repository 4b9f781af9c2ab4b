// lra_mma.cuh -- tensor-core streaming kernels for the two sweeps of the LRA update (bf16, rank 16 or 32).
//
// Both sweeps are HBM-bound (6*n*r*2 bytes per update); on CUDA cores the r x r Grams (sweep 1) and the
// balancing rotations (sweep 2) make them compute-bound by 5-7x.  Here every warp streams its own 16-row chunks
//   global --cp.async 16 B--> warp-private smem ring (XOR-swizzled rows [U | V]) --ldmatrix--> mma.sync.m16n8k16 (bf16, fp32 acc)
// so the arithmetic rides on the tensor cores and the kernel runs at memory speed.  Warp-level mma.sync is the right tool
// here: the operands are 16 x 64 slivers with a 2-4 K-step reduction, far below a tcgen05 tile, and the roofline is HBM.
#pragma once
#include "common.cuh"

namespace psgd {

// parameter block (floats) written by k_lra_small:  [Au RP^2][Av RP^2][LV_NVEC vectors of RP][LS_NSCAL scalars][EuT bf16 RP^2][EvT bf16 RP^2]
enum { LV_AUC1 = 0, LV_AVC2, LV_AVS1, LV_AUS2, LV_AVATU, LV_AVBTU, LV_WA, LV_WB, LV_ATU, LV_BTU, LV_P1, LV_P2,
       LV_C1, LV_C2, LV_S1, LV_S2, LV_NVEC };
enum { LS_STEP = 0, LS_STEP_D = 1, LS_MAX_PHH = 2, LS_MAX_VINV = 3, LS_INV_RHO = 4, LS_RHO = 5, LS_NSCAL = 16 };

__host__ __device__ inline size_t lra_acc_floats(int RP) { return (size_t)3 * RP * RP + 4 * RP + 2; }
__host__ __device__ inline size_t lra_par_vec_off(int RP) { return (size_t)2 * RP * RP; }
__host__ __device__ inline size_t lra_par_scal_off(int RP) { return (size_t)2 * RP * RP + (size_t)LV_NVEC * RP; }
__host__ __device__ inline size_t lra_par_et_off(int RP) { return lra_par_scal_off(RP) + LS_NSCAL; }   // bf16 EuT then EvT
__host__ __device__ inline size_t lra_par_floats(int RP) { return lra_par_et_off(RP) + (size_t)RP * RP; }

__device__ __forceinline__ uint32_t smem_u32_generic(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t* r) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t addr, uint32_t* r) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float2 unpack_bf16(uint32_t u) { return __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&u)); }

// warp-private smem tile of one 16-row chunk: row = [U (RP bf16) | V (RP bf16)] = CPR 16-byte chunks; chunk c of row r sits at
// physical chunk c ^ swz(r) so that the 8 rows of an ldmatrix 8x8 block hit 8 different bank groups.
constexpr int LRA_STAGES = 8;
template <int RP> struct LraTile {
  static constexpr int CPR = RP / 4;            // 16-byte chunks per row (8 for RP=32, 4 for RP=16)
  static constexpr int ROW_BYTES = CPR * 16;
  static constexpr int VEC_OFF = 16 * ROW_BYTES;   // then one 128-byte line: d[16] | h[16] | v[16] (bf16) | pad
  static constexpr int BYTES = VEC_OFF + 128;
  __device__ static __forceinline__ int swz(int row) { return CPR == 8 ? (row & 7) : ((row >> 1) & 3); }
  __device__ static __forceinline__ uint32_t off(int row, int chunk) { return row * ROW_BYTES + ((chunk ^ swz(row)) << 4); }
};

// per-lane constants of the chunk loader: piece i of this lane copies 16 bytes from (U or V) + chunk_base + src_off[i] to tile + dst_off[i]
template <int RP> struct LraLoader {
  using Tl = LraTile<RP>;
  static constexpr int NP = 16 * Tl::CPR / 32;   // pieces per lane: 4 (RP=32) / 2 (RP=16)
  int src_off[NP];      // element offset inside the chunk
  uint32_t dst_off[NP];
  uint32_t is_v;        // bit i: piece i comes from V
  int row[NP];
  uint32_t vdst;        // vector piece (lanes 0..5): smem offset, source selector, element offset
  int vwhich, vpiece;
  __device__ __forceinline__ void init(int lane) {
    is_v = 0;
#pragma unroll
    for (int i = 0; i < NP; ++i) {
      const int p = lane + 32 * i;
      const int r = p / Tl::CPR, lc = p % Tl::CPR;
      const bool v = lc >= Tl::CPR / 2;
      row[i] = r;
      src_off[i] = r * RP + (v ? lc - Tl::CPR / 2 : lc) * 8;
      dst_off[i] = Tl::off(r, lc);
      is_v |= (v ? 1u : 0u) << i;
    }
    vwhich = lane >> 1; vpiece = lane & 1;
    vdst = Tl::VEC_OFF + vwhich * 32 + vpiece * 16;
  }
  // rows >= n are zero-filled; the full-chunk fast path carries no per-piece predicate
  __device__ __forceinline__ void issue(const bf16* __restrict__ U, const bf16* __restrict__ V, const bf16* __restrict__ d,
                                        const bf16* __restrict__ hvec, const bf16* __restrict__ vvec, long long n, long long ck, uint32_t tile,
                                        int lane) const {
    const long long r0 = ck * 16;
    const bf16* ub = U + r0 * RP;
    const bf16* vb = V + r0 * RP;
    if (r0 + 16 <= n) {
#pragma unroll
      for (int i = 0; i < NP; ++i) cp_async16(tile + dst_off[i], (((is_v >> i) & 1u) ? vb : ub) + src_off[i], 16);
      if (lane < 6) {
        const bf16* vec = vwhich == 0 ? d : (vwhich == 1 ? hvec : vvec);
        cp_async16(tile + vdst, vec + r0 + vpiece * 8, 16);
      }
    } else {
#pragma unroll
      for (int i = 0; i < NP; ++i) {
        const bool ok = r0 + row[i] < n;
        cp_async16(tile + dst_off[i], ok ? (const void*)((((is_v >> i) & 1u) ? vb : ub) + src_off[i]) : (const void*)U, ok ? 16 : 0);
      }
      if (lane < 6) {
        const bf16* vec = vwhich == 0 ? d : (vwhich == 1 ? hvec : vvec);
        long long valid = (n - (r0 + vpiece * 8)) * 2;
        valid = valid < 0 ? 0 : (valid > 16 ? 16 : valid);
        cp_async16(tile + vdst, valid > 0 ? (const void*)(vec + r0 + vpiece * 8) : (const void*)U, (int)valid);
      }
    }
  }
};

// ------------------------------------------------------------------------------------------------
// sweep 1: U^T U, V^T V, V^T U (RP x RP each) and the projections U^T x, V^T x for x1 = d.h, x2 = v/d, plus |x1|^2, |x2|^2
// (psgd.py:1006 and the r-sized right-hand sides of 1017-1052).  block = 8 warps, each warp streams its own chunks.
// accumulators per warp: tiles of the 2RP x 2RP Gram of W = [U V] that are needed (upper blocks of UtU / VtV, all of VtU) + 2RP x 8
// for the x columns.  Only RP = 32 and RP = 16 are instantiated.
// ------------------------------------------------------------------------------------------------
template <int RP>
__global__ void __launch_bounds__(256, 1) k_lra_gram_mma(const bf16* __restrict__ U, const bf16* __restrict__ V, const bf16* __restrict__ d,
                                                         const bf16* __restrict__ hvec, const bf16* __restrict__ vvec, long long n,
                                                         float* __restrict__ acc_out) {
  using Tl = LraTile<RP>;
  constexpr int STAGES = LRA_STAGES;
  constexpr int MT = RP / 8;        // 16-row m-tiles of W^T: 2RP / 16
  constexpr int NT = RP / 4;        // 8-col n-tiles of W: 2RP / 8
  constexpr int HM = MT / 2, HN = NT / 2;   // tiles belonging to U
  extern __shared__ __align__(128) uint8_t smem_lra[];
  __shared__ float blk_acc[3 * RP * RP + 4 * RP + 2];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  for (int e = threadIdx.x; e < 3 * RP * RP + 4 * RP + 2; e += blockDim.x) blk_acc[e] = 0.f;
  __syncthreads();
  const uint32_t ring = smem_u32_generic(smem_lra) + warp * STAGES * Tl::BYTES;
  LraLoader<RP> loader;
  loader.init(lane);

  uint32_t ldoff[MT];   // ldmatrix.x4.trans row addresses of this lane: matrices 0,1 = k rows 0-7, 2,3 = k rows 8-15; chunk 2c + (mi & 1)
#pragma unroll
  for (int c = 0; c < MT; ++c) {
    const int mi = lane >> 3, rr = lane & 7;
    ldoff[c] = Tl::off((mi >> 1) * 8 + rr, 2 * c + (mi & 1));
  }
  // accumulators: UtU tiles (mt < HM, nt in [2*mt, HN)), VtV tiles (mt >= HM, nt in [HN + 2*(mt-HM), NT)), VtU (mt >= HM, nt < HN), x (all mt)
  float aUU[HM][HN][4], aVV[HM][HN][4], aVU[HM][HN][4], aX[MT][4];
#pragma unroll
  for (int i = 0; i < HM; ++i)
#pragma unroll
    for (int j = 0; j < HN; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) { aUU[i][j][e] = 0.f; aVV[i][j][e] = 0.f; aVU[i][j][e] = 0.f; }
#pragma unroll
  for (int i = 0; i < MT; ++i)
#pragma unroll
    for (int e = 0; e < 4; ++e) aX[i][e] = 0.f;
  float sq = 0.f;

  const long long nchunks = (n + 15) / 16;
  const long long gw = (long long)blockIdx.x * 8 + warp, tw = (long long)gridDim.x * 8;
  // prologue: STAGES-1 chunks in flight
  long long ck_issue = gw;
#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (ck_issue < nchunks) loader.issue(U, V, d, hvec, vvec, n, ck_issue, ring + s * Tl::BYTES, lane);
    cp_async_commit();
    ck_issue += tw;
  }
  int stage = 0;
  for (long long ck = gw; ck < nchunks; ck += tw) {
    {  // keep the ring full
      const int s_issue = (stage + STAGES - 1) % STAGES;
      if (ck_issue < nchunks) loader.issue(U, V, d, hvec, vvec, n, ck_issue, ring + s_issue * Tl::BYTES, lane);
      cp_async_commit();
      ck_issue += tw;
    }
    // x columns: lanes with g == 0 carry x1 = d*h, g == 1 carry x2 = v/d (rows 2t, 2t+1, 2t+8, 2t+9 of the chunk), others zero
    cp_async_wait<STAGES - 1>();
    __syncwarp();
    const uint32_t tile = ring + stage * Tl::BYTES;
    // x columns: lanes 0-15 form x1 = d*h of row `lane`, lanes 16-31 x2 = v/d of row `lane-16`; the B fragment of the x tile
    // (column g: 0 -> x1, 1 -> x2, rest 0; rows 2t, 2t+1, 2t+8, 2t+9) is gathered with 4 shuffles
    uint32_t xb0, xb1;
    {
      const bf16* vl = reinterpret_cast<const bf16*>(smem_lra + warp * STAGES * Tl::BYTES + stage * Tl::BYTES + Tl::VEC_OFF);
      const int xr = lane & 15;
      float x = 0.f;
      if (ck * 16 + xr < n) {
        const float dd = __bfloat162float(vl[xr]);
        x = (lane < 16) ? __bfloat162float(__float2bfloat16_rn(dd * __bfloat162float(vl[16 + xr])))
                        : __bfloat162float(__float2bfloat16_rn(__bfloat162float(vl[32 + xr]) / dd));
      }
      sq = fmaf(x, x, sq);
      const int srcb = (g & 1) * 16 + 2 * t;
      float x0 = __shfl_sync(0xffffffffu, x, srcb), x1 = __shfl_sync(0xffffffffu, x, srcb + 1);
      float x2 = __shfl_sync(0xffffffffu, x, srcb + 8), x3 = __shfl_sync(0xffffffffu, x, srcb + 9);
      if (g >= 2) { x0 = 0.f; x1 = 0.f; x2 = 0.f; x3 = 0.f; }
      xb0 = pack_bf16(x0, x1);
      xb1 = pack_bf16(x2, x3);
    }
    // F[c] = transposed 8x8 blocks (k rows 0-7 | 8-15) x (column chunks 2c, 2c+1): A fragment of m-tile c and B fragments of n-tiles 2c, 2c+1
    uint32_t F[MT][4];
#pragma unroll
    for (int c = 0; c < MT; ++c) ldsm_x4_trans(tile + ldoff[c], F[c]);
    __syncwarp();
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const uint32_t b0 = F[nt >> 1][nt & 1], b1 = F[nt >> 1][2 + (nt & 1)];
        if (mt < HM && nt < HN) { if (nt >= 2 * mt) mma16816(aUU[mt][nt], F[mt], b0, b1); }
        else if (mt >= HM && nt >= HN) { if (nt - HN >= 2 * (mt - HM)) mma16816(aVV[mt - HM][nt - HN], F[mt], b0, b1); }
        else if (mt >= HM && nt < HN) mma16816(aVU[mt - HM][nt], F[mt], b0, b1);
      }
      mma16816(aX[mt], F[mt], xb0, xb1);
    }
    stage = (stage + 1) % STAGES;
  }
  cp_async_wait<0>();
  // ---- flush: warp -> block (smem atomics) -> global atomics ----
  float* bUU = blk_acc; float* bVV = blk_acc + RP * RP; float* bVU = blk_acc + 2 * RP * RP; float* bP = blk_acc + 3 * RP * RP;
#pragma unroll
  for (int mt = 0; mt < HM; ++mt)
#pragma unroll
    for (int nt = 0; nt < HN; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int row = 16 * mt + g + 8 * (e >> 1), col = 8 * nt + 2 * t + (e & 1);
        if (nt >= 2 * mt) {
          atomicAdd(&bUU[row * RP + col], aUU[mt][nt][e]);
          atomicAdd(&bVV[row * RP + col], aVV[mt][nt][e]);
          if (nt >= 2 * mt + 2) {   // strictly-upper block: mirror into the block that was skipped
            atomicAdd(&bUU[col * RP + row], aUU[mt][nt][e]);
            atomicAdd(&bVV[col * RP + row], aVV[mt][nt][e]);
          }
        }
        atomicAdd(&bVU[row * RP + col], aVU[mt][nt][e]);
      }
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int row = 16 * mt + g + 8 * (e >> 1), col = 2 * t + (e & 1);   // col 0: x1, col 1: x2
      if (col < 2) {
        const int isV = row >= RP;
        // layout: [Utx1][Vtx1][Utx2][Vtx2]
        atomicAdd(&bP[(col * 2 + isV) * RP + (row - isV * RP)], aX[mt][e]);
      }
    }
  {
    float v = sq;
#pragma unroll
    for (int o = 1; o < 16; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((lane & 15) == 0) atomicAdd(&bP[4 * RP + (lane >> 4)], v);
  }
  __syncthreads();
  for (int e = threadIdx.x; e < 3 * RP * RP + 4 * RP + 2; e += blockDim.x) atomicAdd(&acc_out[e], blk_acc[e]);
}

// ------------------------------------------------------------------------------------------------
// sweep 2: balancing rotation U' = (U - U Eu)/rho, V' = (V + V Ev) rho on tensor cores (Eu = E - E^2/2, Ev = E + E^2/2: the identity part
// is applied exactly in fp32, only the small correction goes through bf16 -- same rounding structure as psgd.py:1012-1015), per-row
// terms of the d update (psgd.py:1017-1029) and the rank-2 update of U or V (psgd.py:1036-1052), written back in place.
// ------------------------------------------------------------------------------------------------
template <int RP>
__global__ void __launch_bounds__(256, 1) k_lra_rotate_mma(bf16* __restrict__ U, bf16* __restrict__ V, const bf16* __restrict__ d,
                                                           const bf16* __restrict__ hvec, const bf16* __restrict__ vvec, long long n,
                                                           const float* __restrict__ par, int update_U, float* __restrict__ dd_out,
                                                           float* __restrict__ scal_out) {
  using Tl = LraTile<RP>;
  constexpr int STAGES = LRA_STAGES;
  constexpr int KS = RP / 16;   // k-steps of the rotation products
  constexpr int NT = RP / 8;    // 8-col n-tiles of one factor
  constexpr int HC = Tl::CPR / 2;
  extern __shared__ __align__(128) uint8_t smem_lra[];
  __shared__ float wvec[2][RP];   // rank-2 update directions: (w_a, w_b) for U, (atU, btU) for V
  __shared__ float red[32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const float* pvec = par + lra_par_vec_off(RP);
  const float* pscal = par + lra_par_scal_off(RP);
  const bf16* EuT = reinterpret_cast<const bf16*>(par + lra_par_et_off(RP));
  const bf16* EvT = EuT + RP * RP;
  for (int e = threadIdx.x; e < RP; e += blockDim.x) {
    wvec[0][e] = update_U ? pvec[LV_WA * RP + e] : pvec[LV_ATU * RP + e];
    wvec[1][e] = update_U ? pvec[LV_WB * RP + e] : pvec[LV_BTU * RP + e];
  }
  __syncthreads();
  const float step = pscal[LS_STEP], inv_rho = pscal[LS_INV_RHO], rho = pscal[LS_RHO];
  // B fragments of Eu, Ev (K x N "col"): b0 = (k = 16ks + 2t, +1 ; n = 8nt + g), b1 = k + 8.  E*T is stored N x K so the pair is one word.
  uint32_t bu[NT][KS][2], bv[NT][KS][2];
#pragma unroll
  for (int nt = 0; nt < NT; ++nt)
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      const int nn = 8 * nt + g, kk = 16 * ks + 2 * t;
      bu[nt][ks][0] = *reinterpret_cast<const uint32_t*>(EuT + nn * RP + kk);
      bu[nt][ks][1] = *reinterpret_cast<const uint32_t*>(EuT + nn * RP + kk + 8);
      bv[nt][ks][0] = *reinterpret_cast<const uint32_t*>(EvT + nn * RP + kk);
      bv[nt][ks][1] = *reinterpret_cast<const uint32_t*>(EvT + nn * RP + kk + 8);
    }
  // the per-row dot products ride on the tensor cores too: one extra 8-column tile per factor whose columns are the (hi, lo) bf16
  // split of the rotated vectors -- U'_i . c = U_i . (Au c) --  U tile: [Au c1 | Au s2 | 0 0];  V tile: [Av c2 | Av s1 | Av atU | Av btU]
  uint32_t du[KS][2], dv[KS][2];
  {
    const int vi = g >> 1, lo = g & 1;
    const int uvec = vi == 0 ? LV_AUC1 : LV_AUS2;
    const int vvecid = vi == 0 ? LV_AVC2 : (vi == 1 ? LV_AVS1 : (vi == 2 ? LV_AVATU : LV_AVBTU));
#pragma unroll
    for (int ks = 0; ks < KS; ++ks)
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        const int kk = 16 * ks + 2 * t + 8 * hf;
        float f[2], w[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const float xu = (vi < 2) ? pvec[uvec * RP + kk + e] : 0.f;
          const float xv = (update_U && vi >= 2) ? 0.f : pvec[vvecid * RP + kk + e];
          const float hu = __bfloat162float(__float2bfloat16_rn(xu)), hv = __bfloat162float(__float2bfloat16_rn(xv));
          f[e] = lo ? xu - hu : hu;
          w[e] = lo ? xv - hv : hv;
        }
        du[ks][hf] = pack_bf16(f[0], f[1]);
        dv[ks][hf] = pack_bf16(w[0], w[1]);
      }
  }
  const uint32_t ring = smem_u32_generic(smem_lra) + warp * STAGES * Tl::BYTES;
  LraLoader<RP> loader;
  loader.init(lane);
  uint8_t* ring_gen = smem_lra + warp * STAGES * Tl::BYTES;
  uint32_t ldoff_u[KS], ldoff_v[KS];   // ldmatrix.x4 (row-major A): matrices (rows 0-7, chunk 2ks), (rows 8-15, 2ks), (rows 0-7, 2ks+1), (rows 8-15, 2ks+1)
#pragma unroll
  for (int ks = 0; ks < KS; ++ks) {
    const int mi = lane >> 3, rr = lane & 7;
    const int row = (mi & 1) * 8 + rr, chunk = 2 * ks + (mi >> 1);
    ldoff_u[ks] = Tl::off(row, chunk);
    ldoff_v[ks] = Tl::off(row, HC + chunk);
  }
  uint32_t stoff[NT][2];   // where element pair (row g + 8q, cols 8nt + 2t, +1) of U lives in the tile (V: chunk + HC)
#pragma unroll
  for (int nt = 0; nt < NT; ++nt)
#pragma unroll
    for (int q = 0; q < 2; ++q) stoff[nt][q] = Tl::off(g + 8 * q, nt) + 4 * t;
  float mx1 = 0.f, mx2 = 0.f;
  const long long nchunks = (n + 15) / 16;
  const long long gw = (long long)blockIdx.x * 8 + warp, tw = (long long)gridDim.x * 8;
  long long ck_issue = gw;
#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (ck_issue < nchunks) loader.issue(U, V, d, hvec, vvec, n, ck_issue, ring + s * Tl::BYTES, lane);
    cp_async_commit();
    ck_issue += tw;
  }
  int stage = 0;
  for (long long ck = gw; ck < nchunks; ck += tw) {
    // the slot refilled now is the one whose coalesced stores were issued last iteration (they read smem synchronously) -> safe
    {
      const int s_issue = (stage + STAGES - 1) % STAGES;
      if (ck_issue < nchunks) loader.issue(U, V, d, hvec, vvec, n, ck_issue, ring + s_issue * Tl::BYTES, lane);
      cp_async_commit();
      ck_issue += tw;
    }
    cp_async_wait<STAGES - 1>();
    __syncwarp();
    const uint32_t tile = ring + stage * Tl::BYTES;
    uint8_t* tile_gen = ring_gen + stage * Tl::BYTES;
    // per-row inputs of rows g and g+8 (streamed with the tile)
    float dd[2], hh[2], vv[2];
    bool rok[2];
    {
      const bf16* vl = reinterpret_cast<const bf16*>(tile_gen + Tl::VEC_OFF);
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int lr = g + 8 * q;
        rok[q] = ck * 16 + lr < n;
        dd[q] = rok[q] ? __bfloat162float(vl[lr]) : 1.f;
        hh[q] = rok[q] ? __bfloat162float(vl[16 + lr]) : 0.f;
        vv[q] = rok[q] ? __bfloat162float(vl[32 + lr]) : 0.f;
      }
    }
    uint32_t au[KS][4], av[KS][4];
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) { ldsm_x4(tile + ldoff_u[ks], au[ks]); ldsm_x4(tile + ldoff_v[ks], av[ks]); }
    float cu[NT][4], cv[NT][4], dotu[4] = {0.f, 0.f, 0.f, 0.f}, dotv[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) { mma16816(dotu, au[ks], du[ks][0], du[ks][1]); mma16816(dotv, av[ks], dv[ks][0], dv[ks][1]); }
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
#pragma unroll
      for (int e = 0; e < 4; ++e) { cu[nt][e] = 0.f; cv[nt][e] = 0.f; }
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) { mma16816(cu[nt], au[ks], bu[nt][ks][0], bu[nt][ks][1]); mma16816(cv[nt], av[ks], bv[nt][ks][0], bv[nt][ks][1]); }
    }
    // dot tile: lane t holds vector t's (hi, lo) columns for rows g (c0, c1) and g+8 (c2, c3); broadcast inside the quad
    float a_[2], b_[2], ca[2], cb[2];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const float su = dotu[2 * q] + dotu[2 * q + 1], sv = dotv[2 * q] + dotv[2 * q + 1];
      const int qb = lane & ~3;
      const float duc1 = __shfl_sync(0xffffffffu, su, qb), dus2 = __shfl_sync(0xffffffffu, su, qb + 1);
      const float dvc2 = __shfl_sync(0xffffffffu, sv, qb), dvs1 = __shfl_sync(0xffffffffu, sv, qb + 1);
      const float dva = __shfl_sync(0xffffffffu, sv, qb + 2), dvb = __shfl_sync(0xffffffffu, sv, qb + 3);
      const float x1 = __bfloat162float(__float2bfloat16_rn(dd[q] * hh[q]));
      const float x2 = __bfloat162float(__float2bfloat16_rn(vv[q] / dd[q]));
      const float a = x1 + duc1;                    // Qh_i          psgd.py:1017
      const float Ph = dd[q] * (a + dvc2);          // Ph_i          psgd.py:1018
      const float b = x2 - dvs1;                    // invQtv_i      psgd.py:1024
      const float invPv = (b - dus2) / dd[q];       // invPv_i       psgd.py:1025-1026
      const float Phh = Ph * hh[q], vinv = vv[q] * invPv;
      if (rok[q]) {
        mx1 = fmaxf(mx1, fabsf(Phh)); mx2 = fmaxf(mx2, fabsf(vinv));
        if (t == 0) dd_out[ck * 16 + g + 8 * q] = Phh - vinv;
      }
      a_[q] = a; b_[q] = b;
      ca[q] = update_U ? step * a : step * (a + dva);
      cb[q] = update_U ? step * b : step * (b + dvb);
    }
    __syncwarp();   // all lanes have consumed their ldmatrix data before the tile is overwritten
    // rotation (identity part exact, correction from the MMA), rank-2 update, write back into the tile at the fragment positions
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      const int ks = nt >> 1, hi = nt & 1;
      const float wa0 = wvec[0][8 * nt + 2 * t], wa1 = wvec[0][8 * nt + 2 * t + 1];
      const float wb0 = wvec[1][8 * nt + 2 * t], wb1 = wvec[1][8 * nt + 2 * t + 1];
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const float2 u0 = unpack_bf16(au[ks][2 * hi + q]);
        const float2 v0 = unpack_bf16(av[ks][2 * hi + q]);
        float un0 = (u0.x - cu[nt][2 * q]) * inv_rho, un1 = (u0.y - cu[nt][2 * q + 1]) * inv_rho;
        float vn0 = (v0.x + cv[nt][2 * q]) * rho, vn1 = (v0.y + cv[nt][2 * q + 1]) * rho;
        const float w0 = ca[q] * wa0 - cb[q] * wb0, w1 = ca[q] * wa1 - cb[q] * wb1;
        if (update_U) { un0 -= w0; un1 -= w1; } else { vn0 -= w0; vn1 -= w1; }
        *reinterpret_cast<uint32_t*>(tile_gen + stoff[nt][q]) = pack_bf16(un0, un1);
        *reinterpret_cast<uint32_t*>(tile_gen + (stoff[nt][q] ^ (uint32_t(HC) << 4))) = pack_bf16(vn0, vn1);
      }
    }
    __syncwarp();
    // coalesced 16-byte stores, mirror image of the loader
    if (ck * 16 + 16 <= n) {
#pragma unroll
      for (int i = 0; i < LraLoader<RP>::NP; ++i) {
        const uint4 val = *reinterpret_cast<const uint4*>(tile_gen + loader.dst_off[i]);
        bf16* dst = (((loader.is_v >> i) & 1u) ? V : U) + ck * 16 * RP + loader.src_off[i];
        *reinterpret_cast<uint4*>(dst) = val;
      }
    } else {
#pragma unroll
      for (int i = 0; i < LraLoader<RP>::NP; ++i) {
        if (ck * 16 + loader.row[i] < n) {
          const uint4 val = *reinterpret_cast<const uint4*>(tile_gen + loader.dst_off[i]);
          bf16* dst = (((loader.is_v >> i) & 1u) ? V : U) + ck * 16 * RP + loader.src_off[i];
          *reinterpret_cast<uint4*>(dst) = val;
        }
      }
    }
    __syncwarp();
    stage = (stage + 1) % STAGES;
  }
  cp_async_wait<0>();
  mx1 = block_max(mx1, red);
  if (threadIdx.x == 0) atomic_max_nonneg(&scal_out[LS_MAX_PHH], mx1);
  mx2 = block_max(mx2, red);
  if (threadIdx.x == 0) atomic_max_nonneg(&scal_out[LS_MAX_VINV], mx2);
}

// ------------------------------------------------------------------------------------------------
// apply sweeps (psgd.py:1055-1063), bf16, rank 16/32: fully coalesced mapping -- lane = (row within a group of 32/PPR rows, 16-byte
// piece of that row); row dot products are reduced across the PPR lanes of a row with shuffles, the r-vector accumulators are split by
// piece so every lane carries only 8 of them.
//   mode 0: p1 += V_i * (d_i g_i)      mode 1: g2_i = d_i g_i + U_i . p1 (fp32, stored) ; p2 += U_i * g2_i      mode 2: out_i = d_i (g2_i + V_i . p2)
// ------------------------------------------------------------------------------------------------
template <int RP>
__global__ void __launch_bounds__(256) k_lra_apply_bf16(const bf16* __restrict__ Mtx, const bf16* __restrict__ d, const bf16* __restrict__ g,
                                                        float* __restrict__ g2, bf16* __restrict__ out, long long n, int mode,
                                                        const float* __restrict__ pin, float* __restrict__ pout, float* sumsq) {
  constexpr int PPR = RP / 8;          // 16-byte pieces per row
  constexpr int RPI = 32 / PPR;        // rows per warp instruction
  constexpr int UNR = 4;
  __shared__ float pacc[RP];
  const int lane = threadIdx.x & 31;
  const int piece = lane % PPR, rsub = lane / PPR;
  if (threadIdx.x < RP) pacc[threadIdx.x] = 0.f;
  float pv[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) pv[c] = (mode > 0) ? pin[piece * 8 + c] : 0.f;
  __syncthreads();
  float acc[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) acc[c] = 0.f;
  float ssq = 0.f;
  const long long gwarp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const long long ngroups = (n + RPI - 1) / RPI;
  for (long long grp0 = gwarp * UNR; grp0 < ngroups; grp0 += nwarps * UNR) {
    uint4 raw[UNR];
    float dd[UNR], gg[UNR], g2v[UNR];
    bool ok[UNR];
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const long long row = (grp0 + u) * RPI + rsub;
      ok[u] = row < n;
      raw[u] = ok[u] ? *reinterpret_cast<const uint4*>(Mtx + row * RP + piece * 8) : make_uint4(0u, 0u, 0u, 0u);
      dd[u] = ok[u] ? __bfloat162float(d[row]) : 0.f;
      gg[u] = (ok[u] && mode < 2) ? __bfloat162float(g[row]) : 0.f;
      g2v[u] = (ok[u] && mode == 2) ? g2[row] : 0.f;
    }
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const long long row = (grp0 + u) * RPI + rsub;
      float x[8];
      const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&raw[u]);
#pragma unroll
      for (int c = 0; c < 4; ++c) { float2 f = __bfloat1622float2(h2[c]); x[2 * c] = f.x; x[2 * c + 1] = f.y; }
      if (mode == 0) {
        const float y = __bfloat162float(__float2bfloat16_rn(dd[u] * gg[u]));
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[c] = fmaf(x[c], y, acc[c]);
      } else {
        float dot = 0.f;
#pragma unroll
        for (int c = 0; c < 8; ++c) dot = fmaf(x[c], pv[c], dot);
#pragma unroll
        for (int o = 1; o < PPR; o <<= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
        if (mode == 1) {
          const float y = __bfloat162float(__float2bfloat16_rn(dd[u] * gg[u])) + dot;
          if (ok[u] && piece == 0) g2[row] = y;
#pragma unroll
          for (int c = 0; c < 8; ++c) acc[c] = fmaf(x[c], ok[u] ? y : 0.f, acc[c]);
        } else {
          const bf16 o = __float2bfloat16_rn(dd[u] * (g2v[u] + dot));
          if (ok[u] && piece == 0) { out[row] = o; const float f = __bfloat162float(o); ssq = fmaf(f, f, ssq); }
        }
      }
    }
  }
  if (mode < 2) {
    // reduce over the lanes that share a piece (lane bits above log2(PPR)), then block (smem), then global
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      float v = acc[c];
#pragma unroll
      for (int o = PPR; o < 32; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (rsub == 0) atomicAdd(&pacc[piece * 8 + c], v);
    }
    __syncthreads();
    if (threadIdx.x < RP) atomicAdd(&pout[threadIdx.x], pacc[threadIdx.x]);
  } else if (sumsq) {
    float v = warp_sum(ssq);
    if (lane == 0) atomicAdd(sumsq, v);
  }
}

// ------------------------------------------------------------------------------------------------
// ------------------------------------------------------------------------------------------------
// apply sweeps fed by the bulk-copy engine: one producer thread issues cp.async.bulk (TMA 1-D) copies of whole 256-row blocks of the
// factor (16 KB for r = 32) plus the block's d / g / g2 slices into an 8-stage mbarrier ring; 8 consumer warps do the row dot products
// from shared memory.  ~130 KB in flight per SM for one instruction per 16 KB: the per-thread-load versions topped out at 3.2-3.8 TB/s.
// n_full = number of whole 256-row blocks; the (< 256 rows) remainder is handled by k_lra_apply_bf16 on offset pointers.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void lra_mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void lra_mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void lra_mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void lra_mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  long long t0 = 0;
  uint32_t spins = 0;
  while (true) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (done) break;
    if (++spins == 1024u) t0 = clock64();
    if (spins > 1024u && (spins & 1023u) == 0u && clock64() - t0 > 8000000000LL) __trap();   // never hang the GPU on a pipeline bug
  }
}
__device__ __forceinline__ void lra_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

constexpr int LRA_APPLY_CW = 16;   // consumer warps of k_lra_apply_tma: a pure-read sweep needs >= 16 resident warps per SM to cover the
                                   // dot-product latency chains (8 warps top out near 4 TB/s on this part, tools/read_bw_probe.py)
template <int RP>
__global__ void __launch_bounds__(32 * (LRA_APPLY_CW + 1), 1) k_lra_apply_tma(const bf16* __restrict__ Mtx, const bf16* __restrict__ d, const bf16* __restrict__ g,
                                                          float* __restrict__ g2, bf16* __restrict__ out, long long n_full, int mode,
                                                          const float* __restrict__ pin, float* __restrict__ pout, float* sumsq) {
  constexpr int ROWS = 256;
  constexpr int PPR = RP / 8;
  constexpr int MAT_BYTES = ROWS * RP * 2;
  constexpr int OFF_D = MAT_BYTES, OFF_G = OFF_D + ROWS * 2, OFF_G2 = OFF_G + ROWS * 2;
  constexpr int TILE = OFF_G2 + ROWS * 4;
  constexpr int STAGES = 8;
  constexpr int NCW = LRA_APPLY_CW;
  constexpr int RPW = ROWS / NCW;     // rows per consumer warp and stage
  constexpr int NP = RPW * PPR / 32;  // 16-byte pieces per lane for a warp's rows
  static_assert(NP >= 1 && RPW * PPR == NP * 32, "consumer warp tiling");
  extern __shared__ __align__(128) uint8_t smem_lra[];
  __shared__ __align__(8) uint64_t bars[2 * STAGES];
  __shared__ float pacc[RP];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t sbase = smem_u32_generic(smem_lra);
  const uint32_t bbase = smem_u32_generic(bars);
  if (threadIdx.x == 0) {
    for (int s0 = 0; s0 < STAGES; ++s0) { lra_mbar_init(bbase + 8 * s0, 1); lra_mbar_init(bbase + 8 * (STAGES + s0), NCW); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < RP) pacc[threadIdx.x] = 0.f;
  __syncthreads();
  if (warp == 0) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (long long blk = blockIdx.x; blk < n_full; blk += gridDim.x) {
        lra_mbar_wait(bbase + 8 * (STAGES + stage), phase ^ 1u);
        const uint32_t tile = sbase + stage * TILE, fb = bbase + 8 * stage;
        const long long r0 = blk * ROWS;
        const uint32_t vbytes = (mode == 2) ? ROWS * 4 : ROWS * 2;
        lra_mbar_expect_tx(fb, MAT_BYTES + ROWS * 2 + vbytes);
        lra_bulk_g2s(tile, Mtx + r0 * RP, MAT_BYTES, fb);
        lra_bulk_g2s(tile + OFF_D, d + r0, ROWS * 2, fb);
        if (mode == 2) lra_bulk_g2s(tile + OFF_G2, g2 + r0, ROWS * 4, fb);
        else lra_bulk_g2s(tile + OFF_G, g + r0, ROWS * 2, fb);
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
    }
  } else {
    const int cw = warp - 1;             // consumer warp -> rows [RPW cw, RPW cw + RPW) of the stage
    const int piece = lane % PPR;
    float pv[8], acc[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) { pv[c] = (mode > 0) ? pin[piece * 8 + c] : 0.f; acc[c] = 0.f; }
    float ssq = 0.f;
    int stage = 0; uint32_t phase = 0;
    for (long long blk = blockIdx.x; blk < n_full; blk += gridDim.x) {
      lra_mbar_wait(bbase + 8 * stage, phase);
      const uint8_t* tile = smem_lra + stage * TILE;
      const bf16* dv = reinterpret_cast<const bf16*>(tile + OFF_D);
      const bf16* gv = reinterpret_cast<const bf16*>(tile + OFF_G);
      const float* g2v = reinterpret_cast<const float*>(tile + OFF_G2);
      const long long r0 = blk * ROWS;
#pragma unroll
      for (int i = 0; i < NP; ++i) {
        const int p = lane + 32 * i;
        const int lrow = cw * RPW + p / PPR;
        const uint4 raw = *reinterpret_cast<const uint4*>(tile + (size_t)lrow * RP * 2 + (p % PPR) * 16);
        float x[8];
        const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
        for (int c = 0; c < 4; ++c) { float2 f = __bfloat1622float2(h2[c]); x[2 * c] = f.x; x[2 * c + 1] = f.y; }
        const float dd = __bfloat162float(dv[lrow]);
        if (mode == 0) {
          const float y = __bfloat162float(__float2bfloat16_rn(dd * __bfloat162float(gv[lrow])));
#pragma unroll
          for (int c = 0; c < 8; ++c) acc[c] = fmaf(x[c], y, acc[c]);
        } else {
          float dot = 0.f;
#pragma unroll
          for (int c = 0; c < 8; ++c) dot = fmaf(x[c], pv[c], dot);
#pragma unroll
          for (int o = 1; o < PPR; o <<= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
          if (mode == 1) {
            const float y = __bfloat162float(__float2bfloat16_rn(dd * __bfloat162float(gv[lrow]))) + dot;
            if (piece == 0) g2[r0 + lrow] = y;
#pragma unroll
            for (int c = 0; c < 8; ++c) acc[c] = fmaf(x[c], y, acc[c]);
          } else if (piece == 0) {
            const bf16 o = __float2bfloat16_rn(dd * (g2v[lrow] + dot));
            out[r0 + lrow] = o;
            const float f = __bfloat162float(o);
            ssq = fmaf(f, f, ssq);
          }
        }
      }
      __syncwarp();
      if (lane == 0) lra_mbar_arrive(bbase + 8 * (STAGES + stage));
      if (++stage == STAGES) { stage = 0; phase ^= 1u; }
    }
    if (mode < 2) {
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        float v = acc[c];
#pragma unroll
        for (int o = PPR; o < 32; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane < PPR) atomicAdd(&pacc[piece * 8 + c], v);
      }
    } else if (sumsq) {
      float v = warp_sum(ssq);
      if (lane == 0) atomicAdd(sumsq, v);
    }
  }
  __syncthreads();
  if (mode < 2 && threadIdx.x < RP) atomicAdd(&pout[threadIdx.x], pacc[threadIdx.x]);
}

}  // namespace psgd
