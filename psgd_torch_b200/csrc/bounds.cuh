// bounds.cuh -- norm_lower_bound_spd / norm_lower_bound_skh (psgd.py:46-93) as ONE persistent cooperative kernel per group of matrices.
//
// The reference evaluates a bound as: nf = max diag (spd) / max |A| (skh); A' = A / nf; j = argmax row norm; V = A'[j] + sgn(<A'[j], V0>) V0
// (32 probes); then four dependent products V <- V A' with a row normalisation after the 1st and 3rd, and returns nf * max row norm.
// Each product needs ALL of A (32 MB at s = 4096) for 1 GFLOP of work: memory-bound, and A (plus the second factor's matrix) fits in the
// 126 MB L2, so the whole evaluation can run out of L2 -- if it is one kernel.  Round 1 issued it as 1 + 4 + 1 launches per pair of
// factors (about 100 us at s = 4096); here every CTA of a persistent grid keeps a 64-column slab of one matrix per step and the steps are
// separated by grid barriers:
//
//   phase I   : per slab, partial dot products <A'[j], V0[p]> (the signs of the probe rotation, psgd.py:63)          -> barrier
//   step 0    : V1 = (A'[j] + sgn V0) A'      the rotated probes are formed on the fly in shared memory               -> barrier
//   step 1    : V2 = normalise(V1) A'         the normalisation (psgd.py:66) is a per-probe factor of the epilogue    -> barrier
//   step 2    : V3 = V2 A'                                                                                             -> barrier
//   step 3    : row norms of normalise(V3) A' only (nothing stored)
//   finish    : the CTA that finishes last turns the norms into the bound and into what its consumer needs (Lipschitz update + step
//               sizes of the dense factor, psgd.py:413-415; the Procrustes normaliser, psgd.py:118)
//
// Work unit = (matrix, 64-column slab): out[32 x 64] = V[32 x s] A[s x 64].  K is split over the 8 warps of the CTA in interleaved 16-row
// slices; every warp streams ITS slices (A: 16 x 64, V: 32 x 16, XOR-swizzled for ldmatrix) through a warp-private cp.async ring and
// multiplies them on mma.sync m16n8k16 (bf16 in, fp32 accumulate; 32 probes are far below a tcgen05 tile and the bound is L2 bandwidth) --
// no block-wide barrier inside the K loop (a first version with block-wide 128-row stages spent its time in per-stage barriers and address
// arithmetic: 92 us per pair of 4096 x 4096 matrices, profiles/r02_ncu_bounds_v1.txt).  Then a cross-warp reduction through shared memory
// and the epilogue: scale, round to bf16, store, per-probe sums of squares (atomics, 32 per unit).
#pragma once
#include "common.cuh"
#include "kron_kernels.cuh"
#include "tc_ptx.cuh"

namespace psgd {

constexpr int NB_W = 64;          // columns of A per work unit
constexpr int NB_KS = 16;         // k rows per warp slice
constexpr int NB_KC = 128;        // k rows all 8 warps cover together (padding granularity of the row kept in shared memory)
constexpr int NB_STAGES = 6;      // slices in flight per warp
constexpr int NB_THREADS = 256;
constexpr int NB_MAX_JOBS = 16;
constexpr int NB_MAX_S = 16384;   // the selected row of A is kept in shared memory (2 bytes per column)
constexpr int NB_A_BYTES = NB_KS * NB_W * 2;    // 2 KB: 16 rows x 128 B
constexpr int NB_V_BYTES = 32 * NB_KS * 2;      // 1 KB: 32 probes x 32 B
constexpr int NB_STAGE_BYTES = NB_A_BYTES + NB_V_BYTES;
constexpr int NB_RING_BYTES = 8 * NB_STAGES * NB_STAGE_BYTES;
constexpr int NB_RED_LD = 72;     // floats per row of the cross-warp reduction buffer (64 + 8: conflict-free float2 stores)
static_assert(8 * 32 * NB_RED_LD * 4 <= NB_RING_BYTES, "reduction buffer aliases the warp rings");

struct NbJob {
  const bf16* A;          // s x s, row-major, ld = s
  const bf16* V0;         // 32 x s probes (psgd.py:62 / 87)
  const float* row_sumsq; // s: squared row norms of A (from the producing kernel's epilogue)
  const float* nf_src;    // max diag (spd) or max |A| (skh)
  bf16* Va; bf16* Vb;     // 32 x s ping-pong
  float* scal;            // SC_* block
  float* rn1; float* rn3; float* rn4; float* dots;   // 32 floats each, zero on entry
  int s;
  int unit0, nunits;
  int mode;               // finish: 0 dense-factor L update + step sizes, 1 Procrustes normaliser, 2 bound only
  float t2, lr, betaL;
  float* L; float* fs;
};

struct NbParams {
  NbJob job[NB_MAX_JOBS];
  int njobs;
  int total_units;
  int dtype;
  float tiny;
  unsigned* barrier;      // grid barrier counter (zero on entry, reset by the finishing CTA)
  unsigned* done;
};

__device__ __forceinline__ unsigned nb_ld_acquire(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// all CTAs of the (cooperative, co-resident) grid; `target` = gridDim.x * number of barriers so far.  Bounded spin: a bug must trap,
// not hang the GPU.
__device__ __forceinline__ void nb_grid_barrier(unsigned* counter, unsigned target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(counter, 1u);
    long long t0 = 0;
    unsigned spins = 0;
    while (nb_ld_acquire(counter) < target) {
      if (++spins == 65536u) t0 = clock64();
      if (spins > 65536u && (spins & 4095u) == 0u && clock64() - t0 > 8000000000LL) __trap();
    }
    __threadfence();
  }
  __syncthreads();
}

__device__ __forceinline__ void nb_cp_async16(uint32_t dst, const void* src, int bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void nb_cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void nb_cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void nb_ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void nb_ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void nb_mma(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// argmax_i row_sumsq[i] (first maximal index, like torch.argmax) by the whole block; result broadcast through shared memory
__device__ __forceinline__ int nb_block_argmax(const float* __restrict__ x, int s, float* sv, int* si) {
  float best = -1.f;
  int bi = 0x7fffffff;
  for (int i = threadIdx.x; i < s; i += NB_THREADS) {
    const float v = __ldcg(x + i);
    if (v > best) { best = v; bi = i; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
  }
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) { sv[w] = best; si[w] = bi; }
  __syncthreads();
  if (w == 0) {
    best = lane < NB_THREADS / 32 ? sv[lane] : -1.f;
    bi = lane < NB_THREADS / 32 ? si[lane] : 0x7fffffff;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    if (lane == 0) si[8] = (bi == 0x7fffffff) ? 0 : bi;
  }
  __syncthreads();
  return si[8];
}

// bound = nf * max_p sqrt(rn[p]) (psgd.py:68) and its consumer -- the body of k_bound_finish for one matrix, run by one thread
__device__ __forceinline__ void nb_finish_job(const NbJob& J, int dtype, float tiny) {
  float v = 0.f;
  for (int p = 0; p < 32; ++p) v = fmaxf(v, __ldcg(J.rn4 + p));
  const float nf = __ldcg(J.nf_src) + tiny;
  const float bound = round_to(dtype, nf * round_to(dtype, sqrtf(v)));
  J.scal[SC_NF] = nf;
  J.scal[SC_INV_NF] = 1.f / nf;
  J.scal[SC_BOUND] = bound;
  if (J.mode == 0) {
    const float ell = round_to(dtype, bound + J.t2);
    const float Ln = fmaxf(J.betaL * (*J.L) + (1.f - J.betaL) * ell, ell);
    *J.L = Ln;
    const float c = J.lr / Ln;
    J.fs[FS_ALPHA] = -c;
    J.fs[FS_BETA] = 1.f + c * J.t2;
  } else if (J.mode == 1) {
    J.fs[FS_INV_SR] = 1.f / (bound + tiny);
  }
}

__global__ void __launch_bounds__(NB_THREADS, 1) k_norm_bounds(const __grid_constant__ NbParams P) {
  extern __shared__ __align__(128) uint8_t nb_smem[];
  bf16* a_row = reinterpret_cast<bf16*>(nb_smem + NB_RING_BYTES);   // row j of A', zero beyond s (up to a multiple of NB_KC)
  float* red = reinterpret_cast<float*>(nb_smem);
  __shared__ float s_sgn[32];
  __shared__ float s_scale[32];
  __shared__ float s_sv[8];
  __shared__ int s_si[9];
  __shared__ int s_last;
  __shared__ int s_jc[NB_MAX_JOBS];   // argmax row of the matrices this CTA has met (phase I -> step 0)
  const uint32_t smem_base = static_cast<uint32_t>(__cvta_generic_to_shared(nb_smem));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const unsigned ncta = gridDim.x;
  unsigned nbar = 0;
  if (tid < NB_MAX_JOBS) s_jc[tid] = -1;
  __syncthreads();

  auto job_of = [&](int u) {
    int ji = 0;
    for (int i = 1; i < P.njobs; ++i)
      if (u >= P.job[i].unit0) ji = i;
    return ji;
  };

  // ------------------------------ phase I: signs of the probe rotation (psgd.py:63) ------------------------------
  {
    int cur = -1, j = 0;
    float inv_nf = 0.f;
    for (int u = blockIdx.x; u < P.total_units; u += ncta) {
      const int ji = job_of(u);
      const NbJob& J = P.job[ji];
      if (ji != cur) {
        j = nb_block_argmax(J.row_sumsq, J.s, s_sv, s_si);
        if (tid == 0) s_jc[ji] = j;
        inv_nf = 1.f / (__ldcg(J.nf_src) + P.tiny);
        cur = ji;
      }
      const int p = tid >> 3, c0 = (u - J.unit0) * NB_W + (tid & 7) * 8;
      float part = 0.f;
      if (c0 < J.s) {
        float a[8], v[8];
        ld8(J.A + (size_t)j * J.s + c0, a);
        ld8(J.V0 + (size_t)p * J.s + c0, v);
#pragma unroll
        for (int t = 0; t < 8; ++t) part += rbf(rbf(a[t] * inv_nf) * v[t]);
      }
      part += __shfl_xor_sync(0xffffffffu, part, 1);
      part += __shfl_xor_sync(0xffffffffu, part, 2);
      part += __shfl_xor_sync(0xffffffffu, part, 4);
      if ((tid & 7) == 0 && part != 0.f) atomicAdd(J.dots + p, part);
    }
  }
  nb_grid_barrier(P.barrier, ncta * (++nbar));

  // ------------------------------ the four products ------------------------------
  for (int st = 0; st < 4; ++st) {
    int cur = -1;
    float inv_nf = 0.f;
    for (int u = blockIdx.x; u < P.total_units; u += ncta) {
      const int ji = job_of(u);
      const NbJob& J = P.job[ji];
      const int s = J.s;
      const int nkb = (s + NB_KC - 1) / NB_KC;
      if (ji != cur) {
        __syncthreads();                     // previous unit's readers of s_scale / a_row are done
        inv_nf = 1.f / (__ldcg(J.nf_src) + P.tiny);
        if (st == 0) {
          const int j = s_jc[ji] >= 0 ? s_jc[ji] : nb_block_argmax(J.row_sumsq, s, s_sv, s_si);
          for (int k = tid; k < nkb * NB_KC; k += NB_THREADS)
            a_row[k] = k < s ? __float2bfloat16_rn(__bfloat162float(J.A[(size_t)j * s + k]) * inv_nf) : __float2bfloat16_rn(0.f);
          if (tid < 32) { const float d = __ldcg(J.dots + tid); s_sgn[tid] = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f); }
          if (blockIdx.x == (unsigned)(J.unit0 % (int)ncta) && tid == 0) reinterpret_cast<int*>(J.scal)[SC_J] = j;
        }
        if (tid < 32) {
          float sc = inv_nf;
          if (st == 1) sc *= fminf(1.f / (sqrtf(__ldcg(J.rn1 + tid)) + P.tiny), 3.0e38f);
          if (st == 3) sc *= fminf(1.f / (sqrtf(__ldcg(J.rn3 + tid)) + P.tiny), 3.0e38f);
          s_scale[tid] = fminf(sc, 3.0e38f);
        }
        __syncthreads();
        cur = ji;
      }
      const bf16* Vsrc = st == 0 ? J.V0 : (st == 2 ? J.Vb : J.Va);
      bf16* Vdst = st == 1 ? J.Vb : J.Va;
      const int col0 = (u - J.unit0) * NB_W;

      // ---- warp-private pipeline over this warp's k slices: slice index ks = warp + 8 i, rows [16 ks, 16 ks + 16) ----
      const int nks = (s + NB_KS - 1) / NB_KS;
      const int my_n = nks > warp ? (nks - warp + 7) / 8 : 0;         // slices of this warp
      const uint32_t ring = smem_base + (uint32_t)warp * NB_STAGES * NB_STAGE_BYTES;
      uint8_t* ring_gen = nb_smem + warp * NB_STAGES * NB_STAGE_BYTES;
      // per-lane copy pattern (fixed for the unit): A slice = 16 rows x 8 chunks of 16 B -> 4 per lane; V slice = 32 probes x 2 chunks -> 2 per lane
      const int ac = lane & 7, ar = lane >> 3;                         // chunk, first row (rows ar + 4 i)
      const bool a_col_ok = col0 + ac * 8 < s;
      const bf16* a_src = J.A + (size_t)(warp * NB_KS + ar) * s + col0 + ac * 8;
      const bf16* v_src = Vsrc + (size_t)lane * s + warp * NB_KS;       // probe = lane
      const size_t a_step = (size_t)8 * NB_KS * s;                     // elements between consecutive slices of this warp
      uint32_t a_dst[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { const int r = ar + 4 * i; a_dst[i] = (uint32_t)(r * 128 + ((ac ^ (r & 7)) << 4)); }
      const uint32_t v_dst0 = NB_A_BYTES + lane * 32 + ((0 ^ ((lane >> 2) & 1)) << 4);
      const uint32_t v_dst1 = NB_A_BYTES + lane * 32 + ((1 ^ ((lane >> 2) & 1)) << 4);

      auto issue = [&](int it) {       // the it-th slice of this warp
        const uint32_t st_base = ring + (uint32_t)(it % NB_STAGES) * NB_STAGE_BYTES;
        const int k0 = (warp + 8 * it) * NB_KS;
        const bf16* ap = a_src + (size_t)it * a_step;
        const bf16* vp = v_src + (size_t)it * (8 * NB_KS);
        if (k0 + NB_KS <= s && a_col_ok) {
#pragma unroll
          for (int i = 0; i < 4; ++i) nb_cp_async16(st_base + a_dst[i], ap + (size_t)(4 * i) * s, 16);
        } else {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const bool ok = a_col_ok && k0 + ar + 4 * i < s;
            nb_cp_async16(st_base + a_dst[i], ok ? (const void*)(ap + (size_t)(4 * i) * s) : (const void*)J.A, ok ? 16 : 0);
          }
        }
        const bool ok0 = k0 < s, ok1 = k0 + 8 < s;
        nb_cp_async16(st_base + v_dst0, ok0 ? (const void*)vp : (const void*)Vsrc, ok0 ? 16 : 0);
        nb_cp_async16(st_base + v_dst1, ok1 ? (const void*)(vp + 8) : (const void*)Vsrc, ok1 ? 16 : 0);
      };

      // ldmatrix addresses (fixed): A-operand = V slice rows (probes) x k 16; B-operand = A slice k rows x 8 column chunks (transposed load)
      uint32_t v_ld[2], a_ld[4];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        const int row = mt * 16 + (lane & 15);
        v_ld[mt] = NB_A_BYTES + row * 32 + (((lane >> 4) ^ ((row >> 2) & 1)) << 4);
      }
#pragma unroll
      for (int np = 0; np < 4; ++np) {
        const int krow = (lane & 7) + ((lane >> 3) & 1) * 8;
        const int nc = np * 2 + (lane >> 4);
        a_ld[np] = krow * 128 + ((nc ^ (krow & 7)) << 4);
      }

      float acc[2][8][4];
#pragma unroll
      for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 8; ++b)
#pragma unroll
          for (int c = 0; c < 4; ++c) acc[a][b][c] = 0.f;

#pragma unroll 1
      for (int it = 0; it < NB_STAGES - 1; ++it) {
        if (it < my_n) issue(it);
        nb_cp_commit();
      }
#pragma unroll 1
      for (int it = 0; it < my_n; ++it) {
        nb_cp_wait<NB_STAGES - 2>();
        const uint32_t st_base = ring + (uint32_t)(it % NB_STAGES) * NB_STAGE_BYTES;
        if (st == 0) {
          // rotated probes V = A'[j] + sgn V0 (psgd.py:63), formed in place on the two chunks this lane itself copied (probe = lane)
          uint8_t* sg_base = ring_gen + (it % NB_STAGES) * NB_STAGE_BYTES;
          const int k0 = (warp + 8 * it) * NB_KS;
          const float sg = s_sgn[lane];
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            bf16* vp = reinterpret_cast<bf16*>(sg_base + (h ? v_dst1 : v_dst0));
            float v[8], a[8], o[8];
            ld8(vp, v);
            ld8(a_row + k0 + 8 * h, a);
#pragma unroll
            for (int t = 0; t < 8; ++t) o[t] = a[t] + sg * v[t];
            st8(vp, o);
          }
        }
        __syncwarp();
        // refill the slot consumed in the previous iteration (every lane of this warp is past its ldmatrix reads: __syncwarp above)
        if (it + NB_STAGES - 1 < my_n) issue(it + NB_STAGES - 1);
        nb_cp_commit();
        uint32_t af[2][4];
        nb_ldsm_x4(st_base + v_ld[0], af[0][0], af[0][1], af[0][2], af[0][3]);
        nb_ldsm_x4(st_base + v_ld[1], af[1][0], af[1][1], af[1][2], af[1][3]);
#pragma unroll
        for (int np = 0; np < 4; ++np) {
          uint32_t b0, b1, b2, b3;
          nb_ldsm_x4_t(st_base + a_ld[np], b0, b1, b2, b3);
#pragma unroll
          for (int mt = 0; mt < 2; ++mt) {
            nb_mma(acc[mt][2 * np], af[mt], b0, b1);
            nb_mma(acc[mt][2 * np + 1], af[mt], b2, b3);
          }
        }
      }
      nb_cp_wait<0>();
      __syncthreads();                       // every warp is done with its ring: the reduction buffer aliases the rings
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 8; ++nt)
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            const int p = mt * 16 + (lane >> 2) + hh * 8;
            const int n = nt * 8 + (lane & 3) * 2;
            *reinterpret_cast<float2*>(&red[(warp * 32 + p) * NB_RED_LD + n]) = make_float2(acc[mt][nt][2 * hh], acc[mt][nt][2 * hh + 1]);
          }
      __syncthreads();
      {
        const int p = tid >> 3, n0 = (tid & 7) * 8;
        float o[8];
#pragma unroll
        for (int t = 0; t < 8; ++t) o[t] = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) {
          const float4 x = *reinterpret_cast<const float4*>(&red[(w * 32 + p) * NB_RED_LD + n0]);
          const float4 y = *reinterpret_cast<const float4*>(&red[(w * 32 + p) * NB_RED_LD + n0 + 4]);
          o[0] += x.x; o[1] += x.y; o[2] += x.z; o[3] += x.w; o[4] += y.x; o[5] += y.y; o[6] += y.z; o[7] += y.w;
        }
        const float sc = s_scale[p];
        float ss = 0.f;
#pragma unroll
        for (int t = 0; t < 8; ++t) { o[t] = rbf(o[t] * sc); ss = fmaf(o[t], o[t], ss); }
        const bool ok = col0 + n0 < s;
        if (ok && st < 3) st8(Vdst + (size_t)p * s + col0 + n0, o);
        if (!ok) ss = 0.f;
        ss += __shfl_xor_sync(0xffffffffu, ss, 1);
        ss += __shfl_xor_sync(0xffffffffu, ss, 2);
        ss += __shfl_xor_sync(0xffffffffu, ss, 4);
        float* rn = st == 0 ? J.rn1 : (st == 2 ? J.rn3 : (st == 3 ? J.rn4 : nullptr));
        if (rn && (tid & 7) == 0) atomicAdd(rn + p, ss);
      }
      __syncthreads();                       // the next unit's copies overwrite the reduction buffer
    }
    if (st < 3) nb_grid_barrier(P.barrier, ncta * (++nbar));
  }

  // ------------------------------ finish: the CTA that arrives last ------------------------------
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    const unsigned old = atomicAdd(P.done, 1u);
    s_last = (old == ncta - 1) ? 1 : 0;
  }
  __syncthreads();
  if (s_last) {
    __threadfence();
    if (tid < P.njobs) nb_finish_job(P.job[tid], P.dtype, P.tiny);
    __syncthreads();
    if (tid == 0) { *P.barrier = 0u; *P.done = 0u; __threadfence(); }
  }
}


// =====================================================================================================================================
// tcgen05 form of the same kernel (matrices whose size is a multiple of 64).  mma.sync m16n8k16 turned out to issue only once per ~32
// cycles per SM sub-partition on this part (profiles/r02_ncu_bounds_hmma.txt: the mma.sync kernel above is bound by exactly that: 4096 HMMA
// per unit and step = 33 k cycles), so the products move to the 5th-generation tensor cores:
//   unit = 128 columns of A:  D[128 x 32] (TMEM, fp32) = A[:, slab]^T (MN-major operand, 3-D TMA box) x V^T (K-major operand, 32 rows)
//   warp 0: TMA producer (8-stage mbarrier ring: A tile 16 KB + V tile 4 KB per 64-row k block), warp 1: single-thread tcgen05.mma issuer
//   (M = 128, N = 32, K = 16 x 4 per block), warps 2-5: epilogue (tcgen05.ld: thread = column of A, registers = probes -> scale, round, per-probe
//   sums of squares by a warp transpose-reduce, coalesced 2-byte stores V_new[p][col]).
// Phase I writes the rotated probes V = A'[j] + sgn V0 to global memory (one more grid barrier than the mma.sync form, whose probes
// are rotated in shared memory) so that step 0 is fed by TMA like the other steps.
// =====================================================================================================================================
constexpr int NT_BM = 128;
constexpr int NT_BK = 256;      // k rows per stage: the issue loops of the single producer / MMA threads cost ~500 cycles per stage whatever its
constexpr int NT_STAGES = 2;    // size (dependent-issue latency of one warp), so stages are big: 80 KB each (64-row stages: 62 GB/s per SM)
constexpr int NT_THREADS = 192;
constexpr int NT_MAX_JOBS = 8;
constexpr int NT_A_BYTES = NT_BM * NT_BK * 2;
constexpr int NT_V_BYTES = 32 * NT_BK * 2;
constexpr int NT_STAGE_BYTES = NT_A_BYTES + NT_V_BYTES;
constexpr int NT_SMEM_BYTES = NT_STAGES * NT_STAGE_BYTES + 1024 + 256;

struct alignas(64) NbTcJob {
  CUtensorMap map_a;    // A as MN-major operand: 3-D {64, s, s / 64}, box {64, NT_BK, 2}
  CUtensorMap map_va;   // Va / Vb as K-major operand: 2-D {s, 32}, box {64, 32} (NT_BK / 64 boxes per stage)
  CUtensorMap map_vb;
  NbJob j;              // unit0 / nunits count 128-column units here
};
struct alignas(64) NbTcParams {
  NbTcJob job[NT_MAX_JOBS];
  int njobs, total_units, dtype;
  float tiny;
  unsigned* barrier;
  unsigned* done;
  int mn_lbo, mn_sbo;
};

__device__ __forceinline__ void nb_fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }

__global__ void __launch_bounds__(NT_THREADS, 1) k_norm_bounds_tc(const __grid_constant__ NbTcParams P) {
  extern __shared__ uint8_t nt_smem_raw[];
  const uint32_t smem_base = (smem_u32(nt_smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = nt_smem_raw + (smem_base - smem_u32(nt_smem_raw));
  const uint32_t bar_base = smem_base + NT_STAGES * NT_STAGE_BYTES;
  auto full_bar = [&](int st_) { return bar_base + 8u * st_; };
  auto empty_bar = [&](int st_) { return bar_base + 8u * (NT_STAGES + st_); };
  const uint32_t tfull_bar = bar_base + 8u * (2 * NT_STAGES);
  const uint32_t tmem_slot = bar_base + 8u * (2 * NT_STAGES + 1);
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + NT_STAGES * NT_STAGE_BYTES + 8 * (2 * NT_STAGES + 1));
  __shared__ float s_scale[32];
  __shared__ float s_dots[32];
  __shared__ float s_sv[8];
  __shared__ int s_si[9];
  __shared__ int s_last;
  __shared__ int s_jc[NT_MAX_JOBS];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const unsigned ncta = gridDim.x;
  unsigned nbar = 0;

  if (tid == 0) {
    for (int i = 0; i < NT_STAGES; ++i) { mbar_init(full_bar(i), 1); mbar_init(empty_bar(i), 1); }
    mbar_init(tfull_bar, 1);
    fence_barrier_init();
    for (int i = 0; i < P.njobs; ++i) { prefetch_tmap(&P.job[i].map_a); prefetch_tmap(&P.job[i].map_va); prefetch_tmap(&P.job[i].map_vb); }
  }
  if (tid < NT_MAX_JOBS) s_jc[tid] = -1;
  if (warp == 1) tmem_alloc(tmem_slot, 32);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;

  auto job_of = [&](int u) {
    int ji = 0;
    for (int i = 1; i < P.njobs; ++i)
      if (u >= P.job[i].j.unit0) ji = i;
    return ji;
  };
  // block-wide argmax of the squared row norms (first maximal index), any block size up to 8 warps
  auto block_argmax = [&](const float* x, int s) {
    float best = -1.f;
    int bi = 0x7fffffff;
    for (int i = tid; i < s; i += NT_THREADS) {
      const float v = __ldcg(x + i);
      if (v > best) { best = v; bi = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    __syncthreads();
    if (lane == 0) { s_sv[warp] = best; s_si[warp] = bi; }
    __syncthreads();
    if (tid == 0) {
      for (int w = 1; w < NT_THREADS / 32; ++w)
        if (s_sv[w] > best || (s_sv[w] == best && s_si[w] < bi)) { best = s_sv[w]; bi = s_si[w]; }
      s_si[8] = (bi == 0x7fffffff) ? 0 : bi;
    }
    __syncthreads();
    return s_si[8];
  };

  // ------------------------------ phase I-a: signs of the probe rotation (psgd.py:63) ------------------------------
  for (int u = blockIdx.x; u < P.total_units; u += ncta) {
    const int ji = job_of(u);
    const NbJob& J = P.job[ji].j;
    int j = s_jc[ji];
    if (j < 0) { j = block_argmax(J.row_sumsq, J.s); if (tid == 0) s_jc[ji] = j; }
    const float inv_nf = 1.f / (__ldcg(J.nf_src) + P.tiny);
    if (tid < 32) s_dots[tid] = 0.f;
    __syncthreads();
    for (int idx = tid; idx < 32 * 16; idx += NT_THREADS) {
      const int p = idx >> 4, c0 = (u - J.unit0) * NT_BM + (idx & 15) * 8;
      if (c0 < J.s) {
        float a[8], v[8], part = 0.f;
        ld8(J.A + (size_t)j * J.s + c0, a);
        ld8(J.V0 + (size_t)p * J.s + c0, v);
#pragma unroll
        for (int t = 0; t < 8; ++t) part += rbf(rbf(a[t] * inv_nf) * v[t]);
        atomicAdd(&s_dots[p], part);
      }
    }
    __syncthreads();
    if (tid < 32 && s_dots[tid] != 0.f) atomicAdd(J.dots + tid, s_dots[tid]);
    __syncthreads();
  }
  nb_grid_barrier(P.barrier, ncta * (++nbar));
  // ------------------------------ phase I-b: V = A'[j] + sgn V0 -> Va ------------------------------
  for (int u = blockIdx.x; u < P.total_units; u += ncta) {
    const int ji = job_of(u);
    const NbJob& J = P.job[ji].j;
    const int j = s_jc[ji];
    const float inv_nf = 1.f / (__ldcg(J.nf_src) + P.tiny);
    if (u == J.unit0 && tid == 0) reinterpret_cast<int*>(J.scal)[SC_J] = j;
    for (int idx = tid; idx < 32 * 16; idx += NT_THREADS) {
      const int p = idx >> 4, c0 = (u - J.unit0) * NT_BM + (idx & 15) * 8;
      if (c0 < J.s) {
        const float d = __ldcg(J.dots + p);
        const float sg = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
        float a[8], v[8], o[8];
        ld8(J.A + (size_t)j * J.s + c0, a);
        ld8(J.V0 + (size_t)p * J.s + c0, v);
#pragma unroll
        for (int t = 0; t < 8; ++t) o[t] = rbf(a[t] * inv_nf) + sg * v[t];
        st8(J.Va + (size_t)p * J.s + c0, o);
      }
    }
  }
  nb_fence_proxy_async();        // generic-proxy stores of Va -> async-proxy (TMA) reads by every CTA after the barrier
  nb_grid_barrier(P.barrier, ncta * (++nbar));
  if (warp == 0) nb_fence_proxy_async();

  // ------------------------------ the four products ------------------------------
  int p_stage = 0; uint32_t p_phase = 0;     // producer's ring position
  int c_stage = 0; uint32_t c_phase = 0;     // MMA issuer's ring position
  uint32_t t_phase = 0;                      // accumulator-full barrier phase (every warp counts the units)
  for (int st = 0; st < 4; ++st) {
    for (int u = blockIdx.x; u < P.total_units; u += ncta) {
      const int ji = job_of(u);
      const NbTcJob& TJ = P.job[ji];
      const NbJob& J = TJ.j;
      const int s = J.s;
      const int nkb = (s + NT_BK - 1) / NT_BK;      // the last block may be partial: TMA zero-fills rows / columns beyond s
      const int slab = u - J.unit0;
      if (tid < 32) {
        float sc = 1.f / (__ldcg(J.nf_src) + P.tiny);
        if (st == 1) sc *= fminf(1.f / (sqrtf(__ldcg(J.rn1 + tid)) + P.tiny), 3.0e38f);
        if (st == 3) sc *= fminf(1.f / (sqrtf(__ldcg(J.rn3 + tid)) + P.tiny), 3.0e38f);
        s_scale[tid] = fminf(sc, 3.0e38f);
      }
      tc_fence_before();
      __syncthreads();
      tc_fence_after();
      if (warp == 0) {
        // ===================== TMA producer =====================
        const CUtensorMap* mv = (st & 1) ? &TJ.map_vb : &TJ.map_va;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(empty_bar(p_stage), p_phase ^ 1u, nullptr);
          if (elect_one()) {
            const uint32_t sa = smem_base + p_stage * NT_STAGE_BYTES;
            mbar_arrive_expect_tx(full_bar(p_stage), NT_STAGE_BYTES);
            tma_load_3d(&TJ.map_a, full_bar(p_stage), sa, 0, kb * NT_BK, slab * (NT_BM / 64));
#pragma unroll
            for (int c = 0; c < NT_BK / 64; ++c) tma_load_2d(mv, full_bar(p_stage), sa + NT_A_BYTES + c * 4096, kb * NT_BK + c * 64, 0);
          }
          __syncwarp();
          if (++p_stage == NT_STAGES) { p_stage = 0; p_phase ^= 1u; }
        }
      } else if (warp == 1) {
        // ===================== MMA issuer =====================
        // instruction descriptor: D = f32 (bit 4), A = bf16 (bit 7), B = bf16 (bit 10), A MN-major (bit 15), N >> 3 at bits 17-22, M >> 4 at 24-28
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (uint32_t(32 >> 3) << 17) | (uint32_t(NT_BM >> 4) << 24);
        // descriptors of stage 0, k step 0; the others differ in the 14-bit start-address field only (16-byte units, no carry: smem < 256 KB).
        // MN-major A tile = [2 chunks][NT_BK k rows][128 B]: LBO = chunk stride, SBO = 8 k rows; K-major V tile = NT_BK / 64 boxes of [32 rows][128 B]
        const uint64_t adesc0 = make_smem_desc(smem_base, (uint32_t)(NT_BK * 128), (uint32_t)P.mn_sbo);
        const uint64_t bdesc0 = make_smem_desc(smem_base + NT_A_BYTES, 0u, 1024u);
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(full_bar(c_stage), c_phase, nullptr);
          tc_fence_after();
          if (elect_one()) {
            const uint64_t so = (uint64_t)((c_stage * NT_STAGE_BYTES) >> 4);
#pragma unroll
            for (int k = 0; k < NT_BK / 16; ++k) {
              const uint64_t adesc = adesc0 + so + (uint64_t)((k * 16 * 128) >> 4);
              const uint64_t bdesc = bdesc0 + so + (uint64_t)(((k >> 2) * 4096 + (k & 3) * 32) >> 4);
              umma_bf16(tmem_base, adesc, bdesc, idesc, (kb | k) != 0 ? 1u : 0u);
            }
            umma_commit(empty_bar(c_stage));
            if (kb + 1 == nkb) umma_commit(tfull_bar);
          }
          __syncwarp();
          if (++c_stage == NT_STAGES) { c_stage = 0; c_phase ^= 1u; }
        }
      } else {
        // ===================== epilogue: warp w owns TMEM lanes 32 (w % 4) .. + 31 = columns of A =====================
        const int quarter = warp & 3;
        mbar_wait_relaxed(tfull_bar, t_phase, nullptr);
        tc_fence_after();
        uint32_t raw[32];
        tmem_ld_32x32(tmem_base + (uint32_t(quarter * 32) << 16), raw);
        tmem_ld_wait();
        const int col = slab * NT_BM + quarter * 32 + lane;
        const bool ok = col < s;
        float v[32];
#pragma unroll
        for (int jx = 0; jx < 32; ++jx) v[jx] = ok ? rbf(__uint_as_float(raw[jx]) * s_scale[jx]) : 0.f;
        if (st < 3 && ok) {
          bf16* dst = ((st & 1) ? J.Va : J.Vb) + col;
#pragma unroll
          for (int jx = 0; jx < 32; ++jx) dst[(size_t)jx * s] = __float2bfloat16_rn(v[jx]);
        }
        float* rn = st == 0 ? J.rn1 : (st == 2 ? J.rn3 : (st == 3 ? J.rn4 : nullptr));
        if (rn) {
          // transpose-reduce: after the 5 halving steps lane l holds the sum over the warp's 32 columns of probe l
#pragma unroll
          for (int jx = 0; jx < 32; ++jx) v[jx] = v[jx] * v[jx];
#define NB_TR_STEP(OFF)                                                \
          {                                                            \
            const bool upper = (lane & (OFF)) != 0;                    \
            _Pragma("unroll") for (int i = 0; i < (OFF); ++i) {        \
              const float send = upper ? v[i] : v[i + (OFF)];          \
              const float keep = upper ? v[i + (OFF)] : v[i];          \
              v[i] = keep + __shfl_xor_sync(0xffffffffu, send, (OFF)); \
            }                                                          \
          }
          NB_TR_STEP(16) NB_TR_STEP(8) NB_TR_STEP(4) NB_TR_STEP(2) NB_TR_STEP(1)
#undef NB_TR_STEP
          if (v[0] != 0.f) atomicAdd(rn + lane, v[0]);
        }
        if (st < 3) nb_fence_proxy_async();   // this thread's stores of V_new are read through TMA (async proxy) after the grid barrier
      }
      t_phase ^= 1u;
      tc_fence_before();
      __syncthreads();       // the accumulator is drained and every role is done with this unit
    }
    if (st < 3) {
      nb_grid_barrier(P.barrier, ncta * (++nbar));
      if (warp == 0) nb_fence_proxy_async();
    }
  }

  // ------------------------------ finish: the CTA that arrives last ------------------------------
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 32); }
  if (tid == 0) {
    __threadfence();
    const unsigned old = atomicAdd(P.done, 1u);
    s_last = (old == ncta - 1) ? 1 : 0;
  }
  __syncthreads();
  if (s_last) {
    __threadfence();
    if (tid < P.njobs) nb_finish_job(P.job[tid].j, P.dtype, P.tiny);
    __syncthreads();
    if (tid == 0) { *P.barrier = 0u; *P.done = 0u; __threadfence(); }
  }
}

}  // namespace psgd
