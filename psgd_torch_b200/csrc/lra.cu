// lra.cu -- LRA preconditioner Q = (I + U V^T) diag(d)   (psgd.py:987-1072), HBM-streaming formulation.
//
// The reference makes ~25 separate passes over the n x r factors.  Here one update is
//   sweep 1 (read U,V,h,v,d):  r x r Grams U^T U, V^T V, V^T U and the r-vectors U^T x, V^T x for x in {d.h, v/d}
//   small   (one CTA, fp32):   everything of size r -- balancing rotation, LU solves, Lipschitz constants, step --
//                              follows algebraically from the sweep-1 products
//   sweep 2 (read+write U,V):  balancing rotation U<-U Au, V<-V Av fused with the rank-2 update of U or V and with
//                              the per-row quantities of the d update
//   d pass  (n-vector only)
// and the apply is three row sweeps (V, U, V).  fp32 arithmetic throughout; storage dtype bf16 or fp32.
#include <stdio.h>
#include <string.h>

#include "common.cuh"
#include "lra_mma.cuh"
#include "lra_tc.cuh"

namespace psgd {

// parameter / accumulator block layouts: see lra_mma.cuh

template <typename T, int RP>
__device__ __forceinline__ void load_row(const T* __restrict__ base, long long row, int r, float* x) {
  const T* p = base + row * (long long)r;
  if (r == RP && (RP * sizeof(T)) % 16 == 0) {
    constexpr int NV = (RP * sizeof(T)) / 16;
    constexpr int PER = 16 / sizeof(T);
#pragma unroll
    for (int q = 0; q < NV; ++q) {
      uint4 u = reinterpret_cast<const uint4*>(p)[q];
      const T* e = reinterpret_cast<const T*>(&u);
#pragma unroll
      for (int t = 0; t < PER; ++t) x[q * PER + t] = to_f<T>(e[t]);
    }
  } else {
#pragma unroll
    for (int c = 0; c < RP; ++c) x[c] = (c < r) ? to_f<T>(p[c]) : 0.f;
  }
}
template <typename T, int RP>
__device__ __forceinline__ void store_row(T* __restrict__ base, long long row, int r, const float* x) {
  T* p = base + row * (long long)r;
  if (r == RP && (RP * sizeof(T)) % 16 == 0) {
    constexpr int NV = (RP * sizeof(T)) / 16;
    constexpr int PER = 16 / sizeof(T);
#pragma unroll
    for (int q = 0; q < NV; ++q) {
      uint4 u;
      T* e = reinterpret_cast<T*>(&u);
#pragma unroll
      for (int t = 0; t < PER; ++t) e[t] = from_f<T>(x[q * PER + t]);
      reinterpret_cast<uint4*>(p)[q] = u;
    }
  } else {
#pragma unroll
    for (int c = 0; c < RP; ++c)
      if (c < r) p[c] = from_f<T>(x[c]);
  }
}

// ------------------------------------------------------------------------------------------------
// sweep 1: Grams and projections.  256 threads; tiles of 64 rows staged in smem as fp32; each thread owns 4x4
// blocks of the three RP x RP Grams (register tiling) and, for tid < 4*RP, one entry of the projection vectors.
// ------------------------------------------------------------------------------------------------
template <typename T, int RP>
__global__ void __launch_bounds__(256) k_lra_sweep1(const T* __restrict__ U, const T* __restrict__ V, const T* __restrict__ d,
                                                    const T* __restrict__ hvec, const T* __restrict__ vvec, long long n, int r,
                                                    float* __restrict__ acc_out) {
  constexpr int TR = 64;
  constexpr int LD = RP + 4;  // keeps float4 alignment, skews banks
  __shared__ __align__(16) float Us[TR][LD];
  __shared__ __align__(16) float Vs[TR][LD];
  __shared__ float x1s[TR], x2s[TR];
  constexpr int NT = RP / 4;                 // 4x4 tiles per side
  constexpr int TILES = 3 * NT * NT;         // over the three Grams
  constexpr int PER = (TILES + 255) / 256;   // tiles per thread
  float acc[PER][16];
#pragma unroll
  for (int t = 0; t < PER; ++t)
#pragma unroll
    for (int e = 0; e < 16; ++e) acc[t][e] = 0.f;
  float pv = 0.f;      // projection entry
  float sq1 = 0.f, sq2 = 0.f;
  const int tid = threadIdx.x;
  const long long ntiles = (n + TR - 1) / TR;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long row0 = tile * TR;
    __syncthreads();
    // stage: thread-per-row loads (64 rows -> threads 0..63 load U, 64..127 load V, 128..191 the vectors)
    if (tid < TR) {
      long long row = row0 + tid;
      float x[RP];
      if (row < n) load_row<T, RP>(U, row, r, x); else {
#pragma unroll
        for (int c = 0; c < RP; ++c) x[c] = 0.f;
      }
#pragma unroll
      for (int c = 0; c < RP; ++c) Us[tid][c] = x[c];
    } else if (tid < 2 * TR) {
      int rr = tid - TR;
      long long row = row0 + rr;
      float x[RP];
      if (row < n) load_row<T, RP>(V, row, r, x); else {
#pragma unroll
        for (int c = 0; c < RP; ++c) x[c] = 0.f;
      }
#pragma unroll
      for (int c = 0; c < RP; ++c) Vs[rr][c] = x[c];
    } else if (tid < 3 * TR) {
      int rr = tid - 2 * TR;
      long long row = row0 + rr;
      float a = 0.f, b = 0.f;
      if (row < n) {
        float dd = to_f<T>(d[row]);
        a = to_f<T>(from_f<T>(dd * to_f<T>(hvec[row])));   // d*h   psgd.py:1017
        b = to_f<T>(from_f<T>(to_f<T>(vvec[row]) / dd));   // v/d   psgd.py:1022
      }
      x1s[rr] = a; x2s[rr] = b;
      sq1 += a * a; sq2 += b * b;
    }
    __syncthreads();
#pragma unroll
    for (int t = 0; t < PER; ++t) {
      const int tl = tid + t * 256;
      if (tl < TILES) {
        const int which = tl / (NT * NT);
        const int ij = tl - which * NT * NT;
        const int ti = ij / NT, tj = ij - ti * NT;
        const float(*A)[LD] = (which == 0) ? Us : Vs;   // UtU: U,U ; VtV: V,V ; VtU: V,U
        const float(*B)[LD] = (which == 1) ? Vs : Us;
#pragma unroll 4
        for (int rr = 0; rr < TR; ++rr) {
          const float4 a4 = *reinterpret_cast<const float4*>(&A[rr][ti * 4]);
          const float4 b4 = *reinterpret_cast<const float4*>(&B[rr][tj * 4]);
          const float av[4] = {a4.x, a4.y, a4.z, a4.w};
          const float bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[t][i * 4 + j] = fmaf(av[i], bv[j], acc[t][i * 4 + j]);
        }
      }
    }
    if (tid < 4 * RP) {
      const int which = tid / RP, c = tid - which * RP;   // 0: U^T x1, 1: V^T x1, 2: U^T x2, 3: V^T x2
      const float(*A)[LD] = (which & 1) ? Vs : Us;
      const float* xs = (which & 2) ? x2s : x1s;
      float s = 0.f;
#pragma unroll 8
      for (int rr = 0; rr < TR; ++rr) s = fmaf(A[rr][c], xs[rr], s);
      pv += s;
    }
  }
  // flush
#pragma unroll
  for (int t = 0; t < PER; ++t) {
    const int tl = tid + t * 256;
    if (tl < TILES) {
      const int which = tl / (NT * NT);
      const int ij = tl - which * NT * NT;
      const int ti = ij / NT, tj = ij - ti * NT;
      float* G = acc_out + (size_t)which * RP * RP;
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) atomicAdd(&G[(ti * 4 + i) * RP + tj * 4 + j], acc[t][i * 4 + j]);
    }
  }
  if (tid < 4 * RP) atomicAdd(&acc_out[(size_t)3 * RP * RP + tid], pv);
  if (tid >= 2 * 64 && tid < 3 * 64) {
    float a = warp_sum(sq1), b = warp_sum(sq2);
    if ((tid & 31) == 0) { atomicAdd(&acc_out[(size_t)3 * RP * RP + 4 * RP], a); atomicAdd(&acc_out[(size_t)3 * RP * RP + 4 * RP + 1], b); }
  }
}

// sweep 1 for ranks <= 4 (BASELINE configs[4] sweeps r = 4): the three 4 x 4 Grams, the four projections and the two norms are 66 sums --
// they fit in one thread's registers, so every thread streams whole rows (8 or 16 bytes each, coalesced across the warp) and the block
// reduces once at the end.  The tiled kernel above leaves 253 of its 256 threads idle at this rank (3 tiles of 4 x 4): 2.76 ms per 2^24
// rows against 0.37 GB of compulsory traffic.
template <typename T>
__global__ void __launch_bounds__(256) k_lra_sweep1_r4(const T* __restrict__ U, const T* __restrict__ V, const T* __restrict__ d,
                                                       const T* __restrict__ hvec, const T* __restrict__ vvec, long long n, int r,
                                                       float* __restrict__ acc_out) {
  constexpr int RP = 4, NS = 3 * RP * RP + 4 * RP + 2;
  __shared__ float red[NS];
  float a[NS];
#pragma unroll
  for (int e = 0; e < NS; ++e) a[e] = 0.f;
  for (int e = threadIdx.x; e < NS; e += blockDim.x) red[e] = 0.f;
  __syncthreads();
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long row = blockIdx.x * (long long)blockDim.x + threadIdx.x; row < n; row += stride) {
    float u[RP], w[RP];
    if (r == RP && sizeof(T) == 2) {   // 8-byte rows
      const uint2 pu = *reinterpret_cast<const uint2*>(U + row * RP), pw = *reinterpret_cast<const uint2*>(V + row * RP);
      const T* eu = reinterpret_cast<const T*>(&pu);
      const T* ew = reinterpret_cast<const T*>(&pw);
#pragma unroll
      for (int c = 0; c < RP; ++c) { u[c] = to_f<T>(eu[c]); w[c] = to_f<T>(ew[c]); }
    } else {
      load_row<T, RP>(U, row, r, u);
      load_row<T, RP>(V, row, r, w);
    }
    const float dd = to_f<T>(d[row]);
    const float x1 = to_f<T>(from_f<T>(dd * to_f<T>(hvec[row])));   // d*h   psgd.py:1017
    const float x2 = to_f<T>(from_f<T>(to_f<T>(vvec[row]) / dd));   // v/d   psgd.py:1022
#pragma unroll
    for (int i = 0; i < RP; ++i) {
#pragma unroll
      for (int j = 0; j < RP; ++j) {
        a[i * RP + j] = fmaf(u[i], u[j], a[i * RP + j]);                       // U^T U
        a[RP * RP + i * RP + j] = fmaf(w[i], w[j], a[RP * RP + i * RP + j]);   // V^T V
        a[2 * RP * RP + i * RP + j] = fmaf(w[i], u[j], a[2 * RP * RP + i * RP + j]);   // V^T U
      }
      a[3 * RP * RP + i] = fmaf(u[i], x1, a[3 * RP * RP + i]);                 // U^T x1, V^T x1, U^T x2, V^T x2
      a[3 * RP * RP + RP + i] = fmaf(w[i], x1, a[3 * RP * RP + RP + i]);
      a[3 * RP * RP + 2 * RP + i] = fmaf(u[i], x2, a[3 * RP * RP + 2 * RP + i]);
      a[3 * RP * RP + 3 * RP + i] = fmaf(w[i], x2, a[3 * RP * RP + 3 * RP + i]);
    }
    a[NS - 2] = fmaf(x1, x1, a[NS - 2]);
    a[NS - 1] = fmaf(x2, x2, a[NS - 1]);
  }
#pragma unroll
  for (int e = 0; e < NS; ++e) {
    const float v = warp_sum(a[e]);
    if ((threadIdx.x & 31) == 0) atomicAdd(&red[e], v);
  }
  __syncthreads();
  for (int e = threadIdx.x; e < NS; e += blockDim.x) atomicAdd(&acc_out[e], red[e]);
}

// ------------------------------------------------------------------------------------------------
// small kernel: one CTA, all r x r / r-vector algebra in fp32 in shared memory  (psgd.py:1006-1052)
// ------------------------------------------------------------------------------------------------
__device__ inline void mm_small(const float* A, const float* B, float* Cm, int r, int RP, bool ta, bool tb) {
  // C = op(A) op(B), all RP-strided r x r, block-cooperative
  for (int e = threadIdx.x; e < r * r; e += blockDim.x) {
    int i = e / r, j = e - i * r;
    float s = 0.f;
    for (int k = 0; k < r; ++k) s = fmaf(ta ? A[k * RP + i] : A[i * RP + k], tb ? B[j * RP + k] : B[k * RP + j], s);
    Cm[i * RP + j] = s;
  }
  __syncthreads();
}
__device__ inline void mv_small(const float* A, const float* x, float* y, int r, int RP, bool ta) {
  for (int i = threadIdx.x; i < r; i += blockDim.x) {
    float s = 0.f;
    for (int k = 0; k < r; ++k) s = fmaf(ta ? A[k * RP + i] : A[i * RP + k], x[k], s);
    y[i] = s;
  }
  __syncthreads();
}
__device__ inline float dot_small(const float* a, const float* b, int r) {  // every thread computes it (r <= 64)
  float s = 0.f;
  for (int k = 0; k < r; ++k) s = fmaf(a[k], b[k], s);
  return s;
}

// grid 1, block 256. dyn smem: 8 RPxRP matrices + 24 vectors
__global__ void k_lra_small(const float* __restrict__ acc, float* __restrict__ par, int r, int RP, float lr, float betaL, int update_U,
                            float* Lu, float* Lv, int dtype) {
  extern __shared__ float sm[];
  const int MM = RP * RP;
  float* Guu = sm; float* Gvv = Guu + MM; float* Gvu = Gvv + MM; float* E = Gvu + MM; float* Au = E + MM; float* Av = Au + MM;
  float* T1 = Av + MM; float* Mx = T1 + MM;   // Mx = I + V'^T U'
  float* vec = Mx + MM;
  float* Utx1 = vec; float* Vtx1 = vec + RP; float* Utx2 = vec + 2 * RP; float* Vtx2 = vec + 3 * RP;
  float* c1 = vec + 4 * RP; float* c2 = vec + 5 * RP; float* s1 = vec + 6 * RP; float* s2 = vec + 7 * RP;
  float* t = vec + 8 * RP; float* atX = vec + 9 * RP; float* btX = vec + 10 * RP; float* tmp = vec + 11 * RP;
  float* tmp2 = vec + 12 * RP; float* upx1 = vec + 13 * RP; float* upx2 = vec + 14 * RP; float* vpx2 = vec + 15 * RP;
  __shared__ int piv[64];
  __shared__ float LUm[64 * 64];
  const int tid = threadIdx.x;
  for (int e = tid; e < 3 * MM; e += blockDim.x) sm[e] = acc[e];
  for (int e = tid; e < 4 * RP; e += blockDim.x) vec[e] = acc[3 * MM + e];
  __syncthreads();
  const float x1sq = acc[3 * MM + 4 * RP], x2sq = acc[3 * MM + 4 * RP + 1];
  // --- balancing (psgd.py:1006-1015) ---
  float trU = 0.f, trV = 0.f;
  for (int k = 0; k < r; ++k) { trU += Guu[k * RP + k]; trV += Gvv[k * RP + k]; }
  const float rho = sqrtf(sqrtf(trU / trV));
  const float rho2 = rho * rho;
  const float den = trU / rho2 + trV * rho2;
  for (int e = tid; e < r * r; e += blockDim.x) {
    int i = e / r, j = e - i * r;
    E[i * RP + j] = 0.1f * (Guu[i * RP + j] / rho2 - Gvv[i * RP + j] * rho2) / den;
  }
  __syncthreads();
  mm_small(E, E, T1, r, RP, false, false);  // T1 = E E
  for (int e = tid; e < r * r; e += blockDim.x) {
    int i = e / r, j = e - i * r;
    float idn = (i == j) ? 1.f : 0.f;
    float e1 = E[i * RP + j], e2 = 0.5f * T1[i * RP + j];
    Au[i * RP + j] = (idn - e1 + e2) / rho;   // U' = (U/rho)(I - E + E2)
    Av[i * RP + j] = (idn + e1 + e2) * rho;   // V' = (V rho)(I + E + E2)
  }
  {  // tensor-core sweep 2 applies the identity part exactly and only the small corrections through bf16 (transposed: N x K)
    bf16* EuT = reinterpret_cast<bf16*>(par + lra_par_et_off(RP));
    bf16* EvT = EuT + RP * RP;
    for (int e = tid; e < RP * RP; e += blockDim.x) {
      int nn = e / RP, kk = e - nn * RP;
      bool in = nn < r && kk < r;
      float e1 = in ? E[kk * RP + nn] : 0.f, e2 = in ? 0.5f * T1[kk * RP + nn] : 0.f;
      EuT[e] = __float2bfloat16_rn(e1 - e2);
      EvT[e] = __float2bfloat16_rn(e1 + e2);
    }
    if (tid == 0) { float* ps = par + lra_par_scal_off(RP); ps[LS_INV_RHO] = 1.f / rho; ps[LS_RHO] = rho; }
  }
  __syncthreads();
  // balanced Grams: G'uu = Au^T Guu Au etc. (reuse E as scratch)
  mm_small(Guu, Au, T1, r, RP, false, false); mm_small(Au, T1, E, r, RP, true, false);
  for (int e = tid; e < r * r; e += blockDim.x) { int i = e / r, j = e - i * r; Guu[i * RP + j] = E[i * RP + j]; }
  __syncthreads();
  mm_small(Gvv, Av, T1, r, RP, false, false); mm_small(Av, T1, E, r, RP, true, false);
  for (int e = tid; e < r * r; e += blockDim.x) { int i = e / r, j = e - i * r; Gvv[i * RP + j] = E[i * RP + j]; }
  __syncthreads();
  mm_small(Gvu, Au, T1, r, RP, false, false); mm_small(Av, T1, E, r, RP, true, false);
  for (int e = tid; e < r * r; e += blockDim.x) {
    int i = e / r, j = e - i * r;
    Gvu[i * RP + j] = E[i * RP + j];
    Mx[i * RP + j] = E[i * RP + j] + ((i == j) ? 1.f : 0.f);   // IpVtU  psgd.py:1020-1021
  }
  __syncthreads();
  // projections of the balanced factors
  mv_small(Au, Utx1, upx1, r, RP, true);   // U'^T x1
  mv_small(Av, Vtx1, c1, r, RP, true);     // c1 = V'^T x1          (Qh = x1 + U' c1)
  mv_small(Au, Utx2, upx2, r, RP, true);   // U'^T x2
  mv_small(Av, Vtx2, vpx2, r, RP, true);   // V'^T x2
  mv_small(Guu, c1, tmp, r, RP, false);
  for (int i = tid; i < r; i += blockDim.x) c2[i] = upx1[i] + tmp[i];   // c2 = U'^T Qh   (Ph = d (Qh + V' c2))
  __syncthreads();
  // --- LU with partial pivoting of Mx (fp32; psgd.py:1023) ---
  for (int e = tid; e < r * r; e += blockDim.x) { int i = e / r, j = e - i * r; LUm[i * 64 + j] = Mx[i * RP + j]; }
  __syncthreads();
  for (int k = 0; k < r; ++k) {
    if (tid == 0) {
      int p = k; float best = fabsf(LUm[k * 64 + k]);
      for (int i = k + 1; i < r; ++i) { float v = fabsf(LUm[i * 64 + k]); if (v > best) { best = v; p = i; } }
      piv[k] = p;
    }
    __syncthreads();
    const int p = piv[k];
    if (p != k) for (int j = tid; j < r; j += blockDim.x) { float a = LUm[k * 64 + j]; LUm[k * 64 + j] = LUm[p * 64 + j]; LUm[p * 64 + j] = a; }
    __syncthreads();
    const float pivv = LUm[k * 64 + k];
    for (int i = k + 1 + tid; i < r; i += blockDim.x) LUm[i * 64 + k] /= pivv;
    __syncthreads();
    for (int e = tid; e < (r - k - 1) * (r - k - 1); e += blockDim.x) {
      int i = k + 1 + e / (r - k - 1), j = k + 1 + e % (r - k - 1);
      LUm[i * 64 + j] -= LUm[i * 64 + k] * LUm[k * 64 + j];
    }
    __syncthreads();
  }
  // s1 = Mx^{-T} (U'^T x2):  (P Mx = L Uu)  =>  Mx^T = Uu^T L^T P  => solve Uu^T y = b, L^T z = y, s1 = P^T z
  if (tid == 0) {
    float y[64];
    for (int i = 0; i < r; ++i) { float s = upx2[i]; for (int k = 0; k < i; ++k) s -= LUm[k * 64 + i] * y[k]; y[i] = s / LUm[i * 64 + i]; }
    for (int i = r - 1; i >= 0; --i) { float s = y[i]; for (int k = i + 1; k < r; ++k) s -= LUm[k * 64 + i] * y[k]; y[i] = s; }
    for (int k = r - 1; k >= 0; --k) { int p = piv[k]; if (p != k) { float a = y[k]; y[k] = y[p]; y[p] = a; } }
    for (int i = 0; i < r; ++i) s1[i] = y[i];
  }
  __syncthreads();
  // t = V'^T invQtv = V'^T x2 - G'vv s1 ;  s2 = Mx^{-1} t
  mv_small(Gvv, s1, tmp, r, RP, false);
  for (int i = tid; i < r; i += blockDim.x) t[i] = vpx2[i] - tmp[i];
  __syncthreads();
  if (tid == 0) {
    float y[64];
    for (int i = 0; i < r; ++i) y[i] = t[i];
    for (int k = 0; k < r; ++k) { int p = piv[k]; if (p != k) { float a = y[k]; y[k] = y[p]; y[p] = a; } }
    for (int i = 0; i < r; ++i) { float s = y[i]; for (int k = 0; k < i; ++k) s -= LUm[i * 64 + k] * y[k]; y[i] = s; }
    for (int i = r - 1; i >= 0; --i) { float s = y[i]; for (int k = i + 1; k < r; ++k) s -= LUm[i * 64 + k] * y[k]; y[i] = s / LUm[i * 64 + i]; }
    for (int i = 0; i < r; ++i) s2[i] = y[i];
  }
  __syncthreads();
  // ||a||^2 = ||x1 + U' c1||^2 ; ||b||^2 = ||x2 - V' s1||^2
  mv_small(Guu, c1, tmp, r, RP, false);
  const float na2 = x1sq + 2.f * dot_small(c1, upx1, r) + dot_small(c1, tmp, r);
  __syncthreads();
  mv_small(Gvv, s1, tmp, r, RP, false);
  const float nb2 = x2sq - 2.f * dot_small(s1, vpx2, r) + dot_small(s1, tmp, r);
  __syncthreads();
  const float na = sqrtf(fmaxf(na2, 0.f)), nb = sqrtf(fmaxf(nb2, 0.f));
  float* pvec = par + lra_par_vec_off(RP);
  float* pscal = par + lra_par_scal_off(RP);
  if (update_U) {  // psgd.py:1036-1043
    mv_small(Mx, c1, atX, r, RP, false);                 // atV = V'^T a = (I + V'^T U') c1
    for (int i = tid; i < r; i += blockDim.x) btX[i] = t[i];   // btV = V'^T b
    __syncthreads();
    mv_small(Gvv, atX, tmp, r, RP, false);
    const float n1 = sqrtf(fmaxf(dot_small(atX, tmp, r), 0.f));
    __syncthreads();
    mv_small(Gvv, btX, tmp, r, RP, false);
    const float n2 = sqrtf(fmaxf(dot_small(btX, tmp, r), 0.f));
    __syncthreads();
    const float ell = round_to(dtype, na * n1 + nb * n2);
    const float Ln = fmaxf(betaL * (*Lu) + (1.f - betaL) * ell, ell);
    mv_small(Mx, atX, tmp, r, RP, true);    // w_a = atV Mx  (row vector) = Mx^T atV
    mv_small(Mx, btX, tmp2, r, RP, true);
    for (int i = tid; i < RP; i += blockDim.x) { pvec[LV_WA * RP + i] = i < r ? tmp[i] : 0.f; pvec[LV_WB * RP + i] = i < r ? tmp2[i] : 0.f; }
    __syncthreads();
    if (tid == 0) { *Lu = Ln; pscal[LS_STEP] = lr / Ln; }
  } else {  // psgd.py:1045-1052
    for (int i = tid; i < r; i += blockDim.x) atX[i] = c2[i];  // atU = U'^T a
    mv_small(Gvu, s1, tmp, r, RP, true);                       // (V'^T U')^T s1 = U'^T V' s1
    for (int i = tid; i < r; i += blockDim.x) btX[i] = upx2[i] - tmp[i];   // btU = U'^T b
    __syncthreads();
    mv_small(Guu, atX, tmp, r, RP, false);
    const float n1 = sqrtf(fmaxf(dot_small(atX, tmp, r), 0.f));
    __syncthreads();
    mv_small(Guu, btX, tmp, r, RP, false);
    const float n2 = sqrtf(fmaxf(dot_small(btX, tmp, r), 0.f));
    __syncthreads();
    const float ell = round_to(dtype, na * n1 + nb * n2);
    const float Ln = fmaxf(betaL * (*Lv) + (1.f - betaL) * ell, ell);
    for (int i = tid; i < RP; i += blockDim.x) { pvec[LV_ATU * RP + i] = i < r ? atX[i] : 0.f; pvec[LV_BTU * RP + i] = i < r ? btX[i] : 0.f; }
    mv_small(Av, atX, tmp, r, RP, false);   // Av atU: V'_i . atU = V_i . (Av atU)
    mv_small(Av, btX, tmp2, r, RP, false);
    for (int i = tid; i < RP; i += blockDim.x) { pvec[LV_AVATU * RP + i] = i < r ? tmp[i] : 0.f; pvec[LV_AVBTU * RP + i] = i < r ? tmp2[i] : 0.f; }
    __syncthreads();
    if (tid == 0) { *Lv = Ln; pscal[LS_STEP] = lr / Ln; }
  }
  for (int i = tid; i < RP; i += blockDim.x) {
    pvec[LV_C1 * RP + i] = i < r ? c1[i] : 0.f; pvec[LV_C2 * RP + i] = i < r ? c2[i] : 0.f;
    pvec[LV_S1 * RP + i] = i < r ? s1[i] : 0.f; pvec[LV_S2 * RP + i] = i < r ? s2[i] : 0.f;
  }
  __syncthreads();
  // row-dot vectors against the UNtransformed rows: U'_i . c = U_i . (Au c)
  mv_small(Au, c1, tmp, r, RP, false);
  for (int i = tid; i < RP; i += blockDim.x) pvec[LV_AUC1 * RP + i] = i < r ? tmp[i] : 0.f;
  __syncthreads();
  mv_small(Av, c2, tmp, r, RP, false);
  for (int i = tid; i < RP; i += blockDim.x) pvec[LV_AVC2 * RP + i] = i < r ? tmp[i] : 0.f;
  __syncthreads();
  mv_small(Av, s1, tmp, r, RP, false);
  for (int i = tid; i < RP; i += blockDim.x) pvec[LV_AVS1 * RP + i] = i < r ? tmp[i] : 0.f;
  __syncthreads();
  mv_small(Au, s2, tmp, r, RP, false);
  for (int i = tid; i < RP; i += blockDim.x) pvec[LV_AUS2 * RP + i] = i < r ? tmp[i] : 0.f;
  __syncthreads();
  for (int e = tid; e < MM; e += blockDim.x) {
    int i = e / RP, j = e - i * RP;
    bool in = i < r && j < r;
    par[e] = in ? Au[e] : 0.f;
    par[MM + e] = in ? Av[e] : 0.f;
  }
}

// ------------------------------------------------------------------------------------------------
// sweep 2: thread-per-row; rotation + rank-2 update of U or V; per-row terms of the d update
// ------------------------------------------------------------------------------------------------
template <typename T, int RP>
__global__ void __launch_bounds__(128) k_lra_sweep2(T* __restrict__ U, T* __restrict__ V, const T* __restrict__ d,
                                                    const T* __restrict__ hvec, const T* __restrict__ vvec, long long n, int r,
                                                    const float* __restrict__ par, int update_U, float* __restrict__ dd_out,
                                                    float* __restrict__ scal_out) {
  extern __shared__ __align__(16) float smp[];
  float* Au = smp; float* Av = smp + RP * RP; float* pvec = Av + RP * RP;
  __shared__ float red[32];
  for (int e = threadIdx.x; e < 2 * RP * RP + LV_NVEC * RP; e += blockDim.x) smp[e] = par[e];
  __syncthreads();
  const float step = par[lra_par_scal_off(RP) + LS_STEP];
  float mx1 = 0.f, mx2 = 0.f;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long row = blockIdx.x * (long long)blockDim.x + threadIdx.x; row < n; row += stride) {
    float u[RP], v[RP];
    load_row<T, RP>(U, row, r, u);
    load_row<T, RP>(V, row, r, v);
    const float dd = to_f<T>(d[row]);
    const float hh = to_f<T>(hvec[row]);
    const float vv = to_f<T>(vvec[row]);
    const float x1 = to_f<T>(from_f<T>(dd * hh));
    const float x2 = to_f<T>(from_f<T>(vv / dd));
    float duc1 = 0.f, dvc2 = 0.f, dvs1 = 0.f, dus2 = 0.f, dva = 0.f, dvb = 0.f;
#pragma unroll
    for (int c = 0; c < RP; ++c) {
      duc1 = fmaf(u[c], pvec[LV_AUC1 * RP + c], duc1);
      dus2 = fmaf(u[c], pvec[LV_AUS2 * RP + c], dus2);
      dvc2 = fmaf(v[c], pvec[LV_AVC2 * RP + c], dvc2);
      dvs1 = fmaf(v[c], pvec[LV_AVS1 * RP + c], dvs1);
    }
    if (!update_U) {
#pragma unroll
      for (int c = 0; c < RP; ++c) { dva = fmaf(v[c], pvec[LV_AVATU * RP + c], dva); dvb = fmaf(v[c], pvec[LV_AVBTU * RP + c], dvb); }
    }
    const float a = x1 + duc1;                 // Qh_i            psgd.py:1017
    const float Ph = dd * (a + dvc2);          // Ph_i            psgd.py:1018
    const float b = x2 - dvs1;                 // invQtv_i        psgd.py:1024
    const float invPv = (b - dus2) / dd;       // invPv_i         psgd.py:1025-1026
    const float Phh = Ph * hh, vinv = vv * invPv;   // psgd.py:1029
    mx1 = fmaxf(mx1, fabsf(Phh)); mx2 = fmaxf(mx2, fabsf(vinv));
    dd_out[row] = Phh - vinv;
    // new rows, 4 output columns at a time (Au/Av read as broadcast float4)
    const float ca = update_U ? step * a : step * (a + dva);
    const float cb = update_U ? step * b : step * (b + dvb);
    float o[RP];
#pragma unroll
    for (int c0 = 0; c0 < RP; c0 += 4) {
      float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int k = 0; k < RP; ++k) {
        const float4 m = *reinterpret_cast<const float4*>(&Au[k * RP + c0]);
        s.x = fmaf(u[k], m.x, s.x); s.y = fmaf(u[k], m.y, s.y); s.z = fmaf(u[k], m.z, s.z); s.w = fmaf(u[k], m.w, s.w);
      }
      o[c0] = s.x; o[c0 + 1] = s.y; o[c0 + 2] = s.z; o[c0 + 3] = s.w;
    }
    if (update_U) {
#pragma unroll
      for (int c = 0; c < RP; ++c) o[c] -= ca * pvec[LV_WA * RP + c] - cb * pvec[LV_WB * RP + c];
    }
    store_row<T, RP>(U, row, r, o);
#pragma unroll
    for (int c0 = 0; c0 < RP; c0 += 4) {
      float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int k = 0; k < RP; ++k) {
        const float4 m = *reinterpret_cast<const float4*>(&Av[k * RP + c0]);
        s.x = fmaf(v[k], m.x, s.x); s.y = fmaf(v[k], m.y, s.y); s.z = fmaf(v[k], m.z, s.z); s.w = fmaf(v[k], m.w, s.w);
      }
      o[c0] = s.x; o[c0 + 1] = s.y; o[c0 + 2] = s.z; o[c0 + 3] = s.w;
    }
    if (!update_U) {
#pragma unroll
      for (int c = 0; c < RP; ++c) o[c] -= ca * pvec[LV_ATU * RP + c] - cb * pvec[LV_BTU * RP + c];
    }
    store_row<T, RP>(V, row, r, o);
  }
  mx1 = block_max(mx1, red);
  if (threadIdx.x == 0) atomic_max_nonneg(&scal_out[LS_MAX_PHH], mx1);
  mx2 = block_max(mx2, red);
  if (threadIdx.x == 0) atomic_max_nonneg(&scal_out[LS_MAX_VINV], mx2);
}

// Ld update + step (psgd.py:1030-1031)
__global__ void k_lra_Ld(float* scal, float lr, float betaL, float* Ld, int dtype) {
  if (threadIdx.x == 0) {
    float ell = round_to(dtype, scal[LS_MAX_PHH] + scal[LS_MAX_VINV]);
    float Ln = fmaxf(betaL * (*Ld) + (1.f - betaL) * ell, ell);
    *Ld = Ln;
    scal[LS_STEP_D] = lr / Ln;
  }
}
// d -= lr/Ld * (Phh - vinvPv) * d   (psgd.py:1032)
template <typename T>
__global__ void k_lra_d_update(T* __restrict__ d, const float* __restrict__ dd, long long n, const float* __restrict__ scal) {
  const float step = scal[LS_STEP_D];
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += stride) {
    float x = to_f<T>(d[i]);
    d[i] = from_f<T>(x - step * to_f<T>(from_f<T>(dd[i])) * x);
  }
}
// h = g + (damping + eps|g|) v   (psgd.py:1071-1072)
template <typename T>
__global__ void k_lra_damp(const T* __restrict__ g, const T* __restrict__ v, T* __restrict__ out, long long n, float damping, float eps) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += stride) {
    float gg = to_f<T>(g[i]);
    float dmp = to_f<T>(from_f<T>(damping + to_f<T>(from_f<T>(eps * fabsf(gg)))));
    out[i] = from_f<T>(gg + to_f<T>(from_f<T>(dmp * to_f<T>(v[i]))));
  }
}

// ------------------------------------------------------------------------------------------------
// apply (psgd.py:1055-1063): three thread-per-row sweeps
//   mode 0: p1 += V_i * (d_i g_i)
//   mode 1: g2_i = d_i g_i + U_i . p1 (stored fp32) ; p2 += U_i * g2_i
//   mode 2: out_i = d_i * (g2_i + V_i . p2) ; sumsq
// ------------------------------------------------------------------------------------------------
template <typename T, int RP>
__global__ void __launch_bounds__(128) k_lra_apply(const T* __restrict__ Mtx, const T* __restrict__ d, const T* __restrict__ g,
                                                   float* __restrict__ g2, T* __restrict__ out, long long n, int r, int mode,
                                                   const float* __restrict__ pin, float* __restrict__ pout, float* sumsq) {
  __shared__ float ps[RP];
  __shared__ float accs[4][RP];
  if (threadIdx.x < RP) ps[threadIdx.x] = (mode > 0 && threadIdx.x < r) ? pin[threadIdx.x] : 0.f;
  __syncthreads();
  float acc[RP];
#pragma unroll
  for (int c = 0; c < RP; ++c) acc[c] = 0.f;
  float ssq = 0.f;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long row = blockIdx.x * (long long)blockDim.x + threadIdx.x; row < n; row += stride) {
    float x[RP];
    load_row<T, RP>(Mtx, row, r, x);
    const float dd = to_f<T>(d[row]);
    if (mode == 0) {
      const float y = to_f<T>(from_f<T>(dd * to_f<T>(g[row])));
#pragma unroll
      for (int c = 0; c < RP; ++c) acc[c] = fmaf(x[c], y, acc[c]);
    } else if (mode == 1) {
      float dot = 0.f;
#pragma unroll
      for (int c = 0; c < RP; ++c) dot = fmaf(x[c], ps[c], dot);
      const float y = to_f<T>(from_f<T>(dd * to_f<T>(g[row]))) + dot;
      g2[row] = y;
#pragma unroll
      for (int c = 0; c < RP; ++c) acc[c] = fmaf(x[c], y, acc[c]);
    } else {
      float dot = 0.f;
#pragma unroll
      for (int c = 0; c < RP; ++c) dot = fmaf(x[c], ps[c], dot);
      T o = from_f<T>(dd * (g2[row] + dot));
      out[row] = o;
      float f = to_f<T>(o);
      ssq = fmaf(f, f, ssq);
    }
  }
  if (mode < 2) {
    // block reduce the RP accumulators: warp shuffle then smem across the 4 warps
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int c = 0; c < RP; ++c) {
      float s = warp_sum(acc[c]);
      if (lane == 0) accs[w][c] = s;
    }
    __syncthreads();
    if (threadIdx.x < RP && threadIdx.x < r) atomicAdd(&pout[threadIdx.x], accs[0][threadIdx.x] + accs[1][threadIdx.x] + accs[2][threadIdx.x] + accs[3][threadIdx.x]);
  } else if (sumsq) {
    float s = warp_sum(ssq);
    if ((threadIdx.x & 31) == 0) atomicAdd(sumsq, s);
  }
}

// ------------------------------------------------------------------------------------------------
// host orchestration
// ------------------------------------------------------------------------------------------------
struct LraWs {
  float* acc; float* par; float* p1; float* p2; float* dd; void* hbuf; char* zero_begin; size_t zero_bytes; size_t total; int RP;
};
static int pad_rank(int r) { int p = 4; while (p < r) p <<= 1; return p; }

static void layout_lra(const psgd_lra_t* l, void* base, LraWs& w) {
  char* b = reinterpret_cast<char*>(base);
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~size_t(255); return b ? (void*)(b + o) : (void*)(o + 256); };
  const int RP = pad_rank(l->r);
  w.RP = RP;
  w.zero_begin = (char*)take(0);
  size_t z0 = off;
  w.acc = (float*)take(lra_acc_floats(RP) * 4);
  w.par = (float*)take(lra_par_floats(RP) * 4);
  w.p1 = (float*)take(64 * 4);
  w.p2 = (float*)take(64 * 4);
  w.zero_bytes = off - z0;
  w.dd = (float*)take((size_t)l->n * 4);
  w.hbuf = take((size_t)l->n * dtype_size(l->dtype));
  w.total = off;
}

#define LRA_DISPATCH(dt, RP, ...)                                                          \
  do {                                                                                     \
    if ((dt) == PSGD_BF16) { typedef bf16 T;                                               \
      switch (RP) { case 4: { constexpr int R_ = 4; __VA_ARGS__; } break; case 8: { constexpr int R_ = 8; __VA_ARGS__; } break; \
        case 16: { constexpr int R_ = 16; __VA_ARGS__; } break; case 32: { constexpr int R_ = 32; __VA_ARGS__; } break;         \
        default: { constexpr int R_ = 64; __VA_ARGS__; } break; }                          \
    } else { typedef float T;                                                              \
      switch (RP) { case 4: { constexpr int R_ = 4; __VA_ARGS__; } break; case 8: { constexpr int R_ = 8; __VA_ARGS__; } break; \
        case 16: { constexpr int R_ = 16; __VA_ARGS__; } break; case 32: { constexpr int R_ = 32; __VA_ARGS__; } break;         \
        default: { constexpr int R_ = 64; __VA_ARGS__; } break; }                          \
    }                                                                                      \
  } while (0)

static int validate_lra(const psgd_lra_t* l) {
  if (!l || l->n < 1 || l->r < 1 || l->r > 64 || !l->U || !l->V || !l->d) return PSGD_ERR_INVALID_ARG;
  if (l->dtype != PSGD_BF16 && l->dtype != PSGD_F32) return PSGD_ERR_INVALID_ARG;
  return PSGD_OK;
}

// stages of one update: a row-sharded preconditioner (rows of U, V, d spread over ranks) all-reduces the sweep-1 sums between SWEEP1 and
// SWEEP2 and the two maxima of the d update between SWEEP2 and FINISH (psgd_torch_b200/lra_sharded.py); a single GPU runs all three
enum { LRA_ST_SWEEP1 = 1, LRA_ST_SWEEP2 = 2, LRA_ST_FINISH = 4, LRA_ST_ALL = 7 };

static int lra_update_impl(Ctx* ctx, const psgd_lra_t* l, const void* v, const void* hv, float lr, float betaL, int update_U, LraWs& w,
                           cudaStream_t st, int stages = LRA_ST_ALL, bool zeroed = false) {
  const int dt = l->dtype, RP = w.RP, r = l->r;
  const long long n = l->n;
  int rc;
  if ((stages & LRA_ST_SWEEP1) && !zeroed) { rc = check_cuda(ctx, cudaMemsetAsync(w.zero_begin, 0, w.zero_bytes, st), "memset"); if (rc) return rc; }
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
  // tensor-core sweeps: bf16, rank exactly 16 or 32 (rows are whole 16-byte pieces); everything else takes the CUDA-core sweeps
  const bool mma_path = dt == PSGD_BF16 && (r == 16 || r == 32) && al16(l->U) && al16(l->V) && al16(l->d) && al16(hv) && al16(v) && ctx->gemm_path != 1;
  const int smem_mma = 8 * LRA_STAGES * (r == 32 ? LraTile<32>::BYTES : LraTile<16>::BYTES);
  if (mma_path) {
    static PerDeviceOnce attr_g;
    if (attr_g.need(ctx->device)) {
      cudaFuncSetAttribute(k_lra_gram_mma<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * LRA_STAGES * LraTile<32>::BYTES);
      cudaFuncSetAttribute(k_lra_gram_mma<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * LRA_STAGES * LraTile<16>::BYTES);
      cudaFuncSetAttribute(k_lra_rotate_mma<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * LRA_STAGES * LraTile<32>::BYTES);
      cudaFuncSetAttribute(k_lra_rotate_mma<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * LRA_STAGES * LraTile<16>::BYTES);
    }
  }
  // sweep 1 on tcgen05 (lra_tc.cuh): bf16, rank 16 / 32 / 64, whole 128 * (64 / r)-row blocks; the remainder (< one block) goes through
  // the kernels below on offset pointers and adds into the same accumulators
  const bool tc_rank = r == 16 || r == 32 || r == 64;
  const int tc_rows = tc_rank ? LT_KR * (64 / r) : 1;
  const bool tc_path = dt == PSGD_BF16 && tc_rank && al16(l->U) && al16(l->V) && al16(l->d) && al16(hv) && al16(v) && ctx->gemm_path != 1 &&
                       ctx->encode_tiled && !(ctx->debug_flags & 1024) && n >= tc_rows;
  long long n_done = 0;     // rows already covered by the tcgen05 kernel
  if ((stages & LRA_ST_SWEEP1) && tc_path) {
    LtParams P;
    memset(&P, 0, sizeof(P));
    P.nblocks = n / tc_rows;
    n_done = P.nblocks * tc_rows;
    const long long prow = n_done / (64 / r);      // 128-byte lines
    if (prow < (1LL << 31)) {
      rc = make_tmap(ctx, &P.map_u, l->U, (int)prow, 64, 64, LT_KR); if (rc) return rc;
      rc = make_tmap(ctx, &P.map_v, l->V, (int)prow, 64, 64, LT_KR); if (rc) return rc;
      P.d = (const bf16*)l->d; P.h = (const bf16*)hv; P.v = (const bf16*)v; P.acc_out = w.acc;
      const int grid = (int)(P.nblocks < (long long)ctx->num_sms ? P.nblocks : (long long)ctx->num_sms);
      static PerDeviceOnce attr_t;
      if (attr_t.need(ctx->device)) {
        cudaFuncSetAttribute(k_lra_gram_tc<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, LtCfg<16>::SMEM_BYTES);
        cudaFuncSetAttribute(k_lra_gram_tc<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, LtCfg<32>::SMEM_BYTES);
        cudaFuncSetAttribute(k_lra_gram_tc<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, LtCfg<64>::SMEM_BYTES);
      }
      if (r == 16) k_lra_gram_tc<16><<<grid, LT_THREADS, LtCfg<16>::SMEM_BYTES, st>>>(P);
      else if (r == 32) k_lra_gram_tc<32><<<grid, LT_THREADS, LtCfg<32>::SMEM_BYTES, st>>>(P);
      else k_lra_gram_tc<64><<<grid, LT_THREADS, LtCfg<64>::SMEM_BYTES, st>>>(P);
      ctx->launches++; rc = check_cuda(ctx, cudaGetLastError(), "k_lra_gram_tc"); if (rc) return rc;
    } else {
      n_done = 0;
    }
  }
  if ((stages & LRA_ST_SWEEP1) && n_done > 0 && n_done < n) {
    // remainder rows: the mma.sync / CUDA-core sweep on the tail slice
    const long long nr = n - n_done;
    const size_t es = dtype_size(dt);
    const char* Ut = (const char*)l->U + (size_t)n_done * r * es; const char* Vt = (const char*)l->V + (size_t)n_done * r * es;
    const char* dt_ = (const char*)l->d + (size_t)n_done * es; const char* ht = (const char*)hv + (size_t)n_done * es;
    const char* vt = (const char*)v + (size_t)n_done * es;
    long long tiles = (nr + 63) / 64;
    int grid1 = (int)(tiles < (long long)ctx->num_sms * 2 ? tiles : (long long)ctx->num_sms * 2);
    LRA_DISPATCH(dt, RP, (k_lra_sweep1<T, R_><<<grid1, 256, 0, st>>>((const T*)Ut, (const T*)Vt, (const T*)dt_, (const T*)ht, (const T*)vt, nr, r, w.acc)));
    ctx->launches++; rc = check_cuda(ctx, cudaGetLastError(), "k_lra_sweep1(tail)"); if (rc) return rc;
  }
  if (!(stages & LRA_ST_SWEEP1) || n_done > 0) {
    // sums of sweep 1 already in w.acc (all-reduced by the caller), or formed above
  } else if (mma_path) {
    long long chunks = (n + 15) / 16;
    int grid1 = (int)((chunks + 7) / 8 < (long long)ctx->num_sms ? (chunks + 7) / 8 : (long long)ctx->num_sms);
    if (r == 32) k_lra_gram_mma<32><<<grid1, 256, smem_mma, st>>>((const bf16*)l->U, (const bf16*)l->V, (const bf16*)l->d, (const bf16*)hv, (const bf16*)v, n, w.acc);
    else k_lra_gram_mma<16><<<grid1, 256, smem_mma, st>>>((const bf16*)l->U, (const bf16*)l->V, (const bf16*)l->d, (const bf16*)hv, (const bf16*)v, n, w.acc);
    ctx->launches++; rc = check_cuda(ctx, cudaGetLastError(), "k_lra_gram_mma"); if (rc) return rc;
  } else {
    long long tiles = (n + 63) / 64;
    int grid1 = (int)(tiles < (long long)ctx->num_sms * 2 ? tiles : (long long)ctx->num_sms * 2);
    if (RP == 4) {   // thread-per-row kernel: everything fits in registers at this rank
      long long nb = (n + 255) / 256;
      int gridr = (int)(nb < (long long)ctx->num_sms * 8 ? nb : (long long)ctx->num_sms * 8);
      if (dt == PSGD_BF16) k_lra_sweep1_r4<bf16><<<gridr, 256, 0, st>>>((const bf16*)l->U, (const bf16*)l->V, (const bf16*)l->d, (const bf16*)hv, (const bf16*)v, n, r, w.acc);
      else k_lra_sweep1_r4<float><<<gridr, 256, 0, st>>>((const float*)l->U, (const float*)l->V, (const float*)l->d, (const float*)hv, (const float*)v, n, r, w.acc);
    } else
    LRA_DISPATCH(dt, RP, (k_lra_sweep1<T, R_><<<grid1, 256, 0, st>>>((const T*)l->U, (const T*)l->V, (const T*)l->d, (const T*)hv, (const T*)v, n, r, w.acc)));
    ctx->launches++; rc = check_cuda(ctx, cudaGetLastError(), "k_lra_sweep1"); if (rc) return rc;
  }
  float* scal = w.par + lra_par_scal_off(RP);
  if (stages & LRA_ST_SWEEP2) {
  size_t smem_small = ((size_t)8 * RP * RP + 24 * RP) * 4;
  static PerDeviceOnce small_attr;
  if (small_attr.need(ctx->device)) cudaFuncSetAttribute(k_lra_small, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
  k_lra_small<<<1, 256, smem_small, st>>>(w.acc, w.par, r, RP, lr, betaL, update_U, l->Lu, l->Lv, dt);
  ctx->launches++; rc = check_cuda(ctx, cudaGetLastError(), "k_lra_small"); if (rc) return rc;
  long long rows_blocks = (n + 127) / 128;
  int grid2 = (int)(rows_blocks < (long long)ctx->num_sms * 8 ? rows_blocks : (long long)ctx->num_sms * 8);
  size_t smem2 = ((size_t)2 * RP * RP + LV_NVEC * RP) * 4;
  // tcgen05 form (lra_tc.cuh) for the whole blocks, the kernels below for the remainder rows
  long long n2_done = 0;
  if (tc_path && !(ctx->debug_flags & 2048)) {
    const long long blocks = n / tc_rows;
    const long long prow = blocks * LT_KR;
    if (prow < (1LL << 31)) {
      LrParams P;
      memset(&P, 0, sizeof(P));
      rc = make_tmap(ctx, &P.map_u, l->U, (int)prow, 64, 64, LT_KR); if (rc) return rc;
      rc = make_tmap(ctx, &P.map_v, l->V, (int)prow, 64, 64, LT_KR); if (rc) return rc;
      P.U = (bf16*)l->U; P.V = (bf16*)l->V; P.d = (const bf16*)l->d; P.h = (const bf16*)hv; P.v = (const bf16*)v;
      P.nblocks = blocks; P.par = w.par; P.update_U = update_U; P.dd_out = w.dd; P.scal_out = scal;
      const int grid = (int)(blocks < (long long)ctx->num_sms ? blocks : (long long)ctx->num_sms);
      static PerDeviceOnce attr_r;
      if (attr_r.need(ctx->device)) {
        cudaFuncSetAttribute(k_lra_rotate_tc<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, LrCfg<16>::SMEM_BYTES);
        cudaFuncSetAttribute(k_lra_rotate_tc<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, LrCfg<32>::SMEM_BYTES);
        cudaFuncSetAttribute(k_lra_rotate_tc<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, LrCfg<64>::SMEM_BYTES);
      }
      if (r == 16) k_lra_rotate_tc<16><<<grid, LR_THREADS, LrCfg<16>::SMEM_BYTES, st>>>(P);
      else if (r == 32) k_lra_rotate_tc<32><<<grid, LR_THREADS, LrCfg<32>::SMEM_BYTES, st>>>(P);
      else k_lra_rotate_tc<64><<<grid, LR_THREADS, LrCfg<64>::SMEM_BYTES, st>>>(P);
      ctx->launches++; rc = check_cuda(ctx, cudaGetLastError(), "k_lra_rotate_tc"); if (rc) return rc;
      n2_done = blocks * tc_rows;
    }
  }
  const long long nr2 = n - n2_done;
  if (nr2 > 0) {
    const size_t es = dtype_size(dt);
    char* Ut = (char*)l->U + (size_t)n2_done * r * es; char* Vt = (char*)l->V + (size_t)n2_done * r * es;
    const char* dtl = (const char*)l->d + (size_t)n2_done * es; const char* ht = (const char*)hv + (size_t)n2_done * es;
    const char* vt = (const char*)v + (size_t)n2_done * es;
    float* ddt = w.dd + n2_done;
    if (mma_path) {
      long long chunks = (nr2 + 15) / 16;
      int gridr = (int)((chunks + 7) / 8 < (long long)ctx->num_sms ? (chunks + 7) / 8 : (long long)ctx->num_sms);
      if (r == 32) k_lra_rotate_mma<32><<<gridr, 256, smem_mma, st>>>((bf16*)Ut, (bf16*)Vt, (const bf16*)dtl, (const bf16*)ht, (const bf16*)vt, nr2, w.par, update_U, ddt, scal);
      else k_lra_rotate_mma<16><<<gridr, 256, smem_mma, st>>>((bf16*)Ut, (bf16*)Vt, (const bf16*)dtl, (const bf16*)ht, (const bf16*)vt, nr2, w.par, update_U, ddt, scal);
    } else {
      long long rb2 = (nr2 + 127) / 128;
      int grid2t = (int)(rb2 < (long long)ctx->num_sms * 8 ? rb2 : (long long)ctx->num_sms * 8);
      LRA_DISPATCH(dt, RP, {
        static PerDeviceOnce attr;
        if (attr.need(ctx->device)) cudaFuncSetAttribute(k_lra_sweep2<T, R_>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
        k_lra_sweep2<T, R_><<<grid2t, 128, smem2, st>>>((T*)Ut, (T*)Vt, (const T*)dtl, (const T*)ht, (const T*)vt, nr2, r, w.par, update_U, ddt, scal);
      });
    }
  ctx->launches++; rc = check_cuda(ctx, cudaGetLastError(), "k_lra_sweep2"); if (rc) return rc;
  }
  }
  if (!(stages & LRA_ST_FINISH)) return PSGD_OK;
  k_lra_Ld<<<1, 32, 0, st>>>(scal, lr, betaL, l->Ld, dt);
  ctx->launches++; rc = check_cuda(ctx, cudaGetLastError(), "k_lra_Ld"); if (rc) return rc;
  int gridd = (int)(((n + 255) / 256) < (long long)ctx->num_sms * 8 ? ((n + 255) / 256) : (long long)ctx->num_sms * 8);
  if (dt == PSGD_BF16) k_lra_d_update<bf16><<<gridd, 256, 0, st>>>((bf16*)l->d, w.dd, n, scal);
  else k_lra_d_update<float><<<gridd, 256, 0, st>>>((float*)l->d, w.dd, n, scal);
  ctx->launches++; return check_cuda(ctx, cudaGetLastError(), "k_lra_d_update");
}

}  // namespace psgd

using namespace psgd;

extern "C" {

size_t psgd_lra_workspace_bytes(psgd_handle_t, const psgd_lra_t* l) {
  if (validate_lra(l)) return 0;
  LraWs w;
  layout_lra(l, nullptr, w);
  return w.total;
}

int psgd_lra_update(psgd_handle_t h, const psgd_lra_t* l, const void* v, const void* hvec, float lr, float betaL, int update_U,
                    void* workspace, size_t workspace_bytes, void* stream) {
  Ctx* ctx = reinterpret_cast<Ctx*>(h);
  if (!ctx || !v || !hvec) return PSGD_ERR_INVALID_ARG;
  int rc = validate_lra(l); if (rc) return rc;
  if (!l->Lu || !l->Lv || !l->Ld) return PSGD_ERR_INVALID_ARG;
  LraWs w;
  layout_lra(l, workspace, w);
  if (!workspace || workspace_bytes < w.total) return PSGD_ERR_WORKSPACE;
  return lra_update_impl(ctx, l, v, hvec, lr, betaL, update_U, w, reinterpret_cast<cudaStream_t>(stream));
}

int psgd_lra_whiten_update(psgd_handle_t h, const psgd_lra_t* l, const void* g, const void* v, float lr, float betaL, float damping,
                           int update_U, void* workspace, size_t workspace_bytes, void* stream) {
  Ctx* ctx = reinterpret_cast<Ctx*>(h);
  if (!ctx || !v || !g) return PSGD_ERR_INVALID_ARG;
  int rc = validate_lra(l); if (rc) return rc;
  if (!l->Lu || !l->Lv || !l->Ld) return PSGD_ERR_INVALID_ARG;
  LraWs w;
  layout_lra(l, workspace, w);
  if (!workspace || workspace_bytes < w.total) return PSGD_ERR_WORKSPACE;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const long long n = l->n;
  int grid = (int)(((n + 255) / 256) < (long long)ctx->num_sms * 8 ? ((n + 255) / 256) : (long long)ctx->num_sms * 8);
  if (l->dtype == PSGD_BF16) k_lra_damp<bf16><<<grid, 256, 0, st>>>((const bf16*)g, (const bf16*)v, (bf16*)w.hbuf, n, damping, dtype_eps(PSGD_BF16));
  else k_lra_damp<float><<<grid, 256, 0, st>>>((const float*)g, (const float*)v, (float*)w.hbuf, n, damping, dtype_eps(PSGD_F32));
  ctx->launches++; rc = check_cuda(ctx, cudaGetLastError(), "k_lra_damp"); if (rc) return rc;
  return lra_update_impl(ctx, l, v, w.hbuf, lr, betaL, update_U, w, st);
}

// psgd.py:1193-1198: the pair (v, hvp) with independent damping noise z on the Hessian-vector product: h = hvp + (damping + eps|hvp|) z
int psgd_lra_newton_update(psgd_handle_t h, const psgd_lra_t* l, const void* v, const void* hvp, const void* z, float lr, float betaL,
                           float damping, int update_U, void* workspace, size_t workspace_bytes, void* stream) {
  Ctx* ctx = reinterpret_cast<Ctx*>(h);
  if (!ctx || !v || !hvp || !z) return PSGD_ERR_INVALID_ARG;
  int rc = validate_lra(l); if (rc) return rc;
  if (!l->Lu || !l->Lv || !l->Ld) return PSGD_ERR_INVALID_ARG;
  LraWs w;
  layout_lra(l, workspace, w);
  if (!workspace || workspace_bytes < w.total) return PSGD_ERR_WORKSPACE;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const long long n = l->n;
  int grid = (int)(((n + 255) / 256) < (long long)ctx->num_sms * 8 ? ((n + 255) / 256) : (long long)ctx->num_sms * 8);
  if (l->dtype == PSGD_BF16) k_lra_damp<bf16><<<grid, 256, 0, st>>>((const bf16*)hvp, (const bf16*)z, (bf16*)w.hbuf, n, damping, dtype_eps(PSGD_BF16));
  else k_lra_damp<float><<<grid, 256, 0, st>>>((const float*)hvp, (const float*)z, (float*)w.hbuf, n, damping, dtype_eps(PSGD_F32));
  ctx->launches++; rc = check_cuda(ctx, cudaGetLastError(), "k_lra_damp"); if (rc) return rc;
  return lra_update_impl(ctx, l, v, w.hbuf, lr, betaL, update_U, w, st);
}

static int lra_apply_impl(psgd_handle_t h, const psgd_lra_t* l, const void* g, void* out, float* sumsq_out, void* workspace,
                          size_t workspace_bytes, void* stream, int modes);

int psgd_lra_precond_grad(psgd_handle_t h, const psgd_lra_t* l, const void* g, void* out, float* sumsq_out, void* workspace,
                          size_t workspace_bytes, void* stream) {
  return lra_apply_impl(h, l, g, out, sumsq_out, workspace, workspace_bytes, stream, 7);
}

// modes: bit 0 = V^T (d g) (zeroes the projections first), bit 1 = y = d g + U p1 and U^T y, bit 2 = out = d (y + V p2).  A row-sharded
// preconditioner all-reduces the projection after bit 0 and after bit 1 (psgd_lra_workspace_offsets [2], [3]) and sumsq_out at the end.
int psgd_lra_precond_grad_staged(psgd_handle_t h, const psgd_lra_t* l, const void* g, void* out, float* sumsq_out, int modes, void* workspace,
                                 size_t workspace_bytes, void* stream) {
  if (!(modes & 7)) return PSGD_ERR_INVALID_ARG;
  return lra_apply_impl(h, l, g, out, sumsq_out, workspace, workspace_bytes, stream, modes);
}

static int lra_apply_impl(psgd_handle_t h, const psgd_lra_t* l, const void* g, void* out, float* sumsq_out, void* workspace,
                          size_t workspace_bytes, void* stream, int modes) {
  Ctx* ctx = reinterpret_cast<Ctx*>(h);
  if (!ctx || !g || !out) return PSGD_ERR_INVALID_ARG;
  int rc = validate_lra(l); if (rc) return rc;
  LraWs w;
  layout_lra(l, workspace, w);
  if (!workspace || workspace_bytes < w.total) return PSGD_ERR_WORKSPACE;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int dt = l->dtype, RP = w.RP, r = l->r;
  const long long n = l->n;
  if (modes & 1) {
    rc = check_cuda(ctx, cudaMemsetAsync(w.p1, 0, 64 * 4, st), "memset"); if (rc) return rc;
    rc = check_cuda(ctx, cudaMemsetAsync(w.p2, 0, 64 * 4, st), "memset"); if (rc) return rc;
    if (sumsq_out) { rc = check_cuda(ctx, cudaMemsetAsync(sumsq_out, 0, 4, st), "memset"); if (rc) return rc; }
  }
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
  if (dt == PSGD_BF16 && (r == 16 || r == 32) && al16(l->U) && al16(l->V) && ctx->gemm_path != 1) {
    const int rpi = r == 32 ? 8 : 16;   // rows per warp instruction
    long long need = ((n + rpi - 1) / rpi + 4 * 8 - 1) / (4 * 8);   // blocks of 8 warps x 4 groups
    int gridc = (int)(need < (long long)ctx->num_sms * 8 ? need : (long long)ctx->num_sms * 8);
    if (gridc < 1) gridc = 1;
    // whole 256-row blocks through the bulk-copy ring, the remainder (< 256 rows) through the direct-load kernel on offset pointers
    const bool tma_ok = al16(l->d) && al16(g) && !(ctx->debug_flags & 64);
    const long long n_full = tma_ok ? n / 256 : 0;
    const long long n_rem = n - n_full * 256;
    const int tile_bytes = 256 * r * 2 + 256 * 2 + 256 * 2 + 256 * 4;
    static PerDeviceOnce attr_a;
    if (attr_a.need(ctx->device)) {
      cudaFuncSetAttribute(k_lra_apply_tma<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * (256 * 32 * 2 + 2048));
      cudaFuncSetAttribute(k_lra_apply_tma<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * (256 * 16 * 2 + 2048));
    }
    int gridt = (int)(n_full < (long long)ctx->num_sms ? n_full : (long long)ctx->num_sms);
    long long need_r = ((n_rem + rpi - 1) / rpi + 4 * 8 - 1) / (4 * 8);
    int gridrem = (int)(need_r < 1 ? 1 : need_r);
    for (int mode = 0; mode < 3; ++mode) {
      if (!(modes & (1 << mode))) continue;
      const bf16* Mx = (const bf16*)(mode == 1 ? l->U : l->V);
      const float* pin = mode == 0 ? nullptr : (mode == 1 ? w.p1 : w.p2);
      float* pout = mode == 0 ? w.p1 : (mode == 1 ? w.p2 : nullptr);
      if (n_full > 0) {
        if (r == 32) k_lra_apply_tma<32><<<gridt, 32 * (LRA_APPLY_CW + 1), 8 * tile_bytes, st>>>(Mx, (const bf16*)l->d, (const bf16*)g, w.dd, (bf16*)out, n_full, mode, pin, pout, sumsq_out);
        else k_lra_apply_tma<16><<<gridt, 32 * (LRA_APPLY_CW + 1), 8 * tile_bytes, st>>>(Mx, (const bf16*)l->d, (const bf16*)g, w.dd, (bf16*)out, n_full, mode, pin, pout, sumsq_out);
        ctx->launches++; rc = check_cuda(ctx, cudaGetLastError(), "k_lra_apply_tma"); if (rc) return rc;
      }
      if (n_rem > 0) {
        const long long o = n_full * 256;
        if (r == 32) k_lra_apply_bf16<32><<<gridrem, 256, 0, st>>>(Mx + o * r, (const bf16*)l->d + o, (const bf16*)g + o, w.dd + o, (bf16*)out + o, n_rem, mode, pin, pout, sumsq_out);
        else k_lra_apply_bf16<16><<<gridrem, 256, 0, st>>>(Mx + o * r, (const bf16*)l->d + o, (const bf16*)g + o, w.dd + o, (bf16*)out + o, n_rem, mode, pin, pout, sumsq_out);
        ctx->launches++; rc = check_cuda(ctx, cudaGetLastError(), "k_lra_apply_bf16"); if (rc) return rc;
      }
    }
    return PSGD_OK;
  }
  long long rb = (n + 127) / 128;
  int grid = (int)(rb < (long long)ctx->num_sms * 8 ? rb : (long long)ctx->num_sms * 8);
  if (modes & 1) {
    LRA_DISPATCH(dt, RP, (k_lra_apply<T, R_><<<grid, 128, 0, st>>>((const T*)l->V, (const T*)l->d, (const T*)g, w.dd, (T*)out, n, r, 0, nullptr, w.p1, nullptr)));
    ctx->launches++; rc = check_cuda(ctx, cudaGetLastError(), "k_lra_apply0"); if (rc) return rc;
  }
  if (modes & 2) {
    LRA_DISPATCH(dt, RP, (k_lra_apply<T, R_><<<grid, 128, 0, st>>>((const T*)l->U, (const T*)l->d, (const T*)g, w.dd, (T*)out, n, r, 1, w.p1, w.p2, nullptr)));
    ctx->launches++; rc = check_cuda(ctx, cudaGetLastError(), "k_lra_apply1"); if (rc) return rc;
  }
  if (modes & 4) {
    LRA_DISPATCH(dt, RP, (k_lra_apply<T, R_><<<grid, 128, 0, st>>>((const T*)l->V, (const T*)l->d, (const T*)g, w.dd, (T*)out, n, r, 2, w.p2, nullptr, sumsq_out)));
    ctx->launches++; rc = check_cuda(ctx, cudaGetLastError(), "k_lra_apply2"); if (rc) return rc;
  }
  return PSGD_OK;
}


// ------------------------------------------------------------------------------------------------
// Row-sharded LRA (rows of U, V, d, g spread over the GPUs of a box): the same kernels, issued stage by stage so that the host can
// all-reduce the few cross-row quantities in between (NCCL over NVLink; psgd_torch_b200/lra_sharded.py).
//   offsets (bytes into the workspace) / counts (floats): [0] sweep-1 sums (SUM), [1] the two maxima of the d update (MAX),
//   [2] apply projection after mode 0 (SUM), [3] apply projection after mode 1 (SUM)
// ------------------------------------------------------------------------------------------------
int psgd_lra_workspace_offsets(psgd_handle_t, const psgd_lra_t* l, size_t* offsets, size_t* counts) {
  if (validate_lra(l) || !offsets || !counts) return PSGD_ERR_INVALID_ARG;
  LraWs w;
  layout_lra(l, nullptr, w);
  auto off = [](const void* p) { return (size_t)(reinterpret_cast<uintptr_t>(p) - 256); };   // sizing mode hands out offset + 256
  offsets[0] = off(w.acc); counts[0] = lra_acc_floats(w.RP);
  offsets[1] = off(w.par + lra_par_scal_off(w.RP) + LS_MAX_PHH); counts[1] = 2;
  offsets[2] = off(w.p1); counts[2] = 64;
  offsets[3] = off(w.p2); counts[3] = 64;
  return PSGD_OK;
}

// stages: 1 = (damping, if whiten) + sweep 1, 2 = r x r algebra + sweep 2, 4 = Ld + d update.  gh = g (whiten != 0: h = g + (damping + eps|g|) v is
// formed here) or h itself.
int psgd_lra_update_staged(psgd_handle_t h, const psgd_lra_t* l, const void* gh, const void* v, float lr, float betaL, float damping,
                           int whiten, int update_U, int stages, void* workspace, size_t workspace_bytes, void* stream) {
  Ctx* ctx = reinterpret_cast<Ctx*>(h);
  if (!ctx || !v || !gh || !(stages & LRA_ST_ALL)) return PSGD_ERR_INVALID_ARG;
  int rc = validate_lra(l); if (rc) return rc;
  if (!l->Lu || !l->Lv || !l->Ld) return PSGD_ERR_INVALID_ARG;
  LraWs w;
  layout_lra(l, workspace, w);
  if (!workspace || workspace_bytes < w.total) return PSGD_ERR_WORKSPACE;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const long long n = l->n;
  const void* hv = gh;
  if (whiten) {
    hv = w.hbuf;
    if (stages & LRA_ST_SWEEP1) {
      int grid = (int)(((n + 255) / 256) < (long long)ctx->num_sms * 8 ? ((n + 255) / 256) : (long long)ctx->num_sms * 8);
      if (l->dtype == PSGD_BF16) k_lra_damp<bf16><<<grid, 256, 0, st>>>((const bf16*)gh, (const bf16*)v, (bf16*)w.hbuf, n, damping, dtype_eps(PSGD_BF16));
      else k_lra_damp<float><<<grid, 256, 0, st>>>((const float*)gh, (const float*)v, (float*)w.hbuf, n, damping, dtype_eps(PSGD_F32));
      ctx->launches++; rc = check_cuda(ctx, cudaGetLastError(), "k_lra_damp"); if (rc) return rc;
    }
  }
  return lra_update_impl(ctx, l, v, hv, lr, betaL, update_U, w, st, stages);
}

}  // extern "C"
