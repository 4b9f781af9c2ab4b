// kron_kernels.cuh -- the bandwidth-bound kernels of the Kron path (everything that is not a big GEMM).
// Each kernel cites the reference lines it implements. All are stream-ordered and branch on device scalars
// only (no host synchronisation anywhere in an update, cf. SURVEY.md 7 hard part 7).
#pragma once
#include "common.cuh"

namespace psgd {

// slots of the per-bound scalar block (floats)
enum { SC_INV_NF = 0, SC_NF = 1, SC_J = 2, SC_BOUND = 3, SC_COUNT = 8 };
// (slots of the per-factor update scalars FS_*: common.cuh)

// 8 bf16 <-> 8 floats through one 16-byte access
__device__ __forceinline__ void ld8(const bf16* p, float* x) {
  uint4 u = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int t = 0; t < 4; ++t) { float2 f = __bfloat1622float2(h[t]); x[2 * t] = f.x; x[2 * t + 1] = f.y; }
}
__device__ __forceinline__ void st8(bf16* p, const float* x) {
  uint4 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int t = 0; t < 4; ++t) h[t] = __floats2bfloat162_rn(x[2 * t], x[2 * t + 1]);
  *reinterpret_cast<uint4*>(p) = u;
}
__device__ __forceinline__ float rbf(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }

// G' = G + (damping + eps*|G|) * N     psgd.py:402-403
template <typename T>
__global__ void k_add_noise(const T* __restrict__ G, const T* __restrict__ Nz, T* __restrict__ out, size_t numel,
                            float damping, float eps) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < numel; i += stride) {
    float g = to_f<T>(G[i]);
    // the reference rounds (damping + eps|G|) and the product to the tensor dtype before the add
    float d = to_f<T>(from_f<T>(damping + to_f<T>(from_f<T>(eps * fabsf(g)))));
    float dn = to_f<T>(from_f<T>(d * to_f<T>(Nz[i])));
    out[i] = from_f<T>(g + dn);
  }
}
// bf16, numel % 8 == 0, 16-byte aligned pointers
__global__ void k_add_noise_bf16x8(const bf16* __restrict__ G, const bf16* __restrict__ Nz, bf16* __restrict__ out, size_t nvec,
                                   float damping, float eps) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < nvec; i += stride) {
    float g[8], z[8], o[8];
    ld8(G + i * 8, g); ld8(Nz + i * 8, z);
#pragma unroll
    for (int t = 0; t < 8; ++t) { float d = rbf(damping + rbf(eps * fabsf(g[t]))); o[t] = g[t] + rbf(d * z[t]); }
    st8(out + i * 8, o);
  }
}

// ------------------------------ in-kernel noise (performance mode: psgd_kron_noise_t pointers == NULL) ------------------------------
// Philox4x32-10 (Salmon et al., SC'11; the generator behind torch's CUDA randn), counter-based: element group g of stream `strm` of a call
// with (seed, offset) always gets the same four 32-bit words, whatever the launch geometry.
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, ctr.x), lo0 = 0xD2511F53u * ctr.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, ctr.z), lo1 = 0xCD9E8D57u * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += 0x9E3779B9u; key.y += 0xBB67AE85u;
  }
  return ctr;
}
// four standard normals from one Philox block (Box-Muller on two pairs of uniforms; fast-math log / sincos: this is damping noise)
__device__ __forceinline__ void philox_normal4(uint64_t seed, uint64_t offset, uint32_t strm, uint64_t group, float* z) {
  const uint4 r = philox4x32_10(make_uint4((uint32_t)group, (uint32_t)(group >> 32), strm, (uint32_t)offset),
                                make_uint2((uint32_t)seed ^ (uint32_t)(offset >> 32), (uint32_t)(seed >> 32)));
  const float u0 = ((float)r.x + 1.0f) * 2.3283064365386963e-10f, u1 = (float)r.y * 2.3283064365386963e-10f;
  const float u2 = ((float)r.z + 1.0f) * 2.3283064365386963e-10f, u3 = (float)r.w * 2.3283064365386963e-10f;
  const float ra = sqrtf(-2.0f * __logf(u0)), rb = sqrtf(-2.0f * __logf(u2));
  float sa, ca, sb, cb;
  __sincosf(6.283185307179586f * u1, &sa, &ca);
  __sincosf(6.283185307179586f * u3, &sb, &cb);
  z[0] = ra * ca; z[1] = ra * sa; z[2] = rb * cb; z[3] = rb * sb;
}
// eight standard normals from one Philox block: 16-bit uniforms (enough for noise that is rounded to bf16 -- 8 mantissa bits -- right
// away; tails to 4.7 sigma).  Halves the integer work per normal: the damping-noise pass is bound by it, not by HBM.
__device__ __forceinline__ void philox_normal8(uint64_t seed, uint64_t offset, uint32_t strm, uint64_t group, float* z) {
  const uint4 r = philox4x32_10(make_uint4((uint32_t)group, (uint32_t)(group >> 32), strm, (uint32_t)offset),
                                make_uint2((uint32_t)seed ^ (uint32_t)(offset >> 32), (uint32_t)(seed >> 32)));
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const float u0 = ((float)(w[t] & 0xffffu) + 1.0f) * 1.52587890625e-05f;     // (0, 1]
    const float ang = (float)(w[t] >> 16) * (6.283185307179586f * 1.52587890625e-05f);
    const float rad = sqrtf(-2.0f * __logf(u0));
    float sn, cs;
    __sincosf(ang, &sn, &cs);
    z[2 * t] = rad * cs; z[2 * t + 1] = rad * sn;
  }
}

// ------------------------------ batched forms: blockIdx.y = unit of a same-shape batch, pointers from a table ------------------------------
constexpr int KB_MAX = 16;   // units per batched call
struct PtrTab { void* p[KB_MAX]; };
struct CPtrTab { const void* p[KB_MAX]; };

template <typename T, bool VEC8>
__global__ void k_add_noise_multi(CPtrTab G, CPtrTab Nz, PtrTab out, size_t numel, float damping, float eps) {
  const T* g = reinterpret_cast<const T*>(G.p[blockIdx.y]);
  const T* z = reinterpret_cast<const T*>(Nz.p[blockIdx.y]);
  T* o = reinterpret_cast<T*>(out.p[blockIdx.y]);
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  if (VEC8) {   // bf16, numel % 8 == 0, 16-byte aligned pointers
    const size_t nvec = numel / 8;
    for (; i < nvec; i += stride) {
      float a[8], b[8], r[8];
      ld8(reinterpret_cast<const bf16*>(g) + i * 8, a); ld8(reinterpret_cast<const bf16*>(z) + i * 8, b);
#pragma unroll
      for (int t = 0; t < 8; ++t) { float d = rbf(damping + rbf(eps * fabsf(a[t]))); r[t] = a[t] + rbf(d * b[t]); }
      st8(reinterpret_cast<bf16*>(o) + i * 8, r);
    }
  } else {
    for (; i < numel; i += stride) {
      float a = to_f<T>(g[i]);
      float d = to_f<T>(from_f<T>(damping + to_f<T>(from_f<T>(eps * fabsf(a)))));
      float dn = to_f<T>(from_f<T>(d * to_f<T>(z[i])));
      o[i] = from_f<T>(a + dn);
    }
  }
}

// G' = G + (damping + eps|G|) N with N drawn in the kernel (stream 0 of the unit's Philox sequence): one read and one write of G instead of
// torch's randn kernel + a three-pass add.  seeds[u] / offsets[u] per unit.
struct U64Tab { uint64_t v[KB_MAX]; };
template <typename T, bool VEC8>
__global__ void k_add_noise_philox_multi(CPtrTab G, PtrTab out, size_t numel, float damping, float eps, U64Tab seeds, U64Tab offsets) {
  const T* g = reinterpret_cast<const T*>(G.p[blockIdx.y]);
  T* o = reinterpret_cast<T*>(out.p[blockIdx.y]);
  const uint64_t seed = seeds.v[blockIdx.y], off = offsets.v[blockIdx.y];
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  if (VEC8) {
    const size_t nvec = numel / 8;
    for (; i < nvec; i += stride) {
      float a[8], z[8], r[8];
      ld8(reinterpret_cast<const bf16*>(g) + i * 8, a);
      philox_normal8(seed, off, 0u, i, z);
#pragma unroll
      for (int t = 0; t < 8; ++t) { const float d = rbf(damping + rbf(eps * fabsf(a[t]))); r[t] = a[t] + rbf(d * rbf(z[t])); }
      st8(reinterpret_cast<bf16*>(o) + i * 8, r);
    }
  } else {
    const size_t ngrp = (numel + 3) / 4;
    for (; i < ngrp; i += stride) {
      float z[4];
      philox_normal4(seed, off, 0u, i, z);
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const size_t e = 4 * i + t;
        if (e < numel) {
          const float a = to_f<T>(g[e]);
          const float d = to_f<T>(from_f<T>(damping + to_f<T>(from_f<T>(eps * fabsf(a)))));
          o[e] = from_f<T>(a + to_f<T>(from_f<T>(d * to_f<T>(from_f<T>(z[t])))));
        }
      }
    }
  }
}

// probe blocks (32 x s standard normals each) of a batch: table entry e = (buffer, stream id of its unit's sequence); blockIdx.y = entry
struct ProbeTab { void* p[4 * KB_MAX]; uint64_t seed[4 * KB_MAX]; uint64_t off[4 * KB_MAX]; uint32_t strm[4 * KB_MAX]; };
template <typename T>
__global__ void k_philox_probes(const __grid_constant__ ProbeTab tab, size_t numel) {
  T* o = reinterpret_cast<T*>(tab.p[blockIdx.y]);
  const uint64_t seed = tab.seed[blockIdx.y], off = tab.off[blockIdx.y];
  const uint32_t strm = tab.strm[blockIdx.y];
  const size_t ngrp = (numel + 3) / 4;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < ngrp; i += (size_t)gridDim.x * blockDim.x) {
    float z[4];
    philox_normal4(seed, off, strm, i, z);
#pragma unroll
    for (int t = 0; t < 4; ++t)
      if (4 * i + t < numel) o[4 * i + t] = from_f<T>(z[t]);
  }
}

template <typename T>
__global__ void k_square_to_f32_multi(CPtrTab q, PtrTab out, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { float v = to_f<T>(reinterpret_cast<const T*>(q.p[blockIdx.y])[i]); reinterpret_cast<float*>(out.p[blockIdx.y])[i] = v * v; }
}

// k_scale2d for a batch (1-D tensors / diag x diag: tiny, the atomics per element are fine)
template <typename T>
__global__ void k_scale2d_multi(CPtrTab X, PtrTab out, int m, int n, CPtrTab rs, CPtrTab cs, PtrTab row_sumsq, PtrTab col_sumsq, PtrTab total_sumsq) {
  const int u = blockIdx.y;
  const T* x = reinterpret_cast<const T*>(X.p[u]);
  T* o = reinterpret_cast<T*>(out.p[u]);
  const float* r_ = reinterpret_cast<const float*>(rs.p[u]);
  const float* c_ = reinterpret_cast<const float*>(cs.p[u]);
  float* rss = reinterpret_cast<float*>(row_sumsq.p[u]);
  float* css = reinterpret_cast<float*>(col_sumsq.p[u]);
  float* tss = reinterpret_cast<float*>(total_sumsq.p[u]);
  const size_t numel = (size_t)m * n;
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  float tot = 0.f;
  for (; i < numel; i += stride) {
    const int r = (int)(i / n), c = (int)(i % n);
    const float v = to_f<T>(x[i]) * (r_ ? r_[r] : 1.f) * (c_ ? c_[c] : 1.f);
    const T ov = from_f<T>(v);
    if (o) o[i] = ov;
    const float f = to_f<T>(ov);
    tot += f * f;
    if (rss) atomicAdd(&rss[r], f * f);
    if (css) atomicAdd(&css[c], f * f);
  }
  if (tss) { tot = warp_sum(tot); if ((threadIdx.x & 31) == 0) atomicAdd(tss, tot); }
}

// q2[i] = float(q[i])^2  (diagonal factor applied twice: Q^T Q)   psgd.py:327 with 1-D q
template <typename T>
__global__ void k_square_to_f32(const T* __restrict__ q, float* __restrict__ out, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { float v = to_f<T>(q[i]); out[i] = v * v; }
}

// out[i,j] = X[i,j] * rs[i] * cs[j] (+ reductions). Used when no dense factor exists (1-D tensors, diag x diag).
template <typename T>
__global__ void k_scale2d(const T* __restrict__ X, T* __restrict__ out, int m, int n, const float* __restrict__ rs,
                          const float* __restrict__ cs, float* row_sumsq, float* col_sumsq, float* total_sumsq) {
  size_t numel = (size_t)m * n;
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  float tot = 0.f;
  for (; i < numel; i += stride) {
    int r = (int)(i / n), c = (int)(i % n);
    float v = to_f<T>(X[i]) * (rs ? rs[r] : 1.f) * (cs ? cs[c] : 1.f);
    T o = from_f<T>(v);
    if (out) out[i] = o;     // out == nullptr: reductions only (sums of squares of X itself)
    float f = to_f<T>(o);
    tot += f * f;
    if (row_sumsq) atomicAdd(&row_sumsq[r], f * f);
    if (col_sumsq) atomicAdd(&col_sumsq[c], f * f);
  }
  if (total_sumsq) { tot = warp_sum(tot); if ((threadIdx.x & 31) == 0) atomicAdd(total_sumsq, tot); }
}

// R = Q^T - Q  (psgd.py:117) with max|R| (psgd.py:84) and row sums of squares (psgd.py:86) fused.
// grid (T, T), T = ceil(s/64), block 256; block (bi <= bj) loads the tile pair Q[bi,bj], Q[bj,bi] once and writes both
// R[bi,bj] and its mirror R[bj,bi] = -R[bi,bj]^T (exactly skew by construction); blocks with bi > bj exit.
template <typename T>
__global__ void __launch_bounds__(256) k_skew(const T* __restrict__ Q, T* __restrict__ R, int s, float* abs_max, float* row_sumsq) {
  const int bi = blockIdx.y, bj = blockIdx.x;
  if (bi > bj) return;
  __shared__ float tA[64][65];  // Q[i0 + y][j0 + x]
  __shared__ float tB[64][65];  // Q[j0 + y][i0 + x]
  __shared__ float red[32];
  const int i0 = bi * 64, j0 = bj * 64, tid = threadIdx.x;
  for (int e = tid; e < 64 * 64; e += 256) {
    const int y = e >> 6, x = e & 63;
    int i = i0 + y, j = j0 + x;
    tA[y][x] = (i < s && j < s) ? to_f<T>(Q[(size_t)i * s + j]) : 0.f;
    i = j0 + y; j = i0 + x;
    tB[y][x] = (i < s && j < s) ? to_f<T>(Q[(size_t)i * s + j]) : 0.f;
  }
  __syncthreads();
  const int w = tid >> 5, lane = tid & 31;
  float am = 0.f;
  for (int y = w; y < 64; y += 8) {   // tile (bi, bj): R[i0+y][j0+x] = Q[j0+x][i0+y] - Q[i0+y][j0+x]
    float rs = 0.f;
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      const int x = lane + 32 * hh;
      const int i = i0 + y, j = j0 + x;
      T o = from_f<T>(tB[x][y] - tA[y][x]);
      float f = to_f<T>(o);
      if (i < s && j < s) R[(size_t)i * s + j] = o; else f = 0.f;
      rs = fmaf(f, f, rs);
      am = fmaxf(am, fabsf(f));
    }
    rs = warp_sum(rs);
    if (lane == 0 && i0 + y < s && row_sumsq) atomicAdd(&row_sumsq[i0 + y], rs);
  }
  if (bi != bj) {
    for (int y = w; y < 64; y += 8) {  // mirror tile (bj, bi): R[j0+y][i0+x] = Q[i0+x][j0+y] - Q[j0+y][i0+x]
      float rs = 0.f;
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const int x = lane + 32 * hh;
        const int i = j0 + y, j = i0 + x;
        T o = from_f<T>(tA[x][y] - tB[y][x]);
        float f = to_f<T>(o);
        if (i < s && j < s) R[(size_t)i * s + j] = o; else f = 0.f;
        rs = fmaf(f, f, rs);
        am = fmaxf(am, fabsf(f));
      }
      rs = warp_sum(rs);
      if (lane == 0 && j0 + y < s && row_sumsq) atomicAdd(&row_sumsq[j0 + y], rs);
    }
  }
  am = block_max(am, red);
  if (tid == 0 && abs_max) atomic_max_nonneg(abs_max, am);
}

// bf16, s % 8 == 0: same tile-pair scheme with 16-byte global accesses (8 columns per thread)
__global__ void __launch_bounds__(256) k_skew_bf16x8(const bf16* __restrict__ Q, bf16* __restrict__ R, int s, float* abs_max, float* row_sumsq) {
  const int bi = blockIdx.y, bj = blockIdx.x;
  if (bi > bj) return;
  __shared__ float tA[64][65];  // Q[i0 + y][j0 + x]
  __shared__ float tB[64][65];  // Q[j0 + y][i0 + x]
  __shared__ float red[32];
  const int i0 = bi * 64, j0 = bj * 64, tid = threadIdx.x;
  const int cg = tid & 7, r0 = tid >> 3;   // column group (8 elements), row within a 32-row half
#pragma unroll
  for (int hh = 0; hh < 2; ++hh) {
    const int y = r0 + 32 * hh;
    float a[8], b[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) { a[e] = 0.f; b[e] = 0.f; }
    if (i0 + y < s && j0 + 8 * cg < s) ld8(Q + (size_t)(i0 + y) * s + j0 + 8 * cg, a);
    if (j0 + y < s && i0 + 8 * cg < s) ld8(Q + (size_t)(j0 + y) * s + i0 + 8 * cg, b);
#pragma unroll
    for (int e = 0; e < 8; ++e) { tA[y][8 * cg + e] = a[e]; tB[y][8 * cg + e] = b[e]; }
  }
  __syncthreads();
  float am = 0.f;
#pragma unroll
  for (int pass = 0; pass < 2; ++pass) {     // pass 0: tile (bi, bj); pass 1: its mirror (bj, bi)
    if (pass == 1 && bi == bj) break;
    const float(*S)[65] = pass == 0 ? tB : tA;   // transposed source
    const float(*D)[65] = pass == 0 ? tA : tB;   // direct source
    const int ri0 = pass == 0 ? i0 : j0, ci0 = pass == 0 ? j0 : i0;
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      const int y = r0 + 32 * hh;
      float o[8];
      float rs = 0.f;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float f = rbf(S[8 * cg + e][y] - D[y][8 * cg + e]);
        o[e] = f;
        rs = fmaf(f, f, rs);
        am = fmaxf(am, fabsf(f));
      }
      const bool ok = ri0 + y < s && ci0 + 8 * cg < s;
      if (ok) st8(R + (size_t)(ri0 + y) * s + ci0 + 8 * cg, o); else rs = 0.f;
      rs += __shfl_xor_sync(0xffffffffu, rs, 1); rs += __shfl_xor_sync(0xffffffffu, rs, 2); rs += __shfl_xor_sync(0xffffffffu, rs, 4);
      if (cg == 0 && ri0 + y < s && row_sumsq) atomicAdd(&row_sumsq[ri0 + y], rs);
    }
  }
  am = block_max(am, red);
  if (tid == 0 && abs_max) atomic_max_nonneg(abs_max, am);
}

// Fused psgd.py:58-63: j = argmax_i rowsumsq[i], nf = *nf_src + tiny, and the rotated probe V[p,:] = A[j,:]/nf + sgn(<A[j,:]/nf, V0[p,:]>) V0[p,:].
// grid = k (one block of 1024 threads per probe row); every block recomputes the (tiny) argmax, block 0 publishes nf, 1/nf and j.
template <typename T>
__global__ void __launch_bounds__(1024) k_probe_init(const T* __restrict__ A, int s, const T* __restrict__ V0, const float* __restrict__ row_sumsq,
                                                     const float* __restrict__ nf_src, float tiny, float* scal, T* __restrict__ V) {
  __shared__ float sv[32];
  __shared__ int si[32];
  __shared__ float red[32];
  __shared__ float sgn_s;
  __shared__ int j_s;
  float best = -1.f;
  int bi = 0x7fffffff;
  for (int i = threadIdx.x; i < s; i += blockDim.x) {
    float v = row_sumsq[i];
    if (v > best) { best = v; bi = i; }
  }
  for (int o = 16; o > 0; o >>= 1) {   // ties -> smallest index (torch.argmax returns the first maximal element)
    float ov = __shfl_xor_sync(0xffffffffu, best, o);
    int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
  }
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) { sv[w] = best; si[w] = bi; }
  __syncthreads();
  if (w == 0) {
    const int nw = (blockDim.x + 31) >> 5;
    best = lane < nw ? sv[lane] : -1.f;
    bi = lane < nw ? si[lane] : 0x7fffffff;
    for (int o = 16; o > 0; o >>= 1) {
      float ov = __shfl_xor_sync(0xffffffffu, best, o);
      int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    if (lane == 0) j_s = (bi == 0x7fffffff) ? 0 : bi;
  }
  __syncthreads();
  const int j = j_s;
  const float nf = *nf_src + tiny;
  const float inv_nf = 1.f / nf;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    scal[SC_NF] = nf;
    scal[SC_INV_NF] = inv_nf;
    reinterpret_cast<int*>(scal)[SC_J] = j;
  }
  const int p = blockIdx.x;
  const T* arow = A + (size_t)j * s;
  const T* v0 = V0 + (size_t)p * s;
  float dot = 0.f;
  for (int c = threadIdx.x; c < s; c += blockDim.x) {
    float a = to_f<T>(from_f<T>(to_f<T>(arow[c]) * inv_nf));
    dot += to_f<T>(from_f<T>(a * to_f<T>(v0[c])));
  }
  dot = block_sum(dot, red);
  if (threadIdx.x == 0) sgn_s = (dot > 0.f) ? 1.f : ((dot < 0.f) ? -1.f : 0.f);
  __syncthreads();
  const float sg = sgn_s;
  for (int c = threadIdx.x; c < s; c += blockDim.x) {
    float a = to_f<T>(from_f<T>(to_f<T>(arow[c]) * inv_nf));
    V[(size_t)p * s + c] = from_f<T>(a + sg * to_f<T>(v0[c]));
  }
}

// bound = nf * max_p sqrt(rn[p]) (psgd.py:68) fused with its consumer:
//   mode 0: dense-factor L update + step size (psgd.py:412-415): fs[FS_ALPHA] = -lr/L, fs[FS_BETA] = 1 + lr/L*t2
//   mode 1: procrustes normaliser (psgd.py:118): fs[FS_INV_SR] = 1 / (bound + tiny)
__global__ void k_bound_finish(const float* __restrict__ rn, int k, float* scal, int dtype, int mode, float t2, float lr, float betaL, float* L,
                               float* fs, float tiny) {
  float v = threadIdx.x < k ? rn[threadIdx.x] : 0.f;
  v = warp_max(v);
  if (threadIdx.x == 0) {
    const float bound = round_to(dtype, scal[SC_NF] * round_to(dtype, sqrtf(v)));
    scal[SC_BOUND] = bound;
    if (mode == 0) {
      const float ell = round_to(dtype, bound + t2);
      const float Ln = fmaxf(betaL * (*L) + (1.f - betaL) * ell, ell);
      *L = Ln;
      const float c = lr / Ln;
      fs[FS_ALPHA] = -c;
      fs[FS_BETA] = 1.f + c * t2;
    } else if (mode == 1) {
      fs[FS_INV_SR] = 1.f / (bound + tiny);
    }
  }
}

// diagonal factor (psgd.py:406-410): ell = max(term1) + t2; L = max(betaL*L+(1-betaL)*ell, ell);
// q *= 1 - lr/L*(term1 - t2).  term1 (fp32, length s) = sums of squares of Pg along the other axes.  One block.
template <typename T>
__global__ void k_diag_update(T* __restrict__ q, const float* __restrict__ term1, int s, float t2, float lr, float betaL,
                              float* L) {
  __shared__ float red[32];
  __shared__ float c_s;
  float mx = -INFINITY;
  for (int i = threadIdx.x; i < s; i += blockDim.x) mx = fmaxf(mx, to_f<T>(from_f<T>(term1[i])));
  mx = block_max(mx, red);
  if (threadIdx.x == 0) {
    float ell = to_f<T>(from_f<T>(mx + t2));
    float Ln = fmaxf(betaL * (*L) + (1.f - betaL) * ell, ell);
    *L = Ln;
    c_s = lr / Ln;
  }
  __syncthreads();
  const float c = c_s;
  for (int i = threadIdx.x; i < s; i += blockDim.x) {
    float t1 = to_f<T>(from_f<T>(term1[i]));
    q[i] = from_f<T>(to_f<T>(q[i]) * (1.f - c * (t1 - t2)));
  }
}

// batched k_diag_update: one block per unit
template <typename T>
__global__ void k_diag_update_multi(PtrTab q_, CPtrTab term1_, int s, float t2, float lr, float betaL, PtrTab L_) {
  __shared__ float red[32];
  __shared__ float c_s;
  T* q = reinterpret_cast<T*>(q_.p[blockIdx.x]);
  const float* term1 = reinterpret_cast<const float*>(term1_.p[blockIdx.x]);
  float* L = reinterpret_cast<float*>(L_.p[blockIdx.x]);
  float mx = -INFINITY;
  for (int i = threadIdx.x; i < s; i += blockDim.x) mx = fmaxf(mx, to_f<T>(from_f<T>(term1[i])));
  mx = block_max(mx, red);
  if (threadIdx.x == 0) {
    float ell = to_f<T>(from_f<T>(mx + t2));
    float Ln = fmaxf(betaL * (*L) + (1.f - betaL) * ell, ell);
    *L = Ln;
    c_s = lr / Ln;
  }
  __syncthreads();
  const float c = c_s;
  for (int i = threadIdx.x; i < s; i += blockDim.x) {
    float t1 = to_f<T>(from_f<T>(term1[i]));
    q[i] = from_f<T>(to_f<T>(q[i]) * (1.f - c * (t1 - t2)));
  }
}

// max |x| over a tensor (balance_kron_precond psgd.py:272)
template <typename T>
__global__ void k_absmax(const T* __restrict__ x, size_t numel, float* out) {
  __shared__ float red[32];
  float m = 0.f;
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < numel; i += stride) m = fmaxf(m, fabsf(to_f<T>(x[i])));
  m = block_max(m, red);
  if (threadIdx.x == 0) atomic_max_nonneg(out, m);
}

// q *= gmean / norm_self, gmean = sqrt(norm_a * norm_b)   psgd.py:273-275 (order-2 Q)
template <typename T>
__global__ void k_balance_scale(T* __restrict__ q, size_t numel, const float* __restrict__ norm_self,
                                const float* __restrict__ norm_other, int dtype) {
  // the factor gmean/norm is ~1.00x: rounding it to bf16 (as the reference's bf16 arithmetic does) quantises it to 2^-7 steps;
  // it is kept in fp32 here (closer to the exact balance, and Q_L (x) Q_R is then preserved to fp32 accuracy)
  (void)dtype;
  const float ns = *norm_self, no = *norm_other;
  const float f = sqrtf(ns * no) / ns;
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < numel; i += stride) q[i] = from_f<T>(to_f<T>(q[i]) * f);
}

// ------------------------------ KWNS4 glue (ddp.py:117-157) ------------------------------
// weight decay + cast + EMA, one pass.   TP: param dtype, TG: grad dtype, TQ: preconditioner dtype
template <typename TP, typename TG, typename TQ>
__global__ void k_kwns4_head(TP* __restrict__ p, const TG* __restrict__ grad, size_t numel, float wd, float lr_params,
                             int decoupled, TQ* __restrict__ ema, TQ* __restrict__ g_out, float beta) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < numel; i += stride) {
    float g = to_f<TG>(grad[i]);
    if (wd > 0.f) {
      if (decoupled) p[i] = from_f<TP>(to_f<TP>(p[i]) * (1.f - wd * lr_params));   // ddp.py:120
      else g = to_f<TG>(from_f<TG>(g + wd * to_f<TP>(p[i])));                     // ddp.py:122
    }
    TQ gq = from_f<TQ>(g);                                                         // ddp.py:127
    if (g_out) g_out[i] = gq;
    if (ema) {                                                                     // ddp.py:142
      float e = to_f<TQ>(from_f<TQ>(to_f<TQ>(ema[i]) * beta));
      ema[i] = from_f<TQ>(e + (1.f - beta) * to_f<TQ>(gq));
    }
  }
}

// clip + clamp + parameter update (ddp.py:153-157)
template <typename TP, typename TQ>
__global__ void k_kwns4_tail(TP* __restrict__ p, TQ* __restrict__ h, size_t numel, float amp_numel, const float* __restrict__ sumsq,
                             float max_avg_amp, float max_elem_amp, float lr_params, int h_dtype) {
  float avg_amp = round_to(h_dtype, sqrtf(round_to(h_dtype, *sumsq / amp_numel)));
  float scale = (avg_amp > max_avg_amp) ? round_to(h_dtype, max_avg_amp / avg_amp) : 1.f;
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < numel; i += stride) {
    float v = to_f<TQ>(h[i]);
    if (scale != 1.f) v = to_f<TQ>(from_f<TQ>(v * scale));
    v = fminf(fmaxf(v, -max_elem_amp), max_elem_amp);
    h[i] = from_f<TQ>(v);
    p[i] = from_f<TP>(to_f<TP>(p[i]) - lr_params * v);
  }
}

// ------------------------------ fp32 products on the bf16 tensor cores ------------------------------
// x = hi + mid + lo with three bf16 pieces is exact for an fp32 mantissa (3 x 8 bits).  A product A B is then the sum of the six piece
// products that matter (hi hi, hi mid, mid hi, hi lo, mid mid, lo hi: the dropped ones are <= 2^-24 relative), i.e. ONE bf16 GEMM over a
// six times longer K axis: A' = [A_s0 | A_s1 | ... | A_s5], B' = [B_t0; ...; B_t5] with fp32 accumulation in TMEM.  This kernel writes one
// operand: out holds six copies of the rows x cols matrix X, copy c being piece seq[c] of X, laid side by side (axis 1: rows x 6 cols)
// or stacked (axis 0: 6 rows x cols).
struct Split3Seq { int s[6]; };
__global__ void k_split3(const float* __restrict__ X, int rows, int cols, int ld, bf16* __restrict__ out, int out_ld, int axis, Split3Seq seq) {
  const size_t numel = (size_t)rows * cols;
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < numel; i += stride) {
    const int r = (int)(i / cols), c = (int)(i % cols);
    const float x = X[(size_t)r * ld + c];
    bf16 p[3];
    p[0] = __float2bfloat16_rn(x);
    const float r1 = x - __bfloat162float(p[0]);
    p[1] = __float2bfloat16_rn(r1);
    p[2] = __float2bfloat16_rn(r1 - __bfloat162float(p[1]));
#pragma unroll
    for (int t = 0; t < 6; ++t) {
      const size_t o = axis == 1 ? (size_t)r * out_ld + (size_t)t * cols + c : ((size_t)t * rows + r) * out_ld + c;
      out[o] = p[seq.s[t]];
    }
  }
}

// ------------------------------ the other Kron geometries (psgd.py:278-391, 422-513, 657-829) ------------------------------
// out[i] = float(q[i]) (op 0) or 1 / float(q[i]) (op 1): fp32 row / column factors for the GEMM epilogue (exprA with a diagonal factor,
// psgd.py:248-249; conjB / q, psgd.py:300)
template <typename T>
__global__ void k_vec_to_f32(const T* __restrict__ q, float* __restrict__ out, int n, int op) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { float v = to_f<T>(q[i]); out[i] = op == 1 ? 1.f / v : v; }
}

// S = T1 + T2 (rounded; row sums of squares and max diagonal fused: the inputs of norm_lower_bound_spd(term1 + term2), psgd.py:315 / 363 /
// 688 ...) and E = T1 - T2 (rounded; upper triangle only when triu: psgd.py:316).  S overwrites T1 and E overwrites T2.
// grid.x = s (one row per block)
template <typename T>
__global__ void __launch_bounds__(256) k_combine_terms(T* __restrict__ T1, T* __restrict__ T2, int s, int triu, float* row_sumsq, float* diag_max) {
  __shared__ float red[32];
  const int i = blockIdx.x;
  float ss = 0.f;
  for (int j = threadIdx.x; j < s; j += blockDim.x) {
    const size_t idx = (size_t)i * s + j;
    const float a = to_f<T>(T1[idx]), b = to_f<T>(T2[idx]);
    const T sm = from_f<T>(a + b);
    const float f = to_f<T>(sm);
    T1[idx] = sm;
    T2[idx] = (triu && j < i) ? from_f<T>(0.f) : from_f<T>(a - b);
    ss = fmaf(f, f, ss);
    if (j == i && diag_max) atomic_max_nonneg(diag_max, f);
  }
  ss = block_sum(ss, red);
  if (threadIdx.x == 0 && row_sumsq) row_sumsq[i] = ss;
}

// bf16, s % 8 == 0, 16-byte aligned: the same with 16-byte accesses
__global__ void __launch_bounds__(256) k_combine_terms_bf16x8(bf16* __restrict__ T1, bf16* __restrict__ T2, int s, int triu, float* row_sumsq,
                                                              float* diag_max) {
  __shared__ float red[32];
  const int i = blockIdx.x;
  float ss = 0.f;
  for (int j0 = threadIdx.x * 8; j0 < s; j0 += blockDim.x * 8) {
    const size_t idx = (size_t)i * s + j0;
    float a[8], b[8], sm[8], e[8];
    ld8(T1 + idx, a); ld8(T2 + idx, b);
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      sm[t] = rbf(a[t] + b[t]);
      e[t] = (triu && j0 + t < i) ? 0.f : a[t] - b[t];
      ss = fmaf(sm[t], sm[t], ss);
      if (j0 + t == i && diag_max) atomic_max_nonneg(diag_max, sm[t]);
    }
    st8(T1 + idx, sm); st8(T2 + idx, e);
  }
  ss = block_sum(ss, red);
  if (threadIdx.x == 0 && row_sumsq) row_sumsq[i] = ss;
}

// Q = (P + P^T) / 2    psgd.py:479 / 509 / 794 / 824.  grid (ceil(s/32), ceil(s/32)), block (32, 8)
template <typename T>
__global__ void k_symmetrize(const T* __restrict__ P, T* __restrict__ Q, int s) {
  __shared__ float tA[32][33];   // P[i0 + y][j0 + x]
  __shared__ float tB[32][33];   // P[j0 + y][i0 + x]
  const int i0 = blockIdx.y * 32, j0 = blockIdx.x * 32;
  for (int y = threadIdx.y; y < 32; y += 8) {
    const int x = threadIdx.x;
    tA[y][x] = (i0 + y < s && j0 + x < s) ? to_f<T>(P[(size_t)(i0 + y) * s + j0 + x]) : 0.f;
    tB[y][x] = (j0 + y < s && i0 + x < s) ? to_f<T>(P[(size_t)(j0 + y) * s + i0 + x]) : 0.f;
  }
  __syncthreads();
  for (int y = threadIdx.y; y < 32; y += 8) {
    const int x = threadIdx.x;
    if (i0 + y < s && j0 + x < s) Q[(size_t)(i0 + y) * s + j0 + x] = from_f<T>(to_f<T>(from_f<T>(tA[y][x] + tB[x][y])) * 0.5f);
  }
}

// diagonal factor, every geometry: term1 = t1[i] * (t1_q2 ? q_i^2 : 1), term2 = (t2v ? t2v[i] : t2s) * (t2_q2 ? q_i^2 : 1);
// ell = max(term1 + term2); L = max(betaL L + (1 - betaL) ell, ell); c = lr / L;
// quad == 0: q *= 1 - c (term1 - term2)   (psgd.py:312, 359, 383, 410, 437 ...);  quad == 1: q *= (1 - c (term1 - term2))^2  (psgd.py:471, 501)
template <typename T>
__global__ void k_diag_update_gen(T* __restrict__ q, const float* __restrict__ t1, const float* __restrict__ t2v, float t2s, int t1_q2, int t2_q2,
                                  int s, float lr, float betaL, float* L, int quad) {
  __shared__ float red[32];
  __shared__ float c_s;
  float mx = -INFINITY;
  for (int i = threadIdx.x; i < s; i += blockDim.x) {
    const float qv = to_f<T>(q[i]);
    const float q2 = to_f<T>(from_f<T>(qv * qv));
    const float a = to_f<T>(from_f<T>(t1[i] * (t1_q2 ? q2 : 1.f)));
    const float b = to_f<T>(from_f<T>((t2v ? t2v[i] : t2s) * (t2_q2 ? q2 : 1.f)));
    mx = fmaxf(mx, to_f<T>(from_f<T>(a + b)));
  }
  mx = block_max(mx, red);
  if (threadIdx.x == 0) {
    const float ell = mx;
    const float Ln = fmaxf(betaL * (*L) + (1.f - betaL) * ell, ell);
    *L = Ln;
    c_s = lr / Ln;
  }
  __syncthreads();
  const float c = c_s;
  for (int i = threadIdx.x; i < s; i += blockDim.x) {
    const float qv = to_f<T>(q[i]);
    const float q2 = to_f<T>(from_f<T>(qv * qv));
    const float a = to_f<T>(from_f<T>(t1[i] * (t1_q2 ? q2 : 1.f)));
    const float b = to_f<T>(from_f<T>((t2v ? t2v[i] : t2s) * (t2_q2 ? q2 : 1.f)));
    const float gain = 1.f - c * (a - b);
    q[i] = from_f<T>(qv * (quad ? gain * gain : gain));
  }
}

// Inverse of the upper-triangular diagonal blocks of Q (leaf of the blocked inversion behind the triangular solves of psgd.py:288-303).
// One CTA of 256 threads per TRI_NB x TRI_NB block, everything in shared memory (fp32):
//   level 0: the eight 16 x 16 diagonal micro-blocks by back substitution, one thread per column (columns are independent: no barriers);
//   levels b = 16, 32, 64: inv([[A11, A12], [0, A22]]) = [[X11, -X11 (A12 X22)], [0, X22]] as two small products per pair, all threads,
//            1 x 4 register tiles, the k ranges trimmed to the triangles.
// Output: Xf (fp32, s x s row-major), only the diagonal blocks are written.
constexpr int TRI_NB = 128;
constexpr int TRI_P = TRI_NB + 4;   // smem pitch (floats): rows stay 16-byte aligned
constexpr size_t TRI_LEAF_SMEM = (size_t)(2 * TRI_NB + TRI_NB / 2) * TRI_P * sizeof(float);
template <typename T>
__global__ void __launch_bounds__(256) k_tri_inv_leaf(const T* __restrict__ Q, int s, float* __restrict__ Xf) {
  extern __shared__ __align__(16) float tri_sm[];
  float* U = tri_sm;                     // TRI_NB x TRI_P: the block's upper triangle (identity beyond the matrix edge)
  float* X = U + TRI_NB * TRI_P;         // its inverse
  float* Tm = X + TRI_NB * TRI_P;        // TRI_NB/2 x TRI_P: A12 X22 of every pair of the current level
  const int o = blockIdx.x * TRI_NB;
  const int tid = threadIdx.x;
  for (int e = tid; e < TRI_NB * TRI_NB; e += 256) {
    const int r = e / TRI_NB, c = e % TRI_NB;
    float v = (r == c) ? 1.f : 0.f;
    if (o + r < s && o + c < s) v = (c >= r) ? to_f<T>(Q[(size_t)(o + r) * s + o + c]) : 0.f;
    U[r * TRI_P + c] = v;
    X[r * TRI_P + c] = 0.f;
  }
  __syncthreads();
  if (tid < TRI_NB) {   // level 0: column jj of micro-block blk
    const int base = (tid >> 4) * 16, jj = tid & 15;
    float x[16];
#pragma unroll
    for (int i = 15; i >= 0; --i) {
      float acc = (i == jj) ? 1.f : 0.f;
#pragma unroll
      for (int k = i + 1; k < 16; ++k) acc = fmaf(-U[(base + i) * TRI_P + base + k], x[k], acc);
      x[i] = (i <= jj) ? acc / U[(base + i) * TRI_P + base + i] : 0.f;
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) X[(base + i) * TRI_P + base + jj] = x[i];
  }
  __syncthreads();
  for (int b = 16; b < TRI_NB; b *= 2) {
    const int tpr = b / 4;                 // 1 x 4 tiles per row
    const int ntiles = (TRI_NB / 2) * tpr; // pairs * b rows = TRI_NB / 2
    for (int tile = tid; tile < ntiles; tile += 256) {      // Tm = A12 X22   (X22 upper triangular: k <= c)
      const int row = tile / tpr, c0 = (tile % tpr) * 4;
      const int p = row / b, r = row % b, r1 = 2 * p * b, c2 = r1 + b;
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
      const int kend = min(c0 + 3, b - 1);
      for (int k = 0; k <= kend; ++k) {
        const float a = U[(r1 + r) * TRI_P + c2 + k];
        const float4 xv = *reinterpret_cast<const float4*>(&X[(c2 + k) * TRI_P + c2 + c0]);
        a0 = fmaf(a, xv.x, a0); a1 = fmaf(a, xv.y, a1); a2 = fmaf(a, xv.z, a2); a3 = fmaf(a, xv.w, a3);
      }
      *reinterpret_cast<float4*>(&Tm[row * TRI_P + c0]) = make_float4(a0, a1, a2, a3);
    }
    __syncthreads();
    for (int tile = tid; tile < ntiles; tile += 256) {      // X12 = -X11 Tm   (X11 upper triangular: k >= r)
      const int row = tile / tpr, c0 = (tile % tpr) * 4;
      const int p = row / b, r = row % b, r1 = 2 * p * b, c2 = r1 + b;
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
      for (int k = r; k < b; ++k) {
        const float a = X[(r1 + r) * TRI_P + r1 + k];
        const float4 tv = *reinterpret_cast<const float4*>(&Tm[(p * b + k) * TRI_P + c0]);
        a0 = fmaf(a, tv.x, a0); a1 = fmaf(a, tv.y, a1); a2 = fmaf(a, tv.z, a2); a3 = fmaf(a, tv.w, a3);
      }
      *reinterpret_cast<float4*>(&X[(r1 + r) * TRI_P + c2 + c0]) = make_float4(-a0, -a1, -a2, -a3);
    }
    __syncthreads();
  }
  for (int e = tid; e < TRI_NB * TRI_NB; e += 256) {
    const int r = e / TRI_NB, c = e % TRI_NB;
    if (o + r < s && o + c < s) Xf[(size_t)(o + r) * s + o + c] = X[r * TRI_P + c];
  }
}

// One level of the blocked inversion in fp32 on CUDA cores, batched over the pairs (blockIdx.z = pair): C = alpha * A * B with
//   STAGE 1: A = A12 (a block of Q, dtype T), B = X22 (fp32, upper triangular: k < n0 + 64 suffices), C = Tm  (slot p, ld b)
//   STAGE 2: A = X11 (fp32, upper triangular: k >= m0), B = Tm, C = X12 (into the inverse), alpha = -1
// 64 x 64 tiles, 256 threads, 4 x 4 per thread.  Used for the low levels (b < TRI_TC_MIN_B: the tcgen05 kernel's launch latency dominates
// there) and for every level of fp32 factors.
template <typename T, int STAGE>
__global__ void __launch_bounds__(256) k_tri_pair_gemm(const T* __restrict__ Q, float* __restrict__ Xf, float* __restrict__ Tm, int s, int b) {
  __shared__ float As[16][68];
  __shared__ float Bs[16][68];
  const int p = blockIdx.z;
  const size_t r1 = (size_t)2 * p * b, c2 = r1 + b;
  const int b2 = min(b, s - (int)c2);                 // columns of this pair's off-diagonal block
  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  if (n0 >= b2) return;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  float* Tp = Tm + (size_t)p * b * b;
  const int K = STAGE == 1 ? b2 : b;
  const int kbeg = STAGE == 1 ? 0 : (m0 / 16) * 16;
  const int kend = STAGE == 1 ? min(K, n0 + 64) : K;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int k0 = kbeg; k0 < kend; k0 += 16) {
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int idx = tid + t * 256;
      {  // A tile: 64 rows x 16 k
        const int mi = idx >> 4, ki = idx & 15;
        const int gm = m0 + mi, gk = k0 + ki;
        float v = 0.f;
        if (gm < b && gk < K) v = STAGE == 1 ? to_f<T>(Q[(r1 + gm) * s + c2 + gk]) : Xf[(r1 + gm) * s + r1 + gk];
        As[ki][mi] = v;
      }
      {  // B tile: 16 k x 64 cols
        const int ki = idx >> 6, ni = idx & 63;
        const int gk = k0 + ki, gn = n0 + ni;
        float v = 0.f;
        if (gk < K && gn < b2) v = STAGE == 1 ? Xf[(c2 + gk) * s + c2 + gn] : Tp[(size_t)gk * b + gn];
        Bs[ki][ni] = v;
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      float a[4], bb[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) bb[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gm = m0 + ty * 4 + i;
    if (gm >= b) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gn = n0 + tx * 4 + j;
      if (gn >= b2) continue;
      if (STAGE == 1) Tp[(size_t)gm * b + gn] = acc[i][j];
      else Xf[(r1 + gm) * s + c2 + gn] = -acc[i][j];
    }
  }
}

// fp32 inverse -> bf16 hi + lo (the whole s x s matrix, zeros included).  8 elements per thread.
__global__ void k_split_full(const float* __restrict__ Xf, bf16* __restrict__ hi, bf16* __restrict__ lo, size_t numel) {
  size_t i = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) * 4;
  const size_t stride = (size_t)gridDim.x * blockDim.x * 4;
  for (; i + 3 < numel; i += stride) {
    const float4 v = *reinterpret_cast<const float4*>(Xf + i);
    const bf16 h0 = __float2bfloat16_rn(v.x), h1 = __float2bfloat16_rn(v.y), h2 = __float2bfloat16_rn(v.z), h3 = __float2bfloat16_rn(v.w);
    __nv_bfloat162 ha = __halves2bfloat162(h0, h1), hb = __halves2bfloat162(h2, h3);
    __nv_bfloat162 la = __floats2bfloat162_rn(v.x - __bfloat162float(h0), v.y - __bfloat162float(h1));
    __nv_bfloat162 lb = __floats2bfloat162_rn(v.z - __bfloat162float(h2), v.w - __bfloat162float(h3));
    *reinterpret_cast<uint2*>(hi + i) = make_uint2(*reinterpret_cast<uint32_t*>(&ha), *reinterpret_cast<uint32_t*>(&hb));
    *reinterpret_cast<uint2*>(lo + i) = make_uint2(*reinterpret_cast<uint32_t*>(&la), *reinterpret_cast<uint32_t*>(&lb));
  }
  if (blockIdx.x == 0 && threadIdx.x == 0)
    for (size_t t = numel & ~size_t(3); t < numel; ++t) {
      const bf16 h = __float2bfloat16_rn(Xf[t]);
      hi[t] = h; lo[t] = __float2bfloat16_rn(Xf[t] - __bfloat162float(h));
    }
}

// fp32 blocks -> bf16 hi / lo splits, batched over the pairs of one level of the blocked inversion.
// pair p: src = src0 + p * src_stride (rows x cols(p), ld src_ld), dst = hi0/lo0 + p * dst_stride (ld dst_ld);
// cols(p) = min(b, s - (2 p b + b)).  grid (ceil(rows*b / 256), pairs)
__global__ void k_split_hilo(const float* __restrict__ src0, size_t src_stride, int src_ld, bf16* __restrict__ hi0, bf16* __restrict__ lo0,
                             size_t dst_stride, int dst_ld, int rows, int b, int s) {
  const int p = blockIdx.y;
  const int cols = min(b, s - (2 * p * b + b));
  const float* src = src0 + (size_t)p * src_stride;
  bf16* hi = hi0 + (size_t)p * dst_stride;
  bf16* lo = lo0 + (size_t)p * dst_stride;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  const int r = e / b, c = e % b;
  if (r < rows && c < cols) {
    const float v = src[(size_t)r * src_ld + c];
    const bf16 h = __float2bfloat16_rn(v);
    hi[(size_t)r * dst_ld + c] = h;
    lo[(size_t)r * dst_ld + c] = __float2bfloat16_rn(v - __bfloat162float(h));
  }
}

// procrustes_step3 finish (psgd.py:149-155): if tr_RQ > 0 and tr_RRRQ < 0:
//   a = min((-tr2 - sqrt(tr2^2 - 1.5 tr1 tr3)) / (0.75 tr3), max_step);  Q = Qn + a (RQ + 0.5 a (RRQ + 0.25 a RRRQ));  else Q = Qn
template <typename T>
__global__ void k_procrustes3_finish(const T* __restrict__ Qn, const T* __restrict__ RQ, const T* __restrict__ RRQ, const T* __restrict__ RRRQ,
                                     T* __restrict__ Q, size_t numel, const float* __restrict__ fs, float max_step) {
  const float tr1 = fs[FS_TR1], tr2 = fs[FS_TR2], tr3 = fs[FS_TR3];
  const bool go = tr1 > 0.f && tr3 < 0.f;
  float a = 0.f;
  if (go) a = fminf((-tr2 - sqrtf(tr2 * tr2 - 1.5f * tr1 * tr3)) / (0.75f * tr3), max_step);
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < numel; i += stride) {
    float v = to_f<T>(Qn[i]);
    if (go) v += a * (to_f<T>(RQ[i]) + 0.5f * a * (to_f<T>(RRQ[i]) + 0.25f * a * to_f<T>(RRRQ[i])));
    Q[i] = from_f<T>(v);
  }
}

}  // namespace psgd
