// gemm_simt.cu -- generic fp32-FFMA GEMM with the fused PSGD epilogue.
// Used for (a) shapes the TMA/tcgen05 path cannot take (tiny, unaligned, e.g. the 2x2 / 10x10 plumbing
// configs), (b) fp32 preconditioners (exact fp32 products, 1e-5 parity with the reference's SGEMM), and
// (c) as the on-device cross-check of the tcgen05 kernel in tests.  Any shape, any transposition.
#include "common.cuh"

namespace psgd {

constexpr int SM_T = 64;  // tile M = tile N
constexpr int SM_K = 16;

template <typename T>
__global__ void __launch_bounds__(256) gemm_simt_kernel(const GemmDesc g) {
  __shared__ float As[SM_K][SM_T + 4];
  __shared__ float Bs[SM_K][SM_T + 4];
  const T* __restrict__ A = reinterpret_cast<const T*>(g.A);
  const T* __restrict__ B = reinterpret_cast<const T*>(g.B);
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * SM_T, n0 = blockIdx.x * SM_T;
  const int M = g.M, N = g.N, K = g.K;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < K; k0 += SM_K) {
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      int idx = tid + t * 256;
      int mi, ki;
      if (!g.ta) { mi = idx >> 4; ki = idx & 15; } else { ki = idx >> 6; mi = idx & 63; }
      int gm = m0 + mi, gk = k0 + ki;
      float v = 0.f;
      if (gm < M && gk < K) v = to_f<T>(g.ta ? A[(size_t)gk * g.lda + gm] : A[(size_t)gm * g.lda + gk]);
      As[ki][mi] = v;
      int ni;
      if (!g.tb) { ki = idx >> 6; ni = idx & 63; } else { ni = idx >> 4; ki = idx & 15; }
      int gn = n0 + ni;
      gk = k0 + ki;
      v = 0.f;
      if (gn < N && gk < K) v = to_f<T>(g.tb ? B[(size_t)gn * g.ldb + gk] : B[(size_t)gk * g.ldb + gn]);
      Bs[ki][ni] = v;
    }
    __syncthreads();
    // blocked summation: the 16 products of a k tile are summed on their own and then added to the running sum.  One long fp32 FMA
    // chain swamps small terms once a large one has entered the accumulator (Q ~ c I + tiny, early in training: Q^T Q lost 1e-5 at
    // s = 2048), which an MKL / cuBLAS-style blocked sum does not.
    float part[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) part[i][j] = 0.f;
#pragma unroll
    for (int kk = 0; kk < SM_K; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) part[i][j] = fmaf(a[i], b[j], part[i][j]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] += part[i][j];
    __syncthreads();
  }

  // ---- fused epilogue ----
  const Epi& e = g.epi;
  float alpha = e.alpha * (e.alpha_ptr ? *e.alpha_ptr : 1.f);
  float beta = e.beta * (e.beta_ptr ? *e.beta_ptr : 1.f);
  float beta2 = e.beta2;
  if (e.pro_fs) { const float a = epi_procrustes_step(e); alpha *= 0.5f * a * a; beta2 = a; }
  float csum[4] = {0.f, 0.f, 0.f, 0.f};
  float tot = 0.f, amax = 0.f, tr = 0.f, dmax = 0.f, dot = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int gm = m0 + ty * 4 + i;
    if (gm >= M) continue;
    float rs = e.row_scale ? e.row_scale[gm] : 1.f;
    if (e.norm_axis == 1) rs *= epi_norm_factor(e, gm);
    float rsum = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int gn = n0 + tx * 4 + j;
      if (gn >= N) continue;
      float v = acc[i][j] * alpha * rs;
      if (e.col_scale) v *= e.col_scale[gn];
      if (e.norm_axis == 2) v *= epi_norm_factor(e, gn);
      if (e.D) v += beta * ld_as_float(e.D, e.d_dtype, (size_t)gm * e.ldd + gn) * (e.d_row_scale ? e.d_row_scale[gm] : 1.f) * (e.d_col_scale ? e.d_col_scale[gn] : 1.f);
      if (e.D2) v += beta2 * ld_as_float(e.D2, e.d_dtype, (size_t)gm * e.ldd2 + gn);
      float r = round_to(e.out_dtype, v);
      if (e.dotm) dot = fmaf(r, ld_as_float(e.dotm, e.d_dtype, (size_t)gm * e.ld_dot + gn), dot);
      if (e.diag_resid && gm == gn) e.diag_resid[gm] = v - r;
      st_from_float(e.C, e.out_dtype, (size_t)gm * e.ldc + gn, v);
      rsum += r * r;
      csum[j] += r * r;
      amax = fmaxf(amax, fabsf(r));
      if (gm == gn) { tr += r; dmax = fmaxf(dmax, r); }
    }
    tot += rsum;
    if (e.row_sumsq) atomicAdd(&e.row_sumsq[gm], rsum);
  }
  if (e.col_sumsq) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int gn = n0 + tx * 4 + j;
      if (gn < N) atomicAdd(&e.col_sumsq[gn], csum[j]);
    }
  }
  if (e.total_sumsq) { float s = warp_sum(tot); if ((tid & 31) == 0) atomicAdd(e.total_sumsq, s); }
  if (e.dot_out) { float s = warp_sum(dot); if ((tid & 31) == 0 && s != 0.f) atomicAdd(e.dot_out, s); }
  if (e.abs_max) { float s = warp_max(amax); if ((tid & 31) == 0) atomic_max_nonneg(e.abs_max, s); }
  if (e.trace && m0 < n0 + SM_T && n0 < m0 + SM_T) { float s = warp_sum(tr); if ((tid & 31) == 0 && s != 0.f) atomicAdd(e.trace, s); }
  if (e.diag_max && m0 < n0 + SM_T && n0 < m0 + SM_T) { float s = warp_max(dmax); if ((tid & 31) == 0) atomic_max_nonneg(e.diag_max, s); }
}

int launch_gemm_simt(Ctx* ctx, const GemmDesc& g, cudaStream_t st) {
  if (g.M <= 0 || g.N <= 0) return PSGD_OK;
  dim3 grid((g.N + SM_T - 1) / SM_T, (g.M + SM_T - 1) / SM_T);
  if (g.in_dtype == PSGD_BF16)
    gemm_simt_kernel<bf16><<<grid, 256, 0, st>>>(g);
  else
    gemm_simt_kernel<float><<<grid, 256, 0, st>>>(g);
  ctx->launches++;
  return check_cuda(ctx, cudaGetLastError(), "gemm_simt");
}

}  // namespace psgd
