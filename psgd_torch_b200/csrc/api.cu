// api.cu -- C-ABI of libpsgd_b200.so (include/psgd_b200.h): context, workspace layout and the stream-ordered
// orchestration of one Kron update / apply out of the GEMM and bandwidth kernels.
#include <stdio.h>
#include <string.h>

#include <new>

#include <stdlib.h>
#include "common.cuh"
#include "kron_kernels.cuh"
#include "bounds.cuh"

namespace psgd {

static inline int ew_blocks_n(Ctx* ctx, size_t numel) {
  size_t b = (numel + 255) / 256, cap = (size_t)ctx->num_sms * 8;
  if (b > cap) b = cap;
  return (int)(b < 1 ? 1 : b);
}

int check_cuda(Ctx* ctx, cudaError_t e, const char* what) {
  if (e == cudaSuccess) return PSGD_OK;
  if (ctx) snprintf(ctx->last_error, sizeof(ctx->last_error), "%s: %s", what, cudaGetErrorString(e));
  return PSGD_ERR_CUDA;
}

// fp32 product on the tensor cores: both operands re-expressed as bf16 triples concatenated along K (k_split3), one tcgen05 GEMM of depth
// 6K with the caller's epilogue: 7.6x the CUDA-core kernel at 4096^3 (0.77 vs 5.8 ms), a whole 4096^2 fp32 update 77 -> 16 ms.
// The splits are exact and the dropped piece products are <= 2^-24, but the tensor core accumulates in fp32 with truncation, which
// biases long sums: measured 4e-6 per 4096-deep product on random data and 9e-6 through the chain of an apply at 2048^2, against 4e-7
// for the CUDA-core kernel.  That sits at north_star's 1e-5 for fp32 and grows with the depth, so -- like TF32 in torch -- it is OPT-IN
// (psgd_set_fp32_tensor_cores); the default fp32 path stays on the CUDA cores.  Returns PSGD_ERR_UNSUPPORTED when not applicable.
static int launch_gemm_f32x3(Ctx* ctx, const GemmDesc& g, cudaStream_t st) {
  if (g.in_dtype != PSGD_F32 || !ctx->fp32_tensor) return PSGD_ERR_UNSUPPORTED;
  if (g.M < 512 || g.N < 256 || g.K < 256 || (g.K % 8) || (g.M % 8) || (g.N % 8)) return PSGD_ERR_UNSUPPORTED;
  if ((double)g.M * g.N * g.K < 512.0 * 512.0 * 512.0) return PSGD_ERR_UNSUPPORTED;
  const int K6 = 6 * g.K;
  const size_t need[2] = {(size_t)g.M * K6 * 2, (size_t)g.N * K6 * 2};
  for (int i = 0; i < 2; ++i) {
    if (ctx->x3_cap[i] >= need[i]) continue;
    if (ctx->x3_buf[i]) { cudaStreamSynchronize(st); cudaFree(ctx->x3_buf[i]); ctx->x3_buf[i] = nullptr; ctx->x3_cap[i] = 0; }
    const size_t cap = need[i] + need[i] / 4;
    if (cudaMalloc(&ctx->x3_buf[i], cap) != cudaSuccess) { cudaGetLastError(); return PSGD_ERR_UNSUPPORTED; }
    ctx->x3_cap[i] = cap;
  }
  GemmDesc t = g;
  t.in_dtype = PSGD_BF16;
  t.K = K6;
  t.A = ctx->x3_buf[0]; t.B = ctx->x3_buf[1];
  t.lda = g.ta ? g.M : K6;     // ta: A stored K x M -> six copies stacked (6K x M); else M x K -> side by side (M x 6K)
  t.ldb = g.tb ? K6 : g.N;     // tb: B stored N x K -> side by side (N x 6K); else K x N -> stacked (6K x N)
  if (!tc_eligible(t)) return PSGD_ERR_UNSUPPORTED;
  // small terms first: (lo,hi) (hi,lo) (mid,mid) (mid,hi) (hi,mid) (hi,hi)
  const Split3Seq sa = {{2, 0, 1, 1, 0, 0}}, sb = {{0, 2, 1, 0, 1, 0}};
  const int a_rows = g.ta ? g.K : g.M, a_cols = g.ta ? g.M : g.K;
  const int b_rows = g.tb ? g.N : g.K, b_cols = g.tb ? g.K : g.N;
  k_split3<<<ew_blocks_n(ctx, (size_t)a_rows * a_cols), 256, 0, st>>>((const float*)g.A, a_rows, a_cols, g.lda, (bf16*)ctx->x3_buf[0], t.lda, g.ta ? 0 : 1, sa);
  ctx->launches++;
  k_split3<<<ew_blocks_n(ctx, (size_t)b_rows * b_cols), 256, 0, st>>>((const float*)g.B, b_rows, b_cols, g.ldb, (bf16*)ctx->x3_buf[1], t.ldb, g.tb ? 1 : 0, sb);
  ctx->launches++;
  int rc = check_cuda(ctx, cudaGetLastError(), "k_split3"); if (rc) return rc;
  return launch_gemm_tc_group(ctx, &t, 1, st);
}

int launch_gemm(Ctx* ctx, const GemmDesc& g, cudaStream_t st) {
  if (g.M <= 0 || g.N <= 0) return PSGD_OK;
  if (ctx->gemm_path == 1) return launch_gemm_simt(ctx, g, st);
  if (tc_eligible(g)) return launch_gemm_tc_group(ctx, &g, 1, st);
  {
    const int rc = launch_gemm_f32x3(ctx, g, st);
    if (rc != PSGD_ERR_UNSUPPORTED) return rc;
  }
  if (ctx->gemm_path == 2) return PSGD_ERR_UNSUPPORTED;
  return launch_gemm_simt(ctx, g, st);
}

// two independent GEMMs; one grouped tcgen05 launch when both qualify
int launch_gemm_pair(Ctx* ctx, const GemmDesc& g0, const GemmDesc& g1, cudaStream_t st) {
  if (ctx->gemm_path != 1 && tc_eligible(g0) && tc_eligible(g1)) {
    GemmDesc gs[2] = {g0, g1};
    return launch_gemm_tc_group(ctx, gs, 2, st);
  }
  int rc = launch_gemm(ctx, g0, st);
  if (rc) return rc;
  return launch_gemm(ctx, g1, st);
}

static GemmDesc gemm_desc(int dt, const void* A, int lda, int ta, const void* B, int ldb, int tb, int M, int N, int K,
                          void* C, int ldc) {
  GemmDesc g;
  g.A = A; g.B = B; g.lda = lda; g.ldb = ldb; g.ta = ta; g.tb = tb; g.M = M; g.N = N; g.K = K; g.in_dtype = dt; g.sym = 0;
  g.epi = make_epi(C, ldc, dt);
  return g;
}

static inline int ew_blocks(Ctx* ctx, size_t numel, int threads = 256) {
  size_t b = (numel + threads - 1) / threads;
  size_t cap = (size_t)ctx->num_sms * 8;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

#define LAUNCH_CHECK(ctx, what)                                  \
  do {                                                           \
    (ctx)->launches++;                                           \
    int rc__ = check_cuda((ctx), cudaGetLastError(), what);      \
    if (rc__) return rc__;                                       \
  } while (0)

#define DISPATCH_T(dt, ...)                      \
  do {                                           \
    if ((dt) == PSGD_BF16) { typedef bf16 T; __VA_ARGS__; } \
    else { typedef float T; __VA_ARGS__; }      \
  } while (0)

// ---------------------------------------------------------------------------------------------
// workspace layout
// ---------------------------------------------------------------------------------------------
struct Bump {
  char* base;
  size_t off;
  explicit Bump(void* b) : base(reinterpret_cast<char*>(b)), off(0) {}
  void* take(size_t bytes) {
    size_t o = off;
    off += (bytes + 255) & ~size_t(255);
    return base ? base + o : reinterpret_cast<void*>(o + 256);  // sizing mode: fake non-null pointers
  }
};

struct BoundWs {          // scratch of one norm_lower_bound_* evaluation
  float* scal;            // SC_COUNT
  float* rn1; float* rn3; float* rn4;  // 32 each
  float* sc1; float* sc3;              // 32 each
};

struct FactorWs {
  float* row_sumsq;   // s  (Gram row norms^2)
  float* diag_max;    // 1
  float* r_row_sumsq; // s  (R row norms^2)
  float* r_abs_max;   // 1
  float* fs;          // FS_COUNT
  BoundWs b_spd, b_skh;
  float* term1;       // s (diagonal factor: sums of squares)
};

struct KronWs {
  // zeroed region
  char* zero_begin; size_t zero_bytes;
  FactorWs f[2];
  float* bal;       // 2 floats (max|QL|, max|QR|)
  // buffers
  float* qsq[2];    // fp32 squares of diagonal factors
  float* presid[2]; // fp32 rounding residual of diag(P_L), diag(P_R)
  void* B0; void* B1; void* B2;   // m x n
  void* S[2][4];    // per dense factor, s x s each: T (Gram, later R), Qn, RQ, RRQ   (the Qn slots double as P_L / P_R in the chain)
  void* Va[2]; void* Vb[2];       // per dense factor: 32 x s probe buffers
  void* P0[2][2];                 // per dense factor: probe blocks drawn on the device (performance mode): [factor][spd, skh]
  size_t total;
};

static void layout_bound(Bump& b, BoundWs& w) {
  w.scal = (float*)b.take(SC_COUNT * 4);
  w.rn1 = (float*)b.take(32 * 4); w.rn3 = (float*)b.take(32 * 4); w.rn4 = (float*)b.take(32 * 4);
  w.sc1 = (float*)b.take(32 * 4); w.sc3 = (float*)b.take(32 * 4);
}

// zeroed part (reductions, scalars) and buffers of one unit; a batch lays out every unit's zeroed part first so that ONE memset clears them
static void layout_kron_zero(Bump& b, const psgd_kron_t* k, KronWs& w) {
  const int sdim[2] = {k->m, k->has_r ? k->n : 0};
  w.zero_begin = (char*)b.take(0);
  size_t z0 = b.off;
  for (int i = 0; i < 2; ++i) {
    FactorWs& f = w.f[i];
    size_t s = sdim[i] > 0 ? sdim[i] : 1;
    f.row_sumsq = (float*)b.take(s * 4);
    f.diag_max = (float*)b.take(4);
    f.r_row_sumsq = (float*)b.take(s * 4);
    f.r_abs_max = (float*)b.take(4);
    f.fs = (float*)b.take(FS_COUNT * 4);
    layout_bound(b, f.b_spd);
    layout_bound(b, f.b_skh);
    f.term1 = (float*)b.take(s * 4);
  }
  w.bal = (float*)b.take(2 * 4);
  w.zero_bytes = b.off - z0;
}

static void layout_kron_bufs(Bump& b, const psgd_kron_t* k, KronWs& w) {
  const int es = dtype_size(k->dtype);
  const size_t m = k->m, n = k->has_r ? k->n : 1;
  const int sdim[2] = {k->m, k->has_r ? k->n : 0};
  const int dense[2] = {k->kind_l == PSGD_DENSE, k->has_r && k->kind_r == PSGD_DENSE};
  w.qsq[0] = (float*)b.take(m * 4);
  w.qsq[1] = (float*)b.take(n * 4);
  w.presid[0] = (float*)b.take(m * 4);
  w.presid[1] = (float*)b.take(n * 4);
  w.B0 = b.take(m * n * es); w.B1 = b.take(m * n * es); w.B2 = b.take(m * n * es);
  for (int i = 0; i < 2; ++i) {
    const size_t sd = dense[i] ? (size_t)sdim[i] : 0;
    for (int j = 0; j < 4; ++j) w.S[i][j] = b.take(sd * sd * es);
    w.Va[i] = b.take(32 * sd * es); w.Vb[i] = b.take(32 * sd * es);
    w.P0[i][0] = b.take(32 * sd * es); w.P0[i][1] = b.take(32 * sd * es);
  }
}

static void layout_kron(const psgd_kron_t* k, void* base, KronWs& w) {
  Bump b(base);
  layout_kron_zero(b, k, w);
  layout_kron_bufs(b, k, w);
  w.total = b.off;
}

// n units: [zeroed parts of all units][buffers of unit 0][buffers of unit 1]...
static size_t layout_kron_batch(const psgd_kron_t* ks, int n, void* base, KronWs* w, char** zero_begin, size_t* zero_bytes) {
  Bump b(base);
  char* z0p = (char*)b.take(0);
  const size_t z0 = b.off;
  for (int u = 0; u < n; ++u) layout_kron_zero(b, ks + u, w[u]);
  if (zero_begin) *zero_begin = z0p;
  if (zero_bytes) *zero_bytes = b.off - z0;
  for (int u = 0; u < n; ++u) { layout_kron_bufs(b, ks + u, w[u]); w[u].total = b.off; }
  return b.off;
}

// any number of independent GEMMs: runs of tcgen05-eligible problems share grouped launches (TC_GROUP_MAX per launch), the rest go one by one
constexpr int TC_GROUP_MAX = 4;
static int launch_gemm_group(Ctx* ctx, const GemmDesc* gs, int n, cudaStream_t st) {
  int i = 0;
  while (i < n) {
    if (gs[i].M <= 0 || gs[i].N <= 0) { ++i; continue; }
    if (ctx->gemm_path != 1 && tc_eligible(gs[i])) {
      int j = i + 1;
      while (j < n && j - i < TC_GROUP_MAX && gs[j].M > 0 && gs[j].N > 0 && tc_eligible(gs[j])) ++j;
      int rc = launch_gemm_tc_group(ctx, gs + i, j - i, st); if (rc) return rc;
      i = j;
    } else {
      int rc = launch_gemm(ctx, gs[i], st); if (rc) return rc;
      ++i;
    }
  }
  return PSGD_OK;
}

// ---------------------------------------------------------------------------------------------
// norm_lower_bound_{spd,skh}  (psgd.py:46-93) for up to two matrices at once (the left and right factor of one update
// advance together so that their 32-probe products share one grouped launch).  row_sumsq / nf_src already reduced.
// ---------------------------------------------------------------------------------------------
struct BoundJob {
  const void* A; int s; const void* V0; const float* row_sumsq; const float* nf_src; BoundWs* w; void* Va; void* Vb;
};

// finish: mode 0 -> dense-factor L update + step (needs t2, lr, betaL, L, fs); mode 1 -> procrustes normaliser (fs); mode 2 -> bound only
struct BoundFinish { int mode; float t2, lr, betaL; float* L; float* fs; };

// the whole evaluation of up to NB_MAX_JOBS bounds as one persistent cooperative kernel (bounds.cuh)
static bool bound_fusable(const Ctx* ctx, int dt, const BoundJob& J) {
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
  return ctx->nb_sync && ctx->gemm_path != 1 && !(ctx->debug_flags & (4 | 256)) && dt == PSGD_BF16 && J.s >= 128 && (J.s % 8) == 0 &&
         J.s <= NB_MAX_S && al16(J.A) && al16(J.V0) && al16(J.Va) && al16(J.Vb);
}

static int run_bounds_fused(Ctx* ctx, int dt, const BoundJob* jb, int n, const BoundFinish* fin, cudaStream_t st) {
  NbParams P;
  memset(&P, 0, sizeof(P));
  int units = 0, smax = 0;
  for (int j = 0; j < n; ++j) {
    const BoundJob& J = jb[j];
    NbJob& o = P.job[j];
    o.A = (const bf16*)J.A; o.V0 = (const bf16*)J.V0; o.row_sumsq = J.row_sumsq; o.nf_src = J.nf_src;
    o.Va = (bf16*)J.Va; o.Vb = (bf16*)J.Vb; o.scal = J.w->scal; o.rn1 = J.w->rn1; o.rn3 = J.w->rn3; o.rn4 = J.w->rn4; o.dots = J.w->sc1;
    o.s = J.s; o.unit0 = units; o.nunits = (J.s + NB_W - 1) / NB_W;
    units += o.nunits;
    if (J.s > smax) smax = J.s;
    const BoundFinish f = fin ? fin[j] : BoundFinish{2, 0.f, 0.f, 0.f, nullptr, nullptr};
    o.mode = f.mode; o.t2 = f.t2; o.lr = f.lr; o.betaL = f.betaL; o.L = f.L; o.fs = f.fs;
  }
  P.njobs = n; P.total_units = units; P.dtype = dt; P.tiny = dtype_tiny(dt);
  P.barrier = ctx->nb_sync; P.done = ctx->nb_sync + 1;
  const int smem = NB_RING_BYTES + ((smax + NB_KC - 1) / NB_KC) * NB_KC * 2;
  static PerDeviceOnce attr;
  if (attr.need(ctx->device)) {
    cudaError_t e = cudaFuncSetAttribute(k_norm_bounds, cudaFuncAttributeMaxDynamicSharedMemorySize, NB_RING_BYTES + NB_MAX_S * 2);
    if (e != cudaSuccess) return check_cuda(ctx, e, "cudaFuncSetAttribute(k_norm_bounds)");
  }
  const int grid = units < ctx->num_sms ? units : ctx->num_sms;
  void* args[1] = {&P};
  cudaError_t e = cudaLaunchCooperativeKernel((const void*)k_norm_bounds, dim3(grid), dim3(NB_THREADS), args, (size_t)smem, st);
  ctx->launches++;
  return check_cuda(ctx, e == cudaSuccess ? cudaGetLastError() : e, "k_norm_bounds");
}

// tcgen05 form (bounds.cuh: k_norm_bounds_tc): matrices whose size is a multiple of 64
static bool bound_fusable_tc(const Ctx* ctx, int dt, const BoundJob& J) {
  return bound_fusable(ctx, dt, J) && (J.s % 64) == 0 && ctx->encode_tiled && !(ctx->debug_flags & 512);
}

static int run_bounds_fused_tc(Ctx* ctx, int dt, const BoundJob* jb, int n, const BoundFinish* fin, cudaStream_t st) {
  NbTcParams P;
  memset(&P, 0, sizeof(P));
  int units = 0;
  for (int j = 0; j < n; ++j) {
    const BoundJob& J = jb[j];
    NbTcJob& T = P.job[j];
    int rc = make_tmap_mn3(ctx, &T.map_a, J.A, J.s, J.s, J.s, NT_BM / 64, NT_BK); if (rc) return rc;
    rc = make_tmap(ctx, &T.map_va, J.Va, 32, J.s, J.s, 32); if (rc) return rc;
    rc = make_tmap(ctx, &T.map_vb, J.Vb, 32, J.s, J.s, 32); if (rc) return rc;
    NbJob& o = T.j;
    o.A = (const bf16*)J.A; o.V0 = (const bf16*)J.V0; o.row_sumsq = J.row_sumsq; o.nf_src = J.nf_src;
    o.Va = (bf16*)J.Va; o.Vb = (bf16*)J.Vb; o.scal = J.w->scal; o.rn1 = J.w->rn1; o.rn3 = J.w->rn3; o.rn4 = J.w->rn4; o.dots = J.w->sc1;
    o.s = J.s; o.unit0 = units; o.nunits = (J.s + NT_BM - 1) / NT_BM;
    units += o.nunits;
    const BoundFinish f = fin ? fin[j] : BoundFinish{2, 0.f, 0.f, 0.f, nullptr, nullptr};
    o.mode = f.mode; o.t2 = f.t2; o.lr = f.lr; o.betaL = f.betaL; o.L = f.L; o.fs = f.fs;
  }
  P.njobs = n; P.total_units = units; P.dtype = dt; P.tiny = dtype_tiny(dt);
  P.barrier = ctx->nb_sync; P.done = ctx->nb_sync + 1;
  P.mn_lbo = ctx->mn_lbo; P.mn_sbo = ctx->mn_sbo;
  static PerDeviceOnce attr;
  if (attr.need(ctx->device)) {
    cudaError_t e = cudaFuncSetAttribute(k_norm_bounds_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, NT_SMEM_BYTES);
    if (e != cudaSuccess) return check_cuda(ctx, e, "cudaFuncSetAttribute(k_norm_bounds_tc)");
  }
  const int grid = units < ctx->num_sms ? units : ctx->num_sms;
  void* args[1] = {&P};
  cudaError_t e = cudaLaunchCooperativeKernel((const void*)k_norm_bounds_tc, dim3(grid), dim3(NT_THREADS), args, (size_t)NT_SMEM_BYTES, st);
  ctx->launches++;
  return check_cuda(ctx, e == cudaSuccess ? cudaGetLastError() : e, "k_norm_bounds_tc");
}

static int run_bounds_unfused(Ctx* ctx, int dt, BoundJob* jb, int n, const BoundFinish* fin, cudaStream_t st);

static int run_bounds(Ctx* ctx, int dt, BoundJob* jb, int n, const BoundFinish* fin, cudaStream_t st) {
  BoundJob fj[NB_MAX_JOBS], tj[NT_MAX_JOBS];
  BoundFinish ff[NB_MAX_JOBS], tf[NT_MAX_JOBS];
  int nf = 0, nt = 0;
  for (int j = 0; j < n; ++j) {
    const BoundFinish f = fin ? fin[j] : BoundFinish{2, 0.f, 0.f, 0.f, nullptr, nullptr};
    if (bound_fusable_tc(ctx, dt, jb[j])) {
      tj[nt] = jb[j]; tf[nt++] = f;
      if (nt == NT_MAX_JOBS) { int rc = run_bounds_fused_tc(ctx, dt, tj, nt, tf, st); if (rc) return rc; nt = 0; }
    } else if (bound_fusable(ctx, dt, jb[j])) {
      fj[nf] = jb[j]; ff[nf++] = f;
      if (nf == NB_MAX_JOBS) { int rc = run_bounds_fused(ctx, dt, fj, nf, ff, st); if (rc) return rc; nf = 0; }
    } else {
      BoundJob one = jb[j];
      int rc = run_bounds_unfused(ctx, dt, &one, 1, &f, st); if (rc) return rc;
    }
  }
  if (nt) { int rc = run_bounds_fused_tc(ctx, dt, tj, nt, tf, st); if (rc) return rc; }
  if (nf) { int rc = run_bounds_fused(ctx, dt, fj, nf, ff, st); if (rc) return rc; }
  return PSGD_OK;
}

// round-1 form: probe initialisation, four grouped GEMM launches, finish -- kept for fp32 / small / unaligned matrices (SIMT products)
static int run_bounds_unfused(Ctx* ctx, int dt, BoundJob* jb, int n, const BoundFinish* fin, cudaStream_t st) {
  const float tiny = dtype_tiny(dt);
  bool tc_form[2];
  for (int j = 0; j < n; ++j) {
    tc_form[j] = ctx->gemm_path != 1 && !(ctx->debug_flags & 4) && dt == PSGD_BF16 && jb[j].s >= 128 && (jb[j].s % 8) == 0;
    DISPATCH_T(dt, (k_probe_init<T><<<32, 1024, 0, st>>>((const T*)jb[j].A, jb[j].s, (const T*)jb[j].V0, jb[j].row_sumsq, jb[j].nf_src, tiny,
                                                          jb[j].w->scal, (T*)jb[j].Va)));
    LAUNCH_CHECK(ctx, "k_probe_init");
  }
  GemmDesc g[2];
  // four products W <- W A / nf with a row normalisation after the 1st and 3rd (psgd.py:64-67).
  // tensor-core formulation: the probes stay transposed, X = W^T (s x 32), X_new = A^T X: the s-long dimension is the UMMA M
  // dimension, the 32 probes ride in N, A is read through MN-major descriptors (no transposed copy); probe norms are column
  // sums of squares and the normalisation a column factor computed in the consumer's epilogue from those sums (no glue kernel).
  // SIMT formulation (small / fp32): W (32 x s) row-major as written.
  for (int step = 0; step < 4; ++step) {
    for (int j = 0; j < n; ++j) {
      const BoundJob& J = jb[j];
      BoundWs& w = *J.w;
      void* src = (step & 1) ? J.Vb : J.Va;
      void* dst = (step & 1) ? J.Va : J.Vb;
      const int s = J.s;
      const int axis = tc_form[j] ? 2 : 1;
      if (tc_form[j]) {
        if (step == 0) g[j] = gemm_desc(dt, J.A, s, 1, src, s, 1, s, 32, s, dst, 32);   // V1 is stored 32 x s
        else g[j] = gemm_desc(dt, J.A, s, 1, src, 32, 0, s, 32, s, dst, 32);
      } else {
        g[j] = gemm_desc(dt, src, s, 0, J.A, s, 0, 32, s, s, dst, s);
      }
      Epi& e = g[j].epi;
      if (step == 0 || step == 2) {
        e.alpha_ptr = w.scal + SC_INV_NF;
      } else {
        e.norm_axis = axis; e.norm_sumsq = (step == 1) ? w.rn1 : w.rn3; e.norm_inv_nf = w.scal + SC_INV_NF; e.norm_tiny = tiny;
      }
      float* sums = step == 0 ? w.rn1 : (step == 2 ? w.rn3 : (step == 3 ? w.rn4 : nullptr));
      if (sums) { if (tc_form[j]) e.col_sumsq = sums; else e.row_sumsq = sums; }
    }
    int rc = launch_gemm_group(ctx, g, n, st); if (rc) return rc;
  }
  for (int j = 0; j < n; ++j) {
    const BoundFinish f = fin ? fin[j] : BoundFinish{2, 0.f, 0.f, 0.f, nullptr, nullptr};
    k_bound_finish<<<1, 32, 0, st>>>(jb[j].w->rn4, 32, jb[j].w->scal, dt, f.mode, f.t2, f.lr, f.betaL, f.L, f.fs, tiny);
    LAUNCH_CHECK(ctx, "k_bound_finish");
  }
  return PSGD_OK;
}

// one dense factor travelling through the update
struct DenseItem {
  int s; void* q; float* L; float t2;
  void* T;     // Gram term1 on entry; reused for R = Qn^T - Qn
  void* Qn; void* RQ; void* RRQ; void* Va; void* Vb;
  const void* v_spd; const void* v_skh;
  FactorWs* f;
};

// procrustes_step2 (psgd.py:101-124) on it[i].Qn -> writes it[i].q
constexpr int KB_ITEMS = 2 * KB_MAX;   // dense factors travelling together through one batched update
static int run_procrustes(Ctx* ctx, int dt, DenseItem* it, int n, float max_step, cudaStream_t st) {
  if (n > KB_ITEMS) return PSGD_ERR_INVALID_ARG;
  BoundJob jb[KB_ITEMS];
  GemmDesc g[KB_ITEMS];
  for (int i = 0; i < n; ++i) {
    const int s = it[i].s;
    dim3 grid((s + 63) / 64, (s + 63) / 64);
    if (dt == PSGD_BF16 && s % 8 == 0 && (reinterpret_cast<uintptr_t>(it[i].Qn) & 15u) == 0 && (reinterpret_cast<uintptr_t>(it[i].T) & 15u) == 0) {
      k_skew_bf16x8<<<grid, 256, 0, st>>>((const bf16*)it[i].Qn, (bf16*)it[i].T, s, it[i].f->r_abs_max, it[i].f->r_row_sumsq);
    } else {
      DISPATCH_T(dt, (k_skew<T><<<grid, 256, 0, st>>>((const T*)it[i].Qn, (T*)it[i].T, s, it[i].f->r_abs_max, it[i].f->r_row_sumsq)));
    }
    LAUNCH_CHECK(ctx, "k_skew");
    jb[i] = BoundJob{it[i].T, s, it[i].v_skh, it[i].f->r_row_sumsq, it[i].f->r_abs_max, &it[i].f->b_skh, it[i].Va, it[i].Vb};
  }
  BoundFinish fin[KB_ITEMS];
  for (int i = 0; i < n; ++i) fin[i] = BoundFinish{1, 0.f, 0.f, 0.f, nullptr, it[i].f->fs};
  int rc = run_bounds(ctx, dt, jb, n, fin, st); if (rc) return rc;
  for (int i = 0; i < n; ++i) {
    // RQ = R Qn / |R| with tr(RQ) and <RQ, R> (= -tr(R RQ): R is skew) from the epilogue    psgd.py:119-120
    g[i] = gemm_desc(dt, it[i].T, it[i].s, 0, it[i].Qn, it[i].s, 0, it[i].s, it[i].s, it[i].s, it[i].RQ, it[i].s);
    g[i].epi.alpha_ptr = it[i].f->fs + FS_INV_SR; g[i].epi.trace = it[i].f->fs + FS_TR1;
    g[i].epi.dotm = it[i].T; g[i].epi.ld_dot = it[i].s; g[i].epi.dot_out = it[i].f->fs + FS_DOT;
  }
  rc = launch_gemm_group(ctx, g, n, st); if (rc) return rc;
  for (int i = 0; i < n; ++i) {
    // Q = Qn + a (RQ + a/2 RRQ), RRQ = R RQ / |R| never materialised: the step length a follows from the two traces the previous
    // product left on the device (psgd.py:121-124; Epi::pro_fs)
    g[i] = gemm_desc(dt, it[i].T, it[i].s, 0, it[i].RQ, it[i].s, 0, it[i].s, it[i].s, it[i].s, it[i].q, it[i].s);
    g[i].epi.alpha_ptr = it[i].f->fs + FS_INV_SR; g[i].epi.pro_fs = it[i].f->fs; g[i].epi.pro_max_step = max_step;
    g[i].epi.D = it[i].Qn; g[i].epi.ldd = it[i].s; g[i].epi.d_dtype = dt; g[i].epi.beta = 1.f;
    g[i].epi.D2 = it[i].RQ; g[i].epi.ldd2 = it[i].s;
  }
  return launch_gemm_group(ctx, g, n, st);
}

// psgd.py:412-416 for up to two dense factors: bound of term1, L update, Newton-Schulz step, procrustes
static int run_dense_factors(Ctx* ctx, int dt, DenseItem* it, int n, float lr, float betaL, cudaStream_t st) {
  if (n <= 0) return PSGD_OK;
  if (n > KB_ITEMS) return PSGD_ERR_INVALID_ARG;
  BoundJob jb[KB_ITEMS];
  GemmDesc g[KB_ITEMS];
  for (int i = 0; i < n; ++i)
    jb[i] = BoundJob{it[i].T, it[i].s, it[i].v_spd, it[i].f->row_sumsq, it[i].f->diag_max, &it[i].f->b_spd, it[i].Va, it[i].Vb};
  BoundFinish fin[KB_ITEMS];
  for (int i = 0; i < n; ++i) fin[i] = BoundFinish{0, it[i].t2, lr, betaL, it[i].L, it[i].f->fs};
  int rc = run_bounds(ctx, dt, jb, n, fin, st); if (rc) return rc;
  for (int i = 0; i < n; ++i) {
    // Qn = Q - lr/L (term1 Q - t2 Q)    psgd.py:415   (out of place: Q is an operand of the product)
    g[i] = gemm_desc(dt, it[i].T, it[i].s, 0, it[i].q, it[i].s, 0, it[i].s, it[i].s, it[i].s, it[i].Qn, it[i].s);
    g[i].epi.alpha_ptr = it[i].f->fs + FS_ALPHA; g[i].epi.D = it[i].q; g[i].epi.ldd = it[i].s; g[i].epi.d_dtype = dt; g[i].epi.beta = 1.f;
    g[i].epi.beta_ptr = it[i].f->fs + FS_BETA;
  }
  rc = launch_gemm_group(ctx, g, n, st); if (rc) return rc;
  return run_procrustes(ctx, dt, it, n, 0.125f, st);
}

// ---------------------------------------------------------------------------------------------
// out = (Q_L^T Q_L) X (Q_R^T Q_R)   psgd.py:322-327 / 403, min-flop contraction order
// reductions (on out): row_sumsq / col_sumsq / total_sumsq, any may be null
// ---------------------------------------------------------------------------------------------
// A batched call records each unit's products level by level (products of one level are independent across units and share grouped
// launches); plan == nullptr launches at once.  Levels: 0 symmetric products P = Q^T Q, 1-2 left side, 3-4 right side.
struct ChainPlan {
  GemmDesc g[5][2];
  int cnt[5];
  bool squares_done;   // the batch driver has already formed the fp32 squares of the diagonal factors (k_square_to_f32_multi)
};

static int chain_emit(Ctx* ctx, ChainPlan* plan, int level, const GemmDesc& g, cudaStream_t st) {
  if (!plan) return launch_gemm(ctx, g, st);
  plan->g[level][plan->cnt[level]++] = g;
  return PSGD_OK;
}

static int run_chain(Ctx* ctx, const psgd_kron_t* k, KronWs& w, const void* X, void* out, float* row_sumsq, float* col_sumsq,
                     float* total_sumsq, cudaStream_t st, ChainPlan* plan = nullptr) {
  const int dt = k->dtype;
  const int m = k->m, n = k->has_r ? k->n : 1;
  const bool dl = k->kind_l == PSGD_DENSE, dr = k->has_r && k->kind_r == PSGD_DENSE;
  int rc;
  if (!(plan && plan->squares_done)) {
    if (!dl) { DISPATCH_T(dt, (k_square_to_f32<T><<<(m + 255) / 256, 256, 0, st>>>((const T*)k->QL, w.qsq[0], m))); LAUNCH_CHECK(ctx, "k_square"); }
    if (k->has_r && !dr) { DISPATCH_T(dt, (k_square_to_f32<T><<<(n + 255) / 256, 256, 0, st>>>((const T*)k->QR, w.qsq[1], n))); LAUNCH_CHECK(ctx, "k_square"); }
  }
  const float* rs = dl ? nullptr : w.qsq[0];
  const float* cs = (!k->has_r || dr) ? nullptr : w.qsq[1];
  if (!dl && !dr) {
    if (plan) return PSGD_ERR_INVALID_ARG;   // the batch driver handles all-diagonal units itself (k_scale2d_multi)
    size_t numel = (size_t)m * n;
    DISPATCH_T(dt, (k_scale2d<T><<<ew_blocks(ctx, numel), 256, 0, st>>>((const T*)X, (T*)out, m, n, rs, cs, row_sumsq, col_sumsq,
                                                                       total_sumsq)));
    LAUNCH_CHECK(ctx, "k_scale2d");
    return PSGD_OK;
  }
  void* PL = w.S[0][1];
  void* PR = w.S[1][1];
  // ---- left side ----
  // scratch: B0 and B2 (out is B1 or caller memory, X is B0 or caller memory).  X is dead once the first product
  // of the chain form has consumed it, so B0 may be overwritten afterwards; the P-first form reads X last.
  const void* Y = X;  // result of the left stage
  GemmDesc g;
  auto set_final = [&](GemmDesc& gd) {
    gd.epi.row_scale = rs; gd.epi.col_scale = cs;
    gd.epi.row_sumsq = row_sumsq; gd.epi.col_sumsq = col_sumsq; gd.epi.total_sumsq = total_sumsq;
    if (gd.epi.D) {  // the residual term resid * X must carry the diagonal factor's q^2 scaling too
      if (rs && !gd.epi.d_row_scale) gd.epi.d_row_scale = rs;
      if (cs && !gd.epi.d_col_scale) gd.epi.d_col_scale = cs;
    }
  };
  // P-first (P = Q^T Q is symmetric: the tensor-core path computes its upper 128-blocks only, ~1.06 s^3 instead of 2 s^3) costs
  // 1.06 m^3 + 2 m^2 n on the left against 4 m^2 n for the chain Q_L^T (Q_L X): take it whenever m < 1.88 n (same on the right).
  // bf16 only: P = Q^T Q in fp32 has no exact-diagonal correction and nothing to gain (the fp32 kernels compute symmetric products in
  // full), while forming P squares Q's dynamic range -- measured 2e-5 instead of 2e-6 on the apply at s = 2048 while Q is still ~ c I
  const bool pl = dl && (double)m < 1.88 * (double)n && !(ctx->debug_flags & 2) && dt == PSGD_BF16;
  const bool pr = dr && (double)n < 1.88 * (double)m && !(ctx->debug_flags & 2) && dt == PSGD_BF16;
  {
    GemmDesc sy[2];
    int ns = 0;
    if (pl) { sy[ns] = gemm_desc(dt, k->QL, m, 1, k->QL, m, 0, m, m, m, PL, m); sy[ns].sym = 1; sy[ns].epi.diag_resid = w.presid[0]; ++ns; }
    if (pr) { sy[ns] = gemm_desc(dt, k->QR, n, 1, k->QR, n, 0, n, n, n, PR, n); sy[ns].sym = 1; sy[ns].epi.diag_resid = w.presid[1]; ++ns; }
    if (plan) { for (int i = 0; i < ns; ++i) plan->g[0][plan->cnt[0]++] = sy[i]; }
    else { rc = launch_gemm_group(ctx, sy, ns, st); if (rc) return rc; }
  }
  if (dl) {
    if (pl) {
      void* dstL = dr ? ((X == w.B0) ? w.B2 : w.B0) : out;
      g = gemm_desc(dt, PL, m, 0, X, n, 0, m, n, m, dstL, n);
      // (P_bf16 + diag(resid)) X: the diagonal of P is large and nearly constant, its bf16 rounding would be a systematic bias
      g.epi.D = X; g.epi.ldd = n; g.epi.d_dtype = dt; g.epi.beta = 1.f; g.epi.d_row_scale = w.presid[0];
      if (!dr) set_final(g);
      rc = chain_emit(ctx, plan, 1, g, st); if (rc) return rc;
      Y = dstL;
    } else {      // chain: Q_L X then Q_L^T (.): 4 m^2 n
      void* dstL = dr ? w.B0 : out;
      g = gemm_desc(dt, k->QL, m, 0, X, n, 0, m, n, m, w.B2, n);
      rc = chain_emit(ctx, plan, 1, g, st); if (rc) return rc;
      g = gemm_desc(dt, k->QL, m, 1, w.B2, n, 0, m, n, m, dstL, n);
      if (!dr) set_final(g);
      rc = chain_emit(ctx, plan, 2, g, st); if (rc) return rc;
      Y = dstL;
    }
  }
  // ---- right side ----
  if (dr) {
    if (pr) {
      g = gemm_desc(dt, Y, n, 0, PR, n, 0, m, n, n, out, n);
      g.epi.D = Y; g.epi.ldd = n; g.epi.d_dtype = dt; g.epi.beta = 1.f; g.epi.d_col_scale = w.presid[1];
      set_final(g);
      rc = chain_emit(ctx, plan, 3, g, st); if (rc) return rc;
    } else {      // Y Q_R^T then (.) Q_R
      void* tr = (Y == w.B2) ? w.B0 : w.B2;
      g = gemm_desc(dt, Y, n, 0, k->QR, n, 1, m, n, n, tr, n);
      rc = chain_emit(ctx, plan, 3, g, st); if (rc) return rc;
      g = gemm_desc(dt, tr, n, 0, k->QR, n, 0, m, n, n, out, n);
      set_final(g);
      rc = chain_emit(ctx, plan, 4, g, st); if (rc) return rc;
    }
  }
  return PSGD_OK;
}

// (kron_i Q_i^T Q_i) X for every unit of a same-shape batch: X[u] -> out[u]; optional per-unit reductions on the outputs
struct ChainIO { const void* X; void* out; float* row_sumsq; float* col_sumsq; float* total_sumsq; };

static int run_chain_batch(Ctx* ctx, const psgd_kron_t* ks, int n, KronWs* w, const ChainIO* io, cudaStream_t st) {
  const psgd_kron_t& k0 = ks[0];
  const int dt = k0.dtype;
  const int m = k0.m, nn = k0.has_r ? k0.n : 1;
  const bool dl = k0.kind_l == PSGD_DENSE, dr = k0.has_r && k0.kind_r == PSGD_DENSE;
  if (n == 1) return run_chain(ctx, ks, w[0], io[0].X, io[0].out, io[0].row_sumsq, io[0].col_sumsq, io[0].total_sumsq, st);
  // squares of the diagonal factors, one launch per side
  if (!dl) {
    CPtrTab q; PtrTab o;
    for (int u = 0; u < n; ++u) { q.p[u] = ks[u].QL; o.p[u] = w[u].qsq[0]; }
    DISPATCH_T(dt, (k_square_to_f32_multi<T><<<dim3((m + 255) / 256, n), 256, 0, st>>>(q, o, m)));
    LAUNCH_CHECK(ctx, "k_square_multi");
  }
  if (k0.has_r && !dr) {
    CPtrTab q; PtrTab o;
    for (int u = 0; u < n; ++u) { q.p[u] = ks[u].QR; o.p[u] = w[u].qsq[1]; }
    DISPATCH_T(dt, (k_square_to_f32_multi<T><<<dim3((nn + 255) / 256, n), 256, 0, st>>>(q, o, nn)));
    LAUNCH_CHECK(ctx, "k_square_multi");
  }
  if (!dl && !dr) {   // 1-D tensors / diag x diag: one scaling kernel for the whole batch
    CPtrTab X, rs, cs; PtrTab out, rss, css, tss;
    for (int u = 0; u < n; ++u) {
      X.p[u] = io[u].X; out.p[u] = io[u].out; rs.p[u] = w[u].qsq[0]; cs.p[u] = k0.has_r ? w[u].qsq[1] : nullptr;
      rss.p[u] = io[u].row_sumsq; css.p[u] = io[u].col_sumsq; tss.p[u] = io[u].total_sumsq;
    }
    const size_t numel = (size_t)m * nn;
    int bx = (int)((numel + 255) / 256); if (bx > ctx->num_sms * 8) bx = ctx->num_sms * 8; if (bx < 1) bx = 1;
    DISPATCH_T(dt, (k_scale2d_multi<T><<<dim3(bx, n), 256, 0, st>>>(X, out, m, nn, rs, cs, rss, css, tss)));
    LAUNCH_CHECK(ctx, "k_scale2d_multi");
    return PSGD_OK;
  }
  ChainPlan plans[KB_MAX];
  for (int u = 0; u < n; ++u) {
    memset(plans[u].cnt, 0, sizeof(plans[u].cnt));
    plans[u].squares_done = true;
    int rc = run_chain(ctx, ks + u, w[u], io[u].X, io[u].out, io[u].row_sumsq, io[u].col_sumsq, io[u].total_sumsq, st, &plans[u]);
    if (rc) return rc;
  }
  GemmDesc lvl[2 * KB_MAX];
  for (int level = 0; level < 5; ++level) {
    int c = 0;
    for (int u = 0; u < n; ++u)
      for (int i = 0; i < plans[u].cnt[level]; ++i) lvl[c++] = plans[u].g[level][i];
    int rc = launch_gemm_group(ctx, lvl, c, st); if (rc) return rc;
  }
  return PSGD_OK;
}

static int validate_kron(const psgd_kron_t* k) {
  if (!k || k->m < 1 || !k->QL || !k->LL) return PSGD_ERR_INVALID_ARG;
  if (k->dtype != PSGD_BF16 && k->dtype != PSGD_F32) return PSGD_ERR_INVALID_ARG;
  if (k->has_r && (k->n < 1 || !k->QR || !k->LR)) return PSGD_ERR_INVALID_ARG;
  return PSGD_OK;
}

static int run_balance(Ctx* ctx, const psgd_kron_t* k, KronWs& w, cudaStream_t st) {
  if (!k->has_r) return PSGD_OK;  // order-1: nothing to balance (psgd.py:271)
  const int dt = k->dtype;
  size_t nl = k->kind_l == PSGD_DENSE ? (size_t)k->m * k->m : (size_t)k->m;
  size_t nr = k->kind_r == PSGD_DENSE ? (size_t)k->n * k->n : (size_t)k->n;
  int rc = check_cuda(ctx, cudaMemsetAsync(w.bal, 0, 8, st), "memset"); if (rc) return rc;
  DISPATCH_T(dt, (k_absmax<T><<<ew_blocks(ctx, nl), 256, 0, st>>>((const T*)k->QL, nl, w.bal)));
  LAUNCH_CHECK(ctx, "k_absmax");
  DISPATCH_T(dt, (k_absmax<T><<<ew_blocks(ctx, nr), 256, 0, st>>>((const T*)k->QR, nr, w.bal + 1)));
  LAUNCH_CHECK(ctx, "k_absmax");
  DISPATCH_T(dt, (k_balance_scale<T><<<ew_blocks(ctx, nl), 256, 0, st>>>((T*)k->QL, nl, w.bal, w.bal + 1, dt)));
  LAUNCH_CHECK(ctx, "k_balance_scale");
  DISPATCH_T(dt, (k_balance_scale<T><<<ew_blocks(ctx, nr), 256, 0, st>>>((T*)k->QR, nr, w.bal + 1, w.bal, dt)));
  LAUNCH_CHECK(ctx, "k_balance_scale");
  return PSGD_OK;
}

#include "kron_geom.cuh"

}  // namespace psgd

using namespace psgd;

// =================================================================================================
// extern "C"
// =================================================================================================
extern "C" {

int psgd_abi_version(void) { return PSGD_B200_ABI_VERSION; }

const char* psgd_status_string(int s) {
  switch (s) {
    case PSGD_OK: return "ok";
    case PSGD_ERR_INVALID_ARG: return "invalid argument";
    case PSGD_ERR_UNSUPPORTED: return "unsupported shape/dtype for the requested path";
    case PSGD_ERR_WORKSPACE: return "workspace too small";
    case PSGD_ERR_CUDA: return "CUDA error (see psgd_last_error)";
    case PSGD_ERR_NOT_SM100: return "device is not compute capability 10.x (B200); this engine has no other code path";
    default: return "unknown status";
  }
}

const char* psgd_last_error(psgd_handle_t h) { return h ? reinterpret_cast<Ctx*>(h)->last_error : ""; }

int psgd_create(psgd_handle_t* out, int device) {
  if (!out) return PSGD_ERR_INVALID_ARG;
  *out = nullptr;
  cudaDeviceProp prop;
  cudaError_t e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) return PSGD_ERR_CUDA;
  if (prop.major != 10) return PSGD_ERR_NOT_SM100;
  Ctx* ctx = new (std::nothrow) Ctx();
  if (!ctx) return PSGD_ERR_CUDA;
  memset(ctx, 0, sizeof(Ctx));
  ctx->device = device;
  ctx->num_sms = prop.multiProcessorCount;
  ctx->num_sms_hw = prop.multiProcessorCount;
  ctx->gemm_path = 0;
  ctx->mn_lbo = 8192;
  ctx->mn_sbo = 1024;
  if (const char* dbg = getenv("PSGD_B200_DEBUG_FLAGS")) ctx->debug_flags = (int)strtol(dbg, nullptr, 0);   // experiments only (see psgd_debug_set_flags)
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess) fn = nullptr;
  ctx->encode_tiled = fn;
  {  // split-K workspace: partial-tile slots + arrival counters; all launches of a handle are expected on one stream at a time
    int prev = 0;
    cudaGetDevice(&prev);
    cudaSetDevice(device);
    ctx->ws_slots = 320;   // partial tiles of 128 x 256 fp32 (42 MB)
    const size_t bytes = (size_t)ctx->ws_slots * 128 * 256 * sizeof(float);
    if (cudaMalloc(&ctx->ws, bytes) == cudaSuccess && cudaMalloc(&ctx->ws_count, ctx->ws_slots * sizeof(int)) == cudaSuccess) {
      cudaMemset(ctx->ws, 0, bytes);
      cudaMemset(ctx->ws_count, 0, ctx->ws_slots * sizeof(int));
    } else {
      ctx->ws = nullptr; ctx->ws_count = nullptr; ctx->ws_slots = 0;   // split-K simply stays off
      cudaGetLastError();
    }
    cudaSetDevice(prev);
  }
  {  // grid-barrier / completion counters of the fused norm-bound kernel (self-resetting)
    int prev = 0;
    cudaGetDevice(&prev);
    cudaSetDevice(device);
    int coop = 0;
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device);
    if (coop && cudaMalloc(&ctx->nb_sync, 256) == cudaSuccess) cudaMemset(ctx->nb_sync, 0, 256);
    else { ctx->nb_sync = nullptr; cudaGetLastError(); }
    cudaSetDevice(prev);
  }
  *out = reinterpret_cast<psgd_handle_t>(ctx);
  return PSGD_OK;
}

void psgd_destroy(psgd_handle_t h) {
  Ctx* ctx = reinterpret_cast<Ctx*>(h);
  if (!ctx) return;
  if (ctx->nb_sync) cudaFree(ctx->nb_sync);
  if (ctx->ws) cudaFree(ctx->ws);
  if (ctx->ws_count) cudaFree(ctx->ws_count);
  for (int i = 0; i < 2; ++i) if (ctx->x3_buf[i]) cudaFree(ctx->x3_buf[i]);
  delete ctx;
}

int psgd_set_gemm_path(psgd_handle_t h, int p) {
  if (!h || p < 0 || p > 2) return PSGD_ERR_INVALID_ARG;
  reinterpret_cast<Ctx*>(h)->gemm_path = p;
  return PSGD_OK;
}

int psgd_set_sm_limit(psgd_handle_t h, int sms) {
  Ctx* ctx = reinterpret_cast<Ctx*>(h);
  if (!ctx) return PSGD_ERR_INVALID_ARG;
  ctx->num_sms = (sms <= 0 || sms > ctx->num_sms_hw) ? ctx->num_sms_hw : (sms < 2 ? 2 : sms & ~1);   // even: the 2-CTA GEMM runs CTA pairs
  return PSGD_OK;
}

int psgd_peer_enable(int peer_device) {
  int cur = -1;
  if (cudaGetDevice(&cur) != cudaSuccess) return PSGD_ERR_CUDA;
  if (peer_device == cur) return PSGD_OK;
  int can = 0;
  if (cudaDeviceCanAccessPeer(&can, cur, peer_device) != cudaSuccess || !can) { cudaGetLastError(); return PSGD_ERR_UNSUPPORTED; }
  const cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
  if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); return PSGD_ERR_CUDA; }
  cudaGetLastError();
  return PSGD_OK;
}

int psgd_peer_copy_async(void* dst, const void* src, size_t bytes, void* stream) {
  if (!dst || !src) return PSGD_ERR_INVALID_ARG;
  if (bytes == 0) return PSGD_OK;
  return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, reinterpret_cast<cudaStream_t>(stream)) == cudaSuccess ? PSGD_OK : PSGD_ERR_CUDA;
}

int64_t psgd_launch_count(psgd_handle_t h) { return h ? reinterpret_cast<Ctx*>(h)->launches : 0; }

int psgd_set_fp32_tensor_cores(psgd_handle_t h, int on) {
  if (!h) return PSGD_ERR_INVALID_ARG;
  reinterpret_cast<Ctx*>(h)->fp32_tensor = on ? 1 : 0;
  return PSGD_OK;
}

int psgd_timing_enable(psgd_handle_t h, int on) {
  Ctx* ctx = reinterpret_cast<Ctx*>(h);
  if (!ctx) return PSGD_ERR_INVALID_ARG;
  const int want = on > 1 ? on : 8192;      // on > 1: size of the event pool (launches that can be timed between two resets)
  if (on && ctx->ev_begin && ctx->ev_capacity < want) {
    for (int i = 0; i < ctx->ev_capacity; ++i) { cudaEventDestroy(ctx->ev_begin[i]); cudaEventDestroy(ctx->ev_end[i]); }
    delete[] ctx->ev_begin; delete[] ctx->ev_end;
    ctx->ev_begin = nullptr; ctx->ev_end = nullptr;
  }
  if (on && !ctx->ev_begin) {
    ctx->ev_capacity = want;
    ctx->ev_begin = new (std::nothrow) cudaEvent_t[ctx->ev_capacity];
    ctx->ev_end = new (std::nothrow) cudaEvent_t[ctx->ev_capacity];
    if (!ctx->ev_begin || !ctx->ev_end) return PSGD_ERR_CUDA;
    for (int i = 0; i < ctx->ev_capacity; ++i) { cudaEventCreate(&ctx->ev_begin[i]); cudaEventCreate(&ctx->ev_end[i]); }
  }
  ctx->timing_on = on ? 1 : 0;
  ctx->timing_count = 0;
  ctx->timing_seen = 0;
  ctx->timing_flops = 0.0;
  ctx->timing_flops_exec = 0.0;
  return PSGD_OK;
}

int psgd_timing_read(psgd_handle_t h, int* launches, double* total_ms, double* flops) {
  Ctx* ctx = reinterpret_cast<Ctx*>(h);
  if (!ctx || !launches || !total_ms || !flops) return PSGD_ERR_INVALID_ARG;
  double ms = 0.0;
  for (int i = 0; i < ctx->timing_count; ++i) {
    float t = 0.f;
    if (cudaEventElapsedTime(&t, ctx->ev_begin[i], ctx->ev_end[i]) != cudaSuccess) return PSGD_ERR_CUDA;
    ms += t;
  }
  *launches = ctx->timing_count; *total_ms = ms; *flops = ctx->timing_flops;
  return PSGD_OK;
}

int64_t psgd_timing_gemm_launches(psgd_handle_t h) { return h ? reinterpret_cast<Ctx*>(h)->timing_seen : 0; }

double psgd_timing_executed_flops(psgd_handle_t h) { return h ? reinterpret_cast<Ctx*>(h)->timing_flops_exec : 0.0; }

int psgd_debug_set_flags(psgd_handle_t h, int flags) {
  if (!h) return PSGD_ERR_INVALID_ARG;
  reinterpret_cast<Ctx*>(h)->debug_flags = flags;
  return PSGD_OK;
}

int psgd_debug_set_tile_n(psgd_handle_t h, int bn) {
  if (!h || (bn != 0 && bn != 128 && bn != 256)) return PSGD_ERR_INVALID_ARG;
  reinterpret_cast<Ctx*>(h)->force_bn = bn;
  return PSGD_OK;
}

int psgd_debug_set_mn_desc(psgd_handle_t h, int lbo, int sbo) {
  if (!h) return PSGD_ERR_INVALID_ARG;
  reinterpret_cast<Ctx*>(h)->mn_lbo = lbo;
  reinterpret_cast<Ctx*>(h)->mn_sbo = sbo;
  return PSGD_OK;
}

size_t psgd_kron_workspace_bytes(psgd_handle_t, const psgd_kron_t* k) {
  if (validate_kron(k)) return 0;
  KronWs w;
  layout_kron(k, nullptr, w);
  return w.total;
}

}  // extern "C"

namespace psgd {
static int validate_batch(const psgd_kron_t* ks, int n) {
  if (!ks || n < 1 || n > KB_MAX) return PSGD_ERR_INVALID_ARG;
  for (int u = 0; u < n; ++u) {
    int rc = validate_kron(ks + u); if (rc) return rc;
    if (ks[u].m != ks[0].m || ks[u].n != ks[0].n || ks[u].kind_l != ks[0].kind_l || ks[u].kind_r != ks[0].kind_r ||
        ks[u].dtype != ks[0].dtype || ks[u].has_r != ks[0].has_r)
      return PSGD_ERR_INVALID_ARG;   // one batch = one shape bucket
  }
  return PSGD_OK;
}

// psgd.py:394-419 for n units of one shape bucket, every stage issued once for the whole batch
static int kron_update_batch(Ctx* ctx, const psgd_kron_t* ks, int n, const void* const* Gs, float lr, float betaL, float damping,
                             const psgd_kron_noise_t* noises, const int* do_balance, KronWs* w, char* zero_begin, size_t zero_bytes,
                             cudaStream_t st) {
  const psgd_kron_t& k0 = ks[0];
  const int dt = k0.dtype;
  const int m = k0.m, nn = k0.has_r ? k0.n : 1;
  const size_t numel = (size_t)m * nn;
  const bool dense[2] = {k0.kind_l == PSGD_DENSE, k0.has_r && k0.kind_r == PSGD_DENSE};
  // parity mode: the host supplies every random number; performance mode (NULL pointers): drawn on the device (Philox).  One mode per batch.
  const bool dev_noise = noises[0].N == nullptr;
  const void* probes[KB_MAX][2][2];   // [unit][factor][spd, skh]
  bool gen_probes = false;
  for (int u = 0; u < n; ++u) {
    if (!Gs[u] || (noises[u].N == nullptr) != dev_noise) return PSGD_ERR_INVALID_ARG;
    const void* given[2][2] = {{noises[u].V0_spd_l, noises[u].V0_skh_l}, {noises[u].V0_spd_r, noises[u].V0_skh_r}};
    for (int i = 0; i < 2; ++i)
      for (int b = 0; b < 2; ++b) {
        probes[u][i][b] = given[i][b];
        if (dense[i] && !given[i][b]) { probes[u][i][b] = w[u].P0[i][b]; gen_probes = true; }
      }
  }
  int rc = check_cuda(ctx, cudaMemsetAsync(zero_begin, 0, zero_bytes, st), "memset"); if (rc) return rc;

  // G' = G + (damping + eps|G|) N      psgd.py:402-403
  {
    CPtrTab G, N; PtrTab O;
    U64Tab seeds, offs;
    bool vec = dt == PSGD_BF16 && numel % 8 == 0;
    for (int u = 0; u < n; ++u) {
      G.p[u] = Gs[u]; N.p[u] = noises[u].N; O.p[u] = w[u].B0;
      seeds.v[u] = noises[u].philox_seed; offs.v[u] = noises[u].philox_offset;
      vec = vec && (reinterpret_cast<uintptr_t>(Gs[u]) & 15u) == 0 && (reinterpret_cast<uintptr_t>(noises[u].N) & 15u) == 0;
    }
    int bx = ew_blocks(ctx, vec ? numel / 8 : numel);
    if (n > 1) { bx = (bx + n - 1) / n; if (bx < 1) bx = 1; }
    if (dev_noise) {
      if (vec) k_add_noise_philox_multi<bf16, true><<<dim3(bx, n), 256, 0, st>>>(G, O, numel, damping, dtype_eps(dt), seeds, offs);
      else DISPATCH_T(dt, (k_add_noise_philox_multi<T, false><<<dim3(bx, n), 256, 0, st>>>(G, O, numel, damping, dtype_eps(dt), seeds, offs)));
    } else {
      if (vec) k_add_noise_multi<bf16, true><<<dim3(bx, n), 256, 0, st>>>(G, N, O, numel, damping, dtype_eps(dt));
      else DISPATCH_T(dt, (k_add_noise_multi<T, false><<<dim3(bx, n), 256, 0, st>>>(G, N, O, numel, damping, dtype_eps(dt))));
    }
    LAUNCH_CHECK(ctx, "k_add_noise");
  }
  if (gen_probes) {   // the probe blocks the host did not supply: one launch for the whole batch
    ProbeTab tab;
    int ne = 0;
    for (int u = 0; u < n; ++u)
      for (int i = 0; i < 2; ++i)
        for (int b = 0; b < 2; ++b)
          if (dense[i] && probes[u][i][b] == w[u].P0[i][b]) {
            tab.p[ne] = w[u].P0[i][b]; tab.seed[ne] = noises[u].philox_seed; tab.off[ne] = noises[u].philox_offset;
            tab.strm[ne] = 1u + 2u * i + b;
            ++ne;
          }
    // all entries of one call have the same size only if m == n or one factor is dense; generate per factor size otherwise
    for (int i = 0; i < 2; ++i) {
      if (!dense[i]) continue;
      ProbeTab t2; int c = 0;
      for (int e = 0; e < ne; ++e)
        if ((int)((tab.strm[e] - 1u) / 2u) == i) { t2.p[c] = tab.p[e]; t2.seed[c] = tab.seed[e]; t2.off[c] = tab.off[e]; t2.strm[c] = tab.strm[e]; ++c; }
      if (!c) continue;
      const size_t pn = (size_t)32 * (i == 0 ? m : nn);
      int bx = (int)((pn / 4 + 255) / 256); if (bx < 1) bx = 1; if (bx > 64) bx = 64;
      DISPATCH_T(dt, (k_philox_probes<T><<<dim3(bx, c), 256, 0, st>>>(t2, pn)));
      LAUNCH_CHECK(ctx, "k_philox_probes");
    }
  }
  // Pg = P G'  with the sums of squares the diagonal factors need fused into the last product
  ChainIO io[KB_MAX];
  for (int u = 0; u < n; ++u)
    io[u] = ChainIO{w[u].B0, w[u].B1, dense[0] ? nullptr : w[u].f[0].term1, (k0.has_r && !dense[1]) ? w[u].f[1].term1 : nullptr, nullptr};
  rc = run_chain_batch(ctx, ks, n, w, io, st); if (rc) return rc;

  // Grams of the dense factors (psgd.py:405) -- grouped launches over factors and units
  {
    GemmDesc gg[2 * KB_MAX];
    int ng = 0;
    for (int u = 0; u < n; ++u) {
      void* Pg = w[u].B1;
      if (dense[0]) {
        gg[ng] = gemm_desc(dt, Pg, nn, 0, Pg, nn, 1, m, m, nn, w[u].S[0][0], m); gg[ng].sym = 1;
        gg[ng].epi.row_sumsq = w[u].f[0].row_sumsq; gg[ng].epi.diag_max = w[u].f[0].diag_max; ++ng;
      }
      if (dense[1]) {
        gg[ng] = gemm_desc(dt, Pg, nn, 1, Pg, nn, 0, nn, nn, m, w[u].S[1][0], nn); gg[ng].sym = 1;
        gg[ng].epi.row_sumsq = w[u].f[1].row_sumsq; gg[ng].epi.diag_max = w[u].f[1].diag_max; ++ng;
      }
    }
    rc = launch_gemm_group(ctx, gg, ng, st); if (rc) return rc;
  }

  DenseItem items[KB_ITEMS];
  int nd = 0;
  for (int i = 0; i < 2; ++i) {
    if (i == 1 && !k0.has_r) break;
    const int s = i == 0 ? m : nn;
    const float t2 = (float)((double)numel / (double)s);  // psgd.py:407 / 412
    if (!dense[i]) {   // diagonal factors of the whole batch: one launch (psgd.py:406-410)
      PtrTab q, L; CPtrTab t1;
      for (int u = 0; u < n; ++u) { q.p[u] = i == 0 ? ks[u].QL : ks[u].QR; L.p[u] = i == 0 ? ks[u].LL : ks[u].LR; t1.p[u] = w[u].f[i].term1; }
      DISPATCH_T(dt, (k_diag_update_multi<T><<<n, 1024, 0, st>>>(q, t1, s, t2, lr, betaL, L)));
      LAUNCH_CHECK(ctx, "k_diag_update");
      continue;
    }
    for (int u = 0; u < n; ++u) {
      DenseItem& d = items[nd++];
      FactorWs& f = w[u].f[i];
      d.s = s; d.q = i == 0 ? ks[u].QL : ks[u].QR; d.L = i == 0 ? ks[u].LL : ks[u].LR; d.t2 = t2;
      d.T = w[u].S[i][0]; d.Qn = w[u].S[i][1]; d.RQ = w[u].S[i][2]; d.RRQ = w[u].S[i][3]; d.Va = w[u].Va[i]; d.Vb = w[u].Vb[i];
      d.v_spd = probes[u][i][0];
      d.v_skh = probes[u][i][1];
      d.f = &f;
    }
  }
  rc = run_dense_factors(ctx, dt, items, nd, lr, betaL, st); if (rc) return rc;
  for (int u = 0; u < n; ++u)
    if (do_balance && do_balance[u]) { rc = run_balance(ctx, ks + u, w[u], st); if (rc) return rc; }
  return PSGD_OK;
}
}  // namespace psgd

extern "C" {

size_t psgd_kron_batch_workspace_bytes(psgd_handle_t, const psgd_kron_t* ks, int n) {
  if (validate_batch(ks, n)) return 0;
  KronWs w[KB_MAX];
  return layout_kron_batch(ks, n, nullptr, w, nullptr, nullptr);
}

int psgd_kron_whiten_q0p5eq1p5_update_batched(psgd_handle_t h, const psgd_kron_t* ks, int n, const void* const* Gs, float lr, float betaL,
                                              float damping, const psgd_kron_noise_t* noises, const int* do_balance, void* workspace,
                                              size_t workspace_bytes, void* stream) {
  Ctx* ctx = reinterpret_cast<Ctx*>(h);
  if (!ctx || !Gs || !noises) return PSGD_ERR_INVALID_ARG;
  int rc = validate_batch(ks, n); if (rc) return rc;
  KronWs w[KB_MAX];
  char* zb = nullptr; size_t zn = 0;
  const size_t total = layout_kron_batch(ks, n, workspace, w, &zb, &zn);
  if (!workspace || workspace_bytes < total) return PSGD_ERR_WORKSPACE;
  return kron_update_batch(ctx, ks, n, Gs, lr, betaL, damping, noises, do_balance, w, zb, zn, reinterpret_cast<cudaStream_t>(stream));
}

int psgd_kron_whiten_q0p5eq1p5_update(psgd_handle_t h, const psgd_kron_t* k, const void* G, float lr, float betaL, float damping,
                                      const psgd_kron_noise_t* noise, int do_balance, void* workspace, size_t workspace_bytes,
                                      void* stream) {
  return psgd_kron_whiten_q0p5eq1p5_update_batched(h, k, 1, &G, lr, betaL, damping, noise, &do_balance, workspace, workspace_bytes, stream);
}

int psgd_kron_precond_grad_batched(psgd_handle_t h, const psgd_kron_t* ks, int n, const void* const* Xs, void* const* Hs, float* sumsq_out,
                                   void* workspace, size_t workspace_bytes, void* stream) {
  Ctx* ctx = reinterpret_cast<Ctx*>(h);
  if (!ctx || !Xs || !Hs) return PSGD_ERR_INVALID_ARG;
  int rc = validate_batch(ks, n); if (rc) return rc;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  KronWs w[KB_MAX];
  const size_t total = layout_kron_batch(ks, n, workspace, w, nullptr, nullptr);
  if (!workspace || workspace_bytes < total) return PSGD_ERR_WORKSPACE;
  if (sumsq_out) { rc = check_cuda(ctx, cudaMemsetAsync(sumsq_out, 0, 4 * (size_t)n, st), "memset"); if (rc) return rc; }
  ChainIO io[KB_MAX];
  for (int u = 0; u < n; ++u) {
    if (!Xs[u] || !Hs[u]) return PSGD_ERR_INVALID_ARG;
    io[u] = ChainIO{Xs[u], Hs[u], nullptr, nullptr, sumsq_out ? sumsq_out + u : nullptr};
  }
  return run_chain_batch(ctx, ks, n, w, io, st);
}

int psgd_kron_precond_grad(psgd_handle_t h, const psgd_kron_t* k, const void* X, void* H, float* sumsq_out, void* workspace,
                           size_t workspace_bytes, void* stream) {
  return psgd_kron_precond_grad_batched(h, k, 1, &X, &H, sumsq_out, workspace, workspace_bytes, stream);
}

int psgd_kron_balance(psgd_handle_t h, const psgd_kron_t* k, void* workspace, size_t workspace_bytes, void* stream) {
  Ctx* ctx = reinterpret_cast<Ctx*>(h);
  if (!ctx) return PSGD_ERR_INVALID_ARG;
  int rc = validate_kron(k); if (rc) return rc;
  KronWs w;
  layout_kron(k, workspace, w);
  if (!workspace || workspace_bytes < w.total) return PSGD_ERR_WORKSPACE;
  return run_balance(ctx, k, w, reinterpret_cast<cudaStream_t>(stream));
}

// ------------------------------- helpers exposed like the reference exposes them -------------------------------
struct HelperWs {
  float* row_sumsq; float* nf; FactorWs f; void* R; void* RQ; void* RRQ; void* RRRQ; void* Qn; void* Va; void* Vb;
  char* zero_begin; size_t zero_bytes; size_t total;
};
static void layout_helper(int s, int dt, void* base, HelperWs& w) {
  Bump b(base);
  const int es = dtype_size(dt);
  w.zero_begin = (char*)b.take(0);
  size_t z0 = b.off;
  w.row_sumsq = (float*)b.take((size_t)s * 4);
  w.nf = (float*)b.take(4);
  w.f.row_sumsq = w.row_sumsq; w.f.diag_max = w.nf;
  w.f.r_row_sumsq = (float*)b.take((size_t)s * 4);
  w.f.r_abs_max = (float*)b.take(4);
  w.f.fs = (float*)b.take(FS_COUNT * 4);
  layout_bound(b, w.f.b_spd);
  layout_bound(b, w.f.b_skh);
  w.f.term1 = nullptr;
  w.zero_bytes = b.off - z0;
  w.R = b.take((size_t)s * s * es); w.RQ = b.take((size_t)s * s * es); w.RRQ = b.take((size_t)s * s * es);
  w.Qn = b.take((size_t)s * s * es);
  w.Va = b.take((size_t)32 * s * es); w.Vb = b.take((size_t)32 * s * es);
  w.RRRQ = b.take((size_t)s * s * es);
  w.total = b.off;
}

size_t psgd_helper_workspace_bytes(psgd_handle_t, int s, int dtype) {
  if (s < 1) return 0;
  HelperWs w;
  layout_helper(s, dtype, nullptr, w);
  return w.total;
}

}  // extern "C"
// row sums of squares + max diag / max abs of an s x s matrix (standalone bound entry points only)
template <typename T>
__global__ void k_rowstats(const T* __restrict__ A, int s, float* row_sumsq, float* diag_max, float* abs_max) {
  __shared__ float red[32];
  int i = blockIdx.x;
  float ss = 0.f, am = 0.f;
  for (int c = threadIdx.x; c < s; c += blockDim.x) { float v = to_f<T>(A[(size_t)i * s + c]); ss += v * v; am = fmaxf(am, fabsf(v)); }
  ss = block_sum(ss, red);
  if (threadIdx.x == 0) row_sumsq[i] = ss;
  am = block_max(am, red);
  if (threadIdx.x == 0) {
    if (abs_max) atomic_max_nonneg(abs_max, am);
    if (diag_max) atomic_max_nonneg(diag_max, to_f<T>(A[(size_t)i * s + i]));
  }
}
__global__ void k_copy_scalar(const float* src, float* dst) { if (threadIdx.x == 0) *dst = *src; }

static int bound_entry(psgd_handle_t h, int dt, const void* A, int s, const void* V0, float* out, void* workspace, size_t wsb,
                       void* stream, bool spd) {
  Ctx* ctx = reinterpret_cast<Ctx*>(h);
  if (!ctx || !A || !V0 || !out || s < 1) return PSGD_ERR_INVALID_ARG;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  HelperWs w;
  layout_helper(s, dt, workspace, w);
  if (!workspace || wsb < w.total) return PSGD_ERR_WORKSPACE;
  int rc = check_cuda(ctx, cudaMemsetAsync(w.zero_begin, 0, w.zero_bytes, st), "memset"); if (rc) return rc;
  DISPATCH_T(dt, (k_rowstats<T><<<s, 256, 0, st>>>((const T*)A, s, w.row_sumsq, spd ? w.nf : nullptr, spd ? nullptr : w.nf)));
  LAUNCH_CHECK(ctx, "k_rowstats");
  BoundJob jb{A, s, V0, w.row_sumsq, w.nf, &w.f.b_spd, w.Va, w.Vb};
  rc = run_bounds(ctx, dt, &jb, 1, nullptr, st); if (rc) return rc;
  k_copy_scalar<<<1, 32, 0, st>>>(w.f.b_spd.scal + SC_BOUND, out);
  LAUNCH_CHECK(ctx, "k_copy_scalar");
  return PSGD_OK;
}

extern "C" {
int psgd_norm_lower_bound_spd(psgd_handle_t h, int dt, const void* A, int s, const void* V0, float* out, void* ws, size_t wsb, void* stream) {
  return bound_entry(h, dt, A, s, V0, out, ws, wsb, stream, true);
}
int psgd_norm_lower_bound_skh(psgd_handle_t h, int dt, const void* A, int s, const void* V0, float* out, void* ws, size_t wsb, void* stream) {
  return bound_entry(h, dt, A, s, V0, out, ws, wsb, stream, false);
}

// one factor's share of psgd.py:404-416 given its Gram / sums of squares (building block of the order >= 3 host composition)
int psgd_kron_factor_update(psgd_handle_t h, int dt, int kind, int s, void* q, float* L, const void* term1, float t2, float lr, float betaL,
                            const void* V0_spd, const void* V0_skh, void* workspace, size_t wsb, void* stream) {
  Ctx* ctx = reinterpret_cast<Ctx*>(h);
  if (!ctx || !q || !L || !term1 || s < 1) return PSGD_ERR_INVALID_ARG;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (kind == PSGD_DIAG) {
    DISPATCH_T(dt, (k_diag_update<T><<<1, 1024, 0, st>>>((T*)q, (const float*)term1, s, t2, lr, betaL, L)));
    LAUNCH_CHECK(ctx, "k_diag_update");
    return PSGD_OK;
  }
  if (!V0_spd || !V0_skh) return PSGD_ERR_INVALID_ARG;
  HelperWs w;
  layout_helper(s, dt, workspace, w);
  if (!workspace || wsb < w.total) return PSGD_ERR_WORKSPACE;
  int rc = check_cuda(ctx, cudaMemsetAsync(w.zero_begin, 0, w.zero_bytes, st), "memset"); if (rc) return rc;
  rc = check_cuda(ctx, cudaMemcpyAsync(w.R, term1, (size_t)s * s * dtype_size(dt), cudaMemcpyDeviceToDevice, st), "memcpy"); if (rc) return rc;
  DISPATCH_T(dt, (k_rowstats<T><<<s, 256, 0, st>>>((const T*)w.R, s, w.row_sumsq, w.nf, nullptr)));
  LAUNCH_CHECK(ctx, "k_rowstats");
  DenseItem it;
  it.s = s; it.q = q; it.L = L; it.t2 = t2; it.T = w.R; it.Qn = w.Qn; it.RQ = w.RQ; it.RRQ = w.RRQ; it.Va = w.Va; it.Vb = w.Vb;
  it.v_spd = V0_spd; it.v_skh = V0_skh; it.f = &w.f;
  return run_dense_factors(ctx, dt, &it, 1, lr, betaL, st);
}

int psgd_procrustes_step2(psgd_handle_t h, int dt, void* Q, int s, const void* V0, float max_step_size, void* workspace, size_t wsb,
                          void* stream) {
  Ctx* ctx = reinterpret_cast<Ctx*>(h);
  if (!ctx || !Q || !V0 || s < 1) return PSGD_ERR_INVALID_ARG;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  HelperWs w;
  layout_helper(s, dt, workspace, w);
  if (!workspace || wsb < w.total) return PSGD_ERR_WORKSPACE;
  int rc = check_cuda(ctx, cudaMemsetAsync(w.zero_begin, 0, w.zero_bytes, st), "memset"); if (rc) return rc;
  rc = check_cuda(ctx, cudaMemcpyAsync(w.Qn, Q, (size_t)s * s * dtype_size(dt), cudaMemcpyDeviceToDevice, st), "memcpy"); if (rc) return rc;
  DenseItem it;
  it.s = s; it.q = Q; it.L = nullptr; it.t2 = 0.f; it.T = w.R; it.Qn = w.Qn; it.RQ = w.RQ; it.RRQ = w.RRQ; it.Va = w.Va; it.Vb = w.Vb;
  it.v_spd = nullptr; it.v_skh = V0; it.f = &w.f;
  return run_procrustes(ctx, dt, &it, 1, max_step_size, st);
}

}  // extern "C"
// ------------------------------------------- KWNS4 glue -------------------------------------------
template <typename TP, typename TG>
static int head_dispatch_q(Ctx* ctx, int64_t numel, void* p, const void* grad, float wd, float lr, int dec, void* ema, void* g_out,
                           int pre_dtype, float beta, cudaStream_t st) {
  int blocks = ew_blocks(ctx, (size_t)numel);
  if (pre_dtype == PSGD_BF16)
    k_kwns4_head<TP, TG, bf16><<<blocks, 256, 0, st>>>((TP*)p, (const TG*)grad, (size_t)numel, wd, lr, dec, (bf16*)ema, (bf16*)g_out, beta);
  else
    k_kwns4_head<TP, TG, float><<<blocks, 256, 0, st>>>((TP*)p, (const TG*)grad, (size_t)numel, wd, lr, dec, (float*)ema, (float*)g_out, beta);
  LAUNCH_CHECK(ctx, "k_kwns4_head");
  return PSGD_OK;
}

extern "C" {
int psgd_kwns4_head(psgd_handle_t h, int64_t numel, void* p, int p_dtype, const void* grad, int grad_dtype, float weight_decay,
                    float lr_params, int decoupled, void* ema, void* g_out, int pre_dtype, float beta, void* stream) {
  Ctx* ctx = reinterpret_cast<Ctx*>(h);
  if (!ctx || numel < 0 || !p || !grad) return PSGD_ERR_INVALID_ARG;
  if (numel == 0) return PSGD_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (p_dtype == PSGD_BF16 && grad_dtype == PSGD_BF16) return head_dispatch_q<bf16, bf16>(ctx, numel, p, grad, weight_decay, lr_params, decoupled, ema, g_out, pre_dtype, beta, st);
  if (p_dtype == PSGD_BF16 && grad_dtype == PSGD_F32) return head_dispatch_q<bf16, float>(ctx, numel, p, grad, weight_decay, lr_params, decoupled, ema, g_out, pre_dtype, beta, st);
  if (p_dtype == PSGD_F32 && grad_dtype == PSGD_BF16) return head_dispatch_q<float, bf16>(ctx, numel, p, grad, weight_decay, lr_params, decoupled, ema, g_out, pre_dtype, beta, st);
  return head_dispatch_q<float, float>(ctx, numel, p, grad, weight_decay, lr_params, decoupled, ema, g_out, pre_dtype, beta, st);
}

int psgd_kwns4_tail(psgd_handle_t h, int64_t numel, int64_t amp_numel, void* p, int p_dtype, void* hbuf, int h_dtype, const float* sumsq,
                    float max_avg_amp, float max_elem_amp, float lr_params, void* stream) {
  Ctx* ctx = reinterpret_cast<Ctx*>(h);
  if (!ctx || numel < 0 || !p || !hbuf || !sumsq) return PSGD_ERR_INVALID_ARG;
  if (numel == 0) return PSGD_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  int blocks = ew_blocks(ctx, (size_t)numel);
  if (p_dtype == PSGD_BF16 && h_dtype == PSGD_BF16) k_kwns4_tail<bf16, bf16><<<blocks, 256, 0, st>>>((bf16*)p, (bf16*)hbuf, (size_t)numel, (float)amp_numel, sumsq, max_avg_amp, max_elem_amp, lr_params, h_dtype);
  else if (p_dtype == PSGD_BF16) k_kwns4_tail<bf16, float><<<blocks, 256, 0, st>>>((bf16*)p, (float*)hbuf, (size_t)numel, (float)amp_numel, sumsq, max_avg_amp, max_elem_amp, lr_params, h_dtype);
  else if (h_dtype == PSGD_BF16) k_kwns4_tail<float, bf16><<<blocks, 256, 0, st>>>((float*)p, (bf16*)hbuf, (size_t)numel, (float)amp_numel, sumsq, max_avg_amp, max_elem_amp, lr_params, h_dtype);
  else k_kwns4_tail<float, float><<<blocks, 256, 0, st>>>((float*)p, (float*)hbuf, (size_t)numel, (float)amp_numel, sumsq, max_avg_amp, max_elem_amp, lr_params, h_dtype);
  LAUNCH_CHECK(ctx, "k_kwns4_tail");
  return PSGD_OK;
}

// ------------------------------------------- GEMM building block -------------------------------------------
int psgd_gemm(psgd_handle_t h, int path, int in_dtype, int out_dtype, int trans_a, int trans_b, int M, int N, int K, const void* A,
              int lda, const void* B, int ldb, void* C, int ldc, float alpha, const void* D, int ldd, float beta, void* stream) {
  Ctx* ctx = reinterpret_cast<Ctx*>(h);
  if (!ctx || !A || !B || !C || M < 0 || N < 0 || K < 0) return PSGD_ERR_INVALID_ARG;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  GemmDesc g = gemm_desc(in_dtype, A, lda, trans_a, B, ldb, trans_b, M, N, K, C, ldc);
  g.epi.out_dtype = out_dtype;
  g.epi.alpha = alpha;
  if (D) { g.epi.D = D; g.epi.ldd = ldd; g.epi.d_dtype = in_dtype; g.epi.beta = beta; }
  int saved = ctx->gemm_path;
  ctx->gemm_path = path;
  int rc = launch_gemm(ctx, g, st);
  ctx->gemm_path = saved;
  return rc;
}

}  // extern "C"

// ------------------------------------------- the other geometries / Newton pairs -------------------------------------------
namespace psgd {
// procrustes_step3 (psgd.py:127-155) on it.Qn -> writes it.q
static int run_procrustes3(Ctx* ctx, int dt, DenseItem& it, void* RRRQ, float max_step, cudaStream_t st) {
  const int s = it.s;
  dim3 grid((s + 63) / 64, (s + 63) / 64);
  DISPATCH_T(dt, (k_skew<T><<<grid, 256, 0, st>>>((const T*)it.Qn, (T*)it.T, s, it.f->r_abs_max, it.f->r_row_sumsq)));
  LAUNCH_CHECK(ctx, "k_skew");
  BoundJob jb{it.T, s, it.v_skh, it.f->r_row_sumsq, it.f->r_abs_max, &it.f->b_skh, it.Va, it.Vb};
  BoundFinish fin{1, 0.f, 0.f, 0.f, nullptr, it.f->fs};
  int rc = run_bounds(ctx, dt, &jb, 1, &fin, st); if (rc) return rc;
  const void* src[3] = {it.Qn, it.RQ, it.RRQ};
  void* dst[3] = {it.RQ, it.RRQ, RRRQ};
  const int tr[3] = {FS_TR1, FS_TR2, FS_TR3};
  for (int j = 0; j < 3; ++j) {   // RQ = R Q / |R|, RRQ = R RQ / |R|, RRRQ = R RRQ / |R|
    GemmDesc g = gemm_desc(dt, it.T, s, 0, src[j], s, 0, s, s, s, dst[j], s);
    g.epi.alpha_ptr = it.f->fs + FS_INV_SR; g.epi.trace = it.f->fs + tr[j];
    rc = launch_gemm(ctx, g, st); if (rc) return rc;
  }
  const size_t numel = (size_t)s * s;
  DISPATCH_T(dt, (k_procrustes3_finish<T><<<ew_blocks(ctx, numel), 256, 0, st>>>((const T*)it.Qn, (const T*)it.RQ, (const T*)it.RRQ, (const T*)RRRQ,
                                                                                  (T*)it.q, numel, it.f->fs, max_step)));
  LAUNCH_CHECK(ctx, "k_procrustes3_finish");
  return PSGD_OK;
}
}  // namespace psgd

extern "C" {

size_t psgd_kron_update_workspace_bytes(psgd_handle_t, const psgd_kron_t* k, int dq) {
  if (validate_kron(k) || dq < 0 || dq > PSGD_DQ_PRO4P) return 0;
  GeomWs g;
  layout_geom(k, dq, nullptr, g);
  return g.total;
}

int psgd_kron_update(psgd_handle_t h, const psgd_kron_t* k, int dq, const void* X, const void* V, float lr, float betaL, float damping,
                     const psgd_kron_noise_t* noise, int stages, void* workspace, size_t workspace_bytes, void* stream) {
  Ctx* ctx = reinterpret_cast<Ctx*>(h);
  if (!ctx || !noise || dq < 0 || dq > PSGD_DQ_PRO4P) return PSGD_ERR_INVALID_ARG;
  int rc = validate_kron(k); if (rc) return rc;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  GeomWs g;
  layout_geom(k, dq, workspace, g);
  if (!workspace || workspace_bytes < g.total) return PSGD_ERR_WORKSPACE;
  if (stages & PSGD_STAGE_PREPARE) {
    if (!X || (!noise->N && !V)) return PSGD_ERR_INVALID_ARG;
    rc = geom_prepare(ctx, k, dq, X, V, damping, noise, g, st); if (rc) return rc;
  }
  // `newton` must be told to the factor stages too: a staged caller passes V (any non-null pointer) to every stage
  const bool newton = V != nullptr;
  if (stages & PSGD_STAGE_FACTOR_L) { rc = geom_factor(ctx, k, dq, newton, 0, lr, betaL, noise, g, st); if (rc) return rc; }
  if ((stages & PSGD_STAGE_FACTOR_R) && k->has_r) { rc = geom_factor(ctx, k, dq, newton, 1, lr, betaL, noise, g, st); if (rc) return rc; }
  if (stages & PSGD_STAGE_BALANCE) { rc = run_balance(ctx, k, g.k, st); if (rc) return rc; }
  return PSGD_OK;
}

int psgd_kron_apply_factors(psgd_handle_t h, const psgd_kron_t* k, const void* X, void* out, float* sumsq_out, void* workspace,
                            size_t workspace_bytes, void* stream) {
  Ctx* ctx = reinterpret_cast<Ctx*>(h);
  if (!ctx || !X || !out) return PSGD_ERR_INVALID_ARG;
  int rc = validate_kron(k); if (rc) return rc;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  GeomWs g;
  layout_geom(k, PSGD_DQ_QUAD4P, workspace, g);
  if (!workspace || workspace_bytes < g.total) return PSGD_ERR_WORKSPACE;
  if (sumsq_out) { rc = check_cuda(ctx, cudaMemsetAsync(sumsq_out, 0, 4, st), "memset"); if (rc) return rc; }
  return run_apply_factors(ctx, k, g, X, out, nullptr, nullptr, sumsq_out, st);
}

int psgd_kron_solve_factors(psgd_handle_t h, const psgd_kron_t* k, const void* V, void* out, void* workspace, size_t workspace_bytes,
                            void* stream) {
  Ctx* ctx = reinterpret_cast<Ctx*>(h);
  if (!ctx || !V || !out) return PSGD_ERR_INVALID_ARG;
  int rc = validate_kron(k); if (rc) return rc;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  GeomWs g;
  layout_geom(k, PSGD_DQ_EQ, workspace, g);
  if (!workspace || workspace_bytes < g.total) return PSGD_ERR_WORKSPACE;
  TriJob jobs[2];
  int nj = 0;
  if (k->kind_l == PSGD_DENSE) jobs[nj++] = TriJob{k->QL, k->m, &g.tri[0]};
  if (k->has_r && k->kind_r == PSGD_DENSE) jobs[nj++] = TriJob{k->QR, k->n, &g.tri[1]};
  rc = run_tri_inverse(ctx, k->dtype, jobs, nj, st); if (rc) return rc;
  return run_inverse_apply(ctx, k, g, V, out, nullptr, nullptr, st);
}

// One factor's bound + Lipschitz update + step for any geometry, given its contractions (the building block of the host-side composition
// for tensors of order >= 3; the fused psgd_kron_update forms the same quantities itself for order <= 2).
//   kind = PSGD_DENSE: term1 (s x s, dtype), term2 (s x s, dtype) or NULL (then term2 = t2 * I);
//   kind = PSGD_DIAG : term1 (fp32, length s), term2 (fp32, length s) or NULL (then term2 = t2).
int psgd_kron_factor_step(psgd_handle_t h, int dt, int dq, int kind, int s, void* q, float* L, const void* term1, const void* term2, float t2,
                          float lr, float betaL, const void* V0_spd, const void* V0_skh, void* workspace, size_t wsb, void* stream) {
  Ctx* ctx = reinterpret_cast<Ctx*>(h);
  if (!ctx || !q || !L || !term1 || s < 1 || dq < 0 || dq > PSGD_DQ_PRO4P) return PSGD_ERR_INVALID_ARG;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const bool quad = dq == PSGD_DQ_QUAD || dq == PSGD_DQ_QUAD4P;
  if (kind == PSGD_DIAG) {
    DISPATCH_T(dt, (k_diag_update_gen<T><<<1, 1024, 0, st>>>((T*)q, (const float*)term1, (const float*)term2, t2, 0, 0, s,
                                                            dq == PSGD_DQ_QUAD ? 0.5f * lr : lr, betaL, L, quad ? 1 : 0)));
    LAUNCH_CHECK(ctx, "k_diag_update_gen");
    return PSGD_OK;
  }
  HelperWs w;
  layout_helper(s, dt, workspace, w);
  if (!workspace || wsb < w.total) return PSGD_ERR_WORKSPACE;
  const size_t bytes = (size_t)s * s * dtype_size(dt);
  int rc = check_cuda(ctx, cudaMemsetAsync(w.zero_begin, 0, w.zero_bytes, st), "memset"); if (rc) return rc;
  rc = check_cuda(ctx, cudaMemcpyAsync(w.R, term1, bytes, cudaMemcpyDeviceToDevice, st), "memcpy"); if (rc) return rc;
  if (term2) {
    rc = check_cuda(ctx, cudaMemcpyAsync(w.RRRQ, term2, bytes, cudaMemcpyDeviceToDevice, st), "memcpy"); if (rc) return rc;
    DISPATCH_T(dt, (k_combine_terms<T><<<s, 256, 0, st>>>((T*)w.R, (T*)w.RRRQ, s, dq == PSGD_DQ_EQ ? 1 : 0, w.row_sumsq, w.nf)));
    LAUNCH_CHECK(ctx, "k_combine_terms");
  } else {
    DISPATCH_T(dt, (k_rowstats<T><<<s, 256, 0, st>>>((const T*)w.R, s, w.row_sumsq, w.nf, nullptr)));
    LAUNCH_CHECK(ctx, "k_rowstats");
  }
  DenseStep d;
  d.s = s; d.q = q; d.L = L; d.t2 = t2; d.S = w.R; d.E = term2 ? w.RRRQ : nullptr; d.Qn = w.Qn; d.RQ = w.RQ; d.RRQ = w.RRQ; d.Va = w.Va; d.Vb = w.Vb;
  d.v_spd = V0_spd; d.v_skh = V0_skh; d.f = &w.f;
  return geom_dense_step(ctx, dt, dq, d, lr, betaL, st);
}

int psgd_procrustes_step3(psgd_handle_t h, int dt, void* Q, int s, const void* V0, float max_step_size, void* workspace, size_t wsb,
                          void* stream) {
  Ctx* ctx = reinterpret_cast<Ctx*>(h);
  if (!ctx || !Q || !V0 || s < 1) return PSGD_ERR_INVALID_ARG;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  HelperWs w;
  layout_helper(s, dt, workspace, w);
  if (!workspace || wsb < w.total) return PSGD_ERR_WORKSPACE;
  int rc = check_cuda(ctx, cudaMemsetAsync(w.zero_begin, 0, w.zero_bytes, st), "memset"); if (rc) return rc;
  rc = check_cuda(ctx, cudaMemcpyAsync(w.Qn, Q, (size_t)s * s * dtype_size(dt), cudaMemcpyDeviceToDevice, st), "memcpy"); if (rc) return rc;
  DenseItem it;
  it.s = s; it.q = Q; it.L = nullptr; it.t2 = 0.f; it.T = w.R; it.Qn = w.Qn; it.RQ = w.RQ; it.RRQ = w.RRQ; it.Va = w.Va; it.Vb = w.Vb;
  it.v_spd = nullptr; it.v_skh = V0; it.f = &w.f;
  return run_procrustes3(ctx, dt, it, w.RRRQ, max_step_size, st);
}

int psgd_symmetry_gap(psgd_handle_t h, int dt, const void* Q, int s, float* out2, void* workspace, size_t wsb, void* stream) {
  Ctx* ctx = reinterpret_cast<Ctx*>(h);
  if (!ctx || !Q || !out2 || s < 1) return PSGD_ERR_INVALID_ARG;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  HelperWs w;
  layout_helper(s, dt, workspace, w);
  if (!workspace || wsb < w.total) return PSGD_ERR_WORKSPACE;
  int rc = check_cuda(ctx, cudaMemsetAsync(out2, 0, 8, st), "memset"); if (rc) return rc;
  dim3 grid((s + 63) / 64, (s + 63) / 64);
  DISPATCH_T(dt, (k_skew<T><<<grid, 256, 0, st>>>((const T*)Q, (T*)w.R, s, out2, nullptr)));
  LAUNCH_CHECK(ctx, "k_skew");
  const size_t numel = (size_t)s * s;
  DISPATCH_T(dt, (k_absmax<T><<<ew_blocks(ctx, numel), 256, 0, st>>>((const T*)Q, numel, out2 + 1)));
  LAUNCH_CHECK(ctx, "k_absmax");
  return PSGD_OK;
}

}  // extern "C"

// ------------------------------------------- measurement aid: pure-read HBM stream -------------------------------------------
// What a read-only sweep can reach on this part, independent of the LRA kernels' own structure (tools/read_bw_probe.py).
template <int UNROLL, bool NOALLOC>
__global__ void __launch_bounds__(512) k_read_probe(const uint4* __restrict__ x, size_t nvec, unsigned* out) {
  unsigned acc = 0;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  for (; i + (UNROLL - 1) * stride < nvec; i += UNROLL * stride) {
    uint4 v[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const uint4* p = x + i + u * stride;
      if (NOALLOC) asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v[u].x), "=r"(v[u].y), "=r"(v[u].z), "=r"(v[u].w) : "l"(p));
      else v[u] = *p;
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) acc ^= v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
  }
  for (; i < nvec; i += stride) { uint4 v = x[i]; acc ^= v.x ^ v.y ^ v.z ^ v.w; }
  if (acc == 0x12345678u) out[0] = acc;   // keeps the loads alive
}

extern "C" int psgd_debug_read_probe(psgd_handle_t h, const void* x, size_t bytes, int blocks, int threads, int unroll, int noalloc,
                                     void* scratch, void* stream) {
  Ctx* ctx = reinterpret_cast<Ctx*>(h);
  if (!ctx || !x || !scratch || blocks < 1 || threads < 32 || threads > 512) return PSGD_ERR_INVALID_ARG;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const size_t nvec = bytes / 16;
  const uint4* p = (const uint4*)x;
  unsigned* o = (unsigned*)scratch;
#define PROBE(U, NA) k_read_probe<U, NA><<<blocks, threads, 0, st>>>(p, nvec, o)
  if (noalloc) { if (unroll == 16) PROBE(16, true); else if (unroll == 8) PROBE(8, true); else if (unroll == 4) PROBE(4, true); else PROBE(1, true); }
  else { if (unroll == 16) PROBE(16, false); else if (unroll == 8) PROBE(8, false); else if (unroll == 4) PROBE(4, false); else PROBE(1, false); }
#undef PROBE
  LAUNCH_CHECK(ctx, "k_read_probe");
  return PSGD_OK;
}
