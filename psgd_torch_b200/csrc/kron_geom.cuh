// kron_geom.cuh -- the Kron geometries beyond Q0.5EQ1.5 and the Newton-pair updates, orchestrated out of the
// same GEMM / bound / bandwidth kernels.  Included by api.cu inside namespace psgd (it uses api.cu's static helpers).
//
//   reference                                                   here
//   update_precond_kron_eq            psgd.py:278-319           run_geom_update(dq = EQ): exprA chain, blocked triangular inverse, triu step
//   update_precond_kron_whiten_*      psgd.py:330-513           run_geom_update(V == nullptr)
//   update_precond_kron_newton_*      psgd.py:657-829           run_geom_update(V != nullptr)
//   procrustes_step3                  psgd.py:127-155           run_procrustes3
//   exprA(*Q, X)                      psgd.py:248-249           run_apply_factors
#pragma once

// ---------------------------------------------------------------------------------------------
// workspace of one generic update: the KronWs of api.cu followed by the extras below
// ---------------------------------------------------------------------------------------------
struct TriWs {           // blocked inversion of one upper-triangular dense factor
  float* Xf;             // the inverse in fp32, s x s (leaves and low levels; every level for fp32 factors)
  bf16* Xhi; bf16* Xlo;  // bf16 path: hi / lo split of the inverse, s x s each
  float* Wf;             // fp32 pair temporaries (slot p at p * b * b, ld b)
  bf16* Wt; bf16* Whi; bf16* Wlo; bf16* Z2;
};

struct GeomWs {
  KronWs k;
  void* C0; void* C1;    // m x n
  void* T2[2];           // per dense factor: term2, later E = term1 - term2 (s x s)
  float* t2vec[2];       // per factor (diagonal): term2 sums of squares (zeroed at PREPARE)
  float* qf[2];          // per factor: q (or 1/q) as fp32
  TriWs tri[2];
  char* zero2_begin; size_t zero2_bytes;
  size_t total;
};

static void layout_geom(const psgd_kron_t* k, int dq, void* base, GeomWs& g) {
  layout_kron(k, base, g.k);
  Bump b(base);
  b.off = g.k.total;
  const int es = dtype_size(k->dtype);
  const size_t m = k->m, n = k->has_r ? k->n : 1;
  const int sdim[2] = {k->m, k->has_r ? k->n : 0};
  const int dense[2] = {k->kind_l == PSGD_DENSE, k->has_r && k->kind_r == PSGD_DENSE};
  g.zero2_begin = (char*)b.take(0);
  const size_t z0 = b.off;
  for (int i = 0; i < 2; ++i) g.t2vec[i] = (float*)b.take((size_t)(sdim[i] > 0 ? sdim[i] : 1) * 4);
  g.zero2_bytes = b.off - z0;
  for (int i = 0; i < 2; ++i) g.qf[i] = (float*)b.take((size_t)(sdim[i] > 0 ? sdim[i] : 1) * 4);
  g.C0 = b.take(m * n * es); g.C1 = b.take(m * n * es);
  for (int i = 0; i < 2; ++i) {
    const size_t sd = dense[i] ? (size_t)sdim[i] : 0;
    g.T2[i] = b.take(sd * sd * es);
    TriWs& t = g.tri[i];
    t.Xf = nullptr; t.Xhi = t.Xlo = t.Wt = t.Whi = t.Wlo = t.Z2 = nullptr; t.Wf = nullptr;
    if (dq == PSGD_DQ_EQ && sd > 0) {
      size_t half = 64;   // pair temporaries: slots of b x b (ld b), the widest level decides
      for (size_t bb = TRI_NB; bb < sd; bb *= 2) {
        size_t pairs = 0;
        while (2 * pairs * bb + bb < sd) ++pairs;
        if (pairs * bb * bb > half) half = pairs * bb * bb;
      }
      if (k->dtype == PSGD_BF16) {
        t.Xhi = (bf16*)b.take(sd * sd * 2); t.Xlo = (bf16*)b.take(sd * sd * 2);
        t.Wt = (bf16*)b.take(half * 2); t.Whi = (bf16*)b.take(half * 2); t.Wlo = (bf16*)b.take(half * 2); t.Z2 = (bf16*)b.take(half * 2);
      }
      t.Xf = (float*)b.take(sd * sd * 4);
      t.Wf = (float*)b.take(half * 4);
    }
  }
  g.total = b.off;
}

// ---------------------------------------------------------------------------------------------
// exprA(*Q, X) = Q_L X Q_R^T (diagonal factors: scaling by q)   psgd.py:248-249
// reductions on the output as in run_chain
// ---------------------------------------------------------------------------------------------
static int run_apply_factors(Ctx* ctx, const psgd_kron_t* k, GeomWs& gw, const void* X, void* out, float* row_sumsq, float* col_sumsq,
                             float* total_sumsq, cudaStream_t st) {
  const int dt = k->dtype;
  const int m = k->m, n = k->has_r ? k->n : 1;
  const bool dl = k->kind_l == PSGD_DENSE, dr = k->has_r && k->kind_r == PSGD_DENSE;
  if (!dl) { DISPATCH_T(dt, (k_vec_to_f32<T><<<(m + 255) / 256, 256, 0, st>>>((const T*)k->QL, gw.qf[0], m, 0))); LAUNCH_CHECK(ctx, "k_vec_to_f32"); }
  if (k->has_r && !dr) { DISPATCH_T(dt, (k_vec_to_f32<T><<<(n + 255) / 256, 256, 0, st>>>((const T*)k->QR, gw.qf[1], n, 0))); LAUNCH_CHECK(ctx, "k_vec_to_f32"); }
  const float* rs = dl ? nullptr : gw.qf[0];
  const float* cs = (!k->has_r || dr) ? nullptr : gw.qf[1];
  if (!dl && !dr) {
    const size_t numel = (size_t)m * n;
    DISPATCH_T(dt, (k_scale2d<T><<<ew_blocks(ctx, numel), 256, 0, st>>>((const T*)X, (T*)out, m, n, rs, cs, row_sumsq, col_sumsq, total_sumsq)));
    LAUNCH_CHECK(ctx, "k_scale2d");
    return PSGD_OK;
  }
  auto set_final = [&](GemmDesc& gd) {
    gd.epi.row_scale = rs; gd.epi.col_scale = cs;
    gd.epi.row_sumsq = row_sumsq; gd.epi.col_sumsq = col_sumsq; gd.epi.total_sumsq = total_sumsq;
  };
  GemmDesc g;
  int rc;
  const void* Y = X;
  if (dl) {
    void* dst = dr ? gw.k.B2 : out;
    g = gemm_desc(dt, k->QL, m, 0, X, n, 0, m, n, m, dst, n);
    if (!dr) set_final(g);
    rc = launch_gemm(ctx, g, st); if (rc) return rc;
    Y = dst;
  }
  if (dr) {
    g = gemm_desc(dt, Y, n, 0, k->QR, n, 1, m, n, n, out, n);
    set_final(g);
    rc = launch_gemm(ctx, g, st); if (rc) return rc;
  }
  return PSGD_OK;
}

// ---------------------------------------------------------------------------------------------
// Blocked inverse of upper-triangular factors Q (s x s): leaves of TRI_NB inverted in shared memory (k_tri_inv_leaf), then level by level
//   inv([[A11, A12], [0, A22]]) = [[X11, -X11 A12 X22], [0, X22]].
// Levels below TRI_TC_MIN_B (and every level of fp32 factors) run in fp32 on CUDA cores, batched over the pairs (k_tri_pair_gemm): the
// products are tiny there and a tcgen05 launch costs 12-16 us whatever its size.  From TRI_TC_MIN_B on, bf16 factors switch to the tensor
// cores: the inverse is carried as a hi + lo pair of bf16 matrices (~16 mantissa bits), so every product is a bf16 GEMM yet the solve
// stays well below the bf16 rounding the reference applies to its fp32 solve (psgd.py:291):
//   W   = A12 X22  = A12 Xhi22 + [A12 Xlo22]_bf16                         (fp32 out, split into Whi + Wlo)
//   X12 = -X11 W   = -(Xhi11 Whi) - [Xhi11 Wlo + [Xlo11 Whi]_bf16]_bf16   (fp32 out, split into the hi / lo storage)
// The pairs of both factors of one update share the grouped launches (four problems per launch).
// ---------------------------------------------------------------------------------------------
constexpr int TRI_TC_MIN_B = 512;
struct TriJob { const void* Q; int s; TriWs* t; };

static int run_tri_inverse(Ctx* ctx, int dt, const TriJob* jobs, int nj, cudaStream_t st) {
  static PerDeviceOnce attr_done[2];
  int rc;
  int smax = 0;
  for (int jx = 0; jx < nj; ++jx) {
    const TriJob& J = jobs[jx];
    TriWs& t = *J.t;
    const int s = J.s;
    if (s > smax) smax = s;
    rc = check_cuda(ctx, cudaMemsetAsync(t.Xf, 0, (size_t)s * s * 4, st), "memset"); if (rc) return rc;
    const int ai = dt == PSGD_BF16 ? 0 : 1;
    if (attr_done[ai].need(ctx->device)) {
      if (dt == PSGD_BF16) cudaFuncSetAttribute(k_tri_inv_leaf<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TRI_LEAF_SMEM);
      else cudaFuncSetAttribute(k_tri_inv_leaf<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TRI_LEAF_SMEM);
    }
    DISPATCH_T(dt, (k_tri_inv_leaf<T><<<(s + TRI_NB - 1) / TRI_NB, 256, TRI_LEAF_SMEM, st>>>((const T*)J.Q, s, t.Xf)));
    LAUNCH_CHECK(ctx, "k_tri_inv_leaf");
    for (int b = TRI_NB; b < s && (dt == PSGD_F32 || b < TRI_TC_MIN_B); b *= 2) {
      int pairs = 0;
      while (2 * pairs * b + b < s) ++pairs;
      dim3 grid((b + 63) / 64, (b + 63) / 64, pairs);
      DISPATCH_T(dt, (k_tri_pair_gemm<T, 1><<<grid, 256, 0, st>>>((const T*)J.Q, t.Xf, t.Wf, s, b)));
      LAUNCH_CHECK(ctx, "k_tri_pair_gemm");
      DISPATCH_T(dt, (k_tri_pair_gemm<T, 2><<<grid, 256, 0, st>>>((const T*)J.Q, t.Xf, t.Wf, s, b)));
      LAUNCH_CHECK(ctx, "k_tri_pair_gemm");
    }
  }
  if (dt == PSGD_F32) return PSGD_OK;
  for (int jx = 0; jx < nj; ++jx) {
    const size_t numel = (size_t)jobs[jx].s * jobs[jx].s;
    k_split_full<<<ew_blocks(ctx, numel / 4 + 1), 256, 0, st>>>(jobs[jx].t->Xf, jobs[jx].t->Xhi, jobs[jx].t->Xlo, numel);
    LAUNCH_CHECK(ctx, "k_split_full");
  }
  // ---- tensor-core levels: hi / lo products, the pairs of all factors grouped ----
  for (int b = TRI_TC_MIN_B; b < smax; b *= 2) {
    struct Item { const char* Q; int s; TriWs* t; int p; };
    Item items[64];
    int ni = 0;
    int pairs_of[2] = {0, 0};
    for (int jx = 0; jx < nj; ++jx) {
      int pairs = 0;
      while (2 * pairs * b + b < jobs[jx].s) ++pairs;
      pairs_of[jx] = pairs;
      for (int p = 0; p < pairs && ni < 64; ++p) items[ni++] = Item{reinterpret_cast<const char*>(jobs[jx].Q), jobs[jx].s, jobs[jx].t, p};
    }
    const size_t slot = (size_t)b * b;
    auto for_items = [&](auto&& make) -> int {
      GemmDesc gs[4];
      int ng = 0;
      for (int it = 0; it < ni; ++it) {
        gs[ng++] = make(items[it]);
        if (ng == 4 || it + 1 == ni) { int r = launch_gemm_group(ctx, gs, ng, st); if (r) return r; ng = 0; }
      }
      return PSGD_OK;
    };
    auto b2_of = [&](const Item& I) { int c = I.s - (2 * I.p * b + b); return c < b ? c : b; };
    rc = for_items([&](const Item& I) {   // Wt = A12 Xlo22
      const size_t s = I.s, r1 = (size_t)2 * I.p * b, c2 = r1 + b; const int b2 = b2_of(I); TriWs& t = *I.t;
      return gemm_desc(dt, I.Q + (r1 * s + c2) * 2, I.s, 0, t.Xlo + c2 * s + c2, I.s, 0, b, b2, b2, t.Wt + I.p * slot, b); });
    if (rc) return rc;
    rc = for_items([&](const Item& I) {   // Wf = A12 Xhi22 + Wt   (fp32 out)
      const size_t s = I.s, r1 = (size_t)2 * I.p * b, c2 = r1 + b; const int b2 = b2_of(I); TriWs& t = *I.t;
      GemmDesc g = gemm_desc(dt, I.Q + (r1 * s + c2) * 2, I.s, 0, t.Xhi + c2 * s + c2, I.s, 0, b, b2, b2, t.Wf + I.p * slot, b);
      g.epi.out_dtype = PSGD_F32; g.epi.D = t.Wt + I.p * slot; g.epi.ldd = b; g.epi.d_dtype = PSGD_BF16; g.epi.beta = 1.f;
      return g; });
    if (rc) return rc;
    for (int jx = 0; jx < nj; ++jx) {
      if (!pairs_of[jx]) continue;
      TriWs& t = *jobs[jx].t;
      dim3 grid((unsigned)((slot + 255) / 256), pairs_of[jx]);
      k_split_hilo<<<grid, 256, 0, st>>>(t.Wf, slot, b, t.Whi, t.Wlo, slot, b, b, b, jobs[jx].s);
      LAUNCH_CHECK(ctx, "k_split_hilo");
    }
    rc = for_items([&](const Item& I) {   // Wt = Xlo11 Whi
      const size_t s = I.s, r1 = (size_t)2 * I.p * b; const int b2 = b2_of(I); TriWs& t = *I.t;
      return gemm_desc(dt, t.Xlo + r1 * s + r1, I.s, 0, t.Whi + I.p * slot, b, 0, b, b2, b, t.Wt + I.p * slot, b); });
    if (rc) return rc;
    rc = for_items([&](const Item& I) {   // Z2 = Xhi11 Wlo + Wt
      const size_t s = I.s, r1 = (size_t)2 * I.p * b; const int b2 = b2_of(I); TriWs& t = *I.t;
      GemmDesc g = gemm_desc(dt, t.Xhi + r1 * s + r1, I.s, 0, t.Wlo + I.p * slot, b, 0, b, b2, b, t.Z2 + I.p * slot, b);
      g.epi.D = t.Wt + I.p * slot; g.epi.ldd = b; g.epi.d_dtype = PSGD_BF16; g.epi.beta = 1.f;
      return g; });
    if (rc) return rc;
    rc = for_items([&](const Item& I) {   // Wf = -Xhi11 Whi - Z2   (fp32 out)
      const size_t s = I.s, r1 = (size_t)2 * I.p * b; const int b2 = b2_of(I); TriWs& t = *I.t;
      GemmDesc g = gemm_desc(dt, t.Xhi + r1 * s + r1, I.s, 0, t.Whi + I.p * slot, b, 0, b, b2, b, t.Wf + I.p * slot, b);
      g.epi.out_dtype = PSGD_F32; g.epi.alpha = -1.f; g.epi.D = t.Z2 + I.p * slot; g.epi.ldd = b; g.epi.d_dtype = PSGD_BF16; g.epi.beta = -1.f;
      return g; });
    if (rc) return rc;
    for (int jx = 0; jx < nj; ++jx) {
      if (!pairs_of[jx]) continue;
      TriWs& t = *jobs[jx].t;
      dim3 grid((unsigned)((slot + 255) / 256), pairs_of[jx]);
      k_split_hilo<<<grid, 256, 0, st>>>(t.Wf, slot, b, t.Xhi + b, t.Xlo + b, (size_t)2 * b * jobs[jx].s + 2 * b, jobs[jx].s, b, b, jobs[jx].s);
      LAUNCH_CHECK(ctx, "k_split_hilo");
    }
  }
  return PSGD_OK;
}

// conjB = Q_L^{-T} V Q_R^{-1} (diagonal factors: division)   psgd.py:297-303, each solve rounded to the tensor dtype like the reference.
// Needs run_tri_inverse of the dense factors first.  Sums of squares of the result for the diagonal factors' term2.
static int run_inverse_apply(Ctx* ctx, const psgd_kron_t* k, GeomWs& gw, const void* V, void* out, float* row_sumsq, float* col_sumsq,
                             cudaStream_t st) {
  const int dt = k->dtype;
  const int m = k->m, n = k->has_r ? k->n : 1;
  const bool dl = k->kind_l == PSGD_DENSE, dr = k->has_r && k->kind_r == PSGD_DENSE;
  if (!dl) { DISPATCH_T(dt, (k_vec_to_f32<T><<<(m + 255) / 256, 256, 0, st>>>((const T*)k->QL, gw.qf[0], m, 1))); LAUNCH_CHECK(ctx, "k_vec_to_f32"); }
  if (k->has_r && !dr) { DISPATCH_T(dt, (k_vec_to_f32<T><<<(n + 255) / 256, 256, 0, st>>>((const T*)k->QR, gw.qf[1], n, 1))); LAUNCH_CHECK(ctx, "k_vec_to_f32"); }
  const float* rs = dl ? nullptr : gw.qf[0];
  const float* cs = (!k->has_r || dr) ? nullptr : gw.qf[1];
  if (!dl && !dr) {
    const size_t numel = (size_t)m * n;
    DISPATCH_T(dt, (k_scale2d<T><<<ew_blocks(ctx, numel), 256, 0, st>>>((const T*)V, (T*)out, m, n, rs, cs, row_sumsq, col_sumsq, nullptr)));
    LAUNCH_CHECK(ctx, "k_scale2d");
    return PSGD_OK;
  }
  auto set_final = [&](GemmDesc& gd) {
    gd.epi.row_scale = rs; gd.epi.col_scale = cs;
    gd.epi.row_sumsq = row_sumsq; gd.epi.col_sumsq = col_sumsq;
    if (gd.epi.D) { gd.epi.d_row_scale = rs; gd.epi.d_col_scale = cs; }
  };
  GemmDesc g;
  int rc;
  const void* Y = V;
  void* tmp = gw.k.B2;   // lo-correction temporary
  if (dl) {
    void* dst = dr ? gw.C0 : out;
    if (dt == PSGD_BF16) {
      g = gemm_desc(dt, gw.tri[0].Xlo, m, 1, V, n, 0, m, n, m, tmp, n);
      rc = launch_gemm(ctx, g, st); if (rc) return rc;
      g = gemm_desc(dt, gw.tri[0].Xhi, m, 1, V, n, 0, m, n, m, dst, n);
      g.epi.D = tmp; g.epi.ldd = n; g.epi.d_dtype = dt; g.epi.beta = 1.f;
    } else {
      g = gemm_desc(dt, gw.tri[0].Xf, m, 1, V, n, 0, m, n, m, dst, n);
    }
    if (!dr) set_final(g);
    rc = launch_gemm(ctx, g, st); if (rc) return rc;
    Y = dst;
  }
  if (dr) {
    if (dt == PSGD_BF16) {
      g = gemm_desc(dt, Y, n, 0, gw.tri[1].Xlo, n, 0, m, n, n, tmp, n);
      rc = launch_gemm(ctx, g, st); if (rc) return rc;
      g = gemm_desc(dt, Y, n, 0, gw.tri[1].Xhi, n, 0, m, n, n, out, n);
      g.epi.D = tmp; g.epi.ldd = n; g.epi.d_dtype = dt; g.epi.beta = 1.f;
    } else {
      g = gemm_desc(dt, Y, n, 0, gw.tri[1].Xf, n, 0, m, n, n, out, n);
    }
    set_final(g);
    rc = launch_gemm(ctx, g, st); if (rc) return rc;
  }
  return PSGD_OK;
}

// exprGs[i](X, X*) for a dense factor: rows (i = 0: X X^T, m x m) or columns (i = 1: X^T X, n x n) Gram of an m x n matrix
static GemmDesc gram_desc(int dt, const void* X, int m, int n, int i, void* out) {
  GemmDesc g = (i == 0) ? gemm_desc(dt, X, n, 0, X, n, 1, m, m, n, out, m) : gemm_desc(dt, X, n, 1, X, n, 0, n, n, m, out, n);
  g.sym = 1;
  return g;
}

static inline bool dq_fits_p(int dq) { return dq == PSGD_DQ_QUAD4P || dq == PSGD_DQ_PRO4P; }
static inline bool dq_uses_expr_a(int dq) { return dq == PSGD_DQ_EQ || dq_fits_p(dq); }

// ---------------------------------------------------------------------------------------------
// PREPARE: everything before the per-factor loop of the reference -- damped input, Pg (or A), conjB, all term1 / term2 contractions.
// Afterwards, per factor i: dense: S[i][0] = term1 (scalar term2) or term1 + term2 with E = term1 - term2 in T2[i]; row_sumsq / diag_max of
// S[i][0] in k.f[i]; diagonal: k.f[i].term1 and t2vec[i].
// ---------------------------------------------------------------------------------------------
static int geom_prepare(Ctx* ctx, const psgd_kron_t* k, int dq, const void* X, const void* V, float damping, const psgd_kron_noise_t* noise,
                        GeomWs& gw, cudaStream_t st) {
  KronWs& w = gw.k;
  const int dt = k->dtype;
  const int m = k->m, n = k->has_r ? k->n : 1;
  const size_t numel = (size_t)m * n;
  const bool dense[2] = {k->kind_l == PSGD_DENSE, k->has_r && k->kind_r == PSGD_DENSE};
  const int nf = k->has_r ? 2 : 1;
  const bool newton = V != nullptr;
  const bool matrix_t2 = newton || dq == PSGD_DQ_EQ || dq == PSGD_DQ_QEP;
  int rc = check_cuda(ctx, cudaMemsetAsync(w.zero_begin, 0, w.zero_bytes, st), "memset"); if (rc) return rc;
  rc = check_cuda(ctx, cudaMemsetAsync(gw.zero2_begin, 0, gw.zero2_bytes, st), "memset"); if (rc) return rc;
  if (dq == PSGD_DQ_QEP) { rc = run_balance(ctx, k, w, st); if (rc) return rc; }   // psgd.py:347 / 674: not optional for this geometry

  // H = X + (damping + eps |X|) N      psgd.py:334-335, 352-353, 661 ...   (noise->N == NULL: the raw pair update of psgd.py:278, H = X)
  const void* H = X;
  if (noise->N) {
    DISPATCH_T(dt, (k_add_noise<T><<<ew_blocks(ctx, numel), 256, 0, st>>>((const T*)X, (const T*)noise->N, (T*)w.B0, numel, damping, dtype_eps(dt))));
    LAUNCH_CHECK(ctx, "k_add_noise");
    H = w.B0;
  }
  if (dq == PSGD_DQ_EQ && !newton) V = noise->N;   // psgd.py:334: the probe doubles as the damping noise

  // Pg = P H (exprP) or A = exprA(*Q, H); sums of squares for the diagonal factors' term1 fused into the last product
  void* Pg = w.B1;
  float* rsq = dense[0] ? nullptr : w.f[0].term1;
  float* csq = (k->has_r && !dense[1]) ? w.f[1].term1 : nullptr;
  if (dq_uses_expr_a(dq)) rc = run_apply_factors(ctx, k, gw, H, Pg, rsq, csq, nullptr, st);
  else rc = run_chain(ctx, k, w, H, Pg, rsq, csq, nullptr, st);
  if (rc) return rc;

  // ---- term1 of the dense factors ----
  for (int i = 0; i < nf; ++i) {
    if (!dense[i]) continue;
    const void* src = Pg;
    if (dq == PSGD_DQ_QEP) {   // exprQs[i](q, Pg): psgd.py:355 / 678
      GemmDesc g = (i == 0) ? gemm_desc(dt, k->QL, m, 0, Pg, n, 0, m, n, m, gw.C0, n) : gemm_desc(dt, Pg, n, 0, k->QR, n, 1, m, n, n, gw.C0, n);
      rc = launch_gemm(ctx, g, st); if (rc) return rc;
      src = gw.C0;
    }
    GemmDesc g = gram_desc(dt, src, m, n, i, w.S[i][0]);
    if (!matrix_t2) { g.epi.row_sumsq = w.f[i].row_sumsq; g.epi.diag_max = w.f[i].diag_max; }
    rc = launch_gemm(ctx, g, st); if (rc) return rc;
  }
  if (!matrix_t2) return PSGD_OK;

  // ---- term2 ----
  const float t2s[2] = {(float)((double)numel / (double)m), (float)((double)numel / (double)n)};
  if (dq == PSGD_DQ_EQ) {
    TriJob jobs[2];
    int nj = 0;
    for (int i = 0; i < nf; ++i)
      if (dense[i]) jobs[nj++] = TriJob{i == 0 ? k->QL : k->QR, i == 0 ? m : n, &gw.tri[i]};
    rc = run_tri_inverse(ctx, dt, jobs, nj, st); if (rc) return rc;
    rc = run_inverse_apply(ctx, k, gw, V, gw.C1, dense[0] ? nullptr : gw.t2vec[0], (k->has_r && !dense[1]) ? gw.t2vec[1] : nullptr, st);
    if (rc) return rc;
    for (int i = 0; i < nf; ++i)
      if (dense[i]) { GemmDesc g = gram_desc(dt, gw.C1, m, n, i, gw.T2[i]); rc = launch_gemm(ctx, g, st); if (rc) return rc; }
  } else if (!newton) {   // whitening QEP: term2 = numel/s q q^T (dense), numel/s q^2 (diagonal: formed inside k_diag_update_gen)
    for (int i = 0; i < nf; ++i) {
      if (!dense[i]) continue;
      const int s = i == 0 ? m : n;
      const void* q = i == 0 ? k->QL : k->QR;
      GemmDesc g = gemm_desc(dt, q, s, 0, q, s, 1, s, s, s, gw.T2[i], s);
      g.sym = 1; g.epi.alpha = t2s[i];
      rc = launch_gemm(ctx, g, st); if (rc) return rc;
    }
  } else {                // Newton pairs: term2 = exprGs[i](V, V*) (QEP: of exprQs[i](q, V))   psgd.py:679-680, 703, 728 ...
    if (!dense[0] || (k->has_r && !dense[1])) {
      DISPATCH_T(dt, (k_scale2d<T><<<ew_blocks(ctx, numel), 256, 0, st>>>((const T*)V, (T*)nullptr, m, n, nullptr, nullptr,
                                                                         dense[0] ? nullptr : gw.t2vec[0],
                                                                         (k->has_r && !dense[1]) ? gw.t2vec[1] : nullptr, nullptr)));
      LAUNCH_CHECK(ctx, "k_scale2d");
    }
    for (int i = 0; i < nf; ++i) {
      if (!dense[i]) continue;
      const void* src = V;
      if (dq == PSGD_DQ_QEP) {
        GemmDesc g = (i == 0) ? gemm_desc(dt, k->QL, m, 0, V, n, 0, m, n, m, gw.C0, n) : gemm_desc(dt, V, n, 0, k->QR, n, 1, m, n, n, gw.C0, n);
        rc = launch_gemm(ctx, g, st); if (rc) return rc;
        src = gw.C0;
      }
      GemmDesc g = gram_desc(dt, src, m, n, i, gw.T2[i]);
      rc = launch_gemm(ctx, g, st); if (rc) return rc;
    }
  }
  // S = term1 + term2 (with the bound's row norms / max diagonal), E = term1 - term2 (upper triangle for dQ = E*Q)
  for (int i = 0; i < nf; ++i) {
    if (!dense[i]) continue;
    const int s = i == 0 ? m : n;
    const int triu = dq == PSGD_DQ_EQ ? 1 : 0;
    if (dt == PSGD_BF16 && s % 8 == 0 && ((reinterpret_cast<uintptr_t>(w.S[i][0]) | reinterpret_cast<uintptr_t>(gw.T2[i])) & 15u) == 0) {
      k_combine_terms_bf16x8<<<s, 256, 0, st>>>((bf16*)w.S[i][0], (bf16*)gw.T2[i], s, triu, w.f[i].row_sumsq, w.f[i].diag_max);
    } else {
      DISPATCH_T(dt, (k_combine_terms<T><<<s, 256, 0, st>>>((T*)w.S[i][0], (T*)gw.T2[i], s, triu, w.f[i].row_sumsq, w.f[i].diag_max)));
    }
    LAUNCH_CHECK(ctx, "k_combine_terms");
  }
  return PSGD_OK;
}

// ---------------------------------------------------------------------------------------------
// one factor's bound, Lipschitz update and step
// ---------------------------------------------------------------------------------------------
struct DenseStep {          // one dense factor entering its step: S = term1 (scalar term2) or term1 + term2, E = term1 - term2 (or null)
  int s; void* q; float* L; float t2;
  void* S; void* E; void* Qn; void* RQ; void* RRQ; void* Va; void* Vb;
  const void* v_spd; const void* v_skh;
  FactorWs* f;              // row_sumsq / diag_max of S already reduced
};

// norm bound of S, Lipschitz update, then the geometry's step on q (psgd.py:315-316, 362-364, 386-388, 413-416, 440-449, 474-479 ...)
static int geom_dense_step(Ctx* ctx, int dt, int dq, const DenseStep& d, float lr, float betaL, cudaStream_t st) {
  const int s = d.s;
  FactorWs& f = *d.f;
  const bool matrix_t2 = d.E != nullptr;
  const bool quad = dq == PSGD_DQ_QUAD || dq == PSGD_DQ_QUAD4P;
  const float lr_eff = dq == PSGD_DQ_QUAD ? 0.5f * lr : lr;   // psgd.py:470 / 476: lr/2/L
  if (!d.v_spd) return PSGD_ERR_INVALID_ARG;
  BoundJob jb{d.S, s, d.v_spd, f.row_sumsq, f.diag_max, &f.b_spd, d.Va, d.Vb};
  BoundFinish fin{0, matrix_t2 ? 0.f : d.t2, lr_eff, betaL, d.L, f.fs};
  int rc = run_bounds(ctx, dt, &jb, 1, &fin, st); if (rc) return rc;
  // fs[FS_ALPHA] = -c, fs[FS_BETA] = 1 + c t2 (1 when term2 is a matrix): one product gives beta * q + alpha * (M q) = q - c (term1 - term2) q
  const void* M = matrix_t2 ? d.E : d.S;
  auto step_desc = [&](const void* src, void* dst, bool left) {
    GemmDesc g = left ? gemm_desc(dt, M, s, 0, src, s, 0, s, s, s, dst, s) : gemm_desc(dt, src, s, 0, M, s, 0, s, s, s, dst, s);
    g.epi.alpha_ptr = f.fs + FS_ALPHA; g.epi.D = src; g.epi.ldd = s; g.epi.d_dtype = dt; g.epi.beta = 1.f; g.epi.beta_ptr = f.fs + FS_BETA;
    return g;
  };
  GemmDesc g = step_desc(d.q, d.Qn, dq != PSGD_DQ_QEQ);   // QEQ: q - c q (term1 - term2)   psgd.py:388 / 713
  rc = launch_gemm(ctx, g, st); if (rc) return rc;
  if (dq == PSGD_DQ_Q0P5EQ1P5) {                         // psgd.py:416 / 738
    if (!d.v_skh) return PSGD_ERR_INVALID_ARG;
    DenseItem it;
    it.s = s; it.q = d.q; it.L = d.L; it.t2 = d.t2; it.T = d.S; it.Qn = d.Qn; it.RQ = d.RQ; it.RRQ = d.RRQ; it.Va = d.Va; it.Vb = d.Vb;
    it.v_spd = d.v_spd; it.v_skh = d.v_skh; it.f = &f;
    return run_procrustes(ctx, dt, &it, 1, 0.125f, st);
  }
  if (quad) {                                            // p = p - c p (term1 - term2); q = (p + p^T)/2   psgd.py:477-479 / 792-794
    g = step_desc(d.Qn, d.RQ, false);
    rc = launch_gemm(ctx, g, st); if (rc) return rc;
    dim3 grid((s + 31) / 32, (s + 31) / 32), block(32, 8);
    DISPATCH_T(dt, (k_symmetrize<T><<<grid, block, 0, st>>>((const T*)d.RQ, (T*)d.q, s)));
    LAUNCH_CHECK(ctx, "k_symmetrize");
    return PSGD_OK;
  }
  return check_cuda(ctx, cudaMemcpyAsync(d.q, d.Qn, (size_t)s * s * dtype_size(dt), cudaMemcpyDeviceToDevice, st), "memcpy");
}

static int geom_factor(Ctx* ctx, const psgd_kron_t* k, int dq, bool newton, int i, float lr, float betaL, const psgd_kron_noise_t* noise,
                       GeomWs& gw, cudaStream_t st) {
  KronWs& w = gw.k;
  const int dt = k->dtype;
  const int m = k->m, n = k->has_r ? k->n : 1;
  const size_t numel = (size_t)m * n;
  const int s = i == 0 ? m : n;
  const bool dense = i == 0 ? k->kind_l == PSGD_DENSE : (k->has_r && k->kind_r == PSGD_DENSE);
  void* q = i == 0 ? k->QL : k->QR;
  float* L = i == 0 ? k->LL : k->LR;
  FactorWs& f = w.f[i];
  const bool matrix_t2 = newton || dq == PSGD_DQ_EQ || dq == PSGD_DQ_QEP;
  const float t2 = (float)((double)numel / (double)s);
  const bool quad = dq == PSGD_DQ_QUAD || dq == PSGD_DQ_QUAD4P;
  const float lr_eff = dq == PSGD_DQ_QUAD ? 0.5f * lr : lr;   // psgd.py:470 / 476: lr/2/L
  int rc;
  if (!dense) {
    const bool qep = dq == PSGD_DQ_QEP;
    // whitening: term2 = numel/s (QEP: numel/s q^2); Newton / EQ: sums of squares in t2vec (QEP: times q^2); QEP term1 = q^2 * sumsq(Pg)
    const float* t2v = (newton || dq == PSGD_DQ_EQ) ? gw.t2vec[i] : nullptr;
    DISPATCH_T(dt, (k_diag_update_gen<T><<<1, 1024, 0, st>>>((T*)q, f.term1, t2v, t2, qep ? 1 : 0, qep ? 1 : 0, s, lr_eff, betaL, L, quad ? 1 : 0)));
    LAUNCH_CHECK(ctx, "k_diag_update_gen");
    return PSGD_OK;
  }
  DenseStep d;
  d.s = s; d.q = q; d.L = L; d.t2 = t2; d.S = w.S[i][0]; d.E = matrix_t2 ? gw.T2[i] : nullptr; d.Qn = w.S[i][1]; d.RQ = w.S[i][2]; d.RRQ = w.S[i][3];
  d.Va = w.Va[i]; d.Vb = w.Vb[i]; d.f = &f;
  d.v_spd = i == 0 ? noise->V0_spd_l : noise->V0_spd_r;
  d.v_skh = i == 0 ? noise->V0_skh_l : noise->V0_skh_r;
  return geom_dense_step(ctx, dt, dq, d, lr, betaL, st);
}
