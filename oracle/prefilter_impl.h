/*
 * TEST INFRASTRUCTURE — NOT PRODUCT CODE.
 *
 * Type-generic body of the prefilter oracle; included twice by oracle.c with
 * REAL = float / double and NAME(x) = x##_f32 / x##_f64.
 *
 * A literal CPU restatement of Interpolations.jl v0.16.3 (paths relative to the
 * reference checkout):
 *   - system assembly      src/b-splines/cubic.jl:101-264, quadratic.jl:71-189
 *   - lut!                 src/b-splines/b-splines.jl:257  (LinearAlgebra.lu!(Tridiagonal, NoPivot()))
 *   - Woodbury constructor WoodburyMatrices (un-vendored, compat 0.4/0.5/1.0):
 *                          Cp = inv(inv(C) + V*(A\U))
 *   - A_ldiv_B_md!         src/filter1d.jl:8-85 (LU sweeps :21-73, Woodbury 6-step :77-85)
 * Operation order and rounding points follow those files; compiled with
 * -ffp-contract=off so no multiply-add is fused (Julia does not contract).
 * PARITY UNPINNED at the last ulp for the third-party parts (Woodbury's `Cp`, and
 * AxisAlgorithms' summation order in `_A_mul_B_md`), which are not in the checkout.
 */

/* convert(T, SimpleRatio(num, den)) == T(num)/T(den)   (Ratios.jl; cubic.jl:102-103) */
#define RATIO(num, den) ((REAL)(num) / (REAL)(den))

/* prefiltering_system + lut!.  dl[n-1], d[n], du[n-1]; Woodbury triplets (urow, vcol, cval), *k of them.
 * returns 0, or 2 (unsupported: the reference has no method, prefiltering.jl:124). */
static int NAME(orc_system)(int64_t n, int degree, int bc, int grid, REAL *dl, REAL *d, REAL *du,
                            int *k, int *urow, int *vcol, REAL *cval) {
  REAL a, b;
  *k = 0;
  if (degree == B200ITP_CUBIC) { /* cubic.jl:101-106 */
    a = RATIO(1, 6);
    b = RATIO(2, 3);
  } else if (degree == B200ITP_QUADRATIC) { /* quadratic.jl:71-76 */
    a = RATIO(1, 8);
    b = RATIO(3, 4);
  } else {
    return 2;
  }
  if (n < 2) return 2;
  for (int64_t i = 0; i < n - 1; i++) dl[i] = du[i] = a;
  for (int64_t i = 0; i < n; i++) d[i] = b;

#define SETROWS(dv, ov)            \
  do {                             \
    d[0] = d[n - 1] = (REAL)(dv);  \
    du[0] = dl[n - 2] = (REAL)(ov); \
  } while (0)
#define WB(i, j, v)        \
  do {                     \
    urow[*k] = (int)(i);   \
    vcol[*k] = (int)(j);   \
    cval[*k] = (REAL)(v);  \
    (*k)++;                \
  } while (0)

  if (degree == B200ITP_CUBIC) {
    if (bc == B200ITP_BC_FLAT && grid == B200ITP_ONGRID) { /* cubic.jl:114-124 */
      SETROWS(-1, 0);
      WB(1, 3, 1);
      WB(n, n - 2, 1);
    } else if (bc == B200ITP_BC_FLAT && grid == B200ITP_ONCELL) { /* cubic.jl:140-157 */
      SETROWS(-9, 11);
      WB(1, 3, -3);
      WB(n, n - 2, -3);
      WB(1, 4, 1);
      WB(n, n - 3, 1);
    } else if (bc == B200ITP_BC_LINE && grid == B200ITP_ONCELL) { /* cubic.jl:169-186 */
      SETROWS(3, -7);
      WB(1, 3, 5);
      WB(n, n - 2, 5);
      WB(1, 4, -1);
      WB(n, n - 3, -1);
    } else if (bc == B200ITP_BC_LINE && grid == B200ITP_ONGRID) { /* cubic.jl:194-208 */
      SETROWS(1, -2);
      WB(1, 3, 1);
      WB(n, n - 2, 1);
    } else if (bc == B200ITP_BC_PERIODIC) { /* cubic.jl:218-228 */
      WB(1, n, du[0]);
      WB(n, 1, dl[n - 2]);
    } else if (bc == B200ITP_BC_FREE) { /* cubic.jl:242-264 */
      SETROWS(-1, 4);
      WB(1, 3, -6);
      WB(1, 4, 4);
      WB(1, 5, -1);
      WB(n, n - 2, -6);
      WB(n, n - 3, 4);
      WB(n, n - 4, -1);
    } else {
      return 2; /* Cubic(Reflect), Cubic(Throw): no method */
    }
  } else {
    if ((bc == B200ITP_BC_FLAT || bc == B200ITP_BC_REFLECT) && grid == B200ITP_ONCELL) {
      SETROWS(-1, 1); /* quadratic.jl:84-89, bare LU */
    } else if ((bc == B200ITP_BC_FLAT || bc == B200ITP_BC_REFLECT) && grid == B200ITP_ONGRID) {
      SETROWS(-1, 0); /* quadratic.jl:118-129 */
      WB(1, 3, 1);
      WB(n, n - 2, 1);
    } else if (bc == B200ITP_BC_LINE) { /* quadratic.jl:138-149 */
      SETROWS(1, -2);
      WB(1, 3, 1);
      WB(n, n - 2, 1);
    } else if (bc == B200ITP_BC_FREE) { /* quadratic.jl:158-170 */
      SETROWS(1, -3);
      WB(1, 3, 3);
      WB(1, 4, -1);
      WB(n, n - 2, 3);
      WB(n, n - 3, -1);
    } else if (bc == B200ITP_BC_PERIODIC) { /* quadratic.jl:180-189 */
      WB(1, n, du[0]);
      WB(n, 1, dl[n - 2]);
    } else if (bc == B200ITP_BC_INPLACE && grid == B200ITP_ONCELL) { /* quadratic.jl:91-95 */
      d[0] = d[n - 1] = RATIO(7, 8);
    } else if (bc == B200ITP_BC_INPLACEQ && grid == B200ITP_ONCELL) { /* quadratic.jl:98-110 */
      d[0] = d[n - 1] = RATIO(9, 8);
      dl[n - 2] = du[0] = RATIO(-1, 4);
      WB(1, 3, RATIO(1, 8));
      WB(n, n - 2, RATIO(1, 8));
    } else {
      return 2;
    }
  }
#undef SETROWS
#undef WB
  for (int j = 0; j < *k; j++)
    if (vcol[j] < 1 || vcol[j] > n) return 4; /* array too short for this BC */

  /* lut!(dl, d, du) = lu!(Tridiagonal(dl, d, du), NoPivot())  (b-splines.jl:257) */
  for (int64_t i = 0; i < n - 1; i++) {
    if (d[i] != 0) {
      REAL fact = dl[i] / d[i];
      dl[i] = fact;
      REAL t = fact * du[i];
      d[i + 1] = d[i + 1] - t;
    }
  }
  return 0;
}

/* LU solve the way LAPACK ?gttrs does it (used by `A \ U` inside the Woodbury constructor) */
static void NAME(orc_gttrs)(int64_t n, const REAL *dl, const REAL *d, const REAL *du, REAL *b) {
  for (int64_t i = 0; i < n - 1; i++) {
    REAL t = dl[i] * b[i];
    b[i + 1] = b[i + 1] - t;
  }
  b[n - 1] = b[n - 1] / d[n - 1];
  for (int64_t i = n - 2; i >= 0; i--) {
    REAL t = du[i] * b[i + 1];
    b[i] = (b[i] - t) / d[i];
  }
}

/* dense k x k inverse, Gaussian elimination with partial pivoting, in place (column-major) */
static int NAME(orc_inv)(int k, REAL *M) {
  REAL aug[B200ITP_MAX_WOODBURY][2 * B200ITP_MAX_WOODBURY];
  for (int i = 0; i < k; i++)
    for (int j = 0; j < k; j++) {
      aug[i][j] = M[i + j * k];
      aug[i][k + j] = (i == j) ? (REAL)1 : (REAL)0;
    }
  for (int c = 0; c < k; c++) {
    int p = c;
    for (int r = c + 1; r < k; r++)
      if (fabs((double)aug[r][c]) > fabs((double)aug[p][c])) p = r;
    if (aug[p][c] == 0) return 1;
    if (p != c)
      for (int j = 0; j < 2 * k; j++) {
        REAL t = aug[p][j];
        aug[p][j] = aug[c][j];
        aug[c][j] = t;
      }
    REAL piv = aug[c][c];
    for (int j = 0; j < 2 * k; j++) aug[c][j] = aug[c][j] / piv;
    for (int r = 0; r < k; r++) {
      if (r == c) continue;
      REAL f = aug[r][c];
      if (f == 0) continue;
      for (int j = 0; j < 2 * k; j++) {
        REAL t = f * aug[c][j];
        aug[r][j] = aug[r][j] - t;
      }
    }
  }
  for (int i = 0; i < k; i++)
    for (int j = 0; j < k; j++) M[i + j * k] = aug[i][k + j];
  return 0;
}

/* Woodbury(A, U, C, V): Cp = inv(inv(C) + V*(A\U)), k x k column-major */
static int NAME(orc_woodbury_cp)(int64_t n, const REAL *dl, const REAL *d, const REAL *du, int k,
                                 const int *urow, const int *vcol, const REAL *cval, REAL *Cp) {
  if (k == 0) return 0;
  REAL *z = (REAL *)malloc(sizeof(REAL) * (size_t)n);
  if (!z) return -1;
  for (int i = 0; i < k * k; i++) Cp[i] = 0;
  for (int j = 0; j < k; j++) { /* column j of A\U */
    for (int64_t i = 0; i < n; i++) z[i] = 0;
    z[urow[j] - 1] = 1;
    NAME(orc_gttrs)(n, dl, d, du, z);
    for (int i = 0; i < k; i++) Cp[i + j * k] = z[vcol[i] - 1]; /* (V*Z)[i,j] */
  }
  for (int j = 0; j < k; j++) Cp[j + j * k] = ((REAL)1 / cval[j]) + Cp[j + j * k]; /* inv(C) + V*Z */
  free(z);
  return NAME(orc_inv)(k, Cp);
}

/* filter1d.jl:21-73 for one slab: x viewed as [R1, n] with row stride R1 (R1 == 1 is the dim-1 case).
 * src === dest, b == 0 (adding b[i] == 0 is kept: -0.0 + 0.0 == +0.0 like the reference). */
static void NAME(orc_lu_sweep)(REAL *x, int64_t R1, int64_t i1lo, int64_t i1hi, int64_t n,
                               const REAL *dl, const REAL *dinv, const REAL *du, int addb) {
  const REAL zero = 0;
  if (addb)
    for (int64_t I1 = i1lo; I1 < i1hi; I1++) x[I1] = x[I1] + zero;
  for (int64_t i = 1; i < n; i++) {
    REAL *xi = x + i * R1, *xm = x + (i - 1) * R1;
    const REAL l = dl[i - 1];
    if (addb) {
      for (int64_t I1 = i1lo; I1 < i1hi; I1++) {
        REAL t = l * xm[I1];
        xi[I1] = (xi[I1] + zero) - t;
      }
    } else {
      for (int64_t I1 = i1lo; I1 < i1hi; I1++) {
        REAL t = l * xm[I1];
        xi[I1] = xi[I1] - t;
      }
    }
  }
  {
    REAL *xn = x + (n - 1) * R1;
    const REAL r = dinv[n - 1];
    for (int64_t I1 = i1lo; I1 < i1hi; I1++) xn[I1] = xn[I1] * r;
  }
  for (int64_t i = n - 2; i >= 0; i--) {
    REAL *xi = x + i * R1, *xp = x + (i + 1) * R1;
    const REAL u = du[i], r = dinv[i];
    for (int64_t I1 = i1lo; I1 < i1hi; I1++) {
      REAL t = u * xp[I1];
      xi[I1] = (xi[I1] - t) * r;
    }
  }
}

/* A_ldiv_B_md!(coefs, M, coefs, dim, zeros(n))  (filter1d.jl:8-18,77-85), dim 0-based */
static int NAME(orc_solve_axis)(REAL *A, int ndims, const int64_t *size, int dim, const REAL *dl,
                                const REAL *d, const REAL *du, int k, const int *urow,
                                const int *vcol, const REAL *Cp, int nthreads) {
  int64_t n = size[dim], R1 = 1, R2 = 1;
  for (int i = 0; i < dim; i++) R1 *= size[i];
  for (int i = dim + 1; i < ndims; i++) R2 *= size[i];
  REAL *dinv = (REAL *)malloc(sizeof(REAL) * (size_t)n);
  if (!dinv) return -1;
  for (int64_t i = 0; i < n; i++) dinv[i] = (REAL)1 / d[i]; /* filter1d.jl:31,51 */
  /* work items: (I2, block of I1) */
  int64_t blk = 256;
  if (R1 < blk) blk = R1;
  int64_t nb1 = (R1 + blk - 1) / blk;
  int64_t nitems = nb1 * R2;
  int err = 0;
  (void)nthreads;
#pragma omp parallel num_threads(nthreads > 0 ? nthreads : 1)
  {
    REAL *tmp3 = k > 0 ? (REAL *)malloc(sizeof(REAL) * (size_t)(blk * n)) : NULL;
    REAL tmp1[B200ITP_MAX_WOODBURY], tmp2[B200ITP_MAX_WOODBURY];
    if (k > 0 && !tmp3) {
#pragma omp atomic write
      err = -1;
    }
#pragma omp for schedule(dynamic, 4)
    for (int64_t it = 0; it < nitems; it++) {
      if (k > 0 && !tmp3) continue;
      int64_t I2 = it / nb1, b1 = it % nb1;
      int64_t lo = b1 * blk, hi = lo + blk;
      if (hi > R1) hi = R1;
      REAL *x = A + I2 * R1 * n;
      /* step 1: LU solve in place, with the offset b == zeros (filter1d.jl:78) */
      NAME(orc_lu_sweep)(x, R1, lo, hi, n, dl, dinv, du, 1);
      if (k == 0) continue;
      int64_t w = hi - lo;
      for (int64_t I1 = lo; I1 < hi; I1++) {
        /* step 2: tmp1 = V*dest ; step 3: tmp2 = Cp*tmp1 (filter1d.jl:79-80) */
        for (int j = 0; j < k; j++) {
          REAL t = (REAL)1 * x[(int64_t)(vcol[j] - 1) * R1 + I1];
          tmp1[j] = (REAL)0 + t;
        }
        for (int i = 0; i < k; i++) tmp2[i] = 0;
        for (int j = 0; j < k; j++)
          for (int i = 0; i < k; i++) {
            REAL t = Cp[i + j * k] * tmp1[j];
            tmp2[i] = tmp2[i] + t;
          }
        /* step 4: tmp3 = U*tmp2 (full length, zero except rows urow) (filter1d.jl:81) */
        REAL *t3 = tmp3 + (I1 - lo);
        for (int64_t i = 0; i < n; i++) t3[i * w] = 0;
        for (int j = 0; j < k; j++) {
          REAL t = (REAL)1 * tmp2[j];
          t3[(int64_t)(urow[j] - 1) * w] = t3[(int64_t)(urow[j] - 1) * w] + t;
        }
      }
      /* step 5: tmp3 = A \ tmp3 (second full sweep, no offset) (filter1d.jl:83) */
      NAME(orc_lu_sweep)(tmp3, w, 0, w, n, dl, dinv, du, 0);
      /* step 6: dest -= tmp3 (filter1d.jl:84) */
      for (int64_t i = 0; i < n; i++) {
        REAL *xi = x + i * R1 + lo;
        const REAL *ti = tmp3 + i * w;
        for (int64_t q = 0; q < w; q++) xi[q] = xi[q] - ti[q];
      }
    }
    free(tmp3);
  }
  free(dinv);
  return err;
}

/* copy_with_padding (prefiltering.jl:17-28): dst zero-filled, interior = src */
static void NAME(orc_pad)(const REAL *src, REAL *dst, int ndims, const int64_t *ssz, const int *pad) {
  int64_t dsz[B200ITP_MAX_DIMS], total = 1, stot = 1;
  for (int i = 0; i < ndims; i++) {
    dsz[i] = ssz[i] + 2 * pad[i];
    total *= dsz[i];
    stot *= ssz[i];
  }
  for (int64_t i = 0; i < total; i++) dst[i] = 0;
  int64_t idx[B200ITP_MAX_DIMS] = {0, 0, 0, 0};
  int64_t nrow = stot / ssz[0];
  for (int64_t r = 0; r < nrow; r++) {
    int64_t off = pad[0], mul = dsz[0];
    for (int i = 1; i < ndims; i++) {
      off += (idx[i] + pad[i]) * mul;
      mul *= dsz[i];
    }
    memcpy(dst + off, src + r * ssz[0], sizeof(REAL) * (size_t)ssz[0]);
    for (int i = 1; i < ndims; i++) {
      if (++idx[i] < ssz[i]) break;
      idx[i] = 0;
    }
  }
}

/* prefilter!(TW, ret, it) (prefiltering.jl:49-59,108-122) on an already padded array */
static int NAME(orc_prefilter_inplace)(REAL *coefs, int ndims, const int64_t *size,
                                       const int *degree, const int *bc, const int *grid,
                                       int nthreads) {
  for (int dim = 0; dim < ndims; dim++) {
    if (degree[dim] != B200ITP_CUBIC && degree[dim] != B200ITP_QUADRATIC) continue;
    int64_t n = size[dim];
    REAL *dl = (REAL *)malloc(sizeof(REAL) * (size_t)(3 * n + 3));
    if (!dl) return -1;
    REAL *d = dl + n, *du = d + n + 1;
    int k, urow[B200ITP_MAX_WOODBURY], vcol[B200ITP_MAX_WOODBURY];
    REAL cval[B200ITP_MAX_WOODBURY], Cp[B200ITP_MAX_WOODBURY * B200ITP_MAX_WOODBURY];
    int rc = NAME(orc_system)(n, degree[dim], bc[dim], grid[dim], dl, d, du, &k, urow, vcol, cval);
    if (rc == 2) { /* no method: M === nothing, axis left as is (prefiltering.jl:124, filter1d.jl:6) */
      free(dl);
      continue;
    }
    if (rc) {
      free(dl);
      return rc;
    }
    rc = NAME(orc_woodbury_cp)(n, dl, d, du, k, urow, vcol, cval, Cp);
    if (!rc) rc = NAME(orc_solve_axis)(coefs, ndims, size, dim, dl, d, du, k, urow, vcol, Cp, nthreads);
    free(dl);
    if (rc) return rc;
  }
  return 0;
}

#undef RATIO
