// Host-side assembly of the prefiltering systems and of the compact per-axis plan the
// kernels consume.
//
// Reference (paths relative to the Interpolations.jl v0.16.3 checkout):
//   inner_system_diags / prefiltering_system : src/b-splines/cubic.jl:101-264, quadratic.jl:71-189
//   lut!                                     : src/b-splines/b-splines.jl:257
//   Woodbury(A,U,C,V).Cp = inv(inv(C)+V*(A\U)): WoodburyMatrices (third party), used at src/filter1d.jl:77-85
// In a Julia deployment these numbers come from the reference's own functions and enter through
// b200itp_prefilter_axis; this restatement serves hosts without Julia (b200itp_prefilter).
#pragma once

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <limits>
#include <utility>
#include <vector>

#include "../../include/b200itp.h"

namespace b200itp {

template <typename T>
struct HostSystem {
  int64_t n = 0;
  std::vector<T> dl, d, du;  // LU factors (dl: multipliers, d: pivots, du: untouched super-diagonal)
  int k = 0;
  int urow[B200ITP_MAX_WOODBURY], vcol[B200ITP_MAX_WOODBURY];
  T cval[B200ITP_MAX_WOODBURY];
  T Cp[B200ITP_MAX_WOODBURY * B200ITP_MAX_WOODBURY];  // column-major k x k
};

// Edge rows of each boundary condition: diagonal entry, first off-diagonal entry, and the
// off-tridiagonal entries (column counted from the edge, value) that go through Woodbury.
struct EdgeRow {
  double diag, off;
  int nextra;
  int col[3];
  double val[3];
};

inline bool edge_row(int degree, int bc, int grid, EdgeRow &e) {
  auto set = [&](double dg, double of, std::initializer_list<std::pair<int, double>> ex) {
    e.diag = dg;
    e.off = of;
    e.nextra = 0;
    for (auto &p : ex) {
      e.col[e.nextra] = p.first;
      e.val[e.nextra] = p.second;
      ++e.nextra;
    }
    return true;
  };
  if (degree == B200ITP_CUBIC) {
    switch (bc) {
      case B200ITP_BC_FLAT:  // cubic.jl:114-124 (OnGrid), :140-157 (OnCell)
        return grid == B200ITP_ONGRID ? set(-1, 0, {{3, 1}}) : set(-9, 11, {{3, -3}, {4, 1}});
      case B200ITP_BC_LINE:  // cubic.jl:194-208 (OnGrid), :169-186 (OnCell)
        return grid == B200ITP_ONGRID ? set(1, -2, {{3, 1}}) : set(3, -7, {{3, 5}, {4, -1}});
      case B200ITP_BC_FREE:  // cubic.jl:242-264
        return set(-1, 4, {{3, -6}, {4, 4}, {5, -1}});
      default:
        return false;  // Reflect/Throw: the reference has no method (prefiltering.jl:124)
    }
  }
  if (degree == B200ITP_QUADRATIC) {
    switch (bc) {
      case B200ITP_BC_FLAT:
      case B200ITP_BC_REFLECT:  // quadratic.jl:84-89 (OnCell, bare LU), :118-129 (OnGrid)
        return grid == B200ITP_ONCELL ? set(-1, 1, {}) : set(-1, 0, {{3, 1}});
      case B200ITP_BC_LINE:  // quadratic.jl:138-149
        return set(1, -2, {{3, 1}});
      case B200ITP_BC_FREE:  // quadratic.jl:158-170
        return set(1, -3, {{3, 3}, {4, -1}});
      case B200ITP_BC_INPLACE:  // quadratic.jl:91-95 (bare LU)
        return grid == B200ITP_ONCELL ? set(0.875, 0.125, {}) : false;
      case B200ITP_BC_INPLACEQ:  // quadratic.jl:98-110
        return grid == B200ITP_ONCELL ? set(1.125, -0.25, {{3, 0.125}}) : false;
      default:
        return false;
    }
  }
  return false;
}

// returns 0, B200ITP_E_UNSUPPORTED (no method in the reference) or B200ITP_E_SIZE
template <typename T>
int assemble_system(int64_t n, int degree, int bc, int grid, HostSystem<T> &S) {
  if (degree != B200ITP_CUBIC && degree != B200ITP_QUADRATIC) return B200ITP_E_UNSUPPORTED;
  if (n < 2) return B200ITP_E_SIZE;
  // convert(T, SimpleRatio(p, q)) == T(p)/T(q): rounded once in the working type
  const T off = degree == B200ITP_CUBIC ? T(1) / T(6) : T(1) / T(8);
  const T dia = degree == B200ITP_CUBIC ? T(2) / T(3) : T(3) / T(4);
  S.n = n;
  S.dl.assign(n - 1, off);
  S.du.assign(n - 1, off);
  S.d.assign(n, dia);
  S.k = 0;
  auto wb = [&](int64_t i, int64_t j, T v) {
    S.urow[S.k] = (int)i;
    S.vcol[S.k] = (int)j;
    S.cval[S.k] = v;
    ++S.k;
  };
  if (bc == B200ITP_BC_PERIODIC) {  // cubic.jl:218-228, quadratic.jl:180-189
    wb(1, n, S.du[0]);
    wb(n, 1, S.dl[n - 2]);
  } else {
    EdgeRow e;
    if (!edge_row(degree, bc, grid, e)) return B200ITP_E_UNSUPPORTED;
    S.d[0] = S.d[n - 1] = (T)e.diag;
    S.du[0] = S.dl[n - 2] = (T)e.off;
    // the reference lists (1,c,v),(n,n+1-c,v) pairs in the order of its sparse_factors calls;
    // only the pairing matters for the solution
    if (degree == B200ITP_CUBIC && bc == B200ITP_BC_FREE) {
      for (int q = 0; q < e.nextra; ++q) wb(1, e.col[q], (T)e.val[q]);
      for (int q = 0; q < e.nextra; ++q) wb(n, n + 1 - e.col[q], (T)e.val[q]);
    } else {
      for (int q = 0; q < e.nextra; ++q) {
        wb(1, e.col[q], (T)e.val[q]);
        wb(n, n + 1 - e.col[q], (T)e.val[q]);
      }
    }
    for (int j = 0; j < S.k; ++j)
      if (S.vcol[j] < 1 || S.vcol[j] > n) return B200ITP_E_SIZE;
  }
  // lu!(Tridiagonal(dl, d, du), NoPivot())
  for (int64_t i = 0; i + 1 < n; ++i) {
    if (S.d[i] != T(0)) {
      const T f = S.dl[i] / S.d[i];
      S.dl[i] = f;
      const T t = f * S.du[i];
      S.d[i + 1] = S.d[i + 1] - t;
    }
  }
  return 0;
}

// x <- A \ x for the LU-factored tridiagonal, long-double accumulation free version in T
template <typename T, typename V>
void lu_solve_inplace(const std::vector<T> &dl, const std::vector<T> &d, const std::vector<T> &du, std::vector<V> &x) {
  const int64_t n = (int64_t)d.size();
  for (int64_t i = 1; i < n; ++i) x[i] = x[i] - (V)dl[i - 1] * x[i - 1];
  x[n - 1] = x[n - 1] / (V)d[n - 1];
  for (int64_t i = n - 2; i >= 0; --i) x[i] = (x[i] - (V)du[i] * x[i + 1]) / (V)d[i];
}

// k x k inverse (Gauss-Jordan, partial pivoting); returns false if singular
template <typename V>
bool small_inverse(int k, V *M /*column-major*/) {
  V a[B200ITP_MAX_WOODBURY][2 * B200ITP_MAX_WOODBURY];
  for (int i = 0; i < k; ++i)
    for (int j = 0; j < k; ++j) {
      a[i][j] = M[i + j * k];
      a[i][k + j] = i == j ? V(1) : V(0);
    }
  for (int c = 0; c < k; ++c) {
    int p = c;
    for (int r = c + 1; r < k; ++r)
      if (std::fabs((double)a[r][c]) > std::fabs((double)a[p][c])) p = r;
    if (a[p][c] == V(0)) return false;
    if (p != c)
      for (int j = 0; j < 2 * k; ++j) std::swap(a[p][j], a[c][j]);
    const V piv = a[c][c];
    for (int j = 0; j < 2 * k; ++j) a[c][j] /= piv;
    for (int r = 0; r < k; ++r) {
      if (r == c) continue;
      const V f = a[r][c];
      if (f == V(0)) continue;
      for (int j = 0; j < 2 * k; ++j) a[r][j] -= f * a[c][j];
    }
  }
  for (int i = 0; i < k; ++i)
    for (int j = 0; j < k; ++j) M[i + j * k] = a[i][k + j];
  return true;
}

// W.Cp = inv(inv(C) + V*(A\U))
template <typename T>
int woodbury_cp(HostSystem<T> &S) {
  const int k = S.k;
  if (k == 0) return 0;
  std::vector<T> z(S.n);
  for (int j = 0; j < k; ++j) {
    std::fill(z.begin(), z.end(), T(0));
    z[S.urow[j] - 1] = T(1);
    lu_solve_inplace(S.dl, S.d, S.du, z);
    for (int i = 0; i < k; ++i) S.Cp[i + j * k] = z[S.vcol[i] - 1];
  }
  for (int j = 0; j < k; ++j) S.Cp[j + j * k] += T(1) / S.cval[j];
  return small_inverse<T>(k, S.Cp) ? 0 : B200ITP_E_BADARG;
}

// ---- the compact plan passed to the kernels by value ------------------------------------------
constexpr int kHead = 32;  // rows with tabulated LU coefficients; they are constant afterwards
constexpr int kZ = 48;     // extent of the decaying Woodbury correction columns

template <typename T>
struct AxisPlan {
  int n;          // line length
  int P;          // head rows
  int nv;         // number of V picks (Woodbury rank), 0: bare LU
  int Hz;         // rows touched by the correction at each end
  T Lh[kHead];    // forward multiplier of row r (1-based r = index+1); Lh[0] unused
  T Uh[kHead];    // super-diagonal entry of row r
  T Dh[kHead];    // reciprocal pivot of row r
  T Linf, Uinf, Dinf;
  T Llast, Dlast;  // row n
  int vrow[B200ITP_MAX_WOODBURY];
  T W1[B200ITP_MAX_WOODBURY], Wn[B200ITP_MAX_WOODBURY];
  T z1[kZ];  // (A\e_1)[r], r = 1..Hz
  T zn[kZ];  // (A\e_n)[n-Hz+1 .. n]
};

// Full-length tables for the general (any n, any system) kernel.
template <typename T>
struct FullTables {
  std::vector<T> L, U, Dinv, z1, zn;
};

// Builds plan + full tables from LU factors and Woodbury data given in type T.
// `fast_ok` tells whether the streaming kernels' assumptions hold (coefficients constant after the
// head, corrections decayed within kZ rows, V picks near the ends).
template <typename T>
int build_plan(int64_t n, const T *dl, const T *d, const T *du, int k, const int32_t *urow, const int32_t *vcol,
               const T *Cp, AxisPlan<T> &p, FullTables<T> &F, bool &fast_ok) {
  if (n < 2 || n > 0x7fffffff) return B200ITP_E_SIZE;
  if (k < 0 || k > B200ITP_MAX_WOODBURY) return B200ITP_E_BADARG;
  F.L.assign(n, T(0));
  F.U.assign(n, T(0));
  F.Dinv.resize(n);
  for (int64_t r = 0; r < n; ++r) {
    if (d[r] == T(0)) return B200ITP_E_BADARG;
    F.Dinv[r] = T(1) / d[r];  // filter1d.jl:31,51 `dinv = 1 ./ d`
    if (r > 0) F.L[r] = dl[r - 1];
    if (r + 1 < n) F.U[r] = du[r];
  }
  // correction columns z1 = A\e_1 and zn = A\e_n, weights W1/Wn so that
  //   x_final = x - z1*(W1 . x[vrow]) - zn*(Wn . x[vrow])      (filter1d.jl:79-84 without the temporaries)
  F.z1.assign(n, T(0));
  F.zn.assign(n, T(0));
  p.nv = k;
  for (int j = 0; j < B200ITP_MAX_WOODBURY; ++j) {
    p.vrow[j] = 0;
    p.W1[j] = p.Wn[j] = T(0);
  }
  if (k > 0) {
    std::vector<T> dlv(dl, dl + (n - 1)), dv(d, d + n), duv(du, du + (n - 1));
    F.z1[0] = T(1);
    F.zn[n - 1] = T(1);
    lu_solve_inplace(dlv, dv, duv, F.z1);
    lu_solve_inplace(dlv, dv, duv, F.zn);
    for (int i = 0; i < k; ++i) {
      if (urow[i] != 1 && urow[i] != n) return B200ITP_E_UNSUPPORTED;
      if (vcol[i] < 1 || vcol[i] > n) return B200ITP_E_BADARG;
    }
    for (int j = 0; j < k; ++j) {
      p.vrow[j] = vcol[j];
      T w1 = 0, wn = 0;
      for (int i = 0; i < k; ++i) {  // tmp3[urow[i]] += (Cp*tmp1)[i]
        if (urow[i] == 1) w1 += Cp[i + j * k];
        if (urow[i] == n) wn += Cp[i + j * k];
      }
      p.W1[j] = w1;
      p.Wn[j] = wn;
    }
  }
  // compact plan
  p.n = (int)n;
  p.P = (int)std::min<int64_t>(n, kHead);
  for (int r = 0; r < kHead; ++r) {
    const bool in = r < n;
    p.Lh[r] = in ? F.L[r] : T(0);
    p.Uh[r] = in ? F.U[r] : T(0);
    p.Dh[r] = in ? F.Dinv[r] : T(1);
  }
  const int64_t rs = std::min<int64_t>(n - 2 >= 0 ? n - 2 : 0, kHead);  // a representative steady row (0-based)
  p.Linf = F.L[rs];
  p.Uinf = F.U[rs];
  p.Dinf = F.Dinv[rs];
  p.Llast = F.L[n - 1];
  p.Dlast = F.Dinv[n - 1];
  fast_ok = n >= 128;
  if (fast_ok) {
    const T tol = T(8) * std::numeric_limits<T>::epsilon();
    for (int64_t r = kHead; r + 1 < n; ++r) {
      if (std::fabs(F.L[r] - p.Linf) > tol * std::fabs(p.Linf) || std::fabs(F.Dinv[r] - p.Dinf) > tol * std::fabs(p.Dinf) ||
          F.U[r] != p.Uinf) {
        fast_ok = false;
        break;
      }
    }
    // both recurrences must contract fast enough for the block look-ahead to be exact to round-off
    const double cf = std::fabs((double)p.Linf), cb = std::fabs((double)p.Uinf * (double)p.Dinf);
    if (!(cf < 0.3 && cb < 0.3)) fast_ok = false;
  }
  p.Hz = 0;
  for (int r = 0; r < kZ; ++r) p.z1[r] = p.zn[r] = T(0);
  if (k > 0 && fast_ok) {
    double m1 = 0, mn = 0;
    for (int64_t r = 0; r < n; ++r) {
      m1 = std::max(m1, std::fabs((double)F.z1[r]));
      mn = std::max(mn, std::fabs((double)F.zn[r]));
    }
    const double thr = 1e-3 * (double)std::numeric_limits<T>::epsilon();  // below round-off of the result
    int64_t h1 = 0, hn = 0;
    for (int64_t r = 0; r < n; ++r)
      if (std::fabs((double)F.z1[r]) > thr * m1) h1 = r + 1;
    for (int64_t r = n - 1; r >= 0; --r)
      if (std::fabs((double)F.zn[r]) > thr * mn) hn = n - r;
    const int64_t hz = std::max(h1, hn);
    if (hz > kZ) {
      fast_ok = false;
    } else {
      p.Hz = (int)hz;
      for (int r = 0; r < p.Hz; ++r) {
        p.z1[r] = F.z1[r];
        p.zn[r] = F.zn[n - p.Hz + r];
      }
      for (int j = 0; j < k; ++j)
        if (!(p.vrow[j] <= 8 || p.vrow[j] > n - 8)) fast_ok = false;
    }
  }
  return 0;
}

}  // namespace b200itp
