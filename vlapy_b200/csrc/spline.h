// spline.h -- semi-Lagrangian advection operators (vlapy/core/vlasov.py:42-80 get_vdfdx_sl, :168-210 get_edfdv_sl).
//
// The reference pads f with ONE periodic ghost line on either side of the advected axis, fits a bicubic
// scipy.interpolate.RectBivariateSpline (FITPACK, s = 0) and evaluates it at the feet of the characteristics,
// (x - v dt, v) resp. (x, v - e dt).  The evaluation points lie on grid lines of the OTHER axis, where a
// tensor-product interpolating spline reduces to the 1-D interpolating spline of that line (oracle:
// nak_spline_shift); FITPACK's knot choice for s = 0 makes it the not-a-knot cubic spline, and points outside the
// padded axis are clamped to its ends (bispeu -> fpbisp).  So per line of n + 2 padded values y on a uniform grid:
//     r_k = 6 (y_{k+1} - 2 y_k + y_{k-1}) / h^2                                   k = 1 .. n
//     M_1 = r_1 / 6,  M_n = r_n / 6                                               (not-a-knot on a uniform grid)
//     M_{k-1} + 4 M_k + M_{k+1} = r_k                                             k = 2 .. n-1   (second derivatives)
//     M_0 = 2 M_1 - M_2,  M_{n+1} = 2 M_n - M_{n-1}
//     S(q) = (1-s) y_c + s y_{c+1} + h^2/6 (((1-s)^3 - (1-s)) M_c + (s^3 - s) M_{c+1}),   q in cell c, s = (q - x_c)/h
// Phase programs (vpfp_common.h); tests/emul runs the same source on the host.
//   columns (v df/dx): SplineColSweepProg -- one thread per v column marches the constant (1,4,1) system down and up
//                      the x axis (coalesced across columns; the elimination factors are the same for every column and
//                      come from a table), then SplineEvalProg<COLS>;
//   rows    (e df/dv): SplineRowRhsProg writes M_1, M_n and the right-hand sides, the (1,4,1) systems of all rows are
//                      solved by tridiag.h TridiagProg with broadcast diagonals, then SplineEvalProg<ROWS>.
#pragma once
#include "vpfp_common.h"

// padded line value k (0 .. n+1) of a periodic axis of n cells: ghost = the cell at the other end
VPFP_HD long spl_wrap(long k, long n) { return k == 0 ? n - 1 : (k == n + 1 ? 0 : k - 1); }

struct SplineColSweepProg {
  const double* f; long ld;      // (nx, nv)
  double* M; long ldm;           // (nx + 2, nv): second derivatives of every column's padded spline
  const double* cp;              // elimination factors c'_k of the (1,4,1) system, nx - 2 entries
  double h;                      // grid spacing of the advected (x) axis
  int nx, nv;

  VPFP_HD void phase(int, long blk, int tid, int nthr, unsigned char*) const {
    const long j = blk * nthr + tid;
    if (j >= nv) return;
    const long n = nx, n2 = n + 2;
    const double s6 = 6.0 / (h * h);
    auto Y = [&](long k) { return f[spl_wrap(k, n) * ld + j]; };
    auto R = [&](long k) { return s6 * (Y(k + 1) - 2.0 * Y(k) + Y(k - 1)); };
    const double M1 = R(1) / 6.0, Mn = R(n) / 6.0;
    M[1 * ldm + j] = M1;
    M[n * ldm + j] = Mn;
    const long m = n2 - 4;                       // unknowns M_2 .. M_{n-1}
    if (m > 0) {
      double ym = Y(1), y0 = Y(2), yp, dp = 0.0;
      for (long k = 0; k < m; ++k) {             // unknown k is M_{k+2}
        yp = Y(k + 3);
        double rhs = s6 * (yp - 2.0 * y0 + ym);
        if (k == 0) rhs -= M1;
        if (k == m - 1) rhs -= Mn;
        dp = (rhs - dp) * cp[k];                 // c'_k = 1 / (4 - c'_{k-1}), d'_k = (rhs_k - d'_{k-1}) c'_k
        M[(k + 2) * ldm + j] = dp;
        ym = y0; y0 = yp;
      }
      double mk = dp;                            // M_{n-1}
      for (long k = m - 2; k >= 0; --k) {
        mk = M[(k + 2) * ldm + j] - cp[k] * mk;
        M[(k + 2) * ldm + j] = mk;
      }
    }
    M[0 * ldm + j] = 2.0 * M1 - M[2 * ldm + j];
    M[(n + 1) * ldm + j] = 2.0 * Mn - M[(n - 1) * ldm + j];
  }
};

struct SplineRowRhsProg {
  const double* f; long ld;      // (nx, nv)
  double* M; long ldm;           // (nx, nv + 2): M_1, M_n in place, right-hand sides of M_2 .. M_{n-1} in slots 2 .. n-1
  double h;
  int nx, nv, cblocks;

  VPFP_HD void phase(int, long blk, int tid, int nthr, unsigned char*) const {
    const long row = blk / cblocks;
    const long k = (blk % cblocks) * nthr + tid + 1;          // 1 .. n
    const long n = nv;
    if (k > n) return;
    const double* y = f + row * ld;
    const double s6 = 6.0 / (h * h);
    const double r = s6 * (y[spl_wrap(k + 1, n)] - 2.0 * y[spl_wrap(k, n)] + y[spl_wrap(k - 1, n)]);
    double* Mr = M + row * ldm;
    if (k == 1 || k == n) { Mr[k] = r / 6.0; return; }
    double rhs = r;
    if (k == 2) rhs -= s6 * (y[spl_wrap(2, n)] - 2.0 * y[spl_wrap(1, n)] + y[spl_wrap(0, n)]) / 6.0;              // - M_1
    if (k == n - 1) rhs -= s6 * (y[spl_wrap(n + 1, n)] - 2.0 * y[spl_wrap(n, n)] + y[spl_wrap(n - 1, n)]) / 6.0;  // - M_n
    Mr[k] = rhs;
  }
};

// evaluation at the feet of the characteristics.  COLS: q = x_i - v_j dt along x (line = column j);
// ROWS: q = v_j - e_i dt along v (line = row i).
template <int COLS>
struct SplineEvalProg {
  const double* f; long ld;
  const double* M; long ldm;     // COLS: (nx + 2, nv); ROWS: (nx, nv + 2) with slots 0 and n + 1 not yet filled
  double* out; long ld_out;
  const double* ax;              // the advected axis (x for COLS, v for ROWS)
  const double* c;               // COLS: v[nv]; ROWS: e[nx]
  double dt;
  int nx, nv, cblocks;

  VPFP_HD void phase(int, long blk, int tid, int nthr, unsigned char*) const {
    const long i = blk / cblocks;
    const long j = (blk % cblocks) * nthr + tid;
    if (j >= nv) return;
    const long n = COLS ? nx : nv;
    const double h = ax[2] - ax[1];
    const double a0 = ax[0] - h, a1 = ax[n - 1] + h;                      // ends of the padded axis (vlasov.py:35-37)
    double q = COLS ? ax[i] - c[j] * dt : ax[j] - c[i] * dt;
    q = q < a0 ? a0 : (q > a1 ? a1 : q);                                  // FITPACK clamps (fpbisp)
    long cell = (long)floor((q - a0) / h);
    cell = cell < 0 ? 0 : (cell > n ? n : cell);
    const double s = (q - (a0 + (double)cell * h)) / h, u = 1.0 - s;
    double yc, yn, Mc, Mn;
    if (COLS) {
      yc = f[spl_wrap(cell, n) * ld + j];
      yn = f[spl_wrap(cell + 1, n) * ld + j];
      Mc = M[cell * ldm + j];
      Mn = M[(cell + 1) * ldm + j];
    } else {
      const double* y = f + i * ld;
      const double* Mr = M + i * ldm;
      yc = y[spl_wrap(cell, n)];
      yn = y[spl_wrap(cell + 1, n)];
      Mc = (cell == 0) ? 2.0 * Mr[1] - Mr[2] : Mr[cell];
      Mn = (cell + 1 == n + 1) ? 2.0 * Mr[n] - Mr[n - 1] : Mr[cell + 1];
    }
    out[i * ld_out + j] = u * yc + s * yn + (h * h / 6.0) * ((u * u * u - u) * Mc + (s * s * s - s) * Mn);
  }
};
