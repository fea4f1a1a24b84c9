// tridiag.h -- the reference's TWO-STAGE collision interface on the device (phase programs, vpfp_common.h):
//   DiagProg     f -> (a, b, c): the three diagonals of the Lenard-Bernstein / Dougherty operator for every x
//                (vlapy/core/collisions.py:44-81, 104-158 via get_batched_array_maker :292-317)
//   TridiagProg  (a, b, c, d) -> x: one tridiagonal system per row with GENERAL diagonals, no pivoting
//                (vlapy/core/collisions.py:222-265, get_batched_tridiag_solver / _batched_tridiag_solver_)
// The product step never materialises the diagonals (fp_reg.cuh / fp_fast.cuh / rowops.h FpProg build them
// from two scalars per row); these two kernels serve callers of the reference's explicit interface
// (get_matrix_solver(nx, nv, name)(a, b, c, f), tests/test_collisions.py of the reference).
//
// TridiagProg is the partition method of rowops.h FpProg with the diagonals read from memory: a row is cut into
// P chunks; the interior of a chunk (all cells but its last, the "separator") is eliminated exactly by an LU sweep
// down and a UL sweep up, which leaves six spike end values per chunk; the P separators form a tridiagonal system
// solved by parallel cyclic reduction; the interiors are back-substituted with their two neighbours known.
// Layout: a, c are (rows, nv - 1) with pitch lda (a[i-1] couples row i to x[i-1], c[i] couples row i to x[i+1], as
// the reference stores them), b and d are (rows, nv).
#pragma once
#include "rowops.h"

#ifndef TRIDIAG_MAXM
#define TRIDIAG_MAXM 64
#endif

struct TridiagProg {
  const double* a; long lda;
  const double* b; long ldb;
  const double* c; long ldc;
  const double* d; long ldd;
  double* x; long ldx;
  int rows, nv;
  int P, m;   // chunks and nominal chunk length (the last chunk takes the remainder, m <= len < 2m)

  VPFP_HD int npcr() const { return ilog2(P); }
  VPFP_HD int nphases() const { return 1 + 1 + 1 + npcr() + 1 + 1 + 1; }
  VPFP_HD long smem_bytes(int nthr) const {
    const long S = (P > nthr ? P : nthr);
    return ((long)nv + 10L * S) * (long)sizeof(double);
  }
  VPFP_HD int cstart(int j) const { return j * m; }
  VPFP_HD int cend(int j) const { return (j == P - 1) ? nv - 1 : j * m + m - 1; }

  VPFP_HD void phase(int ph, long blk, int tid, int nthr, unsigned char* smem) const {
    double* row = reinterpret_cast<double*>(smem);   // right-hand side, then the solution
    double* W = row + nv;
    const long S = (P > nthr ? P : nthr);
    const double* A = a + blk * lda - 1;   // A[i] = coefficient of x[i-1] in row i (i >= 1)
    const double* B = b + blk * ldb;
    const double* C = c + blk * ldc;       // C[i] = coefficient of x[i+1] in row i (i <= nv - 2)
    int p = ph;
    if (p == 0) {
      const double* src = d + blk * ldd;
      for (int j = tid; j < nv; j += nthr) row[j] = src[j];
      return;
    }
    p -= 1;
    // ---- chunk interiors -> six spike end values per chunk
    if (p == 0) {
      double* chq = W + 4L * S;  // [6][S]
      for (int j = tid; j < P; j += nthr) {
        const int s = cstart(j), e = cend(j);
        // LU sweep down the interior s..e-1: z = eliminated right-hand side, g = eliminated unit spike of x[s-1]
        double rp = 1.0 / ldg(B + s), z = row[s], g = 1.0;
        for (int i = s + 1; i <= e - 1; ++i) {
          const double l = ldg(A + i) * rp;
          rp = 1.0 / (ldg(B + i) - l * ldg(C + i - 1));
          z = row[i] - l * z;
          g = -l * g;
        }
        const double y_last = z * rp, w_last = rp, u_last = g * rp;
        // UL sweep up the interior e-1..s
        double rq = 1.0 / ldg(B + e - 1), t = row[e - 1], h = 1.0;
        for (int i = e - 2; i >= s; --i) {
          const double r = ldg(C + i) * rq;
          rq = 1.0 / (ldg(B + i) - r * ldg(A + i + 1));
          t = row[i] - r * t;
          h = -r * h;
        }
        chq[0 * S + j] = rq;       // u_first: x[s] per unit of (-A[s] x[s-1])
        chq[1 * S + j] = u_last;
        chq[2 * S + j] = h * rq;   // w_first: x[s] per unit of (-C[e-1] x[e])
        chq[3 * S + j] = w_last;
        chq[4 * S + j] = t * rq;   // y_first
        chq[5 * S + j] = y_last;
      }
      return;
    }
    p -= 1;
    // ---- separator system
    if (p == 0) {
      const double* chq = W + 4L * S;
      double* R = W;  // [4][S]: ra, rb, rc, rd
      for (int j = tid; j < P; j += nthr) {
        const int s = cstart(j), e = cend(j);
        const double As = (s > 0) ? ldg(A + s) : 0.0;
        const double Ce1 = ldg(C + e - 1);
        const double Ae = ldg(A + e);
        double ra = -Ae * As * chq[1 * S + j];
        double rb = ldg(B + e) - Ae * Ce1 * chq[3 * S + j];
        double rc = 0.0;
        double rd = row[e] - Ae * chq[5 * S + j];
        if (j + 1 < P) {
          const int s2 = cstart(j + 1), e2 = cend(j + 1);
          const double Ce = ldg(C + e);
          rb -= Ce * ldg(A + s2) * chq[0 * S + j + 1];
          rc = -Ce * ldg(C + e2 - 1) * chq[2 * S + j + 1];
          rd -= Ce * chq[4 * S + j + 1];
        }
        R[0 * S + j] = ra; R[1 * S + j] = rb; R[2 * S + j] = rc; R[3 * S + j] = rd;
      }
      return;
    }
    p -= 1;
    // ---- parallel cyclic reduction, ping-pong between W[0,4S) and W[4S,8S)
    if (p < npcr()) {
      const int st = 1 << p;
      const double* src = W + ((p & 1) ? 4L * S : 0);
      double* dst = W + ((p & 1) ? 0 : 4L * S);
      for (int j = tid; j < P; j += nthr) {
        double ra = src[j], rb = src[S + j], rc = src[2 * S + j], rd = src[3 * S + j];
        double na = 0.0, nc = 0.0;
        const int im = j - st, ip = j + st;
        if (im >= 0) {
          const double al = -ra / src[S + im];
          na = al * src[im];
          rb += al * src[2 * S + im];
          rd += al * src[3 * S + im];
        }
        if (ip < P) {
          const double ga = -rc / src[S + ip];
          nc = ga * src[2 * S + ip];
          rb += ga * src[ip];
          rd += ga * src[3 * S + ip];
        }
        dst[j] = na; dst[S + j] = rb; dst[2 * S + j] = nc; dst[3 * S + j] = rd;
      }
      return;
    }
    p -= npcr();
    if (p == 0) {
      const double* src = W + ((npcr() & 1) ? 4L * S : 0);
      double* xr = W + 8L * S;
      for (int j = tid; j < P; j += nthr) xr[j] = src[3 * S + j] / src[S + j];
      return;
    }
    p -= 1;
    // ---- interiors with known neighbours (Thomas on s..e-1), in place in `row`
    if (p == 0) {
      const double* xr = W + 8L * S;
      for (int j = tid; j < P; j += nthr) {
        const int s = cstart(j), e = cend(j);
        const double xl = (j > 0) ? xr[j - 1] : 0.0;
        const double xe = xr[j];
        double rpv[2 * TRIDIAG_MAXM];
        double d0 = row[s] - ((s > 0) ? ldg(A + s) * xl : 0.0);
        if (e - 1 == s) d0 -= ldg(C + s) * xe;
        double rp = 1.0 / ldg(B + s), z = d0;
        rpv[0] = rp;
        row[s] = z;
        for (int i = s + 1; i <= e - 1; ++i) {
          const double l = ldg(A + i) * rp;
          rp = 1.0 / (ldg(B + i) - l * ldg(C + i - 1));
          double di = row[i];
          if (i == e - 1) di -= ldg(C + i) * xe;
          z = di - l * z;
          rpv[i - s] = rp;
          row[i] = z;
        }
        double xv = row[e - 1] * rpv[e - 1 - s];
        row[e - 1] = xv;
        for (int i = e - 2; i >= s; --i) {
          xv = (row[i] - ldg(C + i) * xv) * rpv[i - s];
          row[i] = xv;
        }
        row[e] = xe;
      }
      return;
    }
    p -= 1;
    {
      double* dst = x + blk * ldx;
      for (int j = tid; j < nv; j += nthr) dst[j] = row[j];
    }
  }
};

// f -> (a, b, c).  One CTA per row: the row moment(s) by a deterministic tree, then the diagonals elementwise with the
// reference's own association:  a = nu*dt*(-T/dv**2 + (v[:-1]-vbar)/2/dv),  b = 1 + nu*dt*(2*T/dv**2),
// c = nu*dt*(-T/dv**2 - (v[1:]-vbar)/2/dv)   (LB: vbar = 0, T = int f v^2; DG: vbar = int f v, T = int f (v-vbar)^2).
struct DiagProg {
  const double* f; long ld;
  const double* v;
  double nu, dt, dv;
  int op;   // 0 = lb, 1 = dg
  double* a; long lda;
  double* b; long ldb;
  double* c; long ldc;
  int rows, nv;

  VPFP_HD int nphases(int nthr) const { return 2 * (1 + tree_phases(nthr)) + 1; }
  VPFP_HD long smem_bytes(int nthr) const { return (2L * nthr + 4) * (long)sizeof(double); }

  VPFP_HD void phase(int ph, long blk, int tid, int nthr, unsigned char* smem) const {
    double* W = reinterpret_cast<double*>(smem);
    double* scal = W + 2L * nthr;
    const double* src = f + blk * ld;
    const int nt = tree_phases(nthr);
    int p = ph;
    if (p == 0) {
      double acc = 0.0;
      for (int j = tid; j < nv; j += nthr) {
        const double w = trapz_w(j, nv, dv, 3);
        acc += (op == 0) ? w * src[j] * v[j] * v[j] : w * src[j] * v[j];
      }
      W[tid] = acc;
      return;
    }
    p -= 1;
    if (p < nt) { tree_step(p, nthr, 1, tid, nthr, W); return; }
    p -= nt;
    if (p == 0) {
      const double first = W[0];
      double acc = 0.0;
      if (op == 1)
        for (int j = tid; j < nv; j += nthr) {
          const double dd = v[j] - first;
          acc += trapz_w(j, nv, dv, 3) * src[j] * dd * dd;
        }
      W[nthr + tid] = acc;
      if (tid == 0) scal[0] = first;
      return;
    }
    p -= 1;
    if (p < nt) { tree_step(p, nthr, 1, tid, nthr, W + nthr); return; }
    {
      const double T = (op == 0) ? scal[0] : W[nthr];
      const double vbar = (op == 0) ? 0.0 : scal[0];
      const double nudt = nu * dt;
      const double tdv = -T / (dv * dv);
      const double bd = 1.0 + nudt * (2.0 * T / (dv * dv));
      double* pa = a + blk * lda;
      double* pb = b + blk * ldb;
      double* pc = c + blk * ldc;
      for (int j = tid; j < nv; j += nthr) {
        pb[j] = bd;
        if (j < nv - 1) {
          pa[j] = nudt * (tdv + (v[j] - vbar) / 2.0 / dv);
          pc[j] = nudt * (tdv - (v[j + 1] - vbar) / 2.0 / dv);
        }
      }
    }
  }
};
