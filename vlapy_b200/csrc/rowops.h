// rowops.h -- per-x-row programs: v-moments, the implicit Fokker-Planck solve, the cd2 stencil,
// low x-Fourier modes, the driver and the series means.  Same phase/barrier structure as
// advect.h (see vpfp_common.h); one CTA owns one row unless stated otherwise.
#pragma once
#include "vpfp_common.h"

// trapezoid weight of local column j (np.trapz(..., dx=dv): interior dv, global end cells dv/2)
VPFP_HD double trapz_w(int j, int ncols, double dv, int edge_flags) {
  if ((j == 0 && (edge_flags & 1)) || (j == ncols - 1 && (edge_flags & 2))) return 0.5 * dv;
  return dv;
}

// Deterministic tree reduction of nval interleaved arrays part[k*P + i], i < P, one halving per
// phase; after ceil(log2 P) phases part[k*P] holds the sums.
VPFP_HD int tree_phases(int P) { return ilog2(P); }
VPFP_HD void tree_step(int step, int P, int nval, int tid, int nthr, double* part) {
  // halving with ceil so that any P works: active width w -> (w+1)/2
  int w = P;
  for (int s = 0; s < step; ++s) w = (w + 1) / 2;
  int h = (w + 1) / 2;
  for (int i = tid; i < h * nval; i += nthr) {
    int k = i / h, j = i - k * h;
    if (j + h < w) part[(long)k * P + j] += part[(long)k * P + j + h];
  }
}

// ---------------------------------------------------------------------------------------------
// v-moments of a row: vlapy/core/step.py:164-171 (p = 0..5), :216-224 (f^2, f ln f),
// vlapy/core/field.py:27-36 (p = 0).
// ---------------------------------------------------------------------------------------------
struct MomentsProg {
  const double* f;
  long ld;
  const double* v;
  double dv;
  double* out;
  long out_ld;
  int nmom, rows, ncols, edge_flags;

  VPFP_HD int nphases(int nthr) const { return 2 + tree_phases(nthr); }
  VPFP_HD long smem_bytes(int nthr) const { return (long)8 * nthr * sizeof(double); }

  VPFP_HD static void accumulate(double acc[8], double w, double fv, double vv, int nmom) {
    double t = w * fv;
    acc[0] += t;
    if (nmom > 1) {
      double p = t * vv;
      acc[1] += p;
      p *= vv; acc[2] += p;
      p *= vv; acc[3] += p;
      p *= vv; acc[4] += p;
      p *= vv; acc[5] += p;
    }
    if (nmom > 6) {
      acc[6] += t * fv;
      acc[7] += t * log(fv);
    }
  }

  VPFP_HD void phase(int ph, long blk, int tid, int nthr, unsigned char* smem) const {
    double* part = reinterpret_cast<double*>(smem);
    const int nt = tree_phases(nthr);
    if (ph == 0) {
      double acc[8];
      for (int k = 0; k < 8; ++k) acc[k] = 0.0;
      const double* row = f + blk * ld;
      for (int j = tid; j < ncols; j += nthr)
        accumulate(acc, trapz_w(j, ncols, dv, edge_flags), row[j], v[j], nmom);
      for (int k = 0; k < 8; ++k) part[(long)k * nthr + tid] = acc[k];
    } else if (ph <= nt) {
      tree_step(ph - 1, nthr, 8, tid, nthr, part);
    } else {
      if (tid < nmom) out[(long)tid * out_ld + blk] = part[(long)tid * nthr];
    }
  }
};

// ---------------------------------------------------------------------------------------------
// Implicit Fokker-Planck step for one row (vlapy/core/collisions.py:44-81, 104-158, 232-263 through
// vlapy/core/step.py:102-108).  The diagonals are functions of two row scalars (T = v0t_sq, vbar):
//   sub   A_i = nu dt (-T/dv^2 + (v[i-1] - vbar)/2/dv)     (multiplies x[i-1], i >= 1)
//   diag  Bd  = 1 + nu dt 2 T/dv^2
//   super C_i = nu dt (-T/dv^2 - (v[i+1] - vbar)/2/dv)     (multiplies x[i+1], i <= nv-2)
// and are never materialised.  The row lives in shared memory; it is cut into P chunks (one per
// thread).  Each chunk's last cell is a separator; the chunk interiors are eliminated exactly
// (one LU sweep down, one UL sweep up) which leaves a tridiagonal system in the P separators,
// solved by parallel cyclic reduction; the interiors are then solved with known neighbours.
// The matrix is strictly diagonally dominant (Bd - |A| - |C| = 1), so this reordering of the
// reference's Thomas sweep is stable and agrees with it to rounding.
// ---------------------------------------------------------------------------------------------
#define FP_MAXM 32

struct FpProg {
  const double* fin;
  long ld_in;
  double* fout;
  long ld_out;
  const double* v;
  double nu, dt, dv;
  int op;  // 0 = lb, 1 = dg
  double* mom_out;  // nullable, (8, rows)
  long mom_ld;
  int rows, nv;
  int P, m;  // chunks and nominal chunk length (last chunk takes the remainder, m <= len < 2m)

  VPFP_HD int npcr() const { return ilog2(P); }
  VPFP_HD int ntree(int nthr) const { return tree_phases(nthr); }
  // phase map
  VPFP_HD int nphases(int nthr) const {
    int nt = ntree(nthr);
    return 1 + nt + 1 + nt + 1 + 1 + npcr() + 1 + 1 + 1 + nt + 1;
  }
  VPFP_HD long smem_bytes(int nthr) const {
    long scratch = 10L * (P > nthr ? P : nthr);
    return ((long)nv + scratch + 8) * sizeof(double);
  }

  struct Coef {
    double nudt, tdv, bd, vbar, hdv;
  };
  VPFP_HD Coef coef(double T, double vbar) const {
    Coef c;
    c.nudt = nu * dt;
    c.tdv = -T / (dv * dv);
    c.bd = 1.0 + c.nudt * (2.0 * T / (dv * dv));
    c.vbar = vbar;
    c.hdv = dv;
    return c;
  }
  VPFP_HD double cA(const Coef& c, int i) const { return c.nudt * (c.tdv + (v[i - 1] - c.vbar) / 2.0 / c.hdv); }
  VPFP_HD double cC(const Coef& c, int i) const { return c.nudt * (c.tdv - (v[i + 1] - c.vbar) / 2.0 / c.hdv); }
  VPFP_HD int cstart(int j) const { return j * m; }
  VPFP_HD int cend(int j) const { return (j == P - 1) ? nv - 1 : j * m + m - 1; }

  VPFP_HD void phase(int ph, long blk, int tid, int nthr, unsigned char* smem) const {
    double* row = reinterpret_cast<double*>(smem);
    double* W = row + nv;                      // scratch, 10*S doubles, S = max(P, nthr)
    const int S = (P > nthr ? P : nthr);
    double* scal = W + 10L * S;                // [0] = first moment result, [1] = second
    const int nt = ntree(nthr);
    int p = ph;
    // ---- 0: load row + first moment partials (lb: int f v^2; dg: int f v)
    if (p == 0) {
      const double* src = fin + blk * ld_in;
      double acc = 0.0;
      for (int j = tid; j < nv; j += nthr) {
        double fv = src[j];
        row[j] = fv;
        double w = trapz_w(j, nv, dv, 3);
        acc += (op == 0) ? w * fv * v[j] * v[j] : w * fv * v[j];
      }
      W[tid] = acc;
      return;
    }
    p -= 1;
    if (p < nt) { tree_step(p, nthr, 1, tid, nthr, W); return; }
    p -= nt;
    // ---- second moment (dg: thermal spread about vbar); lb: nothing to do
    if (p == 0) {
      double first = W[0];
      double acc = 0.0;
      if (op == 1) {
        for (int j = tid; j < nv; j += nthr) {
          double d = v[j] - first;
          acc += trapz_w(j, nv, dv, 3) * row[j] * d * d;
        }
      }
      W[S + tid] = acc;
      if (tid == 0) scal[0] = first;
      return;
    }
    p -= 1;
    if (p < nt) { tree_step(p, nthr, 1, tid, nthr, W + S); return; }
    p -= nt;
    // ---- chunk interiors -> six spike end values per chunk
    if (p == 0) {
      if (tid == 0) scal[1] = W[S];
      const double T = (op == 0) ? scal[0] : W[S];
      const double vbar = (op == 0) ? 0.0 : scal[0];
      const Coef c = coef(T, vbar);
      double* chq = W + 4L * S;  // [6][S]
      for (int j = tid; j < P; j += nthr) {
        const int s = cstart(j), e = cend(j);
        // LU sweep down the interior s..e-1
        double rp = 1.0 / c.bd, z = row[s], g = 1.0;
        for (int i = s + 1; i <= e - 1; ++i) {
          double l = cA(c, i) * rp;
          rp = 1.0 / (c.bd - l * cC(c, i - 1));
          z = row[i] - l * z;
          g = -l * g;
        }
        double y_last = z * rp, w_last = rp, u_last = g * rp;
        // UL sweep up the interior e-1..s
        double rq = 1.0 / c.bd, t = row[e - 1], h = 1.0;
        for (int i = e - 2; i >= s; --i) {
          double r = cC(c, i) * rq;
          rq = 1.0 / (c.bd - r * cA(c, i + 1));
          t = row[i] - r * t;
          h = -r * h;
        }
        chq[0 * S + j] = rq;       // u_first
        chq[1 * S + j] = u_last;
        chq[2 * S + j] = h * rq;   // w_first
        chq[3 * S + j] = w_last;
        chq[4 * S + j] = t * rq;   // y_first
        chq[5 * S + j] = y_last;
      }
      return;
    }
    p -= 1;
    // ---- assemble the separator system
    if (p == 0) {
      const double T = (op == 0) ? scal[0] : scal[1];
      const double vbar = (op == 0) ? 0.0 : scal[0];
      const Coef c = coef(T, vbar);
      const double* chq = W + 4L * S;
      double* R = W;  // [4][S]: ra, rb, rc, rd
      for (int j = tid; j < P; j += nthr) {
        const int s = cstart(j), e = cend(j);
        const double As = (s > 0) ? cA(c, s) : 0.0;
        const double Ce1 = cC(c, e - 1);
        const double Ae = cA(c, e);
        double ra = -Ae * As * chq[1 * S + j];
        double rb = c.bd - Ae * Ce1 * chq[3 * S + j];
        double rc = 0.0;
        double rd = row[e] - Ae * chq[5 * S + j];
        if (j + 1 < P) {
          const int s2 = cstart(j + 1), e2 = cend(j + 1);
          const double Ce = cC(c, e);
          rb -= Ce * cA(c, s2) * chq[0 * S + j + 1];
          rc = -Ce * cC(c, e2 - 1) * chq[2 * S + j + 1];
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
        int im = j - st, ip = j + st;
        if (im >= 0) {
          double al = -ra / src[S + im];
          na = al * src[im];
          rb += al * src[2 * S + im];
          rd += al * src[3 * S + im];
        }
        if (ip < P) {
          double ga = -rc / src[S + ip];
          nc = ga * src[2 * S + ip];
          rb += ga * src[ip];
          rd += ga * src[3 * S + ip];
        }
        dst[j] = na; dst[S + j] = rb; dst[2 * S + j] = nc; dst[3 * S + j] = rd;
      }
      return;
    }
    p -= npcr();
    // ---- separator values
    if (p == 0) {
      const double* src = W + ((npcr() & 1) ? 4L * S : 0);
      double* xr = W + 8L * S;
      for (int j = tid; j < P; j += nthr) xr[j] = src[3 * S + j] / src[S + j];
      return;
    }
    p -= 1;
    // ---- interiors with known neighbours (Thomas on s..e-1), in place in `row`
    if (p == 0) {
      const double T = (op == 0) ? scal[0] : scal[1];
      const double vbar = (op == 0) ? 0.0 : scal[0];
      const Coef c = coef(T, vbar);
      const double* xr = W + 8L * S;
      for (int j = tid; j < P; j += nthr) {
        const int s = cstart(j), e = cend(j);
        const double xl = (j > 0) ? xr[j - 1] : 0.0;
        const double xe = xr[j];
        double rpv[2 * FP_MAXM];
        double d0 = row[s] - ((s > 0) ? cA(c, s) * xl : 0.0);
        if (e - 1 == s) d0 -= cC(c, s) * xe;
        double rp = 1.0 / c.bd, z = d0;
        rpv[0] = rp;
        row[s] = z;
        for (int i = s + 1; i <= e - 1; ++i) {
          double l = cA(c, i) * rp;
          rp = 1.0 / (c.bd - l * cC(c, i - 1));
          double di = row[i];
          if (i == e - 1) di -= cC(c, i) * xe;
          z = di - l * z;
          rpv[i - s] = rp;
          row[i] = z;
        }
        double x = row[e - 1] * rpv[e - 1 - s];
        row[e - 1] = x;
        for (int i = e - 2; i >= s; --i) {
          x = (row[i] - cC(c, i) * x) * rpv[i - s];
          row[i] = x;
        }
        row[e] = xe;
      }
      return;
    }
    p -= 1;
    // ---- store + moments of the new row
    if (p == 0) {
      double* dst = fout + blk * ld_out;
      double acc[8];
      for (int k = 0; k < 8; ++k) acc[k] = 0.0;
      for (int j = tid; j < nv; j += nthr) {
        double x = row[j];
        dst[j] = x;
        if (mom_out) MomentsProg::accumulate(acc, trapz_w(j, nv, dv, 3), x, v[j], 8);
      }
      if (mom_out)
        for (int k = 0; k < 8; ++k) W[(long)k * nthr + tid] = acc[k];
      return;
    }
    p -= 1;
    if (p < nt) { if (mom_out) tree_step(p, nthr, 8, tid, nthr, W); return; }
    p -= nt;
    if (mom_out && tid < 8) mom_out[(long)tid * mom_ld + blk] = W[(long)tid * nthr];
  }
};

// ---------------------------------------------------------------------------------------------
// e df/dv by centred differences (vlapy/core/vlasov.py:153-163; np.gradient edge_order=2).
// Elementwise: blk indexes a (row, column block) pair.
// ---------------------------------------------------------------------------------------------
struct Cd2Prog {
  const double* fin;
  long ld_in;
  double* fout;
  long ld_out;
  const double* e;
  double dt, dv;
  int rows, nv, cblocks;  // cblocks column blocks of nthr*4 columns

  VPFP_HD int nphases() const { return 1; }
  VPFP_HD void phase(int, long blk, int tid, int nthr, unsigned char*) const {
    long r = blk / cblocks;
    int cb = (int)(blk % cblocks);
    const double* src = fin + r * ld_in;
    double* dst = fout + r * ld_out;
    const double er = e[r];
    for (int u = 0; u < 4; ++u) {
      int j = (cb * 4 + u) * nthr + tid;
      if (j >= nv) continue;
      double g;
      if (j == 0) g = -(3.0 * src[0] - 4.0 * src[1] + src[2]) / (2.0 * dv);
      else if (j == nv - 1) g = (3.0 * src[nv - 1] - 4.0 * src[nv - 2] + src[nv - 3]) / (2.0 * dv);
      else g = (src[j + 1] - src[j - 1]) / (2.0 * dv);
      dst[j] = src[j] - er * g * dt;
    }
  }
};

// ---------------------------------------------------------------------------------------------
// Lowest x-Fourier modes of f for every v (vlapy/core/step.py:130-135): stage 1 sums a chunk of
// x rows per CTA for a block of columns, stage 2 adds the chunk partials in fixed order.
// ---------------------------------------------------------------------------------------------
struct XmodesProg {
  const double* f;
  long ld;
  double* partial;  // [batch][xchunks][nmodes][ncols][2]
  int nmodes, batch, nx, ncols, xchunks, cblocks;
  int x_offset, nx_total;  // rows are cells x_offset .. x_offset+nx-1 of a grid of nx_total (x-sharding)

  VPFP_HD int nphases() const { return 1; }
  VPFP_HD void phase(int, long blk, int tid, int nthr, unsigned char*) const {
    int cb = (int)(blk % cblocks);
    long r = blk / cblocks;
    int xc = (int)(r % xchunks);
    int b = (int)(r / xchunks);
    int j = cb * nthr + tid;
    if (j >= ncols) return;
    int x0 = (int)((long)nx * xc / xchunks), x1 = (int)((long)nx * (xc + 1) / xchunks);
    const double* col = f + ((long)b * nx) * ld + j;
    // mode 0 and modes 1.. in one sweep over the rows; four rows in flight per iteration
    double s0 = 0.0;
    double sr[4], si[4];
    double wr[4], wi[4], cr[4], ci[4];
    const int nm = nmodes < 5 ? nmodes : 5;
    for (int m = 1; m < nm; ++m) {
      const double ang = -2.0 * 3.14159265358979323846 * (double)m / (double)nx_total;
      sincos_hd(ang * (double)(x0 + x_offset), &wi[m - 1], &wr[m - 1]);
      sincos_hd(ang, &ci[m - 1], &cr[m - 1]);
      sr[m - 1] = 0.0; si[m - 1] = 0.0;
    }
    int x = x0;
    for (; x + 4 <= x1; x += 4) {
      double v0 = col[(long)x * ld], v1 = col[(long)(x + 1) * ld], v2 = col[(long)(x + 2) * ld],
             v3 = col[(long)(x + 3) * ld];
      double vv[4] = {v0, v1, v2, v3};
      s0 += (v0 + v1) + (v2 + v3);
      for (int m = 1; m < nm; ++m) {
        double a = wr[m - 1], bb = wi[m - 1];
        for (int q = 0; q < 4; ++q) {
          sr[m - 1] += vv[q] * a;
          si[m - 1] += vv[q] * bb;
          double na = a * cr[m - 1] - bb * ci[m - 1];
          bb = a * ci[m - 1] + bb * cr[m - 1];
          a = na;
        }
        wr[m - 1] = a; wi[m - 1] = bb;
      }
    }
    for (; x < x1; ++x) {
      double val = col[(long)x * ld];
      s0 += val;
      for (int m = 1; m < nm; ++m) {
        sr[m - 1] += val * wr[m - 1];
        si[m - 1] += val * wi[m - 1];
        double na = wr[m - 1] * cr[m - 1] - wi[m - 1] * ci[m - 1];
        wi[m - 1] = wr[m - 1] * ci[m - 1] + wi[m - 1] * cr[m - 1];
        wr[m - 1] = na;
      }
    }
    long o = ((((long)b * xchunks + xc) * nmodes + 0) * ncols + j) * 2;
    partial[o] = s0;
    partial[o + 1] = 0.0;
    for (int m = 1; m < nm; ++m) {
      o = ((((long)b * xchunks + xc) * nmodes + m) * ncols + j) * 2;
      partial[o] = sr[m - 1];
      partial[o + 1] = si[m - 1];
    }
  }
};

// Two modes (k = 0, 1: what vlapy/core/step.py:130-135 stores), two adjacent columns per thread
// (16-byte loads) and eight rows in flight.  Inside a group of eight rows the phases are constants,
//     sum_q f[xg+q] W^(xg+q) = W^xg * sum_q f[xg+q] W^q,      W = exp(-2 pi i / nx_total),
// so a row costs one add and two FMAs per column, and the running phase W^xg is advanced once per group.
// Same partial-sum layout as XmodesProg (second stage: XmodesReduceProg).
struct Xmodes2Prog {
  const double* f;
  long ld;
  double* partial;  // [batch][xchunks][2][ncols][2]
  int batch, nx, ncols, xchunks, cblocks;
  int x_offset, nx_total;

  VPFP_HD int nphases() const { return 1; }
  VPFP_HD void phase(int, long blk, int tid, int nthr, unsigned char*) const {
    const int cb = (int)(blk % cblocks);
    const long rr = blk / cblocks;
    const int xc = (int)(rr % xchunks);
    const int b = (int)(rr / xchunks);
    const int j = 2 * (cb * nthr + tid);
    if (j >= ncols) return;
    const int x0 = (int)((long)nx * xc / xchunks), x1 = (int)((long)nx * (xc + 1) / xchunks);
    const double* col = f + ((long)b * nx) * ld + j;
    const double ang = -2.0 * 3.14159265358979323846 / (double)nx_total;
#ifndef XMODES_UNROLL
#define XMODES_UNROLL 2    // groups of eight rows in flight (measured at 16384^2 with the 128-thread launch: 1 -> 0.377,
                           // 2 -> 0.370, 4 -> 0.427 ms; under prog_kernel's 1024-thread bound -- 64 registers -- it was 0.41-0.43)
#endif
    double cr[8], ci[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) sincos_hd(ang * (double)q, &ci[q], &cr[q]);
    double gr, gi, wr, wi;                     // W^8 and the running W^xg
    sincos_hd(ang * 8.0, &gi, &gr);
    sincos_hd(ang * (double)(x0 + x_offset), &wi, &wr);
    double s0a = 0.0, s0b = 0.0, sra = 0.0, sia = 0.0, srb = 0.0, sib = 0.0;
    int x = x0;
    constexpr int XU = XMODES_UNROLL;
#pragma unroll XU
    for (; x + 8 <= x1; x += 8) {
      double va[8], vb[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const cplx v2 = *reinterpret_cast<const cplx*>(col + (long)(x + q) * ld);
        va[q] = v2.x; vb[q] = v2.y;
      }
      double ta = 0.0, tb = 0.0, ra = 0.0, ia = 0.0, rb = 0.0, ib = 0.0;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        ta += va[q]; tb += vb[q];
        ra += va[q] * cr[q]; ia += va[q] * ci[q];
        rb += vb[q] * cr[q]; ib += vb[q] * ci[q];
      }
      s0a += ta; s0b += tb;
      sra += ra * wr - ia * wi; sia += ra * wi + ia * wr;
      srb += rb * wr - ib * wi; sib += rb * wi + ib * wr;
      const double nw = wr * gr - wi * gi;
      wi = wr * gi + wi * gr;
      wr = nw;
    }
    for (; x < x1; ++x) {
      const cplx v2 = *reinterpret_cast<const cplx*>(col + (long)x * ld);
      s0a += v2.x; s0b += v2.y;
      sra += v2.x * wr; sia += v2.x * wi;
      srb += v2.y * wr; sib += v2.y * wi;
      const double nw = wr * cr[1] - wi * ci[1];
      wi = wr * ci[1] + wi * cr[1];
      wr = nw;
    }
    long o = ((((long)b * xchunks + xc) * 2 + 0) * ncols + j) * 2;
    partial[o] = s0a; partial[o + 1] = 0.0; partial[o + 2] = s0b; partial[o + 3] = 0.0;
    o = ((((long)b * xchunks + xc) * 2 + 1) * ncols + j) * 2;
    partial[o] = sra; partial[o + 1] = sia; partial[o + 2] = srb; partial[o + 3] = sib;
  }
};

struct XmodesReduceProg {
  const double* partial;
  double* out;  // [batch][nmodes][ncols][2]
  int nmodes, batch, ncols, xchunks;
  VPFP_HD int nphases() const { return 1; }
  VPFP_HD void phase(int, long blk, int tid, int nthr, unsigned char*) const {
    long i = blk * nthr + tid;  // over batch*nmodes*ncols*2
    long per = (long)nmodes * ncols * 2;
    if (i >= per * batch) return;
    long b = i / per, r = i % per;
    double s = 0.0;
    for (int xc = 0; xc < xchunks; ++xc) s += partial[((long)b * xchunks + xc) * per + r];
    out[i] = s;
  }
};

// ---------------------------------------------------------------------------------------------
// Ponderomotive driver (vlapy/field_driver.py:24-50), elementwise over x.
// ---------------------------------------------------------------------------------------------
#define DRIVER_MAX_PULSES 8
struct DriverProg {
  const double* x;
  double* out;
  double t;
  const double* t_dev;   // optional: time base read on the device (CUDA-graph replay), then the
  int ninc;              // increments are added one by one in the reference's order
  double inc[6];
  int nx, npulse;
  double pulses[DRIVER_MAX_PULSES * 7];  // k0, w0, a0, t_L, t_R, t_wL, t_wR

  VPFP_HD int nphases() const { return 1; }
  VPFP_HD void phase(int, long blk, int tid, int nthr, unsigned char*) const {
    long i = blk * nthr + tid;
    if (i >= nx) return;
    double t = this->t;
    if (t_dev) {
      t = t_dev[0];
      for (int k = 0; k < ninc; ++k) t = t + inc[k];
    }
    double total = 0.0;
    for (int p = 0; p < npulse; ++p) {
      const double* q = pulses + 7 * p;
      double env = 0.5 * (tanh((t - q[3]) / q[5]) - tanh((t - q[4]) / q[6]));
      total += env * q[0] * q[2] * sin(q[0] * x[i] - q[1] * t);
    }
    out[i] = total;
  }
};

// Ensemble driver: simulation b has its own grid row x[b*nx ...] and pulse parameters
// pulses[b][npulse][7] (device memory); one time for all.
struct DriverBatchProg {
  const double* x;
  const double* pulses;
  double* out;
  double t;
  const double* t_dev;
  int ninc;
  double inc[6];
  int nx, npulse, batch;

  VPFP_HD int nphases() const { return 1; }
  VPFP_HD void phase(int, long blk, int tid, int nthr, unsigned char*) const {
    long i = blk * nthr + tid;
    if (i >= (long)nx * batch) return;
    double t = this->t;
    if (t_dev) {
      t = t_dev[0];
      for (int k = 0; k < ninc; ++k) t = t + inc[k];
    }
    const double* q0 = pulses + (i / nx) * npulse * 7;
    double total = 0.0;
    for (int p = 0; p < npulse; ++p) {
      const double* q = q0 + 7 * p;
      double env = 0.5 * (tanh((t - q[3]) / q[5]) - tanh((t - q[4]) / q[6]));
      total += env * q[0] * q[2] * sin(q[0] * x[i] - q[1] * t);
    }
    out[i] = total;
  }
};

// ---------------------------------------------------------------------------------------------
// Series means of one step (vlapy/core/step.py:202-224): single CTA.
// out = [mean n, mean j, mean T, mean e^2, mean de^2, mean int f^2, mean int f ln f]
// ---------------------------------------------------------------------------------------------
struct SeriesProg {
  const double* mom;
  long mom_ld;
  const double* e;
  const double* de;
  double* out;
  int nx;
  VPFP_HD int nphases(int nthr) const { return 2 + tree_phases(nthr); }
  VPFP_HD long smem_bytes(int nthr) const { return (long)7 * nthr * sizeof(double); }
  // blk = simulation index of an ensemble (rows blk*nx .. blk*nx + nx - 1 of the moment arrays)
  VPFP_HD void phase(int ph, long blk, int tid, int nthr, unsigned char* smem) const {
    double* part = reinterpret_cast<double*>(smem);
    const int nt = tree_phases(nthr);
    const long o = blk * nx;
    if (ph == 0) {
      // one CTA per simulation: the loop is latency bound, so two cells per thread (1024 threads) are in flight
      double a[7] = {0, 0, 0, 0, 0, 0, 0};
      int i = tid;
      for (; i + nthr < nx; i += 2 * nthr) {
        double v[2][7];
        for (int q = 0; q < 2; ++q) {
          const long c = o + i + (long)q * nthr;
          v[q][0] = mom[0 * mom_ld + c];
          v[q][1] = mom[1 * mom_ld + c];
          v[q][2] = mom[2 * mom_ld + c];
          v[q][3] = e[c];
          v[q][4] = de ? de[c] : 0.0;
          v[q][5] = mom[6 * mom_ld + c];
          v[q][6] = mom[7 * mom_ld + c];
        }
        for (int q = 0; q < 2; ++q) {
          a[0] += v[q][0]; a[1] += v[q][1]; a[2] += v[q][2];
          a[3] += v[q][3] * v[q][3]; a[4] += v[q][4] * v[q][4];
          a[5] += v[q][5]; a[6] += v[q][6];
        }
      }
      for (; i < nx; i += nthr) {
        a[0] += mom[0 * mom_ld + o + i];
        a[1] += mom[1 * mom_ld + o + i];
        a[2] += mom[2 * mom_ld + o + i];
        a[3] += e[o + i] * e[o + i];
        a[4] += de ? de[o + i] * de[o + i] : 0.0;
        a[5] += mom[6 * mom_ld + o + i];
        a[6] += mom[7 * mom_ld + o + i];
      }
      for (int k = 0; k < 7; ++k) part[(long)k * nthr + tid] = a[k];
    } else if (ph <= nt) {
      tree_step(ph - 1, nthr, 7, tid, nthr, part);
    } else if (tid < 7) {
      out[blk * 7 + tid] = part[(long)tid * nthr] / (double)nx;
    }
  }
};
