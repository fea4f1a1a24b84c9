// advect.h -- the spectral advection / Poisson "program" (FFT -> pointwise multiply -> inverse FFT).
//
// Replaces the reference's exponential integrators (vlapy/core/vlasov.py:94-108 and :123-138)
// and its spectral Poisson solve (vlapy/core/field.py:39-63).
//
// Data model.  Two real sequences are packed into one complex sequence z = a + i b of length N:
//   COLS mode (x-FFT, vdfdx): a, b are two ADJACENT v-columns of f, so one 16-byte load of
//        f[x][2s], f[x][2s+1] is exactly (re, im): packing costs nothing and every access is a
//        full 16 B (B consecutive packed columns = one 128 B line).
//   ROWS mode (v-FFT, edfdv, Poisson): a, b are two ADJACENT rows (re from row 2s, im from 2s+1).
// The transform of z gives both spectra; the pointwise step un-mixes them on the fly
//   Y[k]   = S Z[k]   + D conj(Z[N-k]),     S = (Pa + Pb)/2N, D = (Pa - Pb)/2N
//   Y[N-k] = S*Z[N-k] + D* conj(Z[k])
// with Pa, Pb the phase factors of the two channels for bin k <= N/2 (the Nyquist factor is
// replaced by its real part, which is what np.real(...) of the reference amounts to).
//
// Decomposition.  N = N1 * N2 (n = n1 N2 + n2, k = k1 + N1 k2).  N1 == 1: one kernel (pass 0) does
// load -> DIF FFT -> pointwise -> DIT inverse -> store with the whole sequence in shared memory.
// N1 > 1 (sequences that do not fit a CTA: x-columns are strided by the row pitch): three passes,
//   pass 1: length-N1 DIF over n1 for fixed n2, times W_N^(n2 k1)            (slot layout)
//   pass 2: for the group pair {k1, N1-k1}: length-N2 DIF over n2, pointwise, length-N2 DIT
//           inverse, times conj W_N^(n2 k1)
//   pass 3: length-N1 DIT inverse over the k1 slots -> natural order.
// DIF leaves bit-reversed order and DIT consumes it, so nothing is ever reordered: slot p of a
// length-L transform holds frequency brev(p).  All passes run in place on f_out.
//
// Shared-memory tile: T[g][l][b], index (g*L + l)*Bp + b, lanes run along b (the batch of packed
// sequences or, for ROWS pass 1/3, the n2 sub-range), so butterflies are bank-conflict free and
// twiddles are warp-uniform whenever B >= 32.  L and B are powers of two.
#pragma once
#include "vpfp_common.h"

enum { ADV_COLS = 0, ADV_ROWS = 1 };
enum { OP_PHASE = 0, OP_POISSON = 1 };

struct AdvectProg {
  int mode, pass, op;
  int N, N1, N2, lN1, lN2;
  int L, lL, G, B, lB, Bp;  // in-tile transform length L = 1<<lL, groups, batch B = 1<<lB, padded pitch
  int nsim, nseq, nrows;    // COLS: nseq = ncols/2 per sim; ROWS: nrows = total rows, nseq = ceil(nrows/2)
  int tiles_b;              // batch tiles (COLS: packed columns; ROWS pass 0/2: row pairs; pass 1/3: n2 blocks)
  int T1;                   // pass 2: number of group-pair tiles (N1/2); 1 otherwise
  const double* fin;
  long ld_in;
  double* fout;
  long ld_out;
  const double* kvec;  // wavenumbers [nsim][N] (COLS) or [N] (ROWS); OP_POISSON: one_over_kx [nrows][N]
  const double* cvec;  // COLS: v[ncols]; ROWS: e[nrows]; OP_POISSON: unused
  const double* addv;  // OP_POISSON: driver rows [nrows][N] added on the final store (nullable)
  double* phantom;     // ROWS, three-pass, odd row count: N doubles standing in for the missing
                       // partner row of the last packed pair (its intermediate spectrum is complex)
  double dt;
  double inv_n;
  const cplx* tw;  // exp(-2 pi i m / N), m in [0, N)

  // ------------------------------------------------------------------ geometry
  VPFP_HD int nstages() const { return lL / 2 + (lL & 1); }
  VPFP_HD int nphases() const {
    if (pass == 1 || pass == 3) return 2 + nstages();
    return 3 + 2 * nstages();
  }
  VPFP_HD int tile_elems() const { return G * L * B; }
  VPFP_HD long smem_bytes() const { return (long)G * L * Bp * (long)sizeof(cplx); }
  VPFP_HD long ntiles() const {
    if (mode == ADV_COLS) {
      if (pass == 0) return (long)nsim * tiles_b;
      if (pass == 2) return (long)nsim * T1 * tiles_b;
      return (long)nsim * N2 * tiles_b;
    }
    if (pass == 0) return tiles_b;
    if (pass == 2) return (long)tiles_b * T1;
    return (long)nseq * tiles_b;
  }

  struct Tile {
    int sim, bt, n2, t1, seq;  // which of these are meaningful depends on mode/pass
    int k1[2];                 // pass 2: frequency k1 of each group
    int p1[2];                 // pass 2: its slot brev(k1)
    int self;                  // groups are self-paired (k1 == N1 - k1 mod N1)
  };

  VPFP_HD Tile decode(long blk) const {
    Tile t;
    t.sim = 0; t.bt = 0; t.n2 = 0; t.t1 = 0; t.seq = 0;
    t.p1[0] = t.p1[1] = 0; t.k1[0] = t.k1[1] = 0; t.self = 1;
    if (mode == ADV_COLS) {
      t.bt = (int)(blk % tiles_b);
      long r = blk / tiles_b;
      if (pass == 0) {
        t.sim = (int)r;
      } else if (pass == 2) {
        t.t1 = (int)(r % T1);
        t.sim = (int)(r / T1);
      } else {
        t.n2 = (int)(r % N2);
        t.sim = (int)(r / N2);
      }
    } else {
      if (pass == 0) {
        t.bt = (int)blk;
      } else if (pass == 2) {
        t.t1 = (int)(blk % T1);
        t.bt = (int)(blk / T1);
      } else {
        t.bt = (int)(blk % tiles_b);
        t.seq = (int)(blk / tiles_b);
      }
    }
    if (pass == 2) {
      if (t.t1 == 0) {
        t.k1[0] = 0; t.k1[1] = N1 / 2; t.self = 1;
      } else {
        t.k1[0] = t.t1; t.k1[1] = N1 - t.t1; t.self = 0;
      }
      t.p1[0] = (int)brev_bits((unsigned)t.k1[0], lN1);
      t.p1[1] = (int)brev_bits((unsigned)t.k1[1], lN1);
    }
    return t;
  }

  // tile element (g, l, b) -> packed sequence id and position n inside the sequence
  VPFP_HD void locate(const Tile& t, int g, int l, int b, int* seq, long* n) const {
    if (pass == 0) { *seq = t.bt * B + b; *n = l; return; }
    if (pass == 2) { *seq = t.bt * B + b; *n = (long)t.p1[g] * N2 + l; return; }
    if (mode == ADV_COLS) { *seq = t.bt * B + b; *n = (long)l * N2 + t.n2; return; }
    *seq = t.seq;
    *n = (long)l * N2 + (long)t.bt * B + b;
  }

  // which tile index runs along the lanes of a warp in the global load/store loops: the one that
  // is contiguous in global memory (the packed-column / n2 batch, or for ROWS pass 0/2 the
  // sequence position itself).
  VPFP_HD bool lanes_along_l() const { return mode == ADV_ROWS && (pass == 0 || pass == 2); }

  VPFP_HD void split_idx(int idx, int* g, int* l, int* b) const {
    if (lanes_along_l()) {
      *l = idx & (L - 1);
      *b = (idx >> lL) & (B - 1);
    } else {
      *b = idx & (B - 1);
      *l = (idx >> lB) & (L - 1);
    }
    *g = idx >> (lL + lB);
  }

  // ------------------------------------------------------------------ global <-> tile
  VPFP_HD cplx gload(const double* base, long ld, const Tile& t, int seq, long n) const {
    if (seq >= nseq) return cmake(0.0, 0.0);
    if (mode == ADV_COLS) {
      const double* p = base + ((long)t.sim * N + n) * ld + 2 * (long)seq;
      return *reinterpret_cast<const cplx*>(p);
    }
    long ra = 2 * (long)seq, rb = ra + 1;
    double re = base[ra * ld + n];
    double im = (rb < nrows) ? base[rb * ld + n] : ((pass >= 2 && phantom) ? phantom[n] : 0.0);
    return cmake(re, im);
  }

  VPFP_HD void gstore(double* base, long ld, const Tile& t, int seq, long n, cplx val) const {
    if (seq >= nseq) return;
    if (mode == ADV_COLS) {
      double* p = base + ((long)t.sim * N + n) * ld + 2 * (long)seq;
      *reinterpret_cast<cplx*>(p) = val;
      return;
    }
    long ra = 2 * (long)seq, rb = ra + 1;
    base[ra * ld + n] = val.x;
    if (rb < nrows) base[rb * ld + n] = val.y;
    else if (pass != 0 && pass != 3 && phantom) phantom[n] = val.y;
  }

  // ------------------------------------------------------------------ FFT stages on the tile
  VPFP_HD cplx twiddle(long m) const {  // exp(-2 pi i m / N)
    const double* p = reinterpret_cast<const double*>(tw + m);
    return cmake(ldg(p), ldg(p + 1));
  }

  VPFP_HD void radix2_span1(int tid, int nthr, cplx* T) const {
    const int work = G * (L / 2) * B;
    for (int idx = tid; idx < work; idx += nthr) {
      int b = idx & (B - 1);
      int r = idx >> lB;  // runs over g*(L/2) + pair
      cplx* p = T + (long)(2 * r) * Bp + b;
      cplx a = p[0], bb = p[Bp];
      p[0] = cadd(a, bb);
      p[Bp] = csub(a, bb);
    }
  }

  // forward DIF stage s (0-based). Radix-4 with quarter-span q = L >> 2(s+1); the last stage of an
  // odd-log2 length is a twiddle-free radix-2 with span 1.
  VPFP_HD void fwd_stage(int s, int tid, int nthr, cplx* T) const {
    const int n4 = lL / 2;
    if (s >= n4) { radix2_span1(tid, nthr, T); return; }
    const int q = L >> (2 * (s + 1));
    const long tstride = (long)N / (4 * q);  // W_{4q}^j = tw[j * N/(4q)]
    const int work = G * (L / 4) * B;
    for (int idx = tid; idx < work; idx += nthr) {
      int b = idx & (B - 1);
      int r = (idx >> lB) & (L / 4 - 1);
      int g = idx >> (lB + lL - 2);
      int j = r & (q - 1);
      int i = ((r - j) << 2) + j;
      cplx* p = T + ((long)g * L + i) * Bp + b;
      const long st = (long)q * Bp;
      cplx a = p[0], bb = p[st], c = p[2 * st], d = p[3 * st];
      cplx apc = cadd(a, c), amc = csub(a, c), bpd = cadd(bb, d), bmd = csub(bb, d);
      cplx y0 = cadd(apc, bpd);
      cplx y2 = csub(apc, bpd);
      cplx y1 = cadd(amc, cmul_mi(bmd));  // (a-c) - i (b-d)
      cplx y3 = cadd(amc, cmul_i(bmd));   // (a-c) + i (b-d)
      if (j != 0) {
        cplx w1 = twiddle(j * tstride), w2 = twiddle(2 * j * tstride), w3 = twiddle(3 * j * tstride);
        y1 = cmul(y1, w1);
        y2 = cmul(y2, w2);
        y3 = cmul(y3, w3);
      }
      // bit-reversed placement of the four outputs: (y0, y2, y1, y3)
      p[0] = y0; p[st] = y2; p[2 * st] = y1; p[3 * st] = y3;
    }
  }

  // inverse DIT stage s (0-based): mirror image of fwd_stage with conjugated twiddles.
  VPFP_HD void inv_stage(int s, int tid, int nthr, cplx* T) const {
    const int odd = lL & 1;
    if (odd && s == 0) { radix2_span1(tid, nthr, T); return; }
    const int s4 = s - odd;
    const int q = (odd ? 2 : 1) << (2 * s4);
    const long tstride = (long)N / (4 * q);
    const int work = G * (L / 4) * B;
    for (int idx = tid; idx < work; idx += nthr) {
      int b = idx & (B - 1);
      int r = (idx >> lB) & (L / 4 - 1);
      int g = idx >> (lB + lL - 2);
      int j = r & (q - 1);
      int i = ((r - j) << 2) + j;
      cplx* p = T + ((long)g * L + i) * Bp + b;
      const long st = (long)q * Bp;
      cplx a = p[0], bb = p[st], c = p[2 * st], d = p[3 * st];
      if (j != 0) {
        cplx w1 = twiddle(j * tstride), w2 = twiddle(2 * j * tstride), w3 = twiddle(3 * j * tstride);
        bb = cmulc(bb, w2);
        c = cmulc(c, w1);
        d = cmulc(d, w3);
      }
      cplx s0 = cadd(a, bb), s1 = csub(a, bb), s2 = cadd(c, d), s3 = csub(c, d);
      p[0] = cadd(s0, s2);
      p[2 * st] = csub(s0, s2);
      p[st] = cadd(s1, cmul_i(s3));
      p[3 * st] = csub(s1, cmul_i(s3));
    }
  }

  // ------------------------------------------------------------------ pointwise
  // Phase factor of bin k (0 <= k <= N/2) for advection constant c, the reference's rounding:
  // theta = (K[k] * dt) * c, P = cos(theta) - i sin(theta); Nyquist keeps only the real part.
  VPFP_HD cplx phase_exact(double Kk, double c, bool nyq) const {
    double th = mul_rn(mul_rn(Kk, dt), c);
    double s, co;
    sincos_hd(th, &s, &co);
    return nyq ? cmake(co, 0.0) : cmake(co, -s);
  }

  VPFP_HD void pointwise(const Tile& t, int tid, int nthr, cplx* T) const {
    // work items = bins k <= N/2 of the tile's groups: (g, k2 < L/2, b) plus, for the group with
    // k1 == 0 (always g == 0 of a self tile), the Nyquist bin k2 = L/2.
    const int half = L / 2;
    const int main_work = G * half * B;
    const int work = main_work + (t.self ? B : 0);
    for (int idx = tid; idx < work; idx += nthr) {
      int b, kk2, g;
      if (idx < main_work) {
        b = idx & (B - 1);
        kk2 = (idx >> lB) & (half - 1);
        g = idx >> (lB + lL - 1);
      } else {
        b = idx - main_work; kk2 = half; g = 0;
      }
      const int k1 = (pass == 2) ? t.k1[g] : 0;
      const long k = (long)k1 + (long)N1 * kk2;
      const long kp = (N - k) & (long)(N - 1);
      const int gp = t.self ? g : 1 - g;
      const int p = (int)brev_bits((unsigned)kk2, lL);
      const int pp = (int)brev_bits((unsigned)(kp >> lN1), lL);
      cplx* zp = T + ((long)g * L + p) * Bp + b;
      cplx* zq = T + ((long)gp * L + pp) * Bp + b;
      const int seq = t.bt * B + b;
      if (seq >= nseq) continue;
      const long ra = 2 * (long)seq, rb = ra + 1;
      const bool selfpair = (kp == k);
      cplx Pa, Pb;
      if (op == OP_POISSON) {
        // E_k = i * ook[k] * rho_k; np.real() keeps the Hermitian part, so the effective
        // multiplier is M[k] = i (ook[k] - ook[N-k]) / 2, each packed channel with its own row.
        double oa = 0.5 * (kvec[ra * N + k] - kvec[ra * N + kp]);
        double ob = (rb < nrows) ? 0.5 * (kvec[rb * N + k] - kvec[rb * N + kp]) : 0.0;
        Pa = cmake(0.0, oa);
        Pb = cmake(0.0, ob);
      } else {
        double ca, cb, Kk;
        if (mode == ADV_COLS) {
          ca = cvec[ra];
          cb = cvec[rb];
          Kk = kvec[(long)t.sim * N + k];
        } else {
          ca = cvec[ra];
          cb = (rb < nrows) ? cvec[rb] : 0.0;
          Kk = kvec[k];
        }
        const bool nyq = (2 * k == N);
        Pa = phase_exact(Kk, ca, nyq);
        Pb = phase_exact(Kk, cb, nyq);
      }
      const cplx Z = *zp, Zp = *zq;
      const cplx S = cscale(cadd(Pa, Pb), 0.5 * inv_n), D = cscale(csub(Pa, Pb), 0.5 * inv_n);
      const cplx Y = cadd(cmul(S, Z), cmul(D, cconj(Zp)));
      if (!selfpair) *zq = cadd(cmul(cconj(S), Zp), cmul(cconj(D), cconj(Z)));
      *zp = Y;
    }
  }

  // ------------------------------------------------------------------ phases
  VPFP_HD void load_phase(const Tile& t, int tid, int nthr, cplx* T) const {
    const bool first = (pass == 0 || pass == 1);
    const double* src = first ? fin : fout;
    const long ld = first ? ld_in : ld_out;
    const int work = tile_elems();
    for (int idx = tid; idx < work; idx += nthr) {
      int g, l, b, seq; long n;
      split_idx(idx, &g, &l, &b);
      locate(t, g, l, b, &seq, &n);
      cplx val = gload(src, ld, t, seq, n);
      if (op == OP_POISSON && first && seq < nseq) {
        // net charge 1 - n (field.py:61-63); fin holds the density rows
        val.x = 1.0 - val.x;
        val.y = (2 * (long)seq + 1 < nrows) ? 1.0 - val.y : 0.0;
      }
      T[((long)g * L + l) * Bp + b] = val;
    }
  }

  VPFP_HD void store_phase(const Tile& t, int tid, int nthr, cplx* T) const {
    const int work = tile_elems();
    for (int idx = tid; idx < work; idx += nthr) {
      int g, l, b, seq; long n;
      split_idx(idx, &g, &l, &b);
      locate(t, g, l, b, &seq, &n);
      cplx val = T[((long)g * L + l) * Bp + b];
      if (pass == 1) {
        // slot l holds k1 = brev(l); the four-step twiddle W_N^(n2 k1)
        long n2 = n & (long)(N2 - 1);
        long m = n2 * (long)brev_bits((unsigned)l, lN1);
        if (m != 0) val = cmul(val, twiddle(m));
      } else if (pass == 2) {
        long m = (long)l * t.k1[g];
        if (m != 0) val = cmulc(val, twiddle(m));
      }
      if (op == OP_POISSON && (pass == 0 || pass == 3) && addv != nullptr && seq < nseq) {
        long ra = 2 * (long)seq, rb = ra + 1;
        val.x += addv[ra * N + n];
        if (rb < nrows) val.y += addv[rb * N + n];
      }
      gstore(fout, ld_out, t, seq, n, val);
    }
  }

  VPFP_HD void phase(int ph, long blk, int tid, int nthr, unsigned char* smem) const {
    cplx* T = reinterpret_cast<cplx*>(smem);
    const Tile t = decode(blk);
    const int ns = nstages();
    if (ph == 0) { load_phase(t, tid, nthr, T); return; }
    if (pass == 1) {
      if (ph <= ns) fwd_stage(ph - 1, tid, nthr, T);
      else store_phase(t, tid, nthr, T);
      return;
    }
    if (pass == 3) {
      if (ph <= ns) inv_stage(ph - 1, tid, nthr, T);
      else store_phase(t, tid, nthr, T);
      return;
    }
    if (ph <= ns) fwd_stage(ph - 1, tid, nthr, T);
    else if (ph == ns + 1) pointwise(t, tid, nthr, T);
    else if (ph <= 2 * ns + 1) inv_stage(ph - ns - 2, tid, nthr, T);
    else store_phase(t, tid, nthr, T);
  }
};

// ---------------------------------------------------------------------------------------------
// Host-side planning shared by the CUDA launcher and the emulator: pick N1/N2, tile batch and CTA
// size for a transform of length N.  max_single = longest sequence handled by one CTA.
// ---------------------------------------------------------------------------------------------
struct AdvectPlan {
  int N1, N2;
  int B[4];        // tile batch per pass (index = pass id 0..3)
  int threads[4];
};

inline AdvectPlan make_advect_plan(int mode, int N, int max_single_cols = 2048, int max_single_rows = 8192) {
  AdvectPlan pl;
  const int lN = ilog2(N);
  const int max_single = (mode == ADV_COLS) ? max_single_cols : max_single_rows;
  if (N <= max_single) {
    pl.N1 = 1;
    pl.N2 = N;
  } else {
    int l2 = (lN + 1) / 2;  // N2 >= N1
    pl.N2 = 1 << l2;
    pl.N1 = N >> l2;
  }
  for (int p = 0; p < 4; ++p) { pl.B[p] = 1; pl.threads[p] = 128; }
  auto thr = [](long elems) {
    long t = elems / 8;
    if (t < 64) t = 64;
    if (t > 1024) t = 1024;
    return (int)t;
  };
  if (mode == ADV_COLS) {
    // lanes along packed columns: 8 packed columns = one 128-byte line
    int b0 = 8;
    while ((long)pl.N2 * b0 * 16 > 128 * 1024 && b0 > 1) b0 >>= 1;
    pl.B[0] = b0; pl.B[1] = 8; pl.B[2] = 8; pl.B[3] = 8;
    pl.threads[0] = thr((long)pl.N2 * b0);
    pl.threads[1] = pl.threads[3] = thr((long)pl.N1 * 8);
    pl.threads[2] = thr(2L * pl.N2 * 8);
  } else {
    int b0 = 1;
    while ((long)pl.N2 * b0 * 2 <= 2048 && b0 < 8) b0 <<= 1;   // small rows: several row pairs per CTA
    pl.B[0] = b0;
    pl.B[1] = pl.B[3] = (pl.N2 >= 16) ? 16 : pl.N2;              // n2 block: 16 doubles = 128 B per row
    pl.B[2] = 4;
    pl.threads[0] = thr((long)pl.N2 * b0);
    pl.threads[1] = pl.threads[3] = thr((long)pl.N1 * pl.B[1]);
    pl.threads[2] = thr(2L * pl.N2 * 4);
  }
  return pl;
}

// Fill the geometry fields of an AdvectProg for a given pass from a plan.
inline void advect_set_pass(AdvectProg& a, const AdvectPlan& pl, int pass) {
  a.pass = pass;
  a.N1 = pl.N1; a.N2 = pl.N2; a.lN1 = ilog2(pl.N1); a.lN2 = ilog2(pl.N2);
  a.B = pl.B[pass]; a.lB = ilog2(a.B);
  const bool lanes_l = (a.mode == ADV_ROWS && (pass == 0 || pass == 2));
  a.Bp = (lanes_l && a.B > 1) ? a.B + 1 : a.B;   // odd pitch keeps the transposing store conflict-free
  if (pass == 0) { a.L = a.N; a.G = 1; a.T1 = 1; }
  else if (pass == 2) { a.L = pl.N2; a.G = 2; a.T1 = pl.N1 / 2; }
  else { a.L = pl.N1; a.G = 1; a.T1 = 1; }
  a.lL = ilog2(a.L);
  if (a.mode == ADV_COLS) a.tiles_b = (a.nseq + a.B - 1) / a.B;
  else if (pass == 0 || pass == 2) a.tiles_b = (a.nseq + a.B - 1) / a.B;
  else a.tiles_b = pl.N2 / a.B;
  a.inv_n = 1.0 / (double)a.N;
}
