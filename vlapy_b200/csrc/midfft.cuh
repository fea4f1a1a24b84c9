// midfft.cuh -- single-pass advection for MID-SIZE transforms, L in {256, 512, 1024, 2048}: the whole packed sequence
// is transformed, phase-multiplied and transformed back by ONE kernel with 16 complex values per thread in
// registers, so the operator costs one read and one write of f (16 B/cell).  Serves the grids of BASELINE configs
// 1, 2 and 4 (C1 32 x 512, C2 256 x 2048, C4 1024 x (256 x 512)):
//     ROWS  e df/dv, vlapy/core/vlasov.py:123-138, nv = L   (two adjacent rows packed as one complex sequence)
//     COLS  v df/dx, vlapy/core/vlasov.py:94-108,  nx = L   (two adjacent v-columns packed: one 16-byte load)
// Before, these sizes took either the generic shared-memory program of advect.h (one launch, but radix-2/4 stages
// through shared memory: ~1 TB/s) or the three register passes of advect_fast.cuh (48 B/cell).
//
// Decomposition L = R1 * R2 * 8, n = n1 (L/R1) + n2 8 + n3, k = k1 + R1 k2 + R1 R2 k3:
//   stage A: radix-R1 over n1 for fixed r = (n2, n3), twiddle W_L^(r k1)                      -> y[k1][r]
//   stage B: radix-R2 over n2 for fixed (k1, n3), twiddle W_L^(R1 n3 k2), IN PLACE             -> z[s = k1 + R1 k2][n3]
//   stage C: radix-8 over n3; a thread owns sub-transforms s and S - s (S = L/8), i.e. BOTH members of every pair
//            (k, L - k): bin s + S k3 pairs with (S - s) + S (7 - k3); thread 0 owns the self-paired s = 0 and S/2.
// The pointwise step un-mixes the two packed real channels on the pair (advect_fast.cuh pass 2):
//   U = Z + conj Z', V = Z - conj Z', X1 = Pa U, X2 = Pb V, Y[k] = X1 + X2, Y[L-k] = conj(X1 - X2), with the phase
//   factors P(j) = exp(-i phi j) / (2L) = T0[j & 15] T12[j >> 4] from two small tables per channel (phi = (K[1] dt) c,
//   the reference's two roundings; uniform fftfreq grids: VPFP_PHASE_TABLE).  The Nyquist factor keeps its real part
//   (np.real of the reference).
// The kernel body is a phase program whose per-thread registers persist across barriers (as rowfft.cuh); tests/emul
// runs the same source thread by thread on the host.
#pragma once
#include "advect.h"
#include "butterflies.h"

namespace midfft {

using fast::fft8;
using fast::fftR;

struct Args {
  int nsim, nseq, nrows;     // COLS: nseq = ncols / 2 per simulation; ROWS: nrows rows, nseq = ceil(nrows / 2)
  const double* fin; long ld_in;
  double* fout; long ld_out;
  const double* kvec;        // COLS: [nsim][L]; ROWS: [L]   (only K[1] is used)
  const double* cvec;        // COLS: v[ncols]; ROWS: e[nrows]
  double dt;
  const cplx* tw;            // exp(-2 pi i m / L), L entries
  // fused charge density (COLS, Prog<..., DENS = true>): dens_partial[(column tile * nsim + sim) * L + x] = weighted sum
  // (trapezoid, vlapy/core/field.py:27-36) over the 2 CB columns of the tile; summed over the tiles by
  // fast::dens_reduce_kernel.  edge_flags: bit 0 / 1 = the first / last column is an end of the global v axis.
  double* dens_partial; double dv; int edge_flags;
  // Poisson mode (Prog<..., POISSON = true>, ROWS): fin = n [nrows][L], fout = e, kvec = one_over_kx [nrows][L],
  // addv = driver field [nrows][L] or null
  const double* addv;
};

// asynchronous copies global -> shared (host emulation: plain copies)
VPFP_HD void cp_async(void* smem, const void* gmem, int bytes16) {
#if defined(__CUDA_ARCH__)
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  if (bytes16) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gmem) : "memory");
  else asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(sa), "l"(gmem) : "memory");
#else
  memcpy(smem, gmem, bytes16 ? 16 : 8);
#endif
}
VPFP_HD void cp_async_commit() {
#if defined(__CUDA_ARCH__)
  asm volatile("cp.async.commit_group;\n" ::: "memory");
#endif
}
VPFP_HD void cp_async_wait() {
#if defined(__CUDA_ARCH__)
  asm volatile("cp.async.wait_group 0;\n" ::: "memory");
#endif
}

// POISSON_: the spectral field solve of vlapy/core/field.py:39-88 for nx = L in one launch -- two density rows (two
// simulations of an ensemble) packed as one complex sequence, loads 1 - n, the pointwise step multiplies by
// i (one_over_kx[k] - one_over_kx[L-k]) / 2 (the Hermitian part of the reference's multiplier: what survives its np.real)
// instead of a phase, the store adds the driver row.  Before: the generic shared-memory program, 16 us at nx = 256.
template <int L_, int R1_, int R2_, int MODE_, int CB_, bool DENS_ = false, bool POISSON_ = false>
struct Prog {
  static constexpr int L = L_, R1 = R1_, R2 = R2_, MODE = MODE_, CB = CB_;
  static constexpr bool DENS = DENS_, POISSON = POISSON_;
  static_assert(!POISSON_ || (MODE_ == ADV_ROWS && !DENS_), "Poisson mode packs density rows");
  static_assert(!DENS_ || MODE_ == ADV_COLS, "the fused density belongs to v df/dx");
  static_assert(R1 * R2 * 8 == L, "L = R1 * R2 * 8");
  static constexpr int S = L / 8;              // stage-C sub-transforms
  static constexpr int TPC = L / 16;           // threads per sequence
  static constexpr int NT = CB * TPC;          // threads per CTA
  static constexpr int LA = L / R1;            // points per k1 row
  static constexpr int PITCH = LA + 1;         // odd pitch (in 16-byte units) of a k1 row
  static constexpr int XELEMS = R1 * PITCH;    // exchange buffer of one sequence
  static constexpr int NQA = 16 / R1, NQB = 16 / R2;
  static constexpr int NT12 = L / 32 + 1;      // table over j >> 4, j = 0 .. L/2
  static constexpr int NPH = DENS ? 8 : 5;
  static constexpr int DP = CB * TPC + 1;      // pitch of the density exchange buffer D[8][DP] (doubles): two rounds of 8
                                               // output slots, so that the buffer does not cost a resident CTA
  static constexpr int NTWB = R2 * 8;          // stage-B twiddles
  static constexpr bool PREFETCH = (MODE == ADV_COLS);   // ROWS would prefetch 8-byte pieces (2-way bank conflicts): direct loads
  static constexpr long SMEM_BYTES = (long)sizeof(cplx) * ((long)XELEMS * CB + L + NTWB + 2L * CB * (16 + NT12)) +
                                     (DENS ? (long)sizeof(double) * 8 * DP : 0);

  struct Regs {
    cplx x[16];
  };

  Args a;

  VPFP_HD long ntiles() const {
    const long tb = (a.nseq + CB - 1) / CB;
    return (MODE == ADV_COLS) ? (long)a.nsim * tb : tb;
  }
  VPFP_HD static cplx* xbuf(unsigned char* smem) { return reinterpret_cast<cplx*>(smem); }
  // twiddles laid out the way the lanes read them (ROWS: lanes run along r resp. n3; a single exp(-2 pi i m / L)
  // table read at m = r k1 and m = R1 n3 k2 cost up to 8-way bank conflicts: 47 % of all shared-memory wavefronts)
  VPFP_HD static cplx* twa(unsigned char* smem) { return xbuf(smem) + (long)XELEMS * CB; }   // [R1][L/R1]: W_L^(r k1)
  VPFP_HD static cplx* twb(unsigned char* smem) { return twa(smem) + L; }                    // [R2][8]: W_L^(R1 n3 k2)
  VPFP_HD static cplx* t0(unsigned char* smem) { return twb(smem) + NTWB; }            // [2][CB][16]
  VPFP_HD static cplx* t12(unsigned char* smem) { return t0(smem) + 2 * CB * 16; }     // [2][CB][NT12]
  VPFP_HD static double* dbuf(unsigned char* smem) { return reinterpret_cast<double*>(t12(smem) + 2 * CB * NT12); }
  // phase-table index of entry i (of n) of channel ch of sequence b: the lanes of a warp run along b (COLS) or share b
  // and differ in i (ROWS)
  VPFP_HD static int tix(int ch, int b, int i, int n) { return (MODE == ADV_COLS) ? (ch * n + i) * CB + b : (ch * CB + b) * n + i; }
  // exchange-buffer index of slot `slot` of sequence b: lanes run along b (COLS) or along the slot (ROWS)
  VPFP_HD static int xi(int b, int slot) { return (MODE == ADV_COLS) ? slot * CB + b : b * XELEMS + slot; }
  VPFP_HD static void roles(int tid, int* b, int* u) {
    if (MODE == ADV_COLS) { *b = tid % CB; *u = tid / CB; }
    else { *u = tid % TPC; *b = tid / TPC; }
  }

  // once per CTA: the twiddle tables
  VPFP_HD void init(int tid, unsigned char* smem) const {
    cplx* TWA = twa(smem);
    cplx* TWB = twb(smem);
    for (int j = tid; j < L; j += NT) TWA[j] = a.tw[((j % LA) * (j / LA)) % L];
    for (int j = tid; j < NTWB; j += NT) TWB[j] = a.tw[((j & 7) * (j >> 3) * R1) % L];
  }

  struct Tile {
    int sim, seq0;
  };
  VPFP_HD Tile decode(long tile) const {
    Tile t;
    const long tb = (a.nseq + CB - 1) / CB;
    t.sim = (MODE == ADV_COLS) ? (int)(tile / tb) : 0;
    t.seq0 = (int)(tile % tb) * CB;
    return t;
  }
  // advection constants of the two packed channels of sequence seq
  VPFP_HD void consts(int seq, double* ca, double* cb) const {
    const long ra = 2 * (long)seq, rb = ra + 1;
    *ca = 0.0; *cb = 0.0;
    if (seq >= a.nseq) return;
    *ca = a.cvec[ra];
    if (MODE == ADV_COLS || rb < a.nrows) *cb = a.cvec[rb];
  }

  VPFP_HD cplx gload(const Tile& t, int seq, int n) const {
    if (seq >= a.nseq) return cmake(0.0, 0.0);
    if (MODE == ADV_COLS) {
      const double* p = a.fin + ((long)t.sim * L + n) * a.ld_in + 2L * seq;
#if defined(__CUDA_ARCH__)
      const double2 v2 = *reinterpret_cast<const double2*>(p);
      return cmake(v2.x, v2.y);
#else
      return cmake(p[0], p[1]);
#endif
    }
    const long ra = 2 * (long)seq, rb = ra + 1;
    if (POISSON) return cmake(1.0 - a.fin[ra * a.ld_in + n], (rb < a.nrows) ? 1.0 - a.fin[rb * a.ld_in + n] : 0.0);
    return cmake(a.fin[ra * a.ld_in + n], (rb < a.nrows) ? a.fin[rb * a.ld_in + n] : 0.0);
  }
  VPFP_HD void gstore(const Tile& t, int seq, int n, cplx v) const {
    if (seq >= a.nseq) return;
    if (MODE == ADV_COLS) {
      double* p = a.fout + ((long)t.sim * L + n) * a.ld_out + 2L * seq;
#if defined(__CUDA_ARCH__)
      *reinterpret_cast<double2*>(p) = make_double2(v.x, v.y);
#else
      p[0] = v.x; p[1] = v.y;
#endif
      return;
    }
    const long ra = 2 * (long)seq, rb = ra + 1;
    if (POISSON && a.addv != nullptr) {            // total field = self-consistent field + driver (field.py:66-88)
      v.x += a.addv[ra * a.ld_out + n];
      if (rb < a.nrows) v.y += a.addv[rb * a.ld_out + n];
    }
    a.fout[ra * a.ld_out + n] = v.x;
    if (rb < a.nrows) a.fout[rb * a.ld_out + n] = v.y;
  }

  // The NEXT tile of the CTA travels global -> shared memory with cp.async while the current one is finished: a thread
  // reads last (inverse stage A) and first (stage A of the next tile) the same 16 slots of the exchange buffer,
  // (k1, r) for its own r, so it copies the raw points n = k1 (L/R1) + r exactly there and needs no barrier for them
  // (the trick of rowfft.cuh and of pass 2 in advect_fast.cuh).
  VPFP_HD void prefetch_own(long tile, int b, int u, unsigned char* smem) const {
    cplx* X = xbuf(smem);
    const Tile t = decode(tile);
    const int seq = t.seq0 + b;
#pragma unroll
    for (int q = 0; q < NQA; ++q) {
      const int rr = u + TPC * q;
#pragma unroll
      for (int n1 = 0; n1 < R1; ++n1) {
        cplx* dst = X + xi(b, n1 * PITCH + rr);
        const int n = n1 * LA + rr;
        if (seq >= a.nseq) { *dst = cmake(0.0, 0.0); continue; }
        if (MODE == ADV_COLS) {
          cp_async(dst, a.fin + ((long)t.sim * L + n) * a.ld_in + 2L * seq, 1);
        } else {
          const long ra = 2 * (long)seq, rb = ra + 1;
          cp_async(&dst->x, a.fin + ra * a.ld_in + n, 0);
          if (rb < a.nrows) cp_async(&dst->y, a.fin + rb * a.ld_in + n, 0);
          else dst->y = 0.0;
        }
      }
    }
    cp_async_commit();
  }

  // one (k, L-k) pair: Zr = Z[kbin], Zpr = Z[L-kbin] in; the phase-multiplied, re-packed pair out
  VPFP_HD static void pair_op(cplx& Zr, cplx& Zpr, const int kbin, const bool selfpair, const int b, const cplx* T0,
                              const cplx* T12) {
    const bool neg = (2 * kbin > L);
    const bool nyq = (2 * kbin == L);
    const int j = neg ? L - kbin : kbin;                 // |signed frequency index|, 0 .. L/2
    cplx Pa = cmul(T0[tix(0, b, j & 15, 16)], T12[tix(0, b, j >> 4, NT12)]);
    cplx Pb = cmul(T0[tix(1, b, j & 15, 16)], T12[tix(1, b, j >> 4, NT12)]);
    if (neg) { Pa = cconj(Pa); Pb = cconj(Pb); }
    if (nyq) { Pa.y = 0.0; Pb.y = 0.0; }
    const cplx Z = Zr, Zp = Zpr;
    const cplx U = cadd(Z, cconj(Zp)), V = csub(Z, cconj(Zp));
    const cplx X1 = cmul(Pa, U), X2 = cmul(Pb, V);
    Zr = cadd(X1, X2);
    if (!selfpair) Zpr = cconj(csub(X1, X2));
  }

  // fused density: the weighted sum of the two columns of this thread for its output slots j0 .. j0 + 7 -> D[j - j0][tid]
  VPFP_HD void dens_park(int j0, int seq, int tid, const cplx* x, unsigned char* smem) const {
    double* D = dbuf(smem);
    const int ncols = 2 * a.nseq;
    const bool valid = seq < a.nseq;
    const double wa = valid ? ((2 * seq == 0 && (a.edge_flags & 1)) ? 0.5 * a.dv : a.dv) : 0.0;
    const double wb = valid ? ((2 * seq + 1 == ncols - 1 && (a.edge_flags & 2)) ? 0.5 * a.dv : a.dv) : 0.0;
#pragma unroll
    for (int j = 0; j < 8; ++j) D[j * DP + tid] = wa * x[j0 + j].x + wb * x[j0 + j].y;
  }

  // Poisson mode: the pair (k, L-k) of the two packed density rows ra, ra + 1 times i m(k) / (2L), m(k) = (ook[k] - ook[L-k]) / 2
  VPFP_HD void pair_poisson(cplx& Zr, cplx& Zpr, const int kbin, const bool selfpair, const int seq) const {
    const bool neg = (2 * kbin > L);
    const int j = neg ? L - kbin : kbin;                 // |signed frequency index|, 0 .. L/2
    const long ra = 2 * (long)seq, rb = ra + 1;
    double ma = 0.0, mb = 0.0;
    if (seq < a.nseq) {
      const double* ka = a.kvec + ra * L;
      ma = 0.5 * (ka[j] - ka[(L - j) % L]);
      if (rb < a.nrows) mb = 0.5 * (ka[L + j] - ka[L + (L - j) % L]);
    }
    const double sc = (neg ? -0.5 : 0.5) / (double)L;
    const cplx Pa = cmake(0.0, ma * sc), Pb = cmake(0.0, mb * sc);
    const cplx Z = Zr, Zp = Zpr;
    const cplx U = cadd(Z, cconj(Zp)), V = csub(Z, cconj(Zp));
    const cplx X1 = cmul(Pa, U), X2 = cmul(Pb, V);
    Zr = cadd(X1, X2);
    if (!selfpair) Zpr = cconj(csub(X1, X2));
  }

  // prefetched: this tile was brought into the thread's own slots by prefetch_own; nexttile: the tile this CTA handles
  // after this one (< 0: none)
  VPFP_HD void phase(int ph, long tile, long nexttile, bool prefetched, int tid, Regs& r, unsigned char* smem) const {
    cplx* X = xbuf(smem);
    const cplx* TWA = twa(smem);
    const cplx* TWB = twb(smem);
    cplx* x = r.x;
    int b, u;
    roles(tid, &b, &u);
    const Tile t = decode(tile);
    const int seq = t.seq0 + b;
    switch (ph) {
      case 0: {
        // ---- phase tables of this tile: T0[j] = exp(-i phi j) / (2L), j < 16; T12[i] = exp(-i phi 16 i)
        const double* K = a.kvec + ((MODE == ADV_COLS) ? (long)t.sim * L : 0);
        const double kdt = POISSON ? 0.0 : mul_rn(K[1], a.dt);
        cplx* T0 = t0(smem);
        cplx* T12 = t12(smem);
        constexpr int PER = 16 + NT12;
        for (int w = tid; w < (POISSON ? 0 : 2 * CB * PER); w += NT) {
          const int i = w % PER, cb_ = (w / PER) % CB, ch = w / (PER * CB);
          double ca, cbv;
          consts(t.seq0 + cb_, &ca, &cbv);
          const double phi = mul_rn(kdt, ch ? cbv : ca);
          const double j = (i < 16) ? (double)i : 16.0 * (double)(i - 16);
          double sn, cs;
          sincos_hd(phi * j, &sn, &cs);
          if (i < 16) T0[tix(ch, cb_, i, 16)] = cmake(cs * (0.5 / (double)L), -sn * (0.5 / (double)L));
          else T12[tix(ch, cb_, i - 16, NT12)] = cmake(cs, -sn);
        }
        // ---- stage A: radix-R1 over n1 for r = u + TPC q
        if (PREFETCH && prefetched) {
          cp_async_wait();
#pragma unroll
          for (int q = 0; q < NQA; ++q) {
            const int rr = u + TPC * q;
#pragma unroll
            for (int n1 = 0; n1 < R1; ++n1) x[q * R1 + n1] = X[xi(b, n1 * PITCH + rr)];
          }
        } else {
#pragma unroll
          for (int q = 0; q < NQA; ++q) {
            const int rr = u + TPC * q;
#pragma unroll
            for (int n1 = 0; n1 < R1; ++n1) x[q * R1 + n1] = gload(t, seq, n1 * LA + rr);
          }
        }
#pragma unroll
        for (int q = 0; q < NQA; ++q) {
          const int rr = u + TPC * q;
          fftR<R1, -1>(x + q * R1);
#pragma unroll
          for (int k1 = 1; k1 < R1; ++k1) x[q * R1 + k1] = cmul(x[q * R1 + k1], TWA[k1 * LA + rr]);
#pragma unroll
          for (int k1 = 0; k1 < R1; ++k1) X[xi(b, k1 * PITCH + rr)] = x[q * R1 + k1];
        }
      } break;
      case 1: {
        // ---- stage B, in place: radix-R2 over n2 for (k1, n3) = c = u + TPC q
#pragma unroll
        for (int q = 0; q < NQB; ++q) {
          const int c = u + TPC * q, k1 = c >> 3, n3 = c & 7;
#pragma unroll
          for (int n2 = 0; n2 < R2; ++n2) x[q * R2 + n2] = X[xi(b, k1 * PITCH + n2 * 8 + n3)];
          fftR<R2, -1>(x + q * R2);
#pragma unroll
          for (int k2 = 1; k2 < R2; ++k2) x[q * R2 + k2] = cmul(x[q * R2 + k2], TWB[k2 * 8 + n3]);
#pragma unroll
          for (int k2 = 0; k2 < R2; ++k2) X[xi(b, k1 * PITCH + k2 * 8 + n3)] = x[q * R2 + k2];
        }
      } break;
      case 2: {
        // ---- stage C for the sub-transforms sA, sB of this thread, pointwise on the pairs, inverse stage C
        const bool special = (u == 0);
        const int sA = special ? 0 : u, sB = special ? S / 2 : S - u;
        const int slotA = (sA % R1) * PITCH + (sA / R1) * 8, slotB = (sB % R1) * PITCH + (sB / R1) * 8;
#pragma unroll
        for (int n3 = 0; n3 < 8; ++n3) {
          x[n3] = X[xi(b, slotA + n3)];
          x[8 + n3] = X[xi(b, slotB + n3)];
        }
        fft8<-1>(x);
        fft8<-1>(x + 8);
        const cplx* T0 = t0(smem);
        const cplx* T12 = t12(smem);
        if (POISSON) {
          if (!special) {
#pragma unroll
            for (int pr = 0; pr < 8; ++pr) pair_poisson(x[pr], x[15 - pr], sA + S * pr, false, seq);
          } else {
            pair_poisson(x[0], x[0], 0, true, seq);                        // DC: multiplier 0
            pair_poisson(x[4], x[4], S * 4, true, seq);                    // Nyquist: multiplier 0
#pragma unroll
            for (int pr = 1; pr < 4; ++pr) pair_poisson(x[pr], x[8 - pr], S * pr, false, seq);
#pragma unroll
            for (int pr = 0; pr < 4; ++pr) pair_poisson(x[8 + pr], x[15 - pr], S / 2 + S * pr, false, seq);
          }
        } else if (!special) {
          // A[k3] (bin sA + S k3) pairs with B[7 - k3]
#pragma unroll
          for (int pr = 0; pr < 8; ++pr) pair_op(x[pr], x[15 - pr], sA + S * pr, false, b, T0, T12);
        } else {
          pair_op(x[0], x[0], 0, true, b, T0, T12);                       // DC
          pair_op(x[4], x[4], S * 4, true, b, T0, T12);                   // Nyquist
#pragma unroll
          for (int pr = 1; pr < 4; ++pr) pair_op(x[pr], x[8 - pr], S * pr, false, b, T0, T12);
#pragma unroll
          for (int pr = 0; pr < 4; ++pr) pair_op(x[8 + pr], x[15 - pr], S / 2 + S * pr, false, b, T0, T12);
        }
        fft8<1>(x);
        fft8<1>(x + 8);
#pragma unroll
        for (int n3 = 0; n3 < 8; ++n3) {
          X[xi(b, slotA + n3)] = x[n3];
          X[xi(b, slotB + n3)] = x[8 + n3];
        }
      } break;
      case 3: {
        // ---- inverse stage B, in place
#pragma unroll
        for (int q = 0; q < NQB; ++q) {
          const int c = u + TPC * q, k1 = c >> 3, n3 = c & 7;
#pragma unroll
          for (int k2 = 0; k2 < R2; ++k2) {
            cplx val = X[xi(b, k1 * PITCH + k2 * 8 + n3)];
            if (k2 > 0) val = cmulc(val, TWB[k2 * 8 + n3]);
            x[q * R2 + k2] = val;
          }
          fftR<R2, 1>(x + q * R2);
#pragma unroll
          for (int n2 = 0; n2 < R2; ++n2) X[xi(b, k1 * PITCH + n2 * 8 + n3)] = x[q * R2 + n2];
        }
      } break;
      case 5:
      case 7: {
        // ---- fused density: thread (b, u) adds, in a fixed order, the CB column pairs of the slots j = b, b + CB, ...
        // of this round (slots 0..7 in phase 5, 8..15 in phase 7)
        if (!DENS) break;
        const double* D = dbuf(smem);
        const long tb = (a.nseq + CB - 1) / CB;
        double* dst = a.dens_partial + ((tile % tb) * a.nsim + t.sim) * (long)L;
        const int round = (ph == 7);
#pragma unroll
        for (int jg = 0; jg < 8 / CB; ++jg) {
          const int jl = b + CB * jg, j = jl + 8 * round, q = j / R1, n1 = j % R1;
          double sum = 0.0;
#pragma unroll
          for (int bb = 0; bb < CB; ++bb) sum += D[jl * DP + u * CB + bb];
          dst[n1 * LA + u + TPC * q] = sum;
        }
      } break;
      case 6: {
        if (DENS) dens_park(8, seq, tid, x, smem);
      } break;
      default: {
        // ---- inverse stage A, store; the thread's slots are free once they are in registers: next tile into them
#pragma unroll
        for (int q = 0; q < NQA; ++q) {
          const int rr = u + TPC * q;
#pragma unroll
          for (int k1 = 0; k1 < R1; ++k1) x[q * R1 + k1] = X[xi(b, k1 * PITCH + rr)];
        }
        if (PREFETCH && nexttile >= 0) prefetch_own(nexttile, b, u, smem);
#pragma unroll
        for (int q = 0; q < NQA; ++q) {
          const int rr = u + TPC * q;
#pragma unroll
          for (int k1 = 1; k1 < R1; ++k1) x[q * R1 + k1] = cmulc(x[q * R1 + k1], TWA[k1 * LA + rr]);
          fftR<R1, 1>(x + q * R1);
#pragma unroll
          for (int n1 = 0; n1 < R1; ++n1) gstore(t, seq, n1 * LA + rr, x[q * R1 + n1]);
        }
        if (DENS) dens_park(0, seq, tid, x, smem);
      } break;
    }
  }
};

#if defined(__CUDACC__)
template <class P>
__global__ void __launch_bounds__(P::NT, (512 / P::NT > 0 ? 512 / P::NT : 1)) midfft_kernel(const P prog) {   // <= 128 registers
  extern __shared__ __align__(16) unsigned char smem_raw[];
  typename P::Regs r;
  const int tid = (int)threadIdx.x;
  prog.init(tid, smem_raw);
  __syncthreads();
  const long nt = prog.ntiles();
  bool prefetched = false;
  for (long tile = blockIdx.x; tile < nt; tile += gridDim.x) {
    const long nxt = (tile + gridDim.x < nt) ? tile + gridDim.x : -1;
#pragma unroll
    for (int ph = 0; ph < P::NPH; ++ph) {
      prog.phase(ph, tile, nxt, prefetched, tid, r, smem_raw);
      __syncthreads();
    }
    prefetched = true;
  }
}
#endif

}  // namespace midfft
