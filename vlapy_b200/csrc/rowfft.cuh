// rowfft.cuh -- single-pass e df/dv: one CTA transforms one whole x-row on chip, so the operator
// costs ONE HBM read and ONE write of f (16 B/cell).  Replaces vlapy/core/vlasov.py:123-138
//     f_new[x, :] = Re ifft_v( exp(-i kv dt e[x]) fft_v f[x, :] )
// for nv = N in {4096, 8192, 16384} on uniform fftfreq wavenumber grids (VPFP_PHASE_TABLE).
//
// The row is real, so it is transformed as ONE complex sequence of M = N/2 points
//     z[m] = f[2m] + i f[2m+1]
// (a 16-byte load is one point), and the spectrum of the real row is recovered pairwise from
// Z[k], Z[M-k] (E = (Z[k] + conj Z[M-k])/2, O = (Z[k] - conj Z[M-k])/(2i), X[k] = E + W_N^k O).
// The phase multiply and the packing for the inverse transform are done on the same pair, so
// between the forward and the inverse transform the data never leaves registers:
//     A = E + W^k O = X[k],  B = E - W^k O = conj X[M-k],
//     Yk = P_k A,  Q = conj(P_{M-k}) B                       (Q = Y[k+M], Hermitian output)
//     Z'[k] = (Yk + Q)/2 + i conj(W^k) (Yk - Q)/2,   Z'[M-k] = conj( (Yk + Q)/2 - i conj(W^k)(Yk - Q)/2 )
// np.real of the reference drops the imaginary part that the Nyquist factor would produce:
// P_M is replaced by its real part (bin k = 0 carries X[0] and X[M], both real).
//
// Decomposition M = R1 * R2 * 16, every thread owns V = 32 complex values in registers:
//   stage 1: radix-R1 over m1 (m = m1*L2 + r, r = 0..L2-1, L2 = R2*16), twiddle W_M^(r k1)
//   stage 2: radix-R2 over m2 (r = m2*16 + m3),                          twiddle W_L2^(m3 k2)
//   stage 3: radix-16 over m3 for sub-transform s = k1 + R1 k2; a thread owns s and its partner
//            S - s (S = R1 R2), i.e. BOTH members of every pair (k, M-k): bin k = s + S k3 pairs
//            with (S - s) + S (15 - k3).
// Two shared-memory exchanges on the way in and two on the way out; T = M/32 threads per row.
// Stage-1 twiddles are powers of one per-thread constant (generated in registers), stage-2
// twiddles a 256-entry table in shared memory, phase factors P_k = exp(-i phi k) come from three
// small geometric tables per row (phi = (K[1] dt) e[x]):  P(s + S k3) = Lo[s & 31] Hi[s >> 5] G[k3].
//
// The kernel body is a phase program (vpfp_common.h) whose per-thread registers persist across
// barriers; tests/emul runs the same source thread by thread on the host.
#pragma once
#include "butterflies.h"
#include "vpfp_common.h"

namespace rowfft {

#ifndef ROWFFT_SPECIAL_BRANCH
#define ROWFFT_SPECIAL_BRANCH 0   // measured slower (1.607 ms against 1.551 ms, profiles/ab_r01_v15_rowfft_special_branch.txt)
#endif

using fast::fft16;
using fast::fft8;

struct Args {
  const double* fin; long ld_in;
  double* fout; long ld_out;
  const double* kvec;   // [N] wavenumbers (uniform fftfreq grid: only K[1] is used)
  const double* cvec;   // [nrows] e[x]
  double dt;
  int nrows;
  const cplx* twN;      // exp(-2 pi i m / N), N entries
  // optional scatter of the result to peer GPUs (multi-GPU layout change fused into the store):
  // element (row, n) goes to the v-shard of rank q = n / part at [my_rank*nrows + row][n % part]
  int peer_mode, nparts, my_rank, lpart;   // part = 1 << lpart
  double* peer[8];
  const double* addv;   // POISSON: driver rows [nrows][N] added to the result (nullable)
};

// 16-byte asynchronous copy global -> shared (host emulation: plain copy)
VPFP_HD void cp_async16(void* smem, const void* gmem) {
#if defined(__CUDA_ARCH__)
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gmem) : "memory");
#else
  memcpy(smem, gmem, 16);
#endif
}
VPFP_HD void cp_async_commit_wait(bool wait) {
#if defined(__CUDA_ARCH__)
  if (wait) asm volatile("cp.async.wait_group 0;\n" ::: "memory");
  else asm volatile("cp.async.commit_group;\n" ::: "memory");
#else
  (void)wait;
#endif
}

// cos(2 pi j / 32), j = 0..16 (no recursion: folds to an immediate once the caller's loop is unrolled)
VPFP_HD constexpr double cos32(int j) {
  return j == 0 ? 1.0 : j == 1 ? 0.98078528040323044913 : j == 2 ? 0.92387953251128675613
       : j == 3 ? 0.83146961230254523708 : j == 4 ? 0.70710678118654752440 : j == 5 ? 0.55557023301960222474
       : j == 6 ? 0.38268343236508977173 : j == 7 ? 0.19509032201612826785 : j == 8 ? 0.0
       : j == 9 ? -0.19509032201612826785 : j == 10 ? -0.38268343236508977173 : j == 11 ? -0.55557023301960222474
       : j == 12 ? -0.70710678118654752440 : j == 13 ? -0.83146961230254523708 : j == 14 ? -0.92387953251128675613
       : j == 15 ? -0.98078528040323044913 : -1.0;
}
VPFP_HD constexpr double sin32(int j) { return j <= 8 ? cos32(8 - j) : cos32(j - 8); }   // sin(2 pi j/32), j = 0..16

template <int DIR>
VPFP_HD void fft32(cplx* x) {
  // radix-2 step then two radix-16 transforms (DIF); W32^j constants
  cplx a[16], b[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    a[j] = cadd(x[j], x[j + 16]);
    const cplx d = csub(x[j], x[j + 16]);
    if (j == 0) b[j] = d;
    else if (j == 8) b[j] = fast::rot_i<DIR>(d);
    else b[j] = fast::mul_w16<DIR>(d, cos32(j), sin32(j));
  }
  fft16<DIR>(a);
  fft16<DIR>(b);
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    x[2 * k] = a[k];
    x[2 * k + 1] = b[k];
  }
}

template <int R, int DIR>
VPFP_HD void fftR(cplx* x) {
  if (R == 32) fft32<DIR>(x);
  else if (R == 16) fft16<DIR>(x);
  else fft8<DIR>(x);
}

// one (k, M-k) pair: Z = Z[k], Zp = Z[M-k] in; Z'[k], Z'[M-k] out.  Wk = W_N^k, Pk / Pmk the phase
// factors of bins k and M-k, both pre-scaled by 1/(4M).
VPFP_HD void pair_op(cplx& Z, cplx& Zp, const cplx Wk, const cplx Pk, const cplx Pmk) {
  const cplx U = cadd(Z, cconj(Zp)), Vv = csub(Z, cconj(Zp));
  const cplx Tw = cmul_mi(cmul(Wk, Vv));           // -i W^k (Z - conj Zp) = 2 W^k O
  const cplx A = cadd(U, Tw), B = csub(U, Tw);     // 2 X[k], 2 conj X[M-k]
  const cplx Yk = cmul(Pk, A);
  const cplx Q = cmulc(B, Pmk);
  const cplx Ye = cadd(Yk, Q);
  const cplx D = cmul_i(cmulc(csub(Yk, Q), Wk));
  Z = cadd(Ye, D);
  Zp = cconj(csub(Ye, D));
}

// PEER_: the result is scattered to peer GPUs (Args::peer_mode) -- a template parameter, so that the
// single-GPU kernel does not carry the registers and code of the scatter path.
// POISSON_: the same transform pair as the spectral Poisson solve of vlapy/core/field.py:39-63 on a row of densities,
//   E = Re ifft( i one_over_kx fft(1 - n) ) + driver:  the row is loaded as 1 - n, the multiplier of bin k is
//   i ook[k] (kvec = one_over_kx; real and antisymmetric, so the result is real; bins 0 and N/2 contribute nothing),
//   the driver row is added on the store.  One CTA, one launch, for nx in {4096, 8192, 16384} (the generic path took
//   three dependent launches of ~20 us).
template <int R1_, int R2_, bool PEER_ = false, bool POISSON_ = false>
struct Prog {
  static constexpr int R1 = R1_, R2 = R2_, V = 32;
  static constexpr int S = R1 * R2;        // stage-3 sub-transforms
  static constexpr int T = S / 2;          // threads per row
  static constexpr int M = S * 16;         // complex points per row
  static constexpr int N = 2 * M;          // real cells per row
  static constexpr int L2 = R2 * 16;
  static constexpr int NQ1 = V / R1, NQ2 = V / R2;
  static constexpr int NHI = S / 32;
  static constexpr int NPH = 8;
  // Barriers that are NOT needed after a phase: the pointwise phase (3) reads the rows sA, sB of the exchange buffer
  // and phase 4 writes the same two rows; the last phase (7) reads, and the first phase of the next row reads and
  // writes, only the thread's own prefetch slots, and the row tables written at the end of phase 7 are not read
  // before phase 3.  Without the barrier after phase 7 the warps that are done storing start on the next row while
  // the others still drain their stores (the wait at that barrier was 5 % of the kernel in the ncu source view).
  VPFP_HD static constexpr bool sync_after(int ph) { return ph != 3 && ph != 7; }
  // shared memory: exchange buffer (two layouts), stage-2 twiddles, per-row phase tables
  static constexpr int X_ELEMS = (S * 17 > M) ? S * 17 : M;
  static constexpr int NTAB = 16 + 32 + NHI;
  static constexpr long SMEM_BYTES = (long)sizeof(cplx) * (X_ELEMS + L2 + NTAB) + 16;

  struct Regs {
    cplx x[V];
    cplx w1[NQ1];     // W_M^r for the stage-1 butterflies of this thread
    cplx w4[NQ1];     // (W_M^r)^4
    cplx wA, wB;      // W_N^sA, W_N^sB
  };

  Args a;

  VPFP_HD static cplx* xbuf(unsigned char* smem) { return reinterpret_cast<cplx*>(smem); }
  VPFP_HD static cplx* tw2(unsigned char* smem) { return xbuf(smem) + X_ELEMS; }
  VPFP_HD static cplx* tabs(unsigned char* smem) { return tw2(smem) + L2; }   // G[16], Lo[32], Hi[NHI]
  VPFP_HD static double* cosM(unsigned char* smem) { return reinterpret_cast<double*>(tabs(smem) + NTAB); }

  // once per CTA: per-thread constants and the stage-2 twiddle table
  VPFP_HD void init(int tid, Regs& r, unsigned char* smem) const {
#pragma unroll
    for (int q = 0; q < NQ1; ++q) {
      const cplx w = a.twN[2 * (tid + T * q)];
      const cplx w2 = cmul(w, w);
      r.w1[q] = w;
      r.w4[q] = cmul(w2, w2);
    }
    const int sA = (tid == 0) ? 0 : tid, sB = (tid == 0) ? T : S - tid;
    r.wA = a.twN[sA];
    r.wB = a.twN[sB];
    cplx* TW2 = tw2(smem);
    // stage-2 twiddles W_L2^(m3 k2) laid out [k2][m3]: the lanes of a quarter-warp read consecutive m3 of one k2
    // (a single table indexed m3 * k2 cost up to 8-way bank conflicts for even k2: ~10 % of the kernel's wavefronts)
    for (int j = tid; j < L2; j += T) TW2[j] = a.twN[(long)((j & 15) * (j >> 4)) * (N / L2)];
  }

  // sin / cos of pi*t without a slow path (exact reduction mod 2)
  VPFP_HD static void sincospi_hd(double t, double* sn, double* cs) {
#if defined(__CUDA_ARCH__)
    sincospi(t, sn, cs);
#else
    const double two = 2.0;
    double fr = t - two * floor(t / two);      // exact for |t| < 2^52
    *sn = sin(3.14159265358979323846 * fr);
    *cs = cos(3.14159265358979323846 * fr);
#endif
  }

  // The next row travels global -> shared memory with cp.async while the current row is finished:
  // thread tid copies the complex points m = tid + T k, k = 0..V-1, to X[m].  These are exactly the
  // points the thread itself reads last from X in the final phase (inverse stage 1) and first in the
  // next row's first phase (stage 1: m = m1 L2 + tid + T q), so the copy needs no barrier on either
  // side, only the thread's own cp.async.wait_group.
  static_assert(L2 % T == 0, "a thread's stage-1 points are the points it prefetches");
  VPFP_HD void prefetch_copy(long row, int tid, unsigned char* smem) const {
    cplx* X = xbuf(smem);
    const double* src = a.fin + row * a.ld_in;
#pragma unroll
    for (int k = 0; k < V; ++k) cp_async16(X + tid + T * k, src + 2L * (tid + T * k));
    cp_async_commit_wait(false);
  }
  // phase slope of a row over pi: (K[1] dt e[row]) / pi, the reference's two roundings of (k dt) e
  VPFP_HD double phi_over_pi(long row) const {
    if (POISSON_) return 0.0;
    return mul_rn(mul_rn(a.kvec[1], a.dt), a.cvec[row]) * 0.31830988618379067154;
  }
  // first row of a CTA: copy and phase tables.  For the following rows the last phase of the previous row
  // issues the copy and prepares the tables itself (they are read only by the pointwise phase, which every
  // thread has left by then), with the loads of K[1] and e[row] issued at the top of that phase: they used to
  // stall the first phase of every row (4 % of the kernel in the ncu source view).
  VPFP_HD void prefetch_row(long row, int tid, unsigned char* smem) const {
    prefetch_copy(row, tid, smem);
    row_tables(tid, phi_over_pi(row), smem);
  }

  // phase tables of the row in registers: G[j] = exp(-i phi S j), Lo[j] = exp(-i phi j),
  // Hi[j] = exp(-i phi 32 j)/(4M), cos(phi M); phi_pi = (K[1] dt e[row]) / pi.  Entry w is computed by lane
  // w / NWARP of warp w % NWARP (every warp pays for one sincos).  Written one phase before the row starts
  // (prefetch_row), read in its fourth phase.
  VPFP_HD void row_tables(int tid, const double phi_pi, unsigned char* smem) const {
    if (POISSON_) return;
    constexpr int NWARP = (T >= 32) ? T / 32 : 1;
    cplx* G = tabs(smem);
    const int w = (tid & 31) * NWARP + (tid >> 5);
    if ((tid & 31) * NWARP < NTAB + 1 && w < NTAB + 1) {
      double k, sc = 1.0;
      if (w < 16) k = (double)(S * w);
      else if (w < 48) k = (double)(w - 16);
      else if (w < NTAB) { k = (double)(32 * (w - 48)); sc = 0.25 / (double)M; }
      else k = (double)M;
      double sn, cs;
      sincospi_hd(phi_pi * k, &sn, &cs);
      if (w < NTAB) G[w] = cmake(cs * sc, -sn * sc);
      else *cosM(smem) = cs;
    }
  }

  // x[k] *= w^k (CONJ: conj(w)^k), k = 0..R1-1, w = r.w1[q]; powers as w^(4a) w^b
  template <bool CONJ>
  VPFP_HD static void twiddle1(cplx* x, const cplx w, const cplx w4) {
    const cplx b1 = w, b2 = cmul(w, w), b3 = cmul(b2, w);
    cplx A = cmake(1.0, 0.0);
#pragma unroll
    for (int g = 0; g < R1 / 4; ++g) {
      if (g == 0) {
        x[1] = CONJ ? cmulc(x[1], b1) : cmul(x[1], b1);
        x[2] = CONJ ? cmulc(x[2], b2) : cmul(x[2], b2);
        x[3] = CONJ ? cmulc(x[3], b3) : cmul(x[3], b3);
        A = w4;
      } else {
        const cplx t1 = cmul(A, b1), t2 = cmul(A, b2), t3 = cmul(A, b3);
        x[4 * g] = CONJ ? cmulc(x[4 * g], A) : cmul(x[4 * g], A);
        x[4 * g + 1] = CONJ ? cmulc(x[4 * g + 1], t1) : cmul(x[4 * g + 1], t1);
        x[4 * g + 2] = CONJ ? cmulc(x[4 * g + 2], t2) : cmul(x[4 * g + 2], t2);
        x[4 * g + 3] = CONJ ? cmulc(x[4 * g + 3], t3) : cmul(x[4 * g + 3], t3);
        if (g + 1 < R1 / 4) A = cmul(A, w4);
      }
    }
  }

  VPFP_HD void store_pair(long row, int m, cplx val) const {
    const long n = 2L * m;
    if (a.peer_mode) {
      const int q = (int)(n >> a.lpart);
      const long part = 1L << a.lpart;
      *reinterpret_cast<cplx*>(a.peer[q] + ((long)a.my_rank * a.nrows + row) * part + (n & (part - 1))) = val;
      return;
    }
    *reinterpret_cast<cplx*>(a.fout + row * a.ld_out + n) = val;
  }

  // stage 3 for the sub-transforms sA, sB of the thread, pointwise step on the pairs, inverse stage 3.
  // MS = false: the caller knows that this thread is not thread 0 (ROWFFT_SPECIAL_BRANCH: a warp-uniform branch
  // around the select-based special case of thread 0, so that warps 1.. do not carry its selects).
  template <bool MS>
  VPFP_HD void pointwise_phase(int tid, long row, Regs& r, unsigned char* smem) const {
    cplx* X = xbuf(smem);
    cplx* G = tabs(smem);
    cplx* LO = G + 16;
    cplx* HI = LO + 32;
    cplx* x = r.x;
    {
        // ---- stage 3 for sub-transforms sA, sB; pointwise on pairs; inverse stage 3
        const bool special = MS && (tid == 0);
        const int sA = special ? 0 : tid, sB = special ? T : S - tid;
#pragma unroll
        for (int m3 = 0; m3 < 16; ++m3) {
          x[m3] = X[sA * 17 + m3];
          x[16 + m3] = X[sB * 17 + m3];
        }
        fft16<-1>(x);
        fft16<-1>(x + 16);
        const cplx bA = POISSON_ ? cmake(0.0, 0.0) : cmul(LO[sA & 31], HI[sA >> 5]);
        const cplx bB = POISSON_ ? cmake(0.0, 0.0) : cmul(LO[sB & 31], HI[sB >> 5]);
        const double* ook = a.kvec + (POISSON_ ? row * (long)N : 0);
        // Pair slots j = 0..15 hold (x[j], x[31-j]).  Ordinary threads: bin sA + S j with its partner
        // sB + S (15-j).  Thread 0 owns the two self-paired sub-transforms s = 0 and s = S/2; its
        // registers are permuted (selects, no divergent copy of the pair code) so that the same slots
        // hold   j = 0: (A[8], A[8])   1..7: (A[j], A[16-j])   8..15: (B[j-8], B[23-j]),
        // and bin 0 (X[0] and X[M], both real) is finished separately.
        cplx dc = x[0];
        if (special) {
          cplx y[V];
          y[0] = x[8]; y[31] = x[8];
#pragma unroll
          for (int i = 1; i < 8; ++i) y[i] = x[i];
#pragma unroll
          for (int i = 8; i < 16; ++i) y[i] = x[i + 8];
#pragma unroll
          for (int i = 16; i < 24; ++i) y[i] = x[i + 8];
#pragma unroll
          for (int i = 24; i < 31; ++i) y[i] = x[i - 15];
#pragma unroll
          for (int i = 0; i < V; ++i) x[i] = y[i];
        }
        const cplx wHi = special ? r.wB : r.wA;      // W_N^s of slots 8..15
        const cplx pHi = special ? bB : bA;          // base of P_k for slots 8..15
        const cplx qLo = special ? bA : bB;          // base of P_{M-k} for slots 0..7
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int jw = (j < 8) ? j : j - 8;                                       // thread 0: bin index inside its set
          const int gk = special ? (j == 0 ? 8 : jw) : j;                             // P_k = base G^gk
          const int gm = special ? (j == 0 ? 8 : (j < 8 ? 16 - j : 23 - j)) : 15 - j;  // P_{M-k} = base G^gm
          cplx w32 = (j < 8) ? cmake(cos32(j), -sin32(j))
                             : cmake(special ? cos32(jw) : cos32(j), special ? -sin32(jw) : -sin32(j));
          if (j == 0 && special) w32 = cmake(0.0, -1.0);                             // W_N^(M/2) = -i
          const cplx Wk = cmul((j < 8) ? r.wA : wHi, w32);
          cplx Pk, Pm;
          if (POISSON_) {
            // multipliers i ook[k] / (4M) of bin k and of its partner M - k (both in the lower half of the spectrum)
            const int sK = (j < 8) ? sA : (special ? sB : sA), sM = (j < 8) ? (special ? sA : sB) : sB;
            // (np.real of the reference keeps the Hermitian part of the multiplier: (ook[k] - ook[N-k]) / 2, which is
            // ook[k] itself for the antisymmetric one_over_kx of vlapy/initializers.py:85-87)
            const int kk = sK + S * gk, km = sM + S * gm;
            Pk = cmake(0.0, (ook[kk] - ook[N - kk]) * (0.125 / (double)M));
            Pm = cmake(0.0, (ook[km] - ook[N - km]) * (0.125 / (double)M));
          } else {
            Pk = cmul((j < 8) ? bA : pHi, G[gk]);
            Pm = cmul((j < 8) ? qLo : bB, G[gm]);
          }
          pair_op(x[j], x[31 - j], Wk, Pk, Pm);
        }
        if (special) {
          cplx y[V];
          {  // bin 0: Y[0] = X[0], Y[M] = Re(P_M) X[M]
            const double sc = 0.5 / (double)M;
            const double y0 = (dc.x + dc.y) * sc, ym = (dc.x - dc.y) * (*cosM(smem)) * sc;
            y[0] = POISSON_ ? cmake(0.0, 0.0) : cmake(y0 + ym, y0 - ym);   // Poisson: ook[0] = 0, Nyquist imaginary
          }
          y[8] = x[0];
#pragma unroll
          for (int i = 1; i < 8; ++i) y[i] = x[i];
#pragma unroll
          for (int i = 8; i < 16; ++i) y[i + 8] = x[i];
#pragma unroll
          for (int i = 16; i < 24; ++i) y[i + 8] = x[i];
#pragma unroll
          for (int i = 24; i < 31; ++i) y[i - 15] = x[i];
#pragma unroll
          for (int i = 0; i < V; ++i) x[i] = y[i];
        }
        fft16<1>(x);
        fft16<1>(x + 16);
    }
  }

  // nextrow: the row this CTA handles after `row` (< 0: none); its tables are prepared in the last phase
  VPFP_HD void phase(int ph, long row, long nextrow, int tid, Regs& r, unsigned char* smem) const {
    cplx* X = xbuf(smem);
    cplx* TW2 = tw2(smem);
    cplx* x = r.x;
    switch (ph) {
      case 0: {
        // ---- stage 1 on the row that prefetch_row brought into X (its phase tables are in place as well)
        cp_async_commit_wait(true);
#pragma unroll
        for (int q = 0; q < NQ1; ++q) {
          const int rr = tid + T * q;
#pragma unroll
          for (int m1 = 0; m1 < R1; ++m1) x[q * R1 + m1] = X[m1 * L2 + rr];
          if (POISSON_) {                                   // net charge 1 - n (vlapy/core/field.py:61-63)
#pragma unroll
            for (int m1 = 0; m1 < R1; ++m1) x[q * R1 + m1] = cmake(1.0 - x[q * R1 + m1].x, 1.0 - x[q * R1 + m1].y);
          }
          fftR<R1, -1>(x + q * R1);
          twiddle1<false>(x + q * R1, r.w1[q], r.w4[q]);
#pragma unroll
          for (int k1 = 0; k1 < R1; ++k1) X[k1 * L2 + rr] = x[q * R1 + k1];
        }
      } break;
      case 1: {
        // ---- stage 2: (k1, m3) butterflies over m2
#pragma unroll
        for (int q = 0; q < NQ2; ++q) {
          const int c = tid + T * q, k1 = c >> 4, m3 = c & 15;
#pragma unroll
          for (int m2 = 0; m2 < R2; ++m2) x[q * R2 + m2] = X[k1 * L2 + m2 * 16 + m3];
          fftR<R2, -1>(x + q * R2);
#pragma unroll
          for (int k2 = 1; k2 < R2; ++k2) x[q * R2 + k2] = cmul(x[q * R2 + k2], TW2[k2 * 16 + m3]);
        }
      } break;
      case 2: {
#pragma unroll
        for (int q = 0; q < NQ2; ++q) {
          const int c = tid + T * q, k1 = c >> 4, m3 = c & 15;
#pragma unroll
          for (int k2 = 0; k2 < R2; ++k2) X[(k1 + R1 * k2) * 17 + m3] = x[q * R2 + k2];
        }
      } break;
      case 3: {
#if ROWFFT_SPECIAL_BRANCH
        if (tid < 32) pointwise_phase<true>(tid, row, r, smem);
        else pointwise_phase<false>(tid, row, r, smem);
#else
        pointwise_phase<true>(tid, row, r, smem);
#endif
      } break;
      case 4: {
        const bool special = (tid == 0);
        const int sA = special ? 0 : tid, sB = special ? T : S - tid;
#pragma unroll
        for (int m3 = 0; m3 < 16; ++m3) {
          X[sA * 17 + m3] = x[m3];
          X[sB * 17 + m3] = x[16 + m3];
        }
      } break;
      case 5: {
        // ---- inverse stage 2
#pragma unroll
        for (int q = 0; q < NQ2; ++q) {
          const int c = tid + T * q, k1 = c >> 4, m3 = c & 15;
#pragma unroll
          for (int k2 = 0; k2 < R2; ++k2) {
            cplx val = X[(k1 + R1 * k2) * 17 + m3];
            if (k2 > 0) val = cmulc(val, TW2[k2 * 16 + m3]);
            x[q * R2 + k2] = val;
          }
          fftR<R2, 1>(x + q * R2);
        }
      } break;
      case 6: {
#pragma unroll
        for (int q = 0; q < NQ2; ++q) {
          const int c = tid + T * q, k1 = c >> 4, m3 = c & 15;
#pragma unroll
          for (int m2 = 0; m2 < R2; ++m2) X[k1 * L2 + m2 * 16 + m3] = x[q * R2 + m2];
        }
      } break;
      default: {
        // ---- inverse stage 1, store; the thread's part of X is free once it is in registers
        const double phi_next = (nextrow >= 0) ? phi_over_pi(nextrow) : 0.0;   // loads issued here, used at the end
#pragma unroll
        for (int q = 0; q < NQ1; ++q) {
          const int rr = tid + T * q;
#pragma unroll
          for (int k1 = 0; k1 < R1; ++k1) x[q * R1 + k1] = X[k1 * L2 + rr];
        }
        if (nextrow >= 0) prefetch_copy(nextrow, tid, smem);
#pragma unroll
        for (int q = 0; q < NQ1; ++q) {
          const int rr = tid + T * q;
          twiddle1<true>(x + q * R1, r.w1[q], r.w4[q]);
          fftR<R1, 1>(x + q * R1);
          if (POISSON_ && a.addv != nullptr) {               // total field = self-consistent field + driver (field.py:66-88)
            const cplx* drv = reinterpret_cast<const cplx*>(a.addv + row * (long)N) + rr;
#pragma unroll
            for (int m1 = 0; m1 < R1; ++m1) x[q * R1 + m1] = cadd(x[q * R1 + m1], drv[m1 * L2]);
          }
          if (!PEER_) {
            // one base address, compile-time offsets: a store must not wait for the address registers
            // of the previous one (they are held until the load/store unit has taken the store)
            cplx* dst = reinterpret_cast<cplx*>(a.fout + row * a.ld_out) + rr;
#pragma unroll
            for (int m1 = 0; m1 < R1; ++m1) dst[m1 * L2] = x[q * R1 + m1];
          } else {
#pragma unroll
            for (int m1 = 0; m1 < R1; ++m1) store_pair(row, m1 * L2 + rr, x[q * R1 + m1]);
          }
        }
        if (nextrow >= 0) row_tables(tid, phi_next, smem);
      } break;
    }
  }
};

#if defined(__CUDACC__)
template <class P>
__global__ void __launch_bounds__(P::T, 1) rowfft_kernel(const P prog) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  typename P::Regs r;
  const int tid = (int)threadIdx.x;
  prog.init(tid, r, smem_raw);
  if (blockIdx.x < prog.a.nrows) prog.prefetch_row(blockIdx.x, tid, smem_raw);
  __syncthreads();
  for (long row = blockIdx.x; row < prog.a.nrows; row += gridDim.x) {
    long nxt = row + gridDim.x;
    if (nxt >= prog.a.nrows) nxt = -1;
#pragma unroll
    for (int ph = 0; ph < P::NPH; ++ph) {
      prog.phase(ph, row, nxt, tid, r, smem_raw);
      if (P::sync_after(ph)) __syncthreads();
    }
  }
}
#endif

}  // namespace rowfft
