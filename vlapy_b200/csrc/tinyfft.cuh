// tinyfft.cuh -- v df/dx (vlapy/core/vlasov.py:94-108) and the spectral field solve (vlapy/core/field.py:39-88) for
// nx = 16 / 32, the grid of the reference's Landau-damping test (32 x 512, BASELINE config 1): the whole x transform of
// one packed sequence lives in the registers of ONE thread -- no shared memory, no barrier, one launch.  Before, these
// lengths took the generic shared-memory program (radix-2/4 stages separated by barriers: 12 us and 9 us per launch at
// 32 x 512, a sixth and an eighth of the CUDA-graph step).
//   ColsProg    a thread owns two adjacent v-columns packed as one complex sequence (one 16-byte load per row; the lanes of
//               a warp read 512 contiguous bytes of every row), transforms, un-mixes the pair (k, N-k), multiplies by the
//               phase factors P(k) = exp(-i (K[1] dt) c k) / (2N) -- powers of P(1), at most 16 multiplications deep --,
//               transforms back.  The Nyquist factor keeps its real part (np.real of the reference).
//   PoissonProg a thread owns two adjacent density rows (two simulations of an ensemble) packed the same way; multiplier
//               i (ook[k] - ook[N-k]) / 2 per row, the Hermitian part of the reference's i one_over_kx (what survives its
//               np.real); the driver row is added by the store.
// Phase programs (vpfp_common.h): tests/emul runs the same source on the host.
#pragma once
#include "rowfft.cuh"

namespace tiny {

VPFP_HD cplx ld_pair(const double* p) {       // 16-byte aligned pair of adjacent columns
#if defined(__CUDA_ARCH__)
  const double2 v = *reinterpret_cast<const double2*>(p);
  return cmake(v.x, v.y);
#else
  return cmake(p[0], p[1]);
#endif
}
VPFP_HD void st_pair(double* p, cplx v) {
#if defined(__CUDA_ARCH__)
  *reinterpret_cast<double2*>(p) = make_double2(v.x, v.y);
#else
  p[0] = v.x; p[1] = v.y;
#endif
}

template <int N, int DIR>
VPFP_HD void fftN(cplx* x) {
  if (N == 32) rowfft::fft32<DIR>(x);
  else fast::fft16<DIR>(x);
}

// the pair (k, N-k) of two packed real channels: Zr = Z[k], Zpr = Z[N-k] in; multiplied by Pa / Pb (channel a / b, both
// pre-scaled by 1 / (2N)) and re-packed out
VPFP_HD void pair(cplx& Zr, cplx& Zpr, const cplx Pa, const cplx Pb, const bool selfpair) {
  const cplx Z = Zr, Zp = Zpr;
  const cplx U = cadd(Z, cconj(Zp)), V = csub(Z, cconj(Zp));
  const cplx X1 = cmul(Pa, U), X2 = cmul(Pb, V);
  Zr = cadd(X1, X2);
  if (!selfpair) Zpr = cconj(csub(X1, X2));
}

template <int N>
struct ColsProg {
  int nsim, nseq;            // nseq = ncols / 2 packed column pairs per simulation
  const double* fin; long ld_in;
  double* fout; long ld_out;
  const double* kvec;        // [nsim][N] (only K[1] is used: uniform fftfreq grid)
  const double* cvec;        // v[ncols]
  double dt;

  VPFP_HD void phase(int, long blk, int tid, int nthr, unsigned char*) const {
    const long g = blk * nthr + tid;
    if (g >= (long)nsim * nseq) return;
    const int sim = (int)(g / nseq), seq = (int)(g % nseq);
    const double* src = fin + ((long)sim * N) * ld_in + 2L * seq;
    double* dst = fout + ((long)sim * N) * ld_out + 2L * seq;
    cplx x[N];
#pragma unroll
    for (int n = 0; n < N; ++n) x[n] = ld_pair(src + n * ld_in);
    fftN<N, -1>(x);
    const double kdt = mul_rn(kvec[(long)sim * N + 1], dt);
    double sa, ca, sb, cb;
    sincos_hd(mul_rn(kdt, cvec[2 * seq]), &sa, &ca);
    sincos_hd(mul_rn(kdt, cvec[2 * seq + 1]), &sb, &cb);
    const cplx Wa = cmake(ca, -sa), Wb = cmake(cb, -sb);      // exp(-i phi)
    const double sc = 0.5 / (double)N;
    cplx Pa = cmake(sc, 0.0), Pb = cmake(sc, 0.0);
    pair(x[0], x[0], Pa, Pb, true);
#pragma unroll
    for (int k = 1; k < N / 2; ++k) {
      Pa = cmul(Pa, Wa); Pb = cmul(Pb, Wb);
      pair(x[k], x[N - k], Pa, Pb, false);
    }
    Pa = cmul(Pa, Wa); Pb = cmul(Pb, Wb);
    Pa.y = 0.0; Pb.y = 0.0;                                    // Nyquist: real part
    pair(x[N / 2], x[N / 2], Pa, Pb, true);
    fftN<N, 1>(x);
#pragma unroll
    for (int n = 0; n < N; ++n) st_pair(dst + n * ld_out, x[n]);
  }
};

template <int N>
struct PoissonProg {
  int nrows;                 // density rows (simulations); a thread takes rows 2 g and 2 g + 1
  const double* n;           // [nrows][N]
  const double* ook;         // [nrows][N]
  const double* driver;      // [nrows][N] or null
  double* e;                 // [nrows][N]

  VPFP_HD void phase(int, long blk, int tid, int nthr, unsigned char*) const {
    const long g = blk * nthr + tid;
    const long ra = 2 * g, rb = ra + 1;
    if (ra >= nrows) return;
    const bool hb = rb < nrows;
    cplx x[N];
#pragma unroll
    for (int i = 0; i < N; ++i) x[i] = cmake(1.0 - n[ra * N + i], hb ? 1.0 - n[rb * N + i] : 0.0);
    fftN<N, -1>(x);
    const double sc = 0.5 / (double)N;
    x[0] = cmake(0.0, 0.0);                                    // multiplier 0 at k = 0 and at the Nyquist bin
    x[N / 2] = cmake(0.0, 0.0);
#pragma unroll
    for (int k = 1; k < N / 2; ++k) {
      const double ma = 0.5 * (ook[ra * N + k] - ook[ra * N + N - k]);
      const double mb = hb ? 0.5 * (ook[rb * N + k] - ook[rb * N + N - k]) : 0.0;
      pair(x[k], x[N - k], cmake(0.0, ma * sc), cmake(0.0, mb * sc), false);
    }
    fftN<N, 1>(x);
#pragma unroll
    for (int i = 0; i < N; ++i) {
      double va = x[i].x, vb = x[i].y;
      if (driver != nullptr) { va += driver[ra * N + i]; if (hb) vb += driver[rb * N + i]; }
      e[ra * N + i] = va;
      if (hb) e[rb * N + i] = vb;
    }
  }
};

}  // namespace tiny
