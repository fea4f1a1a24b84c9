// vpfp_emul.cpp -- HOST EMULATION of the CUDA phase programs, TEST INFRASTRUCTURE ONLY.
//
// Compiles the very same program sources as the GPU library (vlapy_b200/csrc/advect.h, rowops.h)
// with g++ and runs every CTA's phases sequentially over its threads, so that the tile index
// arithmetic and the numerics can be checked against the oracle on the CPU-only build box
// (tests/test_emul_kernels.py).  It is never loaded by the product package, is not a fallback,
// and is not built by __graft_entry__.build() as part of the product library.
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include <utility>

#include "../../vlapy_b200/csrc/advect.h"
#include "../../vlapy_b200/csrc/rowops.h"
#include "../../vlapy_b200/csrc/tridiag.h"
#include "../../vlapy_b200/csrc/spline.h"

template <class Prog>
static void run_prog(const Prog& prog, long nblocks, int threads, long smem_bytes, int nph) {
  std::vector<unsigned char> smem((size_t)(smem_bytes > 0 ? smem_bytes : 16) + 64);
  unsigned char* base = smem.data();
  base += (16 - ((uintptr_t)base & 15)) & 15;
  for (long blk = 0; blk < nblocks; ++blk)
    for (int ph = 0; ph < nph; ++ph)
      for (int tid = 0; tid < threads; ++tid) prog.phase(ph, blk, tid, threads, base);
}

static std::vector<cplx> make_tw(int N) {
  std::vector<cplx> h((size_t)N);
  const long double two_pi = 6.283185307179586476925286766559005768L;
  for (int m = 0; m < N; ++m) {
    long double a = two_pi * (long double)m / (long double)N;
    h[m].x = (double)cosl(a);
    h[m].y = (double)(-sinl(a));
  }
  h[0].x = 1.0; h[0].y = 0.0;
  if (N % 4 == 0) { h[N / 4].x = 0.0; h[N / 4].y = -1.0; h[3 * N / 4].x = 0.0; h[3 * N / 4].y = 1.0; }
  if (N % 2 == 0) { h[N / 2].x = -1.0; h[N / 2].y = 0.0; }
  return h;
}

static void run_advect(AdvectProg a, int max_cols, int max_rows) {
  const AdvectPlan pl = make_advect_plan(a.mode, a.N, max_cols, max_rows);
  std::vector<cplx> tw = make_tw(a.N);
  a.tw = tw.data();
  if (pl.N1 == 1) {
    advect_set_pass(a, pl, 0);
    run_prog(a, a.ntiles(), pl.threads[0], a.smem_bytes(), a.nphases());
    return;
  }
  std::vector<double> phantom((size_t)a.N, 0.0);
  if (a.mode == ADV_ROWS && (a.nrows & 1)) a.phantom = phantom.data();
  for (int pass = 1; pass <= 3; ++pass) {
    advect_set_pass(a, pl, pass);
    run_prog(a, a.ntiles(), pl.threads[pass], a.smem_bytes(), a.nphases());
  }
}

extern "C" {

// max_single_* let the tests force the three-pass decomposition at small sizes.
int emul_edfdv_exp(const double* f_in, long ld_in, double* f_out, long ld_out, const double* e,
                   const double* kv, double dt, int rows, int nv, int max_single) {
  AdvectProg a;
  memset(&a, 0, sizeof(a));
  a.mode = ADV_ROWS; a.op = OP_PHASE; a.N = nv;
  a.nsim = 1; a.nrows = rows; a.nseq = (rows + 1) / 2;
  a.fin = f_in; a.ld_in = ld_in; a.fout = f_out; a.ld_out = ld_out;
  a.kvec = kv; a.cvec = e; a.addv = nullptr; a.dt = dt;
  run_advect(a, max_single, max_single);
  return 0;
}

int emul_vdfdx_exp(const double* f_in, long ld_in, double* f_out, long ld_out, const double* kx,
                   const double* v, double dt, int batch, int nx, int ncols, int max_single) {
  AdvectProg a;
  memset(&a, 0, sizeof(a));
  a.mode = ADV_COLS; a.op = OP_PHASE; a.N = nx;
  a.nsim = batch; a.nrows = nx; a.nseq = ncols / 2;
  a.fin = f_in; a.ld_in = ld_in; a.fout = f_out; a.ld_out = ld_out;
  a.kvec = kx; a.cvec = v; a.addv = nullptr; a.dt = dt;
  run_advect(a, max_single, max_single);
  return 0;
}

int emul_poisson(const double* n, const double* ook, const double* driver, double* e, int batch,
                 int nx, int max_single) {
  AdvectProg a;
  memset(&a, 0, sizeof(a));
  a.mode = ADV_ROWS; a.op = OP_POISSON; a.N = nx;
  a.nsim = 1; a.nrows = batch; a.nseq = (batch + 1) / 2;
  a.fin = n; a.ld_in = nx; a.fout = e; a.ld_out = nx;
  a.kvec = ook; a.cvec = nullptr; a.addv = driver; a.dt = 0.0;
  run_advect(a, max_single, max_single);
  return 0;
}

int emul_edfdv_cd2(const double* f_in, long ld_in, double* f_out, long ld_out, const double* e,
                   double dt, double dv, int rows, int nv) {
  Cd2Prog p;
  p.fin = f_in; p.ld_in = ld_in; p.fout = f_out; p.ld_out = ld_out; p.e = e;
  p.dt = dt; p.dv = dv; p.rows = rows; p.nv = nv;
  const int threads = 256;
  p.cblocks = (nv + threads * 4 - 1) / (threads * 4);
  run_prog(p, (long)rows * p.cblocks, threads, 0, 1);
  return 0;
}

int emul_moments(const double* f, long ld, const double* v, double dv, double* out, long out_ld,
                 int nmom, int rows, int ncols, int edge_flags) {
  MomentsProg p;
  p.f = f; p.ld = ld; p.v = v; p.dv = dv; p.out = out; p.out_ld = out_ld;
  p.nmom = nmom; p.rows = rows; p.ncols = ncols; p.edge_flags = edge_flags;
  int threads = 256;
  while (threads > 32 && threads * 2 > ncols) threads >>= 1;
  run_prog(p, rows, threads, p.smem_bytes(threads), p.nphases(threads));
  return 0;
}

int emul_fp_step(const double* f_in, long ld_in, double* f_out, long ld_out, const double* v,
                 double nu, double dt, double dv, int op, double* moments_out, long mom_ld,
                 int rows, int nv, int m_override) {
  FpProg p;
  p.fin = f_in; p.ld_in = ld_in; p.fout = f_out; p.ld_out = ld_out; p.v = v;
  p.nu = nu; p.dt = dt; p.dv = dv; p.op = op; p.mom_out = moments_out; p.mom_ld = mom_ld;
  p.rows = rows; p.nv = nv;
  int m = nv / 256;
  if (m < 4) m = 4;
  if (m > 16) m = 16;
  if (m_override > 0) m = m_override;
  p.m = m;
  p.P = nv / m;
  int threads = ((p.P + 31) / 32) * 32;
  if (threads > 1024) threads = 1024;
  run_prog(p, rows, threads, p.smem_bytes(threads), p.nphases(threads));
  return 0;
}

// launch geometry as vpfp_fp_diagonals / vpfp_tridiag_solve (vlapy_b200/csrc/vpfp_cuda.cu)
int emul_fp_diagonals(const double* f, long ld, const double* v, double nu, double dt, double dv, int op,
                      double* a, double* b, double* c, int rows, int nv) {
  DiagProg p;
  p.f = f; p.ld = ld; p.v = v; p.nu = nu; p.dt = dt; p.dv = dv; p.op = op;
  p.a = a; p.lda = nv - 1; p.b = b; p.ldb = nv; p.c = c; p.ldc = nv - 1; p.rows = rows; p.nv = nv;
  int threads = 256;
  while (threads > 32 && threads * 2 > nv) threads >>= 1;
  run_prog(p, rows, threads, p.smem_bytes(threads), p.nphases(threads));
  return 0;
}

int emul_tridiag_solve(const double* a, const double* b, const double* c, const double* d, double* x, int rows,
                       int nv, int m_override) {
  TridiagProg p;
  p.a = a; p.lda = nv - 1; p.b = b; p.ldb = nv; p.c = c; p.ldc = nv - 1; p.d = d; p.ldd = nv; p.x = x; p.ldx = nv;
  p.rows = rows; p.nv = nv;
  int m = nv / 256;
  if (m < 4) m = 4;
  if (m > 16) m = 16;
  if (m_override > 0) m = m_override;
  p.m = m;
  p.P = nv / m;
  int threads = ((p.P + 31) / 32) * 32;
  if (threads > 1024) threads = 1024;
  run_prog(p, rows, threads, p.smem_bytes(threads), p.nphases());
  return 0;
}

// semi-Lagrangian operators: launch geometry as vpfp_vdfdx_sl / vpfp_edfdv_sl (vlapy_b200/csrc/vpfp_cuda.cu)
static std::vector<double> spline_tab(int n) {
  std::vector<double> h(3 * (size_t)n);
  double cp = 0.0;
  for (int k = 0; k < n; ++k) { cp = 1.0 / (4.0 - cp); h[k] = cp; h[n + k] = 1.0; h[2 * (size_t)n + k] = 4.0; }
  return h;
}

int emul_vdfdx_sl(const double* f_in, double* f_out, const double* x, const double* v, double dt, double dx, int nx,
                  int nv) {
  std::vector<double> M((size_t)(nx + 2) * nv), tab = spline_tab(nx);
  SplineColSweepProg sw;
  sw.f = f_in; sw.ld = nv; sw.M = M.data(); sw.ldm = nv; sw.cp = tab.data(); sw.h = dx; sw.nx = nx; sw.nv = nv;
  run_prog(sw, (nv + 63) / 64, 64, 0, 1);
  SplineEvalProg<1> ev;
  ev.f = f_in; ev.ld = nv; ev.M = M.data(); ev.ldm = nv; ev.out = f_out; ev.ld_out = nv;
  ev.ax = x; ev.c = v; ev.dt = dt; ev.nx = nx; ev.nv = nv; ev.cblocks = (nv + 255) / 256;
  run_prog(ev, (long)nx * ev.cblocks, 256, 0, 1);
  return 0;
}

int emul_tridiag_solve_ld(const double* a, long lda, const double* b, long ldb, const double* c, long ldc,
                          const double* d, long ldd, double* x, long ldx, int rows, int nv) {
  TridiagProg p;
  p.a = a; p.lda = lda; p.b = b; p.ldb = ldb; p.c = c; p.ldc = ldc; p.d = d; p.ldd = ldd; p.x = x; p.ldx = ldx;
  p.rows = rows; p.nv = nv;
  int m = nv / 256;
  if (m < 4) m = 4;
  if (m > 16) m = 16;
  p.m = m;
  p.P = nv / m;
  int threads = ((p.P + 31) / 32) * 32;
  if (threads > 1024) threads = 1024;
  run_prog(p, rows, threads, p.smem_bytes(threads), p.nphases());
  return 0;
}

int emul_edfdv_sl(const double* f_in, double* f_out, const double* e, const double* v, double dt, double dv, int nx,
                  int nv) {
  if (nv - 2 < 8) return 1;
  const long ldm = nv + 2;
  std::vector<double> M((size_t)nx * ldm), tab = spline_tab(nv);
  SplineRowRhsProg rh;
  rh.f = f_in; rh.ld = nv; rh.M = M.data(); rh.ldm = ldm; rh.h = dv; rh.nx = nx; rh.nv = nv; rh.cblocks = (nv + 255) / 256;
  run_prog(rh, (long)nx * rh.cblocks, 256, 0, 1);
  emul_tridiag_solve_ld(tab.data() + nv, 0, tab.data() + 2 * (long)nv, 0, tab.data() + nv, 0, M.data() + 2, ldm,
                        M.data() + 2, ldm, nx, nv - 2);
  SplineEvalProg<0> ev;
  ev.f = f_in; ev.ld = nv; ev.M = M.data(); ev.ldm = ldm; ev.out = f_out; ev.ld_out = nv;
  ev.ax = v; ev.c = e; ev.dt = dt; ev.nx = nx; ev.nv = nv; ev.cblocks = (nv + 255) / 256;
  run_prog(ev, (long)nx * ev.cblocks, 256, 0, 1);
  return 0;
}

int emul_xmodes(const double* f, long ld, double* out, int nmodes, int batch, int nx, int ncols) {
  XmodesProg p;
  p.f = f; p.ld = ld; p.nmodes = nmodes; p.batch = batch; p.nx = nx; p.ncols = ncols;
  p.x_offset = 0; p.nx_total = nx;
  const int threads = 128;
  p.cblocks = (ncols + threads - 1) / threads;
  int rows_per_chunk = 128;                        // the chunk rule of vpfp_xmodes_partial
  {
    const long col_ctas = (long)batch * ((ncols / 2 + threads - 1) / threads);
    while (rows_per_chunk > 8 && col_ctas * ((nx + rows_per_chunk - 1) / rows_per_chunk) < 148) rows_per_chunk >>= 1;
  }
  int xch = nx / rows_per_chunk;
  if (xch < 1) xch = 1;
  if (xch > 32) xch = 32;
  p.xchunks = xch;
  std::vector<double> partial((size_t)batch * xch * nmodes * ncols * 2);
  p.partial = partial.data();
  if (nmodes == 2 && (ncols & 1) == 0 && (ld & 1) == 0 && ((uintptr_t)f & 15) == 0) {   // as vpfp_xmodes_partial
    Xmodes2Prog p2;
    p2.f = f; p2.ld = ld; p2.partial = p.partial; p2.batch = batch; p2.nx = nx; p2.ncols = ncols;
    p2.xchunks = xch; p2.cblocks = (ncols / 2 + threads - 1) / threads;
    p2.x_offset = 0; p2.nx_total = nx;
    run_prog(p2, (long)batch * xch * p2.cblocks, threads, 0, 1);
  } else {
    run_prog(p, (long)batch * xch * p.cblocks, threads, 0, 1);
  }
  XmodesReduceProg r;
  r.partial = p.partial; r.out = out; r.nmodes = nmodes; r.batch = batch; r.ncols = ncols; r.xchunks = xch;
  long total = (long)batch * nmodes * ncols * 2;
  run_prog(r, (total + 255) / 256, 256, 0, 1);
  return 0;
}

int emul_driver(const double* x, double t, const double* pulses, int npulse, double* out, int nx) {
  DriverProg p;
  p.x = x; p.out = out; p.t = t; p.t_dev = nullptr; p.ninc = 0; p.nx = nx; p.npulse = npulse;
  for (int i = 0; i < npulse * 7; ++i) p.pulses[i] = pulses[i];
  run_prog(p, (nx + 255) / 256, 256, 0, 1);
  return 0;
}

int emul_series(const double* moments, long mom_ld, const double* e, const double* de, double* out, int nx) {
  SeriesProg p;
  p.mom = moments; p.mom_ld = mom_ld; p.e = e; p.de = de; p.out = out; p.nx = nx;
  int threads = 256;
  while (threads > 32 && threads > nx) threads >>= 1;
  run_prog(p, 1, threads, p.smem_bytes(threads), p.nphases(threads));
  return 0;
}
}

// ---- register butterflies of the fast kernels (vlapy_b200/csrc/butterflies.h)
#include "../../vlapy_b200/csrc/butterflies.h"
extern "C" int emul_butterfly(double* xy, int radix, int dir) {
  cplx x[16];
  for (int i = 0; i < radix; ++i) { x[i].x = xy[2 * i]; x[i].y = xy[2 * i + 1]; }
  if (radix == 4) { if (dir < 0) fast::fft4<-1>(x[0], x[1], x[2], x[3]); else fast::fft4<1>(x[0], x[1], x[2], x[3]); }
  else if (radix == 8) { if (dir < 0) fast::fft8<-1>(x); else fast::fft8<1>(x); }
  else if (radix == 16) { if (dir < 0) fast::fft16<-1>(x); else fast::fft16<1>(x); }
  else return 1;
  for (int i = 0; i < radix; ++i) { xy[2 * i] = x[i].x; xy[2 * i + 1] = x[i].y; }
  return 0;
}

// ---- mid-size single-pass kernels (vlapy_b200/csrc/midfft.cuh): per-thread registers persist across phases
#include "../../vlapy_b200/csrc/midfft.cuh"
template <class P>
static void run_midfft(const P& prog) {
  std::vector<typename P::Regs> regs((size_t)P::NT);
  std::vector<unsigned char> smem((size_t)P::SMEM_BYTES + 64);
  unsigned char* base = smem.data();
  base += (16 - ((uintptr_t)base & 15)) & 15;
  const char* oe = getenv("VPFP_EMUL_ORDER");
  const int order = oe ? atoi(oe) : 0;
  for (int tid = 0; tid < P::NT; ++tid) prog.init(tid, base);
  // two emulated CTAs (grid = 2) walk the tiles: every CTA prefetches its next tile into its own slots
  const long grid = 2;
  for (long cta = 0; cta < grid; ++cta) {
    bool prefetched = false;
    for (long tile = cta; tile < prog.ntiles(); tile += grid) {
      const long nxt = (tile + grid < prog.ntiles()) ? tile + grid : -1;
      for (int ph = 0; ph < P::NPH; ++ph)
        for (int i = 0; i < P::NT; ++i) {
          int tid = i;
          if (order == 1) tid = P::NT - 1 - i;
          else if (order == 2) tid = (i & 1) ? P::NT / 2 + i / 2 : i / 2;
          prog.phase(ph, tile, nxt, prefetched, tid, regs[tid], base);
        }
      prefetched = true;
    }
  }
}

extern "C" int emul_midfft(int mode, const double* f_in, long ld_in, double* f_out, long ld_out, const double* kvec,
                           const double* cvec, double dt, int nsim, int nrows_or_nx, int ncols) {
  midfft::Args a;
  memset(&a, 0, sizeof(a));
  int N;
  if (mode == ADV_COLS) { N = nrows_or_nx; a.nsim = nsim; a.nrows = N; a.nseq = ncols / 2; }
  else { N = ncols; a.nsim = 1; a.nrows = nrows_or_nx; a.nseq = (nrows_or_nx + 1) / 2; }
  std::vector<cplx> tw = make_tw(N);
  a.fin = f_in; a.ld_in = ld_in; a.fout = f_out; a.ld_out = ld_out; a.kvec = kvec; a.cvec = cvec; a.dt = dt;
  a.tw = tw.data();
  if (mode == ADV_COLS) {
    if (N == 256) { midfft::Prog<256, 8, 4, ADV_COLS, 8> p; p.a = a; run_midfft(p); }
    else if (N == 512) { midfft::Prog<512, 8, 8, ADV_COLS, 8> p; p.a = a; run_midfft(p); }
    else if (N == 1024) { midfft::Prog<1024, 16, 8, ADV_COLS, 4> p; p.a = a; run_midfft(p); }
    else if (N == 2048) { midfft::Prog<2048, 16, 16, ADV_COLS, 4> p; p.a = a; run_midfft(p); }
    else return 1;
  } else {
    if (N == 256) { midfft::Prog<256, 8, 4, ADV_ROWS, 8> p; p.a = a; run_midfft(p); }
    else if (N == 512) { midfft::Prog<512, 8, 8, ADV_ROWS, 8> p; p.a = a; run_midfft(p); }
    else if (N == 1024) { midfft::Prog<1024, 16, 8, ADV_ROWS, 4> p; p.a = a; run_midfft(p); }
    else if (N == 2048) { midfft::Prog<2048, 16, 16, ADV_ROWS, 2> p; p.a = a; run_midfft(p); }
    else return 1;
  }
  return 0;
}

// ---- nx = 16 / 32: whole transform in the registers of one thread (vlapy_b200/csrc/tinyfft.cuh)
#include "../../vlapy_b200/csrc/tinyfft.cuh"
extern "C" int emul_tiny_cols(const double* f_in, long ld_in, double* f_out, long ld_out, const double* kvec, const double* cvec,
                              double dt, int nsim, int nx, int ncols) {
  const long work = (long)nsim * (ncols / 2);
  if (nx == 32) {
    tiny::ColsProg<32> p;
    p.nsim = nsim; p.nseq = ncols / 2; p.fin = f_in; p.ld_in = ld_in; p.fout = f_out; p.ld_out = ld_out; p.kvec = kvec; p.cvec = cvec; p.dt = dt;
    run_prog(p, (work + 127) / 128, 128, 0, 1);
  } else if (nx == 16) {
    tiny::ColsProg<16> p;
    p.nsim = nsim; p.nseq = ncols / 2; p.fin = f_in; p.ld_in = ld_in; p.fout = f_out; p.ld_out = ld_out; p.kvec = kvec; p.cvec = cvec; p.dt = dt;
    run_prog(p, (work + 127) / 128, 128, 0, 1);
  } else return 1;
  return 0;
}
extern "C" int emul_tiny_poisson(const double* n, const double* ook, const double* driver, double* e, int batch, int nx) {
  const long work = (batch + 1) / 2;
  if (nx == 32) {
    tiny::PoissonProg<32> p;
    p.nrows = batch; p.n = n; p.ook = ook; p.driver = driver; p.e = e;
    run_prog(p, (work + 127) / 128, 128, 0, 1);
  } else if (nx == 16) {
    tiny::PoissonProg<16> p;
    p.nrows = batch; p.n = n; p.ook = ook; p.driver = driver; p.e = e;
    run_prog(p, (work + 127) / 128, 128, 0, 1);
  } else return 1;
  return 0;
}

// spectral Poisson solve through the mid-size kernel in Poisson mode (two density rows per packed sequence)
extern "C" int emul_midfft_poisson(const double* n, const double* ook, const double* driver, double* e, int batch, int nx) {
  midfft::Args a;
  memset(&a, 0, sizeof(a));
  a.nsim = 1; a.nrows = batch; a.nseq = (batch + 1) / 2;
  std::vector<cplx> tw = make_tw(nx);
  a.fin = n; a.ld_in = nx; a.fout = e; a.ld_out = nx; a.kvec = ook; a.addv = driver; a.tw = tw.data();
  if (nx == 256) { midfft::Prog<256, 8, 4, ADV_ROWS, 8, false, true> p; p.a = a; run_midfft(p); }
  else if (nx == 512) { midfft::Prog<512, 8, 8, ADV_ROWS, 4, false, true> p; p.a = a; run_midfft(p); }
  else if (nx == 1024) { midfft::Prog<1024, 16, 8, ADV_ROWS, 4, false, true> p; p.a = a; run_midfft(p); }
  else if (nx == 2048) { midfft::Prog<2048, 16, 16, ADV_ROWS, 2, false, true> p; p.a = a; run_midfft(p); }
  else return 1;
  return 0;
}

// v df/dx with the charge density fused into the store phase (Prog<..., DENS = true>); the sum over the column tiles
// (fast::dens_reduce_kernel on the device) is done here on the host, in tile order
extern "C" int emul_midfft_cols_density(const double* f_in, long ld_in, double* f_out, long ld_out, const double* kvec,
                                        const double* cvec, double dt, int nsim, int nx, int ncols, double dv,
                                        int edge_flags, double* n_out) {
  midfft::Args a;
  memset(&a, 0, sizeof(a));
  const int N = nx;
  a.nsim = nsim; a.nrows = N; a.nseq = ncols / 2;
  std::vector<cplx> tw = make_tw(N);
  a.fin = f_in; a.ld_in = ld_in; a.fout = f_out; a.ld_out = ld_out; a.kvec = kvec; a.cvec = cvec; a.dt = dt;
  a.tw = tw.data();
  const int CB = (N == 1024) ? 4 : 8;
  const int tiles = (a.nseq + CB - 1) / CB;
  std::vector<double> partial((size_t)tiles * nsim * N, -1.0);
  a.dens_partial = partial.data(); a.dv = dv; a.edge_flags = edge_flags;
  if (N == 256) { midfft::Prog<256, 8, 4, ADV_COLS, 8, true> p; p.a = a; run_midfft(p); }
  else if (N == 512) { midfft::Prog<512, 8, 8, ADV_COLS, 8, true> p; p.a = a; run_midfft(p); }
  else if (N == 1024) { midfft::Prog<1024, 16, 8, ADV_COLS, 4, true> p; p.a = a; run_midfft(p); }
  else return 1;
  const long n = (long)nsim * N;
  for (long i = 0; i < n; ++i) {
    double s2 = 0.0;
    for (int t = 0; t < tiles; ++t) s2 += partial[(size_t)t * n + i];
    n_out[i] = s2;
  }
  return 0;
}

// ---- single-pass row kernel (vlapy_b200/csrc/rowfft.cuh): per-thread registers persist across phases
#include "../../vlapy_b200/csrc/rowfft.cuh"
template <class P>
static void run_rowfft(const P& prog) {
  std::vector<typename P::Regs> regs((size_t)P::T);
  std::vector<unsigned char> smem((size_t)P::SMEM_BYTES + 64);
  unsigned char* base = smem.data();
  base += (16 - ((uintptr_t)base & 15)) & 15;
  for (int tid = 0; tid < P::T; ++tid) prog.init(tid, regs[tid], base);
  for (int tid = 0; tid < P::T; ++tid) prog.prefetch_row(0, tid, base);
  // threads of a phase may run in any order: VPFP_EMUL_ORDER=1 reverses it, 2 interleaves halves
  // (a race between threads inside a phase shows up as a result that depends on the order)
  const char* oe = getenv("VPFP_EMUL_ORDER");
  const int order = oe ? atoi(oe) : 0;
  // the (row, phase) steps between two barriers form one interval: every thread runs the whole interval before the
  // next thread starts (P::sync_after(ph) == false: no barrier after that phase, e.g. the last phase of a row and
  // the first phase of the next one in rowfft.cuh)
  std::vector<std::pair<long, int>> interval;
  for (long row = 0; row < prog.a.nrows; ++row)
    for (int ph = 0; ph < P::NPH; ++ph) {
      interval.push_back({row, ph});
      const bool last = (row + 1 == prog.a.nrows) && (ph + 1 == P::NPH);
      if (!P::sync_after(ph) && !last) continue;
      for (int i = 0; i < P::T; ++i) {
        int tid = i;
        if (order == 1) tid = P::T - 1 - i;
        else if (order == 2) tid = (i & 1) ? P::T / 2 + i / 2 : i / 2;
        for (auto& st : interval)
          prog.phase(st.second, st.first, st.first + 1 < prog.a.nrows ? st.first + 1 : -1, tid, regs[tid], base);
      }
      interval.clear();
    }
}

extern "C" int emul_poisson_rowfft(const double* n, const double* ook, const double* driver, double* e, int batch, int nx) {
  std::vector<cplx> tw = make_tw(nx);
  rowfft::Args a;
  memset(&a, 0, sizeof(a));
  a.fin = n; a.ld_in = nx; a.fout = e; a.ld_out = nx; a.kvec = ook; a.cvec = nullptr; a.dt = 0.0;
  a.nrows = batch; a.twN = tw.data(); a.addv = driver;
  if (nx == 16384) { rowfft::Prog<32, 16, false, true> p; p.a = a; run_rowfft(p); }
  else if (nx == 8192) { rowfft::Prog<16, 16, false, true> p; p.a = a; run_rowfft(p); }
  else if (nx == 4096) { rowfft::Prog<8, 16, false, true> p; p.a = a; run_rowfft(p); }
  else return 1;
  return 0;
}

extern "C" int emul_edfdv_rowfft(const double* f_in, long ld_in, double* f_out, long ld_out, const double* e,
                                 const double* kv, double dt, int rows, int nv) {
  std::vector<cplx> tw = make_tw(nv);
  rowfft::Args a;
  memset(&a, 0, sizeof(a));
  a.fin = f_in; a.ld_in = ld_in; a.fout = f_out; a.ld_out = ld_out; a.kvec = kv; a.cvec = e; a.dt = dt;
  a.nrows = rows; a.twN = tw.data();
  if (nv == 16384) { rowfft::Prog<32, 16> p; p.a = a; run_rowfft(p); }
  else if (nv == 8192) { rowfft::Prog<16, 16> p; p.a = a; run_rowfft(p); }
  else if (nv == 4096) { rowfft::Prog<8, 16> p; p.a = a; run_rowfft(p); }
  else return 1;
  return 0;
}
